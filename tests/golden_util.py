"""Load the committed reference fixtures (tests/golden/reference_cases.{npz,json}; made by oracle/make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    store = np.load(os.path.join(HERE, "reference_cases.npz"))
    with open(os.path.join(HERE, "reference_cases.json")) as f:
        metas = json.load(f)
    cases = {}
    for key, meta in metas.items():
        cells = []
        for v in range(meta["n_cells"]):
            trs = []
            for t, kind in enumerate(meta["transitions"][v]):
                trs.append(tuple([kind[0], store[f"{key}/cell{v}/tr{t}"]] + kind[1:]))
            cells.append(dict(cardinality=meta["cards"][v], struct=store[f"{key}/cell{v}/struct"], transitions=trs))
        prior = dict(struct=store[f"{key}/prior/struct"],
                     dists=[store[f"{key}/prior/dist{v}"] for v in range(meta["n_cells"])])
        seq = [[(None if y < 0 else int(y)) for y in pos] for pos in store[f"{key}/sequence"]]
        grads = {k[len(key) + 6:]: store[k] for k in store.files if k.startswith(f"{key}/grad/")}
        cases[key] = dict(cells=cells, prior=prior, sequence=seq, Z=float(store[f"{key}/Z"]),
                          charts=[store[f"{key}/chart{v}"] for v in range(meta["n_cells"])],
                          grads=grads, meta=meta)
    return cases


def build_model(case, prob_rep):
    """An rbn_b200 SequentialRBN with exactly the (normalised) brick probabilities of a golden case."""
    from rbn_b200 import pcfg, sequential
    cells = []
    for c in case["cells"]:
        trs = []
        for tr in c["transitions"]:
            if tr[0] == "term":
                trs.append(pcfg.DiscreteTerminalTransition(weights=tr[1], term_idx=tr[2], prob_rep=prob_rep))
            else:
                trs.append(pcfg.DiscreteBinaryNonTerminalTransition(weights=tr[1], left_idx=tr[2], right_idx=tr[3],
                                                                    prob_rep=prob_rep))
        cells.append(pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(c["cardinality"]), weights=c["struct"],
                                       transitions=trs, prob_rep=prob_rep))
    prior = pcfg.DiscretePrior(struc_weights=case["prior"]["struct"], prior_weights=case["prior"]["dists"],
                               prob_rep=prob_rep)
    return sequential.SequentialRBN(cells=cells, prior=prior)


def case_prob_rep(case):
    from rbn_b200 import util
    if any(k.endswith("log_p") for k in case["grads"]):
        return util.LogProb
    return util.Prob
