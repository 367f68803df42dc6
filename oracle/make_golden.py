"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden
The reference is imported through oracle/reference_loader.py (three no-arithmetic stand-ins for its missing
third-party imports).  Charts are produced under torch.set_default_dtype(float64) (oracle tier T1, SURVEY.md
section 8c) except where `shipped_dtype` says the shipped fp32 chart was kept (tier T0).
Each case stores: the normalised brick probabilities (what `.p` returns), the input sequence, Z, the
per-variable chart in the reference's TMap layout, and the gradients d Z / d parameter exactly as
`Z.backward()` leaves them on the Parameters (i.e. AFTER the project_grad hooks, util.py:147-156,200-202).
"""
import json
import os

import numpy as np
import torch

from oracle.reference_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def describe(ref, model):
    """bricks description (plain arrays) of a reference SequentialRBN."""
    pcfg = ref["pcfg"]
    cells = []
    for c in model._cells:
        trs = []
        for t in c._transitions:
            p = t.transition_probabilities.p.detach().numpy().astype(np.float64)
            if isinstance(t, pcfg.DiscreteTerminalTransition):
                trs.append(["term", p, int(t.term_idx)])
            else:
                trs.append(["bin", p, int(t.left_idx), int(t.right_idx)])
        cells.append(dict(cardinality=int(c.variable.cardinality),
                          struct=c.transition_probabilities.p.detach().numpy().astype(np.float64),
                          transitions=trs))
    prior = dict(struct=model._prior.structural_distribution.p.detach().numpy().astype(np.float64),
                 dists=[d.p.detach().numpy().astype(np.float64) for d in model._prior.prior_distributions])
    return cells, prior


def pack_case(store, key, ref, model, sequence, Z, extra=None):
    cells, prior = describe(ref, model)
    meta = dict(n_cells=len(cells), cards=[c["cardinality"] for c in cells], transitions=[], extra=extra or {})
    for v, c in enumerate(cells):
        store[f"{key}/cell{v}/struct"] = c["struct"]
        kinds = []
        for t, tr in enumerate(c["transitions"]):
            store[f"{key}/cell{v}/tr{t}"] = tr[1]
            kinds.append([tr[0]] + [int(x) for x in tr[2:]])
        meta["transitions"].append(kinds)
    store[f"{key}/prior/struct"] = prior["struct"]
    for v, d in enumerate(prior["dists"]):
        store[f"{key}/prior/dist{v}"] = d
    seq = [[(-1 if y is None else int(y)) for y in np.atleast_1d(np.asarray(pos, dtype=object))] for pos in sequence]
    store[f"{key}/sequence"] = np.asarray(seq, dtype=np.int64)
    store[f"{key}/Z"] = np.asarray(float(Z.detach()))
    for v, ch in enumerate(model.inside_chart):
        store[f"{key}/chart{v}"] = ch.arr.detach().numpy().astype(np.float64)
    if Z.requires_grad and float(Z.detach()) > 0:
        model.zero_grad()
        Z.backward()
        for name, p in model.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            store[f"{key}/grad/{name}"] = g.detach().numpy().astype(np.float64)
    return meta


def random_rbn(ref, rng, cards, V, n_tvars, prob_rep, none_prob=0.0):
    """Random multi-cell RBN: every cell has one terminal transition per terminal variable it can emit and one
    binary transition for every ordered pair of child variables (dense in structure, random in weights)."""
    pcfg = ref["pcfg"]
    n = len(cards)
    cells = []
    for v, K in enumerate(cards):
        trs = []
        for t in range(n_tvars):
            trs.append(pcfg.DiscreteTerminalTransition(weights=rng.random((V, K)) + 0.05, term_idx=t, prob_rep=prob_rep))
        for li in range(n):
            for ri in range(n):
                w = rng.random((cards[li], cards[ri], K)) + 0.01
                trs.append(pcfg.DiscreteBinaryNonTerminalTransition(weights=w, left_idx=li, right_idx=ri, prob_rep=prob_rep))
        cells.append(pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(K), weights=rng.random((len(trs), K)) + 0.05,
                                       transitions=trs, prob_rep=prob_rep))
    prior = pcfg.DiscretePrior(struc_weights=rng.random(n) + 0.1, prior_weights=[rng.random(K) + 0.1 for K in cards],
                               prob_rep=prob_rep)
    return ref["sequential"].SequentialRBN(cells=cells, prior=prior)


def random_sequence(rng, L, V, n_tvars, none_prob):
    seq = []
    for _ in range(L):
        pos = [int(rng.integers(0, V)) if rng.random() >= none_prob else None for _ in range(n_tvars)]
        if all(p is None for p in pos):
            pos[int(rng.integers(0, n_tvars))] = int(rng.integers(0, V))
        seq.append(pos)
    return seq


def main():
    ref = load_reference()
    pcfg, util = ref["pcfg"], ref["util"]
    os.makedirs(OUT, exist_ok=True)
    store, metas = {}, {}

    # ---- the reference's own test grammars (tests/test_pcfg.py:21-35) in shipped dtype (fp32 chart) -------------
    g = pcfg.AbstractedPCFG(non_terminals="SAB", terminals="ab", start="S", rules=[
        ("S --> A B", 1), ("S --> B A", 1), ("A --> B A", 1), ("B --> A B", 1), ("A --> a", 1), ("B --> b", 1)],
        prob_rep=util.Prob)
    for s in ["aaaa", "bbbb", "bbba", "aaab"]:
        Z = g.inside(sequence=s)
        metas[f"binary/{s}"] = pack_case(store, f"binary/{s}", ref, g, g.tokenise(s), Z, dict(shipped_dtype=True))

    torch.set_default_dtype(torch.float64)
    # ---- plot_PCFG.py / test_word_pcfg grammar (tests/test_base.py:250-295) ------------------------------------
    subjects = ["I", "You", "We", "They"]; verbs = ["run", "drink", "sleep"]
    adverb_non_gradable = ["a-lot", "alone"]; adverb_gradable = ["fast", "slowly", "quickly"]
    grade = ["very", "veeery", "really"]; verb_qualifier = ["rarely", "do-not", "never", "always"]
    nts = ["start", "subject", "verb", "gradable_adverb", "non_gradable_adverb", "verb_qualifier", "grade"]
    terms = subjects + verbs + adverb_non_gradable + adverb_gradable + grade + verb_qualifier
    rules = [("start --> subject verb", 1), ("verb --> verb_qualifier verb", 1), ("verb --> verb gradable_adverb", 1),
             ("verb --> verb non_gradable_adverb", 1), ("gradable_adverb --> grade gradable_adverb", 1),
             ("grade --> grade grade", 1)] + \
            [(f"{n} --> {t}", 1) for n, ts in zip(nts[1:], [subjects, verbs, adverb_gradable, adverb_non_gradable,
                                                            verb_qualifier, grade]) for t in ts]
    for rep_name, rep in [("Prob", util.Prob), ("LogProb", util.LogProb)]:
        wg = pcfg.AbstractedPCFG(non_terminals=nts, terminals=terms, rules=rules, start="start", prob_rep=rep)
        for k, sent in enumerate(["I run", "You never run", "We run very very slowly", "They always run alone",
                                  "I never sleep really very quickly", "You do-not drink very quickly",
                                  "I You", "run fast"]):
            Z = wg.inside(sequence=sent.split())
            metas[f"word{rep_name}/{k}"] = pack_case(store, f"word{rep_name}/{k}", ref, wg, wg.tokenise(sent.split()), Z,
                                                     dict(sentence=sent, terminals=terms, non_terminals=nts))

    # ---- random single-cell grammars (the AbstractedPCFG shape), fp64, with gradients ------------------------------
    rng = np.random.default_rng(20261019)
    k = 0
    for rep_name, rep in [("LogProb", util.LogProb), ("Prob", util.Prob)]:
        for (N, V, L) in [(2, 2, 1), (3, 4, 2), (3, 4, 5), (5, 6, 9), (8, 5, 12), (10, 19, 20), (16, 7, 7)]:
            m = random_rbn(ref, rng, [N], V, 1, rep)
            seq = random_sequence(rng, L, V, 1, 0.0)
            Z = m.inside(sequence=seq)
            metas[f"rand1_{rep_name}/{k}"] = pack_case(store, f"rand1_{rep_name}/{k}", ref, m, seq, Z)
            k += 1
    # ---- random multi-cell RBNs with several terminal variables and None terminals --------------------------------
    k = 0
    for rep_name, rep in [("LogProb", util.LogProb), ("Prob", util.Prob)]:
        for (cards, V, nt, L, nonep) in [([3, 4], 5, 1, 5, 0.0), ([2, 3], 3, 2, 6, 0.4), ([1, 1], 1, 2, 4, 0.5),
                                         ([2, 3, 2], 4, 2, 7, 0.3)]:
            m = random_rbn(ref, rng, cards, V, nt, rep)
            seq = random_sequence(rng, L, V, nt, nonep)
            Z = m.inside(sequence=seq)
            metas[f"randmulti_{rep_name}/{k}"] = pack_case(store, f"randmulti_{rep_name}/{k}", ref, m, seq, Z)
            k += 1
    # ---- sparse grammar with zero rules (structural zeros propagate; ungrammatical spans are exact zeros) ---------
    for k in range(3):
        N, V, L = 6, 4, 8
        W = rng.random((N, N, N)) * (rng.random((N, N, N)) < 0.15)
        W[0, 0, :] += 0.01  # every parent keeps at least one binary rule
        T = rng.random((V, N)) * (rng.random((V, N)) < 0.6)
        T[0, :] += 0.01
        trs = [pcfg.DiscreteTerminalTransition(weights=T), pcfg.DiscreteBinaryNonTerminalTransition(weights=W)]
        cell = pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(N), weights=rng.random((2, N)) + 0.05, transitions=trs)
        pw = np.zeros(N); pw[0] = 1
        m = ref["sequential"].SequentialRBN(cells=[cell], prior=pcfg.DiscretePrior(struc_weights=np.ones(1), prior_weights=[pw]))
        seq = random_sequence(rng, L, V, 1, 0.0)
        Z = m.inside(sequence=seq)
        metas[f"sparse/{k}"] = pack_case(store, f"sparse/{k}", ref, m, seq, Z)

    np.savez_compressed(os.path.join(OUT, "reference_cases.npz"), **store)
    with open(os.path.join(OUT, "reference_cases.json"), "w") as f:
        json.dump(metas, f, indent=1)
    print(f"wrote {len(metas)} cases, {len(store)} arrays to {OUT}")


if __name__ == "__main__":
    main()
