"""Host-side API mirror, grammar flattening and the C-ABI library surface (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from rbn_b200 import _lib, base, pcfg, sequential, util
from rbn_b200.engine import InsideEngine
from rbn_b200.tmap import TMap
from oracle import inside_oracle as O
from tests.golden_util import build_model, case_prob_rep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _concrete(cls, *args, **kwargs):
    sub = type("Concrete", (cls,), {})
    sub.__abstractmethods__ = frozenset()
    return sub(*args, **kwargs)


def test_abstract_hooks_raise_like_the_reference():
    """/root/reference/tests/test_base.py:22-45."""
    x = _concrete(base.RBN)
    for attr in ("inside_chart", "terminal_chart", "prior", "root_location"):
        with pytest.raises(NotImplementedError):
            getattr(x, attr)
    with pytest.raises(NotImplementedError):
        x.inside_schedule()
    with pytest.raises(NotImplementedError):
        x.update_inside_chart(None, None, None)
    assert x.init_inside() is None
    with pytest.raises(NotImplementedError):
        _concrete(base.Transition).inside_marginals(None, None, None)
    with pytest.raises(NotImplementedError):
        _concrete(base.Prior).marginal_likelihood(None, None)
    v = _concrete(base.NonTermVar)
    with pytest.raises(NotImplementedError):
        v.get_chart()
    with pytest.raises(NotImplementedError):
        v.mixture()
    c = _concrete(base.Cell, None)
    with pytest.raises(NotImplementedError):
        c.transitions()
    with pytest.raises(NotImplementedError):
        c.inside_mixture(None)


def test_tmap_layout_pinned_by_reference_tests():
    """/root/reference/tests/test_pcfg.py:14-19 (n=4 row order) and tests/test_util.py:18-22 (n=3, [0,2] -> row 1)."""
    t = TMap(np.arange(10))
    order = [(0, 4), (0, 3), (1, 4), (0, 2), (1, 3), (2, 4), (0, 1), (1, 2), (2, 3), (3, 4)]
    assert [t[s] for s in order] == list(range(10))
    assert TMap(np.arange(6))[0, 2] == 1
    assert TMap.size_from_n(7) == 28
    with pytest.raises(ValueError):
        TMap(np.arange(5))


def _binary_pcfg(rep):
    return pcfg.AbstractedPCFG(non_terminals="SAB", terminals="ab", start="S", rules=[
        ("S --> A B", 1), ("S --> B A", 1), ("A --> B A", 1), ("B --> A B", 1), ("A --> a", 1), ("B --> b", 1)],
        prob_rep=rep)


def test_parameter_names_shapes_dtypes_match_reference(golden):
    """state_dict keys are the reference's; the fixture's gradient keys ARE the reference's
    named_parameters()."""
    g = _binary_pcfg(util.Prob)
    ref_names = set(golden["binary/aaab"]["grads"])
    assert {n for n, _ in g.named_parameters()} == ref_names
    assert all(p.dtype == torch.float64 for p in g.parameters())
    shapes = {n: tuple(p.shape) for n, p in g.named_parameters()}
    assert shapes["_cells.0._transitions.1.transition_probabilities.p"] == (3, 3, 3)
    assert shapes["_cells.0._transitions.0.transition_probabilities.p"] == (2, 3)
    assert shapes["_cells.0.transition_probabilities.p"] == (2, 3)
    g2 = _binary_pcfg(util.LogProb)
    assert {n for n, _ in g2.named_parameters()} == {n[:-1] + "log_p" for n in ref_names}
    assert g.tokenise("ab") == [[0], [1]]
    with pytest.raises(KeyError):
        g.tokenise("abc")


def test_front_end_weights_match_reference(golden):
    """AbstractedPCFG rule parsing + normalisation reproduces the brick probabilities the reference built."""
    g = _binary_pcfg(util.Prob)
    c = golden["binary/aaab"]["cells"][0]
    np.testing.assert_array_equal(g._cells[0].transition_probabilities.p.detach().numpy(), c["struct"])
    np.testing.assert_array_equal(g._cells[0]._transitions[0].transition_probabilities.p.detach().numpy(), c["transitions"][0][1])
    np.testing.assert_array_equal(g._cells[0]._transitions[1].transition_probabilities.p.detach().numpy(), c["transitions"][1][1])
    # word grammar of examples/plot_PCFG.py / tests/test_base.py:250-295
    meta = golden["wordProb/0"]["meta"]["extra"]
    nts, terms = meta["non_terminals"], meta["terminals"]
    groups = [terms[0:4], terms[4:7], terms[9:12], terms[7:9], terms[15:19], terms[12:15]]
    rules = [("start --> subject verb", 1), ("verb --> verb_qualifier verb", 1), ("verb --> verb gradable_adverb", 1),
             ("verb --> verb non_gradable_adverb", 1), ("gradable_adverb --> grade gradable_adverb", 1),
             (("grade", ("grade", "grade")), 1)] + [(f"{n} --> {t}", 1) for n, ts in zip(nts[1:], groups) for t in ts]
    for rep, key in [(util.Prob, "wordProb/0"), (util.LogProb, "wordLogProb/0")]:
        wg = pcfg.AbstractedPCFG(non_terminals=nts, terminals=terms, rules=rules, start="start", prob_rep=rep)
        c = golden[key]["cells"][0]
        np.testing.assert_allclose(wg._cells[0].transition_probabilities.p.detach().numpy(), c["struct"], rtol=1e-15)
        np.testing.assert_allclose(wg._cells[0]._transitions[0].transition_probabilities.p.detach().numpy(), c["transitions"][0][1], rtol=1e-15)
        np.testing.assert_allclose(wg._cells[0]._transitions[1].transition_probabilities.p.detach().numpy(), c["transitions"][1][1], rtol=1e-15)


def test_constructor_validation_matches_reference_exceptions():
    with pytest.raises(ValueError):
        pcfg.DiscreteBinaryNonTerminalTransition(weights=np.ones((2, 2)))
    with pytest.raises(ValueError):
        pcfg.DiscreteBinaryNonTerminalTransition(weights=-np.ones((2, 2, 2)))
    with pytest.raises(ValueError):
        pcfg.DiscreteTerminalTransition(weights=np.array([[0.0, 1.0], [0.0, 1.0]]))
    with pytest.raises(ValueError):
        pcfg.DiscreteTerminalTransition(weights=np.ones(3))
    with pytest.raises(ValueError):
        pcfg.DiscretePrior(struc_weights=np.ones(2), prior_weights=[np.ones(2)])
    with pytest.raises(ValueError):
        pcfg.DiscretePrior(struc_weights=np.ones((1, 1)), prior_weights=[np.ones(2)])
    with pytest.raises(ValueError):
        pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(2), weights=np.ones((1, 2)), transitions=[])
    with pytest.raises(ValueError):
        pcfg.StaticCell(variable=pcfg.DiscreteNonTermVar(3), weights=np.ones(1),
                        transitions=[pcfg.DiscreteTerminalTransition(weights=np.ones((2, 3)))][:1] * 2)
    with pytest.raises(ValueError):
        pcfg.AbstractedPCFG("SA", "a", [("S -> A A", 1)], "S")
    with pytest.raises(ValueError):
        pcfg.AbstractedPCFG("SA", "a", [(("S", ("A", "A", "A")), 1)], "S")
    with pytest.raises(RuntimeError):      # a non-terminal without any rule
        pcfg.AbstractedPCFG("SA", "a", [("S --> A A", 1)], "S")
    with pytest.raises(ValueError):
        util.LogProb()
    with pytest.raises(TypeError):
        util.Prob(p=np.ones(3, dtype=np.int64))
    with pytest.raises(RuntimeError):
        util.Prob(p=np.zeros(3))
    with pytest.raises(ValueError):
        pcfg.DiscreteNonTermVar(2, chart_type="nope").get_chart(3)


def test_prob_and_logprob_constraints_and_gradient_projection():
    p = util.Prob(p=np.array([[1.0, 3.0], [1.0, 1.0]]), dim=0)
    np.testing.assert_allclose(p.p.detach().numpy(), [[0.5, 0.75], [0.5, 0.25]])
    (p.p * torch.tensor([[1.0, 0.0], [0.0, 0.0]])).sum().backward()
    np.testing.assert_allclose(p.p.grad.numpy(), [[0.5, 0.0], [-0.5, 0.0]])       # g - mean over dim 0
    lp = util.LogProb(p=np.array([1.0, 1.0, 2.0]))
    np.testing.assert_allclose(lp.p.detach().numpy(), [0.25, 0.25, 0.5])
    with torch.no_grad():
        lp.log_p.add_(1.0)
    lp.enforce_constraints()
    np.testing.assert_allclose(lp.p.detach().numpy(), [0.25, 0.25, 0.5])
    m = _binary_pcfg(util.LogProb)
    with torch.no_grad():
        for q in m.parameters():
            q.add_(0.3)
    m.enforce_constraints()
    W = m._cells[0]._transitions[1].transition_probabilities.p
    np.testing.assert_allclose(W.sum(dim=(0, 1)).detach().numpy(), np.ones(3), rtol=1e-12)
    assert m.remap(m._cells[0]._transitions[1].transition_probabilities.log_p) is not None


def test_normalize_non_zero():
    a = np.array([[1.0, 0.0], [3.0, 0.0]])
    out = util.normalize_non_zero(a.copy(), axis=0)
    np.testing.assert_allclose(out, [[0.25, 0.0], [0.75, 0.0]])
    out = util.normalize_non_zero(a.copy(), axis=0, make_zeros_uniform=True)
    np.testing.assert_allclose(out, [[0.25, 0.5], [0.75, 0.5]])
    t = util.normalize_non_zero(torch.tensor([[2.0, 2.0], [0.0, 0.0]]), axis=1, make_zeros_uniform=True)
    np.testing.assert_allclose(t.numpy(), [[0.5, 0.5], [0.5, 0.5]])
    np.testing.assert_allclose(util.normalize_non_zero(np.ones((2, 2)), axis=None), np.full((2, 2), 0.25))
    with pytest.raises(TypeError):
        util.normalize_non_zero(np.ones(3, dtype=int), axis=0)
    with pytest.raises(ValueError):
        util.normalize_non_zero(-np.ones(3), axis=0)


def test_flatten_matches_oracle_flatten(golden):
    """The product's differentiable flattening equals the oracle's restatement on every golden model."""
    for key, c in golden.items():
        m = build_model(c, case_prob_rep(c))
        W, T, pr = InsideEngine(m).flatten()
        Wo, To, po, off, toff = O.flatten_bricks(c["cells"], c["prior"])
        np.testing.assert_allclose(W.detach().numpy(), Wo.numpy(), rtol=1e-13, atol=0, err_msg=key)
        np.testing.assert_allclose(T.detach().numpy(), To.numpy(), rtol=1e-13, atol=0, err_msg=key)
        np.testing.assert_allclose(pr.detach().numpy(), po.numpy(), rtol=1e-13, atol=0, err_msg=key)
        tok, lengths = InsideEngine(m).pack_sequences([c["sequence"]])
        np.testing.assert_array_equal(tok[0].numpy(), O.flatten_sequence(c["sequence"], toff))
        assert int(lengths[0]) == len(c["sequence"])


def test_pack_sequences_ragged_and_numpy_positions():
    m = _binary_pcfg(util.Prob)
    e = InsideEngine(m)
    tok, lengths = e.pack_sequences([[[0], [1], [1]], np.array([[1], [0]]), [[1]]])
    assert tok.shape == (3, 3, 1) and lengths.tolist() == [3, 2, 1]
    assert tok[1, :, 0].tolist() == [1, 0, -1] and tok[2, :, 0].tolist() == [1, -1, -1]
    with pytest.raises(IndexError):
        e.pack_sequences([[[5]]])
    with pytest.raises(ValueError):
        e.pack_sequences([[]])


def test_host_hooks_refuse_to_compute_on_cpu():
    m = _binary_pcfg(util.Prob)
    cell = m._cells[0]
    with pytest.raises(NotImplementedError):
        cell._transitions[1].inside_marginals((0, 2), None, None)
    with pytest.raises(NotImplementedError):
        cell.inside_mixture([])
    with pytest.raises(NotImplementedError):
        m.prior.marginal_likelihood((0, 1), None)
    assert InsideEngine.supports(m)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure on a box without a GPU")
def test_inside_fails_loudly_without_gpu():
    m = _binary_pcfg(util.Prob)
    with pytest.raises(_lib.RbnError):
        m.inside(sequence="aaab")


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rbn_b200.h")).read()
    declared = set(re.findall(r"\b(rbn_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"librbn_b200.so does not export {name}"
    assert declared == set(_lib.exported_symbols())
    assert lib.rbn_abi_version() == _lib.ABI_VERSION == 3
    # descriptor validation happens without touching a device
    bad = _lib.RbnDesc(N=0, V=1, B=1, L=1, n_tvars=1, path=0, chunk_items=0, flags=0)
    assert lib.rbn_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"bad descriptor" in lib.rbn_last_error()
    ok = _lib.RbnDesc(N=10, V=19, B=1, L=20, n_tvars=1, path=0, chunk_items=0, flags=0)
    assert lib.rbn_workspace_bytes(ctypes.byref(ok)) > 0
    assert lib.rbn_resolve_path(ctypes.byref(ok)) == _lib.PATH_SIMT
    # the sorted-ragged entry point validates its HOST array before touching a device: active[w] must be non-increasing,
    # within [0, B], and B at w = 0, 1
    desc = _lib.RbnDesc(N=10, V=19, B=3, L=4, n_tvars=1, path=0, chunk_items=0, flags=0)
    fake = ctypes.c_void_p(1024)                     # never dereferenced: validation fails first
    for bad_active, msg in (([3, 3, 2, 3, 1], b"non-increasing"), ([3, 2, 2, 1, 0], b"must equal B"), ([3, 3, 4, 1, 0], b"non-increasing")):
        arr = (ctypes.c_int32 * 5)(*bad_active)
        rc = lib.rbn_inside_forward_sorted(ctypes.byref(desc), fake, fake, fake, fake, fake, fake, 1 << 40, fake, fake, fake,
                                           ctypes.cast(arr, ctypes.c_void_p), None)
        assert rc == -1 and msg in lib.rbn_last_error(), (bad_active, lib.rbn_last_error())


def test_engine_flattens_unmodified_reference_objects():
    """the engine works on the reference's own bricks (duck-typed)."""
    from oracle.reference_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present")
    ref = load_reference()
    rp, ru = ref["pcfg"], ref["util"]
    g = rp.AbstractedPCFG(non_terminals="SAB", terminals="ab", start="S", rules=[
        ("S --> A B", 1), ("S --> B A", 1), ("A --> B A", 1), ("B --> A B", 1), ("A --> a", 1), ("B --> b", 1)],
        prob_rep=ru.Prob)
    ours = _binary_pcfg(util.Prob)
    Wr, Tr, pr = InsideEngine(g).flatten()
    Wo, To, po = InsideEngine(ours).flatten()
    np.testing.assert_array_equal(Wr.detach().numpy(), Wo.detach().numpy())
    np.testing.assert_array_equal(Tr.detach().numpy(), To.detach().numpy())
    np.testing.assert_array_equal(pr.detach().numpy(), po.detach().numpy())
    tok, lengths = InsideEngine(g).pack_sequences([g.tokenise("aaab")])
    assert tok[0, :, 0].tolist() == [0, 0, 0, 1]


def _two_terminal_variable_model():
    """One nonterminal variable (3 values) emitting TWO terminal variables (alphabets of 2 and 4 symbols)."""
    rng = np.random.default_rng(0)
    cell = pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(3), weights=rng.random((3, 3)),
                             transitions=[pcfg.DiscreteTerminalTransition(weights=rng.random((2, 3)), term_idx=0),
                                          pcfg.DiscreteTerminalTransition(weights=rng.random((4, 3)), term_idx=1),
                                          pcfg.DiscreteBinaryNonTerminalTransition(weights=rng.random((3, 3, 3)))])
    prior = pcfg.DiscretePrior(struc_weights=np.ones(1), prior_weights=[rng.random(3)])
    return sequential.SequentialRBN(cells=[cell], prior=prior)


def test_batched_tokens_are_per_variable_values_and_range_checked():
    """log_inside_batch(tokens=...) takes PER-VARIABLE terminal values: they are shifted into the flat alphabet
    like pack_sequences does, and a value outside its own variable's alphabet raises IndexError instead of silently
    reading another variable's rows of T."""
    e = InsideEngine(_two_terminal_variable_model())
    seqs = [[[1, 3], [0, None], [1, 0]], [[0, 2]]]
    flat, lengths = e.pack_sequences(seqs)
    raw = torch.tensor([[[1, 3], [0, -1], [1, 0]], [[0, 2], [-1, -1], [-1, -1]]])
    packed, _ = e.pack_tokens(raw, lengths)
    assert packed.tolist() == flat.tolist()                      # second variable's values moved up by the first's 2 rows
    assert packed[0, 0].tolist() == [1, 2 + 3] and packed[0, 1].tolist() == [0, -1]
    with pytest.raises(IndexError):
        e.pack_tokens(torch.tensor([[[2, 0]]]))                  # variable 0 has only 2 symbols
    with pytest.raises(ValueError):
        e.pack_tokens(torch.tensor([[0, 1]]))                    # one value per position, the model has two variables
    # the flat functional entry checks the range against the rows of T before anything touches a device
    from rbn_b200 import inside_logZ
    W, T, pr = torch.rand(4, 4, 4), torch.rand(5, 4), torch.rand(4)
    with pytest.raises(IndexError):
        inside_logZ(W, T, pr, torch.tensor([[0, 5, 1]]))
    with pytest.raises(ValueError):
        inside_logZ(W, T, pr, torch.tensor([[0, 1, 1]]), lengths=torch.tensor([4]))


def test_very_long_sequences_are_not_padded_onto_the_tensor_core_path():
    from rbn_b200 import engine
    assert engine._padded_cardinality(40, 100000, L=128) == 64
    assert engine._padded_cardinality(40, 100000, L=130) == 64   # spans wider than 129 tokens: windows of 128 splits
    assert engine._padded_cardinality(40, 100000, L=1026) == 40  # beyond the cap of the tensor-core path
    big = _lib.RbnDesc(N=64, V=4, B=2, L=1026, n_tvars=1, path=_lib.PATH_TCGEN05, chunk_items=0, flags=0)
    lib = _lib.load()
    assert lib.rbn_resolve_path(ctypes.byref(big)) == -4         # RBN_E_UNSUPPORTED
    big.path = _lib.PATH_AUTO
    assert lib.rbn_resolve_path(ctypes.byref(big)) == _lib.PATH_SIMT
    ok = _lib.RbnDesc(N=64, V=4, B=2, L=129, n_tvars=1, path=_lib.PATH_AUTO, chunk_items=0, flags=0)
    assert lib.rbn_resolve_path(ctypes.byref(ok)) == _lib.PATH_TCGEN05
    # the split-sum cache only adds whole spans, never more than the batch has
    base = lib.rbn_workspace_bytes(ctypes.byref(ok))
    ok.cache_bytes = 10 ** 15
    full = lib.rbn_workspace_bytes(ctypes.byref(ok))
    spans = 2 * 129 * 128 // 2
    assert 0 < full - base - spans * (64 * 64 * 2 + 4 + 64 * 2 + 64 * 4) < 4096 + 1024
    ok.cache_bytes = -1
    assert lib.rbn_workspace_bytes(ctypes.byref(ok)) == 0
