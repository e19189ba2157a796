"""CPU: the restatement (oracle/restate) must reproduce the compiled reference's outputs held
in tests/golden bit for bit -- this is what pins the oracle (prompt section 3)."""
import numpy as np
import pytest

from util import CASES, DNN_CASES, ORACLE_ONLY_CASES, Golden, atoms_equal


@pytest.mark.parametrize("case", CASES + ORACLE_ONLY_CASES)
def test_gmm_restatement_bit_exact(case, oracle_lib):
    g = Golden(case)
    for u, x in zip(g.utts, g.feats):
        sc = oracle_lib.gmm_score(g.ds, x)
        assert sc.shape == u.outprob.shape
        assert np.array_equal(sc.view(np.uint32), u.outprob.view(np.uint32))


@pytest.mark.parametrize("case", CASES + ORACLE_ONLY_CASES)
def test_beam_restatement_identical_trellis(case, oracle_lib):
    g = Golden(case)
    for u in g.utts:
        r = oracle_lib.beam_decode(g.ds, u.outprob)
        ok, why = atoms_equal(r["atoms"], u.atoms)
        assert ok, why
        assert r["status"] == u.status
        assert r["words"] == u.words
        assert np.float32(r["score"]) == np.float32(u.score)


def test_addlog_table_matches_definition(oracle_lib):
    t = oracle_lib.addlog_table()
    i = np.array([0, 1, 1000, 250000, 499999])
    f = -(np.float32(15) * i.astype(np.float32) / np.float32(500000))
    want = np.log(1 + np.exp(f.astype(np.float64))).astype(np.float32)
    assert np.array_equal(t[i], want)


def test_cdset_scores_consistent_with_states(oracle_lib):
    g = Golden("small_b100")
    st = g.utts[0].outprob
    cd = oracle_lib.cdset_score(g.ds, st)
    off, ids = g.blob["am.cd_off"], g.blob["am.cd_states"]
    # N-best average (N=3) of a set is bounded by its max and its mean of top-3 in float64
    for c in range(0, cd.shape[1], 37):
        v = np.sort(st[5, ids[off[c]:off[c + 1]]])[::-1][:3]
        assert abs(cd[5, c] - v.astype(np.float64).mean()) < 1e-3


@pytest.mark.parametrize("case", DNN_CASES)
def test_dnn_restatement_bit_exact_and_beam_on_dnn_scores(case, oracle_lib):
    """DNN-HMM: the restatement of dnn_calc_outprob (x86 FMA GEMV order, logistic table, addlog
    softmax, prior) equals the compiled reference bit for bit; the beam restatement runs on it."""
    g = Golden(case)
    for u, x in zip(g.utts, g.feats):
        sc = oracle_lib.dnn_score(g.ds, x)
        assert np.array_equal(sc.view(np.uint32), u.outprob.view(np.uint32))
        r = oracle_lib.beam_decode(g.ds, u.outprob)
        ok, why = atoms_equal(r["atoms"], u.atoms)
        assert ok, why
        assert r["words"] == u.words and r["status"] == u.status


def test_dnn_restatement_at_full_shape_agrees_with_fp64(oracle_lib):
    """The checker of the full-shape K2 test (tests/test_gpu_dnn.py) is the restatement of dnn_calc_outprob run at
    528 -> 7 x 2048 -> 3000; here it is held against a float64 numpy forward pass with the exact logistic (the
    reference's table logistic and fp32 FMA accumulation differ from it by a few 1e-5 in log10 units)."""
    import numpy as np
    from julius_b200 import desc, synth
    from util import full_dnn_blob
    blob = full_dnn_blob()
    ds = desc.Descriptors(blob)
    x = synth.sample_dnn_input(np.random.default_rng(1), 6, 528)
    got = oracle_lib.dnn_score(ds, x)
    h = x.astype(np.float64)
    L = int(blob["dnn.n_layers"][0])
    for i in range(L):
        w = blob[f"dnn.l{i}.w"].reshape(int(blob[f"dnn.l{i}.out"][0]), int(blob[f"dnn.l{i}.in"][0])).astype(np.float64)
        h = h @ w.T + blob[f"dnn.l{i}.b"].astype(np.float64)
        if i < L - 1:
            h = 1.0 / (1.0 + np.exp(-h))
    m = h.max(1, keepdims=True)
    want = (h - (m + np.log(np.exp(h - m).sum(1, keepdims=True)))) / np.log(10.0) - blob["dnn.state_prior"].astype(np.float64)
    assert np.abs(got - want).max() < 2e-4
