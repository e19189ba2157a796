"""GPU: frame-synchronous operation (jb200_stream_*) and the sliced batch pipeline.

A stream is the reference's call sequence get_back_trellis_init / _proceed(t)... / _end (pass1.c:112-254): the utterance
arrives piece by piece and its length is only known at the end.  Whatever the feed sizes, the word trellis must be the
one the compiled reference produced for the whole utterance (golden fixtures), atom for atom and bit for bit; the same
holds for a batch cut into time slices (scoring of slice c+1 beside the token passing of slice c)."""
import numpy as np
import pytest

from julius_b200 import capi
from util import CASES, DNN_CASES, Golden, atoms_equal

pytestmark = pytest.mark.gpu


def _check(r, u):
    assert r["overflow"] == 0
    ok, why = atoms_equal(r["atoms"], u.atoms)
    assert ok, why
    assert r["status"] == u.status
    assert r["words"] == u.words
    assert np.float32(r["score"]) == np.float32(u.score)


def _pieces(T, sizes):
    """cut [0, T) into pieces of the given sizes, repeating the last size"""
    out, t, i = [], 0, 0
    while t < T:
        n = min(sizes[min(i, len(sizes) - 1)], T - t)
        out.append((t, t + n)); t += n; i += 1
    return out


@pytest.mark.parametrize("case", CASES + ["small_dfa", "small_tr"])
@pytest.mark.parametrize("sizes", [[1], [7], [1, 2, 64], [1000]])
def test_one_stream_equals_the_reference_whatever_the_feed_size(case, sizes):
    g = Golden(case)
    am = capi.GmmScorer(g.ds, mode=capi.GMM_EXACT)
    dec = capi.Decoder(g.ds, am, max_utts=2, max_frames=4096)
    for x, u in list(zip(g.feats, g.utts))[:2]:
        dec.stream_open(1)
        pcs = _pieces(len(x), sizes)
        for k, (a, b) in enumerate(pcs):
            dec.stream_feed([x[a:b]], last=[k == len(pcs) - 1])
            st = dec.stream_status(0)
            assert st["frames"] == b and st["ended"] == (k == len(pcs) - 1)
        _check(dec.stream_result(0), u)


@pytest.mark.parametrize("case", ["small_b100", "small_mp"])
def test_end_of_input_may_come_without_frames(case):
    """real-time input: the host learns that the utterance is over after the last frame was handed in"""
    g = Golden(case)
    am = capi.GmmScorer(g.ds, mode=capi.GMM_EXACT)
    dec = capi.Decoder(g.ds, am, max_utts=1, max_frames=2048)
    x, u = g.feats[0], g.utts[0]
    dec.stream_open(1)
    for a in range(0, len(x), 50):
        dec.stream_feed([x[a:a + 50]])
    dec.stream_feed([None], last=[1])
    _check(dec.stream_result(0), u)


@pytest.mark.parametrize("case", ["small_b100", "small_iwsp"])
def test_streams_advance_independently_and_restart(case):
    g = Golden(case)
    n = len(g.feats)
    am = capi.GmmScorer(g.ds, mode=capi.GMM_EXACT)
    dec = capi.Decoder(g.ds, am, max_utts=n, max_frames=n * 1024)
    dec.stream_open(n)
    pos = [0] * n
    step = [3 + 5 * s for s in range(n)]               # every stream at its own pace
    done = [False] * n
    while not all(done):
        chunks, last = [], []
        for s in range(n):
            if done[s]:
                chunks.append(None); last.append(0); continue
            a, b = pos[s], min(pos[s] + step[s], len(g.feats[s]))
            chunks.append(g.feats[s][a:b]); last.append(1 if b == len(g.feats[s]) else 0)
            pos[s] = b
        dec.stream_feed(chunks, last=last)
        for s in range(n):
            if last[s]:
                done[s] = True
                _check(dec.stream_result(s), g.utts[s])
    # the next utterance on stream 0 (another one than before), while stream 1 is abandoned half way and restarted
    dec.stream_restart(0)
    dec.stream_feed([g.feats[1][:40]] + [None] * (n - 1))
    dec.stream_restart(0)                              # abandoned: its node slots must be wiped
    dec.stream_feed([g.feats[1]] + [None] * (n - 1), last=[1] + [0] * (n - 1))
    _check(dec.stream_result(0), g.utts[1])


@pytest.mark.parametrize("case", DNN_CASES)
def test_dnn_stream(case):
    """DNN-HMM scoring per feed (K2 on a handful of frames at a time) under the 1e-4 tolerance: same words, same trellis
    structure as the one-batch decode of the same vectors on the device"""
    g = Golden(case)
    am = capi.GmmScorer(g.ds, gmm_desc=g.ds.cd_only_gmm())
    dnn = capi.DnnScorer(g.ds)
    dec = capi.Decoder(g.ds, am, max_utts=1, max_frames=2048)
    dec.attach_dnn(dnn)
    x = g.feats[0]
    whole = dec.decode([x])[0]
    dec.stream_open(1)
    pcs = _pieces(len(x), [1, 9, 33])
    for k, (a, b) in enumerate(pcs):
        dec.stream_feed([x[a:b]], last=[k == len(pcs) - 1])
    r = dec.stream_result(0)
    ok, why = atoms_equal(r["atoms"], whole["atoms"])
    assert ok, why
    assert r["words"] == whole["words"]


@pytest.mark.parametrize("case", ["small_b100", "small_mp"])
def test_interim_result_is_the_best_word_ending_at_the_last_frame(case):
    """bt_current_max (beam.c:876-921): best trellis word of end time t-1 after frame t, traced back to the start"""
    g = Golden(case)
    am = capi.GmmScorer(g.ds, mode=capi.GMM_EXACT)
    dec = capi.Decoder(g.ds, am, max_utts=1, max_frames=2048)
    x, u = g.feats[0], g.utts[0]
    dec.stream_open(1)
    seen = 0
    for a in range(0, len(x) - 1, 10):
        b = min(a + 10, len(x) - 1)
        dec.stream_feed([x[a:b]], interim=True)
        p = dec.stream_partial(0)
        assert p["frame"] == b - 2
        at = u.atoms[u.atoms["end"] == b - 2]
        if len(at) == 0:
            assert p["words"] == []
            continue
        mx = at["backscore"].max()
        assert np.float32(p["score"]) == np.float32(mx)
        cands = np.nonzero((u.atoms["end"] == b - 2) & (u.atoms["backscore"] == mx))[0]
        seqs = []
        for c in cands:
            w, k = [], int(c)
            while True:
                w.append(int(u.atoms["wid"][k]))
                if u.atoms["begin"][k] <= 0:
                    break
                k = int(u.atoms["last"][k])
                if k < 0:
                    break
            seqs.append(w[::-1])
        assert p["words"] in seqs
        seen += 1
    assert seen > 3
    dec.stream_feed([x[len(x) - 1:]], last=[1])
    _check(dec.stream_result(0), u)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("frames", [16, 100])
def test_sliced_batch_pipeline_is_bit_identical(case, frames):
    g = Golden(case)
    am = capi.GmmScorer(g.ds, mode=capi.GMM_EXACT)
    dec = capi.Decoder(g.ds, am, max_utts=8, max_frames=8192)
    dec.set_pipeline(frames)
    lens = [len(x) for x in g.feats]
    feats = list(g.feats) + [g.feats[0][:frames], g.feats[0][:frames + 1], g.feats[0][:1]]      # ragged: slice edges
    res = dec.decode(feats)
    assert dec.pipeline_info()["slices"] == (max(lens) + frames - 1) // frames
    for r, u in zip(res, g.utts):
        _check(r, u)
    from oracle import ffi
    for r, n in zip(res[len(g.utts):], (frames, frames + 1, 1)):
        o = ffi.beam_decode(g.ds, g.utts[0].outprob[:n])
        ok, why = atoms_equal(r["atoms"], o["atoms"])
        assert ok, f"T={n}: {why}"
    # and again on the same decoder, unsliced
    dec.set_pipeline(0)
    for r, u in zip(dec.decode(g.feats), g.utts):
        _check(r, u)
