"""Sparse consumers of the score matrix: per-utterance pdf subsets (forced alignment,
gmm-align-compiled.cpp:119-128) and (frame, pdf) arcs (lattice rescoring, lat/lattice-functions.cc:1214-1360), through the
C ABI.  Checked against the oracle's dense DiagGmm::LogLikelihoods + LogSumExp and against the dense GPU matrix (bit-equal:
the sparse forms only extract)."""
import numpy as np
import pytest

from tests.common import assert_ll_close
from tests.gpu_common import oracle_feats, oracle_mfcc_batch, oracle_stats
from voicebridge_b200 import capi, host, synth

pytestmark = pytest.mark.gpu


def _subsets(rng, P, n_utts, lo, hi):
    return [rng.choice(P, size=int(rng.integers(lo, hi + 1)), replace=False).astype(np.int32) for _ in range(n_utts)]


@pytest.mark.parametrize("kernel", [0, 1])
def test_subset_and_gather_vs_oracle(orc, kernel):
    m = synth.make_model(180, 1500, 39, 21)
    lens = [0, 37, 300, 1, 129, 256]                 # ragged, with an empty utterance
    fo = np.zeros(len(lens) + 1, np.int64)
    fo[1:] = np.cumsum(lens)
    X = synth.make_feats(m, int(fo[-1]), 22)
    am = host.AmDiagGmmGpu.from_model(m)
    am.set_kernel(kernel)
    rc, want = orc.gmm_loglikes(m, X)
    assert rc == 0
    dense = am.score(X)
    rng = np.random.default_rng(5)
    subs = _subsets(rng, 180, len(lens), 1, 60)
    subs[2] = np.arange(180, dtype=np.int32)[::-1]   # a whole-model subset, reversed order
    subs[4] = np.zeros(0, np.int32)                  # an utterance that asks for nothing
    got = am.score_subset(X, fo, subs)
    for u, g in enumerate(got):
        rows = slice(int(fo[u]), int(fo[u + 1]))
        assert g.shape == (lens[u], len(subs[u]))
        if g.size:
            assert np.array_equal(g, dense[rows][:, subs[u]])
            assert_ll_close(g, want[rows][:, subs[u]], what="subset of utterance %d" % u)
    n = 5000
    fr = rng.integers(0, int(fo[-1]), n).astype(np.int32)
    pd = rng.integers(0, 180, n).astype(np.int32)
    arcs = am.score_gather(X, fr, pd)
    assert np.array_equal(arcs, dense[fr, pd])
    assert_ll_close(arcs, want[fr, pd], what="arcs")
    assert am.score_gather(X, fr[:0], pd[:0]).shape == (0,)


def test_sparse_errors():
    m = synth.make_model(20, 80, 39, 3)
    am = host.AmDiagGmmGpu.from_model(m)
    X = synth.make_feats(m, 50, 4)
    with pytest.raises(capi.VbgpuError):     # pdf id out of range
        am.score_subset(X, [0, 50], [np.array([3, 20], np.int32)])
    with pytest.raises(capi.VbgpuError):     # offsets do not cover T
        am.score_subset(X, [0, 40], [np.array([3], np.int32)])
    with pytest.raises(capi.VbgpuError):     # frame out of range
        am.score_gather(X, np.array([50], np.int32), np.array([0], np.int32))


def test_sparse_many_slabs_at_model_scale(orc):
    """P = 4000 / N = 40000 with more frames than one slab of the dense matrix: every slab boundary is exercised; spot
    columns against the oracle on a sub-model, everything against the dense matrix of the same frames."""
    import torch
    m = synth.make_model(4000, 40000, 39, 11)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    T = 256 * (sms // 2) * 8 + 77            # > one 2 GiB slab of [frames x 4012] floats
    X = synth.make_feats(m, T, 12)
    am = host.AmDiagGmmGpu.from_model(m)
    rng = np.random.default_rng(1)
    n_utts = 40
    cuts = np.sort(rng.choice(np.arange(1, T), n_utts - 1, replace=False))
    fo = np.r_[0, cuts, T].astype(np.int64)
    subs = _subsets(rng, 4000, n_utts, 100, 300)
    got = am.score_subset(X, fo, subs)
    probe = [0, 7, 19, n_utts - 1]
    for u in probe:
        rows = slice(int(fo[u]), int(fo[u + 1]))
        dense = am.score(X[rows])
        assert np.array_equal(got[u], dense[:, subs[u]])
    fr = rng.integers(0, T, 20000).astype(np.int32)
    pd = rng.integers(0, 4000, 20000).astype(np.int32)
    arcs = am.score_gather(X, fr, pd)
    sel = np.argsort(fr)[:3000]              # the arcs of the first frames: compare with the oracle on a sub-model
    rows = np.unique(fr[sel])
    pdfs = np.unique(pd[sel])[:60]
    sub_off = np.zeros(len(pdfs) + 1, np.int32)
    idx = []
    for i, p in enumerate(pdfs):
        g0, g1 = m.pdf_offsets[p], m.pdf_offsets[p + 1]
        idx.extend(range(g0, g1))
        sub_off[i + 1] = sub_off[i] + (g1 - g0)
    idx = np.array(idx)
    sub = synth.GmmModel(sub_off, m.weights[idx], m.means[idx], m.iv[idx], m.miv[idx], m.gconsts[idx])
    rc, want = orc.gmm_loglikes(sub, X[rows])
    assert rc == 0
    row_pos = {int(r): i for i, r in enumerate(rows)}
    pdf_pos = {int(p): i for i, p in enumerate(pdfs)}
    k = [i for i in sel if int(pd[i]) in pdf_pos]
    assert len(k) > 20
    assert_ll_close(arcs[k], np.array([want[row_pos[int(fr[i])], pdf_pos[int(pd[i])]] for i in k]), what="arcs at scale")


def test_pipeline_subset_and_gather_from_pcm(orc):
    """PCM in, only the requested log-likelihoods out (the e2e_align path of bench.py)."""
    o = capi.default_mfcc_opts(dither=0.0, use_energy=0)
    n_spk = 3
    pcm, so, u2s = synth.make_corpus(n_spk, 2, 0.4, 0.9, 77)
    fopts = capi.default_feat_opts()
    fm = synth.make_fmllr(n_spk, 39, 6)
    mf_want, fo = oracle_mfcc_batch(orc, o, pcm, so)
    st = oracle_stats(orc, mf_want, fo, u2s, n_spk)
    feats_want = oracle_feats(orc, mf_want, fo, u2s, st, fopts, None, fm)
    model = synth.make_model_from_feats(feats_want, 60, 500, 3)
    rc, ll_want = orc.gmm_loglikes(model, feats_want)
    assert rc == 0
    mfcc, fp, am = host.Mfcc(o), host.FeaturePipeline(fopts, 13), host.AmDiagGmmGpu.from_model(model)
    pipe = host.ScoringPipeline(mfcc, fp, am)
    dense = pipe.score(pcm, so, u2s, n_spk, fmllr=fm)
    rng = np.random.default_rng(9)
    subs = _subsets(rng, 60, len(so) - 1, 5, 25)
    sub_o = np.zeros(len(subs) + 1, np.int64)
    sub_o[1:] = np.cumsum([len(x) for x in subs])
    out, oo = pipe.score_subset(pcm, so, sub_o, np.concatenate(subs), u2s, n_spk, fmllr=fm)
    for u in range(len(subs)):
        rows = slice(int(fo[u]), int(fo[u + 1]))
        blk = out[oo[u]:oo[u + 1]].reshape(int(fo[u + 1] - fo[u]), len(subs[u]))
        assert np.array_equal(blk, dense[rows][:, subs[u]])
        assert_ll_close(blk, ll_want[rows][:, subs[u]], what="pipeline subset, utterance %d" % u)
    fr = rng.integers(0, int(fo[-1]), 800).astype(np.int32)
    pd = rng.integers(0, 60, 800).astype(np.int32)
    arcs = pipe.score_gather(pcm, so, fr, pd, u2s, n_spk, fmllr=fm)
    assert np.array_equal(arcs, dense[fr, pd])


def test_transition_accumulators_share_the_reduce_buffer():
    """gmm-acc-stats-ali.cpp:92 keeps transition accumulators beside the GMM statistics and gmm-sum-accs.cpp:48 sums both:
    here they live behind tot_frames in the ONE buffer, so Add / the all-reduce cover them."""
    m = synth.make_model(30, 200, 39, 5)
    am = host.AmDiagGmmGpu.from_model(m)
    n_tids = 75
    rng = np.random.default_rng(2)
    tids = np.repeat(rng.integers(1, n_tids + 1, 400), rng.integers(1, 9, 400)).astype(np.int32)
    a, b = host.AccumAmDiagGmmGpu(am, num_tids=n_tids), host.AccumAmDiagGmmGpu(am, num_tids=n_tids)
    a.AccumulateTransitions(tids)
    a.AccumulateTransitions(tids[:100])
    want = np.bincount(tids, minlength=n_tids + 1) + np.bincount(tids[:100], minlength=n_tids + 1)
    assert np.array_equal(a.transition_accs(), want.astype(np.float64))
    X = synth.make_feats(m, 300, 6)
    ali = synth.make_alignment(30, 300, 7)
    b.AccumulateForUtterance(X, ali)
    b.AccumulateTransitions(tids[:50])
    a.Add(2.0, b)                                  # AccumAmDiagGmm::Add(scale, other) + the transition part
    assert np.array_equal(a.transition_accs(), want + 2.0 * np.bincount(tids[:50], minlength=n_tids + 1))
    ptr, n = a.buffer()
    assert n == m.num_gauss * (2 * 39 + 1) + 2 + n_tids + 1
    with pytest.raises(capi.VbgpuError):           # transition-id 0 / beyond NumTransitionIds()
        a.AccumulateTransitions(np.array([0, 3], np.int32))
    with pytest.raises(capi.VbgpuError):
        host.AccumAmDiagGmmGpu(am).AccumulateTransitions(tids)


def test_handle_scratch_is_ordered_across_streams(orc):
    """A handle's scratch (batch layout, frame2utt, statistics) is filled on the stream of the call that first sees a
    layout; the next call may come on ANOTHER stream with the same offsets (a cache hit).  The library orders the two with an
    event (common.h: StreamOrder): results on stream B right after stream A must equal the single-stream ones."""
    import torch
    o = capi.default_mfcc_opts(dither=0.0, use_energy=0)
    mfcc, fp = host.Mfcc(o), host.FeaturePipeline(capi.default_feat_opts(), 13)
    pcm, so, u2s = synth.make_corpus(4, 8, 2.0, 4.0, 31, fast=True)
    fo = mfcc.frame_offsets(so)
    T = int(fo[-1])
    mf, _ = mfcc.compute_batch(pcm[:int(so[4])], so[:5])
    fs = fp.run(mf, fo[:5], cmvn_stats=fp.cmvn_stats(mf, fo[:5]))
    model = synth.make_model_from_feats(fs, 200, 1500, 3)
    am = host.AmDiagGmmGpu.from_model(model)
    pipe = host.ScoringPipeline(mfcc, fp, am)
    dev = torch.device("cuda", 0)
    d_pcm = torch.from_numpy(pcm).to(dev)
    nc = am.NumCols()
    want = torch.empty((T, nc), dtype=torch.float32, device=dev)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(sa):
        pipe.score_cols_dev(d_pcm, so, u2s, 4, None, 0, want, nc, stream=sa)
    torch.cuda.synchronize()
    for trial in range(3):
        so2 = so.copy()
        if trial:  # a NEW layout is uploaded on stream A, then used from stream B immediately
            so2 = np.r_[so[:-1], so[-1] - 160 * trial]
        T2 = int(mfcc.frame_offsets(so2)[-1])
        a = torch.empty((T2, nc), dtype=torch.float32, device=dev)
        b = torch.empty((T2, nc), dtype=torch.float32, device=dev)
        pipe.score_cols_dev(d_pcm, so2, u2s, 4, None, 0, a, nc, stream=sa)
        pipe.score_cols_dev(d_pcm, so2, u2s, 4, None, 0, b, nc, stream=sb)
        torch.cuda.synchronize()
        assert torch.equal(a, b)
        if not trial:
            assert torch.equal(a, want)
