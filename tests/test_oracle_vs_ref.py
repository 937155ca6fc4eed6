"""Pins the plain-C oracle against the reference's own code compiled from /root/reference (oracle/_ref/libvbref.so).
Skipped where that library has not been built."""
import numpy as np
import pytest

from tests.common import (assert_acc_close, assert_feats_close, assert_fmllr_close, assert_ll_close, assert_stats_close,
                          fmllr_truth)
from oracle import pyoracle as po
from voicebridge_b200 import synth

VARIANTS = [
    dict(), dict(use_energy=1), dict(use_energy=1, raw_energy=0), dict(use_energy=1, energy_floor=1e9),
    dict(snip_edges=0), dict(samp_freq=8000.0), dict(htk_compat=1), dict(htk_compat=1, use_energy=1),
    dict(window_type=1), dict(window_type=2), dict(window_type=3), dict(window_type=4),
    dict(remove_dc_offset=0), dict(preemph_coeff=0.0), dict(cepstral_lifter=0.0), dict(num_bins=30, num_ceps=20),
    dict(low_freq=100.0, high_freq=-400.0), dict(htk_mode=1), dict(frame_length_ms=20.0, frame_shift_ms=5.0),
    dict(samp_freq=44100.0),
]


@pytest.mark.parametrize("kw", VARIANTS, ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()) or "default")
def test_mfcc(orc, ref, kw):
    o = po.default_opts(dither=0.0, use_energy=0)
    for k, v in kw.items():
        setattr(o, k, v)
    w = synth.make_wave(int(o.samp_freq * 1.3), 5, o.samp_freq).astype(np.float32)
    a, b = orc.mfcc(o, w), ref.mfcc(o, w)
    assert a.shape == b.shape and a.shape[0] > 0
    assert_feats_close(a, b)
    assert orc.num_frames(len(w), o) == ref.num_frames(len(w), o)


@pytest.mark.parametrize("warp", [0.85, 0.9, 1.1, 1.2])
def test_mfcc_vtln(orc, ref, warp):
    o = po.default_opts(dither=0.0, use_energy=0)
    w = synth.make_wave(16000, 6).astype(np.float32)
    assert_feats_close(orc.mfcc(o, w, warp), ref.mfcc(o, w, warp))
    oa, la, wa = orc.mel_banks(o, warp)
    ob, lb, wb = ref.mel_banks(o, warp)
    assert np.array_equal(oa, ob) and np.array_equal(la, lb) and np.abs(wa - wb).max() < 1e-6


def test_tables(orc, ref):
    for fs in (8000.0, 16000.0):
        o = po.default_opts(samp_freq=fs)
        oa, la, wa = orc.mel_banks(o)
        ob, lb, wb = ref.mel_banks(o)
        assert np.array_equal(oa, ob) and np.array_equal(la, lb) and np.array_equal(wa, wb)
        assert int(la.sum()) == (480 if fs == 16000.0 else 241)
        for wt in range(5):
            o.window_type = wt
            assert np.array_equal(orc.window_table(o), ref.window_table(o))


@pytest.mark.parametrize("n", [0, 1, 79, 80, 199, 200, 399, 400, 401, 559, 560, 16000, 123457])
def test_num_frames(orc, ref, n):
    for snip in (0, 1):
        for fs in (8000.0, 16000.0):
            o = po.default_opts(snip_edges=snip, samp_freq=fs)
            assert orc.num_frames(n, o) == ref.num_frames(n, o)


def test_short_and_edge_utterances(orc, ref):
    o = po.default_opts(dither=0.0, use_energy=0, snip_edges=0)
    for n in (81, 150, 399, 400, 401, 1000):  # shorter than a window: reflection is applied repeatedly
        w = synth.make_wave(n, n).astype(np.float32)
        a, b = orc.mfcc(o, w), ref.mfcc(o, w)
        assert a.shape == b.shape
        assert_feats_close(a, b)


def test_feature_pipeline_steps(orc, ref):
    o = po.default_opts(dither=0.0, use_energy=0)
    x = orc.mfcc(o, synth.make_wave(16000 * 2, 3).astype(np.float32))
    st = orc.cmvn_acc(x)
    assert_stats_close(st, ref.cmvn_acc(x), 1e-12)
    for nv in (False, True):
        assert_feats_close(orc.cmvn_apply(st, x, nv), ref.cmvn_apply(st, x, nv), 1e-6)
    c = orc.cmvn_apply(st, x)
    for order, window in ((2, 2), (1, 2), (2, 3), (3, 1), (0, 2)):
        assert_feats_close(orc.deltas(c, order, window), ref.deltas(c, order, window), 1e-6)
    for l, r in ((3, 3), (4, 4), (0, 2), (5, 0)):
        assert np.array_equal(orc.splice(c, l, r), ref.splice(c, l, r))
    sp = orc.splice(c, 3, 3)
    for cols in (91, 92):
        m = synth.make_lda(40, cols, cols)
        assert_feats_close(orc.transform(sp, m), ref.transform(sp, m), 1e-5)
    # very short utterance: every delta/splice tap is clamped
    for T in (1, 2, 5):
        assert_feats_close(orc.deltas(c[:T]), ref.deltas(c[:T]), 1e-6)
        assert np.array_equal(orc.splice(c[:T], 3, 3), ref.splice(c[:T], 3, 3))


@pytest.mark.parametrize("P,N,D,seed", [(11, 60, 39, 1), (50, 400, 39, 2), (30, 200, 40, 3), (7, 7, 13, 4)])
def test_scoring_and_stats(orc, ref, P, N, D, seed):
    m = synth.make_model(P, N, D, seed)
    g_ref, miv_ref, iv_ref = ref.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    g_orc, miv_orc, _ = orc.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    assert np.array_equal(miv_ref, miv_orc) and np.abs(g_ref - g_orc).max() <= 1e-4
    m = synth.GmmModel(m.pdf_offsets, m.weights, m.means, iv_ref, miv_ref, g_ref)
    X = synth.make_feats(m, 150, seed + 10)
    rc1, a = orc.gmm_loglikes(m, X)
    rc2, b = ref.gmm_loglikes(m, X)
    rc3, c = ref.gmm_loglikes_matrix(m, X)
    assert rc1 == 0 and rc2 == 0 and rc3 == 0
    assert_ll_close(a, b, 2e-4)
    assert_ll_close(c, b, 2e-4, "matrix form vs decodable form")
    ali = synth.make_alignment(P, 150, seed)
    w = np.random.default_rng(seed).uniform(0.1, 1.0, 150).astype(np.float32)
    for weights, f2 in ((None, None), (w, None), (w, synth.make_feats(m, 150, seed + 20))):
        r1 = orc.acc_ali(m, X, ali, weights, f2)
        r2 = ref.acc_ali(m, X, ali, weights, f2)
        assert r1[0] == 0 and r2[0] == 0
        assert_acc_close(r1[1:4], r2[1:4])
        assert abs(r1[4] - r2[4]) <= 1e-6 * abs(r2[4]) and abs(r1[5] - r2[5]) <= 1e-6 * r2[5]


# ------------------------------------------------------------------------------------------ fMLLR statistics (§8f n1)
@pytest.mark.parametrize("D,weighted", [(39, False), (40, True), (13, True)])
def test_fmllr_stats(orc, ref, D, weighted):
    m = synth.make_model(25, 180, D, 51)
    gc, miv, iv = ref.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    m = synth.GmmModel(m.pdf_offsets, m.weights, m.means, iv, miv, gc)
    T = 1200  # FmllrOptions::min_count is 500 (fmllr-diag-gmm.h:45)
    X = synth.make_feats(m, T, 52)
    ali = synth.make_alignment(25, T, 53)
    w = np.random.default_rng(54).uniform(0.2, 1.0, T).astype(np.float32) if weighted else None
    ra, ba, Ka, Ga, la = orc.fmllr_acc(m, X, ali, w)
    rb, bb, Kb, Gb, lb = ref.fmllr_acc(m, X, ali, w)
    assert ra == 0 and rb == 0
    assert abs(la - lb) <= 1e-4 * abs(lb)
    truth = fmllr_truth(m, X, ali, w)
    assert_fmllr_close((bb, Kb, Gb), truth, what="compiled reference")
    assert_fmllr_close((ba, Ka, Ga), truth, what="oracle")
    # ... and the reference's own solver lands on the same transform from either set of statistics
    r1, x1, i1, c1 = ref.fmllr_update(ba, Ka, Ga)
    r2, x2, i2, c2 = ref.fmllr_update(bb, Kb, Gb)
    assert r1 == 0 and r2 == 0 and c1 > 0
    assert np.abs(x1 - x2).max() <= 1e-3
    assert np.abs(x1 - np.eye(D, D + 1)).max() > 1e-3  # the update really moved the transform


# ------------------------------------------------------------------------------------------ filterbank front end (§8f n4)
FBANK_VARIANTS = [
    dict(), dict(use_energy=1), dict(use_energy=1, htk_compat=1), dict(use_energy=1, raw_energy=0, energy_floor=1e9),
    dict(samp_freq=8000.0), dict(num_bins=30), dict(htk_mode=1), dict(snip_edges=0), dict(low_freq=100.0, high_freq=-400.0),
]


@pytest.mark.parametrize("use_log,use_power", [(1, 1), (0, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("kw", FBANK_VARIANTS, ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()) or "default")
def test_fbank(orc, ref, kw, use_log, use_power):
    o = po.default_opts(dither=0.0, use_energy=0)
    for k, v in kw.items():
        setattr(o, k, v)
    w = synth.make_wave(int(o.samp_freq * 1.1), 8, o.samp_freq).astype(np.float32)
    a, b = orc.fbank(o, w, 1.0, use_log, use_power), ref.fbank(o, w, 1.0, use_log, use_power)
    assert a.shape == b.shape == (orc.num_frames(len(w), o), o.num_bins + (1 if o.use_energy else 0))
    assert_feats_close(a, b, what="fbank")
    if "low_freq" not in kw:  # (vtln_low = 100 must lie above low_freq, mel-computations.cc:152-224)
        assert_feats_close(orc.fbank(o, w, 0.9, use_log, use_power), ref.fbank(o, w, 0.9, use_log, use_power),
                           what="fbank vtln")


# ------------------------------------------------------------------------------------------------ PLP front end (§8f n4)
PLP_VARIANTS = [
    dict(), dict(use_energy=1), dict(use_energy=1, htk_compat=1), dict(samp_freq=8000.0, num_bins=15), dict(cepstral_lifter=0.0),
    dict(htk_mode=1), dict(snip_edges=0), dict(num_ceps=9), dict(use_energy=1, raw_energy=0, energy_floor=1e9),
]


@pytest.mark.parametrize("kw", PLP_VARIANTS, ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()) or "default")
def test_plp(orc, ref, kw):
    o = po.default_opts(dither=0.0, use_energy=0)
    for k, v in kw.items():
        setattr(o, k, v)
    w = synth.make_wave(int(o.samp_freq * 1.1), 8, o.samp_freq).astype(np.float32)
    for extra in (dict(), dict(lpc_order=14, compress_factor=0.5, cepstral_scale=10.0)):
        a, b = orc.plp(o, w, 1.0, **extra), ref.plp(o, w, 1.0, **extra)
        assert a.shape == b.shape == (orc.num_frames(len(w), o), o.num_ceps)
        assert_feats_close(a, b, what="plp")
    assert_feats_close(orc.plp(o, w, 0.9), ref.plp(o, w, 0.9), what="plp vtln")


# ------------------------------------------------------------------------------------------ MLLT statistics (§8f n1)
@pytest.mark.parametrize("D,weighted", [(40, False), (13, True)])
def test_mllt_stats(orc, ref, D, weighted):
    from tests.common import mllt_truth
    m = synth.make_model(20, 150, D, 3)
    gc, miv, iv = ref.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    m = synth.GmmModel(m.pdf_offsets, m.weights, m.means, iv, miv, gc)
    T = 700
    X = synth.make_feats(m, T, 4)
    ali = synth.make_alignment(20, T, 5)
    w = np.random.default_rng(1).uniform(0.2, 1.0, T).astype(np.float32) if weighted else None
    ra, ba, Ga, la = orc.mllt_acc(m, X, ali, w)
    rb, bb, Gb, lb = ref.mllt_acc(m, X, ali, w)
    assert ra == 0 and rb == 0 and abs(la - lb) <= 1e-5 * abs(lb)
    tb, tG, SG = mllt_truth(m, X, ali, w)
    for name, (b, G) in (("oracle", (ba, Ga)), ("compiled reference", (bb, Gb))):
        assert abs(b - tb) <= 1e-4 * tb, name
        assert (np.abs(G - tG) / np.maximum(SG, 1e-30)).max() <= 1e-4, name
    r1, M1, i1, c1 = ref.mllt_update(ba, Ga)
    r2, M2, i2, c2 = ref.mllt_update(bb, Gb)
    assert r1 == 0 and r2 == 0 and np.abs(M1 - M2).max() <= 1e-3 and np.abs(M2 - np.eye(D)).max() > 1e-3


def test_component_posteriors(orc, ref):
    m = synth.make_model(25, 200, 39, 61)
    gc, miv, iv = ref.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    m = synth.GmmModel(m.pdf_offsets, m.weights, m.means, iv, miv, gc)
    X = synth.make_feats(m, 300, 62)
    ali = synth.make_alignment(25, 300, 63)
    w = np.random.default_rng(64).uniform(0.2, 1.0, 300).astype(np.float32)
    ra, pa, oa, la = orc.component_posteriors(m, X, ali, w)
    rb, pb, ob, lb = ref.component_posteriors(m, X, ali, w)
    assert ra == 0 and rb == 0 and np.array_equal(oa, ob)
    assert np.abs(pa - pb).max() <= 1e-5 and np.abs(la - lb).max() <= 1e-3


# ---- Kaldi pitch (feat/pitch-functions.cc, feat/resample.cc) ------------------------------------------------------
PITCH_CASES = [
    (3.0, dict()), (7.3, dict()), (0.31, dict()), (0.11, dict()), (0.02, dict()), (2.0, dict(snip_edges=0)),
    (2.0, dict(preemph_coeff=0.5)), (3.0, dict(samp_freq=8000.0)), (2.0, dict(min_f0=60.0, max_f0=300.0, delta_pitch=0.01)),
    (6.0, dict(recompute_frame=100)), (6.0, dict(recompute_frame=100000)), (2.0, dict(frame_shift_ms=5.0, frame_length_ms=20.0)),
    (2.0, dict(resample_freq=3000.0, lowpass_cutoff=800.0, upsample_filter_width=3, lowpass_filter_width=2)),
    (2.0, dict(nccf_ballast=100.0, soft_min_f0=30.0, penalty_factor=0.3)), (1.5, dict(samp_freq=22050.0)),
]


@pytest.mark.parametrize("secs,kw", PITCH_CASES, ids=lambda v: str(v))
def test_pitch_restatement_vs_compiled_reference(orc, ref, secs, kw):
    from tests.common import assert_pitch_close
    o = po.default_pitch_opts(**kw)
    w = synth.make_pitch_wave(int(secs * o.samp_freq), 31 + int(secs * 10), o.samp_freq).astype(np.float32)
    a, b = orc.pitch(o, w), ref.pitch(o, w)
    assert len(a) == len(b) == orc.lib.orc_pitch_num_frames(po.C.byref(o), po.C.c_int64(len(w)))
    assert_pitch_close(a, b, what="oracle vs reference %s" % kw, nccf_atol=1e-5)


def test_pitch_energy_correction_path(orc, ref):
    """A loud burst in the last samples moves the mean-square energy by more than 1% between the first call and the
    flush: RecomputeBacktraces rescales every frame of the first call (pitch-functions.cc:961-1002)."""
    from tests.common import assert_pitch_close
    o = po.default_pitch_opts()
    w = synth.make_pitch_wave(16000, 5).astype(np.float32) * 0.3
    w[-12:] = 32000.0
    a, b = orc.pitch(o, w), ref.pitch(o, w)
    assert_pitch_close(a, b, what="energy correction", nccf_atol=1e-5)
    w[-12:] = 0.0  # sanity: the burst really changes the early frames, i.e. the path was exercised
    quiet = ref.pitch(o, w)
    assert (quiet[:90, 1] != b[:90, 1]).sum() > 10


@pytest.mark.parametrize("kw", [dict(), dict(delay=3, add_raw_log_pitch=1, normalization_left_context=10, delta_window=3),
                                dict(add_pov_feature=0, add_delta_pitch=0), dict(pov_offset=1.0, pitch_scale=1.0)],
                         ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()) or "default")
def test_process_pitch_restatement_vs_compiled_reference(orc, ref, kw):
    from tests.common import assert_process_pitch_close
    raw = ref.pitch(po.default_pitch_opts(), synth.make_pitch_wave(16000 * 4, 77).astype(np.float32))
    pp = po.default_process_pitch_opts(**kw)
    assert_process_pitch_close(orc.process_pitch(pp, raw), ref.process_pitch(pp, raw))
    one = raw[:1]
    assert_process_pitch_close(orc.process_pitch(pp, one), ref.process_pitch(pp, one))


@pytest.mark.parametrize("seed", range(24))
def test_pitch_random_options_vs_compiled_reference(orc, ref, seed):
    from tests.common import assert_pitch_close, random_pitch_opts
    rng = np.random.default_rng(1000 + seed)
    kw = random_pitch_opts(rng)
    o = po.default_pitch_opts(**kw)
    w = synth.make_pitch_wave(int(rng.uniform(0.2, 3.0) * o.samp_freq), seed, o.samp_freq).astype(np.float32)
    assert_pitch_close(orc.pitch(o, w), ref.pitch(o, w), what=str(kw), nccf_atol=1e-5)


@pytest.mark.parametrize("orig,new,n", [(16000, 8000, 16000), (44100, 16000, 30001), (48000, 16000, 5000), (22050, 16000, 12345),
                                        (16000, 8000, 7), (11025, 8000, 3000)])
def test_downsample_waveform_port_matches_the_reference(orc, ref, orig, new, n):
    """oracle.c:orc_downsample_waveform against the reference's DownsampleWaveForm (feat/resample.cc:368-376)."""
    w = (np.random.default_rng(n).standard_normal(n) * 3000).astype(np.float32)
    a, b = orc.downsample_waveform(orig, new, w), ref.downsample_waveform(orig, new, w)
    assert len(a) == len(b)
    assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
