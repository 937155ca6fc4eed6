"""The plain-C oracle against the committed golden vectors (runs anywhere; no GPU, no /root/reference)."""
import numpy as np
import pytest

from tests.common import assert_acc_close, assert_feats_close, assert_ll_close, assert_stats_close
from oracle import pyoracle as po
from voicebridge_b200 import synth
from tests.golden.make_golden import HTK_CASES, HTK_VTLN


@pytest.mark.parametrize("case", sorted(HTK_CASES))
def test_htk_known_answers(orc, golden, case):
    """UnitTestHTKCompare1..6 (feat/feature-mfcc-test.cc:112-650): MFCC + deltas of test.wav vs HTK's HCopy output,
    |diff| <= 1.0, 10 frames skipped at each end — the reference's own known-answer test for this path."""
    htk, _ = golden
    o = po.default_opts(dither=0.0)
    for k, v in HTK_CASES[case].items():
        setattr(o, k, v)
    raw = orc.mfcc(o, htk["pcm"].astype(np.float32), HTK_VTLN.get(case, 1.0))
    feats = orc.deltas(raw, 2, 2)
    want = htk["htk%d" % case]
    assert feats.shape == want.shape
    assert np.abs(feats[10:-10] - want[10:-10]).max() <= 1.0


def test_oracle_matches_reference_dumps(orc, golden):
    """Every step of the path against outputs dumped from the compiled reference (tests/golden/make_golden.py)."""
    htk, g = golden
    wav = htk["pcm"].astype(np.float32)
    assert_feats_close(orc.mfcc(po.default_opts(dither=0.0, use_energy=0), wav), g["mfcc16"], what="mfcc16")
    assert_feats_close(orc.mfcc(po.default_opts(dither=0.0), wav), g["mfcc16_energy"], what="mfcc16_energy")
    assert_feats_close(orc.mfcc(po.default_opts(dither=0.0, use_energy=0, snip_edges=0), wav), g["mfcc16_nosnip"])
    assert_feats_close(orc.mfcc(po.default_opts(dither=0.0, use_energy=0), wav, 0.9), g["mfcc16_vtln09"])
    assert_feats_close(orc.mfcc(po.default_opts(dither=0.0, use_energy=0, samp_freq=8000.0),
                                g["wave8"].astype(np.float32)), g["mfcc8"], what="mfcc8")
    x = g["mfcc16"]
    st = orc.cmvn_acc(x)
    assert_stats_close(st, g["cmvn_stats"], 1e-12, "cmvn stats")
    assert_feats_close(orc.cmvn_apply(g["cmvn_stats"], x, False), g["cmvn_mean"], 1e-6)
    assert_feats_close(orc.cmvn_apply(g["cmvn_stats"], x, True), g["cmvn_meanvar"], 1e-6)
    assert_feats_close(orc.deltas(g["cmvn_mean"], 2, 2), g["delta"], 1e-6)
    assert_feats_close(orc.deltas(g["cmvn_mean"], 1, 3), g["delta_o1_w3"], 1e-6)
    sp = orc.splice(g["cmvn_mean"], 3, 3)
    assert_feats_close(orc.transform(sp, g["lda_mat"]), g["lda"], 1e-5)
    assert_feats_close(orc.transform(sp, g["lda_aff_mat"]), g["lda_aff"], 1e-5)
    assert_feats_close(orc.transform(g["delta"], g["fmllr_mat"]), g["fmllr"], 1e-5)


def _golden_model(g):
    return synth.GmmModel(g["pdf_offsets"], g["weights"], g["means"], g["iv"], g["miv"], g["gconsts"])


def test_oracle_scoring_and_stats_match_reference_dumps(orc, golden):
    _, g = golden
    m = _golden_model(g)
    gc, miv, iv = orc.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    assert np.abs(gc - g["gconsts"]).max() <= 1e-4 and np.array_equal(miv, g["miv"])
    rc, ll = orc.gmm_loglikes(m, g["delta"])
    assert rc == 0
    assert_ll_close(ll, g["loglikes"], 2e-4)
    rc, occ, mean, var, tl, tf = orc.acc_ali(m, g["delta"], g["ali"])
    assert rc == 0
    assert_acc_close((occ, mean, var), (g["acc_occ"], g["acc_mean"], g["acc_var"]))
    assert abs(tl - g["acc_tot"][0]) <= 1e-6 * abs(g["acc_tot"][0]) and tf == g["acc_tot"][1]
    rc, occ, mean, var, tl, tf = orc.acc_ali(m, g["delta"], g["ali"], g["ali_w"], g["fmllr"])
    assert rc == 0
    assert_acc_close((occ, mean, var), (g["acc2_occ"], g["acc2_mean"], g["acc2_var"]))
    assert abs(tl - g["acc2_tot"][0]) <= 1e-6 * abs(g["acc2_tot"][0])
    assert abs(tf - g["acc2_tot"][1]) <= 1e-6 * g["acc2_tot"][1]


def test_oracle_edge_cases(orc):
    o = po.default_opts(dither=0.0, use_energy=0)
    assert orc.num_frames(399, o) == 0 and orc.num_frames(400, o) == 1 and orc.num_frames(559, o) == 1
    assert orc.num_frames(560, o) == 2
    o2 = po.default_opts(dither=0.0, use_energy=0, snip_edges=0)
    assert orc.num_frames(100, o2) == 1 and orc.num_frames(79, o2) == 0
    assert orc.mfcc(o, np.zeros(100, np.float32)).shape == (0, 13)
    # silence: every mel energy is floored at FLT_EPSILON -> finite, identical rows
    z = orc.mfcc(o, np.zeros(1000, np.float32))
    assert np.isfinite(z).all() and np.abs(z - z[0]).max() == 0.0
    # a pdf whose only Gaussian has zero weight scores -inf and is reported (KALDI_ERR in the reference)
    m = synth.make_model(3, 6, 5, 1)
    gc = m.gconsts.copy()
    gc[m.pdf_offsets[1]:m.pdf_offsets[2]] = -np.inf
    m2 = synth.GmmModel(m.pdf_offsets, m.weights, m.means, m.iv, m.miv, gc)
    rc, ll = orc.gmm_loglikes(m2, synth.make_feats(m, 4, 2))
    assert rc == -2 and (~np.isfinite(ll[:, 1])).all() and np.isfinite(ll[:, [0, 2]]).all()


def test_pitch_oracle_matches_reference_dumps(orc):
    """tests/golden/pitch_golden.npz was written by the compiled reference (tests/golden/make_pitch_golden.py)."""
    import os
    from tests.common import assert_pitch_close, assert_process_pitch_close
    from tests.golden.make_pitch_golden import PROCESS_VARIANT
    from oracle import pyoracle as po
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = dict(np.load(os.path.join(d, "pitch_golden.npz")))
    pcm = np.load(os.path.join(d, "htk_golden.npz"))["pcm"].astype(np.float32)
    assert_pitch_close(orc.pitch(po.default_pitch_opts(), pcm), g["raw16"], "test.wav", nccf_atol=1e-5)
    assert_pitch_close(orc.pitch(po.default_pitch_opts(snip_edges=0), pcm), g["raw16_nosnip"], "no snip", nccf_atol=1e-5)
    assert_pitch_close(orc.pitch(po.default_pitch_opts(), pcm[:9000]), g["raw16_short"], "short", nccf_atol=1e-5)
    o8 = po.default_pitch_opts(samp_freq=8000.0, min_f0=60.0, max_f0=350.0)
    assert_pitch_close(orc.pitch(o8, g["wave8"].astype(np.float32)), g["raw8"], "8 kHz", nccf_atol=1e-5)
    assert_process_pitch_close(orc.process_pitch(po.default_process_pitch_opts(), g["raw16"]), g["proc16"])
    assert_process_pitch_close(orc.process_pitch(po.default_process_pitch_opts(**PROCESS_VARIANT), g["raw16"]),
                               g["proc16_variant"])


def test_downsample_oracle_matches_reference_dumps(orc):
    """tests/golden/downsample_golden.npz was written by the compiled reference's DownsampleWaveForm and ComputeFeatures
    (tests/golden/make_downsample_golden.py): pins oracle.c:orc_downsample_waveform where oracle/_ref is not at hand."""
    import os
    from oracle import pyoracle as po
    from tests.golden.make_downsample_golden import PAIRS, noise
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = dict(np.load(os.path.join(d, "downsample_golden.npz")))
    pcm = np.load(os.path.join(d, "htk_golden.npz"))["pcm"].astype(np.float32)
    cases = [(16000, 8000, pcm, "speech_16k_to_8k"), (16000, 11025, pcm, "speech_16k_to_11025")]
    cases += [(o, n_, noise(n, n), "noise_%d_to_%d_%d" % (o, n_, n)) for o, n_, n in PAIRS]
    for orig, new, w, key in cases:
        got, want = orc.downsample_waveform(orig, new, w), g[key]
        assert got.shape == want.shape, key
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), key
    o = po.default_opts(dither=0.0, samp_freq=8000.0)
    assert_feats_close(orc.mfcc(o, orc.downsample_waveform(16000, 8000, pcm)), g["mfcc_of_speech_8k"],
                       what="MFCC of the down-sampled test.wav")
