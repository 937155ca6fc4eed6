"""fMLLR sufficient statistics on the device against the float64 restatement, the plain-C oracle and
the reference's own FmllrDiagGmmAccs; the reference's solver must land on the same transform from either statistics."""
import numpy as np
import pytest

from tests.common import assert_fmllr_close, fmllr_truth
from voicebridge_b200 import capi, host, synth

pytestmark = pytest.mark.gpu


def _model(chk, P, N, D, seed):
    m = synth.make_model(P, N, D, seed)
    gc, miv, iv = chk.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    return synth.GmmModel(m.pdf_offsets, m.weights, m.means, iv, miv, gc)


@pytest.mark.parametrize("D,weighted,T", [(39, False, 1200), (40, True, 1500), (13, True, 700), (39, True, 1)])
def test_fmllr_stats_single_speaker(orc, D, weighted, T):
    m = _model(orc, 25, 180, D, 51)
    X = synth.make_feats(m, T, 52)
    ali = synth.make_alignment(25, T, 53)
    w = np.random.default_rng(54).uniform(0.2, 1.0, T).astype(np.float32) if weighted else None
    acc = host.FmllrDiagGmmAccsGpu(host.AmDiagGmmGpu.from_model(m))
    tl = acc.AccumulateForUtterances(X, ali, weights=w)
    rc, ob, oK, oG, ol = orc.fmllr_acc(m, X, ali, w)
    assert rc == 0 and abs(tl - ol) <= 1e-4 * abs(ol)
    assert_fmllr_close(acc.stats(0), fmllr_truth(m, X, ali, w), what="GPU")
    # a second call adds; SetZero clears
    acc.AccumulateForUtterances(X, ali, weights=w)
    b2, K2, G2 = acc.stats(0)
    assert abs(b2 - 2 * ob) <= 1e-4 * max(ob, 1.0) and np.abs(G2 - 2 * oG).max() <= 2e-4 * np.abs(oG).max()
    acc.SetZero()
    b0, K0, G0 = acc.stats(0)
    assert b0 == 0.0 and not K0.any() and not G0.any()


def test_fmllr_stats_many_speakers_vs_reference(ref):
    """A packed batch of ragged utterances from 5 speakers (one of them silent): every speaker's statistics equal the
    reference's FmllrDiagGmmAccs run over that speaker's frames alone, and FmllrDiagGmmAccs::Update gives the same
    transform from the device statistics as from its own."""
    D, P = 39, 40
    m = _model(ref, P, 300, D, 61)
    rng = np.random.default_rng(62)
    lens = [300, 1, 257, 0, 900, 512, 77, 640, 256, 1100]
    u2s = np.array([0, 0, 1, 1, 2, 2, 4, 4, 4, 0], np.int32)  # speaker 3 has no data
    fo = np.zeros(len(lens) + 1, np.int64)
    fo[1:] = np.cumsum(lens)
    T = int(fo[-1])
    X = synth.make_feats(m, T, 63)
    ali = synth.make_alignment(P, T, 64)
    acc = host.FmllrDiagGmmAccsGpu(host.AmDiagGmmGpu.from_model(m), n_spk=5)
    acc.AccumulateForUtterances(X, ali, frame_offsets=fo, utt2spk=u2s)
    for s in range(5):
        rows = np.concatenate([np.arange(fo[u], fo[u + 1]) for u in range(len(lens)) if u2s[u] == s] or [np.zeros(0, int)])
        b, K, G = acc.stats(s)
        if len(rows) == 0:
            assert b == 0.0 and not K.any() and not G.any()
            continue
        assert_fmllr_close((b, K, G), fmllr_truth(m, X[rows], ali[rows]), what="GPU, speaker %d" % s)
        rc, rb, rK, rG, _ = ref.fmllr_acc(m, X[rows], ali[rows])
        assert rc == 0
        r1, x_gpu, i1, c1 = ref.fmllr_update(b, K, G)
        r2, x_ref, i2, c2 = ref.fmllr_update(rb, rK, rG)
        assert r1 == 0 and r2 == 0 and abs(c1 - c2) <= 1e-3 * c2
        assert np.abs(x_gpu - x_ref).max() <= 1e-3
        if c2 > 500:
            assert np.abs(x_ref - np.eye(D, D + 1)).max() > 1e-3 and abs(i1 - i2) <= 1e-3 * max(abs(i2), 1.0)


def test_fmllr_stats_errors(orc):
    m = _model(orc, 5, 20, 39, 71)
    am = host.AmDiagGmmGpu.from_model(m)
    acc = host.FmllrDiagGmmAccsGpu(am, n_spk=2)
    X = synth.make_feats(m, 50, 72)
    ali = synth.make_alignment(5, 50, 73)
    bad = ali.copy()
    bad[7] = 99  # invalid pdf id: reported, the frame adds nothing
    with pytest.raises(capi.VbgpuError) as e:
        acc.AccumulateForUtterances(X, bad)
    assert e.value.code == capi.ERR_NUMERIC
    with pytest.raises(capi.VbgpuError):  # speaker outside [0, n_spk)
        acc.AccumulateForUtterances(X, ali, frame_offsets=[0, 50], utt2spk=[2])
    with pytest.raises(capi.VbgpuError):  # frame_offsets beyond T
        acc.AccumulateForUtterances(X, ali, frame_offsets=[0, 60], utt2spk=[0])
    big = synth.make_model(3, 6, 41, 74)
    with pytest.raises(capi.VbgpuError):  # D > 40
        host.FmllrDiagGmmAccsGpu(host.AmDiagGmmGpu.from_model(big))


def test_fmllr_stats_full_size_properties():
    """BASELINE cfg 3 model (P=4000, N=40000, D=39), 200 000 frames, 16 speakers: size-independent properties.
    beta_s = frames of speaker s (posteriors sum to one); the statistics of two calls over disjoint speaker sets equal
    those of one call over everything; G[i](D, D) = sum_t b_t[i] is a sum of positive inverse variances."""
    D, P, n_spk = 39, 4000, 16
    m = synth.make_model(P, 40000, D, 11)
    rng = np.random.default_rng(5)
    lens = rng.integers(900, 1600, size=160)
    fo = np.zeros(len(lens) + 1, np.int64)
    fo[1:] = np.cumsum(lens)
    T = int(fo[-1])
    u2s = np.repeat(np.arange(n_spk, dtype=np.int32), len(lens) // n_spk)
    X = synth.make_feats(m, T, 12)
    ali = synth.make_alignment(P, T, 13)
    am = host.AmDiagGmmGpu.from_model(m)
    one = host.FmllrDiagGmmAccsGpu(am, n_spk=n_spk)
    one.AccumulateForUtterances(X, ali, frame_offsets=fo, utt2spk=u2s)
    two = host.FmllrDiagGmmAccsGpu(am, n_spk=n_spk)
    half = len(lens) // 2  # a speaker boundary: 80 utterances = 8 speakers
    two.AccumulateForUtterances(X[:fo[half]], ali[:fo[half]], frame_offsets=fo[:half + 1], utt2spk=u2s[:half])
    two.AccumulateForUtterances(X[fo[half]:], ali[fo[half]:], frame_offsets=fo[half:] - fo[half], utt2spk=u2s[half:])
    jj, kk = np.tril_indices(D + 1)
    last = np.flatnonzero((jj == D) & (kk == D))[0]
    for s in range(n_spk):
        b1, K1, G1 = one.stats(s)
        b2, K2, G2 = two.stats(s)
        frames = int(lens[u2s == s].sum())
        assert abs(b1 - frames) <= 1e-5 * frames
        assert abs(b1 - b2) <= 1e-9 * b1
        assert np.abs(K1 - K2).max() <= 1e-9 * np.abs(K1).max() and np.abs(G1 - G2).max() <= 1e-9 * np.abs(G1).max()
        assert (G1[:, last] > 0).all()  # sum_t b_t[i] of positive inverse variances
        assert np.isfinite(G1).all() and np.isfinite(K1).all()


# ------------------------------------------------------------------------------------------ MLLT statistics (gmm-acc-mllt)
@pytest.mark.parametrize("D,weighted,T", [(40, False, 900), (13, True, 700), (39, True, 1)])
def test_mllt_stats(orc, ref, D, weighted, T):
    """MlltAccs on the device (the fMLLR G contraction over one pseudo-frame per (frame, Gaussian)) against the float64
    restatement, the oracle, and through the reference's own MlltAccs::Update."""
    from tests.common import mllt_truth
    m = _model(orc, 20, 150, D, 3)
    X = synth.make_feats(m, T, 4)
    ali = synth.make_alignment(20, T, 5)
    w = np.random.default_rng(1).uniform(0.2, 1.0, T).astype(np.float32) if weighted else None
    acc = host.MlltAccsGpu(host.AmDiagGmmGpu.from_model(m))
    tl = acc.AccumulateForUtterance(X, ali, w)
    rc, ob, oG, ol = orc.mllt_acc(m, X, ali, w)
    assert rc == 0 and abs(tl - ol) <= 1e-4 * abs(ol)
    tb, tG, SG = mllt_truth(m, X, ali, w)
    assert abs(acc.beta - tb) <= 1e-4 * max(tb, 1.0)
    assert (np.abs(acc.G - tG) / np.maximum(SG, 1e-30)).max() <= 1e-4
    if T >= 700:
        r1, M1, i1, c1 = ref.mllt_update(acc.beta, acc.G)
        r2, M2, i2, c2 = ref.mllt_update(ob, oG)
        assert r1 == 0 and r2 == 0 and np.abs(M1 - M2).max() <= 1e-3 and abs(i1 - i2) <= 1e-3 * max(abs(i2), 1.0)
    acc.AccumulateForUtterance(X, ali, w)  # a second call adds
    assert abs(acc.beta - 2 * tb) <= 2e-4 * max(tb, 1.0)


def test_mllt_stats_many_slabs_and_bad_ids(orc):
    m = _model(orc, 30, 900, 39, 8)  # ~30 Gaussians per pdf: 60 000 frames make several row slabs
    T = 60000
    X = synth.make_feats(m, T, 9)
    ali = synth.make_alignment(30, T, 10)
    am = host.AmDiagGmmGpu.from_model(m)
    whole = host.MlltAccsGpu(am)
    whole.AccumulateForUtterance(X, ali)
    parts = host.MlltAccsGpu(am)
    for a, b in ((0, 7), (7, 25000), (25000, T)):
        parts.AccumulateForUtterance(X[a:b], ali[a:b])
    assert abs(whole.beta - T) <= 1e-5 * T and abs(whole.beta - parts.beta) <= 1e-9 * T
    assert np.abs(whole.G - parts.G).max() <= 1e-6 * np.abs(whole.G).max()
    bad = ali[:100].copy()
    bad[3] = -1
    with pytest.raises(capi.VbgpuError) as e:
        host.MlltAccsGpu(am).AccumulateForUtterance(X[:100], bad)
    assert e.value.code == capi.ERR_NUMERIC


def test_component_posteriors(orc):
    """gmm-post-to-gpost: Gaussian-level posteriors of every frame's aligned pdf, pdfs of 1 .. 100 Gaussians."""
    m = _model(orc, 40, 900, 39, 61)
    T = 500
    X = synth.make_feats(m, T, 62)
    ali = synth.make_alignment(40, T, 63)
    w = np.random.default_rng(64).uniform(0.2, 1.0, T).astype(np.float32)
    am = host.AmDiagGmmGpu.from_model(m)
    for weights in (None, w):
        post, offs, ll = am.ComponentPosteriors(X, ali, m.pdf_offsets, weights)
        rc, po_, oo, lo = orc.component_posteriors(m, X, ali, weights)
        assert rc == 0 and np.array_equal(offs, oo)
        assert np.abs(post - po_).max() <= 1e-5 and np.abs(ll - lo).max() <= 1e-3
        sums = np.add.reduceat(post, offs[:-1])
        assert np.abs(sums - (1.0 if weights is None else weights)).max() <= 1e-5
