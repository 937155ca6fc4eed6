"""Shared tolerances (BASELINE.json north_star) and helpers for the parity tests."""
import numpy as np

FEAT_RTOL = 1e-4    # features within 1e-4 relative
LL_ATOL = 1e-3      # log-likelihoods within 1e-3 absolute
STATS_RTOL = 1e-4   # statistics within 1e-4 relative

# Raw element-wise maxima seen by the assert_* helpers of this session, BEFORE any of the documented floors / scales is
# applied: tests/conftest.py prints them after the run (and every failure message carries its own), so that the tolerances
# below can be read against unprocessed numbers.  key -> (max abs err, max element-wise rel err, magnitude at that element).
RAW = {}
USED = {}   # key -> largest asserted error as a fraction of its tolerance (1.0 = at the limit)


def _used(kind, frac):
    USED[kind] = max(USED.get(kind, 0.0), float(frac))


def _record(kind, a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(a) & np.isfinite(b)
    if not fin.any():
        return "no finite elements"
    d = np.abs(a - b)[fin]
    rel = d / np.maximum(np.abs(b[fin]), 1e-300)
    i = int(np.argmax(rel))
    prev = RAW.get(kind, (0.0, 0.0, 0.0, 0))
    RAW[kind] = (max(prev[0], float(d.max())), max(prev[1], float(rel[i])),
                 float(np.abs(b[fin])[i]) if rel[i] >= prev[1] else prev[2], prev[3] + 1)
    return "raw: max abs err %.3g, max element-wise rel err %.3g (at |ref| = %.3g)" % (d.max(), rel[i], np.abs(b[fin])[i])


def assert_feats_close(a, b, rtol=FEAT_RTOL, what="features"):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    # "1e-4 relative": relative to the element, floored at that coefficient's natural scale (its RMS over the
    # utterance, at least 1.0).  Cepstra are sums of ~23 log-mel terms of magnitude 10..20 that cancel, so an element
    # that happens to land near zero still carries the rounding noise of its O(10) terms: the compiled reference and
    # its plain-C restatement already differ by 1.1e-4 "elementwise-relative" on feat/test_data/test.wav, both in FP32.
    raw = _record("features", a, b)
    scale = np.maximum(np.sqrt((b * b).mean(axis=0, keepdims=True)), 1.0) if b.ndim == 2 else 1.0
    err = np.abs(a - b) / np.maximum(np.abs(b), scale)
    i = np.unravel_index(np.argmax(err), err.shape)
    _used("features", err.max() / rtol)
    assert err.max() <= rtol, "%s: max rel err %.3g at %s (%r vs %r); %s" % (what, err.max(), i, a[i], b[i], raw)


def assert_ll_close(a, b, atol=LL_ATOL, what="loglikes"):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), what + ": finiteness pattern differs"
    err = np.abs(a[fin] - b[fin])
    raw = _record("loglikes", a, b)
    if err.size:
        _used("loglikes", err.max() / atol)
    assert err.size == 0 or err.max() <= atol, "%s: max abs err %.3g (tolerance %.1g) at max |ll| %.4g; %s" % (
        what, err.max(), atol, np.abs(b[fin]).max(), raw)


def assert_stats_close(a, b, rtol=STATS_RTOL, what="stats"):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    # relative to the element, floored at 1e-3 of the largest statistic (sums that cancel to ~0 carry the
    # rounding of their O(max) terms)
    raw = _record("statistics", a, b)
    floor = 1e-3 * np.abs(b).max() + 1e-300
    err = np.abs(a - b) / np.maximum(np.abs(b), floor)
    _used("statistics", err.max() / rtol)
    assert err.max() <= rtol, "%s: max rel err %.3g; %s" % (what, err.max(), raw)


def fmllr_truth(model, X, ali, weights=None):
    """float64 restatement of the fMLLR statistics (fmllr-diag-gmm.cc:30-45,562-583) and, per element, the sum of the
    ABSOLUTE accumulated terms.  K and G entries are signed sums of hundreds of O(10..100) terms that cancel to O(1), and
    the reference forms the posteriors (hence a, b) in FP32: "within 1e-4 relative" is measured against the magnitude
    that was accumulated, the only scale an FP32 implementation (the reference itself included) can be accurate to.
    Returns beta, K, G, scale_K, scale_G with G in SpMatrix packing."""
    X = np.asarray(X, np.float64)
    T, D = X.shape
    gc, miv, iv = (np.asarray(v, np.float64) for v in (model.gconsts, model.miv, model.iv))
    jj, kk = np.tril_indices(D + 1)  # row-major lower triangle = SpMatrix packing
    K, SK = np.zeros((D, D + 1)), np.zeros((D, D + 1))
    G, SG = np.zeros((D, len(jj))), np.zeros((D, len(jj)))
    beta = 0.0
    for t in range(T):
        g0, g1 = model.pdf_offsets[ali[t]], model.pdf_offsets[ali[t] + 1]
        x = X[t]
        ll = gc[g0:g1] + miv[g0:g1] @ x - 0.5 * (iv[g0:g1] @ (x * x))
        post = np.exp(ll - ll.max())
        post *= (1.0 if weights is None else float(weights[t])) / post.sum()
        a, b = post @ miv[g0:g1], post @ iv[g0:g1]
        xp = np.append(x, 1.0)
        z = xp[jj] * xp[kk]
        beta += post.sum()
        K += np.outer(a, xp)
        SK += np.outer(np.abs(a), np.abs(xp))
        G += np.outer(b, z)
        SG += np.outer(b, np.abs(z))
    return beta, K, G, SK, SG


def assert_fmllr_close(got, truth, rtol=STATS_RTOL, what="fMLLR stats"):
    """got = (beta, K, G); truth = fmllr_truth(...)."""
    beta, K, G = got
    tb, tK, tG, SK, SG = truth
    assert abs(beta - tb) <= rtol * max(tb, 1.0), "%s: beta %r vs %r" % (what, beta, tb)
    eK = (np.abs(np.asarray(K) - tK) / np.maximum(SK, 1e-30)).max()
    eG = (np.abs(np.asarray(G) - tG) / np.maximum(SG, 1e-30)).max()
    assert eK <= rtol, "%s: K off by %.3g of the accumulated magnitude" % (what, eK)
    assert eG <= rtol, "%s: G off by %.3g of the accumulated magnitude" % (what, eG)


def mllt_truth(model, X, ali, weights=None):
    """float64 restatement of MlltAccs::AccumulateFromPosteriors (mllt.cc:131-160, rand_prune = 0): beta, G[D, D(D+1)/2]
    and the per-element sum of absolute accumulated terms (see fmllr_truth for why)."""
    X = np.asarray(X, np.float64)
    T, D = X.shape
    gc, miv, iv = (np.asarray(v, np.float64) for v in (model.gconsts, model.miv, model.iv))
    rr, cc = np.tril_indices(D)
    G, SG = np.zeros((D, len(rr))), np.zeros((D, len(rr)))
    beta = 0.0
    for t in range(T):
        g0, g1 = model.pdf_offsets[ali[t]], model.pdf_offsets[ali[t] + 1]
        x = X[t]
        ll = gc[g0:g1] + miv[g0:g1] @ x - 0.5 * (iv[g0:g1] @ (x * x))
        post = np.exp(ll - ll.max())
        post *= (1.0 if weights is None else float(weights[t])) / post.sum()
        off = miv[g0:g1] / iv[g0:g1] - x            # [M, D]
        z = off[:, rr] * off[:, cc]                  # [M, pairs]
        a = iv[g0:g1] * post[:, None]                # [M, D]
        G += a.T @ z
        SG += a.T @ np.abs(z)
        beta += post.sum()
    return beta, G, SG


def recipe_opts(po_or_capi, **kw):
    """The recipes' MFCC config: Kaldi defaults + --use-energy=false, and --dither=0 for parity."""
    base = dict(dither=0.0, use_energy=0)
    base.update(kw)
    if hasattr(po_or_capi, "default_mfcc_opts"):
        return po_or_capi.default_mfcc_opts(**base)
    return po_or_capi.default_opts(**base)


def copy_opts(src, dst_cls):
    """Copy an MfccOpts between the oracle's and the library's ctypes classes (identical layout)."""
    d = dst_cls()
    for name, _ in src._fields_:
        setattr(d, name, getattr(src, name))
    return d


def assert_acc_close(a, b, rtol=STATS_RTOL, what="acc"):
    """EM statistics (occ, mean_acc, var_acc) within `rtol` relative.  occ and var_acc are sums of positive terms and
    are judged elementwise.  mean_acc = sum_t gamma*x cancels in sign, so an element is judged relative to the mass of
    its terms, bounded by Cauchy-Schwarz: |sum gamma x| <= sqrt(sum gamma * sum gamma x^2)."""
    occ_a, mean_a, var_a = [np.asarray(v, np.float64) for v in a]
    occ_b, mean_b, var_b = [np.asarray(v, np.float64) for v in b]
    raw = "; ".join("%s %s" % (n, _record("EM " + n, x, y)) for n, x, y in (("occ", occ_a, occ_b), ("mean", mean_a, mean_b),
                                                                            ("var", var_a, var_b)))
    tiny = 1e-6 * max(occ_b.max(), 1e-300)
    e = np.abs(occ_a - occ_b) / np.maximum(occ_b, tiny)
    assert e.max() <= rtol, "%s occ: max rel err %.3g; %s" % (what, e.max(), raw)
    vfloor = np.maximum(var_b, tiny * np.abs(var_b).max() / max(occ_b.max(), 1e-300))
    e = np.abs(var_a - var_b) / np.maximum(vfloor, 1e-300)
    assert e.max() <= rtol, "%s var: max rel err %.3g; %s" % (what, e.max(), raw)
    mass = np.sqrt(np.maximum(occ_b, tiny)[:, None] * vfloor)
    e = np.abs(mean_a - mean_b) / np.maximum(np.maximum(np.abs(mean_b), mass), 1e-300)
    assert e.max() <= rtol, "%s mean: max rel err %.3g; %s" % (what, e.max(), raw)


def assert_pitch_close(got, want, what="pitch", max_diff_frac=0.10, max_rel=0.02, nccf_atol=1e-4):
    """(NCCF, pitch) rows of ComputeKaldiPitch.  The pitch column is a discrete Viterbi path over lag states 0.5% apart:
    where two paths are tied to within FP32 rounding (noise-only stretches) a different summation order picks the
    neighbouring state for a run of frames — the compiled reference and its plain-C restatement already do that to each
    other (BLAS dot products vs sequential sums).  So: identical states on at least 90% of the frames, never more than
    2% (four states) apart, and the NCCF within 1e-4 absolute wherever the state is the same."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return
    rel = np.abs(got[:, 1] - want[:, 1]) / want[:, 1]
    same = rel <= 1e-6
    assert (~same).mean() <= max_diff_frac, "%s: %d of %d frames on a different lag state" % (what, (~same).sum(), len(same))
    assert rel.max() <= max_rel, "%s: pitch differs by %.3g relative at frame %d" % (what, rel.max(), int(rel.argmax()))
    if same.any():
        err = np.abs(got[same, 0] - want[same, 0])
        assert err.max() <= nccf_atol, "%s: NCCF differs by %.3g" % (what, err.max())
    return int((~same).sum())


def assert_process_pitch_close(got, want, what="processed pitch", atol=2e-5, rtol=1e-4):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return
    err = np.abs(got - want) - rtol * np.abs(want)
    i = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= atol, "%s: |diff| %.3g at %s (%r vs %r)" % (what, np.abs(got - want)[i], i, got[i], want[i])


def random_pitch_opts(rng):
    """A random but valid PitchExtractionOptions (resample.cc:42-47 / pitch-functions.cc:715-766 constraints hold)."""
    sf = float(rng.choice([8000, 16000, 22050, 44100]))
    rf = float(rng.choice([2000, 3000, 4000, 5000]))
    return dict(samp_freq=sf, resample_freq=rf, lowpass_cutoff=float(rng.uniform(0.2, 0.5) * rf),
                lowpass_filter_width=int(rng.integers(1, 4)), upsample_filter_width=int(rng.choice([3, 5, 7])),
                min_f0=float(rng.uniform(40, 80)), max_f0=float(rng.uniform(250, 500)),
                delta_pitch=float(rng.choice([0.005, 0.01, 0.02])), penalty_factor=float(rng.uniform(0.05, 0.3)),
                nccf_ballast=float(rng.uniform(1000, 10000)), soft_min_f0=float(rng.uniform(5, 30)),
                frame_shift_ms=float(rng.choice([5.0, 10.0, 12.5])), frame_length_ms=float(rng.choice([20.0, 25.0, 32.0])),
                snip_edges=int(rng.integers(0, 2)), preemph_coeff=float(rng.choice([0.0, 0.3])),
                recompute_frame=int(rng.choice([20, 500])))
