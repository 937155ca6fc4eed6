"""Bring-up aid for the tensor-core scorer: compares it with the FP32 SIMT kernel and the CPU oracle on a few shapes.
Usage: python tools/tc_bringup.py [case ...]   (run under `timeout`; a protocol bug traps instead of hanging)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voicebridge_b200 import capi, host, synth  # noqa: E402

CASES = {
    "tiny": (11, 60, 39, 100),
    "mid": (130, 2000, 39, 1000),
    "bigpdf": (3, 700, 39, 50),
    "d40": (30, 200, 40, 257),
    "d13": (7, 7, 13, 64),
    "wide": (400, 4000, 39, 5000),
    "full": (4000, 40000, 39, 20000),
}


def run(name):
    P, N, D, T = CASES[name]
    m = synth.make_model(P, N, D, 5)
    X = synth.make_feats(m, T, 15)
    am = host.AmDiagGmmGpu.from_model(m)
    am.set_kernel(1)
    want = am.score(X)
    try:
        am.set_kernel(2)
    except capi.VbgpuError as e:
        print(name, "tc unavailable:", e)
        return
    t0 = time.time()
    got = am.score(X)
    dt = time.time() - t0
    err = np.abs(got - want)
    bad = np.argwhere(~(err <= 1e-3))
    print("%-7s P=%d N=%d D=%d T=%d  max|tc-simt| = %.3e  nonfinite=%d  mismatches=%d  (%.3fs)" % (
        name, P, m.num_gauss, D, T, np.nanmax(err), int((~np.isfinite(got)).sum()), len(bad), dt), flush=True)
    if len(bad):
        print("   first mismatches (t, p, got, want):", [(int(t), int(p), float(got[t, p]), float(want[t, p])) for t, p in bad[:6]])
        rows = np.unique(bad[:, 0])
        cols = np.unique(bad[:, 1])
        print("   rows affected: %d (first %s)  pdfs affected: %d (first %s)" % (len(rows), rows[:8], len(cols), cols[:8]))


if __name__ == "__main__":
    for c in (sys.argv[1:] or ["tiny", "d13", "d40", "bigpdf", "mid", "wide"]):
        if c in CASES:
            run(c)


def slab_check():
    """Scores the same frames in one call and in two calls (different work-unit splits): bits must agree."""
    m = synth.make_model(3000, 9000, 39, 4)
    X = synth.make_feats(m, 29843, 7)
    am = host.AmDiagGmmGpu.from_model(m)
    am.set_kernel(2)
    whole = am.score(X)
    again = am.score(X)
    cut = 22272
    parts = np.concatenate([am.score(X[:cut]), am.score(X[cut:])])
    am.set_kernel(1)
    simt = am.score(X)
    print("slab_check: run-to-run equal:", np.array_equal(whole, again), " whole vs 2 calls equal:",
          np.array_equal(whole, parts), " max|tc-simt| %.3e" % np.abs(whole - simt).max(), flush=True)
    d = np.argwhere(whole != parts)
    if len(d):
        print("   differing entries: %d, rows %s..., pdfs %s..., max diff %.3e" % (
            len(d), np.unique(d[:, 0])[:10], np.unique(d[:, 1])[:10], np.abs(whole - parts).max()))
        t, pp = d[0]
        print("   e.g. (t=%d,p=%d): whole %.6f parts %.6f simt %.6f" % (t, pp, whole[t, pp], parts[t, pp], simt[t, pp]))


if __name__ == "__main__" and "slab" in sys.argv[1:]:
    slab_check()
