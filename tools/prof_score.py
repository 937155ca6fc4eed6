"""Short driver for ncu: the bench model (cfg 3: P=4000, N=40000, D=39) scored on a few waves of resident features.
Usage: python tools/prof_score.py [frames] [iters] [kernel] [layout]
       kernel: 0 auto, 1 simt, 2 tcgen05;  layout: cols (device column order, default) | pdf (model order: + gather kernel)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voicebridge_b200 import capi, host, synth  # noqa: E402

if os.environ.get("VBGPU_LIB"):  # bring-up: an experimental build of the library
    capi.LIB_PATH = os.path.abspath(os.environ["VBGPU_LIB"])


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 4
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    opts = capi.default_mfcc_opts(dither=0.0, use_energy=0)
    mfcc = host.Mfcc(opts)
    fp = host.FeaturePipeline(capi.default_feat_opts(), 13)
    w = synth.make_wave(int(8 * bench.SAMP), bench.SEED + 1000, bench.SAMP)
    mf, mfo = mfcc.compute_batch(w, [0, len(w)])
    fs = fp.run(mf, mfo, cmvn_stats=fp.cmvn_stats(mf, mfo))
    model = bench.make_bench_model(fs)
    am = host.AmDiagGmmGpu.from_model(model)
    if kernel:
        am.set_kernel(kernel)
    reps = (T + fs.shape[0] - 1) // fs.shape[0]
    X = np.tile(fs, (reps, 1))[:T]
    X = X + np.random.default_rng(3).normal(0, 0.3, X.shape).astype(np.float32)
    d_feats = torch.zeros((T, 40), dtype=torch.float32, device="cuda")
    d_feats[:, :39] = torch.from_numpy(X).cuda()
    layout = sys.argv[4] if len(sys.argv) > 4 else "cols"
    ncols = am.NumCols() if layout == "cols" else bench.P_PDFS
    d_ll = torch.empty((T, ncols), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    run = am.score_cols_dev if layout == "cols" else am.score_dev
    run(d_feats, T, 40, d_ll, ncols, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        run(d_feats, T, 40, d_ll, ncols, s)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * 79 * bench.N_GAUSS * T
    print("score[%s, %d cols, note=%r]: T=%d  %.3f ms/launch  %.1f TFLOP/s algorithmic  (%.2f us per 256-frame tile-row per SM)" % (
        layout, ncols, am.plan_note(), T, ms, fl / ms / 1e9, ms * 1e3 / (T / 256.0 / 148.0)), flush=True)
    if os.environ.get("VBGPU_TC_DEBUG"):
        print("debug counter (VBGPU_TC_DEBUG=%s): %d over %d launches" % (os.environ["VBGPU_TC_DEBUG"], am.bad_count(), iters + 1))


if __name__ == "__main__":
    main()
