"""Short driver for ncu: fMLLR statistics of 32 speakers over resident features, bench model (cfg 3).
Usage: python tools/prof_fmllr.py [frames] [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voicebridge_b200 import capi, host, synth  # noqa: E402


def model_offsets(am):
    return am._pdf_offsets


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1262745
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mfcc = host.Mfcc(capi.default_mfcc_opts(dither=0.0, use_energy=0))
    fp = host.FeaturePipeline(capi.default_feat_opts(), 13)
    w = synth.make_wave(int(8 * bench.SAMP), bench.SEED + 1000, bench.SAMP)
    mf, mfo = mfcc.compute_batch(w, [0, len(w)])
    fs = fp.run(mf, mfo, cmvn_stats=fp.cmvn_stats(mf, mfo))
    bm = bench.make_bench_model(fs)
    am = host.AmDiagGmmGpu.from_model(bm)
    am._pdf_offsets = np.asarray(bm.pdf_offsets)
    n_spk, n_utts = 32, 1024
    fo = np.linspace(0, T, n_utts + 1).astype(np.int64)
    u2s = np.repeat(np.arange(n_spk, dtype=np.int32), n_utts // n_spk)
    X = np.tile(fs, ((T + len(fs) - 1) // len(fs), 1))[:T]
    d_feats = torch.zeros((T, 40), dtype=torch.float32, device="cuda")
    d_feats[:, :39] = torch.from_numpy(X).cuda()
    d_ali = torch.from_numpy(synth.make_alignment(am.NumPdfs(), T, 7)).cuda()
    acc = host.FmllrDiagGmmAccsGpu(am, n_spk=n_spk)
    s = torch.cuda.current_stream()
    acc.accumulate_dev(d_feats, T, 40, d_ali, fo, u2s, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        acc.accumulate_dev(d_feats, T, 40, d_ali, fo, u2s, stream=s)
    e1.record(s)
    torch.cuda.synchronize()
    print("fmllr stats: T=%d  %.3f ms/call  (%d bad)" % (T, e0.elapsed_time(e1) / iters, am.bad_count()), flush=True)
    if os.environ.get("VBGPU_PROF_MLLT"):  # gmm-acc-mllt through the host entry point (H2D of the features included)
        import time
        Tm = min(T, 262144)
        Xh = np.ascontiguousarray(d_feats[:Tm].cpu().numpy())
        ah = np.ascontiguousarray(d_ali[:Tm].cpu().numpy())
        mllt = host.MlltAccsGpu(am)
        mllt.AccumulateForUtterance(Xh, ah)
        t0 = time.perf_counter()
        mllt.AccumulateForUtterance(Xh, ah)
        dt = time.perf_counter() - t0
        rows = int(np.diff(model_offsets(am))[ah].sum())
        print("mllt stats: T=%d (%d (frame, Gaussian) rows)  %.1f ms/call host-to-host = %.2f M frames/s" % (
            Tm, rows, dt * 1e3, Tm / dt / 1e6), flush=True)


if __name__ == "__main__":
    main()
