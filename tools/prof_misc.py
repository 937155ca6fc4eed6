"""Short driver for ncu: one EM pass PCM + alignment -> statistics on the bench batch (mfcc_kernel, cmvn_stats_kernel,
cmvn_norm_kernel, feat_kernel, acc_hist / acc_scan / acc_scatter / acc_bucket_kernel, acc_transitions_kernel) and one
sparse-consumer call (subset_kernel).   Usage: python tools/prof_misc.py [speakers]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voicebridge_b200 import capi, host, synth  # noqa: E402

if os.environ.get("VBGPU_LIB"):  # an experimental build of the library
    capi.LIB_PATH = os.path.abspath(os.environ["VBGPU_LIB"])


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    opts = capi.default_mfcc_opts(dither=0.0, use_energy=0)
    mfcc = host.Mfcc(opts)
    fp = host.FeaturePipeline(capi.default_feat_opts(), 13)
    pcm, so, u2s, n_spk, audio_s = bench.rank_corpus(1, 0)
    fo = mfcc.frame_offsets(so)
    T = int(fo[-1])
    w = synth.make_wave(int(8 * bench.SAMP), bench.SEED + 1000, bench.SAMP)
    mf, mfo = mfcc.compute_batch(w, [0, len(w)])
    fs = fp.run(mf, mfo, cmvn_stats=fp.cmvn_stats(mf, mfo))
    model = bench.make_bench_model(fs)
    am = host.AmDiagGmmGpu.from_model(model)
    pipe = host.ScoringPipeline(mfcc, fp, am)
    d_pcm = torch.from_numpy(pcm).to(dev)
    d_fm = torch.from_numpy(synth.make_fmllr(n_spk, bench.DIM, bench.SEED + 6)).to(dev)
    pdf, tid = bench.synth_alignment(bench.P_PDFS, fo, bench.SEED + 21)
    d_pdf, d_tid = torch.from_numpy(pdf).to(dev), torch.from_numpy(tid).to(dev)
    acc = host.AccumAmDiagGmmGpu(am, num_tids=2 * bench.P_PDFS)
    for _ in range(2):
        pipe.accumulate_dev(acc, d_pcm, so, u2s, n_spk, d_fm, bench.DIM + 1, d_pdf, stream=stream)
        acc.accumulate_transitions_dev(d_tid, T, stream=stream)
    torch.cuda.synchronize()
    n_utts = len(so) - 1
    sub_o = np.arange(n_utts + 1, dtype=np.int64) * 200
    sub_p = np.concatenate([np.random.default_rng(u).choice(bench.P_PDFS, 200, replace=False) for u in range(n_utts)]).astype(np.int32)
    pipe.score_subset(pcm, so, sub_o, sub_p, u2s, n_spk)
    print("ok: %d frames, statistics checksum %.17g" % (T, float(acc.as_tensor().double().abs().sum().item())))


if __name__ == "__main__":
    main()
