"""Features of the bench batch (PCM -> MFCC -> CMVN -> deltas -> fMLLR) from the library named by VBGPU_LIB (or the in-tree
one): saves a checksum file so that two builds can be compared bit for bit.  Usage: python tools/cmp_feats.py out.npy"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voicebridge_b200 import capi, host, synth  # noqa: E402

if os.environ.get("VBGPU_LIB"):
    capi.LIB_PATH = os.path.abspath(os.environ["VBGPU_LIB"])


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    mfcc = host.Mfcc(capi.default_mfcc_opts(dither=0.0, use_energy=0))
    fp = host.FeaturePipeline(capi.default_feat_opts(), 13)
    w = synth.make_wave(int(8 * bench.SAMP), bench.SEED + 1000, bench.SAMP)
    mf, mfo = mfcc.compute_batch(w, [0, len(w)])
    fs = fp.run(mf, mfo, cmvn_stats=fp.cmvn_stats(mf, mfo))
    am = host.AmDiagGmmGpu.from_model(synth.make_model(50, 200, bench.DIM, 3))
    pipe = host.ScoringPipeline(mfcc, fp, am)
    pcm, so, u2s, n_spk, audio_s = bench.rank_corpus(1, 0)
    fo = mfcc.frame_offsets(so)
    T = int(fo[-1])
    d_pcm = torch.from_numpy(pcm).to(dev)
    d_fm = torch.from_numpy(synth.make_fmllr(n_spk, bench.DIM, bench.SEED + 6)).to(dev)
    d_feats = torch.zeros((T, 40), dtype=torch.float32, device=dev)
    n_cols = am.NumCols()
    d_ll = torch.empty((T, n_cols), dtype=torch.float32, device=dev)
    pipe.score_cols_dev(d_pcm, so, u2s, n_spk, d_fm, bench.DIM + 1, d_ll, n_cols, d_feats, 40, stream)
    torch.cuda.synchronize()
    np.save(sys.argv[1], d_feats.cpu().numpy())
    print("saved", sys.argv[1], float(d_feats.double().abs().sum().item()))


if __name__ == "__main__":
    main()
