"""Times the MFCC launch (mfcc_kernel) and the whole front end (MFCC + CMVN + splice/LDA/fMLLR) on the bench batch with
CUDA events, and prints a checksum of the features so two library builds can be compared (VBGPU_LIB=...).
Usage: python tools/time_front.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voicebridge_b200 import capi, host  # noqa: E402

if os.environ.get("VBGPU_LIB"):  # an experimental build of the library
    capi.LIB_PATH = os.path.abspath(os.environ["VBGPU_LIB"])


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    pcm, so, u2s, n_spk, audio_s = bench.rank_corpus(1, 0)
    for name, kw in (("mfcc 8k/int16", dict(dither=0.0, use_energy=0)), ("mfcc snip_edges=false", dict(dither=0.0, use_energy=1, snip_edges=0))):
        mfcc = host.Mfcc(capi.default_mfcc_opts(**kw))
        fo = mfcc.frame_offsets(so)
        T = int(fo[-1])
        d_pcm = torch.from_numpy(pcm).to(dev)
        d_out = torch.empty((T, 16), dtype=torch.float32, device=dev)
        for _ in range(3):
            mfcc.compute_dev(d_pcm, so, d_out, 16, stream=stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            mfcc.compute_dev(d_pcm, so, d_out, 16, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        h = d_out[:, :13].double()
        print("%-24s %8.3f ms / batch of %d frames   sum %.10e  abs-sum %.10e" % (name, e0.elapsed_time(e1) / reps, T, h.sum().item(), h.abs().sum().item()))
        np.save("gpurun_out/front_%s_%s.npy" % (os.environ.get("VBGPU_LIB", "cur").split("/")[-1], "snip" if "snip" in name else "std"), d_out[:200000].cpu().numpy())


if __name__ == "__main__":
    main()
