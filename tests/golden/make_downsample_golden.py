"""Writes tests/golden/downsample_golden.npz: DownsampleWaveForm (feat/resample.cc:368-376) of the reference's OWN code
(oracle/_ref/libvbref.so, compiled from /root/reference by `make -C oracle ref`) on the 16 kHz speech of
feat/test_data/test.wav (its samples are in htk_golden.npz) taken down to 8 kHz and 11.025 kHz, and on seeded noise for the
rate pairs the recipes meet (44.1 k / 48 k / 22.05 k -> 16 k), plus the MFCCs of the 8 kHz rendering through the reference's
ComputeFeatures(wave, 16000, ...) with allow_downsample (feat/feature-common-inl.h:29-55).

Run in the build container only:  python -m tests.golden.make_downsample_golden
"""
import os

import numpy as np

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
PAIRS = [(44100, 16000, 30001), (48000, 16000, 5000), (22050, 16000, 12345), (16000, 8000, 7)]


def noise(n, seed):
    return (np.random.default_rng(seed).standard_normal(n) * 3000).astype(np.float32)


def main():
    ref = po.load("ref")
    pcm = np.load(os.path.join(HERE, "htk_golden.npz"))["pcm"].astype(np.float32)
    g = {"speech_16k_to_8k": ref.downsample_waveform(16000, 8000, pcm),
         "speech_16k_to_11025": ref.downsample_waveform(16000, 11025, pcm)}
    for orig, new, n in PAIRS:
        g["noise_%d_to_%d_%d" % (orig, new, n)] = ref.downsample_waveform(orig, new, noise(n, n))
    o = po.default_opts(dither=0.0, samp_freq=8000.0)
    g["mfcc_of_speech_8k"] = ref.mfcc(o, g["speech_16k_to_8k"])
    np.savez_compressed(os.path.join(HERE, "downsample_golden.npz"), **g)
    print("downsample_golden.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
