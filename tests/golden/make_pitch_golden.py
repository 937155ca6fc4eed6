"""Writes tests/golden/pitch_golden.npz: Kaldi-pitch outputs of the reference's OWN code (oracle/_ref/libvbref.so,
compiled from /root/reference by `make -C oracle ref`) — ComputeKaldiPitch and ProcessPitch on the 16 kHz speech of
feat/test_data/test.wav (its samples are already in htk_golden.npz) and on a synthetic 8 kHz wave.

Run in the build container only:  python -m tests.golden.make_pitch_golden
"""
import os

import numpy as np

from oracle import pyoracle as po
from voicebridge_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))

PROCESS_VARIANT = dict(delay=2, add_raw_log_pitch=1, normalization_left_context=40, normalization_right_context=20,
                       delta_window=3, pov_offset=0.5)


def main():
    ref = po.load("ref")
    pcm = np.load(os.path.join(HERE, "htk_golden.npz"))["pcm"].astype(np.float32)
    g = {}
    g["raw16"] = ref.pitch(po.default_pitch_opts(), pcm)
    g["raw16_nosnip"] = ref.pitch(po.default_pitch_opts(snip_edges=0), pcm)
    g["raw16_short"] = ref.pitch(po.default_pitch_opts(), pcm[:9000])  # < recompute_frame: the energy correction path
    g["proc16"] = ref.process_pitch(po.default_process_pitch_opts(), g["raw16"])
    g["proc16_variant"] = ref.process_pitch(po.default_process_pitch_opts(**PROCESS_VARIANT), g["raw16"])
    w8 = synth.make_pitch_wave(8000 * 3, 21, 8000.0)
    g["wave8"] = w8
    g["raw8"] = ref.pitch(po.default_pitch_opts(samp_freq=8000.0, min_f0=60.0, max_f0=350.0), w8.astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "pitch_golden.npz"), **g)
    print("pitch_golden.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
