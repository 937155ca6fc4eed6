"""Generate the committed golden fixtures under tests/golden/ (run in the build container only).

  htk_golden.npz   the reference's own known-answer vectors for the MFCC path: feat/test_data/test.wav (16 kHz) and the
                   six HTK feature files test.wav.fea_htk.{1..6} that feat/feature-mfcc-test.cc:112-650 compares against
                   (tolerance 1.0 absolute, 10 edge frames skipped), plus the option set of each compare.
  ref_golden.npz   outputs of the reference's OWN code (oracle/_ref/libvbref.so, compiled from /root/reference) on
                   seeded inputs for every step of the path: MFCC, CMVN, deltas, splice+LDA, fMLLR, dense log-likelihoods,
                   EM statistics.  These pin the steps the reference's unit tests do not pin.

Usage:  python tests/golden/make_golden.py      (needs /root/reference and `make -C oracle ref`)
"""
import os
import struct
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from voicebridge_b200 import synth  # noqa: E402

TEST_DATA = "/root/reference/kaldi-master/src/feat/test_data"

# Option sets of UnitTestHTKCompare1..6 (feature-mfcc-test.cc:133-143,217-226,301-311,386-394,469-480,557-567)
HTK_CASES = {
    1: dict(preemph_coeff=0.0, window_type=1, remove_dc_offset=0, low_freq=0.0, htk_mode=1, htk_compat=1, use_energy=0),
    2: dict(preemph_coeff=0.0, window_type=1, remove_dc_offset=0, low_freq=0.0, htk_mode=1, htk_compat=1, use_energy=1),
    3: dict(preemph_coeff=0.0, window_type=1, remove_dc_offset=0, low_freq=20.0, htk_mode=1, htk_compat=1, use_energy=1),
    4: dict(window_type=1, remove_dc_offset=0, low_freq=0.0, htk_mode=1, htk_compat=1, use_energy=1),
    5: dict(window_type=1, remove_dc_offset=0, low_freq=0.0, vtln_low=100.0, vtln_high=7500.0, htk_mode=1, htk_compat=1,
            use_energy=1),
    6: dict(preemph_coeff=0.97, window_type=1, remove_dc_offset=0, num_bins=24, low_freq=125.0, high_freq=7800.0,
            htk_compat=1, use_energy=0),
}
HTK_VTLN = {5: 1.1}


def read_htk(path):
    with open(path, "rb") as f:
        n, period, size, kind = struct.unpack(">iihh", f.read(12))
        data = np.frombuffer(f.read(n * size), dtype=">f4").reshape(n, size // 4).astype(np.float32)
    return data


def main():
    w = wave.open(os.path.join(TEST_DATA, "test.wav"), "rb")
    assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
    pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16)
    out = {"pcm": pcm}
    for k in HTK_CASES:
        out["htk%d" % k] = read_htk(os.path.join(TEST_DATA, "test.wav.fea_htk.%d" % k))
    np.savez_compressed(os.path.join(HERE, "htk_golden.npz"), **out)
    print("htk_golden.npz: pcm", pcm.shape, {k: out["htk%d" % k].shape for k in HTK_CASES})

    ref = po.load("ref")
    g = {}
    # (1) standard recipe front end on test.wav and on a synthetic 8 kHz wave
    o16 = po.default_opts(dither=0.0, use_energy=0)
    g["mfcc16"] = ref.mfcc(o16, pcm.astype(np.float32))
    o16e = po.default_opts(dither=0.0)  # Kaldi defaults: use_energy + raw_energy
    g["mfcc16_energy"] = ref.mfcc(o16e, pcm.astype(np.float32))
    o16ns = po.default_opts(dither=0.0, use_energy=0, snip_edges=0)
    g["mfcc16_nosnip"] = ref.mfcc(o16ns, pcm.astype(np.float32))
    g["mfcc16_vtln09"] = ref.mfcc(o16, pcm.astype(np.float32), 0.9)
    w8 = synth.make_wave(8000 * 2, 11, 8000.0)
    g["wave8"] = w8
    o8 = po.default_opts(dither=0.0, use_energy=0, samp_freq=8000.0)
    g["mfcc8"] = ref.mfcc(o8, w8.astype(np.float32))
    # (2) CMVN, deltas, splice + LDA, fMLLR
    x = g["mfcc16"]
    st = ref.cmvn_acc(x)
    g["cmvn_stats"] = st
    g["cmvn_mean"] = ref.cmvn_apply(st, x, False)
    g["cmvn_meanvar"] = ref.cmvn_apply(st, x, True)
    g["delta"] = ref.deltas(g["cmvn_mean"], 2, 2)
    g["delta_o1_w3"] = ref.deltas(g["cmvn_mean"], 1, 3)
    lda = synth.make_lda(40, 91, 3)
    g["lda_mat"] = lda
    g["lda"] = ref.transform(ref.splice(g["cmvn_mean"], 3, 3), lda)
    lda_aff = synth.make_lda(40, 92, 5)
    g["lda_aff_mat"] = lda_aff
    g["lda_aff"] = ref.transform(ref.splice(g["cmvn_mean"], 3, 3), lda_aff)
    fm = synth.make_fmllr(1, 39, 4)[0]
    g["fmllr_mat"] = fm
    g["fmllr"] = ref.transform(g["delta"], fm)
    # (3) model: scoring + statistics on features matched to the model
    m = synth.make_model_from_feats(g["delta"], 40, 300, 7)
    gc, miv, iv = ref.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    g["pdf_offsets"], g["weights"], g["means"] = m.pdf_offsets, m.weights, m.means
    g["gconsts"], g["miv"], g["iv"] = gc, miv, iv
    m.gconsts, m.miv, m.iv = gc, miv, iv
    rc, ll = ref.gmm_loglikes(m, g["delta"])
    assert rc == 0
    g["loglikes"] = ll
    ali = synth.make_alignment(40, len(g["delta"]), 3)
    wts = np.random.default_rng(5).uniform(0.2, 1.0, len(ali)).astype(np.float32)
    g["ali"], g["ali_w"] = ali, wts
    rc, occ, mean, var, tl, tf = ref.acc_ali(m, g["delta"], ali)
    assert rc == 0
    g["acc_occ"], g["acc_mean"], g["acc_var"], g["acc_tot"] = occ, mean, var, np.array([tl, tf])
    rc, occ, mean, var, tl, tf = ref.acc_ali(m, g["delta"], ali, wts, g["fmllr"])
    assert rc == 0
    g["acc2_occ"], g["acc2_mean"], g["acc2_var"], g["acc2_tot"] = occ, mean, var, np.array([tl, tf])
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **g)
    print("ref_golden.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
