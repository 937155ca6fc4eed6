"""Yes-no scenario (BASELINE configs[0]): train a monophone diag-GMM system and decode held-out utterances with the
reference's OWN decoders, once on the reference's CPU decodable and once on the GPU decodable of include/vbgpu_kaldi.h.
1-best word sequences and alignments must be IDENTICAL; features / log-likelihoods / EM statistics within the
BASELINE tolerances.  The driver (oracle/ref_yesno.cc) is compiled in the build container against the reference's
sources and travels as a binary (oracle/_ref/ref_yesno); nothing here reads /root/reference at run time."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_yesno")

needs_bin = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/ref_yesno not built (needs /root/reference)")


def run(mode):
    r = subprocess.run([BIN, mode], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_bin
def test_yesno_scenario_on_the_reference():
    """The synthesised scenario is a real recognition task for the reference itself (its README quotes 2 % WER)."""
    out = run("cpu")
    assert out["pdfs"] == 9 and out["gaussians"] == 120
    assert out["wer_reference"] <= 0.05


@needs_bin
@pytest.mark.gpu
def test_yesno_identical_transcripts_and_alignments():
    out = run("gpu")
    n = out["test_utts"]
    assert out["transcripts_identical"] == n, out
    assert out["decode_alignments_identical"] == n, out
    assert out["forced_alignments_identical"] == n, out
    assert out["wer_gpu"] == out["wer_reference"]
    assert out["max_feat_rel_err"] <= 1e-4, out        # features: 1e-4 relative
    assert out["max_loglike_abs_err"] <= 1e-3, out     # log-likelihoods: 1e-3 absolute
    assert out["max_stats_rel_err"] <= 1e-4, out       # EM statistics: 1e-4 relative
    assert out["max_acc_loglike_rel_err"] <= 1e-5, out
    # SURVEY 8f n3 / n1 through the C++ adaptors: one scoring call for the whole set, per-utterance decodable views;
    # fMLLR statistics per speaker and the reference's own solver on them
    assert out["batch_forced_alignments_identical"] == n, out
    assert out["fmllr_stats_rel_err"] <= 1e-4, out
    assert out["fmllr_xform_rel_err"] <= 1e-3, out
    # gmm-rescore-lattice: the reference's RescoreLattice over the batch's decodable views
    assert out["rescored_lattice_arcs"] > 1000 and out["rescored_arc_abs_err"] <= 1e-3, out
    assert out["rescored_best_paths_identical"] == n, out
    # the sparse consumers: forced alignment on each utterance's own pdf subset (vbgpu_gmm_score_subset), lattice rescoring
    # on the arcs' (frame, pdf) pairs only (vbgpu_gmm_score_gather) — same alignments / best paths, a fraction of the floats
    assert out["subset_forced_alignments_identical"] == n, out
    assert 0 < out["subset_floats"] <= out["dense_floats"], out  # (the 9-pdf yes-no graphs touch every pdf; see test_gpu_sparse)
    assert out["gather_arcs"] > 1000 and out["gather_arc_abs_err"] <= 1e-3, out
    assert out["gather_best_paths_identical"] == n, out
    # Kaldi pitch through vbgpu::GpuPitch vs the reference's ComputeKaldiPitch / ProcessPitch (tests.common.assert_pitch_close)
    assert out["pitch_frames"] > 1000 and out["pitch_frames_identical"] >= 0.9 * out["pitch_frames"], out
    assert out["pitch_max_rel_err"] <= 0.02 and out["pitch_nccf_abs_err"] <= 1e-4, out
    assert out["process_pitch_abs_err"] <= 1e-4, out
    # --allow-downsample: a 16 kHz wave through ComputeFeatures of the reference and of the GPU adaptor; refused without the switch
    assert out["downsample_mfcc_rel_err"] <= 1e-4 and out["downsample_refused_without_switch"] == 1, out
