"""Device-resident throughput of the other BASELINE configs on one GPU (the driver's bench line is cfg 3, bench.py):
  cfg 2  tri-delta          P=2000  N=10000 D=39   PCM -> loglikes
  cfg 3  delta+SAT (fMLLR)  P=4000  N=40000 D=39   PCM -> loglikes   (same as bench.py, for cross-checking)
  cfg 4  LDA+MLLT           P=2500  N=15000 D=40   PCM -> splice+-3 -> 40x91 -> loglikes
  cfg 5  EM accumulation    N=10000 / 40000        feats -> stats and PCM -> stats, alignments = random pdf per ~7 frames
  n1     fMLLR statistics   cfg 2 / cfg 3 models   feats + alignment -> per-speaker beta, K, G (32 speakers)
  n4     other front ends   MFCC / fbank / PLP device-resident, Kaldi pitch host to host
Usage: python tools/bench_configs.py [steps]      -> one JSON line per config."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voicebridge_b200 import capi, host, synth  # noqa: E402


def timed(fn, steps, stream):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    opts = capi.default_mfcc_opts(dither=0.0, use_energy=0)
    mfcc = host.Mfcc(opts)
    pcm, so, u2s, n_spk, audio_s = bench.rank_corpus(1, 0)
    fo = mfcc.frame_offsets(so)
    T = int(fo[-1])
    d_pcm = torch.from_numpy(pcm).to(dev)
    w = synth.make_wave(int(8 * bench.SAMP), bench.SEED + 1000, bench.SAMP)
    mf, mfo = mfcc.compute_batch(w, [0, len(w)])

    cfgs = {
        "cfg2_tri_delta": dict(P=2000, N=10000, lda=False, fmllr=False),
        "cfg3_delta_sat": dict(P=4000, N=40000, lda=False, fmllr=True),
        "cfg4_lda_mllt": dict(P=2500, N=15000, lda=True, fmllr=False),
    }
    models = {}
    for name, c in cfgs.items():
        if c["lda"]:
            fo_ = capi.default_feat_opts()
            fo_.mode = 1
            lda = (np.random.default_rng(7).standard_normal((40, 91)) / np.sqrt(91.0)).astype(np.float32)
            fp = host.FeaturePipeline(fo_, 13, transform=lda)
            D = 40
        else:
            fp = host.FeaturePipeline(capi.default_feat_opts(), 13)
            D = 39
        fs = fp.run(mf, mfo, cmvn_stats=fp.cmvn_stats(mf, mfo))
        model = synth.make_model_from_feats(fs, c["P"], c["N"], bench.SEED)
        am = host.AmDiagGmmGpu.from_model(model)
        pipe = host.ScoringPipeline(mfcc, fp, am)
        d_fm = torch.from_numpy(synth.make_fmllr(n_spk, D, bench.SEED + 6)).to(dev) if c["fmllr"] else None
        d_ll = torch.empty((T, c["P"]), dtype=torch.float32, device=dev)
        d_feats = torch.empty((T, 40), dtype=torch.float32, device=dev)

        def step(pipe=pipe, d_fm=d_fm, D=D, d_ll=d_ll, P=c["P"], d_feats=d_feats):
            pipe.score_dev(d_pcm, so, u2s, n_spk, d_fm, D + 1 if d_fm is not None else 0, d_ll, P, d_feats, 40, stream)
        ms = timed(step, steps, stream)
        ms_score = timed(lambda am=am, d_feats=d_feats, d_ll=d_ll, P=c["P"]: am.score_dev(d_feats, T, 40, d_ll, P, stream),
                         steps, stream)
        flops = 2.0 * (2 * D + 1) * c["N"] * T
        print(json.dumps({"config": name, "pdfs": c["P"], "gaussians": c["N"], "dim": D, "frames": T,
                          "ms_per_step": ms, "audio_s_per_s": audio_s / (ms * 1e-3), "scoring_ms": ms_score,
                          "scoring_tflops_algorithmic": flops / (ms_score * 1e-3) / 1e12,
                          "nonfinite": am.bad_count()}), flush=True)
        models[name] = (model, am, pipe, d_feats, d_fm, D)
        del d_ll
        torch.cuda.empty_cache()

    # cfg 5: EM accumulation (one pass over the same corpus; statistics for the aligned pdf of every frame)
    for name in ("cfg2_tri_delta", "cfg3_delta_sat"):
        model, am, pipe, d_feats, d_fm, D = models[name]
        P = am.NumPdfs()
        ali = synth.make_alignment(P, T, 7)
        d_ali = torch.from_numpy(ali).to(dev)
        acc = host.AccumAmDiagGmmGpu(am)
        ms_f = timed(lambda: acc.accumulate_dev(d_feats, T, 40, d_ali, stream=stream), steps, stream)
        ms_p = timed(lambda: pipe.accumulate_dev(acc, d_pcm, so, u2s, n_spk, d_fm, D + 1 if d_fm is not None else 0, d_ali,
                                                 stream=stream), steps, stream)
        print(json.dumps({"config": "cfg5_em_" + name, "gaussians": am.NumGauss(), "frames": T,
                          "feats_to_stats_ms": ms_f, "feats_to_stats_audio_s_per_s": audio_s / (ms_f * 1e-3),
                          "feats_to_stats_gbs": T * 164 / (ms_f * 1e-3) / 1e9,
                          "pcm_to_stats_ms": ms_p, "pcm_to_stats_audio_s_per_s": audio_s / (ms_p * 1e-3),
                          "nonfinite": am.bad_count()}), flush=True)

    # fMLLR statistics of all 32 speakers of the batch (gmm-est-fmllr's accumulation)
    for name in ("cfg2_tri_delta", "cfg3_delta_sat"):
        model, am, pipe, d_feats, d_fm, D = models[name]
        ali = synth.make_alignment(am.NumPdfs(), T, 7)
        d_ali = torch.from_numpy(ali).to(dev)
        fm = host.FmllrDiagGmmAccsGpu(am, n_spk=n_spk)
        ms_f = timed(lambda: fm.accumulate_dev(d_feats, T, 40, d_ali, fo, u2s, stream=stream), steps, stream)
        flops = 2.0 * T * D * (D + 1) * (D + 2) / 2  # G: one multiply-add per (frame, i, j >= k)
        print(json.dumps({"config": "fmllr_stats_" + name, "gaussians": am.NumGauss(), "frames": T, "speakers": n_spk,
                          "feats_to_fmllr_stats_ms": ms_f, "audio_s_per_s": audio_s / (ms_f * 1e-3),
                          "g_tflops_fp32": flops / (ms_f * 1e-3) / 1e12, "nonfinite": am.bad_count()}), flush=True)


    # the other front ends on the same batch, device-resident PCM -> features
    fronts = (("mfcc_13", mfcc), ("fbank_23_log", host.Fbank(capi.default_mfcc_opts(dither=0.0, use_energy=0))),
              ("fbank_23_log_energy", host.Fbank(capi.default_mfcc_opts(dither=0.0, use_energy=1))),
              ("plp_13", host.Plp(opts)))
    for name, fe in fronts:
        dim = fe.Dim()
        st = (dim + 3) // 4 * 4
        d_out = torch.empty((T, st), dtype=torch.float32, device=dev)
        ms_f = timed(lambda fe=fe, d_out=d_out, st=st: fe.compute_dev(d_pcm, so, d_out, st, False, stream), steps, stream)
        print(json.dumps({"config": "frontend_" + name, "frames": T, "dim": dim, "ms": ms_f, "audio_s_per_s": audio_s / (ms_f * 1e-3),
                          "hbm_gbs_algorithmic": T * (320 + 4 * st) / (ms_f * 1e-3) / 1e9}), flush=True)
    pit = host.Pitch()
    for name, proc in (("pitch_raw", None), ("pitch_processed", capi.default_process_pitch_opts())):
        t = []
        import time
        pin = torch.from_numpy(pcm).pin_memory().numpy()
        pit.compute_batch(pin, so, proc)
        for _ in range(3):
            t0 = time.perf_counter()
            pit.compute_batch(pin, so, proc)
            t.append(time.perf_counter() - t0)
        ms_f = 1e3 * float(np.median(t))
        print(json.dumps({"config": "frontend_" + name, "ms_host_to_host": ms_f, "audio_s_per_s": audio_s / (ms_f * 1e-3),
                          "lag_states": pit.NumStates()}), flush=True)


if __name__ == "__main__":
    main()
