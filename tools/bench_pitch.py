"""Kaldi-pitch throughput on one B200: a packed batch of utterances through
vbgpu_pitch_compute_i16 (host buffers in, (NCCF, pitch) or processed rows out), timed by wall clock around the call
(the call is synchronous and includes the copies), next to the reference's ComputeKaldiPitch on one host core.

    python tools/bench_pitch.py [n_utts] [reps]        -> one JSON line
Per-kernel times: run the same command under `ncu --metrics gpu__time_duration.sum` (profiles/r1_pitch_launches.csv).
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from voicebridge_b200 import capi, host, synth  # noqa: E402


def main():
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    pcm, so, _ = synth.make_corpus(n_utts // 8, 8, 5.0, 20.0, 3, fast=True)
    pageable = pcm
    try:  # pinned host PCM, as bench.py's e2e leg does
        import torch
        pcm = torch.from_numpy(pcm).pin_memory().numpy()
        host_mem = "pinned"
    except Exception:  # noqa: BLE001
        host_mem = "pageable"
    audio_s = float(so[-1]) / 16000.0
    p = host.Pitch()
    pp = capi.default_process_pitch_opts()
    out = {"metric": "audio_sec_per_sec_kaldi_pitch", "unit": "audio-s/s", "n_utts": n_utts, "audio_s": audio_s,
           "lag_states": p.NumStates(), "host_pcm": host_mem}
    for name, proc in (("raw", None), ("processed", pp)):
        rows, ro = p.compute_batch(pcm, so, proc)  # warm-up: allocations
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            p.compute_batch(pcm, so, proc)
            t.append(time.perf_counter() - t0)
        out[name] = {"ms_per_call": 1e3 * float(np.median(t)), "value": audio_s / float(np.median(t)), "rows": int(ro[-1])}
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        p.compute_batch(pageable, so, None)
        t.append(time.perf_counter() - t0)
    out["raw_pageable_pcm"] = {"ms_per_call": 1e3 * float(np.median(t)), "value": audio_s / float(np.median(t))}
    try:  # the reference's own ComputeKaldiPitch on one host core, bounded sample
        from oracle import pyoracle as po
        ref = po.load("ref")
        o = po.default_pitch_opts()
        n = 0
        t0 = time.perf_counter()
        for u in range(min(n_utts, 6)):
            ref.pitch(o, pcm[so[u]:so[u + 1]].astype(np.float32))
            n += int(so[u + 1] - so[u])
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / 16000.0 / dt, "unit": "audio-s/s", "cores": 1, "kind": "reference",
                               "sample": "%d utterances, %.0f audio-s, ComputeKaldiPitch" % (min(n_utts, 6), n / 16000.0)}
    except Exception as e:  # noqa: BLE001
        out["cpu_baseline"] = {"unavailable": str(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
