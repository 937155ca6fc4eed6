"""torchrun diagnostic: is one rank slow for every kernel (device-level) or only for its own data?"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
import bench
from voicebridge_b200 import capi, host, synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream()

def timed(fn, n):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def gather(x):
    t = torch.zeros(world, dtype=torch.float64, device=dev); t[rank] = x
    dist.all_reduce(t); return [round(v, 2) for v in t.tolist()]

a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
mm = gather(timed(lambda: torch.matmul(a, b), 50))
mfcc = host.Mfcc(capi.default_mfcc_opts(dither=0.0, use_energy=0), device=local)
fp = host.FeaturePipeline(capi.default_feat_opts(), 13, device=local)
w = synth.make_wave(int(8 * bench.SAMP), bench.SEED + 1000, bench.SAMP)
mf, mfo = mfcc.compute_batch(w, [0, len(w)])
fs = fp.run(mf, mfo, cmvn_stats=fp.cmvn_stats(mf, mfo))
model = bench.make_bench_model(fs)
am = host.AmDiagGmmGpu.from_model(model, device=local)
n_cols = am.NumCols()
T = 1262745
X = np.tile(fs, ((T + fs.shape[0] - 1) // fs.shape[0], 1))[:T]
d_feats = torch.zeros((T, 40), dtype=torch.float32, device=dev); d_feats[:, :39] = torch.from_numpy(X).to(dev)
d_ll = torch.empty((T + 20000, n_cols), dtype=torch.float32, device=dev)
same = gather(timed(lambda: am.score_cols_dev(d_feats, T, 40, d_ll, n_cols, stream), 5))
pipe = host.ScoringPipeline(mfcc, fp, am)
pcm, so, u2s, n_spk, audio_s = bench.rank_corpus(world, rank)
fo = mfcc.frame_offsets(so); T2 = int(fo[-1])
d_pcm = torch.from_numpy(pcm).to(dev)
d_fm = torch.from_numpy(synth.make_fmllr(n_spk, bench.DIM, bench.SEED + 6)).to(dev)
d_f2 = torch.zeros((T2, 40), dtype=torch.float32, device=dev)
pipe.score_cols_dev(d_pcm, so, u2s, n_spk, d_fm, bench.DIM + 1, d_ll, n_cols, d_f2, 40, stream)
own = gather(timed(lambda: am.score_cols_dev(d_f2, T2, 40, d_ll, n_cols, stream), 5))
mm2 = gather(timed(lambda: torch.matmul(a, b), 50))
col = torch.from_numpy(am.col_of_pdf().astype(np.int64)).to(dev)
low = gather(float((d_ll[:T2][:, col].min(dim=1).values < -27000).sum()))
if rank == 0:
    print("matmul 8192^3 bf16 ms by rank   :", mm)
    print("score, same features, ms by rank:", same)
    print("score, own corpus, ms by rank   :", own)
    print("matmul again                    :", mm2)
    print("frames below the sunk level     :", low)
dist.destroy_process_group()
