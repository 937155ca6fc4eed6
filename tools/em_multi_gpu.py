"""EM statistics accumulation over N GPUs with ONE all-reduce per pass (BASELINE configs[4] / SURVEY cfg 5 at reduced size).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/em_multi_gpu.py

Every rank accumulates its own shard of frames (feats -> stats) with acc_kernel, then the FP64 buffer
[occ | mean | var | tot_like | tot_frames] is summed in place twice, for comparison:
  (a) vbgpu_acc_allreduce on a raw ncclComm_t created here with ctypes (the C-ABI path a C++ host uses), and
  (b) torch.distributed.all_reduce on the zero-copy tensor view (host.AccumAmDiagGmmGpu.AllReduce).
Rank 0 re-accumulates ALL shards on its own GPU and checks the merged statistics (1e-10 relative: FP64 sums).
Prints one JSON line (rank 0)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voicebridge_b200 import capi, host, synth  # noqa: E402

P, N, D = 2000, 10000, 39
FRAMES_PER_RANK = 2_000_000


def nccl_comm(rank, world, device):
    """A raw ncclComm_t: the unique id is made on rank 0 and shared through torch.distributed."""
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
    lib = C.CDLL(cands[0] if cands else "libnccl.so.2", mode=C.RTLD_GLOBAL)

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]
    uid = UniqueId()
    if rank == 0:
        assert lib.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone().cuda(device)
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    comm = C.c_void_p()
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert lib.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    return lib, comm


def shard_data(model, rank):
    X = synth.make_feats(model, 20000, 100 + rank)
    reps = FRAMES_PER_RANK // len(X)
    ali = synth.make_alignment(P, len(X) * reps, 7 + rank)
    return np.tile(X, (reps, 1)), ali


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    model = synth.make_model(P, N, D, 5)
    am = host.AmDiagGmmGpu.from_model(model, device=local)
    X, ali = shard_data(model, rank)
    T = len(X)
    d_x = torch.zeros((T, 40), dtype=torch.float32, device="cuda")
    d_x[:, :D] = torch.from_numpy(X).cuda()
    d_ali = torch.from_numpy(ali).cuda()
    acc = host.AccumAmDiagGmmGpu(am)
    s = torch.cuda.current_stream()
    lib, comm = nccl_comm(rank, world, local)

    def one_pass(reduce):
        acc.SetZero()
        acc.accumulate_dev(d_x, T, 40, d_ali, stream=s)
        if reduce == "capi":
            capi.check(capi.lib().vbgpu_acc_allreduce(acc.h, comm, s.cuda_stream))
        elif reduce == "torch":
            acc.AllReduce()

    for _ in range(2):
        one_pass("capi")
    torch.cuda.synchronize()
    dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(s)
    acc.SetZero()
    acc.accumulate_dev(d_x, T, 40, d_ali, stream=s)
    e[1].record(s)
    capi.check(capi.lib().vbgpu_acc_allreduce(acc.h, comm, s.cuda_stream))
    e[2].record(s)
    torch.cuda.synchronize()
    t_acc, t_red = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    tm = torch.tensor([t_acc, t_red], dtype=torch.float64, device="cuda")
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    got_capi = acc.as_tensor().clone()
    one_pass("torch")
    torch.cuda.synchronize()
    got_torch = acc.as_tensor().clone()

    # rank 0: all shards on one GPU
    ok, err = True, 0.0
    if rank == 0:
        ref = host.AccumAmDiagGmmGpu(am)
        for r in range(world):
            Xr, ar = shard_data(model, r)
            dx = torch.zeros((len(Xr), 40), dtype=torch.float32, device="cuda")
            dx[:, :D] = torch.from_numpy(Xr).cuda()
            ref.accumulate_dev(dx, len(Xr), 40, torch.from_numpy(ar).cuda(), stream=s)
        torch.cuda.synchronize()
        want = ref.as_tensor()
        scale = want.abs().max().item()
        err = max((got_capi - want).abs().max().item(), (got_torch - want).abs().max().item()) / scale
        ok = err < 1e-10 and got_capi[-1].item() == world * T
        print(json.dumps({"n_gpus": world, "pdfs": P, "gaussians": N, "dim": D, "frames_per_gpu": T,
                          "accumulate_ms": tm[0].item(), "allreduce_ms": tm[1].item(),
                          "allreduce_bytes": int(got_capi.numel() * 8),
                          "audio_s_per_s_feats_to_stats": world * T / 100.0 / ((tm[0].item() + tm[1].item()) * 1e-3),
                          "hbm_gbs_per_gpu": T * 164 / (tm[0].item() * 1e-3) / 1e9,
                          "merged_vs_single_gpu_rel_err": err, "ok": bool(ok)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
