"""world_size-2 test of the N>1 path on CPU (gloo): speakers are sharded over ranks with no data-path collective for
scoring, and the EM statistics are merged by ONE sum all-reduce of the flat [occ|mean|var|tot_like|tot_frames] buffer.
Per-rank compute is done by the oracle here (the GPU kernels are covered by the -m gpu tests); what is under test is the
host logic: the partition, the buffer layout, and that reduce(shards) == single-rank result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _corpus():
    from voicebridge_b200 import synth
    m = synth.make_model(12, 60, 13, 5)
    rng = np.random.default_rng(1)
    n_utts = 10
    u2s = np.array([0, 0, 1, 1, 1, 2, 3, 3, 4, 4], np.int32)
    lens = rng.integers(20, 60, n_utts)
    fo = np.zeros(n_utts + 1, np.int64)
    fo[1:] = np.cumsum(lens)
    X = synth.make_feats(m, int(fo[-1]), 9)
    ali = synth.make_alignment(12, int(fo[-1]), 4)
    return m, u2s, lens, fo, X, ali


def _flat(occ, mean, var, tl, tf):
    return np.concatenate([occ, mean.ravel(), var.ravel(), [tl, tf]])


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from voicebridge_b200 import shard
    orc = po.load("orc")
    m, u2s, lens, fo, X, ali = _corpus()
    mine = shard.shard_speakers(u2s, lens, world)[rank]
    rows = np.concatenate([np.arange(fo[u], fo[u + 1]) for u in mine]) if len(mine) else np.zeros(0, np.int64)
    rc, occ, mean, var, tl, tf = orc.acc_ali(m, X[rows], ali[rows])
    assert rc == 0
    buf = torch.from_numpy(_flat(occ, mean, var, tl, tf))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # the one collective of an EM iteration
    _, ll = orc.gmm_loglikes(m, X[rows])        # scoring: shard-local, no collective
    np.save(os.path.join(outdir, "acc_%d.npy" % rank), buf.numpy())
    np.save(os.path.join(outdir, "ll_%d.npy" % rank), ll)
    np.save(os.path.join(outdir, "rows_%d.npy" % rank), rows)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_accumulate_allreduce_equals_single_rank(tmp_path, orc):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    m, u2s, lens, fo, X, ali = _corpus()
    rc, occ, mean, var, tl, tf = orc.acc_ali(m, X, ali)
    want = _flat(occ, mean, var, tl, tf)
    got0, got1 = np.load(tmp_path / "acc_0.npy"), np.load(tmp_path / "acc_1.npy")
    assert np.array_equal(got0, got1)                       # every rank holds the merged statistics
    assert np.allclose(got0, want, rtol=1e-12, atol=1e-12)  # FP64 sums: order-of-addition noise only
    assert got0[-1] == fo[-1]                               # tot_frames
    _, ll = orc.gmm_loglikes(m, X)
    seen = np.zeros(len(X), bool)
    for r in range(world):
        rows = np.load(tmp_path / ("rows_%d.npy" % r))
        assert not seen[rows].any()
        seen[rows] = True
        assert np.array_equal(np.load(tmp_path / ("ll_%d.npy" % r)), ll[rows])
    assert seen.all()
