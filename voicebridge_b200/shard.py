"""Speaker-level sharding of a corpus over ranks (one process per GPU).

Mirrors the reference's job split: `SplitData` keeps every speaker inside one job because CMVN statistics and
fMLLR transforms are per speaker (VB/scr/utils/split_data.cpp:17-27).  Scoring needs no exchange between shards;
training merges the per-rank accumulators with ONE sum all-reduce per EM iteration
(AccumAmDiagGmmGpu.AllReduce, replacing gmm-sum-accs over x.JOBID.acc files, VB/src/gmmbin/gmm-sum-accs.cpp:44-50).
"""
import numpy as np


def shard_speakers(utt2spk, utt_frames, world_size):
    """Greedy longest-processing-time bin packing of speakers by total frames.

    Returns a list (one entry per rank) of sorted utterance-index arrays.  Deterministic; every utterance appears
    exactly once; utterances of one speaker never straddle ranks."""
    utt2spk = np.asarray(utt2spk, np.int64)
    utt_frames = np.asarray(utt_frames, np.int64)
    assert utt2spk.shape == utt_frames.shape and world_size >= 1
    if len(utt2spk) == 0:
        return [np.zeros(0, np.int64) for _ in range(world_size)]
    n_spk = int(utt2spk.max()) + 1
    load = np.bincount(utt2spk, weights=utt_frames, minlength=n_spk)
    order = np.lexsort((np.arange(n_spk), -load))  # heaviest first, ties by speaker id
    rank_load = np.zeros(world_size)
    owner = np.zeros(n_spk, np.int64)
    for s in order:
        r = int(np.argmin(rank_load))  # first minimum: deterministic
        owner[s] = r
        rank_load[r] += load[s]
    return [np.flatnonzero(owner[utt2spk] == r) for r in range(world_size)]


def take_shard(pcm, sample_offsets, utt2spk, utts):
    """Pack the utterances `utts` of a corpus into a contiguous sub-corpus with dense local speaker ids.

    Returns (pcm, sample_offsets, utt2spk_local, n_spk_local, global_speaker_ids)."""
    sample_offsets = np.asarray(sample_offsets, np.int64)
    utt2spk = np.asarray(utt2spk, np.int64)
    lens = sample_offsets[1:] - sample_offsets[:-1]
    so = np.zeros(len(utts) + 1, np.int64)
    so[1:] = np.cumsum(lens[utts])
    out = np.empty(int(so[-1]), dtype=pcm.dtype)
    for i, u in enumerate(utts):
        out[so[i]:so[i + 1]] = pcm[sample_offsets[u]:sample_offsets[u + 1]]
    spk_ids, local = np.unique(utt2spk[utts], return_inverse=True)
    return out, so, local.astype(np.int32), len(spk_ids), spk_ids


def imbalance(shards, utt_frames):
    """max rank load / mean rank load (1.0 = perfect); the scaling bound of a no-collective sharded job."""
    utt_frames = np.asarray(utt_frames, np.float64)
    loads = np.array([utt_frames[s].sum() for s in shards])
    return float(loads.max() / max(loads.mean(), 1e-30))
