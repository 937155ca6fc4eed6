"""Host-side logic that needs no GPU: synthetic generators, sharding, frame bookkeeping."""
import numpy as np

from oracle import pyoracle as po
from voicebridge_b200 import shard, synth


def test_model_generator_shapes_and_gconsts(orc):
    m = synth.make_model(200, 1000, 39, 1)
    assert m.num_pdfs == 200 and m.num_gauss == 1000 and m.dim == 39
    sizes = np.diff(m.pdf_offsets)
    assert sizes.min() >= 1 and sizes.sum() == 1000 and sizes.max() > sizes.min()
    for p in (0, 17, 199):
        s = slice(m.pdf_offsets[p], m.pdf_offsets[p + 1])
        assert abs(m.weights[s].sum() - 1.0) < 1e-5
    # numpy ComputeGconsts used for synthetic models agrees with the oracle's restatement
    assert np.abs(orc.gconsts(m.weights, m.miv, m.iv) - m.gconsts).max() < 1e-4


def test_corpus_and_frame_offsets(orc):
    pcm, so, u2s = synth.make_corpus(3, 4, 0.5, 1.5, 7)
    assert pcm.dtype == np.int16 and len(so) == 13 and so[-1] == len(pcm) and len(u2s) == 12
    o = po.default_opts()
    frames = [orc.num_frames(int(so[i + 1] - so[i]), o) for i in range(12)]
    assert min(frames) >= 48 and max(frames) <= 148


def test_shard_speakers_properties():
    rng = np.random.default_rng(0)
    u2s = np.repeat(np.arange(37), rng.integers(1, 40, 37))
    rng.shuffle(u2s)
    frames = rng.integers(100, 3000, len(u2s))
    for ws in (1, 2, 4, 8):
        sh = shard.shard_speakers(u2s, frames, ws)
        allu = np.concatenate(sh)
        assert sorted(allu.tolist()) == list(range(len(u2s)))           # a partition
        owners = {}
        for r, s in enumerate(sh):
            for spk in set(u2s[s].tolist()):
                assert owners.setdefault(spk, r) == r                   # a speaker never straddles ranks
        assert shard.imbalance(sh, frames) < 1.15
        sh2 = shard.shard_speakers(u2s, frames, ws)
        assert all(np.array_equal(a, b) for a, b in zip(sh, sh2))       # deterministic


def test_take_shard_packs_contiguously():
    pcm, so, u2s = synth.make_corpus(4, 3, 0.2, 0.4, 3)
    utts = np.array([1, 4, 5, 9])
    p2, so2, local, n_spk, spk_ids = shard.take_shard(pcm, so, u2s, utts)
    assert n_spk == 3 and list(spk_ids) == [0, 1, 3] and list(local) == [0, 1, 1, 2]
    for i, u in enumerate(utts):
        assert np.array_equal(p2[so2[i]:so2[i + 1]], pcm[so[u]:so[u + 1]])


def test_paste_feats_length_tolerance():
    """paste-feats' AppendFeats (VB/src/featbin/paste-feats.cpp:25-65): trim to the shortest within the tolerance, drop
    the utterance beyond it or when one input is empty."""
    from voicebridge_b200 import host
    a = np.arange(10 * 3, dtype=np.float32).reshape(10, 3)
    b = np.arange(12 * 2, dtype=np.float32).reshape(12, 2) + 100
    out = host.paste_feats([a, b], 2)
    assert out.shape == (10, 5) and np.array_equal(out[:, :3], a) and np.array_equal(out[:, 3:], b[:10])
    assert host.paste_feats([a, b], 1) is None
    assert host.paste_feats([a, b[:0]], 20) is None
    assert host.paste_feats([a, a], 0).shape == (10, 6)
