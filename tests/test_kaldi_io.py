"""Wire / disk formats: vbgpu_io_* against the reference's own writers and readers (compiled from
/root/reference into oracle/_ref/libvbref.so) — bit-exact in both directions."""
import numpy as np
import pytest

from voicebridge_b200 import capi, kaldi_io as kio, synth


def _feats(rows, cols, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((rows, cols)) * np.linspace(20.0, 0.5, cols) + rng.standard_normal(cols) * 3).astype(np.float32)


@pytest.mark.parametrize("kind,name", [(0, "FM"), (1, "DM"), (2, "CM"), (3, "CM2"), (4, "CM3")])
@pytest.mark.parametrize("rows,cols", [(141, 13), (3, 39), (1000, 40), (1, 1), (5, 7)])
def test_read_matrix_objects_written_by_the_reference(ref, kind, name, rows, cols):
    m = _feats(rows, cols, 100 * kind + rows)
    b = ref.io_write_matrix(m, kind)
    info = kio.object_info(b)
    # the automatic method keeps matrices of 8 rows or fewer in the 2-byte format (compressed-matrix.cc:88-103)
    assert (info.rows, info.cols) == (rows, cols) and info.total_bytes == len(b)
    assert kio.KINDS[info.kind] == name or (kind == 2 and kio.KINDS[info.kind] in ("CM", "CM2"))
    got = kio.read_matrix(b)
    want = ref.io_read_matrix(b)  # Matrix<BaseFloat>::Read, which expands CompressedMatrix objects itself
    assert got.dtype == np.float32 and np.array_equal(got, want)
    if kind == 0:
        assert np.array_equal(got, m) and kio.write_matrix(m) == b  # and our writer emits the same bytes


def test_empty_and_truncated_objects(ref):
    e = ref.io_write_matrix(np.zeros((0, 0), np.float32), 0)
    assert kio.read_matrix(e).shape == (0, 0) and kio.write_matrix(np.zeros((0, 0), np.float32)) == e
    b = ref.io_write_matrix(_feats(20, 13, 1), 2)
    for cut in (1, 3, 10, len(b) - 1):
        with pytest.raises(capi.VbgpuError):
            kio.read_matrix(b[:cut])
    with pytest.raises(capi.VbgpuError):
        kio.read_matrix(b"\0BXX 123")


def test_int32_vectors_both_ways(ref):
    for n in (0, 1, 7, 1234):
        v = np.random.default_rng(n).integers(-5, 50000, n).astype(np.int32)
        b = ref.io_write_int32_vector(v)
        assert np.array_equal(kio.read_int32_vector(b), v)
        mine = kio.write_int32_vector(v)
        assert mine == b and np.array_equal(ref.io_read_int32_vector(mine), v)


def test_archive_of_mixed_objects(ref):
    """A binary archive as copy-feats / ali-to-pdf write it: keys, compressed and plain matrices, alignments."""
    mats = {"spk1_utt1": (_feats(50, 13, 1), 2), "spk1_utt2": (_feats(9, 13, 2), 0), "s2-u1": (_feats(300, 40, 3), 4)}
    ali = np.arange(77, dtype=np.int32)
    ark = b"".join(kio.write_ark_entry(k, ref.io_write_matrix(m, kind)) for k, (m, kind) in mats.items())
    ark += kio.write_ark_entry("ali1", ref.io_write_int32_vector(ali))
    seen = []
    for key, info, off in kio.read_ark(ark):
        seen.append(key)
        if key in mats:
            assert np.array_equal(kio.read_matrix(ark, off), ref.io_read_matrix(ref.io_write_matrix(*mats[key])))
        else:
            assert kio.KINDS[info.kind] == "IV" and np.array_equal(kio.read_int32_vector(ark, off), ali)
    assert seen == list(mats) + ["ali1"]
    with pytest.raises(capi.VbgpuError):
        list(kio.read_ark(ark[:-5]))


def test_model_file(ref):
    n_phones, D = 4, 13
    m = synth.make_model(3 * n_phones, 40, D, 5)
    b = ref.io_write_mdl(m, n_phones)
    got = kio.read_mdl(b)
    want = ref.io_read_mdl(b, D)
    assert got["dim"] == D and len(got["pdf_offsets"]) == want["num_pdfs"] + 1
    assert np.array_equal(got["pdf_offsets"], m.pdf_offsets)
    assert len(got["tid2pdf"]) == len(want["tid2pdf"]) and len(got["tid2pdf"]) > 1
    assert np.array_equal(got["tid2pdf"][1:], want["tid2pdf"][1:])
    assert np.array_equal(got["trans_log_probs"][1:], want["trans_log_probs"][1:])
    for k in ("gconsts", "weights", "means_invvars", "inv_vars"):  # gconsts: recomputed on read, on both sides
        assert np.array_equal(got[k], want[k]), k
    # a bare AmDiagGmm (no transition model) is accepted too: cut the file at <DIMENSION>
    bare = b[b.index(b"<DIMENSION>"):]
    g2 = kio.read_mdl(bare)
    assert len(g2["tid2pdf"]) == 0 and np.array_equal(g2["gconsts"], got["gconsts"])
    with pytest.raises(capi.VbgpuError):
        kio.read_mdl(b[: len(b) // 2])


def test_statistics_file_read_back_by_the_reference(ref):
    m = synth.make_model(6, 30, 13, 9)
    N, D = len(m.gconsts), 13
    rng = np.random.default_rng(3)
    occ, mean, var = rng.uniform(0, 50, N), rng.standard_normal((N, D)) * 100, rng.uniform(0, 1e4, (N, D))
    trans = rng.uniform(0, 9, 25)
    b = kio.write_acc(m.pdf_offsets, occ, mean, var, -12345.678, 4321.0, trans_accs=trans)
    tr, o2, m2, v2, tl, tf = ref.io_read_acc(m, b, len(trans))
    # (TotLogLike() / TotCount() hand the stored doubles back as BaseFloat, mle-am-diag-gmm.h:81-83)
    assert np.array_equal(tr, trans) and tl == float(np.float32(-12345.678)) and tf == 4321.0
    # AccumDiagGmm::Write narrows to float (mle-diag-gmm.cc:86-101)
    assert np.array_equal(o2, occ.astype(np.float32).astype(np.float64))
    assert np.array_equal(m2, mean.astype(np.float32).astype(np.float64))
    assert np.array_equal(v2, var.astype(np.float32).astype(np.float64))
    b0 = kio.write_acc(m.pdf_offsets, occ, mean, var, 1.0, 2.0)  # without transition accs
    assert ref.io_read_acc(m, b0, 0)[4:] == (1.0, 2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
def test_matrix_to_device_is_bit_exact(ref, kind):
    import torch
    for rows, cols, stride in ((1237, 13, 16), (50, 40, 40), (33, 65, 68)):
        b = ref.io_write_matrix(_feats(rows, cols, 7 + kind), kind)
        want = ref.io_read_matrix(b)
        d_out = torch.full((rows, stride), -7.0, dtype=torch.float32, device="cuda")
        scratch = torch.empty(len(b) + 16, dtype=torch.uint8, device="cuda")
        kio.matrix_to_device(b, d_out, stride, scratch)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        assert np.array_equal(got[:, :cols], want)
        assert (got[:, cols:] == -7.0).all()  # the stride padding is left alone


@pytest.mark.gpu
def test_scorer_from_model_file(ref):
    """final.mdl bytes -> device scorer + tid2pdf: the DecodableInterface view agrees with the reference's own reading
    of the same file."""
    from voicebridge_b200 import host
    from tests.common import assert_ll_close
    n_phones, D = 5, 13
    m = synth.make_model(3 * n_phones, 60, D, 15)
    b = ref.io_write_mdl(m, n_phones)
    am, tid2pdf = host.AmDiagGmmGpu.from_mdl(b)
    want = ref.io_read_mdl(b, D)
    assert np.array_equal(tid2pdf[1:], want["tid2pdf"][1:])
    X = synth.make_feats(m, 64, 16)
    gc, miv, iv = ref.model_params(m.pdf_offsets, m.weights, m.means, m.iv)
    rc, ll = ref.gmm_loglikes(synth.GmmModel(m.pdf_offsets, m.weights, m.means, iv, miv, gc), X)
    assert rc == 0
    assert_ll_close(am.score(X), ll)


def test_parsers_survive_corrupt_input(ref):
    """Truncations and byte flips of real objects must end in a VbgpuError (or a clean parse), never in a crash or an
    out-of-bounds read: the parsers run on files a job did not write itself."""
    rng = np.random.default_rng(11)
    m = synth.make_model(6, 20, 13, 5)
    blobs = {
        "mdl": ref.io_write_mdl(m, 2),
        "cm": ref.io_write_matrix(_feats(40, 13, 1), 2),
        "fm": ref.io_write_matrix(_feats(7, 5, 2), 0),
        "iv": ref.io_write_int32_vector(np.arange(50, dtype=np.int32)),
    }
    ark = b"".join(kio.write_ark_entry("k%d" % i, blobs[k]) for i, k in enumerate(("cm", "fm", "iv")))

    def attempt(fn, b):
        try:
            fn(b)
        except capi.VbgpuError:
            pass

    for name, b in blobs.items():
        fn = {"mdl": kio.read_mdl, "cm": kio.read_matrix, "fm": kio.read_matrix, "iv": kio.read_int32_vector}[name]
        for cut in sorted(set(rng.integers(0, len(b), 60).tolist() + [0, 1, 2, 3, len(b) - 1])):
            with pytest.raises(capi.VbgpuError):
                fn(b[:cut])
        for _ in range(150):
            c = bytearray(b)
            for pos in rng.integers(0, min(len(c), 400), rng.integers(1, 4)):
                c[pos] = rng.integers(0, 256)
            attempt(fn, bytes(c))
    for _ in range(150):
        c = bytearray(ark)
        for pos in rng.integers(0, len(c), rng.integers(1, 4)):
            c[pos] = rng.integers(0, 256)
        attempt(lambda x: [kio.read_matrix(x, off) if info.kind <= 5 else None for _, info, off in kio.read_ark(x)], bytes(c))


@pytest.mark.gpu
def test_wav_files_to_feature_archive_like_make_mfcc(ref):
    """compute-mfcc-feats at file level: RIFF images in, a binary feature archive out, read back by the reference's
    own matrix reader and compared with the reference's own MFCCs of the same waveforms."""
    from oracle import pyoracle as po
    from tests.common import assert_feats_close
    from tests.test_wave import riff
    from voicebridge_b200 import host
    o = capi.default_mfcc_opts(dither=0.0, use_energy=0)
    mfcc = host.Mfcc(o)
    waves = {"utt%d" % i: synth.make_wave(6000 + 3211 * i, 40 + i) for i in range(4)}
    ark = b""
    for key, w in waves.items():
        wd = host.WaveData(riff(w))                      # WaveData::Read
        assert wd.SampFreq() == 16000.0
        feats = mfcc.ComputeFeatures(wd.Data()[0], wd.SampFreq())
        ark += kio.write_ark_entry(key, kio.write_matrix(feats))
    oo = po.default_opts(dither=0.0, use_energy=0)
    seen = 0
    for key, info, off in kio.read_ark(ark):
        back = ref.io_read_matrix(ark[off:off + info.total_bytes])   # the reference parses what we wrote
        assert_feats_close(back, ref.mfcc(oo, waves[key].astype(np.float32)))
        seen += 1
    assert seen == len(waves)
