"""WAV container parsing (vbgpu_wave_*, the WaveData::Read step): against Python's own `wave` writer, hand-built
RIFF/RIFX images covering what the reference accepts and rejects (feat/wave-reader.cc:119-310), and the compiled
reference where it is available.  Host byte handling: no GPU needed."""
import ctypes as C
import io
import struct
import wave

import numpy as np
import pytest

from voicebridge_b200 import capi, host


def riff(samples, rate=16000, channels=1, fmt_tag=1, extra_chunks=(), big=False, riff_size=None, data_size=None,
         bits=16, truncate=0, fmt_extra=b""):
    e = ">" if big else "<"
    pcm = np.asarray(samples, np.int16).astype(e + "i2").tobytes()
    block = channels * bits // 8
    if fmt_tag == 0xFFFE:
        guid = struct.pack(e + "IIII", 0x00000001, 0x00100000, 0xAA000080, 0x719B3800)
        fmt = struct.pack(e + "HHIIHH", fmt_tag, channels, rate, rate * block, block, bits) + \
            struct.pack(e + "HHI", 22, bits, 3) + guid
    else:
        fmt = struct.pack(e + "HHIIHH", fmt_tag, channels, rate, rate * block, block, bits) + fmt_extra
    body = b"WAVE" + b"fmt " + struct.pack(e + "I", len(fmt)) + fmt
    for tag, payload in extra_chunks:
        body += tag + struct.pack(e + "I", len(payload)) + payload
    body += b"data" + struct.pack(e + "I", len(pcm) if data_size is None else data_size) + pcm
    img = (b"RIFX" if big else b"RIFF") + struct.pack(e + "I", len(body) if riff_size is None else riff_size) + body
    return img[:len(img) - truncate] if truncate else img


RNG = np.random.default_rng(5)
MONO = RNG.integers(-32768, 32767, size=4001).astype(np.int16)
STEREO = RNG.integers(-32768, 32767, size=(2, 1500)).astype(np.int16)

CASES = {
    "mono": (riff(MONO), MONO[None, :], 16000),
    "stereo_8k": (riff(STEREO.T.reshape(-1), rate=8000, channels=2), STEREO, 8000),
    "fact_and_list_chunks": (riff(MONO, extra_chunks=[(b"fact", b"\0\0\0\0"), (b"LIST", b"x" * 26)]), MONO[None, :], 16000),
    "extensible_pcm": (riff(MONO, fmt_tag=0xFFFE), MONO[None, :], 16000),
    "long_fmt_chunk": (riff(MONO, fmt_extra=b"\0\0"), MONO[None, :], 16000),
    "rifx_big_endian": (riff(STEREO.T.reshape(-1), channels=2, big=True), STEREO, 16000),
    "stream_mode_sox": (riff(MONO, data_size=0x7FFFF000), MONO[None, :], 16000),
    "stream_mode_zero": (riff(MONO, riff_size=0, data_size=0), MONO[None, :], 16000),
    "stream_mode_ffffffff": (riff(MONO, riff_size=0xFFFFFFFF, data_size=0xFFFFFFFF), MONO[None, :], 16000),
    "truncated_data": (riff(MONO, truncate=801), MONO[None, :3600], 16000),
    "odd_trailing_byte": (riff(MONO) + b"\0", MONO[None, :], 16000),
}
BAD = {
    "not_riff": b"RIFQ" + riff(MONO)[4:],
    "not_wave": riff(MONO)[:8] + b"WAVX" + riff(MONO)[12:],
    "float_format": riff(MONO, fmt_tag=3),
    "eight_bit": riff(MONO, bits=8),
    "no_data": riff([]),
    "header_only": riff(MONO)[:30],
    "no_channels": riff(MONO, channels=0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_wave_read(name):
    img, want, rate = CASES[name]
    w = host.WaveData(img)
    assert w.SampFreq() == rate
    assert w.Data().shape == want.shape
    assert np.array_equal(w.Data(), want)          # bit-exact: byte work
    assert abs(w.Duration() - want.shape[1] / rate) < 1e-9


@pytest.mark.parametrize("name", sorted(BAD))
def test_wave_rejects(name):
    with pytest.raises(capi.VbgpuError) as e:
        host.WaveData(BAD[name])
    assert e.value.code == capi.ERR_INVALID


def test_wave_against_python_wave_module(tmp_path):
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as f:
        f.setnchannels(2)
        f.setsampwidth(2)
        f.setframerate(22050)
        f.writeframes(STEREO.T.astype("<i2").tobytes())
    w = host.WaveData.Read(path)
    assert w.SampFreq() == 22050 and np.array_equal(w.Data(), STEREO)


def test_wave_channel_selection_errors():
    img = CASES["stereo_8k"][0]
    buf = np.frombuffer(img, np.uint8)
    info = capi.WaveInfo()
    assert capi.lib().vbgpu_wave_parse(buf.ctypes.data, buf.size, C.byref(info)) == 0
    out = np.zeros(info.num_samples, np.int16)
    assert capi.lib().vbgpu_wave_channel_i16(buf.ctypes.data, buf.size, C.byref(info), -1, out.ctypes.data) == 0
    assert np.array_equal(out, STEREO[0])          # --channel=-1 -> first channel
    assert capi.lib().vbgpu_wave_channel_i16(buf.ctypes.data, buf.size, C.byref(info), 2, out.ctypes.data) == capi.ERR_INVALID


def test_wave_matches_compiled_reference(ref):
    """Same accept / reject decisions and the same samples as the reference's WaveData::Read."""
    fn = ref.lib.ref_wave_read
    fn.restype = C.c_int64
    for name, img in list((k, v[0]) for k, v in CASES.items()) + list(BAD.items()):
        sf, ch = C.c_float(0), C.c_int32(0)
        data = np.zeros(20000, np.float32)
        n = fn(img, C.c_int64(len(img)), C.byref(sf), C.byref(ch), data.ctypes.data_as(C.c_void_p), C.c_int64(data.size))
        try:
            w = host.WaveData(img)
        except capi.VbgpuError:
            assert n < 0, "%s: the reference accepts what we reject" % name
            continue
        assert n >= 0, "%s: the reference rejects what we accept" % name
        assert (w.SampFreq(), w.Data().shape) == (sf.value, (ch.value, n)), name
        assert np.array_equal(w.Data().astype(np.float32), data[:ch.value * n].reshape(ch.value, n)), name
