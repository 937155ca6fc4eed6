// wave.cu — RIFF/RIFX 16-bit PCM container parsing: the host-side I/O step in front of the path (WaveInfo::Read /
// WaveData::Read, feat/wave-reader.cc:119-310).  Pure byte handling on an in-memory file image, no device work: the
// samples stay int16 (the reference widens them to float without rescaling, wave-reader.cc:302-309; the MFCC kernel does
// that widening on the device).  Accepts what the reference accepts: PCM (format 1) and WAVE_FORMAT_EXTENSIBLE with the
// PCM sub-format, 16 bits per sample, any number of channels, chunks between "fmt " and "data" skipped, "stream mode"
// sizes (0, 0xFFFFFFFF, SoX's 0x7FFFF000) meaning "data runs to the end of the image", a truncated data chunk.
#include "common.h"

namespace {

struct Cursor {
  const uint8_t *p;
  size_t n, pos;
  bool swap, ok;
  void tag(char *t) {
    if (pos + 4 > n) {
      ok = false;
      t[0] = 0;
      return;
    }
    std::memcpy(t, p + pos, 4);
    t[4] = 0;
    pos += 4;
  }
  uint32_t u32() {
    if (pos + 4 > n) {
      ok = false;
      return 0;
    }
    const uint8_t *b = p + pos;
    pos += 4;
    return swap ? ((uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3])
                : ((uint32_t)b[3] << 24 | (uint32_t)b[2] << 16 | (uint32_t)b[1] << 8 | b[0]);
  }
  uint16_t u16() {
    if (pos + 2 > n) {
      ok = false;
      return 0;
    }
    const uint8_t *b = p + pos;
    pos += 2;
    return swap ? (uint16_t)(b[0] << 8 | b[1]) : (uint16_t)(b[1] << 8 | b[0]);
  }
};

}  // namespace

using vb::fail;

extern "C" {

int vbgpu_wave_parse(const void *bytes, size_t n_bytes, vbgpu_wave_info *info) {
  VB_CHECK(bytes && info, "null argument");
  std::memset(info, 0, sizeof(*info));
  Cursor c{static_cast<const uint8_t *>(bytes), n_bytes, 0, false, true};
  char t[5];
  c.tag(t);
  if (std::strcmp(t, "RIFF") == 0) c.swap = false;
  else if (std::strcmp(t, "RIFX") == 0) c.swap = true;
  else return fail(VBGPU_ERR_INVALID, "WaveData: expected RIFF or RIFX, got %s", t);
  const uint32_t riff_size = c.u32();
  c.tag(t);
  VB_CHECK(c.ok && std::strcmp(t, "WAVE") == 0, "WaveData: expected WAVE, got %s", t);
  c.tag(t);
  VB_CHECK(c.ok && std::strcmp(t, "fmt ") == 0, "WaveData: expected fmt chunk, got %s", t);
  const uint32_t fmt_size = c.u32();
  const uint16_t format = c.u16(), channels = c.u16();
  const uint32_t rate = c.u32(), byte_rate = c.u32();
  const uint16_t block_align = c.u16(), bits = c.u16();
  VB_CHECK(c.ok, "WaveData: unexpected end of file or read error");
  uint32_t fmt_read = 16;
  if (format == 1) {
    VB_CHECK(fmt_size >= 16, "WaveData: expect PCM format data to have fmt chunk of at least size 16.");
  } else if (format == 0xFFFE) {  // WAVE_FORMAT_EXTENSIBLE, PCM sub-format only (wave-reader.cc:150-172)
    const uint16_t extra = c.u16();
    VB_CHECK(c.ok && fmt_size >= 40 && extra >= 22, "WaveData: malformed WAVE_FORMAT_EXTENSIBLE format data.");
    c.u16();
    c.u32();
    const uint32_t g1 = c.u32(), g2 = c.u32(), g3 = c.u32(), g4 = c.u32();
    fmt_read = 40;
    VB_CHECK(c.ok && g1 == 0x00000001u && g2 == 0x00100000u && g3 == 0xAA000080u && g4 == 0x719B3800u,
             "WaveData: unsupported WAVE_FORMAT_EXTENSIBLE format.");
  } else {
    return fail(VBGPU_ERR_INVALID, "WaveData: can read only PCM data, format id in file is: %d", (int)format);
  }
  VB_CHECK(fmt_size >= fmt_read && c.pos + (fmt_size - fmt_read) <= c.n, "WaveData: unexpected end of file or read error");
  c.pos += fmt_size - fmt_read;
  VB_CHECK(channels != 0, "WaveData: no channels present");
  VB_CHECK(bits == 16, "WaveData: unsupported bits_per_sample = %d", (int)bits);
  VB_CHECK(byte_rate == rate * 2u * channels, "Unexpected byte rate %u vs. %u * 2 * %d", byte_rate, rate, (int)channels);
  VB_CHECK(block_align == channels * 2, "Unexpected block_align: %d vs. %d * 2", (int)block_align, (int)channels);
  c.tag(t);
  while (c.ok && std::strcmp(t, "data") != 0) {  // "fact", "LIST", ... are skipped (wave-reader.cc:196-210)
    const uint32_t sz = c.u32();
    VB_CHECK(c.ok && c.pos + sz <= c.n, "WaveData: unexpected end of file or read error");
    c.pos += sz;
    c.tag(t);
  }
  VB_CHECK(c.ok, "WaveData: unexpected end of file or read error");
  const uint32_t data_size = c.u32();
  VB_CHECK(c.ok, "WaveData: unexpected end of file or read error");
  const bool stream = riff_size == 0 || riff_size == 0xFFFFFFFFu || data_size == 0 || data_size == 0xFFFFFFFFu ||
                      data_size == 0x7FFFF000u;
  size_t avail = c.n - c.pos;
  if (!stream && data_size < avail) avail = data_size;  // a shorter image is a truncated file: keep what is there
  VB_CHECK(avail > 0, "WaveData: empty file (no data)");
  info->samp_freq = static_cast<float>(rate);
  info->num_channels = channels;
  info->num_samples = static_cast<int64_t>(avail / block_align);
  info->data_offset = static_cast<int64_t>(c.pos);
  info->reverse_bytes = c.swap ? 1 : 0;
  return 0;
}

int vbgpu_wave_channel_i16(const void *bytes, size_t n_bytes, const vbgpu_wave_info *info, int32_t channel, int16_t *out) {
  VB_CHECK(bytes && info && out, "null argument");
  if (channel < 0) channel = 0;  // compute-mfcc-feats --channel=-1: the first channel (compute-mfcc-feats.cpp:120-135)
  VB_CHECK(channel < info->num_channels, "File has %d channels; channel %d requested", info->num_channels, channel);
  VB_CHECK(info->data_offset >= 0 && (size_t)info->data_offset + (size_t)info->num_samples * info->num_channels * 2 <= n_bytes,
           "wave info does not match the image");
  const uint8_t *d = static_cast<const uint8_t *>(bytes) + info->data_offset + 2 * channel;
  const size_t step = 2 * (size_t)info->num_channels;
  for (int64_t i = 0; i < info->num_samples; i++, d += step)
    out[i] = info->reverse_bytes ? (int16_t)(uint16_t)(d[0] << 8 | d[1]) : (int16_t)(uint16_t)(d[1] << 8 | d[0]);
  return 0;
}

}  // extern "C"
