// frontend.cu — fused MFCC front end for sm_100a: framing -> [dither] -> DC removal -> raw log-energy ->
// pre-emphasis -> window -> real FFT -> power spectrum -> mel filterbank -> log -> DCT -> lifter -> C0/energy.
//
// Replaces OfflineFeatureTpl<MfccComputer>::Compute (feat/feature-common-inl.h:61-98), ExtractWindow/ProcessWindow
// (feat/feature-window.cc:133-220), SplitRadixRealFft::Compute (matrix/srfft.cc:362-431), ComputePowerSpectrum
// (feat/feature-functions.cc:29-51), MelBanks::Compute (feat/mel-computations.cc:228-253) and MfccComputer::Compute
// (feat/feature-mfcc.cc:28-80) of the reference with ONE kernel: a warp owns a frame; the N-point real FFT is an
// (N/2)-point complex FFT held entirely in registers (E = N/64 complex values per lane: a radix-E pass inside the lane,
// then five radix-2 passes across lanes with warp shuffles); only the 257-bin power spectrum touches shared memory.
// HBM traffic per frame is the algorithmic minimum: `shift` new int16 samples in (overlap is served by L1/L2) and
// one 64-byte MFCC row out.
#include <cfloat>
#include <cmath>

#include "common.h"

namespace vb {

// ---------------------------------------------------------------------------------------------------------------
// Host-side tables (computed once per handle).
// ---------------------------------------------------------------------------------------------------------------
static inline float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }      // mel-computations.h:85-87
static inline float inv_mel_scale(float m) { return 700.0f * (expf(m / 1127.0f) - 1.0f); }  // mel-computations.h:81-83

// Piecewise-linear VTLN warp, mel-computations.cc:152-224.
static float vtln_warp_mel(float vlow, float vhigh, float low, float high, float warp, float mel) {
  float f = inv_mel_scale(mel), out;
  if (f < low || f > high) {
    out = f;
  } else {
    float l = vlow * fmaxf(1.0f, warp), h = vhigh * fminf(1.0f, warp), scale = 1.0f / warp;
    float Fl = scale * l, Fh = scale * h;
    float sl = (Fl - low) / (l - low), sr = (high - Fh) / (high - h);
    if (f < l) out = low + sl * (f - low);
    else if (f < h) out = scale * f;
    else out = high + sr * (f - high);
  }
  return mel_scale(out);
}

// Triangular mel filters as (first bin, length, weights) triples, mel-computations.cc:33-144.
static int build_mel(const vbgpu_mfcc_opts &o, int npad, float warp, std::vector<int32_t> *off, std::vector<int32_t> *len,
                     std::vector<float> *w, int pitch) {
  const int B = o.num_bins, nfft = npad / 2;
  const float fs = o.samp_freq, nyq = 0.5f * fs;
  const float low = o.low_freq, high = o.high_freq > 0.0f ? o.high_freq : nyq + o.high_freq;
  if (low < 0.0f || low >= nyq || high <= 0.0f || high > nyq || high <= low)
    return fail(VBGPU_ERR_INVALID, "bad mel options: low-freq %g high-freq %g nyquist %g", low, high, nyq);
  const float bin_width = fs / npad, mlow = mel_scale(low), mhigh = mel_scale(high), delta = (mhigh - mlow) / (B + 1);
  float vlow = o.vtln_low, vhigh = o.vtln_high;
  if (vhigh < 0.0f) vhigh += nyq;
  if (warp != 1.0f && (vlow < 0.0f || vlow <= low || vlow >= high || vhigh <= 0.0f || vhigh >= high || vhigh <= vlow))
    return fail(VBGPU_ERR_INVALID, "bad vtln-low %g / vtln-high %g", vlow, vhigh);
  off->assign(B, 0);
  len->assign(B, 0);
  w->assign((size_t)B * pitch, 0.0f);
  std::vector<float> tmp(nfft);
  for (int b = 0; b < B; b++) {
    float lm = mlow + b * delta, cm = mlow + (b + 1) * delta, rm = mlow + (b + 2) * delta;
    if (warp != 1.0f) {
      lm = vtln_warp_mel(vlow, vhigh, low, high, warp, lm);
      cm = vtln_warp_mel(vlow, vhigh, low, high, warp, cm);
      rm = vtln_warp_mel(vlow, vhigh, low, high, warp, rm);
    }
    int first = -1, last = -1;
    for (int i = 0; i < nfft; i++) {
      float mel = mel_scale(bin_width * i);
      tmp[i] = 0.0f;
      if (mel > lm && mel < rm) {
        tmp[i] = mel <= cm ? (mel - lm) / (cm - lm) : (rm - mel) / (rm - cm);
        if (first < 0) first = i;
        last = i;
      }
    }
    if (first < 0) return fail(VBGPU_ERR_INVALID, "empty mel bin %d: num-mel-bins too large", b);
    int n = last + 1 - first;
    if (n > pitch) return fail(VBGPU_ERR_INVALID, "mel bin %d spans %d FFT bins (> %d)", b, n, pitch);
    (*off)[b] = first;
    (*len)[b] = n;
    for (int i = 0; i < n; i++) (*w)[(size_t)b * pitch + i] = tmp[first + i];
    if (o.htk_mode && b == 0 && mlow != 0.0f) (*w)[0] = 0.0f;
  }
  return 0;
}

}  // namespace vb

// ---------------------------------------------------------------------------------------------------------------
// Device code.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct MfccParams {
  const void *pcm;
  const int64_t *sample_offsets, *frame_offsets;
  const int32_t *frame2utt, *utt_mel;  // utt_mel nullable: per-utterance mel-table index
  int64_t total_frames;
  int32_t L, shift, snip_edges, remove_dc, use_energy, raw_energy, htk_compat, htk_mode, B, C, use_lifter, mel_pitch, n_mel;
  int32_t fbank, use_log_fbank, use_power;  // fbank != 0: FbankComputer's tail (feature-fbank.cc:97-121) instead of the DCT
  int32_t plp, lpc_order;                   // plp != 0: PlpComputer's tail (feature-plp.cc:143-188)
  float compress_factor, cepstral_scale;
  const float *idft;                        // [lpc_order + 1][B + 2]  (InitIdftBases)
  const float *eq_loud;                     // [n_mel][B]              (GetEqualLoudnessVector per mel table)
  float preemph, energy_floor, log_energy_floor, dither;
  uint32_t seed;
  const float *window;  // [npad], zero beyond L
  const float2 *tw;     // [npad/2] : exp(-2 pi i k / npad)
  const int32_t *mel_off, *mel_len;
  const float *mel_w;   // [n_mel][B][mel_pitch]
  const float *dct;     // [C][B]
  const float *lifter;  // [C]
  float *out;
  int32_t out_stride;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int E>
__device__ __forceinline__ constexpr int bitrev(int x) {
  int r = 0;
  for (int b = 1; b < E; b <<= 1) {
    r = (r << 1) | (x & 1);
    x >>= 1;
  }
  return r;
}

// exp(-2 pi i j / 32), j in [0,16): folds to literals once the caller's loops are unrolled.
__device__ __forceinline__ float2 w32(int j) {
  switch (j) {
    case 0: return make_float2(1.0f, -0.0f);
    case 1: return make_float2(0.98078528040323043f, -0.19509032201612825f);
    case 2: return make_float2(0.92387953251128674f, -0.38268343236508978f);
    case 3: return make_float2(0.83146961230254524f, -0.55557023301960218f);
    case 4: return make_float2(0.70710678118654757f, -0.70710678118654757f);
    case 5: return make_float2(0.55557023301960229f, -0.83146961230254524f);
    case 6: return make_float2(0.38268343236508984f, -0.92387953251128674f);
    case 7: return make_float2(0.19509032201612833f, -0.98078528040323043f);
    case 8: return make_float2(0.0f, -1.0f);
    case 9: return make_float2(-0.19509032201612819f, -0.98078528040323043f);
    case 10: return make_float2(-0.38268343236508973f, -0.92387953251128674f);
    case 11: return make_float2(-0.55557023301960196f, -0.83146961230254546f);
    case 12: return make_float2(-0.70710678118654746f, -0.70710678118654757f);
    case 13: return make_float2(-0.83146961230254535f, -0.55557023301960218f);
    case 14: return make_float2(-0.92387953251128674f, -0.38268343236508989f);
    default: return make_float2(-0.98078528040323043f, -0.19509032201612861f);
  }
}

// In-register decimation-in-frequency FFT of size E; result for frequency k sits in v[bitrev<E>(k)].
template <int E>
__device__ __forceinline__ void local_fft(float2 (&v)[E]) {
#pragma unroll
  for (int len = E; len >= 2; len >>= 1) {
    const int half = len >> 1;
#pragma unroll
    for (int base = 0; base < E; base += len) {
#pragma unroll
      for (int k = 0; k < half; k++) {
        float2 a = v[base + k], c = v[base + k + half];
        v[base + k] = cadd(a, c);
        float2 d = csub(a, c);
        const int j = k * (32 / len);  // W_len^k = W_32^(k*32/len)
        if (j == 0) v[base + k + half] = d;
        else if (j == 8) v[base + k + half] = make_float2(d.y, -d.x);
        else v[base + k + half] = cmul(d, w32(j));
      }
    }
  }
}

__device__ __forceinline__ uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}

// Two N(0,1) samples for the sample pair (2j, 2j+1) of frame t (Box-Muller on two hashed uniforms).  Stands in for the
// reference's per-frame RandGauss() dither (feature-window.cc:90-98), which uses libc rand() and is not reproducible.
__device__ __forceinline__ float2 gauss_pair(uint32_t seed, uint32_t t, uint32_t j) {
  uint32_t h1 = hash3(seed, t, 2 * j), h2 = hash3(seed ^ 0xA511E9B3u, t, 2 * j + 1);
  float u1 = ((h1 >> 8) + 1) * (1.0f / 16777217.0f), u2 = (h2 >> 8) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * __logf(u1)), s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// Edge frames of snip_edges=false mirror the signal (feature-window.cc:195-211).  Rare, so kept out of line: the frame
// loop is one long unrolled body and every inlined copy of this loop costs instruction-cache space.
__device__ __noinline__ int64_t reflect_index(int64_t k, int64_t ns) {
  while (k < 0 || k >= ns) k = (k < 0) ? -k - 1 : 2 * ns - 1 - k;
  return k;
}
template <typename SampleT>
__device__ __forceinline__ float load_sample(const SampleT *p, int64_t k, int64_t ns, bool reflect) {
  if (reflect) k = reflect_index(k, ns);
  return static_cast<float>(p[k]);
}

// Samples (i, i+1) of a frame that lies wholly inside its utterance.  `wide`: the pair sits on a 2*sizeof(SampleT)
// boundary (warp-uniform), so one load fetches both; otherwise two scalar loads.  All loads of a frame are independent
// and predicated, not branched, so they are in flight together (the generic path below serialises on its reflect check).
__device__ __forceinline__ void load_pair(const int16_t *q, bool wide, bool in0, bool in1, float &x0, float &x1) {
  x0 = 0.0f;
  x1 = 0.0f;
  if (wide) {
    if (in1) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(q));
      x0 = static_cast<float>(static_cast<int16_t>(w & 0xffffu));
      x1 = static_cast<float>(static_cast<int16_t>(w >> 16));
    } else if (in0) {
      x0 = static_cast<float>(__ldg(q));
    }
  } else {
    if (in0) x0 = static_cast<float>(__ldg(q));
    if (in1) x1 = static_cast<float>(__ldg(q + 1));
  }
}
__device__ __forceinline__ void load_pair(const float *q, bool wide, bool in0, bool in1, float &x0, float &x1) {
  x0 = 0.0f;
  x1 = 0.0f;
  if (wide) {
    if (in1) {
      const float2 w = __ldg(reinterpret_cast<const float2 *>(q));
      x0 = w.x;
      x1 = w.y;
    } else if (in0) {
      x0 = __ldg(q);
    }
  } else {
    if (in0) x0 = __ldg(q);
    if (in1) x1 = __ldg(q + 1);
  }
}

constexpr int kWarpsPerBlock = 8;

// Shared-memory map of mfcc_kernel (byte offsets, every array on a 16-byte boundary), shared by the kernel and its launcher.
struct MfccSmem {
  int tw, wl, wp, win, dct, lift, lm, moff, mlen, melw, ps, total, ds;
};
__host__ __device__ inline MfccSmem mfcc_smem(int E, int C, int B, int n_mel, int mel_pitch) {
  const int n = 32 * E, NPAD = 64 * E, PS = n + 8;
  MfccSmem L;
  // row stride of the DCT table / log-mel vector: 4 * odd >= B, so a warp's float4 reads of 32 rows are conflict-free
  L.ds = (B + 3) / 4 * 4;
  if ((L.ds / 4) % 2 == 0) L.ds += 4;
  int o = 0;
  auto take = [&o](int bytes) {
    const int at = o;
    o += (bytes + 15) / 16 * 16;
    return at;
  };
  L.tw = take(8 * n);
  L.wl = take(8 * E * 32);
  L.wp = take(8 * E * 32);
  L.win = take(4 * NPAD);
  L.dct = take(4 * 32 * L.ds);
  L.lift = take(4 * C);
  L.lm = take(4 * kWarpsPerBlock * L.ds);
  L.moff = take(4 * n_mel * B);
  L.mlen = take(4 * n_mel * B);
  L.melw = take(4 * B * mel_pitch);
  L.ps = take(4 * kWarpsPerBlock * PS);
  L.total = o;
  return L;
}

// E complex values per lane; n = 32E complex points; frame padded to NPAD = 64E real samples.
// PLP is a separate instantiation: its tail calls double-precision log() and would otherwise cost the MFCC kernel registers.
// Four resident CTAs (32 warps) per SM: the kernel is latency-bound, and measured on B200 (bench batch, 1.26 M frames)
// 3 CTAs at 80 registers run 2.35 ms, 4 at 64 registers 2.00 ms, 5 at 48 registers (spilling) 2.06 ms.
template <int E, typename SampleT, bool DITHER, bool PLP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4) mfcc_kernel(const MfccParams p) {
  constexpr int n = 32 * E, NPAD = 64 * E, PS = n + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const MfccSmem L = mfcc_smem(E, p.C, p.B, p.n_mel, p.mel_pitch);
  float2 *s_tw = reinterpret_cast<float2 *>(smem_raw + L.tw);      // [n]
  // Per-lane twiddles, [m][lane] so that a warp reads consecutive words (the plain table is indexed with lane-dependent
  // strides, up to 16-way bank conflicts): s_wl = factor applied after the in-lane radix-E pass, s_wp = post-pass factor.
  float2 *s_wl = reinterpret_cast<float2 *>(smem_raw + L.wl);      // [E][32]
  float2 *s_wp = reinterpret_cast<float2 *>(smem_raw + L.wp);      // [E][32]
  float *s_win = reinterpret_cast<float *>(smem_raw + L.win);      // [NPAD]
  float *s_dct = reinterpret_cast<float *>(smem_raw + L.dct);      // [32][ds]: rows >= C and columns >= B are zero
  float *s_lift = reinterpret_cast<float *>(smem_raw + L.lift);    // [C]
  float *s_lm = reinterpret_cast<float *>(smem_raw + L.lm);        // [warps][ds]: log mel energies of the frame in flight
  int32_t *s_moff = reinterpret_cast<int32_t *>(smem_raw + L.moff);  // [n_mel*B]
  int32_t *s_mlen = reinterpret_cast<int32_t *>(smem_raw + L.mlen);  // [n_mel*B]
  float *s_melw = reinterpret_cast<float *>(smem_raw + L.melw);    // [B*mel_pitch] (table 0 only)
  float *s_ps = reinterpret_cast<float *>(smem_raw + L.ps);        // [warps][PS]: power spectrum, 8 zero words behind it
  const int ds = L.ds;

  for (int i = threadIdx.x; i < n; i += blockDim.x) s_tw[i] = p.tw[i];
  for (int i = threadIdx.x; i < NPAD; i += blockDim.x) s_win[i] = p.window[i];
  for (int i = threadIdx.x; i < 32 * ds; i += blockDim.x) {
    const int c = i / ds, b = i - c * ds;
    s_dct[i] = (c < p.C && b < p.B) ? p.dct[c * p.B + b] : 0.0f;
  }
  for (int i = threadIdx.x; i < p.C; i += blockDim.x) s_lift[i] = p.lifter[i];
  for (int i = threadIdx.x; i < kWarpsPerBlock * ds; i += blockDim.x) s_lm[i] = 0.0f;
  for (int i = threadIdx.x; i < p.n_mel * p.B; i += blockDim.x) {
    s_moff[i] = p.mel_off[i];
    s_mlen[i] = p.mel_len[i];
  }
  for (int i = threadIdx.x; i < p.B * p.mel_pitch; i += blockDim.x) s_melw[i] = p.mel_w[i];
  for (int i = threadIdx.x; i < kWarpsPerBlock * PS; i += blockDim.x) s_ps[i] = 0.0f;
  __syncthreads();
  for (int i = threadIdx.x; i < E * 32; i += blockDim.x) {
    const int m = i >> 5, l = i & 31, k1 = bitrev<E>(m);
    int idx = 2 * l * k1;  // W_n^(l*k1) = W_N^(2*l*k1), N = 2n
    const bool neg = idx >= n;
    if (neg) idx -= n;
    float2 w = s_tw[idx];
    if (neg) w = make_float2(-w.x, -w.y);
    s_wl[i] = w;
    s_wp[i] = s_tw[k1 + E * (int)(__brev((unsigned)l) >> 27)];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *ps = s_ps + warp * PS;
  const int k2 = __brev((unsigned)lane) >> 27;  // after the cross-lane DIF passes this lane holds frequency index k2

  // Per-lane twiddles of the five cross-lane stages: W_(2*half)^(lane mod half) = tw[(lane & (half-1)) * n / half].
  float2 wst[5];
  float sgn[5];
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const int half = 16 >> s;
    const bool upper = (lane & half) != 0;
    wst[s] = upper ? s_tw[(lane & (half - 1)) * (n / half)] : make_float2(1.0f, 0.0f);
    sgn[s] = upper ? -1.0f : 1.0f;
  }

  const SampleT *pcm = static_cast<const SampleT *>(p.pcm);

  for (int64_t t = (int64_t)blockIdx.x * kWarpsPerBlock + warp; t < p.total_frames;
       t += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int u = p.frame2utt[t];
    const int64_t s0 = p.sample_offsets[u], ns = p.sample_offsets[u + 1] - s0;
    const int64_t r = t - p.frame_offsets[u];
    const int64_t start = p.snip_edges ? r * p.shift : r * p.shift + p.shift / 2 - p.L / 2;  // feature-window.cc:28-39
    const bool reflect = !(start >= 0 && start + p.L <= ns);
    const SampleT *up = pcm + s0;
    const int mt = p.utt_mel ? p.utt_mel[u] : 0;

    // ---- gather: lane owns complex points j = lane + 32 m, i.e. samples (2j, 2j+1) ------------------------------
    // sample i = 2 (lane + 32 m) lies inside the frame iff 64 m < hl; sample i + 1 iff 64 m + 1 < hl
    const int hl = p.L - 2 * lane;
    float a0[E], a1[E];
    float sum = 0.0f;
    if (!reflect) {  // every frame of snip_edges=true and all but the edge frames otherwise
      const SampleT *fp = up + start;
      const bool wide = (reinterpret_cast<uintptr_t>(fp) & (2 * sizeof(SampleT) - 1)) == 0;
#pragma unroll
      for (int m = 0; m < E; m++) load_pair(fp + 2 * (lane + 32 * m), wide, 64 * m < hl, 64 * m + 1 < hl, a0[m], a1[m]);
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = 2 * (lane + 32 * m);
        a0[m] = (64 * m < hl) ? load_sample(up, start + i, ns, true) : 0.0f;
        a1[m] = (64 * m + 1 < hl) ? load_sample(up, start + i + 1, ns, true) : 0.0f;
      }
    }
#pragma unroll
    for (int m = 0; m < E; m++) {
      if (DITHER) {  // feature-window.cc:139-140 (a separate instantiation: the Box-Muller code is large)
        float2 g = gauss_pair(p.seed, (uint32_t)t, (uint32_t)(lane + 32 * m));
        if (64 * m < hl) a0[m] += g.x * p.dither;
        if (64 * m + 1 < hl) a1[m] += g.y * p.dither;
      }
      sum += a0[m] + a1[m];
    }
    if (p.remove_dc) {  // feature-window.cc:142-143
      // -Sum()/frame_length, one rounded division (feature-window.cc:144); __fdiv_rn also keeps the compiler from
      // contracting it into the additions below, which it did for one sample type and not the other.
      // Added to the zero padding as well: everything past the frame is multiplied by the window's zeros below, and
      // the raw energy masks it.
      const float neg_mean = -__fdiv_rn(warp_sum(sum), static_cast<float>(p.L));
#pragma unroll
      for (int m = 0; m < E; m++) {
        a0[m] += neg_mean;
        a1[m] += neg_mean;
      }
    }
    float log_energy = 0.0f;
    if (p.use_energy && p.raw_energy) {  // feature-window.cc:145-149
      float e = 0.0f;
#pragma unroll
      for (int m = 0; m < E; m++) {
        const float x0 = (64 * m < hl) ? a0[m] : 0.0f, x1 = (64 * m + 1 < hl) ? a1[m] : 0.0f;
        e += x0 * x0 + x1 * x1;
      }
      log_energy = logf(fmaxf(warp_sum(e), FLT_EPSILON));
    }
    // ---- pre-emphasis (feature-window.cc:101-107) + window; the previous sample of (2j) lives in lane-1 ----------
    float2 v[E];
    float e_win = 0.0f;
#pragma unroll
    for (int m = 0; m < E; m++) {
      float prev = __shfl_up_sync(0xffffffffu, a1[m], 1);
      const float wrap = __shfl_sync(0xffffffffu, a1[m > 0 ? m - 1 : 0], 31);
      if (lane == 0) prev = (m > 0) ? wrap : a0[0];
      const int i = 2 * (lane + 32 * m);
      float y0 = a0[m] - p.preemph * prev;
      float y1 = a1[m] - p.preemph * a0[m];
      const float2 w = *reinterpret_cast<const float2 *>(s_win + i);
      y0 *= w.x;
      y1 *= w.y;
      v[m] = make_float2(y0, y1);
      e_win += y0 * y0 + y1 * y1;
    }
    if (p.use_energy && !p.raw_energy)  // feature-mfcc.cc:37-39
      log_energy = logf(fmaxf(warp_sum(e_win), FLT_MIN));

    // ---- (N/2)-point complex FFT: radix-E inside the lane, twiddle, then 32-point DIF across lanes --------------
    local_fft<E>(v);
#pragma unroll
    for (int m = 1; m < E; m++) {
      v[m] = cmul(v[m], s_wl[m * 32 + lane]);
    }
    // Upper lanes need (o - v) * w, lower lanes v + o: one form, o + sg * v times (upper ? w : 1), with the same
    // roundings as the two separate expressions and no selects.  The last stage's twiddle is W_2^0 = 1.
#pragma unroll
    for (int s = 0; s < 5; s++) {
      const int half = 16 >> s;
      const float sg = sgn[s];
#pragma unroll
      for (int m = 0; m < E; m++) {
        const float2 o = shfl_xor2(v[m], half);
        const float2 t = make_float2(fmaf(sg, v[m].x, o.x), fmaf(sg, v[m].y, o.y));
        v[m] = (s < 4) ? cmul(t, wst[s]) : t;
      }
    }
    // now v[m] = Z[k1 + E*k2], k1 = bitrev<E>(m)

    // ---- real post-pass + power spectrum (srfft.cc:362-431, feature-functions.cc:29-51) -------------------------
    // X[k] = (Z[k] + conj Z[n-k])/2 + W_N^k (Z[k] - conj Z[n-k])/(2i);  P[0] = (Re Z0 + Im Z0)^2;  P[n] is never read.
    const int lane_k0 = __brev((unsigned)((32 - k2) & 31)) >> 27;  // lane holding Z[E*(32-k2)] in register 0
    const bool root = p.fbank && !p.use_power;  // power_spectrum.ApplyPow(0.5), feature-fbank.cc:97-98
    // The lane ends up with the E consecutive bins E*k2 .. E*k2 + E - 1 (pw[k1]) and stores them as float4.
    float pw[E];
    // k1 = 0 and k1 = E/2 pair with the same register of another lane: one bin per lane
#pragma unroll
    for (int m = 0; m < (E > 1 ? 2 : 1); m++) {
      const int k1 = bitrev<E>(m);
      const float2 z = v[m];
      const float2 zp = (k1 == 0) ? shfl2(v[0], lane_k0) : shfl_xor2(v[m], 31);
      float pk;
      if (k1 == 0 && k2 == 0) {
        pk = (z.x + z.y) * (z.x + z.y);
      } else {
        const float er = 0.5f * (z.x + zp.x), ei = 0.5f * (z.y - zp.y);
        const float orr = 0.5f * (z.y + zp.y), oi = -0.5f * (z.x - zp.x);
        const float2 w = s_wp[m * 32 + lane];
        const float xr = er + (orr * w.x - oi * w.y), xi = ei + (orr * w.y + oi * w.x);
        pk = xr * xr + xi * xi;
      }
      pw[k1] = pk;
    }
    // 0 < k1 < E/2: bins k and n - k share E = (Z[k] + conj Z[n-k])/2 and T = W^k (Z[k] - conj Z[n-k])/(2i);
    // X[k] = E + T, X[n-k] = conj(E - T).  This lane owns Z[k], the lane^31 holds Z[n-k] in register bitrev(E - k1), and
    // bin n - k = (E - k1) + E (31 - k2) belongs to that lane's block: it gets it back with one more shuffle, while the
    // mirrored pair is the other lane's job, so every bin is computed exactly once.
#pragma unroll
    for (int k1 = 1; k1 < E / 2; k1++) {
      const int m = bitrev<E>(k1), mp = bitrev<E>(E - k1);
      const float2 z = v[m];
      const float2 zp = shfl_xor2(v[mp], 31);
      const float er = 0.5f * (z.x + zp.x), ei = 0.5f * (z.y - zp.y);
      const float orr = 0.5f * (z.y + zp.y), oi = -0.5f * (z.x - zp.x);
      const float2 w = s_wp[m * 32 + lane];
      const float tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
      const float xr = er + tr, xi = ei + ti, yr = er - tr, yi = ei - ti;
      pw[k1] = xr * xr + xi * xi;
      pw[E - k1] = __shfl_xor_sync(0xffffffffu, yr * yr + yi * yi, 31);
    }
    if (root) {
#pragma unroll
      for (int k1 = 0; k1 < E; k1++) pw[k1] = sqrtf(pw[k1]);
    }
    if constexpr (E >= 4) {
#pragma unroll
      for (int j = 0; j < E / 4; j++)
        *reinterpret_cast<float4 *>(ps + E * k2 + 4 * j) = make_float4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
    } else {
      *reinterpret_cast<float2 *>(ps + E * k2) = make_float2(pw[0], pw[1]);
    }
    __syncwarp();

    // ---- mel filterbank (lane = bin), floor, log (mel-computations.cc:228-253, feature-mfcc.cc:51-55) -----------
    // Filters start on a multiple of four bins (zero weights in front) and are a whole number of float4 long; the sum
    // runs over the same products in the same order as the reference's dot product, with exact zeros around them.
    float logmel = 0.0f;
    if (lane < p.B) {
      const int off = s_moff[mt * p.B + lane], len4 = s_mlen[mt * p.B + lane];
      const float4 *q = reinterpret_cast<const float4 *>(ps + off);
      float e = 0.0f;
      if (mt == 0) {  // table 0 is staged in shared memory
        const float4 *w = reinterpret_cast<const float4 *>(s_melw + lane * p.mel_pitch);
        for (int i = 0; i < len4; i++) {
          const float4 a = w[i], b = q[i];
          e += a.x * b.x;
          e += a.y * b.y;
          e += a.z * b.z;
          e += a.w * b.w;
        }
      } else {
        const float4 *w = reinterpret_cast<const float4 *>(p.mel_w + ((size_t)mt * p.B + lane) * p.mel_pitch);
        for (int i = 0; i < len4; i++) {
          const float4 a = __ldg(w + i), b = q[i];
          e += a.x * b.x;
          e += a.y * b.y;
          e += a.z * b.z;
          e += a.w * b.w;
        }
      }
      if (p.htk_mode && e < 1.0f) e = 1.0f;
      logmel = (PLP || (p.fbank && !p.use_log_fbank)) ? e : logf(fmaxf(e, FLT_EPSILON));
    }
    __syncwarp();  // ps is rewritten by the next frame

    if constexpr (PLP) {  // PlpComputer::Compute, feature-plp.cc:143-188 (lane = mel bin, then lane = autocorrelation / LPC index)
      const int n = p.lpc_order, B2 = p.B + 2;
      float m = 0.0f;
      if (lane < p.B) m = powf(logmel * __ldg(p.eq_loud + mt * p.B + lane), p.compress_factor);  // MulElements, ApplyPow
      // autocorrelation: idft_bases . [m_0, m_0 .. m_{B-1}, m_{B-1}]  (first and last element duplicated)
      float ac = 0.0f;
      for (int j = 0; j < B2; j++) {
        const float v = __shfl_sync(0xffffffffu, m, j == 0 ? 0 : (j == B2 - 1 ? p.B - 1 : j - 1));
        if (lane <= n) ac = fmaf(__ldg(p.idft + lane * B2 + j), v, ac);
      }
      // Durbin (mel-computations.cc:269-300): lane j holds pLP[j]; the j-loops run across lanes
      float lp = 0.0f, E = __shfl_sync(0xffffffffu, ac, 0);
      for (int i = 0; i < n; i++) {
        const float acv = __shfl_sync(0xffffffffu, ac, (i - lane) & 31);     // pAC[i - j]
        const float sum = warp_sum(lane < i ? lp * acv : 0.0f);
        const float ki = (__shfl_sync(0xffffffffu, ac, i + 1) + sum) / E;
        float c = 1.0f - ki * ki;
        if (c < 1.0e-5f) c = 1.0e-5f;
        E *= c;
        const float lpr = __shfl_sync(0xffffffffu, lp, (i - lane - 1) & 31);  // pLP[i - j - 1]
        lp = lane < i ? lp - ki * lpr : (lane == i ? -ki : lp);
      }
      float res = (float)(-log(1.0 / (double)E));  // ComputeLpc: -Log(1.0 / ans)
      res = fmaxf(res, FLT_MIN);
      // Lpc2Cepstrum (mel-computations.cc:302-311): float products, double sum
      float cep = 0.0f;
      for (int i = 0; i < n; i++) {
        const float cv = __shfl_sync(0xffffffffu, cep, (i - lane - 1) & 31);  // pCepst[i - j - 1]
        double term = lane < i ? (double)(((float)(i - lane) * lp) * cv) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
        const float ci = (float)(-(double)__shfl_sync(0xffffffffu, lp, i) - term / (double)(float)(i + 1));
        if (lane == i) cep = ci;
      }
      float c = __shfl_up_sync(0xffffffffu, cep, 1);  // feature[k] = raw_cepstrum[k - 1], k >= 1
      if (lane == 0) c = res;
      if (lane < p.C && p.use_lifter) c *= s_lift[lane];
      if (p.cepstral_scale != 1.0f) c *= p.cepstral_scale;
      if (p.use_energy) {
        if (p.energy_floor > 0.0f && log_energy < p.log_energy_floor) log_energy = p.log_energy_floor;
        if (lane == 0) c = log_energy;
      }
      float *orow = p.out + t * p.out_stride;
      if (!p.htk_compat) {
        if (lane < p.C) orow[lane] = c;
      } else {
        if (lane == 0) orow[p.C - 1] = c;
        else if (lane < p.C) orow[lane - 1] = c;
      }
      if (lane >= p.C && lane < p.out_stride) orow[lane] = 0.0f;
      continue;
    }

    if (p.fbank) {  // FbankComputer::Compute, feature-fbank.cc:100-121: the mel energies are the feature; energy first or last
      float *orow = p.out + t * p.out_stride;
      const int mel_offset = (p.use_energy && !p.htk_compat) ? 1 : 0, dim = p.B + (p.use_energy ? 1 : 0);
      if (lane < p.B) orow[mel_offset + lane] = logmel;
      if (p.use_energy && lane == 0) {
        if (p.energy_floor > 0.0f && log_energy < p.log_energy_floor) log_energy = p.log_energy_floor;
        orow[p.htk_compat ? p.B : 0] = log_energy;
      }
      for (int c = dim + lane; c < p.out_stride; c += 32) orow[c] = 0.0f;  // keep the stride padding defined
      continue;
    }

    // ---- DCT, lifter, C0/energy, HTK order (feature-mfcc.cc:57-79) ----------------------------------------------
    // lane = cepstral index; the log-mel vector goes through shared memory so that both operands arrive as float4
    float *lm = s_lm + warp * ds;
    if (lane < p.B) lm[lane] = logmel;
    __syncwarp();
    float c = 0.0f;
    {
      const float4 *d4 = reinterpret_cast<const float4 *>(s_dct + lane * ds), *l4 = reinterpret_cast<const float4 *>(lm);
      for (int b = 0; b < ds / 4; b++) {
        const float4 a = d4[b], x = l4[b];
        c += a.x * x.x;
        c += a.y * x.y;
        c += a.z * x.z;
        c += a.w * x.w;
      }
    }
    __syncwarp();  // lm is rewritten by the next frame
    if (lane < p.C && p.use_lifter) c *= s_lift[lane];
    if (p.use_energy) {
      if (p.energy_floor > 0.0f && log_energy < p.log_energy_floor) log_energy = p.log_energy_floor;
      if (lane == 0) c = log_energy;
    }
    float *orow = p.out + t * p.out_stride;
    if (!p.htk_compat) {
      if (lane < p.C) orow[lane] = c;
    } else {
      if (lane == 0) orow[p.C - 1] = p.use_energy ? c : c * 1.41421356237309504880f;
      else if (lane < p.C) orow[lane - 1] = c;
    }
    if (lane >= p.C && lane < p.out_stride) orow[lane] = 0.0f;  // keep the stride padding defined
  }
}

template <int E, typename SampleT>
int launch_e(const MfccParams &p, int n_mel, int device, cudaStream_t s) {
  const size_t smem = (size_t)mfcc_smem(E, p.C, p.B, n_mel, p.mel_pitch).total;
  auto kern = p.plp ? (p.dither != 0.0f ? mfcc_kernel<E, SampleT, true, true> : mfcc_kernel<E, SampleT, false, true>)
                    : (p.dither != 0.0f ? mfcc_kernel<E, SampleT, true, false> : mfcc_kernel<E, SampleT, false, false>);
  if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks_needed = (p.total_frames + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t cap = (int64_t)vb::num_sms(device) * 8;  // persistent-style grid: 8 resident CTAs per SM
  int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  if (grid < 1) return 0;
  kern<<<grid, kWarpsPerBlock * 32, smem, s>>>(p);
  VB_CUDA(cudaGetLastError());
  return 0;
}

template <typename SampleT>
int launch_t(const MfccParams &p, int npad, int n_mel, int device, cudaStream_t s) {
  switch (npad) {
    case 128: return launch_e<2, SampleT>(p, n_mel, device, s);
    case 256: return launch_e<4, SampleT>(p, n_mel, device, s);
    case 512: return launch_e<8, SampleT>(p, n_mel, device, s);
    case 1024: return launch_e<16, SampleT>(p, n_mel, device, s);
    case 2048: return launch_e<32, SampleT>(p, n_mel, device, s);
    default: return vb::fail(VBGPU_ERR_INVALID, "padded window size %d not in {128,...,2048}", npad);
  }
}

}  // namespace

namespace vb {

// Build (or extend) the device tables.  warps: distinct VTLN factors, warps[0] == 1.0.
int mfcc_build_tables(vbgpu_mfcc_t h) {
  const vbgpu_mfcc_opts &o = h->opts;
  const int L = h->L, npad = h->npad, n = npad / 2, B = o.num_bins, C = o.num_ceps;
  std::vector<float> win(npad, 0.0f);
  const double a = 6.283185307179586476925286766559005 / (L - 1);  // feature-window.cc:109-131
  for (int i = 0; i < L; i++) {
    double x = (double)i, v;
    switch (o.window_type) {
      case 0: v = pow(0.5 - 0.5 * cos(a * x), 0.85); break;
      case 1: v = 0.54 - 0.46 * cos(a * x); break;
      case 2: v = 0.5 - 0.5 * cos(a * x); break;
      case 3: v = 1.0; break;
      case 4: v = o.blackman_coeff - 0.5 * cos(a * x) + (0.5 - o.blackman_coeff) * cos(2 * a * x); break;
      default: return fail(VBGPU_ERR_INVALID, "invalid window type %d", o.window_type);
    }
    win[i] = (float)v;
  }
  std::vector<float2> tw(n);
  for (int k = 0; k < n; k++) {
    double ang = -6.283185307179586476925286766559005 * k / npad;
    tw[k] = make_float2((float)cos(ang), (float)sin(ang));
  }
  std::vector<float> dct((size_t)C * B), lift(C, 1.0f);
  {  // matrix-functions.cc:592-608
    float nz = (float)sqrt(1.0 / (float)B);
    for (int j = 0; j < B; j++) dct[j] = nz;
    nz = (float)sqrt(2.0 / (float)B);
    for (int k = 1; k < C; k++)
      for (int j = 0; j < B; j++) dct[(size_t)k * B + j] = (float)(nz * cos(M_PI / B * (j + 0.5) * k));
  }
  if (o.cepstral_lifter != 0.0f)  // mel-computations.cc:255-261
    for (int i = 0; i < C; i++) lift[i] = (float)(1.0 + 0.5 * o.cepstral_lifter * sin(M_PI * i / o.cepstral_lifter));

  // mel tables as the kernel reads them: a filter starts at the multiple of four bins below its first bin (`lead` zero
  // weights in front) and is len4 float4 long; pitch = 4 * odd, so the float4 reads of a warp's 32 rows do not collide
  int pitch = 4;
  std::vector<std::vector<int32_t>> offs(h->warps.size()), lens(h->warps.size());
  std::vector<std::vector<float>> ws(h->warps.size());
  for (size_t i = 0; i < h->warps.size(); i++) {
    VB_TRY(build_mel(o, npad, h->warps[i], &offs[i], &lens[i], &ws[i], n));
    for (int b = 0; b < B; b++) {
      const int span = (offs[i][b] % 4 + lens[i][b] + 3) / 4 * 4;
      pitch = span > pitch ? span : pitch;
    }
  }
  if ((pitch / 4) % 2 == 0) pitch += 4;
  h->mel_pitch = pitch;
  std::vector<int32_t> off_all, len_all;
  std::vector<float> w_all((size_t)h->warps.size() * B * pitch, 0.0f);
  for (size_t i = 0; i < h->warps.size(); i++)
    for (int b = 0; b < B; b++) {
      const int lead = offs[i][b] % 4;
      off_all.push_back(offs[i][b] - lead);
      len_all.push_back((lead + lens[i][b] + 3) / 4);
      for (int k = 0; k < lens[i][b]; k++) w_all[((size_t)i * B + b) * pitch + lead + k] = ws[i][(size_t)b * n + k];
    }
  cudaStream_t s = h->stream;
  if (h->plp) {  // PlpComputer ctor (feature-plp.cc:25-50): InitIdftBases (feature-functions.cc:188-203) and, per mel table,
                 // GetEqualLoudnessVector (mel-computations.cc:313-326) at the filters' centre frequencies (:89-104)
    const int nb = h->lpc_order + 1, dim = B + 2;
    std::vector<float> idft((size_t)nb * dim), eq((size_t)h->warps.size() * B);
    const float angle = (float)(M_PI / (float)(dim - 1)), scale = (float)(1.0f / (2.0 * (float)(dim - 1)));
    for (int i = 0; i < nb; i++) {
      idft[(size_t)i * dim] = (float)(1.0 * scale);
      for (int j = 1; j < dim - 1; j++) idft[(size_t)i * dim + j] = (float)(2.0 * scale * cos(angle * (float)i * (float)j));
      idft[(size_t)i * dim + dim - 1] = (float)(scale * cos(angle * (float)i * (float)(dim - 1)));
    }
    const float nyq = 0.5f * o.samp_freq, low = o.low_freq, high = o.high_freq > 0.0f ? o.high_freq : nyq + o.high_freq;
    const float mlow = mel_scale(low), delta = (mel_scale(high) - mlow) / (B + 1);
    float vlow = o.vtln_low, vhigh = o.vtln_high;
    if (vhigh < 0.0f) vhigh += nyq;
    for (size_t i = 0; i < h->warps.size(); i++)
      for (int b = 0; b < B; b++) {
        float cm = mlow + (b + 1) * delta;
        if (h->warps[i] != 1.0f) cm = vtln_warp_mel(vlow, vhigh, low, high, h->warps[i], cm);
        const float f0 = inv_mel_scale(cm), fsq = f0 * f0, fsub = (float)(fsq / (fsq + 1.6e5));
        eq[i * B + b] = (float)(fsub * fsub * ((fsq + 1.44e6) / (fsq + 9.61e6)));
      }
    VB_TRY(h->d_idft.reserve(idft.size() * 4));
    VB_TRY(h->d_eq_loud.reserve(eq.size() * 4));
    VB_CUDA(cudaMemcpy(h->d_idft.p, idft.data(), idft.size() * 4, cudaMemcpyHostToDevice));
    VB_CUDA(cudaMemcpy(h->d_eq_loud.p, eq.data(), eq.size() * 4, cudaMemcpyHostToDevice));
  }
  VB_TRY(h->d_window.reserve(win.size() * 4));
  VB_TRY(h->d_tw.reserve(tw.size() * 8));
  VB_TRY(h->d_dct.reserve(dct.size() * 4));
  VB_TRY(h->d_lifter.reserve(lift.size() * 4));
  VB_TRY(h->d_mel_off.reserve(off_all.size() * 4));
  VB_TRY(h->d_mel_len.reserve(len_all.size() * 4));
  VB_TRY(h->d_mel_w.reserve(w_all.size() * 4));
  VB_CUDA(cudaMemcpyAsync(h->d_window.p, win.data(), win.size() * 4, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaMemcpyAsync(h->d_tw.p, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaMemcpyAsync(h->d_dct.p, dct.data(), dct.size() * 4, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaMemcpyAsync(h->d_lifter.p, lift.data(), lift.size() * 4, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaMemcpyAsync(h->d_mel_off.p, off_all.data(), off_all.size() * 4, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaMemcpyAsync(h->d_mel_len.p, len_all.data(), len_all.size() * 4, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaMemcpyAsync(h->d_mel_w.p, w_all.data(), w_all.size() * 4, cudaMemcpyHostToDevice, s));
  VB_CUDA(cudaStreamSynchronize(s));  // the host vectors die here
  return 0;
}

int mfcc_launch(vbgpu_mfcc_t h, const void *d_pcm, bool is_f32, float *d_out, int32_t out_stride, cudaStream_t s) {
  const vbgpu_mfcc_opts &o = h->opts;
  MfccParams p;
  p.pcm = d_pcm;
  p.sample_offsets = h->layout.d_sample_offsets.as<int64_t>();
  p.frame_offsets = h->layout.d_frame_offsets.as<int64_t>();
  p.frame2utt = h->layout.d_frame2utt.as<int32_t>();
  p.utt_mel = h->warps.size() > 1 ? h->layout.d_utt_aux.as<int32_t>() : nullptr;
  p.total_frames = h->layout.total_frames;
  p.L = h->L;
  p.shift = h->shift;
  p.snip_edges = o.snip_edges;
  p.remove_dc = o.remove_dc_offset;
  p.use_energy = o.use_energy;
  p.raw_energy = o.raw_energy;
  p.htk_compat = o.htk_compat;
  p.htk_mode = o.htk_mode;
  p.B = o.num_bins;
  p.C = o.num_ceps;
  p.use_lifter = o.cepstral_lifter != 0.0f;
  p.mel_pitch = h->mel_pitch;
  p.n_mel = (int)h->warps.size();
  p.plp = h->plp;
  p.lpc_order = h->lpc_order;
  p.compress_factor = h->compress_factor;
  p.cepstral_scale = h->cepstral_scale;
  p.idft = h->d_idft.as<float>();
  p.eq_loud = h->d_eq_loud.as<float>();
  p.fbank = h->fbank;
  p.use_log_fbank = h->use_log_fbank;
  p.use_power = h->use_power;
  p.preemph = o.preemph_coeff;
  p.energy_floor = o.energy_floor;
  p.log_energy_floor = h->log_energy_floor;
  p.dither = o.dither;
  p.seed = h->dither_seed++;
  p.window = h->d_window.as<float>();
  p.tw = h->d_tw.as<float2>();
  p.mel_off = h->d_mel_off.as<int32_t>();
  p.mel_len = h->d_mel_len.as<int32_t>();
  p.mel_w = h->d_mel_w.as<float>();
  p.dct = h->d_dct.as<float>();
  p.lifter = h->d_lifter.as<float>();
  p.out = d_out;
  p.out_stride = out_stride;
  return is_f32 ? launch_t<float>(p, h->npad, p.n_mel, h->device, s) : launch_t<int16_t>(p, h->npad, p.n_mel, h->device, s);
}

}  // namespace vb
