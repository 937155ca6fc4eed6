/* oracle/oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's CPU algorithm for PCM -> MFCC -> CMVN/deltas/splice/
 * transform -> diag-GMM log-likelihoods / EM statistics.  Each function cites the reference code it
 * follows (paths relative to /root/reference/kaldi-master/src).  Arithmetic types follow the
 * reference (float where it uses BaseFloat, double where it uses double).  BLAS calls of the
 * reference (sdot/sgemv/sgemm/dger) are restated as plain loops, so results agree with the compiled
 * reference to summation-order noise (~1e-6 relative), not bit-for-bit.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif
#define ORC_2PI 6.283185307179586476925286766559005   /* base/kaldi-math.h:52 */
#define ORC_LOG_2PI 1.8378770664093454835606594728112 /* base/kaldi-math.h:60 */

void orc_mfcc_opts_default(orc_mfcc_opts *o) {
  /* feature-window.h:51-62, mel-computations.h:55-57, feature-mfcc.h:49-58 */
  o->samp_freq = 16000.0f;
  o->frame_shift_ms = 10.0f;
  o->frame_length_ms = 25.0f;
  o->dither = 1.0f;
  o->preemph_coeff = 0.97f;
  o->remove_dc_offset = 1;
  o->window_type = 0;
  o->round_to_power_of_two = 1;
  o->blackman_coeff = 0.42f;
  o->snip_edges = 1;
  o->num_bins = 23;
  o->low_freq = 20.0f;
  o->high_freq = 0.0f;
  o->vtln_low = 100.0f;
  o->vtln_high = -500.0f;
  o->htk_mode = 0;
  o->num_ceps = 13;
  o->use_energy = 1;
  o->energy_floor = 0.0f;
  o->raw_energy = 1;
  o->cepstral_lifter = 22.0f;
  o->htk_compat = 0;
}

/* feature-window.h:92-101 (the products are evaluated in double, as `samp_freq * 0.001 * ms`) */
int32_t orc_window_shift(const orc_mfcc_opts *o) { return (int32_t)(o->samp_freq * 0.001 * o->frame_shift_ms); }
int32_t orc_window_size(const orc_mfcc_opts *o) { return (int32_t)(o->samp_freq * 0.001 * o->frame_length_ms); }
int32_t orc_padded_window_size(const orc_mfcc_opts *o) {
  int32_t n = orc_window_size(o);
  if (!o->round_to_power_of_two) return n;
  /* base/kaldi-math.cc:31-40 RoundUpToNearestPowerOfTwo */
  int32_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

/* feature-window.cc:28-39 */
int64_t orc_first_sample_of_frame(int32_t frame, const orc_mfcc_opts *o) {
  int64_t shift = orc_window_shift(o);
  if (o->snip_edges) return (int64_t)frame * shift;
  int64_t mid = shift * frame + shift / 2;
  return mid - orc_window_size(o) / 2;
}

/* feature-window.cc:41-87 with flush == true (the offline computer's call, feature-common-inl.h:78) */
int32_t orc_num_frames(int64_t num_samples, const orc_mfcc_opts *o) {
  int64_t shift = orc_window_shift(o), len = orc_window_size(o);
  if (o->snip_edges) {
    if (num_samples < len) return 0;
    return (int32_t)(1 + (num_samples - len) / shift);
  }
  return (int32_t)((num_samples + shift / 2) / shift);
}

/* feature-window.cc:109-131 */
int orc_window_table(const orc_mfcc_opts *o, float *window) {
  int32_t L = orc_window_size(o);
  if (L <= 0) return -1;
  double a = ORC_2PI / (L - 1);
  for (int32_t i = 0; i < L; i++) {
    double x = (double)i;
    switch (o->window_type) {
      case 2: window[i] = (float)(0.5 - 0.5 * cos(a * x)); break;
      case 1: window[i] = (float)(0.54 - 0.46 * cos(a * x)); break;
      case 0: window[i] = (float)pow(0.5 - 0.5 * cos(a * x), 0.85); break;
      case 3: window[i] = 1.0f; break;
      case 4:
        window[i] = (float)(o->blackman_coeff - 0.5 * cos(a * x) + (0.5 - o->blackman_coeff) * cos(2 * a * x));
        break;
      default: return -1;
    }
  }
  return 0;
}

/* mel-computations.h:81-87 */
static float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }
static float inv_mel_scale(float m) { return 700.0f * (expf(m / 1127.0f) - 1.0f); }

/* mel-computations.cc:152-213 */
static float vtln_warp_freq(float vlow, float vhigh, float low, float high, float warp, float freq) {
  if (freq < low || freq > high) return freq;
  float one = 1.0f;
  float l = vlow * (one > warp ? one : warp);
  float h = vhigh * (one < warp ? one : warp);
  float scale = 1.0f / warp;
  float Fl = scale * l, Fh = scale * h;
  float scale_left = (Fl - low) / (l - low);
  float scale_right = (high - Fh) / (high - h);
  if (freq < l) return low + scale_left * (freq - low);
  if (freq < h) return scale * freq;
  return high + scale_right * (freq - high);
}
/* mel-computations.cc:215-224 */
static float vtln_warp_mel(float vlow, float vhigh, float low, float high, float warp, float mel) {
  return mel_scale(vtln_warp_freq(vlow, vhigh, low, high, warp, inv_mel_scale(mel)));
}

/* mel-computations.cc:33-144 */
int orc_mel_banks(const orc_mfcc_opts *o, float vtln_warp, int32_t *offsets, int32_t *lens, float *weights) {
  int32_t B = o->num_bins;
  if (B < 3) return -1;
  float fs = o->samp_freq;
  int32_t Npad = orc_padded_window_size(o);
  if (Npad % 2 != 0) return -1;
  int32_t nfft = Npad / 2;
  float nyq = 0.5f * fs;
  float low = o->low_freq, high = (o->high_freq > 0.0f) ? o->high_freq : nyq + o->high_freq;
  if (low < 0.0f || low >= nyq || high <= 0.0f || high > nyq || high <= low) return -1;
  float bin_width = fs / Npad;
  float mel_low = mel_scale(low), mel_high = mel_scale(high);
  float delta = (mel_high - mel_low) / (B + 1);
  float vlow = o->vtln_low, vhigh = o->vtln_high;
  if (vhigh < 0.0f) vhigh += nyq;
  if (vtln_warp != 1.0f &&
      (vlow < 0.0f || vlow <= low || vlow >= high || vhigh <= 0.0f || vhigh >= high || vhigh <= vlow))
    return -1;
  memset(weights, 0, sizeof(float) * (size_t)B * (size_t)nfft);
  float *tmp = (float *)malloc(sizeof(float) * (size_t)nfft);
  for (int32_t b = 0; b < B; b++) {
    float lm = mel_low + b * delta, cm = mel_low + (b + 1) * delta, rm = mel_low + (b + 2) * delta;
    if (vtln_warp != 1.0f) {
      lm = vtln_warp_mel(vlow, vhigh, low, high, vtln_warp, lm);
      cm = vtln_warp_mel(vlow, vhigh, low, high, vtln_warp, cm);
      rm = vtln_warp_mel(vlow, vhigh, low, high, vtln_warp, rm);
    }
    int32_t first = -1, last = -1;
    for (int32_t i = 0; i < nfft; i++) {
      tmp[i] = 0.0f;
      float freq = bin_width * i;
      float mel = mel_scale(freq);
      if (mel > lm && mel < rm) {
        float w = (mel <= cm) ? (mel - lm) / (cm - lm) : (rm - mel) / (rm - cm);
        tmp[i] = w;
        if (first == -1) first = i;
        last = i;
      }
    }
    if (first == -1) { free(tmp); return -2; } /* "You may have set --num-mel-bins too large" */
    offsets[b] = first;
    lens[b] = last + 1 - first;
    for (int32_t i = 0; i < lens[b]; i++) weights[(size_t)b * nfft + i] = tmp[first + i];
    if (o->htk_mode && b == 0 && mel_low != 0.0f) weights[0] = 0.0f; /* :132-134 */
  }
  free(tmp);
  return 0;
}

/* MelBanks::center_freqs_ (mel-computations.cc:89-104) and GetEqualLoudnessVector (:313-326) */
static int equal_loudness(const orc_mfcc_opts *o, float vtln_warp, float *eq) {
  int32_t B = o->num_bins;
  float nyq = 0.5f * o->samp_freq;
  float low = o->low_freq, high = (o->high_freq > 0.0f) ? o->high_freq : nyq + o->high_freq;
  float mel_low = mel_scale(low), mel_high = mel_scale(high);
  float delta = (mel_high - mel_low) / (B + 1);
  float vlow = o->vtln_low, vhigh = o->vtln_high;
  if (vhigh < 0.0f) vhigh += nyq;
  for (int32_t b = 0; b < B; b++) {
    float cm = mel_low + (b + 1) * delta;
    if (vtln_warp != 1.0f) cm = vtln_warp_mel(vlow, vhigh, low, high, vtln_warp, cm);
    float f0 = inv_mel_scale(cm);
    float fsq = f0 * f0;
    float fsub = (float)(fsq / (fsq + 1.6e5));
    eq[b] = (float)(fsub * fsub * ((fsq + 1.44e6) / (fsq + 9.61e6)));
  }
  return 0;
}

/* InitIdftBases, feature-functions.cc:188-203 */
static void idft_bases(int32_t n_bases, int32_t dim, float *m) {
  float angle = (float)(M_PI / (float)(dim - 1));
  float scale = (float)(1.0f / (2.0 * (float)(dim - 1)));
  for (int32_t i = 0; i < n_bases; i++) {
    m[i * dim] = (float)(1.0 * scale);
    float i_fl = (float)i;
    for (int32_t j = 1; j < dim - 1; j++) m[i * dim + j] = (float)(2.0 * scale * cos(angle * i_fl * (float)j));
    m[i * dim + dim - 1] = (float)(scale * cos(angle * i_fl * (float)(dim - 1)));
  }
}

/* Durbin, mel-computations.cc:269-300 */
static float durbin(int n, const float *ac, float *lp, float *tmp) {
  float E = ac[0];
  for (int i = 0; i < n; i++) {
    float ki = ac[i + 1];
    for (int j = 0; j < i; j++) ki += lp[j] * ac[i - j];
    ki = ki / E;
    float c = 1 - ki * ki;
    if (c < 1.0e-5) c = 1.0e-5;
    E *= c;
    tmp[i] = -ki;
    for (int j = 0; j < i; j++) tmp[j] = lp[j] - ki * lp[i - j - 1];
    for (int j = 0; j <= i; j++) lp[j] = tmp[j];
  }
  return E;
}

/* Lpc2Cepstrum, mel-computations.cc:302-311 */
static void lpc2cepstrum(int n, const float *lpc, float *cep) {
  for (int i = 0; i < n; i++) {
    double sum = 0.0;
    for (int j = 0; j < i; j++) sum += (float)(i - j) * lpc[j] * cep[i - j - 1];
    cep[i] = (float)(-lpc[i] - sum / (float)(i + 1));
  }
}

/* matrix/matrix-functions.cc:592-608 (float normalizer, double cosine) */
void orc_dct_matrix(int32_t K, int32_t N, float *M) {
  float normalizer = (float)sqrt(1.0 / (float)N);
  for (int32_t j = 0; j < N; j++) M[j] = normalizer;
  normalizer = (float)sqrt(2.0 / (float)N);
  for (int32_t k = 1; k < K; k++)
    for (int32_t n = 0; n < N; n++) M[k * N + n] = (float)(normalizer * cos((double)M_PI / N * (n + 0.5) * k));
}

/* mel-computations.cc:255-261 */
void orc_lifter_coeffs(float Q, int32_t n, float *c) {
  for (int32_t i = 0; i < n; i++) c[i] = (float)(1.0 + 0.5 * Q * sin(M_PI * i / Q));
}

/* Forward real FFT of N floats (N a power of two), output packed as srfft.cc:362-431 does:
 * [Re0, Re(N/2), Re1, Im1, ...], sign exp(-2*pi*i*k*n/N), unscaled.  The reference uses a float
 * split-radix kernel; any exact FFT agrees with it to float rounding, so this is
 * an iterative radix-2 complex FFT of length N/2 on (even, odd) pairs + the standard real post-pass,
 * in float with double-evaluated twiddles. */
static void real_fft_forward(float *x, int32_t N, float *scratch /* N floats */) {
  int32_t n = N / 2, logn = 0;
  while ((1 << logn) < n) logn++;
  float *re = scratch, *im = scratch + n;
  for (int32_t i = 0; i < n; i++) {
    int32_t r = 0;
    for (int32_t b = 0; b < logn; b++)
      if (i & (1 << b)) r |= 1 << (logn - 1 - b);
    re[r] = x[2 * i];
    im[r] = x[2 * i + 1];
  }
  for (int32_t len = 2; len <= n; len <<= 1) {
    int32_t half = len / 2;
    for (int32_t k = 0; k < half; k++) {
      double ang = -ORC_2PI * k / len;
      float wr = (float)cos(ang), wi = (float)sin(ang);
      for (int32_t s = k; s < n; s += len) {
        int32_t t = s + half;
        float tr = re[t] * wr - im[t] * wi, ti = re[t] * wi + im[t] * wr;
        re[t] = re[s] - tr;
        im[t] = im[s] - ti;
        re[s] += tr;
        im[s] += ti;
      }
    }
  }
  /* X[k] = E[k] + W^k O[k],  E = (Z[k] + conj Z[n-k])/2,  O = (Z[k] - conj Z[n-k])/(2i) */
  x[0] = re[0] + im[0];
  x[1] = re[0] - im[0];
  for (int32_t k = 1; k < n; k++) {
    int32_t j = n - k;
    float er = 0.5f * (re[k] + re[j]), ei = 0.5f * (im[k] - im[j]);
    float orr = 0.5f * (im[k] + im[j]), oi = -0.5f * (re[k] - re[j]);
    double ang = -ORC_2PI * k / N;
    float wr = (float)cos(ang), wi = (float)sin(ang);
    x[2 * k] = er + (orr * wr - oi * wi);
    x[2 * k + 1] = ei + (orr * wi + oi * wr);
  }
}

/* fbank = 0: MfccComputer::Compute (feature-mfcc.cc:28-80); fbank = 1: FbankComputer::Compute (feature-fbank.cc:73-123),
 * which shares everything up to the mel energies. */
typedef struct { int32_t lpc_order; float compress_factor, cepstral_scale; } plp_extra;
static int frontend_impl(const orc_mfcc_opts *o, int fbank, int use_log_fbank, int use_power, const plp_extra *plp,
                         const float *wave, int64_t n_samp, float vtln_warp, float *out, int32_t out_stride) {
  if (o->dither != 0.0f) return -3;           /* rand()-based dither is not reproducible; parity runs use 0 */
  if (!o->round_to_power_of_two) return -3;   /* the non-pow2 RealFft branch (feature-mfcc.cc:43-44) is not restated */
  int32_t L = orc_window_size(o), Npad = orc_padded_window_size(o), B = o->num_bins, C = o->num_ceps;
  int32_t T = orc_num_frames(n_samp, o);
  if (T == 0) return 0;
  int32_t nfft = Npad / 2;
  float *window = (float *)malloc(sizeof(float) * L);
  int32_t *offs = (int32_t *)malloc(sizeof(int32_t) * B), *lens = (int32_t *)malloc(sizeof(int32_t) * B);
  float *melw = (float *)malloc(sizeof(float) * (size_t)B * nfft);
  float *dct = (float *)malloc(sizeof(float) * C * B), *lift = (float *)malloc(sizeof(float) * C);
  float *frame = (float *)malloc(sizeof(float) * Npad), *scr = (float *)malloc(sizeof(float) * Npad);
  float *mel = (float *)malloc(sizeof(float) * (B + 2));
  float *eq = NULL, *idft = NULL, ac[64], lpc[64], ltmp[64], cep[64];
  if (plp) { /* PlpComputer ctor, feature-plp.cc:25-50 */
    if (plp->lpc_order < 1 || plp->lpc_order > 60 || C > plp->lpc_order + 1) { free(mel); return -3; }
    eq = (float *)malloc(sizeof(float) * B);
    idft = (float *)malloc(sizeof(float) * (plp->lpc_order + 1) * (B + 2));
    equal_loudness(o, vtln_warp, eq);
    idft_bases(plp->lpc_order + 1, B + 2, idft);
  }
  int rc = orc_window_table(o, window);
  if (rc == 0) rc = orc_mel_banks(o, vtln_warp, offs, lens, melw);
  if (rc != 0) goto done;
  orc_dct_matrix(C, B, dct);
  if (o->cepstral_lifter != 0.0f) orc_lifter_coeffs(o->cepstral_lifter, C, lift);
  float log_energy_floor = (o->energy_floor > 0.0f) ? logf(o->energy_floor) : 0.0f;

  for (int32_t r = 0; r < T; r++) {
    /* ExtractWindow, feature-window.cc:162-220 */
    int64_t start = orc_first_sample_of_frame(r, o);
    if (start >= 0 && start + L <= n_samp) {
      for (int32_t s = 0; s < L; s++) frame[s] = wave[start + s];
    } else {
      for (int32_t s = 0; s < L; s++) {
        int64_t k = s + start;
        while (k < 0 || k >= n_samp) k = (k < 0) ? -k - 1 : 2 * n_samp - 1 - k;
        frame[s] = wave[k];
      }
    }
    for (int32_t s = L; s < Npad; s++) frame[s] = 0.0f;
    /* ProcessWindow, feature-window.cc:133-156 */
    if (o->remove_dc_offset) {
      float sum = 0.0f;
      for (int32_t s = 0; s < L; s++) sum += frame[s];
      float m = -sum / L;
      for (int32_t s = 0; s < L; s++) frame[s] += m;
    }
    float log_energy = 0.0f;
    if (o->use_energy && o->raw_energy) {
      float e = 0.0f;
      for (int32_t s = 0; s < L; s++) e += frame[s] * frame[s];
      log_energy = logf(e > FLT_EPSILON ? e : FLT_EPSILON);
    }
    if (o->preemph_coeff != 0.0f) { /* :101-107 */
      for (int32_t s = L - 1; s > 0; s--) frame[s] -= o->preemph_coeff * frame[s - 1];
      frame[0] -= o->preemph_coeff * frame[0];
    }
    for (int32_t s = 0; s < L; s++) frame[s] *= window[s];
    /* MfccComputer::Compute, feature-mfcc.cc:28-80 */
    if (o->use_energy && !o->raw_energy) {
      float e = 0.0f;
      for (int32_t s = 0; s < Npad; s++) e += frame[s] * frame[s];
      log_energy = logf(e > FLT_MIN ? e : FLT_MIN);
    }
    real_fft_forward(frame, Npad, scr);
    { /* ComputePowerSpectrum, feature-functions.cc:29-51 */
      float first = frame[0] * frame[0], last = frame[1] * frame[1];
      for (int32_t i = 1; i < nfft; i++) {
        float re = frame[2 * i], im = frame[2 * i + 1];
        frame[i] = re * re + im * im;
      }
      frame[0] = first;
      frame[nfft] = last;
    }
    if (fbank && !use_power) /* power_spectrum.ApplyPow(0.5), feature-fbank.cc:97-98 */
      for (int32_t i = 0; i <= nfft; i++) frame[i] = sqrtf(frame[i]);
    for (int32_t b = 0; b < B; b++) { /* MelBanks::Compute, mel-computations.cc:228-253 */
      float e = 0.0f;
      const float *w = melw + (size_t)b * nfft;
      for (int32_t i = 0; i < lens[b]; i++) e += w[i] * frame[offs[b] + i];
      if (o->htk_mode && e < 1.0f) e = 1.0f;
      if (plp) { mel[b] = e; continue; }
      if (!fbank || use_log_fbank) {
        if (e < FLT_EPSILON) e = FLT_EPSILON; /* feature-mfcc.cc:54, feature-fbank.cc:109 */
        e = logf(e);                          /* :55, :110 */
      }
      mel[b] = e;
    }
    float *feat = out + (size_t)r * out_stride;
    if (plp) { /* PlpComputer::Compute, feature-plp.cc:143-188 */
      int32_t n = plp->lpc_order;
      for (int32_t b = B - 1; b >= 0; b--) mel[b + 1] = powf(mel[b] * eq[b], plp->compress_factor); /* MulElements, ApplyPow */
      mel[0] = mel[1];
      mel[B + 1] = mel[B];
      for (int32_t i = 0; i <= n; i++) { /* AddMatVec(1.0, idft_bases, kNoTrans, mel_dup, 0.0) */
        float s = 0.0f;
        for (int32_t j = 0; j < B + 2; j++) s += idft[i * (B + 2) + j] * mel[j];
        ac[i] = s;
      }
      for (int32_t i = 0; i < n; i++) lpc[i] = 0.0f;
      float E = durbin(n, ac, lpc, ltmp);
      float res = (float)(-log(1.0 / E)); /* ComputeLpc: -Log(1.0 / ans) */
      if (res < FLT_MIN) res = FLT_MIN;
      lpc2cepstrum(n, lpc, cep);
      feat[0] = res;
      for (int32_t k = 1; k < C; k++) feat[k] = cep[k - 1];
      if (o->cepstral_lifter != 0.0f)
        for (int32_t k = 0; k < C; k++) feat[k] *= lift[k];
      if (plp->cepstral_scale != 1.0f)
        for (int32_t k = 0; k < C; k++) feat[k] *= plp->cepstral_scale;
      if (o->use_energy) {
        if (o->energy_floor > 0.0f && log_energy < log_energy_floor) log_energy = log_energy_floor;
        feat[0] = log_energy;
      }
      if (o->htk_compat) {
        float e = feat[0];
        for (int32_t i = 0; i < C - 1; i++) feat[i] = feat[i + 1];
        feat[C - 1] = e;
      }
      continue;
    }
    if (fbank) { /* feature-fbank.cc:100-121: energy first, or last with htk_compat */
      int32_t mel_offset = (o->use_energy && !o->htk_compat) ? 1 : 0;
      for (int32_t b = 0; b < B; b++) feat[mel_offset + b] = mel[b];
      if (o->use_energy) {
        if (o->energy_floor > 0.0f && log_energy < log_energy_floor) log_energy = log_energy_floor;
        feat[o->htk_compat ? B : 0] = log_energy;
      }
      continue;
    }
    for (int32_t k = 0; k < C; k++) { /* :59 */
      float s = 0.0f;
      for (int32_t b = 0; b < B; b++) s += dct[k * B + b] * mel[b];
      feat[k] = s;
    }
    if (o->cepstral_lifter != 0.0f)
      for (int32_t k = 0; k < C; k++) feat[k] *= lift[k];
    if (o->use_energy) {
      if (o->energy_floor > 0.0f && log_energy < log_energy_floor) log_energy = log_energy_floor;
      feat[0] = log_energy;
    }
    if (o->htk_compat) { /* :70-79 */
      float e = feat[0];
      for (int32_t i = 0; i < C - 1; i++) feat[i] = feat[i + 1];
      if (!o->use_energy) e *= (float)1.41421356237309504880;
      feat[C - 1] = e;
    }
  }
  rc = T;
done:
  free(window); free(offs); free(lens); free(melw); free(dct); free(lift); free(frame); free(scr); free(mel);
  free(eq); free(idft);
  return rc;
}

int orc_mfcc_compute(const orc_mfcc_opts *o, const float *wave, int64_t n_samp, float vtln_warp, float *out,
                     int32_t out_stride) {
  return frontend_impl(o, 0, 1, 1, NULL, wave, n_samp, vtln_warp, out, out_stride);
}

/* OfflineFeatureTpl<FbankComputer>: out has num_bins (+1 with use_energy) columns; num_ceps / cepstral_lifter unused. */
int orc_fbank_compute(const orc_mfcc_opts *o, int32_t use_log_fbank, int32_t use_power, const float *wave, int64_t n_samp,
                      float vtln_warp, float *out, int32_t out_stride) {
  return frontend_impl(o, 1, use_log_fbank, use_power, NULL, wave, n_samp, vtln_warp, out, out_stride);
}

/* OfflineFeatureTpl<PlpComputer> (feat/feature-plp.cc): num_ceps columns (C0 = LPC residual log-energy, or the frame
 * log-energy with use_energy); frame / mel / energy / lifter / htk_compat options from o. */
int orc_plp_compute(const orc_mfcc_opts *o, int32_t lpc_order, float compress_factor, float cepstral_scale,
                    const float *wave, int64_t n_samp, float vtln_warp, float *out, int32_t out_stride) {
  plp_extra p = {lpc_order, compress_factor, cepstral_scale};
  return frontend_impl(o, 0, 1, 1, &p, wave, n_samp, vtln_warp, out, out_stride);
}

/* transform/cmvn.cc:30-62 (weight 1.0 per frame) */
void orc_cmvn_acc(const float *feats, int32_t T, int32_t D, int32_t stride, double *stats) {
  double *mean = stats, *var = stats + (D + 1);
  for (int32_t t = 0; t < T; t++) {
    const float *x = feats + (size_t)t * stride;
    mean[D] += 1.0;
    for (int32_t d = 0; d < D; d++) {
      mean[d] += x[d] * 1.0;
      var[d] += x[d] * x[d] * 1.0; /* float product, then promoted: `*feats_ptr * *feats_ptr * weight` */
    }
  }
}

/* transform/cmvn.cc:64-113 */
int orc_cmvn_apply(const double *stats, int32_t D, int32_t norm_vars, float *feats, int32_t T, int32_t stride) {
  double count = stats[D];
  if (count < 1.0) return -1;
  float *off = (float *)malloc(sizeof(float) * D), *scl = (float *)malloc(sizeof(float) * D);
  for (int32_t d = 0; d < D; d++) {
    double mean = stats[d] / count, offset, scale;
    if (!norm_vars) {
      scale = 1.0;
      offset = -mean;
    } else {
      double var = stats[(D + 1) + d] / count - mean * mean, floor = 1.0e-20;
      if (var < floor) var = floor;
      scale = 1.0 / sqrt(var);
      if (scale != scale || 1 / scale == 0.0) { free(off); free(scl); return -2; }
      offset = -(mean * scale);
    }
    off[d] = (float)offset;
    scl[d] = (float)scale;
  }
  for (int32_t t = 0; t < T; t++) {
    float *x = feats + (size_t)t * stride;
    if (norm_vars)
      for (int32_t d = 0; d < D; d++) x[d] *= scl[d];
    for (int32_t d = 0; d < D; d++) x[d] += off[d];
  }
  free(off); free(scl);
  return 0;
}

/* DeltaFeatures ctor, feat/feature-functions.cc:54-86.  scales row i has lens[i] = 2*i*window+1 taps,
 * rows are laid out with pitch (2*order*window+1). */
int orc_delta_scales(int32_t order, int32_t window, float *scales, int32_t *lens) {
  if (order < 0 || order >= 1000 || window <= 0 || window >= 1000) return -1;
  int32_t pitch = 2 * order * window + 1;
  memset(scales, 0, sizeof(float) * (size_t)(order + 1) * pitch);
  scales[0] = 1.0f;
  lens[0] = 1;
  for (int32_t i = 1; i <= order; i++) {
    const float *prev = scales + (size_t)(i - 1) * pitch;
    float *cur = scales + (size_t)i * pitch;
    int32_t prev_off = (lens[i - 1] - 1) / 2, cur_off = prev_off + window;
    lens[i] = lens[i - 1] + 2 * window;
    float normalizer = 0.0f;
    for (int32_t j = -window; j <= window; j++) {
      normalizer += j * j;
      for (int32_t k = -prev_off; k <= prev_off; k++) cur[j + k + cur_off] += (float)j * prev[k + prev_off];
    }
    for (int32_t k = 0; k < lens[i]; k++) cur[k] *= (float)(1.0 / normalizer); /* Scale(1.0/normalizer) */
  }
  return 0;
}

/* DeltaFeatures::Process + ComputeDeltas, feature-functions.cc:88-111,160-171 */
void orc_deltas(int32_t order, int32_t window, const float *in, int32_t T, int32_t D, int32_t in_stride, float *out,
                int32_t out_stride) {
  int32_t pitch = 2 * order * window + 1;
  float *scales = (float *)malloc(sizeof(float) * (size_t)(order + 1) * pitch);
  int32_t *lens = (int32_t *)malloc(sizeof(int32_t) * (order + 1));
  orc_delta_scales(order, window, scales, lens);
  for (int32_t t = 0; t < T; t++) {
    float *orow = out + (size_t)t * out_stride;
    for (int32_t k = 0; k < D * (order + 1); k++) orow[k] = 0.0f;
    for (int32_t i = 0; i <= order; i++) {
      const float *sc = scales + (size_t)i * pitch;
      int32_t max_off = (lens[i] - 1) / 2;
      float *o = orow + i * D;
      for (int32_t j = -max_off; j <= max_off; j++) {
        int32_t f = t + j;
        if (f < 0) f = 0;
        else if (f >= T) f = T - 1;
        float s = sc[j + max_off];
        if (s != 0.0f) {
          const float *x = in + (size_t)f * in_stride;
          for (int32_t d = 0; d < D; d++) o[d] += s * x[d]; /* saxpy */
        }
      }
    }
  }
  free(scales); free(lens);
}

/* SpliceFrames, feature-functions.cc:205-226 */
void orc_splice(const float *in, int32_t T, int32_t D, int32_t in_stride, int32_t left, int32_t right, float *out,
                int32_t out_stride) {
  int32_t N = 1 + left + right;
  for (int32_t t = 0; t < T; t++)
    for (int32_t j = 0; j < N; j++) {
      int32_t t2 = t + j - left;
      if (t2 < 0) t2 = 0;
      if (t2 >= T) t2 = T - 1;
      memcpy(out + (size_t)t * out_stride + (size_t)j * D, in + (size_t)t2 * in_stride, sizeof(float) * D);
    }
}

/* transform-feats.cpp:95-107 */
int orc_transform(const float *in, int32_t T, int32_t D, int32_t in_stride, const float *mat, int32_t rows,
                  int32_t cols, float *out, int32_t out_stride) {
  if (cols != D && cols != D + 1) return -1;
  for (int32_t t = 0; t < T; t++) {
    const float *x = in + (size_t)t * in_stride;
    float *y = out + (size_t)t * out_stride;
    for (int32_t r = 0; r < rows; r++) {
      const float *m = mat + (size_t)r * cols;
      float s = 0.0f;
      for (int32_t d = 0; d < D; d++) s += x[d] * m[d];
      if (cols == D + 1) s += m[D]; /* AddVecToRows after the GEMM */
      y[r] = s;
    }
  }
  return 0;
}

/* DiagGmm::ComputeGconsts, gmm/diag-gmm.cc:114-152 */
int orc_gconsts(int32_t M, int32_t D, const float *weights, const float *miv, const float *iv, float *gconsts) {
  float offset = (float)(-0.5 * ORC_LOG_2PI * D);
  int bad = 0;
  for (int32_t m = 0; m < M; m++) {
    float gc = logf(weights[m]) + offset;
    for (int32_t d = 0; d < D; d++) {
      float a = iv[(size_t)m * D + d], b = miv[(size_t)m * D + d];
      /* `gc += 0.5 * Log(iv) - 0.5 * miv * miv / iv` : right-hand side in double, += rounds to float */
      gc = (float)(gc + (0.5 * logf(a) - 0.5 * b * b / a));
    }
    if (isnan(gc)) return -1;
    if (isinf(gc)) {
      bad++;
      if (gc > 0) gc = -gc;
    }
    gconsts[m] = gc;
  }
  return bad;
}

/* per-Gaussian loglikes of one pdf for one frame: diag-gmm.cc:528-543 / decodable-am-diag-gmm.cc:58-62 */
static void pdf_loglikes(int32_t M, int32_t D, const float *gc, const float *miv, const float *iv, const float *x,
                         const float *xsq, float *ll) {
  for (int32_t m = 0; m < M; m++) {
    float a = 0.0f, b = 0.0f;
    const float *mr = miv + (size_t)m * D, *vr = iv + (size_t)m * D;
    for (int32_t d = 0; d < D; d++) a += mr[d] * x[d];
    for (int32_t d = 0; d < D; d++) b += vr[d] * xsq[d];
    float v = gc[m];
    v = 1.0f * a + 1.0f * v;  /* sgemv(alpha=1, beta=1) */
    v = -0.5f * b + 1.0f * v; /* sgemv(alpha=-0.5, beta=1) */
    ll[m] = v;
  }
}

/* VectorBase<float>::LogSumExp, matrix/kaldi-vector.cc:757-775 */
static float log_sum_exp(const float *v, int32_t n, float prune) {
  float mx = v[0];
  for (int32_t i = 1; i < n; i++)
    if (v[i] > mx) mx = v[i];
  float cutoff = mx + logf(FLT_EPSILON);
  if (prune > 0.0f && mx - prune > cutoff) cutoff = mx - prune;
  double sum = 0.0;
  for (int32_t i = 0; i < n; i++)
    if (v[i] >= cutoff) sum += expf(v[i] - mx);
  return (float)(mx + log(sum)); /* Log(double) then narrowed to Real */
}

int orc_gmm_loglikes(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                     const float *iv, const float *feats, int32_t T, int32_t stride, float prune, float *out,
                     int32_t out_stride) {
  int32_t maxM = 0;
  for (int32_t p = 0; p < P; p++)
    if (pdf_offsets[p + 1] - pdf_offsets[p] > maxM) maxM = pdf_offsets[p + 1] - pdf_offsets[p];
  float *ll = (float *)malloc(sizeof(float) * (maxM > 0 ? maxM : 1));
  float *xsq = (float *)malloc(sizeof(float) * D);
  int rc = 0;
  for (int32_t t = 0; t < T; t++) {
    const float *x = feats + (size_t)t * stride;
    for (int32_t d = 0; d < D; d++) xsq[d] = x[d] * x[d]; /* ApplyPow(2.0) */
    for (int32_t p = 0; p < P; p++) {
      int32_t g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
      pdf_loglikes(M, D, gconsts + g0, miv + (size_t)g0 * D, iv + (size_t)g0 * D, x, xsq, ll);
      float s = log_sum_exp(ll, M, prune);
      if (isnan(s) || isinf(s)) rc = -2; /* decodable-am-diag-gmm.cc:65-66 is a KALDI_ERR */
      out[(size_t)t * out_stride + p] = s;
    }
  }
  free(ll); free(xsq);
  return rc;
}

static int acc_impl(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                    const float *iv, const float *feats1, const float *feats2, int32_t T, int32_t stride,
                    const int32_t *pdf_ids, const float *weights, double *occ, double *mean_acc, double *var_acc,
                    double *tot_like, double *tot_frames) {
  int32_t maxM = 0;
  for (int32_t p = 0; p < P; p++)
    if (pdf_offsets[p + 1] - pdf_offsets[p] > maxM) maxM = pdf_offsets[p + 1] - pdf_offsets[p];
  float *post = (float *)malloc(sizeof(float) * (maxM > 0 ? maxM : 1));
  float *xsq = (float *)malloc(sizeof(float) * D);
  int rc = 0;
  for (int32_t t = 0; t < T; t++) {
    int32_t p = pdf_ids[t];
    if (p < 0 || p >= P) { rc = -1; break; }
    float w = weights ? weights[t] : 1.0f;
    const float *x = feats1 + (size_t)t * stride, *y = feats2 + (size_t)t * stride;
    for (int32_t d = 0; d < D; d++) xsq[d] = x[d] * x[d];
    int32_t g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
    pdf_loglikes(M, D, gconsts + g0, miv + (size_t)g0 * D, iv + (size_t)g0 * D, x, xsq, post);
    /* ApplySoftMax, kaldi-vector.cc:852-859 (float sum) */
    float mx = post[0];
    for (int32_t m = 1; m < M; m++)
      if (post[m] > mx) mx = post[m];
    float sum = 0.0f;
    for (int32_t m = 0; m < M; m++) sum += (post[m] = expf(post[m] - mx));
    float inv = (float)(1.0 / sum);
    for (int32_t m = 0; m < M; m++) post[m] *= inv;
    float log_like = mx + logf(sum);
    if (isnan(log_like) || isinf(log_like)) { rc = -2; break; } /* diag-gmm.cc:609-610 */
    for (int32_t m = 0; m < M; m++) post[m] *= w; /* posteriors.Scale(frame_posterior), mle-diag-gmm.cc:200 */
    /* AccumulateFromPosteriors, mle-diag-gmm.cc:171-189 */
    for (int32_t m = 0; m < M; m++) {
      double g = (double)post[m];
      occ[g0 + m] += g;
      double *ma = mean_acc + (size_t)(g0 + m) * D, *va = var_acc + (size_t)(g0 + m) * D;
      for (int32_t d = 0; d < D; d++) {
        double yd = (double)y[d];
        ma[d] += g * yd;
        va[d] += g * (yd * yd);
      }
    }
    *tot_like += log_like * w; /* mle-am-diag-gmm.cc:76-77: float product, added to double */
    *tot_frames += w;
  }
  free(post); free(xsq);
  return rc;
}

int orc_acc_ali(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                const float *iv, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids,
                const float *weights, double *occ, double *mean_acc, double *var_acc, double *tot_like,
                double *tot_frames) {
  return acc_impl(P, D, pdf_offsets, gconsts, miv, iv, feats, feats, T, stride, pdf_ids, weights, occ, mean_acc,
                  var_acc, tot_like, tot_frames);
}

int orc_acc_ali_twofeats(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                         const float *iv, const float *feats1, const float *feats2, int32_t T, int32_t stride,
                         const int32_t *pdf_ids, const float *weights, double *occ, double *mean_acc,
                         double *var_acc, double *tot_like, double *tot_frames) {
  return acc_impl(P, D, pdf_offsets, gconsts, miv, iv, feats1, feats2, T, stride, pdf_ids, weights, occ, mean_acc,
                  var_acc, tot_like, tot_frames);
}

/* FmllrDiagGmmAccs over an alignment (gmm-est-fmllr.cpp:40-55 -> transform/fmllr-diag-gmm.cc:110-121
 * AccumulateForGmm -> :30-45 AccumulateFromPosteriors -> :562-583 CommitSingleFrameStats, update_type "full").
 * One (pdf, weight) per frame.  beta, K[D][D+1] and G[D][(D+1)(D+2)/2] (SpMatrix packing: row-major lower triangle)
 * are ADDED to.  tot_like receives the sum of the frames' log-likelihoods (AccumulateForGmm's return value). */
int orc_fmllr_acc(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                  const float *iv, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids,
                  const float *weights, double *beta, double *K, double *G, double *tot_like) {
  int32_t maxM = 0;
  for (int32_t p = 0; p < P; p++)
    if (pdf_offsets[p + 1] - pdf_offsets[p] > maxM) maxM = pdf_offsets[p + 1] - pdf_offsets[p];
  float *post = (float *)malloc(sizeof(float) * (maxM > 0 ? maxM : 1));
  float *xsq = (float *)malloc(sizeof(float) * D), *a = (float *)malloc(sizeof(float) * D),
        *b = (float *)malloc(sizeof(float) * D);
  double *xp = (double *)malloc(sizeof(double) * (D + 1));
  const int32_t np = (D + 1) * (D + 2) / 2;
  int rc = 0;
  for (int32_t t = 0; t < T; t++) {
    int32_t p = pdf_ids[t];
    if (p < 0 || p >= P) { rc = -1; break; }
    float w = weights ? weights[t] : 1.0f;
    const float *x = feats + (size_t)t * stride;
    for (int32_t d = 0; d < D; d++) xsq[d] = x[d] * x[d];
    int32_t g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
    /* ComponentPosteriors, diag-gmm.cc:601-615 */
    pdf_loglikes(M, D, gconsts + g0, miv + (size_t)g0 * D, iv + (size_t)g0 * D, x, xsq, post);
    float mx = post[0];
    for (int32_t m = 1; m < M; m++)
      if (post[m] > mx) mx = post[m];
    float sum = 0.0f;
    for (int32_t m = 0; m < M; m++) sum += (post[m] = expf(post[m] - mx));
    float inv = (float)(1.0 / sum);
    for (int32_t m = 0; m < M; m++) post[m] *= inv;
    float log_like = mx + logf(sum);
    if (isnan(log_like) || isinf(log_like)) { rc = -2; break; }
    for (int32_t m = 0; m < M; m++) post[m] *= w; /* posterior.Scale(weight), fmllr-diag-gmm.cc:118 */
    /* AccumulateFromPosteriors: count += posterior.Sum(); a += means_invvars^T post; b += inv_vars^T post (float) */
    double count = 0.0;
    { double ps = 0.0; for (int32_t m = 0; m < M; m++) ps += post[m]; count = (double)(float)ps; } /* Sum(): double, returned as Real (kaldi-vector.cc:692-696) */
    for (int32_t d = 0; d < D; d++) a[d] = b[d] = 0.0f;
    for (int32_t m = 0; m < M; m++) {
      const float *mr = miv + (size_t)(g0 + m) * D, *vr = iv + (size_t)(g0 + m) * D;
      for (int32_t d = 0; d < D; d++) { a[d] += post[m] * mr[d]; b[d] += post[m] * vr[d]; }
    }
    *tot_like += log_like;
    if (count == 0.0) continue; /* CommitSingleFrameStats returns early */
    for (int32_t d = 0; d < D; d++) xp[d] = (double)x[d];
    xp[D] = 1.0;
    *beta += count;
    for (int32_t i = 0; i < D; i++)
      for (int32_t k = 0; k <= D; k++) K[(size_t)i * (D + 1) + k] += (double)a[i] * xp[k]; /* K_.AddVecVec */
    for (int32_t i = 0; i < D; i++) {                                                      /* G_[i].AddSp(b(i), scatter) */
      double *g = G + (size_t)i * np, bi = (double)b[i];
      for (int32_t j = 0; j <= D; j++)
        for (int32_t k = 0; k <= j; k++) g[j * (j + 1) / 2 + k] += bi * (xp[j] * xp[k]);
    }
  }
  free(post); free(xsq); free(a); free(b); free(xp);
  return rc;
}

/* MlltAccs over an alignment (gmm-acc-mllt.cpp:100-112 -> transform/mllt.cc:162-170 AccumulateFromGmm -> :131-160
 * AccumulateFromPosteriors), rand_prune = 0.  beta and G[D][D(D+1)/2] (SpMatrix packing) are ADDED to; tot_like receives
 * the sum of loglike * weight (the driver's tot_like_this_file). */
int orc_mllt_acc(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                 const float *iv, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids,
                 const float *weights, double *beta, double *G, double *tot_like) {
  int32_t maxM = 0;
  for (int32_t p = 0; p < P; p++)
    if (pdf_offsets[p + 1] - pdf_offsets[p] > maxM) maxM = pdf_offsets[p + 1] - pdf_offsets[p];
  float *post = (float *)malloc(sizeof(float) * (maxM > 0 ? maxM : 1));
  float *xsq = (float *)malloc(sizeof(float) * D), *mean = (float *)malloc(sizeof(float) * D);
  double *off = (double *)malloc(sizeof(double) * D);
  const int32_t np = D * (D + 1) / 2;
  int rc = 0;
  for (int32_t t = 0; t < T; t++) {
    int32_t p = pdf_ids[t];
    if (p < 0 || p >= P) { rc = -1; break; }
    float w = weights ? weights[t] : 1.0f;
    const float *x = feats + (size_t)t * stride;
    for (int32_t d = 0; d < D; d++) xsq[d] = x[d] * x[d];
    int32_t g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
    pdf_loglikes(M, D, gconsts + g0, miv + (size_t)g0 * D, iv + (size_t)g0 * D, x, xsq, post);
    float mx = post[0];
    for (int32_t m = 1; m < M; m++)
      if (post[m] > mx) mx = post[m];
    float sum = 0.0f;
    for (int32_t m = 0; m < M; m++) sum += (post[m] = expf(post[m] - mx));
    float inv = (float)(1.0 / sum);
    for (int32_t m = 0; m < M; m++) post[m] *= inv;
    float log_like = mx + logf(sum);
    if (isnan(log_like) || isinf(log_like)) { rc = -2; break; }
    for (int32_t m = 0; m < M; m++) post[m] *= w; /* posteriors.Scale(weight) */
    double this_beta = 0.0;
    for (int32_t m = 0; m < M; m++) {
      float po = post[m]; /* RandPrune(post, 0.0) == post */
      if (po == 0.0f) continue;
      const float *mr = miv + (size_t)(g0 + m) * D, *vr = iv + (size_t)(g0 + m) * D;
      for (int32_t d = 0; d < D; d++) { mean[d] = mr[d] / vr[d]; mean[d] += -1.0f * x[d]; off[d] = (double)mean[d]; }
      for (int32_t j = 0; j < D; j++) { /* G_[j].AddSp(inv_var(j) * posterior, offset offset^T) */
        double a = (double)(vr[j] * po), *g = G + (size_t)j * np;
        for (int32_t r = 0; r < D; r++)
          for (int32_t c = 0; c <= r; c++) g[r * (r + 1) / 2 + c] += a * (off[r] * off[c]);
      }
      this_beta += po;
    }
    *beta += this_beta;
    *tot_like += log_like * w;
  }
  free(post); free(xsq); free(mean); free(off);
  return rc;
}

/* DiagGmm::ComponentPosteriors (diag-gmm.cc:601-615: LogLikelihoods + ApplySoftMax) of each frame's aligned pdf, then
 * Scale(weight) as gmm-post-to-gpost does.  post: the frames' posterior vectors back to back; loglikes[T]. */
int orc_component_posteriors(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                             const float *iv, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids,
                             const float *weights, float *post_out, float *loglikes) {
  float *xsq = (float *)malloc(sizeof(float) * D);
  int rc = 0;
  size_t o = 0;
  for (int32_t t = 0; t < T; t++) {
    int32_t p = pdf_ids[t];
    if (p < 0 || p >= P) { rc = -1; break; }
    const float *x = feats + (size_t)t * stride;
    for (int32_t d = 0; d < D; d++) xsq[d] = x[d] * x[d];
    int32_t g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
    float *post = post_out + o;
    pdf_loglikes(M, D, gconsts + g0, miv + (size_t)g0 * D, iv + (size_t)g0 * D, x, xsq, post);
    float mx = post[0];
    for (int32_t m = 1; m < M; m++)
      if (post[m] > mx) mx = post[m];
    float sum = 0.0f;
    for (int32_t m = 0; m < M; m++) sum += (post[m] = expf(post[m] - mx));
    float inv = (float)(1.0 / sum);
    for (int32_t m = 0; m < M; m++) post[m] *= inv;
    float log_like = mx + logf(sum);
    if (isnan(log_like) || isinf(log_like)) { rc = -2; break; }
    float w = weights ? weights[t] : 1.0f;
    for (int32_t m = 0; m < M; m++) post[m] *= w;
    if (loglikes) loglikes[t] = log_like;
    o += M;
  }
  free(xsq);
  return rc;
}

/* ===================================================================================================
 * Kaldi pitch: feat/resample.cc + feat/pitch-functions.cc, offline path (see oracle.h).
 * =================================================================================================== */
#define ORC_2PI 6.283185307179586476925286766559005
#define ORC_PI 3.1415926535897932384626433832795

void orc_pitch_opts_default(orc_pitch_opts *o) { /* pitch-functions.h:103-123 */
  o->samp_freq = 16000; o->frame_shift_ms = 10; o->frame_length_ms = 25; o->preemph_coeff = 0; o->min_f0 = 50;
  o->max_f0 = 400; o->soft_min_f0 = 10; o->penalty_factor = 0.1f; o->lowpass_cutoff = 1000; o->resample_freq = 4000;
  o->delta_pitch = 0.005f; o->nccf_ballast = 7000; o->lowpass_filter_width = 1; o->upsample_filter_width = 5;
  o->recompute_frame = 500; o->snip_edges = 1;
}
void orc_process_pitch_opts_default(orc_process_pitch_opts *o) { /* pitch-functions.h:241-255 */
  o->pitch_scale = 2; o->pov_scale = 2; o->pov_offset = 0; o->delta_pitch_scale = 10; o->delta_pitch_noise_stddev = 0.005f;
  o->normalization_left_context = 75; o->normalization_right_context = 75; o->delta_window = 2; o->delay = 0;
  o->add_pov_feature = 1; o->add_normalized_log_pitch = 1; o->add_delta_pitch = 1; o->add_raw_log_pitch = 0;
}

/* LinearResample::FilterFunc / ArbitraryResample::FilterFunc (resample.cc:213-226, 318-331): Hanning-windowed sinc. */
static float resample_filter(float t, float cutoff, int32_t num_zeros) {
  float window, filter;
  if (fabs(t) < num_zeros / (2.0 * cutoff)) window = (float)(0.5 * (1 + cos(ORC_2PI * cutoff / num_zeros * t)));
  else window = 0.0f;
  if (t != 0) filter = (float)(sin(ORC_2PI * cutoff * t) / (ORC_PI * t));
  else filter = (float)(2 * cutoff);
  return filter * window;
}
static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

typedef struct {
  int32_t in_hz, out_hz, in_unit, out_unit, num_zeros, max_w;
  float cutoff;
  int32_t *first; int32_t *nw; float *w; /* per output phase: first input index, #weights, weights[max_w] */
} lin_resamp;

static void lin_resamp_init(lin_resamp *r, int32_t in_hz, int32_t out_hz, float cutoff, int32_t num_zeros) { /* resample.cc:34-55,82-106 */
  r->in_hz = in_hz; r->out_hz = out_hz; r->cutoff = cutoff; r->num_zeros = num_zeros;
  int32_t base = (int32_t)gcd64(in_hz, out_hz);
  r->in_unit = in_hz / base; r->out_unit = out_hz / base;
  double window_width = num_zeros / (2.0 * cutoff);
  r->first = (int32_t *)malloc(sizeof(int32_t) * r->out_unit);
  r->nw = (int32_t *)malloc(sizeof(int32_t) * r->out_unit);
  r->max_w = 0;
  for (int32_t i = 0; i < r->out_unit; i++) {
    double output_t = i / (double)out_hz, min_t = output_t - window_width, max_t = output_t + window_width;
    int32_t lo = (int32_t)ceil(min_t * in_hz), hi = (int32_t)floor(max_t * in_hz);
    r->first[i] = lo; r->nw[i] = hi - lo + 1;
    if (r->nw[i] > r->max_w) r->max_w = r->nw[i];
  }
  r->w = (float *)calloc((size_t)r->out_unit * r->max_w, sizeof(float));
  for (int32_t i = 0; i < r->out_unit; i++) {
    double output_t = i / (double)out_hz;
    for (int32_t j = 0; j < r->nw[i]; j++) {
      double input_t = (r->first[i] + j) / (double)in_hz, delta_t = input_t - output_t;
      r->w[(size_t)i * r->max_w + j] = resample_filter((float)delta_t, cutoff, num_zeros) / in_hz;
    }
  }
}
static void lin_resamp_free(lin_resamp *r) { free(r->first); free(r->nw); free(r->w); }

static int64_t lin_resamp_num_out(const lin_resamp *r, int64_t n_in, int flush) { /* resample.cc:57-80 */
  int64_t tick_freq = (int64_t)r->in_hz / gcd64(r->in_hz, r->out_hz) * r->out_hz;
  int64_t ticks_per_in = tick_freq / r->in_hz;
  int64_t len = n_in * ticks_per_in;
  if (!flush) {
    float window_width = (float)(r->num_zeros / (2.0 * r->cutoff));
    int32_t wt = (int32_t)floor(window_width * (int32_t)tick_freq);
    len -= wt;
  }
  if (len <= 0) return 0;
  int64_t ticks_per_out = tick_freq / r->out_hz;
  int64_t last = len / ticks_per_out;
  if (last * ticks_per_out == len) last--;
  return last + 1;
}

/* One output sample of LinearResample::Resample (resample.cc:120-160).  valid_lo/valid_hi: the input indexes the call
 * can see (the first call sees [0, n); the flush call sees only the kept remainder [n - R, n)). */
static float lin_resamp_sample(const lin_resamp *r, const float *in, int64_t valid_lo, int64_t valid_hi, int64_t samp_out) {
  int64_t unit = samp_out / r->out_unit;
  int32_t ph = (int32_t)(samp_out - unit * r->out_unit);
  int64_t first = r->first[ph] + unit * r->in_unit;
  const float *w = r->w + (size_t)ph * r->max_w;
  double acc = 0.0; /* the reference uses float VecVec / a float running sum; see the tolerance note in the tests */
  for (int32_t i = 0; i < r->nw[ph]; i++) {
    int64_t idx = first + i;
    if (idx >= valid_lo && idx < valid_hi) acc += (double)w[i] * in[idx];
  }
  return (float)acc;
}

/* DownsampleWaveForm (feat/resample.cc:368-376): what OfflineFeatureTpl::ComputeFeatures does to a wave whose rate is above
 * the options' when allow_downsample is set (feat/feature-common-inl.h:29-55).  One flushed LinearResample call with
 * cutoff 0.99 * new_freq / 2 and 6 zeros.  out nullable: returns the number of output samples (<0: bad arguments). */
int64_t orc_downsample_waveform(float orig_freq, float new_freq, const float *wave, int64_t n, float *out) {
  if (!(new_freq < orig_freq) || !(new_freq >= 1.0f)) return -1;
  lin_resamp r;
  float cutoff = (float)(0.99 * 0.5 * new_freq);
  lin_resamp_init(&r, (int32_t)orig_freq, (int32_t)new_freq, cutoff, 6);
  int64_t n_out = lin_resamp_num_out(&r, n, 1);
  if (out)
    for (int64_t i = 0; i < n_out; i++) out[i] = lin_resamp_sample(&r, wave, 0, n, i);
  lin_resamp_free(&r);
  return n_out;
}

typedef struct {
  orc_pitch_opts o;
  lin_resamp lr;
  int32_t first_lag, last_lag, n_meas, S, win, shift, full;
  float *lags;                 /* [S] */
  int32_t *up_first, *up_n;    /* ArbitraryResample: [S] */
  float *up_w; int32_t up_max; /* [S][up_max] */
} pitch_plan;

static int pitch_plan_init(pitch_plan *p, const orc_pitch_opts *o) { /* pitch-functions.cc:715-766 */
  memset(p, 0, sizeof *p);
  p->o = *o;
  if (!(o->samp_freq > 0 && o->resample_freq > 0 && o->lowpass_cutoff > 0 && o->lowpass_cutoff * 2 <= o->samp_freq &&
        o->lowpass_cutoff * 2 <= o->resample_freq && o->lowpass_filter_width > 0 && o->upsample_filter_width > 0 &&
        o->min_f0 > 0 && o->max_f0 > o->min_f0 && o->delta_pitch > 0))
    return -1;
  lin_resamp_init(&p->lr, (int32_t)o->samp_freq, (int32_t)o->resample_freq, o->lowpass_cutoff, o->lowpass_filter_width);
  double outer_min_lag = 1.0 / o->max_f0 - (o->upsample_filter_width / (2.0 * o->resample_freq));
  double outer_max_lag = 1.0 / o->min_f0 + (o->upsample_filter_width / (2.0 * o->resample_freq));
  p->first_lag = (int32_t)ceil(o->resample_freq * outer_min_lag);
  p->last_lag = (int32_t)floor(o->resample_freq * outer_max_lag);
  p->n_meas = p->last_lag + 1 - p->first_lag;
  p->win = (int32_t)(o->resample_freq * o->frame_length_ms / 1000.0);   /* NccfWindowSize, pitch-functions.h:226-228 */
  p->shift = (int32_t)(o->resample_freq * o->frame_shift_ms / 1000.0);  /* NccfWindowShift */
  p->full = p->win + p->last_lag;
  /* SelectLags, pitch-functions.cc:157-167 */
  float min_lag = (float)(1.0 / o->max_f0), max_lag = (float)(1.0 / o->min_f0);
  int32_t S = 0;
  for (float lag = min_lag; lag <= max_lag; lag = (float)(lag * (1.0 + o->delta_pitch))) S++;
  p->S = S;
  p->lags = (float *)malloc(sizeof(float) * S);
  S = 0;
  for (float lag = min_lag; lag <= max_lag; lag = (float)(lag * (1.0 + o->delta_pitch))) p->lags[S++] = lag;
  /* ArbitraryResample(num_measured_lags, resample_freq, resample_freq/2, lags - first_lag/resample_freq, upsample_filter_width)
   * resample.cc:229-243, 278-309 */
  float samp_rate_in = o->resample_freq, cutoff = (float)(o->resample_freq * 0.5);
  int32_t nz = o->upsample_filter_width;
  float off = -p->first_lag / o->resample_freq;
  float filter_width = (float)(nz / (2.0 * cutoff));
  p->up_first = (int32_t *)malloc(sizeof(int32_t) * S);
  p->up_n = (int32_t *)malloc(sizeof(int32_t) * S);
  p->up_max = 0;
  for (int32_t i = 0; i < S; i++) {
    float t = p->lags[i] + off, t_min = t - filter_width, t_max = t + filter_width;
    int32_t lo = (int32_t)ceil(samp_rate_in * t_min), hi = (int32_t)floor(samp_rate_in * t_max);
    if (lo < 0) lo = 0;
    if (hi >= p->n_meas) hi = p->n_meas - 1;
    p->up_first[i] = lo; p->up_n[i] = hi - lo + 1;
    if (p->up_n[i] > p->up_max) p->up_max = p->up_n[i];
  }
  p->up_w = (float *)calloc((size_t)S * p->up_max, sizeof(float));
  for (int32_t i = 0; i < S; i++) {
    float t = p->lags[i] + off;
    for (int32_t j = 0; j < p->up_n[i]; j++) {
      float delta_t = t - (p->up_first[i] + j) / samp_rate_in;
      p->up_w[(size_t)i * p->up_max + j] = resample_filter(delta_t, cutoff, nz) / samp_rate_in;
    }
  }
  return 0;
}
static void pitch_plan_free(pitch_plan *p) { lin_resamp_free(&p->lr); free(p->lags); free(p->up_first); free(p->up_n); free(p->up_w); }

/* OnlinePitchFeatureImpl::NumFramesAvailable (pitch-functions.cc:768-792) */
static int32_t pitch_frames_available(const pitch_plan *p, int64_t n_down, int finished) {
  int32_t frame_length = p->win;
  if (!finished) frame_length += p->last_lag;
  if (n_down < frame_length) return 0;
  if (!p->o.snip_edges) {
    if (finished) return (int32_t)(n_down * 1.0f / p->shift + 0.5f);
    return (int32_t)((n_down - frame_length / 2) * 1.0f / p->shift + 0.5f);
  }
  return (int32_t)((n_down - frame_length) / p->shift + 1);
}

int32_t orc_pitch_num_frames(const orc_pitch_opts *o, int64_t n_samp) {
  pitch_plan p;
  if (pitch_plan_init(&p, o) != 0) return -1;
  int64_t n2 = lin_resamp_num_out(&p.lr, n_samp, 1);
  int32_t F = pitch_frames_available(&p, n2, 1);
  pitch_plan_free(&p);
  return F;
}

static int approx_equal_f(float a, float b, float tol) { /* base/kaldi-math.h ApproxEqual */
  if (a == b) return 1;
  float diff = fabsf(a - b);
  if (isinf(diff) || diff != diff) return 0;
  return diff <= tol * (fabsf(a) + fabsf(b));
}

int orc_pitch_compute(const orc_pitch_opts *o, const float *wave, int64_t n_samp, float *out, int32_t out_stride) {
  pitch_plan p;
  if (pitch_plan_init(&p, o) != 0) return -1;
  const int32_t S = p.S, M = p.n_meas, W = p.win, full = p.full;
  /* ---- AcceptWaveform(wave) then InputFinished() -> AcceptWaveform(empty) with flush (pitch-functions.cc:1046-1062, 928-931):
   * the down-sampled signal is d[0, n1) from the first call and d[n1, n2) from the flush, which only sees the last
   * R = ceil(in_hz * num_zeros / cutoff) input samples (LinearResample::SetRemainder, resample.cc:171-186). */
  const int64_t n1 = lin_resamp_num_out(&p.lr, n_samp, 0), n2 = lin_resamp_num_out(&p.lr, n_samp, 1);
  const int64_t R = (int64_t)ceilf((float)(p.lr.in_hz * p.lr.num_zeros) / p.lr.cutoff);
  float *d = (float *)malloc(sizeof(float) * (size_t)(n2 > 0 ? n2 : 1));
  for (int64_t i = 0; i < n1; i++) d[i] = lin_resamp_sample(&p.lr, wave, 0, n_samp, i);
  for (int64_t i = n1; i < n2; i++) d[i] = lin_resamp_sample(&p.lr, wave, n_samp - R < 0 ? 0 : n_samp - R, n_samp, i);
  /* signal statistics per call (1052-1058): float VecVec / Sum() results added to doubles */
  double sumsq[2], sum[2];
  int64_t cnt[2];
  {
    double a = 0, b = 0;
    for (int64_t i = 0; i < n1; i++) { a += (double)d[i] * d[i]; b += d[i]; }
    sumsq[0] = (double)(float)a; sum[0] = (double)(float)b; cnt[0] = n1;
    a = 0; b = 0;
    for (int64_t i = n1; i < n2; i++) { a += (double)d[i] * d[i]; b += d[i]; }
    sumsq[1] = sumsq[0] + (double)(float)a; sum[1] = sum[0] + (double)(float)b; cnt[1] = n2;
  }
  const int32_t F1 = pitch_frames_available(&p, n1, 0), F = pitch_frames_available(&p, n2, 1);
  if (F <= 0) { free(d); pitch_plan_free(&p); return 0; }
  float *nccf_pitch = (float *)malloc(sizeof(float) * (size_t)F * S), *nccf_pov = (float *)malloc(sizeof(float) * (size_t)F * S);
  float *avg_norm = (float *)malloc(sizeof(float) * F), *ms_frame = (float *)malloc(sizeof(float) * F);
  float *win = (float *)malloc(sizeof(float) * full), *ip = (float *)malloc(sizeof(float) * M), *np_ = (float *)malloc(sizeof(float) * M);
  float *m_pitch = (float *)malloc(sizeof(float) * M), *m_pov = (float *)malloc(sizeof(float) * M);
  for (int32_t f = 0; f < F; f++) {
    const int call = f < F1 ? 0 : 1;
    const int64_t avail = call == 0 ? n1 : n2;
    int64_t start = o->snip_edges ? (int64_t)f * p.shift : (int64_t)((f + 0.5) * p.shift) - full / 2; /* 1089-1095 */
    /* ExtractFrame (839-901): zero outside [0, avail); pre-emphasis on the copied part only */
    int64_t lo = start < 0 ? -start : 0, hi = start + full > avail ? avail - start : full;
    for (int32_t i = 0; i < full; i++) win[i] = (i >= lo && i < hi) ? d[start + i] : 0.0f;
    if (o->preemph_coeff != 0.0f) {
      for (int64_t i = hi - 1; i > lo; i--) win[i] -= o->preemph_coeff * win[i - 1];
      if (hi > lo) win[lo] = (float)(win[lo] * (1.0 - o->preemph_coeff));
    }
    const double mean_square = sumsq[call] / cnt[call] - pow(sum[call] / cnt[call], 2.0); /* 1110-1111 */
    /* ComputeCorrelation (102-121) */
    { double s = 0; for (int32_t i = 0; i < W; i++) s += win[i]; float mean = -(float)s / W; for (int32_t i = 0; i < full; i++) win[i] += mean; }
    double e1d = 0; for (int32_t i = 0; i < W; i++) e1d += (double)win[i] * win[i];
    const float e1 = (float)e1d;
    for (int32_t lag = p.first_lag; lag <= p.last_lag; lag++) {
      double e2 = 0, s = 0;
      for (int32_t i = 0; i < W; i++) { e2 += (double)win[lag + i] * win[lag + i]; s += (double)win[i] * win[lag + i]; }
      ip[lag - p.first_lag] = (float)s;
      np_[lag - p.first_lag] = e1 * (float)e2;
    }
    const double ballast_pitch = pow(mean_square * W, 2) * o->nccf_ballast; /* 1115-1118 */
    { double s = 0; for (int32_t l = 0; l < M; l++) s += np_[l]; avg_norm[f] = (float)((double)(float)s / M); }
    ms_frame[f] = (float)mean_square;
    /* ComputeNccf (131-150) with ballast (pitch) and without (pov) */
    for (int32_t l = 0; l < M; l++) {
      float den = (float)pow(np_[l] + (float)ballast_pitch, 0.5);
      m_pitch[l] = den != 0.0f ? ip[l] / den : 0.0f;
      den = (float)pow(np_[l] + 0.0f, 0.5);
      m_pov[l] = den != 0.0f ? ip[l] / den : 0.0f;
    }
    /* ArbitraryResample::Resample (245-262) */
    for (int32_t i = 0; i < S; i++) {
      double a = 0, b = 0;
      const float *w = p.up_w + (size_t)i * p.up_max;
      for (int32_t j = 0; j < p.up_n[i]; j++) { a += (double)w[j] * m_pitch[p.up_first[i] + j]; b += (double)w[j] * m_pov[p.up_first[i] + j]; }
      nccf_pitch[(size_t)f * S + i] = (float)a;
      nccf_pov[(size_t)f * S + i] = (float)b;
    }
  }
  /* ---- RecomputeBacktraces (945-1035): runs when the utterance ends before recompute_frame, or at frame
   * recompute_frame - 1; it rescales the first min(F, recompute_frame) frames if any of them saw a mean-square energy more
   * than 1% away from the current one, and restarts the Viterbi from zero.  In the offline path only the first call's frames
   * can differ, and a recompute inside the first call is a no-op, so one Viterbi over the final values reproduces it. */
  if (F1 > 0 && F1 < o->recompute_frame) {
    const double mean = sum[1] / (double)cnt[1];
    const float mean_square = (float)(sumsq[1] / (double)cnt[1] - mean * mean);
    int must = 0;
    const int32_t nre = F < o->recompute_frame ? F : o->recompute_frame;
    for (int32_t f = 0; f < nre; f++) if (!approx_equal_f(ms_frame[f], mean_square, 0.01f)) must = 1;
    if (must) {
      const float new_ballast = (float)(pow(mean_square * W, 2) * o->nccf_ballast);
      for (int32_t f = 0; f < nre; f++) {
        const float old_ballast = (float)(pow(ms_frame[f] * W, 2) * o->nccf_ballast);
        const float scale = powf((old_ballast + avg_norm[f]) / (new_ballast + avg_norm[f]), 0.5f);
        for (int32_t i = 0; i < S; i++) nccf_pitch[(size_t)f * S + i] *= scale;
      }
    }
  }
  /* ---- Viterbi: PitchFrameInfo::ComputeBacktraces, exhaustive form (306-348, 471-473), ComputeLocalCost (178-188) */
  const float delta_pitch_sq = (float)pow(log(1.0 + o->delta_pitch), 2.0), factor = delta_pitch_sq * o->penalty_factor;
  float *fwd = (float *)calloc(S, sizeof(float)), *nxt = (float *)malloc(sizeof(float) * S);
  int32_t *bp = (int32_t *)malloc(sizeof(int32_t) * (size_t)F * S);
  for (int32_t f = 0; f < F; f++) {
    const float *nc = nccf_pitch + (size_t)f * S;
    for (int32_t i = 0; i < S; i++) {
      float best = INFINITY; int32_t bj = -1;
      for (int32_t j = 0; j < S; j++) {
        float c = (j - i) * (j - i) * factor + fwd[j];
        if (c < best) { best = c; bj = j; }
      }
      float local = 1.0f + (-1.0f) * nc[i];
      local = o->soft_min_f0 * p.lags[i] * nc[i] + 1.0f * local;
      nxt[i] = best + local;
      bp[(size_t)f * S + i] = bj;
    }
    float mn = nxt[0];
    for (int32_t i = 1; i < S; i++) if (nxt[i] < mn) mn = nxt[i];
    for (int32_t i = 0; i < S; i++) fwd[i] = nxt[i] - mn; /* 1174-1176 */
  }
  int32_t best = 0;
  for (int32_t i = 1; i < S; i++) if (fwd[i] < fwd[best]) best = i; /* Min(&index): first minimum */
  for (int32_t f = F - 1; f >= 0; f--) { /* SetBestState (486-512), GetFrame (921-926) */
    out[(size_t)f * out_stride + 0] = nccf_pov[(size_t)f * S + best];
    out[(size_t)f * out_stride + 1] = (float)(1.0 / p.lags[best]);
    best = bp[(size_t)f * S + best];
  }
  free(d); free(nccf_pitch); free(nccf_pov); free(avg_norm); free(ms_frame); free(win); free(ip); free(np_);
  free(m_pitch); free(m_pov); free(fwd); free(nxt); free(bp);
  pitch_plan_free(&p);
  return F;
}

/* NccfToPovFeature / NccfToPov (pitch-functions.cc:44-53, 78-87) */
static float nccf_to_pov_feature(float n) {
  if (n > 1.0f) n = 1.0f; else if (n < -1.0f) n = -1.0f;
  return (float)(pow((1.0001 - n), 0.15) - 1.0);
}
static float nccf_to_pov(float n) {
  float ndash = fabsf(n);
  if (ndash > 1.0f) ndash = 1.0f;
  float r = (float)(-5.2 + 5.4 * exp(7.5 * (ndash - 1.0)) + 4.8 * ndash - 2.0 * exp(-10.0 * ndash) +
                    4.2 * exp(20.0 * (ndash - 1.0)));
  return (float)(1.0 / (1 + exp(-1.0 * r)));
}

int orc_process_pitch(const orc_process_pitch_opts *o, const float *in, int32_t T, int32_t in_stride, float *out,
                      int32_t out_stride) {
  if (o->delta_pitch_noise_stddev != 0.0f && o->add_delta_pitch) return -2; /* the reference's noise comes from rand() */
  const int32_t dim = (o->add_pov_feature ? 1 : 0) + (o->add_normalized_log_pitch ? 1 : 0) + (o->add_delta_pitch ? 1 : 0) +
                      (o->add_raw_log_pitch ? 1 : 0);
  if (dim == 0) return -1;
  if (T == 0) return 0;
  const int32_t rows = T + o->delay; /* NumFramesReady with the input finished, 1569-1579 */
  /* delta scales of ComputeDeltas(order 1, window w) (feature-functions.cc:77-110): scales[k + w] = k / sum k^2 */
  const int32_t w = o->delta_window;
  float norm = 0; for (int32_t k = -w; k <= w; k++) norm += (float)(k * k);
  const float inv_norm = (float)(1.0 / norm);
  for (int32_t t = 0; t < rows; t++) {
    const int32_t f = t < o->delay ? 0 : t - o->delay; /* 1416 */
    int32_t idx = 0;
    float *row = out + (size_t)t * out_stride;
    const float nccf = in[(size_t)f * in_stride], log_pitch = logf(in[(size_t)f * in_stride + 1]);
    if (o->add_pov_feature) row[idx++] = o->pov_scale * nccf_to_pov_feature(nccf) + o->pov_offset; /* 1431-1437 */
    if (o->add_normalized_log_pitch) { /* 1476-1485, 1502-1567: pov-weighted mean of log-pitch over the window, in double */
      int32_t b = f - o->normalization_left_context, e = f + o->normalization_right_context + 1;
      if (b < 0) b = 0;
      if (e > T) e = T;
      double sp = 0, slp = 0;
      for (int32_t g = b; g < e; g++) {
        float pov = nccf_to_pov(in[(size_t)g * in_stride]), lp = logf(in[(size_t)g * in_stride + 1]);
        sp += pov; slp += pov * lp;
      }
      float avg = (float)(slp / sp);
      row[idx++] = (log_pitch - avg) * o->pitch_scale;
    }
    if (o->add_delta_pitch) { /* 1439-1466 */
      float dlt = 0;
      for (int32_t k = -w; k <= w; k++) {
        if (k == 0) continue;
        int32_t g = f + k;
        if (g < 0) g = 0;
        if (g > T - 1) g = T - 1;
        dlt += (k * inv_norm) * logf(in[(size_t)g * in_stride + 1]);
      }
      row[idx++] = (dlt + 0.0f) * o->delta_pitch_scale;
    }
    if (o->add_raw_log_pitch) row[idx++] = log_pitch;
  }
  return rows;
}
