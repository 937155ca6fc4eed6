// featpipe.cu — per-speaker CMVN statistics and the fused feature pipeline
//     apply-cmvn -> add-deltas                      ("delta" systems, decode_gmm.cpp:395-450)
//     apply-cmvn -> splice-feats -> transform-feats ("lda" systems,   decode_gmm.cpp:519-546)
//     [-> per-speaker fMLLR transform-feats]        (decode_gmm.cpp:552-571)
// Replaces AccCmvnStats/ApplyCmvn (transform/cmvn.cc:30-113), ComputeDeltas (feat/feature-functions.cc:88-111,160-171),
// SpliceFrames (:205-226) and the sgemm + AddVecToRows of transform-feats.cpp:95-107.  The reference chains these
// through temp files; here a CTA stages a tile of MFCC rows (+halo) in shared memory once and every later stage
// reads it from there, so HBM sees 64 B in and 160 B out per frame.
#include "common.h"

namespace {

constexpr int kTile = 64;      // output frames per CTA
constexpr int kThreads = 256;

// ---- per-speaker statistics ---------------------------------------------------------------------------------------
// A group of 16 lanes walks a contiguous run of frames; lane d < D accumulates sum and sum of squares for column d
// in double (as AccCmvnStats does, cmvn.cc:30-48), lane D the count; partial sums are flushed with double atomics
// whenever the speaker changes.  stats layout: [spk][2][D+1].
__global__ void __launch_bounds__(kThreads) cmvn_stats_kernel(const float *__restrict__ feats, int32_t stride,
                                                              int32_t D, const int32_t *__restrict__ frame2utt,
                                                              const int32_t *__restrict__ utt2spk, int64_t T,
                                                              int32_t frames_per_group, double *__restrict__ stats) {
  const int lanes = 16;  // D+1 <= 16 handled per pass; larger D loops
  const int group = (blockIdx.x * kThreads + threadIdx.x) / lanes, lane = threadIdx.x % lanes;
  const int64_t t0 = (int64_t)group * frames_per_group;
  if (t0 >= T) return;
  const int64_t t1 = (t0 + frames_per_group < T) ? t0 + frames_per_group : T;
  for (int d0 = 0; d0 < D + 1; d0 += lanes) {
    const int d = d0 + lane;
    double s1 = 0.0, s2 = 0.0;
    int cur = -1;
    for (int64_t t = t0; t < t1; t++) {
      const int u = frame2utt[t];
      const int spk = utt2spk ? utt2spk[u] : u;
      if (spk != cur) {
        if (cur >= 0 && d <= D) {
          atomicAdd(&stats[((size_t)cur * 2 + 0) * (D + 1) + d], s1);
          if (d < D) atomicAdd(&stats[((size_t)cur * 2 + 1) * (D + 1) + d], s2);
        }
        s1 = s2 = 0.0;
        cur = spk;
      }
      if (d < D) {
        const float x = feats[t * stride + d];
        s1 += (double)x;       // *mean_ptr += *feats_ptr * weight
        s2 += (double)(x * x); // *var_ptr += *feats_ptr * *feats_ptr * weight  (float product)
      } else if (d == D) {
        s1 += 1.0;
      }
    }
    if (cur >= 0 && d <= D) {
      atomicAdd(&stats[((size_t)cur * 2 + 0) * (D + 1) + d], s1);
      if (d < D) atomicAdd(&stats[((size_t)cur * 2 + 1) * (D + 1) + d], s2);
    }
  }
}

// stats -> (offset, scale) floats per speaker, exactly as ApplyCmvn derives them (cmvn.cc:84-108).
// norm layout: [spk][2][D] = offsets | scales.  bad[0] counts speakers with count < 1 or non-finite scale.
__global__ void cmvn_norm_kernel(const double *__restrict__ stats, int32_t n_spk, int32_t D, int32_t norm_means,
                                 int32_t norm_vars, float *__restrict__ norm, int32_t *__restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_spk * D) return;
  const int spk = i / D, d = i % D;
  const double *st = stats + (size_t)spk * 2 * (D + 1);
  const double count = st[D];
  double offset = 0.0, scale = 1.0;
  if (norm_means || norm_vars) {
    if (count < 1.0) {
      if (d == 0) atomicAdd(bad, 1);
    } else {
      const double mean = st[d] / count;
      if (!norm_vars) {
        offset = -mean;
      } else {
        double var = st[(D + 1) + d] / count - mean * mean;
        if (var < 1.0e-20) var = 1.0e-20;
        scale = 1.0 / sqrt(var);
        if (scale != scale || 1.0 / scale == 0.0) atomicAdd(bad, 1);
        offset = -(mean * scale);
      }
    }
  }
  norm[((size_t)spk * 2 + 0) * D + d] = (float)offset;
  norm[((size_t)spk * 2 + 1) * D + d] = (float)scale;
}

struct FeatParams {
  const float *in;
  int32_t in_stride, D;  // MFCC dim
  const int64_t *frame_offsets;
  const int32_t *frame2utt, *utt2spk;
  int64_t T;
  const float *norm;  // [spk][2][D] or null (no CMVN)
  int32_t norm_vars;
  int32_t mode, halo;
  int32_t order;
  const float *delta_scales;  // [(order+1)][pitch], pitch = 2*halo+1, centred
  int32_t left, right;
  const float *transform;  // [t_rows][t_cols]
  int32_t t_rows, t_cols, mid_dim, out_dim;
  const float *fmllr;  // [spk][out_dim][fmllr_cols] or null
  int32_t fmllr_cols;
  int32_t tab_size;  // floats of s_tab (the staged fMLLR matrix follows it)
  float *out;
  int32_t out_stride;
};

// One CTA = kTile consecutive frames of the packed batch (tiles may straddle utterances; every access is clamped to
// the owning utterance, feature-functions.cc:100-104,216-218).
__global__ void __launch_bounds__(kThreads) feat_kernel(const FeatParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, rows = kTile + 2 * p.halo;
  const int ms = (p.mid_dim + 4) / 4 * 4;         // row stride of s_mid: mid_dim values, a 1.0 (affine column), zeros
  float *s_x = smem;                              // [rows][D]      normalised MFCC rows, global frame tile0-halo+r
  float *s_mid = s_x + (rows * D + 3) / 4 * 4;    // [kTile][ms]    deltas or spliced+transformed (16-byte aligned rows)
  float *s_tab = s_mid + kTile * ms;              // delta scales, or the global transform
  const int64_t tile0 = (int64_t)blockIdx.x * kTile;
  const int tid = threadIdx.x;

  // ---- stage 0: tables ----
  if (p.mode == 0) {
    const int n = (p.order + 1) * (2 * p.halo + 1);
    for (int i = tid; i < n; i += kThreads) s_tab[i] = p.delta_scales[i];
  } else {
    const int n = p.t_rows * p.t_cols;
    for (int i = tid; i < n; i += kThreads) s_tab[i] = p.transform[i];
  }
  // ---- stage 1: load + CMVN (cmvn.cc:104-108: MulColsVec(scale) then AddVecToRows(offset)) ----
  for (int i = tid; i < rows * D; i += kThreads) {
    const int r = i / D, d = i % D;
    const int64_t t = tile0 - p.halo + r;
    float x = 0.0f;
    if (t >= 0 && t < p.T) {
      x = p.in[t * p.in_stride + d];
      if (p.norm) {
        const int u = p.frame2utt[t];
        const int spk = p.utt2spk ? p.utt2spk[u] : u;
        const float *nm = p.norm + (size_t)spk * 2 * D;
        if (p.norm_vars) x *= nm[D + d];
        x += nm[d];
      }
    }
    s_x[i] = x;
  }
  __syncthreads();

  // ---- stage 2 ----
  if (p.mode == 0) {
    // deltas: out block i of frame f = sum_j scales_i[j] * x[clamp(f+j)]  (saxpy order of DeltaFeatures::Process)
    const int pitch = 2 * p.halo + 1;
    for (int i = tid; i < kTile * D; i += kThreads) {
      const int f = i / D, d = i % D;
      const int64_t t = tile0 + f;
      if (t >= p.T) continue;
      const int u = p.frame2utt[t];
      // rows of s_x that belong to the frame's utterance (the tile holds at most halo frames either side of it)
      const int64_t f0 = p.frame_offsets[u] - tile0 + p.halo, f1 = p.frame_offsets[u + 1] - tile0 + p.halo;
      const int lo = f0 < 0 ? 0 : (int)f0, hi = f1 > rows ? rows - 1 : (int)f1 - 1, r = f + p.halo;
      for (int o = 0; o <= p.order; o++) {
        const int maxoff = o * (p.halo / (p.order > 0 ? p.order : 1));  // window * o
        float acc = 0.0f;
        for (int j = -maxoff; j <= maxoff; j++) {
          const float sc = s_tab[o * pitch + p.halo + j];
          if (sc != 0.0f) acc += sc * s_x[min(max(r + j, lo), hi) * D + d];
        }
        s_mid[f * ms + o * D + d] = acc;
      }
    }
  } else {
    // splice + global transform: y[o] = sum_k M[o][k] * spliced[k] (+ M[o][K] if affine), transform-feats.cpp:95-107
    const int K = D * (p.left + p.right + 1);
    for (int i = tid; i < kTile * p.t_rows; i += kThreads) {
      const int f = i / p.t_rows, o = i % p.t_rows;
      const int64_t t = tile0 + f;
      if (t >= p.T) continue;
      const int u = p.frame2utt[t];
      const int64_t f0 = p.frame_offsets[u], f1 = p.frame_offsets[u + 1];
      const float *m = s_tab + o * p.t_cols;
      float acc = 0.0f;
      for (int j = 0; j <= p.left + p.right; j++) {
        int64_t tt = t + j - p.left;
        tt = tt < f0 ? f0 : (tt >= f1 ? f1 - 1 : tt);
        const float *x = s_x + (int)(tt - tile0 + p.halo) * D;
        for (int d = 0; d < D; d++) acc += m[j * D + d] * x[d];
      }
      if (p.t_cols == K + 1) acc += m[K];
      s_mid[f * ms + o] = acc;
    }
  }
  for (int i = tid; i < kTile * (ms - p.mid_dim); i += kThreads) {  // [mid_dim] = 1 (the affine column's operand), then zeros
    const int f = i / (ms - p.mid_dim), c = i - f * (ms - p.mid_dim);
    s_mid[f * ms + p.mid_dim + c] = c == 0 ? 1.0f : 0.0f;
  }
  __syncthreads();

  // ---- stage 3: optional per-speaker fMLLR, then the coalesced store ----
  const int OD = p.out_dim;
  // The tile's frames almost always belong to one speaker: its matrix is staged TRANSPOSED in shared memory
  // (s_fm[d][o], odd pitch) so that threads o = 0..OD-1 of a frame read consecutive words; a tile that straddles a
  // speaker change reads the matrices from global memory instead.
  float *s_fm = s_tab + p.tab_size;
  const int fm_pitch = OD | 1;
  bool staged = false;
  if (p.fmllr) {
    const int64_t t_last = (tile0 + kTile <= p.T ? tile0 + kTile : p.T) - 1;
    const int u0 = p.frame2utt[tile0], u1 = p.frame2utt[t_last];
    const int spk0 = p.utt2spk ? p.utt2spk[u0] : u0, spk1 = p.utt2spk ? p.utt2spk[u1] : u1;
    staged = (spk0 == spk1) && (u1 - u0 <= 1);  // at most two utterances in the tile: every frame is that speaker's
    if (staged) {
      const float *a = p.fmllr + (size_t)spk0 * OD * p.fmllr_cols;
      for (int i = tid; i < OD * ms; i += kThreads) {  // rows d >= fmllr_cols are zero
        const int o = i / ms, d = i - o * ms;
        s_fm[d * fm_pitch + o] = d < p.fmllr_cols ? a[o * p.fmllr_cols + d] : 0.0f;
      }
    }
    __syncthreads();
  }
  if (staged) {
    // y[f][o] = sum_d A[o][d] x[f][d] (+ A[o][mid_dim] * 1): a thread takes one output column of FOUR frames, so a matrix
    // element is read once per four products and the frames' rows arrive as float4 (same order of additions per output
    // as the one-product-at-a-time form: bit-identical)
    const int OS = p.out_stride, quads = kTile / 4;
    for (int i = tid; i < OS * quads; i += kThreads) {
      const int q = i / OS, o = i - q * OS;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (o < OD) {
        const float *w = s_fm + o;
        const float4 *x0 = reinterpret_cast<const float4 *>(s_mid + (4 * q) * ms);
        for (int d4 = 0; d4 < ms / 4; d4++) {
          const float w0 = w[(4 * d4) * fm_pitch], w1 = w[(4 * d4 + 1) * fm_pitch], w2 = w[(4 * d4 + 2) * fm_pitch],
                      w3 = w[(4 * d4 + 3) * fm_pitch];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float4 x = x0[k * (ms / 4) + d4];
            acc[k] = fmaf(w0, x.x, acc[k]);
            acc[k] = fmaf(w1, x.y, acc[k]);
            acc[k] = fmaf(w2, x.z, acc[k]);
            acc[k] = fmaf(w3, x.w, acc[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int64_t t = tile0 + 4 * q + k;
        if (t < p.T) p.out[t * OS + o] = acc[k];
      }
    }
    return;
  }
  for (int i = tid; i < kTile * p.out_stride; i += kThreads) {
    const int f = i / p.out_stride, o = i % p.out_stride;
    const int64_t t = tile0 + f;
    if (t >= p.T) continue;
    float y = 0.0f;
    if (o < OD) {
      const float *x = s_mid + f * ms;
      if (p.fmllr) {
        float acc = 0.0f;
        {  // (a tile that straddles a speaker change: the matrices come from global memory)
          const int u = p.frame2utt[t];
          const int spk = p.utt2spk ? p.utt2spk[u] : u;
          const float *a = p.fmllr + ((size_t)spk * OD + o) * p.fmllr_cols;
          for (int d = 0; d < p.mid_dim; d++) acc += a[d] * x[d];
          if (p.fmllr_cols == p.mid_dim + 1) acc += a[p.mid_dim];
        }
        y = acc;
      } else {
        y = x[o];
      }
    }
    p.out[t * p.out_stride + o] = y;
  }
}

}  // namespace

namespace vb {

int feat_launch_stats(vbgpu_feat_t h, const float *d_feats, int32_t stride, double *d_stats, int32_t n_spk,
                      cudaStream_t s) {
  const int64_t T = h->layout.total_frames;
  if (T == 0) return 0;
  const int frames_per_group = 256, groups_per_block = kThreads / 16;
  const int64_t groups = (T + frames_per_group - 1) / frames_per_group;
  const int grid = (int)((groups + groups_per_block - 1) / groups_per_block);
  cmvn_stats_kernel<<<grid, kThreads, 0, s>>>(d_feats, stride, h->in_dim, h->layout.d_frame2utt.as<int32_t>(),
                                              h->layout.h_utt2spk.empty() ? nullptr : h->layout.d_utt2spk.as<int32_t>(),
                                              T, frames_per_group, d_stats);
  VB_CUDA(cudaGetLastError());
  return 0;
}

// d_norm: [n_spk][2][D] floats followed by one int32 "bad" counter.
int feat_compute_norm(vbgpu_feat_t h, const double *d_stats, int32_t n_spk, cudaStream_t s) {
  const int D = h->in_dim;
  const size_t nfloats = (size_t)n_spk * 2 * D;
  VB_TRY(h->d_norm.reserve(nfloats * 4 + 16));
  int32_t *bad = reinterpret_cast<int32_t *>(h->d_norm.as<float>() + nfloats);
  VB_CUDA(cudaMemsetAsync(bad, 0, 4, s));
  const int n = n_spk * D;
  cmvn_norm_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_stats, n_spk, D, h->opts.norm_means, h->opts.norm_vars,
                                                   h->d_norm.as<float>(), bad);
  VB_CUDA(cudaGetLastError());
  return 0;
}

int feat_launch(vbgpu_feat_t h, const float *d_in, int32_t in_stride, const float *d_fmllr, int32_t fmllr_cols,
                float *d_out, int32_t out_stride, cudaStream_t s) {
  const int64_t T = h->layout.total_frames;
  if (T == 0) return 0;
  FeatParams p;
  p.in = d_in;
  p.in_stride = in_stride;
  p.D = h->in_dim;
  p.frame_offsets = h->layout.d_frame_offsets.as<int64_t>();
  p.frame2utt = h->layout.d_frame2utt.as<int32_t>();
  p.utt2spk = h->layout.h_utt2spk.empty() ? nullptr : h->layout.d_utt2spk.as<int32_t>();
  p.T = T;
  p.norm = (h->opts.norm_means || h->opts.norm_vars) ? h->d_norm.as<float>() : nullptr;
  p.norm_vars = h->opts.norm_vars;
  p.mode = h->opts.mode;
  p.halo = h->halo;
  p.order = h->opts.delta_order;
  p.delta_scales = h->d_delta_scales.as<float>();
  p.left = h->opts.splice_left;
  p.right = h->opts.splice_right;
  p.transform = h->d_transform.as<float>();
  p.t_rows = h->t_rows;
  p.t_cols = h->t_cols;
  p.mid_dim = h->mid_dim;
  p.out_dim = h->out_dim;
  p.fmllr = d_fmllr;
  p.fmllr_cols = fmllr_cols;
  p.out = d_out;
  p.out_stride = out_stride;
  const int rows = kTile + 2 * h->halo;
  const size_t tab = p.mode == 0 ? (size_t)(p.order + 1) * (2 * p.halo + 1) : (size_t)p.t_rows * p.t_cols;
  p.tab_size = (int32_t)tab;
  const size_t ms = (size_t)(p.mid_dim + 4) / 4 * 4;  // as in the kernel
  const size_t fm = d_fmllr ? (size_t)(p.out_dim | 1) * ms : 0;
  const size_t smem = sizeof(float) * (((size_t)rows * p.D + 3) / 4 * 4 + (size_t)kTile * ms + tab + fm);
  if (smem > 200 * 1024) return fail(VBGPU_ERR_INVALID, "feature pipeline needs %zu bytes of shared memory", smem);
  if (smem > 48 * 1024)
    VB_CUDA(cudaFuncSetAttribute(feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)((T + kTile - 1) / kTile);
  feat_kernel<<<grid, kThreads, smem, s>>>(p);
  VB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vb
