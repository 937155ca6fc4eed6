// accum.cu — EM sufficient statistics from alignments: AccumAmDiagGmm::AccumulateForGmm[Twofeats]
// (gmm/mle-am-diag-gmm.cc:69-97) -> AccumDiagGmm::AccumulateFromDiag / AccumulateFromPosteriors
// (gmm/mle-diag-gmm.cc:171-204) -> DiagGmm::ComponentPosteriors (gmm/diag-gmm.cc:601-615) + ApplySoftMax
// (matrix/kaldi-vector.cc:852-859).
//
// A warp owns a frame.  Lanes first split the aligned pdf's Gaussians between them (one FP32 dot product each, the
// reference's arithmetic), the softmax is a warp reduction, then lanes switch to the feature dimension and add
// gamma*x, gamma*x^2 (formed in double, as the reference does) into the FP64 accumulators with red.global.add.f64.
// The accumulator is ONE buffer [occ | mean | var | tot_like | tot_frames] so that the cross-GPU reduce is one
// all-reduce.
#include <cfloat>

#include "common.h"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxD = 128;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kWarps * 32) acc_kernel(const float *__restrict__ feats,
                                                          const float *__restrict__ feats2, int64_t T, int32_t stride,
                                                          int32_t D, int32_t DP, const int32_t *__restrict__ pdf_ids,
                                                          const float *__restrict__ weights,
                                                          const float *__restrict__ rows,  // [N][2*DP]
                                                          const float *__restrict__ gconsts,
                                                          const int32_t *__restrict__ pdf_offsets, int32_t P,
                                                          int32_t N, double *__restrict__ acc,
                                                          unsigned long long *bad) {
  __shared__ float s_x[kWarps][2 * kMaxD];   // x | x^2 of the warp's frame (posterior features)
  __shared__ float s_post[kWarps][32];
  __shared__ double s_like[kWarps], s_cnt[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *occ = acc, *mean = acc + N, *var = acc + N + (size_t)N * D;
  double like = 0.0, cnt = 0.0;
  unsigned long long nbad = 0;

  for (int64_t t = (int64_t)blockIdx.x * kWarps + warp; t < T; t += (int64_t)gridDim.x * kWarps) {
    const int p = pdf_ids[t];
    if (p < 0 || p >= P) {  // invalid alignment entry: counted as an error, frame skipped
      if (lane == 0) nbad++;
      continue;
    }
    const float w = weights ? weights[t] : 1.0f;
    const float *xr = feats + t * stride;
    for (int d = lane; d < D; d += 32) {
      const float v = xr[d];
      s_x[warp][d] = v;
      s_x[warp][kMaxD + d] = v * v;
    }
    __syncwarp();
    const int g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
    const float *yr = (feats2 ? feats2 : feats) + t * stride;

    // pass over the Gaussians in chunks of 32 (lane = Gaussian): online max / sum across chunks
    float run_max = -INFINITY, run_sum = 0.0f;
    for (int c0 = 0; c0 < M; c0 += 32) {
      const int m = c0 + lane;
      float ll = -INFINITY;
      if (m < M) {
        const float *r = rows + (size_t)(g0 + m) * (2 * DP);
        float a = 0.0f, b = 0.0f;
        for (int d = 0; d < D; d++) a = fmaf(r[d], s_x[warp][d], a);
        for (int d = 0; d < D; d++) b = fmaf(r[DP + d], s_x[warp][kMaxD + d], b);
        ll = (gconsts[g0 + m] + a) + b;
      }
      const float cmax = warp_max(ll);
      const float nmax = fmaxf(run_max, cmax);
      const float e = (m < M && nmax > -INFINITY) ? __expf(ll - nmax) : 0.0f;
      const float csum = warp_sumf(e);
      run_sum = (run_max > -INFINITY ? run_sum * __expf(run_max - nmax) : 0.0f) + csum;
      run_max = nmax;
    }
    const float log_like = run_max + __logf(run_sum);  // ApplySoftMax returns max + Log(sum)
    if (!(fabsf(log_like) <= FLT_MAX)) {               // diag-gmm.cc:609-610 raises KALDI_ERR
      if (lane == 0) nbad++;
      __syncwarp();
      continue;
    }
    const float inv_sum = 1.0f / run_sum;

    // second pass: posteriors chunk by chunk, then lanes switch to dimensions for the FP64 accumulation
    for (int c0 = 0; c0 < M; c0 += 32) {
      const int m = c0 + lane;
      float post = 0.0f;
      if (m < M) {
        const float *r = rows + (size_t)(g0 + m) * (2 * DP);
        float a = 0.0f, b = 0.0f;
        for (int d = 0; d < D; d++) a = fmaf(r[d], s_x[warp][d], a);
        for (int d = 0; d < D; d++) b = fmaf(r[DP + d], s_x[warp][kMaxD + d], b);
        const float ll = (gconsts[g0 + m] + a) + b;
        post = __expf(ll - run_max) * inv_sum * w;  // Exp(x-max), Scale(1/sum), Scale(frame_posterior)
        atomicAdd(&occ[g0 + m], (double)post);      // occupancy_.AddVec(1.0, post_d)
      }
      s_post[warp][lane] = post;
      __syncwarp();
      const int mc = min(32, M - c0);
      for (int d = lane; d < D; d += 32) {
        const double yd = (double)yr[d], yd2 = yd * yd;  // data_d.ApplyPow(2.0) in double
        for (int k = 0; k < mc; k++) {
          const double g = (double)s_post[warp][k];
          if (g != 0.0) {
            atomicAdd(&mean[(size_t)(g0 + c0 + k) * D + d], g * yd);
            atomicAdd(&var[(size_t)(g0 + c0 + k) * D + d], g * yd2);
          }
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      like += (double)(log_like * w);  // total_log_like_ += log_like * weight (float product)
      cnt += (double)w;
    }
  }
  if (lane == 0) {
    s_like[warp] = like;
    s_cnt[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0, c = 0.0;
    for (int i = 0; i < kWarps; i++) {
      l += s_like[i];
      c += s_cnt[i];
    }
    const size_t tail = (size_t)N + 2 * (size_t)N * D;
    if (c != 0.0 || l != 0.0) {
      atomicAdd(&acc[tail], l);
      atomicAdd(&acc[tail + 1], c);
    }
  }
  if (nbad) atomicAdd(bad, nbad);
}

__global__ void axpy_kernel(double *__restrict__ dst, const double *__restrict__ src, double scale, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] += scale * src[i];
}

}  // namespace

namespace vb {

int acc_launch(vbgpu_acc_t h, const float *d_feats, const float *d_feats2, int64_t T, int32_t stride,
               const int32_t *d_ids, const float *d_w, cudaStream_t s) {
  if (T == 0) return 0;
  vbgpu_gmm_t g = h->model;
  if (g->D > kMaxD) return fail(VBGPU_ERR_INVALID, "feature dim %d > %d", g->D, kMaxD);
  int64_t blocks = (T + kWarps - 1) / kWarps;
  int64_t cap = (int64_t)num_sms(h->device) * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  acc_kernel<<<grid, kWarps * 32, 0, s>>>(d_feats, d_feats2, T, stride, g->D, g->DP, d_ids, d_w, g->d_rows.as<float>(),
                                          g->d_gconsts.as<float>(), g->d_pdf_offsets.as<int32_t>(), g->P, g->N,
                                          h->d_acc.as<double>(), g->d_bad.as<unsigned long long>());
  VB_CUDA(cudaGetLastError());
  return 0;
}

int acc_axpy(double *d_dst, const double *d_src, double scale, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  axpy_kernel<<<grid, 256, 0, s>>>(d_dst, d_src, scale, n);
  VB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vb
