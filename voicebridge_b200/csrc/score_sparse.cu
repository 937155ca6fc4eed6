// score_sparse.cu — consumers that need a SUBSET of the log-likelihood matrix.
//
// The dense matrix is 16 kB per frame (P = 4000); shipped to the host it is the PCIe line, not the GPU, that sets the pace
// (round 1: 52 GB/s, 12.5x below the device-resident rate).  Its consumers read far less:
//   * forced alignment (VB/src/gmmbin/gmm-align-compiled.cpp:119-128, decoder/decoder-wrappers.cc AlignUtteranceWrapper)
//     only ever asks for the pdfs of the utterance's own training graph — a few hundred of the P columns;
//   * lattice rescoring (lat/lattice-functions.cc:1214-1360 RescoreCompactLatticeInternal / RescoreLattice) asks for one
//     (frame, pdf) pair per arc.
// Both are served from the dense device matrix (scored in slabs of frames by the tensor-core kernel, device column order)
// by the two kernels below; only the compact result crosses PCIe.
#include <algorithm>

#include "common.h"

namespace {

// Utterance u owns rows [frame_offsets[u], frame_offsets[u+1]) and asks for columns cols[sub_offsets[u] .. sub_offsets[u+1]);
// its block of the output starts at float out_offsets[u] and is [rows x n_u] row-major.  One warp per row of the slab; the
// reads gather inside one 16 kB row (L1/L2 resident: the slab was just written), the writes are contiguous.
__global__ void __launch_bounds__(256) subset_kernel(const float *__restrict__ slab, int32_t slab_stride, int64_t t0, int64_t t1,
                                                     const int32_t *__restrict__ frame2utt,
                                                     const int64_t *__restrict__ frame_offsets,
                                                     const int64_t *__restrict__ sub_offsets, const int32_t *__restrict__ cols,
                                                     const int64_t *__restrict__ out_offsets, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = t0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= t1) return;
  const int u = frame2utt[row];
  const int64_t s0 = sub_offsets[u], n = sub_offsets[u + 1] - s0;
  const float *src = slab + (row - t0) * slab_stride;
  float *dst = out + out_offsets[u] + (row - frame_offsets[u]) * n;
  for (int64_t k = lane; k < n; k += 32) dst[k] = src[__ldg(cols + s0 + k)];
}

// out[i] = slab[frame[i]][col[i]] for the arcs whose frame lies in this slab.
__global__ void __launch_bounds__(256) gather_kernel(const float *__restrict__ slab, int32_t slab_stride, int64_t t0, int64_t t1,
                                                     const int32_t *__restrict__ frames, const int32_t *__restrict__ cols,
                                                     int64_t n, float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = frames[i];
    if (f >= t0 && f < t1) out[i] = slab[(f - t0) * slab_stride + cols[i]];
  }
}

}  // namespace

namespace vb {

int sparse_subset_launch(const float *d_slab, int32_t slab_stride, int64_t t0, int64_t t1, const int32_t *d_frame2utt,
                         const int64_t *d_frame_offsets, const int64_t *d_sub_offsets, const int32_t *d_cols,
                         const int64_t *d_out_offsets, float *d_out, cudaStream_t s) {
  if (t1 <= t0) return 0;
  const int64_t blocks = (t1 - t0 + 7) / 8;
  subset_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_slab, slab_stride, t0, t1, d_frame2utt, d_frame_offsets, d_sub_offsets, d_cols,
                                                  d_out_offsets, d_out);
  VB_CUDA(cudaGetLastError());
  return 0;
}

int sparse_gather_launch(const float *d_slab, int32_t slab_stride, int64_t t0, int64_t t1, const int32_t *d_frames,
                         const int32_t *d_cols, int64_t n, float *d_out, cudaStream_t s) {
  if (t1 <= t0 || n == 0) return 0;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  gather_kernel<<<blocks, 256, 0, s>>>(d_slab, slab_stride, t0, t1, d_frames, d_cols, n, d_out);
  VB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vb
