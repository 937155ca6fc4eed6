// score_tc.cu — dense Gaussian scoring on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM) with the
// per-pdf log-sum-exp fused into the epilogue.
//
//   loglikes[t][p] = LogSumExp_{m in pdf p}( gconst_m + means_invvars_m . x_t - 0.5 inv_vars_m . x_t^2 )
//
// which DiagGmm::LogLikelihoods (gmm/diag-gmm.cc:528-562) + VectorBase::LogSumExp (matrix/kaldi-vector.cc:757-775)
// compute per (frame, pdf) behind DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased (gmm/decodable-am-diag-gmm.cc:28-72).
// Here it is ONE contraction  Y[T x N] = A[T x K] . B[K x N]  with  K = 2D+2:
//     A row t    = [ x_0 s1_0, x_0^2 s2_0, x_1 s1_1, x_1^2 s2_1, ..., 1, 1 ]              (x already centred, see below)
//     B column g = [ miv'_0/s1_0, -0.5 iv_0/s2_0, ...,            gconst'_a, gconst'_b ] * log2(e)
// so that Y is the per-Gaussian log-likelihood in log2 units and the epilogue is exp2/log2 only.
//
// Precision.  The tensor cores take 11-bit mantissas; a single pass is ~0.05 abs off.  Both operands are
// therefore split in two fp16 halves (v = hi + lo, 22 bits) and three products are accumulated in the FP32 TMEM tile:
// lo.hi + hi.lo + hi.hi (lo.lo is below 2^-22 relative).  fp16 has the mantissa of TF32 at twice the MMA rate; its
// narrow exponent range is handled by exact power-of-two scales per dimension (s1, s2, chosen from the model) and by
// centring the features on the mean of the model means (c; folded exactly into miv' = miv - iv c and
// gconst' = gconst + miv.c - 0.5 iv.c^2, computed in double; gconst' is carried in three fp16 pieces over the two
// constant-one columns of A).
//
// Layout of the Gaussians (round 2).  The epilogue is the co-bottleneck of this kernel (one MUFU.EX2 per frame and
// Gaussian), so the columns are laid out for IT, not in model order:
//   * pdfs are sorted by size and taken 16 at a time into a GROUP: 16 slots x S rows, slot j / row m at column
//     m*16 + j of the group, every member padded to S Gaussians with dummy columns (score -40000).  The four epilogue
//     warps of a TMEM lane quarter own four adjacent slots each: S tcgen05.ld.x4 loads bring a warp exactly its
//     4 x S values, the log-sum-exp over the rows is straight-line code with compile-time S, and the
//     four results of a frame are one 16-byte store.  No part tables, no carries, no per-pdf branches.
//   * pdfs with 11..20 / 21..40 Gaussians span 2 / 4 adjacent slots (W = 2, 4: 8 / 4 pdfs per group); larger ones are
//     cut into virtual pdfs of <= 40 whose partial results a small merge kernel combines afterwards.
//   * a PANEL (one UMMA N, one TMA bulk copy) is 1..10 groups of the same (S, W), N = 16*S*groups <= 160 columns.
// The output therefore comes out in DEVICE COLUMN ORDER: column col_of_pdf[p] of the matrix holds pdf p (n_cols >= P
// columns: padding members and the extra pieces of cut pdfs take columns too).  Consumers index through that map — a
// decodable already goes through tid2pdf — and score_tc_launch() offers the model's pdf order through a gather kernel.
//
// Kernel shape (persistent, one CTA per SM, 18 warps, warp-specialised):
//   warp 16    producer : cp.async.bulk (TMA engine, 1-D) of pre-tiled B panels [N Gaussians x K] into a 2-stage
//                         shared-memory ring (mbarrier complete_tx).
//   warp 17    MMA      : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=160, K=16) — 3*K/16 per
//                         accumulator — for the CTA's TWO 128-frame tiles in turn (they share every B panel); the
//                         accumulators form a 3-deep ring of 160-column TMEM slots.
//   warps 0-15 epilogue : build the fp16 hi/lo A panels of the CTA's 256 frames once per work unit (straight from the
//                         FP32 features), then per (panel, frame tile): tcgen05.ld the warp's slots, hand the TMEM slot
//                         back as soon as the values are in registers, log-sum-exp, store.
// Work unit = (256-frame tile, range of B panels).  Large batches use one range (all panels); small batches and the
// tiles of the last partial wave split the panels over CTAs to fill the GPU.
//
// Frames outside the fp16 plan (|x - c| beyond ~32x the model's radius) are flagged while the A panel is built and
// re-scored by an FP32 SIMT kernel afterwards, so outliers get the reference's finite answer instead of an error.
#include <cuda_fp16.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.h"

namespace {

constexpr int kRowsMt = 128;   // frames per accumulator (UMMA M)
constexpr int kMt = 2;         // frame tiles per CTA
constexpr int kEpiWarps = 16;  // 4 per TMEM lane quarter
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kNmax = 160;     // columns per panel (UMMA N <= kNmax, multiple of 16) = width of a TMEM slot
constexpr int kSlots = 3;      // TMEM accumulator ring
constexpr int kStages = 2;     // B panel ring in shared memory
constexpr int kSmax = kNmax / 16;      // rows of a group
constexpr int kChunkMax = 4 * kSmax;   // largest (virtual) pdf
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kDummy = -40000.0f;    // log2-domain score of padding columns / zero-weight Gaussians
constexpr double kGcMax = 4096.0;      // |gconst'| (log2 units) beyond which FP32 accumulation cannot hold 1e-3

template <int KS>
struct Cfg {  // KS = 16-wide K steps per split; K = 16*KS >= 2D+2
  static constexpr int kc_half = 2 * KS;       // 16-byte K chunks per split
  static constexpr int kc = 4 * KS;            // hi + lo
  static constexpr int a_bytes = kc * 2048;    // one 128-row A panel: [kc][16 row groups][8 rows x 16 B]
  static constexpr int b_stage = kc * 16 * kNmax;  // largest B panel: [kc][N/8 column groups][8 columns x 16 B]
  static constexpr int off_b = kMt * a_bytes;
  static constexpr int off_bar = off_b + kStages * b_stage;
  static constexpr int smem_bytes = off_bar + 256;
};

// Panel header (one int4 per panel, read through the read-only path one panel ahead):
//   x = byte offset of the panel in the image / 16      y = N | S << 16 | W << 24
//   z = number of groups                                w = first output column of the panel
struct TcParams {
  const float *feats;
  int64_t T;
  int32_t stride, D;
  const uint8_t *bimg;
  const int4 *hdr;
  const float *centre, *s1, *s2;  // [D]
  int32_t n_panels, n_splits;
  int64_t n_units, n_whole;
  float *out;
  int32_t ll_stride, vec_ok;
  unsigned long long *bad;
  uint8_t *rowflag;  // [T]: 1 = re-score this frame in FP32 (outside the fp16 plan)
  uint32_t dbg;      // bring-up only (VBGPU_TC_DEBUG): bit 0 = the epilogue skips loads, math and stores,
                     // bit 1 = no MMAs are issued, bit 3 = no global stores
};

// ---- PTX helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (a CUDA error the host reports), never in a hung GPU.  The
// suspend-time hint lets the hardware park the thread instead of re-issuing the poll (polling warps share their
// sub-partition's issue slots with the epilogue).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (!done && spin > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- TMEM -> registers: N consecutive accumulator columns of this thread's lane (32x32b shape: lane = frame) -------------
__device__ __forceinline__ void tmem_ld1(uint32_t t, float *v) {
  uint32_t r0;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(t) : "memory");
  v[0] = __uint_as_float(r0);
}
__device__ __forceinline__ void tmem_ld2(uint32_t t, float *v) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(t) : "memory");
  v[0] = __uint_as_float(r0), v[1] = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_ld4(uint32_t t, float *v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(t)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t t, float *v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(t)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t t, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(t)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}

// ---- one group: 16 slots x S rows of one accumulator; this warp owns slots 4*cls .. 4*cls+3 ----------------------------
// taddr = TMEM address of (the thread's lane, row 0, slot 4*cls).  S loads of four adjacent columns bring exactly the warp's
// values; as soon as they are in registers the warp may hand the TMEM slot back (release != 0: this was the warp's last
// group of the panel), BEFORE the arithmetic — the MMA warp then only ever waits for loads, not for exponentials.
// W = slots per pdf: res[] receives 4 / W log-likelihoods (natural log).
template <int S, int W>
__device__ __forceinline__ void group_lse(uint32_t taddr, uint32_t release, int lane, float (&res)[4 / W]) {
  float v[S][4];
#pragma unroll
  for (int m = 0; m < S; m++) tmem_ld4(taddr + 16u * m, v[m]);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  // branch-free hand-back: fence + warp sync on every group, a predicated arrive
  tc_fence_before();
  __syncwarp();
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(release),
      "r"((lane == 0 && release != 0u) ? 1u : 0u)
      : "memory");
#pragma unroll
  for (int m = 0; m < S; m++)
#pragma unroll
    for (int j = 0; j < 4; j++) asm volatile("" : "+f"(v[m][j]));  // pin every consumer behind the wait
  if constexpr (S == 1 && W == 1) {
#pragma unroll
    for (int j = 0; j < 4; j++) res[j] = v[0][j] * kLn2;
    return;
  }
  // per-slot maxima (FMNMX3: two new values per instruction)
  float M[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float mx = v[0][j];
#pragma unroll
    for (int m = 1; m + 1 < S; m += 2) mx = max3f(mx, v[m][j], v[m + 1][j]);
    if constexpr ((S & 1) == 0) mx = fmaxf(mx, v[S - 1][j]);
    M[j] = mx;
  }
  if constexpr (W == 2) {
    M[0] = M[1] = fmaxf(M[0], M[1]);
    M[2] = M[3] = fmaxf(M[2], M[3]);
  } else if constexpr (W == 4) {
    M[0] = M[1] = M[2] = M[3] = max3f(fmaxf(M[0], M[1]), M[2], M[3]);
  }
  // sums of 2^(v - M): subtraction and summation as packed FP32 pairs over adjacent slots, one MUFU.EX2 per value
  const float2 nm01 = make_float2(-M[0], -M[1]), nm23 = make_float2(-M[2], -M[3]);
  float2 s01, s23;
#pragma unroll
  for (int m = 0; m < S; m++) {
    const float2 d01 = __fadd2_rn(make_float2(v[m][0], v[m][1]), nm01);
    const float2 d23 = __fadd2_rn(make_float2(v[m][2], v[m][3]), nm23);
    const float2 e01 = make_float2(ex2f(d01.x), ex2f(d01.y));
    const float2 e23 = make_float2(ex2f(d23.x), ex2f(d23.y));
    s01 = (m == 0) ? e01 : __fadd2_rn(s01, e01);
    s23 = (m == 0) ? e23 : __fadd2_rn(s23, e23);
  }
  if constexpr (W == 1) {
    res[0] = (M[0] + lg2f(s01.x)) * kLn2;
    res[1] = (M[1] + lg2f(s01.y)) * kLn2;
    res[2] = (M[2] + lg2f(s23.x)) * kLn2;
    res[3] = (M[3] + lg2f(s23.y)) * kLn2;
  } else if constexpr (W == 2) {
    res[0] = (M[0] + lg2f(s01.x + s01.y)) * kLn2;
    res[1] = (M[2] + lg2f(s23.x + s23.y)) * kLn2;
  } else {
    res[0] = (M[0] + lg2f((s01.x + s01.y) + (s23.x + s23.y))) * kLn2;
  }
}

// All groups of one (panel, frame tile) for this warp.  orow = &out[frame][first column of the panel + cls * 4 / W].
template <int S, int W>
__device__ __forceinline__ void run_groups(uint32_t taddr, int ng, uint32_t rel_bar, int lane, float *orow, bool live,
                                           bool vec, bool no_store) {
#pragma unroll 1
  for (int g = 0; g < ng; g++) {
    float res[4 / W];
    group_lse<S, W>(taddr + (uint32_t)(g * 16 * S), (g + 1 == ng) ? rel_bar : 0u, lane, res);
    float *o = orow + g * (16 / W);
    if (live && !no_store) {
      if constexpr (W == 1) {
        if (vec) {
          *reinterpret_cast<float4 *>(o) = make_float4(res[0], res[1], res[2], res[3]);
        } else {
          o[0] = res[0], o[1] = res[1], o[2] = res[2], o[3] = res[3];
        }
      } else if constexpr (W == 2) {
        if (vec) {
          *reinterpret_cast<float2 *>(o) = make_float2(res[0], res[1]);
        } else {
          o[0] = res[0], o[1] = res[1];
        }
      } else {
        o[0] = res[0];
      }
    }
  }
}

// Shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO))):
// a core matrix is 8 rows x 16 bytes stored contiguously (128 B); LBO = byte distance between the two core matrices
// of one K=16 step, SBO = byte distance between 8-row groups.  Bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1 (sm_100), [61,64) layout = 0 (no swizzle).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// Instruction descriptor for kind::f16: D = F32 (bit 4), A = B = F16 (0), both K-major (0), N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((uint32_t)(kRowsMt >> 4) << 24);
}

struct UnitRange {
  int64_t mtile;
  int32_t t0, t1;
};
// Units 0 .. n_whole-1 are whole frame tiles (all panels); the frame tiles after them are each cut into n_splits units by
// panel range.  Large batches: whole tiles fill the full waves and only the tiles of the last, partial wave are cut, so that
// the tail keeps every SM busy.  Small batches: n_whole = 0, every tile is cut.
__device__ __forceinline__ UnitRange unit_range(const TcParams &p, int64_t u) {
  UnitRange r;
  if (u < p.n_whole) {
    r.mtile = u;
    r.t0 = 0;
    r.t1 = p.n_panels;
    return r;
  }
  u -= p.n_whole;
  r.mtile = p.n_whole + u / p.n_splits;
  const int split = (int)(u - (r.mtile - p.n_whole) * p.n_splits);
  r.t0 = (int)(((int64_t)split * p.n_panels) / p.n_splits);
  r.t1 = (int)(((int64_t)(split + 1) * p.n_panels) / p.n_splits);
  return r;
}

// barrier slots
enum { kBarFull = 0, kBarEmpty = kBarFull + kStages, kBarAccFull = kBarEmpty + kStages, kBarAccEmpty = kBarAccFull + kSlots,
       kBarAReady = kBarAccEmpty + kSlots, kNumBars };
static_assert(kNumBars * 8 <= 128, "barrier block");

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) score_tc_kernel(const TcParams p) {
  using C = Cfg<KS>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::off_bar);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::off_bar + 128);
  uint32_t dbg;
  asm volatile("mov.u32 %0, %1;" : "=r"(dbg) : "r"(p.dbg));

  if (threadIdx.x == kEpiWarps * 32) {
    for (int i = 0; i < kStages; i++) mbar_init(BAR(kBarFull + i), 1);
    for (int i = 0; i < kStages; i++) mbar_init(BAR(kBarEmpty + i), 1);
    for (int i = 0; i < kSlots; i++) mbar_init(BAR(kBarAccFull + i), 1);
    for (int i = 0; i < kSlots; i++) mbar_init(BAR(kBarAccEmpty + i), kEpiWarps);
    mbar_init(BAR(kBarAReady), kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + 1) {  // TMEM: all 512 columns (the CTA owns the SM: > 180 KB of shared memory)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

  if (warp == kEpiWarps) {
    // ================================================= producer =================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const UnitRange ur = unit_range(p, u);
        int4 hn = __ldg(p.hdr + ur.t0);
        for (int t = ur.t0; t < ur.t1; t++, it++) {
          const int4 h = hn;
          if (t + 1 < ur.t1) hn = __ldg(p.hdr + t + 1);
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          const uint32_t bytes = (uint32_t)(h.y & 0xffff) * (uint32_t)(C::kc * 16);
          mbar_wait(BAR(kBarEmpty + s), ph ^ 1);
          mbar_expect_tx(BAR(kBarFull + s), bytes);
          bulk_g2s(smem_u32(smem + C::off_b + s * C::b_stage), p.bimg + (size_t)(uint32_t)h.x * 16, bytes,
                   BAR(kBarFull + s));
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ================================================= MMA issuer ===============================================
    // The whole warp walks the loop converged (every value below is warp-uniform, so descriptors and addresses are
    // formed in uniform registers); one elected lane issues the tensor-core instructions and the commits.
    uint32_t itp = 0, iti = 0, un = 0;
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + C::off_b);
    for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x, un++) {
      const UnitRange ur = unit_range(p, u);
      int4 hn = __ldg(p.hdr + ur.t0);
      mbar_wait(BAR(kBarAReady), un & 1);  // the A panels of this unit are in shared memory
      for (int t = ur.t0; t < ur.t1; t++, itp++) {
        const int4 h = hn;
        if (t + 1 < ur.t1) hn = __ldg(p.hdr + t + 1);
        const uint32_t n = (uint32_t)(h.y & 0xffff);
        const uint32_t s = itp % kStages, ph = (itp / kStages) & 1;
        const uint32_t idesc = make_idesc(n);
        const uint32_t kstep = 2u * n;               // one K=16 step = two chunks of n*16 bytes, in 16-byte units
        const uint32_t lo_off = C::kc_half * n;      // the lo half of the panel, in 16-byte units
        mbar_wait(BAR(kBarFull + s), ph);
        const uint64_t bdesc = make_desc(b_base + s * C::b_stage, n * 16u, 128);
#pragma unroll 1
        for (int mt = 0; mt < kMt; mt++, iti++) {
          const uint32_t slot = iti % kSlots, sph = (iti / kSlots) & 1;
          mbar_wait(BAR(kBarAccEmpty + slot), sph ^ 1);  // every epilogue warp has read this slot's previous tile
          tc_fence_after();
          if (elect_one()) {
            if (!(dbg & 2u)) {
              const uint32_t d = tmem_base + slot * (uint32_t)kNmax;
              const uint64_t adesc = make_desc(a_base + mt * C::a_bytes, 2048, 128);
#pragma unroll
              for (int prod = 0; prod < 3; prod++) {  // lo.hi, hi.lo, hi.hi (small terms first)
                const uint32_t ao = (prod == 0) ? (uint32_t)(C::kc_half * 128) : 0u, bo = (prod == 1) ? lo_off : 0u;
#pragma unroll
                for (int k = 0; k < KS; k++)
                  tc_mma_f16(d, adesc + (ao + k * 256u), bdesc + (bo + k * kstep), idesc, (prod | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(BAR(kBarAccFull + slot));
            if (mt == kMt - 1) tc_commit(BAR(kBarEmpty + s));  // the B panel (and every MMA before it) is done
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================= epilogue =================================================
    // Warp w serves TMEM lanes 32*(w&3)..+31, i.e. frame (w&3)*32+lane of each frame tile, and slots 4*(w>>2)..+3 of
    // every group.
    const int q = warp & 3, cls = warp >> 2;
    uint32_t iti = 0;
    // Non-finite results can only come from non-finite features: the model image is validated on the host, the operands
    // are bounded and every sum of exponentials is >= 1.  They are counted where the features are read.
    unsigned long long nbad = 0;
    uint32_t vec_in;  // read through an opaque move: keeps the compiler from cloning the loops per loop-invariant flag
    asm volatile("mov.u32 %0, %1;" : "=r"(vec_in) : "r"(p.vec_ok));
    const bool vec = (vec_in & 1u) != 0, no_store = (dbg & 8u) != 0;
#pragma unroll 1
    for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const UnitRange ur = unit_range(p, u);
      int4 hn = __ldg(p.hdr + ur.t0);
      // ---- A panel: thread -> (row = tid & 255, K half = tid >> 8) -> fp16 hi/lo, K-major core-matrix layout.  The
      //      previous unit's MMAs completed before its last accumulator was published (tcgen05.commit covers all
      //      earlier MMAs), and every epilogue warp has waited for that accumulator.
      {
        const int row = threadIdx.x & 255, kh = threadIdx.x >> 8, mt = row >> 7, rowl = row & 127;
        const int64_t trow = ur.mtile * (kMt * kRowsMt) + row;
        const int d0 = kh * 4 * KS;  // this thread's dims: d0 .. d0 + 4*KS - 1  (KS chunks of 4 dims)
        float x[4 * KS];
        const float *xr = p.feats + trow * p.stride;
#pragma unroll
        for (int d = 0; d < 4 * KS; d++) x[d] = 0.0f;
        if (trow < p.T) {
          if (vec_in & 2) {  // rows are 16-byte aligned and the stride covers the padded row
#pragma unroll
            for (int d4 = 0; d4 < KS; d4++)
              if (d0 + d4 * 4 < p.D) {
                const float4 v = *reinterpret_cast<const float4 *>(xr + d0 + d4 * 4);
                x[d4 * 4 + 0] = v.x;
                x[d4 * 4 + 1] = v.y;
                x[d4 * 4 + 2] = v.z;
                x[d4 * 4 + 3] = v.w;
              }
          } else {
#pragma unroll
            for (int d = 0; d < 4 * KS; d++)
              if (d0 + d < p.D) x[d] = xr[d0 + d];
          }
        }
        bool outlier = false;
        uint8_t *arow = smem + mt * C::a_bytes + (rowl >> 3) * 128 + (rowl & 7) * 16 + kh * KS * 2048;
#pragma unroll
        for (int kc = 0; kc < KS; kc++) {
          __half2 hi[4], lo[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; e2++) {
            const int d = d0 + kc * 4 + e2;  // K index 2d -> x_d * s1_d, 2d+1 -> x_d^2 * s2_d, 2D and 2D+1 -> 1
            float v0 = 0.0f, v1 = 0.0f;
            if (d < p.D) {
              const float xv = x[kc * 4 + e2];
              const float xc = xv - __ldg(p.centre + d);
              v0 = xc * __ldg(p.s1 + d);
              v1 = (xc * xc) * __ldg(p.s2 + d);
              if (!(fabsf(xv) <= FLT_MAX)) {  // NaN / Inf feature: the reference fails on the frame's log-likelihoods
                nbad++;
                v0 = v1 = 0.0f;
              } else if (!(fabsf(v0) <= 65504.0f) || !(fabsf(v1) <= 65504.0f)) {  // outside the fp16 plan
                outlier = true;
                v0 = fminf(fmaxf(v0, -65504.0f), 65504.0f);
                v1 = fminf(v1, 65504.0f);
              }
            } else if (d == p.D) {
              v0 = 1.0f, v1 = 1.0f;
            }
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            hi[e2] = __halves2half2(h0, h1);
            lo[e2] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
          }
          *reinterpret_cast<uint4 *>(arow + kc * 2048) = *reinterpret_cast<uint4 *>(hi);
          *reinterpret_cast<uint4 *>(arow + (kc + C::kc_half) * 2048) = *reinterpret_cast<uint4 *>(lo);
        }
        if (outlier && trow < p.T) p.rowflag[trow] = 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(kBarAReady));
      }

      // ---- panels ----
      const int64_t trow0 = ur.mtile * (kMt * kRowsMt) + q * 32 + lane;
#pragma unroll 1
      for (int t = ur.t0; t < ur.t1; t++) {
        const int4 h = hn;
        if (t + 1 < ur.t1) hn = __ldg(p.hdr + t + 1);
        const int S = (h.y >> 16) & 0xff, W = (h.y >> 24) & 0xff, ng = h.z;
        const int key = (S - 1) + (W == 1 ? 0 : W == 2 ? kSmax : 2 * kSmax);
        const int per = (W == 1) ? 4 : (W == 2) ? 2 : 1;  // output columns of this warp per group
#pragma unroll 1
        for (int mt = 0; mt < kMt; mt++, iti++) {
          const uint32_t slot = iti % kSlots, sph = (iti / kSlots) & 1;
          const int64_t trow = trow0 + mt * kRowsMt;
          const bool live = trow < p.T;
          float *orow = p.out + trow * p.ll_stride + h.w + cls * per;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)kNmax + 4u * cls;
          const uint32_t rel = BAR(kBarAccEmpty + slot);
          mbar_wait(BAR(kBarAccFull + slot), sph);
          tc_fence_after();
          if (dbg & 1u) {  // bring-up: hand the slot straight back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(rel);
            continue;
          }
          switch (key) {
#define VB_CASE(S_, W_, K_) \
  case K_: run_groups<S_, W_>(taddr, ng, rel, lane, orow, live, vec, no_store); break;
#define VB_CASES(W_, B_)                                                                                               \
  VB_CASE(1, W_, B_ + 0) VB_CASE(2, W_, B_ + 1) VB_CASE(3, W_, B_ + 2) VB_CASE(4, W_, B_ + 3) VB_CASE(5, W_, B_ + 4)     \
  VB_CASE(6, W_, B_ + 5) VB_CASE(7, W_, B_ + 6) VB_CASE(8, W_, B_ + 7) VB_CASE(9, W_, B_ + 8) VB_CASE(10, W_, B_ + 9)
            VB_CASES(1, 0)
            VB_CASES(2, kSmax)
            VB_CASES(4, 2 * kSmax)
#undef VB_CASES
#undef VB_CASE
            default: __trap();
          }
        }
      }
      __syncwarp();
    }
    if (nbad) atomicAdd(p.bad, nbad);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
static_assert(kSmax == 10, "the dispatch table above lists S = 1..10");

// ---- small companions of the tensor-core kernel -----------------------------------------------------------------------
// Frames flagged by the A-panel builder (outside the fp16 plan) are re-scored here with the FP32 arithmetic of
// score_simt.cu (the parity anchor): warp = frame, lanes stride over the pdfs.  Launched after every tensor-core launch;
// without flagged frames it reads T bytes and exits.
__global__ void __launch_bounds__(256) score_fix_kernel(const uint8_t *__restrict__ rowflag, const float *__restrict__ feats,
                                                        int64_t T, int32_t stride, int32_t D, int32_t DP,
                                                        const float *__restrict__ rows, const float *__restrict__ gconsts,
                                                        const int32_t *__restrict__ pdf_offsets,
                                                        const int32_t *__restrict__ col_of_pdf, int32_t P,
                                                        float *__restrict__ out, int32_t out_stride,
                                                        unsigned long long *bad) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  unsigned long long nbad = 0;
  for (int64_t c = warp; c * 32 < T; c += n_warps) {
    const int64_t t = c * 32 + lane;
    unsigned mask = __ballot_sync(0xffffffffu, t < T && rowflag[t] != 0);
    while (mask) {
      const int r = __ffs(mask) - 1;
      mask &= mask - 1;
      const float *x = feats + (c * 32 + r) * stride;
      for (int pdf = lane; pdf < P; pdf += 32) {
        const int g0 = pdf_offsets[pdf], g1 = pdf_offsets[pdf + 1];
        float mx = -INFINITY, sum = 0.0f;
        for (int g = g0; g < g1; g++) {
          const float *row = rows + (size_t)g * 2 * DP;
          float a = 0.0f, b = 0.0f;
          for (int d = 0; d < D; d++) a = fmaf(__ldg(row + d), x[d], a);
          for (int d = 0; d < D; d++) b = fmaf(__ldg(row + DP + d), x[d] * x[d], b);
          const float ll = (__ldg(gconsts + g) + a) + b;
          if (ll > mx) {
            sum = sum * __expf(mx - ll) + 1.0f;
            mx = ll;
          } else if (ll > -INFINITY) {
            sum += __expf(ll - mx);
          }
        }
        const float res = (g1 > g0) ? mx + __logf(sum) : -INFINITY;
        if (!(fabsf(res) <= FLT_MAX)) nbad++;
        out[(c * 32 + r) * out_stride + col_of_pdf[pdf]] = res;
      }
    }
  }
  if (nbad) atomicAdd(bad, nbad);
}

// A pdf cut into virtual pdfs: out[t][main] = LogSumExp(out[t][main], out[t][extra_0], ...).  merge = [n][2] (main, extra),
// the entries of one pdf adjacent; a thread handles one (frame, entry run).
__global__ void __launch_bounds__(256) score_merge_kernel(float *__restrict__ out, int64_t T, int32_t out_stride,
                                                          const int32_t *__restrict__ merge, int32_t n_merge) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float *row = out + t * out_stride;
  int i = 0;
  while (i < n_merge) {
    const int main_col = merge[2 * i];
    float mx = row[main_col], sum = 1.0f;
    for (; i < n_merge && merge[2 * i] == main_col; i++) {
      const float v = row[merge[2 * i + 1]];
      if (v > mx) {
        sum = sum * __expf(mx - v) + 1.0f;
        mx = v;
      } else {
        sum += __expf(v - mx);
      }
    }
    row[main_col] = mx + __logf(sum);
  }
}

// Device column order -> the model's pdf order: dst[t][p] = src[t][col_of_pdf[p]].  One CTA per 8 frames; the reads
// gather inside a row that the CTA has in L1, the writes are coalesced.
__global__ void __launch_bounds__(256) score_gather_pdf_kernel(const float *__restrict__ src, int32_t src_stride,
                                                               const int32_t *__restrict__ col_of_pdf, int32_t P, int64_t T,
                                                               float *__restrict__ dst, int32_t dst_stride) {
  const int64_t t0 = (int64_t)blockIdx.x * 8;
  for (int r = 0; r < 8 && t0 + r < T; r++) {
    const float *s = src + (t0 + r) * src_stride;
    float *d = dst + (t0 + r) * dst_stride;
    for (int pdf = threadIdx.x; pdf < P; pdf += blockDim.x) d[pdf] = s[__ldg(col_of_pdf + pdf)];
  }
}

// ---- host-side state ------------------------------------------------------------------------------------------------
struct GaussPos {
  uint32_t panel;  // index into the panel tables
  uint16_t col;    // column inside the panel
};
struct TcState {
  int KS = 0, n_panels = 0, n_cols = 0, n_merge = 0;
  std::vector<uint8_t> h_bimg;        // kept for gconst updates
  std::vector<uint64_t> panel_off;    // byte offset of every panel in the image
  std::vector<uint16_t> panel_n;      // columns of every panel
  std::vector<GaussPos> gpos;         // where every Gaussian sits
  std::vector<double> gshift;         // gconst' - gconst  (centring term), per Gaussian
  std::vector<int32_t> col_of_pdf;    // output column of every pdf
  vb::DevBuf d_bimg, d_hdr, d_centre, d_s1, d_s2, d_col_of_pdf, d_merge, d_rowflag, d_scratch;
  bool attr_set = false;
};

// Byte offset of element (column n, K index k) of the hi half inside a panel of N columns; the lo half follows
// 2*KS chunks later.
inline size_t elem_off(int N, int n, int k) { return ((size_t)(k / 8) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; }

inline void put_split(uint8_t *panel, int N, int KS, int n, int k, double v) {
  const __half hi = __double2half(v);
  const __half lo = __double2half(v - (double)__half2float(hi));
  const size_t off = elem_off(N, n, k);
  std::memcpy(panel + off, &hi, 2);
  std::memcpy(panel + off + (size_t)2 * KS * 16 * N, &lo, 2);
}
// gconst' in three fp16 pieces: hi + lo at K index 2D, the remainder at K index 2D+1 (hi half only).
inline void put_gconst(uint8_t *panel, int N, int KS, int n, int D, double gc) {
  const __half hi = __double2half(gc);
  const double r1 = gc - (double)__half2float(hi);
  const __half lo = __double2half(r1);
  const double r2 = r1 - (double)__half2float(lo);
  const __half third = __double2half(r2), zero = __double2half(0.0);
  const size_t off = elem_off(N, n, 2 * D), off2 = elem_off(N, n, 2 * D + 1), lo_half = (size_t)2 * KS * 16 * N;
  std::memcpy(panel + off, &hi, 2);
  std::memcpy(panel + off + lo_half, &lo, 2);
  std::memcpy(panel + off2, &third, 2);
  std::memcpy(panel + off2 + lo_half, &zero, 2);
}

struct TcHostImage {
  std::vector<int4> hdr;
  std::vector<int32_t> merge;
  std::vector<float> centre, s1, s2;
};

// The layout of a model for the tensor-core kernel (pure host code).  Returns null, or the reason the model is off the plan.
const char *build_layout(int D, int N, int P, const std::vector<int32_t> &po, const float *gconsts, const float *miv,
                         const float *iv, int32_t stride, TcState *st, TcHostImage *img) {
  int KS = (2 * D + 2 + 15) / 16;
  if (KS < 2) KS = 2;
  if (KS > 6) return "feature dimension above 47";
  for (int pdf = 0; pdf < P; pdf++) {
    bool any = false;
    for (int g = po[pdf]; g < po[pdf + 1]; g++) any |= (gconsts[g] > -INFINITY);
    if (!any) return "a pdf has no Gaussian with a finite gconst";
  }
  // centre and scales from the model: means mu = miv/iv, sigma = iv^-1/2
  std::vector<double> c(D, 0.0), R(D, 0.0);
  for (int g = 0; g < N; g++)
    for (int d = 0; d < D; d++) {
      const double v = iv[(size_t)g * stride + d];
      if (!(v > 0.0) || !std::isfinite(v) || !std::isfinite((double)miv[(size_t)g * stride + d])) return "non-positive or non-finite variance / mean";
      c[d] += (double)miv[(size_t)g * stride + d] / v;
    }
  std::vector<float> cf(D), s1(D), s2(D);
  for (int d = 0; d < D; d++) cf[d] = (float)(c[d] / N), c[d] = (double)cf[d];
  for (int g = 0; g < N; g++)
    for (int d = 0; d < D; d++) {
      const double v = iv[(size_t)g * stride + d], mu = (double)miv[(size_t)g * stride + d] / v;
      R[d] = std::max(R[d], std::fabs(mu - c[d]) + 3.0 / std::sqrt(v));
    }
  for (int d = 0; d < D; d++) {
    const int e = (R[d] > 0.0 && std::isfinite(R[d])) ? (int)std::ceil(std::log2(R[d])) : 1000;
    if (e < -40 || e > 40) return "model radius outside the fp16 scaling plan";
    s1[d] = (float)std::ldexp(1.0, 5 - e);            // |x - c| <= R  ->  |a| <= 32
    s2[d] = (float)std::ldexp(1.0, 2 * (5 - e) - 4);  // (x - c)^2 s2 <= 64
  }

  // ---- virtual pdfs: a pdf above kChunkMax Gaussians is cut into near-equal pieces ----
  struct VPdf {
    int pdf, g0, size, piece;
  };
  std::vector<VPdf> vp;
  for (int pdf = 0; pdf < P; pdf++) {
    const int M = po[pdf + 1] - po[pdf];
    const int k = (M + kChunkMax - 1) / kChunkMax;
    int g = po[pdf];
    for (int i = 0; i < k; i++) {
      const int sz = M / k + (i < M % k ? 1 : 0);
      vp.push_back({pdf, g, sz, i});
      g += sz;
    }
  }
  // ---- groups: by slots per pdf (W = 1, 2, 4), then by size; 16 / W members per group ----
  struct Group {
    int S, W;
    std::vector<int> members;  // indices into vp
  };
  std::vector<Group> groups;
  for (int W : {1, 2, 4}) {
    std::vector<int> idx;
    for (int i = 0; i < (int)vp.size(); i++) {
      const int sz = vp[i].size, w = sz <= kSmax ? 1 : sz <= 2 * kSmax ? 2 : 4;
      if (w == W) idx.push_back(i);
    }
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return vp[a].size < vp[b].size; });
    const int per = 16 / W;
    for (size_t i = 0; i < idx.size(); i += per) {
      Group gr;
      gr.W = W;
      gr.members.assign(idx.begin() + i, idx.begin() + std::min(idx.size(), i + per));
      gr.S = (vp[gr.members.back()].size + W - 1) / W;
      groups.push_back(gr);
    }
  }
  // ---- panels: runs of groups of the same (S, W), at most kNmax columns ----
  st->KS = KS;
  st->gpos.resize(N);
  st->gshift.assign(N, 0.0);
  st->col_of_pdf.assign(P, -1);
  std::vector<int4> &hdr = img->hdr;
  std::vector<int32_t> &merge = img->merge;  // (main column, extra column)
  std::vector<int32_t> vcol(vp.size(), -1);
  std::vector<std::pair<int, int>> panel_groups;  // [first group, count)
  {
    size_t gi = 0;
    int out_col = 0;
    uint64_t off = 0;
    while (gi < groups.size()) {
      const int S = groups[gi].S, W = groups[gi].W, fit = std::max(1, kNmax / (16 * S));
      int ng = 1;
      while (ng < fit && gi + ng < groups.size() && groups[gi + ng].S == S && groups[gi + ng].W == W) ng++;
      const int Np = 16 * S * ng;
      st->panel_off.push_back(off);
      st->panel_n.push_back((uint16_t)Np);
      hdr.push_back(make_int4((int)(off / 16), Np | (S << 16) | (W << 24), ng, out_col));
      panel_groups.push_back({(int)gi, ng});
      for (int g = 0; g < ng; g++)
        for (size_t j = 0; j < groups[gi + g].members.size(); j++) vcol[groups[gi + g].members[j]] = out_col + g * (16 / W) + (int)j;
      out_col += ng * (16 / W);
      off += (uint64_t)4 * KS * 16 * Np;
      gi += ng;
    }
    st->n_panels = (int)hdr.size();
    st->n_cols = (out_col + 3) / 4 * 4;
    st->h_bimg.assign(off, 0);
    hdr.push_back(make_int4(0, 16 | (1 << 16) | (1 << 24), 1, 0));  // padding entry: the kernel may prefetch one past the end
  }
  for (size_t i = 0; i < vp.size(); i++) {
    if (vp[i].piece == 0) {
      st->col_of_pdf[vp[i].pdf] = vcol[i];
    }
  }
  for (size_t i = 0; i < vp.size(); i++)
    if (vp[i].piece > 0) {
      merge.push_back(st->col_of_pdf[vp[i].pdf]);
      merge.push_back(vcol[i]);
    }
  st->n_merge = (int)merge.size() / 2;
  // ---- B image ----
  const double L2E = 1.4426950408889634074;
  const char *why = nullptr;
  for (int pi = 0; pi < st->n_panels && !why; pi++) {
    const int Np = st->panel_n[pi];
    uint8_t *panel = st->h_bimg.data() + st->panel_off[pi];
    std::vector<int> gauss_of_col(Np, -1);
    for (int g = 0; g < panel_groups[pi].second; g++) {
      const Group &gr = groups[panel_groups[pi].first + g];
      for (size_t j = 0; j < gr.members.size(); j++) {
        const VPdf &v = vp[gr.members[j]];
        for (int k = 0; k < v.size; k++) gauss_of_col[g * 16 * gr.S + (k / gr.W) * 16 + gr.W * (int)j + k % gr.W] = v.g0 + k;
      }
    }
    for (int n = 0; n < Np && !why; n++) {
      const int g = gauss_of_col[n];
      double gc = kDummy;
      if (g >= 0) {
        st->gpos[g] = {(uint32_t)pi, (uint16_t)n};
        double shift = 0.0;
        for (int d = 0; d < D; d++) {
          const double v = iv[(size_t)g * stride + d], mv = miv[(size_t)g * stride + d];
          const double b1 = (mv - v * c[d]) * L2E / (double)s1[d], b2 = -0.5 * v * L2E / (double)s2[d];
          if (std::fabs(b1) > 60000.0 || std::fabs(b2) > 60000.0) why = "model parameters outside the fp16 range";
          put_split(panel, Np, KS, n, 2 * d, b1);
          put_split(panel, Np, KS, n, 2 * d + 1, b2);
          shift += mv * c[d] - 0.5 * v * c[d] * c[d];
        }
        st->gshift[g] = shift;
        if (gconsts[g] > -INFINITY) {  // a zero-weight Gaussian keeps the dummy score
          gc = ((double)gconsts[g] + shift) * L2E;
          if (!(std::fabs(gc) <= kGcMax)) why = "a centred gconst is too large for FP32 accumulation at 1e-3";
        }
      }
      put_gconst(panel, Np, KS, n, D, gc);
    }
  }
  img->centre = cf, img->s1 = s1, img->s2 = s2;
  return why;
}

template <int KS>
int launch_ks(const TcParams &p, TcState *st, int grid, cudaStream_t s) {
  using C = Cfg<KS>;
  if (!st->attr_set) {
    VB_CUDA(cudaFuncSetAttribute(score_tc_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes));
    st->attr_set = true;
  }
  score_tc_kernel<KS><<<grid, kThreads, C::smem_bytes, s>>>(p);
  VB_CUDA(cudaGetLastError());
  return 0;
}

void note_fallback(vbgpu_gmm_t h, const char *why) {
  h->tc_note = why;
  if (!getenv("VBGPU_QUIET"))
    fprintf(stderr, "vbgpu: model (P=%d N=%d D=%d) is scored by the FP32 SIMT kernel, not tcgen05: %s\n", h->P, h->N, h->D, why);
}

}  // namespace

namespace vb {

void score_tc_release(vbgpu_gmm_t h) {
  TcState *st = static_cast<TcState *>(h->tc);
  if (!st) return;
  for (DevBuf *b : {&st->d_bimg, &st->d_hdr, &st->d_centre, &st->d_s1, &st->d_s2, &st->d_col_of_pdf, &st->d_merge,
                    &st->d_rowflag, &st->d_scratch})
    b->release();
  delete st;
  h->tc = nullptr;
}

bool score_tc_available(vbgpu_gmm_t h) { return h->tc != nullptr; }
int32_t score_tc_num_cols(vbgpu_gmm_t h) { return h->tc ? static_cast<TcState *>(h->tc)->n_cols : h->P; }
const int32_t *score_tc_col_of_pdf(vbgpu_gmm_t h) {
  return h->tc ? static_cast<TcState *>(h->tc)->col_of_pdf.data() : nullptr;
}
const int32_t *score_tc_col_of_pdf_dev(vbgpu_gmm_t h) {
  return h->tc ? static_cast<TcState *>(h->tc)->d_col_of_pdf.as<int32_t>() : nullptr;
}

// Builds the tensor-core image of the model.  Leaves h->tc null (the SIMT kernel serves the model, h->tc_note says why)
// when the model does not fit: D > 47, a pdf with no finite gconst, or values outside the fp16 plan.
int score_tc_prepare(vbgpu_gmm_t h, const float *gconsts, const float *miv, const float *iv, int32_t stride) {
  h->tc_note = "";
  if (getenv("VBGPU_DISABLE_TC")) {
    h->tc_note = "VBGPU_DISABLE_TC is set";
    return 0;
  }
  TcState *st = new TcState;
  TcHostImage img;
  const char *why = build_layout(h->D, h->N, h->P, h->h_pdf_offsets, gconsts, miv, iv, stride, st, &img);
  if (why) {
    delete st;
    note_fallback(h, why);
    return 0;
  }
  int rc = 0;
  auto up = [&](DevBuf &b, const void *src, size_t bytes) {
    if (rc == 0) rc = b.reserve(std::max<size_t>(bytes, 16));
    if (rc == 0 && bytes && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(VBGPU_ERR_CUDA, "upload of the tensor-core model image failed");
  };
  up(st->d_bimg, st->h_bimg.data(), st->h_bimg.size());
  up(st->d_hdr, img.hdr.data(), img.hdr.size() * sizeof(int4));
  up(st->d_centre, img.centre.data(), h->D * 4);
  up(st->d_s1, img.s1.data(), h->D * 4);
  up(st->d_s2, img.s2.data(), h->D * 4);
  up(st->d_col_of_pdf, st->col_of_pdf.data(), (size_t)h->P * 4);
  up(st->d_merge, img.merge.data(), img.merge.size() * 4);
  h->tc = st;
  if (rc < 0) score_tc_release(h);
  return rc;
}

// Host-only view of the layout (no device needed): what tests/test_tc_layout.py decodes and checks against the oracle.
int score_tc_debug_layout(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *miv,
                          const float *iv, int32_t stride, int32_t *info, uint8_t *image, int64_t image_cap, int32_t *hdr,
                          int32_t hdr_cap, int32_t *col_of_pdf, int32_t *merge, int32_t merge_cap, float *centre, float *s1,
                          float *s2) {
  TcState st;
  TcHostImage img;
  std::vector<int32_t> po(pdf_offsets, pdf_offsets + P + 1);
  const char *why = build_layout(D, po[P], P, po, gconsts, miv, iv, stride, &st, &img);
  if (why) return fail(VBGPU_ERR_INVALID, "not on the tensor-core plan: %s", why);
  info[0] = st.KS, info[1] = st.n_panels, info[2] = st.n_cols, info[3] = st.n_merge;
  info[4] = (int32_t)(st.h_bimg.size() >> 4);
  if (image && (int64_t)st.h_bimg.size() <= image_cap) std::memcpy(image, st.h_bimg.data(), st.h_bimg.size());
  if (hdr && st.n_panels * 4 <= hdr_cap) std::memcpy(hdr, img.hdr.data(), (size_t)st.n_panels * 16);
  if (col_of_pdf) std::memcpy(col_of_pdf, st.col_of_pdf.data(), (size_t)P * 4);
  if (merge && (int)img.merge.size() <= merge_cap) std::memcpy(merge, img.merge.data(), img.merge.size() * 4);
  if (centre) std::memcpy(centre, img.centre.data(), D * 4);
  if (s1) std::memcpy(s1, img.s1.data(), D * 4);
  if (s2) std::memcpy(s2, img.s2.data(), D * 4);
  return 0;
}

int score_tc_update_gconsts(vbgpu_gmm_t h, const float *gconsts) {
  TcState *st = static_cast<TcState *>(h->tc);
  if (!st) return 0;
  const std::vector<int32_t> &po = h->h_pdf_offsets;
  for (int pdf = 0; pdf < h->P; pdf++) {
    bool any = false;
    for (int g = po[pdf]; g < po[pdf + 1]; g++) any |= (gconsts[g] > -INFINITY);
    if (!any) {  // a pdf without a finite gconst: only the SIMT kernel reproduces the reference's -inf
      score_tc_release(h);
      note_fallback(h, "a pdf has no Gaussian with a finite gconst");
      return 0;
    }
  }
  const double L2E = 1.4426950408889634074;
  for (int g = 0; g < h->N; g++) {
    double gc = kDummy;
    if (gconsts[g] > -INFINITY) {
      gc = ((double)gconsts[g] + st->gshift[g]) * L2E;
      if (!(std::fabs(gc) <= kGcMax)) {
        score_tc_release(h);
        note_fallback(h, "a centred gconst is too large for FP32 accumulation at 1e-3");
        return 0;
      }
    }
    const GaussPos gp = st->gpos[g];
    put_gconst(st->h_bimg.data() + st->panel_off[gp.panel], st->panel_n[gp.panel], st->KS, gp.col, h->D, gc);
  }
  VB_CUDA(cudaMemcpy(st->d_bimg.p, st->h_bimg.data(), st->h_bimg.size(), cudaMemcpyHostToDevice));
  return 0;
}

// Scores T frames.  native != 0: d_ll is [T x ll_stride] in DEVICE COLUMN ORDER (ll_stride >= score_tc_num_cols());
// native == 0: d_ll is [T x ll_stride] in the model's pdf order (the kernel writes a scratch matrix, a gather kernel
// produces d_ll).
int score_tc_launch(vbgpu_gmm_t h, const float *d_feats, int64_t T, int32_t stride, float *d_ll, int32_t ll_stride,
                    int native, cudaStream_t s) {
  TcState *st = static_cast<TcState *>(h->tc);
  if (!st) return fail(VBGPU_ERR_INVALID, "tensor-core scorer unavailable for this model");
  if (T == 0) return 0;
  float *out = d_ll;
  int32_t out_stride = ll_stride;
  if (!native) {
    out_stride = st->n_cols;
    VB_TRY(st->d_scratch.reserve((size_t)T * out_stride * 4));
    out = st->d_scratch.as<float>();
  } else if (ll_stride < st->n_cols) {
    return fail(VBGPU_ERR_INVALID, "ll_stride %d < %d device columns", ll_stride, st->n_cols);
  }
  VB_TRY(st->d_rowflag.reserve((size_t)T));
  VB_CUDA(cudaMemsetAsync(st->d_rowflag.p, 0, (size_t)T, s));
  const int sms = num_sms(h->device);
  const int64_t n_mtiles = (T + kMt * kRowsMt - 1) / (kMt * kRowsMt);
  // split the panels over CTAs when there are too few frame tiles: pick the split with the best last-wave fill
  int best = 1;
  int64_t n_whole = 0;
  if (n_mtiles < 4LL * sms) {
    double best_eff = 0.0;
    const int max_split = std::min(st->n_panels, 64);
    for (int k = 1; k <= max_split; k++) {
      const int64_t units = n_mtiles * k, waves = (units + sms - 1) / sms;
      // each extra split repeats the A-panel build: charge it as ~2 panels of work per unit
      const double work = (double)st->n_panels / k + 2.0;
      const double eff = ((double)st->n_panels / k) / work * (double)units / (double)(waves * sms);
      if (eff > best_eff * 1.02) best_eff = eff, best = k;
    }
  } else {
    // many tiles: whole tiles for the full waves; the r tiles of the last, partial wave are cut into floor(sms / r) panel
    // ranges each so that the tail runs on (almost) every SM for 1/k of a tile's time instead of on r SMs for all of it
    const int64_t r = n_mtiles % sms;
    int k = r > 0 ? (int)std::min<int64_t>(std::min(st->n_panels, 16), sms / r) : 1;
    if (getenv("VBGPU_TC_NO_TAIL_SPLIT")) k = 1;  // bring-up: A/B of the tail split
    if (k >= 2) best = k, n_whole = n_mtiles - r;
    else n_whole = n_mtiles;
  }
  TcParams p;
  p.feats = d_feats;
  p.T = T;
  p.stride = stride;
  p.D = h->D;
  p.bimg = st->d_bimg.as<uint8_t>();
  p.hdr = st->d_hdr.as<int4>();
  p.centre = st->d_centre.as<float>();
  p.s1 = st->d_s1.as<float>();
  p.s2 = st->d_s2.as<float>();
  p.n_panels = st->n_panels;
  p.n_splits = best;
  p.n_whole = n_whole;
  p.n_units = n_whole + (n_mtiles - n_whole) * best;
  p.out = out;
  p.ll_stride = out_stride;
  const int padded = (h->D + 3) / 4 * 4;
  p.vec_ok = (((reinterpret_cast<uintptr_t>(out) & 15) == 0 && out_stride % 4 == 0) ? 1 : 0) |
             (((reinterpret_cast<uintptr_t>(d_feats) & 15) == 0 && stride % 4 == 0 && stride >= padded) ? 2 : 0);
  p.bad = h->d_bad.as<unsigned long long>();
  p.rowflag = st->d_rowflag.as<uint8_t>();
  const char *dbg_env = getenv("VBGPU_TC_DEBUG");
  p.dbg = dbg_env ? (uint32_t)atoi(dbg_env) : 0u;
  const int grid = (int)std::min<int64_t>(p.n_units, sms);
  int rc;
  switch (st->KS) {
    case 2: rc = launch_ks<2>(p, st, grid, s); break;
    case 3: rc = launch_ks<3>(p, st, grid, s); break;
    case 4: rc = launch_ks<4>(p, st, grid, s); break;
    case 5: rc = launch_ks<5>(p, st, grid, s); break;
    case 6: rc = launch_ks<6>(p, st, grid, s); break;
    default: return fail(VBGPU_ERR_INVALID, "unsupported K for the tensor-core scorer");
  }
  VB_TRY(rc);
  if (st->n_merge > 0) {
    score_merge_kernel<<<(unsigned)((T + 255) / 256), 256, 0, s>>>(out, T, out_stride, st->d_merge.as<int32_t>(), st->n_merge);
    VB_CUDA(cudaGetLastError());
  }
  {
    const int64_t chunks = (T + 31) / 32;
    const int blocks = (int)std::min<int64_t>((chunks + 7) / 8, 4LL * sms);
    score_fix_kernel<<<blocks, 256, 0, s>>>(st->d_rowflag.as<uint8_t>(), d_feats, T, stride, h->D, h->DP,
                                            h->d_rows.as<float>(), h->d_gconsts.as<float>(), h->d_pdf_offsets.as<int32_t>(),
                                            st->d_col_of_pdf.as<int32_t>(), h->P, out, out_stride, p.bad);
    VB_CUDA(cudaGetLastError());
  }
  if (!native) {
    score_gather_pdf_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(out, out_stride, st->d_col_of_pdf.as<int32_t>(), h->P, T,
                                                                   d_ll, ll_stride);
    VB_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace vb
