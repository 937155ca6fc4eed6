// score_tc.cu — dense Gaussian scoring on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM) with the
// per-pdf log-sum-exp fused into the epilogue.
//
//   loglikes[t][p] = LogSumExp_{m in pdf p}( gconst_m + means_invvars_m . x_t - 0.5 inv_vars_m . x_t^2 )
//
// which DiagGmm::LogLikelihoods (gmm/diag-gmm.cc:528-562) + VectorBase::LogSumExp (matrix/kaldi-vector.cc:757-775)
// compute per (frame, pdf) behind DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased (gmm/decodable-am-diag-gmm.cc:28-72).
// Here it is ONE contraction  Y[T x N] = A[T x K] . B[K x N]  with  K = 2D+1:
//     A row t    = [ x_0 s1_0, x_0^2 s2_0, x_1 s1_1, x_1^2 s2_1, ..., 1 ]               (x already centred, see below)
//     B column g = [ miv'_0/s1_0, -0.5 iv_0/s2_0, ...,                  gconst' ] * log2(e)
// so that Y is the per-Gaussian log-likelihood in log2 units and the epilogue is exp2/log2 only.
//
// Precision.  The tensor cores take 11-bit mantissas; a single pass is ~0.05 abs off.  Both operands are
// therefore split in two fp16 halves (v = hi + lo, 22 bits) and three products are accumulated in the FP32 TMEM tile:
// lo.hi + hi.lo + hi.hi (lo.lo is below 2^-22 relative).  fp16 has the mantissa of TF32 at twice the MMA rate; its
// narrow exponent range is handled by exact power-of-two scales per dimension (s1, s2, chosen from the model) and by
// centring the features on the mean of the model means (c; folded exactly into miv' = miv - iv c and
// gconst' = gconst + miv.c - 0.5 iv.c^2, computed in double).  Measured against a float64 restatement the result is
// as accurate as the reference's own FP32 BLAS path (~6e-5 max abs; budget 1e-3).
//
// Kernel shape (persistent, one CTA per SM, 18 warps, warp-specialised):
//   warp 16    producer : cp.async.bulk (TMA engine, 1-D) of pre-tiled B panels [128 Gaussians x K] + their part tables
//                         into a 3-stage shared-memory ring (mbarrier complete_tx).
//   warp 17    MMA      : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) — 3*K/16 per
//                         accumulator — for TWO 128-frame accumulators that share every B panel; accumulators are
//                         double-buffered in TMEM (4 x 128 columns = all 512 columns).
//   warps 0-15 epilogue : build the fp16 hi/lo A panel of the CTA's 256 frames once per work unit (straight from the FP32
//                         features), then per B panel: tcgen05.ld the accumulator (lane = frame), segmented two-pass
//                         log-sum-exp over the Gaussians of each pdf, 16-byte stores of 4 consecutive pdfs per frame.
//                         Four warps share a TMEM lane quarter (4 per SM sub-partition: the MUFU pipe, 16 ex2/clk/SM, is
//                         the epilogue's floor and needs that many warps in flight); they split the pdfs by aligned
//                         groups of four, and each walks only ITS parts through a per-panel table sorted by owner.
// Work unit = (256-frame tile, range of B panels).  Large batches use one range (all panels); small batches split the
// panels over CTAs at "super-block" boundaries (every 4 panels a pdf boundary is forced by padding) to fill the GPU.
#include <cuda_fp16.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.h"

namespace {

constexpr int kTileN = 128;    // Gaussians per B panel (UMMA N)
constexpr int kRowsMt = 128;   // frames per accumulator (UMMA M)
constexpr int kMt = 2;         // accumulators per CTA
constexpr int kEpiWarps = 16;  // 4 per TMEM lane quarter
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kTabRing = 7;    // ring of segment tables (lifetime analysis in DESIGN.md: >= 6)
// Part table of one panel: 4 x { u32 first part | n parts << 16 ; i32 first pdf of the panel } (one pair per owner
// class) then u32 parts[<=128], grouped by owner class, column order inside a class.
//   part = column | len << 8 | (pdf - first pdf) << 16 | ends-its-pdf << 24 | continues-a-pdf << 25.
// A part is a run of <= 16 columns of one pdf that a single power-of-two tcgen05.ld covers without leaving the panel.
constexpr int kTabBytes = 544;
constexpr uint32_t kPartEnds = 1u << 24, kPartCont = 1u << 25;
constexpr int kSbTiles = 4;    // panels per super-block
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
constexpr float kDummy = -40000.0f;  // log2-domain score of padding columns / zero-weight Gaussians

template <int KS>
struct Cfg {  // KS = 16-wide K steps per split; K = 16*KS >= 2D+1
  static constexpr int kc_half = 2 * KS;     // 16-byte K chunks per split
  static constexpr int kc = 4 * KS;          // hi + lo
  static constexpr int a_bytes = kc * 2048;  // one 128-row A panel: [kc][16 row groups][8 rows x 16 B]
  static constexpr int b_bytes = kc * 2048;  // one 128-column B panel, same canonical K-major layout
  static constexpr int stages = (KS <= 5) ? 3 : 2;
  static constexpr int off_b = kMt * a_bytes;
  static constexpr int off_tab = off_b + stages * b_bytes;
  static constexpr int off_stg = off_tab + ((kTabRing * kTabBytes + 127) / 128) * 128;
  static constexpr int off_carry = off_stg + kEpiWarps * 2 * 4 * 32 * 4;  // [warp][2][32] partial log-sum-exps of a cut pdf
  static constexpr int off_bar = off_carry + kEpiWarps * 2 * 32 * 4;
  static constexpr int smem_bytes = off_bar + 256;
};

struct TcParams {
  const float *feats;
  int64_t T;
  int32_t stride, D;
  const uint8_t *bimg;  // [n_tiles][b_bytes]
  const uint8_t *tabs;  // [n_tiles][kTabBytes]
  const float *centre, *s1, *s2;  // [D]
  const int32_t *sb_tile, *sb_pdf;  // [n_sb+1]: first panel / first pdf of each super-block
  int32_t n_sb, n_splits;
  int64_t n_units, n_whole;
  float *out;
  int32_t ll_stride, vec_ok;
  unsigned long long *bad;
  uint32_t lbo, sbo;
  uint32_t dbg;  // bring-up only (VBGPU_TC_DEBUG): bit 0 = epilogue skips the math, bit 1 = no MMAs are issued,
                 // bit 2 = accumulators are released after a warp's last part instead of after that part's loads (A/B)
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
// L (1, 2, 4, 8 or 16) accumulator columns starting at column address t.
template <int L>
__device__ __forceinline__ void tmem_ld(uint32_t t, float (&v)[L]) {
  if constexpr (L == 16) tmem_ld16(t, v);
  else if constexpr (L == 8) tmem_ld8(t, v);
  else if constexpr (L == 4) tmem_ld4(t, v);
  else if constexpr (L == 2) tmem_ld2(t, v);
  else tmem_ld1(t, v);
}
__host__ __device__ constexpr int ld_width(int S) { return S <= 1 ? 1 : S <= 2 ? 2 : S <= 4 ? 4 : S <= 8 ? 8 : 16; }

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}
template <int S, int L>
__device__ __forceinline__ float max_n(const float (&v)[L]) {
  float m0 = v[0];
  if constexpr (S >= 8) {  // two chains
    float m1 = v[S / 2];
#pragma unroll
    for (int i = 1; i + 1 < S / 2; i += 2) m0 = max3f(m0, v[i], v[i + 1]);
    if constexpr (((S / 2) & 1) == 0) m0 = fmaxf(m0, v[S / 2 - 1]);
#pragma unroll
    for (int i = S / 2 + 1; i + 1 < S; i += 2) m1 = max3f(m1, v[i], v[i + 1]);
    if constexpr (((S - S / 2) & 1) == 0) m1 = fmaxf(m1, v[S - 1]);
    return fmaxf(m0, m1);
  } else {
#pragma unroll
    for (int i = 1; i + 1 < S; i += 2) m0 = max3f(m0, v[i], v[i + 1]);
    if constexpr ((S & 1) == 0) m0 = fmaxf(m0, v[S - 1]);
    return m0;
  }
}
// sum_i 2^(v_i - M) over the first S of L values: the subtraction and the summation run as packed FP32 pairs
// (add.rn.f32x2, two lanes per issue slot); the exponentials are the MUFU pipe's, one per value.
template <int S, int L>
__device__ __forceinline__ float sum_ex2(const float (&v)[L], float M) {
  static_assert(S >= 2, "S == 1 needs no exponential");
  const float2 nm = make_float2(-M, -M);
  float2 acc;
#pragma unroll
  for (int i = 0; i + 1 < S; i += 2) {
    const float2 d = __fadd2_rn(make_float2(v[i], v[i + 1]), nm);
    const float2 e = make_float2(ex2f(d.x), ex2f(d.y));
    acc = (i == 0) ? e : __fadd2_rn(acc, e);
  }
  float s = acc.x + acc.y;
  if constexpr ((S & 1) != 0) s += ex2f(v[S - 1] - M);
  return s;
}
// Log-sum-exp pieces (max, sum of 2^(y - max)) of S accumulator columns, for the thread's frame in BOTH accumulators
// (two independent dependency chains per thread).  One power-of-two load per accumulator; the host never emits a part
// whose load would leave its panel.
// release != 0: this is the warp's last part of the panel — once its values are in registers the warp no longer needs the
// accumulators, so it hands the buffer back to the MMA warp BEFORE doing the part's arithmetic (the MMA warp waits for the
// slowest of the 16 warps: every cycle shaved off the hold time is a cycle of tensor-pipe time won).
template <int S>
__device__ __forceinline__ void seg_lse2(uint32_t tA, uint32_t tB, float &MA, float &sA, float &MB, float &sB,
                                         uint32_t release, int lane) {
  constexpr int L = ld_width(S);
  float a[L], b[L];
  tmem_ld<L>(tA, a);
  tmem_ld<L>(tB, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  // branch-free: the fence and the warp sync cost two issue slots on every part, a predicated arrive replaces the branch
  tc_fence_before();
  __syncwarp();
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(release),
      "r"((lane == 0 && release != 0u) ? 1u : 0u)
      : "memory");
#pragma unroll
  for (int i = 0; i < S; i++) {  // pin every consumer behind the wait
    asm volatile("" : "+f"(a[i]));
    asm volatile("" : "+f"(b[i]));
  }
  if constexpr (S == 1) {
    MA = a[0], MB = b[0], sA = 1.0f, sB = 1.0f;
  } else {
    MA = max_n<S, L>(a);
    MB = max_n<S, L>(b);
    sA = sum_ex2<S, L>(a, MA);
    sB = sum_ex2<S, L>(b, MB);
  }
}
// (M, s) += (M2, s2) in the log domain: one of the two rescale factors is 2^0, so one ex2 serves.
__device__ __forceinline__ void lse_merge(float &M, float &s, float M2, float s2) {
  const float d = M2 - M, e = ex2f(-fabsf(d));
  s = (d > 0.0f) ? fmaf(s, e, s2) : fmaf(s2, e, s);
  M = fmaxf(M, M2);
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
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kRowsMt >> 4) << 24);

struct UnitRange {
  int64_t mtile;
  int32_t t0, t1, p0, p1;
};
// Units 0 .. n_whole-1 are whole frame tiles (all panels); the frame tiles after them are each cut into n_splits units by
// panel range.  Large batches: whole tiles fill the full waves and only the tiles of the last, partial wave are cut, so that
// the tail keeps every SM busy.  Small batches: n_whole = 0, every tile is cut.
__device__ __forceinline__ UnitRange unit_range(const TcParams &p, int64_t u) {
  UnitRange r;
  if (u < p.n_whole) {
    r.mtile = u;
    r.t0 = __ldg(p.sb_tile);
    r.t1 = __ldg(p.sb_tile + p.n_sb);
    r.p0 = __ldg(p.sb_pdf);
    r.p1 = __ldg(p.sb_pdf + p.n_sb);
    return r;
  }
  u -= p.n_whole;
  r.mtile = p.n_whole + u / p.n_splits;
  const int split = (int)(u - (r.mtile - p.n_whole) * p.n_splits);
  const int sb0 = (int)(((int64_t)split * p.n_sb) / p.n_splits), sb1 = (int)(((int64_t)(split + 1) * p.n_sb) / p.n_splits);
  r.t0 = __ldg(p.sb_tile + sb0);
  r.t1 = __ldg(p.sb_tile + sb1);
  r.p0 = __ldg(p.sb_pdf + sb0);
  r.p1 = __ldg(p.sb_pdf + sb1);
  return r;
}

// barrier slots
enum { kBarFull = 0, kBarEmpty = 3, kBarAccFull = 6, kBarAccEmpty = 8, kBarAReady = 10 };

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
    for (int i = 0; i < 3; i++) mbar_init(BAR(kBarFull + i), 1);
    for (int i = 0; i < 3; i++) mbar_init(BAR(kBarEmpty + i), 1);
    for (int i = 0; i < 2; i++) mbar_init(BAR(kBarAccFull + i), 2);  // tcgen05.commit + the issuing thread's own arrive
    for (int i = 0; i < 2; i++) mbar_init(BAR(kBarAccEmpty + i), kEpiWarps);
    mbar_init(BAR(kBarAReady), kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + 1) {  // TMEM: all 512 columns (the CTA owns the SM: ~220 KB of shared memory)
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
        for (int t = ur.t0; t < ur.t1; t++, it++) {
          const uint32_t s = it % C::stages, ph = (it / C::stages) & 1;
          mbar_wait(BAR(kBarEmpty + s), ph ^ 1);
          mbar_expect_tx(BAR(kBarFull + s), C::b_bytes + kTabBytes);
          bulk_g2s(smem_u32(smem + C::off_b + s * C::b_bytes), p.bimg + (size_t)t * C::b_bytes, C::b_bytes,
                   BAR(kBarFull + s));
          bulk_g2s(smem_u32(smem + C::off_tab + (it % kTabRing) * kTabBytes), p.tabs + (size_t)t * kTabBytes, kTabBytes,
                   BAR(kBarFull + s));
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ================================================= MMA issuer ===============================================
    // The whole warp walks the loop converged (every value below is warp-uniform, so descriptors and addresses are
    // formed in uniform registers); one elected lane issues the tensor-core instructions and the commits.
    {
      uint32_t it = 0, un = 0;
      const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + C::off_b);
      for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x, un++) {
        const UnitRange ur = unit_range(p, u);
        mbar_wait(BAR(kBarAReady), un & 1);  // the A panels of this unit are in shared memory
        for (int t = ur.t0; t < ur.t1; t++, it++) {
          const uint32_t s = it % C::stages, ph = (it / C::stages) & 1, buf = it & 1, aph = (it >> 1) & 1;
          mbar_wait(BAR(kBarFull + s), ph);
          mbar_wait(BAR(kBarAccEmpty + buf), aph ^ 1);  // the epilogue has drained both accumulators of this buffer
          tc_fence_after();
          const uint64_t bdesc = make_desc(b_base + s * C::b_bytes, 2048, 128);
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < kMt; mt++) {
              if (dbg & 2u) break;
              const uint32_t d = tmem_base + (uint32_t)((mt * 2 + buf) * kTileN);
              const uint64_t adesc = make_desc(a_base + mt * C::a_bytes, 2048, 128);
#pragma unroll
              for (int prod = 0; prod < 3; prod++) {  // lo.hi, hi.lo, hi.hi (small terms first)
                const uint32_t ao = (prod == 0) ? C::kc_half * 2048u : 0u, bo = (prod == 1) ? C::kc_half * 2048u : 0u;
#pragma unroll
                for (int k = 0; k < KS; k++)  // one K=16 step = two 2048-byte chunks
                  tc_mma_f16(d, adesc + ((ao + k * 4096u) >> 4), bdesc + ((bo + k * 4096u) >> 4), kIdesc,
                             (prod | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(BAR(kBarAccFull + buf));
            // A plain (release) arrive by this thread as well: it has acquired the stage's `full` barrier, so the
            // epilogue's acquire of accfull also orders it after the bulk copy of the panel's part table.
            mbar_arrive(BAR(kBarAccFull + buf));
            tc_commit(BAR(kBarEmpty + s));  // the B panel (and every MMA before it) is done: free the stage
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================= epilogue =================================================
    // Warp w serves TMEM lanes 32*(w&3)..+31, i.e. frame (w&3)*32+lane of BOTH accumulators (two frames per thread,
    // two independent chains).  The four warps of a lane quarter split the pdfs by aligned groups of four:
    // cls = w>>2 owns the pdfs with ((pdf >> 2) & 3) == cls, so each warp produces whole 16-byte output groups.
    const int q = warp & 3, cls = warp >> 2;
    float *stgA = reinterpret_cast<float *>(smem + C::off_stg) + warp * 256, *stgB = stgA + 128;  // [4 pdfs][32 lanes]
    uint32_t it = 0;
    // Non-finite results can only come from non-finite (or unrepresentably large) features: the model image is
    // validated on the host, the operands are bounded and every sum of exponentials is >= 1.  They are counted where
    // the features are read, not per stored value.
    unsigned long long nbad = 0;
#pragma unroll 1
    for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const UnitRange ur = unit_range(p, u);
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
        uint32_t vec_in;  // read through an opaque move: keeps the compiler from cloning the whole unit loop per flag
        asm volatile("mov.u32 %0, %1;" : "=r"(vec_in) : "r"(p.vec_ok));
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
        uint8_t *arow = smem + mt * C::a_bytes + (rowl >> 3) * 128 + (rowl & 7) * 16 + kh * KS * 2048;
#pragma unroll
        for (int kc = 0; kc < KS; kc++) {
          __half2 hi[4], lo[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; e2++) {
            const int d = d0 + kc * 4 + e2;  // K index 2d -> x_d * s1_d, 2d+1 -> x_d^2 * s2_d, 2D -> 1
            float v0 = 0.0f, v1 = 0.0f;
            if (d < p.D) {
              const float xc = x[kc * 4 + e2] - __ldg(p.centre + d);
              v0 = xc * __ldg(p.s1 + d);
              v1 = (xc * xc) * __ldg(p.s2 + d);
            } else if (d == p.D) {
              v0 = 1.0f;
            }
            if (!(fabsf(v0) <= 65504.0f) || !(fabsf(v1) <= 65504.0f)) nbad++;  // NaN/Inf, or outside the fp16 plan
            v0 = fminf(fmaxf(v0, -65504.0f), 65504.0f);
            v1 = fminf(fmaxf(v1, -65504.0f), 65504.0f);
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            hi[e2] = __halves2half2(h0, h1);
            lo[e2] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
          }
          *reinterpret_cast<uint4 *>(arow + kc * 2048) = *reinterpret_cast<uint4 *>(hi);
          *reinterpret_cast<uint4 *>(arow + (kc + C::kc_half) * 2048) = *reinterpret_cast<uint4 *>(lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(kBarAReady));
      }

      // ---- panels ----
      const int64_t trowA = ur.mtile * (kMt * kRowsMt) + q * 32 + lane, trowB = trowA + kRowsMt;
      float *orowA = p.out + trowA * p.ll_stride, *orowB = p.out + trowB * p.ll_stride;
      const int p_lo = ur.p0;
      // The earlier parts of a pdf whose last part is still to come (rare: a pdf cut by a panel edge or longer than 16)
      // wait in shared memory as ONE number per accumulator, L = M + log2(s); kept in registers they were loop-carried
      // state that cost every part eight register moves.
      float *carry = reinterpret_cast<float *>(smem + C::off_carry) + warp * 64;
      uint32_t vec_out;  // opaque read: keeps the compiler from cloning the panel loop per loop-invariant flag
      asm volatile("mov.u32 %0, %1;" : "=r"(vec_out) : "r"(p.vec_ok));
      const bool liveA = trowA < p.T, liveB = trowB < p.T;
      const bool fast = (vec_out & 1u) && liveB;  // (liveB implies liveA)
      auto store_group = [&](int pb, int n) {  // pdfs pb..pb+n-1 of the staged group, clipped to this unit's range
        if (fast && n == 4 && pb >= p_lo) {
          *reinterpret_cast<float4 *>(orowA + pb) = make_float4(stgA[lane], stgA[32 + lane], stgA[64 + lane], stgA[96 + lane]);
          *reinterpret_cast<float4 *>(orowB + pb) = make_float4(stgB[lane], stgB[32 + lane], stgB[64 + lane], stgB[96 + lane]);
        } else {
          for (int k = 0; k < n; k++)
            if (pb + k >= p_lo) {
              if (liveA) orowA[pb + k] = stgA[k * 32 + lane];
              if (liveB) orowB[pb + k] = stgB[k * 32 + lane];
            }
        }
      };
#pragma unroll 1
      for (int t = ur.t0; t < ur.t1; t++, it++) {
        const uint32_t buf = it & 1, aph = (it >> 1) & 1;
        const uint32_t *tab = reinterpret_cast<const uint32_t *>(smem + C::off_tab + (it % kTabRing) * kTabBytes);
        mbar_wait(BAR(kBarAccFull + buf), aph);
        tc_fence_after();
        const uint32_t tA = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kTileN), tB = tA + 2 * kTileN;
        const uint2 hdr = *reinterpret_cast<const uint2 *>(tab + 2 * cls);
        const int first_pdf = (int)hdr.y;
        const uint32_t *part = tab + 8 + (hdr.x & 0xffffu);
        const int n_parts = (dbg & 1u) ? 0 : (int)(hdr.x >> 16);
        uint32_t e_next = n_parts > 0 ? part[0] : 0u;
#pragma unroll 1
        for (int i = 0; i < n_parts; i++) {
          const uint32_t e = e_next;
          if (i + 1 < n_parts) e_next = part[i + 1];
          const uint32_t col = e & 255u;
          const int len = (int)((e >> 8) & 255u), pdf = first_pdf + (int)((e >> 16) & 255u);
          const uint32_t rel = (i + 1 == n_parts && !(dbg & 4u)) ? BAR(kBarAccEmpty + buf) : 0u;  // last part: early release
          float MA, sA, MB, sB;
          switch (len) {
#define VB_CASE(S) case S: seg_lse2<S>(tA + col, tB + col, MA, sA, MB, sB, rel, lane); break;
            VB_CASE(1) VB_CASE(2) VB_CASE(3) VB_CASE(4) VB_CASE(5) VB_CASE(6) VB_CASE(7) VB_CASE(8)
            VB_CASE(9) VB_CASE(10) VB_CASE(11) VB_CASE(12) VB_CASE(13) VB_CASE(14) VB_CASE(15)
            default: seg_lse2<16>(tA + col, tB + col, MA, sA, MB, sB, rel, lane); break;
#undef VB_CASE
          }
          if (e & kPartCont) {  // the earlier parts of this pdf (rare: a pdf cut by a panel edge or longer than 16)
            lse_merge(MA, sA, carry[lane], 1.0f);
            lse_merge(MB, sB, carry[32 + lane], 1.0f);
          }
          if (e & kPartEnds) {
            stgA[(pdf & 3) * 32 + lane] = (MA + lg2f(sA)) * kLn2;
            stgB[(pdf & 3) * 32 + lane] = (MB + lg2f(sB)) * kLn2;
            if ((pdf & 3) == 3) store_group(pdf - 3, 4);
          } else {
            carry[lane] = MA + lg2f(sA), carry[32 + lane] = MB + lg2f(sB);
          }
        }
        // a warp without parts in this panel hands the buffer back here (the others did so inside their last part)
        if (n_parts == 0 || (dbg & 4u)) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(kBarAccEmpty + buf));
        }
      }
      // the unit's last group of four may be incomplete: its owner stores what exists
      if (!(dbg & 1u) && (ur.p1 & 3) != 0 && (((ur.p1 >> 2) & 3) == cls)) store_group(ur.p1 & ~3, ur.p1 & 3);
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

// ---- host-side state ------------------------------------------------------------------------------------------------
struct TcState {
  int KS = 0, n_tiles = 0, n_sb = 0;
  size_t b_bytes = 0;
  std::vector<uint8_t> h_bimg;       // kept for gconst updates
  std::vector<int32_t> col_of_gauss; // global column of every Gaussian
  std::vector<double> gshift;        // gconst' - gconst  (centring term), per Gaussian
  vb::DevBuf d_bimg, d_tabs, d_centre, d_s1, d_s2, d_sb_tile, d_sb_pdf;
  bool attr_set = false;
  uint32_t lbo = 2048, sbo = 128;  // descriptor strides: K-chunk (leading) and 8-row-group (stride) byte offsets, verified on B200
};

inline void put_half_pair(uint8_t *img, size_t b_bytes, int KS, int col, int k, double v) {
  // element (column n, K index k) of the hi half lives in chunk kc = k/8, the lo half in chunk kc + 2*KS
  const int tile = col / kTileN, n = col % kTileN;
  const __half hi = __double2half(v);
  const __half lo = __double2half(v - (double)__half2float(hi));
  uint8_t *base = img + (size_t)tile * b_bytes;
  const size_t off_hi = ((size_t)(k / 8) * 16 + n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
  const size_t off_lo = off_hi + (size_t)2 * KS * 2048;
  std::memcpy(base + off_hi, &hi, 2);
  std::memcpy(base + off_lo, &lo, 2);
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

}  // namespace

namespace vb {

void score_tc_release(vbgpu_gmm_t h) {
  TcState *st = static_cast<TcState *>(h->tc);
  if (!st) return;
  for (DevBuf *b : {&st->d_bimg, &st->d_tabs, &st->d_centre, &st->d_s1, &st->d_s2, &st->d_sb_tile, &st->d_sb_pdf})
    b->release();
  delete st;
  h->tc = nullptr;
}

bool score_tc_available(vbgpu_gmm_t h) { return h->tc != nullptr; }

// Builds the tensor-core image of the model.  Leaves h->tc null (the SIMT kernel serves the model) when the model
// does not fit: D > 47, a pdf with more than 512 Gaussians or with no finite gconst, or values outside the fp16 plan.
int score_tc_prepare(vbgpu_gmm_t h, const float *gconsts, const float *miv, const float *iv, int32_t stride) {
  const int D = h->D, N = h->N, P = h->P;
  if (getenv("VBGPU_DISABLE_TC")) return 0;
  int KS = (2 * D + 1 + 15) / 16;
  if (KS < 2) KS = 2;
  if (KS > 6) return 0;
  const std::vector<int32_t> &po = h->h_pdf_offsets;
  for (int pdf = 0; pdf < P; pdf++) {
    if (po[pdf + 1] - po[pdf] > kSbTiles * kTileN) return 0;
    bool any = false;
    for (int g = po[pdf]; g < po[pdf + 1]; g++) any |= (gconsts[g] > -INFINITY);
    if (!any) return 0;
  }
  // centre and scales from the model: means mu = miv/iv, sigma = iv^-1/2
  std::vector<double> c(D, 0.0), R(D, 0.0);
  for (int g = 0; g < N; g++)
    for (int d = 0; d < D; d++) {
      const double v = iv[(size_t)g * stride + d];
      if (!(v > 0.0) || !std::isfinite(v) || !std::isfinite((double)miv[(size_t)g * stride + d])) return 0;
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
    if (!(R[d] > 0.0) || !std::isfinite(R[d])) return 0;
    const int e = (int)std::ceil(std::log2(R[d]));
    if (e < -40 || e > 40) return 0;
    s1[d] = (float)std::ldexp(1.0, 5 - e);                  // |x - c| <= R  ->  |a| <= 32
    s2[d] = (float)std::ldexp(1.0, 2 * (5 - e) - 4);        // (x - c)^2 s2 <= 64
  }
  // column layout: pdfs in order, never straddling a super-block (4 panels)
  TcState *st = new TcState;
  st->KS = KS;
  st->b_bytes = (size_t)4 * KS * 2048;
  st->col_of_gauss.resize(N);
  std::vector<int32_t> sb_tile(1, 0), sb_pdf(1, 0);
  const int sb_cols = kSbTiles * kTileN;
  int64_t col = 0;
  for (int pdf = 0; pdf < P; pdf++) {
    const int M = po[pdf + 1] - po[pdf];
    const int64_t sb_start = (int64_t)sb_tile.back() * kTileN;
    if (col - sb_start + M > sb_cols) {  // close the super-block at the next panel boundary
      col = (col + kTileN - 1) / kTileN * kTileN;
      sb_tile.push_back((int32_t)(col / kTileN));
      sb_pdf.push_back(pdf);
    }
    for (int m = 0; m < M; m++) st->col_of_gauss[po[pdf] + m] = (int32_t)(col + m);
    col += M;
  }
  const int n_tiles = (int)((col + kTileN - 1) / kTileN);
  sb_tile.push_back(n_tiles);
  sb_pdf.push_back(P);
  st->n_tiles = n_tiles;
  st->n_sb = (int)sb_tile.size() - 1;
  // B image + segment tables
  st->h_bimg.assign((size_t)n_tiles * st->b_bytes, 0);
  std::vector<uint8_t> tabs((size_t)n_tiles * kTabBytes, 0);
  std::vector<int32_t> gauss_of_col((size_t)n_tiles * kTileN, -1);
  for (int g = 0; g < N; g++) gauss_of_col[st->col_of_gauss[g]] = g;
  std::vector<char> is_end((size_t)n_tiles * kTileN, 0);
  std::vector<char> is_start((size_t)n_tiles * kTileN, 0);
  for (int pdf = 0; pdf < P; pdf++) is_end[st->col_of_gauss[po[pdf + 1] - 1]] = 1, is_start[st->col_of_gauss[po[pdf]]] = 1;
  st->gshift.assign(N, 0.0);
  const double L2E = 1.4426950408889634074;
  bool ok = true;
  for (int64_t cc = 0; cc < (int64_t)n_tiles * kTileN && ok; cc++) {
    const int g = gauss_of_col[cc];
    double gc = kDummy;
    if (g >= 0) {
      double shift = 0.0;
      for (int d = 0; d < D; d++) {
        const double v = iv[(size_t)g * stride + d], mv = miv[(size_t)g * stride + d];
        const double b1 = (mv - v * c[d]) * L2E / (double)s1[d], b2 = -0.5 * v * L2E / (double)s2[d];
        if (std::fabs(b1) > 60000.0 || std::fabs(b2) > 60000.0) ok = false;
        put_half_pair(st->h_bimg.data(), st->b_bytes, KS, (int)cc, 2 * d, b1);
        put_half_pair(st->h_bimg.data(), st->b_bytes, KS, (int)cc, 2 * d + 1, b2);
        shift += mv * c[d] - 0.5 * v * c[d] * c[d];
      }
      st->gshift[g] = shift;
      gc = std::max(((double)gconsts[g] + shift) * L2E, (double)kDummy);
      if (!(gc < 60000.0)) ok = false;
    }
    put_half_pair(st->h_bimg.data(), st->b_bytes, KS, (int)cc, 2 * D, gc);
  }
  if (!ok) {
    delete st;
    return 0;
  }
  // part tables: the runs of columns of one pdf in every panel, cut so that (a) a part has at most 16 columns and
  // (b) the power-of-two load that covers it stays inside the panel; grouped by owner class ((pdf >> 2) & 3), column
  // order inside a class.  Padding columns sit at the tail of a panel and are not listed.
  {
    std::vector<int32_t> pdf_of_col((size_t)n_tiles * kTileN, -1);
    for (int pdf = 0; pdf < P; pdf++)
      for (int g = po[pdf]; g < po[pdf + 1]; g++) pdf_of_col[st->col_of_gauss[g]] = pdf;
    for (int t = 0; t < n_tiles; t++) {
      std::vector<uint32_t> cls_parts[4];
      int first_pdf = -1;
      int n = 0;
      while (n < kTileN) {
        const int pdf = pdf_of_col[(size_t)t * kTileN + n];
        if (pdf < 0) break;
        if (first_pdf < 0) first_pdf = pdf;
        int run = 1;
        while (n + run < kTileN && pdf_of_col[(size_t)t * kTileN + n + run] == pdf) run++;
        const bool pdf_ends_here = is_end[(size_t)t * kTileN + n + run - 1] != 0;
        int c = n, left = run;
        while (left > 0) {
          int len = std::min(left, 16);
          while (c + ld_width(len) > kTileN) len--;  // len = 1 always fits
          const bool ends = pdf_ends_here && len == left;
          const bool cont = c > n || !is_start[(size_t)t * kTileN + n];  // an earlier part of this pdf exists
          cls_parts[(pdf >> 2) & 3].push_back((uint32_t)c | ((uint32_t)len << 8) | ((uint32_t)(pdf - first_pdf) << 16) |
                                              (ends ? kPartEnds : 0u) | (cont ? kPartCont : 0u));
          c += len;
          left -= len;
        }
        n += run;
      }
      uint32_t *tb = reinterpret_cast<uint32_t *>(tabs.data() + (size_t)t * kTabBytes);
      uint32_t k = 0;
      for (int c4 = 0; c4 < 4; c4++) {
        tb[2 * c4] = k | ((uint32_t)cls_parts[c4].size() << 16);
        tb[2 * c4 + 1] = (uint32_t)std::max(first_pdf, 0);
        for (uint32_t e : cls_parts[c4]) tb[8 + k++] = e;
      }
    }
  }
  int rc = 0;
  auto up = [&](DevBuf &b, const void *src, size_t bytes) {
    if (rc == 0) rc = b.reserve(bytes);
    if (rc == 0 && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(VBGPU_ERR_CUDA, "upload of the tensor-core model image failed");
  };
  up(st->d_bimg, st->h_bimg.data(), st->h_bimg.size());
  up(st->d_tabs, tabs.data(), tabs.size());
  up(st->d_centre, cf.data(), D * 4);
  up(st->d_s1, s1.data(), D * 4);
  up(st->d_s2, s2.data(), D * 4);
  up(st->d_sb_tile, sb_tile.data(), sb_tile.size() * 4);
  up(st->d_sb_pdf, sb_pdf.data(), sb_pdf.size() * 4);
  h->tc = st;
  if (rc < 0) score_tc_release(h);
  return rc;
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
      return 0;
    }
  }
  const double L2E = 1.4426950408889634074;
  for (int g = 0; g < h->N; g++) {
    const double gc = std::max(((double)gconsts[g] + st->gshift[g]) * L2E, (double)kDummy);
    if (!(gc < 60000.0)) {
      score_tc_release(h);
      return 0;
    }
    put_half_pair(st->h_bimg.data(), st->b_bytes, st->KS, st->col_of_gauss[g], 2 * h->D, gc);
  }
  VB_CUDA(cudaMemcpy(st->d_bimg.p, st->h_bimg.data(), st->h_bimg.size(), cudaMemcpyHostToDevice));
  return 0;
}

int score_tc_launch(vbgpu_gmm_t h, const float *d_feats, int64_t T, int32_t stride, float *d_ll, int32_t ll_stride,
                    cudaStream_t s) {
  TcState *st = static_cast<TcState *>(h->tc);
  if (!st) return fail(VBGPU_ERR_INVALID, "tensor-core scorer unavailable for this model");
  if (T == 0) return 0;
  const int sms = num_sms(h->device);
  const int64_t n_mtiles = (T + kMt * kRowsMt - 1) / (kMt * kRowsMt);
  // split the panels over CTAs when there are too few frame tiles: pick the split with the best last-wave fill
  int best = 1;
  int64_t n_whole = 0;
  if (n_mtiles < 4LL * sms) {
    double best_eff = 0.0;
    const int max_split = std::min(st->n_sb, 64);
    for (int k = 1; k <= max_split; k++) {
      const int64_t units = n_mtiles * k, waves = (units + sms - 1) / sms;
      // each extra split repeats the A-panel build: charge it as ~2 panels of work per unit
      const double work = (double)st->n_tiles / k + 2.0;
      const double eff = ((double)st->n_tiles / k) / work * (double)units / (double)(waves * sms);
      if (eff > best_eff * 1.02) best_eff = eff, best = k;
    }
  } else {
    // many tiles: whole tiles for the full waves; the r tiles of the last, partial wave are cut into floor(sms / r) panel
    // ranges each so that the tail runs on (almost) every SM for 1/k of a tile's time instead of on r SMs for all of it
    const int64_t r = n_mtiles % sms;
    int k = r > 0 ? (int)std::min<int64_t>(std::min(st->n_sb, 16), sms / r) : 1;
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
  p.tabs = st->d_tabs.as<uint8_t>();
  p.centre = st->d_centre.as<float>();
  p.s1 = st->d_s1.as<float>();
  p.s2 = st->d_s2.as<float>();
  p.sb_tile = st->d_sb_tile.as<int32_t>();
  p.sb_pdf = st->d_sb_pdf.as<int32_t>();
  p.n_sb = st->n_sb;
  p.n_splits = best;
  p.n_whole = n_whole;
  p.n_units = n_whole + (n_mtiles - n_whole) * best;
  p.out = d_ll;
  p.ll_stride = ll_stride;
  const int padded = (h->D + 3) / 4 * 4;
  p.vec_ok = (((reinterpret_cast<uintptr_t>(d_ll) & 15) == 0 && ll_stride % 4 == 0) ? 1 : 0) |
             (((reinterpret_cast<uintptr_t>(d_feats) & 15) == 0 && stride % 4 == 0 && stride >= padded) ? 2 : 0);
  p.bad = h->d_bad.as<unsigned long long>();
  p.lbo = st->lbo;
  p.sbo = st->sbo;
  const char *dbg_env = getenv("VBGPU_TC_DEBUG");
  p.dbg = dbg_env ? (uint32_t)atoi(dbg_env) : 0u;
  const int grid = (int)std::min<int64_t>(p.n_units, sms);
  switch (st->KS) {
    case 2: return launch_ks<2>(p, st, grid, s);
    case 3: return launch_ks<3>(p, st, grid, s);
    case 4: return launch_ks<4>(p, st, grid, s);
    case 5: return launch_ks<5>(p, st, grid, s);
    case 6: return launch_ks<6>(p, st, grid, s);
    default: return fail(VBGPU_ERR_INVALID, "unsupported K for the tensor-core scorer");
  }
}

}  // namespace vb
