// gemm.cu -- K3/K4: the Linear layers of the acoustic model as tcgen05 tensor-core GEMMs.
//
// Replaces, fused in one kernel per layer,
//   SpliceLayer + NarrowLayer   src/nnet.cc:50-75,182-202   (implicit: one K-slab per tap, the
//                                                            same activation rows at a row offset)
//   LinearLayer                 src/nnet.cc:22-36 -> MatMat src/matrix.cc:300-323 (cblas_sgemm)
//   Quantize'd variant          MatMat_U8U8F32 src/matrix.cc:389-420 -> gemmlowp
//                               eight_bit_int_gemm.cc:107-141,383-391, internal/unpack.h:118-125
//   bias / ReLU / BatchNorm     src/nnet.cc:34,149-160,106-117
//   FindMinMax of the result    src/matrix.cc:329-345 (input of the next layer's Quantize)
//
//   LogSoftmax + prior + argmax src/nnet.cc:137-146, src/vector.cc:110-122, src/am.cc:109-112 (kModeLsm: fused
//                               into the output layer's epilogue, int8 models)
//
// Structure (Blackwell-native): persistent CTAs, one per SM (a pair per TPC with cta_group::2):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D tiles (128B swizzle) of A [128 rows x 128 B] and the
//               CTA's share of B into a 5-stage shared-memory ring, mbarrier-signalled;
//   warp 1      MMA issuer: one thread issues tcgen05.mma (cta_group::2: M = 256 over the pair, N = 256,
//               K = 32 bytes per instruction) with the accumulator in tensor memory; two 256-column
//               accumulator stages so tile i+1 is multiplied while tile i drains;
//   warps 2..   epilogue: tcgen05.ld of the accumulator (one TMEM lane quadrant per warp), the exact fp32 chain
//               of the reference (un-fused multiplies/adds, its order), swizzled staging tile, TMA store.
//               kModeClassic: 8 warps on 32-column chunks (every kind / output format); kModeWide / kModeLsm:
//               16 warps on 16-column pieces with packed f32x2 arithmetic (int8);
//   last warp   parameter prefetcher: per-tile epilogue parameters one tile ahead in two shared-memory slots.
// Data paths: kind::i8 (u8 x u8 -> s32, the zero-point algebra of gemmlowp applied to the exact raw sums),
// kind::f16 (bf16 -> fp32; bf16x3 = hi/lo split operands) and kind::tf32 (optionally three passes hi*hi +
// hi*lo + lo*hi for fp32-class accuracy).

#include "gemm.h"

#include <cuda_runtime.h>
#include <float.h>

#include <mutex>

namespace ce {
namespace {

constexpr int kABytes = kTileM * kTileKBytes;            // 16 KB: 128 rows of A per CTA per stage
// Warps of a CTA: warp 0 TMA, warp 1 MMA, then the epilogue warps (Cfg::kEpiWarps: 2 per TMEM lane
// quadrant, one per column half of the tile; 4 per quadrant in the fused output layer, whose epilogue is
// most of its work), then the epilogue's parameter prefetcher.
constexpr int kOutTileBytes = 32 * 128;                  // 32 rows x 128 B staged per TMA store
// Per-tile epilogue parameters, staged by the prefetch warp one tile ahead (two slots): bias, bn scale,
// bn offset, the per-column integer correction (x4 in granule mode: one per 32-row quadrant), then per
// accumulator row of this CTA the row's integer correction and its FindMinMax flag, then per quadrant
// c_scale and the utterance.
constexpr int kParamCols = 8 * kTileN;                   // words of per-column arrays (slot 7: log prior, LSM)
constexpr int kParamPrior = 7 * kTileN;
constexpr int kParamRowCorr = kParamCols;                // int32[128]
constexpr int kParamRowFlag = kParamRowCorr + kTileM;    // int32[128]
constexpr int kParamScale = kParamRowFlag + kTileM;      // float[4]
constexpr int kParamUtt = kParamScale + 4;               // int32[4]
constexpr int kParamSlotWords = kParamUtt + 4 + 24;      // padded to a multiple of 32 words
constexpr int kParamSlots = 2;
constexpr int kParamBytes = kParamSlots * kParamSlotWords * 4;
constexpr int kXchgBytes = 2 * 4 * kTileM * 8;              // LSM: [stats | argmax][column part][row] 8-byte pairs
constexpr int kTmemCols = 512;
constexpr int kAccStages = 2;

// CG = 1: one CTA computes a 128 x 256 tile.  CG = 2 (cta_group::2): a pair of CTAs on one TPC
// computes a 256 x 256 tile with one tcgen05.mma M=256 -- each CTA stages its own 128 rows of A
// and HALF of the B tile (128 of the 256 weight rows), so per CTA the L2 -> smem traffic and the
// shared-memory operand reads of the tensor core drop by a third, and the smaller stages leave
// room for a deeper ring.
template <int CG, bool WIDE = false>                    // WIDE: the 16-warp epilogue (kModeLsm, kModeWide)
struct Cfg {
  static constexpr int kEpiWarps = WIDE ? 16 : 8;
  static constexpr int kThreads = 64 + 32 * kEpiWarps + 32;
  static constexpr int kParts = kEpiWarps / 4;           // column parts of a tile, one epilogue warp each
  static constexpr int kPartCols = kTileN / kParts;
  static constexpr int kStgBytes = WIDE ? 32 * 64 : kOutTileBytes;   // staging per epilogue warp
  static constexpr int kMaxRegs = WIDE ? 96 : 128;       // warps (rounded up to 4) x 32 x kMaxRegs <= 64 K registers
  static constexpr int kStages = (CG == 2) ? 5 : 3;
  static constexpr int kBRows = kTileN / CG;             // weight rows staged per CTA
  static constexpr int kBBytes = kBRows * kTileKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOffStage = kStages * kStageBytes;   // output staging (1024-aligned)
  static constexpr int kOffParams = kOffStage + kEpiWarps * kStgBytes;
  static constexpr int kOffXchg = kOffParams + kParamBytes;    // LSM: row statistics / argmax exchange
  static constexpr int kOffBars = kOffXchg + kXchgBytes;
  static constexpr int kSmemBytes = 1024 /*alignment slack*/ + kOffBars + 256;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(2 * kStages + 2 * kAccStages + 2 * kParamSlots + 1 <= 32, "barrier block");
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits: a protocol bug becomes a trap (reported as a CUDA error), never a hung GPU.
// mbar_wait: for the warps whose wake-up latency is on the critical path (MMA issuer, epilogue); the
// clock is read once per 256 failed tries, so the polling loop is three instructions.
// mbar_wait_relaxed: for the warps that run ahead by design (TMA producer, parameter prefetcher): the
// try carries a suspend-time hint, the warp sleeps in hardware instead of taking issue slots from the
// epilogue warps of its scheduler (ncu: the polling loops were a fifth of the instructions of an
// epilogue-bound launch).
__device__ __forceinline__ bool mbar_try_wait_suspend(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(400u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 255u) == 0 && clock64() - t0 > 20000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_suspend(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// The same with an L2 evict-first policy: a stream of results nobody on this GPU reads again must not
// push the operands out of L2.
__device__ __forceinline__ void tma_store_2d_stream(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {   // smem of all my stores has been read
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- cluster (cta_group::2) helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank`.
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Relaxed: this arrive only hands a TMEM accumulator stage back to the MMA thread -- no memory written
// by this warp is consumed through it (tcgen05.fence::before_thread_sync orders the TMEM reads), and
// a release at cluster scope costs a full memory barrier per tile and warp.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data lands in THIS CTA's smem, the bytes are counted on the leader's barrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
// tcgen05.commit of the pair: arrives on the barrier at this smem offset in BOTH CTAs.
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(bar), "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

#define CE_TC_MMA(GROUP, KINDSTR)                                                          \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                            \
               "tcgen05.mma.cta_group::" GROUP ".kind::" KINDSTR " [%0], %1, %2, %3, p;\n\t}"  \
               :                                                                          \
               : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)       \
               : "memory")

template <int KIND, int CG>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                       uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    if (KIND == kKindI8) CE_TC_MMA("1", "i8");
    else if (KIND == kKindBF16 || KIND == kKindBF16X3) CE_TC_MMA("1", "f16");
    else CE_TC_MMA("1", "tf32");
  } else {
    if (KIND == kKindI8) CE_TC_MMA("2", "i8");
    else if (KIND == kKindBF16 || KIND == kKindBF16X3) CE_TC_MMA("2", "f16");
    else CE_TC_MMA("2", "tf32");
  }
}

// 32 lanes x 32 columns of 32-bit accumulators: thread i gets lane (base + i), v[j] = column j.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 16 columns, asynchronous: the registers are valid after tmem_ld16_wait (which names them, so
// that no use can be scheduled above it).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),
                 "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO); LBO is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);          // start address
  d |= (uint64_t)1 << 16;                                // leading byte offset (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
  return d;
}

template <int KIND, int CG>
__device__ __forceinline__ uint32_t make_idesc() {
  const uint32_t c_fmt = (KIND == kKindI8) ? 2u : 1u;                 // S32 : F32
  const uint32_t ab_fmt = (KIND == kKindI8) ? 0u : (KIND == kKindTF32) ? 2u : 1u;   // U8, TF32, BF16
  return (c_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) |
         ((uint32_t)((kTileM * CG) >> 4) << 24);                      // M = 128 (one CTA) / 256 (pair)
}

__device__ __forceinline__ float ex2_approx(float x) {   // 2^x: MUFU.EX2, tiny results flush to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Two fp32 operations in one instruction (sm_100 add / sub / mul / fma .f32x2): each half is an ordinary
// IEEE round-to-nearest fp32 operation, so the results equal the scalar __fadd_rn / __fmul_rn chain.
#define CE_F32X2_OP(NAME, OP)                                                                      \
  __device__ __forceinline__ float2 NAME(float2 a, float2 b) {                                     \
    float2 r;                                                                                      \
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"       \
        OP " rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"                                           \
        : "=f"(r.x), "=f"(r.y)                                                                     \
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                                 \
    return r;                                                                                      \
  }
CE_F32X2_OP(add2, "add.rn.f32x2")
CE_F32X2_OP(sub2, "sub.rn.f32x2")
#undef CE_F32X2_OP
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ---------------------------------------------------------------------------
// epilogue math: 32 accumulator columns of one row, reference order of operations
// ---------------------------------------------------------------------------
struct RowConst {
  int32_t row_corr;     // zp_b * sum_k A[row][k]   (u8 path)
  float c_scale;        // scale_a * scale_b          (u8 path)
};

template <int KIND, bool RELU, bool BN, bool MM>
__device__ __forceinline__ void epi_math(const uint32_t (&raw)[32], float (&v)[32], const float *sp,
                                         int corr_off, int pcol, const RowConst rc, float &vmin,
                                         float &vmax) {
  const float4 *b4 = reinterpret_cast<const float4 *>(sp + pcol);
  const float4 *s4 = reinterpret_cast<const float4 *>(sp + kTileN + pcol);
  const float4 *o4 = reinterpret_cast<const float4 *>(sp + 2 * kTileN + pcol);
  const int4 *c4 = reinterpret_cast<const int4 *>(sp + corr_off + pcol);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 bb = b4[q];
    float4 ss = make_float4(1.f, 1.f, 1.f, 1.f), oo = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BN) {
      ss = s4[q];
      oo = o4[q];
    }
    int4 cc = make_int4(0, 0, 0, 0);
    if (KIND == kKindI8) cc = c4[q];
    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
    const float sv[4] = {ss.x, ss.y, ss.z, ss.w};
    const float ov[4] = {oo.x, oo.y, oo.z, oo.w};
    const int cv[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * q + e;
      float x;
      if (KIND == kKindI8) {
        // acc = sum (A - zpA)(B - zpB) = raw - zpB*rowsum(A) - zpA*colsum(B) + K*zpA*zpB
        const int32_t a = (int32_t)raw[j] + (cv[e] - rc.row_corr);
        x = __fmul_rn(__int2float_rn(a), rc.c_scale);                 // eight_bit_int_gemm.cc:389
      } else {
        x = __uint_as_float(raw[j]);
      }
      x = __fadd_rn(x, bv[e]);                                         // nnet.cc:34
      if (RELU) {
        // nnet.cc:156 `if (x < 0) x = 0`: a NaN stays a NaN, which is max.NaN (one instruction instead of
        // compare + select; -0.0 becomes +0.0, equal under every comparison the path makes)
        asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(x) : "f"(x));
      }
      if (BN) {
        x = __fmul_rn(x, sv[e]);                                       // nnet.cc:114
        x = __fadd_rn(x, ov[e]);                                       // nnet.cc:115
      }
      if (MM) {
        vmin = fminf(vmin, x);                                         // NaNs ignored, as the
        vmax = fmaxf(vmax, x);                                         // comparisons of matrix.cc:337-340
      }
      v[j] = x;
    }
  }
}

// ---------------------------------------------------------------------------
// fused output layer (LSM): 16 accumulator columns of one row at a time
// ---------------------------------------------------------------------------
// The layer's value in the reference's order of operations (as epi_math), two columns per instruction.
// (An output layer with ReLU / BatchNorm behind it, or the accumulator dump, takes the un-fused path.)
template <int KIND>
__device__ __forceinline__ void lsm_value(const uint32_t (&raw)[16], const float *sp, int corr_off, int pcol,
                                          const RowConst rc, float neg_zero, float2 (&v2)[8]) {
  const float2 nz2 = make_float2(neg_zero, neg_zero);
  const float4 *b4 = reinterpret_cast<const float4 *>(sp + pcol);
  const int4 *c4 = reinterpret_cast<const int4 *>(sp + corr_off + pcol);
  const float2 sc2 = make_float2(rc.c_scale, rc.c_scale);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 bb = b4[q];
    float2 x0, x1;
    if (KIND == kKindI8) {
      const int4 cc = c4[q];
      x0 = make_float2(__int2float_rn((int32_t)raw[4 * q] + (cc.x - rc.row_corr)),
                       __int2float_rn((int32_t)raw[4 * q + 1] + (cc.y - rc.row_corr)));
      x1 = make_float2(__int2float_rn((int32_t)raw[4 * q + 2] + (cc.z - rc.row_corr)),
                       __int2float_rn((int32_t)raw[4 * q + 3] + (cc.w - rc.row_corr)));
      // the product, rounded on its own (eight_bit_int_gemm.cc:389): x * s + (-0.0) is exactly RN(x * s).
      // ptxas contracts mul.rn.f32x2 + add.rn.f32x2 (and fma with a literal -0.0 + add) into ONE FFMA2,
      // which would skip that rounding -- so the -0.0 comes from a kernel argument it cannot see through.
      x0 = fma2(x0, sc2, nz2);
      x1 = fma2(x1, sc2, nz2);
    } else {
      x0 = make_float2(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]));
      x1 = make_float2(__uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3]));
    }
    v2[2 * q] = add2(x0, make_float2(bb.x, bb.y));                       // nnet.cc:34
    v2[2 * q + 1] = add2(x1, make_float2(bb.z, bb.w));
  }
}

__device__ __forceinline__ float min16(const float2 (&v2)[8]) {
  const float p0 = fminf(fminf(v2[0].x, v2[0].y), fminf(v2[1].x, v2[1].y));
  const float p1 = fminf(fminf(v2[2].x, v2[2].y), fminf(v2[3].x, v2[3].y));
  const float p2 = fminf(fminf(v2[4].x, v2[4].y), fminf(v2[5].x, v2[5].y));
  const float p3 = fminf(fminf(v2[6].x, v2[6].y), fminf(v2[7].x, v2[7].y));
  return fminf(fminf(p0, p1), fminf(p2, p3));
}
__device__ __forceinline__ float max16(const float2 (&v2)[8]) {
  const float p0 = fmaxf(fmaxf(v2[0].x, v2[0].y), fmaxf(v2[1].x, v2[1].y));
  const float p1 = fmaxf(fmaxf(v2[2].x, v2[2].y), fmaxf(v2[3].x, v2[3].y));
  const float p2 = fmaxf(fmaxf(v2[4].x, v2[4].y), fmaxf(v2[5].x, v2[5].y));
  const float p3 = fmaxf(fmaxf(v2[6].x, v2[6].y), fmaxf(v2[7].x, v2[7].y));
  return fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
}

// First sweep: online log-sum-exp over the row's columns in the fixed order the tiles arrive in
// (m: running maximum, s: sum of exp(x - m)); exp(x - m) = 2^(x log2e - m log2e), one FFMA and one
// MUFU.EX2 per element.  n_ok (< 16 only in the last piece of a ragged row) is a multiple of 4.
__device__ __forceinline__ void lsm_stats(float2 (&v2)[8], int n_ok, float &m, float &s) {
  const float kLog2e = 1.4426950408889634f;
  if (n_ok < 16) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (2 * j >= n_ok) v2[j] = make_float2(-FLT_MAX, -FLT_MAX);
  }
  const float cmax = max16(v2);
  if (cmax > m) {
    s *= ex2_approx((m - cmax) * kLog2e);
    m = cmax;
  }
  const float nm = -m * kLog2e;
  const float2 nm2 = make_float2(nm, nm), l2 = make_float2(kLog2e, kLog2e);
  float2 a[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 e = fma2(v2[j], l2, nm2);
    a[j] = make_float2(ex2_approx(e.x), ex2_approx(e.y));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 e = fma2(v2[4 + j], l2, nm2);
    a[j] = add2(a[j], make_float2(ex2_approx(e.x), ex2_approx(e.y)));
  }
  const float2 t = add2(add2(a[0], a[1]), add2(a[2], a[3]));
  s += t.x + t.y;
}

// Second sweep: the finished row (x - logsumexp) - log prior, and the row's first maximum.
__device__ __forceinline__ void lsm_finish(float2 (&v2)[8], const float *prior, bool softmax, float lse, int n_ok,
                                           int col0, float &best, int &best_i) {
  const float4 *lp4 = reinterpret_cast<const float4 *>(prior);
  const float2 lse2 = make_float2(lse, lse);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 lp = lp4[q];
    float2 x0 = v2[2 * q], x1 = v2[2 * q + 1];
    if (softmax) {                                                       // x -= max + log(sum)    vector.cc:120
      x0 = sub2(x0, lse2);
      x1 = sub2(x1, lse2);
    }
    v2[2 * q] = sub2(x0, make_float2(lp.x, lp.y));                       // AddVec(-1, log_prior_) am.cc:111
    v2[2 * q + 1] = sub2(x1, make_float2(lp.z, lp.w));
  }
  if (n_ok < 16) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (2 * j >= n_ok) v2[j] = make_float2(-FLT_MAX, -FLT_MAX);
  }
  // the piece's maximum first; its position only when it beats the row's best so far (first maximum wins:
  // columns ascend, and only a strictly larger value replaces the best)
  const float cmax = max16(v2);
  if (cmax > best) {
    best = cmax;
    int at = 15;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      if (v2[j].y == cmax) at = 2 * j + 1;
      if (v2[j].x == cmax) at = 2 * j;
    }
    best_i = col0 + at;
  }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
// MODE: kModeClassic = 8 epilogue warps on 32-column chunks (every kind and output format);
//       kModeLsm     = the output layer fused with LogSoftmax + prior + argmax (16 epilogue warps);
//       kModeWide    = a plain int8 layer (fp32 result, bias / ReLU / BatchNorm / FindMinMax) with the fused
//                      layer's epilogue: 16 warps, 16-column pieces, packed arithmetic, 64-byte TMA stores.
constexpr int kModeClassic = 0, kModeLsm = 1, kModeWide = 2;
template <int KIND, int CG, bool GRAN = false, int MODE = kModeClassic>
__global__ void __maxnreg__((Cfg<CG, MODE != kModeClassic>::kMaxRegs))
gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_b1,
            const __grid_constant__ CUtensorMap map_o0, const __grid_constant__ CUtensorMap map_o1,
            const GemmArgs p) {
  constexpr bool LSM = MODE == kModeLsm;
  using C = Cfg<CG, MODE != kModeClassic>;
  constexpr int kEpiWarps = C::kEpiWarps;
  constexpr int kStages = C::kStages;
  constexpr int kBBytes = C::kBBytes;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by OFFSET from the __shared__ array, not by integer arithmetic on a generic
  // pointer: the compiler then knows every access below is shared memory and emits LDS / STS instead
  // of generic loads and stores (which also wait on the long scoreboard).
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *smem_a = smem;                                   // [kStages][kABytes]
  unsigned char *smem_b = smem + kStages * kABytes;               // [kStages][kBBytes]
  unsigned char *smem_out = smem + C::kOffStage;                  // [kEpiWarps][kOutTileBytes]
  float *sp_all = reinterpret_cast<float *>(smem + C::kOffParams);   // [kParamSlots][kParamSlotWords]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kOffBars);
  // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base slot.
  // With CG == 2 the same block exists in both CTAs; `full` and `tmem_empty` are used in the
  // leader (cluster rank 0) only, `empty` and `tmem_full` in both (multicast commits).
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages);
  const uint32_t bar_tempty = smem_u32(bars + 2 * kStages + kAccStages);
  const uint32_t bar_pfull = smem_u32(bars + 2 * kStages + 2 * kAccStages);      // parameter slots
  const uint32_t bar_pempty = bar_pfull + 8 * kParamSlots;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2 * kAccStages + 2 * kParamSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;      // position in the CTA pair
  const bool leader = rank == 0;
  const int group_id = blockIdx.x / CG;                          // tile-processing unit (CTA or pair)
  const int n_groups = gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kEpiWarps * CG);
    }
    for (int i = 0; i < kParamSlots; ++i) {
      mbar_init(bar_pfull + 8 * i, 1);
      mbar_init(bar_pempty + 8 * i, kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                       // the peer's barriers exist before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr int kGroupM = kTileM * CG;                   // rows per tile of the processing unit
  const int m_tiles = (p.M + kGroupM - 1) / kGroupM;
  const int n_tiles = (p.N + kTileN - 1) / kTileN;
  const int total_tiles = m_tiles * n_tiles;
  constexpr int kEltBytes = (KIND == kKindI8) ? 1 : (KIND == kKindTF32) ? 4 : 2;
  constexpr int kTileK = kTileKBytes / kEltBytes;
  const int kb_per_tap = p.c_pad / kTileK;
  const int steps_per_pass = p.n_taps * kb_per_tap;
  const int n_steps = p.n_pass * steps_per_pass;
  // Work units.  Plain GEMM: one tile (m, n) per unit, units round-robin over the CTA groups.  LSM (the
  // output layer fused with LogSoftmax / prior / argmax): one unit = ALL n tiles of one m tile, walked
  // twice when the softmax is on -- a first sweep that only reduces every row's maximum and sum of
  // exponentials, a second one that recomputes the accumulators and writes the finished rows -- so the
  // logits never exist in memory.
  const int n_units = LSM ? m_tiles : total_tiles;
  const int n_subs = LSM ? ((p.lsm_softmax && !p.lsm_single) ? 2 * n_tiles : n_tiles) : 1;

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own operand slices) ===============
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = (CG == 2) ? map_to_cta(bar_full, 0) : bar_full;   // leader's barriers
      for (int unit = group_id; unit < n_units; unit += n_groups)
      for (int sub = 0; sub < n_subs; ++sub) {
        const int mt = LSM ? unit : unit / n_tiles;
        const int nt = LSM ? (sub >= n_tiles ? sub - n_tiles : sub) : unit % n_tiles;
        const int m0 = mt * kGroupM + (int)rank * kTileM;
        const int n0 = nt * kTileN + (int)rank * C::kBRows;
        for (int ps = 0; ps < p.n_pass; ++ps) {
          const CUtensorMap *ma = p.pass_a[ps] ? &map_a1 : &map_a0;
          const CUtensorMap *mb = p.pass_b[ps] ? &map_b1 : &map_b0;
          for (int tap = 0; tap < p.n_taps; ++tap) {
            const int row = m0 + p.tap_off[tap];
            for (int kb = 0; kb < kb_per_tap; ++kb) {
#ifdef CE_PRODUCER_RELAXED
              mbar_wait_relaxed(bar_empty + 8 * stage, phase ^ 1);
#else
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
#endif
              if (p.debug & 4) {
                if (leader) mbar_arrive(bar_full + 8 * stage);
              } else if (CG == 1) {
                mbar_expect_tx(bar_full + 8 * stage, C::kStageBytes);
                tma_load_2d(smem_u32(smem_a + stage * kABytes), ma, bar_full + 8 * stage, kb * kTileK, row);
                tma_load_2d(smem_u32(smem_b + stage * kBBytes), mb, bar_full + 8 * stage,
                            tap * p.c_pad + kb * kTileK, n0);
              } else {
                if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * C::kStageBytes);   // both CTAs' bytes
                tma_load_2d_pair(smem_u32(smem_a + stage * kABytes), ma, full0 + 8 * stage, kb * kTileK, row);
                tma_load_2d_pair(smem_u32(smem_b + stage * kBBytes), mb, full0 + 8 * stage,
                                 tap * p.c_pad + kb * kTileK, n0);
              }
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc<KIND, CG>();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = group_id; unit < n_units; unit += n_groups)
      for (int sub = 0; sub < n_subs; ++sub) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kTileN);
        for (int step = 0; step < n_steps; ++step) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t da = make_desc(smem_u32(smem_a + stage * kABytes));
          const uint64_t db = make_desc(smem_u32(smem_b + stage * kBBytes));
          if (KIND == kKindBF16X3) {
            // atom = [32 hi | 32 lo]: K steps 0,1 are hi, 2,3 lo.  Small terms first: lo*hi, hi*lo, hi*hi.
            constexpr int ka[6] = {2, 3, 0, 1, 0, 1};
            constexpr int kb[6] = {0, 1, 2, 3, 0, 1};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              tc_mma<KIND, CG>(tmem_d, da + (uint64_t)(2 * ka[i]), db + (uint64_t)(2 * kb[i]), idesc,
                               (step | i) != 0 ? 1u : 0u);
            }
          } else if (!(p.debug & 2)) {
#pragma unroll
            for (int k = 0; k < kTileKBytes / 32; ++k) {
              // +32 bytes along K inside the swizzle atom = +2 in the (addr >> 4) field
              tc_mma<KIND, CG>(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                               (step | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem slot (in both CTAs of a pair) when the MMAs retire
          if (CG == 1) tc_commit(bar_empty + 8 * stage); else tc_commit_pair(bar_empty + 8 * stage);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete (in both CTAs of a pair)
        if (CG == 1) tc_commit(bar_tfull + 8 * acc); else tc_commit_pair(bar_tfull + 8 * acc);
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                            // which part of the tile's 256 columns (128 each; LSM: 64)
    unsigned char *stg = smem_out + ew * C::kStgBytes;
    const uint32_t stg_u32 = smem_u32(stg);
    const int flags = (p.relu ? 1 : 0) | (p.bn_scale ? 2 : 0) | (p.minmax ? 4 : 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int pslot = 0;                                       // parameter slot of the tile being drained
    uint32_t pphase = 0;
    // CE_GPU_GEMM_PROF: where an epilogue warp's cycles go, summed over the grid into p.dbg[0..7]:
    // waiting for parameters, waiting for a full accumulator, tcgen05.ld, math, staging + store, the tile's
    // tail (TMEM hand-back, min/max reduction), the whole loop, tiles
    // (compiled in only with -DCE_GEMM_PROF: the counters cost the epilogue registers)
#ifdef CE_GEMM_PROF
    const bool prof_on = p.dbg != nullptr;
#else
    constexpr bool prof_on = false;
#endif
    long long pr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long pr_t0 = prof_on ? clock64() : 0;
    const uint32_t tempty0 = (CG == 2) ? map_to_cta(bar_tempty, 0) : bar_tempty;
    // granule mode: every 32-row quadrant of the tile may belong to another utterance, so the
    // per-column integer correction (it contains the utterance's zero point) exists once per quadrant
    const int corr_off = (3 + (GRAN ? quad : 0)) * kTileN;
    // LSM: the row's running maximum / sum of exponentials (first sweep), its log-sum-exp and the best
    // (value, pdf) so far (second sweep); both column halves of a row meet through shared memory
    float lsm_m = -FLT_MAX, lsm_s = 0.0f, lsm_lse = 0.0f, lsm_best = -FLT_MAX;
    int lsm_best_i = 0x7fffffff;
    float2 *xchg_stats = reinterpret_cast<float2 *>(smem + C::kOffXchg);            // [4][kTileM]
    float2 *xchg_best = xchg_stats + 4 * kTileM;                                    // [4][kTileM]
    for (int unit = group_id; unit < n_units; unit += n_groups)
    for (int sub = 0; sub < n_subs; ++sub) {
      const int mt = LSM ? unit : unit / n_tiles;
      const int nt = LSM ? (sub >= n_tiles ? sub - n_tiles : sub) : unit % n_tiles;
      const int m0 = mt * kGroupM + (int)rank * kTileM;  // this CTA's 128 rows
      const int n0 = nt * kTileN;
      const int my_row = m0 + quad * 32 + lane;          // the accumulator row this thread reads

      // ---- per-tile parameters: staged one tile ahead by the prefetch warp (no global load and no CTA
      //      barrier on this path: the dependent tile -> utterance -> parameters loads used to cost the
      //      epilogue ~2.5 us a tile) ----
      long long tq = prof_on ? clock64() : 0;
      mbar_wait(bar_pfull + 8 * pslot, pphase);
      if (prof_on) pr[0] += clock64() - tq;
      const float *sp = sp_all + pslot * kParamSlotWords;
      const int32_t *spi = reinterpret_cast<const int32_t *>(sp);
      RowConst rc;
      rc.row_corr = spi[kParamRowCorr + quad * 32 + lane];
      rc.c_scale = sp[kParamScale + (GRAN ? quad : 0)];
      const int utt = spi[kParamUtt + (GRAN ? quad : 0)];
      const int spi_row_flag = spi[kParamRowFlag + quad * 32 + lane];    // LSM: the row's output row
      const bool use_row = spi_row_flag != 0;            // takes part in the fused FindMinMax
      float vmin = FLT_MAX, vmax = -FLT_MAX;

      tq = prof_on ? clock64() : 0;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      if (prof_on) pr[1] += clock64() - tq;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileN + half * C::kPartCols);

      if constexpr (MODE == kModeWide) {
        // ---- a plain int8 layer, 16 columns at a time (the fused output layer's epilogue without its
        //      two sweeps): value, ReLU, BatchNorm, this tile's share of FindMinMax, 64-byte TMA stores ----
        constexpr int kPieces = C::kPartCols / 16;
        const float neg_zero = __int_as_float((int)0x80000000 | p.lsm_zero);
        const float2 nz2 = make_float2(neg_zero, neg_zero);
        const int colh = n0 + half * C::kPartCols;       // first column of this warp's part of the tile
        const int n_piece = min(kPieces, (p.n_store - colh + 15) >> 4);   // warp-uniform
        const int sw = (lane >> 1) & 3;
        if (!(p.debug & 1)) {
#pragma unroll
          for (int k = 0; k < kPieces; ++k) {
            if (k < n_piece) {
              uint32_t r[16];
              tmem_ld16_issue(taddr + (uint32_t)(k * 16), r);
              tmem_ld16_wait(r);
              const int pcol = half * C::kPartCols + k * 16;
              const int col0 = n0 + pcol;
              float2 v2[8];
              lsm_value<KIND>(r, sp, corr_off, pcol, rc, neg_zero, v2);
              if (flags & 1) {                                           // nnet.cc:156 (see epi_math)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(v2[j].x) : "f"(v2[j].x));
                  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(v2[j].y) : "f"(v2[j].y));
                }
              }
              if (flags & 2) {                                           // nnet.cc:114-115, product rounded on its own
                const float4 *s4 = reinterpret_cast<const float4 *>(sp + kTileN + pcol);
                const float4 *o4 = reinterpret_cast<const float4 *>(sp + 2 * kTileN + pcol);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 ss = s4[q], oo = o4[q];
                  v2[2 * q] = add2(fma2(v2[2 * q], make_float2(ss.x, ss.y), nz2), make_float2(oo.x, oo.y));
                  v2[2 * q + 1] = add2(fma2(v2[2 * q + 1], make_float2(ss.z, ss.w), nz2), make_float2(oo.z, oo.w));
                }
              }
              const int n_ok = p.N - col0;               // < 16: columns [N, n_store) are the next layer's K padding
              if (flags & 4) {
                if (n_ok >= 16) {
                  vmin = fminf(vmin, min16(v2));
                  vmax = fmaxf(vmax, max16(v2));
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    if (2 * j < n_ok) {
                      vmin = fminf(vmin, v2[j].x);
                      vmax = fmaxf(vmax, v2[j].x);
                    }
                    if (2 * j + 1 < n_ok) {
                      vmin = fminf(vmin, v2[j].y);
                      vmax = fmaxf(vmax, v2[j].y);
                    }
                  }
                }
              }
              if (n_ok < 16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (2 * j >= n_ok) v2[j].x = 0.0f;
                  if (2 * j + 1 >= n_ok) v2[j].y = 0.0f;
                }
              }
              if (lane == 0) tma_store_wait_read();      // the previous piece's store has read the tile
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                *reinterpret_cast<float4 *>(stg + lane * 64 + ((q ^ sw) << 4)) =
                    make_float4(v2[2 * q].x, v2[2 * q].y, v2[2 * q + 1].x, v2[2 * q + 1].y);
              }
              fence_async_smem();
              __syncwarp();
              if (lane == 0) tma_store_2d(&map_o0, stg_u32, col0, m0 + quad * 32);
            }
          }
        }
      } else
      if constexpr (LSM) {
        // ---- output layer fused with LogSoftmax (src/nnet.cc:137-146 -> ApplyLogSoftMax, src/vector.cc:110-122),
        //      the prior (src/am.cc:109-112) and the per-frame argmax; 16 columns at a time, the next
        //      piece's tcgen05.ld in flight while this one is worked on ----
        constexpr int kPieces = C::kPartCols / 16;
        const bool single = p.lsm_single != 0;           // one sweep: plain result + the row's log-sum-exp
        const bool stats = p.lsm_softmax != 0 && !single && sub < n_tiles;
        const int out_row = spi_row_flag;                // where my accumulator row goes; -1: nowhere
        const float neg_zero = __int_as_float((int)0x80000000 | p.lsm_zero);
        const int colh = n0 + half * C::kPartCols;       // first column of this warp's part of the tile
        const int n_piece = min(kPieces, (p.N - colh + 15) >> 4);   // warp-uniform; N % 4 == 0
        // all 32 rows of this warp go to consecutive output rows (blocks never share a 32-row granule):
        // the staged pieces leave as TMA stores
        const bool full = __all_sync(0xffffffffu, out_row >= 0) && !(p.debug & 32);
        const int out_row0 = __shfl_sync(0xffffffffu, out_row, 0);
        float4 *const out4 = (p.debug & 8) ? nullptr : reinterpret_cast<float4 *>(p.out_f32);   // 8: timing probe
        const unsigned char *const stg_rd = stg + (lane >> 2) * 64 + (((lane & 3) ^ ((lane >> 3) & 3)) << 4);
        unsigned char *const stg_wr = stg + lane * 64;
        const int sw = (lane >> 1) & 3;
        auto piece = [&](const uint32_t(&raw)[16], const int k) {
          const int pcol = half * C::kPartCols + k * 16; // column within the tile
          const int col0 = n0 + pcol;
          const int n_ok = min(16, p.N - col0);          // < 16: the last piece of a ragged row
          float2 v2[8];
          lsm_value<KIND>(raw, sp, corr_off, pcol, rc, neg_zero, v2);
          if (stats) {
            if (!(p.debug & 16)) lsm_stats(v2, n_ok, lsm_m, lsm_s);
            else lsm_s += v2[0].x;
            return;
          }
          if (single) {
            float2 t2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t2[j] = v2[j];
            lsm_stats(t2, n_ok, lsm_m, lsm_s);
          } else {
            lsm_finish(v2, sp + kParamPrior + pcol, p.lsm_softmax != 0, lsm_lse, n_ok, col0, lsm_best, lsm_best_i);
          }
          if (out4) {
            // registers -> swizzled staging tile (32 rows x 16 columns) -> 16-byte row pieces, 8 rows per
            // warp store: every row goes to its own output row (the compact frame index of its
            // utterance), or nowhere
            if (lane == 0) tma_store_wait_read();        // the previous piece's store has read the tile
            __syncwarp();                                // ... and so have the plain stores' reads
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              *reinterpret_cast<float4 *>(stg_wr + ((q ^ sw) << 4)) =
                  make_float4(v2[2 * q].x, v2[2 * q].y, v2[2 * q + 1].x, v2[2 * q + 1].y);
            }
            if (full) {                                  // warp-uniform
              fence_async_smem();
              __syncwarp();
              if (lane == 0) tma_store_2d_stream(&map_o0, stg_u32, col0, out_row0);   // columns >= N are clipped
            } else {
              __syncwarp();
              // (the first and the last group of an utterance's block: rows 8 i + lane / 4 of the tile,
              //  16-byte piece lane % 4, each to its own output row or nowhere)
              const bool col_ok = 4 * (lane & 3) < n_ok;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int orow = __shfl_sync(0xffffffffu, out_row, 8 * i + (lane >> 2));
                const float4 val = *reinterpret_cast<const float4 *>(stg_rd + i * 512);
                if (orow >= 0 && col_ok)
                  __stcs(reinterpret_cast<float4 *>(p.out_f32 + (int64_t)orow * p.ld_out + col0) + (lane & 3), val);
              }
            }
          }
        };
        if (!(p.debug & 1)) {
#ifdef CE_LSM_PIPELINED_LD
          uint32_t r[2][16];
          if (n_piece > 0) tmem_ld16_issue(taddr, r[0]);
#pragma unroll
          for (int k = 0; k < kPieces; ++k) {
            if (k < n_piece) {
              tmem_ld16_wait(r[k & 1]);
              if (k + 1 < kPieces && k + 1 < n_piece) tmem_ld16_issue(taddr + (uint32_t)((k + 1) * 16), r[(k + 1) & 1]);
              piece(r[k & 1], k);
            }
          }
#else
          // (one piece in flight per warp: with four warps per scheduler the other warps cover the
          //  tcgen05.ld, and a second buffer costs 16 of the 96 registers)
#pragma unroll
          for (int k = 0; k < kPieces; ++k) {
            if (k < n_piece) {
              uint32_t r[16];
              tmem_ld16_issue(taddr + (uint32_t)(k * 16), r);
              tmem_ld16_wait(r);
              piece(r, k);
            }
          }
#endif
        }
      } else
      for (int c = 0; c < 4; ++c) {
        const int pcol = half * 128 + c * 32;            // column within the tile
        const int col0 = n0 + pcol;
        if (col0 >= p.n_store) break;                    // warp-uniform
        uint32_t raw[32];
        const long long tc0 = prof_on ? clock64() : 0;
        tmem_ld32(taddr + (uint32_t)(c * 32), raw);
        const long long tc1 = prof_on ? clock64() : 0;
        if (p.debug & 1) continue;
        float v[32];
        float cmin = FLT_MAX, cmax = -FLT_MAX;           // this chunk's share of FindMinMax
        switch (flags) {
          case 0: epi_math<KIND, false, false, false>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          case 1: epi_math<KIND, true, false, false>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          case 2: epi_math<KIND, false, true, false>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          case 3: epi_math<KIND, true, true, false>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          case 4: epi_math<KIND, false, false, true>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          case 5: epi_math<KIND, true, false, true>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          case 6: epi_math<KIND, false, true, true>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
          default: epi_math<KIND, true, true, true>(raw, v, sp, corr_off, pcol, rc, cmin, cmax); break;
        }
        if (col0 + 32 > p.N || p.out_acc) {              // rare: ragged N, or the debug dump
          cmin = FLT_MAX;
          cmax = -FLT_MAX;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (KIND == kKindI8 && p.out_acc && my_row < p.M && col < p.N) {
              const int32_t cj = reinterpret_cast<const int32_t *>(sp)[corr_off + pcol + j];
              p.out_acc[(int64_t)my_row * p.ld_out + col] = (int32_t)raw[j] + (cj - rc.row_corr);
            }
            if (col >= p.N) {
              v[j] = 0.0f;                               // K padding of the next layer
            } else {
              cmin = fminf(cmin, v[j]);
              cmax = fmaxf(cmax, v[j]);
            }
          }
        }
        vmin = fminf(vmin, cmin);
        vmax = fmaxf(vmax, cmax);

        // ---- registers -> swizzled staging tile -> one TMA store per 32 x 128 B ----
        const long long tc2 = prof_on ? clock64() : 0;
        if (prof_on) {
          pr[2] += tc1 - tc0;
          pr[3] += tc2 - tc1;
        }
        const int sw = lane & 7;
        if (KIND == kKindBF16X3 && p.out_bf16) {
          // the next layer's operand: 32 columns -> one 128-byte atom [32 hi | 32 lo] per row
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t wh[4], wl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = v[8 * q + 2 * e], x1 = v[8 * q + 2 * e + 1];
              const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
              const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
              const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
              wh[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              wl[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            *reinterpret_cast<uint4 *>(stg + lane * 128 + ((q ^ sw) << 4)) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            *reinterpret_cast<uint4 *>(stg + lane * 128 + (((4 + q) ^ sw) << 4)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&map_o0, stg_u32, 2 * col0, m0 + quad * 32);
        } else if (KIND == kKindBF16 && p.out_bf16) {
          if ((c & 1) == 0) {                            // a new 64-column tile: previous store must
            if (lane == 0) tma_store_wait_read();        // have finished reading the buffer
            __syncwarp();
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
              w[e] = *reinterpret_cast<uint32_t *>(&h2);
            }
            const int piece = (c & 1) * 4 + q;
            *reinterpret_cast<uint4 *>(stg + lane * 128 + ((piece ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          const bool flush = (c & 1) == 1 || col0 + 32 >= p.n_store;
          if (flush) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(&map_o0, stg_u32, col0 - (c & 1) * 32, m0 + quad * 32);
          }
        } else {
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = (KIND == kKindTF32 && p.round_tf32) ? round_tf32(v[j]) : v[j];
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            *reinterpret_cast<float4 *>(stg + lane * 128 + ((q ^ sw) << 4)) =
                make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&map_o0, stg_u32, col0, m0 + quad * 32);
          if (KIND == kKindTF32 && p.out_lo) {           // 3xTF32 consumers: the exact remainder
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              *reinterpret_cast<float4 *>(stg + lane * 128 + ((q ^ sw) << 4)) =
                  make_float4(v[4 * q] - h[4 * q], v[4 * q + 1] - h[4 * q + 1], v[4 * q + 2] - h[4 * q + 2],
                              v[4 * q + 3] - h[4 * q + 3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(&map_o1, stg_u32, col0, m0 + quad * 32);
          }
        }
        if (prof_on) pr[4] += clock64() - tc2;
      }
      tq = prof_on ? clock64() : 0;
      // accumulator drained: hand the TMEM stage back to the MMA warp, the parameter slot to the prefetcher
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(bar_tempty + 8 * acc); else mbar_arrive_cluster(tempty0 + 8 * acc);
        mbar_arrive(bar_pempty + 8 * pslot);
      }
      if (++pslot == kParamSlots) {
        pslot = 0;
        pphase ^= 1;
      }
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }

      if constexpr (LSM) {
        const float kLog2e = 1.4426950408889634f;
        const int xr = quad * 32 + lane;
        if (p.lsm_softmax && sub == n_tiles - 1) {       // end of the first (or only) sweep: the row's log-sum-exp
          xchg_stats[half * kTileM + xr] = make_float2(lsm_m, lsm_s);
          asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * C::kParts) : "memory");   // this row quadrant's warps
          float mm = -FLT_MAX;
#pragma unroll
          for (int h = 0; h < C::kParts; ++h) mm = fmaxf(mm, xchg_stats[h * kTileM + xr].x);
          float ss = 0.0f;
#pragma unroll
          for (int h = 0; h < C::kParts; ++h) {          // column parts in ascending order
            const float2 o = xchg_stats[h * kTileM + xr];
            ss += o.y * ex2_approx((o.x - mm) * kLog2e);
          }
          lsm_lse = mm + logf(ss);
          lsm_m = -FLT_MAX;
          lsm_s = 0.0f;
          if (p.lsm_single && half == 0 && spi_row_flag >= 0) p.lsm_lse_out[spi_row_flag] = lsm_lse;
        }
        if (sub == n_subs - 1) {                         // end of the unit: argmax of the finished row
          const int out_row = spi_row_flag;
          xchg_best[half * kTileM + xr] = make_float2(lsm_best, __int_as_float(lsm_best_i));
          asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * C::kParts) : "memory");
          if (half == 0 && p.lsm_argmax && out_row >= 0) {
#pragma unroll
            for (int h = 1; h < C::kParts; ++h) {
              const float2 o = xchg_best[h * kTileM + xr];
              const int oi = __float_as_int(o.y);
              if (o.x > lsm_best || (o.x == lsm_best && oi < lsm_best_i)) {
                lsm_best = o.x;
                lsm_best_i = oi;
              }
            }
            p.lsm_argmax[out_row] = lsm_best_i == 0x7fffffff ? 0 : lsm_best_i;
          }
          lsm_best = -FLT_MAX;
          lsm_best_i = 0x7fffffff;
        }
      }

      if (p.minmax) {
        if (!use_row) {
          vmin = FLT_MAX;
          vmax = -FLT_MAX;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
          vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if (lane == 0 && vmin <= vmax) {
          atomicMin(p.minmax + 2 * utt, OrderedFromFloat(vmin));
          atomicMax(p.minmax + 2 * utt + 1, OrderedFromFloat(vmax));
        }
      }
      if (prof_on) {
        pr[5] += clock64() - tq;
        pr[7] += 1;
      }
    }
    if (lane == 0) tma_store_wait_all();                 // global writes done before the CTA exits
    if (prof_on && lane == 0) {
      pr[6] = clock64() - pr_t0;
      for (int i = 0; i < 8; ++i) atomicAdd(p.dbg + i, (unsigned long long)pr[i]);
    }
  } else {
    // ===================== parameter prefetcher (warp 10) =====================
    // One tile ahead of the epilogue: everything its warps need per tile -- per-column bias /
    // batch-norm / integer-correction arrays, per-row corrections and FindMinMax flags, per-quadrant
    // scale and utterance -- goes into one of two shared-memory slots, signalled by an mbarrier.
    // Everything is LOADED (into registers) before the wait for a free slot, so the global-memory latency
    // of a tile's parameters overlaps the epilogue of the tile before; the per-row and per-utterance
    // values of the fused output layer are the same for all column tiles of a unit and are fetched once.
    const int n_gran = (p.M + kRowGran - 1) / kRowGran;
    int pslot = 0;
    uint32_t pphase = 0;
    const bool per_quad = GRAN || (LSM && p.gran != 0);  // (the float kinds have no per-utterance arithmetic and
    // no GRAN instantiation, but the fused output layer still needs every quadrant's utterance for the
    // row mapping of a packed row space)
    int utt_q[4] = {0, 0, 0, 0};
    int32_t zp_q[4] = {0, 0, 0, 0};
    float scale_q[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    int32_t row_corr[4] = {0, 0, 0, 0}, row_flag[4] = {0, 0, 0, 0};
    for (int unit = group_id; unit < n_units; unit += n_groups)
    for (int sub = 0; sub < n_subs; ++sub) {
      const int mt = LSM ? unit : unit / n_tiles;
      const int nt = LSM ? (sub >= n_tiles ? sub - n_tiles : sub) : unit % n_tiles;
      const int m0 = mt * kGroupM + (int)rank * kTileM;  // this CTA's 128 rows
      const int n0 = nt * kTileN;
      // per column: 8 columns a lane (parameter arrays are padded to kTileN)
      float c_bias[kTileN / 32], c_bns[kTileN / 32], c_bno[kTileN / 32], c_prior[kTileN / 32];
      int32_t c_sum[kTileN / 32];
#pragma unroll
      for (int j = 0; j < kTileN / 32; ++j) {
        const int c = lane + 32 * j;
        c_bias[j] = p.bias ? __ldg(p.bias + n0 + c) : 0.0f;
        c_bns[j] = p.bn_scale ? __ldg(p.bn_scale + n0 + c) : 1.0f;
        c_bno[j] = p.bn_offset ? __ldg(p.bn_offset + n0 + c) : 0.0f;
        c_prior[j] = (LSM && p.lsm_prior && n0 + c < p.N) ? __ldg(p.lsm_prior + n0 + c) : 0.0f;
        c_sum[j] = (KIND == kKindI8) ? __ldg(p.b_colsum + n0 + c) : 0;
      }
      if (!LSM || sub == 0) {
        // per quadrant (granule mode) or per tile: utterance and its activation quantisation
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          utt_q[g] = p.tile_utt ? __ldg(p.tile_utt + min(m0 / kRowGran + (per_quad ? g : 0), n_gran - 1)) : 0;
          zp_q[g] = 0;
        }
        if (KIND == kKindI8) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!GRAN && g > 0) {
              zp_q[g] = zp_q[0];
              scale_q[g] = scale_q[0];
              continue;
            }
            const QParam q = p.qa[utt_q[g]];
            zp_q[g] = q.zero_point;
            scale_q[g] = __fmul_rn(q.scale, p.scale_b);                 // matrix.cc:403 (float * float)
          }
        }
        // per row: 4 rows a lane, row group j = accumulator quadrant j.  All row-sum loads of the tile
        // (4 rows x taps) are issued before the first one is used: as a loop over the taps they ran one
        // after the other and the first layer's epilogue (5 taps) waited a fifth of its time for them.
        int32_t rs[4][kMaxTaps];
        if (KIND == kKindI8) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int t = 0; t < kMaxTaps; ++t) {
              const int r = m0 + 32 * j + lane + p.tap_off[t < p.n_taps ? t : 0];
              rs[j][t] = (t < p.n_taps && r >= 0 && r < p.M) ? __ldg(p.a_rowsum + r) : 0;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = m0 + 32 * j + lane;
          int32_t rsum = 0;
          if (KIND == kKindI8) {
#pragma unroll
            for (int t = 0; t < kMaxTaps; ++t) rsum += rs[j][t];
          }
          row_corr[j] = p.zp_b * rsum;
          int flag = 0;
          if (LSM) {
            // the output row of this accumulator row: the frame's index in the caller's compact matrix
            // (or the row itself, lsm_rowspace), -1 for context / padding rows
            int pos = row, P = p.M;
            int64_t base = 0;
            if (p.utts) {
              const int u = utt_q[per_quad ? j : 0];
              const UttRows ur = p.utts[u];
              pos = row - ur.row_off;
              P = ur.rows;
              if (p.lsm_out_row_off) base = p.lsm_out_row_off[u];
            }
            const bool ok = row < p.M && pos >= p.lsm_left && pos < P - p.lsm_right;
            flag = !ok ? -1 : p.lsm_rowspace ? row : (int)(base + (pos - p.lsm_left));
          } else if (p.minmax) {
            int pos = row, P = p.M;
            if (p.utts) {
              const UttRows ur = p.utts[utt_q[GRAN ? j : 0]];
              pos = row - ur.row_off;
              P = ur.rows;
            }
            if (pos >= p.mm_lo && pos < P - p.mm_hi) {
              if (p.next_n_taps == 0) {
                flag = 1;
              } else {
                for (int t = 0; t < p.next_n_taps; ++t) {
                  const int o = pos - p.next_tap_off[t];
                  if (o >= p.next_lo && o < P - p.next_hi) flag = 1;
                }
              }
            }
          }
          row_flag[j] = flag;
        }
      }

      mbar_wait_relaxed(bar_pempty + 8 * pslot, pphase ^ 1);
      float *sp = sp_all + pslot * kParamSlotWords;
      int32_t *spi = reinterpret_cast<int32_t *>(sp);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (lane == g) {
          sp[kParamScale + g] = scale_q[g];
          spi[kParamUtt + g] = utt_q[g];
        }
      }
#pragma unroll
      for (int j = 0; j < kTileN / 32; ++j) {
        const int c = lane + 32 * j;
        sp[c] = c_bias[j];
        sp[kTileN + c] = c_bns[j];
        sp[2 * kTileN + c] = c_bno[j];
        if (LSM) sp[kParamPrior + c] = c_prior[j];
#pragma unroll
        for (int g = 0; g < (GRAN ? 4 : 1); ++g)
          spi[(3 + g) * kTileN + c] = (KIND == kKindI8) ? p.k_true * zp_q[g] * p.zp_b - zp_q[g] * c_sum[j] : 0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        spi[kParamRowCorr + 32 * j + lane] = row_corr[j];
        spi[kParamRowFlag + 32 * j + lane] = row_flag[j];
      }
      __syncwarp();                                      // every lane's stores before the arrive (release)
      if (lane == 0) mbar_arrive(bar_pfull + 8 * pslot);
      if (++pslot == kParamSlots) {
        pslot = 0;
        pphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                       // the peer is done with this CTA's smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                   : "memory");
    } else {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                   : "memory");
    }
  }
}

// ---------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int GetEncodeFn(EncodeTiledFn *out) {
  static std::mutex mu;
  static EncodeTiledFn fn = nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || sym == nullptr) {
      SetError("cuTensorMapEncodeTiled is not available from the driver");
      return CE_GPU_ECUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  *out = fn;
  return CE_GPU_OK;
}

// cuTensorMapEncodeTiled costs microseconds of host time, and a streaming micro-batch call makes 42 of them
// for a millisecond of kernels -- with the same operands call after call.  Encoded maps are therefore
// kept per host thread, keyed by everything that goes into them.
struct MapKey {
  const void *base;
  int64_t rows, cols, ld;
  int box_rows, tag;                                     // tag: operand kind, or 16 + bf16 for output maps
  bool operator==(const MapKey &o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && tag == o.tag;
  }
};
struct MapCache {
  static constexpr int kEntries = 128;
  MapKey key[kEntries];
  CUtensorMap map[kEntries];
  int n = 0, next = 0;
  bool Find(const MapKey &k, CUtensorMap *out) const {
    for (int i = 0; i < n; ++i)
      if (key[i] == k) {
        *out = map[i];
        return true;
      }
    return false;
  }
  void Put(const MapKey &k, const CUtensorMap &m) {
    const int i = n < kEntries ? n++ : (next++ % kEntries);
    key[i] = k;
    map[i] = m;
  }
};
MapCache &Maps() {
  static thread_local MapCache c;
  return c;
}

// 2-D map over a row-major [rows x cols] matrix, box = [box_rows x 128 bytes], 128B swizzle.
int MakeMap(int kind, const void *base, int64_t rows, int64_t cols, int box_rows, CUtensorMap *map) {
  const MapKey mk = {base, rows, cols, cols, box_rows, kind};
  if (Maps().Find(mk, map)) return CE_GPU_OK;
  EncodeTiledFn fn;
  CE_CHECK(GetEncodeFn(&fn));
  const int elt = KindEltBytes(kind);
  const CUtensorMapDataType dt = kind == kKindI8     ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : kind == kKindTF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                     : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (cols * elt) % 16 != 0) {
    SetError("GEMM operand is not 16-byte aligned (base %p, row pitch %lld bytes)", base,
             (long long)(cols * elt));
    return CE_GPU_EINVAL;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(cols * elt)};
  cuuint32_t box[2] = {(cuuint32_t)(kTileKBytes / elt), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    SetError("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld cols %lld kind %d)", (int)r,
             (long long)rows, (long long)cols, kind);
    return CE_GPU_ECUDA;
  }
  Maps().Put(mk, *map);
  return CE_GPU_OK;
}

// 2-D map for the TMA stores of the epilogue: [rows x cols] window of a row-major matrix with row
// stride ld (elements), box = 32 rows x 128 bytes, 128B swizzle (box64: 32 rows x 64 bytes, 64B swizzle).
int MakeOutMap(bool bf16, const void *base, int64_t rows, int64_t cols, int64_t ld, CUtensorMap *map,
               bool box64 = false) {
  const MapKey mk = {base, rows, cols, ld, 32, 16 + (bf16 ? 1 : 0) + (box64 ? 2 : 0)};
  if (Maps().Find(mk, map)) return CE_GPU_OK;
  EncodeTiledFn fn;
  CE_CHECK(GetEncodeFn(&fn));
  const int elt = bf16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * elt) % 16 != 0) {
    SetError("GEMM output is not 16-byte aligned (base %p, row pitch %lld bytes)", base,
             (long long)(ld * elt));
    return CE_GPU_EINVAL;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * elt)};
  cuuint32_t box[2] = {(cuuint32_t)((box64 ? 64 : 128) / elt), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    SetError("cuTensorMapEncodeTiled (output) failed with CUresult %d (rows %lld cols %lld ld %lld)",
             (int)r, (long long)rows, (long long)cols, (long long)ld);
    return CE_GPU_ECUDA;
  }
  Maps().Put(mk, *map);
  return CE_GPU_OK;
}

template <int KIND, int CG, bool GRAN = false, int MODE = kModeClassic>
int LaunchKind(const GemmOperands &ops, const GemmArgs &args, cudaStream_t s) {
  constexpr bool LSM = MODE == kModeLsm;
  using C = Cfg<CG, MODE != kModeClassic>;
  CUtensorMap ma0, ma1, mb0, mb1, mo0, mo1;
  const bool out_is_bf16 = (KIND == kKindBF16 || KIND == kKindBF16X3) && args.out_bf16 != nullptr;
  const void *o0 = out_is_bf16 ? static_cast<const void *>(args.out_bf16) : static_cast<const void *>(args.out_f32);
  if (o0 == nullptr && !LSM) {
    SetError("GemmLaunch: no output buffer");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(MakeMap(KIND, ops.a[0], ops.rows_a, args.c_pad, kTileM, &ma0));
  if (LSM) {
    // whole 32-row groups of finished rows leave through TMA stores of 32 x 16 columns, the others
    // through plain stores
    if (args.out_f32) CE_CHECK(MakeOutMap(false, args.out_f32, args.lsm_out_rows, args.N, args.ld_out, &mo0, true));
    else mo0 = ma0;
  } else if (MODE == kModeWide) {
    CE_CHECK(MakeOutMap(false, o0, args.M, args.n_store, args.ld_out, &mo0, true));
  } else {
    CE_CHECK(MakeOutMap(out_is_bf16, o0, args.M,
                        (KIND == kKindBF16X3 && out_is_bf16) ? 2 * args.n_store : args.n_store, args.ld_out, &mo0));
  }
  if (args.out_lo && !LSM) {
    CE_CHECK(MakeOutMap(false, args.out_lo, args.M, args.n_store, args.ld_out, &mo1));
  } else {
    mo1 = mo0;                                           // never used by the kernel
  }
  CE_CHECK(MakeMap(KIND, ops.a[1] ? ops.a[1] : ops.a[0], ops.rows_a, args.c_pad, kTileM, &ma1));
  CE_CHECK(MakeMap(KIND, ops.b[0], ops.rows_b, ops.k_total, C::kBRows, &mb0));
  CE_CHECK(MakeMap(KIND, ops.b[1] ? ops.b[1] : ops.b[0], ops.rows_b, ops.k_total, C::kBRows, &mb1));
  static thread_local bool configured[64] = {false};
  int dev = 0;
  CE_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    CE_CUDA(cudaFuncSetAttribute(gemm_kernel<KIND, CG, GRAN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::kSmemBytes));
    configured[dev] = true;
  }
  const int group_m = kTileM * CG;
  const int m_tiles = (args.M + group_m - 1) / group_m;
  const int n_tiles = (args.N + kTileN - 1) / kTileN;
  const int64_t tiles = (int64_t)m_tiles * n_tiles;
  if (tiles <= 0) return CE_GPU_OK;
  const int groups = (int)std::min<int64_t>(LSM ? m_tiles : tiles, SmCount() / CG);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(LSM ? kProfFinalize : kProfGemm, s);    // the fused output layer IS the log-softmax launch now
  CE_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<KIND, CG, GRAN, MODE>, ma0, ma1, mb0, mb1, mo0, mo1, args));
  CE_LAUNCHED();
  return CE_GPU_OK;
}

}  // namespace

int GemmLaunch(int kind, const GemmOperands &ops, const GemmArgs &args_in, cudaStream_t s) {
  GemmArgs args = args_in;
  static const int debug_bits = getenv("CE_GPU_GEMM_DEBUG") ? atoi(getenv("CE_GPU_GEMM_DEBUG")) : 0;
  args.debug = debug_bits;
  if (args.c_pad <= 0 || (args.c_pad * KindEltBytes(kind)) % kTileKBytes != 0 || args.n_taps < 1 ||
      args.n_taps > kMaxTaps || args.n_pass < 1 || args.n_pass > 3) {
    SetError("GemmLaunch: bad geometry (c_pad %d, taps %d, passes %d)", args.c_pad, args.n_taps,
             args.n_pass);
    return CE_GPU_EINVAL;
  }
  const bool x3_operand_out = kind == kKindBF16X3 && args.out_bf16 != nullptr;
  if (args.n_store < args.N || (x3_operand_out ? 2 * args.n_store : args.n_store) > args.ld_out ||
      (x3_operand_out && args.n_store % 32 != 0)) {
    SetError("GemmLaunch: n_store %d outside [N %d, ld_out %lld]", args.n_store, args.N,
             (long long)args.ld_out);
    return CE_GPU_EINVAL;
  }
  if (args.ld_out % 4 != 0) {
    SetError("GemmLaunch: output row stride %lld is not a multiple of 4", (long long)args.ld_out);
    return CE_GPU_EINVAL;
  }
  // built with -DCE_GEMM_PROF and run with CE_GPU_GEMM_PROF=1 (a debugging aid: synchronous, one line per
  // launch): where the epilogue warps' time goes, averaged per warp
#ifdef CE_GEMM_PROF
  static const bool prof = getenv("CE_GPU_GEMM_PROF") != nullptr;
#else
  static const bool prof = false;
#endif
  static unsigned long long *prof_buf = nullptr;
  if (prof) {
    if (!prof_buf) CE_CUDA(cudaMalloc(&prof_buf, 64));
    CE_CUDA(cudaMemsetAsync(prof_buf, 0, 64, s));
    args.dbg = prof_buf;
  }
  struct ProfPrint {
    bool on; cudaStream_t s; unsigned long long *buf; int M, N, taps;
    ~ProfPrint() {
      if (!on) return;
      unsigned long long h[8];
      if (cudaStreamSynchronize(s) != cudaSuccess || cudaMemcpy(h, buf, 64, cudaMemcpyDeviceToHost) != cudaSuccess) return;
      constexpr int kEpiWarps = 8;
      const double w = 1965.0 * SmCount() * kEpiWarps;   // cycles -> us per epilogue warp (at the maximum clock)
      fprintf(stderr, "gemm M %d N %d taps %d: epilogue warp us: loop %.1f = wait-params %.1f + wait-accumulator %.1f + "
              "tcgen05.ld %.1f + math %.1f + stage/store %.1f + tail %.1f + rest; %.1f tiles\n", M, N, taps, h[6] / w,
              h[0] / w, h[1] / w, h[2] / w, h[3] / w, h[4] / w, h[5] / w, (double)h[7] / (SmCount() * kEpiWarps));
    }
  } prof_print{prof, s, prof_buf, args.M, args.N, args.n_taps};
  static const int cta_group = getenv("CE_GPU_CTA_GROUP") ? atoi(getenv("CE_GPU_CTA_GROUP")) : 2;
  // granule mode only changes the int8 epilogue (the float kinds carry no per-utterance parameters)
  const bool gran = args.gran != 0 && kind == kKindI8 && args.tile_utt != nullptr;
  if (args.lsm) {
    if (args.relu || args.bn_scale || args.out_acc) {
      SetError("GemmLaunch: the fused LogSoftmax output takes a plain Linear layer (no ReLU / BatchNorm / accumulator dump)");
      return CE_GPU_EINVAL;
    }
    if (args.N % 4 != 0 || (args.out_f32 && (args.ld_out % 4 != 0 || (reinterpret_cast<uintptr_t>(args.out_f32) & 15) != 0))) {
      SetError("GemmLaunch: the fused LogSoftmax output needs N %% 4 == 0 and 16-byte aligned rows (N %d, ld %lld)",
               args.N, (long long)args.ld_out);
      return CE_GPU_EINVAL;
    }
    if (cta_group == 1) {
      switch (kind) {
        case kKindI8: return gran ? LaunchKind<kKindI8, 1, true, kModeLsm>(ops, args, s) : LaunchKind<kKindI8, 1, false, kModeLsm>(ops, args, s);
        case kKindBF16: return LaunchKind<kKindBF16, 1, false, kModeLsm>(ops, args, s);
        case kKindTF32: return LaunchKind<kKindTF32, 1, false, kModeLsm>(ops, args, s);
        case kKindBF16X3: return LaunchKind<kKindBF16X3, 1, false, kModeLsm>(ops, args, s);
      }
    } else {
      switch (kind) {
        case kKindI8: return gran ? LaunchKind<kKindI8, 2, true, kModeLsm>(ops, args, s) : LaunchKind<kKindI8, 2, false, kModeLsm>(ops, args, s);
        case kKindBF16: return LaunchKind<kKindBF16, 2, false, kModeLsm>(ops, args, s);
        case kKindTF32: return LaunchKind<kKindTF32, 2, false, kModeLsm>(ops, args, s);
        case kKindBF16X3: return LaunchKind<kKindBF16X3, 2, false, kModeLsm>(ops, args, s);
      }
    }
  }
  // kModeWide: plain int8 layers through the fused output layer's 16-warp epilogue.  Bit-exact (the whole GPU
  // suite passes with it on every layer, CE_GPU_WIDE_EPILOGUE=1), but only worth it where the epilogue is the
  // bound: layers whose K fits two pipeline stages (the first layer, K = 200: 160 -> 148 us per 131072 rows,
  // GEMMs 6.37 -> 6.20 ms per step) -- the default, CE_GPU_WIDE_EPILOGUE=2.  On the hidden layers, which are bound
  // by the multiplications, it is a little faster per launch under ncu (280 -> 274 us) and SLOWER in the timed
  // step, which runs into the power cap (9.61 against 9.26 ms with it on every layer).  0 = never.
  static const int wide_env = getenv("CE_GPU_WIDE_EPILOGUE") ? atoi(getenv("CE_GPU_WIDE_EPILOGUE")) : 2;
  const bool wide_on = wide_env == 1 || (wide_env == 2 && (int64_t)args.n_taps * args.c_pad <= 2 * kTileKBytes);
  if (wide_on && kind == kKindI8 && cta_group != 1 && args.out_f32 && !args.out_acc && !args.out_lo)
    return gran ? LaunchKind<kKindI8, 2, true, kModeWide>(ops, args, s) : LaunchKind<kKindI8, 2, false, kModeWide>(ops, args, s);
  if (cta_group == 1) {
    switch (kind) {
      case kKindI8: return gran ? LaunchKind<kKindI8, 1, true>(ops, args, s) : LaunchKind<kKindI8, 1>(ops, args, s);
      case kKindBF16: return LaunchKind<kKindBF16, 1>(ops, args, s);
      case kKindTF32: return LaunchKind<kKindTF32, 1>(ops, args, s);
      case kKindBF16X3: return LaunchKind<kKindBF16X3, 1>(ops, args, s);
    }
  } else {
    switch (kind) {
      case kKindI8: return gran ? LaunchKind<kKindI8, 2, true>(ops, args, s) : LaunchKind<kKindI8, 2>(ops, args, s);
      case kKindBF16: return LaunchKind<kKindBF16, 2>(ops, args, s);
      case kKindTF32: return LaunchKind<kKindTF32, 2>(ops, args, s);
      case kKindBF16X3: return LaunchKind<kKindBF16X3, 2>(ops, args, s);
    }
  }
  SetError("GemmLaunch: unknown kind %d", kind);
  return CE_GPU_EINVAL;
}

}  // namespace ce
