// K1 v2 on the 5th-generation tensor cores: 128 points per CTA (UMMA M=256 over a CTA pair),
// activation hi-halves in TENSOR MEMORY (A operand read by tcgen05.mma straight from TMEM),
// lo-halves in shared memory, biases and pose-align point terms folded into K=16 UMMAs so the
// epilogues carry no per-feature parameters.
//
// Same contract as k1_tc.cu (asdf_tc_eval); replaces utils/mesh.py:46-63,96-115,
// utils/utils.py:376-430,561-572, networks/model.py:285-350, utils/mesh.py:207-247.
//
// Why a second kernel: with 64 rows per CTA (k1_tc.cu) every weight byte feeds half as many MACs,
// so the weight ring (L2 -> SMEM) and the tensor core's shared-memory operand reads both run at
// twice the rate per FLOP and the kernel is bound by shared-memory bandwidth and by the
// barrier round trip of the ring.  Here a CTA owns 128 rows: weight traffic per FLOP
// halves, and two of the three split-precision products read A from TMEM, not from SMEM.
//
//   TMEM   [  0,128) ACC0   [128,256) ACC1   (fp32 accumulators of one 128-wide N block each)
//          [256,512) AHI    fp16 pairs, column 256 + k/2 holds (k even | k odd << 16) of hi(t*x[k])
//   SMEM   ALO  8 slots x [128 rows x 64 k] fp16 lo halves, K-major 128B swizzle       128 KiB
//          AP   [128 rows x 16 k] point operand (cp*p_hi, c1, cp*p_lo, ...), no swizzle      4 KiB
//          RING 5 x (hi, lo) pairs of [64 rows x 64 k] weight tiles (this CTA's half)      80 KiB
//
// Per N block (128 output features):  UMMA(AP, Ptile)            bias + point term, K=16
//                                     per 64-wide K chunk:  4 x UMMA(AHI, Bhi) + 4 x UMMA(ALO, Bhi)
//                                                           4 x UMMA(AHI, Blo)
// Epilogue of every layer: v = relu(acc * inv) -> fp16 hi/lo split -> hi to AHI (tcgen05.st),
// lo to ALO; layer 3: dot with w4, tanh, store, bbox.
#include "common.cuh"
#include <cuda_fp16.h>

namespace asdf {
namespace tc2 {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr int kRows = 128;                  // points per CTA
constexpr int kPtsPerTile = 256;            // per CTA pair
constexpr int kTileBytes = 64 * 64 * 2;     // weight tile: 64 rows x 64 k fp16 = 8 KiB
constexpr int kSlotBytes = kRows * 64 * 2;  // ALO slot / AP: 128 rows x 64 k fp16 = 16 KiB
constexpr int kRing = 5;                    // ring slots of one (hi, lo) tile pair each
constexpr int kMainTilesPerDecoder = 128;   // 32 (L1) + 32 (L2) + 64 (L3)
constexpr int kPTilesPerDecoder = 14;       // 4 (L0) + 2 (L1) + 4 (L2) + 4 (L3) N blocks
constexpr int kSlotTileBytes = 2 * kTileBytes;   // a ring slot holds a (hi, lo) pair (16 KiB) or one P tile
constexpr int kFillsPerItem = kMainTilesPerDecoder / 2 + kPTilesPerDecoder;   // ring fills per work item
constexpr int64_t kWeightBytes = (int64_t)2 * 2 * kMainTilesPerDecoder * kTileBytes;   // [dec][rank][tile]
constexpr int kStaticParamFloats = 512 + 8;            // w4[512] | b4, inv1, inv2, inv3, pad
constexpr int64_t kSampleBytes = (int64_t)2 * 2 * kPTilesPerDecoder * kTileBytes + 64;  // P tiles + 16 floats

constexpr int kOffALo = 0;
constexpr int kApBytes = kRows * 16 * 2;                          // point operand: 128 rows x 16 k, no swizzle (4 KiB)
constexpr int kOffAP = kOffALo + 8 * kSlotBytes;                 // 131072
constexpr int kOffRing = kOffAP + kApBytes;                      // 135168 (1024-aligned)
constexpr int kOffW4 = kOffRing + kRing * kSlotTileBytes;        // 217088
constexpr int kOffRed = kOffW4 + 512 * 4;
constexpr int kOffBar = kOffRed + 2 * kRows * 4;
constexpr int kBarFull = 0;
constexpr int kBarFullLocal = kBarFull + kRing;
constexpr int kBarEmpty = kBarFullLocal + kRing;
constexpr int kBarAFull = kBarEmpty + kRing;           // [8] K positions
constexpr int kBarTmemFull = kBarAFull + 8;            // [2]
constexpr int kBarTmemEmpty = kBarTmemFull + 2;        // [2]
constexpr int kBarApFull = kBarTmemEmpty + 2;          // [1]
constexpr int kNumBars = kBarApFull + 1;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KiB dynamic shared memory limit");

constexpr uint32_t kIdesc = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);   // f32 acc, f16 x f16, N=128, M=256
constexpr uint32_t kAhiCol = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {   // see k1_tc.cu on the scope
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" :: "r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {     // K-major, 128B swizzle, SBO = 1024 B
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d), "l"(a), "l"(b), "r"(kIdesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
               :: "r"(d), "r"(a_tmem), "l"(b), "r"(kIdesc), "r"(acc), "r"(0u) : "memory");
}
// Lean forms used by the issuer: operands are the LOW descriptor words (address >> 4); the high
// word (SBO = 1024 B, version 1, 128B swizzle) is the constant 0x40004040.  All operands are
// warp-uniform so ptxas keeps them in uniform registers (no R2UR waterfall per UMMA).
// The whole (converged) issuer warp executes these wrappers and elect.sync picks the issuing lane:
// ptxas then emits the UTC*MMA directly (a `lane == 0` predicate made it wrap every UMMA in a
// VOTEU / ELECT / BRA.U.ANY loop over the active lanes, ~50 cycles per UMMA; see k1_tc3.cu).
__device__ __forceinline__ void umma_ss_lo(uint32_t issue, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
               "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
               :: "r"(d), "r"(a_lo), "r"(b_lo), "r"(kIdesc), "r"(acc), "r"(0x40004040u), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma_ts_lo(uint32_t issue, uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
               "mov.b64 db, {%2, %5};\n\t"
               "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, {%6, %6, %6, %6, %6, %6, %6, %6}, p;\n\t}"
               :: "r"(d), "r"(a_tmem), "r"(b_lo), "r"(kIdesc), "r"(acc), "r"(0x40004040u), "r"(0u), "r"(issue) : "memory");
}
// A = point operand in the no-swizzle K-major layout: core matrices of 8 rows x 16 B, the two K
// halves 128 B apart (LBO), 8-row groups 256 B apart (SBO); B = SW128 tile as above.
__device__ __forceinline__ void umma_ap(uint32_t issue, uint32_t d, uint32_t ap_addr, uint32_t b_lo, uint32_t acc) {
  const uint32_t a_lo = ((ap_addr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
               "mov.b64 da, {%1, %6};\n\tmov.b64 db, {%2, %5};\n\t"
               "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
               :: "r"(d), "r"(a_lo), "r"(b_lo), "r"(kIdesc), "r"(acc), "r"(0x40004040u),
                  "r"((256u >> 4) | (1u << 14)), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma_commit_both_if(uint32_t issue, uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
               :: "r"(bar), "h"((uint16_t)3), "r"(issue) : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr & 0x3FFFFu) >> 4; }
__device__ __forceinline__ void umma_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* w) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
         "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ float lds_f1(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct Args {
  asdf_query q;
  const uint8_t* stat;     // main weight tiles [dec][rank][128] then static params [dec][520 floats]
  const uint8_t* samp;     // P tiles [dec][rank][14] then 16 floats: inv0[2], cp, c1
  float* out_hand;
  float* out_obj;
  int32_t* bbox;
  long long* dbg;          // optional int64[32] of cycle counters of CTA pair 0 (tools/tc_phase_timing.py)
};

// N-block schedule of one work item: layer, number of 64-wide K chunks, K position of chunk j
__device__ __forceinline__ int nb_layer(int g) { return g < 4 ? 0 : (g < 6 ? 1 : (g < 10 ? 2 : 3)); }
__device__ __forceinline__ int layer_chunks(int layer) { return layer == 0 ? 0 : (layer == 2 ? 4 : 8); }

template <bool kDebug>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) tc2_eval_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;          // == %cluster_ctarank for __cluster_dims__(2,1,1); provably uniform
  const bool leader = rank == 0;
  auto bar = [&](int i) { return sbase + kOffBar + 8 * i; };

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar(kBarFull + i), 2);
      mbar_init(bar(kBarFullLocal + i), 1);
      mbar_init(bar(kBarEmpty + i), 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(bar(kBarAFull + i), 16);     // 4 warps x 2 lanes x 2 CTAs
    for (int i = 0; i < 2; ++i) { mbar_init(bar(kBarTmemFull + i), 1); mbar_init(bar(kBarTmemEmpty + i), 16); }
    mbar_init(bar(kBarApFull), 8);                                    // 4 warps x 2 CTAs
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(sbase + kOffTmemPtr) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

  const int64_t total = a.q.end - a.q.begin;
  const int64_t n_tiles = (total + kPtsPerTile - 1) / kPtsPerTile;
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // =========================== weight-stream producer ===========================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      auto push = [&](const uint8_t* src, uint32_t bytes) {
        mbar_wait(bar(kBarEmpty + slot), phase ^ 1);
        const uint32_t fb = bar((leader ? kBarFull : kBarFullLocal) + slot);
        mbar_expect_tx(fb, bytes);
        bulk_g2s(sbase + kOffRing + slot * kSlotTileBytes, src, bytes, fb);
        if (++slot == kRing) { slot = 0; phase ^= 1; }
      };
      for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
        for (int dec = 0; dec < 2; ++dec) {
          const uint8_t* mt = a.stat + (int64_t)(dec * 2 + rank) * kMainTilesPerDecoder * kTileBytes;
          const uint8_t* pt = a.samp + (int64_t)(dec * 2 + rank) * kPTilesPerDecoder * kTileBytes;
          for (int g = 0; g < kPTilesPerDecoder; ++g) {
            push(pt, kTileBytes); pt += kTileBytes;
            const int n = layer_chunks(nb_layer(g));
            for (int i = 0; i < n; ++i) { push(mt, kSlotTileBytes); mt += kSlotTileBytes; }   // hi + lo in one copy
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (!leader) {
      // ======================= peer CTA: relay "my half of the tile landed" =======================
      if (lane == 0) {
        uint32_t slot = 0, phase = 0;
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
          for (int i = 0; i < 2 * kFillsPerItem; ++i) {
            mbar_wait(bar(kBarFullLocal + slot), phase);
            mbar_arrive_cluster(bar(kBarFull + slot), 0);
            if (++slot == kRing) { slot = 0; phase ^= 1; }
          }
        }
      }
      __syncwarp();
    } else {
      // =================================== UMMA issuer ===================================
      // The whole warp walks the schedule (all values warp-uniform -> uniform registers); lane 0
      // issues the tcgen05 instructions.
      const uint32_t issue = lane == 0 ? 1u : 0u;
      uint32_t slot = 0, phase = 0, a_phase = 0, ap_phase = 0;
      uint32_t nblk = 0;                                   // global N-block counter -> TMEM buffer + parities
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0);
      const uint32_t alo_lo = desc_lo(sb + kOffALo), ring_lo = desc_lo(sb + kOffRing), ap_addr = sb + kOffAP;
      const uint32_t bar0 = sb + kOffBar;
      long long w_ring = 0, w_a = 0, w_acc = 0, t_begin = kDebug ? clock64() : 0;
      auto take = [&]() -> uint32_t {                      // wait for the next ring tile, return its descriptor word
        const long long t0 = kDebug ? clock64() : 0;
        mbar_wait(bar0 + 8 * (kBarFull + slot), phase);
        if (kDebug) w_ring += clock64() - t0;
        tc_fence_after();
        return ring_lo + slot * (kSlotTileBytes >> 4);
      };
      auto release = [&]() {
        umma_commit_both_if(issue, bar0 + 8 * (kBarEmpty + slot));
        if (++slot == kRing) { slot = 0; phase ^= 1; }
      };
      for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
        mbar_wait(bar0 + 8 * kBarApFull, ap_phase); ap_phase ^= 1;
        tc_fence_after();
        for (int dec = 0; dec < 2; ++dec) {
          for (int g = 0; g < kPTilesPerDecoder; ++g, ++nblk) {
            const int layer = nb_layer(g);
            const bool first_nb = g == 0 || g == 4 || g == 6 || g == 10;
            const uint32_t buf = nblk & 1u, use = nblk >> 1;
            const uint32_t d_tmem = tmem_u + buf * 128;
            {
              const long long t0 = kDebug ? clock64() : 0;
              mbar_wait(bar0 + 8 * (kBarTmemEmpty + buf), (use & 1u) ^ 1u);
              if (kDebug) w_acc += clock64() - t0;
            }
            tc_fence_after();
            {   // bias + point term: K = 16
              const uint32_t b = take();
              umma_ap(issue, d_tmem, ap_addr, b, 0u);
              release();
            }
            const int nch = layer_chunks(layer);
            for (int j = 0; j < nch; ++j) {
              const int pos = layer == 3 ? ((j + 4) & 7) : j;
              if (first_nb) {
                const long long t0 = kDebug ? clock64() : 0;
                mbar_wait(bar0 + 8 * (kBarAFull + pos), (a_phase >> pos) & 1u);
                if (kDebug) w_a += clock64() - t0;
                a_phase ^= 1u << pos;
                tc_fence_after();
              }
              const uint32_t ahi = tmem_u + kAhiCol + pos * 32;
              const uint32_t alo = alo_lo + pos * (kSlotBytes >> 4);
              const uint32_t b = take();                       // (hi, lo) tile pair
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_ts_lo(issue, d_tmem, ahi + ks * 8, b + ks * 2, 1u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_ss_lo(issue, d_tmem, alo + ks * 2, b + ks * 2, 1u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_ts_lo(issue, d_tmem, ahi + ks * 8, b + (kTileBytes >> 4) + ks * 2, 1u);
              release();
            }
            umma_commit_both_if(issue, bar0 + 8 * (kBarTmemFull + buf));
          }
        }
      }
      if (kDebug && cluster_id == 0 && issue) {
        a.dbg[0] = clock64() - t_begin; a.dbg[1] = w_a; a.dbg[2] = w_ring; a.dbg[3] = w_acc;
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // =================================== epilogue warps ===================================
    const int e = warp - kEpiWarp0;
    const int q = warp & 3;                        // TMEM lane quadrant
    const int ch = e >> 2;                         // 64-column half of the 128-column accumulator
    const int row = q * 32 + lane;                 // 0..127
    const int et = threadIdx.x - kEpiWarp0 * 32;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t a_lo = sbase + kOffALo, sW4 = sbase + kOffW4, sRed = sbase + kOffRed;
    const float* sparams = reinterpret_cast<const float*>(a.stat + kWeightBytes);
    const float* sscal = reinterpret_cast<const float*>(a.samp + (int64_t)2 * 2 * kPTilesPerDecoder * kTileBytes);
    const float cp = __ldg(sscal + 2), c1 = __ldg(sscal + 3);
    uint32_t nblk = 0;
    auto wait_full = [&](uint32_t n) {
      mbar_wait(bar(kBarTmemFull + (n & 1u)), (n >> 1) & 1u);
      tc_fence_after();
    };
    // relu + fp16 split of 32 accumulator columns -> 16 hi words, 16 lo words (pairs k, k+1)
    auto split32 = [&](const float* acc, float inv, uint32_t* hi, uint32_t* lo) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x = fminf(fmaxf(acc[2 * i] * inv, 0.f), 60000.f), y = fminf(fmaxf(acc[2 * i + 1] * inv, 0.f), 60000.f);
        const __half2 h = __floats2half2_rn(x, y);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
      }
    };
    // write the 64 features [64*cidx', ...) of this thread's row: position chunk `pos`, halves h = 0,1
    auto store_half = [&](int pos, int h, const uint32_t* hi, const uint32_t* lo) {
      tmem_st16(tmem_base + lane_addr + kAhiCol + pos * 32 + h * 16, hi);
      const uint32_t base = a_lo + pos * kSlotBytes + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        sts_u4(base + ((((h * 4 + g) ^ (row & 7))) << 4), make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]));
    };
    auto publish = [&](int pos) {
      tmem_st_wait();
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane < 2) mbar_arrive_cluster(bar(kBarAFull + pos), 0);
    };
    auto free_acc = [&](uint32_t n) {              // this warp is done reading accumulator buffer n&1
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar(kBarTmemEmpty + (n & 1u)), 0);
    };

    long long ph[16];
    for (int z = 0; z < 16; ++z) ph[z] = 0;
    const bool stamp = kDebug && cluster_id == 0 && rank == 0 && et == 0;
    long long tlast = kDebug ? clock64() : 0;
#define ASDF_STAMP2(k) do { if (stamp) { const long long _t = clock64(); ph[k] += _t - tlast; tlast = _t; } } while (0)
    for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
      const int64_t i = a.q.begin + t * kPtsPerTile + rank * kRows + row;
      const bool live = i < a.q.end;
      // ---------------- point operand AP (rows written by the ch == 0 warps) ----------------
      if (ch == 0) {
        float px = 0.f, py = 0.f, pz = 0.f;
        if (live) {
          if (a.q.mode == ASDF_QUERY_POINTS) {
            const float* r = a.q.points_dev + (size_t)i * a.q.point_stride;
            px = __ldg(r); py = __ldg(r + 1); pz = __ldg(r + 2);
          } else {
            grid_point(i, a.q.N, a.q.mode, a.q.voxel, a.q.origin[0], a.q.origin[1], a.q.origin[2], px, py, pz);
          }
        }
        const float sx = px * cp, sy = py * cp, sz = pz * cp;
        const __half2 hxy = __floats2half2_rn(sx, sy), hz1 = __floats2half2_rn(sz, c1);
        const float2 fxy = __half22float2(hxy);
        const float fz = __low2float(hz1);
        const __half2 lxy = __floats2half2_rn(sx - fxy.x, sy - fxy.y), lz0 = __floats2half2_rn(sz - fz, 0.f);
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(&hxy), w1 = *reinterpret_cast<const uint32_t*>(&hz1);
        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(&lxy), w3 = *reinterpret_cast<const uint32_t*>(&lz0);
        const uint32_t base = sbase + kOffAP + (row >> 3) * 256 + (row & 7) * 16;
        sts_u4(base, make_uint4(w0, w1, w2, w3));             // k 0..7 : p_hi, c1, p_lo, 0
        sts_u4(base + 128, make_uint4(w0, w1, 0u, 0u));       // k 8..15: p_hi, c1, 0
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar(kBarApFull), 0);
      }
      for (int dec = 0; dec < 2; ++dec) {
        epi_bar_sync();                            // previous item's readers of sW4 / sRed are done
        {
          const float* sp = sparams + (size_t)dec * kStaticParamFloats;
          sts_f1(sW4 + 4 * et, __ldg(sp + et)); sts_f1(sW4 + 4 * (et + 256), __ldg(sp + et + 256));
        }
        epi_bar_sync();
        const float* sp = sparams + (size_t)dec * kStaticParamFloats + 512;
        const float b4 = __ldg(sp), inv1 = __ldg(sp + 1), inv2 = __ldg(sp + 2), inv3 = __ldg(sp + 3);
        const float inv0 = __ldg(sscal + dec);
        float part = 0.f;
        for (int g = 0; g < kPTilesPerDecoder; ++g, ++nblk) {
          const int layer = nb_layer(g);
          const int nb = g - (layer == 0 ? 0 : (layer == 1 ? 4 : (layer == 2 ? 6 : 10)));
          const uint32_t acc_addr = tmem_base + lane_addr + (nblk & 1u) * 128 + ch * 64;
          ASDF_STAMP2(8 + layer);                 // epilogue work attributed to the previous phase of this layer
          wait_full(nblk);
          ASDF_STAMP2(layer);                     // waiting for the accumulator of this layer
          if (layer < 3) {
            // feature chunk 2*nb + ch of the layer output -> K position of the next layer's input
            const int cidx = 2 * nb + ch;
            const int pos = layer == 2 ? ((cidx + 4) & 7) : cidx;
            const float inv = layer == 0 ? inv0 : (layer == 1 ? inv1 : inv2);
            // blocks whose target positions are still being read by this layer's remaining UMMAs
            const bool hold = (layer == 1 && nb == 0) || (layer == 2 && nb == 2);
            uint32_t hi[2][16], lo[2][16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float acc[32];
              tmem_ld32(acc_addr + h * 32, acc);
              tmem_ld_wait();
              split32(acc, inv, hi[h], lo[h]);
            }
            free_acc(nblk);
            ASDF_STAMP2(8 + layer);
            if (hold) wait_full(nblk + 1);          // all UMMAs of this layer have retired
            ASDF_STAMP2(4 + layer);
            store_half(pos, 0, hi[0], lo[0]);
            store_half(pos, 1, hi[1], lo[1]);
            publish(pos);
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float acc[32];
              tmem_ld32(acc_addr + h * 32, acc);
              tmem_ld_wait();
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 w = lds_f4(sW4 + 4 * (128 * nb + 64 * ch + 32 * h + 4 * j4));
                part = fmaf(fmaxf(acc[4 * j4 + 0] * inv3, 0.f), w.x, part);
                part = fmaf(fmaxf(acc[4 * j4 + 1] * inv3, 0.f), w.y, part);
                part = fmaf(fmaxf(acc[4 * j4 + 2] * inv3, 0.f), w.z, part);
                part = fmaf(fmaxf(acc[4 * j4 + 3] * inv3, 0.f), w.w, part);
              }
            }
            free_acc(nblk);
          }
        }
        ASDF_STAMP2(11);
        sts_f1(sRed + 4 * (ch * kRows + row), part);
        epi_bar_sync();
        if (ch == 0) {
          const float val = tanhf(lds_f1(sRed + 4 * row) + lds_f1(sRed + 4 * (kRows + row)) + b4);
          if (live) (dec == 0 ? a.out_hand : a.out_obj)[i - a.q.begin] = val;
          if (a.bbox && a.q.mode != ASDF_QUERY_POINTS && (a.q.bbox_mask >> dec & 1))
            bbox_update(a.bbox + 6 * dec, live && val < 0.f, i, a.q.N);
        }
        ASDF_STAMP2(12);
      }
    }
    if (stamp) for (int z = 0; z < 16; ++z) a.dbg[8 + z] = ph[z];
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace tc2
}  // namespace asdf

extern "C" int64_t asdf_tc2_static_bytes(void) {
  return asdf::tc2::kWeightBytes + (int64_t)2 * asdf::tc2::kStaticParamFloats * 4;
}
extern "C" int64_t asdf_tc2_sample_bytes(void) { return asdf::tc2::kSampleBytes; }

extern "C" int asdf_tc2_eval_debug(const void* static_dev, const void* sample_dev, const asdf_query* q,
                                   float* out_hand_dev, float* out_obj_dev, int32_t* bbox_dev, void* stream,
                                   void* debug_dev);

extern "C" int asdf_tc2_eval(const void* static_dev, const void* sample_dev, const asdf_query* q,
                             float* out_hand_dev, float* out_obj_dev, int32_t* bbox_dev, void* stream) {
  return asdf_tc2_eval_debug(static_dev, sample_dev, q, out_hand_dev, out_obj_dev, bbox_dev, stream, nullptr);
}

extern "C" int asdf_tc2_eval_debug(const void* static_dev, const void* sample_dev, const asdf_query* q,
                                   float* out_hand_dev, float* out_obj_dev, int32_t* bbox_dev, void* stream,
                                   void* debug_dev) {
  using namespace asdf;
  ASDF_REQUIRE(static_dev && sample_dev && q && out_hand_dev && out_obj_dev, "asdf_tc2_eval: null argument");
  ASDF_REQUIRE(q->end >= q->begin, "negative query range");
  if (q->mode == ASDF_QUERY_POINTS) {
    ASDF_REQUIRE(q->points_dev && q->point_stride >= 3, "points query needs xyz rows");
  } else {
    ASDF_REQUIRE(q->mode == ASDF_QUERY_GRID_REFERENCE || q->mode == ASDF_QUERY_GRID_REGULAR, "bad query mode");
    ASDF_REQUIRE(q->N >= 2 && q->begin >= 0 && q->end <= (int64_t)q->N * q->N * q->N, "grid range outside N^3");
  }
  if (q->end == q->begin) return ASDF_OK;
  static bool configured = false;
  if (!configured) {
    ASDF_CUDA_CHECK(cudaFuncSetAttribute(tc2::tc2_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemBytes));
    ASDF_CUDA_CHECK(cudaFuncSetAttribute(tc2::tc2_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemBytes));
    configured = true;
  }
  int dev = 0, sms = 0;
  ASDF_CUDA_CHECK(cudaGetDevice(&dev));
  ASDF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (q->end - q->begin + tc2::kPtsPerTile - 1) / tc2::kPtsPerTile;
  int64_t clusters = sms / 2;
  if (n_tiles < clusters) clusters = n_tiles;
  tc2::Args a;
  a.q = *q; a.stat = (const uint8_t*)static_dev; a.samp = (const uint8_t*)sample_dev;
  a.out_hand = out_hand_dev; a.out_obj = out_obj_dev; a.bbox = bbox_dev; a.dbg = (long long*)debug_dev;
  if (debug_dev)
    tc2::tc2_eval_kernel<true><<<(unsigned)(2 * clusters), tc2::kThreads, tc2::kSmemBytes, (cudaStream_t)stream>>>(a);
  else
    tc2::tc2_eval_kernel<false><<<(unsigned)(2 * clusters), tc2::kThreads, tc2::kSmemBytes, (cudaStream_t)stream>>>(a);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}
