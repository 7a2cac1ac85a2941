// K1 on the 5th-generation tensor cores: fused grid-generation + pose-align + two 5-layer
// 512-wide decoders, activations never leave the SM pair.
//
// Replaces the hot loops utils/mesh.py:46-63 and :96-115 (per chunk: H2D copy,
// kinematic_embedding utils/utils.py:376-430, decode_sdf_multi_output :561-572,
// SeparateDecoder.forward networks/model.py:285-350, two D2H copies) and the nonzero()-based
// bounding box of utils/mesh.py:207-247.
//
// Shape of the computation (per decoder, after the host-side folding of packer.py):
//   x1 = relu(M0 p + B0)               3 -> 512   CUDA cores, written straight into the A operand
//   x2 = relu(W1 x1 + b1)            512 -> 256   tcgen05.mma  (h = 250 padded to 256)
//   x3 = relu(W2 x2 + M2 p + B2)     256 -> 512   tcgen05.mma, point term in the epilogue
//   x4 = relu(W3 x3 + b3)            512 -> 512   tcgen05.mma
//   sdf = tanh(w4 . x4 + b4)         512 -> 1     in the layer-3 epilogue
// Precision: every product is issued three times in fp16 (a_hi b_hi + a_lo b_hi + a_hi b_lo with
// a = a_hi + a_lo, b = b_hi + b_lo, fp32 accumulation in TMEM), which keeps |sdf - fp32 reference|
// around 1e-7 (contract: 1e-5); single-pass fp16/bf16 misses the contract.
//
// Mapping: a CTA pair (cluster of 2, tcgen05 cta_group::2) owns a tile of 128 consecutive grid
// points, 64 per CTA.  One UMMA is M=128 (64 rows per CTA) x N=256 x K=16; each CTA supplies its
// own 64 activation rows (A, K-major, 128B swizzle, written by the epilogue warps) and one half
// (128 rows) of the weight tile (B, streamed L2 -> SMEM with cp.async.bulk from a pre-swizzled
// packed stream).  Accumulators: 4 TMEM buffers of 128 columns ("2x2" layout: lanes 0-63 hold
// n<128, lanes 64-127 hold n>=128).
//
// Warp roles per CTA (384 threads): warp 0 weight-stream producer, warp 1 UMMA issuer (leader CTA)
// / full-barrier relay (peer CTA), warp 2 TMEM allocator, warps 4-11 epilogue (activation,
// fp16 split, operand write-back, final dot + tanh, bbox).
#include "common.cuh"
#include <cuda_fp16.h>

namespace asdf {
namespace tc {

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;               // first epilogue warp
constexpr int kEpiThreads = 256;
constexpr int kPtsPerCta = 64;
constexpr int kPtsPerTile = 128;           // per CTA pair
constexpr int kChunkK = 64;                // K elements per swizzle atom row (128 B of fp16)
constexpr int kTileBytes = 128 * kChunkK * 2;       // one B tile: 128 rows x 64 k  = 16 KiB
constexpr int kASlotBytes = kPtsPerCta * kChunkK * 2;  // one A chunk: 64 rows x 64 k =  8 KiB
constexpr int kNumASlots = 8;
constexpr int kRing = 5;                   // B tiles in flight
constexpr int kTilesPerDecoder = 64;       // 16 (L1) + 16 (L2) + 32 (L3)
constexpr int kStaticParamFloats = 256 + 1024 + 8;   // b1*t | (b3, w4) pairs | b4, 1/s1, 1/s2, 1/(s3 t)
constexpr int64_t kWeightBytes = (int64_t)2 * 2 * kTilesPerDecoder * kTileBytes;   // [dec][cta][tile]
constexpr int kSampleFloatsPerDecoder = 2 * 512 * 4;   // M0B0[512][4] | M2B2[512][4] (both x act_scale)

// shared memory map (bytes, relative to a 1024-aligned base)
constexpr int kOffAHi = 0;
constexpr int kOffALo = kOffAHi + kNumASlots * kASlotBytes;          //  65536
constexpr int kOffRing = kOffALo + kNumASlots * kASlotBytes;         // 131072
constexpr int kOffM2 = kOffRing + kRing * kTileBytes;                // 212992
constexpr int kOffB1 = kOffM2 + 512 * 16;
constexpr int kOffB3W4 = kOffB1 + 256 * 4;
constexpr int kOffRed = kOffB3W4 + 512 * 8;                          // [4][64] partial sums
constexpr int kOffMisc = kOffRed + 4 * 64 * 4;                       // 8 floats of scalars
constexpr int kOffPts = kOffMisc + 64;                               // [64] float4 points of this CTA's rows
constexpr int kOffBar = kOffPts + 64 * 16;
// barriers (8 B each)
constexpr int kBarFull = 0;                       // [kRing]   leader: own tx + peer relay arrive
constexpr int kBarFullLocal = kBarFull + kRing;   // [kRing]   peer CTA: own tx only
constexpr int kBarEmpty = kBarFullLocal + kRing;  // [kRing]   UMMA commit (multicast)
constexpr int kBarAFull = kBarEmpty + kRing;      // [8]       A chunk written by both CTAs
constexpr int kBarTmemFull = kBarAFull + kNumASlots;  // [4]   accumulator complete (multicast)
constexpr int kNumBars = kBarTmemFull + 4;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;   // + slack for the 1024 B alignment

static_assert(kSmemBytes <= 232448, "exceeds the 227 KiB dynamic shared memory limit");

// instruction descriptor: D=f32, A=B=f16, both K-major, N=256, M=128 (cta_group::2)
constexpr uint32_t kIdesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster.  Plain (CTA-scope release)
// arrive as in CUTLASS' ClusterBarrier::arrive: the payload it orders lives in this CTA's shared
// memory and has already been made visible to the async proxy (fence.proxy.async / TMA
// complete_tx); `.release.cluster` would compile to MEMBAR.ALL.GPU on every call.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      :: "r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle shared memory matrix descriptor (SBO = 1024 B between 8-row groups)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all previously issued UMMAs retired) on the barrier at this offset in both CTAs
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

// explicit shared-state-space accesses (32-bit shared addresses): keeps ptxas on LDS/STS
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ float lds_f1(uint32_t a) {
  float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kEpiThreads) : "memory"); }

// ------------------------------------------------------------------------------------------
// operand write-back: 8 consecutive k of one row, split into fp16 hi + lo, 128B-swizzled K-major
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8_store(uint32_t a_hi_slot, uint32_t a_lo_slot, int row, int k8, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = fminf(v[2 * i], 60000.f), b = fminf(v[2 * i + 1], 60000.f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
  const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((k8 ^ (row & 7)) << 4);
  sts_u4(a_hi_slot + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
  sts_u4(a_lo_slot + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
}

struct Args {
  asdf_tc_desc d;
  asdf_query q;
  const uint8_t* stat;     // packed weight stream followed by the static parameter block
  const float* samp;       // [2][kSampleFloatsPerDecoder]
  float* out_hand;
  float* out_obj;
  int32_t* bbox;
  long long* dbg;          // optional phase timing (cluster 0): see asdf_tc_desc.debug_dev
};

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <bool kDebug>   // kDebug: clock64() phase stamps + experiment flags (tools/tc_phase_timing.py)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) tc_eval_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];    // 128B-swizzle atoms need 1024 B alignment
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  auto bar = [&](int i) { return sbase + kOffBar + 8 * i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar(kBarFull + i), 2);        // own expect_tx arrive + peer relay arrive
      mbar_init(bar(kBarFullLocal + i), 1);
      mbar_init(bar(kBarEmpty + i), 1);
    }
    for (int i = 0; i < kNumASlots; ++i) mbar_init(bar(kBarAFull + i), 16);  // see publish()
    for (int i = 0; i < 4; ++i) mbar_init(bar(kBarTmemFull + i), 1);
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
  const uint32_t tmem_base = *tmem_ptr_s;

  const int64_t total = a.q.end - a.q.begin;
  const int64_t n_tiles = (total + kPtsPerTile - 1) / kPtsPerTile;
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // =========================== weight-stream producer ===========================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const bool dbg_nocopy = kDebug && (a.dbg[15] & 1);   // timing experiment: skip the copies
      for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
        for (int dec = 0; dec < 2; ++dec) {
          const uint8_t* src = a.stat + ((int64_t)(dec * 2 + rank) * kTilesPerDecoder) * kTileBytes;
          for (int i = 0; i < kTilesPerDecoder; ++i) {
            mbar_wait(bar(kBarEmpty + slot), phase ^ 1);
            const uint32_t fb = bar((leader ? kBarFull : kBarFullLocal) + slot);
            if (dbg_nocopy) {
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(fb) : "memory");
              if (++slot == kRing) { slot = 0; phase ^= 1; }
              continue;
            }
            mbar_expect_tx(fb, kTileBytes);
            bulk_g2s(sbase + kOffRing + slot * kTileBytes, src + (int64_t)i * kTileBytes, kTileBytes, fb);
            if (++slot == kRing) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      if (!leader) {
        // ======================= peer CTA: relay "my half of the tile landed" =======================
        uint32_t slot = 0, phase = 0;
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
          for (int i = 0; i < 2 * kTilesPerDecoder; ++i) {
            mbar_wait(bar(kBarFullLocal + slot), phase);
            mbar_arrive_cluster(bar(kBarFull + slot), 0);
            if (++slot == kRing) { slot = 0; phase ^= 1; }
          }
        }
      } else {
        // =================================== UMMA issuer ===================================
        uint32_t slot = 0, phase = 0, a_phase = 0 /* bit per A slot */;
        const uint32_t a_hi = sbase + kOffAHi, a_lo = sbase + kOffALo, ring = sbase + kOffRing;
        long long w_a = 0, w_b = 0, t_begin = kDebug ? clock64() : 0;
        // one K chunk (64) of one N block: hi tile then lo tile of the ring
        auto chunk = [&](int a_slot, uint32_t d_tmem, bool first) {
          const uint32_t ah = a_hi + a_slot * kASlotBytes, al = a_lo + a_slot * kASlotBytes;
          long long t0 = kDebug ? clock64() : 0;
          mbar_wait(bar(kBarFull + slot), phase);
          if (kDebug) w_b += clock64() - t0;
          tc_fence_after();
          uint32_t b = ring + slot * kTileBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_cg2(d_tmem, smem_desc(ah + ks * 32), smem_desc(b + ks * 32), kIdesc, (first && ks == 0) ? 0u : 1u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_cg2(d_tmem, smem_desc(al + ks * 32), smem_desc(b + ks * 32), kIdesc, 1u);
          umma_commit_both(bar(kBarEmpty + slot));
          if (++slot == kRing) { slot = 0; phase ^= 1; }
          if (kDebug) t0 = clock64();
          mbar_wait(bar(kBarFull + slot), phase);
          if (kDebug) w_b += clock64() - t0;
          tc_fence_after();
          b = ring + slot * kTileBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_cg2(d_tmem, smem_desc(ah + ks * 32), smem_desc(b + ks * 32), kIdesc, 1u);
          umma_commit_both(bar(kBarEmpty + slot));
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        };
        auto wait_a = [&](int a_slot) {
          const long long t0 = kDebug ? clock64() : 0;
          mbar_wait(bar(kBarAFull + a_slot), (a_phase >> a_slot) & 1u);
          if (kDebug) w_a += clock64() - t0;
          a_phase ^= 1u << a_slot;
          tc_fence_after();
        };
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
          for (int dec = 0; dec < 2; ++dec) {
            // layer 1: x1 (slots 0..7) -> buffer 0
            for (int kc = 0; kc < 8; ++kc) { wait_a(kc); chunk(kc, tmem_base + 0 * 128, kc == 0); }
            umma_commit_both(bar(kBarTmemFull + 0));
            // layer 2: x2 (slots 0..3) -> buffers 1, 2
            for (int nb = 0; nb < 2; ++nb) {
              for (int kc = 0; kc < 4; ++kc) { if (nb == 0) wait_a(kc); chunk(kc, tmem_base + (1 + nb) * 128, kc == 0); }
              umma_commit_both(bar(kBarTmemFull + 1 + nb));
            }
            // layer 3: x3 (k<256 in slots 4..7, k>=256 in slots 0..3) -> buffers 3, 0
            for (int nb = 0; nb < 2; ++nb) {
              for (int j = 0; j < 8; ++j) {
                const int s = (j + 4) & 7;
                if (nb == 0) wait_a(s);
                chunk(s, tmem_base + (nb == 0 ? 3 : 0) * 128, j == 0);
              }
              umma_commit_both(bar(kBarTmemFull + (nb == 0 ? 3 : 0)));
            }
          }
        }
        if (kDebug && cluster_id == 0) {
          a.dbg[0] = clock64() - t_begin; a.dbg[1] = w_a; a.dbg[2] = w_b;
        }
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // =================================== epilogue warps ===================================
    const int e = warp - kEpiWarp0;            // 0..7
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int ch = e >> 2;                     // 0: 32-column groups {0,2}; 1: groups {1,3} of a buffer
    const int row = (q & 1) * 32 + lane;       // point row of this thread's TMEM lane (0..63)
    const int nhalf = q >> 1;                  // accumulator n-half held by this lane quadrant
    const int et = threadIdx.x - kEpiWarp0 * 32;   // 0..255
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t sM2 = sbase + kOffM2, sB1 = sbase + kOffB1, sB3W4 = sbase + kOffB3W4;
    const uint32_t sRed = sbase + kOffRed, sMisc = sbase + kOffMisc, sPts = sbase + kOffPts;
    const uint32_t a_hi = sbase + kOffAHi, a_lo = sbase + kOffALo;
    const float* sparams = reinterpret_cast<const float*>(a.stat + kWeightBytes);
    uint32_t tphase = 0;                       // bit per TMEM buffer
    auto wait_tmem = [&](int buf) {
      mbar_wait(bar(kBarTmemFull + buf), (tphase >> buf) & 1u);
      tphase ^= 1u << buf;
      tc_fence_after();
    };
    // publish "this warp's share of A slot s is written": every slot phase collects 16 arrivals
    // (x1: 1 warp x 2 CTAs x 8 lanes;  x2 / x3: 4 warps x 2 CTAs x 2 lanes)
    auto publish = [&](int s, int lanes) {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane < lanes) mbar_arrive_cluster(bar(kBarAFull + s), 0);
    };
    // the 8 KiB buffer sM2 is time-shared: (M0,B0) rows during x1 generation, (M2,B2) afterwards
    auto load_m0 = [&](int dec) {
      const float4* g4 = reinterpret_cast<const float4*>(a.samp + (size_t)dec * kSampleFloatsPerDecoder);
      sts_f4(sM2 + 16 * et, __ldg(g4 + et)); sts_f4(sM2 + 16 * (et + 256), __ldg(g4 + et + 256));
    };
    auto load_params = [&](int dec) {
      // b1 + M2B2 + (b3,w4) + scalars of decoder `dec`; all 256 epilogue threads cooperate
      const float* samp = a.samp + (size_t)dec * kSampleFloatsPerDecoder;
      const float* sp = sparams + (size_t)dec * kStaticParamFloats;
      const float4* g4 = reinterpret_cast<const float4*>(samp + 2048);
      sts_f4(sM2 + 16 * et, __ldg(g4 + et)); sts_f4(sM2 + 16 * (et + 256), __ldg(g4 + et + 256));
      sts_f1(sB1 + 4 * et, __ldg(sp + et));
      const float2* g2 = reinterpret_cast<const float2*>(sp + 256);
      sts_f2(sB3W4 + 8 * et, __ldg(g2 + et)); sts_f2(sB3W4 + 8 * (et + 256), __ldg(g2 + et + 256));
      if (et < 8) sts_f1(sMisc + 4 * et, __ldg(sp + 256 + 1024 + et));
    };
    auto point_of = [&](int64_t i, float& x, float& y, float& z) {
      x = y = z = 0.f;
      if (i < a.q.end) {
        if (a.q.mode == ASDF_QUERY_POINTS) {
          const float* r = a.q.points_dev + (size_t)i * a.q.point_stride;
          x = __ldg(r); y = __ldg(r + 1); z = __ldg(r + 2);
        } else {
          grid_point(i, a.q.N, a.q.mode, a.q.voxel, a.q.origin[0], a.q.origin[1], a.q.origin[2], x, y, z);
        }
      }
    };

    long long ph[12];
    for (int z = 0; z < 12; ++z) ph[z] = 0;
    const bool stamp = kDebug && cluster_id == 0 && rank == 0 && et == 0;
#define ASDF_STAMP(k) do { if (stamp) { const long long _t = clock64(); ph[k] += _t - tlast; tlast = _t; } } while (0)
    long long tlast = kDebug ? clock64() : 0;
    load_m0(0);                                // first work item; later ones are prefetched in the layer-3 epilogue
    for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
      const int64_t ibase = a.q.begin + t * kPtsPerTile + rank * kPtsPerCta;
      const int64_t i = ibase + row;
      const bool live = i < a.q.end;
      float px, py, pz;
      point_of(i, px, py, pz);                 // point of this thread's TMEM lane
      epi_bar_sync();                          // previous tile's x1 generation is done with sPts
      if (et < kPtsPerCta) sts_f4(sPts + 16 * et, make_float4(px, py, pz, 0.f));   // row == et for warps 4,5
      epi_bar_sync();
      for (int dec = 0; dec < 2; ++dec) {
        // ---------------- x1 = relu(M0 p + B0) -> A slots 0..7 ----------------
        // warp e owns A slot e (64 features); a lane owns 8 of them (one 16-byte chunk) for all rows.
        // Its 8 (M0,B0) rows stay in registers; the rows' points come from shared memory.
        {
          float4 m[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = lds_f4(sM2 + 16 * (e * 64 + (lane & 7) * 8 + j));
#pragma unroll 2
          for (int it = 0; it < 16; ++it) {
            const int r = (lane >> 3) + 4 * it;
            const float4 p = lds_f4(sPts + 16 * r);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j] = fmaxf(fmaf(m[j].x, p.x, fmaf(m[j].y, p.y, fmaf(m[j].z, p.z, m[j].w))), 0.f);
            split8_store(a_hi + e * kASlotBytes, a_lo + e * kASlotBytes, r, lane & 7, v);
          }
          publish(e, 8);
        }
        ASDF_STAMP(1);
        epi_bar_sync();                       // everybody is done with the previous decoder's parameters
        load_params(dec);
        epi_bar_sync();
        const float inv1 = lds_f1(sMisc + 4), inv2 = lds_f1(sMisc + 8), inv3 = lds_f1(sMisc + 12), b4 = lds_f1(sMisc);
        // ---------------- layer-1 epilogue: x2 -> A slots 0..3 ----------------
        ASDF_STAMP(2);
        wait_tmem(0);
        ASDF_STAMP(3);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {        // chunk inside this lane quadrant's n-half
          const int s = nhalf * 2 + cc;
          const int col0 = cc * 64 + ch * 32;   // this warp's 32 columns of the chunk
          float acc[32];
          tmem_ld32(tmem_base + lane_addr + 0 * 128 + col0, acc);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j] = fmaxf(fmaf(acc[g * 8 + j], inv1, lds_f1(sB1 + 4 * (nhalf * 128 + col0 + g * 8 + j))), 0.f);
            split8_store(a_hi + s * kASlotBytes, a_lo + s * kASlotBytes, row, ch * 4 + g, v);
          }
          publish(s, 2);
        }
        // ---------------- layer-2 epilogue: x3 -> A slots 4..7 (nb=0), 0..3 (nb=1) ----------------
        ASDF_STAMP(4);
#pragma unroll 1
        for (int nb = 0; nb < 2; ++nb) {
          wait_tmem(1 + nb);
          ASDF_STAMP(5);
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int s = (nb == 0 ? 4 : 0) + nhalf * 2 + cc;
            const int col0 = cc * 64 + ch * 32;
            float acc[32];
            tmem_ld32(tmem_base + lane_addr + (1 + nb) * 128 + col0, acc);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 m = lds_f4(sM2 + 16 * (nb * 256 + nhalf * 128 + col0 + g * 8 + j));
                const float pt = fmaf(m.x, px, fmaf(m.y, py, fmaf(m.z, pz, m.w)));
                v[j] = fmaxf(fmaf(acc[g * 8 + j], inv2, pt), 0.f);
              }
              split8_store(a_hi + s * kASlotBytes, a_lo + s * kASlotBytes, row, ch * 4 + g, v);
            }
            publish(s, 2);
          }
          ASDF_STAMP(6);
        }
        // ---------------- layer-3 epilogue: partial dot with w4 ----------------
        epi_bar_sync();                       // every warp is done with (M2,B2): prefetch the next item's (M0,B0)
        load_m0(dec ^ 1);
        float part = 0.f;
#pragma unroll 1
        for (int nb = 0; nb < 2; ++nb) {
          const int buf = nb == 0 ? 3 : 0;
          wait_tmem(buf);
          ASDF_STAMP(7);
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int col0 = cc * 64 + ch * 32;
            float acc[32];
            tmem_ld32(tmem_base + lane_addr + buf * 128 + col0, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 bw = lds_f2(sB3W4 + 8 * (nb * 256 + nhalf * 128 + col0 + j));
              part = fmaf(fmaxf(fmaf(acc[j], inv3, bw.x), 0.f), bw.y, part);
            }
          }
          ASDF_STAMP(8);
        }
        tc_fence_before();
        sts_f1(sRed + 4 * ((nhalf * 2 + ch) * 64 + row), part);
        epi_bar_sync();
        if (et < kPtsPerCta) {
          // threads 0..63 are warps 4,5: row == et for them (q = 0,1 -> rows 0..31, 32..63)
          const float s4 = lds_f1(sRed + 4 * et) + lds_f1(sRed + 4 * (64 + et)) + lds_f1(sRed + 4 * (128 + et)) +
                           lds_f1(sRed + 4 * (192 + et));
          const float val = tanhf(s4 + b4);
          if (live) (dec == 0 ? a.out_hand : a.out_obj)[i - a.q.begin] = val;
          if (a.bbox && a.q.mode != ASDF_QUERY_POINTS && (a.q.bbox_mask >> dec & 1))
            bbox_update(a.bbox + 6 * dec, live && val < 0.f, i, a.q.N);
        }
        ASDF_STAMP(9);
      }
    }
    if (stamp) for (int z = 0; z < 12; ++z) a.dbg[4 + z] = ph[z];
  }

  // ---------------------------------- teardown ----------------------------------
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// self test: D[128 x 256] = A[128 x 64] . B[256 x 64]^T through exactly the same operand layouts,
// descriptors, cta_group::2 UMMA, commit and TMEM read-back as the production kernel.
//   a_rows: [128][64] fp16 row-major (rows 0..63 -> CTA 0, 64..127 -> CTA 1)
//   b_tiles: [2][16 KiB] pre-swizzled tiles (tile c holds B rows 128c .. 128c+127)
//   d_out: [128][256] f32
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_selftest_kernel(const __half* __restrict__ a_rows, const uint8_t* __restrict__ b_tiles, float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  auto bar = [&](int i) { return sbase + kOffBar + 8 * i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);
  if (warp == 1 && lane == 0) {
    mbar_init(bar(kBarFull), 2);
    mbar_init(bar(kBarFullLocal), 1);
    mbar_init(bar(kBarAFull), 8);      // 4 staging warps x 2 CTAs
    mbar_init(bar(kBarTmemFull), 1);
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
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t fb = bar(rank == 0 ? kBarFull : kBarFullLocal);
      mbar_expect_tx(fb, kTileBytes);
      bulk_g2s(sbase + kOffRing, b_tiles + (size_t)rank * kTileBytes, kTileBytes, fb);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      if (rank != 0) {
        mbar_wait(bar(kBarFullLocal), 0);
        mbar_arrive_cluster(bar(kBarFull), 0);
      } else {
        mbar_wait(bar(kBarAFull), 0);
        mbar_wait(bar(kBarFull), 0);
        tc_fence_after();
        for (int ks = 0; ks < 4; ++ks)
          umma_f16_cg2(tmem_base, smem_desc(sbase + kOffAHi + ks * 32), smem_desc(sbase + kOffRing + ks * 32),
                       kIdesc, ks == 0 ? 0u : 1u);
        umma_commit_both(bar(kBarTmemFull));
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    const int e = warp - kEpiWarp0, q = warp & 3, ch = e >> 2;
    const int row = (q & 1) * 32 + lane, nhalf = q >> 1;
    // stage A: warps with nhalf==0 write k chunks [ch*4, ch*4+4) of their 32 rows (plain fp16, no split)
    if (nhalf == 0) {
      for (int k8 = ch * 4; k8 < ch * 4 + 4; ++k8) {
        const uint4 v = *reinterpret_cast<const uint4*>(a_rows + (size_t)(rank * 64 + row) * 64 + k8 * 8);
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((k8 ^ (row & 7)) << 4);
        sts_u4(sbase + kOffAHi + off, v);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0 && nhalf == 0) mbar_arrive_cluster(bar(kBarAFull), 0);   // 4 warps x 2 CTAs = 8 arrivals
    mbar_wait(bar(kBarTmemFull), 0);
    tc_fence_after();
    for (int c32 = 0; c32 < 2; ++c32) {
      float acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 64 + c32 * 32, acc);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j)
        d_out[(size_t)(rank * 64 + row) * 256 + nhalf * 128 + ch * 64 + c32 * 32 + j] = acc[j];
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace tc
}  // namespace asdf

extern "C" int64_t asdf_tc_static_bytes(void) {
  return asdf::tc::kWeightBytes + (int64_t)2 * asdf::tc::kStaticParamFloats * 4;
}
extern "C" int64_t asdf_tc_sample_floats(void) { return 2 * asdf::tc::kSampleFloatsPerDecoder; }

extern "C" int asdf_tc_eval(const asdf_tc_desc* desc, const void* static_dev, const float* sample_dev,
                            const asdf_query* q, float* out_hand_dev, float* out_obj_dev, int32_t* bbox_dev,
                            void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(desc && static_dev && sample_dev && q && out_hand_dev && out_obj_dev, "asdf_tc_eval: null argument");
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
    ASDF_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
    ASDF_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
    configured = true;
  }
  int dev = 0, sms = 0;
  ASDF_CUDA_CHECK(cudaGetDevice(&dev));
  ASDF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (q->end - q->begin + tc::kPtsPerTile - 1) / tc::kPtsPerTile;
  int64_t clusters = sms / 2;
  if (n_tiles < clusters) clusters = n_tiles;
  tc::Args a;
  a.d = *desc; a.q = *q; a.stat = (const uint8_t*)static_dev; a.samp = sample_dev;
  a.out_hand = out_hand_dev; a.out_obj = out_obj_dev; a.bbox = bbox_dev;
  a.dbg = (long long*)desc->debug_dev;
  if (a.dbg)
    tc::tc_eval_kernel<true><<<(unsigned)(2 * clusters), tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>(a);
  else
    tc::tc_eval_kernel<false><<<(unsigned)(2 * clusters), tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>(a);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_tc_selftest(const void* a_rows_dev, const void* b_tiles_dev, float* d_out_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(a_rows_dev && b_tiles_dev && d_out_dev, "asdf_tc_selftest: null argument");
  static bool configured = false;
  if (!configured) {
    ASDF_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
    configured = true;
  }
  tc::tc_selftest_kernel<<<2, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>(
      (const __half*)a_rows_dev, (const uint8_t*)b_tiles_dev, d_out_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}
