// Probe: tcgen05.mma with A from TMEM (cta_group::2, M=256, N=128/256), A written by tcgen05.st.
// Checks the packed-fp16 TMEM layout of A (lane = row, 32-bit column = k/2, low half = even k)
// and measures the rate of the fp16x3 sequence  Ahi(tmem).Bhi + Alo(smem).Bhi + Ahi(tmem).Blo.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
               :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(bar), "h"((uint16_t)3) : "memory");
}

// a: [256][64] fp16 row-major; b_tiles: [2][64 rows][64 k] pre-swizzled (8 KB each); d: [256][128] f32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __half* a, const uint8_t* b_tiles, float* d, int iters, int N, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 100 * 1024, tptr = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ctarank();
  // B tile (this CTA's 64 rows) -> smem offset 0 (8 KB); A_lo-style smem operand (all ones) at 16 KB (128 rows x 128 B)
  for (int i = threadIdx.x; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(b_tiles + rank * 8192)[i];
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 16384)[i] = 0x3c003c00u;
  // a second, larger B image for N=256 rate runs (128 rows) at 32 KB: just ones
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 32768)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(tptr) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  csync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 100 * 1024 + 16);
  // ---- A (this thread's row) -> TMEM columns [256, 288): 64 fp16 = 32 packed words, low half = even k
  {
    const int row = rank * 128 + warp * 32 + lane;
    uint32_t w[32];
    for (int j = 0; j < 32; ++j) w[j] = reinterpret_cast<const uint32_t*>(a + (size_t)row * 64)[j];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
                    "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]),
                    "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]),
                    "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  csync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc128 = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
  const uint32_t idescN = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((256u >> 4) << 24);
  if (warp == 0 && lane == 0 && rank == 0) {
    // correctness: D[256 x 128] = A . B^T, K = 64
    for (int ks = 0; ks < 4; ++ks) mma_ts(tmem, tmem + 256 + ks * 8, smem_desc(sbase + ks * 32), idesc128, ks ? 1u : 0u);
    commit(bar);
    mbar_wait(bar, 0);
    // rate: fp16x3 sequence into columns [128, 128+N) (garbage values, timing only)
    const long long t0 = clock64();
    const uint32_t boff = N == 256 ? 32768u : 0u;
    for (int i = 0; i < iters; ++i)
      for (int ks = 0; ks < 4; ++ks) {
        mma_ts(tmem + 128, tmem + 256 + ks * 8, smem_desc(sbase + boff + ks * 32), idescN, 1u);
        mma_ss(tmem + 128, smem_desc(sbase + 16384 + ks * 32), smem_desc(sbase + boff + ks * 32), idescN, 1u);
        mma_ts(tmem + 128, tmem + 256 + ks * 8, smem_desc(sbase + boff + ks * 32), idescN, 1u);
      }
    commit(bar);
    mbar_wait(bar, 1);
    cyc[0] = clock64() - t0;
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  csync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read back D columns [0,128) of this thread's row
  {
    const int row = rank * 128 + warp * 32 + lane;
    for (int c32 = 0; c32 < 4; ++c32) {
      uint32_t r[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                   "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                     "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                     "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c32 * 32) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) d[(size_t)row * 128 + c32 * 32 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  csync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

int main() {
  std::vector<__half> a(256 * 64), b(128 * 64);
  for (auto& x : a) x = __float2half((float)((rand() % 17) - 8));
  for (auto& x : b) x = __float2half((float)((rand() % 9) - 4));
  std::vector<uint8_t> tiles(2 * 8192);
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < 64; ++r)
      for (int k = 0; k < 64; ++k) {
        const int off = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2;
        *reinterpret_cast<__half*>(&tiles[c * 8192 + off]) = b[(c * 64 + r) * 64 + k];
      }
  __half* da; uint8_t* db; float* dd; long long* dc;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, tiles.size()); cudaMalloc(&dd, 256 * 128 * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, tiles.data(), tiles.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int N : {128, 256}) {
    const int iters = 500;
    cudaMemset(dd, 0xff, 256 * 128 * 4);
    probe<<<2, 128, 120 * 1024>>>(da, db, dd, iters, N, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<float> d(256 * 128);
    long long cyc;
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < 128; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)__half2float(a[m * 64 + k]) * (double)__half2float(b[n * 64 + k]);
        maxerr = fmax(maxerr, fabs(ref - d[m * 128 + n]));
      }
    printf("TS cg2 M=256: layout check max|err| = %g   fp16x3 rate at N=%d: %.1f cycles per UMMA (floor %d)\n",
           maxerr, N, (double)cyc / (iters * 12.0), N / 2);
  }
  return 0;
}
