// Probe: mixing kind::f16 (A from TMEM) and kind::f8f6f4 (e4m3, A from shared memory, K=32) UMMAs into ONE fp32
// accumulator (cta_group::2, M=256, N=128): numerical check + cycles per 64-wide K chunk for the
// "fp16 main product + fp8 correction products" schedule of k1_tc3.cu versus the fp16x3 schedule.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

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


__device__ __forceinline__ void mma8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// a16: [256][64] fp16 row-major; a8: [2][128 rows][128 B] pre-swizzled fp8 (16 KB per CTA);
// b16: [2][64 rows][64 k] pre-swizzled fp16 (8 KB per CTA); b8: [2][64 rows][128 B] pre-swizzled fp8 (8 KB per CTA)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __half* a16, const uint8_t* a8, const uint8_t* b16, const uint8_t* b8, float* d, int iters, int variant, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 100 * 1024, tptr = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ctarank();
  const int pair = blockIdx.x >> 1;
  // smem: B16 at 0 (8 KB), B8 at 8 KB (8 KB), A8 at 16 KB (16 KB), fp16 A_lo stand-in at 32 KB (16 KB)
  for (int i = threadIdx.x; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(b16 + rank * 8192)[i];
  for (int i = threadIdx.x; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 8192)[i] = reinterpret_cast<const uint32_t*>(b8 + rank * 8192)[i];
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 16384)[i] = reinterpret_cast<const uint32_t*>(a8 + rank * 16384)[i];
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
  {
    const int row = rank * 128 + warp * 32 + lane;
    uint32_t w[32];
    for (int j = 0; j < 32; ++j) w[j] = reinterpret_cast<const uint32_t*>(a16 + (size_t)row * 64)[j];
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
  const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);   // same word for f16 x f16 and e4m3 x e4m3
  if (warp == 0 && lane == 0 && rank == 0) {
    // correctness: D = A16.B16^T (K = 64 fp16, A from TMEM) + A8.B8^T (K = 128 e4m3, A from smem)
    for (int ks = 0; ks < 4; ++ks) mma_ts(tmem, tmem + 256 + ks * 8, smem_desc(sbase + ks * 32), idesc, ks ? 1u : 0u);
    for (int ks = 0; ks < 4; ++ks) mma8_ss(tmem, smem_desc(sbase + 16384 + ks * 32), smem_desc(sbase + 8192 + ks * 32), idesc, 1u);
    commit(bar);
    mbar_wait(bar, 0);
    const long long t0 = clock64();
    const uint32_t D = tmem + 128;
    for (int i = 0; i < iters; ++i) {
      if (variant == 0) {            // v2: fp16x3
        for (int ks = 0; ks < 4; ++ks) mma_ts(D, tmem + 256 + ks * 8, smem_desc(sbase + ks * 32), idesc, 1u);
        for (int ks = 0; ks < 4; ++ks) mma_ss(D, smem_desc(sbase + 32768 + ks * 32), smem_desc(sbase + ks * 32), idesc, 1u);
        for (int ks = 0; ks < 4; ++ks) mma_ts(D, tmem + 256 + ks * 8, smem_desc(sbase + 8192 + ks * 32), idesc, 1u);
      } else if (variant == 1) {     // v3 grouped: 4 x f16 TS then 4 x f8 SS
        for (int ks = 0; ks < 4; ++ks) mma_ts(D, tmem + 256 + ks * 8, smem_desc(sbase + ks * 32), idesc, 1u);
        for (int ks = 0; ks < 4; ++ks) mma8_ss(D, smem_desc(sbase + 16384 + ks * 32), smem_desc(sbase + 8192 + ks * 32), idesc, 1u);
      } else if (variant == 2) {     // v3 interleaved
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(D, tmem + 256 + ks * 8, smem_desc(sbase + ks * 32), idesc, 1u);
          mma8_ss(D, smem_desc(sbase + 16384 + ks * 32), smem_desc(sbase + 8192 + ks * 32), idesc, 1u);
        }
      } else if (variant == 3) {     // f8 SS only
        for (int r = 0; r < 2; ++r)
          for (int ks = 0; ks < 4; ++ks) mma8_ss(D, smem_desc(sbase + 16384 + ks * 32), smem_desc(sbase + 8192 + ks * 32), idesc, 1u);
      } else {                       // f16 TS only
        for (int r = 0; r < 2; ++r)
          for (int ks = 0; ks < 4; ++ks) mma_ts(D, tmem + 256 + ks * 8, smem_desc(sbase + ks * 32), idesc, 1u);
      }
    }
    commit(bar);
    mbar_wait(bar, 1);
    cyc[pair] = clock64() - t0;
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  csync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (pair == 0) {
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

static inline int swz(int r, int byte) { return (r / 8) * 1024 + (r % 8) * 128 + ((((byte / 16) ^ (r % 8))) * 16) + (byte % 16); }

int main() {
  std::vector<__half> a(256 * 64), b(128 * 64);
  for (auto& x : a) x = __float2half((float)((rand() % 17) - 8));
  for (auto& x : b) x = __float2half((float)((rand() % 9) - 4));
  const float vals[8] = {0.5f, -1.5f, 2.f, 0.0625f, -3.5f, 1.f, -0.25f, 7.f};     // exactly representable in e4m3
  std::vector<float> a8f(256 * 128), b8f(128 * 128);
  for (auto& x : a8f) x = vals[rand() % 8];
  for (auto& x : b8f) x = vals[rand() % 8];
  std::vector<uint8_t> b16t(2 * 8192), b8t(2 * 8192), a8t(2 * 16384);
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < 64; ++r) {
      for (int k = 0; k < 64; ++k) *reinterpret_cast<__half*>(&b16t[c * 8192 + swz(r, 2 * k)]) = b[(c * 64 + r) * 64 + k];
      for (int k = 0; k < 128; ++k) b8t[c * 8192 + swz(r, k)] = __nv_cvt_float_to_fp8(b8f[(c * 64 + r) * 128 + k], __NV_SATFINITE, __NV_E4M3);
    }
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 128; ++k) a8t[c * 16384 + swz(r, k)] = __nv_cvt_float_to_fp8(a8f[(c * 128 + r) * 128 + k], __NV_SATFINITE, __NV_E4M3);
  __half* da; uint8_t *da8, *db16, *db8; float* dd; long long* dc;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&da8, a8t.size()); cudaMalloc(&db16, b16t.size()); cudaMalloc(&db8, b8t.size());
  cudaMalloc(&dd, 256 * 128 * 4); cudaMalloc(&dc, 8 * 74);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(da8, a8t.data(), a8t.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db16, b16t.data(), b16t.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db8, b8t.data(), b8t.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const char* names[5] = {"v2 fp16x3 (8 TS16 + 4 SS16)", "v3 grouped (4 TS16 + 4 SS8)", "v3 interleaved", "8 x SS8", "8 x TS16"};
  for (int grid : {2, 148})
    for (int variant = 0; variant < 5; ++variant) {
      const int iters = 2000;
      cudaMemset(dd, 0xff, 256 * 128 * 4);
      probe<<<grid, 128, 120 * 1024>>>(da, da8, db16, db8, dd, iters, variant, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
      std::vector<float> d(256 * 128);
      std::vector<long long> cyc(74);
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(cyc.data(), dc, 8 * (grid / 2), cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 256; ++m)
        for (int n = 0; n < 128; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)__half2float(a[m * 64 + k]) * (double)__half2float(b[n * 64 + k]);
          for (int k = 0; k < 128; ++k) ref += (double)a8f[m * 128 + k] * (double)b8f[n * 128 + k];
          maxerr = fmax(maxerr, fabs(ref - d[m * 128 + n]));
        }
      long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      printf("grid %3d  %-30s mixed-kind accumulate max|err| = %g   %.1f cycles per 64-k chunk\n", grid, names[variant], maxerr, (double)mx / iters);
    }
  return 0;
}
