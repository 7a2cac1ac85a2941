// Probe: how fast can one SM stream an L2-resident buffer into shared memory with cp.async.bulk,
// as a function of ring depth R, tile size T and whether all SMs read in lockstep?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu && ./stream_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* buf, size_t buf_bytes, int R, int T, int n_tiles, int stagger,
                                               long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + R * T;
  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t tiles_in_buf = buf_bytes / T;
    size_t pos = stagger ? (size_t)blockIdx.x * 37 % tiles_in_buf : 0;
    long long t0 = clock64();
    // prime the ring
    for (int i = 0; i < R && i < n_tiles; ++i) {
      expect_tx(bars + 8 * i, T);
      bulk(sbase + i * T, buf + ((pos + i) % tiles_in_buf) * T, T, bars + 8 * i);
    }
    uint32_t phase = 0; int slot = 0;
    for (int i = 0; i < n_tiles; ++i) {
      mbar_wait(bars + 8 * slot, phase);
      if (i + R < n_tiles) {          // consumer is instantaneous: refill right away
        expect_tx(bars + 8 * slot, T);
        bulk(sbase + slot * T, buf + ((pos + i + R) % tiles_in_buf) * T, T, bars + 8 * slot);
      }
      if (++slot == R) { slot = 0; phase ^= 1; }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t buf_bytes = 4 << 20;
  uint8_t* buf; cudaMalloc(&buf, buf_bytes); cudaMemset(buf, 1, buf_bytes);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("%6s %3s %8s %8s | %12s %12s\n", "T", "R", "stagger", "grid", "B/cyc/SM", "cyc/tile");
  for (int grid : {148, 1}) for (int stagger : {0, 1}) for (int T : {8192, 16384, 32768}) for (int R : {2, 4, 5, 6, 8, 12}) {
    if ((size_t)R * T > 200 * 1024) continue;
    const int n_tiles = 4096;
    probe<<<grid, 128, 226 * 1024>>>(buf, buf_bytes, R, T, n_tiles, stagger, cyc);
    probe<<<grid, 128, 226 * 1024>>>(buf, buf_bytes, R, T, n_tiles, stagger, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    printf("%6d %3d %8d %8d | %12.1f %12.1f\n", T, R, stagger, grid, (double)n_tiles * T / avg, avg / n_tiles);
  }
  return 0;
}
