// Probe: how does tcgen05.mma (kind::f16, fp32 accumulate) round?  D[0][0] of a 128x16x16 UMMA whose only
// non-zero operand rows are A[0][:] and B[0][:], after a scripted sequence of accumulating UMMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o acc_round_probe acc_round_probe.cu && ./acc_round_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int kMaxSteps = 8;
struct Script { int n; float a[kMaxSteps][16]; float b[kMaxSteps][16]; };

// one CTA; step s: A row 0 = a[s], B row 0 = b[s]; D (+)= A.B^T
__global__ void __launch_bounds__(128, 1) probe(Script sc, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t offB = 16 * 1024, bar = sbase + 32 * 1024, tptr = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(tptr) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 32 * 1024 + 16);
  const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t parity = 0;
  for (int s = 0; s < sc.n; ++s) {
    for (int i = threadIdx.x; i < 8 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;     // A (16K) + B (16K)
    __syncthreads();
    if (threadIdx.x < 16) {
      reinterpret_cast<__half*>(smem)[threadIdx.x] = __float2half_rn(sc.a[s][threadIdx.x]);          // row 0: unswizzled
      reinterpret_cast<__half*>(smem + offB)[threadIdx.x] = __float2half_rn(sc.b[s][threadIdx.x]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mma(tmem, smem_desc(sbase), smem_desc(sbase + offB), idesc, s > 0 ? 1u : 0u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
      mbar_wait(bar, parity);
    }
    parity ^= 1;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (lane == 0) out[0] = __uint_as_float(v);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tmem) : "memory");
}

static float run(const Script& sc) {
  float* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  probe<<<1, 128, 40 * 1024>>>(sc, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  float h; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost); cudaFree(out);
  return h;
}

static void show(const char* what, float got, double exact) {
  const float rn = (float)exact;
  float rz = rn; if (fabs((double)rz) > fabs(exact)) rz = nextafterf(rz, 0.f);
  float rd = rn; if ((double)rd > exact) rd = nextafterf(rd, -INFINITY);
  printf("%-58s got %.9g (%a)  exact %.12g  RN %a RZ %a RD %a -> %s\n", what, got, got, exact, rn, rz, rd,
         got == rn && got != rz ? "RN" : (got == rz && got != rn ? "RZ" : (got == rn && got == rz ? "exact/ambiguous" : (got == rd ? "RD" : "OTHER"))));
}

int main() {
  const float u = ldexpf(1.f, -24);
  auto step = [](Script& s, int i) { memset(s.a[i], 0, sizeof s.a[i]); memset(s.b[i], 0, sizeof s.b[i]); };
  {   // across instructions
    for (float sign : {1.f, -1.f}) for (float frac : {0.5f, 0.75f, 1.5f, 1.75f}) for (float isign : {1.f, -1.f}) {
      Script s; s.n = 2; step(s, 0); step(s, 1);
      s.a[0][0] = sign; s.b[0][0] = 1.f;
      s.a[1][0] = isign * frac * 2.f; s.b[1][0] = u;           // increment = +-frac * 2^-23 (ulp of 1.0 is 2^-23)
      char w[96]; snprintf(w, sizeof w, "D=%+.0f then += %+.2f ulp (separate UMMA)", sign, isign * frac);
      show(w, run(s), (double)sign + (double)isign * frac * 2.0 * u);
    }
  }
  {   // inside one instruction: sixteen small products on top of an accumulator
    for (float frac : {0.375f, 0.125f, 0.03125f}) {
      Script s; s.n = 2; step(s, 0); step(s, 1);
      s.a[0][0] = 1.f; s.b[0][0] = 1.f;
      for (int k = 0; k < 16; ++k) { s.a[1][k] = frac * 2.f; s.b[1][k] = u; }
      char w[96]; snprintf(w, sizeof w, "D=1 then += 16 x %.5f ulp in ONE UMMA", frac);
      show(w, run(s), 1.0 + 16.0 * frac * 2.0 * u);
    }
  }
  {   // inside one instruction, no prior accumulator: one big product + fifteen small ones
    for (float frac : {0.375f, 0.125f, 0.03125f, 0.0078125f}) {
      Script s; s.n = 1; step(s, 0);
      s.a[0][0] = 1.f; s.b[0][0] = 1.f;
      for (int k = 1; k < 16; ++k) { s.a[0][k] = frac * 2.f; s.b[0][k] = u; }
      char w[96]; snprintf(w, sizeof w, "one UMMA: 1.0 + 15 x %.7f ulp", frac);
      show(w, run(s), 1.0 + 15.0 * frac * 2.0 * u);
    }
  }
  {   // big accumulator, product needing more than 24 bits itself: (1+2^-10)^2 on top of 2^12
    Script s; s.n = 2; step(s, 0); step(s, 1);
    s.a[0][0] = 64.f; s.b[0][0] = 64.f;
    s.a[1][0] = 1.f + ldexpf(1.f, -10); s.b[1][0] = 1.f + ldexpf(1.f, -10);
    show("D=4096 then += (1+2^-10)^2", run(s), 4096.0 + (1.0 + ldexp(1.0, -10)) * (1.0 + ldexp(1.0, -10)));
  }
  return 0;
}
