// Developer microbenchmark: how long does one tcgen05.mma (M=128, kind::f16, K=16) take as a function of N,
// of the operand layout (no-swizzle vs 128B-swizzle K-major), and of whether consecutive MMAs accumulate into
// the same TMEM columns?  One CTA per SM, one issuing thread; cycles = clock64 around L MMAs + commit + wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mma_probe tools/probe/mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_swz128(uint32_t saddr) {  // K-major, 128B swizzle: SBO = 1024 B, LBO unused
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__global__ void __launch_bounds__(128, 1) probe(int N, int L, int nacc, int swz, int a_step, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x < 32) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    const bool leader = pred != 0;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    uint64_t ad[8];
    uint32_t dd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t aa = a0 + (uint32_t)(i * a_step);
      ad[i] = swz ? desc_swz128(aa) : desc_noswz(aa, 2176, 128);
      dd[i] = tb + (uint32_t)((i % nacc) * N);
    }
    const uint64_t bd = swz ? desc_swz128(b0) : desc_noswz(b0, N * 16, 128);
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (leader) {
        for (int i = 0; i < L; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) mma(dd[j], ad[j], bd, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      long long t1 = clock64();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
      parity ^= 1;
      long long t2 = clock64();
      if (blockIdx.x == 0 && rep == 2 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int L = 256;
  printf("%5s %5s %4s %7s | %10s %10s\n", "N", "nacc", "swz", "a_step", "issue/MMA", "total/MMA");
  for (int swz = 0; swz < 2; ++swz)
    for (int N : {16, 32, 48, 64, 96, 128, 256})
      for (int nacc : {1, 4})
        for (int a_step : {0, 8704}) {
          if (nacc * N > 512) continue;
          probe<<<148, 128, 200 * 1024>>>(N, L, nacc, swz, a_step, out);
          cudaError_t le = cudaGetLastError();
          if (le != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(le)); return 1; }
          long long h[2];
          cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          printf("%5d %5d %4d %7d | %10.1f %10.1f\n", N, nacc, swz, a_step, (double)h[0] / L, (double)h[1] / L);
        }
  return 0;
}
