// Developer probe: tcgen05.mma with the A operand in TENSOR MEMORY (M = 128, kind::f16, K = 16).
// (1) checks the A layout assumption used by csrc/conv_mux.cu's TS path: lane = row m, 32-bit column c holds
//     K elements (2c, 2c+1) (low half = 2c), written with tcgen05.st.32x32b.x8;
// (2) times L back-to-back MMAs for several N, next to the shared-memory-A form.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/ts_probe tools/probe/ts_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int N, int L, int mode, float* dout, long long* tout) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B[n][k] = ((n*2 + k) % 5) - 2, K-major no-swizzle: chunk c = k/8 at c*LBO, row n at (n/8)*128 + (n%8)*16
  __half* Bs = reinterpret_cast<__half*>(smem);
  const uint32_t B_LBO = (uint32_t)N * 16;
  for (int e = tid; e < N * 16; e += 128) {
    const int n = e / 16, k = e % 16;
    const size_t off = (size_t)(k / 8) * B_LBO + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
    *reinterpret_cast<__half*>(smem + off) = __float2half((float)(((n * 2 + k) % 5) - 2));
  }
  // A in shared memory too (for the SS timing): [2 chunks][128 rows][16 B]
  unsigned char* As = smem + 32 * 1024;
  for (int e = tid; e < 128 * 16; e += 128) {
    const int m = e / 16, k = e % 16;
    const size_t off = (size_t)(k / 8) * 2048 + (size_t)(m / 8) * 128 + (size_t)(m % 8) * 16 + (size_t)(k % 8) * 2;
    *reinterpret_cast<__half*>(As + off) = __float2half((float)(((m * 3 + k * 5) % 7) - 3));
  }
  (void)Bs;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  const uint32_t a_col = 256;  // A lives at columns [256, 264)
  // thread m writes its row of A: column c = (A[m][2c], A[m][2c+1])
  {
    const int m = tid;
    uint32_t r[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const __half lo = __float2half((float)(((m * 3 + (2 * c) * 5) % 7) - 3));
      const __half hi = __float2half((float)(((m * 3 + (2 * c + 1) * 5) % 7) - 3));
      r[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + a_col;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t bd = desc_noswz(smem_u32(smem), B_LBO, 128);
  const uint64_t ad = desc_noswz(smem_u32(As), 2048, 128);
  uint32_t parity = 0;
  if (warp == 0) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    const bool leader = pred != 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      if (leader) {
        for (int i = 0; i < L; ++i) {
          if (mode == 0) mma_ts(tb, tb + a_col, bd, idesc, i > 0 ? 1u : 0u);
          else mma_ss(tb, ad, bd, idesc, i > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(&bar)), "r"(parity)
                     : "memory");
      }
      parity ^= 1;
      const long long t1 = clock64();
      if (tid == 0) tout[rep] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read D: thread m reads N columns
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) dout[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  float* dout;
  long long* tout;
  cudaMalloc(&dout, 128 * 256 * sizeof(float));
  cudaMalloc(&tout, 8 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  float* h = (float*)malloc(128 * 256 * sizeof(float));
  printf("   N    L mode |  clk/MMA  | max |D - ref| (L=1 check)\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 48, 96, 144, 256}) {
      // correctness with L = 1
      probe<<<1, 128, 64 * 1024>>>(N, 1, mode, dout, tout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d mode=%d: %s\n", N, mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, dout, 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < 16; ++k) ref += (double)(((m * 3 + k * 5) % 7) - 3) * (double)(((n * 2 + k) % 5) - 2);
          const double d = fabs((double)h[m * N + n] - ref);
          if (d > maxerr) maxerr = d;
        }
      long long t[3];
      const int L = 64;
      probe<<<1, 128, 64 * 1024>>>(N, L, mode, dout, tout);
      cudaDeviceSynchronize();
      cudaMemcpy(t, tout, sizeof(t), cudaMemcpyDeviceToHost);
      printf("%4d %4d %s | %8.1f  | %g\n", N, L, mode == 0 ? "TS" : "SS", (double)t[2] / L, maxerr);
    }
  return 0;
}
