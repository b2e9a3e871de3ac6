// Developer microbenchmark + correctness probe: what does an M=128 N=48 K=16 kind::f16 tcgen05.mma cost when its A tile
// starts 1 or 2 pixels into a staged row (the dx taps of the row-streaming convs), for
//   layout 0  K-major no-swizzle planes (conv_row.cu / conv_mux.cu today): pixel stride 16 B, start += 16*dx
//             -> every 8-row core matrix straddles two 128-B lines
//   layout 1  K-major SWIZZLE_64B rows: one pixel = 64 B = [hi 16 ch | lo 16 ch], 16-B chunk c of pixel p stored at
//             p*64 + ((c ^ ((p >> 1) & 3)) << 4); start += 64*dx (+32 for the lo K-slice), SBO = 512
// and is the swizzled, shifted read CORRECT (the XOR must follow absolute address bits), with which base-offset field?
// Every configuration is checked against integers computed on the host.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/swz_probe tools/probe/swz_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_swz64(uint32_t saddr, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7u) << 49) | (4ull << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__host__ __device__ inline int a_val(int p, int k) { return (p * 7 + k * 3) % 13 - 6; }   // pixel p, K index k (0..31)
__host__ __device__ inline int b_val(int n, int k) { return (n * 5 + k) % 7 - 3; }        // row n, K index k (0..15)

constexpr int N = 48, PLANE = 136;

// layout: 0 no-swizzle planes, 1 swizzle-64B rows.  dx: pixel shift, part: K-slice (0 = hi, 1 = lo), bo_mode: base offset field
// 0 -> 0, 1 -> (start >> 7) & 7.  out[0..1] = cycles, D -> dout[128][48]
__global__ void __launch_bounds__(128, 1) probe(int layout, int dx, int part, int bo_mode, int L, long long* out, float* dout) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char* As = smem;               // 8704 B
  unsigned char* Bs = smem + 16 * 1024;   // 2 K-chunks x 48 rows x 16 B
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < PLANE * 32; e += 128) {
    const int p = e / 32, k = e % 32;
    const __half h = __int2half_rn(a_val(p, k));
    uint32_t off;
    if (layout == 0) {
      // planes: [part][K-chunk][pixel][8 ch]
      const int pt = k / 16, kc = (k % 16) / 8, c8 = k % 8;
      off = (uint32_t)(((pt * 2 + kc) * PLANE + p) * 16 + c8 * 2);
    } else {
      const int c = k / 8, c8 = k % 8;
      off = (uint32_t)(p * 64 + ((c ^ ((p >> 1) & 3)) << 4) + c8 * 2);
    }
    *reinterpret_cast<__half*>(As + off) = h;
  }
  for (int e = threadIdx.x; e < N * 16; e += 128) {
    const int n = e / 16, k = e % 16;
    *reinterpret_cast<__half*>(Bs + ((k / 8) * N + n) * 16 + (k % 8) * 2) = __int2half_rn(b_val(n, k));
  }
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
  const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
  uint64_t ad;
  if (layout == 0) {
    ad = desc_noswz(a0 + (uint32_t)(part * 2 * PLANE * 16 + dx * 16), PLANE * 16, 128);
  } else {
    const uint32_t sa = a0 + (uint32_t)(dx * 64 + part * 32);
    ad = desc_swz64(sa, bo_mode ? (sa >> 7) & 7u : 0u);
  }
  const uint64_t bd = desc_noswz(b0, N * 16, 128);
  uint32_t parity = 0;
  if (threadIdx.x < 32) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    const bool leader = pred != 0;
    // rep 0..2: timing (accumulating garbage into columns 64..), then one clean MMA into columns 0..47
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64();
      if (leader) {
        if (rep < 3) {
          for (int i = 0; i < L; ++i) mma(tb + 64u + (uint32_t)((i & 3) * N), ad, bd, idesc, 1u);
        } else {
          mma(tb, ad, bd, idesc, 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
      parity ^= 1;
      long long t2 = clock64();
      if (rep == 2 && threadIdx.x == 0) out[0] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tb + ((uint32_t)(w * 32) << 16) + (uint32_t)c0)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) dout[(w * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* out;
  float* dout;
  cudaMalloc(&out, 16);
  cudaMalloc(&dout, 128 * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int L = 256;
  static float h[128 * N];
  printf("%6s %3s %4s %3s | %9s | %s\n", "layout", "dx", "part", "bo", "clk/MMA", "check");
  for (int layout = 0; layout < 2; ++layout)
    for (int dx = 0; dx < 3; ++dx)
      for (int part = 0; part < 2; ++part)
        for (int bo = 0; bo < (layout ? 2 : 1); ++bo) {
          cudaMemset(dout, 0, sizeof(h));
          probe<<<1, 128, 64 * 1024>>>(layout, dx, part, bo, L, out, dout);
          long long t;
          cudaError_t e = cudaMemcpy(&t, out, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
          int bad = 0, first = -1;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
              int ref = 0;
              for (int k = 0; k < 16; ++k) ref += a_val(m + dx, part * 16 + k) * b_val(n, k);
              if (h[m * N + n] != (float)ref) { if (first < 0) first = m * N + n; ++bad; }
            }
          printf("%6d %3d %4d %3d | %9.1f | %s", layout, dx, part, bo, (double)t / L, bad ? "WRONG" : "ok");
          if (bad) printf(" (%d of %d, first at m=%d n=%d)", bad, 128 * N, first / N, first % N);
          printf("\n");
        }
  return 0;
}
