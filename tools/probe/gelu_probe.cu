// Developer microbenchmark: cycles for W warps per SM to apply GN+GELU+fp16 split to 16 values per thread.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../pbml_mantle_convection_b200/csrc/common.cuh"
#include "../../pbml_mantle_convection_b200/csrc/tc05.cuh"
using namespace pbmc;
template <int MODE>
__global__ void probe(const float* in, uint4* out, long long* cyc, int iters) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = in[threadIdx.x * 16 + i];
  __syncthreads();
  long long t0 = clock64();
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = fmaf(v[i], 1.01f, 0.001f * it);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = gelu_erf(w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) gelu_erf2(w[i], w[i + 1]);
    }
    uint4 h0, l0, h1, l1;
    split_f16(w, h0, l0);
    split_f16(w + 8, h1, l1);
    acc.x ^= h0.x ^ l0.y ^ h1.z ^ l1.w; acc.y ^= h0.y ^ l0.x ^ h1.w ^ l1.z; acc.z ^= h0.z ^ l0.w ^ h1.x ^ l1.y; acc.w ^= h0.w ^ l0.z ^ h1.y ^ l1.x;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* in; uint4* out; long long* cyc;
  cudaMalloc(&in, 1024 * 16 * 4); cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 1024 * 16 * 4);
  for (int mode = 0; mode < 2; ++mode)
  for (int warps : {1, 4, 12, 16, 32}) {
    if (mode == 0) probe<0><<<148, warps * 32>>>(in, out, cyc, 200); else probe<1><<<148, warps * 32>>>(in, out, cyc, 200);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mode %d warps/SM %2d: %.1f clk per iteration (16 GN+GELU+split per thread) -> %.1f clk per warp-iteration-per-SMSP\n", warps, mode, warps, h / 200.0, h / 200.0 / ((warps + 3) / 4));
  }
  return 0;
}
