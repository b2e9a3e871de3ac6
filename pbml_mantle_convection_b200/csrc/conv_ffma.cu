// fp32 CUDA-core (FFMA) direct convolution, k in {3,5}, 'same', stride 1, over a list of
// blocked sources with the producers' GroupNorm(+GELU) folded into the tile load and the
// padding mode (zeros / replicate / reflect) folded into the load coordinates.
//
// Restates (reference repo paths):
//   SymmetricConv2d.forward            symmetric_layers_torch.py:113-138   (filters pre-expanded at pack time)
//   nn.Conv2d heads conv[1..3]         pytorch_networks_convae.py:1263-1309
//   FluidLayer conv -> GroupNorm -> GELU  :790-799  (GN+GELU of the PRODUCER applied on load here;
//                                          this layer's own GN statistics are reduced in the epilogue)
//   torch.cat over levels + inputs     :1327, :1332  (sources are read in place; no concat tensor)
//
// This is the generic / reference-accuracy kernel (any C_in, C_out, both kernel sizes).  The
// tensor-core path (conv_umma.cu) covers the hot 16-channel shapes.
#include "common.cuh"

namespace pbmc {

constexpr int CF_TX = 16;               // threads along x
constexpr int CF_TY = 8;                // threads along y
constexpr int CF_PX = 4;                // pixels per thread, interleaved by CF_TX (conflict-free LDS.128)
constexpr int CF_TW = CF_TX * CF_PX;    // 64-pixel wide tile
constexpr int CF_TH = CF_TY;            // 8 rows
constexpr int CF_THREADS = CF_TX * CF_TY;
constexpr int CF_STAGE = 4;             // channel blocks (16 channels) staged in smem per K-step

struct ConvFfmaParams {
  pbmc_src src[PBMC_MAX_SRC];
  int nsrc, B, H, W;
  int cout_blks, cin_blks, pad_mode, epi_act;
  const float* wpk;
  const float* bias;
  float* out;
  double* out_stats;
  double* out_chan_sum;
};

template <int KS, int COB>
__global__ void __launch_bounds__(CF_THREADS, (KS == 3 ? 3 : 2)) conv_ffma_kernel(const ConvFfmaParams p) {
  constexpr int P = KS / 2;
  constexpr int TWP = CF_TW + 2 * P;
  constexpr int THP = CF_TH + 2 * P;
  constexpr int NCO = COB * 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);                 // [CF_STAGE][THP][TWP]
  float* w_s = reinterpret_cast<float*>(tile + CF_STAGE * THP * TWP); // [CF_STAGE][KS*KS][4][16]
  float* xf_a = w_s + CF_STAGE * KS * KS * 4 * 16;                    // [cin_blks*4]
  float* xf_b = xf_a + p.cin_blks * 4;
  double* red = reinterpret_cast<double*>(xf_b + p.cin_blks * 4 + ((p.cin_blks * 8) & 1));  // 8B aligned

  const int tid = threadIdx.x;
  const int tx = tid % CF_TX, ty = tid / CF_TX;
  const int n_cg = (p.cout_blks + 3) / 4;
  const int b = blockIdx.z / n_cg;
  const int cg = blockIdx.z % n_cg;
  const int x0 = blockIdx.x * CF_TW, y0 = blockIdx.y * CF_TH;
  const int H = p.H, W = p.W;
  const size_t plane = (size_t)H * W;

  // ---- fused-GroupNorm coefficients for every input channel of batch element b
  {
    int c0 = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const pbmc_src& S = p.src[s];
      for (int c = tid; c < S.nblk * 4; c += CF_THREADS) {
        float a = 1.f, bb = 0.f;
        if (S.xform == PBMC_XFORM_GN_GELU || S.xform == PBMC_XFORM_GN)
          gn_coeffs(S.stats + ((size_t)b * S.nblk + (c >> 2)) * 2, S.inv_count, S.gamma[c], S.beta[c], a, bb);
        xf_a[c0 + c] = a;
        xf_b[c0 + c] = bb;
      }
      c0 += S.nblk * 4;
    }
  }
  __syncthreads();

  float acc[CF_PX][NCO];
#pragma unroll
  for (int j = 0; j < CF_PX; ++j)
#pragma unroll
    for (int c = 0; c < NCO; ++c) acc[j][c] = 0.f;

  int gblk0 = 0;  // global channel-block index of the current source's first block
  for (int s = 0; s < p.nsrc; ++s) {
    const pbmc_src S = p.src[s];
    for (int cb = 0; cb < S.nblk; cb += CF_STAGE) {
      const int nb = min(CF_STAGE, S.nblk - cb);
      // ---- stage load: input tile (+halo) with transform and padding folded in
      const int n_el = nb * THP * TWP;
      for (int e = tid; e < n_el; e += CF_THREADS) {
        const int c = e % TWP;
        const int r = (e / TWP) % THP;
        const int k = e / (TWP * THP);
        const int sy = pad_index(y0 + r - P, H, p.pad_mode);
        const int sx = pad_index(x0 + c - P, W, p.pad_mode);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sy >= 0 && sx >= 0) {
          v = ldg4(S.ptr + (((size_t)b * S.nblk + cb + k) * plane + (size_t)sy * W + sx) * 4);
          const int ch = (gblk0 + cb + k) * 4;
          v = xform4(v, xf_a + ch, xf_b + ch, S.xform);
        }
        tile[e] = v;
      }
      // ---- stage weights: [nb][KS*KS][4][16] contiguous in the packed image
      {
        const float4* wsrc = reinterpret_cast<const float4*>(
            p.wpk + ((size_t)cg * p.cin_blks + gblk0 + cb) * (KS * KS * 4 * 16));
        float4* wdst = reinterpret_cast<float4*>(w_s);
        const int n4 = nb * KS * KS * 4 * 4;
        for (int e = tid; e < n4; e += CF_THREADS) wdst[e] = __ldg(wsrc + e);
      }
      __syncthreads();
      // ---- FFMA
      for (int k = 0; k < nb; ++k) {
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          float4 in[CF_PX][KS];
          const float4* trow = tile + (k * THP + ty + dy) * TWP + tx;
#pragma unroll
          for (int j = 0; j < CF_PX; ++j)
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) in[j][dx] = trow[j * CF_TX + dx];
#pragma unroll
          for (int dx = 0; dx < KS; ++dx) {
            const float* wt = w_s + ((k * KS * KS + dy * KS + dx) * 4) * 16;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              float w[NCO];
#pragma unroll
              for (int q = 0; q < COB; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(wt + ci * 16 + q * 4);
                w[q * 4 + 0] = w4.x; w[q * 4 + 1] = w4.y; w[q * 4 + 2] = w4.z; w[q * 4 + 3] = w4.w;
              }
#pragma unroll
              for (int j = 0; j < CF_PX; ++j) {
                const float4 iv = in[j][dx];
                const float xin = ci == 0 ? iv.x : (ci == 1 ? iv.y : (ci == 2 ? iv.z : iv.w));
#pragma unroll
                for (int c = 0; c < NCO; ++c) acc[j][c] = fmaf(xin, w[c], acc[j][c]);
              }
            }
          }
        }
      }
      __syncthreads();
    }
    gblk0 += S.nblk;
  }

  // ---- epilogue: bias, activation, store, statistics
  const int gy = y0 + ty;
  float s1[COB], s2[COB], cs[NCO];
#pragma unroll
  for (int q = 0; q < COB; ++q) s1[q] = s2[q] = 0.f;
#pragma unroll
  for (int c = 0; c < NCO; ++c) cs[c] = 0.f;
#pragma unroll
  for (int q = 0; q < COB; ++q) {
    const int ob = cg * 4 + q;
    if (ob >= p.cout_blks) continue;
    const float4 bias = ldg4(p.bias + ob * 4);
#pragma unroll
    for (int j = 0; j < CF_PX; ++j) {
      const int gx = x0 + tx + j * CF_TX;
      if (gy < H && gx < W) {
        float4 v = make_float4(acc[j][q * 4 + 0] + bias.x, acc[j][q * 4 + 1] + bias.y, acc[j][q * 4 + 2] + bias.z,
                               acc[j][q * 4 + 3] + bias.w);
        if (p.epi_act == PBMC_ACT_GELU) {
          v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        }
        *reinterpret_cast<float4*>(p.out + (((size_t)b * p.cout_blks + ob) * plane + (size_t)gy * W + gx) * 4) = v;
        s1[q] += (v.x + v.y) + (v.z + v.w);
        s2[q] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        cs[q * 4 + 0] += v.x; cs[q * 4 + 1] += v.y; cs[q * 4 + 2] += v.z; cs[q * 4 + 3] += v.w;
      }
    }
  }
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NWARP = CF_THREADS / 32;
  if (p.out_stats != nullptr) {
#pragma unroll
    for (int q = 0; q < COB; ++q) {
      const double a = warp_sum((double)s1[q]);
      const double c = warp_sum((double)s2[q]);
      if (lane == 0) { red[(warp * COB + q) * 2] = a; red[(warp * COB + q) * 2 + 1] = c; }
    }
    __syncthreads();
    if (tid < COB * 2) {
      const int q = tid >> 1;
      const int ob = cg * 4 + q;
      if (ob < p.cout_blks) {
        double t = 0.0;
        for (int w = 0; w < NWARP; ++w) t += red[(w * COB + q) * 2 + (tid & 1)];
        atomicAdd(p.out_stats + ((size_t)b * p.cout_blks + ob) * 2 + (tid & 1), t);
      }
    }
    __syncthreads();
  }
  if (p.out_chan_sum != nullptr) {
#pragma unroll
    for (int c = 0; c < NCO; ++c) {
      const double a = warp_sum((double)cs[c]);
      if (lane == 0) red[warp * NCO + c] = a;
    }
    __syncthreads();
    if (tid < NCO && cg * 16 + tid < p.cout_blks * 4) {
      double t = 0.0;
      for (int w = 0; w < NWARP; ++w) t += red[w * NCO + tid];
      atomicAdd(p.out_chan_sum + (size_t)b * p.cout_blks * 4 + cg * 16 + tid, t);
    }
  }
}

template <int KS, int COB>
static int launch_ffma(const ConvFfmaParams& p, cudaStream_t st) {
  constexpr int P = KS / 2;
  const size_t smem = (size_t)CF_STAGE * (CF_TH + 2 * P) * (CF_TW + 2 * P) * sizeof(float4) +
                      (size_t)CF_STAGE * KS * KS * 4 * 16 * sizeof(float) + (size_t)p.cin_blks * 8 * sizeof(float) + 8 +
                      (size_t)(CF_THREADS / 32) * 16 * sizeof(double);
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    PBMC_CUDA(cudaFuncSetAttribute(conv_ffma_kernel<KS, COB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  if (smem > 160 * 1024) return PBMC_ERR_UNSUPPORTED;
  const int n_cg = (p.cout_blks + 3) / 4;
  dim3 grid(cdiv(p.W, CF_TW), cdiv(p.H, CF_TH), p.B * n_cg);
  if (grid.y > 65535 || grid.z > 65535) return PBMC_ERR_BAD_SHAPE;
  conv_ffma_kernel<KS, COB><<<grid, CF_THREADS, smem, st>>>(p);
  PBMC_CHECK_LAUNCH("conv_ffma_kernel");
  return PBMC_OK;
}

int conv_ffma_dispatch(const pbmc_conv_desc& d, cudaStream_t st) {
  ConvFfmaParams p;
  int cin = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    p.src[s] = d.src[s];
    cin += d.src[s].nblk;
  }
  p.nsrc = d.nsrc; p.B = d.B; p.H = d.H; p.W = d.W;
  p.cout_blks = (d.cout + 3) / 4;
  p.cin_blks = cin;
  p.pad_mode = d.pad_mode; p.epi_act = d.epi_act;
  p.wpk = d.wpk; p.bias = d.bias; p.out = d.out; p.out_stats = d.out_stats; p.out_chan_sum = d.out_chan_sum;
  const int cob = p.cout_blks >= 3 ? 4 : p.cout_blks;
  if (d.ksize == 3) {
    if (cob == 1) return launch_ffma<3, 1>(p, st);
    if (cob == 2) return launch_ffma<3, 2>(p, st);
    return launch_ffma<3, 4>(p, st);
  } else if (d.ksize == 5) {
    if (cob == 1) return launch_ffma<5, 1>(p, st);
    if (cob == 2) return launch_ffma<5, 2>(p, st);
    return launch_ffma<5, 4>(p, st);
  }
  return PBMC_ERR_UNSUPPORTED;
}

}  // namespace pbmc
