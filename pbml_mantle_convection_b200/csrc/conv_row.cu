// Row-streaming, warp-specialised tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
// Same operator contract as conv_ffma.cu / conv_umma.cu (sources with the producer's GroupNorm+GELU
// fused into the load, padding folded into the load coordinates, bias / GELU / GroupNorm statistics /
// zero-mean sums in the epilogue), restating
//   SymmetricConv2d.forward symmetric_layers_torch.py:113-138, nn.Conv2d heads
//   pytorch_networks_convae.py:1263-1309, FluidLayer :790-799, concat :1327/:1332.
//
// GEMM view (k = KS, P = k/2).  A CTA owns a strip of 128 output columns x `rpc` output rows and
// streams the strip's input rows through a shared-memory ring.  For ONE input row r (128 + k - 1
// positions, staged once as fp16 hi|lo channel-chunk planes = UMMA K-major / no-swizzle canonical
// layout) and one 16-channel input group, k tcgen05.mma (x3 for the hi/lo split) with
//     M = 128 output columns, N = k*16 = (dy, c_out), K = 16 input channels
// accumulate   D_r[x][(dy, co)] += sum_{dx, ci} X[r][x + dx][ci] * W[dy][dx][ci][co]
// where the dx shift is the A descriptor's start address (+dx*16 B) and ALL k vertical taps ride in
// the N dimension: the A tile is read from shared memory once per (dx, pass) instead of once per tap
// (tcgen05.mma with N = 16 is shared-memory-read bound, see conv_umma.cu).  Output row y is then
//     out[y][x][co] = bias + sum_dy D_{y+dy}[x][(dy, co)]
// i.e. k TMEM loads from the SAME lane -- no shuffles, no junk lanes, no halo columns in M.
// D_r live in a TMEM ring of ND accumulators (N columns each).
//
// Warp roles (14 warps):  0-3 epilogue (TMEM lane quarter == warp),  4-11 producers (two groups of
// 128 threads, one thread per position, alternate stages, register prefetch of the next stage),
// 12 MMA issuer (one elected lane) + TMEM allocator,  13 halo producer (the k-1 extra positions).
// Pipelines: a_full/a_empty (producers <-> MMA, NSTAGE smem stages; one stage = one input row of one
// 16-channel group), d_full/d_empty (MMA <-> epilogue, ND accumulators).
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace pbmc {

constexpr int CR_MMA_WARP = 12;
constexpr int CR_HALO_WARP = 13;
constexpr int CR_THREADS = 14 * 32;
constexpr int CR_MAXG = 24;
constexpr int CR_SMEM_HDR = 2176;

struct ConvRowParams {
  pbmc_src src[PBMC_MAX_SRC];
  int nsrc, B, H, W;
  int cout_blks, pad_mode, epi_act;
  int ngroups;  // 16-channel K groups over the concatenated sources
  int cin_ch;   // padded channel count of the concatenation (sum nblk * 4)
  int rpc;      // output rows per CTA
  const void* wpk;  // [group][dx][part][2 K-chunks][N = k*16 rows (dy, c_out)][8 c_in] 16-bit
  const float* bias;
  float* out;
  double* out_stats;
  double* out_chan_sum;
};

struct RowGroup {
  const float* base;  // first block of the group for this sample
  int nb;             // real 4-channel blocks (1..4); the rest of the 16 channels are zero
  int xform;
  int chan0;          // first channel in the xf_a / xf_b tables
  int pad;
};

template <int KS, int PARTS>
struct RowGeom {
  static constexpr int P = KS / 2;
  static constexpr int N = KS * 16;
  static constexpr int PWS = 128 + KS - 1;
  static constexpr int PLANE = (PWS + 7) / 8 * 8;   // positions per K-chunk plane (16 B each)
  static constexpr int PART_BYTES = 2 * PLANE * 16;  // two K chunks (8 channels each)
  static constexpr int STAGE_BYTES = PARTS * PART_BYTES;
  static constexpr int NSTAGE = 4;
  static constexpr int ND = KS == 3 ? 5 : 6;
  static constexpr uint32_t TMEM_COLS = ND * N <= 256 ? 256 : 512;
  static constexpr int B_TILE = 2 * N * 16;  // one (group, dx, part) operand: [2 chunks][N rows][16 B]
  static constexpr int B_GROUP = KS * PARTS * B_TILE;
  static_assert(ND * N <= 512, "TMEM has 512 columns");
  static_assert(ND >= KS + 1, "the MMA must be able to run ahead of the epilogue");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// the wait names the destination registers as in/out operands so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait16(uint32_t r[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__host__ __device__ constexpr uint32_t row_idesc(uint32_t fmt, uint32_t n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ float xform1(float v, float a, float b, int xform) {
  if (xform == PBMC_XFORM_NONE) return v;
  if (xform != PBMC_XFORM_GELU) v = fmaf(v, a, b);
  if (xform != PBMC_XFORM_GN) v = gelu_erf(v);
  return v;
}

template <int KS, int PARTS>
__global__ void __launch_bounds__(CR_THREADS, 1) conv_row_kernel(const __grid_constant__ ConvRowParams p) {
  using G = RowGeom<KS, PARTS>;
  constexpr int P = G::P, N = G::N, NSTAGE = G::NSTAGE, ND = G::ND, PLANE = G::PLANE;
  constexpr uint32_t FMT = PARTS == 2 ? 0u : 1u;  // fp16 hi|lo split, or one bf16 pass
  constexpr uint32_t IDESC = row_idesc(FMT, N);
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  RowGroup* gtab = reinterpret_cast<RowGroup*>(smem + 512);
  double* red = reinterpret_cast<double*>(smem + 1152);
  unsigned char* Bs = smem + CR_SMEM_HDR;
  unsigned char* As = Bs + (size_t)p.ngroups * G::B_GROUP;
  float* xf_a = reinterpret_cast<float*>(As + NSTAGE * G::STAGE_BYTES);
  float* xf_b = xf_a + p.cin_ch;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z;
  const int x0 = blockIdx.y * 128;
  const int y0 = blockIdx.x * p.rpc;
  const int H = p.H, W = p.W;
  const int nrows = min(p.rpc, H - y0);
  const int nin = nrows + KS - 1;
  const int NG = p.ngroups;
  const int total = nin * NG;
  const size_t plane_px = (size_t)H * W;
  const uint32_t bar0 = smem_u32(smem);
  auto a_full = [&](int s) { return bar0 + (uint32_t)s * 8u; };
  auto a_empty = [&](int s) { return bar0 + (uint32_t)(NSTAGE + s) * 8u; };
  auto d_full = [&](int d) { return bar0 + (uint32_t)(2 * NSTAGE + d) * 8u; };
  auto d_empty = [&](int d) { return bar0 + (uint32_t)(2 * NSTAGE + ND + d) * 8u; };

  // ---- one-time setup
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(a_full(s), 5);   // 4 producer warps + the halo warp
      mbar_init(a_empty(s), 1);  // tcgen05.commit
    }
    for (int d = 0; d < ND; ++d) {
      mbar_init(d_full(d), 1);   // tcgen05.commit
      mbar_init(d_empty(d), 4);  // 4 epilogue warps
    }
    fence_mbar_init();
    int g = 0, c0 = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const pbmc_src& S = p.src[s];
      for (int cb = 0; cb < S.nblk; cb += 4, ++g) {
        RowGroup gi;
        gi.base = S.ptr + ((size_t)b * S.nblk + cb) * plane_px * 4;
        gi.nb = min(4, S.nblk - cb);
        gi.xform = S.xform;
        gi.chan0 = c0 + cb * 4;
        gi.pad = 0;
        gtab[g] = gi;
      }
      c0 += S.nblk * 4;
    }
  }
  if (warp == CR_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), G::TMEM_COLS);
  {
    int c0 = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const pbmc_src& S = p.src[s];
      for (int c = tid; c < S.nblk * 4; c += CR_THREADS) {
        float a = 1.f, bb = 0.f;
        if (S.xform == PBMC_XFORM_GN_GELU || S.xform == PBMC_XFORM_GN)
          gn_coeffs(S.stats + ((size_t)b * S.nblk + (c >> 2)) * 2, S.inv_count, S.gamma[c], S.beta[c], a, bb);
        xf_a[c0 + c] = a;
        xf_b[c0 + c] = bb;
      }
      c0 += S.nblk * 4;
    }
    const uint4* wsrc = reinterpret_cast<const uint4*>(p.wpk);
    uint4* wdst = reinterpret_cast<uint4*>(Bs);
    for (int e = tid; e < NG * (G::B_GROUP / 16); e += CR_THREADS) wdst[e] = __ldg(wsrc + e);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================================================================ epilogue
    const int col = warp * 32 + lane, gx = x0 + col;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float bias[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < p.cout_blks) bq = ldg4(p.bias + q * 4);
      bias[q * 4 + 0] = bq.x; bias[q * 4 + 1] = bq.y; bias[q * 4 + 2] = bq.z; bias[q * 4 + 3] = bq.w;
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, cs[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) cs[c] = 0.f;
    for (int yo = 0; yo < nrows; ++yo) {
      const int rl = yo + KS - 1;  // the last input row this output row needs (commits are in order)
      mbar_wait(d_full(rl % ND), (uint32_t)(rl / ND) & 1u);
      tc_fence_after();
      uint32_t r[KS][16];
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) tmem_ld16_issue(lane_addr + (uint32_t)(((yo + dy) % ND) * N + dy * 16), r[dy]);
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) tmem_ld_wait16(r[dy]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty(yo % ND));  // D_yo has now been read by all of its k consumers
      const int gy = y0 + yo;
      if (gx < W) {
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) {
          if (qb < p.cout_blks) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = __uint_as_float(r[0][qb * 4 + e]);
#pragma unroll
              for (int dy = 1; dy < KS; ++dy) a += __uint_as_float(r[dy][qb * 4 + e]);
              a += bias[qb * 4 + e];
              if (p.epi_act == PBMC_ACT_GELU) a = gelu_erf(a);
              o[e] = a;
              cs[qb * 4 + e] += a;
            }
            *reinterpret_cast<float4*>(p.out + (((size_t)b * p.cout_blks + qb) * plane_px + (size_t)gy * W + gx) * 4) =
                make_float4(o[0], o[1], o[2], o[3]);
            s1[qb] += (o[0] + o[1]) + (o[2] + o[3]);
            s2[qb] += (o[0] * o[0] + o[1] * o[1]) + (o[2] * o[2] + o[3] * o[3]);
          }
        }
      }
    }
    if (p.out_stats != nullptr) {
#pragma unroll
      for (int qb = 0; qb < 4; ++qb) {
        const double a = warp_sum((double)s1[qb]);
        const double c2 = warp_sum((double)s2[qb]);
        if (lane == 0) { red[(warp * 4 + qb) * 2] = a; red[(warp * 4 + qb) * 2 + 1] = c2; }
      }
      epi_bar_sync();
      if (tid < 8 && (tid >> 1) < p.cout_blks) {
        double t = 0.0;
        for (int w = 0; w < 4; ++w) t += red[(w * 4 + (tid >> 1)) * 2 + (tid & 1)];
        atomicAdd(p.out_stats + ((size_t)b * p.cout_blks + (tid >> 1)) * 2 + (tid & 1), t);
      }
      epi_bar_sync();
    }
    if (p.out_chan_sum != nullptr) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const double a = warp_sum((double)cs[c]);
        if (lane == 0) red[warp * 16 + c] = a;
      }
      epi_bar_sync();
      if (tid < p.cout_blks * 4) {
        double t = 0.0;
        for (int w = 0; w < 4; ++w) t += red[w * 16 + tid];
        atomicAdd(p.out_chan_sum + (size_t)b * p.cout_blks * 4 + tid, t);
      }
    }
  } else if (warp < CR_MMA_WARP) {
    // ================================================================ producers (one thread per position)
    const int pg = (warp - 4) >> 2;
    const int i = ((warp - 4) & 3) * 32 + lane;
    const int gxp = x0 - P + i;
    const int sx = pad_index(gxp, W, p.pad_mode);
    const bool col_ok = gxp < W + P && sx >= 0;  // columns past the image feed masked outputs only
    float4 cur[4], nxt[4];
    bool ok_cur = false, ok_nxt = false;
    auto load_stage = [&](int st, float4 (&r)[4]) -> bool {
      const int ri = st / NG, g = st - ri * NG;
      const int sy = pad_index(y0 - P + ri, H, p.pad_mode);
      const bool ok = col_ok && sy >= 0;
      const RowGroup gi = gtab[g];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && j < gi.nb) r[j] = ldg4(gi.base + ((size_t)j * plane_px + (size_t)sy * W + sx) * 4);
      }
      return ok;
    };
    int st = pg;
    if (st < total) ok_cur = load_stage(st, cur);
    for (; st < total; st += 2) {
      if (st + 2 < total) ok_nxt = load_stage(st + 2, nxt);
      const int ri = st / NG, g = st - ri * NG;
      const RowGroup gi = gtab[g];
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 t = cur[j];
        if (ok_cur && j < gi.nb) t = xform4(t, xf_a + gi.chan0 + 4 * j, xf_b + gi.chan0 + 4 * j, gi.xform);
        v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
      const int slot = st % NSTAGE;
      uint4* stage = reinterpret_cast<uint4*>(As + (size_t)slot * G::STAGE_BYTES);
      mbar_wait(a_empty(slot), ((uint32_t)(st / NSTAGE) & 1u) ^ 1u);
      if (PARTS == 2) {
        uint4 h0, l0, h1, l1;
        split_f16(v, h0, l0);
        split_f16(v + 8, h1, l1);
        stage[i] = h0;
        stage[PLANE + i] = h1;
        stage[2 * PLANE + i] = l0;
        stage[3 * PLANE + i] = l1;
      } else {
        stage[i] = pack_bf16(v);
        stage[PLANE + i] = pack_bf16(v + 8);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(slot));
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
      ok_cur = ok_nxt;
    }
  } else if (warp == CR_MMA_WARP) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs);
      constexpr uint32_t A_LBO = PLANE * 16, B_LBO = N * 16, SBO = 128;
      int st = 0;
      for (int ri = 0; ri < nin; ++ri) {
        const int ds = ri % ND;
        mbar_wait(d_empty(ds), ((uint32_t)(ri / ND) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(ds * N);
        for (int g = 0; g < NG; ++g, ++st) {
          const int slot = st % NSTAGE;
          mbar_wait(a_full(slot), (uint32_t)(st / NSTAGE) & 1u);
          tc_fence_after();
          const uint32_t a_st = a_base + (uint32_t)slot * G::STAGE_BYTES;
          const uint32_t b_g = b_base + (uint32_t)g * G::B_GROUP;
#pragma unroll
          for (int dx = 0; dx < KS; ++dx) {
            const uint64_t a_hi = umma_desc(a_st + (uint32_t)dx * 16, A_LBO, SBO);
            const uint64_t b_hi = umma_desc(b_g + (uint32_t)(dx * PARTS) * G::B_TILE, B_LBO, SBO);
            umma_ss<1>(dcol, a_hi, b_hi, IDESC, (g == 0 && dx == 0) ? 0u : 1u);
            if (PARTS == 2) {
              const uint64_t a_lo = umma_desc(a_st + G::PART_BYTES + (uint32_t)dx * 16, A_LBO, SBO);
              const uint64_t b_lo = umma_desc(b_g + (uint32_t)(dx * PARTS + 1) * G::B_TILE, B_LBO, SBO);
              umma_ss<1>(dcol, a_lo, b_hi, IDESC, 1u);
              umma_ss<1>(dcol, a_hi, b_lo, IDESC, 1u);
            }
          }
          umma_commit(a_empty(slot));  // frees the smem stage once these MMAs have read it
        }
        umma_commit(d_full(ds));  // D_ri complete
      }
    }
    __syncwarp();
  } else {
    // ================================================================ halo producer: positions 128 .. 128+k-2
    constexpr int ITEMS = ((KS - 1) * 16 + 31) / 32;
    float cur[ITEMS], nxt[ITEMS];
    auto load_stage = [&](int st, float (&r)[ITEMS]) {
      const int ri = st / NG, g = st - ri * NG;
      const int sy = pad_index(y0 - P + ri, H, p.pad_mode);
      const RowGroup gi = gtab[g];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int item = lane + 32 * k, e = item >> 4, ch = item & 15;
        const int gxp = x0 - P + 128 + e;
        const int sx = pad_index(gxp, W, p.pad_mode);
        const bool ok = e < KS - 1 && gxp < W + P && sx >= 0 && sy >= 0 && (ch >> 2) < gi.nb;
        float val = 0.f;
        if (ok) {
          val = __ldg(gi.base + ((size_t)(ch >> 2) * plane_px + (size_t)sy * W + sx) * 4 + (ch & 3));
          val = xform1(val, xf_a[gi.chan0 + ch], xf_b[gi.chan0 + ch], gi.xform);
        }
        r[k] = val;
      }
    };
    if (total > 0) load_stage(0, cur);
    for (int st = 0; st < total; ++st) {
      if (st + 1 < total) load_stage(st + 1, nxt);
      const int slot = st % NSTAGE;
      unsigned char* stage = As + (size_t)slot * G::STAGE_BYTES;
      mbar_wait(a_empty(slot), ((uint32_t)(st / NSTAGE) & 1u) ^ 1u);
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int item = lane + 32 * k, e = item >> 4, ch = item & 15;
        if (e < KS - 1) {
          const size_t off = ((size_t)(ch >> 3) * PLANE + 128 + e) * 16 + (size_t)(ch & 7) * 2;
          if (PARTS == 2) {
            const __half h = __float2half_rn(cur[k]);
            const __half l = __float2half_rn(cur[k] - __half2float(h));
            *reinterpret_cast<__half*>(stage + off) = h;
            *reinterpret_cast<__half*>(stage + G::PART_BYTES + off) = l;
          } else {
            *reinterpret_cast<__nv_bfloat16*>(stage + off) = __float2bfloat16_rn(cur[k]);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(slot));
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) cur[k] = nxt[k];
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == CR_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, G::TMEM_COLS);
  }
}

// rows per CTA: minimise waves * (rows + halo + fixed per-CTA cost in row units)
static int choose_rpc(int units, int H, int ks) {
  static const int forced = getenv("PBMC_ROW_RPC") ? atoi(getenv("PBMC_ROW_RPC")) : 0;  // developer knob
  if (forced > 0) return forced < H ? forced : H;
  int best = 1;
  double best_cost = 1e30;
  for (int r = 1; r <= 64 && r <= H; ++r) {
    const long ctas = (long)units * cdiv(H, r);
    const long waves = (ctas + 147) / 148;
    const double cost = (double)waves * (r + ks - 1 + 6.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = r; }
  }
  return best;
}

template <int KS, int PARTS>
static int launch_row(ConvRowParams& p, cudaStream_t st) {
  using G = RowGeom<KS, PARTS>;
  const size_t smem = CR_SMEM_HDR + (size_t)p.ngroups * G::B_GROUP + (size_t)G::NSTAGE * G::STAGE_BYTES + (size_t)p.cin_ch * 8;
  if (smem > 227 * 1024) return PBMC_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    PBMC_CUDA(cudaFuncSetAttribute(conv_row_kernel<KS, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int nstrips = cdiv(p.W, 128);
  p.rpc = choose_rpc(nstrips * p.B, p.H, KS);
  dim3 grid(cdiv(p.H, p.rpc), nstrips, p.B);
  if (grid.y > 65535 || grid.z > 65535) return PBMC_ERR_BAD_SHAPE;
  conv_row_kernel<KS, PARTS><<<grid, CR_THREADS, smem, st>>>(p);
  PBMC_CHECK_LAUNCH("conv_row_kernel");
  return PBMC_OK;
}

static int row_groups(const pbmc_conv_desc& d) {
  int g = 0;
  for (int s = 0; s < d.nsrc; ++s) g += (d.src[s].nblk + 3) / 4;
  return g;
}

bool conv_row_supported(const pbmc_conv_desc& d) {
  if (d.ksize != 3 && d.ksize != 5) return false;
  if (d.cout > 16) return false;
  const int ng = row_groups(d);
  if (ng > CR_MAXG) return false;
  int cin = 0;
  for (int s = 0; s < d.nsrc; ++s) cin += d.src[s].nblk * 4;
  const size_t n = (size_t)d.ksize * 16, b_group = (size_t)d.ksize * 2 * (2 * n * 16);
  const size_t plane = ((128 + d.ksize - 1) + 7) / 8 * 8;
  const size_t smem = CR_SMEM_HDR + ng * b_group + 4 * (2 * 2 * plane * 16) + (size_t)cin * 8;
  return smem <= 227 * 1024;
}

// wpk_row holds two operand images back to back (ops.pack_conv_weight_row):
//   [fp16 hi|lo : groups * k * 2 * (2*N*16) B][bf16 : groups * k * (2*N*16) B]
int conv_row_dispatch(const pbmc_conv_desc& d, cudaStream_t st) {
  ConvRowParams p;
  int cin = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    p.src[s] = d.src[s];
    cin += d.src[s].nblk * 4;
  }
  p.nsrc = d.nsrc; p.B = d.B; p.H = d.H; p.W = d.W;
  p.cout_blks = (d.cout + 3) / 4;
  p.pad_mode = d.pad_mode; p.epi_act = d.epi_act;
  p.ngroups = row_groups(d);
  p.cin_ch = cin;
  p.rpc = 1;
  p.bias = d.bias; p.out = d.out; p.out_stats = d.out_stats; p.out_chan_sum = d.out_chan_sum;
  const char* base = reinterpret_cast<const char*>(d.wpk_row);
  if (!base) return PBMC_ERR_NULL_POINTER;
  if (!aligned16(base)) return PBMC_ERR_MISALIGNED;
  const size_t n = (size_t)d.ksize * 16;
  const size_t off_bf16 = (size_t)p.ngroups * d.ksize * 2 * (2 * n * 16);
  if (d.impl == PBMC_CONV_ROW_F16X2) {
    p.wpk = base;
    return d.ksize == 3 ? launch_row<3, 2>(p, st) : launch_row<5, 2>(p, st);
  }
  if (d.impl == PBMC_CONV_ROW_BF16) {
    p.wpk = base + off_bf16;
    return d.ksize == 3 ? launch_row<3, 1>(p, st) : launch_row<5, 1>(p, st);
  }
  return PBMC_ERR_UNSUPPORTED;
}

}  // namespace pbmc
