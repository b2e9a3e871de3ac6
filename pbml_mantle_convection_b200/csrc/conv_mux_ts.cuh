// conv_ts_kernel: the mux kernel (conv_mux.cu) with the A operand of every tcgen05.mma in TENSOR MEMORY.
//
// Why: with A in shared memory the N = 48 MMA is operand-fetch bound -- tools/probe/ts_probe measures 69.7 clk per
// dependent M128 N48 K16 MMA with A in shared memory against 37.0 clk with A in TMEM -- and while it runs it uses
// the whole 128 B/clk shared-memory pipe, which the staging warps' LDG/LDS/STS/STG share (in-kernel timelines,
// tools/muxtrace.py: 640-690 clk per 128-pixel row = 50 KB of operand re-reads + ~30 KB of L1/smem traffic, for
// a plain conv with no arithmetic to speak of).  The pixel tile [128 x 16 ch] is needed at the three horizontal
// shifts dx: in shared memory that was one staged row read at three start addresses; a TMEM operand has lane =
// row of A, so the worker that owns TMEM lane L writes positions L, L+1, L+2 (its own packed values and its two
// right neighbours', by warp shuffle; the two lanes at the end of a warp read them from a small exchange buffer)
// as three operand images, 8 columns each for hi and lo.  Nothing but the filters (9 KB) lives in shared memory.
//
// TMEM map (512 columns): D ring = ND x 48 columns at 0, A ring = NA x 48 columns at 256 (row ri uses A slot
// ri % NA = its group, image (dx, part) at + dx*16 + part*8).
// Barriers: a_full[NA] (4 warps) / a_empty[NA] (tcgen05.commit), d_full[ND] (commit) / d_empty[ND] (4 warps x KS rows).
#pragma once

namespace pbmc {

constexpr int CT_NA = CM_SETS;  // A ring: one slot per worker group
constexpr int CT_ND = 5;        // accumulator ring
constexpr uint32_t CT_A_BASE = 256;
constexpr int CT_XCHG = 2048;   // exchange buffer: [2 buffers][groups][5 sources][2 positions][64 B]
constexpr int CT_XCHG_BYTES = 2 * CM_SETS * 5 * 2 * 64;
constexpr int CT_BS = (CT_XCHG + CT_XCHG_BYTES + 127) / 128 * 128;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint4 a, uint4 b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <uint32_t KIND_F16>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint4 shfl_down4(uint4 v, int d) {
  return make_uint4(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d), __shfl_down_sync(0xffffffffu, v.z, d),
                    __shfl_down_sync(0xffffffffu, v.w, d));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 q;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
  return q;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 q) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
}

#ifdef PBMC_ROW_TRACE
// trace builds: a wait that gives up after ~1 ms and says who was waiting for what (the kernel then runs on with wrong data)
__device__ __forceinline__ void ts_wait_dbg(uint32_t bar, uint32_t parity, int code, int idx, unsigned long long* trace) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(2000u)
        : "memory");
    if (!done && ++spins > (1u << 12)) {
      // give the host something to read (the trace buffer may be pinned host memory), then fail loudly
      if (trace != nullptr && (threadIdx.x & 31) == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        trace[4000 + (threadIdx.x >> 5)] = 0x8000000000000000ull | ((unsigned long long)code << 32) | ((unsigned long long)idx << 8) | parity;
        __threadfence_system();
      }
      __nanosleep(1000000);
      __trap();
    }
  }
}
#define ts_wait(bar, parity, code, idx) ts_wait_dbg(bar, parity, code, idx, p.trace)
#else
__device__ __forceinline__ void ts_wait(uint32_t bar, uint32_t parity, int, int) { mbar_wait_parked(bar, parity); }
#endif

template <int PARTS>
__global__ void __maxnreg__(CM_MAXNREG) conv_ts_kernel(const __grid_constant__ ConvMuxParams p) {
  constexpr int KS = 3, P = 1, N = CM_N, ND = CT_ND, NA = CT_NA;
  constexpr int B_TILE = 2 * N * 16, B_GROUP = KS * PARTS * B_TILE;
  constexpr uint32_t FMT = PARTS == 2 ? 0u : 1u;  // fp16 hi|lo split, or one bf16 pass
  constexpr uint32_t IDESC = row_idesc(FMT, N);
  static_assert(8 * (2 * NA + 2 * ND) <= 448, "barrier area");
  static_assert(ND * N <= 256 && NA * 48 <= 256, "TMEM map");
  static_assert(NA == ND && ND == CM_SETS, "row ri: A slot = D slot = staging / reading group");
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 448);
  double* red = reinterpret_cast<double*>(smem + 512);    // 20 warps x 8 doubles
  float* bias_s = reinterpret_cast<float*>(smem + 1792);  // 16 floats (zero padded)
  float* xf_a = reinterpret_cast<float*>(smem + 1856);    // GroupNorm scale / shift of the 16 input channels
  float* xf_b = reinterpret_cast<float*>(smem + 1920);
  unsigned char* Bs = smem + CT_BS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z;
  const int x0 = blockIdx.y * 128;
  const int y0 = blockIdx.x * p.rpc;
  const int H = p.H, W = p.W;
  const int nrows = min(p.rpc, H - y0);
  const int nin = nrows + KS - 1;
  const int nb = min(p.nblk, 4);
  const size_t plane_px = (size_t)H * W;
  const uint32_t bar0 = smem_u32(smem);
  auto a_full = [&](uint32_t s) { return bar0 + s * 8u; };
  auto a_empty = [&](uint32_t s) { return bar0 + (uint32_t)(NA + s) * 8u; };
  auto d_full = [&](uint32_t d) { return bar0 + (uint32_t)(2 * NA + d) * 8u; };
  auto d_empty = [&](uint32_t d) { return bar0 + (uint32_t)(2 * NA + ND + d) * 8u; };

  // ---- worker geometry (needed before the set-up barrier: the first rows are requested right away)
  const int g = warp >> 2, wq = warp & 3;
  const int i = wq * 32 + lane;  // position in the row = input column x0 - 1 + i  (= TMEM lane of the dx = 0 image)
  const int gxp = x0 - P + i;
  const int sx = pad_index(gxp, W, p.pad_mode);
  const bool col_ok = gxp < W + P && sx >= 0;  // columns past the image feed masked outputs only
  const int hch = lane & 15, he = lane >> 4;   // positions 128, 129: warp 0 of a group, lane = (position, channel)
  const int gxh = x0 - P + 128 + he;
  const int hsx = pad_index(gxh, W, p.pad_mode);
  const bool h_on = warp < CM_WORKERS && wq == 0 && gxh < W + P && hsx >= 0 && (hch >> 2) < nb;
  const float* in_b = p.in + (size_t)b * p.nblk * plane_px * 4;
  const float* cbase = in_b + (size_t)(sx < 0 ? 0 : sx) * 4;
  const float* hbase = in_b + (size_t)(hch >> 2) * plane_px * 4 + (size_t)(hsx < 0 ? 0 : hsx) * 4 + (hch & 3);
  const size_t pstride = plane_px * 4, rstride = (size_t)W * 4;
  struct Row {
    float4 v0, v1, v2, v3;
    float h;
    bool ok;
  };
  auto load_row = [&](int ri, Row& R) {
    const int sy = pad_index(y0 - P + ri, H, p.pad_mode);
    R.ok = sy >= 0;
    const size_t ro = (size_t)(sy < 0 ? 0 : sy) * rstride;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    R.v0 = R.v1 = R.v2 = R.v3 = z;
    R.h = 0.f;
    if (col_ok && sy >= 0) {
      const float* c = cbase + ro;
      R.v0 = ldg4(c);
      if (nb > 1) R.v1 = ldg4(c + pstride);
      if (nb > 2) R.v2 = ldg4(c + 2 * pstride);
      if (nb > 3) R.v3 = ldg4(c + 3 * pstride);
    }
    if (h_on && sy >= 0) R.h = __ldg(hbase + ro);
  };

  // ---- one-time setup
  if (tid == 0) {
    CM_TR(0);
    for (int s = 0; s < NA; ++s) {
      mbar_init(a_full(s), 4);   // the 4 warps of the group that stages the row
      mbar_init(a_empty(s), 1);  // tcgen05.commit
    }
    for (int d = 0; d < ND; ++d) {
      mbar_init(d_full(d), 1);        // tcgen05.commit
      mbar_init(d_empty(d), 4 * KS);  // 4 warps x the KS output rows that read D_d
    }
    fence_mbar_init();
  }
  if (warp == CM_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 512);
  {
    if (tid < 16) bias_s[tid] = tid < p.cout_blks * 4 ? __ldg(p.bias + tid) : 0.f;
    const uint4* wsrc = reinterpret_cast<const uint4*>(p.wpk);
    uint4* wdst = reinterpret_cast<uint4*>(Bs);
    for (int e = tid; e < B_GROUP / 16; e += CM_THREADS) wdst[e] = __ldg(wsrc + e);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // no-op unless launched with programmatic serialization
  Row ra;
  if (warp < CM_WORKERS && g < nin) load_row(g, ra);  // the group's first row is requested before the set-up barrier
  const bool do_x = p.xform == PBMC_XFORM_GN_GELU;
  if (tid < 16) {
    float a = 1.f, bb = 0.f;
    if (do_x && tid < nb * 4) gn_coeffs(p.stats + ((size_t)b * p.nblk + (tid >> 2)) * 2, p.inv_count, p.gamma[tid], p.beta[tid], a, bb);
    xf_a[tid] = a;
    xf_b[tid] = bb;
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  fence_proxy_async_smem();  // the filters were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) CM_TR(1);

  if (warp < CM_WORKERS) {
    // ================================================================ workers: stage rows into TMEM, drain accumulators
    // Group g stages input rows g, g+5, ... (always into A slot g) and owns output rows g, g+5, ...  Order of a
    // group's work: rows are staged in order; an output row is taken early when its accumulators are already
    // complete, and MUST be taken before the staging step that (transitively) needs its accumulator slot back
    // (row ri needs the MMAs of row ri-NA retired, whose accumulator slot is released by output rows <= ri-NA-ND).
    const float h_a = xf_a[hch], h_b = xf_b[hch];
    const bool tr_lane = lane == 0 && wq == 0;
    (void)tr_lane;
    // Stagger: the five groups run the same cycle (arithmetic of GroupNorm+GELU, then shuffles / TMEM stores, then
    // loads, then TMEM loads / global stores of an output row); started together they stay in lock step and every
    // phase is bound by one pipe while the others idle.  Group g starts g * p.stagger clocks late.
    if (g > 0 && p.stagger > 0) {
      const long long t_go = clock64() + (long long)g * p.stagger;
      while (clock64() < t_go) __nanosleep(32);
    }
    const uint32_t xfa_addr = smem_u32(xf_a), xfb_addr = smem_u32(xf_b);
    const uint32_t xchg0 = smem_u32(smem + CT_XCHG);
    const uint32_t a_taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + CT_A_BASE + (uint32_t)g * 48u;
    int nstaged = 0;
    auto stage_row = [&](int ri, const Row& R) {
      if (tr_lane) CM_TR(100 + g * 64 + 4 * (ri / CM_SETS));
      float v[16] = {R.v0.x, R.v0.y, R.v0.z, R.v0.w, R.v1.x, R.v1.y, R.v1.z, R.v1.w,
                     R.v2.x, R.v2.y, R.v2.z, R.v2.w, R.v3.x, R.v3.y, R.v3.z, R.v3.w};
      float hv = R.h;
      if (do_x) {
        // GroupNorm + GELU, branch-free: out-of-image taps are masked back to zero afterwards
        const bool keep = R.ok && col_ok;
        const bool all_keep = __all_sync(0xffffffffu, keep) && nb == 4;  // interior warp: nothing to mask (warp-uniform)
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (j < nb) {
            float4 a0, b0, a1, b1;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w) : "r"(xfa_addr + j * 16));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(xfb_addr + j * 16));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w) : "r"(xfa_addr + j * 16 + 16));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(xfb_addr + j * 16 + 16));
            float x8[8];
            x8[0] = fmaf(v[4 * j + 0], a0.x, b0.x); x8[1] = fmaf(v[4 * j + 1], a0.y, b0.y);
            x8[2] = fmaf(v[4 * j + 2], a0.z, b0.z); x8[3] = fmaf(v[4 * j + 3], a0.w, b0.w);
            x8[4] = fmaf(v[4 * j + 4], a1.x, b1.x); x8[5] = fmaf(v[4 * j + 5], a1.y, b1.y);
            x8[6] = fmaf(v[4 * j + 6], a1.z, b1.z); x8[7] = fmaf(v[4 * j + 7], a1.w, b1.w);
            gelu_erf2n<4>(x8);
            if (all_keep) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[4 * j + e] = x8[e];
            } else {
              const bool keep1 = keep && j + 1 < nb;  // absent blocks of a partial group stay exactly zero
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[4 * j + e] = keep ? x8[e] : 0.f;
                v[4 * j + 4 + e] = keep1 ? x8[4 + e] : 0.f;
              }
            }
          }
        }
        if (h_on) {
          float hx[2] = {fmaf(hv, h_a, h_b), 0.f};
          gelu_erf2n<1>(hx);
          hv = R.ok ? hx[0] : 0.f;
        }
      }
      if (tr_lane) CM_TR(101 + g * 64 + 4 * (ri / CM_SETS));
      // pack: h0|h1 = the 16 channels as fp16 (TMEM columns 0-7 of the image), l0|l1 = the fp16 remainders
      uint4 h0, h1, l0, l1;
      if (PARTS == 2) {
        split_f16(v, h0, l0);
        split_f16(v + 8, h1, l1);
      } else {
        h0 = pack_bf16(v);
        h1 = pack_bf16(v + 8);
        l0 = l1 = make_uint4(0u, 0u, 0u, 0u);
      }
      // exchange buffer (double buffered by staging step): [source 0..3 = warp of the group, 4 = positions 128/129][2][64 B]
      const uint32_t xb = xchg0 + (uint32_t)(((nstaged & 1) * CM_SETS + g) * 5) * 128u;
      if (lane < 2) {
        const uint32_t e = xb + (uint32_t)(wq * 2 + lane) * 64u;
        sts128(e, h0);
        sts128(e + 16, h1);
        if (PARTS == 2) {
          sts128(e + 32, l0);
          sts128(e + 48, l1);
        }
      }
      if (wq == 0) {
        const uint32_t ha = xb + (uint32_t)(4 * 2 + he) * 64u + (uint32_t)hch * 2u;
        if (PARTS == 2) {
          const __half hh = __float2half_rn(hv);
          const __half hl = __float2half_rn(hv - __half2float(hh));
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(ha), "h"(__half_as_ushort(hh)) : "memory");
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(ha + 32u), "h"(__half_as_ushort(hl)) : "memory");
        } else {
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(ha), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(hv))) : "memory");
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
      // the A slot must have been read by the MMAs of the row staged into it NA rows ago
      if (ri >= NA) ts_wait(a_empty((uint32_t)g), (((uint32_t)ri / NA) & 1u) ^ 1u, 1, ri);
      tc_fence_after();
      // image dx = 0: this thread's own position
      tmem_st8(a_taddr + 0u, h0, h1);
      if (PARTS == 2) tmem_st8(a_taddr + 8u, l0, l1);
      // images dx = 1, 2: the right neighbours' values; the last two lanes of a warp take them from the exchange buffer
      const uint32_t nx = xb + (uint32_t)((wq + 1) * 2) * 64u;  // positions 32 (wq+1) + {0, 1}
#pragma unroll
      for (int dx = 1; dx <= 2; ++dx) {
        uint4 sh0 = shfl_down4(h0, dx), sh1 = shfl_down4(h1, dx);
        uint4 sl0 = l0, sl1 = l1;
        if (PARTS == 2) { sl0 = shfl_down4(l0, dx); sl1 = shfl_down4(l1, dx); }
        if (lane + dx >= 32) {
          const uint32_t e = nx + (uint32_t)(lane + dx - 32) * 64u;
          sh0 = lds128(e);
          sh1 = lds128(e + 16);
          if (PARTS == 2) { sl0 = lds128(e + 32); sl1 = lds128(e + 48); }
        }
        tmem_st8(a_taddr + (uint32_t)(dx * 16), sh0, sh1);
        if (PARTS == 2) tmem_st8(a_taddr + (uint32_t)(dx * 16 + 8), sl0, sl1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full((uint32_t)g));
      ++nstaged;
      if (tr_lane) CM_TR(102 + g * 64 + 4 * (ri / CM_SETS));
    };

    const int q = wq;
    const int col = q * 32 + lane, gx = x0 + col;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float cs[4] = {0.f, 0.f, 0.f, 0.f};  // zero-mean sums: only the c_out <= 4 head conv asks for them
    const bool want_cs = p.out_chan_sum != nullptr;
    const int cout_blks = p.cout_blks;
    const bool epi_gelu = p.epi_act == PBMC_ACT_GELU;
    const uint32_t bias_addr = smem_u32(bias_s);
    float* const obase = p.out + (((size_t)b * cout_blks) * plane_px + (size_t)y0 * W + gx) * 4;
    const size_t blk_stride = plane_px * 4;
    const bool col_in = gx < W;
    auto worker_loop = [&](auto lean_tag) {
      constexpr bool LEAN = decltype(lean_tag)::value;
      // Output row yo reads D_yo .. D_{yo+KS-1} and waits for all KS of them.  A parity wait is only sound when the
      // waiter sees EVERY phase of the barrier in order (one phase ahead, try_wait passes spuriously; two behind, it
      // never passes).  Accumulator slot s is read by the groups s-2, s-1, s (mod 5), every time it is used, in
      // order -- provided the rows above the strip (yo = -2, -1: they only exist as readers of D_0, D_1) are
      // processed too: groups 3 and 4 start with such a "virtual" row, which waits and releases but loads nothing.
      auto epi_ready = [&](int yo) {
        bool ok = true;
#pragma unroll
        for (int dy = 0; dy < KS; ++dy)
          if (yo + dy >= 0) ok = ok && mbar_test(d_full((uint32_t)(yo + dy) % ND), ((uint32_t)(yo + dy) / ND) & 1u);
        return ok;
      };
      auto epi_row = [&](int yo) {
#pragma unroll
        for (int dy = 0; dy < KS; ++dy)
          if (yo + dy >= 0) ts_wait(d_full((uint32_t)(yo + dy) % ND), ((uint32_t)(yo + dy) / ND) & 1u, 2, yo * 10 + dy);
        tc_fence_after();
        if (yo < 0) {
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int dy = 0; dy < KS; ++dy)
              if (yo + dy >= 0) mbar_arrive(d_empty((uint32_t)(yo + dy) % ND));
          }
          return;
        }
        if (tr_lane) CM_TR(1200 + 3 * yo);
        const uint32_t s_lo = (uint32_t)yo % ND;
        float* orow = obase + (size_t)yo * W * 4;
        // two halves of 8 output channels: 24 live accumulator registers instead of 48
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          uint32_t r[KS][8];
          uint32_t sl = s_lo;
#pragma unroll
          for (int dy = 0; dy < KS; ++dy) {
            tmem_ld8_issue(lane_addr + sl * (uint32_t)N + (uint32_t)(dy * 16 + hq * 8), r[dy]);
            if (++sl == (uint32_t)ND) sl = 0;
          }
#pragma unroll
          for (int dy = 0; dy < KS; ++dy) tmem_ld_wait8(r[dy]);
          if (hq == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              sl = s_lo;
#pragma unroll
              for (int dy = 0; dy < KS; ++dy) {
                mbar_arrive(d_empty(sl));
                if (++sl == (uint32_t)ND) sl = 0;
              }
            }
            if (tr_lane) CM_TR(1201 + 3 * yo);
          }
          if (col_in) {
#pragma unroll
            for (int qh = 0; qh < 2; ++qh) {
              const int qb = 2 * hq + qh;
              if (LEAN || qb < cout_blks) {
                float4 bq;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bq.x), "=f"(bq.y), "=f"(bq.z), "=f"(bq.w) : "r"(bias_addr + qb * 16));
                const float bias4[4] = {bq.x, bq.y, bq.z, bq.w};
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a = __uint_as_float(r[0][qh * 4 + e]);
#pragma unroll
                  for (int dy = 1; dy < KS; ++dy) a += __uint_as_float(r[dy][qh * 4 + e]);
                  a += bias4[e];
                  if (!LEAN && epi_gelu) a = gelu_erf(a);
                  o[e] = a;
                }
                *reinterpret_cast<float4*>(orow + qb * blk_stride) = make_float4(o[0], o[1], o[2], o[3]);
                s1[qb] += (o[0] + o[1]) + (o[2] + o[3]);
                s2[qb] = fmaf(o[0], o[0], fmaf(o[1], o[1], fmaf(o[2], o[2], fmaf(o[3], o[3], s2[qb]))));
                if (!LEAN && qb == 0 && want_cs) { cs[0] += o[0]; cs[1] += o[1]; cs[2] += o[2]; cs[3] += o[3]; }
              }
            }
          }
        }
        if (tr_lane) CM_TR(1202 + 3 * yo);
      };
      int ri = g, yo = g + (KS - 1) >= CM_SETS ? g - CM_SETS : g;  // groups 3, 4: the virtual rows -2, -1 first
      while (ri < nin) {
        // output rows whose accumulator slots the MMAs ahead of row ri are waiting for
        while (yo < nrows && yo + NA + ND <= ri) {
          epi_row(yo);
          yo += CM_SETS;
        }
        stage_row(ri, ra);
        ri += CM_SETS;
        if (ri < nin) load_row(ri, ra);
        if (yo < nrows && epi_ready(yo)) {
          epi_row(yo);
          yo += CM_SETS;
        }
      }
      for (; yo < nrows; yo += CM_SETS) epi_row(yo);
    };
    if (cout_blks == 4 && !epi_gelu && !want_cs)
      worker_loop(std::true_type{});
    else
      worker_loop(std::false_type{});
    if (tid == 0) CM_TR(3);
    if (p.out_stats != nullptr) {
#pragma unroll
      for (int qb = 0; qb < 4; ++qb) {
        const double a = warp_sum((double)s1[qb]);
        const double c2 = warp_sum((double)s2[qb]);
        if (lane == 0) { red[(warp * 4 + qb) * 2] = a; red[(warp * 4 + qb) * 2 + 1] = c2; }
      }
      cm_worker_bar();
      if (tid < 8 && (tid >> 1) < p.cout_blks) {
        double t = 0.0;
        for (int w = 0; w < CM_WORKERS; ++w) t += red[(w * 4 + (tid >> 1)) * 2 + (tid & 1)];
        atomicAdd(p.out_stats + ((size_t)b * p.cout_blks + (tid >> 1)) * 2 + (tid & 1), t);
      }
      cm_worker_bar();
    }
    if (want_cs) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double a = warp_sum((double)cs[c]);
        if (lane == 0) red[warp * 4 + c] = a;
      }
      cm_worker_bar();
      if (tid < 4) {
        double t = 0.0;
        for (int w = 0; w < CM_WORKERS; ++w) t += red[w * 4 + tid];
        atomicAdd(p.out_chan_sum + (size_t)b * 4 + tid, t);
      }
    }
  } else {
    // ================================================================ MMA issuers: row ri by issuer ri % CM_NMMA
    const bool leader = elect_one();
    constexpr uint32_t B_LBO = N * 16, SBO = 128;
    const uint64_t b_desc0 = umma_desc(smem_u32(Bs), B_LBO, SBO);
    // Issuer = slot % CM_NMMA: every a_full / d_empty barrier then has ONE waiter, which sees all its phases in order.
    for (int ri = 0; ri < nin; ++ri) {
      const uint32_t ds = (uint32_t)ri % ND, as = (uint32_t)ri % NA;
      if ((int)(as % CM_NMMA) != warp - CM_MMA_WARP) continue;
      if (ri >= ND) ts_wait(d_empty(ds), (((uint32_t)ri / ND) & 1u) ^ 1u, 3, ri);  // first ND rows: the ring is free
      ts_wait(a_full(as), ((uint32_t)ri / NA) & 1u, 4, ri);
      tc_fence_after();
      if (leader) {
        CM_TR(1400 + 2 * ri);
        const uint32_t dcol = tmem_base + ds * (uint32_t)N;
        const uint32_t acol = tmem_base + CT_A_BASE + as * 48u;
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const uint32_t a_hi = acol + (uint32_t)(dx * 16);
          const uint64_t b_hi = b_desc0 + (uint64_t)(dx * PARTS * (B_TILE >> 4));
          umma_ts<1>(dcol, a_hi, b_hi, IDESC, (uint32_t)dx);
          if (PARTS == 2) {
            umma_ts<1>(dcol, a_hi + 8u, b_hi, IDESC, 1u);
            umma_ts<1>(dcol, a_hi, b_hi + (uint64_t)(B_TILE >> 4), IDESC, 1u);
          }
        }
        umma_commit(a_empty(as));  // the A slot may be overwritten once these MMAs have read it
        umma_commit(d_full(ds));   // D_ri complete
        CM_TR(1401 + 2 * ri);
      }
      __syncwarp();
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (tid == 0) CM_TR(2);
  if (warp == CM_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pbmc
