// Fused residual pair of the vocoder's C = 64 stacks (the last up-sampling stage: 443 k time steps per clip) as ONE kernel:
//     x_new = x + conv_b(lrelu(conv_a(lrelu(x)) + bias_a)) + bias_b          (oracle: vocoder_generator, the `res.s.i` pair)
// conv_a: k = 3, dilation d, zero padding;  conv_b: k = 3, dilation 1, zero padding.  hi-only fp16 operands, fp32
// accumulation (the vocoder's 1-term mode).  What two separate GEMM launches move through HBM in between - the activated
// intermediate h, written by "a" and read back by "b" (4 of 16 bytes per element) - stays in shared memory.
//
//   tile = 126 output rows t0 .. t0+125 of one clip, m0 = t0 - 1.
//   P1 (conv_a)   A = lrelu(x) rows m0 + (tap-1) d + [0,128) by TMA (out-of-range rows zero filled), accumulator 1 = h rows
//                 m0 .. m0+127 before bias / activation
//   E1            + bias_a, LeakyReLU, rows outside the clip forced to zero (conv_b's zero padding), fp16, written to the
//                 shared-memory tile H in the K-major SWIZZLE_128B layout a TMA load would have produced (h row k sits at
//                 buffer row k + 1: the three taps of conv_b are row-shifted views, start rows 0/1/2 - legal with
//                 base_offset = 0, tools/probe_desc_shift.cu)
//   P2 (conv_b)   A = H views, accumulator 2 = output rows m0 .. m0+127, of which 1..126 are valid
//   E2            + bias_b + x (hi + lo planes), raw hi/lo planes of x_new and the activated hi plane for the next pair
//
// Round 1's version ran E1 and E2 on the same warps, one tile at a time: every tile paid the whole TMA -> P1 -> E1 -> P2 ->
// E2 latency chain and the kernel measured 30 % SLOWER than the two launches it replaces (3.2 vs 2.5 ms per pair).  Here each
// step has its own warps and every inter-step buffer is double buffered, so the steps of consecutive tiles overlap and the
// throughput is set by the slowest role (E2, the HBM traffic) instead of the sum:
//   warp 0       TMA producer: one 3-tap A stage (48 KB) per tile, two stages; both weight matrices (2 x 3 x 8 KB) are
//                loaded once and stay resident
//   warp 1       MMA issuer, software pipelined P1(j+1) before P2(j); 12 MMAs per barrier round trip, descriptors
//                precomputed per stage (the issue thread's per-chunk instruction count bounds narrow tiles)
//   warps 2-5    E1: accumulator 1 -> H (two H buffers)
//   warps 6-13   E2: residual (fp32 stream of the stack; hi/lo planes for its first pair) prefetched one tile ahead,
//                accumulator 2 -> fp32 x_new + the activated fp16 plane of the next pair
// Every wait is bounded (ptx.cuh).
#include "gemm.cuh"
#include "ptx.cuh"

namespace vf {

namespace {
constexpr int PAIR_C = 64;
constexpr int PAIR_ROWS = 126;                 // valid output rows per tile
constexpr int PAIR_A_TAP = 128 * 128;          // one A box: 128 rows x 64 channels fp16
constexpr int PAIR_A_STAGE = 3 * PAIR_A_TAP;   // the three dilated taps of one tile
constexpr int PAIR_W_TAP = PAIR_C * 128;       // one weight tile: 64 output channels x 64 input channels fp16
constexpr int PAIR_H_BUF = 136 * 128;          // 130 rows used, rounded up to whole 1024-byte swizzle atoms
constexpr int PAIR_E1_WARPS = 4, PAIR_E2_WARPS = 8;
constexpr int PAIR_E1_THREADS = 32 * PAIR_E1_WARPS, PAIR_E2_THREADS = 32 * PAIR_E2_WARPS;
constexpr int PAIR_THREADS = 64 + PAIR_E1_THREADS + PAIR_E2_THREADS;
constexpr int PAIR_SMEM = 2 * PAIR_A_STAGE + 6 * PAIR_W_TAP + 2 * PAIR_H_BUF + PAIR_E2_WARPS * 4096 + 256 + 2 * PAIR_C * 4 + 1024;
}  // namespace

template <bool F32_IN>
__global__ void __launch_bounds__(PAIR_THREADS, 1) pair_tc_kernel(const __grid_constant__ PairParams P) {
  constexpr int C = PAIR_C;
  constexpr uint32_t IDESC = make_idesc_f16(GEMM_BM, C);
  constexpr uint32_t DHI = make_smem_desc_hi(128);

  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem;                                   // [2][3][128 x 128 B]
  uint8_t* w_base = a_base + 2 * PAIR_A_STAGE;              // [Wa tap 0..2][Wb tap 0..2], 8 KB each
  uint8_t* h_base = w_base + 6 * PAIR_W_TAP;                // [2][136 x 128 B]
  uint8_t* stg_base = h_base + 2 * PAIR_H_BUF;              // E2: 8 x 4 KB staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + PAIR_E2_WARPS * 4096);
  uint64_t* w_full = bars;            // [1]
  uint64_t* a_full = bars + 1;        // [2]
  uint64_t* a_empty = bars + 3;       // [2]
  uint64_t* acc1_full = bars + 5;     // [2]
  uint64_t* acc1_empty = bars + 7;    // [2]
  uint64_t* h_ready = bars + 9;       // [2]
  uint64_t* h_free = bars + 11;       // [2]
  uint64_t* acc2_full = bars + 13;    // [2]
  uint64_t* acc2_empty = bars + 15;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 17);
  float* s_bias_a = reinterpret_cast<float*>(bars + 32);   // [C] (16-byte aligned)
  float* s_bias_b = s_bias_a + C;                          // [C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = P.n_img * P.tiles_per_img;
  const int n_local = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  constexpr int TMEM_COLS = 4 * C;               // accumulator 1 x 2 at columns 0 / C, accumulator 2 x 2 at 2C / 3C

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1);
      mbar_init(acc1_full + i, 1); mbar_init(acc1_empty + i, PAIR_E1_THREADS);
      mbar_init(h_ready + i, PAIR_E1_THREADS); mbar_init(h_free + i, 1);
      mbar_init(acc2_full + i, 1); mbar_init(acc2_empty + i, PAIR_E2_THREADS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&P.a_map);
    tma_prefetch_desc(&P.wa_map);
    tma_prefetch_desc(&P.wb_map);
  }
  if (warp == 1) tmem_alloc_dyn(tmem_holder, TMEM_COLS);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_bias_a[i] = __ldg(P.bias_a + i);
    s_bias_b[i] = __ldg(P.bias_b + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(w_full, 6 * PAIR_W_TAP);
      for (int t = 0; t < 3; ++t) {
        tma_load_2d(w_base + t * PAIR_W_TAP, &P.wa_map, w_full, t * C, 0);
        tma_load_2d(w_base + (3 + t) * PAIR_W_TAP, &P.wb_map, w_full, t * C, 0);
      }
      bool ok = true;
      for (int j = 0; j < n_local && ok; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int img = (int)fast_div_pair((uint32_t)tile, (uint32_t)P.tiles_per_img, P.magic_t);
        const int m0 = (tile - img * P.tiles_per_img) * PAIR_ROWS - 1;
        const int b = j & 1;
        const uint32_t pj = (j >> 1) & 1;
        if (!mbar_wait(a_empty + b, pj ^ 1u, P.err, ERR_PIPE_PRODUCER)) { ok = false; break; }
        uint8_t* st = a_base + b * PAIR_A_STAGE;
        mbar_expect_tx(a_full + b, PAIR_A_STAGE);
        tma_load_3d(st, &P.a_map, a_full + b, 0, m0 - P.dil, img);
        tma_load_3d(st + PAIR_A_TAP, &P.a_map, a_full + b, 0, m0, img);
        tma_load_3d(st + 2 * PAIR_A_TAP, &P.a_map, a_full + b, 0, m0 + P.dil, img);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t da0 = make_smem_desc_lo(smem_u32(a_base));
      const uint32_t dw0 = make_smem_desc_lo(smem_u32(w_base));
      const uint32_t dh0 = make_smem_desc_lo(smem_u32(h_base));
      bool ok = mbar_wait(w_full, 0, P.err, ERR_PIPE_MMA);
      for (int j = 0; j <= n_local && ok; ++j) {
        if (j < n_local) {          // ---- P1(j): conv_a into accumulator 1[b]
          const int b = j & 1;
          const uint32_t pj = (j >> 1) & 1;
          if (!mbar_wait(acc1_empty + b, pj ^ 1u, P.err, ERR_PIPE_MMA)) { ok = false; break; }
          if (!mbar_wait(a_full + b, pj, P.err, ERR_PIPE_MMA)) { ok = false; break; }
          tc_fence_after();
          const uint32_t da = da0 + (uint32_t)(b * (PAIR_A_STAGE >> 4));
          const uint32_t d = tmem_base + b * C;
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_lo(d, da + t * (PAIR_A_TAP >> 4) + 2 * k, dw0 + t * (PAIR_W_TAP >> 4) + 2 * k, DHI, IDESC, (t | k) ? 1u : 0u);
          umma_commit(a_empty + b);
          umma_commit(acc1_full + b);
        }
        if (j >= 1) {               // ---- P2(j-1): conv_b on H[b] into accumulator 2[b]
          const int i = j - 1, b = i & 1;
          const uint32_t pi = (i >> 1) & 1;
          if (!mbar_wait(h_ready + b, pi, P.err, ERR_PIPE_MMA)) { ok = false; break; }
          if (!mbar_wait(acc2_empty + b, pi ^ 1u, P.err, ERR_PIPE_MMA)) { ok = false; break; }
          tc_fence_after();
          const uint32_t dh = dh0 + (uint32_t)(b * (PAIR_H_BUF >> 4));
          const uint32_t d = tmem_base + 2 * C + b * C;
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k)      // view of tap t starts at buffer row t
              umma_f16_lo(d, dh + t * (128 >> 4) + 2 * k, dw0 + (3 + t) * (PAIR_W_TAP >> 4) + 2 * k, DHI, IDESC, (t | k) ? 1u : 0u);
          umma_commit(acc2_full + b);
          umma_commit(h_free + b);
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + PAIR_E1_WARPS) {
    // ------------------------------------------------------------------ E1: accumulator 1 -> H
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const int jrow = q * 32 + lane;            // this thread's accumulator row = h row
    const int brow = jrow + 1;                 // its row in the H buffer
    const float slope_h = P.slope_h;
    float amax = 0.f;
    bool ok = true;
    // buffer rows 0 and 129 of both H buffers are read by the first / last tap of accumulator rows 0 and 127 (never stored,
    // but they must stay finite for the overflow guard): zero them once, E1 only ever writes rows 1..128
    if (warp == 2 && lane < 16) {
      uint8_t* hb = h_base + (size_t)(lane >> 3) * PAIR_H_BUF;
      *reinterpret_cast<uint4*>(hb + (lane & 7) * 16) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(hb + 129 * 128 + (lane & 7) * 16) = make_uint4(0, 0, 0, 0);
    }
    for (int j = 0; j < n_local && ok; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int img = (int)fast_div_pair((uint32_t)tile, (uint32_t)P.tiles_per_img, P.magic_t);
      const int m0 = (tile - img * P.tiles_per_img) * PAIR_ROWS - 1;
      const int b = j & 1;
      const uint32_t pj = (j >> 1) & 1;
      const int t = m0 + jrow;
      const bool in_clip = t >= 0 && t < P.L;
      if (!mbar_wait(acc1_full + b, pj, P.err, ERR_PIPE_EPILOGUE)) { ok = false; break; }
      tc_fence_after();
      if (!mbar_wait(h_free + b, pj ^ 1u, P.err, ERR_PIPE_EPILOGUE)) { ok = false; break; }   // P2(j-2) has read H[b]
      uint8_t* rowp = h_base + (size_t)b * PAIR_H_BUF + (size_t)brow * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {              // two passes of 32 columns
        float v[32];
        tmem_ld_32x32(tmem_base + lane_bits + b * C + c * 32, v);
        if (c == 1) {
          tc_fence_before();
          mbar_arrive(acc1_empty + b);           // accumulator 1[b] is in registers: P1(j+2) may start
        }
        const float4* bp = reinterpret_cast<const float4*>(s_bias_a + c * 32);
        uint32_t hw[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = bp[i];
          float a0 = v[4 * i] + b4.x, a1 = v[4 * i + 1] + b4.y, a2 = v[4 * i + 2] + b4.z, a3 = v[4 * i + 3] + b4.w;
          a0 = fmaxf(a0, a0 * slope_h); a1 = fmaxf(a1, a1 * slope_h); a2 = fmaxf(a2, a2 * slope_h); a3 = fmaxf(a3, a3 * slope_h);
          if (!in_clip) { a0 = a1 = a2 = a3 = 0.f; }     // conv_b pads h with zeros outside the clip
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fmaxf(fabsf(a2), fabsf(a3))));
          const __half2 p0 = __floats2half2_rn(a0, a1), p1 = __floats2half2_rn(a2, a3);
          hw[2 * i] = *reinterpret_cast<const uint32_t*>(&p0);
          hw[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&p1);
        }
        // K-major SWIZZLE_128B image: 16-byte chunk index XOR (buffer row & 7)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(rowp + (((c * 4 + i) ^ (brow & 7)) << 4)) = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
      }
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async proxy
      mbar_arrive(h_ready + b);
    }
    if (!(amax <= 65504.f) && P.err) atomicCAS(P.err, 0, ERR_FP16_OVERFLOW);
  } else {
    // ------------------------------------------------------------------ E2: accumulator 2 + bias_b + x -> x_new, activated plane
    // The residual stream x of a fused stack is fp32 (F32_IN / out_f32: same 4 bytes per element as the hi/lo planes the
    // un-fused layers exchange, but no hi/lo split and no half -> float conversions in this role: ~640 -> ~450 instructions per
    // tile and warp, 2.62 -> 2.33 ms per pair).  The first pair of a stack still reads planes.  Measured alternatives: 16 warps
    // of 16 columns (four per scheduler, half the dependent chain each) are SLOWER, 2.72 ms - the stores of a warp then cover
    // 64-byte pieces of 256-byte rows, and ncu shows the kernel waiting on its global stores at 58 % DRAM utilisation.
    const int ew = warp - (2 + PAIR_E1_WARPS);
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // which 32-column chunk this warp takes (two warps share a lane quarter)
    float4* stg_f = reinterpret_cast<float4*>(stg_base) + (size_t)ew * 256;   // 4 KB per warp
    uint4* stg_h = reinterpret_cast<uint4*>(stg_f);                           // fp16 tiles alias it: two 32 x 64-byte tiles
    uint4* stg_l = stg_h + 128;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const int h_row = lane >> 2, h_c16 = lane & 3;          // fp16 row-major role: 8 rows x 64 B per instruction
    const int f_row = lane >> 3, f_c16 = lane & 7;          // fp32 row-major role: 4 rows x 128 B per instruction
    const int so_h0 = lane * 4, so_hx = (lane >> 1) & 3;
    const int sr_h0 = h_row * 4, sr_hx = h_c16;
    const int so_f0 = lane * 8, so_fx = lane & 7;
    const int sr_f0 = f_row * 8;
#define PSO_H(i) (so_h0 + ((i) ^ so_hx))
#define PSR_H(i) (32 * (i) + sr_h0 + (sr_hx ^ ((h_row >> 1) & 3)))
#define PSO_F(i) (so_f0 + ((i) ^ so_fx))
#define PSR_F(i) (32 * (i) + sr_f0 + (f_c16 ^ ((4 * (i) + f_row) & 7)))
    const float slope_out = P.slope_out;
    const bool want_f = P.out_f32 != nullptr;
    const int wrow0 = q * 32;                  // first accumulator row of this warp
    float amax = 0.f;
    bool ok = true;
    // Residual of a tile in the row-major role, fetched ONE TILE AHEAD (it is consumed right after the accumulator read).
    uint4 xr[8];                               // F32_IN: 8 x float4 (rows 4i + f_row); planes: [0,4) hi, [4,8) lo (rows 8i + h_row)
    long base_next = 0;                        // element offset of this warp's first row and column chunk (next tile)
    int m0_next = 0;
    auto row_ok = [&](int m0, int jr) { return jr >= 1 && jr <= PAIR_ROWS && m0 + jr < P.L; };
    auto load_resid = [&](int jj) {
      const int tile = blockIdx.x + jj * gridDim.x;
      const int img = (int)fast_div_pair((uint32_t)tile, (uint32_t)P.tiles_per_img, P.magic_t);
      m0_next = (tile - img * P.tiles_per_img) * PAIR_ROWS - 1;
      base_next = ((long)img * P.L + (m0_next + wrow0)) * C + half * 32;     // row -1 of a clip's first tile is never dereferenced
      if (F32_IN) {
        const float4* g = reinterpret_cast<const float4*>(P.resid_f32 + base_next + (long)f_row * C) + f_c16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xr[i] = make_uint4(0, 0, 0, 0);
          if (row_ok(m0_next, wrow0 + 4 * i + f_row)) xr[i] = __ldg(reinterpret_cast<const uint4*>(g + i * C));   // 4 rows further = 4 C floats = C float4
        }
      } else {
        const uint4* gh = reinterpret_cast<const uint4*>(P.resid_hi + base_next + (long)h_row * C) + h_c16;
        const uint4* gl = reinterpret_cast<const uint4*>(P.resid_lo + base_next + (long)h_row * C) + h_c16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xr[i] = make_uint4(0, 0, 0, 0); xr[4 + i] = make_uint4(0, 0, 0, 0);
          if (row_ok(m0_next, wrow0 + 8 * i + h_row)) {
            xr[i] = __ldg(gh + i * C);                                      // 8 rows further = 8 C halves = C uint4
            xr[4 + i] = __ldg(gl + i * C);
          }
        }
      }
    };
    if (n_local > 0) load_resid(0);
    for (int j = 0; j < n_local && ok; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int img = (int)fast_div_pair((uint32_t)tile, (uint32_t)P.tiles_per_img, P.magic_t);
      const int m0 = m0_next;
      const int b = j & 1;
      const uint32_t pj = (j >> 1) & 1;
      const long in_base = base_next;
      const long out_base = ((long)img * P.out_img_rows + P.out_row0 + (m0 + wrow0)) * C + half * 32;
      if (!mbar_wait(acc2_full + b, pj, P.err, ERR_PIPE_EPILOGUE)) { ok = false; break; }
      tc_fence_after();
      float v[32];
      tmem_ld_32x32(tmem_base + lane_bits + 2 * C + b * C + half * 32, v);
      tc_fence_before();
      mbar_arrive(acc2_empty + b);
      const float4* bp = reinterpret_cast<const float4*>(s_bias_b + half * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = bp[i];
        v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
      }
      // coalesced row-major residual -> staging -> own row
      __syncwarp();
      if (F32_IN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) stg_f[PSR_F(i)] = *reinterpret_cast<const float4*>(&xr[i]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 x = stg_f[PSO_F(i)];
          v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          stg_h[PSR_H(i)] = xr[i];
          stg_l[PSR_H(i)] = xr[4 + i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 yh = stg_h[PSO_H(i)], yl = stg_l[PSO_H(i)];
          const __half2* ph2 = reinterpret_cast<const __half2*>(&yh);
          const __half2* pl2 = reinterpret_cast<const __half2*>(&yl);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fh = __half22float2(ph2[k]), fl = __half22float2(pl2[k]);
            v[8 * i + 2 * k] += fh.x + fl.x;
            v[8 * i + 2 * k + 1] += fh.y + fl.y;
          }
        }
      }
      if (j + 1 < n_local) load_resid(j + 1);   // in flight while this tile is packed and stored
      if (want_f) {                             // x_new, fp32 stream of the stack
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) stg_f[PSO_F(i)] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        __syncwarp();
        float4* g = reinterpret_cast<float4*>(P.out_f32 + in_base + (long)f_row * C) + f_c16;    // the clip's own row pitch
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (row_ok(m0, wrow0 + 4 * i + f_row)) g[i * C] = stg_f[PSR_F(i)];
      }
      {                                         // activated plane for the next pair / stage
        uint32_t hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = fmaxf(v[2 * i], v[2 * i] * slope_out), a1 = fmaxf(v[2 * i + 1], v[2 * i + 1] * slope_out);
          amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));     // every row is finite (H border rows are zeroed)
          const __half2 hh = __floats2half2_rn(a0, a1);
          hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) stg_h[PSO_H(i)] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
        __syncwarp();
        uint4* g = reinterpret_cast<uint4*>(P.out_a + out_base + (long)h_row * C) + h_c16;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (row_ok(m0, wrow0 + 8 * i + h_row)) g[i * C] = stg_h[PSR_H(i)];
      }
    }
    if (!(amax <= 65504.f) && P.err) atomicCAS(P.err, 0, ERR_FP16_OVERFLOW);
#undef PSO_H
#undef PSR_H
#undef PSO_F
#undef PSR_F
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, TMEM_COLS);
  }
}

size_t pair_tc_smem_bytes(int C, int /*stages*/) { return C == PAIR_C ? (size_t)PAIR_SMEM : 0; }

template <bool F32_IN>
static cudaError_t launch_pair_t(const PairParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pair_tc_kernel<F32_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  pair_tc_kernel<F32_IN><<<p.grid, PAIR_THREADS, PAIR_SMEM, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pair_tc(const PairParams& p, cudaStream_t stream) {
  if (p.C != PAIR_C) return cudaErrorInvalidValue;
  return p.resid_f32 ? launch_pair_t<true>(p, stream) : launch_pair_t<false>(p, stream);
}

}  // namespace vf
