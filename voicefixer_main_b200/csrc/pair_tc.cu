// EXPERIMENTAL (opt-in, VF_TUNE_FUSED_PAIR=1; not yet validated on hardware - see DESIGN.md §9.1).
//
// One residual pair of the vocoder stacks with C = 64 or 128 channels as ONE kernel:
//     x_new = x + conv_b(lrelu(conv_a(lrelu(x)) + bias_a)) + bias_b          (oracle: vocoder_generator, the `res.s.i` pair)
// conv_a: k = 3, dilation d, zero padding;  conv_b: k = 3, dilation 1, zero padding.  hi-only fp16 operands, fp32
// accumulation (the vocoder's 1-term mode).  What the two separate GEMM launches move through HBM in between - the
// activated intermediate h, written by "a" and read back by "b" - stays in shared memory:
//
//   tile = 126 output rows t0 .. t0+125 of one clip.  m0 = t0 - 1.
//   phase 1 (conv_a)  A = lrelu(x) rows m0 + (tap-1) d + [0,128) by TMA (out-of-range rows are zero filled), W_a tiles by
//                     TMA, accumulator 1 = h rows m0 .. m0+127 (before bias / activation)
//   epilogue 1        + bias_a, LeakyReLU, rows outside the clip forced to zero (conv_b's zero padding), fp16, written to
//                     the shared-memory tile H in the K-major SWIZZLE_128B layout a TMA load would have produced
//                     (h row k sits at buffer row k + 1: the three taps of conv_b are row-shifted views, start rows 0/1/2 -
//                     legal with base_offset = 0, tools/probe_desc_shift.cu)
//   phase 2 (conv_b)  A = H views, W_b tiles by TMA, accumulator 2 = output rows m0 .. m0+127, of which 1..126 are valid
//   epilogue 2        + bias_b + x (hi + lo planes), raw hi/lo planes of x_new and the activated hi plane for the next pair,
//                     row-major stores through the same swizzled staging tiles as gemm_tc.cu
//
// Roles and barriers follow gemm_tc.cu (warp 0 TMA producer, warp 1 MMA issuer, 8 epilogue warps; every wait bounded).
// Phase 1 of tile i+1 overlaps epilogue 2 of tile i (separate accumulators); H is single-buffered and guarded by
// h_ready (epilogue -> MMA) / h_free (MMA -> epilogue).
#include "gemm.cuh"
#include "ptx.cuh"

namespace vf {

namespace {
constexpr int PAIR_ROWS = 126;                 // valid output rows per tile
constexpr int PAIR_A_SLOT = 128 * 128;         // one A box: 128 rows x 64 channels fp16
constexpr int PAIR_H_ATOM = 136 * 128;         // 130 rows used, rounded up to whole 1024-byte swizzle atoms
constexpr int PAIR_EPI_WARPS = 8;
constexpr int PAIR_EPI_THREADS = 32 * PAIR_EPI_WARPS;
}  // namespace

template <int C>
__global__ void __launch_bounds__(64 + PAIR_EPI_THREADS, C == 64 ? 2 : 1) pair_tc_kernel(const __grid_constant__ PairParams P) {
  constexpr int KA = C / 64;                   // 64-channel K atoms per tap
  constexpr int B_TILE = C * 128;              // one weight tile: C rows x 64 channels fp16
  constexpr int STAGE = PAIR_A_SLOT + B_TILE;
  constexpr int NCHUNK = C / 32;               // 32-column epilogue chunks
  constexpr uint32_t IDESC = make_idesc_f16(GEMM_BM, C);
  constexpr uint32_t DHI = make_smem_desc_hi(128);

  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stages = P.stages;
  uint8_t* h_base = smem + (size_t)stages * STAGE;                       // KA atoms of PAIR_H_ATOM bytes
  uint8_t* stg_base = h_base + KA * PAIR_H_ATOM;                         // 8 x 4 KB staging
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + PAIR_EPI_WARPS * 4096);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* acc_full = empty_bar + stages;      // [2]: accumulator 1 / 2 holds a finished tile
  uint64_t* acc_empty = acc_full + 2;           // [2]: ... has been read by every epilogue thread
  uint64_t* h_ready = acc_empty + 2;            // H written and visible to the tensor core
  uint64_t* h_free = h_ready + 1;               // phase 2 of the previous tile has finished reading H
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(h_free + 1);
  float* s_bias_a = reinterpret_cast<float*>(tmem_holder + 4);           // [C] (16-byte aligned)
  float* s_bias_b = s_bias_a + C;                                        // [C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = P.n_img * P.tiles_per_img;
  constexpr int TMEM_COLS = 2 * C;              // accumulator 1 at column 0, accumulator 2 at column C (128 or 256: powers of two)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, PAIR_EPI_THREADS); }
    mbar_init(h_ready, PAIR_EPI_THREADS);
    mbar_init(h_free, 1);
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
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        const int img = (int)fast_div_pair((uint32_t)tile, (uint32_t)P.tiles_per_img, P.magic_t);
        const int m0 = (tile - img * P.tiles_per_img) * PAIR_ROWS - 1;
#pragma unroll 1
        for (int c = 0; c < 6 * KA && ok; ++c) {            // 3 taps x KA atoms of conv_a, then of conv_b
          const int phase2 = c >= 3 * KA;
          const int cc = phase2 ? c - 3 * KA : c;
          const int tap = cc / KA, ka = cc - tap * KA;
          if (!mbar_wait(empty_bar + s, ph ^ 1, P.err, ERR_PIPE_PRODUCER)) { ok = false; break; }
          uint8_t* st = smem + (size_t)s * STAGE;
          if (!phase2) {
            mbar_expect_tx(full_bar + s, PAIR_A_SLOT + B_TILE);
            tma_load_3d(st, &P.a_map, full_bar + s, ka * 64, m0 + (tap - 1) * P.dil, img);
            tma_load_2d(st + PAIR_A_SLOT, &P.wa_map, full_bar + s, tap * C + ka * 64, 0);
          } else {
            mbar_expect_tx(full_bar + s, B_TILE);
            tma_load_2d(st + PAIR_A_SLOT, &P.wb_map, full_bar + s, tap * C + ka * 64, 0);
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0, it = 0;
      bool ok = true;
      const uint32_t h_lo = make_smem_desc_lo(smem_u32(h_base));
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++it) {
        const uint32_t tp = it & 1u;
        // ---- phase 1: conv_a into accumulator 1
        if (!mbar_wait(acc_empty + 0, tp ^ 1u, P.err, ERR_PIPE_MMA)) { ok = false; break; }
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 3 * KA && ok; ++c) {
          if (!mbar_wait(full_bar + s, ph, P.err, ERR_PIPE_MMA)) { ok = false; break; }
          tc_fence_after();
          const uint32_t da = make_smem_desc_lo(smem_u32(smem + (size_t)s * STAGE));
          const uint32_t db = da + (PAIR_A_SLOT >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_lo(tmem_base, da + 2 * k, db + 2 * k, DHI, IDESC, (c > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar + s);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (!ok) break;
        umma_commit(acc_full + 0);
        // ---- phase 2: conv_b on the shared-memory tile H into accumulator 2
        if (!mbar_wait(h_ready, tp, P.err, ERR_PIPE_MMA)) { ok = false; break; }
        if (!mbar_wait(acc_empty + 1, tp ^ 1u, P.err, ERR_PIPE_MMA)) { ok = false; break; }
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 3 * KA && ok; ++c) {
          const int tap = c / KA, ka = c - tap * KA;
          if (!mbar_wait(full_bar + s, ph, P.err, ERR_PIPE_MMA)) { ok = false; break; }
          tc_fence_after();
          const uint32_t db = make_smem_desc_lo(smem_u32(smem + (size_t)s * STAGE)) + (PAIR_A_SLOT >> 4);
          const uint32_t da = h_lo + (uint32_t)(ka * (PAIR_H_ATOM >> 4) + tap * (128 >> 4));   // view starts at buffer row `tap`
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_lo(tmem_base + C, da + 2 * k, db + 2 * k, DHI, IDESC, (c > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar + s);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (!ok) break;
        umma_commit(acc_full + 1);
        umma_commit(h_free);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // which column chunks this warp takes (two warps share a lane quarter)
    uint4* stg_h = reinterpret_cast<uint4*>(stg_base) + (size_t)ew * 256;     // 4 KB per warp: two 32 x 64-byte tiles
    uint4* stg_l = stg_h + 128;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const int h_row = lane >> 2, h_c16 = lane & 3;          // row-major role: 8 rows x 64 B per instruction
    const int so_h0 = lane * 4, so_hx = (lane >> 1) & 3;
    const int sr_h0 = h_row * 4, sr_hx = h_c16;
#define PSO_H(i) (so_h0 + ((i) ^ so_hx))
#define PSR_H(i) (32 * (i) + sr_h0 + (sr_hx ^ ((h_row >> 1) & 3)))
    const float slope_h = P.slope_h, slope_out = P.slope_out;
    const bool want_r = P.out_r_hi != nullptr;
    float amax = 0.f;
    bool ok = true;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++it) {
      const uint32_t tp = it & 1u;
      const int img = (int)fast_div_pair((uint32_t)tile, (uint32_t)P.tiles_per_img, P.magic_t);
      const int m0 = (tile - img * P.tiles_per_img) * PAIR_ROWS - 1;
      const int jrow = q * 32 + lane;            // this thread's accumulator row
      const int t = m0 + jrow;                   // its time index
      const bool in_clip = t >= 0 && t < P.L;

      // ---- epilogue 1: accumulator 1 -> H
      if (!mbar_wait(acc_full + 0, tp, P.err, ERR_PIPE_EPILOGUE)) { ok = false; break; }
      tc_fence_after();
      if (!mbar_wait(h_free, tp ^ 1u, P.err, ERR_PIPE_EPILOGUE)) { ok = false; break; }
#pragma unroll 1
      for (int j = half; j < NCHUNK; j += 2) {
        float v[32];
        tmem_ld_32x32(tmem_base + lane_bits + j * 32, v);
        if (j + 2 >= NCHUNK) {                   // last chunk of this warp is in registers: hand accumulator 1 back
          tc_fence_before();
          mbar_arrive(acc_empty + 0);
        }
        const float4* bp = reinterpret_cast<const float4*>(s_bias_a + j * 32);
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
        // K-major SWIZZLE_128B image: h row jrow -> buffer row jrow + 1; 16-byte chunk index XOR (buffer row & 7)
        const int brow = jrow + 1;
        uint8_t* rowp = h_base + (size_t)(j >> 1) * PAIR_H_ATOM + (size_t)brow * 128;
        const int c0 = (j & 1) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(rowp + (((c0 + i) ^ (brow & 7)) << 4)) = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
      }
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async proxy
      mbar_arrive(h_ready);

      // ---- epilogue 2: accumulator 2 + bias_b + x -> x_new planes, activated plane
      if (!mbar_wait(acc_full + 1, tp, P.err, ERR_PIPE_EPILOGUE)) { ok = false; break; }
      tc_fence_after();
      const int wrow0 = q * 32;                                  // first accumulator row of this warp
      // element offsets of this warp's first row (row -1 of the first tile is never dereferenced: guarded by jr >= 1)
      const long in_base = ((long)img * P.L + (m0 + wrow0)) * C;
      const long out_base = ((long)img * P.out_img_rows + P.out_row0 + (m0 + wrow0)) * C;
#pragma unroll 1
      for (int j = half; j < NCHUNK; j += 2) {
        float v[32];
        tmem_ld_32x32(tmem_base + lane_bits + C + j * 32, v);
        if (j + 2 >= NCHUNK) {
          tc_fence_before();
          mbar_arrive(acc_empty + 1);
        }
        const float4* bp = reinterpret_cast<const float4*>(s_bias_b + j * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = bp[i];
          v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
        }
        // residual x (hi + lo planes): coalesced row-major loads -> staging -> own row
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + h_row, jr = wrow0 + rr, tt = m0 + jr;
          uint4 xh = make_uint4(0, 0, 0, 0), xl = make_uint4(0, 0, 0, 0);
          if (jr >= 1 && jr <= PAIR_ROWS && tt < P.L) {
            const long o = in_base + (long)rr * C + j * 32;
            xh = __ldg(reinterpret_cast<const uint4*>(P.resid_hi + o) + h_c16);
            xl = __ldg(reinterpret_cast<const uint4*>(P.resid_lo + o) + h_c16);
          }
          stg_h[PSR_H(i)] = xh;
          stg_l[PSR_H(i)] = xl;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 xh = stg_h[PSO_H(i)], xl = stg_l[PSO_H(i)];
          const __half2* ph2 = reinterpret_cast<const __half2*>(&xh);
          const __half2* pl2 = reinterpret_cast<const __half2*>(&xl);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fh = __half22float2(ph2[k]), fl = __half22float2(pl2[k]);
            v[8 * i + 2 * k] += fh.x + fl.x;
            v[8 * i + 2 * k + 1] += fh.y + fl.y;
          }
        }
        if (want_r) {                           // raw hi/lo planes of x_new
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            const float2 f = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
            hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            stg_h[PSO_H(i)] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            stg_l[PSO_H(i)] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + h_row, jr = wrow0 + rr, tt = m0 + jr;
            if (jr >= 1 && jr <= PAIR_ROWS && tt < P.L) {
              const long o = in_base + (long)rr * C + j * 32;              // x_new planes have the clip's own row pitch
              reinterpret_cast<uint4*>(P.out_r_hi + o)[h_c16] = stg_h[PSR_H(i)];
              reinterpret_cast<uint4*>(P.out_r_lo + o)[h_c16] = stg_l[PSR_H(i)];
            }
          }
        }
        {                                         // activated plane for the next pair / stage
          uint32_t hi[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a0 = fmaxf(v[2 * i], v[2 * i] * slope_out), a1 = fmaxf(v[2 * i + 1], v[2 * i + 1] * slope_out);
            if (in_clip && jrow >= 1 && jrow <= PAIR_ROWS) amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
            const __half2 hh = __floats2half2_rn(a0, a1);
            hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) stg_h[PSO_H(i)] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + h_row, jr = wrow0 + rr, tt = m0 + jr;
            if (jr >= 1 && jr <= PAIR_ROWS && tt < P.L)
              reinterpret_cast<uint4*>(P.out_a + out_base + (long)rr * C + j * 32)[h_c16] = stg_h[PSR_H(i)];
          }
        }
      }
    }
    if (!(amax <= 65504.f) && P.err) atomicCAS(P.err, 0, ERR_FP16_OVERFLOW);
#undef PSO_H
#undef PSR_H
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, TMEM_COLS);
  }
}

size_t pair_tc_smem_bytes(int C, int stages) {
  const size_t stage = PAIR_A_SLOT + (size_t)C * 128;
  return stages * stage + (size_t)(C / 64) * PAIR_H_ATOM + PAIR_EPI_WARPS * 4096 + (2 * stages + 8) * 8 + 32 + 2 * C * 4 + 1024;
}

template <int C>
static cudaError_t launch_pair_c(const PairParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pair_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  pair_tc_kernel<C><<<p.grid, 64 + PAIR_EPI_THREADS, pair_tc_smem_bytes(C, p.stages), stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pair_tc(const PairParams& p, cudaStream_t stream) {
  if (p.C == 64) return launch_pair_c<64>(p, stream);
  if (p.C == 128) return launch_pair_c<128>(p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace vf
