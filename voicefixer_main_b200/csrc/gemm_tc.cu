// tcgen05 implementation of the flat-shift multi-tap GEMM (see gemm.cuh).
//
// One CTA = one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0 : TMA producer  - streams A (activation rows, shifted per tap) and B (packed weights) tiles
//            into a `stages`-deep shared-memory ring (SWIZZLE_128B rows for BK = 64, SWIZZLE_64B for BK = 32)
//   warp 1 : TMEM owner + single-thread tcgen05.mma issuer; accumulator 128 lanes x BN fp32 columns in TMEM
//   warps 2-5 : epilogue - tcgen05.ld one TMEM lane (= output row) per thread, fused bias / residual /
//            BN-affine / activation / hi-lo split / concat placement / 1x1 head, vectorised stores
// Several CTAs co-reside per SM when the tile is small (BN = 32: 20 KB per stage, 32 TMEM columns), which
// overlaps one CTA's epilogue with another's main loop.
#include "gemm.cuh"
#include "ptx.cuh"

namespace vf {

template <int BN, int BK>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ GemmTcParams P) {
  constexpr int A_BYTES = GEMM_BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  constexpr int ROW_BYTES = BK * 2;
  constexpr int KSTEPS = BK / 16;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const GemmProblem& pr = P.prob;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int img = blockIdx.x / pr.m_tiles;
  const int m0 = (blockIdx.x - img * pr.m_tiles) * GEMM_BM;
  const int n0 = blockIdx.y * BN;
  const bool three = pr.terms == 3;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&P.a_hi[0]);
    tma_prefetch_desc(&P.b_hi);
    if (three) {
      tma_prefetch_desc(&P.a_lo[0]);
      tma_prefetch_desc(&P.b_lo);
    }
  }
  if (warp == 1) tmem_alloc<BN>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = (three ? 2u : 1u) * (A_BYTES + B_BYTES);
      int it = 0;
      bool ok = true;
      for (int t = 0; t < pr.ntaps && ok; ++t) {
        const GemmTap tap = pr.taps[t];
        for (int c = 0; c < tap.nch; c += BK, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          if (!mbar_wait(empty_bar + s, ph ^ 1, pr.epi.err, ERR_PIPE_PRODUCER)) { ok = false; break; }
          uint8_t* st = smem + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(full_bar + s, tx_bytes);
          tma_load_3d(st, &P.a_hi[tap.src], full_bar + s, tap.c_off + c, m0 + tap.a_off, img);
          tma_load_2d(st + 2 * A_BYTES, &P.b_hi, full_bar + s, tap.k_off + c, n0);
          if (three) {
            tma_load_3d(st + A_BYTES, &P.a_lo[tap.src], full_bar + s, tap.c_off + c, m0 + tap.a_off, img);
            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &P.b_lo, full_bar + s, tap.k_off + c, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN);
      int it = 0;
      bool ok = true;
      for (int t = 0; t < pr.ntaps && ok; ++t) {
        const int nch = pr.taps[t].nch;
        for (int c = 0; c < nch; c += BK, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          if (!mbar_wait(full_bar + s, ph, pr.epi.err, ERR_PIPE_MMA)) { ok = false; break; }
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES;
          const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint64_t da_hi = make_smem_desc(a_hi + k * 32, ROW_BYTES);
            const uint64_t db_hi = make_smem_desc(b_hi + k * 32, ROW_BYTES);
            umma_f16(tmem_base, da_hi, db_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
            if (three) {
              umma_f16(tmem_base, da_hi, make_smem_desc(b_lo + k * 32, ROW_BYTES), idesc, 1u);
              umma_f16(tmem_base, make_smem_desc(a_lo + k * 32, ROW_BYTES), db_hi, idesc, 1u);
            }
          }
          umma_commit(empty_bar + s);   // frees the smem slot once these MMAs have read it
        }
      }
      umma_commit(tmem_full_bar);       // accumulator complete
    }
    __syncwarp();
  } else {
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int r = m0 + q * 32 + lane;
    if (mbar_wait(tmem_full_bar, 0, pr.epi.err, ERR_PIPE_EPILOGUE)) {
      tc_fence_after();
      float head_acc = 0.f;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 32, v);
        epilogue_chunk(pr.epi, img, r, n0 + j * 32, v, head_acc);
      }
      epilogue_head(pr.epi, img, r, head_acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

template <int BN, int BK>
static cudaError_t launch_one(const GemmTcParams& p, cudaStream_t stream) {
  constexpr int STAGE_BYTES = 2 * (GEMM_BM * BK * 2 + BN * BK * 2);
  const size_t smem = (size_t)p.stages * STAGE_BYTES + (2 * p.stages + 1) * 8 + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(p.prob.n_img * p.prob.m_tiles, p.prob.N / BN);
  gemm_tc_kernel<BN, BK><<<grid, 192, smem, stream>>>(p);
  return cudaGetLastError();
}

int gemm_tc_stage_bytes(int bn, int bk) { return 2 * (GEMM_BM * bk * 2 + bn * bk * 2); }

cudaError_t launch_gemm_tc(const GemmTcParams& p, int bn, int bk, cudaStream_t stream) {
  if (bk == 64) {
    if (bn == 128) return launch_one<128, 64>(p, stream);
    if (bn == 64) return launch_one<64, 64>(p, stream);
    if (bn == 32) return launch_one<32, 64>(p, stream);
  } else if (bk == 32) {
    if (bn == 128) return launch_one<128, 32>(p, stream);
    if (bn == 64) return launch_one<64, 32>(p, stream);
    if (bn == 32) return launch_one<32, 32>(p, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vf
