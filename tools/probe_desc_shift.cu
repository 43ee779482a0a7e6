// Probe (not product code): can a tcgen05 shared-memory descriptor start on an arbitrary ROW of a TMA-written
// SWIZZLE_128B / SWIZZLE_64B K-major tile?  That is what "halo" loads need: load 128+h activation rows once and
// issue the 3 horizontally adjacent conv taps as MMAs on row-shifted views instead of 3 TMA loads.
// For every row shift r = 0..8 it tries base_offset = 0 and base_offset = f(address) and reports which matches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/probe tools/probe_desc_shift.cu -lcuda && gpurun_out/probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../voicefixer_main_b200/csrc/ptx.cuh"
using namespace vf;

constexpr int ROWS = 144, N = 64;

template <int BK>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int shift,
                             int base_mode, float* out) {
  extern __shared__ __align__(16) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  constexpr int ROWB = BK * 2;
  uint8_t* sa = smem;                     // ROWS x ROWB
  uint8_t* sb = smem + 32768;             // N x ROWB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_dyn(holder, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * ROWB + N * ROWB);
    tma_load_2d(sa, &ta, bar, 0, 0);
    tma_load_2d(sb, &tb, bar, 0, 0);
    mbar_wait(bar, 0, nullptr, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sa) + shift * ROWB;
    uint64_t da = make_smem_desc(a_addr, ROWB);
    uint32_t bo = 0;
    if (base_mode == 1) bo = (a_addr >> 7) & 7;                 // PTX formula for the 128B pattern
    if (base_mode == 2) bo = (a_addr / ROWB) & 7;               // row index inside the 8-row atom
    da |= static_cast<uint64_t>(bo) << 49;
    const uint64_t db = make_smem_desc(smem_u32(sb), ROWB);
    constexpr uint32_t idesc = make_idesc_f16(128, N);
#pragma unroll
    for (int k = 0; k < BK / 16; ++k) umma_f16(tm, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(bar + 1);
  }
  __syncthreads();
  mbar_wait(bar + 1, 0, nullptr, 0);
  tc_fence_after();
  float v[32];
  for (int j = 0; j < N / 32; ++j) {
    tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + j * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_dyn(tm, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BK>
void run(EncodeFn enc) {
  std::vector<__half> ha(ROWS * BK), hb(N * BK);
  std::vector<float> fa(ROWS * BK), fb(N * BK);
  unsigned s = 7;
  for (size_t i = 0; i < ha.size(); ++i) { s = s * 1664525u + 1013904223u; ha[i] = __float2half_rn(((s >> 9) & 1023) / 512.f - 1.f); fa[i] = __half2float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { s = s * 1664525u + 1013904223u; hb[i] = __float2half_rn(((s >> 9) & 1023) / 512.f - 1.f); fb[i] = __half2float(hb[i]); }
  __half *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint64_t dA[2] = {(cuuint64_t)BK, ROWS}, dB[2] = {(cuuint64_t)BK, N};
  cuuint64_t st[1] = {(cuuint64_t)BK * 2};
  cuuint32_t bA[2] = {(cuuint32_t)BK, ROWS}, bB[2] = {(cuuint32_t)BK, N}, es[2] = {1, 1};
  enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, dA, st, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dB, st, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> out(128 * N);
  for (int shift = 0; shift <= 9; ++shift)
    for (int mode = 0; mode < 3; ++mode) {
      cudaMemset(dout, 0, 128 * N * 4);
      probe_kernel<BK><<<1, 128, 60 * 1024>>>(ta, tb, shift, mode, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("BK=%d shift=%d mode=%d CUDA error %s\n", BK, shift, mode, cudaGetErrorString(e)); return; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double md = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < BK; ++k) ref += (double)fa[(m + shift) * BK + k] * fb[n * BK + k];
          md = std::fmax(md, std::fabs(ref - out[m * N + n]));
        }
      printf("BK=%d row_shift=%d base_offset_mode=%d max_abs_err=%.3e %s\n", BK, shift, mode, md, md < 1e-2 ? "OK" : "WRONG");
    }
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
  run<64>((EncodeFn)fn);
  run<32>((EncodeFn)fn);
  return 0;
}
