// Probe (not product code): 5-D TMA store of a transposed-conv output tile [C, phase, q, image, plane] with a box that starts
// at q = -1 (partly out of bounds) - legal?  Also the 4-D form with a negative row.  Prints the CUDA error of each case and
// checks the bytes that landed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/probe5 tools/probe_tma5d.cu -lcuda && gpurun_out/probe5
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../voicefixer_main_b200/csrc/ptx.cuh"
using namespace vf;

__global__ void k5(const __grid_constant__ CUtensorMap m, int c0, int p, int q, int img, int planes) {
  __shared__ __align__(1024) __half tile[2 * 32 * 32];
  for (int i = threadIdx.x; i < planes * 1024; i += 32) tile[i] = __float2half((float)(i % 1024 / 32 + 1));   // value = box row + 1
  fence_proxy_async();
  __syncwarp();
  if (threadIdx.x == 0) {
    tma_store_5d(&m, tile, c0, p, q, img, 0);
    tma_store_commit();
    tma_store_wait_all();
  }
}
__global__ void k4(const __grid_constant__ CUtensorMap m, int c0, int row, int img) {
  __shared__ __align__(1024) __half tile[2 * 32 * 32];
  for (int i = threadIdx.x; i < 2048; i += 32) tile[i] = __float2half((float)(i % 1024 / 32 + 1));
  fence_proxy_async();
  __syncwarp();
  if (threadIdx.x == 0) {
    tma_store_4d(&m, tile, c0, row, img, 0);
    tma_store_commit();
    tma_store_wait_all();
  }
}

int main() {
  const int C = 64, s = 7, Lq = 40, L = s * Lq, n_img = 2, planes = 2;
  const size_t cnt = (size_t)n_img * L * C;
  __half* d;
  cudaMalloc(&d, 2 * cnt * sizeof(__half));
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Enc enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  for (int order = 0; order < 2; ++order) {
    CUtensorMap m;
    cuuint64_t dimsA[5] = {(cuuint64_t)C, (cuuint64_t)s, (cuuint64_t)Lq, (cuuint64_t)n_img, (cuuint64_t)planes};
    cuuint64_t strA[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * s, (cuuint64_t)L * C * 2, cnt * 2};
    cuuint32_t boxA[5] = {32, 1, 32, 1, (cuuint32_t)planes};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, d, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     order ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 5d (swizzle %s) -> %d\n", order ? "none" : "64B", (int)r);
    const int qs[3] = {0, 20, -1};
    for (int qi = 0; qi < 3; ++qi) {
      cudaMemset(d, 0, 2 * cnt * sizeof(__half));
      k5<<<1, 32>>>(m, 32, 3, qs[qi], 1, planes);
      cudaError_t e = cudaDeviceSynchronize();
      printf("  5d store q=%d: %s\n", qs[qi], cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      std::vector<__half> h(2 * cnt);
      cudaMemcpy(h.data(), d, 2 * cnt * sizeof(__half), cudaMemcpyDeviceToHost);
      int nz = 0, bad = 0;
      for (size_t i = 0; i < 2 * cnt; ++i) {
        const float v = __half2float(h[i]);
        if (v == 0.f) continue;
        ++nz;
        const size_t j = i % cnt;
        const int c = j % C, t = (j / C) % L, img = j / ((size_t)C * L);
        const int q = t / s, p = t % s;
        if (!(img == 1 && p == 3 && c >= 32 && q - qs[qi] >= 0 && q - qs[qi] < 32 && v == (float)(q - qs[qi] + 1))) ++bad;
      }
      printf("    non-zero %d (expect %d), misplaced %d\n", nz, planes * 32 * (qs[qi] < 0 ? 31 : (qs[qi] + 32 > Lq ? Lq - qs[qi] : 32)), bad);
    }
  }
  {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)n_img, 2};
    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2, cnt * 2};
    cuuint32_t box[4] = {32, 32, 1, 2};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 4d -> %d\n", (int)r);
    for (int row : {0, -1}) {
      k4<<<1, 32>>>(m, 0, row, 0);
      cudaError_t e = cudaDeviceSynchronize();
      printf("  4d store row=%d: %s\n", row, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
