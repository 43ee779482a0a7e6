// Probe (not product code): cycles per tcgen05.mma (M=128, K=16, fp16) issued by one thread, as a function of N
// and of the number of independent TMEM accumulators the chain alternates between.  Answers: is a chain of
// small-N MMAs into ONE accumulator latency-bound, and would interleaving tiles / accumulators help?
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../voicefixer_main_b200/csrc/ptx.cuh"
using namespace vf;

template <int N>
__global__ void probe(int reps, int naccs, int both_ops, long long* out) {
  extern __shared__ __align__(16) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // halfs = 1.0
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc_dyn(holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  if (threadIdx.x < 32 && elect_one()) {
    const uint64_t da = make_smem_desc(smem_u32(smem), 128), db = make_smem_desc(smem_u32(smem) + 16384, 128);
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tm + (r % naccs) * N;
      umma_f16(d, da + 2 * (r & 3), db + 2 * (r & 3), idesc, 1u);
      if (both_ops) umma_f16(d, da + 2 * ((r + 1) & 3), db + 2 * (r & 3), idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0, nullptr, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc_dyn(tm, 512);
}

template <int N>
void run(long long* d) {
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int both = 0; both < 2; ++both)
    for (int naccs = 1; naccs <= 512 / N && naccs <= 4; naccs *= 2) {
      const int reps = 4096;
      long long h[2];
      probe<N><<<1, 128, 70 * 1024>>>(reps, naccs, both, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d error %s\n", N, cudaGetErrorString(e)); return; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const int n_mma = reps * (both ? 2 : 1);
      printf("N=%3d accumulators=%d mma_per_iter=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (nominal floor %d)\n", N, naccs,
             both ? 2 : 1, (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * N / 256);
    }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<32>(d); run<64>(d); run<128>(d); run<256>(d);
  return 0;
}
