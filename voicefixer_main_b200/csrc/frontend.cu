// Stage A: fused STFT magnitude + mel projection (+ log10).
//
// Replaces, per frame, the reference chain
//   FDomainHelper.wav_to_spectrogram_phase  tools/pytorch/modules/fDomainHelper.py:60-89
//     (torchlibrosa STFT: reflect pad 1024, periodic-hann windowed 2048-point DFT as two conv1d, hop 441)
//   MelScale.forward                        tools/pytorch/mel_scale.py:52-64 (gsr_voicefixer.py:180)
//   to_log                                  tools/pytorch/pytorch_util.py:157-159
// with one CTA per (frame, clip): windowed load straight from the un-padded waveform (reflection by index
// math), a 1024-point complex radix-4 Stockham FFT in shared memory on the even/odd-packed real frame,
// the real-FFT split, |.| with the 1e-8 power clamp, and the triangular mel filterbank applied in its
// sparse form (2018 non-zeros instead of a 1025x128 dense matmul).  The 1025-bin spectrogram is only
// written when the caller asks for it (the GSR path discards it, eval_gsr_voicefixer.py:51).
// HBM traffic on the GSR path: N*4 bytes in, T*128*4 bytes out per clip (2.277 MB for 10 s).
#include "kernels.cuh"

namespace vf {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(256) frontend_kernel(FrontendParams p) {
  __shared__ float2 buf0[1024];
  __shared__ float2 buf1[1024];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* x = p.wav + (size_t)b * p.n;

  // windowed, reflect-padded frame packed as z[n] = x[2n] + i x[2n+1]
  for (int n = tid; n < 1024; n += 256) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      long g = (long)t * 441 + 2 * n + e - 1024;
      if (g < 0) g = -g;
      if (g >= p.n) g = 2L * (p.n - 1) - g;
      v[e] = __ldg(x + g) * __ldg(p.window + 2 * n + e);
    }
    buf0[n] = make_float2(v[0], v[1]);
  }
  __syncthreads();

  // 5 radix-4 Stockham passes (Ns = 1, 4, 16, 64, 256), natural-order output
  float2* src = buf0;
  float2* dst = buf1;
#pragma unroll
  for (int pass = 0; pass < 5; ++pass) {
    const int ns = 1 << (2 * pass);
    const int k = tid & (ns - 1);
    const int tw_step = k * (256 / ns);
    float2 a = src[tid];
    float2 bq = src[tid + 256];
    float2 c = src[tid + 512];
    float2 d = src[tid + 768];
    if (pass > 0) {
      bq = cmul(bq, __ldg(p.tw1024 + tw_step));
      c = cmul(c, __ldg(p.tw1024 + 2 * tw_step));
      d = cmul(d, __ldg(p.tw1024 + 3 * tw_step));
    }
    const float2 s0 = make_float2(a.x + c.x, a.y + c.y), s1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 s2 = make_float2(bq.x + d.x, bq.y + d.y), s3 = make_float2(bq.x - d.x, bq.y - d.y);
    const int o = ((tid - k) << 2) + k;
    dst[o] = make_float2(s0.x + s2.x, s0.y + s2.y);
    dst[o + ns] = make_float2(s1.x + s3.y, s1.y - s3.x);          // s1 - i*s3
    dst[o + 2 * ns] = make_float2(s0.x - s2.x, s0.y - s2.y);
    dst[o + 3 * ns] = make_float2(s1.x - s3.y, s1.y + s3.x);      // s1 + i*s3
    __syncthreads();
    float2* tmp = src; src = dst; dst = tmp;
  }
  // src == buf1 now holds Z[0..1023]; buf0 is free and becomes the magnitude row
  float* mag = reinterpret_cast<float*>(buf0);
  const size_t frame = (size_t)b * p.T + t;
  for (int k = tid; k <= 1024; k += 256) {
    const float2 zk = src[k & 1023];
    const float2 zr = src[(1024 - k) & 1023];
    const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));      // (Zk + conj Zr)/2
    const float2 o = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));     // -i (Zk - conj Zr)/2
    const float2 wo = cmul(o, __ldg(p.tw2048 + k));
    const float re = e.x + wo.x, im = e.y + wo.y;
    const float m = sqrtf(fmaxf(re * re + im * im, 1e-8f));       // fDomainHelper.py:62, eps = 1e-8
    mag[k] = m;
    if (p.sp_out) {
      p.sp_out[frame * 1025 + k] = m;
      if (p.cos_out) {
        p.cos_out[frame * 1025 + k] = re / m;
        p.sin_out[frame * 1025 + k] = im / m;
      }
    }
  }
  __syncthreads();
  if (tid < 128) {
    const int f0 = __ldg(p.fb_f0 + tid), len = __ldg(p.fb_len + tid);
    const float* w = p.fb_val + __ldg(p.fb_ofs + tid);
    float acc = 0.f;
    for (int j = 0; j < len; ++j) acc = fmaf(mag[f0 + j], __ldg(w + j), acc);
    if (p.mel_out) p.mel_out[frame * 128 + tid] = acc;
    if (p.logmel_out) p.logmel_out[frame * 128 + tid] = log10f(fmaxf(acc, 1e-8f));
  }
}

cudaError_t launch_frontend(const FrontendParams& p, cudaStream_t stream) {
  dim3 grid(p.T, p.batch);
  frontend_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace vf
