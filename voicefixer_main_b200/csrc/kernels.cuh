// Parameter blocks and launchers of the non-GEMM kernels on the hot path.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vf {

struct FrontendParams {
  const float* wav;      // [batch, n]
  long n;
  int batch, T;          // T = 1 + n / 441
  const float* window;   // [2048] periodic hann
  const float2* tw1024;  // e^{-2 pi i j / 1024}
  const float2* tw2048;  // e^{-2 pi i k / 2048}, k = 0..1024
  const int* fb_f0;      // [128] first non-zero frequency bin of each mel filter
  const int* fb_len;     // [128]
  const int* fb_ofs;     // [128] offset into fb_val
  const float* fb_val;
  float* sp_out;         // [batch, T, 1025] or null
  float* cos_out;        // with sp_out, or null
  float* sin_out;
  float* mel_out;        // [batch, T, 128] linear mel or null
  float* logmel_out;     // [batch, T, 128] log10(clip(mel, 1e-8)) or null
};
cudaError_t launch_frontend(const FrontendParams& p, cudaStream_t stream);

struct PlanePtr {
  __half* hi;
  __half* lo;
};

// encoder_block1.conv_block1: BN(1ch) -> LeakyReLU -> Conv3x3 1->32, then bn2 -> LeakyReLU of the same
// block fused in (the only consumer), plus the 1->32 1x1 shortcut of the raw input (modules.py:263-271).
struct UnetFirstParams {
  const float* logmel;   // [batch, T, 128]; bins 0..126 feed the UNet (unet.py:78)
  int batch, T, Tp;      // Tp = T padded to a multiple of 64 with zero *input* rows (unet.py:75-77)
  float bn1_scale, bn1_shift;
  const float* w1;       // [32][9]
  const float* bn2_scale;  // [32]
  const float* bn2_shift;
  const float* w_sc;     // [32] shortcut weight
  const float* b_sc;     // [32] shortcut bias
  float slope;
  PlanePtr a2;           // [batch, Tp*128, 32] act(bn2(conv1(...))), pad column zero
  float* sc_raw;         // [batch, Tp*128, 32] fp32 shortcut(x), pad column zero
  int* err;
};
cudaError_t launch_unet_first(const UnetFirstParams& p, cudaStream_t stream);

// avg_pool2d(2,2) (modules.py:183) + the consumer's BN/LeakyReLU + hi/lo split.
struct PoolParams {
  const float* in;       // [batch, H*Wp, C] fp32
  int batch, H, Wp, C;
  PlanePtr out_r;        // [batch, (H/2)*(Wp/2), C] pooled raw
  PlanePtr out_a;        // act(scale*pooled+shift)
  float* out_raw;        // fp32 pooled or null
  const float* a_scale;
  const float* a_shift;
  float slope;
  int* err;
};
cudaError_t launch_pool(const PoolParams& p, cudaStream_t stream);

// to_log / from_log (tools/pytorch/pytorch_util.py:157-163) for the stage-level API.
cudaError_t launch_to_log(const float* in, float* out, size_t n, int* neg_count, cudaStream_t stream);
cudaError_t launch_from_log(const float* in, float* out, size_t n, cudaStream_t stream);

// Vocoder prologue: mel / w -> dB -> normalise -> [T + tail, 128] planes with the constant tail.
struct VocCondParams {
  const float* mel;      // [batch, T, 128]; linear mel, or log10 mel when is_log (from_log fused)
  int is_log;
  int batch, T, Tv;
  const float* weight;       // [128] per-bin mel weight (divided out)
  float amp_floor, ref_db, min_db, tail_value;
  const float* band_sums;    // [batch][2] (target, estimate) low-band sums from launch_band_energy, or null:
                             // amp_to_original_f (tools/utils.py:50-55) scales the estimate by target/estimate
  PlanePtr out;          // [batch, Tv, 128]
};
cudaError_t launch_voc_condition(const VocCondParams& p, cudaStream_t stream);

// amp_to_original_f, reduction half: per clip, sums over frames and mel bins [5, int(128*0.2)) of the noisy
// linear mel (target) and of from_log(restored log-mel) (estimate).  sums must be zeroed before the launch.
cudaError_t launch_band_energy(const float* mel_target_lin, const float* logmel_est, int batch, int T, float* sums,
                               cudaStream_t stream);

// nn.ReflectionPad1d(3): rows [3, L+3) of each image are already written; fill 3 + 3 mirrored rows.
cudaError_t launch_reflect_fill(PlanePtr planes, int batch, int L, int C, int pad, cudaStream_t stream);

// Tail: ReflectionPad(3) (pre-filled) + Conv1d(C -> 1, k7) + tanh, plus the per-clip peak |out|.
struct VocTailParams {
  PlanePtr in;           // [batch, L + 6, C]
  int batch, L, C, terms;
  const float* w;        // [7][C]
  float bias;
  float* wav;            // [batch, L]
  unsigned int* peak_bits;   // [batch] max |out| as float bits (non-negative floats order like uints)
};
cudaError_t launch_voc_tail(const VocTailParams& p, cudaStream_t stream);

// eval_gsr_voicefixer.py:68-72: out /= max|out| if it exceeds 1; trim_center (tools/utils.py:57-70).
struct FinalizeParams {
  const float* wav;      // [batch, L]
  const unsigned int* peak_bits;
  int batch;
  long L, n, skip;       // out[b, i] = wav[b, skip + i], i < n
  float* out;            // [batch, out_ld]
  long out_ld, out_off;
};
cudaError_t launch_finalize(const FinalizeParams& p, cudaStream_t stream);
cudaError_t launch_pcm16(const float* in, int16_t* out, size_t n, cudaStream_t stream);

}  // namespace vf
