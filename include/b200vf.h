/* b200vf.h - C ABI of libb200vf.so: the B200-native VoiceFixer inference hot path.
 *
 * The reference (haoheliu/voicefixer_main) has no FFI for this path; its boundary is the Python object
 * protocol that eval_gsr_voicefixer.py:handler() consumes.  Each entry point below
 * replaces one reference interface and is what a ctypes / cgo / JNI binding of that interface would bind:
 *
 *   vf_create / vf_load_weights   Model(hp, ...).load_from_checkpoint(ckpt); model.eval(); model.to(device)
 *                                 eval_gsr_voicefixer.py:31-35,40; config keys config/vctk_base_voicefixer_unet.json:68-78
 *   vf_frontend                   VoiceFixer.pre  models/gsr_voicefixer.py:178-181
 *                                 = FDomainHelper.wav_to_spectrogram_phase tools/pytorch/modules/fDomainHelper.py:67-89
 *                                 + MelScale.forward tools/pytorch/mel_scale.py:52-64
 *   vf_unet_mel                   VoiceFixer.forward models/gsr_voicefixer.py:183-193 -> Generator.forward :86-91
 *                                 -> UNetResComplex_100Mb.forward models/components/unet.py:60-103
 *   vf_vocoder                    model.vocoder(mel) eval_gsr_voicefixer.py:66 (third-party voicefixer.Vocoder)
 *   vf_restore / vf_restore_host  one iteration of the segment loop of handler(), eval_gsr_voicefixer.py:49-74:
 *                                 pre -> model -> from_log -> vocoder -> peak normalise -> trim_center
 *   vf_to_log / vf_from_log       tools/pytorch/pytorch_util.py:157-163
 *   vf_to_pcm16                   the int16 conversion of save_wave, tools/file/wav.py:22-24
 *
 * Conventions: every function returns 0 on success or a negative VF_E* code and never throws; the message is
 * available from vf_last_error().  All tensor arguments are contiguous fp32.  Unless a name ends in `_host`,
 * pointers are DEVICE pointers owned by the caller (e.g. PyTorch tensors); the library never frees or retains
 * them.  `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); calls are
 * asynchronous with respect to the host and contain no hidden synchronisation, except where stated.  One
 * context per device; calls on one context must be serialised by the caller.  There is no CPU fallback: without
 * a CUDA device every call fails with VF_ENODEVICE.
 */
#ifndef B200VF_H_
#define B200VF_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define VF_API __attribute__((visibility("default")))
#else
#define VF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define VF_OK 0
#define VF_EINVAL (-1)     /* bad argument / unsupported shape */
#define VF_ENODEVICE (-2)  /* no usable CUDA device */
#define VF_ECUDA (-3)      /* CUDA runtime / driver error */
#define VF_ESTATE (-4)     /* weights not loaded, missing tensor, ... */
#define VF_EDEVICE (-5)    /* sticky device-side error flag (fp16 range overflow, pipeline time-out) */
#define VF_EASSERT (-6)    /* reference assertion would have fired (to_log on negative input) */

typedef struct vf_ctx vf_ctx;

/* Geometry of the front end (reference config "model"/"data" keys) and of the vocoder restatement. */
typedef struct vf_config {
  int sample_rate;        /* 44100 */
  int n_fft;              /* 2048  (window_size) */
  int hop;                /* 441   (hop_size) */
  int n_mels;             /* 128   (mel_freq_bins) */
  /* vocoder generator (voicefixer_main_b200/arch.py:VocoderConfig) */
  int voc_cond_channels;  /* 512 */
  int voc_cond_layers;    /* 5 */
  int voc_channels;       /* 1024 */
  int voc_num_stages;     /* 4 */
  int voc_scales[8];      /* 7,7,3,3 */
  int voc_depth[8];       /* 8,8,8,8 */
  float voc_stage_slope;  /* 0.2 */
  float voc_res_slope;    /* 0.01 */
  float voc_min_db;       /* -115 */
  float voc_ref_db;       /* 20 */
  float voc_amp_floor;    /* 1e-5 */
  float voc_tail_value;   /* -4 */
  int voc_tail_base;      /* 4 */
  double voc_mel_weight_a;
  double voc_mel_weight_b;
} vf_config;

/* Fills *cfg with the reference defaults listed above. */
VF_API void vf_default_config(vf_config* cfg);

typedef struct vf_tensor_desc {
  const char* name;      /* reference state-dict key, e.g. "generator.analysis_module.encoder_block1.conv_block1.bn1.weight",
                            "mel.fb", or "vocoder.<key>" (arch.py:vocoder_keys) */
  const void* data;      /* fp32, contiguous */
  int ndim;
  int64_t shape[4];
  int on_device;         /* 0: host pointer, 1: device pointer */
} vf_tensor_desc;

VF_API int vf_create(vf_ctx** out, int device, const vf_config* cfg);
VF_API void vf_destroy(vf_ctx* ctx);
VF_API const char* vf_last_error(vf_ctx* ctx);   /* ctx may be NULL: error of the last failed vf_create */

/* Copies and packs the tensors (BN folded to per-channel affine, conv weights to K-major fp16 hi/lo
 * matrices, mel filterbank to its sparse form).  Synchronous.  Missing keys -> VF_ESTATE. */
VF_API int vf_load_weights(vf_ctx* ctx, const vf_tensor_desc* descs, int n);

/* wav [B,N] -> mel_out [B,T,128] linear mel (T = 1 + N/hop); optional sp/cos/sin [B,T,1025] (NULL to skip). */
VF_API int vf_frontend(vf_ctx* ctx, const float* wav, int batch, int64_t n_samples, float* mel_out, float* sp_out,
                float* cos_out, float* sin_out, void* stream);

/* mel_lin [B,T,128] (non-negative) -> logmel_out [B,T,128] = unet(log10 mel) + log10 mel.
 * Negative inputs are counted on the device; vf_check_errors() then reports VF_EASSERT (to_log's assert). */
VF_API int vf_unet_mel(vf_ctx* ctx, const float* mel_lin, int batch, int frames, float* logmel_out, void* stream);

/* mel_lin [B,T,128] -> wav_out [B,L], L = vf_vocoder_out_len(ctx, T). */
VF_API int vf_vocoder(vf_ctx* ctx, const float* mel_lin, int batch, int frames, float* wav_out, void* stream);
VF_API int64_t vf_vocoder_out_len(vf_ctx* ctx, int frames);

/* Fused stages A -> B -> C + peak normalise + centre trim: wav [B,N] -> wav_out [B,N] (device pointers). */
VF_API int vf_restore(vf_ctx* ctx, const float* wav, int batch, int64_t n_samples, float* wav_out, void* stream);
/* Same through HOST buffers (pinned for true asynchrony): H2D copy, vf_restore, D2H copy on `stream`.
 * The caller synchronises the stream before reading out_host. */
VF_API int vf_restore_host(vf_ctx* ctx, const float* wav_host, int batch, int64_t n_samples, float* out_host, void* stream);
/* Copies the intermediate results of the last vf_restore of this (batch, n_samples) into caller buffers
 * [B,T,128] (either may be NULL): the linear mel of stage A and the restored log10 mel of stage B. */
VF_API int vf_restore_stages(vf_ctx* ctx, int batch, int64_t n_samples, float* mel_lin_out, float* log_mel_out,
                             void* stream);

VF_API int vf_to_log(vf_ctx* ctx, const float* in, float* out, int64_t n, void* stream);
VF_API int vf_from_log(vf_ctx* ctx, const float* in, float* out, int64_t n, void* stream);
/* fp32 samples -> 16-bit PCM exactly as save_wave does it: x * 2^15, truncation toward zero through a 32-bit integer,
 * low 16 bits kept (so +1.0 wraps to -32768 like numpy's astype(np.short) on the reference's hosts).  `out` is a
 * device buffer of n int16. */
VF_API int vf_to_pcm16(vf_ctx* ctx, const float* in, int16_t* out, int64_t n, void* stream);

/* Device memory the plan for (batch, n_samples) holds (activations + packed weights). */
VF_API int vf_workspace_bytes(vf_ctx* ctx, int batch, int64_t n_samples, size_t* bytes);

/* Synchronises `stream`, reads and clears the sticky device flags.  VF_OK, VF_EDEVICE or VF_EASSERT. */
VF_API int vf_check_errors(vf_ctx* ctx, void* stream);

/* Options: "vocoder_terms" (1 or 3 fp16 split terms; "unet_terms" accepts only 3), "unify_energy" (1: vf_restore applies
 * amp_to_original_f, tools/utils.py:50-55, as handler() does when meta["unify_energy"] is set),
 * "validate_simt" (1: run every GEMM on the SIMT validation kernel instead of tcgen05 - tests only). */
VF_API int vf_set_option(vf_ctx* ctx, const char* key, int value);
/* Number of kernels this context has launched since creation. */
VF_API int64_t vf_launch_count(vf_ctx* ctx);

/* Stage timing: wraps the stages of subsequent vf_restore calls in CUDA events on the call's stream.
 * vf_stage_times synchronises and returns milliseconds of the last vf_restore: [frontend, unet, vocoder, tail]. */
VF_API int vf_enable_stage_timing(vf_ctx* ctx, int enable);
VF_API int vf_stage_times(vf_ctx* ctx, float ms[4]);

/* Per-launch profile: with op timing enabled, the next vf_restore records a CUDA event before every kernel of
 * the UNet and vocoder launch chains.  vf_op_info(i) synchronises and returns the device time of launch i
 * together with its ALGORITHMIC flops / minimum HBM bytes (the reference op's own counts), the tcgen05 tile
 * (bn, bk, fp16 split terms; 0 for non-GEMM kernels) and a label such as "enc3.b2.conv1".  For roofline reporting only. */
VF_API int vf_enable_op_timing(vf_ctx* ctx, int enable);
VF_API int vf_op_count(vf_ctx* ctx);
VF_API int vf_op_info(vf_ctx* ctx, int i, float* ms, double* flops, double* bytes, int* bn, int* bk, int* terms,
                      char* label, int label_cap);

/* Self-test of one flat-shift GEMM configuration: random fp16 hi/lo planes through the tcgen05 kernel and
 * the SIMT validation kernel; returns max |difference| and max |value|.  Synchronous. */
VF_API int vf_selftest_gemm(vf_ctx* ctx, int n_img, int rows, int cin, int cout, int ntaps, int dilation, int terms,
                     double* max_abs_diff, double* max_abs_ref);

#ifdef __cplusplus
}
#endif
#endif /* B200VF_H_ */
