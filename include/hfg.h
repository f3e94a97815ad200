/* hfg.h -- C ABI of the B200-native HiFiGAN generator engine (libhfg_b200.so).
 *
 * This is the drop-in boundary for iris-tts's vocoder hot path.  The reference
 * has no FFI for this path: its "plugin API" is Python function signatures that
 * end in torch library calls.  Each entry point below names the reference code
 * it replaces (paths relative to the reference tree):
 *
 *   hfg_create            HiFiGANModel.__init__ graph construction
 *                         src/iris/hifigan_pretrained.py:77-121
 *                         (and the Keras twin, src/iris/vocoder.py:59-101)
 *   hfg_set_weight_norm   nn.utils.weight_norm re-parametrisation, folded ONCE
 *                         here instead of on every forward
 *                         src/iris/hifigan_pretrained.py:49,55,92,100,119
 *   hfg_set_weight        load_state_dict of already-folded tensors
 *                         src/iris/hifigan_pretrained.py:190 ; Keras
 *                         load_weights, src/iris/vocoder.py:167-170
 *   hfg_finalize          .eval().to(device)   src/iris/hifigan_pretrained.py:202-204
 *   hfg_forward           HiFiGANModel.forward src/iris/hifigan_pretrained.py:123-143
 *                         incl. the H2D/D2H copies of HiFiGANGenerator.__call__
 *                         (:228, :235) when host pointers are passed
 *   hfg_run_layer         one F.conv1d / F.conv_transpose1d call of that forward
 *                         (:67, :69, :124, :128, :140) -- for per-layer parity
 *   hfg_get_tap           forward hooks on submodules (test-only visibility)
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success or a negative hfg_status; the message for the last failure on the
 * calling thread is hfg_last_error().  Nothing throws across the ABI.  A handle
 * is not thread-safe (one CUDA stream per handle); use one handle per thread or
 * per device.  There is no CPU fallback: without a CUDA device hfg_create fails.
 */
#ifndef HFG_H_
#define HFG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFG_ABI_VERSION 2
#define HFG_MAX_UPSAMPLES 8
#define HFG_MAX_KERNELS 8
#define HFG_MAX_DILATIONS 8

typedef enum hfg_status {
    HFG_OK = 0,
    HFG_ERR_INVALID = -1,   /* bad argument / shape / name            */
    HFG_ERR_CUDA = -2,      /* a CUDA runtime or driver call failed   */
    HFG_ERR_STATE = -3,     /* call order (e.g. forward before finalize) */
    HFG_ERR_NOMEM = -4,     /* device or host allocation failed       */
    HFG_ERR_UNSUPPORTED = -5
} hfg_status;

/* Arithmetic the convolutions run in.  Outputs are always fp32. */
typedef enum hfg_precision {
    HFG_PREC_FP32 = 0,    /* fp32 FFMA on CUDA cores, fp32 activations (bit-faithful class)  */
    HFG_PREC_BF16 = 1,    /* tcgen05 kind::f16, bf16 operands, fp32 TMEM accumulation        */
    HFG_PREC_BF16X3 = 2,  /* tcgen05, operands split hi+lo bf16, 3 MMAs: fp32-class accuracy */
    HFG_PREC_FP16 = 3     /* tcgen05 kind::f16, fp16 operands (11-bit significand: TF32-class accuracy at the bf16 mode's speed
                             and bytes), fp32 TMEM accumulation; activations saturate at +-65504                             */
} hfg_precision;

/* hfg_forward flags */
#define HFG_MEL_ON_DEVICE 1u    /* mel is a device pointer (default: host)   */
#define HFG_WAVE_ON_DEVICE 2u   /* wave is a device pointer (default: host)  */
#define HFG_KEEP_TAPS 4u        /* keep copies of intermediate activations for hfg_get_tap */
#define HFG_NO_SYNC 8u          /* enqueue and return; call hfg_sync before reading the result.  Host pointers are allowed if
                                   they are PAGE-LOCKED and stay valid until hfg_sync: the waveform then reaches the host on a
                                   second stream, so the D2H copy of forward i overlaps the kernels of forward i+1 */

/* The six constructor arguments of the reference generator
 * (hifigan_pretrained.py:77-85 / vocoder.py:59-68). */
typedef struct hfg_config {
    int32_t in_channels;                 /* 80 */
    int32_t upsample_initial_channel;    /* 512 (V1) */
    int32_t num_upsamples;
    int32_t upsample_rates[HFG_MAX_UPSAMPLES];
    int32_t upsample_kernel_sizes[HFG_MAX_UPSAMPLES];
    int32_t num_kernels;
    int32_t resblock_kernel_sizes[HFG_MAX_KERNELS];
    int32_t num_dilations[HFG_MAX_KERNELS];
    int32_t resblock_dilations[HFG_MAX_KERNELS][HFG_MAX_DILATIONS];
} hfg_config;

typedef struct hfg_engine hfg_engine;

int hfg_abi_version(void);
const char* hfg_last_error(void);

/* Number of CUDA devices visible (0 if none / no driver). */
int hfg_device_count(void);

int hfg_create(const hfg_config* cfg, int device, hfg_engine** out);
void hfg_destroy(hfg_engine* e);

/* Layer names are the reference's state-dict prefixes: "conv_pre", "ups.<i>",
 * "resblocks.<n>.convs1.<m>", "resblocks.<n>.convs2.<m>", "conv_post".
 * Host fp32 pointers in torch layout: Conv1d weight [C_out][C_in][k],
 * ConvTranspose1d weight [C_in][C_out][k]; g has dim-0 entries; bias [C_out]. */
int hfg_set_weight_norm(hfg_engine* e, const char* layer, const float* g, const float* v, const float* bias);
int hfg_set_weight(hfg_engine* e, const char* layer, const float* w, const float* bias);
/* Shape query so callers can validate checkpoints: dims = {d0, d1, k}. */
int hfg_layer_shape(const hfg_engine* e, const char* layer, int32_t dims[3], int32_t* is_transposed);
int hfg_num_layers(const hfg_engine* e);
int hfg_layer_name(const hfg_engine* e, int index, char* buf, size_t buflen);

/* Repack + upload; every layer must have been set. */
int hfg_finalize(hfg_engine* e);

/* mel [B][in_channels][T] fp32 -> wave [B][T*hop] fp32.
 * Limits (HFG_ERR_UNSUPPORTED beyond them, nothing launched): B <= 65535; a stage's per-item plane < 2^31 elements, which for
 * the V1 generator means T < 262144 frames per call -- longer input is synthesized in chunks (receptive field: 15 frames). */
int hfg_forward(hfg_engine* e, const float* mel, int32_t B, int32_t T, float* wave,
                int32_t precision, uint32_t flags);
/* Ragged batch: item b is lengths[b] mel frames long (1 <= lengths[b] <= T, host array of B entries); whatever mel holds at or
 * behind frame lengths[b] of item b is ignored.  The first lengths[b]*hop samples of wave row b equal, bit for bit, what
 * hfg_forward returns for that item alone (B = 1, T = lengths[b]) -- every layer zero-pads at the item's OWN end, as the
 * reference does for a dense batch of that length (src/iris/hifigan_pretrained.py:49-59, 92-94; its __call__ :221-242 takes one
 * dense [B, 80, T] array and has no lengths, so ragged input costs it one forward per distinct length).  The rest of row b is
 * unspecified.  Every precision; not with HFG_KEEP_TAPS (HFG_ERR_INVALID). */
int hfg_forward_ragged(hfg_engine* e, const float* mel, int32_t B, int32_t T, const int32_t* lengths, float* wave,
                       int32_t precision, uint32_t flags);
int hfg_sync(hfg_engine* e);
int32_t hfg_hop(const hfg_engine* e);

/* Device bytes hfg_forward needs for (B, T) in the given precision. */
size_t hfg_workspace_bytes(const hfg_engine* e, int32_t B, int32_t T, int32_t precision);

/* The engine's CUDA stream (a cudaStream_t) for event timing by the caller. */
void* hfg_stream(hfg_engine* e);
/* Kernels launched by this handle since creation (bench "gpu_launches"). */
uint64_t hfg_launch_count(const hfg_engine* e);
/* Plans whose launches were captured into a CUDA graph / whose capture failed (those keep launching kernel by kernel). */
int hfg_graph_stats(const hfg_engine* e, int32_t* captured, int32_t* failed);

/* Per-launch device timing for roofline reports (bench.py): when enabled, every kernel launch
 * of the following hfg_forward calls is bracketed by CUDA events on the engine's stream.  Records
 * accumulate over forwards until hfg_profile_enable(e, 1) is called again; record i describes launch i: the reference layer it computes
 * (state-dict prefix, or the kernel's own name for layout helpers), the kernel family, its device
 * time, and the layer's algorithmic work F_l = 2*Cin*Cout*k*L*B flop and
 * Q_l = activations in + out + weights bytes (no reference counterpart: the reference has no timers). */
int hfg_profile_enable(hfg_engine* e, int on);
int hfg_profile_count(const hfg_engine* e);
int hfg_profile_get(hfg_engine* e, int i, char* layer, size_t layer_len, char* kernel, size_t kernel_len,
                    float* ms, double* flops, double* bytes);

/* One conv layer in isolation, host pointers, reference layouts:
 * x [B][C_in][L] -> y [B][C_out][L_out]; pre_lrelu applies leaky_relu(.,0.1) to x first. */
int hfg_run_layer(hfg_engine* e, const char* layer, const float* x, int32_t B, int32_t L,
                  int32_t pre_lrelu, float* y, int32_t precision);

/* One ResBlock step in isolation (src/iris/hifigan_pretrained.py:66-70), host pointers, reference layout:
 * y = x + convs2[m](lrelu(convs1[m](lrelu(x)))) for ResBlock n; x, y [B][C][L].  Tensor-core precisions only.
 * Runs the fused pair kernel where the forward's plan would (C <= 64), the two single-conv launches otherwise;
 * *fused (optional) reports which. */
int hfg_run_pair(hfg_engine* e, int32_t resblock, int32_t m, const float* x, int32_t B, int32_t L, float* y,
                 int32_t precision, int32_t* fused);
/* The same step as the LAST one of a ResBlock whose branch sum is folded into its epilogue (:133-137):
 * y = (x + convs2[m](lrelu(convs1[m](lrelu(x)))) + mrf_sum) * out_scale; mrf_sum [B][C][L] is the running sum of the
 * previous branches' outputs (NULL: plain hfg_run_pair), out_scale = 1 or 1/num_kernels. */
int hfg_run_pair_mrf(hfg_engine* e, int32_t resblock, int32_t m, const float* x, const float* mrf_sum, float out_scale,
                     int32_t B, int32_t L, float* y, int32_t precision, int32_t* fused);

/* After hfg_forward(..., HFG_KEEP_TAPS): copy an intermediate activation to
 * host as fp32 [B][C][L] (reference layout).  Names: "conv_pre", "ups.<i>",
 * "resblocks.<3i>" (first ResBlock of each stage), "stage.<i>", "conv_post".
 * Pass out == NULL to query the element count via *n. */
int hfg_get_tap(hfg_engine* e, const char* name, float* out, size_t* n);

/* ---- log-mel front-end: the step immediately before the vocoder -------------------------------------
 *
 *   hfg_logmel_create / _forward   compute_mel_spectrogram  src/iris/data.py:25-67, i.e.
 *                                  librosa.feature.melspectrogram(power=1.0) (data.py:51-62; librosa 0.11.0, uv.lock:872:
 *                                  periodic Hann, center=True with zero padding, Slaney filterbank) and
 *                                  np.log(np.clip(mel, 1e-5, None)) (data.py:65)
 * audio [B][N] fp32 -> mel [B][n_mels][T], T = 1 + N / hop_length: the [B, 80, T] array hfg_forward consumes. */
typedef struct hfg_logmel_config {
    int32_t sample_rate;   /* 22050 */
    int32_t n_fft;         /* 1024 (a power of two in [64, 4096]) */
    int32_t hop_length;    /* 256 */
    int32_t win_length;    /* 1024 (<= n_fft) */
    int32_t n_mels;        /* 80 */
    float fmin;            /* 0 */
    float fmax;            /* 8000; <= 0 means sample_rate / 2 */
    float clip;            /* 1e-5 */
    int32_t log_output;    /* 1: natural log of the clipped magnitude mel (data.py:65); 0: the magnitude mel itself */
} hfg_logmel_config;

typedef struct hfg_logmel hfg_logmel;

#define HFG_LOGMEL_AUDIO_ON_DEVICE 1u
#define HFG_LOGMEL_OUT_ON_DEVICE 2u

int hfg_logmel_create(const hfg_logmel_config* cfg, int device, hfg_logmel** out);
void hfg_logmel_destroy(hfg_logmel* h);
int32_t hfg_logmel_frames(const hfg_logmel* h, int32_t n_samples);
int hfg_logmel_forward(hfg_logmel* h, const float* audio, int32_t B, int32_t N, float* mel, uint32_t flags);

/* Griffin-Lim, the reference's alternative vocoder: scripts/synthesize.py:193  librosa.griffinlim(S, n_iter=60, hop_length, win_length)
 * (librosa 0.11.0: momentum 0.99, random initial phases, center=True) with the STFT geometry of the handle's configuration.
 * mag [B][1 + n_fft/2][T] linear magnitudes (librosa layout), angles0 [B][T][1 + n_fft/2][2] initial unit phasors (cos, sin) -- drawn
 * by the caller, so a result can be reproduced -- -> audio [B][hop_length * (T - 1)].  Host pointers. */
int hfg_griffin_lim(hfg_logmel* h, const float* mag, const float* angles0, int32_t B, int32_t T, int32_t n_iter, float momentum,
                    float* audio);

/* The step before it, scripts/synthesize.py:180-192: linear magnitudes from a log-mel,
 *   mag[b][k][t] = max(0, sum_m proj[k][m] * exp(min(max(logmel[b][m][t], lo), hi)))
 * with proj [n_bins][n_mels] supplied by the caller (the reference solves librosa.feature.inverse.mel_to_stft's non-negative
 * least squares there; the host mirror passes the pseudo-inverse of the Slaney filterbank).  logmel [B][n_mels][T] -> mag [B][n_bins][T].
 * Host pointers; n_bins and n_mels are the handle's 1 + n_fft/2 and any positive mel count. */
int hfg_mel_to_linear(hfg_logmel* h, const float* proj, const float* logmel, int32_t B, int32_t n_mels, int32_t T, float lo, float hi,
                      float* mag);

#ifdef __cplusplus
}
#endif
#endif /* HFG_H_ */
