// Internal declarations shared by the engine and the kernel translation units.
// Nothing here is part of the ABI (see include/hfg.h).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

#include "../../include/hfg.h"

namespace hfg {

constexpr float kLreluSlope = 0.1f;  // hifigan_pretrained.py:66,68,127,139

// One convolution of the generator expressed as the generic channels-last op
//
//   Y[b][m][n] = sum_j sum_ci act(X[b][m + off_j][ci]) * W[j][ci][n],   off_j = tap_off0 + j*tap_step
//
// followed by the scatter  (m, n) -> (t_out = m*ups_s + n/Cout - ups_p, co = n%Cout)
// and the fused epilogue (see each kernel family).
//
// Conv1d (k, dilation d, pad p):   taps=k, off_j = j*d - p, Np=Cout, ups_s=1.
// ConvTranspose1d (k, stride s, pad p), polyphase: taps=ceil(k/s), off_n = -n,
//   Np = s*Cout with W'[n][ci][r*Cout+co] = w[ci][co][r+n*s]; M rows = Lin+taps-1.
//   Because the output is channels-last, (m, n) lands at flat offset
//   (m*s - p)*Cout + n: the scatter is a dense row of s*Cout values.
struct ConvGeom {
    int B, Lin, Cin;
    int Lout, Cout;
    int Mrows;  // GEMM rows per batch item
    int Np;     // GEMM N
    int taps, tap_off0, tap_step;
    int ups_s, ups_p;
};

// ---- fp32 CUDA-core family (kernels_fp32.cu) -------------------------------
// epilogue: v = acc + bias (+ res) ; accumulate: v = y_old + v ; out_div: v /= d ;
//           post_lrelu ; post_tanh ; y = v.      x, y, res are fp32 [B][L][C].
struct ConvParams {
    const void* x;
    const void* w;      // fp32 [taps][Cin][Np]
    const float* bias;  // [Cout]
    void* y;
    const void* res;    // residual, same shape as y, or nullptr
    int B, Lin, Cin;
    int Lout, Cout;
    int Mrows;
    int Np;
    int taps, tap_off0, tap_step;
    int ups_s, ups_p;
    int pre_lrelu;      // apply lrelu to x while staging
    int accumulate;     // v = old y + v
    float out_div;      // v /= out_div when > 0
    int post_lrelu;
    int post_tanh;
};

cudaError_t launch_conv_fp32(const ConvParams& p, cudaStream_t s);
// (lrelu ->) Conv1d(C->1, k, pad (k-1)/2) (-> tanh) ; x [B][L][C] -> wave [B][L]
cudaError_t launch_conv_post_fp32(const float* x, const float* w /*[k][C]*/, const float* bias, float* wave,
                                  int B, int L, int C, int k, int pre_lrelu, int apply_tanh, cudaStream_t s);
cudaError_t launch_conv_post_bf16(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, const float* w, const float* bias,
                                  float* wave, int B, int L, int C, int k, int apply_tanh, cudaStream_t s);
// [B][C][L] <-> [B][L][C] fp32
cudaError_t launch_transpose_cf_to_cl(const float* in, float* out, int B, int C, int L, cudaStream_t s);
cudaError_t launch_transpose_cl_to_cf(const float* in, float* out, int B, int C, int L, cudaStream_t s);
// [B][C][L] fp32 -> channels-last 16-bit planes [B][L][Cpad] (zero padded channels), hi (+ lo) bf16 planes or one fp16 plane
// (f16; the pointer type stays __nv_bfloat16* for every 16-bit plane), optional lrelu
cudaError_t launch_mel_to_cl_bf16(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int L, int Cpad,
                                  int apply_lrelu, int f16, cudaStream_t s);
// xs = (first ? r : xs + r) ; if div > 0: xs /= div   (MRF sum, hifigan_pretrained.py:133-137)
cudaError_t launch_accum_fp32(float* xs, const float* r, size_t n, int first, float div, cudaStream_t s);
// act planes from an fp32 tensor: hi = bf16(lrelu(x)), lo = bf16(lrelu(x) - hi)
cudaError_t launch_act_split(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, int apply_lrelu, cudaStream_t s);

// ---- tcgen05 tensor-core family (kernels_umma.cu) --------------------------
// A operand: activated bf16 planes [B][Lin][CinPad] read by TMA (OOB rows -> 0 == conv zero padding)
// B operand: bf16 weights [taps*Np][CinPad] (K-major) read by TMA
// epilogue:  v = acc + bias (+ res) ; y_raw = v ; if xs_read: v = xs + v ; if out_div > 0: v /= out_div ;
//            if xs_write: xs = v ; y_act(_lo) = split(lrelu(v))
struct UmmaConvParams {
    ConvGeom g;
    int cin_pad;          // channels per row of the A tensor (multiple of kc)
    int kc;               // K elements per chunk: 64 (128-byte rows) or 32 (64-byte rows)
    int npass;            // 1: bf16 / fp16 ; 3: bf16x3 (hi*hi + lo*hi + hi*lo)
    int f16;              // single-plane mode with fp16 operand planes and weights (HFG_PREC_FP16) instead of bf16
    const float* bias;
    const float* res;     // fp32 [B][Lout][Cout] or nullptr
    const __nv_bfloat16* res_hi;   // residual given as ACTIVATED planes (x recovered by inverting lrelu) or nullptr
    const __nv_bfloat16* res_lo;
    // MRF sum folded into the last convs2 of a branch (hifigan_pretrained.py:133-137): v = (acc + bias + x + s_prev) * out_scale,
    // s_prev = inverse-lrelu of these planes (the running sum of the previous branches' outputs); needs res_hi
    const __nv_bfloat16* mrf_hi;
    const __nv_bfloat16* mrf_lo;
    float out_scale;      // used with mrf_hi only: 1, or 1 / num_kernels on the last branch
    float* y_raw;         // fp32 or nullptr
    __nv_bfloat16* y_act; // activated (lrelu) output plane or nullptr
    __nv_bfloat16* y_act_lo;
    float* xs;            // MRF accumulator or nullptr
    int xs_read, xs_write;
    float out_div;
    int reverse;          // conv_umma2: walk tiles last-to-first (L2 reuse between consecutive kernels)
    const int32_t* lens;  // conv_umma2, ragged batch (device, [B] mel frames of len_T per item): rows behind an item's end are written as zeros
    int len_T;
    int len_skip;         // tiles that start this many rows behind an item's end are skipped (>= the widest tap span of any layer)
    int a_per_tap;        // debug/A-B: reload the A tile per tap instead of shifting descriptors
};

struct UmmaLaunch {
    UmmaConvParams p;
    alignas(64) CUtensorMap map_a_hi, map_a_lo, map_w_hi, map_w_lo;
    int n_tile, mt;       // GEMM N per CTA, 128-row subtiles per CTA
    int rows_a;           // rows of the staged A tile (= a_pieces * a_box_rows)
    int a_box_rows, a_pieces;
    int lo;               // smallest tap offset
    int stages_a, stages_w;
    uint32_t a_plane_bytes, w_plane_bytes, tmem_cols;
    dim3 grid;
    size_t smem;
};

// Fills tensor maps + launch geometry.  x planes are [B][Lin][cin_pad] bf16, weights [taps*Np][cin_pad] bf16.
int plan_conv_umma(UmmaLaunch* L, const UmmaConvParams& p, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                   const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo);
cudaError_t launch_conv_umma(const UmmaLaunch& L, cudaStream_t s);

// ---- persistent pipelined variant for plain Conv1d on planes (kernels_umma2.cu) ----
// in: activated planes (A operand), optional residual planes; out: activated planes only.
struct Umma2Launch {
    struct Impl;
    std::shared_ptr<Impl> impl;
    int mt = 0, n_a = 0, n_w = 0, n_e = 0, w_resident = 0, grid = 0;
    size_t smem = 0;
};
bool umma2_supported(const UmmaConvParams& p);
int plan_conv_umma2(Umma2Launch* L, const UmmaConvParams& p, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                    const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int sm_count);
cudaError_t launch_conv_umma2(const Umma2Launch& L, cudaStream_t s);

// ---- fused ResBlock pair (convs1[m] -> lrelu -> convs2[m] -> + x) for C = 32 / 64 on planes (kernels_pair.cu) ----
// hifigan_pretrained.py:66-70 as one kernel: the intermediate stays in shared memory, the residual comes from the staged x tile.
struct PairParams {
    int B, L, C;          // planes are [B][L][C] bf16, C = 32 or 64 (already channel-padded)
    int k1, k2, d;        // c1: k1 taps, dilation d ; c2: k2 taps, dilation 1 ; both 'same'-padded (k1 = k2 for a reference ResBlock;
                          // they differ for the time-folded narrow stages, engine.cu)
    int npass;            // 1: bf16 / fp16 ; 3: bf16x3
    int f16;              // fp16 operand planes (npass == 1 only)
    int reverse;
    const __nv_bfloat16 *mrf_hi, *mrf_lo;   // running MRF sum added in epilogue 2 (see UmmaConvParams), or nullptr
    float out_scale;
    const int32_t* lens;  // ragged batch (device, [B]): item b ends at row lens[b] * len_mul -- c2 sees zeros from there on; or nullptr
    int len_mul;
    int len_skip;         // tiles that start this many rows behind an item's end are skipped (>= the widest tap span of any layer)
    const float* bias1;
    const float* bias2;
    const __nv_bfloat16 *x_hi, *x_lo;
    const __nv_bfloat16 *w1_hi, *w1_lo, *w2_hi, *w2_lo;   // [k*C][C] K-major
    __nv_bfloat16 *y_hi, *y_lo;
};
struct PairLaunch {
    struct Impl;
    std::shared_ptr<Impl> impl;
    int mt = 0, n_x = 0, n_t = 0, n_o = 0, grid = 0;
    size_t smem = 0;
};
bool pair_supported(const PairParams& p);
int plan_conv_pair(PairLaunch* L, const PairParams& p, int sm_count);
cudaError_t launch_conv_pair(const PairLaunch& L, cudaStream_t s);

// raw fp32 [rows][C] = inverse-lrelu(hi (+ lo)) of planes [rows][C_tc], C <= C_tc   (taps; drops padding channels)
cudaError_t launch_planes_to_raw(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* raw, size_t rows, int C_tc, int C,
                                 int f16, cudaStream_t s);
// MRF combine (hifigan_pretrained.py:133-137) on planes: v = ((x0 + x1) + x2 ...) / nk with x_j = inverse-lrelu(plane j);
// writes planes of lrelu(v) and/or raw fp32 v.
struct MrfArgs {
    const __nv_bfloat16* hi[HFG_MAX_KERNELS];
    const __nv_bfloat16* lo[HFG_MAX_KERNELS];   // all nullptr in single-plane mode
    int nk;
    int f16;              // planes are fp16 (single-plane mode)
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    float* out_raw;
};
cudaError_t launch_mrf_combine(const MrfArgs& a, size_t n, cudaStream_t s);
// MRF combine of the LAST stage fused into conv_post: wave = tanh(Conv1d(lrelu(mean_j x_j)))  (:133-141); the mean is rounded
// through the operand-plane format exactly as launch_mrf_combine would store it, so both plans give identical bits.
cudaError_t launch_conv_post_mrf(const MrfArgs& a, const float* w /*[k][C]*/, const float* bias, float* wave, int B, int L, int C,
                                 int k, int apply_tanh, cudaStream_t s);

// Ragged batches: zero the H rows behind each item's own end (lens[b] * mul rows) of a [B][L][row_elems] plane (pair)
cudaError_t launch_zero_tail_rows(__nv_bfloat16* hi, __nv_bfloat16* lo, const int32_t* lens, int mul, int B, int L, int row_elems,
                                  int H, cudaStream_t s);

void set_error(const std::string& msg);

}  // namespace hfg
