// fp32 CUDA-core kernel family (HFG_PREC_FP32): exact-fp32 FFMA convolutions on
// channels-last activations.  This is the bit-faithful-class path and the A/B
// baseline every tensor-core kernel is checked against on the device.
//
// Replaces the ATen library calls of the reference forward:
//   F.conv1d            src/iris/hifigan_pretrained.py:67,69,124,140
//   F.conv_transpose1d  src/iris/hifigan_pretrained.py:128
//   F.leaky_relu        :66,68,127,139   (fused: applied while staging x in smem)
//   x = xt + x          :70              (fused: residual add in the epilogue)
//   xs += .. ; xs / 3   :133-137         (fused: accumulate / divide in the epilogue)
//   torch.tanh          :141             (fused into conv_post)
#include <algorithm>

#include "hfg_internal.h"
#include "umma_ptx.cuh"

namespace hfg {

namespace {

constexpr int kMaxDevices = 64;

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : v * kLreluSlope; }

// ---------------------------------------------------------------------------
// Generic channels-last direct convolution (see ConvParams).
//   block: 256 threads = NTX (along N) x NTY (along M); thread tile 8 rows x TC cols
//   smem : x tile [TILE_M + span][CI] (lrelu applied while staging), W tile [taps][CI][TILE_N]
// ---------------------------------------------------------------------------
constexpr int kCI = 8;   // input channels staged per iteration
constexpr int kTT = 8;   // rows per thread

template <int NTX, int TC>
__global__ void __launch_bounds__(256, 2) conv_cl_fp32_kernel(const ConvParams p) {
    constexpr int NTY = 256 / NTX;
    constexpr int TILE_M = NTY * kTT;
    constexpr int TILE_N = NTX * TC;
    constexpr int NQ = TC / 4;
    extern __shared__ __align__(16) float smem[];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int m0 = blockIdx.x * TILE_M;
    const int n0 = blockIdx.y * TILE_N;
    const int b = blockIdx.z;

    const int taps = p.taps;
    const int last_off = p.tap_off0 + (taps - 1) * p.tap_step;
    const int lo = min(p.tap_off0, last_off);
    const int span = max(p.tap_off0, last_off) - lo;
    const int rows_in = TILE_M + span;

    float* in_s = smem;                       // [rows_in][kCI]
    float* w_s = smem + ((rows_in * kCI + 3) & ~3);  // [taps][kCI][TILE_N]

    const float* __restrict__ x = static_cast<const float*>(p.x) + (size_t)b * p.Lin * p.Cin;
    const float* __restrict__ w = static_cast<const float*>(p.w);

    float acc[kTT][TC];
#pragma unroll
    for (int i = 0; i < kTT; ++i)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[i][c] = 0.f;

    for (int ci0 = 0; ci0 < p.Cin; ci0 += kCI) {
        __syncthreads();
        for (int idx = tid; idx < rows_in * 2; idx += 256) {
            const int r = idx >> 1, h = idx & 1;
            const int t = m0 + lo + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < p.Lin) {
                v = __ldg(reinterpret_cast<const float4*>(x + (size_t)t * p.Cin + ci0 + h * 4));
                if (p.pre_lrelu) { v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w); }
            }
            *reinterpret_cast<float4*>(in_s + r * kCI + h * 4) = v;
        }
        for (int idx = tid; idx < taps * kCI * (TILE_N / 4); idx += 256) {
            const int c4 = idx % (TILE_N / 4);
            const int rc = idx / (TILE_N / 4);
            const int j = rc / kCI, ci = rc % kCI;
            const int n = n0 + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < p.Np) v = __ldg(reinterpret_cast<const float4*>(w + ((size_t)j * p.Cin + ci0 + ci) * p.Np + n));
            *reinterpret_cast<float4*>(w_s + rc * TILE_N + c4 * 4) = v;
        }
        __syncthreads();

        for (int j = 0; j < taps; ++j) {
            const int rbase = p.tap_off0 + j * p.tap_step - lo + ty;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 a[kTT];
#pragma unroll
                for (int i = 0; i < kTT; ++i)
                    a[i] = *reinterpret_cast<const float4*>(in_s + (rbase + i * NTY) * kCI + h * 4);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 wv[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        wv[q] = *reinterpret_cast<const float4*>(w_s + ((j * kCI + h * 4 + c) * TILE_N) + q * NTX * 4 + tx * 4);
#pragma unroll
                    for (int i = 0; i < kTT; ++i) {
                        const float av = c == 0 ? a[i].x : c == 1 ? a[i].y : c == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            acc[i][q * 4 + 0] = fmaf(av, wv[q].x, acc[i][q * 4 + 0]);
                            acc[i][q * 4 + 1] = fmaf(av, wv[q].y, acc[i][q * 4 + 1]);
                            acc[i][q * 4 + 2] = fmaf(av, wv[q].z, acc[i][q * 4 + 2]);
                            acc[i][q * 4 + 3] = fmaf(av, wv[q].w, acc[i][q * 4 + 3]);
                        }
                    }
                }
            }
        }
    }

    // Epilogue: scatter (m, n) -> (t_out, co), bias, residual, accumulate, divide, activations.
    float* __restrict__ y = static_cast<float*>(p.y);
    const float* __restrict__ res = static_cast<const float*>(p.res);
#pragma unroll
    for (int i = 0; i < kTT; ++i) {
        const int m = m0 + ty + i * NTY;
        if (m >= p.Mrows) continue;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int n = n0 + q * NTX * 4 + tx * 4;
            if (n >= p.Np) continue;
            int r = 0, co = n;
            if (p.ups_s > 1) { r = n / p.Cout; co = n - r * p.Cout; }
            const int t_out = m * p.ups_s + r - p.ups_p;
            if (t_out < 0 || t_out >= p.Lout) continue;
            const size_t idx = ((size_t)b * p.Lout + t_out) * p.Cout + co;
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + co));
            float4 v = make_float4(acc[i][q * 4 + 0] + bv.x, acc[i][q * 4 + 1] + bv.y,
                                   acc[i][q * 4 + 2] + bv.z, acc[i][q * 4 + 3] + bv.w);
            if (res) {
                const float4 rv = *reinterpret_cast<const float4*>(res + idx);
                v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
            }
            if (p.accumulate) {
                const float4 ov = *reinterpret_cast<const float4*>(y + idx);
                v.x = ov.x + v.x; v.y = ov.y + v.y; v.z = ov.z + v.z; v.w = ov.w + v.w;
            }
            if (p.out_div > 0.f) {
                v.x = __fdiv_rn(v.x, p.out_div); v.y = __fdiv_rn(v.y, p.out_div);
                v.z = __fdiv_rn(v.z, p.out_div); v.w = __fdiv_rn(v.w, p.out_div);
            }
            if (p.post_lrelu) { v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w); }
            if (p.post_tanh) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
            *reinterpret_cast<float4*>(y + idx) = v;
        }
    }
}

template <int NTX, int TC>
cudaError_t launch_conv_variant(const ConvParams& p, cudaStream_t s) {
    constexpr int NTY = 256 / NTX;
    constexpr int TILE_M = NTY * kTT;
    constexpr int TILE_N = NTX * TC;
    const int last_off = p.tap_off0 + (p.taps - 1) * p.tap_step;
    const int span = abs(last_off - p.tap_off0);
    const int rows_in = TILE_M + span;
    const size_t smem = (size_t)(((rows_in * kCI + 3) & ~3) + p.taps * kCI * TILE_N) * sizeof(float);
    static size_t configured[kMaxDevices] = {};  // per-instantiation, per-device high-water mark
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev % kMaxDevices]) {
        cudaError_t e = cudaFuncSetAttribute(conv_cl_fp32_kernel<NTX, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev % kMaxDevices] = smem;
    }
    dim3 grid((p.Mrows + TILE_M - 1) / TILE_M, (p.Np + TILE_N - 1) / TILE_N, p.B);
    conv_cl_fp32_kernel<NTX, TC><<<grid, 256, smem, s>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// conv_post: lrelu -> Conv1d(C -> 1, k) -> tanh.  Bandwidth-bound: smem halo
// staging with float4 loads, G = C/4 lanes per output sample, warp-shuffle
// reduction over the G lanes, results staged in smem for a coalesced store.
// ---------------------------------------------------------------------------
constexpr int kPostTile = 256;

__device__ __forceinline__ float4 load4(const float* x, const float*, size_t i) {
    return __ldg(reinterpret_cast<const float4*>(x + i));
}
// bf16 planes: value = hi (+ lo when the bf16x3 low-order plane is present)
__device__ __forceinline__ float4 load4(const __nv_bfloat16* x, const __nv_bfloat16* x_lo, size_t i) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(x + i));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    float4 v = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
    if (x_lo) {
        const uint2 ul = __ldg(reinterpret_cast<const uint2*>(x_lo + i));
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&ul.x);
        const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(&ul.y);
        v.x += __low2float(c); v.y += __high2float(c); v.z += __low2float(d); v.w += __high2float(d);
    }
    return v;
}

template <typename TIn>
__global__ void __launch_bounds__(256) conv_post_kernel(const TIn* __restrict__ x, const TIn* __restrict__ x_lo,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ wave, int L, int C, int k, int pre_lrelu,
                                                        int apply_tanh) {
    extern __shared__ __align__(16) float smem[];
    const int pad = (k - 1) / 2;
    const int rows = kPostTile + k - 1;
    float* in_s = smem;                 // [rows][C]
    float* w_s = in_s + rows * C;       // [k][C]
    float* out_s = w_s + k * C;         // [kPostTile]
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * kPostTile;
    const int b = blockIdx.y;
    const size_t xoff = (size_t)b * L * C;
    const int c4n = C / 4;
    for (int idx = tid; idx < rows * c4n; idx += 256) {
        const int r = idx / c4n, c4 = idx - r * c4n;
        const int t = t0 - pad + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < L) {
            v = load4(x, x_lo, xoff + (size_t)t * C + c4 * 4);
            if (pre_lrelu) { v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w); }
        }
        *reinterpret_cast<float4*>(in_s + r * C + c4 * 4) = v;
    }
    for (int idx = tid; idx < k * C; idx += 256) w_s[idx] = __ldg(w + idx);
    __syncthreads();
    const int G = c4n;                       // lanes per output (power of two <= 32)
    const int lane_in_g = tid % G;
    const int groups = 256 / G;
    const float bv = __ldg(bias);
    for (int o = tid / G; o < kPostTile; o += groups) {
        float s = 0.f;
        for (int j = 0; j < k; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(in_s + (o + j) * C + lane_in_g * 4);
            const float4 ww = *reinterpret_cast<const float4*>(w_s + j * C + lane_in_g * 4);
            s = fmaf(a.x, ww.x, s); s = fmaf(a.y, ww.y, s); s = fmaf(a.z, ww.z, s); s = fmaf(a.w, ww.w, s);
        }
        for (int d = G / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane_in_g == 0) {
            s += bv;
            out_s[o] = apply_tanh ? tanhf(s) : s;
        }
    }
    __syncthreads();
    const int t = t0 + tid;
    if (t < L) wave[(size_t)b * L + t] = out_s[tid];
}

// [B][R][Cc] -> [B][Cc][R] tiled transpose (used for mel in / taps out).
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* ib = in + (size_t)b * R * Cc;
    float* ob = out + (size_t)b * R * Cc;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Cc) ? ib[(size_t)r * Cc + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Cc) ob[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

// mel [B][C][L] fp32 -> [B][L][Cpad] bf16 hi (+ lo) planes, channels >= C zero.
__global__ void mel_to_cl_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, int C, int L, int Cpad, int apply_lrelu, int f16) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
    const float* ib = in + (size_t)b * C * L;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && l < L) ? ib[(size_t)c * L + l] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        if (l < L && c < Cpad) {
            float v = tile[threadIdx.x][i];
            if (apply_lrelu) v = lrelu(v);
            const size_t o = ((size_t)b * L + l) * Cpad + c;
            if (f16) {   // one fp16 plane in the same 16-bit storage
                reinterpret_cast<unsigned short*>(hi)[o] = (unsigned short)(ptx::pack_f16(v, 0.f) & 0xffffu);
                continue;
            }
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[o] = h;
            if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

__global__ void accum_fp32_kernel(float* __restrict__ xs, const float* __restrict__ r, size_t n4, int first, float div) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(r)[i];
        if (!first) {
            const float4 o = reinterpret_cast<const float4*>(xs)[i];
            v.x = o.x + v.x; v.y = o.y + v.y; v.z = o.z + v.z; v.w = o.w + v.w;
        }
        if (div > 0.f) {
            v.x = __fdiv_rn(v.x, div); v.y = __fdiv_rn(v.y, div); v.z = __fdiv_rn(v.z, div); v.w = __fdiv_rn(v.w, div);
        }
        reinterpret_cast<float4*>(xs)[i] = v;
    }
}

__global__ void act_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                 size_t n, int apply_lrelu) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = x[i];
        if (apply_lrelu) v = lrelu(v);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

__device__ __forceinline__ float inv_lrelu(float p) { return p > 0.f ? p : p * (1.0f / kLreluSlope); }

__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t i8, float (&f)[8], int f16 = 0) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(hi) + i8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    if (f16) {
#pragma unroll
        for (int t = 0; t < 4; ++t) ptx::unpack2<true>(w[t], f[2 * t], f[2 * t + 1]);
        return;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) { f[2 * t] = __uint_as_float(w[t] << 16); f[2 * t + 1] = __uint_as_float(w[t] & 0xffff0000u); }
    if (lo) {
        const uint4 ul = __ldg(reinterpret_cast<const uint4*>(lo) + i8);
        const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { f[2 * t] += __uint_as_float(wl[t] << 16); f[2 * t + 1] += __uint_as_float(wl[t] & 0xffff0000u); }
    }
}

// MRF sum of NK branch planes at 8 consecutive channels: every plane's 16-byte load is issued before the first use
// (the generic loop over a runtime nk serialises one DRAM latency per branch).  Arithmetic: ((x0 + x1) + x2 ...) as in
// hifigan_pretrained.py:133-136, x_j = inverse-lrelu(plane j).
template <int NK, bool LO, bool F16 = false>
__device__ __forceinline__ void mrf_sum8(const MrfArgs& a, size_t i8, float (&v)[8]) {
    using namespace ptx;
    uint4 h[NK], l[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) {
        h[j] = __ldg(reinterpret_cast<const uint4*>(a.hi[j]) + i8);
        if (LO) l[j] = __ldg(reinterpret_cast<const uint4*>(a.lo[j]) + i8);
    }
    f2 acc[4];   // packed pairs (FADD2 / FMUL2): same values as the scalar form, half the instructions
#pragma unroll
    for (int j = 0; j < NK; ++j) {
        const uint32_t w[4] = {h[j].x, h[j].y, h[j].z, h[j].w};
        const uint32_t wl[4] = {LO ? l[j].x : 0u, LO ? l[j].y : 0u, LO ? l[j].z : 0u, LO ? l[j].w : 0u};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            f2 f = f2_from_h2<F16>(w[t]);
            if (LO) f = f2_add(f, f2_from_bf16x2(wl[t]));
            f = f2_inv_lrelu(f);
            acc[t] = j == 0 ? f : f2_add(acc[t], f);
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) f2_unpack(acc[t], v[2 * t], v[2 * t + 1]);
}
// runtime nk -> the specialised sum (nk = 3 is every shipped config); false: caller runs the generic loop
__device__ __forceinline__ bool mrf_sum8_dispatch(const MrfArgs& a, size_t i8, float (&v)[8]) {
    if (a.nk == 3) { if (a.lo[0]) mrf_sum8<3, true>(a, i8, v); else if (a.f16) mrf_sum8<3, false, true>(a, i8, v); else mrf_sum8<3, false>(a, i8, v); return true; }
    if (a.nk == 2) { if (a.lo[0]) mrf_sum8<2, true>(a, i8, v); else if (a.f16) mrf_sum8<2, false, true>(a, i8, v); else mrf_sum8<2, false>(a, i8, v); return true; }
    return false;
}

// [rows][C_tc] planes -> [rows][C] fp32 (C <= C_tc: drops the zero padding channels of narrow stages)
__global__ void planes_to_raw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, float* __restrict__ raw,
                                     size_t rows, int c8_tc, int c8, int f16) {
    const size_t n8 = rows * (size_t)c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / c8;
        const int c = (int)(i - r * c8);
        float f[8];
        load8(hi, lo, r * c8_tc + c, f, f16);
        float4 a = make_float4(inv_lrelu(f[0]), inv_lrelu(f[1]), inv_lrelu(f[2]), inv_lrelu(f[3]));
        float4 b = make_float4(inv_lrelu(f[4]), inv_lrelu(f[5]), inv_lrelu(f[6]), inv_lrelu(f[7]));
        reinterpret_cast<float4*>(raw)[2 * i] = a;
        reinterpret_cast<float4*>(raw)[2 * i + 1] = b;
    }
}

__global__ void mrf_combine_kernel(const MrfArgs a, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float v[8];
        if (!mrf_sum8_dispatch(a, i, v)) {
            load8(a.hi[0], a.lo[0], i, v, a.f16);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = inv_lrelu(v[t]);
            for (int j = 1; j < a.nk; ++j) {
                float f[8];
                load8(a.hi[j], a.lo[j], i, f, a.f16);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = v[t] + inv_lrelu(f[t]);
            }
        }
        // mean: multiply by 1/nk (the tensor-core modes are not bit-faithful to fp32 anyway; conv_post_mrf uses the same form)
        const float rinv = 1.0f / (float)a.nk;
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = v[t] * rinv;
        if (a.out_raw) {
            reinterpret_cast<float4*>(a.out_raw)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(a.out_raw)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (a.out_hi && a.f16) {
            uint4 h;
            h.x = ptx::pack_f16(lrelu(v[0]), lrelu(v[1])); h.y = ptx::pack_f16(lrelu(v[2]), lrelu(v[3]));
            h.z = ptx::pack_f16(lrelu(v[4]), lrelu(v[5])); h.w = ptx::pack_f16(lrelu(v[6]), lrelu(v[7]));
            reinterpret_cast<uint4*>(a.out_hi)[i] = h;
        } else if (a.out_hi) {
            __nv_bfloat162 h[4];
            float l[8];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float x0 = lrelu(v[2 * t]), x1 = lrelu(v[2 * t + 1]);
                h[t] = __floats2bfloat162_rn(x0, x1);
                l[2 * t] = x0 - __low2float(h[t]);
                l[2 * t + 1] = x1 - __high2float(h[t]);
            }
            reinterpret_cast<uint4*>(a.out_hi)[i] = *reinterpret_cast<const uint4*>(h);
            if (a.out_lo) {
                __nv_bfloat162 hl[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) hl[t] = __floats2bfloat162_rn(l[2 * t], l[2 * t + 1]);
                reinterpret_cast<uint4*>(a.out_lo)[i] = *reinterpret_cast<const uint4*>(hl);
            }
        }
    }
}

// conv_post on the nk branch-output planes of the last stage: stage tile = lrelu(mean_j inverse-lrelu(plane_j)) in fp32
// shared memory, then a sliding-window conv.  Row stride C + 2 floats and channel PAIRS interleaved over the four lanes of an
// output group (lane q owns pairs q, q+4, q+8, ...): the 32 lanes of a warp (8 groups x 4 lanes) then read 16 distinct 8-byte
// bank pairs twice -- the 2-wavefront minimum of an LDS.64.  (Stride C + 4 with 8 consecutive channels per lane put the 8 groups
// on the same banks: ncu counted 27 M bank conflicts per launch and 88 % l1tex utilisation, 3.1 of 6.5 TB/s.)
constexpr int kPostMrfTile = 256;
__global__ void __launch_bounds__(kPostMrfTile) conv_post_mrf_kernel(const MrfArgs a, const float* __restrict__ w, const float* __restrict__ bias,
                                                                    float* __restrict__ wave, int L, int C, int k, int apply_tanh) {
    extern __shared__ __align__(16) float smem[];
    const int pad = (k - 1) / 2;
    const int rows = kPostMrfTile + k - 1;
    const int stride = C + 2;
    float* in_s = smem;                    // [rows][stride]
    float* w_s = in_s + rows * stride;     // [k][C]
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * kPostMrfTile;
    const int b = blockIdx.y;
    const int c8n = C / 8;
    const size_t base8 = (size_t)b * L * c8n;
    const float rinv = 1.0f / (float)a.nk;
    for (int idx = tid; idx < rows * c8n; idx += kPostMrfTile) {
        const int r = idx / c8n, c8 = idx - r * c8n;
        const int t = t0 - pad + r;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < L) {
            const size_t i8 = base8 + (size_t)t * c8n + c8;
            const bool summed = a.nk > 1 && mrf_sum8_dispatch(a, i8, v);
            if (!summed) load8(a.hi[0], a.lo[0], i8, v, a.f16);
            if (a.nk > 1) {
                if (!summed) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = inv_lrelu(v[q]);
                    for (int j = 1; j < a.nk; ++j) {
                        float f[8];
                        load8(a.hi[j], a.lo[j], i8, f, a.f16);
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = v[q] + inv_lrelu(f[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float x = lrelu(v[q] * rinv);
                    if (a.f16) { v[q] = __half2float(__float2half_rn(x)); continue; }
                    const float h = __bfloat162float(__float2bfloat16_rn(x));   // what the operand plane(s) would hold
                    v[q] = a.lo[0] ? h + __bfloat162float(__float2bfloat16_rn(x - h)) : h;
                }
            }
        }
        float* dst = in_s + r * stride + c8 * 8;   // rows are 8-byte aligned (stride even)
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float2*>(dst + 2 * i) = make_float2(v[2 * i], v[2 * i + 1]);
    }
    for (int idx = tid; idx < k * C; idx += kPostMrfTile) w_s[idx] = __ldg(w + idx);
    __syncthreads();
    // Thread = (group of 4 consecutive outputs) x (quarter of the channels): a sliding window over k + 3 staged rows feeds
    // 4 accumulators, so one shared-memory row read serves up to 4 outputs; the 4 channel quarters meet in two shuffles.
    const int q = tid & 3, og = tid >> 2;
    const int cq = C / 4;                      // channels per lane (multiple of 2; 8 for C = 32), as interleaved pairs
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < cq; c += 2) {
        const int ch = 2 * q + 4 * c;          // pair index q + 4*(c/2)
        float2 win[4];                         // rows o .. o+3 of the window for this channel pair
#pragma unroll
        for (int i = 0; i < 3; ++i) win[i + 1] = *reinterpret_cast<const float2*>(in_s + (og * 4 + i) * stride + ch);
        for (int j = 0; j < k; ++j) {
#pragma unroll
            for (int i = 0; i < 3; ++i) win[i] = win[i + 1];
            win[3] = *reinterpret_cast<const float2*>(in_s + (og * 4 + j + 3) * stride + ch);
            const float2 ww = *reinterpret_cast<const float2*>(w_s + j * C + ch);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(win[i].y, ww.y, fmaf(win[i].x, ww.x, acc[i]));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
    }
    const int t = t0 + og * 4 + q;             // lane q of the group writes output q: a warp stores 32 consecutive samples
    if (t < L) {
        const float sres = (q == 0 ? acc[0] : q == 1 ? acc[1] : q == 2 ? acc[2] : acc[3]) + __ldg(bias);
        wave[(size_t)b * L + t] = apply_tanh ? tanhf(sres) : sres;
    }
}

cudaError_t launch_transpose(const float* in, float* out, int B, int R, int Cc, cudaStream_t s) {
    dim3 grid((Cc + 31) / 32, (R + 31) / 32, B);
    dim3 block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, out, R, Cc);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv_fp32(const ConvParams& p, cudaStream_t s) {
    if (p.Cin % kCI != 0 || p.Np % 4 != 0 || p.Cout % 4 != 0) return cudaErrorInvalidValue;
    if (p.Np >= 128) return launch_conv_variant<16, 8>(p, s);   // 128 rows x 128 cols
    if (p.Np >= 64) return launch_conv_variant<16, 4>(p, s);    // 128 rows x 64 cols
    return launch_conv_variant<8, 4>(p, s);                     // 256 rows x 32 cols
}

template <typename TIn>
static cudaError_t launch_conv_post_t(const TIn* x, const TIn* x_lo, const float* w, const float* bias, float* wave,
                                      int B, int L, int C, int k, int pre_lrelu, int apply_tanh, cudaStream_t s) {
    const int c4n = C / 4;
    if (C % 4 != 0 || c4n > 32 || (c4n & (c4n - 1)) != 0) return cudaErrorInvalidValue;
    const size_t smem = (size_t)((kPostTile + k - 1) * C + k * C + kPostTile) * sizeof(float);
    static size_t configured[kMaxDevices] = {};  // per instantiation, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && smem > configured[dev % kMaxDevices]) {
        cudaError_t e = cudaFuncSetAttribute(conv_post_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev % kMaxDevices] = smem;
    }
    dim3 grid((L + kPostTile - 1) / kPostTile, B);
    conv_post_kernel<TIn><<<grid, 256, smem, s>>>(x, x_lo, w, bias, wave, L, C, k, pre_lrelu, apply_tanh);
    return cudaGetLastError();
}

cudaError_t launch_conv_post_fp32(const float* x, const float* w, const float* bias, float* wave,
                                  int B, int L, int C, int k, int pre_lrelu, int apply_tanh, cudaStream_t s) {
    return launch_conv_post_t<float>(x, nullptr, w, bias, wave, B, L, C, k, pre_lrelu, apply_tanh, s);
}
// Tensor-core modes hand conv_post the already-activated bf16 plane(s).
cudaError_t launch_conv_post_bf16(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, const float* w, const float* bias,
                                  float* wave, int B, int L, int C, int k, int apply_tanh, cudaStream_t s) {
    return launch_conv_post_t<__nv_bfloat16>(x_hi, x_lo, w, bias, wave, B, L, C, k, 0, apply_tanh, s);
}

cudaError_t launch_transpose_cf_to_cl(const float* in, float* out, int B, int C, int L, cudaStream_t s) {
    return launch_transpose(in, out, B, C, L, s);
}
cudaError_t launch_transpose_cl_to_cf(const float* in, float* out, int B, int C, int L, cudaStream_t s) {
    return launch_transpose(in, out, B, L, C, s);
}

cudaError_t launch_mel_to_cl_bf16(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int L, int Cpad,
                                  int apply_lrelu, int f16, cudaStream_t s) {
    dim3 grid((L + 31) / 32, (Cpad + 31) / 32, B);
    dim3 block(32, 8);
    mel_to_cl_bf16_kernel<<<grid, block, 0, s>>>(in, hi, f16 ? nullptr : lo, C, L, Cpad, apply_lrelu, f16);
    return cudaGetLastError();
}

cudaError_t launch_accum_fp32(float* xs, const float* r, size_t n, int first, float div, cudaStream_t s) {
    if (n % 4 != 0) return cudaErrorInvalidValue;
    const size_t n4 = n / 4;
    const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
    accum_fp32_kernel<<<blocks, 256, 0, s>>>(xs, r, n4, first, div);
    return cudaGetLastError();
}

cudaError_t launch_planes_to_raw(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* raw, size_t rows, int C_tc, int C,
                                 int f16, cudaStream_t s) {
    if (C % 8 != 0 || C_tc % 8 != 0 || C > C_tc) return cudaErrorInvalidValue;
    const size_t n8 = rows * (size_t)(C / 8);
    const int blocks = (int)std::min<size_t>((n8 + 255) / 256, 148 * 16);
    planes_to_raw_kernel<<<blocks, 256, 0, s>>>(hi, lo, raw, rows, C_tc / 8, C / 8, f16);
    return cudaGetLastError();
}

// Ragged batches (hfg_forward_ragged): item b is lens[b] mel frames long, i.e. lens[b] * mul rows of this [B][L][row_elems]
// 16-bit plane (pair).  The first H rows behind an item's own end are set to zero, so that the next layer's taps read there what
// they read behind the end of a dense batch (TMA's out-of-bounds zero fill == the reference's zero padding,
// hifigan_pretrained.py:49-59,92-94); rows further out never reach a row inside the item (H covers the widest tap span).
__global__ void zero_tail_rows_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, const int32_t* __restrict__ lens,
                                      int mul, int L, int row_elems, int H) {
    const int b = blockIdx.y;
    const long long r0 = (long long)lens[b] * mul;
    if (r0 >= L) return;
    const long long r1 = r0 + H < L ? r0 + H : L;
    const size_t first = ((size_t)b * L + (size_t)r0) * row_elems;          // element offsets; row_elems % 8 == 0: 16-byte aligned
    const size_t n8 = (size_t)(r1 - r0) * row_elems / 8;
    uint4* ph = reinterpret_cast<uint4*>(hi + first);
    uint4* pl = lo ? reinterpret_cast<uint4*>(lo + first) : nullptr;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        ph[i] = z;
        if (pl) pl[i] = z;
    }
}

cudaError_t launch_zero_tail_rows(__nv_bfloat16* hi, __nv_bfloat16* lo, const int32_t* lens, int mul, int B, int L, int row_elems,
                                  int H, cudaStream_t s) {
    if (!hi || !lens || row_elems % 8 != 0 || mul < 1 || H < 1 || B < 1 || B > 65535) return cudaErrorInvalidValue;
    const size_t n8 = (size_t)H * row_elems / 8;
    dim3 grid((unsigned)std::max<size_t>(1, std::min<size_t>((n8 + 255) / 256, 8)), (unsigned)B);
    zero_tail_rows_kernel<<<grid, 256, 0, s>>>(hi, lo, lens, mul, L, row_elems, H);
    return cudaGetLastError();
}

cudaError_t launch_mrf_combine(const MrfArgs& a, size_t n, cudaStream_t s) {
    if (n % 8 != 0 || a.nk < 1 || a.nk > HFG_MAX_KERNELS) return cudaErrorInvalidValue;
    const size_t n8 = n / 8;
    const int blocks = (int)std::min<size_t>((n8 + 255) / 256, 148 * 16);
    mrf_combine_kernel<<<blocks, 256, 0, s>>>(a, n8);
    return cudaGetLastError();
}

cudaError_t launch_conv_post_mrf(const MrfArgs& a, const float* w, const float* bias, float* wave, int B, int L, int C, int k,
                                 int apply_tanh, cudaStream_t s) {
    if (C % 8 != 0 || a.nk < 1 || a.nk > HFG_MAX_KERNELS) return cudaErrorInvalidValue;
    const size_t smem = (size_t)((kPostMrfTile + k - 1) * (C + 2) + k * C) * sizeof(float);
    static size_t configured[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && smem > configured[dev % kMaxDevices]) {
        cudaError_t e = cudaFuncSetAttribute(conv_post_mrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev % kMaxDevices] = smem;
    }
    dim3 grid((L + kPostMrfTile - 1) / kPostMrfTile, B);
    conv_post_mrf_kernel<<<grid, kPostMrfTile, smem, s>>>(a, w, bias, wave, L, C, k, apply_tanh);
    return cudaGetLastError();
}

cudaError_t launch_act_split(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, int apply_lrelu, cudaStream_t s) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    act_split_kernel<<<blocks, 256, 0, s>>>(x, hi, lo, n, apply_lrelu);
    return cudaGetLastError();
}

}  // namespace hfg
