// Log-mel front-end on the GPU: waveform [B][N] -> log-mel [B][n_mels][T], the array the vocoder hot path consumes.
//
// Replaces (reference tree) src/iris/data.py:25-67 compute_mel_spectrogram, i.e.
//   librosa.feature.melspectrogram(y, sr, n_fft, hop_length, win_length, n_mels, fmin, fmax, power=1.0)   data.py:51-62
//   np.log(np.clip(mel, 1e-5, None))                                                                     data.py:65
// with librosa 0.11.0's defaults (uv.lock:872): periodic Hann window centred in the n_fft frame, center=True with
// n_fft/2 ZEROS on both sides (pad_mode='constant'), T = 1 + N / hop frames, |rfft|, Slaney mel filterbank with
// area normalisation (oracle/logmel_oracle.py restates each function and is what the -m gpu tests compare against).
//
// One CTA transforms kFramesPerCta consecutive frames of one utterance, two at a time (two real frames ride one complex
// FFT): windowed frames -> shared memory (bit-reversed), radix-2 FFT in shared memory (fp32, twiddles from a table computed
// in fp64 on the host), the two spectra separated by conjugate symmetry, magnitudes of the n_fft/2+1 bins, sparse
// triangular mel filters (a warp per band, lanes over its contiguous run of bins), log(max(., clip)), and the frames of a
// CTA leave as runs of kFramesPerCta consecutive floats per mel band.  The op is 0.3 % of the vocoder's FLOPs and is
// bound by reading 4 bytes per sample once (frames overlap 4x: the re-reads hit L2) -- CUDA-core fp32 is the right tool.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "hfg_internal.h"

namespace hfg {
namespace {

constexpr int kFramesPerCta = 8;
constexpr int kMelThreads = 256;

struct MelArgs {
    const float* audio;      // [B][N]
    float* out;              // [B][n_mels][T]
    const float* window;     // [n_fft]  (periodic Hann of win_length, zero-padded centrally)
    const float2* twiddle;   // [n_fft/2]  exp(-2 pi i k / n_fft)
    const int* band_start;   // [n_mels]  first bin with a non-zero weight
    const int* band_count;   // [n_mels]
    const int* band_off;     // [n_mels]  offset of the band's weights in `weights`
    const float* weights;
    int n_weights;
    int N, T, n_fft, log2n, hop, n_mels;
    float clip;
    int log_output;
};

// Shared-memory FFT (forward transform, decimation in time on bit-reversed input), fp32.
//  * Data live at PADDED positions fft_pad(i) = i + i / 32: the bit-reversed scatter of the load phase (a warp's 32 stores differ
//    only in address bits >= 5, one bank without padding) and the strided butterflies of the early stages are then spread
//    over the banks (measured: the unpadded radix-2 version spent ~9/10 of the kernel in bank conflicts).
//  * Two radix-2 stages per pass, in registers: a thread owns the 4 points {b, b+h, b+2h, b+3h}, so n_fft = 1024 takes 5 passes
//    (5 CTA barriers) with 256 threads; an odd log2(n_fft) ends with one radix-2 pass.
//  * Twiddles come from a per-stage COMPACT table: entry (h - 1 + k) = exp(-2 pi i k / 2h), k < h, for h = 1, 2, 4 ...; a warp
//    reads consecutive entries (or one, broadcast) instead of a stride of n / 2h through one table.
// Begins and ends with a CTA barrier.
__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 5); }
__host__ __device__ constexpr int fft_padded_len(int n) { return n + n / 32; }

template <int kLog2N>
__device__ __forceinline__ void fft_stages(float* re, float* im, const float2* tws, int tid) {
    constexpr int n_fft = 1 << kLog2N, log2n = kLog2N;     // compile-time: every shift, bound and stride below folds to an immediate
    int st = 1;
#pragma unroll
    for (; st + 1 <= log2n; st += 2) {
        __syncthreads();
        const int h = 1 << (st - 1);
#pragma unroll
        for (int q = tid; q < n_fft / 4; q += kMelThreads) {
            const int k = q & (h - 1);
            const int base = ((q >> (st - 1)) << (st + 1)) + k;
            const int p0 = fft_pad(base), p1 = fft_pad(base + h), p2 = fft_pad(base + 2 * h), p3 = fft_pad(base + 3 * h);
            const float2 wa = tws[h - 1 + k];             // exp(-2 pi i k / 2h): stage st, both butterflies
            const float2 wb = tws[2 * h - 1 + k];         // exp(-2 pi i k / 4h): stage st + 1 (the second butterfly takes wb * -i)
            const float x0r = re[p0], x0i = im[p0], x1r = re[p1], x1i = im[p1];
            const float x2r = re[p2], x2i = im[p2], x3r = re[p3], x3i = im[p3];
            float tr = wa.x * x1r - wa.y * x1i, ti = wa.x * x1i + wa.y * x1r;
            const float y0r = x0r + tr, y0i = x0i + ti, y1r = x0r - tr, y1i = x0i - ti;
            tr = wa.x * x3r - wa.y * x3i; ti = wa.x * x3i + wa.y * x3r;
            const float y2r = x2r + tr, y2i = x2i + ti, y3r = x2r - tr, y3i = x2i - ti;
            tr = wb.x * y2r - wb.y * y2i; ti = wb.x * y2i + wb.y * y2r;
            re[p0] = y0r + tr; im[p0] = y0i + ti;
            re[p2] = y0r - tr; im[p2] = y0i - ti;
            const float ur = wb.x * y3r - wb.y * y3i, ui = wb.x * y3i + wb.y * y3r;   // (ur + i ui) * -i = ui - i ur
            re[p1] = y1r + ui; im[p1] = y1i - ur;
            re[p3] = y1r - ui; im[p3] = y1i + ur;
        }
    }
    if (log2n & 1) {
        st = log2n;
        __syncthreads();
        const int h = 1 << (st - 1);
#pragma unroll
        for (int q = tid; q < n_fft / 2; q += kMelThreads) {
            const int k = q & (h - 1);
            const int i0 = ((q >> (st - 1)) << st) + k;
            const int p0 = fft_pad(i0), p1 = fft_pad(i0 + h);
            const float2 w = tws[h - 1 + k];
            const float xr = re[p1], xi = im[p1];
            const float tr = w.x * xr - w.y * xi, ti = w.x * xi + w.y * xr;
            const float ur = re[p0], ui = im[p0];
            re[p0] = ur + tr; im[p0] = ui + ti;
            re[p1] = ur - tr; im[p1] = ui - ti;
        }
    }
    __syncthreads();
}

template <int kLog2N>
__global__ void __launch_bounds__(kMelThreads) logmel_kernel(const MelArgs a) {
    constexpr int kN = 1 << kLog2N;
    extern __shared__ __align__(16) float smem[];
    float* re = smem;                         // [n_fft] at padded positions
    float* im = re + fft_padded_len(kN);
    float2* tw = reinterpret_cast<float2*>(im + fft_padded_len(kN));   // [n_fft] per-stage compact twiddles
    float* win = reinterpret_cast<float*>(tw + kN);   // [n_fft]
    const int nbins = kN / 2 + 1;
    float* mag = win + kN;               // [kFramesPerCta][nbins]  |X| of the CTA's frames (odd row stride: rows fall on different banks)
    float* wts = mag + kFramesPerCta * nbins; // [n_weights]  the bands' non-zero filter weights, back to back
    int* band = reinterpret_cast<int*>(wts + a.n_weights);   // start | count | offset, n_mels each
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kFramesPerCta;
    const float* x = a.audio + (size_t)b * a.N;
    for (int i = tid; i < kN; i += kMelThreads) { tw[i] = a.twiddle[i]; win[i] = a.window[i]; }
    for (int i = tid; i < a.n_weights; i += kMelThreads) wts[i] = a.weights[i];
    for (int i = tid; i < 3 * a.n_mels; i += kMelThreads) band[i] = a.band_start[i];   // the three tables are one allocation
    // Two real frames per complex FFT: z = x_a + i x_b, and X_a[k] = (Z[k] + conj Z[N-k]) / 2, X_b[k] = (Z[k] - conj Z[N-k]) / 2i.
    for (int f = 0; f < kFramesPerCta; f += 2) {
        const int ta = t0 + f, tb = ta + 1;
        if (ta >= a.T) break;                 // uniform across the CTA
        const bool has_b = tb < a.T;
        __syncthreads();                      // tables loaded / previous pair's spectrum consumed
        const int sa = ta * a.hop - kN / 2, sb = sa + a.hop;
        for (int i = tid; i < kN; i += kMelThreads) {
            const int ia = sa + i, ib = sb + i;
            const float w = win[i];
            const float va = (ia >= 0 && ia < a.N) ? __ldg(x + ia) * w : 0.f;
            const float vb = (has_b && ib >= 0 && ib < a.N) ? __ldg(x + ib) * w : 0.f;
            const int j = fft_pad((int)(__brev((unsigned)i) >> (32 - kLog2N)));
            re[j] = va;
            im[j] = vb;
        }
        fft_stages<kLog2N>(re, im, tw, tid);
        float* mag_a = mag + f * nbins;
        float* mag_b = mag_a + nbins;
        for (int k = tid; k < nbins; k += kMelThreads) {
            const int nk = fft_pad((kN - k) & (kN - 1)), pk = fft_pad(k);
            const float zr = re[pk], zi = im[pk], yr = re[nk], yi = im[nk];
            const float ar = 0.5f * (zr + yr), ai = 0.5f * (zi - yi);      // X_a[k]
            const float br = 0.5f * (zi + yi), bi = -0.5f * (zr - yr);     // X_b[k]
            mag_a[k] = sqrtf(ar * ar + ai * ai);
            mag_b[k] = sqrtf(br * br + bi * bi);
        }
    }
    __syncthreads();
    // Sparse triangular mel filters for all frames of the CTA at once: one thread per (band, frame), serial over the band's
    // contiguous run of bins (13 on average, 2 .. 60); the 8 frames of a band leave as one run of consecutive floats.
    const int nf = min(kFramesPerCta, a.T - t0);
    for (int it = tid; it < a.n_mels * kFramesPerCta; it += kMelThreads) {
        const int m = it / kFramesPerCta, f = it - m * kFramesPerCta;
        if (f >= nf) continue;
        const int s = band[m], c = band[a.n_mels + m];
        const float* w = wts + band[2 * a.n_mels + m];
        const float* mg = mag + f * nbins + s;
        float acc = 0.f;
        for (int i = 0; i < c; ++i) acc = fmaf(w[i], mg[i], acc);
        a.out[((size_t)b * a.n_mels + m) * a.T + t0 + f] = a.log_output ? logf(fmaxf(acc, a.clip)) : acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Griffin-Lim (the reference's alternative vocoder, scripts/synthesize.py:193 -> librosa.griffinlim; oracle/griffinlim_oracle.py)
//   per iteration:  y = istft(S * angles) ;  X = stft(y) ;  angles = X - c * X_prev ;  angles /= |angles| + tiny ;  X_prev = X
// Spectra live frame-major, [B][T][nbins] float2, so that a frame's bins are contiguous.
// ---------------------------------------------------------------------------------------------------------------------
struct GlArgs {
    const float* mag;        // [B][T][nbins]   |S|
    float2* angles;          // [B][T][nbins]   unit phasors
    float2* prev;            // [B][T][nbins]   stft of the previous iteration
    float* y_acc;            // [B][n_fft + hop*(T-1)]  overlap-add accumulator
    float* y;                // [B][hop*(T-1)]
    const float* wss;        // [n_fft + hop*(T-1)]  window sum-square
    const float* window;
    const float2* twiddle;
    int T, n_fft, log2n, hop, N;   // N = hop*(T-1): samples of y
    float mom;               // momentum / (1 + momentum), 0 in the first iteration
};

// y_acc += window * irfft(S * angles), two frames per complex FFT: Z = X_a + i X_b with both spectra Hermitian-extended, and
// ifft(Z) = x_a + i x_b;  ifft(Z) = conj(fft(conj Z)) / n.
template <int kLog2N>
__global__ void __launch_bounds__(kMelThreads) gl_istft_kernel(const GlArgs a) {
    constexpr int kN = 1 << kLog2N;
    extern __shared__ __align__(16) float smem[];
    float* re = smem;
    float* im = re + fft_padded_len(kN);
    float2* tw = reinterpret_cast<float2*>(im + fft_padded_len(kN));
    float* win = reinterpret_cast<float*>(tw + kN);
    float* ola = win + kN;               // [n_fft + (kFramesPerCta-1)*hop]  this CTA's overlap-add
    const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * kFramesPerCta;
    const int nbins = kN / 2 + 1, half = kN / 2;
    const int span = kN + (kFramesPerCta - 1) * a.hop;
    for (int i = tid; i < kN; i += kMelThreads) { tw[i] = a.twiddle[i]; win[i] = a.window[i]; }
    for (int i = tid; i < span; i += kMelThreads) ola[i] = 0.f;
    const float inv_n = 1.0f / (float)kN;
    for (int f = 0; f < kFramesPerCta; f += 2) {
        const int ta = t0 + f, tb = ta + 1;
        if (ta >= a.T) break;
        const bool has_b = tb < a.T;
        const size_t ba = ((size_t)b * a.T + ta) * nbins, bb = ba + nbins;
        __syncthreads();
        for (int k = tid; k < kN; k += kMelThreads) {
            const int kk = k <= half ? k : kN - k;          // Hermitian extension: X[n-k] = conj X[k]
            const float sgn = k <= half ? 1.f : -1.f;
            const float2 ga = a.angles[ba + kk];
            const float ma = a.mag[ba + kk];
            float ar = ma * ga.x, ai = sgn * ma * ga.y;
            float br = 0.f, bi = 0.f;
            if (has_b) {
                const float2 gb = a.angles[bb + kk];
                const float mb = a.mag[bb + kk];
                br = mb * gb.x; bi = sgn * mb * gb.y;
            }
            if (kk == 0 || kk == half) { ai = 0.f; bi = 0.f; }   // irfft ignores the imaginary part of the DC and Nyquist bins
            // Z = X_a + i X_b ; load conj(Z) bit-reversed
            const int j = fft_pad((int)(__brev((unsigned)k) >> (32 - kLog2N)));
            re[j] = ar - bi;
            im[j] = -(ai + br);
        }
        fft_stages<kLog2N>(re, im, tw, tid);
        // ifft(Z) = conj(fft(conj Z)) / n:  x_a = re / n,  x_b = -im / n
        for (int i = tid; i < kN; i += kMelThreads) {
            const float w = win[i] * inv_n;
            ola[f * a.hop + i] += re[fft_pad(i)] * w;
        }
        __syncthreads();
        if (has_b)
            for (int i = tid; i < kN; i += kMelThreads) ola[(f + 1) * a.hop + i] -= im[fft_pad(i)] * win[i] * inv_n;
    }
    __syncthreads();
    const size_t ylen = (size_t)kN + (size_t)a.hop * (a.T - 1);
    const int nf = min(kFramesPerCta, a.T - t0);
    const int used = kN + (nf - 1) * a.hop;
    for (int i = tid; i < used; i += kMelThreads) atomicAdd(a.y_acc + (size_t)b * ylen + (size_t)t0 * a.hop + i, ola[i]);
}

// y = y_acc / window_sumsquare (where that is not ~0), trimmed by n_fft/2 on both sides (center=True)
__global__ void gl_norm_kernel(const GlArgs a) {
    const size_t ylen = (size_t)a.n_fft + (size_t)a.hop * (a.T - 1);
    const int b = blockIdx.y;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += gridDim.x * blockDim.x) {
        const size_t j = (size_t)n + a.n_fft / 2;
        const float w = a.wss[j];
        const float v = a.y_acc[(size_t)b * ylen + j];
        a.y[(size_t)b * a.N + n] = w > 1.17549435e-38f ? v / w : v;
    }
}

// X = stft(y) (two frames per FFT) ; angles = X - mom * prev ; angles /= |angles| + tiny ; prev = X
template <int kLog2N>
__global__ void __launch_bounds__(kMelThreads) gl_stft_update_kernel(const GlArgs a) {
    constexpr int kN = 1 << kLog2N;
    extern __shared__ __align__(16) float smem[];
    float* re = smem;
    float* im = re + fft_padded_len(kN);
    float2* tw = reinterpret_cast<float2*>(im + fft_padded_len(kN));
    float* win = reinterpret_cast<float*>(tw + kN);
    const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * kFramesPerCta;
    const int nbins = kN / 2 + 1;
    const float* x = a.y + (size_t)b * a.N;
    for (int i = tid; i < kN; i += kMelThreads) { tw[i] = a.twiddle[i]; win[i] = a.window[i]; }
    for (int f = 0; f < kFramesPerCta; f += 2) {
        const int ta = t0 + f, tb = ta + 1;
        if (ta >= a.T) break;
        const bool has_b = tb < a.T;
        __syncthreads();
        const int sa = ta * a.hop - kN / 2, sb = sa + a.hop;
        for (int i = tid; i < kN; i += kMelThreads) {
            const int ia = sa + i, ib = sb + i;
            const float w = win[i];
            const float va = (ia >= 0 && ia < a.N) ? __ldg(x + ia) * w : 0.f;
            const float vb = (has_b && ib >= 0 && ib < a.N) ? __ldg(x + ib) * w : 0.f;
            const int j = fft_pad((int)(__brev((unsigned)i) >> (32 - kLog2N)));
            re[j] = va;
            im[j] = vb;
        }
        fft_stages<kLog2N>(re, im, tw, tid);
        for (int k = tid; k < nbins; k += kMelThreads) {
            const int nk = fft_pad((kN - k) & (kN - 1)), pk = fft_pad(k);
            const float zr = re[pk], zi = im[pk], yr = re[nk], yi = im[nk];
            const float2 xa = make_float2(0.5f * (zr + yr), 0.5f * (zi - yi));
            const float2 xb = make_float2(0.5f * (zi + yi), -0.5f * (zr - yr));
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                if (which && !has_b) break;
                const size_t o = ((size_t)b * a.T + (which ? tb : ta)) * nbins + k;
                const float2 X = which ? xb : xa;
                const float2 p = a.prev[o];
                float gr = X.x - a.mom * p.x, gi = X.y - a.mom * p.y;
                const float inv = 1.0f / (sqrtf(gr * gr + gi * gi) + 1.17549435e-38f);
                a.angles[o] = make_float2(gr * inv, gi * inv);
                a.prev[o] = X;
            }
        }
    }
}

// mag[b][k][t] = max(0, sum_m proj[k][m] * exp(clamp(logmel[b][m][t], lo, hi)))   (scripts/synthesize.py:180-192 with a caller-supplied
// projection).  A CTA takes 32 time steps of one item: exp(clamp(.)) of the [n_mels][32] slab goes to shared memory once, then every
// warp walks bins k, lanes over time, the projection row broadcast from L1.  0.1 GFLOP for a 10 s utterance: nothing to tune.
__global__ void __launch_bounds__(256) mel_to_linear_kernel(const float* __restrict__ proj, const float* __restrict__ logmel, float* __restrict__ mag,
                                                            int n_bins, int n_mels, int T, float lo, float hi) {
    extern __shared__ float slab[];            // [n_mels][32]
    const int b = blockIdx.y, t0 = blockIdx.x * 32, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* src = logmel + (size_t)b * n_mels * T;
    for (int i = threadIdx.x; i < n_mels * 32; i += 256) {
        const int m = i >> 5, t = t0 + (i & 31);
        slab[i] = t < T ? expf(fminf(fmaxf(src[(size_t)m * T + t], lo), hi)) : 0.f;
    }
    __syncthreads();
    const int t = t0 + lane;
    for (int k = warp; k < n_bins; k += 8) {
        const float* p = proj + (size_t)k * n_mels;
        float acc = 0.f;
        for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(p + m), slab[m * 32 + lane], acc);
        if (t < T) mag[((size_t)b * n_bins + k) * T + t] = fmaxf(acc, 0.f);
    }
}

// librosa.hz_to_mel / mel_to_hz, Slaney variant (htk=False)
double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_hz / f_sp + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// The FFT kernels are compiled per transform length (n_fft = 64 .. 4096): f receives std::integral_constant<int, log2(n_fft)>.
template <typename F>
cudaError_t mel_dispatch(int log2n, F&& f) {
    switch (log2n) {
        case 6: return f(std::integral_constant<int, 6>{});
        case 7: return f(std::integral_constant<int, 7>{});
        case 8: return f(std::integral_constant<int, 8>{});
        case 9: return f(std::integral_constant<int, 9>{});
        case 10: return f(std::integral_constant<int, 10>{});
        case 11: return f(std::integral_constant<int, 11>{});
        case 12: return f(std::integral_constant<int, 12>{});
        default: return cudaErrorInvalidValue;
    }
}

int mel_fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}

}  // namespace
}  // namespace hfg

using namespace hfg;

struct hfg_logmel {
    hfg_logmel_config cfg;
    int device = 0;
    int log2n = 0;
    cudaStream_t stream = nullptr;
    float* d_window = nullptr;
    float2* d_twiddle = nullptr;
    int* d_band = nullptr;       // start | count | off, n_mels each
    float* d_weights = nullptr;
    float* d_audio = nullptr;    // staging for host pointers
    float* d_out = nullptr;
    size_t audio_cap = 0, out_cap = 0;
    char* d_gl = nullptr;        // Griffin-Lim workspace (one grow-only allocation: cudaMalloc / cudaFree per call cost more than the iteration)
    size_t gl_cap = 0;
    size_t smem = 0;
    int n_weights = 0;
};

namespace {
struct MelDeviceGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit MelDeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        // ALWAYS: a thread that has made no CUDA call yet reports device 0 as current without having a context bound, and the
        // driver entry points the planner calls (cuTensorMapEncodeTiled) then fail with CUDA_ERROR_INVALID_CONTEXT (seen from a
        // worker thread whose page-locked buffers came out of torch's host cache, i.e. with no runtime call before this one)
        err = cudaSetDevice(d);
    }
    ~MelDeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define MCK(expr)                                                                                     \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return mel_fail(HFG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
}  // namespace

extern "C" {

int hfg_logmel_create(const hfg_logmel_config* cfg, int device, hfg_logmel** out) {
    if (!cfg || !out) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_create: null argument");
    *out = nullptr;
    const hfg_logmel_config& c = *cfg;
    int log2n = 0;
    while ((1 << log2n) < c.n_fft) ++log2n;
    if (c.n_fft < 64 || c.n_fft > 4096 || (1 << log2n) != c.n_fft) return mel_fail(HFG_ERR_UNSUPPORTED, "hfg_logmel_create: n_fft must be a power of two in [64, 4096]");
    if (c.win_length <= 0 || c.win_length > c.n_fft || c.hop_length <= 0) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_create: need 0 < win_length <= n_fft and hop_length > 0");
    if (c.n_mels <= 0 || c.n_mels > 512 || c.sample_rate <= 0) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_create: n_mels / sample_rate out of range");
    const double fmax = c.fmax > 0.f ? (double)c.fmax : c.sample_rate / 2.0;
    if (c.fmin < 0.f || fmax <= c.fmin) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_create: need 0 <= fmin < fmax");
    if (hfg_device_count() <= 0) return mel_fail(HFG_ERR_CUDA, "hfg_logmel_create: no CUDA device available (no CPU fallback)");
    if (device < 0 || device >= hfg_device_count()) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_create: device index out of range");
    MelDeviceGuard guard(device);
    MCK(guard.err);
    std::unique_ptr<hfg_logmel> h(new hfg_logmel());
    h->cfg = c; h->device = device; h->log2n = log2n;
    const int n = c.n_fft, nbins = n / 2 + 1;
    // window: scipy.signal.get_window('hann', win_length, fftbins=True), padded centrally to n_fft (librosa util.pad_center)
    std::vector<float> win((size_t)n, 0.f);
    const int lpad = (n - c.win_length) / 2;
    for (int i = 0; i < c.win_length; ++i) win[lpad + i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / c.win_length));
    // per-stage compact twiddle table (fft_stages): entry h - 1 + k = exp(-2 pi i k / 2h), k < h, h = 1, 2, 4 ... n / 2
    std::vector<float2> tw((size_t)n, make_float2(0.f, 0.f));
    for (int hh = 1; hh < n; hh <<= 1)
        for (int k = 0; k < hh; ++k) tw[hh - 1 + k] = make_float2((float)cos(M_PI * k / hh), (float)(-sin(M_PI * k / hh)));
    // librosa.filters.mel(htk=False, norm='slaney')
    std::vector<double> mel_f((size_t)c.n_mels + 2);
    const double m_lo = hz_to_mel(c.fmin), m_hi = hz_to_mel(fmax);
    for (int i = 0; i < c.n_mels + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (c.n_mels + 1));
    std::vector<int> band((size_t)3 * c.n_mels, 0);
    std::vector<float> weights;
    for (int m = 0; m < c.n_mels; ++m) {
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        int first = -1, last = -1;
        std::vector<float> row((size_t)nbins, 0.f);
        for (int k = 0; k < nbins; ++k) {
            const double f = (c.sample_rate / 2.0) * k / (nbins - 1);
            const double lower = (f - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
            const double upper = (mel_f[m + 2] - f) / (mel_f[m + 2] - mel_f[m + 1]);
            const double w = std::max(0.0, std::min(lower, upper)) * enorm;
            row[k] = (float)w;
            if (w > 0.0) { if (first < 0) first = k; last = k; }
        }
        band[m] = first < 0 ? 0 : first;
        band[c.n_mels + m] = first < 0 ? 0 : last - first + 1;
        band[2 * c.n_mels + m] = (int)weights.size();
        for (int k = 0; first >= 0 && k <= last - first; ++k) weights.push_back(row[first + k]);
    }
    if (weights.empty()) weights.push_back(0.f);
    MCK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    MCK(cudaMalloc(&h->d_window, win.size() * sizeof(float)));
    MCK(cudaMalloc(&h->d_twiddle, tw.size() * sizeof(float2)));
    MCK(cudaMalloc(&h->d_band, band.size() * sizeof(int)));
    MCK(cudaMalloc(&h->d_weights, weights.size() * sizeof(float)));
    MCK(cudaMemcpy(h->d_window, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(h->d_twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(h->d_band, band.data(), band.size() * sizeof(int), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(h->d_weights, weights.data(), weights.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->n_weights = (int)weights.size();
    h->smem = (size_t)(2 * fft_padded_len(n) + 2 * n + n + kFramesPerCta * nbins + weights.size() + 3 * c.n_mels) * sizeof(float);
    if (h->smem > 48 * 1024) MCK(mel_dispatch(log2n, [&](auto tag) {
        return cudaFuncSetAttribute(logmel_kernel<decltype(tag)::value>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem);
    }));
    *out = h.release();
    return HFG_OK;
}

void hfg_logmel_destroy(hfg_logmel* h) {
    if (!h) return;
    MelDeviceGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_window); cudaFree(h->d_twiddle); cudaFree(h->d_band); cudaFree(h->d_weights); cudaFree(h->d_audio); cudaFree(h->d_out); cudaFree(h->d_gl);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int32_t hfg_logmel_frames(const hfg_logmel* h, int32_t n_samples) {
    if (!h || n_samples < 0) return 0;
    return 1 + n_samples / h->cfg.hop_length;
}

int hfg_logmel_forward(hfg_logmel* h, const float* audio, int32_t B, int32_t N, float* mel, uint32_t flags) {
    if (!h || !audio || !mel) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_forward: null argument");
    if (B <= 0 || N <= 0) return mel_fail(HFG_ERR_INVALID, "hfg_logmel_forward: B and N must be positive");
    MelDeviceGuard guard(h->device);
    MCK(guard.err);
    const int T = 1 + N / h->cfg.hop_length;
    const size_t n_in = (size_t)B * N, n_out = (size_t)B * h->cfg.n_mels * T;
    const bool in_dev = flags & HFG_LOGMEL_AUDIO_ON_DEVICE, out_dev = flags & HFG_LOGMEL_OUT_ON_DEVICE;
    const float* d_in = audio;
    float* d_o = mel;
    if (!in_dev) {
        if (n_in > h->audio_cap) {
            MCK(cudaStreamSynchronize(h->stream));
            cudaFree(h->d_audio); h->d_audio = nullptr; h->audio_cap = 0;
            MCK(cudaMalloc(&h->d_audio, n_in * sizeof(float)));
            h->audio_cap = n_in;
        }
        MCK(cudaMemcpyAsync(h->d_audio, audio, n_in * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        d_in = h->d_audio;
    }
    if (!out_dev) {
        if (n_out > h->out_cap) {
            MCK(cudaStreamSynchronize(h->stream));
            cudaFree(h->d_out); h->d_out = nullptr; h->out_cap = 0;
            MCK(cudaMalloc(&h->d_out, n_out * sizeof(float)));
            h->out_cap = n_out;
        }
        d_o = h->d_out;
    }
    MelArgs a;
    a.audio = d_in; a.out = d_o; a.window = h->d_window; a.twiddle = h->d_twiddle;
    a.band_start = h->d_band; a.band_count = h->d_band + h->cfg.n_mels; a.band_off = h->d_band + 2 * h->cfg.n_mels;
    a.weights = h->d_weights; a.n_weights = h->n_weights;
    a.N = N; a.T = T; a.n_fft = h->cfg.n_fft; a.log2n = h->log2n; a.hop = h->cfg.hop_length; a.n_mels = h->cfg.n_mels;
    a.clip = h->cfg.clip; a.log_output = h->cfg.log_output;
    dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
    MCK(mel_dispatch(h->log2n, [&](auto tag) {
        logmel_kernel<decltype(tag)::value><<<grid, kMelThreads, h->smem, h->stream>>>(a);
        return cudaGetLastError();
    }));
    if (!out_dev) MCK(cudaMemcpyAsync(mel, d_o, n_out * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    MCK(cudaStreamSynchronize(h->stream));
    return HFG_OK;
}

int hfg_griffin_lim(hfg_logmel* h, const float* mag, const float* angles0, int32_t B, int32_t T, int32_t n_iter, float momentum,
                    float* audio) {
    if (!h || !mag || !angles0 || !audio) return mel_fail(HFG_ERR_INVALID, "hfg_griffin_lim: null argument");
    if (B <= 0 || T < 2 || n_iter < 0 || momentum < 0.f) return mel_fail(HFG_ERR_INVALID, "hfg_griffin_lim: need B > 0, T >= 2, n_iter >= 0, momentum >= 0");
    MelDeviceGuard guard(h->device);
    MCK(guard.err);
    const hfg_logmel_config& c = h->cfg;
    const int n = c.n_fft, nbins = n / 2 + 1, hop = c.hop_length;
    const int N = hop * (T - 1);
    const size_t ylen = (size_t)n + (size_t)hop * (T - 1);
    const size_t nspec = (size_t)B * T * nbins;
    // librosa.filters.window_sumsquare: squared synthesis window overlap-added at every frame position
    std::vector<float> wss(ylen, 0.f);
    {
        std::vector<double> w2((size_t)n, 0.0), acc(ylen, 0.0);
        const int lpad = (n - c.win_length) / 2;
        for (int i = 0; i < c.win_length; ++i) { const double w = 0.5 - 0.5 * cos(2.0 * M_PI * i / c.win_length); w2[lpad + i] = w * w; }
        for (int t = 0; t < T; ++t)
            for (int i = 0; i < n; ++i) acc[(size_t)t * hop + i] += w2[i];
        for (size_t i = 0; i < ylen; ++i) wss[i] = (float)acc[i];
    }
#define GCK(expr)                                                                                                       \
    do {                                                                                                                \
        cudaError_t _e = (expr);                                                                                        \
        if (_e != cudaSuccess) return mel_fail(HFG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)
    // workspace: | mag (librosa layout) | mag (frame-major) | angles | prev | y_acc | y | wss |, each 256-byte aligned
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_mag_cf = 0, o_mag = o_mag_cf + al(nspec * sizeof(float)), o_ang = o_mag + al(nspec * sizeof(float));
    const size_t o_prev = o_ang + al(nspec * sizeof(float2)), o_yacc = o_prev + al(nspec * sizeof(float2));
    const size_t o_y = o_yacc + al((size_t)B * ylen * sizeof(float)), o_wss = o_y + al((size_t)B * N * sizeof(float));
    const size_t need = o_wss + al(ylen * sizeof(float));
    if (need > h->gl_cap) {
        GCK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_gl); h->d_gl = nullptr; h->gl_cap = 0;
        GCK(cudaMalloc(&h->d_gl, need));
        h->gl_cap = need;
    }
    float* d_mag_cf = reinterpret_cast<float*>(h->d_gl + o_mag_cf);
    float* d_mag = reinterpret_cast<float*>(h->d_gl + o_mag);
    float2* d_ang = reinterpret_cast<float2*>(h->d_gl + o_ang);
    float2* d_prev = reinterpret_cast<float2*>(h->d_gl + o_prev);
    float* d_yacc = reinterpret_cast<float*>(h->d_gl + o_yacc);
    float* d_y = reinterpret_cast<float*>(h->d_gl + o_y);
    float* d_wss = reinterpret_cast<float*>(h->d_gl + o_wss);
    cudaStream_t st = h->stream;
    GCK(cudaMemcpyAsync(d_mag_cf, mag, nspec * sizeof(float), cudaMemcpyHostToDevice, st));
    GCK(launch_transpose_cf_to_cl(d_mag_cf, d_mag, B, nbins, T, st));     // [B][nbins][T] (librosa layout) -> frame-major [B][T][nbins]
    GCK(cudaMemcpyAsync(d_ang, angles0, nspec * sizeof(float2), cudaMemcpyHostToDevice, st));
    GCK(cudaMemsetAsync(d_prev, 0, nspec * sizeof(float2), st));
    GCK(cudaMemcpyAsync(d_wss, wss.data(), ylen * sizeof(float), cudaMemcpyHostToDevice, st));
    GlArgs a;
    a.mag = d_mag; a.angles = d_ang; a.prev = d_prev; a.y_acc = d_yacc; a.y = d_y; a.wss = d_wss;
    a.window = h->d_window; a.twiddle = h->d_twiddle;
    a.T = T; a.n_fft = n; a.log2n = h->log2n; a.hop = hop; a.N = N; a.mom = 0.f;
    const size_t smem_i = (size_t)(2 * fft_padded_len(n) + 3 * n + n + (kFramesPerCta - 1) * hop) * sizeof(float);
    const size_t smem_s = (size_t)(2 * fft_padded_len(n) + 3 * n) * sizeof(float);
    if (smem_i > 48 * 1024) GCK(mel_dispatch(h->log2n, [&](auto tag) {
        return cudaFuncSetAttribute(gl_istft_kernel<decltype(tag)::value>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i);
    }));
    if (smem_s > 48 * 1024) GCK(mel_dispatch(h->log2n, [&](auto tag) {
        return cudaFuncSetAttribute(gl_stft_update_kernel<decltype(tag)::value>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
    }));
    const dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
    const dim3 ngrid((unsigned)std::min<size_t>(((size_t)N + 255) / 256, 1024), B);
    for (int it = 0; it <= n_iter; ++it) {
        GCK(cudaMemsetAsync(d_yacc, 0, (size_t)B * ylen * sizeof(float), st));
        GCK(mel_dispatch(h->log2n, [&](auto tag) {
            gl_istft_kernel<decltype(tag)::value><<<grid, kMelThreads, smem_i, st>>>(a);
            return cudaGetLastError();
        }));
        gl_norm_kernel<<<ngrid, 256, 0, st>>>(a);
        if (it == n_iter) break;                                            // the last inverse transform is the result
        a.mom = it == 0 ? 0.f : momentum / (1.0f + momentum);
        GCK(mel_dispatch(h->log2n, [&](auto tag) {
            gl_stft_update_kernel<decltype(tag)::value><<<grid, kMelThreads, smem_s, st>>>(a);
            return cudaGetLastError();
        }));
    }
    GCK(cudaGetLastError());
    GCK(cudaMemcpyAsync(audio, d_y, (size_t)B * N * sizeof(float), cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
#undef GCK
    return HFG_OK;
}

int hfg_mel_to_linear(hfg_logmel* h, const float* proj, const float* logmel, int32_t B, int32_t n_mels, int32_t T, float lo, float hi,
                      float* mag) {
    if (!h || !proj || !logmel || !mag) return mel_fail(HFG_ERR_INVALID, "hfg_mel_to_linear: null argument");
    if (B <= 0 || n_mels <= 0 || T <= 0 || !(lo <= hi)) return mel_fail(HFG_ERR_INVALID, "hfg_mel_to_linear: need B, n_mels, T > 0 and lo <= hi");
    if (B > 65535 || n_mels > 1024) return mel_fail(HFG_ERR_UNSUPPORTED, "hfg_mel_to_linear: B <= 65535 and n_mels <= 1024");
    MelDeviceGuard guard(h->device);
    MCK(guard.err);
    const int n_bins = h->cfg.n_fft / 2 + 1;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_proj = (size_t)n_bins * n_mels * sizeof(float), b_in = (size_t)B * n_mels * T * sizeof(float), b_out = (size_t)B * n_bins * T * sizeof(float);
    const size_t need = al(b_proj) + al(b_in) + al(b_out);
    if (need > h->gl_cap) {     // shares the Griffin-Lim workspace (the two calls follow each other)
        MCK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_gl); h->d_gl = nullptr; h->gl_cap = 0;
        MCK(cudaMalloc(&h->d_gl, need));
        h->gl_cap = need;
    }
    float* d_proj = reinterpret_cast<float*>(h->d_gl);
    float* d_in = reinterpret_cast<float*>(h->d_gl + al(b_proj));
    float* d_out = reinterpret_cast<float*>(h->d_gl + al(b_proj) + al(b_in));
    MCK(cudaMemcpyAsync(d_proj, proj, b_proj, cudaMemcpyHostToDevice, h->stream));
    MCK(cudaMemcpyAsync(d_in, logmel, b_in, cudaMemcpyHostToDevice, h->stream));
    const size_t smem = (size_t)n_mels * 32 * sizeof(float);
    if (smem > 48 * 1024) MCK(cudaFuncSetAttribute(mel_to_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mel_to_linear_kernel<<<dim3((T + 31) / 32, B), 256, smem, h->stream>>>(d_proj, d_in, d_out, n_bins, n_mels, T, lo, hi);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(mag, d_out, b_out, cudaMemcpyDeviceToHost, h->stream));
    MCK(cudaStreamSynchronize(h->stream));
    return HFG_OK;
}

}  // extern "C"
