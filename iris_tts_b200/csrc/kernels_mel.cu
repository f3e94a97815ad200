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
    int N, T, n_fft, log2n, hop, n_mels;
    float clip;
    int log_output;
};

__global__ void __launch_bounds__(kMelThreads) logmel_kernel(const MelArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* re = smem;                         // [n_fft]
    float* im = re + a.n_fft;                 // [n_fft]
    float2* tw = reinterpret_cast<float2*>(im + a.n_fft);   // [n_fft/2]
    float* win = reinterpret_cast<float*>(tw + a.n_fft / 2);   // [n_fft]
    const int nbins = a.n_fft / 2 + 1;
    float* mag_a = win + a.n_fft;             // [nbins]  |X| of the pair's first frame
    float* mag_b = mag_a + nbins;             // [nbins]  ... second frame
    float* mel_s = mag_b + nbins;             // [n_mels][kFramesPerCta]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kFramesPerCta;
    const float* x = a.audio + (size_t)b * a.N;
    for (int i = tid; i < a.n_fft / 2; i += kMelThreads) tw[i] = a.twiddle[i];
    for (int i = tid; i < a.n_fft; i += kMelThreads) win[i] = a.window[i];
    // Two real frames per complex FFT: z = x_a + i x_b, and X_a[k] = (Z[k] + conj Z[N-k]) / 2, X_b[k] = (Z[k] - conj Z[N-k]) / 2i.
    for (int f = 0; f < kFramesPerCta; f += 2) {
        const int ta = t0 + f, tb = ta + 1;
        if (ta >= a.T) break;                 // uniform across the CTA
        const bool has_b = tb < a.T;
        __syncthreads();                      // tables loaded / previous pair's magnitudes consumed
        const int sa = ta * a.hop - a.n_fft / 2, sb = sa + a.hop;
        for (int i = tid; i < a.n_fft; i += kMelThreads) {
            const int ia = sa + i, ib = sb + i;
            const float w = win[i];
            const float va = (ia >= 0 && ia < a.N) ? __ldg(x + ia) * w : 0.f;
            const float vb = (has_b && ib >= 0 && ib < a.N) ? __ldg(x + ib) * w : 0.f;
            const int j = (int)(__brev((unsigned)i) >> (32 - a.log2n));
            re[j] = va;
            im[j] = vb;
        }
        for (int st = 1; st <= a.log2n; ++st) {
            __syncthreads();
            const int half = 1 << (st - 1);
            const int tstep = a.n_fft >> st;
            for (int idx = tid; idx < a.n_fft / 2; idx += kMelThreads) {
                const int k = idx & (half - 1);
                const int i0 = ((idx >> (st - 1)) << st) + k;
                const int i1 = i0 + half;
                const float2 w = tw[k * tstep];
                const float xr = re[i1], xi = im[i1];
                const float tr = w.x * xr - w.y * xi;
                const float ti = w.x * xi + w.y * xr;
                const float ur = re[i0], ui = im[i0];
                re[i0] = ur + tr; im[i0] = ui + ti;
                re[i1] = ur - tr; im[i1] = ui - ti;
            }
        }
        __syncthreads();
        for (int k = tid; k < nbins; k += kMelThreads) {
            const int nk = (a.n_fft - k) & (a.n_fft - 1);
            const float zr = re[k], zi = im[k], yr = re[nk], yi = im[nk];
            const float ar = 0.5f * (zr + yr), ai = 0.5f * (zi - yi);      // X_a[k]
            const float br = 0.5f * (zi + yi), bi = -0.5f * (zr - yr);     // X_b[k]
            mag_a[k] = sqrtf(ar * ar + ai * ai);
            mag_b[k] = sqrtf(br * br + bi * bi);
        }
        __syncthreads();
        // one warp per mel band, lanes over the band's bins (a contiguous run), both frames of the pair at once
        for (int m = warp; m < a.n_mels; m += kMelThreads / 32) {
            const int s = a.band_start[m], c = a.band_count[m];
            const float* w = a.weights + a.band_off[m];
            float acc_a = 0.f, acc_b = 0.f;
            for (int i = lane; i < c; i += 32) {
                const float wi = __ldg(w + i);
                acc_a = fmaf(wi, mag_a[s + i], acc_a);
                acc_b = fmaf(wi, mag_b[s + i], acc_b);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                acc_a += __shfl_xor_sync(0xffffffffu, acc_a, d);
                acc_b += __shfl_xor_sync(0xffffffffu, acc_b, d);
            }
            if (lane == 0) {
                mel_s[m * kFramesPerCta + f] = a.log_output ? logf(fmaxf(acc_a, a.clip)) : acc_a;
                mel_s[m * kFramesPerCta + f + 1] = a.log_output ? logf(fmaxf(acc_b, a.clip)) : acc_b;
            }
        }
    }
    __syncthreads();
    const int nf = min(kFramesPerCta, a.T - t0);
    for (int i = tid; i < a.n_mels * kFramesPerCta; i += kMelThreads) {
        const int m = i / kFramesPerCta, f = i - m * kFramesPerCta;
        if (f < nf) a.out[((size_t)b * a.n_mels + m) * a.T + t0 + f] = mel_s[i];
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
    size_t smem = 0;
};

namespace {
struct MelDeviceGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit MelDeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != d) err = cudaSetDevice(d);
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
    std::vector<float2> tw((size_t)n / 2);
    for (int k = 0; k < n / 2; ++k) tw[k] = make_float2((float)cos(2.0 * M_PI * k / n), (float)(-sin(2.0 * M_PI * k / n)));
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
    h->smem = (size_t)(2 * n + n + n + 2 * nbins + c.n_mels * kFramesPerCta) * sizeof(float);
    if (h->smem > 48 * 1024) MCK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    *out = h.release();
    return HFG_OK;
}

void hfg_logmel_destroy(hfg_logmel* h) {
    if (!h) return;
    MelDeviceGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_window); cudaFree(h->d_twiddle); cudaFree(h->d_band); cudaFree(h->d_weights); cudaFree(h->d_audio); cudaFree(h->d_out);
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
    a.weights = h->d_weights;
    a.N = N; a.T = T; a.n_fft = h->cfg.n_fft; a.log2n = h->log2n; a.hop = h->cfg.hop_length; a.n_mels = h->cfg.n_mels;
    a.clip = h->cfg.clip; a.log_output = h->cfg.log_output;
    dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
    logmel_kernel<<<grid, kMelThreads, h->smem, h->stream>>>(a);
    MCK(cudaGetLastError());
    if (!out_dev) MCK(cudaMemcpyAsync(mel, d_o, n_out * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    MCK(cudaStreamSynchronize(h->stream));
    return HFG_OK;
}

}  // extern "C"
