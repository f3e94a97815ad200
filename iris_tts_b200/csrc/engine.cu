// libhfg_b200: the C-ABI engine behind include/hfg.h.
//
// Host-side graph of the reference generator (src/iris/hifigan_pretrained.py:77-143)
// lowered to a cached plan of CUDA launches per (B, T, precision).  Weight-norm is folded
// once at load (:49,55,92,100,119 re-run it on every forward).  There is no CPU path:
// every entry point that computes requires a CUDA device.
#include <cuda_profiler_api.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "hfg_internal.h"

namespace hfg {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

namespace {

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    return HFG_ERR_CUDA;
}
#define CK(expr)                                              \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);   \
    } while (0)
// Every ABI entry point runs on the engine's device and leaves the caller's current device as it found it
// (a torch process that calls an engine on cuda:1 keeps torch.cuda.current_device() == 0).
struct DeviceGuard {
    int prev = -1, dev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        // ALWAYS: a thread that has made no CUDA call yet reports device 0 as current without having a context bound, and the
        // driver entry points the planner calls (cuTensorMapEncodeTiled) then fail with CUDA_ERROR_INVALID_CONTEXT (seen from a
        // worker thread whose page-locked buffers came out of torch's host cache, i.e. with no runtime call before this one)
        err = cudaSetDevice(d);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define GUARD(e)                       \
    DeviceGuard _guard((e)->device);   \
    CK(_guard.err)
#define RET(expr)                     \
    do {                              \
        int _r = (expr);              \
        if (_r != HFG_OK) return _r;  \
    } while (0)

struct Layer {
    std::string name;
    bool transposed = false;
    int cin = 0, cout = 0, k = 0, dil = 1, stride = 1, pad = 0;
    bool is_post = false;
    bool set = false;
    std::vector<float> w;     // folded, torch layout
    std::vector<float> bias;
    // GEMM form
    int taps = 0, tap_off0 = 0, tap_step = 0, Np = 0, ups_s = 1, ups_p = 0;
    int cin_pad = 0, kc = 64;   // tensor-core family: channels of the input planes (>= 32, multiple of kc)
    int cout_tc = 0, Np_tc = 0;  // tensor-core family: output channels padded to >= 32 (zero weights / bias), GEMM N
    // device packs
    float* d_w32 = nullptr;            // [taps][Cin][Np]  (conv_post: [k][C])
    float* d_bias = nullptr;           // [Cout]
    float* d_bias_tc = nullptr;        // [cout_tc], zero padded
    float* d_w32_tc = nullptr;         // conv_post only: [k][cin_pad], zero padded
    __nv_bfloat16* d_wb_hi = nullptr;  // [taps*Np][cin_pad]
    __nv_bfloat16* d_wb_lo = nullptr;
    __nv_bfloat16* d_wh = nullptr;     // same layout, fp16 values (HFG_PREC_FP16); 16-bit storage shares the pointer type
};

enum StepKind { S_CONV32, S_UMMA, S_UMMA2, S_PAIR, S_POST32, S_ACCUM, S_MEL_CL32, S_MEL_CLBF, S_P2RAW, S_MRF, S_POSTMRF, S_TAP,
                S_STRIP,            // ragged batches: zero the rows behind each item's own end of the planes the previous step wrote
                S_FORK, S_JOIN };   // the branch lanes of a stage start after / end before this point (no kernel)

struct Step {
    StepKind kind;
    ConvParams cp;
    UmmaLaunch ul;
    Umma2Launch u2;
    PairLaunch pl;
    MrfArgs mrf;
    // misc operands
    const float* f_in = nullptr;
    float* f_out = nullptr;
    const __nv_bfloat16* b_in = nullptr;
    const __nv_bfloat16* b_in_lo = nullptr;
    __nv_bfloat16* b_out = nullptr;
    __nv_bfloat16* b_out_lo = nullptr;
    const float* w = nullptr;
    const float* bias = nullptr;
    int B = 0, L = 0, C = 0, k = 0, cpad = 0;
    int flag0 = 0, flag1 = 0, f16 = 0;
    int lane = 0;   // 0: the engine's stream; j > 0: ResBlock branch j of a stage whose branches run concurrently (small inputs)
    float fval = 0.f;
    size_t n = 0;
    std::string tap_name;
    // profiling: reference layer this launch belongs to and its algorithmic work
    std::string label;
    double flops = 0.0, bytes = 0.0;
};

struct Plan {
    std::vector<Step> steps;
    float* mel_dev = nullptr;   // [B][Cin][T] staging when the caller passes a host pointer
    float* wave_dev = nullptr;  // [B][T*hop]
    int32_t* lens_dev = nullptr; // ragged plans: [B] mel frames per item (copied in before every launch of the plan)
    size_t bytes = 0;
    bool keep_taps = false;
    std::vector<size_t> guards;   // HFG_GUARD: canary regions between the workspace buffers
    // CUDA graph of the plan's kernel launches (captured on the second forward of a plan; the H2D / D2H copies stay outside
    // because their host pointers change per call).  Programmatic-dependent-launch edges survive capture.
    cudaGraphExec_t graph = nullptr;
    bool graph_failed = false;
    int runs = 0;
    int kernels = 0;              // launches per forward (steps that are not tap copies)
    ~Plan() { if (graph) cudaGraphExecDestroy(graph); }
};

struct Tap {
    float* dev = nullptr;  // channels-last [B][L][C] (C==1: [B][L])
    int B = 0, C = 0, L = 0;
};

}  // namespace
}  // namespace hfg

using namespace hfg;

struct hfg_engine {
    hfg_config cfg;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::vector<Layer> layers;
    std::map<std::string, int> index;
    // time-folded twins of the layers of the narrow stages (C < 32), see build_folded
    std::map<std::string, Layer> folded;
    int stage_fold[HFG_MAX_UPSAMPLES] = {};   // fold factor of stage i's planes (1: not folded)
    bool finalized = false;
    int hop = 1;
    // workspace
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    std::map<std::tuple<int, int, int, int>, std::unique_ptr<Plan>> plans;
    std::map<std::string, Tap> taps;
    uint64_t launches = 0;
    // per-launch event timing (hfg_profile_*)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    struct ProfRec { std::string label; int kind; double flops, bytes; size_t ev0, ev1; };
    std::vector<ProfRec> prof_recs;   // accumulated over forwards since hfg_profile_enable(1)
    size_t prof_used = 0;             // events consumed from the pool
    std::vector<std::string> ncu_layers;   // HFG_NCU_LAYERS: bracket these layers with cudaProfilerStart/Stop
    // run_layer scratch
    uint8_t* scratch = nullptr;
    size_t scratch_bytes = 0;
    // asynchronous host output (HFG_NO_SYNC with a page-locked host wave pointer): the waveform is parked in one of two device
    // staging buffers and copied to the host on a second stream, so the D2H of forward i overlaps the kernels of forward i+1
    // The nk ResBlock branches of a stage are independent (hifigan_pretrained.py:131-136).  When one kernel cannot fill the GPU
    // (batch 1-2, long-form chunks) branches 1.. run on these streams, forked after the upsampler and joined before the MRF mean;
    // under graph capture they become parallel branches of the plan's graph.
    cudaStream_t lane_stream[HFG_MAX_KERNELS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[HFG_MAX_KERNELS] = {};
    cudaStream_t copy_stream = nullptr;
    float* wave_stage[2] = {nullptr, nullptr};
    size_t wave_stage_bytes[2] = {0, 0};
    cudaEvent_t ev_plan_done[2] = {nullptr, nullptr}, ev_d2h_done[2] = {nullptr, nullptr};
    bool d2h_pending[2] = {false, false};
    uint64_t async_count = 0;
    int graphs_captured = 0, graphs_failed = 0;
};

namespace hfg {
namespace {

int get_padding(int k, int d) { return (k * d - d) / 2; }  // hifigan_pretrained.py:61-62

int env_flag_early(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

int validate_cfg(const hfg_config& c) {
    if (c.in_channels <= 0 || c.upsample_initial_channel <= 0) return fail(HFG_ERR_INVALID, "config: channels must be positive");
    if (c.num_upsamples <= 0 || c.num_upsamples > HFG_MAX_UPSAMPLES) return fail(HFG_ERR_INVALID, "config: num_upsamples out of range");
    if (c.num_kernels <= 0 || c.num_kernels > HFG_MAX_KERNELS) return fail(HFG_ERR_INVALID, "config: num_kernels out of range");
    if (c.upsample_initial_channel % (1 << c.num_upsamples) != 0)
        return fail(HFG_ERR_INVALID, "config: upsample_initial_channel must be divisible by 2^num_upsamples");
    for (int i = 0; i < c.num_upsamples; ++i) {
        const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
        if (u <= 0 || k <= 0) return fail(HFG_ERR_INVALID, "config: upsample rates and kernel sizes must be positive");
        // the reference builds these too (padding = (k - u) // 2, hifigan_pretrained.py:98-104), but their output is not T * hop samples
        // long (k < u leaves gaps, an odd k - u adds one sample per stage): no generator uses them, and this engine does not take them
        if (k < u || (k - u) % 2 != 0) return fail(HFG_ERR_UNSUPPORTED, "config: upsample kernel must be >= rate with even difference");
    }
    for (int j = 0; j < c.num_kernels; ++j) {
        if (c.resblock_kernel_sizes[j] <= 0 || c.resblock_kernel_sizes[j] % 2 == 0)
            return fail(HFG_ERR_INVALID, "config: resblock kernel sizes must be odd");
        if (c.num_dilations[j] <= 0 || c.num_dilations[j] > HFG_MAX_DILATIONS) return fail(HFG_ERR_INVALID, "config: num_dilations out of range");
        for (int m = 0; m < c.num_dilations[j]; ++m)
            if (c.resblock_dilations[j][m] <= 0) return fail(HFG_ERR_INVALID, "config: dilations must be positive");
    }
    const int c_last = c.upsample_initial_channel >> c.num_upsamples;
    if (c_last < 8 || c_last > 128 || (c_last & (c_last - 1)) != 0)
        return fail(HFG_ERR_UNSUPPORTED, "config: final channel count must be a power of two in [8, 128]");
    if (c.in_channels % 8 != 0) return fail(HFG_ERR_UNSUPPORTED, "config: in_channels must be a multiple of 8");
    return HFG_OK;
}

void add_layer(hfg_engine* e, Layer L) {
    if (L.is_post) {
        L.taps = L.k; L.tap_off0 = -L.pad; L.tap_step = 1; L.Np = 1;
    } else if (!L.transposed) {
        L.taps = L.k; L.tap_off0 = -L.pad; L.tap_step = L.dil; L.Np = L.cout; L.ups_s = 1; L.ups_p = 0;
    } else {
        L.taps = (L.k + L.stride - 1) / L.stride; L.tap_off0 = 0; L.tap_step = -1;
        L.Np = L.stride * L.cout; L.ups_s = L.stride; L.ups_p = L.pad;
    }
    // Tensor-core family: stages narrower than 32 channels (V2's tail) are carried with 32 channels; the extra channels have
    // zero weights and zero bias, so they hold lrelu(0) = 0 everywhere and change nothing.
    L.kc = (L.cin % 64 == 0 || L.cin > 64) ? 64 : 32;
    L.cin_pad = std::max(32, (L.cin + L.kc - 1) / L.kc * L.kc);
    L.cout_tc = L.is_post ? 1 : std::max(32, L.cout);
    L.Np_tc = L.transposed ? L.stride * L.cout_tc : L.cout_tc;
    e->index[L.name] = (int)e->layers.size();
    e->layers.push_back(std::move(L));
}

void build_layers(hfg_engine* e) {
    const hfg_config& c = e->cfg;
    const int c0 = c.upsample_initial_channel;
    char buf[64];
    {   // conv_pre  hifigan_pretrained.py:92-94
        Layer L; L.name = "conv_pre"; L.cin = c.in_channels; L.cout = c0; L.k = 7; L.pad = 3;
        add_layer(e, L);
    }
    int hop = 1;
    for (int i = 0; i < c.num_upsamples; ++i) {   // :98-109
        Layer L; snprintf(buf, sizeof buf, "ups.%d", i); L.name = buf; L.transposed = true;
        L.cin = c0 >> i; L.cout = c0 >> (i + 1); L.k = c.upsample_kernel_sizes[i]; L.stride = c.upsample_rates[i];
        L.pad = (L.k - L.stride) / 2;
        hop *= L.stride;
        add_layer(e, L);
    }
    e->hop = hop;
    int n = 0;
    for (int i = 0; i < c.num_upsamples; ++i) {   // :111-116, ResBlock :41-59
        const int ch = c0 >> (i + 1);
        for (int j = 0; j < c.num_kernels; ++j, ++n) {
            const int k = c.resblock_kernel_sizes[j];
            for (int m = 0; m < c.num_dilations[j]; ++m) {
                const int d = c.resblock_dilations[j][m];
                Layer a; snprintf(buf, sizeof buf, "resblocks.%d.convs1.%d", n, m); a.name = buf;
                a.cin = a.cout = ch; a.k = k; a.dil = d; a.pad = get_padding(k, d);
                add_layer(e, a);
                Layer b; snprintf(buf, sizeof buf, "resblocks.%d.convs2.%d", n, m); b.name = buf;
                b.cin = b.cout = ch; b.k = k; b.dil = 1; b.pad = get_padding(k, 1);
                add_layer(e, b);
            }
        }
    }
    {   // conv_post :119
        Layer L; L.name = "conv_post"; L.cin = c0 >> c.num_upsamples; L.cout = 1; L.k = 7; L.pad = 3; L.is_post = true;
        add_layer(e, L);
    }
}

Layer* find_layer(hfg_engine* e, const char* name) {
    if (!name) return nullptr;
    auto it = e->index.find(name);
    return it == e->index.end() ? nullptr : &e->layers[it->second];
}

void free_layer_dev(Layer& L) {
    cudaFree(L.d_w32); cudaFree(L.d_bias); cudaFree(L.d_wb_hi); cudaFree(L.d_wb_lo); cudaFree(L.d_bias_tc); cudaFree(L.d_w32_tc); cudaFree(L.d_wh);
    L.d_w32 = L.d_bias = L.d_bias_tc = L.d_w32_tc = nullptr; L.d_wb_hi = L.d_wb_lo = L.d_wh = nullptr;
}

inline uint16_t f2bf(float f) {   // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
inline uint16_t f2h(float f) {   // fp32 -> fp16, round-to-nearest-even, saturating to +-65504 (like cvt.rn.satfinite.f16.f32)
    uint32_t u; memcpy(&u, &f, 4);
    const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
    const uint32_t ax = u & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          // NaN
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);         // >= 65520 rounds past the largest finite value: saturate
    if (ax < 0x33000001u) return sign;                                 // <= 2^-25: rounds to zero
    int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;                         // 24-bit significand
    int shift = e < -14 ? 13 + (-14 - e) : 13;                         // subnormal results lose extra bits
    uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
    uint32_t h = e < -14 ? q : (((uint32_t)(e + 15) << 10) + (q - 0x400u));   // a carry out of the significand bumps the exponent
    return (uint16_t)(sign | h);
}

int upload_layer(Layer& L) {
    free_layer_dev(L);
    const int Cin = L.cin, Cout = L.cout, k = L.k;
    CK(cudaMalloc(&L.d_bias, sizeof(float) * Cout));
    CK(cudaMemcpy(L.d_bias, L.bias.data(), sizeof(float) * Cout, cudaMemcpyHostToDevice));
    {
        std::vector<float> bt((size_t)L.cout_tc, 0.f);
        std::copy(L.bias.begin(), L.bias.end(), bt.begin());
        CK(cudaMalloc(&L.d_bias_tc, bt.size() * sizeof(float)));
        CK(cudaMemcpy(L.d_bias_tc, bt.data(), bt.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (L.is_post) {
        std::vector<float> pt((size_t)k * L.cin_pad, 0.f);
        for (int j = 0; j < k; ++j)
            for (int ci = 0; ci < Cin; ++ci) pt[(size_t)j * L.cin_pad + ci] = L.w[(size_t)ci * k + j];
        CK(cudaMalloc(&L.d_w32_tc, pt.size() * sizeof(float)));
        CK(cudaMemcpy(L.d_w32_tc, pt.data(), pt.size() * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<float> p((size_t)k * Cin);
        for (int j = 0; j < k; ++j)
            for (int ci = 0; ci < Cin; ++ci) p[(size_t)j * Cin + ci] = L.w[(size_t)ci * k + j];   // w[0][ci][j]
        CK(cudaMalloc(&L.d_w32, p.size() * sizeof(float)));
        CK(cudaMemcpy(L.d_w32, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
        return HFG_OK;
    }
    // generic GEMM form  W'[j][ci][n]
    std::vector<float> p((size_t)L.taps * Cin * L.Np, 0.f);
    if (!L.transposed) {
        for (int co = 0; co < Cout; ++co)
            for (int ci = 0; ci < Cin; ++ci)
                for (int j = 0; j < k; ++j) p[((size_t)j * Cin + ci) * L.Np + co] = L.w[((size_t)co * Cin + ci) * k + j];
    } else {
        const int s = L.stride;
        for (int ci = 0; ci < Cin; ++ci)
            for (int co = 0; co < Cout; ++co)
                for (int kk = 0; kk < k; ++kk) {
                    const int n = kk / s, r = kk % s;
                    p[((size_t)n * Cin + ci) * L.Np + r * Cout + co] = L.w[((size_t)ci * Cout + co) * k + kk];
                }
    }
    CK(cudaMalloc(&L.d_w32, p.size() * sizeof(float)));
    CK(cudaMemcpy(L.d_w32, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
    // tensor-core pack: [taps*Np_tc][cin_pad] bf16 hi/lo, K-major (rows of padded output channels and columns of padded
    // input channels stay zero)
    std::vector<uint16_t> hi((size_t)L.taps * L.Np_tc * L.cin_pad, 0), lo(hi.size(), 0), hf(hi.size(), 0);
    for (int j = 0; j < L.taps; ++j)
        for (int ci = 0; ci < Cin; ++ci)
            for (int n = 0; n < L.Np; ++n) {
                const float v = p[((size_t)j * Cin + ci) * L.Np + n];
                const uint16_t h = f2bf(v);
                const int n_tc = (n / Cout) * L.cout_tc + (n % Cout);   // (phase, co) -> padded column
                const size_t o = ((size_t)j * L.Np_tc + n_tc) * L.cin_pad + ci;
                hi[o] = h;
                lo[o] = f2bf(v - bf2f(h));
                hf[o] = f2h(v);
            }
    CK(cudaMalloc(&L.d_wb_hi, hi.size() * 2));
    CK(cudaMalloc(&L.d_wb_lo, lo.size() * 2));
    CK(cudaMalloc(&L.d_wh, hf.size() * 2));
    CK(cudaMemcpy(L.d_wb_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.d_wb_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.d_wh, hf.data(), hf.size() * 2, cudaMemcpyHostToDevice));
    return HFG_OK;
}

// ---------------------------------------------------------------------------
// Time folding of the narrow stages (V2's 16- and 8-channel tail)
//
// A channels-last plane [L][C] with C < 32 is, byte for byte, the plane [L/f][f*C] with f = 32/C: f consecutive time steps
// form one 32-channel "super-row".  A 'same' Conv1d(C -> C, k, dilation d) on the time axis becomes a 'same' Conv1d(32 -> 32,
// k' = 2*ceil(pad/f) + 1, dilation 1) on super-rows whose weight blocks are the original taps placed by time parity:
//
//   out[f*S + eo] = sum_j w_j . in[f*S + eo + j*d - pad]      with  eo + j*d - pad = f*sigma + ei
//   =>  W'[sigma][eo*C + co][ei*C + ci] = w[co][ci][j]
//
// and the stride-s ConvTranspose1d into such a stage (f_out = s * f_in) becomes a plain conv on super-rows as well
// (kk = eo - f_out*sigma - s*ei + pad).  The narrow stages therefore run on the same tensor-core kernels as C = 32 -- on
// DENSE planes (no zero channels in HBM: 2x / 4x fewer bytes than carrying them padded to 32) and with 2-4x fewer MMA rows.
// leaky_relu, the residual add and the MRF mean are elementwise and do not see the folding.
// ---------------------------------------------------------------------------
int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

int upload_layer(Layer& L);

int build_folded(hfg_engine* e) {
    for (auto& kv : e->folded) free_layer_dev(kv.second);
    e->folded.clear();
    const hfg_config& c = e->cfg;
    const int c0 = c.upsample_initial_channel, NU = c.num_upsamples;
    for (int i = 0; i < NU; ++i) e->stage_fold[i] = 1;
    if (env_flag_early("HFG_FOLD", 1) == 0 || c0 % 32 != 0) return HFG_OK;
    // every narrow stage must fold consistently with the stage before it: f_i == f_{i-1} * rate_i
    bool ok = true, any = false;
    int fold[HFG_MAX_UPSAMPLES];
    for (int i = 0; i < NU; ++i) {
        const int ch = c0 >> (i + 1);
        fold[i] = (ch < 32 && 32 % ch == 0) ? 32 / ch : 1;
        if (ch < 32 && 32 % ch != 0) ok = false;
        const int prev = i > 0 ? fold[i - 1] : 1;
        if (fold[i] > 1) { any = true; if (fold[i] != prev * c.upsample_rates[i]) ok = false; }
        else if (prev != 1) ok = false;
    }
    if (!ok || !any) return HFG_OK;
    for (int i = 0; i < NU; ++i) e->stage_fold[i] = fold[i];

    auto make = [&](const Layer& o, int f_in, int f_out) -> int {
        Layer F;
        F.name = o.name; F.transposed = false; F.cin = f_in * o.cin; F.cout = f_out * o.cout; F.dil = 1; F.stride = 1; F.set = true;
        const int Kin = F.cin, Nout = F.cout;
        int smin = 1 << 30, smax = -(1 << 30);
        struct Item { int sigma, eo, ei, j; };
        std::vector<Item> items;
        if (!o.transposed) {
            for (int j = 0; j < o.k; ++j)
                for (int eo = 0; eo < f_out; ++eo) {
                    const int q = eo + j * o.dil - o.pad;
                    const int sg = floor_div(q, f_out);
                    items.push_back({sg, eo, q - sg * f_out, j});
                }
        } else {
            for (int kk = 0; kk < o.k; ++kk)
                for (int eo = 0; eo < f_out; ++eo)
                    for (int ei = 0; ei < f_in; ++ei) {
                        const int num = eo - o.stride * ei + o.pad - kk;
                        if (((num % f_out) + f_out) % f_out != 0) continue;
                        items.push_back({floor_div(num, f_out), eo, ei, kk});
                    }
        }
        for (const Item& it : items) { smin = std::min(smin, it.sigma); smax = std::max(smax, it.sigma); }
        const int taps = smax - smin + 1;
        F.k = taps; F.pad = -smin;
        F.taps = taps; F.tap_off0 = smin; F.tap_step = 1; F.Np = Nout; F.ups_s = 1; F.ups_p = 0;
        F.kc = 32; F.cin_pad = Kin; F.cout_tc = Nout; F.Np_tc = Nout;
        F.w.assign((size_t)Nout * Kin * taps, 0.f);   // torch Conv1d layout [cout][cin][k]
        for (const Item& it : items)
            for (int co = 0; co < o.cout; ++co)
                for (int ci = 0; ci < o.cin; ++ci) {
                    const float v = o.transposed ? o.w[((size_t)ci * o.cout + co) * o.k + it.j] : o.w[((size_t)co * o.cin + ci) * o.k + it.j];
                    F.w[((size_t)(it.eo * o.cout + co) * Kin + (it.ei * o.cin + ci)) * taps + (it.sigma - smin)] = v;
                }
        F.bias.resize(Nout);
        for (int n = 0; n < Nout; ++n) F.bias[n] = o.bias[n % o.cout];
        RET(upload_layer(F));
        e->folded[o.name] = std::move(F);
        return HFG_OK;
    };
    char buf[64];
    int n = 0;
    for (int i = 0; i < NU; ++i) {
        const int f = fold[i], f_prev = i > 0 ? fold[i - 1] : 1;
        if (f > 1) {
            snprintf(buf, sizeof buf, "ups.%d", i);
            RET(make(e->layers[e->index[buf]], f_prev, f));
        }
        for (int j = 0; j < c.num_kernels; ++j, ++n)
            for (int m = 0; m < c.num_dilations[j] && f > 1; ++m) {
                snprintf(buf, sizeof buf, "resblocks.%d.convs1.%d", n, m);
                RET(make(e->layers[e->index[buf]], f, f));
                snprintf(buf, sizeof buf, "resblocks.%d.convs2.%d", n, m);
                RET(make(e->layers[e->index[buf]], f, f));
            }
    }
    return HFG_OK;
}

ConvGeom geom_of(const Layer& L, int B, int Lin) {
    ConvGeom g;
    g.B = B; g.Lin = Lin; g.Cin = L.cin; g.Cout = L.cout;
    g.Lout = L.transposed ? (Lin - 1) * L.stride - 2 * L.pad + L.k : Lin;
    g.Mrows = L.transposed ? Lin + L.taps - 1 : Lin;
    g.Np = L.Np; g.taps = L.taps; g.tap_off0 = L.tap_off0; g.tap_step = L.tap_step;
    g.ups_s = L.ups_s; g.ups_p = L.ups_p;
    return g;
}

// Geometry of the same layer on the tensor-core family's (channel-padded) planes.
ConvGeom geom_tc(const Layer& L, int B, int Lin) {
    ConvGeom g = geom_of(L, B, Lin);
    g.Cin = L.cin_pad; g.Cout = L.cout_tc; g.Np = L.Np_tc;
    return g;
}

ConvParams conv32(const Layer& L, int B, int Lin, const float* x, float* y, const float* res, int pre_lrelu,
                  int accumulate, float out_div) {
    const ConvGeom g = geom_of(L, B, Lin);
    ConvParams p;
    memset(&p, 0, sizeof p);
    p.x = x; p.w = L.d_w32; p.bias = L.d_bias; p.y = y; p.res = res;
    p.B = B; p.Lin = Lin; p.Cin = L.cin; p.Lout = g.Lout; p.Cout = L.cout; p.Mrows = g.Mrows; p.Np = g.Np;
    p.taps = g.taps; p.tap_off0 = g.tap_off0; p.tap_step = g.tap_step; p.ups_s = g.ups_s; p.ups_p = g.ups_p;
    p.pre_lrelu = pre_lrelu; p.accumulate = accumulate; p.out_div = out_div;
    return p;
}

// ---------------------------------------------------------------------------
// Workspace arena (bump allocator, 256-byte aligned)
// ---------------------------------------------------------------------------
struct Bump {
    uint8_t* base;
    size_t off = 0;
    size_t guard = 0;                      // HFG_GUARD: bytes of canary after every buffer (compute-sanitizer is not available here)
    std::vector<size_t>* guards = nullptr;  // offsets of the canary regions
    template <typename T>
    T* take(size_t count) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += (count * sizeof(T) + 255) / 256 * 256;
        if (guard) {
            if (guards && base) guards->push_back(off);
            off += guard;
        }
        return p;
    }
};

size_t stage_elems_max(const hfg_engine* e, int B, int T, int min_ch = 1) {
    size_t mx = 0;
    size_t L = (size_t)T;
    for (int i = 0; i < e->cfg.num_upsamples; ++i) {
        L *= e->cfg.upsample_rates[i];
        mx = std::max(mx, (size_t)B * L * std::max(min_ch, e->cfg.upsample_initial_channel >> (i + 1)));
    }
    return mx;
}

constexpr size_t kGuardBytes = 4096;
constexpr int kGuardByte = 0xA5;

int env_flag(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

// Builds (or, with base == nullptr, only sizes) the launch plan.
//
// fp32 family: raw fp32 streams everywhere, leaky_relu fused into the consumer's load.
// tensor-core family: activations exist ONLY as activated bf16 planes P(x) = bf16(lrelu(x)) (+ a lo plane in
// bf16x3): every conv reads them through TMA, writes them from its epilogue, and the residual add (:70)
// recovers x by inverting leaky_relu.  The MRF mean (:133-137) is one elementwise pass over the nk branch
// outputs per stage.  Stages narrower than 32 channels (V2's tail) run on the fp32 family inside a
// tensor-core plan; the hand-over is the fp32 output of that MRF pass.
//
// ragged (hfg_forward_ragged): the items of the batch have their own lengths (plan->lens_dev, mel frames).
// Every layer of the reference zero-pads its input at the true end of the sequence (:49-59, :92-94), which a dense plan gets from
// TMA's out-of-bounds zero fill at row L.  A ragged plan makes the same true at every item's OWN end: the kRagged instantiations
// of conv_umma2 and conv_pair write zeros for every output row at or behind it (the pair kernel also for its shared-memory
// intermediate), and the few layers on the first-generation kernel (conv_pre, upsamplers wider than one tile) and the staged mel
// are followed by an S_STRIP step that zeroes the `halo` rows behind the end (halo >= the widest tap span of any layer, so no
// row inside an item ever reads further).  Rows are independent dot products and their bits do not depend on the tile they fall
// in, so an item's samples equal those of a dense forward of that item alone, bit for bit.  The fp32 family (HFG_PREC_FP32, and
// generators whose initial channel count is not a multiple of 32) gets the zero-fill step after every conv.
int build_plan(hfg_engine* e, int B, int T, int prec, bool keep_taps, uint8_t* base, Plan* plan, size_t* bytes_out, bool ragged = false) {
    typedef __nv_bfloat16 bf;
    const hfg_config& c = e->cfg;
    const bool x3 = prec == HFG_PREC_BF16X3;
    const int f16 = prec == HFG_PREC_FP16 ? 1 : 0;   // single-plane mode with fp16 operand planes / weights instead of bf16
    const int npass = x3 ? 3 : 1;
    auto wts_hi = [&](const Layer& L) { return f16 ? L.d_wh : L.d_wb_hi; };
    const int a_per_tap = env_flag("HFG_UMMA_A_PER_TAP", 0);
    // MRF sum folded into the producers (:133-137), HFG_MRF_FOLD=1: the last convs2 of branch j >= 1 adds the running sum of
    // branches 0..j-1 in its epilogue, the last branch also applies 1/nk and writes the stage output -- no separate combine
    // pass.  OFF by default: measured on B200 (profiles/r02_ab_mrf_fold_graph.md) the extra plane stream costs the consuming
    // epilogues what the removed passes saved (bf16 B=16: 9.20 vs 9.19 ms; bf16x3: the two-plane staging slots double and the
    // C = 64 convs2 lose 0.26-0.38 ms each).  The tap plan always keeps the separate pass (it must expose every branch output).
    const bool mrf_fold = env_flag("HFG_MRF_FOLD", 0) != 0 && !keep_taps && c.num_kernels > 1;
    const int snake = env_flag("HFG_SNAKE", 1);   // alternate the tile direction of consecutive convs (L2 reuse)
    int n_umma2 = 0;
    const int c0 = c.upsample_initial_channel;
    const int NU = c.num_upsamples;
    const int nk = c.num_kernels;
    // number of leading stages (after conv_pre) that run on tensor cores; -1: conv_pre is fp32 too
    // (every stage: narrow stages are carried with 32 channels, see add_layer)
    int n_tc = -1;
    if (prec != HFG_PREC_FP32 && c0 % 32 == 0) n_tc = NU;
    const bool any_tc = n_tc >= 0;
    const bool any_32 = n_tc < NU;   // some stage (or everything) runs on the fp32 family

    Bump bump{base};
    if (env_flag("HFG_GUARD", 0)) { bump.guard = kGuardBytes; bump.guards = plan ? &plan->guards : nullptr; }
    const size_t smax = stage_elems_max(e, B, T);
    const size_t smax_tc = stage_elems_max(e, B, T, 32);   // planes of the tensor-core family
    const size_t n_pre = (size_t)B * T * c0;
    const Layer& pre = e->layers[e->index["conv_pre"]];
    const Layer& post = e->layers[e->index["conv_post"]];
    const bool real = base != nullptr;
    auto push = [&](Step&& s) { if (real) plan->steps.push_back(std::move(s)); };
    auto tap = [&](const char* name, const float* src, int C, int L) {
        if (!keep_taps) return;
        Step s{}; s.kind = S_TAP; s.tap_name = name; s.f_in = src; s.B = B; s.C = C; s.L = L;
        push(std::move(s));
    };
    char nm[64];
    auto layer = [&](const char* fmt, int a, int b2) -> const Layer& {
        snprintf(nm, sizeof nm, fmt, a, b2);
        return e->layers[e->index[nm]];
    };

    float* mel_dev = bump.take<float>((size_t)B * c.in_channels * T);
    float* wave_dev = bump.take<float>((size_t)B * T * e->hop);
    int32_t* lens_dev = ragged ? bump.take<int32_t>((size_t)B) : nullptr;
    if (real) { plan->mel_dev = mel_dev; plan->wave_dev = wave_dev; plan->keep_taps = keep_taps; plan->lens_dev = lens_dev; }
    if (ragged && (keep_taps || (any_tc && any_32)))
        return fail(HFG_ERR_UNSUPPORTED, "ragged batches: not available with taps");
    // rows to zero behind an item's end: the widest reach of any layer's taps (in the rows of its own stage), in whole 8-row groups
    int halo_rows = 8;
    for (const Layer& Ly : e->layers) halo_rows = std::max(halo_rows, Ly.transposed ? Ly.k : Ly.dil * (Ly.k - 1) / 2 + 1);
    halo_rows = (halo_rows + 7) / 8 * 8;
    bool strip_bad = false;

    // fp32 streams: the fp32 family's working set, and the tap staging buffer of the tensor-core family
    float* u_raw = any_32 ? bump.take<float>(smax) : nullptr;
    float* r_raw = any_32 ? bump.take<float>(smax) : nullptr;
    float* xs = (any_32 || keep_taps) ? bump.take<float>(smax) : nullptr;
    float* xt32 = any_32 ? bump.take<float>(smax) : nullptr;
    float* tap_tmp = (any_tc && keep_taps) ? bump.take<float>(std::max(smax, n_pre)) : nullptr;
    float* post_tap = keep_taps ? bump.take<float>((size_t)B * T * e->hop) : nullptr;
    float* mel_cl = !any_tc ? bump.take<float>((size_t)B * T * c.in_channels) : nullptr;
    float* x0_raw = n_tc <= 0 ? bump.take<float>(n_pre) : nullptr;
    // tensor-core operand planes
    struct Planes { bf* hi = nullptr; bf* lo = nullptr; };
    auto take_planes = [&](size_t n) { Planes p; p.hi = bump.take<bf>(n); if (x3) p.lo = bump.take<bf>(n); return p; };
    // Concurrent branch lanes for stages one kernel cannot fill: fewer than HFG_BRANCH_PAR_TILES (default 4) x SM-count 128-row
    // tiles.  HFG_BRANCH_PAR = 0: never, 1: every stage, default -1: by that size rule.
    const int par_mode = env_flag("HFG_BRANCH_PAR", -1);
    const double par_tiles = (double)env_flag("HFG_BRANCH_PAR_TILES", 4) * e->sm_count;
    bool par_stage[HFG_MAX_UPSAMPLES] = {};
    bool par_any = false;
    size_t par_elems = 0;
    {
        size_t Ls = (size_t)T;
        for (int i = 0; i < NU; ++i) {
            Ls *= c.upsample_rates[i];
            const double tiles = (double)B * (double)Ls / 128.0;
            par_stage[i] = any_tc && i < n_tc && nk > 1 && !keep_taps && !mrf_fold && par_mode != 0 && (par_mode == 1 || tiles < par_tiles);
            if (par_stage[i]) {
                par_any = true;
                par_elems = std::max(par_elems, (size_t)B * Ls * std::max(32, c0 >> (i + 1)));
            }
        }
    }
    Planes mel_p, x0_p, u_p, xt_p, pp[2], s_p, r_p[HFG_MAX_KERNELS];
    Planes xt_b[HFG_MAX_KERNELS], pp_b[HFG_MAX_KERNELS][2];   // per-branch temporaries of the concurrent lanes (branch 0 uses xt_p / pp)
    if (any_tc) {
        mel_p = take_planes((size_t)B * T * pre.cin_pad);
        x0_p = take_planes(n_pre);
        if (n_tc > 0) {
            u_p = take_planes(smax_tc); xt_p = take_planes(smax_tc); pp[0] = take_planes(smax_tc); pp[1] = take_planes(smax_tc);
            s_p = take_planes(smax_tc);
            for (int j = 0; j < nk; ++j) r_p[j] = take_planes(smax_tc);
            xt_b[0] = xt_p; pp_b[0][0] = pp[0]; pp_b[0][1] = pp[1];
            for (int j = 1; j < nk && par_any; ++j) {
                xt_b[j] = take_planes(par_elems); pp_b[j][0] = take_planes(par_elems); pp_b[j][1] = take_planes(par_elems);
            }
        }
    }
    int cur_lane = 0;   // lane of the steps being emitted
    // ragged plans: zero the rows behind each item's end of planes [B][rows][row_elems] (rows = mul * T)
    auto strip = [&](__nv_bfloat16* hi, __nv_bfloat16* lo, int rows, int row_elems) {
        if (!ragged || !real) return;
        if (rows % T != 0 || row_elems % 8 != 0) { strip_bad = true; return; }
        Step s{}; s.kind = S_STRIP; s.b_out = hi; s.b_out_lo = x3 ? lo : nullptr; s.B = B; s.L = rows; s.C = row_elems; s.k = rows / T;
        s.flag0 = halo_rows; s.lane = cur_lane; s.label = "zero_tail";
        plan->steps.push_back(std::move(s));
    };

    // F_l = 2*Cin*Cout*k*L (L = L_out for Conv1d, L_in for ConvTranspose1d); Q_l = activations in + out + weights
    auto work = [&](Step& s, const Layer& L, int Lin, int act_bytes) {
        const double Lout = L.transposed ? (double)Lin * L.stride : (double)Lin;
        s.label = L.name;
        s.flops = 2.0 * L.cin * L.cout * L.k * (L.transposed ? (double)Lin : Lout) * B;
        s.bytes = ((double)L.cin * Lin + (double)L.cout * Lout) * B * act_bytes + (double)L.cin * L.cout * L.k * act_bytes;
    };
    // one conv on planes: persistent pipelined kernel where it applies, the v1 kernel otherwise
    auto umma = [&](const Layer& L, int Lin, Planes x, Planes res, float* y_raw, Planes y, const Layer* acct = nullptr,
                    int acct_Lin = 0, Planes mrf = Planes(), float out_scale = 1.f) -> int {
        if (!real) return HFG_OK;
        Step s{};
        UmmaConvParams p;
        memset(&p, 0, sizeof p);
        p.g = geom_tc(L, B, Lin);
        p.cin_pad = L.cin_pad; p.kc = L.kc; p.npass = npass; p.f16 = f16;
        p.bias = L.d_bias_tc; p.res_hi = res.hi; p.res_lo = x3 ? res.lo : nullptr;
        p.mrf_hi = mrf.hi; p.mrf_lo = x3 ? mrf.lo : nullptr; p.out_scale = out_scale;
        p.y_raw = y_raw; p.y_act = y.hi; p.y_act_lo = x3 ? y.lo : nullptr;
        p.a_per_tap = a_per_tap;
        p.reverse = snake ? (n_umma2++ & 1) : 0;
        if (ragged) { p.lens = lens_dev; p.len_T = T; p.len_skip = halo_rows; }
        if (acct) work(s, *acct, acct_Lin, x3 ? 4 : 2);   // a time-folded twin: report the reference layer's algorithmic work
        else work(s, L, Lin, x3 ? 4 : 2);   // two bf16 planes carry what an fp32 activation would
        if (umma2_supported(p) && plan_conv_umma2(&s.u2, p, x.hi, x.lo, wts_hi(L), L.d_wb_lo, e->sm_count) == HFG_OK) {
            s.kind = S_UMMA2;
        } else {
            s.kind = S_UMMA;
            RET(plan_conv_umma(&s.ul, p, x.hi, x.lo, wts_hi(L), L.d_wb_lo));
        }
        s.lane = cur_lane;
        const bool masked = s.kind == S_UMMA2;   // conv_umma2 writes the zeros behind every item's end itself
        plan->steps.push_back(std::move(s));
        if (y.hi && !masked) strip(y.hi, y.lo, p.g.Lout, p.g.Cout);
        return HFG_OK;
    };
    // convs1[m] -> lrelu -> convs2[m] -> + x  (:66-70) as one launch where the fused kernel applies (C <= 64); false otherwise
    auto pair = [&](const Layer& c1, const Layer& c2, int Lrows, Planes x, Planes y, char const* label, const Layer* a1 = nullptr,
                    const Layer* a2 = nullptr, int acct_L = 0, Planes mrf = Planes(), float out_scale = 1.f) -> bool {
        if (!real) return false;
        if (c1.cin_pad != c1.cout_tc || c2.cin_pad != c1.cin_pad || c2.cout_tc != c1.cout_tc || c2.dil != 1) return false;
        PairParams p;
        memset(&p, 0, sizeof p);
        p.B = B; p.L = Lrows; p.C = c1.cin_pad; p.k1 = c1.k; p.k2 = c2.k; p.d = c1.dil; p.npass = npass; p.f16 = f16;
        p.bias1 = c1.d_bias_tc; p.bias2 = c2.d_bias_tc;
        p.x_hi = x.hi; p.x_lo = x3 ? x.lo : nullptr;
        p.w1_hi = wts_hi(c1); p.w1_lo = c1.d_wb_lo; p.w2_hi = wts_hi(c2); p.w2_lo = c2.d_wb_lo;
        p.y_hi = y.hi; p.y_lo = x3 ? y.lo : nullptr;
        p.mrf_hi = mrf.hi; p.mrf_lo = x3 ? mrf.lo : nullptr; p.out_scale = out_scale;
        p.reverse = snake ? (n_umma2 & 1) : 0;
        if (ragged) {
            if (Lrows % T != 0) { strip_bad = true; return false; }
            p.lens = lens_dev; p.len_mul = Lrows / T; p.len_skip = halo_rows;
        }
        if (!pair_supported(p)) return false;
        Step s{};
        if (plan_conv_pair(&s.pl, p, e->sm_count) != HFG_OK) return false;
        ++n_umma2;
        Step w2{};
        const Layer& r1 = a1 ? *a1 : c1;   // reference layers for the work model (time-folded twins report the originals)
        const Layer& r2 = a2 ? *a2 : c2;
        const int rL = a1 ? acct_L : Lrows;
        work(s, r1, rL, x3 ? 4 : 2);
        work(w2, r2, rL, x3 ? 4 : 2);
        // algorithmic bytes of the FUSED step: x in, out, both weight sets (the intermediate never leaves the SM)
        const double act_b = x3 ? 4.0 : 2.0;
        s.kind = S_PAIR; s.label = label; s.flops += w2.flops;
        s.bytes = 2.0 * r1.cin * (double)rL * B * act_b + 2.0 * (double)r1.cin * r1.cout * r1.k * act_b;
        s.lane = cur_lane;
        plan->steps.push_back(std::move(s));   // (both epilogues of the pair kernel mask per item: no zero-fill step)
        return true;
    };
    auto c32 = [&](const Layer& L, int Lin, const float* x, float* y, const float* res, int pre_lrelu, int accumulate, float out_div) {
        Step s{}; s.kind = S_CONV32; s.cp = conv32(L, B, Lin, x, y, res, pre_lrelu, accumulate, out_div);
        work(s, L, Lin, 4);
        const int Lout = s.cp.Lout;
        push(std::move(s));
        // ragged plans of the fp32 family: the same zero-fill behind every item's end, on the fp32 stream (two 16-bit elements each)
        strip(reinterpret_cast<__nv_bfloat16*>(y), nullptr, Lout, 2 * L.cout);
    };
    auto accum = [&](const float* r, size_t ne, int j) {
        Step s{}; s.kind = S_ACCUM; s.f_out = xs; s.f_in = r; s.n = ne; s.flag0 = j == 0;
        s.fval = j == nk - 1 ? (float)nk : 0.f;
        push(std::move(s));
    };
    auto to_raw = [&](Planes p, float* raw, size_t rows, int C_tc, int C) {   // [rows][C_tc] planes -> [rows][C] fp32
        Step s{}; s.kind = S_P2RAW; s.b_in = p.hi; s.b_in_lo = x3 ? p.lo : nullptr; s.f_out = raw; s.n = rows; s.cpad = C_tc; s.C = C;
        s.f16 = f16;
        push(std::move(s));
    };
    auto tap_planes = [&](const char* name, Planes p, int C, int L, bool dense = false) {   // dense: time-folded stage, no padding channels
        if (!keep_taps) return;
        to_raw(p, tap_tmp, (size_t)B * L, dense ? C : std::max(32, C), C);
        tap(name, tap_tmp, C, L);
    };

    // ---- conv_pre  (:124) ----
    const float* x_raw = nullptr;   // raw fp32 input of the next upsampler (fp32 family)
    Planes xp;                      // activated planes (tensor-core family)
    if (any_tc) {
        { Step s{}; s.kind = S_MEL_CLBF; s.f_in = mel_dev; s.b_out = mel_p.hi; s.b_out_lo = mel_p.lo; s.B = B; s.C = c.in_channels;
          s.L = T; s.cpad = pre.cin_pad; s.f16 = f16; push(std::move(s)); }
        strip(mel_p.hi, mel_p.lo, T, pre.cin_pad);   // whatever the caller left behind an item's end is not part of it
        RET(umma(pre, T, mel_p, Planes(), x0_raw, n_tc > 0 ? x0_p : Planes()));
        if (n_tc > 0) tap_planes("conv_pre", x0_p, c0, T); else tap("conv_pre", x0_raw, c0, T);
        x_raw = x0_raw; xp = x0_p;
    } else {
        { Step s{}; s.kind = S_MEL_CL32; s.f_in = mel_dev; s.f_out = mel_cl; s.B = B; s.C = c.in_channels; s.L = T; push(std::move(s)); }
        strip(reinterpret_cast<__nv_bfloat16*>(mel_cl), nullptr, T, 2 * c.in_channels);
        c32(pre, T, mel_cl, x0_raw, nullptr, 0, 0, 0.f);
        tap("conv_pre", x0_raw, c0, T);
        x_raw = x0_raw;
    }

    int L = T, n = 0;
    for (int i = 0; i < NU; ++i) {
        const Layer& up = layer("ups.%d", i, 0);
        const int ch = up.cout;
        const int Lin = L;
        L *= up.stride;
        const bool tc_stage = i < n_tc;
        // time-folded stage (C < 32): dense planes [L][ch] read as [L/f][32] by the folded twins of its layers (build_folded)
        const int f = tc_stage ? e->stage_fold[i] : 1;
        const int f_prev = (tc_stage && i > 0) ? e->stage_fold[i - 1] : 1;
        const bool fold = f > 1;
        const int ctc = fold ? ch : up.cout_tc;          // channels per time step of this stage's planes
        const size_t ne = (size_t)B * L * (tc_stage ? ctc : ch);
        auto twin = [&](const Layer& o) -> const Layer& { return fold ? e->folded.at(o.name) : o; };
        if (tc_stage) {
            const bool last_stage = i == NU - 1;
            const bool next_tc = i + 1 < n_tc;
            const bool want_planes = next_tc || last_stage;   // conv_post reads planes after a tensor-core stage
            const bool want_raw = (!last_stage && !next_tc) || keep_taps;   // an fp32-family stage follows (or the stage tap)
            if (fold) RET(umma(twin(up), Lin / f_prev, xp, Planes(), nullptr, u_p, &up, Lin));
            else RET(umma(up, Lin, xp, Planes(), nullptr, u_p));   // lrelu -> ups  (:127-128)
            snprintf(nm, sizeof nm, "ups.%d", i);
            tap_planes(nm, u_p, ch, L, fold);
            const int rows = L / f;
            const bool par = par_stage[i];
            if (par) { Step s{}; s.kind = S_FORK; s.flag0 = nk; push(std::move(s)); }
            for (int j = 0; j < nk; ++j, ++n) {
                Planes xin = u_p;
                const int nd = c.num_dilations[j];
                cur_lane = par ? j : 0;
                const Planes xt_use = par ? xt_b[j] : xt_p;
                const Planes pp_use[2] = {par ? pp_b[j][0] : pp[0], par ? pp_b[j][1] : pp[1]};
                for (int m = 0; m < nd; ++m) {
                    const Layer& o1 = layer("resblocks.%d.convs1.%d", n, m);
                    const Layer& o2 = layer("resblocks.%d.convs2.%d", n, m);
                    const Layer& c1 = twin(o1);
                    const Layer& c2 = twin(o2);
                    const bool last_m = m == nd - 1;
                    // folded MRF sum: the last step of branch j >= 1 adds r_p[j-1] (the running sum); the last branch writes
                    // the stage output s_p = sum / nk
                    const bool fold_in = mrf_fold && last_m && j > 0;
                    const bool fold_out = mrf_fold && last_m && j == nk - 1;
                    Planes xout = fold_out ? s_p : (last_m ? r_p[j] : pp_use[m & 1]);
                    const Planes mrf_in = fold_in ? r_p[j - 1] : Planes();
                    const float osc = fold_out ? 1.0f / (float)nk : 1.0f;
                    snprintf(nm, sizeof nm, "resblocks.%d.pair.%d", n, m);
                    if (!pair(c1, c2, rows, xin, xout, nm, fold ? &o1 : nullptr, fold ? &o2 : nullptr, L, mrf_in, osc)) {
                        RET(umma(c1, rows, xin, Planes(), nullptr, xt_use, fold ? &o1 : nullptr, L));   // :66-67 (+ :68 in the epilogue)
                        RET(umma(c2, rows, xt_use, xin, nullptr, xout, fold ? &o2 : nullptr, L, mrf_in, osc));   // :69-70 (+ :133-137)
                    }
                    xin = xout;
                }
                snprintf(nm, sizeof nm, "resblocks.%d", n);
                tap_planes(nm, r_p[j], ch, L, fold);
            }
            cur_lane = 0;
            if (par) { Step s{}; s.kind = S_JOIN; s.flag0 = nk; push(std::move(s)); }
            const bool fuse_post = last_stage && !keep_taps;   // the MRF mean of the last stage is computed inside conv_post
            if (!fuse_post && !mrf_fold) {   // xs = sum_j r_j ; x = xs / nk  (:133-137)
                Step s{}; s.kind = S_MRF; s.n = ne;
                memset(&s.mrf, 0, sizeof s.mrf);
                s.mrf.f16 = f16;
                for (int j = 0; j < nk; ++j) { s.mrf.hi[j] = r_p[j].hi; s.mrf.lo[j] = x3 ? r_p[j].lo : nullptr; }
                s.mrf.nk = nk;
                s.mrf.out_raw = (want_raw && ctc == ch) ? xs : nullptr;
                s.mrf.out_hi = (want_planes || ctc != ch) ? s_p.hi : nullptr;
                s.mrf.out_lo = ((want_planes || ctc != ch) && x3) ? s_p.lo : nullptr;
                push(std::move(s));
                if (want_raw && ctc != ch) to_raw(s_p, xs, (size_t)B * L, ctc, ch);
            }
            snprintf(nm, sizeof nm, "stage.%d", i);
            tap(nm, xs, ch, L);
            xp = s_p; x_raw = xs;
        } else {
            c32(up, Lin, x_raw, u_raw, nullptr, 1, 0, 0.f);
            snprintf(nm, sizeof nm, "ups.%d", i);
            tap(nm, u_raw, ch, L);
            for (int j = 0; j < nk; ++j, ++n) {
                const float* r = u_raw;
                const int nd = c.num_dilations[j];
                for (int m = 0; m < nd; ++m) {
                    const Layer& c1 = layer("resblocks.%d.convs1.%d", n, m);
                    c32(c1, L, r, xt32, nullptr, 1, 0, 0.f);
                    const Layer& c2 = layer("resblocks.%d.convs2.%d", n, m);
                    const bool last = m == nd - 1;
                    if (!last || keep_taps) {
                        c32(c2, L, xt32, r_raw, r, 1, 0, 0.f);
                        r = r_raw;
                    } else {
                        c32(c2, L, xt32, xs, r, 1, j > 0, j == nk - 1 ? (float)nk : 0.f);
                    }
                }
                if (keep_taps) {
                    snprintf(nm, sizeof nm, "resblocks.%d", n);
                    tap(nm, r_raw, ch, L);
                    accum(r_raw, ne, j);
                }
            }
            snprintf(nm, sizeof nm, "stage.%d", i);
            tap(nm, xs, ch, L);
            x_raw = xs; xp = Planes();
        }
    }
    // ---- lrelu -> conv_post -> tanh  (:139-141) ----
    const bool post_planes = n_tc == NU;
    for (int pass = keep_taps ? 0 : 1; pass < 2; ++pass) {
        Step s{};
        s.w = post.d_w32; s.bias = post.d_bias; s.f_out = pass ? wave_dev : post_tap;
        s.B = B; s.L = L; s.C = post.cin; s.k = post.k; s.flag0 = 1; s.flag1 = pass;
        if (post_planes) {
            const bool dense = e->stage_fold[NU - 1] > 1;   // time-folded last stage: planes carry the real channels only
            s.w = dense ? post.d_w32 : post.d_w32_tc; s.C = dense ? post.cin : post.cin_pad;
            // One kernel for both plans (identical arithmetic order): the production plan hands it the nk branch outputs and
            // it forms the MRF mean itself; the tap plan hands it the already combined stage planes (nk = 1).
            s.kind = S_POSTMRF;
            memset(&s.mrf, 0, sizeof s.mrf);
            s.mrf.f16 = f16;
            if (keep_taps || mrf_fold) { s.mrf.hi[0] = xp.hi; s.mrf.lo[0] = x3 ? xp.lo : nullptr; s.mrf.nk = 1; }
            else { for (int j = 0; j < nk; ++j) { s.mrf.hi[j] = r_p[j].hi; s.mrf.lo[j] = x3 ? r_p[j].lo : nullptr; } s.mrf.nk = nk; }
        } else {
            s.kind = S_POST32; s.f_in = x_raw;
        }
        work(s, post, L, post_planes ? (x3 ? 4 : 2) : 4);
        push(std::move(s));
        if (!pass) tap("conv_post", post_tap, 1, L);
    }
    if (strip_bad) return fail(HFG_ERR_UNSUPPORTED, "ragged batches: a stage of this generator is not a whole number of rows per mel frame");
    if (bytes_out) *bytes_out = bump.off;
    return HFG_OK;
}

int store_tap(hfg_engine* e, const Step& s) {
    Tap& t = e->taps[s.tap_name];
    const size_t n = (size_t)s.B * s.C * s.L;
    if ((size_t)t.B * t.C * t.L != n) {
        cudaFree(t.dev);
        t.dev = nullptr;
        CK(cudaMalloc(&t.dev, n * sizeof(float)));
    }
    t.B = s.B; t.C = s.C; t.L = s.L;
    CK(cudaMemcpyAsync(t.dev, s.f_in, n * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
    return HFG_OK;
}

const char* kind_label(StepKind k) {
    switch (k) {
        case S_CONV32: return "conv_cl_fp32";
        case S_UMMA: return "conv_umma";
        case S_UMMA2: return "conv_umma2";
        case S_PAIR: return "conv_pair";
        case S_P2RAW: return "planes_to_raw";
        case S_MRF: return "mrf_combine";
        case S_POSTMRF: return "conv_post_mrf";
        case S_POST32: return "conv_post";
        case S_ACCUM: return "accum";
        case S_MEL_CL32: return "transpose";
        case S_MEL_CLBF: return "mel_to_cl_bf16";
        case S_STRIP: return "zero_tail_rows";
        default: return "copy";
    }
}

int run_plan(hfg_engine* e, Plan* plan);
inline bool is_kernel_step(StepKind k) { return k != S_TAP && k != S_FORK && k != S_JOIN; }

// Launches the plan's kernels: as ONE CUDA graph once the plan has run before (production path: no per-launch driver work
// on the host, which is what bounds short inputs), directly otherwise (first forward of a shape, profiling, taps, canaries, ncu).
int launch_plan(hfg_engine* e, Plan* plan) {
    static const int use_graph = env_flag("HFG_GRAPH", 1);
    const bool eligible = use_graph && !e->profiling && !plan->keep_taps && plan->guards.empty() && e->ncu_layers.empty();
    if (eligible && !plan->graph && !plan->graph_failed && plan->runs >= 1) {
        cudaGraph_t g = nullptr;
        const uint64_t l0 = e->launches;
        cudaError_t err = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
        if (err == cudaSuccess) {
            const int r = run_plan(e, plan);
            err = cudaStreamEndCapture(e->stream, &g);
            if (r != HFG_OK && err == cudaSuccess) err = cudaErrorUnknown;
        }
        if (err == cudaSuccess) err = cudaGraphInstantiate(&plan->graph, g, 0);
        if (g) cudaGraphDestroy(g);
        e->launches = l0;   // capture enqueued nothing
        if (err != cudaSuccess) {   // not fatal: keep launching kernel by kernel
            if (env_flag("HFG_GRAPH_VERBOSE", 0)) fprintf(stderr, "hfg: graph capture failed (%s); launching directly\n", cudaGetErrorString(err));
            cudaGetLastError();
            plan->graph = nullptr;
            plan->graph_failed = true;
            ++e->graphs_failed;
        } else {
            ++e->graphs_captured;
        }
    }
    ++plan->runs;
    if (eligible && plan->graph) {
        CK(cudaGraphLaunch(plan->graph, e->stream));
        e->launches += (uint64_t)plan->kernels;
        return HFG_OK;
    }
    return run_plan(e, plan);
}

int run_plan(hfg_engine* e, Plan* plan) {
    // per-launch profiling and ncu bracketing serialise everything on the engine's stream; otherwise branch lanes are honoured
    const bool lanes = !e->profiling && e->ncu_layers.empty();
    cudaStream_t st = e->stream;
    size_t nev = e->prof_used;
    if (e->profiling) {
        while (e->prof_events.size() < nev + plan->steps.size() + 1) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            e->prof_events.push_back(ev);
        }
        CK(cudaEventRecord(e->prof_events[nev++], st));
    }
    for (const Step& s : plan->steps) {
        if (s.kind == S_FORK || s.kind == S_JOIN) {
            if (!lanes) continue;
            for (int l = 1; l < s.flag0; ++l) {
                if (s.kind == S_FORK) {
                    if (l == 1) CK(cudaEventRecord(e->ev_fork, e->stream));
                    CK(cudaStreamWaitEvent(e->lane_stream[l], e->ev_fork, 0));
                } else {
                    CK(cudaEventRecord(e->ev_join[l], e->lane_stream[l]));
                    CK(cudaStreamWaitEvent(e->stream, e->ev_join[l], 0));
                }
            }
            continue;
        }
        st = (lanes && s.lane > 0) ? e->lane_stream[s.lane] : e->stream;
        const bool ncu = !e->ncu_layers.empty() && s.kind != S_TAP &&
                         (e->ncu_layers[0] == "*" || std::find(e->ncu_layers.begin(), e->ncu_layers.end(), s.label) != e->ncu_layers.end());
        if (ncu) cudaProfilerStart();
        switch (s.kind) {
            case S_CONV32: CK(launch_conv_fp32(s.cp, st)); break;
            case S_UMMA: CK(launch_conv_umma(s.ul, st)); break;
            case S_UMMA2: CK(launch_conv_umma2(s.u2, st)); break;
            case S_PAIR: CK(launch_conv_pair(s.pl, st)); break;
            case S_P2RAW: CK(launch_planes_to_raw(s.b_in, s.b_in_lo, s.f_out, s.n, s.cpad, s.C, s.f16, st)); break;
            case S_MRF: CK(launch_mrf_combine(s.mrf, s.n, st)); break;
            case S_POSTMRF: CK(launch_conv_post_mrf(s.mrf, s.w, s.bias, s.f_out, s.B, s.L, s.C, s.k, s.flag1, st)); break;
            case S_POST32: CK(launch_conv_post_fp32(s.f_in, s.w, s.bias, s.f_out, s.B, s.L, s.C, s.k, s.flag0, s.flag1, st)); break;
            case S_ACCUM: CK(launch_accum_fp32(s.f_out, s.f_in, s.n, s.flag0, s.fval, st)); break;
            case S_MEL_CL32: CK(launch_transpose_cf_to_cl(s.f_in, s.f_out, s.B, s.C, s.L, st)); break;
            case S_MEL_CLBF: CK(launch_mel_to_cl_bf16(s.f_in, s.b_out, s.b_out_lo, s.B, s.C, s.L, s.cpad, 0, s.f16, st)); break;
            case S_STRIP: CK(launch_zero_tail_rows(s.b_out, s.b_out_lo, plan->lens_dev, s.k, s.B, s.L, s.C, s.flag0, st)); break;
            case S_TAP: RET(store_tap(e, s)); break;   // a copy, not one of our kernels
            case S_FORK: case S_JOIN: break;
        }
        if (ncu) cudaProfilerStop();
        if (s.kind != S_TAP) ++e->launches;
        if (e->profiling) {
            CK(cudaEventRecord(e->prof_events[nev], st));
            e->prof_recs.push_back({s.label.empty() ? std::string(kind_label(s.kind)) : s.label, (int)s.kind, s.flops, s.bytes,
                                    nev - 1, nev});
            ++nev;
        }
    }
    e->prof_used = nev;
    return HFG_OK;
}

int ensure_arena(hfg_engine* e, size_t bytes) {
    if (bytes <= e->arena_bytes) return HFG_OK;
    CK(cudaStreamSynchronize(e->stream));
    e->plans.clear();   // plans hold pointers into the arena
    cudaFree(e->arena);
    e->arena = nullptr;
    e->arena_bytes = 0;
    cudaError_t err = cudaMalloc(&e->arena, bytes);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail(HFG_ERR_NOMEM, std::string("workspace allocation failed: ") + cudaGetErrorString(err));
    }
    e->arena_bytes = bytes;
    return HFG_OK;
}

int check_prec(int prec) {
    if (prec != HFG_PREC_FP32 && prec != HFG_PREC_BF16 && prec != HFG_PREC_BF16X3 && prec != HFG_PREC_FP16)
        return fail(HFG_ERR_INVALID, "unknown precision");
    return HFG_OK;
}

}  // namespace
}  // namespace hfg

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int hfg_abi_version(void) { return HFG_ABI_VERSION; }
const char* hfg_last_error(void) { return g_error.c_str(); }

int hfg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hfg_create(const hfg_config* cfg, int device, hfg_engine** out) {
    if (!cfg || !out) return fail(HFG_ERR_INVALID, "hfg_create: null argument");
    *out = nullptr;
    RET(validate_cfg(*cfg));
    const int n = hfg_device_count();
    if (n <= 0) return fail(HFG_ERR_CUDA, "hfg_create: no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= n) return fail(HFG_ERR_INVALID, "hfg_create: device index out of range");
    DeviceGuard guard(device);
    CK(guard.err);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char buf[128];
        snprintf(buf, sizeof buf, "hfg_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        return fail(HFG_ERR_UNSUPPORTED, buf);
    }
    std::unique_ptr<hfg_engine> e(new hfg_engine());
    e->cfg = *cfg;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    for (int i = 1; i < cfg->num_kernels; ++i) {
        CK(cudaStreamCreateWithFlags(&e->lane_stream[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&e->ev_plan_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->ev_d2h_done[i], cudaEventDisableTiming));
    }
    build_layers(e.get());
    if (const char* nl = getenv("HFG_NCU_LAYERS")) {   // profiling aid: ncu --profile-from-start off captures only these
        std::string cur;
        for (const char* c = nl;; ++c) {
            if (*c == ',' || *c == 0) { if (!cur.empty()) e->ncu_layers.push_back(cur); cur.clear(); if (!*c) break; }
            else cur.push_back(*c);
        }
    }
    *out = e.release();
    return HFG_OK;
}

void hfg_destroy(hfg_engine* e) {
    if (!e) return;
    DeviceGuard guard(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    e->plans.clear();
    for (int i = 0; i < 2; ++i) {
        cudaFree(e->wave_stage[i]);
        if (e->ev_plan_done[i]) cudaEventDestroy(e->ev_plan_done[i]);
        if (e->ev_d2h_done[i]) cudaEventDestroy(e->ev_d2h_done[i]);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (int i = 1; i < HFG_MAX_KERNELS; ++i) {
        if (e->lane_stream[i]) { cudaStreamSynchronize(e->lane_stream[i]); cudaStreamDestroy(e->lane_stream[i]); }
        if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    }
    for (auto& L : e->layers) free_layer_dev(L);
    for (auto& kv : e->folded) free_layer_dev(kv.second);
    for (auto& kv : e->taps) cudaFree(kv.second.dev);
    for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
    cudaFree(e->arena);
    cudaFree(e->scratch);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int hfg_num_layers(const hfg_engine* e) { return e ? (int)e->layers.size() : 0; }

int hfg_layer_name(const hfg_engine* e, int index, char* buf, size_t buflen) {
    if (!e || !buf || index < 0 || index >= (int)e->layers.size()) return fail(HFG_ERR_INVALID, "hfg_layer_name: bad argument");
    const std::string& n = e->layers[index].name;
    if (n.size() + 1 > buflen) return fail(HFG_ERR_INVALID, "hfg_layer_name: buffer too small");
    memcpy(buf, n.c_str(), n.size() + 1);
    return HFG_OK;
}

int hfg_layer_shape(const hfg_engine* e, const char* layer, int32_t dims[3], int32_t* is_transposed) {
    if (!e || !dims) return fail(HFG_ERR_INVALID, "hfg_layer_shape: null argument");
    const Layer* L = find_layer(const_cast<hfg_engine*>(e), layer);
    if (!L) return fail(HFG_ERR_INVALID, std::string("unknown layer: ") + (layer ? layer : "(null)"));
    if (L->transposed) { dims[0] = L->cin; dims[1] = L->cout; } else { dims[0] = L->cout; dims[1] = L->cin; }
    dims[2] = L->k;
    if (is_transposed) *is_transposed = L->transposed ? 1 : 0;
    return HFG_OK;
}

int hfg_set_weight(hfg_engine* e, const char* layer, const float* w, const float* bias) {
    if (!e || !w || !bias) return fail(HFG_ERR_INVALID, "hfg_set_weight: null argument");
    Layer* L = find_layer(e, layer);
    if (!L) return fail(HFG_ERR_INVALID, std::string("unknown layer: ") + (layer ? layer : "(null)"));
    const size_t n = (size_t)L->cin * L->cout * L->k;
    L->w.assign(w, w + n);
    L->bias.assign(bias, bias + L->cout);
    L->set = true;
    e->finalized = false;
    return HFG_OK;
}

int hfg_set_weight_norm(hfg_engine* e, const char* layer, const float* g, const float* v, const float* bias) {
    if (!e || !g || !v || !bias) return fail(HFG_ERR_INVALID, "hfg_set_weight_norm: null argument");
    Layer* L = find_layer(e, layer);
    if (!L) return fail(HFG_ERR_INVALID, std::string("unknown layer: ") + (layer ? layer : "(null)"));
    // torch._weight_norm(v, g, dim=0): w = v * (g / ||v||), norm over every dim but 0.
    // dim 0 is C_out for Conv1d and C_in for ConvTranspose1d.
    const int rows = L->transposed ? L->cin : L->cout;
    const size_t cols = (size_t)(L->transposed ? L->cout : L->cin) * L->k;
    L->w.resize((size_t)rows * cols);
    for (int r = 0; r < rows; ++r) {
        double s = 0.0;
        for (size_t c = 0; c < cols; ++c) { const double x = v[r * cols + c]; s += x * x; }
        const float scale = g[r] / (float)sqrt(s);
        for (size_t c = 0; c < cols; ++c) L->w[r * cols + c] = v[r * cols + c] * scale;
    }
    L->bias.assign(bias, bias + L->cout);
    L->set = true;
    e->finalized = false;
    return HFG_OK;
}

int hfg_finalize(hfg_engine* e) {
    if (!e) return fail(HFG_ERR_INVALID, "hfg_finalize: null engine");
    GUARD(e);
    for (auto& L : e->layers)
        if (!L.set) return fail(HFG_ERR_STATE, "hfg_finalize: layer not set: " + L.name);
    CK(cudaStreamSynchronize(e->stream));
    e->plans.clear();
    for (auto& L : e->layers) RET(upload_layer(L));
    RET(build_folded(e));
    e->finalized = true;
    return HFG_OK;
}

int32_t hfg_hop(const hfg_engine* e) { return e ? e->hop : 0; }

// The kernels index one batch item's plane ([L][C], at most 2 planes) with 32-bit element offsets and put the batch on a grid axis
// (<= 65535): refuse what would wrap instead of computing garbage.  V1: T < 262144 frames (50 minutes) per call -- long-form input
// goes through sharding.synthesize_longform / synthesize_streaming in chunks long before that.
static int check_sizes(const hfg_engine* e, int B, int T) {
    if (B > 65535) return fail(HFG_ERR_UNSUPPORTED, "batch larger than 65535: split it");
    long long L = T, worst = (long long)T * std::max(e->cfg.in_channels, 32);
    int ch = e->cfg.upsample_initial_channel;
    worst = std::max(worst, L * ch);
    for (int i = 0; i < e->cfg.num_upsamples; ++i) {
        L *= e->cfg.upsample_rates[i];
        ch /= 2;
        worst = std::max(worst, L * std::max(ch, 32));
    }
    if (worst >= (1ll << 31)) return fail(HFG_ERR_UNSUPPORTED, "T too large for one call (a stage's per-item plane exceeds 2^31 elements): synthesize in chunks");
    return HFG_OK;
}

size_t hfg_workspace_bytes(const hfg_engine* e, int32_t B, int32_t T, int32_t precision) {
    if (!e || B <= 0 || T <= 0 || check_prec(precision) != HFG_OK || check_sizes(e, B, T) != HFG_OK) return 0;
    size_t bytes = 0;
    if (build_plan(const_cast<hfg_engine*>(e), B, T, precision, false, nullptr, nullptr, &bytes) != HFG_OK) return 0;
    return bytes;
}

void* hfg_stream(hfg_engine* e) { return e ? (void*)e->stream : nullptr; }
uint64_t hfg_launch_count(const hfg_engine* e) { return e ? e->launches : 0; }

int hfg_graph_stats(const hfg_engine* e, int32_t* captured, int32_t* failed) {
    if (!e) return fail(HFG_ERR_INVALID, "hfg_graph_stats: null engine");
    if (captured) *captured = e->graphs_captured;
    if (failed) *failed = e->graphs_failed;
    return HFG_OK;
}

int hfg_profile_enable(hfg_engine* e, int on) {
    if (!e) return fail(HFG_ERR_INVALID, "hfg_profile_enable: null engine");
    // (re)enabling starts a fresh record list; disabling keeps the records readable
    e->profiling = on != 0;
    if (on) { e->prof_recs.clear(); e->prof_used = 0; }
    return HFG_OK;
}

int hfg_profile_count(const hfg_engine* e) { return e ? (int)e->prof_recs.size() : 0; }

int hfg_profile_get(hfg_engine* e, int i, char* layer, size_t layer_len, char* kernel, size_t kernel_len, float* ms,
                    double* flops, double* bytes) {
    if (!e || i < 0 || i >= (int)e->prof_recs.size()) return fail(HFG_ERR_INVALID, "hfg_profile_get: bad index");
    GUARD(e);
    const auto& r = e->prof_recs[i];
    if (layer && layer_len) { strncpy(layer, r.label.c_str(), layer_len - 1); layer[layer_len - 1] = 0; }
    if (kernel && kernel_len) { strncpy(kernel, kind_label((StepKind)r.kind), kernel_len - 1); kernel[kernel_len - 1] = 0; }
    if (ms) {
        CK(cudaEventSynchronize(e->prof_events[r.ev1]));
        CK(cudaEventElapsedTime(ms, e->prof_events[r.ev0], e->prof_events[r.ev1]));
    }
    if (flops) *flops = r.flops;
    if (bytes) *bytes = r.bytes;
    return HFG_OK;
}

int hfg_sync(hfg_engine* e) {
    if (!e) return fail(HFG_ERR_INVALID, "hfg_sync: null engine");
    GUARD(e);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaStreamSynchronize(e->copy_stream));
    e->d2h_pending[0] = e->d2h_pending[1] = false;
    return HFG_OK;
}

static int forward_impl(hfg_engine* e, const float* mel, int32_t B, int32_t T, const int32_t* lengths, float* wave, int32_t precision,
                        uint32_t flags) {
    if (!e || !mel || !wave) return fail(HFG_ERR_INVALID, "hfg_forward: null argument");
    if (B <= 0 || T <= 0) return fail(HFG_ERR_INVALID, "hfg_forward: B and T must be positive");
    RET(check_sizes(e, B, T));
    RET(check_prec(precision));
    if (!e->finalized) return fail(HFG_ERR_STATE, "hfg_forward: call hfg_finalize first");
    const bool mel_dev = flags & HFG_MEL_ON_DEVICE, wave_dev = flags & HFG_WAVE_ON_DEVICE;
    const bool keep = flags & HFG_KEEP_TAPS;
    const bool ragged = lengths != nullptr;
    if (ragged) {
        if (keep) return fail(HFG_ERR_INVALID, "hfg_forward_ragged: HFG_KEEP_TAPS is not available for ragged batches");
        for (int b = 0; b < B; ++b)
            if (lengths[b] < 1 || lengths[b] > T) return fail(HFG_ERR_INVALID, "hfg_forward_ragged: every length must lie in [1, T]");
    }
    // HFG_NO_SYNC with HOST pointers: both must be page-locked and stay valid until hfg_sync (the copies are asynchronous);
    // not combined with taps or the canary mode, which read results back inside the call
    const bool async_host = (flags & HFG_NO_SYNC) && !(mel_dev && wave_dev);
    if (async_host && (keep || env_flag("HFG_GUARD", 0))) return fail(HFG_ERR_INVALID, "hfg_forward: HFG_NO_SYNC with host pointers excludes HFG_KEEP_TAPS / HFG_GUARD");
    GUARD(e);

    const auto key = std::make_tuple((int)B, (int)T, (int)precision, (keep ? 1 : 0) | (ragged ? 2 : 0));
    auto it = e->plans.find(key);
    if (it == e->plans.end()) {
        if (e->plans.size() >= 64) {   // variable-length traffic: bound the cache (plans are cheap to rebuild)
            CK(cudaStreamSynchronize(e->stream));
            e->plans.clear();
        }
        size_t bytes = 0;
        RET(build_plan(e, B, T, precision, keep, nullptr, nullptr, &bytes, ragged));
        RET(ensure_arena(e, bytes));
        std::unique_ptr<Plan> plan(new Plan());
        plan->bytes = bytes;
        // All plans share the arena from offset 0: they run one after another on one stream.
        RET(build_plan(e, B, T, precision, keep, e->arena, plan.get(), nullptr, ragged));
        for (const Step& st : plan->steps) plan->kernels += is_kernel_step(st.kind);
        it = e->plans.emplace(key, std::move(plan)).first;
    }
    Plan* plan = it->second.get();
    const size_t mel_bytes = (size_t)B * e->cfg.in_channels * T * sizeof(float);
    const size_t wave_bytes = (size_t)B * T * e->hop * sizeof(float);
    // HFG_GUARD=1 (debug): canaries between all workspace buffers, refilled before and verified after every forward -- the
    // stand-in for a memcheck pass: a kernel that writes past the end (or before the start) of a plane trips the next canary.
    for (size_t off : plan->guards) CK(cudaMemsetAsync(e->arena + off, kGuardByte, kGuardBytes, e->stream));
    CK(cudaMemcpyAsync(plan->mel_dev, mel, mel_bytes, mel_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    // (pageable source: the runtime stages it before returning, the caller's array may go away at once)
    if (ragged) CK(cudaMemcpyAsync(plan->lens_dev, lengths, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    RET(launch_plan(e, plan));
    if (!plan->guards.empty()) {
        if (env_flag("HFG_GUARD_SELFTEST", 0))   // prove the check itself: clobber one canary byte like a stray store would
            CK(cudaMemsetAsync(e->arena + plan->guards[plan->guards.size() / 2] + 7, 0, 1, e->stream));
        std::vector<uint8_t> host(kGuardBytes);
        CK(cudaStreamSynchronize(e->stream));
        for (size_t gi = 0; gi < plan->guards.size(); ++gi) {
            CK(cudaMemcpy(host.data(), e->arena + plan->guards[gi], kGuardBytes, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < kGuardBytes; ++i)
                if (host[i] != (uint8_t)kGuardByte) {
                    char buf[160];
                    snprintf(buf, sizeof buf, "HFG_GUARD: canary %zu of %zu (arena offset %zu) overwritten at byte %zu", gi,
                             plan->guards.size(), plan->guards[gi], i);
                    return fail(HFG_ERR_CUDA, buf);
                }
        }
    }
    if (async_host && !wave_dev) {
        const int k = (int)(e->async_count++ & 1);
        if (e->wave_stage_bytes[k] < wave_bytes) {
            CK(cudaStreamSynchronize(e->copy_stream));
            CK(cudaStreamSynchronize(e->stream));
            cudaFree(e->wave_stage[k]);
            e->wave_stage[k] = nullptr; e->wave_stage_bytes[k] = 0;
            CK(cudaMalloc(&e->wave_stage[k], wave_bytes));
            e->wave_stage_bytes[k] = wave_bytes;
            e->d2h_pending[k] = false;
        }
        if (e->d2h_pending[k]) CK(cudaStreamWaitEvent(e->stream, e->ev_d2h_done[k], 0));   // staging buffer k is free again
        CK(cudaMemcpyAsync(e->wave_stage[k], plan->wave_dev, wave_bytes, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaEventRecord(e->ev_plan_done[k], e->stream));
        CK(cudaStreamWaitEvent(e->copy_stream, e->ev_plan_done[k], 0));
        CK(cudaMemcpyAsync(wave, e->wave_stage[k], wave_bytes, cudaMemcpyDeviceToHost, e->copy_stream));
        CK(cudaEventRecord(e->ev_d2h_done[k], e->copy_stream));
        e->d2h_pending[k] = true;
        return HFG_OK;
    }
    CK(cudaMemcpyAsync(wave, plan->wave_dev, wave_bytes, wave_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    if (!(flags & HFG_NO_SYNC)) CK(cudaStreamSynchronize(e->stream));
    return HFG_OK;
}

int hfg_forward(hfg_engine* e, const float* mel, int32_t B, int32_t T, float* wave, int32_t precision, uint32_t flags) {
    return forward_impl(e, mel, B, T, nullptr, wave, precision, flags);
}

int hfg_forward_ragged(hfg_engine* e, const float* mel, int32_t B, int32_t T, const int32_t* lengths, float* wave, int32_t precision,
                       uint32_t flags) {
    if (!lengths) return fail(HFG_ERR_INVALID, "hfg_forward_ragged: null lengths");
    return forward_impl(e, mel, B, T, lengths, wave, precision, flags);
}

int hfg_run_layer(hfg_engine* e, const char* layer, const float* x, int32_t B, int32_t L, int32_t pre_lrelu, float* y,
                  int32_t precision) {
    if (!e || !x || !y) return fail(HFG_ERR_INVALID, "hfg_run_layer: null argument");
    if (B <= 0 || L <= 0) return fail(HFG_ERR_INVALID, "hfg_run_layer: B and L must be positive");
    RET(check_prec(precision));
    if (!e->finalized) return fail(HFG_ERR_STATE, "hfg_run_layer: call hfg_finalize first");
    Layer* lay = find_layer(e, layer);
    if (!lay) return fail(HFG_ERR_INVALID, std::string("unknown layer: ") + (layer ? layer : "(null)"));
    GUARD(e);
    const ConvGeom g = geom_of(*lay, B, L);
    const size_t n_in = (size_t)B * lay->cin * L, n_out = (size_t)B * lay->cout * g.Lout;
    const size_t n_in_pad = (size_t)B * lay->cin_pad * L, n_out_pad = (size_t)B * lay->cout_tc * g.Lout;
    const size_t need = (2 * n_in + 2 * n_out) * sizeof(float) + 4 * n_in_pad * 2 + 4 * n_out_pad * 2 + 16 * 256;
    if (need > e->scratch_bytes) {
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(e->scratch);
        e->scratch = nullptr; e->scratch_bytes = 0;
        CK(cudaMalloc(&e->scratch, need));
        e->scratch_bytes = need;
    }
    Bump bump{e->scratch};
    float* x_cf = bump.take<float>(n_in);
    float* x_cl = bump.take<float>(n_in);
    float* y_cl = bump.take<float>(n_out);
    float* y_cf = bump.take<float>(n_out);
    __nv_bfloat16* a_hi = bump.take<__nv_bfloat16>(n_in_pad);
    __nv_bfloat16* a_lo = bump.take<__nv_bfloat16>(n_in_pad);
    __nv_bfloat16* y_hi = bump.take<__nv_bfloat16>(n_out_pad);
    __nv_bfloat16* y_lo = bump.take<__nv_bfloat16>(n_out_pad);
    cudaStream_t st = e->stream;
    CK(cudaMemcpyAsync(x_cf, x, n_in * sizeof(float), cudaMemcpyHostToDevice, st));
    if (lay->is_post) {
        CK(launch_transpose_cf_to_cl(x_cf, x_cl, B, lay->cin, L, st));
        CK(launch_conv_post_fp32(x_cl, lay->d_w32, lay->d_bias, y_cf, B, L, lay->cin, lay->k, pre_lrelu, 0, st));
        e->launches += 2;
    } else if (precision == HFG_PREC_FP32) {
        CK(launch_transpose_cf_to_cl(x_cf, x_cl, B, lay->cin, L, st));
        ConvParams p = conv32(*lay, B, L, x_cl, y_cl, nullptr, pre_lrelu, 0, 0.f);
        CK(launch_conv_fp32(p, st));
        CK(launch_transpose_cl_to_cf(y_cl, y_cf, B, lay->cout, g.Lout, st));
        e->launches += 3;
    } else {
        // The forward's production path for this layer: activated operand planes in (channel-padded like the plan's),
        // activated planes out, inverted back to the raw conv output for the caller.
        const bool x3 = precision == HFG_PREC_BF16X3;
        const int f16 = precision == HFG_PREC_FP16 ? 1 : 0;
        const __nv_bfloat16* w_hi = f16 ? lay->d_wh : lay->d_wb_hi;
        CK(launch_mel_to_cl_bf16(x_cf, a_hi, x3 ? a_lo : nullptr, B, lay->cin, L, lay->cin_pad, pre_lrelu, f16, st));
        UmmaConvParams p;
        memset(&p, 0, sizeof p);
        p.g = geom_tc(*lay, B, L); p.cin_pad = lay->cin_pad; p.kc = lay->kc; p.npass = x3 ? 3 : 1; p.f16 = f16;
        p.bias = lay->d_bias_tc; p.a_per_tap = env_flag("HFG_UMMA_A_PER_TAP", 0);
        p.y_act = y_hi; p.y_act_lo = x3 ? y_lo : nullptr;
        Umma2Launch u2;
        if (umma2_supported(p) && plan_conv_umma2(&u2, p, a_hi, a_lo, w_hi, lay->d_wb_lo, e->sm_count) == HFG_OK) {
            CK(launch_conv_umma2(u2, st));
        } else {
            UmmaLaunch ul;
            RET(plan_conv_umma(&ul, p, a_hi, a_lo, w_hi, lay->d_wb_lo));
            CK(launch_conv_umma(ul, st));
        }
        CK(launch_planes_to_raw(y_hi, x3 ? y_lo : nullptr, y_cl, (size_t)B * g.Lout, lay->cout_tc, lay->cout, f16, st));
        CK(launch_transpose_cl_to_cf(y_cl, y_cf, B, lay->cout, g.Lout, st));
        e->launches += 4;
    }
    CK(cudaMemcpyAsync(y, y_cf, n_out * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return HFG_OK;
}

int hfg_run_pair(hfg_engine* e, int32_t resblock, int32_t m, const float* x, int32_t B, int32_t L, float* y, int32_t precision,
                 int32_t* fused) {
    return hfg_run_pair_mrf(e, resblock, m, x, nullptr, 1.0f, B, L, y, precision, fused);
}

int hfg_run_pair_mrf(hfg_engine* e, int32_t resblock, int32_t m, const float* x, const float* mrf_sum, float out_scale, int32_t B,
                     int32_t L, float* y, int32_t precision, int32_t* fused) {
    if (!e || !x || !y) return fail(HFG_ERR_INVALID, "hfg_run_pair: null argument");
    if (B <= 0 || L <= 0) return fail(HFG_ERR_INVALID, "hfg_run_pair: B and L must be positive");
    RET(check_prec(precision));
    if (precision == HFG_PREC_FP32) return fail(HFG_ERR_UNSUPPORTED, "hfg_run_pair: tensor-core precisions only");
    if (!e->finalized) return fail(HFG_ERR_STATE, "hfg_run_pair: call hfg_finalize first");
    char n1[64], n2[64];
    snprintf(n1, sizeof n1, "resblocks.%d.convs1.%d", resblock, m);
    snprintf(n2, sizeof n2, "resblocks.%d.convs2.%d", resblock, m);
    Layer* c1 = find_layer(e, n1);
    Layer* c2 = find_layer(e, n2);
    if (!c1 || !c2) return fail(HFG_ERR_INVALID, std::string("unknown layer: ") + n1);
    GUARD(e);
    const bool x3 = precision == HFG_PREC_BF16X3;
    const int f16 = precision == HFG_PREC_FP16 ? 1 : 0;
    const int C = c1->cin, Cp = c1->cin_pad;
    const size_t n_raw = (size_t)B * C * L, n_pad = (size_t)B * Cp * L;
    const size_t need = 2 * n_raw * sizeof(float) + 8 * n_pad * 2 + 16 * 256;
    if (need > e->scratch_bytes) {
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(e->scratch);
        e->scratch = nullptr; e->scratch_bytes = 0;
        CK(cudaMalloc(&e->scratch, need));
        e->scratch_bytes = need;
    }
    Bump bump{e->scratch};
    float* x_cf = bump.take<float>(n_raw);
    float* y_cl = bump.take<float>(n_raw);
    __nv_bfloat16* a_hi = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* a_lo = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* t_hi = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* t_lo = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* y_hi = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* y_lo = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* s_hi = bump.take<__nv_bfloat16>(n_pad);
    __nv_bfloat16* s_lo = bump.take<__nv_bfloat16>(n_pad);
    cudaStream_t st = e->stream;
    if (mrf_sum) {   // planes of lrelu(running sum), like the forward keeps them
        CK(cudaMemcpyAsync(x_cf, mrf_sum, n_raw * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(launch_mel_to_cl_bf16(x_cf, s_hi, x3 ? s_lo : nullptr, B, C, L, Cp, 1, f16, st));
        e->launches += 1;
    }
    CK(cudaMemcpyAsync(x_cf, x, n_raw * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(launch_mel_to_cl_bf16(x_cf, a_hi, x3 ? a_lo : nullptr, B, C, L, Cp, 1, f16, st));
    const __nv_bfloat16* w1h = f16 ? c1->d_wh : c1->d_wb_hi;
    const __nv_bfloat16* w2h = f16 ? c2->d_wh : c2->d_wb_hi;
    PairParams pp;
    memset(&pp, 0, sizeof pp);
    pp.B = B; pp.L = L; pp.C = Cp; pp.k1 = c1->k; pp.k2 = c2->k; pp.d = c1->dil; pp.npass = x3 ? 3 : 1; pp.f16 = f16;
    pp.bias1 = c1->d_bias_tc; pp.bias2 = c2->d_bias_tc;
    pp.x_hi = a_hi; pp.x_lo = x3 ? a_lo : nullptr;
    pp.w1_hi = w1h; pp.w1_lo = c1->d_wb_lo; pp.w2_hi = w2h; pp.w2_lo = c2->d_wb_lo;
    pp.y_hi = y_hi; pp.y_lo = x3 ? y_lo : nullptr;
    if (mrf_sum) { pp.mrf_hi = s_hi; pp.mrf_lo = x3 ? s_lo : nullptr; pp.out_scale = out_scale; }
    PairLaunch pl;
    const bool can_fuse = c1->cin_pad == c1->cout_tc && c1->k == c2->k && c2->dil == 1 && pair_supported(pp) &&
                          plan_conv_pair(&pl, pp, e->sm_count) == HFG_OK;
    if (fused) *fused = can_fuse ? 1 : 0;
    if (can_fuse) {
        CK(launch_conv_pair(pl, st));
        e->launches += 1;
    } else {
        for (int which = 0; which < 2; ++which) {
            const Layer* lay = which ? c2 : c1;
            UmmaConvParams p;
            memset(&p, 0, sizeof p);
            p.g = geom_tc(*lay, B, L); p.cin_pad = lay->cin_pad; p.kc = lay->kc; p.npass = x3 ? 3 : 1; p.f16 = f16;
            p.bias = lay->d_bias_tc;
            if (which) { p.res_hi = a_hi; p.res_lo = x3 ? a_lo : nullptr; }
            if (which && mrf_sum) { p.mrf_hi = s_hi; p.mrf_lo = x3 ? s_lo : nullptr; p.out_scale = out_scale; }
            p.y_act = which ? y_hi : t_hi; p.y_act_lo = x3 ? (which ? y_lo : t_lo) : nullptr;
            const __nv_bfloat16* in_hi = which ? t_hi : a_hi;
            const __nv_bfloat16* in_lo = which ? t_lo : a_lo;
            Umma2Launch u2;
            const __nv_bfloat16* wh = which ? w2h : w1h;
            if (umma2_supported(p) && plan_conv_umma2(&u2, p, in_hi, in_lo, wh, lay->d_wb_lo, e->sm_count) == HFG_OK) {
                CK(launch_conv_umma2(u2, st));
            } else {
                UmmaLaunch ul;
                RET(plan_conv_umma(&ul, p, in_hi, in_lo, wh, lay->d_wb_lo));
                CK(launch_conv_umma(ul, st));
            }
        }
        e->launches += 2;
    }
    CK(launch_planes_to_raw(y_hi, x3 ? y_lo : nullptr, y_cl, (size_t)B * L, Cp, C, f16, st));
    CK(launch_transpose_cl_to_cf(y_cl, x_cf, B, C, L, st));
    e->launches += 3;
    CK(cudaMemcpyAsync(y, x_cf, n_raw * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return HFG_OK;
}

int hfg_get_tap(hfg_engine* e, const char* name, float* out, size_t* n) {
    if (!e || !name || !n) return fail(HFG_ERR_INVALID, "hfg_get_tap: null argument");
    auto it = e->taps.find(name);
    if (it == e->taps.end()) return fail(HFG_ERR_INVALID, std::string("no such tap (run hfg_forward with HFG_KEEP_TAPS): ") + name);
    const Tap& t = it->second;
    const size_t cnt = (size_t)t.B * t.C * t.L;
    if (!out) { *n = cnt; return HFG_OK; }
    if (*n < cnt) return fail(HFG_ERR_INVALID, "hfg_get_tap: buffer too small");
    GUARD(e);
    float* tmp = nullptr;
    CK(cudaMalloc(&tmp, cnt * sizeof(float)));
    cudaError_t err = t.C == 1 ? cudaMemcpyAsync(tmp, t.dev, cnt * sizeof(float), cudaMemcpyDeviceToDevice, e->stream)
                               : launch_transpose_cl_to_cf(t.dev, tmp, t.B, t.C, t.L, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out, tmp, cnt * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(tmp);
    if (err != cudaSuccess) return cuda_fail(err, "hfg_get_tap");
    *n = cnt;
    return HFG_OK;
}

}  // extern "C"
