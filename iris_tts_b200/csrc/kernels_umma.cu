// tcgen05 tensor-core kernel family (HFG_PREC_BF16 / HFG_PREC_BF16X3).
//
// One kernel covers every Conv1d and (polyphase) ConvTranspose1d of the generator
//   F.conv1d            src/iris/hifigan_pretrained.py:67,69,124
//   F.conv_transpose1d  src/iris/hifigan_pretrained.py:128
// as an implicit GEMM on channels-last activations (see ConvGeom in hfg_internal.h):
//
//   D[m][n] (TMEM, fp32) = sum_chunk sum_tap  A[m + off_tap][chunk] (smem, bf16) * W[tap][n][chunk] (smem, bf16)
//
//   * A: ONE halo tile per K-chunk ([MT*128 + span] time rows x 64|32 channels), fetched by TMA
//     from the [B][L][C] tensor (3-D map: out-of-range rows are zero-filled, which IS the
//     conv's zero padding, and rows never bleed across batch items).  Each tap is a
//     row-shifted UMMA shared-memory descriptor into that tile - no im2col, no re-fetch.
//   * W: [N_TILE x chunk] K-major tile per (chunk, tap) by TMA, ring-buffered.
//   * tcgen05.mma cta_group::1 kind::f16, M=128, N=N_TILE, K=16, issued by one thread;
//     MT accumulators of N_TILE columns live in TMEM.
//   * BF16X3: operands are split hi+lo bf16 planes; three MMAs (hi*hi, lo*hi, hi*lo)
//     accumulate in fp32 for fp32-class accuracy.
//   * epilogue (4 warps, one TMEM lane = one time row per thread): bias, residual add
//     (:70), MRF accumulate / divide (:133-137), leaky_relu (:66,68,127,139) - writes the
//     raw fp32 stream and/or the activated bf16 operand plane(s) the next conv's TMA reads.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hfg_internal.h"
#include "umma_ptx.cuh"

namespace hfg {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr uint32_t kMaxDynSmem = 227u * 1024u - 2048u;   // leaves room for the static barriers

struct KArgs {
    ConvGeom g;
    int kc;            // K elements per chunk (64 or 32)
    int nchunks;       // cin_pad / kc
    int npass;         // 1 or 3
    int n_tile, mt;
    int lo;            // smallest tap offset
    int rows_a;        // staged A rows
    int a_box_rows;    // rows per TMA piece
    int a_pieces;
    int n_a, n_w;      // ring depths
    int a_per_tap;
    uint32_t a_plane_bytes, w_plane_bytes;   // bytes of one plane of one stage (1024-multiple)
    uint32_t tmem_cols;
    const float* bias;
    const float* res;
    const __nv_bfloat16* res_hi;
    const __nv_bfloat16* res_lo;
    const __nv_bfloat16* mrf_hi;
    const __nv_bfloat16* mrf_lo;
    float out_scale;
    int f16;
    float* y_raw;
    __nv_bfloat16* y_act;
    __nv_bfloat16* y_act_lo;
    float* xs;
    int xs_read, xs_write;
    float out_div;
};

using namespace ptx;

// ---------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) conv_umma_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                                                             const __grid_constant__ CUtensorMap map_a_lo,
                                                             const __grid_constant__ CUtensorMap map_w_hi,
                                                             const __grid_constant__ CUtensorMap map_w_lo,
                                                             const KArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int planes = a.npass > 1 ? 2 : 1;
    const uint32_t row_bytes = (uint32_t)a.kc * 2u;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_stage_bytes = a.a_plane_bytes * planes;
    const uint32_t w_stage_bytes = a.w_plane_bytes * planes;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_w = smem_base + a.n_a * a_stage_bytes;

    const uint32_t bar_a_full = smem_u32(&bars[0]);
    const uint32_t bar_a_empty = smem_u32(&bars[kMaxStages]);
    const uint32_t bar_w_full = smem_u32(&bars[2 * kMaxStages]);
    const uint32_t bar_w_empty = smem_u32(&bars[3 * kMaxStages]);
    const uint32_t bar_acc = smem_u32(&bars[4 * kMaxStages]);

    const int m0 = blockIdx.x * a.mt * 128;
    const int n0 = blockIdx.y * a.n_tile;
    const int b = blockIdx.z;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_w_hi);
        if (planes > 1) { prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < a.n_a; ++i) { mbar_init(bar_a_full + 8 * i, 1); mbar_init(bar_a_empty + 8 * i, 1); }
        for (int i = 0; i < a.n_w; ++i) { mbar_init(bar_w_full + 8 * i, 1); mbar_init(bar_w_empty + 8 * i, 1); }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(&tmem_base_slot), a.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                if (!a.a_per_tap) {
                    mbar_wait(bar_a_empty + 8 * sa, pa ^ 1u);
                    mbar_expect_tx(bar_a_full + 8 * sa, (uint32_t)a.rows_a * row_bytes * planes);
                    for (int pl = 0; pl < planes; ++pl)
                        for (int pc = 0; pc < a.a_pieces; ++pc)
                            tma_load_3d(smem_a + sa * a_stage_bytes + pl * a.a_plane_bytes + pc * a.a_box_rows * row_bytes,
                                        pl ? &map_a_lo : &map_a_hi, bar_a_full + 8 * sa, c * a.kc,
                                        m0 + a.lo + pc * a.a_box_rows, b);
                    if (++sa == a.n_a) { sa = 0; pa ^= 1u; }
                }
                for (int j = 0; j < a.g.taps; ++j) {
                    if (a.a_per_tap) {
                        mbar_wait(bar_a_empty + 8 * sa, pa ^ 1u);
                        mbar_expect_tx(bar_a_full + 8 * sa, (uint32_t)a.rows_a * row_bytes * planes);
                        for (int pl = 0; pl < planes; ++pl)
                            for (int pc = 0; pc < a.a_pieces; ++pc)
                                tma_load_3d(smem_a + sa * a_stage_bytes + pl * a.a_plane_bytes + pc * a.a_box_rows * row_bytes,
                                            pl ? &map_a_lo : &map_a_hi, bar_a_full + 8 * sa, c * a.kc,
                                            m0 + a.g.tap_off0 + j * a.g.tap_step + pc * a.a_box_rows, b);
                        if (++sa == a.n_a) { sa = 0; pa ^= 1u; }
                    }
                    mbar_wait(bar_w_empty + 8 * sw, pw ^ 1u);
                    mbar_expect_tx(bar_w_full + 8 * sw, (uint32_t)a.n_tile * row_bytes * planes);
                    for (int pl = 0; pl < planes; ++pl)
                        tma_load_2d(smem_w + sw * w_stage_bytes + pl * a.w_plane_bytes, pl ? &map_w_lo : &map_w_hi,
                                    bar_w_full + 8 * sw, c * a.kc, j * a.g.Np + n0);
                    if (++sw == a.n_w) { sw = 0; pw ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // Warp-uniform loops; the elected lane issues.  Descriptors advance by one add per MMA.
        {
            const bool leader = elect_one();
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 @17, M>>4 @24
            const uint32_t idesc = make_idesc((uint32_t)a.n_tile, a.f16 != 0);
            const uint32_t dhi = desc_hi(row_bytes);
            const uint32_t sub_step = (128u * row_bytes) >> 4;
            const uint32_t a_pl_step = a.a_plane_bytes >> 4, w_pl_step = a.w_plane_bytes >> 4;
            const bool k4 = a.kc == 64;
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            uint32_t acc = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                if (!a.a_per_tap) mbar_wait(bar_a_full + 8 * sa, pa);
                for (int j = 0; j < a.g.taps; ++j) {
                    if (a.a_per_tap) mbar_wait(bar_a_full + 8 * sa, pa);
                    mbar_wait(bar_w_full + 8 * sw, pw);
                    tc_fence_after();
                    const int shift = a.a_per_tap ? 0 : (a.g.tap_off0 + j * a.g.tap_step - a.lo);
                    const uint32_t a_lo0 = desc_lo(smem_a + sa * a_stage_bytes) + (((uint32_t)shift * row_bytes) >> 4);
                    const uint32_t w_lo0 = desc_lo(smem_w + sw * w_stage_bytes);
                    for (int ps = 0; ps < a.npass; ++ps) {
                        uint32_t a_lo = a_lo0 + (ps == 1 ? a_pl_step : 0u);   // (hi,hi) (lo,hi) (hi,lo)
                        const uint32_t w_lo = w_lo0 + (ps == 2 ? w_pl_step : 0u);
                        uint32_t d = tmem_base;
                        for (int ms = 0; ms < a.mt; ++ms) {
                            if (leader) {
                                if (k4) umma_ksteps<4>(d, a_lo, w_lo, dhi, idesc, acc);
                                else umma_ksteps<2>(d, a_lo, w_lo, dhi, idesc, acc);
                            }
                            a_lo += sub_step;
                            d += (uint32_t)a.n_tile;
                        }
                        acc = 1;
                    }
                    if (leader) umma_commit(bar_w_empty + 8 * sw);
                    if (++sw == a.n_w) { sw = 0; pw ^= 1u; }
                    if (a.a_per_tap) {
                        if (leader) umma_commit(bar_a_empty + 8 * sa);
                        if (++sa == a.n_a) { sa = 0; pa ^= 1u; }
                    }
                }
                if (!a.a_per_tap) {
                    if (leader) umma_commit(bar_a_empty + 8 * sa);
                    if (++sa == a.n_a) { sa = 0; pa ^= 1u; }
                }
            }
            if (leader) umma_commit(bar_acc);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;   // TMEM lane quarter this warp may read
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int Cout = a.g.Cout;
        for (int ms = 0; ms < a.mt; ++ms) {
            const int m = m0 + ms * 128 + q * 32 + lane;
            const bool mvalid = m < a.g.Mrows;
            for (int nc = 0; nc < a.n_tile; nc += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ms * a.n_tile + nc), r);
                tmem_wait_ld();
                const int n = n0 + nc;
                int rr = 0, co = n;
                if (a.g.ups_s > 1) { rr = n / Cout; co = n - rr * Cout; }
                const int t_out = m * a.g.ups_s + rr - a.g.ups_p;
                if (!mvalid || t_out < 0 || t_out >= a.g.Lout) continue;
                const size_t off = ((size_t)b * a.g.Lout + t_out) * Cout + co;
                float v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + co) + i);
                    v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bv.x;
                    v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bv.y;
                    v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bv.z;
                    v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bv.w;
                }
                if (a.res) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 rv = *(reinterpret_cast<const float4*>(a.res + off) + i);
                        v[4 * i + 0] += rv.x; v[4 * i + 1] += rv.y; v[4 * i + 2] += rv.z; v[4 * i + 3] += rv.w;
                    }
                }
                // residual (and the running MRF sum of the previous branches) carried as activated planes: x = inverse-lrelu(hi (+ lo))
                for (int src = 0; src < 2; ++src) {
                    const __nv_bfloat16* ph = src ? a.mrf_hi : a.res_hi;
                    const __nv_bfloat16* pl = src ? a.mrf_lo : a.res_lo;
                    if (!ph) continue;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 u = *(reinterpret_cast<const uint4*>(ph + off) + i);
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                        float f[8];
                        if (a.f16) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) unpack2<true>(w[t], f[2 * t], f[2 * t + 1]);
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t) unpack2<false>(w[t], f[2 * t], f[2 * t + 1]);
                        }
                        if (pl) {
                            const uint4 ul = *(reinterpret_cast<const uint4*>(pl + off) + i);
                            const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) { f[2 * t] += __uint_as_float(wl[t] << 16); f[2 * t + 1] += __uint_as_float(wl[t] & 0xffff0000u); }
                        }
#pragma unroll
                        for (int t = 0; t < 8; ++t) v[8 * i + t] += f[t] > 0.f ? f[t] : f[t] * (1.0f / kLreluSlope);
                    }
                    if (src) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= a.out_scale;
                    }
                }
                if (a.y_raw) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *(reinterpret_cast<float4*>(a.y_raw + off) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                if (a.xs_read) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 ov = *(reinterpret_cast<const float4*>(a.xs + off) + i);
                        v[4 * i + 0] = ov.x + v[4 * i + 0]; v[4 * i + 1] = ov.y + v[4 * i + 1];
                        v[4 * i + 2] = ov.z + v[4 * i + 2]; v[4 * i + 3] = ov.w + v[4 * i + 3];
                    }
                }
                if (a.out_div > 0.f) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __fdiv_rn(v[i], a.out_div);
                }
                if (a.xs_write) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *(reinterpret_cast<float4*>(a.xs + off) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                if (a.y_act) {
                    uint32_t hi[16];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = lrelu(v[i]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) hi[i] = a.f16 ? pack_f16(v[2 * i], v[2 * i + 1]) : pack_bf16(v[2 * i], v[2 * i + 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *(reinterpret_cast<uint4*>(a.y_act + off) + i) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                    if (a.y_act_lo) {
                        uint32_t lo[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&hi[i]);
                            lo[i] = pack_bf16(v[2 * i] - __low2float(h), v[2 * i + 1] - __high2float(h));
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *(reinterpret_cast<uint4*>(a.y_act_lo + off) + i) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---------------------------------------------------------------------------
// Host side: tensor maps + launch geometry
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int kc) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t gdim[3];
    cuuint64_t gstr[2];
    cuuint32_t bx[3], es[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu,%llu box %u,%u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        set_error(buf);
        return false;
    }
    return true;
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

int plan_conv_umma(UmmaLaunch* L, const UmmaConvParams& p, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                   const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo) {
    const ConvGeom& g = p.g;
    if (p.kc != 64 && p.kc != 32) { set_error("conv_umma: kc must be 64 or 32"); return HFG_ERR_INVALID; }
    if (p.cin_pad % p.kc != 0 || g.Np % 32 != 0 || g.Cout % 32 != 0) {
        set_error("conv_umma: channel counts must be multiples of 32");
        return HFG_ERR_UNSUPPORTED;
    }
    if (p.npass != 1 && p.npass != 3) { set_error("conv_umma: npass must be 1 or 3"); return HFG_ERR_INVALID; }
    L->p = p;
    const int planes = p.npass > 1 ? 2 : 1;
    const uint32_t row_bytes = (uint32_t)p.kc * 2u;

    // N tile: largest of 256/128/64/32 dividing Np
    int n_tile = 32;
    for (int c : {256, 128, 64}) if (g.Np % c == 0) { n_tile = c; break; }
    // TMEM columns per CTA: 256 keeps two CTAs co-resident per SM (one's epilogue overlaps the other's MMAs)
    const int cols_budget = env_int("HFG_UMMA_COLS", 256);
    int mt = std::max(1, std::min(cols_budget / n_tile, env_int("HFG_UMMA_MAX_MT", 4)));
    const int mt_need = (g.Mrows + 127) / 128;
    mt = std::min(mt, mt_need);
    L->n_tile = n_tile;
    L->mt = mt;

    const int last_off = g.tap_off0 + (g.taps - 1) * g.tap_step;
    const int lo = std::min(g.tap_off0, last_off);
    const int span = std::max(g.tap_off0, last_off) - lo;
    const int a_per_tap = p.a_per_tap;
    int rows_need = a_per_tap ? mt * 128 : mt * 128 + span;
    int pieces = (rows_need + 255) / 256;
    int box_rows = ((rows_need + pieces - 1) / pieces + 7) / 8 * 8;
    if (a_per_tap) { pieces = mt; box_rows = 128; }
    const int rows_a = pieces * box_rows;
    L->rows_a = rows_a;

    const uint32_t a_plane = round_up((uint32_t)rows_a * row_bytes, 1024);
    const uint32_t w_plane = round_up((uint32_t)n_tile * row_bytes, 1024);
    const uint32_t a_stage = a_plane * planes, w_stage = w_plane * planes;
    const int nchunks = p.cin_pad / p.kc;
    const int a_uses = a_per_tap ? nchunks * g.taps : nchunks;
    const int w_uses = nchunks * g.taps;

    // smem budget: aim for two CTAs per SM when the TMEM budget allows it
    const uint32_t kMaxSmem = kMaxDynSmem;
    const bool want_two = (uint32_t)(mt * n_tile) <= 256u;
    uint32_t budget = want_two ? (kMaxSmem - 1024u) / 2u : kMaxSmem;
    budget -= 1024;  // alignment slack
    int n_a = 1, n_w = 1;
    auto total = [&](int na, int nw) { return (uint32_t)na * a_stage + (uint32_t)nw * w_stage; };
    if (total(1, 2) > budget && want_two) budget = kMaxSmem - 1024;  // does not fit twice: take the whole SM
    if (total(1, 1) > budget) { set_error("conv_umma: tile does not fit in shared memory"); return HFG_ERR_UNSUPPORTED; }
    const int max_a = std::min({a_uses, kMaxStages, env_int("HFG_UMMA_A_STAGES", a_per_tap ? 4 : 2)});
    const int max_w = std::min({w_uses, kMaxStages, env_int("HFG_UMMA_W_STAGES", 4)});
    bool grew = true;
    while (grew) {
        grew = false;
        if (n_w < max_w && total(n_a, n_w + 1) <= budget) { ++n_w; grew = true; }
        if (n_a < max_a && total(n_a + 1, n_w) <= budget && n_w >= std::min(2, max_w)) { ++n_a; grew = true; }
    }
    L->smem = total(n_a, n_w) + 1024;

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(mt * n_tile)) tmem_cols <<= 1;

    // tensor maps
    {
        const uint64_t dims[3] = {(uint64_t)p.cin_pad, (uint64_t)g.Lin, (uint64_t)g.B};
        const uint64_t str[2] = {(uint64_t)p.cin_pad * 2, (uint64_t)g.Lin * p.cin_pad * 2};
        const uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)box_rows, 1};
        if (!encode_map(&L->map_a_hi, x_hi, 3, dims, str, box, p.kc)) return HFG_ERR_CUDA;
        if (!encode_map(&L->map_a_lo, planes > 1 ? x_lo : x_hi, 3, dims, str, box, p.kc)) return HFG_ERR_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p.cin_pad, (uint64_t)g.taps * g.Np};
        const uint64_t str[1] = {(uint64_t)p.cin_pad * 2};
        const uint32_t box[2] = {(uint32_t)p.kc, (uint32_t)n_tile};
        if (!encode_map(&L->map_w_hi, w_hi, 2, dims, str, box, p.kc)) return HFG_ERR_CUDA;
        if (!encode_map(&L->map_w_lo, planes > 1 ? w_lo : w_hi, 2, dims, str, box, p.kc)) return HFG_ERR_CUDA;
    }
    L->grid = dim3((g.Mrows + mt * 128 - 1) / (mt * 128), g.Np / n_tile, g.B);

    L->stages_a = n_a;
    L->stages_w = n_w;
    L->a_box_rows = box_rows;
    L->a_pieces = pieces;
    L->lo = lo;
    L->tmem_cols = tmem_cols;
    L->a_plane_bytes = a_plane;
    L->w_plane_bytes = w_plane;
    return HFG_OK;
}

cudaError_t launch_conv_umma(const UmmaLaunch& L, cudaStream_t s) {
    static size_t configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (L.smem > configured[dev % 64]) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
        if (e != cudaSuccess) return e;
        configured[dev % 64] = kMaxDynSmem;
    }
    KArgs a;
    memset(&a, 0, sizeof a);
    const UmmaConvParams& p = L.p;
    a.g = p.g;
    a.kc = p.kc;
    a.nchunks = p.cin_pad / p.kc;
    a.npass = p.npass;
    a.n_tile = L.n_tile;
    a.mt = L.mt;
    a.lo = L.lo;
    a.rows_a = L.rows_a;
    a.a_box_rows = L.a_box_rows;
    a.a_pieces = L.a_pieces;
    a.n_a = L.stages_a;
    a.n_w = L.stages_w;
    a.a_per_tap = p.a_per_tap;
    a.a_plane_bytes = L.a_plane_bytes;
    a.w_plane_bytes = L.w_plane_bytes;
    a.tmem_cols = L.tmem_cols;
    a.bias = p.bias;
    a.res = p.res;
    a.res_hi = p.res_hi;
    a.res_lo = p.res_lo;
    a.mrf_hi = p.mrf_hi;
    a.mrf_lo = p.mrf_lo;
    a.out_scale = p.mrf_hi ? p.out_scale : 1.0f;
    a.f16 = p.f16 ? 1 : 0;
    a.y_raw = p.y_raw;
    a.y_act = p.y_act;
    a.y_act_lo = p.y_act_lo;
    a.xs = p.xs;
    a.xs_read = p.xs_read;
    a.xs_write = p.xs_write;
    a.out_div = p.out_div;
    conv_umma_kernel<<<L.grid, kThreads, L.smem, s>>>(L.map_a_hi, L.map_a_lo, L.map_w_hi, L.map_w_lo, a);
    return cudaGetLastError();
}

}  // namespace hfg
