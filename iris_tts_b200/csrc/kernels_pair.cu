// conv_pair: one ResBlock (convs1[m], convs2[m]) pair as ONE persistent tcgen05 kernel for the narrow stages (C <= 64).
//
//   xt = lrelu(x); xt = c1(xt); xt = lrelu(xt); xt = c2(xt); x = xt + x      src/iris/hifigan_pretrained.py:66-70
//
// The unfused plan (kernels_umma2.cu) moves five activation streams through HBM per pair (c1: read x, write t; c2: read t,
// read x, write out) and the C <= 64 layers are bound by exactly that (and by the shared-memory operand fetch of small-N
// MMAs), not by the tensor pipe.  Here the intermediate t never leaves the SM:
//
//   x halo tile (TMA, OOB rows = 0)  --c1 taps: row-shifted UMMA descriptors-->  acc1 (TMEM)
//   acc1 --epilogue 1: + b1, lrelu, rows outside [0, L) := 0 (c2's zero padding), bf16 plane(s)-->  t tile (shared, swizzled K-major)
//   t tile  --c2 taps-->  acc2 (TMEM)
//   acc2 --epilogue 2: + b2 + inverse-lrelu(x rows of the SAME x tile), lrelu, bf16 plane(s)-->  staging -> TMA store
//
// Two streams (x in, out) instead of five.  A tile computes R = 128*MT conv rows of both convs; the last k2-1 rows of c2 see
// t rows this tile did not compute and are dropped: V = R - (k2-1) valid output rows per tile (k <= 11: >= 92 % at MT = 1).
// c1 has k1 taps with dilation d, c2 k2 taps with dilation 1 (k1 = k2 for a reference ResBlock; the time-folded narrow stages of
// engine.cu:build_folded give different counts).
//
// c1 and c2 have separate issuer warps ordered only by barriers: c1 of tile i+1 (and i+2) runs while epilogue 1 of tile i+1
// and epilogue 2 of tile i work on their own warps, so the tensor pipe stays busy.  acc1 / acc2 are double-buffered in TMEM (4 * MT * N <= 512 columns).
//
// Arithmetic is the unfused plan's: same bf16 (hi[, lo]) rounding of t, same fp32 accumulation per output row with a fixed
// (tap, K-slice, pass) order that does not depend on the tile a row falls in, nor on B or L.
//
// Warps: 0 = TMA producer (resident weights, x ring), 1 = c2 weight ring (only when c2's weights are streamed), 2 = TMEM allocator,
//        3 = barrier init,
//        4..4+2*MT-1 = MMA issuers (one per conv and 128-row subtile), 8-15 = epilogue 1, 16-23 = epilogue 2 (two groups of four
//        warps each; group g takes the tiles with (i & 1) == g, so a group has two tile intervals for its tile).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hfg_internal.h"
#include "umma_ptx.cuh"

namespace hfg {

namespace {

using namespace ptx;

constexpr int kThreads = 768;   // 24 warps: the epilogues are issue-latency-bound, two groups of each alternate tiles
constexpr int kMaxX = 8, kMaxW2 = 8;
constexpr uint32_t kSmemMax = 227u * 1024u - 2048u;   // dynamic smem; barriers / bias (static, < 2 KB) live outside

struct PairArgs {
    int B, L, N;                 // N = C = 32 or 64 (operand rows of 64 / 128 bytes)
    int k1, k2, d, h1, h2;       // c1: k1 taps, dilation d, halo h1 = d(k1-1)/2 ; c2: k2 taps, dilation 1, halo h2 = (k2-1)/2
    int planes, mt, R, V;        // R = 128*mt conv rows per tile, V = R - (k2-1) valid output rows
    int tiles_per_item, total_tiles;
    int x_rows, x_box_rows, x_pieces;
    int n_x, n_t, n_o;
    int n_w2;                    // 0: c2's weights resident like c1's ; > 0: streamed per tile through a ring of n_w2 tap tiles
    int concat, acc_n, paired, reverse;
    int f16;                     // fp16 operand planes (single-plane mode)
    int has_mrf;                 // epilogue 2 also adds the running MRF sum of the previous branches (an R-row tile per tile of work,
                                 // double-buffered, loaded by warp 3) and scales by out_scale  (hifigan_pretrained.py:133-137)
    float out_scale;
    uint32_t m_plane_bytes, off_m;
    int dbg;                     // HFG_PAIR_DBG (timing experiments only): 1 = epilogue 2 idle, 2 = no MMAs, 3 = epilogue 1 idle, 4 = no TMA stores
    const int32_t* lens;         // ragged batch: item b is lens[b] * len_mul rows long (nullptr: every item is L rows)
    int len_mul;
    int len_skip;                // tiles whose first output row lies this far behind the item's end are not computed (kRagged)
    uint32_t x_plane_bytes, t_plane_bytes, w_plane_bytes, o_plane_bytes;
    uint32_t off_w, off_t, off_o;
    const float* bias1;
    const float* bias2;
};

// hi (and lo = v - hi) planes of 8 consecutive channels (4 pairs) -> 16-byte chunks
template <int kPlanes, bool kF16>
__device__ __forceinline__ void store_planes8(const f2* v, uint32_t addr, uint32_t plane_bytes) {
    uint4 hi;
    hi.x = f2_to_h2<kF16>(v[0]); hi.y = f2_to_h2<kF16>(v[1]); hi.z = f2_to_h2<kF16>(v[2]); hi.w = f2_to_h2<kF16>(v[3]);
    sts128(addr, hi);
    if (kPlanes > 1) {
        uint4 lo;
        lo.x = f2_to_bf16x2(f2_sub(v[0], f2_from_bf16x2(hi.x))); lo.y = f2_to_bf16x2(f2_sub(v[1], f2_from_bf16x2(hi.y)));
        lo.z = f2_to_bf16x2(f2_sub(v[2], f2_from_bf16x2(hi.z))); lo.w = f2_to_bf16x2(f2_sub(v[3], f2_from_bf16x2(hi.w)));
        sts128(addr + plane_bytes, lo);
    }
}

// 16-byte chunk XOR of a row inside a swizzled tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t row_bytes) { return row_bytes == 128 ? (row & 7u) : ((row >> 1) & 3u); }

// All taps of one 128-row subtile of one conv.  The loop is warp-uniform (operands live in uniform registers); only the elected
// lane issues.  KS = K=16 slices per tap, NP = MMA passes per tap: 1 bf16 | 3 bf16x3 (hi,hi)(lo,hi)(hi,lo) | 2 concat (hi,[hi;lo])(lo,hi).
// Per tap the descriptors advance by one add each; everything else is loop-invariant (the generic loop spent ~19 instructions
// and ~100 cycles per MMA on rebuilding them, twice the tensor pipe's own time for N <= 64).
// Ragged batch: tile `tile` lies wholly behind the rows anything inside its item can read.  Every role evaluates this for every tile
// of its CTA and skips the same ones; ring positions and buffer parities count live tiles only.
__device__ __forceinline__ bool pair_tile_skipped(const PairArgs& a, int tile) {
    const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
    const int b = tl / a.tiles_per_item;
    return (tl - b * a.tiles_per_item) * a.V >= __ldg(a.lens + b) * a.len_mul + a.len_skip;
}

template <int KS, int NP>
__device__ __forceinline__ void issue_taps(bool leader, int k, uint32_t d0, uint32_t a_lo, uint32_t w_lo, uint32_t a_tap, uint32_t w_tap,
                                           uint32_t a_pl, uint32_t w_pl, uint32_t dhi, uint32_t id0, uint32_t id1, uint32_t acc_in = 0u) {
#pragma unroll 1
    for (int j = 0; j < k; ++j) {
        if (leader) {
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
                const uint32_t aa = a_lo + (ps == 1 ? a_pl : 0u);
                const uint32_t ww = w_lo + (ps == 2 ? w_pl : 0u);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    umma_bf16_lh(d0, aa + 2u * ks, ww + 2u * ks, dhi, ps == 0 ? id0 : id1, (ps | ks) ? 1u : (acc_in | (uint32_t)(j != 0)));
            }
        }
        a_lo += a_tap;
        w_lo += w_tap;
    }
}

// kRagged: per-item sequence ends (hfg_forward_ragged) as a separate instantiation; the dense kernels carry none of its code.
template <int kPlanes, bool kF16, bool kRagged>
__global__ void __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                 const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                 const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                 const __grid_constant__ CUtensorMap map_y_hi, const __grid_constant__ CUtensorMap map_y_lo,
                 const __grid_constant__ CUtensorMap map_yt_hi, const __grid_constant__ CUtensorMap map_yt_lo,
                 const __grid_constant__ CUtensorMap map_m_hi, const __grid_constant__ CUtensorMap map_m_lo, const PairArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxX + 2 * kMaxW2 + 21];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[2][64];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler too: role values stay in uniform registers
    const int lane = threadIdx.x & 31;
    constexpr int planes = kPlanes;
    constexpr int npass = kPlanes == 2 ? 3 : 1;
    constexpr int CW = 16;                        // epilogue column chunk: 768 threads leave 80 registers per thread
    const uint32_t row_bytes = (uint32_t)a.N * 2u;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t x_stage = a.x_plane_bytes * planes;
    const uint32_t t_stage = a.t_plane_bytes * planes;
    const uint32_t w_stage = a.w_plane_bytes * planes;     // one tap of one conv
    const uint32_t o_slot = a.o_plane_bytes * planes;      // one epilogue-2 warp's 32-row staging box
    const uint32_t smem_x = smem_base;
    const uint32_t smem_w = smem_base + a.off_w;
    const uint32_t smem_t = smem_base + a.off_t;
    const uint32_t smem_o = smem_base + a.off_o;
    const uint32_t smem_m = smem_base + a.off_m;
    const uint32_t m_stage = a.m_plane_bytes * planes;

    uint32_t bp = smem_u32(&bars[0]);
    const uint32_t bar_x_full = bp;   bp += 8 * kMaxX;
    const uint32_t bar_x_empty = bp;  bp += 8 * kMaxX;
    const uint32_t bar_a1_full = bp;  bp += 16;
    const uint32_t bar_a1_empty = bp; bp += 16;
    const uint32_t bar_t_full = bp;   bp += 16;
    const uint32_t bar_t_empty = bp;  bp += 16;
    const uint32_t bar_a2_full = bp;  bp += 16;
    const uint32_t bar_a2_empty = bp; bp += 16;
    const uint32_t bar_w = bp;        bp += 8;
    const uint32_t bar_w2_full = bp;  bp += 8 * kMaxW2;
    const uint32_t bar_w2_empty = bp; bp += 8 * kMaxW2;
    const uint32_t bar_m_full = bp;   bp += 16;
    const uint32_t bar_m_empty = bp;

    if (threadIdx.x < 2 * a.N) {
        const int which = threadIdx.x / a.N, i = threadIdx.x % a.N;
        bias_s[which][i] = which ? a.bias2[i] : a.bias1[i];
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x_hi); prefetch_tmap(&map_w1_hi); prefetch_tmap(&map_w2_hi); prefetch_tmap(&map_y_hi); prefetch_tmap(&map_yt_hi);
        if (planes > 1) { prefetch_tmap(&map_x_lo); prefetch_tmap(&map_w1_lo); prefetch_tmap(&map_w2_lo); prefetch_tmap(&map_y_lo); prefetch_tmap(&map_yt_lo); }
    }
    if (warp == 3 && lane == 0) {
        const uint32_t nmma = (uint32_t)a.mt;
        for (int i = 0; i < a.n_x; ++i) { mbar_init(bar_x_full + 8 * i, 1); mbar_init(bar_x_empty + 8 * i, nmma + 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_a1_full + 8 * i, nmma); mbar_init(bar_a1_empty + 8 * i, 4);
            mbar_init(bar_t_full + 8 * i, 4);     mbar_init(bar_t_empty + 8 * i, nmma);
            mbar_init(bar_a2_full + 8 * i, nmma); mbar_init(bar_a2_empty + 8 * i, 4);
        }
        mbar_init(bar_w, 1);
        for (int i = 0; i < a.n_w2; ++i) { mbar_init(bar_w2_full + 8 * i, 1); mbar_init(bar_w2_empty + 8 * i, nmma); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_m_full + 8 * i, 1); mbar_init(bar_m_empty + 8 * i, 4); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(&tmem_base_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);
    const int acc_cols = a.mt * a.acc_n;                  // columns of one accumulator buffer; acc1[b] at b*acc_cols, acc2[b] at (2+b)*acc_cols
    const int n_my = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===== producer: resident weights of both convs, then the x ring =====
        if (lane == 0) {
            mbar_expect_tx(bar_w, (uint32_t)((a.k1 + (a.n_w2 ? 0 : a.k2)) * planes) * (uint32_t)a.N * row_bytes);
            for (int cv = 0; cv < (a.n_w2 ? 1 : 2); ++cv)
                for (int j = 0; j < (cv ? a.k2 : a.k1); ++j)
                    for (int pl = 0; pl < planes; ++pl)
                        tma_load_2d(smem_w + (uint32_t)(cv * a.k1 + j) * w_stage + pl * a.w_plane_bytes,
                                    cv ? (pl ? &map_w2_lo : &map_w2_hi) : (pl ? &map_w1_lo : &map_w1_hi), bar_w, 0, j * a.N);
            int sx = 0;
            uint32_t px = 0;
            pdl_wait();   // activations of the previous kernel are complete and visible from here on
            for (int t = 0; t < n_my; ++t) {
                const int tile = (int)blockIdx.x + t * (int)gridDim.x;
                if (kRagged && pair_tile_skipped(a, tile)) continue;
                const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
                const int b = tl / a.tiles_per_item;
                const int o0 = (tl - b * a.tiles_per_item) * a.V;
                const int xr0 = o0 - a.h2 - a.h1;
                mbar_wait(bar_x_empty + 8 * sx, px ^ 1u);
                mbar_expect_tx(bar_x_full + 8 * sx, (uint32_t)a.x_rows * row_bytes * planes);
                for (int pl = 0; pl < planes; ++pl)
                    for (int pc = 0; pc < a.x_pieces; ++pc)
                        tma_load_3d(smem_x + sx * x_stage + pl * a.x_plane_bytes + pc * a.x_box_rows * row_bytes,
                                    pl ? &map_x_lo : &map_x_hi, bar_x_full + 8 * sx, 0, xr0 + pc * a.x_box_rows, b);
                if (++sx == a.n_x) { sx = 0; px ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== c2 weight ring (only when both weight sets do not fit beside the tiles): k2 tap tiles per tile of work, from L2 =====
        if (lane == 0 && a.n_w2 > 0) {
            int sw = 0;
            uint32_t pw = 0;
            for (int t = 0; t < n_my; ++t) {
                if (kRagged && pair_tile_skipped(a, (int)blockIdx.x + t * (int)gridDim.x)) continue;
                for (int j = 0; j < a.k2; ++j) {
                    mbar_wait(bar_w2_empty + 8 * sw, pw ^ 1u);
                    mbar_expect_tx(bar_w2_full + 8 * sw, (uint32_t)a.N * row_bytes * planes);
                    for (int pl = 0; pl < planes; ++pl)
                        tma_load_2d(smem_w + (uint32_t)(a.k1 + sw) * w_stage + pl * a.w_plane_bytes, pl ? &map_w2_lo : &map_w2_hi,
                                    bar_w2_full + 8 * sw, 0, j * a.N);
                    if (++sw == a.n_w2) { sw = 0; pw ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===== MRF-sum tiles (only the last step of the second and later branches of a stage): R rows from output row o0 =====
        if (lane == 0 && a.has_mrf) {
            pdl_wait();
            int it = -1;
            for (int t = 0; t < n_my; ++t) {
                const int tile = (int)blockIdx.x + t * (int)gridDim.x;
                if (kRagged && pair_tile_skipped(a, tile)) continue;
                ++it;
                const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
                const int b = tl / a.tiles_per_item;
                const int o0 = (tl - b * a.tiles_per_item) * a.V;
                const int buf = it & 1;
                mbar_wait(bar_m_empty + 8 * buf, (((uint32_t)it >> 1) & 1u) ^ 1u);
                mbar_expect_tx(bar_m_full + 8 * buf, (uint32_t)a.R * row_bytes * planes);
                for (int pl = 0; pl < planes; ++pl)
                    tma_load_3d(smem_m + buf * m_stage + pl * a.m_plane_bytes, pl ? &map_m_lo : &map_m_hi, bar_m_full + 8 * buf, 0, o0, b);
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 8) {
        // ===== MMA issuers: warp 4 + cv*MT + ms issues conv cv (0: c1, 1: c2) of subtile ms =====
        // One thread cannot issue small-N MMAs at the tensor pipe's rate, and c1 of tile i+1 must overlap the
        // epilogues of tile i: c1 and c2 have their own issuers, ordered only by the barriers.
        const int role = warp - 4;
        const int cv = role / a.mt, ms = role % a.mt;
        if (cv < 2) {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc((uint32_t)a.N, kF16);
            const uint32_t idesc2 = make_idesc((uint32_t)(2 * a.N), kF16);
            const bool concat = a.concat != 0;
            const uint32_t id0 = concat ? idesc2 : idesc;
            const uint32_t dhi = desc_hi(row_bytes);
            const uint32_t sub_step = (128u * row_bytes) >> 4;
            const uint32_t a_pl = (cv ? a.t_plane_bytes : a.x_plane_bytes) >> 4, w_pl = a.w_plane_bytes >> 4;
            const bool k4 = a.N == 64;
            const uint32_t a_tap = (uint32_t)(cv ? 1 : a.d) * (row_bytes >> 4);
            const uint32_t w_tap = w_stage >> 4;
            const uint32_t w_lo = desc_lo(smem_w + (uint32_t)(cv * a.k1) * w_stage);
            const uint32_t a_stage = cv ? t_stage : x_stage;
            const uint32_t a_ring = desc_lo(cv ? smem_t : smem_x) + (uint32_t)ms * sub_step;
            const uint32_t d_base = tmem_base + (uint32_t)(cv * 2 * acc_cols + ms * a.acc_n);
            const int k = cv ? a.k2 : a.k1, n_t = a.n_t, n_x = a.n_x, n_w2 = a.n_w2;
            int sw2 = 0;
            uint32_t pw2 = 0;
            const bool dbg_nomma = a.dbg == 2;
            mbar_wait(bar_w, 0);
            int sx = 0;
            uint32_t px = 0;
            int it = -1;   // live tiles so far, minus one
            for (int t = 0; t < n_my; ++t) {
                if (kRagged && pair_tile_skipped(a, (int)blockIdx.x + t * (int)gridDim.x)) continue;
                ++it;
                const int buf = it & 1;
                const uint32_t pb = ((uint32_t)it >> 1) & 1u;
                int slot;
                if (cv == 0) {
                    slot = sx;
                    mbar_wait(bar_a1_empty + 8 * buf, pb ^ 1u);
                    mbar_wait(bar_x_full + 8 * sx, px);
                } else {
                    slot = n_t == 2 ? buf : 0;
                    const uint32_t pt = n_t == 2 ? pb : ((uint32_t)it & 1u);
                    mbar_wait(bar_a2_empty + 8 * buf, pb ^ 1u);
                    mbar_wait(bar_t_full + 8 * slot, pt);
                }
                tc_fence_after();
                const uint32_t d0 = d_base + (uint32_t)(buf * acc_cols);
                const uint32_t a_lo = a_ring + (((uint32_t)slot * a_stage) >> 4);
                const bool go = leader && !dbg_nomma;
                if (cv == 1 && n_w2 > 0) {
                    // streamed c2 weights: one barrier round trip per tap
                    const uint32_t w_ring = desc_lo(smem_w + (uint32_t)a.k1 * w_stage);
                    uint32_t aa = a_lo;
                    for (int j = 0; j < k; ++j) {
                        mbar_wait(bar_w2_full + 8 * sw2, pw2);
                        tc_fence_after();
                        const uint32_t wl = w_ring + (uint32_t)sw2 * w_tap;
                        const uint32_t acc_in = j != 0;
                        if (kPlanes == 1) {
                            if (k4) issue_taps<4, 1>(go, 1, d0, aa, wl, 0u, 0u, a_pl, w_pl, dhi, id0, idesc, acc_in);
                            else issue_taps<2, 1>(go, 1, d0, aa, wl, 0u, 0u, a_pl, w_pl, dhi, id0, idesc, acc_in);
                        } else if (concat) {
                            if (k4) issue_taps<4, 2>(go, 1, d0, aa, wl, 0u, 0u, a_pl, w_pl, dhi, id0, idesc, acc_in);
                            else issue_taps<2, 2>(go, 1, d0, aa, wl, 0u, 0u, a_pl, w_pl, dhi, id0, idesc, acc_in);
                        } else {
                            if (k4) issue_taps<4, 3>(go, 1, d0, aa, wl, 0u, 0u, a_pl, w_pl, dhi, id0, idesc, acc_in);
                            else issue_taps<2, 3>(go, 1, d0, aa, wl, 0u, 0u, a_pl, w_pl, dhi, id0, idesc, acc_in);
                        }
                        if (leader) umma_commit(bar_w2_empty + 8 * sw2);
                        if (++sw2 == n_w2) { sw2 = 0; pw2 ^= 1u; }
                        aa += a_tap;
                    }
                } else if (kPlanes == 1) {
                    if (k4) issue_taps<4, 1>(go, k, d0, a_lo, w_lo, a_tap, w_tap, a_pl, w_pl, dhi, id0, idesc);
                    else issue_taps<2, 1>(go, k, d0, a_lo, w_lo, a_tap, w_tap, a_pl, w_pl, dhi, id0, idesc);
                } else if (concat) {
                    if (k4) issue_taps<4, 2>(go, k, d0, a_lo, w_lo, a_tap, w_tap, a_pl, w_pl, dhi, id0, idesc);
                    else issue_taps<2, 2>(go, k, d0, a_lo, w_lo, a_tap, w_tap, a_pl, w_pl, dhi, id0, idesc);
                } else {
                    if (k4) issue_taps<4, 3>(go, k, d0, a_lo, w_lo, a_tap, w_tap, a_pl, w_pl, dhi, id0, idesc);
                    else issue_taps<2, 3>(go, k, d0, a_lo, w_lo, a_tap, w_tap, a_pl, w_pl, dhi, id0, idesc);
                }
                if (leader) {
                    if (cv == 0) { umma_commit(bar_a1_full + 8 * buf); umma_commit(bar_x_empty + 8 * sx); }
                    else { umma_commit(bar_a2_full + 8 * buf); umma_commit(bar_t_empty + 8 * slot); }
                }
                if (++sx == n_x) { sx = 0; px ^= 1u; }
                __syncwarp();
            }
        }
    } else if (warp >= 8 && warp < 16) {
        // ===== epilogue 1: acc1 -> t tile (activated planes of c1's output, zero outside the sequence) =====
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int grp = (warp - 8) >> 2;             // group g owns the tiles with (it & 1) == g, i.e. accumulator buffer g
        const int nch = a.N / CW;
        // With ONE t buffer the tiles must pass through epilogue 1 in order (the parity wait on bar_t_empty cannot tell "c2 of tile
        // i-1 is done" from "c2 of tile i-3 is done"): group 0 takes every tile and group 1 idles.  With two buffers each group
        // owns one and only ever waits for its own previous tile.
        int it = -1;   // live tiles so far, minus one
        for (int t = 0; t < n_my; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            if (kRagged && pair_tile_skipped(a, tile)) continue;
            ++it;
            if (a.n_t == 1 ? grp != 0 : (it & 1) != grp) continue;   // not this group's tile
            const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
            const int b = tl / a.tiles_per_item;
            const int g0 = (tl - b * a.tiles_per_item) * a.V - a.h2;      // sequence row of t tile row 0
            const int buf = it & 1;
            const int st = a.n_t == 2 ? (it & 1) : 0;
            const uint32_t pt = a.n_t == 2 ? (((uint32_t)it >> 1) & 1u) : ((uint32_t)it & 1u);
            mbar_wait(bar_a1_full + 8 * buf, ((uint32_t)it >> 1) & 1u);
            mbar_wait(bar_t_empty + 8 * st, pt ^ 1u);
            tc_fence_after();
            const uint32_t t_slot = smem_t + st * t_stage;
            // ragged batch: t rows at or behind the item's OWN end are c2's zero padding, as rows >= L are in a dense batch
            const int Lb = kRagged ? min(a.L, __ldg(a.lens + b) * a.len_mul) : a.L;
            const bool edge_tile = g0 < 0 || g0 + a.R > Lb;
            for (int ms = 0; ms < (a.dbg == 3 ? 0 : a.mt); ++ms) {
                const uint32_t row_t = (uint32_t)(ms * 128 + q * 32 + lane);
                const int g = g0 + (int)row_t;
                const bool inside = g >= 0 && g < Lb;
                const uint32_t dst = t_slot + row_t * row_bytes;
                const uint32_t sx = swz(row_t, row_bytes);
                for (int h = 0; h < nch; ++h) {
                    uint32_t r[CW];
                    const int col = h * CW;
                    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + ms * a.acc_n + col);
                    tmem_ld<CW>(tcol, r);
                    tmem_wait_ld();
                    f2 v[CW / 2];
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) {
                        const float2 bv = *reinterpret_cast<const float2*>(&bias_s[0][col + 2 * i]);
                        v[i] = f2_add(f2_bits(r[2 * i], r[2 * i + 1]), f2_pack(bv.x, bv.y));
                    }
                    if (kPlanes == 2 && a.concat) {   // second half of the concatenated accumulator: A_hi x W_lo
                        tmem_ld<CW>(tcol + (uint32_t)a.N, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < CW / 2; ++i) v[i] = f2_add(v[i], f2_bits(r[2 * i], r[2 * i + 1]));
                    }
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) v[i] = f2_lrelu(v[i]);
                    if (edge_tile) {   // warp-uniform: only tiles that overlap a sequence end carry rows outside [0, L)
#pragma unroll
                        for (int i = 0; i < CW / 2; ++i) v[i] = inside ? v[i] : 0ull;
                    }
#pragma unroll
                    for (int c = 0; c < CW / 8; ++c)
                        store_planes8<kPlanes, kF16>(&v[c * 4], dst + ((((uint32_t)(h * (CW / 8) + c)) ^ sx) << 4), a.t_plane_bytes);
                }
            }
            tc_fence_before();      // this warp has read the last of acc1[buf]
            fence_proxy_async();    // t rows written through the generic proxy are read by tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a1_empty + 8 * buf); mbar_arrive(bar_t_full + 8 * st); }
        }
    } else if (warp >= 16) {
        // ===== epilogue 2: acc2 + b2 + x (inverse lrelu of the x tile rows) -> activated planes -> TMA store =====
        const int q = warp & 3;
        const int grp = (warp - 16) >> 2;
        const int nch = a.N / CW;
        // staging box of this warp: 32 rows; C = 32 paired: two 64-byte time rows form one 128-byte row (SWIZZLE_128B)
        const uint32_t orow_bytes = a.paired ? 128u : row_bytes;
        const uint32_t o_row_off = a.paired ? (uint32_t)(lane >> 1) * 128u : (uint32_t)lane * row_bytes;
        const uint32_t o_chunk0 = a.paired ? (uint32_t)(lane & 1) * 4u : 0u;
        const uint32_t o_sx = a.paired ? (uint32_t)((lane >> 1) & 7) : swz((uint32_t)lane, orow_bytes);
        const int rshift = a.paired ? 1 : 0;
        pdl_wait();                                  // before the first global write (WAR against the previous kernel's reads)
        int so = 0;
        int it = -1;   // live tiles so far, minus one
        for (int t = 0; t < n_my; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            if (kRagged && pair_tile_skipped(a, tile)) continue;
            ++it;
            if ((it & 1) != grp) continue;   // the other group's tile
            const int sx = it % a.n_x;
            const uint32_t px = (uint32_t)(it / a.n_x) & 1u;
            const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
            const int b = tl / a.tiles_per_item;
            const int o0 = (tl - b * a.tiles_per_item) * a.V;
            const int buf = it & 1;
            mbar_wait(bar_a2_full + 8 * buf, ((uint32_t)it >> 1) & 1u);
            mbar_wait(bar_x_full + 8 * sx, px);      // completed long ago (c1 consumed it); orders this thread's reads after the TMA writes
            if (a.has_mrf) mbar_wait(bar_m_full + 8 * buf, ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            const uint32_t x_slot = smem_x + sx * x_stage;
            const uint32_t m_slot = smem_m + buf * m_stage;
            const int Lb = kRagged ? __ldg(a.lens + b) * a.len_mul : 0x7fffffff;   // ragged batch: rows from here on are written as zeros
            const f2 scale2 = f2_pack(a.out_scale, a.out_scale);
            for (int ms = 0; ms < (a.dbg == 1 ? 0 : a.mt); ++ms) {
                const int rs = ms * 128 + q * 32;                    // first tile row of this warp's box
                int nvalid = min(32, a.V - rs);
                if (o0 + rs >= a.L) nvalid = 0;
                const uint32_t row_x = (uint32_t)(rs + lane + a.h1 + a.h2);
                const uint32_t xsrc = x_slot + row_x * row_bytes;
                const uint32_t xsw = swz(row_x, row_bytes);
                const uint32_t msrc = m_slot + (uint32_t)(rs + lane) * row_bytes;
                const uint32_t msw = swz((uint32_t)(rs + lane), row_bytes);
                const uint32_t slot = smem_o + (uint32_t)((grp * 4 + q) * a.n_o + so) * o_slot;
                const bool dead = kRagged && o0 + rs + lane >= Lb;
                if (lane == 0) { if (a.n_o == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }   // the store that last used this slot has read it
                __syncwarp();
                for (int h = 0; h < nch; ++h) {
                    uint32_t r[CW];
                    const int col = h * CW;
                    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((2 + buf) * acc_cols + ms * a.acc_n + col);
                    tmem_ld<CW>(tcol, r);
                    f2 res[CW / 2];
#pragma unroll
                    for (int c = 0; c < CW / 8; ++c) {
                        const uint32_t xa = xsrc + ((((uint32_t)(h * (CW / 8) + c)) ^ xsw) << 4);
                        const uint4 ph = lds128(xa);
                        const uint32_t wh[4] = {ph.x, ph.y, ph.z, ph.w};
                        if (planes > 1) {
                            const uint4 pl = lds128(xa + a.x_plane_bytes);
                            const uint32_t wl[4] = {pl.x, pl.y, pl.z, pl.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) res[c * 4 + i] = f2_inv_lrelu(f2_add(f2_from_bf16x2(wh[i]), f2_from_bf16x2(wl[i])));
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) res[c * 4 + i] = f2_inv_lrelu(f2_from_h2<kF16>(wh[i]));
                        }
                    }
                    tmem_wait_ld();
                    f2 v[CW / 2];
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) {
                        const float2 bv = *reinterpret_cast<const float2*>(&bias_s[1][col + 2 * i]);
                        v[i] = f2_add(f2_bits(r[2 * i], r[2 * i + 1]), f2_pack(bv.x, bv.y));
                    }
                    if (kPlanes == 2 && a.concat) {
                        tmem_ld<CW>(tcol + (uint32_t)a.N, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < CW / 2; ++i) v[i] = f2_add(v[i], f2_bits(r[2 * i], r[2 * i + 1]));
                    }
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) v[i] = f2_add(v[i], res[i]);
                    if (a.has_mrf) {   // + running sum of the previous branches' outputs, then the 1/nk of the last branch
#pragma unroll
                        for (int c = 0; c < CW / 8; ++c) {
                            const uint32_t ma = msrc + ((((uint32_t)(h * (CW / 8) + c)) ^ msw) << 4);
                            const uint4 ph = lds128(ma);
                            const uint32_t wh[4] = {ph.x, ph.y, ph.z, ph.w};
                            if (planes > 1) {
                                const uint4 pl = lds128(ma + a.m_plane_bytes);
                                const uint32_t wl[4] = {pl.x, pl.y, pl.z, pl.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    v[c * 4 + i] = f2_mul(f2_add(v[c * 4 + i], f2_inv_lrelu(f2_add(f2_from_bf16x2(wh[i]), f2_from_bf16x2(wl[i])))), scale2);
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    v[c * 4 + i] = f2_mul(f2_add(v[c * 4 + i], f2_inv_lrelu(f2_from_h2<kF16>(wh[i]))), scale2);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) v[i] = (kRagged && dead) ? 0ull : f2_lrelu(v[i]);
#pragma unroll
                    for (int c = 0; c < CW / 8; ++c)
                        store_planes8<kPlanes, kF16>(&v[c * 4], slot + o_row_off + (((o_chunk0 + (uint32_t)(h * (CW / 8) + c)) ^ o_sx) << 4), a.o_plane_bytes);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (nvalid > 0 && a.dbg != 4) {
                        for (int pl = 0; pl < planes; ++pl) {
                            const CUtensorMap* m = nvalid == 32 ? (pl ? &map_y_lo : &map_y_hi) : (pl ? &map_yt_lo : &map_yt_hi);
                            tma_store_3d(m, slot + pl * a.o_plane_bytes, 0, (o0 + rs) >> rshift, b);
                        }
                    }
                    bulk_commit();
                }
                if (++so == a.n_o) so = 0;
            }
            tc_fence_before();   // this warp has read the last of acc2[buf] and of the x slot
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a2_empty + 8 * buf); mbar_arrive(bar_x_empty + 8 * sx); if (a.has_mrf) mbar_arrive(bar_m_empty + 8 * buf); }
        }
        if (lane == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn pair_encode_fn() {
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

bool pair_encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                 uint32_t inner_bytes) {
    EncodeTiledFn fn = pair_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled (pair) failed (%d) rank %d dims %llu,%llu box %u,%u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        set_error(buf);
        return false;
    }
    return true;
}

int penv(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}
uint32_t rup(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

struct PairLaunch::Impl {
    PairArgs a;
    alignas(64) CUtensorMap map_x[2], map_w1[2], map_w2[2], map_y[2], map_yt[2], map_m[2];
    int grid;
    size_t smem;
};

bool pair_supported(const PairParams& p) {
    if (penv("HFG_PAIR", 1) == 0) return false;
    if (p.C != 32 && p.C != 64) return false;
    if (p.k1 < 1 || p.k1 > 31 || p.k1 % 2 == 0 || p.k2 < 1 || p.k2 > 15 || p.k2 % 2 == 0 || p.d < 1) return false;
    if (p.npass != 1 && p.npass != 3) return false;
    if (p.f16 && p.npass != 1) return false;
    const int maxc = penv(p.npass == 3 ? "HFG_PAIR_MAXC_X3" : "HFG_PAIR_MAXC", 64);
    const int maxk = penv(p.npass == 3 ? "HFG_PAIR_MAXK_X3" : "HFG_PAIR_MAXK", 31);
    if (p.C > maxc || std::max(p.k1, p.k2) > maxk) return false;
    return true;
}

int plan_conv_pair(PairLaunch* out, const PairParams& p, int sm_count) {
    out->impl.reset();
    if (!pair_supported(p)) return HFG_ERR_UNSUPPORTED;
    std::shared_ptr<PairLaunch::Impl> I(new PairLaunch::Impl());
    PairArgs& a = I->a;
    memset(&a, 0, sizeof a);
    const int planes = p.npass > 1 ? 2 : 1;
    const int N = p.C;
    const uint32_t row_bytes = (uint32_t)N * 2u;
    a.B = p.B; a.L = p.L; a.N = N; a.k1 = p.k1; a.k2 = p.k2; a.d = p.d;
    a.h1 = p.d * (p.k1 - 1) / 2; a.h2 = (p.k2 - 1) / 2;
    a.planes = planes;
    a.concat = (planes == 2 && N == 32) ? 1 : 0;
    a.acc_n = a.concat ? 2 * N : N;
    a.paired = (N == 32 && p.L % 2 == 0 && penv("HFG_PAIR_PAIRED", 1)) ? 1 : 0;
    a.reverse = p.reverse;
    a.f16 = p.f16 ? 1 : 0;
    a.has_mrf = p.mrf_hi ? 1 : 0;
    a.out_scale = a.has_mrf ? p.out_scale : 1.0f;
    a.dbg = penv("HFG_PAIR_DBG", 0);
    a.bias1 = p.bias1; a.bias2 = p.bias2;
    a.lens = p.lens; a.len_mul = p.len_mul; a.len_skip = std::max(p.len_skip, 8);
    a.w_plane_bytes = rup((uint32_t)N * row_bytes, 1024);
    if (a.concat && a.w_plane_bytes != (uint32_t)N * row_bytes) return HFG_ERR_UNSUPPORTED;
    a.o_plane_bytes = 32u * row_bytes;
    const uint32_t budget = kSmemMax - 1024;   // alignment slack

    // Choose (MT, x ring depth, t buffers, staging slots per warp): cycles per valid output row of a tile interval, the
    // largest of tensor time (measured small-N MMA floors), HBM time and epilogue issue time, inflated when the rings are
    // too shallow for the lifetime of their slots (calibrated on tools/sweep_pair.sh runs).
    auto floor_clk = [](double n) { return std::max(128.0 * n / 256.0, (4096.0 + 32.0 * n) / 128.0); };
    const double step_clk = a.concat ? floor_clk(2.0 * N) + floor_clk(N) : p.npass * floor_clk(N);
    const int ksteps = N / 16;
    const int f_mt = penv("HFG_PAIR_MT", 0), f_nx = penv("HFG_PAIR_NX", 0), f_nt = penv("HFG_PAIR_NT", 0), f_no = penv("HFG_PAIR_NO", 0);
    bool ok = false;
    double best = 1e30;
    // c2's weights: resident like c1's, or -- when both sets leave no room for two t buffers -- streamed per tile through a small
    // ring (k2 tap tiles per tile of work from L2; a barrier round trip per tap in the c2 issuer)
    const int f_w2 = penv("HFG_PAIR_W2RING", -1);
    PairArgs resident_plan = a;
    size_t resident_smem = 0;
    bool resident_ok = false;
    for (int n_w2 : {0, std::min(p.k2, 4)}) {
      if (f_w2 >= 0 && (n_w2 != 0) != (f_w2 != 0)) continue;
      if (n_w2 > 0) {
          if ((ok && a.n_t == 2) || f_nt == 1) break;   // the resident plan already has both t buffers (or one is forced)
          if (planes > 1 && f_w2 < 0) break;            // measured: bf16x3 C = 32, k = 11, d = 5 streamed 0.79 ms vs 0.77 ms as two launches
          // otherwise any streamed plan with two t buffers is preferred; the resident one is kept in case there is none
          resident_plan = a; resident_smem = I->smem; resident_ok = ok;
          ok = false; best = 1e30;
      }
      const uint32_t w_all = (uint32_t)(p.k1 + (n_w2 ? n_w2 : p.k2)) * a.w_plane_bytes * planes;
    for (int mt = 2; mt >= 1; --mt) {
        if (f_mt && mt != f_mt) continue;
        if (4 * mt * a.acc_n > 512) continue;
        const int R = 128 * mt, V = R - (p.k2 - 1);
        if (V < 64) continue;
        const int rows_need = R + 2 * a.h1;
        const int pieces = (rows_need + 255) / 256;
        const int box_rows = ((rows_need + pieces - 1) / pieces + 7) / 8 * 8;
        const uint32_t x_plane = rup((uint32_t)(pieces * box_rows) * row_bytes, 1024);
        const uint32_t t_plane = rup((uint32_t)((R + 2 * a.h2 + 7) / 8 * 8) * row_bytes, 1024);
        const uint32_t m_plane = rup((uint32_t)R * row_bytes, 1024);
        const uint32_t m_all = a.has_mrf ? 2u * m_plane * planes : 0u;
        const double t_mma = (double)mt * (p.k1 + p.k2) * ksteps * step_clk + (n_w2 ? p.k2 * 80.0 : 0.0);
        const double t_hbm = ((double)rows_need + V + (a.has_mrf ? R : 0)) * row_bytes * planes / 20.0;
        const double t_epi = (double)mt * (N / 32) * (planes > 1 ? 1100.0 : 600.0);   // both epilogues share an SM sub-partition
        const double t_int = std::max({t_mma, t_hbm, t_epi}) + 300.0;
        for (int n_t = 2; n_t >= (n_w2 ? 2 : 1); --n_t) {
            if (f_nt && n_t != f_nt) continue;
            // each epilogue-2 warp meets one of its slots again two tiles later (the groups alternate tiles): mt slots suffice
            for (int n_o = std::min(mt, 2); n_o >= 1; --n_o) {
                if (f_no && n_o != f_no) continue;
                const uint32_t fixed = w_all + (uint32_t)n_t * t_plane * planes + 8u * n_o * a.o_plane_bytes * planes + m_all;
                if (fixed + 3 * x_plane * planes > budget) continue;
                const int n_x_max = (int)std::min<uint32_t>((budget - fixed) / (x_plane * planes), (uint32_t)kMaxX);
                for (int n_x = n_x_max; n_x >= 3; --n_x) {
                    if (f_nx && n_x != f_nx) continue;
                    // an x slot lives from its load (~2500 cycles of HBM latency) through c1, epilogue 1, c2 and epilogue 2 of its
                    // tile (the residual is read from it): about three intervals.  Measured: one t buffer costs ~1.4x (epilogue 1 of
                    // tile i+1 waits for c2 of tile i), a staging slot shared by the two steps of a tile ~1.08x.
                    const double f_x = std::min(1.0, n_x / (3.0 + 2500.0 / t_int));
                    const long tiles = (long)((p.L + V - 1) / V) * p.B;
                    const long waves = (tiles + sm_count - 1) / std::max(1, sm_count);
                    const double fill = (double)tiles / (double)(waves * std::max(1, sm_count));
                    const double cost = t_int / f_x / V / fill * (n_t == 1 ? 1.4 : 1.0) * (n_o < mt ? 1.08 : 1.0);
                    if (cost < best - 1e-9) {
                        best = cost; ok = true;
                        a.mt = mt; a.R = R; a.V = V;
                        a.x_rows = pieces * box_rows; a.x_box_rows = box_rows; a.x_pieces = pieces;
                        a.x_plane_bytes = x_plane; a.t_plane_bytes = t_plane;
                        a.n_x = n_x; a.n_t = n_t; a.n_o = n_o; a.n_w2 = n_w2;
                        a.off_w = (uint32_t)n_x * x_plane * planes;
                        a.off_t = a.off_w + w_all;
                        a.off_o = a.off_t + (uint32_t)n_t * t_plane * planes;
                        a.off_m = a.off_o + 8u * n_o * a.o_plane_bytes * planes;
                        a.m_plane_bytes = m_plane;
                        // > half an SM, so exactly one CTA (512 TMEM columns) lives on an SM
                        I->smem = std::max<size_t>((size_t)a.off_m + m_all + 1024, 120u * 1024u);
                    }
                }
            }
        }
    }
    }
    if (!ok && resident_ok) { a = resident_plan; I->smem = resident_smem; ok = true; }
    if (!ok) return HFG_ERR_UNSUPPORTED;
    // One t buffer serialises epilogue 1 behind c2 of the previous tile (and leaves one epilogue-1 group idle): measured slower than
    // the two-launch plan (C = 64, k = 7: 0.256 vs 0.245 ms; bf16x3 C = 32, k = 11, d = 5: 0.80 vs 0.77 ms).  Leave those unfused.
    if (a.n_t == 1 && !f_nt && !penv("HFG_PAIR_ALLOW_NT1", 0)) return HFG_ERR_UNSUPPORTED;
    if (penv("HFG_PAIR_VERBOSE", 0))
        fprintf(stderr, "pair plan C=%d k=%d,%d d=%d planes=%d: mt=%d V=%d n_x=%d n_t=%d n_o=%d n_w2=%d x_rows=%d smem=%zu cost=%.2f\n", N, p.k1, p.k2, p.d,
                planes, a.mt, a.V, a.n_x, a.n_t, a.n_o, a.n_w2, a.x_rows, I->smem, best);
    a.tiles_per_item = (p.L + a.V - 1) / a.V;
    a.total_tiles = a.tiles_per_item * p.B;
    I->grid = std::min(a.total_tiles, std::max(1, sm_count));

    const uint64_t dims3[3] = {(uint64_t)N, (uint64_t)p.L, (uint64_t)p.B};
    const uint64_t str3[2] = {(uint64_t)N * 2, (uint64_t)p.L * N * 2};
    {
        const uint32_t box[3] = {(uint32_t)N, (uint32_t)a.x_box_rows, 1};
        if (!pair_encode(&I->map_x[0], p.x_hi, 3, dims3, str3, box, row_bytes)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_x[1], planes > 1 ? p.x_lo : p.x_hi, 3, dims3, str3, box, row_bytes)) return HFG_ERR_CUDA;
        const uint32_t mbox[3] = {(uint32_t)N, (uint32_t)a.R, 1};
        const void* m0p = a.has_mrf ? (const void*)p.mrf_hi : (const void*)p.x_hi;
        const void* m1p = (a.has_mrf && planes > 1) ? (const void*)p.mrf_lo : m0p;
        if (!pair_encode(&I->map_m[0], m0p, 3, dims3, str3, mbox, row_bytes)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_m[1], m1p, 3, dims3, str3, mbox, row_bytes)) return HFG_ERR_CUDA;
    }
    {
        const uint64_t dims1[2] = {(uint64_t)N, (uint64_t)p.k1 * N};
        const uint64_t dims2[2] = {(uint64_t)N, (uint64_t)p.k2 * N};
        const uint64_t str[1] = {(uint64_t)N * 2};
        const uint32_t box[2] = {(uint32_t)N, (uint32_t)N};
        if (!pair_encode(&I->map_w1[0], p.w1_hi, 2, dims1, str, box, row_bytes)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_w1[1], planes > 1 ? p.w1_lo : p.w1_hi, 2, dims1, str, box, row_bytes)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_w2[0], p.w2_hi, 2, dims2, str, box, row_bytes)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_w2[1], planes > 1 ? p.w2_lo : p.w2_hi, 2, dims2, str, box, row_bytes)) return HFG_ERR_CUDA;
    }
    {
        const int pr = a.paired ? 2 : 1;
        const int tail = 32 - (p.k2 - 1);   // rows of the one partial box per tile (k2-1 is even, so paired boxes stay whole)
        const uint64_t ydims[3] = {(uint64_t)N * pr, (uint64_t)p.L / pr, (uint64_t)p.B};
        const uint64_t ystr[2] = {(uint64_t)N * pr * 2, (uint64_t)p.L * N * 2};
        const uint32_t ybox[3] = {(uint32_t)N * pr, (uint32_t)(32 / pr), 1};
        const uint32_t tbox[3] = {(uint32_t)N * pr, (uint32_t)(tail / pr), 1};
        const uint32_t yb = (uint32_t)N * pr * 2u;
        if (!pair_encode(&I->map_y[0], p.y_hi, 3, ydims, ystr, ybox, yb)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_y[1], planes > 1 ? p.y_lo : p.y_hi, 3, ydims, ystr, ybox, yb)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_yt[0], p.y_hi, 3, ydims, ystr, tbox, yb)) return HFG_ERR_CUDA;
        if (!pair_encode(&I->map_yt[1], planes > 1 ? p.y_lo : p.y_hi, 3, ydims, ystr, tbox, yb)) return HFG_ERR_CUDA;
    }
    out->impl = I;
    out->mt = a.mt; out->n_x = a.n_x; out->n_t = a.n_t; out->n_o = a.n_o; out->smem = I->smem; out->grid = I->grid;
    return HFG_OK;
}

cudaError_t launch_conv_pair(const PairLaunch& L, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev % 64]) {
        cudaError_t e = cudaSuccess;
#define HFG_PAIR_ATTR(P, F)                                                                                                                    \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_pair_kernel<P, F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_pair_kernel<P, F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
        HFG_PAIR_ATTR(1, false); HFG_PAIR_ATTR(1, true); HFG_PAIR_ATTR(2, false);
#undef HFG_PAIR_ATTR
        if (e != cudaSuccess) return e;
        configured[dev % 64] = true;
    }
    const PairLaunch::Impl& I = *L.impl;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(I.grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = I.smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const int use_pdl = penv("HFG_PDL", 1);
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    cudaError_t e;
#define HFG_PAIR_LAUNCH1(P, F, G)                                                                                                       \
    e = cudaLaunchKernelEx(&cfg, conv_pair_kernel<P, F, G>, I.map_x[0], I.map_x[1], I.map_w1[0], I.map_w1[1], I.map_w2[0], I.map_w2[1], \
                           I.map_y[0], I.map_y[1], I.map_yt[0], I.map_yt[1], I.map_m[0], I.map_m[1], I.a)
#define HFG_PAIR_LAUNCH(P, F) \
    do { if (I.a.lens) HFG_PAIR_LAUNCH1(P, F, true); else HFG_PAIR_LAUNCH1(P, F, false); } while (0)
    if (I.a.planes == 2) HFG_PAIR_LAUNCH(2, false);
    else if (I.a.f16) HFG_PAIR_LAUNCH(1, true);
    else HFG_PAIR_LAUNCH(1, false);
#undef HFG_PAIR_LAUNCH
#undef HFG_PAIR_LAUNCH1
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace hfg
