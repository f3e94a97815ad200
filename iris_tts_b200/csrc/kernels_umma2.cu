// conv_umma2: persistent, software-pipelined tcgen05 implicit-GEMM Conv1d on bf16 operand planes.
//
// Same math as conv_umma_kernel (kernels_umma.cu) for the plain (non-transposed) convolutions
//   F.conv1d  src/iris/hifigan_pretrained.py:67,69   (ResBlock convs1 / convs2)
// but organised so that an SM never idles between tiles:
//
//   * one CTA per SM, looping over (batch item, 128*MT-row block) tiles; N = C_out (<= 256) in one tile
//   * weights RESIDENT in shared memory for the whole kernel when they fit (C <= 64, and C = 128 with k = 3),
//     otherwise streamed through a TMA ring as in v1
//   * A halo tiles (one per 64-channel K-chunk) prefetched through an NA-deep TMA ring, taps = row-shifted
//     UMMA descriptors into that tile
//   * TWO accumulator buffers in TMEM (2 * MT * N <= 512 columns): the MMA warp fills buffer (i+1)&1 while
//     the epilogue warps drain buffer i&1
//   * epilogue through shared memory: the residual operand planes arrive by TMA into a staging ring, each
//     thread (one TMEM lane = one time row) combines  acc + bias + x  in place, and the activated bf16
//     plane(s) leave by TMA store (coalesced, asynchronous, rows >= L clipped by the tensor map)
//
// Activations are carried ONLY as activated planes  P(x) = bf16(lrelu(x))  (+ lo = bf16(lrelu(x) - hi) in
// bf16x3 mode).  leaky_relu is a bijection, so the residual  x = xt + x  (:70) recovers x from the plane:
// x = p > 0 ? p : p / slope.  No fp32 activation stream exists in the tensor-core modes.
//
// Warp roles: 0 = A/W TMA producer, 1 = residual TMA producer, 2 = TMEM allocator, 4..7 = epilogue, 8..11 = MMA issuers
// (issuer w owns the 128-row subtile w of every tile: small-N MMAs are issue-bound from one thread).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hfg_internal.h"
#include "umma_ptx.cuh"

namespace hfg {

namespace {

using namespace ptx;

// warps: 0 A/W producer, 1 residual producer, 2 TMEM alloc, 4-7 epilogue group 0, 8-11 MMA issuers, 12-15 epilogue group 1
// (single-plane mode only: the two-plane epilogue needs more than the 128 registers a 512-thread CTA leaves per thread)
constexpr int threads_for(int planes) { return planes == 1 ? 512 : 384; }
constexpr int kMaxA = 6, kMaxW = 12, kMaxE = 8;
constexpr uint32_t kSmemBudget = 227u * 1024u - 4096u;   // dynamic smem; static barriers/bias live outside

struct K2Args {
    int B, L, N;
    int kc, nchunks, taps, tap_off0, tap_step, lo;
    int planes, npass, mt;
    int tiles_per_item, total_tiles;
    int rows_a, a_box_rows, a_pieces;
    int n_a, n_w, n_e, w_resident, has_res;
    int has_mrf;      // the epilogue also adds a second plane set: the running MRF sum of the previous branches (hifigan_pretrained.py:133-136)
    float out_scale;  // with has_mrf: v *= out_scale before the activation (1 / num_kernels on the last branch, :137)
    int ecols, groups;
    int ups;      // polyphase ConvTranspose1d (k = 2s, N = s*C_out <= 256): GEMM row m, column half h lands on output half-row
                  // 2m - 1 + h of [B][2*L_in][N/2]; the two column groups ARE the halves, stored through a 4-D tensor map
    int n_tiles, n_total;   // wide upsampler: the N = s*C_out columns are walked in n_tiles tiles of a.N columns (n_total = n_tiles * a.N)
    int cout;     // bias index = column % cout
    int reverse;  // walk the tiles last-to-first: consecutive kernels alternate, so the part of the input the previous kernel wrote
                  // last (still in the 126 MB L2) is the part this kernel reads first
    int concat;   // bf16x3, 2N <= 128: pass 1 = A_hi x [W_hi ; W_lo] (one MMA of width 2N), pass 2 = A_lo x W_hi; the epilogue adds the halves
    int acc_n;    // TMEM columns of one 128-row subtile accumulator (N, or 2N when concat)
    int paired;   // C = 32: residual / output boxes address two 64-byte time rows as one 128-byte row (full-line TMA requests)
    int f16;   // operand planes are fp16 instead of bf16 (single-plane mode only)
    int dbg;   // HFG_U2_DBG (timing experiments only): 1 = epilogue does no work, 2 = MMA warp issues no MMAs
    const int32_t* lens;   // ragged batch (device, [B] mel frames): output rows at or behind item b's own end, lens[b] * len_mul, are
    int len_mul;           // written as zeros (what the next layer must read there); nullptr: dense batch.  Upsampler: in half-rows
    int len_skip;          // a tile whose first (half-)row lies this far behind the item's end is not computed at all: nothing inside
                           // an item reads that far (>= the widest tap span of any layer), so a ragged batch costs its real frames
    uint32_t a_plane_bytes, w_plane_bytes, e_plane_bytes;
    uint32_t off_w, off_e;
    const float* bias;
    __nv_bfloat16* y_hi;   // raw output planes (upsampler: the one box that starts before the sequence is stored directly)
    __nv_bfloat16* y_lo;
};

__device__ __forceinline__ float inv_lrelu(float p) { return p > 0.f ? p : p * (1.0f / kLreluSlope); }

template <bool F16 = false>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2<F16>(w[i], f[2 * i], f[2 * i + 1]);
}

// One K-chunk of one 128-row subtile: all taps x passes x K=16 slices.  The loop is warp-uniform and lives in uniform registers
// (descriptors advance by one add per tap); only the elected lane issues.  KS = K=16 slices per tap; NP = passes per tap:
// 1 bf16 | 3 bf16x3 (hi,hi)(lo,hi)(hi,lo) | 2 concat (hi,[hi;lo])(lo,hi).  The generic loop this replaces rebuilt descriptors and
// instruction words per MMA (~19 instructions, 100-130 cycles per MMA from one thread).
// Ragged batch: tile (item b, first GEMM row m0) lies wholly behind the rows anything inside the item can read.  Every role
// evaluates this for every tile and skips the same ones (ring positions and accumulator parities count live tiles only).
__device__ __forceinline__ bool tile_skipped(const K2Args& a, int b, int m0) {
    return (a.ups ? 2 * m0 - 1 : m0) >= __ldg(a.lens + b) * a.len_mul + a.len_skip;
}

template <int KS, int NP>
__device__ __forceinline__ void issue_taps_resident(bool leader, int taps, uint32_t d0, uint32_t a_lo, uint32_t w_lo, uint32_t a_tap,
                                                    uint32_t w_tap, uint32_t a_pl, uint32_t w_pl, uint32_t dhi, uint32_t id0, uint32_t id1,
                                                    uint32_t acc_in) {
#pragma unroll 1
    for (int j = 0; j < taps; ++j) {
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
// Same with the W tile of every tap arriving through the TMA ring (one barrier round trip per tap).
template <int KS, int NP>
__device__ __forceinline__ void issue_taps_streamed(bool leader, int taps, uint32_t d0, uint32_t a_lo, uint32_t w_ring, uint32_t a_tap,
                                                    uint32_t w_stage16, uint32_t a_pl, uint32_t w_pl, uint32_t dhi, uint32_t id0, uint32_t id1,
                                                    uint32_t acc_in, uint32_t bar_w_full, uint32_t bar_w_empty, int n_w, int& sw, uint32_t& pw) {
#pragma unroll 1
    for (int j = 0; j < taps; ++j) {
        mbar_wait(bar_w_full + 8 * sw, pw);
        tc_fence_after();
        const uint32_t w_lo = w_ring + (uint32_t)sw * w_stage16;
        if (leader) {
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
                const uint32_t aa = a_lo + (ps == 1 ? a_pl : 0u);
                const uint32_t ww = w_lo + (ps == 2 ? w_pl : 0u);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    umma_bf16_lh(d0, aa + 2u * ks, ww + 2u * ks, dhi, ps == 0 ? id0 : id1, (ps | ks) ? 1u : (acc_in | (uint32_t)(j != 0)));
            }
            umma_commit(bar_w_empty + 8 * sw);
        }
        if (++sw == n_w) { sw = 0; pw ^= 1u; }
        a_lo += a_tap;
    }
}

// kRagged: the per-item masking of hfg_forward_ragged is a separate instantiation -- the dense kernels carry none of its code (a
// select per output value in the issue-bound epilogue cost the bf16 forward 2 % when it was a run-time flag).
template <int kPlanes, bool kHasRes, bool kF16, bool kRagged>
__global__ void __launch_bounds__(threads_for(kPlanes), 1)
conv_umma2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                  const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                  const __grid_constant__ CUtensorMap map_m_hi, const __grid_constant__ CUtensorMap map_m_lo,
                  const __grid_constant__ CUtensorMap map_y_hi, const __grid_constant__ CUtensorMap map_y_lo, const K2Args a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxA + 2 * kMaxW + 2 * kMaxE + 5];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[256];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler too: role values stay in uniform registers
    const int lane = threadIdx.x & 31;
    constexpr int planes = kPlanes;          // 1: bf16 ; 2: bf16x3 (hi + lo operand planes, three MMA passes)
    constexpr int npass = kPlanes == 2 ? 3 : 1;
    constexpr int kEpiGroups = kPlanes == 1 ? 2 : 1;   // epilogue warp groups; steps alternate between them
    const uint32_t row_bytes = (uint32_t)a.kc * 2u;
    const uint32_t erow_bytes = a.paired ? 128u : (uint32_t)a.ecols * 2u;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_stage_bytes = a.a_plane_bytes * planes;
    const uint32_t w_stage_bytes = a.w_plane_bytes * planes;
    const uint32_t e_slot_bytes = a.e_plane_bytes * planes * (a.has_mrf ? 2u : 1u);   // [residual / output planes][MRF-sum planes]
    const uint32_t e_mrf_off = a.e_plane_bytes * planes;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_w = smem_base + a.off_w;
    const uint32_t smem_e = smem_base + a.off_e;

    uint32_t bp = smem_u32(&bars[0]);
    const uint32_t bar_a_full = bp;            bp += 8 * kMaxA;
    const uint32_t bar_a_empty = bp;           bp += 8 * kMaxA;
    const uint32_t bar_w_full = bp;            bp += 8 * kMaxW;
    const uint32_t bar_w_empty = bp;           bp += 8 * kMaxW;
    const uint32_t bar_e_full = bp;            bp += 8 * kMaxE;
    const uint32_t bar_e_empty = bp;           bp += 8 * kMaxE;
    const uint32_t bar_acc_full = bp;          bp += 16;
    const uint32_t bar_acc_empty = bp;         bp += 16;
    const uint32_t bar_wres = bp;

    for (int i = threadIdx.x; i < 256; i += blockDim.x) bias_s[i] = a.bias[i % a.cout];   // column n of the GEMM -> bias[n % cout]
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_y_hi);
        if (kHasRes) { prefetch_tmap(&map_r_hi); if (a.has_mrf) { prefetch_tmap(&map_m_hi); if (planes > 1) prefetch_tmap(&map_m_lo); } }
        if (planes > 1) { prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_lo); prefetch_tmap(&map_y_lo); if (kHasRes) prefetch_tmap(&map_r_lo); }
    }
    if (warp == 3 && lane == 0) {
        const uint32_t nmma = (uint32_t)a.mt;   // one issuing warp per 128-row subtile
        for (int i = 0; i < a.n_a; ++i) { mbar_init(bar_a_full + 8 * i, 1); mbar_init(bar_a_empty + 8 * i, nmma); }
        for (int i = 0; i < a.n_w; ++i) { mbar_init(bar_w_full + 8 * i, 1); mbar_init(bar_w_empty + 8 * i, nmma); }
        for (int i = 0; i < a.n_e; ++i) { mbar_init(bar_e_full + 8 * i, 1); mbar_init(bar_e_empty + 8 * i, 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, nmma); mbar_init(bar_acc_empty + 8 * i, 4 * kEpiGroups); }
        mbar_init(bar_wres, 1);
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
    const int acc_cols = a.mt * a.acc_n;   // columns of one accumulator buffer
    // Programmatic dependent launch: the next kernel of the plan may start its own prologue on SMs this grid has left;
    // nothing above reads or writes an activation.  Weights are static, so the resident-W loads below precede the wait.
    pdl_launch_dependents();

    if (warp == 0) {
        // ===== A / W producer =====
        if (lane == 0) {
            if (a.w_resident) {
                mbar_expect_tx(bar_wres, (uint32_t)(a.nchunks * a.taps * planes) * (uint32_t)a.N * row_bytes);
                for (int c = 0; c < a.nchunks; ++c)
                    for (int j = 0; j < a.taps; ++j)
                        for (int pl = 0; pl < planes; ++pl)
                            tma_load_2d(smem_w + (uint32_t)(c * a.taps + j) * w_stage_bytes + pl * a.w_plane_bytes,
                                        pl ? &map_w_lo : &map_w_hi, bar_wres, c * a.kc, j * a.N);
            }
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            pdl_wait();   // activations of the previous kernel are complete and visible from here on
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
                const int tm = tl / a.n_tiles, n0 = (tl - tm * a.n_tiles) * a.N;   // column tile fastest: neighbours share the A tile in L2
                const int b = tm / a.tiles_per_item;
                const int m0 = (tm - b * a.tiles_per_item) * a.mt * 128;
                if (kRagged && tile_skipped(a, b, m0)) continue;
                for (int c = 0; c < a.nchunks; ++c) {
                    if (a.dbg != 3) {
                        mbar_wait(bar_a_empty + 8 * sa, pa ^ 1u);
                        mbar_expect_tx(bar_a_full + 8 * sa, (uint32_t)a.rows_a * row_bytes * planes);
                        for (int pl = 0; pl < planes; ++pl)
                            for (int pc = 0; pc < a.a_pieces; ++pc)
                                tma_load_3d(smem_a + sa * a_stage_bytes + pl * a.a_plane_bytes + pc * a.a_box_rows * row_bytes,
                                            pl ? &map_a_lo : &map_a_hi, bar_a_full + 8 * sa, c * a.kc, m0 + a.lo + pc * a.a_box_rows, b);
                        if (++sa == a.n_a) { sa = 0; pa ^= 1u; }
                    }
                    if (!a.w_resident) {
                        for (int j = 0; j < a.taps; ++j) {
                            mbar_wait(bar_w_empty + 8 * sw, pw ^ 1u);
                            mbar_expect_tx(bar_w_full + 8 * sw, (uint32_t)a.N * row_bytes * planes);
                            for (int pl = 0; pl < planes; ++pl)
                                tma_load_2d(smem_w + sw * w_stage_bytes + pl * a.w_plane_bytes, pl ? &map_w_lo : &map_w_hi,
                                            bar_w_full + 8 * sw, c * a.kc, j * a.n_total + n0);
                            if (++sw == a.n_w) { sw = 0; pw ^= 1u; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 8 && warp < 12) {
        // ===== MMA issuers =====
        // Issuer warp w owns subtile ms = w of every tile.  The whole warp walks the loops (warp-uniform values); only the
        // elected lane issues tcgen05.mma / tcgen05.commit.  Descriptors advance by one add per MMA.
        const int my_ms = warp - 8;
        if (my_ms < a.mt) {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc((uint32_t)a.N, kF16);
            const uint32_t idesc2 = make_idesc((uint32_t)(2 * a.N), kF16);
            const bool concat = a.concat != 0;
            const uint32_t id0 = concat ? idesc2 : idesc;
            const uint32_t dhi = desc_hi(row_bytes);
            const uint32_t sub_step = (128u * row_bytes) >> 4;          // descriptor step between 128-row subtiles
            const uint32_t a_pl = a.a_plane_bytes >> 4, w_pl = a.w_plane_bytes >> 4;
            const bool k4 = a.kc == 64;
            const bool resident = a.w_resident != 0;
            const int taps = a.taps, nchunks = a.nchunks, n_a = a.n_a, n_w = a.n_w;
            const uint32_t row16 = row_bytes >> 4;
            const uint32_t a_tap = (uint32_t)a.tap_step * row16;        // may be negative (polyphase upsampler): wraps correctly
            const uint32_t a_first = (uint32_t)(a.tap_off0 - a.lo) * row16 + (uint32_t)my_ms * sub_step;
            const uint32_t a_ring = desc_lo(smem_a), a_stage16 = a_stage_bytes >> 4;
            const uint32_t w_ring = desc_lo(smem_w), w_stage16 = w_stage_bytes >> 4;
            const bool go = leader && a.dbg != 2;
            const bool dbg3 = a.dbg == 3;
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            if (resident) mbar_wait(bar_wres, 0);
            int it = 0;   // live tiles of this CTA so far
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                if (kRagged) {
                    const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
                    const int tm = tl / a.n_tiles, b = tm / a.tiles_per_item;
                    if (tile_skipped(a, b, (tm - b * a.tiles_per_item) * a.mt * 128)) continue;
                }
                const int buf = it & 1;
                mbar_wait(bar_acc_empty + 8 * buf, (((uint32_t)it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(buf * acc_cols + my_ms * a.acc_n);
                for (int c = 0; c < nchunks; ++c) {
                    if (!dbg3) mbar_wait(bar_a_full + 8 * sa, pa);
                    tc_fence_after();
                    const uint32_t a_lo = a_ring + (uint32_t)sa * a_stage16 + a_first;
                    const uint32_t acc_in = c != 0;
#define HFG_ISSUE(KS, NP)                                                                                                           \
    do {                                                                                                                            \
        if (resident)                                                                                                               \
            issue_taps_resident<KS, NP>(go, taps, d0, a_lo, w_ring + (uint32_t)(c * taps) * w_stage16, a_tap, w_stage16, a_pl, w_pl, dhi, \
                                        id0, idesc, acc_in);                                                                        \
        else                                                                                                                        \
            issue_taps_streamed<KS, NP>(go, taps, d0, a_lo, w_ring, a_tap, w_stage16, a_pl, w_pl, dhi, id0, idesc, acc_in, bar_w_full,  \
                                        bar_w_empty, n_w, sw, pw);                                                                  \
    } while (0)
                    if (kPlanes == 1) { if (k4) HFG_ISSUE(4, 1); else HFG_ISSUE(2, 1); }
                    else if (concat) { if (k4) HFG_ISSUE(4, 2); else HFG_ISSUE(2, 2); }
                    else { if (k4) HFG_ISSUE(4, 3); else HFG_ISSUE(2, 3); }
#undef HFG_ISSUE
                    if (leader) umma_commit(bar_a_empty + 8 * sa);
                    if (++sa == n_a) { sa = 0; pa ^= 1u; }
                }
                if (leader) umma_commit(bar_acc_full + 8 * buf);
                __syncwarp();
                ++it;
            }
        }
    } else if (warp == 1) {
        // ===== residual producer: one [128 x ecols] box of the residual planes per epilogue step =====
        if (lane == 0 && kHasRes && a.dbg != 1 && a.dbg != 3) {
            int se = 0;
            uint32_t pe = 0;
            pdl_wait();
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;   // (a residual implies n_tiles == 1)
                const int b = tl / a.tiles_per_item;
                const int m0 = (tl - b * a.tiles_per_item) * a.mt * 128;
                if (kRagged && tile_skipped(a, b, m0)) continue;
                for (int ms = 0; ms < a.mt; ++ms) {
                    const int row0 = m0 + ms * 128;
                    if (row0 >= a.L) break;
                    for (int g = 0; g < a.groups; ++g) {
                        mbar_wait(bar_e_empty + 8 * se, pe ^ 1u);
                        mbar_expect_tx(bar_e_full + 8 * se, e_slot_bytes);
                        for (int pl = 0; pl < planes; ++pl)
                            tma_load_3d(smem_e + se * e_slot_bytes + pl * a.e_plane_bytes, pl ? &map_r_lo : &map_r_hi,
                                        bar_e_full + 8 * se, g * a.ecols, a.paired ? row0 >> 1 : row0, b);
                        if (a.has_mrf)
                            for (int pl = 0; pl < planes; ++pl)
                                tma_load_3d(smem_e + se * e_slot_bytes + e_mrf_off + pl * a.e_plane_bytes, pl ? &map_m_lo : &map_m_hi,
                                            bar_e_full + 8 * se, g * a.ecols, a.paired ? row0 >> 1 : row0, b);
                        if (++se == a.n_e) { se = 0; pe ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue =====
        // Each warp owns 32 rows (its TMEM lane quarter) of every 128-row subtile and runs independently of the other
        // three: its rows of the staging box are combined in place and leave through its own 32-row TMA store.
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;               // row of the 128-row subtile this thread owns
        // 16-byte chunk swizzle of this row inside a TMA box: 128-byte rows XOR (row & 7); 64-byte rows XOR ((row >> 1) & 3)
        // (paired: row r is the (r & 1) half of 128-byte row r >> 1)
        const uint32_t sw_xor = a.paired ? (uint32_t)((row >> 1) & 7) : (erow_bytes == 128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3));
        const uint32_t row_off = a.paired ? (uint32_t)(row >> 1) * 128u : (uint32_t)row * erow_bytes;
        const uint32_t chunk0 = a.paired ? (uint32_t)(row & 1) * 4u : 0u;
        const uint32_t warp_off = (uint32_t)(q * 32) * (uint32_t)a.ecols * 2u;
        const int rshift = a.paired ? 1 : 0;
        const int halves = a.ecols / 32;
        // A slot is handed back `depth` of this warp's own steps after its store was issued; it must be back before the warp
        // meets the slot again, i.e. depth <= (own steps between two uses of a slot) - 1.
        // With two groups a warp meets only every other step, and step s may reuse the slot of step s - n_e only if that
        // slot's owner has handed it back: 2 * depth < n_e global steps (one group: depth < n_e).
        const int depth = min(2, kEpiGroups == 2 ? (a.n_e - 1) / 2 : a.n_e - 1);
        pdl_wait();                                  // before the first global write (WAR against the previous kernel's reads)
        int se = 0;
        uint32_t pe = 0;
        int hist[2] = {-1, -1};                      // slots of the last `depth` stores (lane 0)
        const int grp = warp >= 12 ? 1 : 0;          // epilogue group of this warp
        int step = 0;                                // (tile, subtile, column group) steps alternate between the groups
        int it = -1;   // live tiles of this CTA so far, minus one
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            const int tl = a.reverse ? a.total_tiles - 1 - tile : tile;
            const int tm = tl / a.n_tiles, n0 = (tl - tm * a.n_tiles) * a.N;
            const int b = tm / a.tiles_per_item;
            const int m0 = (tm - b * a.tiles_per_item) * a.mt * 128;
            if (kRagged && tile_skipped(a, b, m0)) continue;
            ++it;
            const int boff = n0 % a.cout;                     // bias of tile column c: bias_s[boff + c]
            const int buf = it & 1;
            const int lim = kRagged ? __ldg(a.lens + b) * a.len_mul : 0x7fffffff;   // first (half-)row behind this item's own end
            mbar_wait(bar_acc_full + 8 * buf, ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            if (a.dbg == 1 || a.dbg == 3) {   // timing experiment: drain nothing
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
                continue;
            }
            for (int ms = 0; ms < a.mt; ++ms) {
                const int row0 = m0 + ms * 128;
                if (row0 >= a.L) break;
                for (int g = 0; g < a.groups; ++g) {
                    if (kEpiGroups > 1 && ((step++ & 1) != grp)) {   // the other group's step: only track the ring position
                        if (++se == a.n_e) { se = 0; pe ^= 1u; }
                        continue;
                    }
                    if (kHasRes) mbar_wait(bar_e_full + 8 * se, pe);
                    else mbar_wait(bar_e_empty + 8 * se, pe ^ 1u);
                    const uint32_t slot = smem_e + se * e_slot_bytes;
                    // upsampler: column n0 + g*ecols of the s*C_out GEMM columns -> (half-row parity, column inside the half-row)
                    const int ucol = n0 + g * a.ecols;
                    const int uhalf = a.ups ? ucol / (a.n_total / 2) : 0;
                    const int ucg = a.ups ? ucol - uhalf * (a.n_total / 2) : 0;
                    const bool dead = kRagged && (a.ups ? 2 * (row0 + row) - 1 + uhalf : row0 + row) >= lim;   // behind the item's end
                    for (int h = 0; h < (a.dbg == 5 ? 0 : halves); ++h) {
                        uint32_t r[32];
                        const int col = g * a.ecols + h * 32;
                        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + ms * a.acc_n + col);
                        tmem_ld32(tcol, r);
                        uint32_t addr[4];
#pragma unroll
                        for (int cidx = 0; cidx < 4; ++cidx)
                            addr[cidx] = slot + row_off + (((chunk0 + (uint32_t)(h * 4 + cidx)) ^ sw_xor) << 4);
                        float res[32];
                        if (kHasRes) {
#pragma unroll
                            for (int cidx = 0; cidx < 4; ++cidx) {
                                float f[8];
                                unpack8<kF16>(lds128(addr[cidx]), f);
                                if (planes > 1) {
                                    float fl[8];
                                    unpack8(lds128(addr[cidx] + a.e_plane_bytes), fl);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) f[i] += fl[i];
                                }
#pragma unroll
                                for (int i = 0; i < 8; ++i) res[cidx * 8 + i] = inv_lrelu(f[i]);
                            }
                        }
                        tmem_wait_ld();
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 bv = *reinterpret_cast<const float4*>(&bias_s[boff + col + 4 * i]);
                            v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bv.x;
                            v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bv.y;
                            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bv.z;
                            v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bv.w;
                        }
                        if (kPlanes == 2 && a.concat) {   // second half of the concatenated accumulator: A_hi x W_lo
                            tmem_ld32(tcol + (uint32_t)a.N, r);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
                        }
                        if (kHasRes) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] += res[i];
                            if (a.has_mrf) {   // + running sum of the previous branches' outputs, then the 1/nk of the last branch
#pragma unroll
                                for (int cidx = 0; cidx < 4; ++cidx) {
                                    float f[8];
                                    unpack8<kF16>(lds128(addr[cidx] + e_mrf_off), f);
                                    if (planes > 1) {
                                        float fl[8];
                                        unpack8(lds128(addr[cidx] + e_mrf_off + a.e_plane_bytes), fl);
#pragma unroll
                                        for (int i = 0; i < 8; ++i) f[i] += fl[i];
                                    }
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[cidx * 8 + i] = (v[cidx * 8 + i] + inv_lrelu(f[i])) * a.out_scale;
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = (kRagged && dead) ? 0.f : lrelu(v[i]);
#pragma unroll
                        for (int cidx = 0; cidx < 4; ++cidx) {
                            uint4 hi;
                            hi.x = pack2<kF16>(v[cidx * 8 + 0], v[cidx * 8 + 1]);
                            hi.y = pack2<kF16>(v[cidx * 8 + 2], v[cidx * 8 + 3]);
                            hi.z = pack2<kF16>(v[cidx * 8 + 4], v[cidx * 8 + 5]);
                            hi.w = pack2<kF16>(v[cidx * 8 + 6], v[cidx * 8 + 7]);
                            sts128(addr[cidx], hi);
                            const bool direct = a.ups && uhalf == 0 && row0 + q * 32 == 0;   // box would start at q = -1: TMA stores fault there
                            const bool dwrite = direct && row >= 1 && row < a.L;   // GEMM rows 1 .. L_in of this box (plain stores are not clipped)
                            const size_t doff = dwrite
                                ? ((((size_t)b * (a.L - 1) + (size_t)(row - 1)) * 2 + 1) * (size_t)(a.n_total / 2) + (size_t)(ucg + h * 32 + cidx * 8)) : 0;
                            if (dwrite) *reinterpret_cast<uint4*>(a.y_hi + doff) = hi;
                            if (planes > 1) {
                                float fh[8];
                                unpack8(hi, fh);
                                uint4 lo;
                                lo.x = pack_bf16(v[cidx * 8 + 0] - fh[0], v[cidx * 8 + 1] - fh[1]);
                                lo.y = pack_bf16(v[cidx * 8 + 2] - fh[2], v[cidx * 8 + 3] - fh[3]);
                                lo.z = pack_bf16(v[cidx * 8 + 4] - fh[4], v[cidx * 8 + 5] - fh[5]);
                                lo.w = pack_bf16(v[cidx * 8 + 6] - fh[6], v[cidx * 8 + 7] - fh[7]);
                                sts128(addr[cidx] + a.e_plane_bytes, lo);
                                if (dwrite) *reinterpret_cast<uint4*>(a.y_lo + doff) = lo;
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && a.dbg == 4) {
                        mbar_arrive(bar_e_empty + 8 * se);
                    } else if (lane == 0) {
                        for (int pl = 0; pl < planes; ++pl) {
                            // half h of GEMM rows m..m+31 -> (r, q) = (1, m - 1) for h = 0, (0, m) for h = 1.  TMA stores fault on a
                            // negative start coordinate (measured), so the one box per item that starts at q = -1 was written
                            // with plain stores above (31 rows; GEMM row 0's first half lies before the sequence).
                            if (a.ups && row0 + q * 32 - 1 + uhalf < 0)
                                continue;
                            else if (a.ups)
                                tma_store_4d(pl ? &map_y_lo : &map_y_hi, slot + pl * a.e_plane_bytes + warp_off, ucg, 1 - uhalf, row0 + q * 32 - 1 + uhalf, b);
                            else
                                tma_store_3d(pl ? &map_y_lo : &map_y_hi, slot + pl * a.e_plane_bytes + warp_off, g * a.ecols, (row0 + q * 32) >> rshift, b);
                        }
                        bulk_commit();
                        // hand back the slot whose store was issued `depth` steps ago: its shared-memory reads are done
                        if (depth == 2) {
                            if (hist[1] >= 0) { bulk_wait_read<2>(); mbar_arrive(bar_e_empty + 8 * hist[1]); }
                            hist[1] = hist[0]; hist[0] = se;
                        } else {
                            if (hist[0] >= 0) { bulk_wait_read<1>(); mbar_arrive(bar_e_empty + 8 * hist[0]); }
                            hist[0] = se;
                        }
                    }
                    if (++se == a.n_e) { se = 0; pe ^= 1u; }
                }
            }
            tc_fence_before();   // this warp has read the last of this accumulator buffer
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
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

EncodeTiledFn encode_fn() {
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

bool encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
            uint32_t inner_bytes) {
    EncodeTiledFn fn = encode_fn();
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
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled (umma2) failed (%d) rank %d dims %llu,%llu box %u,%u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        set_error(buf);
        return false;
    }
    return true;
}

int env_i(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}
uint32_t rup(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

struct Umma2Launch::Impl {
    K2Args a;
    alignas(64) CUtensorMap map_a[2], map_w[2], map_r[2], map_m[2], map_y[2];
    int grid;
    size_t smem;
};

bool umma2_supported(const UmmaConvParams& p) {
    const ConvGeom& g = p.g;
    if (env_i("HFG_UMMA_V", 2) < 2) return false;
    if (p.cin_pad != g.Cin) return false;
    if (p.y_raw || p.res || p.xs || !p.y_act) return false;   // planes-only dataflow
    if (p.mrf_hi && !p.res_hi) return false;                   // the MRF sum enters with the residual (last convs2 of a branch)
    if (p.f16 && p.npass != 1) return false;
    if (g.ups_s == 1) {
        if (g.Np != g.Cout || g.Cout > 256 || g.Cout % 32 != 0 || g.Cin != g.Cout) return false;
    } else {   // polyphase upsampler with k = 2s (two taps, pad = s/2) whose whole N = s*C_out fits one tile
        if (env_i("HFG_U2_UPS", 1) == 0) return false;
        if (g.taps != 2 || g.tap_off0 != 0 || g.tap_step != -1 || g.ups_s % 2 != 0 || g.ups_p * 2 != g.ups_s) return false;
        if (g.Cout > 256) return false;   // the bias of a column tile is staged as bias_s[n0 % C_out + c], 256 entries (generators with c0 >= 1024: first-generation kernel)
        if (g.Np != g.ups_s * g.Cout || g.Np % 64 != 0 || p.res_hi) return false;   // halves of <= 64 columns, or
        if (g.Np > 128 && (g.Np % 256 != 0 || env_i("HFG_U2_UPS_WIDE", 1) == 0)) return false;   // wide: 128-column tiles, each inside one half
    }
    return true;
}

int plan_conv_umma2(Umma2Launch* out, const UmmaConvParams& p, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                    const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int sm_count) {
    const ConvGeom& g = p.g;
    out->impl.reset();
    if (!umma2_supported(p)) return HFG_ERR_UNSUPPORTED;
    std::shared_ptr<Umma2Launch::Impl> I(new Umma2Launch::Impl());
    K2Args& a = I->a;
    memset(&a, 0, sizeof a);
    const int planes = p.npass > 1 ? 2 : 1;
    const bool ups = g.ups_s > 1;
    const int N = (ups && g.Np > 128) ? 128 : g.Np;   // C_out, or s * C_out for the polyphase upsampler (wide ones: 128-column tiles)
    const int n_tiles = g.Np / N;
    a.n_tiles = n_tiles; a.n_total = g.Np;
    a.B = g.B; a.L = g.Mrows; a.N = N;   // GEMM rows per item (L_in + 1 for the upsampler: the last row only feeds its first half)
    a.ups = ups ? 1 : 0; a.cout = g.Cout;
    a.taps = g.taps; a.tap_off0 = g.tap_off0; a.tap_step = g.tap_step;
    const int last_off = g.tap_off0 + (g.taps - 1) * g.tap_step;
    a.lo = std::min(g.tap_off0, last_off);
    const int span = std::max(g.tap_off0, last_off) - a.lo;
    a.planes = planes; a.npass = p.npass;
    a.has_res = (p.res_hi != nullptr);
    a.has_mrf = (p.mrf_hi != nullptr) ? 1 : 0;
    a.out_scale = a.has_mrf ? p.out_scale : 1.0f;
    a.paired = (!ups && N == 32 && g.Lin % 2 == 0 && env_i("HFG_U2_PAIRED", 1)) ? 1 : 0;
    a.bias = p.bias;
    a.y_hi = p.y_act; a.y_lo = p.y_act_lo;
    a.dbg = env_i("HFG_U2_DBG", 0);
    a.reverse = p.reverse;
    a.f16 = p.f16 ? 1 : 0;
    if (p.lens) {   // ragged batch: rows (upsampler: output half-rows) per mel frame must be whole
        const int per_T = ups ? 2 * g.Lin : g.Lout;
        if (p.len_T < 1 || per_T % p.len_T != 0) return HFG_ERR_UNSUPPORTED;
        a.lens = p.lens; a.len_mul = per_T / p.len_T; a.len_skip = std::max(p.len_skip, 8);
    }
    // N = 32 always; N = 64 from 7 taps on (measured: k = 11 0.456 -> 0.420 ms, k = 7 0.316 -> 0.277 ms, k = 3 loses); wider layers
    // stream W, where halving MT would double that traffic.  Depends on the layer only: bits never depend on B or L.
    const int concat_maxn = env_i("HFG_U2_CONCAT_MAXN", 0);
    a.concat = (planes == 2 && !ups && (concat_maxn ? N <= concat_maxn : (N <= 32 || (N == 64 && g.taps >= 7)))) ? 1 : 0;
    a.acc_n = a.concat ? 2 * N : N;

    // Choose (K-chunk width, epilogue box width, MT, resident W, ring depths) by a small cost model: cycles per output
    // row = tile interval / rows, where the interval is the largest of the tensor time (measured MMA floors:
    // max(128*N/256, (4096 + 32*N)/128) cycles per K=16 MMA), the HBM time of the tile and the L2 time of the streamed
    // weights, inflated when a ring is too shallow to cover its fetch latency.  Narrow K-chunks / boxes (32 channels)
    // halve the stage sizes, which is what lets the two-plane (bf16x3) mode keep MT >= 2 and real pipelining.
    const int mt_max = std::max(1, std::min({256 / a.acc_n, 4, env_i("HFG_U2_MT", 4), (g.Mrows + 127) / 128}));
    const uint32_t budget = kSmemBudget - 1024;   // alignment slack
    auto floor_clk = [](double n) { return std::max(128.0 * n / 256.0, (4096.0 + 32.0 * n) / 128.0); };
    // tensor cycles per K=16 step of one 128-row subtile, all passes
    const double step_clk = a.concat ? floor_clk(2.0 * N) + floor_clk(N) : p.npass * floor_clk(N);
    const double lat_hbm = 3500.0, lat_l2 = 3000.0, sm_bw = 20.0, l2_bw = 24.0;   // cycles, cycles, bytes/cycle/SM
    const int force_na = env_i("HFG_U2_NA", 0), force_nw = env_i("HFG_U2_NW", 0), force_ne = env_i("HFG_U2_NE", 0);
    const int force_res = env_i("HFG_U2_RESIDENT", -1), force_kc = env_i("HFG_U2_KC", 0), force_ec = env_i("HFG_U2_ECOLS", 0);
    bool ok = false;
    double best = 1e30;
    // The K-chunk width fixes the order in which partial products enter the fp32 accumulator, so it depends on the mode and
    // layer only -- never on B or L: a batch shard must reproduce the bits of the whole batch (tests: batch independence).
    const int kc_fixed = force_kc ? std::min(force_kc, p.kc) : (planes == 2 ? 32 : p.kc);
    for (int kc = kc_fixed; kc == kc_fixed; kc = 0) {
      const uint32_t row_bytes = (uint32_t)kc * 2u;
      const int nchunks = p.cin_pad / kc, ksteps = kc / 16;
      const uint32_t w_plane = rup((uint32_t)N * row_bytes, 1024);
      if (a.concat && w_plane != (uint32_t)N * row_bytes) continue;   // [W_hi ; W_lo] must be contiguous rows
      const uint32_t w_tile = w_plane * planes;
      const uint32_t w_all = (uint32_t)(nchunks * a.taps) * w_tile;
      for (int ecols = std::min(64, N); ecols >= 32; ecols -= 32) {
        if (force_ec && ecols != force_ec && N > 32) continue;
        if (ups && ecols != N / 2) continue;   // the two column groups are the two output half-rows
        const int groups = N / ecols;
        const uint32_t e_plane = 128u * (uint32_t)ecols * 2u;
        const uint32_t e_slot = e_plane * planes * (a.has_mrf ? 2u : 1u);
        for (int mt = mt_max; mt >= 1; --mt) {
            const int rows_need = mt * 128 + span;
            const int pieces = (rows_need + 255) / 256;
            const int box_rows = ((rows_need + pieces - 1) / pieces + 7) / 8 * 8;
            const uint32_t a_plane = rup((uint32_t)(pieces * box_rows) * row_bytes, 1024);
            const uint32_t a_stage = a_plane * planes;
            const int boxes = mt * groups;
            const double t_tile = (double)mt * nchunks * a.taps * ksteps * step_clk;
            const double bytes = (double)pieces * box_rows * nchunks * row_bytes * planes + (a.has_res ? (a.has_mrf ? 3.0 : 2.0) : 1.0) * mt * 128.0 * N * 2.0 * planes;
            // epilogue: ~(250 + 120 per residual plane) issue cycles per 32-column step and warp, 4 warps in parallel
            const double t_epi = (double)mt * (N / 32) * (260.0 + (a.has_res ? (a.has_mrf ? 220.0 : 110.0) : 0.0) * planes + (planes > 1 ? 120.0 : 0.0)) + boxes * 150.0;
            for (int resident = 1; resident >= 0; --resident) {
                if (resident && (w_all > 140u * 1024u || n_tiles > 1)) continue;   // every column tile has its own weights
                if (force_res >= 0 && resident != force_res) continue;
                const double w_bytes_tile = resident ? 0.0 : (double)w_all;
                // streamed weights also cost one barrier round trip per (chunk, tap) in every issuer
                const double t_int = std::max({t_tile, bytes / sm_bw, w_bytes_tile / l2_bw, t_epi}) + 1200.0 + (resident ? 0.0 : nchunks * a.taps * 80.0);
                for (int n_w = resident ? 1 : kMaxW; n_w >= (resident ? 1 : 2); --n_w) {
                    if (!resident && force_nw && n_w != force_nw) continue;
                    const uint32_t w_bytes = resident ? w_all : (uint32_t)n_w * w_tile;
                    for (int n_e = kMaxE; n_e >= 2; --n_e) {
                        if (force_ne && n_e != force_ne) continue;
                        if (planes == 1 && n_e == 2) continue;   // two epilogue groups alternate steps: a slot must outlive one own step
                        const uint32_t fixed = w_bytes + (uint32_t)n_e * e_slot;
                        if (fixed + 2 * a_stage > budget) continue;
                        const int n_a_max = (int)std::min<uint32_t>((budget - fixed) / a_stage, (uint32_t)kMaxA);
                        for (int n_a = n_a_max; n_a >= 2; --n_a) {
                            if (force_na && n_a != force_na) continue;
                            const double f_a = std::min(1.0, (n_a - 1) * (t_int / nchunks) / lat_hbm);
                            const double f_w = resident ? 1.0 : std::min(1.0, (n_w - 1) * (t_int / (nchunks * a.taps)) / lat_l2);
                            const int pend = planes == 1 ? 2 * std::min(2, (n_e - 1) / 2) : std::min(2, n_e - 1);   // slots held by stores in flight
                            const double f_e = a.has_res ? std::max(0.6, std::min(1.0, std::max(0.5, (double)(n_e - pend - 1)) * (t_int / boxes) / lat_hbm))
                                                         : (n_e - pend >= 1 ? 1.0 : 0.5);
                            // small problems: a partly filled last wave of the persistent grid idles SMs (favours smaller tiles)
                            const long tiles = (long)((g.Mrows + mt * 128 - 1) / (mt * 128)) * g.B * n_tiles;
                            const long waves = (tiles + sm_count - 1) / std::max(1, sm_count);
                            const double fill = (double)tiles / (double)(waves * std::max(1, sm_count));
                            const double cost = t_int / std::min({f_a, f_w, f_e}) / (mt * 128.0) / fill;
                            if (cost < best - 1e-9) {
                                best = cost;
                                a.kc = kc; a.nchunks = nchunks; a.ecols = ecols; a.groups = groups;
                                a.e_plane_bytes = e_plane; a.w_plane_bytes = w_plane;
                                a.mt = mt; a.rows_a = pieces * box_rows; a.a_box_rows = box_rows; a.a_pieces = pieces;
                                a.a_plane_bytes = a_plane; a.n_a = n_a; a.n_w = resident ? 1 : n_w; a.n_e = n_e; a.w_resident = resident;
                                a.off_w = (uint32_t)n_a * a_stage;
                                a.off_e = a.off_w + w_bytes;
                                // > half an SM, so exactly one CTA (512 TMEM columns) lives on an SM
                                I->smem = std::max<size_t>((size_t)a.off_e + (size_t)n_e * e_slot + 1024, 120u * 1024u);
                                ok = true;
                            }
                        }
                    }
                }
            }
        }
      }
    }
    if (env_i("HFG_U2_VERBOSE", 0) && ok)
        fprintf(stderr, "umma2 plan N=%d taps=%d planes=%d res=%d: kc=%d ecols=%d mt=%d resident=%d n_a=%d n_w=%d n_e=%d smem=%zu cost=%.2f\n",
                N, a.taps, planes, a.has_res, a.kc, a.ecols, a.mt, a.w_resident, a.n_a, a.n_w, a.n_e, I->smem, best);
    const uint32_t row_bytes = (uint32_t)a.kc * 2u;
    if (!ok) return HFG_ERR_UNSUPPORTED;
    a.tiles_per_item = (g.Mrows + a.mt * 128 - 1) / (a.mt * 128);
    a.total_tiles = a.tiles_per_item * g.B * n_tiles;
    I->grid = std::min(a.total_tiles, std::max(1, sm_count));

    const uint64_t dims3[3] = {(uint64_t)g.Cin, (uint64_t)g.Lin, (uint64_t)g.B};
    const uint64_t str3[2] = {(uint64_t)g.Cin * 2, (uint64_t)g.Lin * g.Cin * 2};
    {
        const uint32_t box[3] = {(uint32_t)a.kc, (uint32_t)a.a_box_rows, 1};
        if (!encode(&I->map_a[0], x_hi, 3, dims3, str3, box, row_bytes)) return HFG_ERR_CUDA;
        if (!encode(&I->map_a[1], planes > 1 ? x_lo : x_hi, 3, dims3, str3, box, row_bytes)) return HFG_ERR_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p.cin_pad, (uint64_t)g.taps * g.Np};
        const uint64_t str[1] = {(uint64_t)p.cin_pad * 2};
        const uint32_t box[2] = {(uint32_t)a.kc, (uint32_t)N};
        if (!encode(&I->map_w[0], w_hi, 2, dims, str, box, row_bytes)) return HFG_ERR_CUDA;
        if (!encode(&I->map_w[1], planes > 1 ? w_lo : w_hi, 2, dims, str, box, row_bytes)) return HFG_ERR_CUDA;
    }
    {
        // residual boxes: 128 time rows; output boxes: one epilogue warp's 32 rows.  Paired (C = 32): [B][L/2][64] view.
        const int pr = a.paired ? 2 : 1;
        const uint64_t edims[3] = {(uint64_t)g.Cin * pr, (uint64_t)g.Lin / pr, (uint64_t)g.B};
        const uint64_t estr[2] = {(uint64_t)g.Cin * pr * 2, (uint64_t)g.Lin * g.Cin * 2};
        const uint32_t box[3] = {(uint32_t)a.ecols * pr, (uint32_t)(128 / pr), 1};
        const uint32_t ybox[3] = {(uint32_t)a.ecols * pr, (uint32_t)(32 / pr), 1};
        const uint32_t eb = (uint32_t)a.ecols * pr * 2u;
        const void* r0 = a.has_res ? (const void*)p.res_hi : (const void*)p.y_act;
        const void* r1 = (a.has_res && planes > 1) ? (const void*)p.res_lo : r0;
        if (!encode(&I->map_r[0], r0, 3, edims, estr, box, eb)) return HFG_ERR_CUDA;
        if (!encode(&I->map_r[1], r1, 3, edims, estr, box, eb)) return HFG_ERR_CUDA;
        const void* m0p = a.has_mrf ? (const void*)p.mrf_hi : r0;
        const void* m1p = (a.has_mrf && planes > 1) ? (const void*)p.mrf_lo : m0p;
        if (!encode(&I->map_m[0], m0p, 3, edims, estr, box, eb)) return HFG_ERR_CUDA;
        if (!encode(&I->map_m[1], m1p, 3, edims, estr, box, eb)) return HFG_ERR_CUDA;
        if (ups) {   // output [B][L_in * s][C_out] viewed as [B][L_in][2][N/2]: (half-row parity r, q = half-row / 2)
            const uint64_t NT = (uint64_t)g.Np;
            const uint64_t ud[4] = {NT / 2, 2, (uint64_t)g.Lin, (uint64_t)g.B};
            const uint64_t us[3] = {NT, NT * 2, (uint64_t)g.Lin * NT * 2};
            const uint32_t ub[4] = {(uint32_t)a.ecols, 1, 32, 1};
            if (!encode(&I->map_y[0], p.y_act, 4, ud, us, ub, eb)) return HFG_ERR_CUDA;
            if (!encode(&I->map_y[1], planes > 1 ? p.y_act_lo : p.y_act, 4, ud, us, ub, eb)) return HFG_ERR_CUDA;
        } else {
            if (!encode(&I->map_y[0], p.y_act, 3, edims, estr, ybox, eb)) return HFG_ERR_CUDA;
            if (!encode(&I->map_y[1], planes > 1 ? p.y_act_lo : p.y_act, 3, edims, estr, ybox, eb)) return HFG_ERR_CUDA;
        }
    }
    out->impl = I;
    out->mt = a.mt; out->n_a = a.n_a; out->n_w = a.n_w; out->n_e = a.n_e; out->w_resident = a.w_resident;
    out->smem = I->smem; out->grid = I->grid;
    return HFG_OK;
}

cudaError_t launch_conv_umma2(const Umma2Launch& L, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev % 64]) {
        cudaError_t e = cudaSuccess;
#define HFG_U2_ATTR(P, R, F)                                                                                                                           \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_umma2_kernel<P, R, F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_umma2_kernel<P, R, F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)
        HFG_U2_ATTR(1, false, false); HFG_U2_ATTR(1, true, false); HFG_U2_ATTR(2, false, false); HFG_U2_ATTR(2, true, false);
        HFG_U2_ATTR(1, false, true); HFG_U2_ATTR(1, true, true);
#undef HFG_U2_ATTR
        if (e != cudaSuccess) return e;
        configured[dev % 64] = true;
    }
    const Umma2Launch::Impl& I = *L.impl;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(I.grid); cfg.blockDim = dim3(threads_for(I.a.planes)); cfg.dynamicSmemBytes = I.smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const int use_pdl = env_i("HFG_PDL", 1);
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
#define HFG_U2_LAUNCH1(P, R, F, G)                                                                                                      \
    e = cudaLaunchKernelEx(&cfg, conv_umma2_kernel<P, R, F, G>, I.map_a[0], I.map_a[1], I.map_w[0], I.map_w[1], I.map_r[0], I.map_r[1], \
                           I.map_m[0], I.map_m[1], I.map_y[0], I.map_y[1], I.a)
#define HFG_U2_LAUNCH(P, R, F) \
    do { if (I.a.lens) HFG_U2_LAUNCH1(P, R, F, true); else HFG_U2_LAUNCH1(P, R, F, false); } while (0)
    cudaError_t e;
    if (I.a.planes == 2) { if (I.a.has_res) HFG_U2_LAUNCH(2, true, false); else HFG_U2_LAUNCH(2, false, false); }
    else if (I.a.f16) { if (I.a.has_res) HFG_U2_LAUNCH(1, true, true); else HFG_U2_LAUNCH(1, false, true); }
    else { if (I.a.has_res) HFG_U2_LAUNCH(1, true, false); else HFG_U2_LAUNCH(1, false, false); }
    if (e != cudaSuccess) return e;
#undef HFG_U2_LAUNCH
#undef HFG_U2_LAUNCH1
    return cudaGetLastError();
}

}  // namespace hfg
