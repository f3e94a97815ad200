// Microbenchmark (development tool): what would a CTA PAIR buy the narrow layers?  Cycles per tcgen05.mma (kind::f16, K = 16,
// SS mode) issued back to back, per SM, for
//   cta_group::1, M = 128, N      (what conv_umma2 / conv_pair issue today: the SM reads A = 4 KB and B = 32 N bytes per MMA), and
//   cta_group::2, M = 256, N      (a CTA pair: each SM reads its own 128 rows of A and HALF of B, 16 N bytes; the other half of B
//                                  arrives from the peer SM).
// Both do 128 x N x 16 multiply-adds per SM per MMA, so cycles/MMA compare one to one.  Operand values are irrelevant (shared
// memory is left uninitialised); descriptors advance by 128 bytes per MMA like a conv's tap loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I iris_tts_b200/csrc tools/umma_microbench2.cu -o tools/umma_mb2
#include <cstdio>
#include <cstdlib>

#include "umma_ptx.cuh"

using namespace hfg::ptx;

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
        : "memory");
}
// collector::a: the first MMA keeps its A tile in the tensor core's collector buffer, the second (same A descriptor) reuses it
template <int kUse>
__device__ __forceinline__ void umma_coll(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    if (kUse == 0)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
            : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

// pair = 0: every CTA on its own (cta_group::1, M = 128).  pair = 1: clusters of two (cta_group::2, M = 256), the leader issues.
// pair = 2: cta_group::1 where every A tile is used by TWO consecutive MMAs (different B: the W_hi / W_lo passes of bf16x3), the
// second one taking A from the collector buffer; cycles are per MMA (two per A tile).
template <int kPair, int KS>
__global__ void __launch_bounds__(256, 1) mb2_kernel(int N, int iters, int row_bytes, long long* out_cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t rank = kPair == 1 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (kPair == 1) cluster_sync_all();
    if (warp == 0) {
        if (kPair == 1) { tmem_alloc2(smem_u32(&tmem_slot), 512); tmem_relinquish2(); }
        else { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (kPair == 1) cluster_sync_all();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const bool leader = elect_one();
        const int M = kPair == 1 ? 256 : 128;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t dhi = desc_hi((uint32_t)row_bytes);
        const uint32_t a0 = desc_lo(base), b0 = desc_lo(base + 96 * 1024);
        long long t0 = clock64();
        if (rank == 0) {
            uint32_t a_lo = a0;
            for (int i = 0; i < iters; ++i) {
                if (leader) {
                    const uint32_t dt = tmem + (uint32_t)((i & 1) * N);
#pragma unroll
                    for (int k = 0; k < KS; ++k) {             // K = 16 slices of one operand row, unrolled like the kernels' issue loops
                        if (kPair == 1) umma2_lh(dt, a_lo + 2u * k, b0 + 2u * k, dhi, idesc, 1u);
                        else if (kPair == 2) {
                            umma_coll<0>(dt, a_lo + 2u * k, b0 + 2u * k, dhi, idesc, 1u);
                            umma_coll<1>(dt, a_lo + 2u * k, b0 + 2048u + 2u * k, dhi, idesc, 1u);   // second B plane 32 KB further
                        } else umma_bf16_lh(dt, a_lo + 2u * k, b0 + 2u * k, dhi, idesc, 1u);
                    }
                }
                a_lo += 8u;                        // + 128 bytes: the next tap's row-shifted view
                if (a_lo > a0 + 2048) a_lo = a0;
            }
            if (leader) {
                if (kPair == 1) umma_commit2(smem_u32(&bar), 3);
                else umma_commit(smem_u32(&bar));
            }
            __syncwarp();
        }
        mbar_wait(smem_u32(&bar), 0);              // the peer CTA's barrier is arrived on by the leader's multicast commit
        long long t1 = clock64();
        if (leader && blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (kPair == 1) cluster_sync_all();
    if (warp == 0) {
        if (kPair == 1) tmem_dealloc2(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
}

}  // namespace

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(mb2_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mb2_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mb2_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mb2_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mb2_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mb2_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    printf("row_bytes  N    cta_group::1 M=128   cta_group::2 M=256   ::1, A reused by 2 MMAs   math floor (128*N/256)   smem bytes per SM and MMA: 1 / 2\n");
    for (int rb : {128, 64}) {
        for (int N : {32, 64, 128, 256}) {
            double per[3] = {0, 0, 0};
            for (int pair = 0; pair < 3; ++pair) {
                if (pair == 2 && N > 128) continue;   // the second B plane would leave the 200 KB
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = pair == 1 ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cudaError_t e;
                if (rb == 128) e = pair == 2 ? cudaLaunchKernelEx(&cfg, mb2_kernel<2, 4>, N, iters, rb, d) : pair ? cudaLaunchKernelEx(&cfg, mb2_kernel<1, 4>, N, iters, rb, d) : cudaLaunchKernelEx(&cfg, mb2_kernel<0, 4>, N, iters, rb, d);
                else e = pair == 2 ? cudaLaunchKernelEx(&cfg, mb2_kernel<2, 2>, N, iters, rb, d) : pair ? cudaLaunchKernelEx(&cfg, mb2_kernel<1, 2>, N, iters, rb, d) : cudaLaunchKernelEx(&cfg, mb2_kernel<0, 2>, N, iters, rb, d);
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error (pair=%d N=%d): %s\n", pair, N, cudaGetErrorString(e)); return 1; }
                long long c;
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                per[pair] = (double)c / ((double)iters * (rb / 32) * (pair == 2 ? 2 : 1));
            }
            printf("%6d   %4d   %10.1f clk        %10.1f clk        %10.1f clk        %8.1f                 %d / %d\n", rb, N, per[0], per[1], per[2], 128.0 * N / 256.0,
                   128 * 32 + N * 32, 128 * 32 + N * 16);
        }
    }
    return 0;
}
