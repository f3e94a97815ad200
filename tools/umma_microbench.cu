// Microbenchmark (development tool): cycles per tcgen05.mma (kind::f16, cta_group::1, M=128, K=16) as a function of N,
// issued back to back by one elected thread from fixed shared-memory descriptors (operand values are irrelevant).
// Answers: is the SS-mode MMA rate bound by tensor math (128*N/256 cycles) or by the shared-memory operand fetch
// (A: 128 rows x 32 B = 4 KB per MMA, B: N x 32 B)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I iris_tts_b200/csrc tools/umma_microbench.cu -o umma_mb
#include <cstdio>
#include <cstdlib>

#include "umma_ptx.cuh"

using namespace hfg::ptx;

// mode 0: same A/B every MMA.  mode 1: A descriptor advances by `a_step16` (x16 bytes) each MMA, wrapping (conv-tap like).
__global__ void __launch_bounds__(256, 1) mb_kernel(int N, int iters, int M, int a_step16, int row_bytes, int spin, int fence, long long* out_cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t dhi = desc_hi((uint32_t)row_bytes);
        const uint32_t sub_step = (128u * (uint32_t)row_bytes) >> 4;
        const uint32_t a0 = desc_lo(base), b0 = desc_lo(base + 96 * 1024);
        long long t0 = clock64();
        uint32_t a_lo = a0;
        for (int i = 0; i < iters; ++i) {
            if (fence) tc_fence_after();
            if (leader) {
                if (row_bytes == 128) umma_ksteps<4>(tmem + (uint32_t)((i & 1) * N), a_lo, b0, dhi, idesc, 1u);
                else {   // the conv kernel's C=32 pattern: 64-byte rows, 2 K-steps per subtile, 2 subtiles per block
                    umma_ksteps<2>(tmem + (uint32_t)((i & 1) * 2 * N), a_lo, b0, dhi, idesc, 1u);
                    umma_ksteps<2>(tmem + (uint32_t)((i & 1) * 2 * N + N), a_lo + sub_step, b0, dhi, idesc, 1u);
                }
            }
            a_lo += (uint32_t)a_step16;
            if (a_lo > a0 + 2048) a_lo = a0;
        }
        if (leader) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (leader && blockIdx.x == 0) *out_cycles = t1 - t0;
        if (leader) mbar_arrive(smem_u32(&bar2));
    } else if (warp >= 4 && spin) {
        mbar_wait(smem_u32(&bar2), 0);   // epilogue-like warps polling a barrier while the MMA warp issues
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;   // x4 MMAs
    printf("M   N   a_step  cycles/MMA  math_floor(128*N/256)  bytes/MMA(A+B)  B/clk\n");
    for (int M : {128, 64}) {
        for (int N : {32, 64, 128, 256}) {
            for (int cfg = 0; cfg < 4; ++cfg) {
                const int step = 8;
                const int rb = 128;
                const int spin = cfg & 1, fence = cfg >> 1;
                if (M == 64) continue;
                mb_kernel<<<148, 256, 200 * 1024>>>(N, iters, M, step, rb, spin, fence, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                long long c;
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                const double per = (double)c / (iters * 4.0);
                const int bytes = M * 32 + N * 32;
                printf("%3d %3d spin%d fence%d %6d rb%3d  %10.1f  %8d  %14d  %6.1f\n", M, N, spin, fence, step, rb, per, 128 * N / 256, bytes, bytes / per);
            }
        }
    }
    return 0;
}
