// Microbenchmark (development tool): what HBM bandwidth can a PERSISTENT, TMA-fed kernel of this engine's shape reach at all?
// conv_umma2 / conv_pair stream their activation planes through a ring of TMA boxes ([rows x 64 bf16], SWIZZLE_128B) with one
// CTA per SM, and their HBM-bound layers sit at 4.0-4.9 TB/s against 6.4 TB/s for the plain-load mrf_combine kernel.  This
// program runs the same data movement with the math removed: one producer thread issues the loads of a ring of `stages` boxes,
// one thread stores every box back (copy mode) or just hands the slot back (load-only mode), over 452 MB in (+ 452 MB out).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I iris_tts_b200/csrc tools/tma_copy_microbench.cu -o tools/tma_copy_mb -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "umma_ptx.cuh"

using namespace hfg::ptx;

namespace {

constexpr int kMaxStages = 12;

__global__ void __launch_bounds__(128, 1)
tma_copy_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, int tiles, int box_rows, int stages,
                int store) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)box_rows * 128u;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = bar_full + 8 * kMaxStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        fence_barrier_init();
        prefetch_tmap(&map_in);
        prefetch_tmap(&map_out);
    }
    __syncthreads();
    if (warp == 0 && lane == 0) {
        int s = 0;
        uint32_t p = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            mbar_wait(bar_empty + 8 * s, p ^ 1u);
            mbar_expect_tx(bar_full + 8 * s, stage_bytes);
            tma_load_3d(base + s * stage_bytes, &map_in, bar_full + 8 * s, 0, t * box_rows, 0);
            if (++s == stages) { s = 0; p ^= 1u; }
        }
    } else if (warp == 1 && lane == 0) {
        int s = 0, prev = -1;
        uint32_t p = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            mbar_wait(bar_full + 8 * s, p);
            if (store) {
                tma_store_3d(&map_out, base + s * stage_bytes, 0, t * box_rows, 0);
                bulk_commit();
                if (prev >= 0) { bulk_wait_read<1>(); mbar_arrive(bar_empty + 8 * prev); }   // the store before this one has read its slot
                prev = s;
            } else {
                mbar_arrive(bar_empty + 8 * s);
            }
            if (++s == stages) { s = 0; p ^= 1u; }
        }
        if (store) bulk_wait<0>();
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool encode(CUtensorMap* m, void* ptr, uint64_t rows, uint32_t box_rows) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
    cuuint64_t dims[3] = {64, rows, 1}, str[2] = {128, rows * 128};
    cuuint32_t box[3] = {64, box_rows, 1}, es[3] = {1, 1, 1};
    return reinterpret_cast<EncodeFn>(fp)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int main() {
    const uint64_t rows = 3530752;   // x 128 bytes = 452 MB: a two-plane C = 64 stage tensor of the headline workload
    const size_t bytes = rows * 128;
    void *in = nullptr, *out = nullptr;
    if (cudaMalloc(&in, bytes) != cudaSuccess || cudaMalloc(&out, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(in, 1, bytes);
    cudaMemset(out, 0, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(tma_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("# persistent TMA copy, %d SMs, %.0f MB in; GB/s counts bytes read + bytes written\n", sms, bytes / 1e6);
    printf("# mode box_rows stage_KB stages ctas_per_sm ms GB/s\n");
    for (int store = 1; store >= 0; --store)
        for (int cps = 1; cps <= 2; ++cps)
            for (int box_rows : {128, 256})
                for (int stages : {2, 3, 4, 6, 8, 12}) {
                    const size_t smem = (size_t)stages * box_rows * 128 + 1024;
                    if (smem > (cps == 1 ? 216u : 100u) * 1024u) continue;
                    // one CTA per SM: ask for more than half of the SM's shared memory so that no second CTA fits
                    const size_t ask = cps == 1 ? (smem > 120u * 1024u ? smem : 120u * 1024u) : smem;
                    CUtensorMap mi, mo;
                    if (!encode(&mi, in, rows, box_rows) || !encode(&mo, out, rows, box_rows)) { printf("encode failed\n"); return 1; }
                    const int tiles = (int)(rows / box_rows);
                    const int grid = sms * cps;
                    for (int i = 0; i < 3; ++i) tma_copy_kernel<<<grid, 128, ask>>>(mi, mo, tiles, box_rows, stages, store);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    const int reps = 10;
                    cudaEventRecord(e0);
                    for (int i = 0; i < reps; ++i) tma_copy_kernel<<<grid, 128, ask>>>(mi, mo, tiles, box_rows, stages, store);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    ms /= reps;
                    printf("%s %d %d %d %d %.4f %.0f\n", store ? "copy" : "load", box_rows, box_rows * 128 / 1024, stages, cps, ms,
                           (store ? 2.0 : 1.0) * bytes / (ms * 1e-3) / 1e9);
                }
    // reference point: the runtime's own device-to-device copy of the same buffers
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpyD2D - - - - %.4f %.0f\n", ms / 10, 2.0 * bytes / (ms / 10 * 1e-3) / 1e9);
    unsigned char probe[256];
    cudaMemcpy(probe, (unsigned char*)out + bytes - 256, 256, cudaMemcpyDeviceToHost);
    printf("# last bytes of out: %d %d (expect 1 1)\n", probe[0], probe[255]);
    return 0;
}
