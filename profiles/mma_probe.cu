// Micro-probe: how many cycles does one tcgen05.mma (kind::f16, M=128, K=16, cta_group::1, SS operands, no-swizzle K-major
// descriptors like conv3d_ring.cu) take as a function of N, of how many accumulators the stream alternates between, and of
// whether the shared-memory pipe is busy with other traffic?  One CTA per SM, one issuing thread; cycles from clock64.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/mma_probe profiles/mma_probe.cu && build/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../estdepth_b200/csrc/tc_ptx.cuh"

using namespace estd;
using namespace estd::tc;

// mode bit 0: other warps hammer shared memory with 16-byte loads/stores meanwhile
template <int NACC>
__global__ void __launch_bounds__(256, 1) probe(int N, int layout, int n_mma, int a_step16, int mode, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 0) {
        const bool leader = elect_one();
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
        const uint32_t idesc = make_idesc(0u, N);
        long long t0 = 0, t1 = 0;
        if (leader) {
            // layout 0: the convolution's halo-tile view (8-row groups one image row apart, K groups 2 planes apart);
            // layout 1: dense core matrices (K group stride 128 B, 8-row group stride 256 B)
            const uint64_t a_desc = layout == 0 ? make_desc(a0, 2 * 9792, 34 * 16) : make_desc(a0, 128, 256);
            const uint64_t b_desc = layout == 0 ? make_desc(b0, N * 16, 128) : make_desc(b0, 128, 256);
            for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up
                t0 = clock64();
                for (int i = 0; i < n_mma; i += 36) {
#pragma unroll
                    for (int j = 0; j < 36; ++j) {       // constant offsets, like the real issue loop
                        const uint32_t acc = tmem_base + (uint32_t)((j % NACC) * N);
                        umma<KIND_F16>(acc, a_desc + (uint64_t)((j % 9) * a_step16), b_desc + (uint64_t)((j % 3) * 4), idesc, 1u);
                    }
                }
                umma_commit(&bar);
                mbar_wait(&bar, (uint32_t)(rep & 1));
                t1 = clock64();
            }
            stop = 1;
            out[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    } else if (warp >= 2 && (mode & 1)) {
        uint4* area = reinterpret_cast<uint4*>(smem + 40 * 1024);
        uint4 acc = make_uint4(0, 0, 0, 0);
        while (!stop) {
            for (int i = tid; i < 2048; i += 192) { uint4 v = area[i]; acc.x ^= v.x; area[i + 2048] = acc; }
        }
        if (acc.x == 0x12345) out[1000] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 2048 * sizeof(long long));
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int n_mma = 36 * 128;
    printf("%5s %6s %6s %5s %5s | %9s  %s\n", "N", "n_acc", "layout", "step", "mode", "clk/MMA", "floor N/2");
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {32, 64, 96, 128, 192, 256})
            for (int n_acc : {1, 2})
                for (int layout : {0, 1})
                    for (int step : {0, 1}) {
                        if (mode == 1 && (step == 0 || n_acc == 1)) continue;
                        if (n_acc == 1) probe<1><<<148, 256, 200 * 1024>>>(N, layout, n_mma, step, mode, out);
                        else            probe<2><<<148, 256, 200 * 1024>>>(N, layout, n_mma, step, mode, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                        long long mx = 0;
                        for (int i = 0; i < 148; ++i) mx = out[i] > mx ? out[i] : mx;
                        printf("%5d %6d %6d %5d %5d | %9.1f  %d\n", N, n_acc, layout, step, mode, (double)mx / n_mma, N / 2);
                    }
    return 0;
}
