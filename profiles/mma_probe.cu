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

// The exact MMA stream of one conv3d_ring.cu stage (2 halves x 9 taps x 3 products x 2 M tiles, N = 96, accumulators at
// columns (2*half + m2) * 96, A_hi / A_lo halo views, W_hi / W_lo blocks of 9 taps) with nothing else running on the SM.
__global__ void __launch_bounds__(256, 1) probe_ring(int n_stage, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 190 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 0) {
        const bool leader = elect_one();
        constexpr int HALO_W = 34, KG = 612 * 16, A_BYTES = 4 * KG, N3 = 96, W_PART = 2 * N3 * 16, W_TAP = 2 * W_PART, STAGE = 94464;
        const uint32_t idesc = make_idesc(0u, N3);
        long long t0 = 0, t1 = 0;
        if (leader) {
            for (int rep = 0; rep < 2; ++rep) {
                t0 = clock64();
                for (int st = 0; st < n_stage; ++st) {
                    const uint32_t a_hi = smem_u32(smem) + (uint32_t)((st & 1) * STAGE);
                    const uint64_t a_hi_desc = make_desc(a_hi, 2 * KG, HALO_W * 16);
                    const uint64_t a_lo_desc = make_desc(a_hi + KG, 2 * KG, HALO_W * 16);
                    const uint64_t w_hi_desc = make_desc(a_hi + A_BYTES, N3 * 16, 128);
                    const uint64_t w_lo_desc = make_desc(a_hi + A_BYTES + W_PART, N3 * 16, 128);
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t acc0 = tmem_base + (uint32_t)(half * 2 * N3);
                        const uint64_t a_base = (uint64_t)(half * 16);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t b_off = (uint64_t)(tap * (W_TAP >> 4));
#pragma unroll
                            for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                                for (int m2 = 0; m2 < 2; ++m2) {
                                    const uint64_t a_off = a_base + (uint64_t)((tap / 3) * HALO_W + 8 * m2 + (tap % 3));
                                    umma<KIND_F16>(acc0 + (uint32_t)(m2 * N3), (prod == 2 ? a_lo_desc : a_hi_desc) + a_off,
                                                   (prod == 1 ? w_lo_desc : w_hi_desc) + b_off, idesc, 1u);
                                }
                            }
                        }
                    }
                }
                umma_commit(&bar);
                mbar_wait(&bar, (uint32_t)(rep & 1));
                t1 = clock64();
            }
            out[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 2048 * sizeof(long long));
    cudaFuncSetAttribute(probe_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    {
        const int n_stage = 40;
        probe_ring<<<148, 256, 200 * 1024>>>(n_stage, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = out[i] > mx ? out[i] : mx;
        printf("ring stage pattern: %.1f clk/MMA (%.0f clk per 108-MMA stage)\n", (double)mx / (n_stage * 108), (double)mx / n_stage);
    }
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
