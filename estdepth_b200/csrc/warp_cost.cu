// K1: fused plane-sweep homography warp + bilinear sample + folded pre0  ->  x0 (cost-volume input).
//
// Reference seam: utils/homo_utils.py:458-504 (homo_warping) + hybrid_models/model_hybrid.py:76,90-94
// (ref_volume repeat, cat, pre0 = 1x1x1 conv 64->32 + BN).  Because bilinear sampling with zero padding
// is linear and maps 0 -> 0, pre0 folds into two 32x32 matvecs at 2-D resolution (estd_premix):
//     x0[c,d,h,w] = (s.W_ref) ref[:,h,w] + b   +   warp_d( (s.W_src) src )[c,h,w]
// so ref_volume, the warped volume and their concat (157+157+315 MB at 480x640/D=64) never exist; the kernel
// reads two L2-resident 2.5 MB maps and streams x0 (157 MB) out once: it is HBM-store bound.
//
// Data layout: maps are map4 [C/4][H][W][4], x0 is vol4 [C/4][D][H][W][4].  One thread owns one target pixel for a
// few consecutive planes and all channel chunks: the homography is evaluated once per (pixel, plane), every global
// access is a 16-byte vector, and for a fixed chunk the 32 lanes of a warp (32 consecutive w) store 512 contiguous bytes.
#include "common.cuh"

namespace estd {

constexpr int kPremixMaxC = 64;

// One warp = 32 consecutive pixels x ONE output chunk (4 channels) of one map; a block = 32 pixels x all chunks, persistent
// over the flat (map, pixel tile) list so that the weights are staged in shared memory once per block.  CIN > 0 (compile-time
// input width): the block stages the tile's CIN x 32 inputs in shared memory -- a few loads per thread, all in flight at once,
// the next tile's issued before this tile's FMAs -- and reads the weights as broadcast 16-byte vectors.
// History: a thread per pixel computing every output was latency bound (75 blocks, 1024 dependent FMA+LDS each: 28 us for
// 2.5 MB); a block per tile paid the weight prologue 20 times per SM (176 us for 5 maps); per-thread register loads of the 32
// inputs were scheduled 3 at a time by ptxas whatever the source order (78 us).
// The accumulation order (ci ascending, bias last) is part of the parity contract and does not depend on the tiling.
template <int CIN>
__global__ void __launch_bounds__(512) premix_kernel(const float* __restrict__ fea, const float* __restrict__ weight,
                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                     int cin_rt, int cout, int HW, int n_maps) {
    extern __shared__ __align__(16) float s_w[];        // [cout][cin] then [cout] bias
    constexpr int XS = CIN > 0 ? CIN : 1;
    constexpr int PER = (XS * 32 + 127) / 128;          // staged elements per thread at the smallest block (128 threads)
    __shared__ float s_x[2][XS][32];
    const int cin = CIN > 0 ? CIN : cin_rt;
    float* s_b = s_w + cout * cin;
    for (int i = threadIdx.x; i < cout * cin; i += blockDim.x) s_w[i] = weight[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int c4 = 4 * (threadIdx.x >> 5);              // blockDim.x == 32 * cout / 4: every warp owns a chunk
    const int tiles = (HW + 31) >> 5;
    const int units = tiles * n_maps;
    const float b0 = s_b[c4], b1 = s_b[c4 + 1], b2 = s_b[c4 + 2], b3 = s_b[c4 + 3];

    if constexpr (CIN > 0) {
        float r[PER];
        auto fetch = [&](int u) {                        // this thread's share of tile u's inputs -> registers
            const int map = u / tiles, p0 = (u - map * tiles) * 32;
            const float* f = fea + (size_t)map * CIN * HW;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int e = (int)threadIdx.x + q * (int)blockDim.x;
                const int ci = e >> 5, p = p0 + (e & 31);
                r[q] = (e < CIN * 32 && p < HW) ? __ldg(f + (size_t)ci * HW + p) : 0.0f;
            }
        };
        auto park = [&](int buf) {
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int e = (int)threadIdx.x + q * (int)blockDim.x;
                if (e < CIN * 32) s_x[buf][e >> 5][e & 31] = r[q];
            }
        };
        int buf = 0;
        if ((int)blockIdx.x < units) { fetch(blockIdx.x); park(0); }
        __syncthreads();
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int un = u + (int)gridDim.x;
            if (un < units) fetch(un);                   // in flight during this tile's FMAs
            const int map = u / tiles;
            const int p = (u - map * tiles) * 32 + lane;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4* wr = reinterpret_cast<const float4*>(s_w + (c4 + k) * CIN);
#pragma unroll
                for (int ci = 0; ci < CIN; ci += 4) {
                    const float4 w = wr[ci >> 2];
                    acc[k] = fmaf(w.x, s_x[buf][ci][lane], acc[k]);
                    acc[k] = fmaf(w.y, s_x[buf][ci + 1][lane], acc[k]);
                    acc[k] = fmaf(w.z, s_x[buf][ci + 2][lane], acc[k]);
                    acc[k] = fmaf(w.w, s_x[buf][ci + 3][lane], acc[k]);
                }
            }
            if (p < HW)
                st4(out + (size_t)map * cout * HW + ((size_t)(c4 >> 2) * HW + p) * 4,
                    make_float4(acc[0] + b0, acc[1] + b1, acc[2] + b2, acc[3] + b3));
            if (un < units) park(buf ^ 1);               // the other buffer: nobody reads it before the barrier below
            __syncthreads();
            buf ^= 1;
        }
    } else {
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int map = u / tiles;
            const int p = (u - map * tiles) * 32 + lane;
            if (p >= HW) continue;
            const float* f = fea + (size_t)map * cin * HW + p;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            float x[kPremixMaxC];
#pragma unroll
            for (int ci = 0; ci < kPremixMaxC; ++ci)
                if (ci < cin) x[ci] = __ldg(f + (size_t)ci * HW);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* wr = s_w + (c4 + k) * cin;
#pragma unroll
                for (int ci = 0; ci < kPremixMaxC; ++ci)
                    if (ci < cin) acc[k] = fmaf(wr[ci], x[ci], acc[k]);
            }
            st4(out + (size_t)map * cout * HW + ((size_t)(c4 >> 2) * HW + p) * 4,
                make_float4(acc[0] + b0, acc[1] + b1, acc[2] + b2, acc[3] + b3));
        }
    }
}

// Tap set of one (pixel, plane) sample: clamped float4 offsets into a source chunk + bilinear weights (0 when the tap
// is out of bounds, so zero padding costs nothing).
struct Taps2 {
    int o_nw, o_ne, o_sw, o_se;
    float w_nw, w_ne, w_sw, w_se;
};

// --- homography (utils/homo_utils.py:469-491) + ATen GridSampler tap arithmetic, fp32, reference operation order ---
template <int ALIGN>
__device__ __forceinline__ Taps2 plane_sweep_taps(const float (&hm)[12], float fx, float fy, float depth, int H, int W) {
    const float qx = fmaf(hm[2], 1.0f, fmaf(hm[1], fy, hm[0] * fx));
    const float qy = fmaf(hm[5], 1.0f, fmaf(hm[4], fy, hm[3] * fx));
    const float qz = fmaf(hm[8], 1.0f, fmaf(hm[7], fy, hm[6] * fx));
    const float px3 = __fadd_rn(__fmul_rn(qx, depth), hm[9]);
    const float py3 = __fadd_rn(__fmul_rn(qy, depth), hm[10]);
    const float pz3 = __fadd_rn(__fmul_rn(qz, depth), hm[11]);
    const float zden = __fadd_rn(pz3, 1e-8f);
    const float px = __fdiv_rn(px3, zden);
    const float py = __fdiv_rn(py3, zden);
    float xn = __fadd_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), -1.0f);
    float yn = __fadd_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), -1.0f);
    xn = force_outside(xn);
    yn = force_outside(yn);
    const float ix = unnormalize(xn, W, ALIGN);
    const float iy = unnormalize(yn, H, ALIGN);
    const float x0f = floorf(ix), y0f = floorf(iy);
    // ATen GridSampler bilinear weights: nw=(x1-ix)(y1-iy), ne=(ix-x0)(y1-iy), sw=(x1-ix)(iy-y0), se=(ix-x0)(iy-y0)
    const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
    const float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
    // NaN / huge coordinates: the float->int conversion saturates and the bounds test rejects the tap
    const int xi = (int)x0f, yi = (int)y0f;
    const bool fin = (ix == ix) && (iy == iy);
    const bool vx0 = fin && (xi >= 0) && (xi < W), vx1 = fin && (xi + 1 >= 0) && (xi + 1 < W);
    const bool vy0 = (yi >= 0) && (yi < H), vy1 = (yi + 1 >= 0) && (yi + 1 < H);
    Taps2 t;
    t.w_nw = (vx0 && vy0) ? wx0 * wy0 : 0.0f;
    t.w_ne = (vx1 && vy0) ? wx1 * wy0 : 0.0f;
    t.w_sw = (vx0 && vy1) ? wx0 * wy1 : 0.0f;
    t.w_se = (vx1 && vy1) ? wx1 * wy1 : 0.0f;
    const int cx0 = min(max(xi, 0), W - 1), cx1 = min(max(xi + 1, 0), W - 1);
    const int cy0 = min(max(yi, 0), H - 1), cy1 = min(max(yi + 1, 0), H - 1);
    t.o_nw = (cy0 * W + cx0) * 4; t.o_ne = (cy0 * W + cx1) * 4;
    t.o_sw = (cy1 * W + cx0) * 4; t.o_se = (cy1 * W + cx1) * 4;
    return t;
}

// Block = 8 rows x 32 columns of target pixels x KP consecutive depth planes per thread.  The target-side value is
// loaded once per chunk and reused for all KP planes, vertically adjacent warps share source rows in L1, and each thread
// keeps 4*KP+1 independent 16-byte loads in flight per chunk.  KP is chosen per launch so that the number of (equal)
// blocks fills whole waves of the 4-blocks-per-SM residency (480x640/D=64: KP=3 -> 1650 blocks = 2.8 waves; KP=4 gave
// 1200 blocks = 2.03 waves, i.e. a third of the time was an almost empty tail).
// History (profiles/README.md): v1 one voxel per thread, one plane per block -> L2-bandwidth bound (L1 hit 42 %);
// a persistent one-plane-at-a-time variant was slower (fewer loads in flight); the kernel is L1-pipe bound (ncu:
// l1tex 68 % of peak, 6 x 16 B through the L1 pipe per 16 B stored).
// (A variant that wrote x0 pre-split -- vol4s, so that pre1 could skip its in-place hi/lo split -- was measured and dropped:
// the results of an even chunk have to wait in registers for the odd chunk that completes their 8-channel group, the kernel
// spills at its 64-register budget and runs at 69 us instead of 41 us, against 5 us saved in the consumer.)
template <int ALIGN, int KP>
__global__ void __launch_bounds__(256, 4) warp_cost_kernel(const float* __restrict__ ref_mix, const float* __restrict__ src_mix,
                                                           const float* __restrict__ homo12,
                                                           const float* __restrict__ depth_values,
                                                           float* __restrict__ x0, int chunks, int D, int H, int W) {
    const int w = blockIdx.x * 32 + (threadIdx.x & 31);
    const int h = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int d0 = blockIdx.z * KP;
    if (w >= W || h >= H) return;
    const int HW = H * W;
    const int p = h * W + w;
    float hm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) hm[i] = __ldg(homo12 + i);
    Taps2 taps[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int d = min(d0 + k, D - 1);
        taps[k] = plane_sweep_taps<ALIGN>(hm, (float)w, (float)h, __ldg(depth_values + d), H, W);
    }
    const size_t plane = (size_t)HW * 4;
    const size_t out_chunk = (size_t)D * HW * 4;
    const float* refp = ref_mix + (size_t)p * 4;
    float* outp = x0 + ((size_t)d0 * HW + p) * 4;
#pragma unroll 1
    for (int j = 0; j < chunks; ++j) {
        const float* s = src_mix + j * plane;
        const float4 r = ldg4(refp + j * plane);
        float4 a[KP], b[KP], c[KP], e[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            a[k] = ldg4(s + taps[k].o_nw); b[k] = ldg4(s + taps[k].o_ne);
            c[k] = ldg4(s + taps[k].o_sw); e[k] = ldg4(s + taps[k].o_se);
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (d0 + k >= D) break;
            const Taps2& t = taps[k];
            float4 o;
            o.x = r.x + fmaf(e[k].x, t.w_se, fmaf(c[k].x, t.w_sw, fmaf(b[k].x, t.w_ne, a[k].x * t.w_nw)));
            o.y = r.y + fmaf(e[k].y, t.w_se, fmaf(c[k].y, t.w_sw, fmaf(b[k].y, t.w_ne, a[k].y * t.w_nw)));
            o.z = r.z + fmaf(e[k].z, t.w_se, fmaf(c[k].z, t.w_sw, fmaf(b[k].z, t.w_ne, a[k].z * t.w_nw)));
            o.w = r.w + fmaf(e[k].w, t.w_se, fmaf(c[k].w, t.w_sw, fmaf(b[k].w, t.w_ne, a[k].w * t.w_nw)));
            __stcs(reinterpret_cast<float4*>(outp + j * out_chunk + (size_t)k * plane), o);   // streamed: never re-read by this kernel
        }
    }
}

template <int KP>
static void launch_warp_cost(int align, dim3 grid, cudaStream_t st, const float* ref, const float* src, const float* homo,
                             const float* dv, float* x0, int chunks, int D, int H, int W) {
    if (align) warp_cost_kernel<1, KP><<<grid, 256, 0, st>>>(ref, src, homo, dv, x0, chunks, D, H, W);
    else       warp_cost_kernel<0, KP><<<grid, 256, 0, st>>>(ref, src, homo, dv, x0, chunks, D, H, W);
}

}  // namespace estd

extern "C" int estd_premix_batch(const float* fea_nchw, const float* weight, const float* bias, float* out_map4,
                                 int n_maps, int cin, int cout, int H, int W, void* stream) {
    ESTD_REQUIRE(fea_nchw && weight && out_map4, "estd_premix: null pointer");
    ESTD_REQUIRE(cin > 0 && cin <= estd::kPremixMaxC && cout > 0 && cout <= 64 && (cout % 4) == 0,
                 "estd_premix: cin=%d cout=%d unsupported (cin<=64, cout<=64, cout%%4==0)", cin, cout);
    ESTD_REQUIRE(n_maps > 0 && n_maps <= 65535 && H > 0 && W > 0 && estd::aligned16(out_map4), "estd_premix: bad shape/alignment");
    const int HW = H * W;
    const size_t smem = (size_t)(cout * cin + cout) * sizeof(float);
    const long long units = (long long)((HW + 31) / 32) * n_maps;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(units < 2ll * sms ? units : 2ll * sms);
    const int threads = 32 * (cout / 4);
    if (cin == 32 && threads >= 128)             // the staged kernel spreads a tile's 32 x 32 inputs over >= 128 threads
        estd::premix_kernel<32><<<grid, threads, smem, (cudaStream_t)stream>>>(fea_nchw, weight, bias, out_map4, cin, cout, HW, n_maps);
    else
        estd::premix_kernel<0><<<grid, threads, smem, (cudaStream_t)stream>>>(fea_nchw, weight, bias, out_map4, cin, cout, HW, n_maps);
    return estd::check_launch("estd_premix");
}

extern "C" int estd_premix(const float* fea_chw, const float* weight, const float* bias, float* out_map4,
                           int cin, int cout, int H, int W, void* stream) {
    return estd_premix_batch(fea_chw, weight, bias, out_map4, 1, cin, cout, H, W, stream);
}

extern "C" int estd_warp_cost(const float* ref_mix_map4, const float* src_mix_map4, const float* homo12,
                              const float* depth_values, float* x0_vol4, int C, int D, int H, int W,
                              int align_corners, void* stream) {
    ESTD_REQUIRE(ref_mix_map4 && src_mix_map4 && homo12 && depth_values && x0_vol4, "estd_warp_cost: null pointer");
    ESTD_REQUIRE(C > 0 && (C % 4) == 0 && D > 0 && H > 1 && W > 1 && (long long)D * H * W < (1ll << 30),
                 "estd_warp_cost: unsupported shape C=%d D=%d H=%d W=%d", C, D, H, W);
    ESTD_REQUIRE(estd::aligned16(ref_mix_map4) && estd::aligned16(src_mix_map4) && estd::aligned16(x0_vol4),
                 "estd_warp_cost: tensors must be 16-byte aligned");
    const int tiles_w = (W + 31) / 32, tiles_h = (H + 7) / 8;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // planes per thread: the value in 2..5 whose block count fills whole waves of 4 resident blocks per SM best
    int best_kp = 4;
    double best_eff = -1.0;
    for (int kp = 5; kp >= 2; --kp) {
        const long long blocks = (long long)tiles_w * tiles_h * ((D + kp - 1) / kp);
        const long long slots = (long long)sms * 4;
        const double eff = (double)D / (double)(((D + kp - 1) / kp) * kp) * (double)blocks / (double)(((blocks + slots - 1) / slots) * slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_kp = kp; }
    }
    const dim3 grid(tiles_w, tiles_h, (D + best_kp - 1) / best_kp);
    ESTD_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "estd_warp_cost: volume too large");
    const cudaStream_t st = (cudaStream_t)stream;
    switch (best_kp) {
        case 2: estd::launch_warp_cost<2>(align_corners, grid, st, ref_mix_map4, src_mix_map4, homo12, depth_values, x0_vol4, C / 4, D, H, W); break;
        case 3: estd::launch_warp_cost<3>(align_corners, grid, st, ref_mix_map4, src_mix_map4, homo12, depth_values, x0_vol4, C / 4, D, H, W); break;
        case 4: estd::launch_warp_cost<4>(align_corners, grid, st, ref_mix_map4, src_mix_map4, homo12, depth_values, x0_vol4, C / 4, D, H, W); break;
        default: estd::launch_warp_cost<5>(align_corners, grid, st, ref_mix_map4, src_mix_map4, homo12, depth_values, x0_vol4, C / 4, D, H, W); break;
    }
    return estd::check_launch("estd_warp_cost");
}
