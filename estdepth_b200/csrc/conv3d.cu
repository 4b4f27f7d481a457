// K2: 3x3x3 convolution (stride 1, pad 1) over vol4 volumes with fused per-channel affine (folded BN or bias),
// activation, up to two residual adds, output scaling, channel-concatenated inputs / split outputs and
// deterministic GroupNorm partial sums.  Exact-fp32 SIMT implementation.
//
// Reference seams: every convbn*_3d / nn.Conv3d(kernel 3) on the hot path -- networks/layers_op.py:16-39,
// hybrid_models/model_hybrid.py:59-60,94-95, hybrid_models/hybrid_depth_decoder.py:84-112,190-200,256,377,
// transformer/epipolar_transformer.py:21,26 -- plus the BatchNorm3d / ReLU / tanh / cat / add passes around them.
//
// Design (B200): persistent grid, one CTA per SM.  The whole packed weight tensor [27][Cin][Cout] stays resident
// in shared memory (110 KB for 32->32); input arrives one 4-channel chunk at a time as a TMA box
// (TD+2)x(TH+2)x(TW+2) x float4 with hardware zero fill for the padding halo, double buffered behind mbarriers.
// A warp owns 8 output channels x (2 planes x 4 rows x 32 columns): lanes run along W so every shared-memory
// activation read is a conflict-free 512-byte row and every weight read is a warp broadcast; each thread keeps
// 64 fp32 accumulators and issues ~21 FFMA per shared-memory load.  Stores are 16-byte, 512 B contiguous per warp.
#include <cuda.h>
#include "common.cuh"
#include "conv3d_common.cuh"

namespace estd {

// ---------------------------------------------------------------- kernel
struct ConvParams {
    const float* weight; const float* scale; const float* shift;
    const float* res0; const float* res1;
    float* out0; float* out1;
    double* gn_partials;
    int in0_chunks, out0_chunks, out_chunks;
    int D, H, W;
    int tiles_h, tiles_w, n_tiles;
    int act_split, act_lo, act_hi;
    float post_scale;
};

constexpr int TD = 2;       // output planes per tile
constexpr int TW = 32;      // output columns per tile (= lanes)

template <int CIN_CHUNKS, int COUT_PAD, int RG>
struct ConvCfg {
    static constexpr int NG = COUT_PAD / 8;                 // warps along output channels
    static constexpr int WARPS = NG * RG;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int TH = 4 * RG;                       // output rows per tile
    static constexpr int CIN = CIN_CHUNKS * 4;
    static constexpr int W_FLOATS = 27 * CIN * COUT_PAD;
    static constexpr int STAGE_F4 = (TD + 2) * (TH + 2) * (TW + 2);
    static constexpr int STAGE_BYTES = STAGE_F4 * 16;
    static constexpr size_t SMEM = (size_t)W_FLOATS * 4 + 2 * (size_t)STAGE_BYTES + 64;
};

template <int CIN_CHUNKS, int COUT_PAD, int RG>
__global__ void __launch_bounds__(ConvCfg<CIN_CHUNKS, COUT_PAD, RG>::THREADS, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const ConvParams p) {
    using Cfg = ConvCfg<CIN_CHUNKS, COUT_PAD, RG>;
    constexpr int TH = Cfg::TH;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_w = reinterpret_cast<float*>(smem_raw);
    unsigned char* s_stage = smem_raw + (size_t)Cfg::W_FLOATS * 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_stage + 2 * (size_t)Cfg::STAGE_BYTES);
    __shared__ double s_red[Cfg::WARPS][2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = warp % Cfg::NG;            // output-channel group: channels [8g, 8g+8)
    const int rg = warp / Cfg::NG;           // row group: rows [4rg, 4rg+4) of the tile

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    {   // resident weights
        const float4* src = reinterpret_cast<const float4*>(p.weight);
        float4* dst = reinterpret_cast<float4*>(s_w);
        for (int i = tid; i < Cfg::W_FLOATS / 4; i += Cfg::THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int n_mine = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_iters = n_mine * CIN_CHUNKS;

    auto tile_origin = [&](int k, int& d0, int& h0, int& w0) {
        const int t = blockIdx.x + k * gridDim.x;
        const int tw = t % p.tiles_w;
        const int th = (t / p.tiles_w) % p.tiles_h;
        const int td = t / (p.tiles_w * p.tiles_h);
        d0 = td * TD; h0 = th * TH; w0 = tw * TW;
    };
    auto issue = [&](int it) {
        int d0, h0, w0;
        tile_origin(it / CIN_CHUNKS, d0, h0, w0);
        const int j = it % CIN_CHUNKS;
        const int s = it & 1;
        void* dst = s_stage + (size_t)s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        if (j < p.in0_chunks) tma_load_4d(dst, &map0, &full[s], 4 * (w0 - 1), h0 - 1, d0 - 1, j);
        else                  tma_load_4d(dst, &map1, &full[s], 4 * (w0 - 1), h0 - 1, d0 - 1, j - p.in0_chunks);
    };

    float acc[TD][4][8];
#pragma unroll
    for (int a = 0; a < TD; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[a][b][c] = 0.0f;
    double gsum = 0.0, gsq = 0.0;

    if (tid == 0 && total_iters > 0) issue(0);

    for (int it = 0; it < total_iters; ++it) {
        if (tid == 0 && it + 1 < total_iters) issue(it + 1);
        mbar_wait(&full[it & 1], (uint32_t)((it >> 1) & 1));
        const int j = it % CIN_CHUNKS;
        const float4* xs = reinterpret_cast<const float4*>(s_stage + (size_t)(it & 1) * Cfg::STAGE_BYTES);
        const float* wj = s_w + j * 4 * COUT_PAD + g * 8;

#pragma unroll 1
        for (int dd = 0; dd < 3; ++dd) {
#pragma unroll 1
            for (int dw = 0; dw < 3; ++dw) {
                float4 xr[TD][6];
#pragma unroll
                for (int od = 0; od < TD; ++od)
#pragma unroll
                    for (int r = 0; r < 6; ++r)
                        xr[od][r] = xs[((od + dd) * (TH + 2) + rg * 4 + r) * (TW + 2) + lane + dw];
#pragma unroll
                for (int dh = 0; dh < 3; ++dh) {
                    const float* wt = wj + ((dd * 3 + dh) * 3 + dw) * (Cfg::CIN * COUT_PAD);
                    float4 wa[4], wb[4];
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        wa[ci] = *reinterpret_cast<const float4*>(wt + ci * COUT_PAD);
                        wb[ci] = *reinterpret_cast<const float4*>(wt + ci * COUT_PAD + 4);
                    }
#pragma unroll
                    for (int od = 0; od < TD; ++od)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float4 x = xr[od][r + dh];
                            float* a = acc[od][r];
                            a[0] = fmaf(x.x, wa[0].x, a[0]); a[1] = fmaf(x.x, wa[0].y, a[1]);
                            a[2] = fmaf(x.x, wa[0].z, a[2]); a[3] = fmaf(x.x, wa[0].w, a[3]);
                            a[4] = fmaf(x.x, wb[0].x, a[4]); a[5] = fmaf(x.x, wb[0].y, a[5]);
                            a[6] = fmaf(x.x, wb[0].z, a[6]); a[7] = fmaf(x.x, wb[0].w, a[7]);
                            a[0] = fmaf(x.y, wa[1].x, a[0]); a[1] = fmaf(x.y, wa[1].y, a[1]);
                            a[2] = fmaf(x.y, wa[1].z, a[2]); a[3] = fmaf(x.y, wa[1].w, a[3]);
                            a[4] = fmaf(x.y, wb[1].x, a[4]); a[5] = fmaf(x.y, wb[1].y, a[5]);
                            a[6] = fmaf(x.y, wb[1].z, a[6]); a[7] = fmaf(x.y, wb[1].w, a[7]);
                            a[0] = fmaf(x.z, wa[2].x, a[0]); a[1] = fmaf(x.z, wa[2].y, a[1]);
                            a[2] = fmaf(x.z, wa[2].z, a[2]); a[3] = fmaf(x.z, wa[2].w, a[3]);
                            a[4] = fmaf(x.z, wb[2].x, a[4]); a[5] = fmaf(x.z, wb[2].y, a[5]);
                            a[6] = fmaf(x.z, wb[2].z, a[6]); a[7] = fmaf(x.z, wb[2].w, a[7]);
                            a[0] = fmaf(x.w, wa[3].x, a[0]); a[1] = fmaf(x.w, wa[3].y, a[1]);
                            a[2] = fmaf(x.w, wa[3].z, a[2]); a[3] = fmaf(x.w, wa[3].w, a[3]);
                            a[4] = fmaf(x.w, wb[3].x, a[4]); a[5] = fmaf(x.w, wb[3].y, a[5]);
                            a[6] = fmaf(x.w, wb[3].z, a[6]); a[7] = fmaf(x.w, wb[3].w, a[7]);
                        }
                }
            }
        }
        __syncthreads();     // everyone is done reading this stage before it is refilled (issue(it+2) next iteration)

        if (j == CIN_CHUNKS - 1) {
            // ---------------- epilogue for this tile ----------------
            int d0, h0, w0;
            tile_origin(it / CIN_CHUNKS, d0, h0, w0);
            const int c0 = g * 8;
            const int act = (c0 < p.act_split) ? p.act_lo : p.act_hi;
            float sc[8], sh[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = __ldg(p.scale + c0 + k); sh[k] = __ldg(p.shift + c0 + k); }
            const size_t vox = (size_t)p.D * p.H * p.W;
            const int w = w0 + lane;
            float tsum = 0.0f, tsq = 0.0f;
#pragma unroll
            for (int od = 0; od < TD; ++od) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int d = d0 + od, h = h0 + rg * 4 + r;
                    const bool ok = (d < p.D) && (h < p.H) && (w < p.W);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int ch = 2 * g + half;
                        float v[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            v[k] = apply_act(fmaf(acc[od][r][half * 4 + k], sc[half * 4 + k], sh[half * 4 + k]), act);
                            acc[od][r][half * 4 + k] = 0.0f;
                        }
                        if (ok && ch < p.out_chunks) {
                            const size_t off = ((size_t)ch * vox + ((size_t)d * p.H + h) * p.W + w) * 4;
                            if (p.res0) { const float4 q = ldg4(p.res0 + off); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
                            if (p.res1) { const float4 q = ldg4(p.res1 + off); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                v[k] *= p.post_scale;
                                tsum += v[k];
                                tsq = fmaf(v[k], v[k], tsq);
                            }
                            float* dst = (ch < p.out0_chunks) ? p.out0 + off
                                                              : p.out1 + (off - (size_t)p.out0_chunks * vox * 4);
                            st4(dst, make_float4(v[0], v[1], v[2], v[3]));
                        }
                    }
                }
            }
            gsum += (double)tsum;
            gsq += (double)tsq;
        }
    }

    if (p.gn_partials) {
        // deterministic: fixed tile->CTA assignment, fixed shuffle tree, fixed warp order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
            gsq += __shfl_xor_sync(0xffffffffu, gsq, o);
        }
        if (lane == 0) { s_red[warp][0] = gsum; s_red[warp][1] = gsq; }
        __syncthreads();
        if (tid == 0) {
            double acc2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int wdx = 0; wdx < Cfg::WARPS; ++wdx) {
                const int grp = (((wdx % Cfg::NG) * 8) < p.act_split) ? 0 : 1;
                acc2[grp][0] += s_red[wdx][0];
                acc2[grp][1] += s_red[wdx][1];
            }
            double* dst = p.gn_partials + (size_t)blockIdx.x * 4;
            dst[0] = acc2[0][0]; dst[1] = acc2[0][1]; dst[2] = acc2[1][0]; dst[3] = acc2[1][1];
        }
    }
}

// ---------------------------------------------------------------- host side
static int make_vol4_map(CUtensorMap* map, const float* base, int chunks, int D, int H, int W, int box_h, int box_d) {
    return make_vol4_tensor_map(map, base, chunks, D, H, W, (TW + 2) * 4, box_h, box_d, 1);
}

template <int CIN_CHUNKS, int COUT_PAD, int RG>
static int tiles_for(const estd_conv3d_desc* d) {
    using Cfg = ConvCfg<CIN_CHUNKS, COUT_PAD, RG>;
    return ((d->D + TD - 1) / TD) * ((d->H + Cfg::TH - 1) / Cfg::TH) * ((d->W + TW - 1) / TW);
}

template <int CIN_CHUNKS, int COUT_PAD, int RG>
static int launch(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    using Cfg = ConvCfg<CIN_CHUNKS, COUT_PAD, RG>;
    const int n_tiles = tiles_for<CIN_CHUNKS, COUT_PAD, RG>(d);
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    *n_ctas = grid;
    if (count_only) return ESTD_OK;
    CUtensorMap map0, map1;
    int rc = make_vol4_map(&map0, d->in0, d->in0_chunks, d->D, d->H, d->W, Cfg::TH + 2, TD + 2);
    if (rc) return rc;
    if (d->in1_chunks > 0) rc = make_vol4_map(&map1, d->in1, d->in1_chunks, d->D, d->H, d->W, Cfg::TH + 2, TD + 2);
    else map1 = map0;
    if (rc) return rc;
    ConvParams p;
    p.weight = d->weight; p.scale = d->scale; p.shift = d->shift;
    p.res0 = d->res0; p.res1 = d->res1;
    p.out0 = d->out0; p.out1 = d->out1;
    p.gn_partials = d->gn_partials;
    p.in0_chunks = d->in0_chunks; p.out0_chunks = d->out0_chunks; p.out_chunks = d->out0_chunks + d->out1_chunks;
    p.D = d->D; p.H = d->H; p.W = d->W;
    p.tiles_h = (d->H + Cfg::TH - 1) / Cfg::TH; p.tiles_w = (d->W + TW - 1) / TW; p.n_tiles = n_tiles;
    p.act_split = d->act_split; p.act_lo = d->act_lo; p.act_hi = d->act_hi;
    p.post_scale = d->post_scale;
    auto kern = conv3d_kernel<CIN_CHUNKS, COUT_PAD, RG>;
    static DeviceOnce attr_set;                  // the opt-in shared-memory limit is a PER-DEVICE function attribute
    if (!attr_set.done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "estd_conv3d: cannot reserve %zu B of shared memory: %s", Cfg::SMEM, cudaGetErrorString(e));
        attr_set.set();
    }
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(map0, map1, p);
    return check_launch("estd_conv3d");
}

static int dispatch(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    ESTD_REQUIRE(d, "estd_conv3d: null descriptor");
    const int cin_chunks = d->in0_chunks + d->in1_chunks;
    ESTD_REQUIRE(d->D > 0 && d->H > 0 && d->W > 0, "estd_conv3d: bad volume %dx%dx%d", d->D, d->H, d->W);
    if (!count_only) {
        ESTD_REQUIRE(d->in0 && (d->weight || d->weight_tc) && d->scale && d->shift && (d->out0 || d->head_out), "estd_conv3d: null pointer");
        const bool ring = !d->planar && (d->precision == ESTD_PREC_3XF16_RING2 || d->precision == ESTD_PREC_3XF16_RING2D || d->precision == ESTD_PREC_3XF16_RING);
        ESTD_REQUIRE(!(d->in0_split || d->in1_split || d->res_split || d->out_split) || ring || d->planar,
                     "estd_conv3d: pre-split (vol4s) tensors are implemented for the plane-ring and planar kernels only");
        ESTD_REQUIRE(!d->out_split || (!d->out1 && (d->out0_chunks % 2) == 0 && d->out0), "estd_conv3d: a pre-split output is one tensor with an even number of chunks");
        ESTD_REQUIRE((!d->in0_split || (d->in0_chunks % 2) == 0) && (!d->in1_split || (d->in1_chunks % 2) == 0), "estd_conv3d: pre-split inputs hold an even number of chunks");
        ESTD_REQUIRE(!d->out_up2 || d->planar, "estd_conv3d: out_up2 is implemented for planar convolutions only");
        ESTD_REQUIRE(!d->head_out || (ring && d->cout_pad == 16 && d->head_w && d->head_b && !d->out_split && !d->out1),
                     "estd_conv3d: the fused logit head needs a plane-ring kernel with cout_pad 16, head_w and head_b");
        ESTD_REQUIRE(d->in0_chunks > 0 && d->in1_chunks >= 0 && (d->in1_chunks == 0 || d->in1), "estd_conv3d: bad input segments");
        ESTD_REQUIRE(d->out0_chunks > 0 && d->out1_chunks >= 0 && (d->out1_chunks == 0 || d->out1), "estd_conv3d: bad output segments");
        ESTD_REQUIRE((d->out0_chunks + d->out1_chunks) * 4 <= (d->cout_pad + 15) / 16 * 16, "estd_conv3d: outputs exceed cout_pad");
        ESTD_REQUIRE(d->act_split >= 0 && (d->act_split % 8) == 0, "estd_conv3d: act_split must be a multiple of 8");
        ESTD_REQUIRE(d->planar || (d->act_lo < ESTD_ACT_ADD_RELU && d->act_hi < ESTD_ACT_ADD_RELU), "estd_conv3d: ESTD_ACT_ADD_RELU / ESTD_ACT_SIGMOID are implemented for planar convolutions only");
        ESTD_REQUIRE(aligned16(d->in0) && (!d->out0 || aligned16(d->out0)) && (!d->weight || aligned16(d->weight)) && (!d->in1 || aligned16(d->in1)) &&
                     (!d->out1 || aligned16(d->out1)) && (!d->res0 || aligned16(d->res0)) && (!d->res1 || aligned16(d->res1)),
                     "estd_conv3d: tensors must be 16-byte aligned");
        ESTD_REQUIRE((d->W * 16) % 16 == 0 && d->W * 4 <= (1 << 30), "estd_conv3d: W too large");
    }
    if (d->planar) return dispatch_planar(d, stream, count_only, n_ctas);
    if (d->precision == ESTD_PREC_3XF16_RING2 || d->precision == ESTD_PREC_3XF16_RING2D) return dispatch_ring2(d, stream, count_only, n_ctas);
    if (d->precision == ESTD_PREC_3XF16_RING) return dispatch_ring(d, stream, count_only, n_ctas);
    if (d->precision == ESTD_PREC_3XTF32 || d->precision == ESTD_PREC_3XF16) return dispatch_tc(d, stream, count_only, n_ctas);
    ESTD_REQUIRE(d->precision == ESTD_PREC_FP32, "estd_conv3d: unknown precision %d", d->precision);
    ESTD_REQUIRE(count_only || d->weight, "estd_conv3d: precision=fp32 needs `weight`");
    if (cin_chunks == 8 && d->cout_pad == 32) return launch<8, 32, 2>(d, stream, count_only, n_ctas);
    if (cin_chunks == 9 && d->cout_pad == 40) return launch<9, 40, 2>(d, stream, count_only, n_ctas);
    if (cin_chunks == 9 && d->cout_pad == 32) return launch<9, 32, 2>(d, stream, count_only, n_ctas);
    if (cin_chunks == 4 && d->cout_pad == 16) return launch<4, 16, 4>(d, stream, count_only, n_ctas);
    if (cin_chunks == 8 && d->cout_pad == 16) return launch<8, 16, 4>(d, stream, count_only, n_ctas);
    return fail(ESTD_EUNSUPPORTED, "estd_conv3d: no kernel for %d input chunks -> cout_pad %d", cin_chunks, d->cout_pad);
}

}  // namespace estd

extern "C" int estd_conv3d_num_ctas(const estd_conv3d_desc* desc) {
    int n = 0;
    int rc = estd::dispatch(desc, nullptr, true, &n);
    return rc ? rc : n;
}

extern "C" int estd_conv3d(const estd_conv3d_desc* desc, void* stream) {
    int n = 0;
    return estd::dispatch(desc, (cudaStream_t)stream, false, &n);
}
