// Planar (2-D) 3x3 convolution, stride 1, dilation 1 or 2, on the tcgen05 tensor cores: the 3x3 layers of the 2-D feeder
// networks (matching-feature net networks/psm_submodule.py:14-54, context decoder hybrid_models/hybrid_depth_decoder.py:
// 17-30,163-184) over a stack of feature maps held as vol4 [C/4][N][H][W][4].  Same error-compensated fp16 split as the
// 3-D kernels (x = x_hi + x_lo, w * 2^k = w_hi + w_lo; x_hi w_hi + x_hi w_lo + x_lo w_hi in fp32 TMEM accumulators).
//
//   GEMM     D[M = 128 pixels (16 rows x 8 columns), N] += A[M, K = 16 channels of one tap] * B[N, K]^T
//   unit     16 x 16 pixels of one map (2 M tiles) x one slice of COUT output channels; persistent CTAs, unit = blockIdx + k*grid
//   A        one TMA box per stage: the halo tile of 16 input channels (4 fp32 chunks), split IN PLACE by 8 warps into
//            x_hi / x_lo K-groups; each tap of each M tile is a shifted start address of a no-swizzle K-major descriptor
//   B        per-stage weight block [9 taps][2 K-groups][W_hi rows | W_lo rows][16 B] by bulk copy
//   MMA      per tap and M tile: D[:, 0:2C] (+)= A_hi x [W_hi | W_lo] (N = 2C) and D[:, C:2C] += A_lo x W_hi (N = C): the
//            small products accumulate apart from the large ones (the tensor core truncates to the accumulator's ulp on
//            every accumulate; sharing the large products between the two halves was measured and is WORSE, because every
//            small addend then also costs one truncation at full magnitude).
//            With C = 64 that is 64 + 48 shared-memory wavefronts of operand fetch for 64 + 32 cycles of math
//            (profiles/mma_probe.cu: an M=128, K=16 MMA costs max(N/2, 32 + N/4) cycles).
//   TMEM     2 M tiles x 2C columns per unit, double buffered: the epilogue of unit u (8 warps: TMEM -> hi + lo -> affine ->
//            activation -> residual -> 16-byte stores) overlaps the MMAs of unit u+1.
//   Pre-split inputs (vol4s, the normal case between planar layers) need no splitter: the issuer waits on the TMA barrier
//            itself and the 8 splitter warps become 8 MORE epilogue warps (each warp then owns half of the slice's channels
//            of its 32 pixels).  ncu (profiles/planar_epilogue_r02.txt): with 8 epilogue warps the 1x1 layers were bound by
//            the epilogue's instruction issue (440 instructions per 16-channel block and warp, 2 warps per scheduler).
//   The number of 16-channel k-steps is a run-time argument (64 ... 2048 input channels use the same kernel), and layers
//   wider than COUT run as cout_pad/COUT slices INSIDE one launch (unit = tile x slice, slices of a tile on adjacent CTAs so
//   that the input tile is shared through L2): the small, deep maps of the context decoder (15x20 ... 30x40 pixels, 256
//   output channels) would otherwise occupy a dozen SMs.
// The x_hi w_hi products of an output run through one accumulator: 9 * Cin/16 accumulating MMAs, ~0.5 ulp of truncation
// each (8e-5 at Cin = 1280 for O(1) outputs; cuDNN's own fp32 Winograd kernels are at 5e-5 .. 1e-4 on such layers).
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "conv3d_common.cuh"
#include "tc_ptx.cuh"

namespace estd {
namespace planar {

using namespace tc;

constexpr int EPI_WARPS = 8, SPLIT_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32, SPLIT_THREADS = SPLIT_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS + SPLIT_THREADS;       // warps 0-3 control, 4-11 epilogue, 12-19 splitters
constexpr int FIRST_SPLIT_WARP = 4 + EPI_WARPS;

// COUT output channels per accumulator (slice), DIL dilation of the taps, TAPS 9 (3x3) or 1 (1x1: the pointwise convolutions
// of the ResNet bottlenecks -- same pipeline without a halo; with a single tap per 16 channels the stage is dominated by the
// TMA fill and the hi/lo split rather than by the MMAs, still several times faster than the fp32 CUDA-core GEMM)
template <int COUT_, int DIL_, int TAPS_>
struct Shape {
    static constexpr int COUT = COUT_, DIL = DIL_, TAPS = TAPS_, MT = 2;
    static constexpr int PAD = (TAPS == 9) ? DIL : 0;
    static constexpr int TILE_H = 16, TILE_W = 8 * MT;
    static constexpr int HALO_H = TILE_H + 2 * PAD, HALO_W = TILE_W + 2 * PAD, HALO_VOX = HALO_H * HALO_W;
    static constexpr int MAX_COUT = (TAPS == 9) ? 512 : 2048;     // widest layer (all slices) one launch takes
    static constexpr int KGROUP_BYTES = HALO_VOX * 16;
    static constexpr int A_BYTES = 4 * KGROUP_BYTES;
    static constexpr int N_ALL = 2 * COUT;
    static constexpr int W_TAP_BYTES = 2 * N_ALL * 16;            // [2 K-groups][N_ALL rows][16 B]
    static constexpr int W_BYTES = TAPS * W_TAP_BYTES;
    static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES + 127) / 128 * 128;
    static constexpr int STAGES = (6 * STAGE_BYTES + 20480 <= 227 * 1024) ? 6 : (4 * STAGE_BYTES + 8192 <= 227 * 1024) ? 4 : 3;
    static constexpr int COLS_PER_UNIT = MT * N_ALL;
    static constexpr int TMEM_COLS = (2 * COLS_PER_UNIT <= 128) ? 128 : (2 * COLS_PER_UNIT <= 256) ? 256 : 512;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 256;
    static_assert(2 * COLS_PER_UNIT <= 512, "two accumulator buffers must fit TMEM");
    static_assert(SMEM + 2 * 4 * MAX_COUT + 2048 <= 227 * 1024, "stages must fit shared memory");
    static_assert(TAPS == 9 || TAPS == 1, "3x3 or 1x1");
    static_assert(COUT % 16 == 0 && COUT <= 64, "bad COUT");
};

struct Params {
    const float* weight_tc;                     // [slices][nks][9 taps][2 K-groups][2*COUT rows][16 bytes]
    int n_slices, cout_total;
    int* status;
    ConvEpilogue ep;
    int in0_chunks, nks;
    int in0_split, in1_split;                   // the input segments are pre-split (vol4s): the splitter warps pass them through
    int all_presplit;                           // every input segment is: no splitter pass, the splitter warps join the epilogue
    int D, H, W;                                // D = number of maps in the stack
    int tiles_h, tiles_w, n_units;
};

template <class S>
__global__ void __launch_bounds__(THREADS, 1)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Params p) {
    constexpr int COUT = S::COUT, STAGES = S::STAGES, N_ALL = S::N_ALL;
    constexpr int HALO_W = S::HALO_W, HALO_VOX = S::HALO_VOX, A_BYTES = S::A_BYTES;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * S::STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES] TMA landed
    uint64_t* ready = bars + STAGES;        // [STAGES] split done
    uint64_t* empty = bars + 2 * STAGES;    // [STAGES] MMAs done reading
    uint64_t* acc_full = bars + 3 * STAGES; // [2]
    uint64_t* acc_empty = acc_full + 2;     // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ __align__(16) float s_scale[S::MAX_COUT], s_shift[S::MAX_COUT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nks = p.nks;
    // helpers: the splitter warps work as a second set of epilogue warps (needs two 16-channel blocks per slice)
    const bool helpers = p.all_presplit && COUT >= 32;
    const int epi_threads = helpers ? EPI_THREADS + SPLIT_THREADS : EPI_THREADS;

    pdl_launch_dependents();                 // the next kernel of the stream may start its prologue as SMs free up
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], SPLIT_THREADS); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], epi_threads); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, S::TMEM_COLS);
    // all threads: a 2048-channel layer read by one warp was 64 dependent round trips before the first TMA could be issued
    for (int i = tid; i < p.cout_total; i += THREADS) { s_scale[i] = p.ep.scale[i]; s_shift[i] = p.ep.shift[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();                              // the prologue above overlapped the previous kernel's tail; activations from here on

    const int n_mine = (p.n_units > (int)blockIdx.x) ? (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto unit_origin = [&](int k, int& d, int& h0, int& w0, int& slice) {
        const int uu = blockIdx.x + k * gridDim.x;
        slice = uu % p.n_slices;
        const int u = uu / p.n_slices;
        const int tw = u % p.tiles_w;
        const int th = (u / p.tiles_w) % p.tiles_h;
        d = u / (p.tiles_w * p.tiles_h);
        h0 = th * S::TILE_H; w0 = tw * S::TILE_W;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int k = 0; k < n_mine; ++k) {
                int d, h0, w0, slice;
                unit_origin(k, d, h0, w0, slice);
                const float* wsl = p.weight_tc + (size_t)slice * nks * (S::W_BYTES / 4);
                for (int ks = 0; ks < nks; ++ks, ++it) {
                    const int s = it % STAGES;
                    if (it >= STAGES) mbar_wait_polls(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
                    unsigned char* stage = smem + (size_t)s * S::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(A_BYTES + S::W_BYTES));
                    const int chunk = 4 * ks;
                    if (chunk < p.in0_chunks) tma_load_4d(stage, &map0, &full[s], 4 * (w0 - S::PAD), h0 - S::PAD, d, chunk);
                    else                      tma_load_4d(stage, &map1, &full[s], 4 * (w0 - S::PAD), h0 - S::PAD, d, chunk - p.in0_chunks);
                    bulk_load(stage + A_BYTES, wsl + (size_t)ks * (S::W_BYTES / 4), (uint32_t)S::W_BYTES, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_all = make_idesc(0u, N_ALL), idesc_hi = make_idesc(0u, COUT);
        const bool leader = elect_one();
        int it = 0;
        for (int k = 0; k < n_mine; ++k) {
            const int buf = k & 1, use = k >> 1;                   // how many times this accumulator buffer was used before
            if (use > 0) { mbar_wait_polls(&acc_empty[buf], (uint32_t)((use - 1) & 1)); tc_fence_after(); }
            const uint32_t acc0 = tmem_base + (uint32_t)(buf * S::COLS_PER_UNIT);
            for (int ks = 0; ks < nks; ++ks, ++it) {
                const int s = it % STAGES;
                mbar_wait_polls(p.all_presplit ? &full[s] : &ready[s], (uint32_t)((it / STAGES) & 1));
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)s * S::STAGE_BYTES);
                const uint64_t a_hi_desc = make_desc(a_hi, 2 * S::KGROUP_BYTES, HALO_W * 16);
                const uint64_t a_lo_desc = make_desc(a_hi + S::KGROUP_BYTES, 2 * S::KGROUP_BYTES, HALO_W * 16);
                const uint64_t b_desc = make_desc(a_hi + A_BYTES, N_ALL * 16, 128);
                const uint32_t later = (ks == 0) ? 0u : 1u;        // the very first MMA of a unit overwrites the accumulator
                if (leader) {
#pragma unroll
                    for (int tap = 0; tap < S::TAPS; ++tap) {
#pragma unroll
                        for (int mt = 0; mt < S::MT; ++mt) {
                            const uint64_t a_off = (uint64_t)((tap / 3) * S::DIL * HALO_W + 8 * mt + (tap % 3) * S::DIL);   // TAPS == 1: 8 * mt
                            const uint64_t b_off = (uint64_t)(tap * (S::W_TAP_BYTES >> 4));
                            const uint32_t acc = acc0 + (uint32_t)(mt * N_ALL);
                            umma<KIND_F16>(acc, a_hi_desc + a_off, b_desc + b_off, idesc_all, tap == 0 ? later : 1u);
                            umma<KIND_F16>(acc + COUT, a_lo_desc + a_off, b_desc + b_off, idesc_hi, 1u);
                        }
                    }
                    umma_commit(&empty[s]);
                    if (ks == nks - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= FIRST_SPLIT_WARP && !p.all_presplit) {
        // ===================== hi/lo splitter =====================
        const int t = tid - FIRST_SPLIT_WARP * 32;
        float amax = 0.0f;
        const int n_stages = n_mine * nks;
        for (int it = 0; it < n_stages; ++it) {
            const int s = it % STAGES;
            if (warp == FIRST_SPLIT_WARP) mbar_wait_polls(&full[s], (uint32_t)((it / STAGES) & 1));
            named_barrier(2, SPLIT_THREADS);
            unsigned char* area = smem + (size_t)s * S::STAGE_BYTES;
            const bool presplit = (4 * (it % nks) < p.in0_chunks) ? (p.in0_split != 0) : (p.in1_split != 0);
            for (int i = t; !presplit && i < 2 * HALO_VOX; i += SPLIT_THREADS) {
                const int pair = i / HALO_VOX, v = i - pair * HALO_VOX;
                float4* c0 = reinterpret_cast<float4*>(area + (size_t)pair * 2 * S::KGROUP_BYTES) + v;
                float4* c1 = c0 + HALO_VOX;
                const float4 a = *c0, b = *c1;
                amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                                         fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
                const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
                const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                const __half2 l0 = __floats2half2_rn(a.x - f0.x, a.y - f0.y), l1 = __floats2half2_rn(a.z - f1.x, a.w - f1.y);
                const __half2 l2 = __floats2half2_rn(b.x - f2.x, b.y - f2.y), l3 = __floats2half2_rn(b.z - f3.x, b.w - f3.y);
                uint4 hv, lv;
                hv.x = h2u(h0); hv.y = h2u(h1); hv.z = h2u(h2); hv.w = h2u(h3);
                lv.x = h2u(l0); lv.y = h2u(l1); lv.z = h2u(l2); lv.w = h2u(l3);
                *reinterpret_cast<uint4*>(c0) = hv;
                *reinterpret_cast<uint4*>(c1) = lv;
            }
            fence_proxy_async();
            mbar_arrive(&ready[s]);
        }
        const bool bad = !(amax <= 65504.0f);
        if (bad && p.status) atomicOr(p.status, 1);
    } else if (warp >= 4 && (warp < FIRST_SPLIT_WARP || helpers)) {
        // ===================== epilogue =====================
        const int e = warp - 4, q = e & 3, mt = (e >> 2) & 1;    // TMEM lane quarter (= warp % 4); M tile
        // channels of the slice this warp handles: all of them, or one half when the splitter warps help
        const int c_lo = helpers ? (e >> 3) * (COUT / 2) : 0, c_hi = helpers ? c_lo + COUT / 2 : COUT;
        const int m = q * 32 + lane;
        const int mh = m >> 3, mw = m & 7;
        const size_t vox = (size_t)p.D * p.H * p.W;
        const ConvEpilogue& ep = p.ep;
        const int act = ep.act_hi;                               // planar layers have one activation (checked by the launcher)
        const bool has_res = ep.res0 != nullptr;
        // residual of one 16-channel block of this thread's pixel in unit k: issued as soon as the previous one is consumed -- for
        // the next block of the unit, or for the first block of this CTA's NEXT unit -- so that its latency hides behind the split
        // and the stores of this block and the TMEM loads of the next (issued at the start of each unit it was fully exposed)
        float4 r0[4];
        auto issue_res = [&](int k, int c0) {
            int d, h0, w0, slice;
            unit_origin(k, d, h0, w0, slice);
            const int h = h0 + mh, w = w0 + 8 * mt + mw;
            const bool ok = (h < p.H) && (w < p.W);
            const size_t pos = ((size_t)d * p.H + h) * p.W + w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = ((slice * COUT + c0) >> 2) + j;
                r0[j] = (ok && ch < ep.out_chunks) ? ldg4(ep.res0 + ((size_t)ch * vox + pos) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (has_res && n_mine > 0) issue_res(0, c_lo);
        float amax = 0.0f;
        for (int k = 0; k < n_mine; ++k) {
            const int buf = k & 1, use = k >> 1;
            int d, h0, w0, slice;
            unit_origin(k, d, h0, w0, slice);
            const int cbase = slice * COUT;                        // first output channel of this slice
            const int h = h0 + mh, w = w0 + 8 * mt + mw;
            const bool ok = (h < p.H) && (w < p.W);
            const size_t pos = ((size_t)d * p.H + h) * p.W + w;
            // output addressing: the same map, or its nearest x2 up-sampled version [2H][2W]
            const size_t ovox = ep.out_up2 ? 4 * vox : vox;
            const size_t opos = ep.out_up2 ? ((size_t)d * 2 * p.H + 2 * h) * (2 * p.W) + 2 * w : pos;
            const size_t orow = (size_t)2 * p.W * 4;               // floats per row of the up-sampled map
            const float mult = s_scale[cbase];                     // uniform within a slice: the per-channel multiplier is folded into the weights
            // truncation-bias compensation (common.cuh): the large-product accumulator received one MMA per k-step for every
            // filter tap inside the map (the small products accumulate apart and need none)
            const float comp = kTruncBiasPerMma * (float)(nks * (S::TAPS == 9 ? taps_inside(h, p.H, S::DIL) * taps_inside(w, p.W, S::DIL) : 1));
            if (e == 0) mbar_wait_polls(&acc_full[buf], (uint32_t)(use & 1));
            named_barrier(3, epi_threads);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * S::COLS_PER_UNIT + mt * N_ALL);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                float a[16], b[16];
                tmem_ld16(t0 + (uint32_t)c0, a);
                tmem_ld16(t0 + (uint32_t)(COUT + c0), b);
                // the 16 offsets of this block in one batch (a shared-memory load queues behind the tensor core's operand fetches)
                float4 sh[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) sh[j] = *reinterpret_cast<const float4*>(s_shift + cbase + c0 + 4 * j);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float s4[4] = {sh[j].x, sh[j].y, sh[j].z, sh[j].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[4 * j + i] = fmaf(fmaf(a[4 * j + i], comp, a[4 * j + i]) + b[4 * j + i], mult, s4[i]);
                }
                if (act == ESTD_ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
                } else if (act == ESTD_ACT_TANH) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
                } else if (act == ESTD_ACT_SIGMOID) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = sigmoidf_acc(v[i]);
                }
                if (has_res) {
                    if (ep.res_split) {                            // vol4s residual: chunks (0,1) = hi / lo of 8 channels, (2,3) of the next 8
                        float t[8];
                        join8(r0[0], r0[1], t);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += t[i];
                        join8(r0[2], r0[3], t);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[8 + i] += t[i];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { v[4 * j] += r0[j].x; v[4 * j + 1] += r0[j].y; v[4 * j + 2] += r0[j].z; v[4 * j + 3] += r0[j].w; }
                    }
                    if (c0 + 16 < c_hi) issue_res(k, c0 + 16);
                    else if (k + 1 < n_mine) issue_res(k + 1, c_lo);
                }
                if (act == ESTD_ACT_ADD_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
                }
                if (ep.post_scale != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] *= ep.post_scale;
                }
                const int ch0 = (cbase + c0) >> 2;                 // first of the block's 4 output chunks
                if (ep.out_split) {
                    // vol4s: x_hi of 8 channels -> chunk c/4, x_lo -> the next chunk
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int ch = ch0 + 2 * g;
                        uint4 hi, lo;
                        split8(v + 8 * g, hi, lo, amax);
                        if (!ok || ch >= ep.out_chunks) continue;
                        float* dst = ep.out0 + ((size_t)ch * ovox + opos) * 4;
                        uint4* dh = reinterpret_cast<uint4*>(dst);
                        uint4* dl = reinterpret_cast<uint4*>(dst + ovox * 4);
                        *dh = hi; *dl = lo;
                        if (ep.out_up2) {
                            dh[1] = hi; dl[1] = lo;
                            uint4* dh2 = reinterpret_cast<uint4*>(dst + orow);
                            uint4* dl2 = reinterpret_cast<uint4*>(dst + ovox * 4 + orow);
                            dh2[0] = hi; dh2[1] = hi; dl2[0] = lo; dl2[1] = lo;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ch = ch0 + j;
                        if (!ok || ch >= ep.out_chunks) continue;
                        const float4 o4 = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        if (ep.out_up2) {                            // nearest x2: the pixel's 2 x 2 block of the [2H][2W] map
                            float* dst = ep.out0 + ((size_t)ch * ovox + opos) * 4;
                            st4(dst, o4); st4(dst + 4, o4); st4(dst + orow, o4); st4(dst + orow + 4, o4);
                            continue;
                        }
                        const size_t off = ((size_t)ch * vox + pos) * 4;
                        float* dst = (ch < ep.out0_chunks) ? ep.out0 + off : ep.out1 + (off - (size_t)ep.out0_chunks * vox * 4);
                        st4(dst, o4);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
        // values of pixels outside the map come from zero-filled tiles and channels beyond out_chunks from zero weights: finite
        if (ep.out_split && !(amax <= 65504.0f) && ep.status) atomicOr(ep.status, 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

template <class S>
static int launch(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    const int tiles_h = (d->H + S::TILE_H - 1) / S::TILE_H, tiles_w = (d->W + S::TILE_W - 1) / S::TILE_W;
    const int n_slices = d->cout_pad / S::COUT;
    const long long n_units = (long long)d->D * tiles_h * tiles_w * n_slices;
    ESTD_REQUIRE(n_units < (1ll << 30), "estd_conv3d(planar): too many tiles");
    const int grid = n_units < sm_count() ? (int)n_units : sm_count();
    *n_ctas = grid;
    if (count_only) return ESTD_OK;
    const int cin_chunks = d->in0_chunks + d->in1_chunks;
    ESTD_REQUIRE(d->weight_tc && aligned16(d->weight_tc), "estd_conv3d(planar): needs a 16-byte aligned weight_tc");
    ESTD_REQUIRE(d->in1_chunks == 0 || (d->in0_chunks % 4) == 0, "estd_conv3d(planar): first input segment must hold a multiple of 4 chunks");
    ESTD_REQUIRE(!d->gn_partials && !d->res1, "estd_conv3d(planar): GroupNorm partial sums / second residual are not implemented for planar convolutions");
    ESTD_REQUIRE(!d->out_up2 || !d->out1, "estd_conv3d(planar): an up-sampled output is one tensor");
    ESTD_REQUIRE(d->act_lo == d->act_hi, "estd_conv3d(planar): one activation per layer (act_lo == act_hi)");
    CUtensorMap map0, map1;
    int rc = make_vol4_tensor_map(&map0, d->in0, d->in0_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, 4);
    if (rc) return rc;
    if (d->in1_chunks > 0) rc = make_vol4_tensor_map(&map1, d->in1, d->in1_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, 4);
    else map1 = map0;
    if (rc) return rc;
    Params p;
    p.weight_tc = d->weight_tc;
    p.n_slices = n_slices; p.cout_total = d->cout_pad;
    p.status = d->status;
    fill_epilogue(&p.ep, d);
    p.in0_chunks = d->in0_chunks;
    p.in0_split = d->in0_split; p.in1_split = d->in1_split;
    p.all_presplit = d->in0_split && (d->in1_chunks == 0 || d->in1_split);
    p.nks = (cin_chunks + 3) / 4;
    p.D = d->D; p.H = d->H; p.W = d->W;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w; p.n_units = (int)n_units;
    auto kern = conv2d_tc_kernel<S>;
    static DeviceOnce attr_set;                  // the opt-in shared-memory limit is a PER-DEVICE function attribute
    if (!attr_set.done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "estd_conv3d(planar): cannot reserve %zu B of shared memory: %s", S::SMEM, cudaGetErrorString(e));
        attr_set.set();
    }
    {
        cudaError_t e = launch_pdl(kern, grid, THREADS, S::SMEM, stream, map0, map1, p);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "%s launch: %s", __FILE__, cudaGetErrorString(e));
    }
    return check_launch("estd_conv3d(planar)");
}

}  // namespace planar

int dispatch_planar(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    using namespace planar;
    ESTD_REQUIRE(d->precision == ESTD_PREC_3XF16 || d->precision == ESTD_PREC_3XF16_RING || d->precision == ESTD_PREC_3XF16_RING2D,
                 "estd_conv3d: planar convolutions are implemented for the fp16 split only");
    const int dil = d->dilation > 0 ? d->dilation : 1;
    const int C = d->cout_pad;
    const int taps = d->planar == 2 ? 1 : 9;
    ESTD_REQUIRE(C <= (taps == 9 ? 512 : 2048), "estd_conv3d(planar): cout_pad %d exceeds what one launch takes (%d)", C, taps == 9 ? 512 : 2048);
#define ESTD_PLANAR(COUT, DIL, TAPS) if ((C == COUT || (COUT == 64 && C > 64 && C % 64 == 0)) && dil == DIL && taps == TAPS) \
        return launch<Shape<COUT, DIL, TAPS>>(d, stream, count_only, n_ctas)
    ESTD_PLANAR(64, 1, 9); ESTD_PLANAR(64, 2, 9); ESTD_PLANAR(32, 1, 9); ESTD_PLANAR(32, 2, 9); ESTD_PLANAR(16, 1, 9);
    ESTD_PLANAR(64, 1, 1); ESTD_PLANAR(32, 1, 1);
#undef ESTD_PLANAR
    return fail(ESTD_EUNSUPPORTED, "estd_conv3d(planar): no kernel for cout_pad %d, dilation %d, %d tap(s) (cout_pad 16/32/64 or a multiple of 64, dilation 1/2)", C, dil, taps);
}

}  // namespace estd
