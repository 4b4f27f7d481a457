// K2 on the 5th-generation tensor cores: 3x3x3 convolution as an implicit GEMM with error-compensated TF32
// ("3xTF32": x = x_hi + x_lo, w = w_hi + w_lo, products x_hi w_hi + x_hi w_lo + x_lo w_hi accumulated in fp32 in
// TMEM).  Single-pass TF32 or BF16 misses the 1e-3 depth gate by 1-3 orders of magnitude; the
// split keeps ~21 mantissa bits per product, i.e. fp32-class accuracy, at 3 MMAs per product.
//
// Mapping (one CTA per SM, persistent over work units; unit = 16 (h) x 32 (w) output voxels of one depth plane):
//   GEMM  D[M = voxels, N = Cout] += A[M, K] * B[N, K]^T,  K = 27 taps x Cin, walked as 3 planes x (Cin/8) k-steps
//   A     never materialised (no im2col): one TMA box per stage brings the halo'd input tile of one input plane and
//         8 channels, [2 chunks][18 rows][34 cols][4 floats] with hardware zero fill for the padding.  In the
//         no-swizzle K-major UMMA layout a core matrix is 8 rows x 16 bytes, so 8 consecutive voxels along w (16 B
//         apart) form its rows, the next 8-row group is the next image row (SBO = row pitch) and the second half of
//         K is the next channel chunk (LBO = chunk pitch): every one of the 9 in-plane taps of every M tile (16 rows
//         x 8 columns = 128 voxels) is just a different 16-byte-granular start address into the same tile.
//   B     per-stage weight block [9 taps][2 k-halves][2*Cout rows (w_hi | w_lo)][4 floats], one bulk copy.
//   MMA   per tap and M tile: D[:, 0:2C] (+)= A_hi * [W_hi | W_lo]  (N = 2C, one instruction for two products) and
//         D[:, C:2C] += A_lo * W_hi (N = C); the epilogue adds the two halves.  Accumulators: 4 M tiles x 2C columns
//         of TMEM, double buffered across units so the epilogue of unit u overlaps the MMAs of unit u+1.
//   warps 0: TMA producer, 1: MMA issuer, 2: TMEM allocator, 4-7: epilogue (tcgen05.ld -> affine/activation/residual ->
//         16-byte stores, GroupNorm partial sums), 8-15: hi/lo splitter (turns the landed fp32 tile into its two
//         low-precision terms in place).  All hand-offs are mbarriers; waits are bounded (trap, not hang).
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "conv3d_common.cuh"
#include "tc_ptx.cuh"

namespace estd {

namespace tc {

constexpr int SPLIT_WARPS = 8, SPLIT_THREADS = SPLIT_WARPS * 32;
constexpr int THREADS = 256 + SPLIT_THREADS;          // warps 0-3 control, 4-7 epilogue, 8.. splitters
constexpr uint32_t TF32_MASK = 0xFFFFE000u;

// Two split arithmetics share the kernel (template parameter KIND):
//   KIND_TF32  stage = 8 channels.  TMA lands 2 fp32 chunks in K-groups 0,1; the splitter masks them to TF32 in place
//              (x_hi) and writes x_lo = tf32(x - x_hi) to K-groups 2,3.  kind::tf32, K = 8 per MMA.
//   KIND_F16   stage = 16 channels.  TMA lands 4 fp32 chunks; the splitter turns each pair of chunks (8 channels of
//              a voxel = 32 B of fp32) IN PLACE into one K-group of x_hi = fp16(x) and one of x_lo = fp16(x - x_hi)
//              (8 x fp16 = 16 B each): same bytes, twice the channels per MMA.  kind::f16, K = 16 per MMA.  Weights
//              are pre-scaled by a power of two so that w_lo stays a normal fp16; activations beyond the fp16 range
//              raise the status flag.  x_hi*w_hi, x_hi*w_lo, x_lo*w_hi are exact in the fp32 accumulator either way.

// Compile-time shape of one specialisation.
//   NKS     k-steps per input plane (8 channels each for TF32, 16 for F16)
//   COUT    padded output channels (multiple of 16); N of the MMAs is 2*COUT ([W_hi | W_lo]) and COUT
//   MT      M tiles (16 rows x 8 columns = 128 voxels each) per work unit: unit = 16 x 8*MT voxels of one plane
//   DIL     dilation of the in-plane taps (1 or 2)
//   PLANAR  true: 1x3x3 filter applied to every plane independently (2-D convolution over a stack of feature maps);
//           false: full 3x3x3 filter (3 input planes per output plane)
template <int KIND_, int NKS_, int COUT_, int MT_, int DIL_, bool PLANAR_>
struct Shape {
    static constexpr int KIND = KIND_, NKS = NKS_, COUT = COUT_, MT = MT_, DIL = DIL_;
    static constexpr bool PLANAR = PLANAR_;
    static constexpr int TILE_H = 16, TILE_W = 8 * MT;
    static constexpr int HALO_H = TILE_H + 2 * DIL, HALO_W = TILE_W + 2 * DIL, HALO_VOX = HALO_H * HALO_W;
    static constexpr int KGROUP_BYTES = HALO_VOX * 16;                     // one 16-byte K-group of the halo tile
    static constexpr int A_BYTES = 4 * KGROUP_BYTES;                       // operand area of a stage (x_hi and x_lo)
    static constexpr int CHUNKS = (KIND == KIND_TF32) ? 2 : 4;             // fp32 chunks landed per stage
    static constexpr uint32_t A_LO_OFFSET = (KIND == KIND_TF32) ? 2 * KGROUP_BYTES : KGROUP_BYTES;
    static constexpr uint32_t A_LBO = (KIND == KIND_TF32) ? KGROUP_BYTES : 2 * KGROUP_BYTES;
    static constexpr uint32_t FORMAT = (KIND == KIND_TF32) ? 2 : 0;        // UMMA a/b format: TF32 / F16
    static constexpr int N_ALL = 2 * COUT;
    static constexpr int W_TAP_BYTES = 2 * N_ALL * 16;                     // [2 K-groups][N_ALL rows][16 B]
    static constexpr int W_BYTES = 9 * W_TAP_BYTES;
    static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES + 127) / 128 * 128;
    static constexpr int STAGES = (3 * STAGE_BYTES <= 224 * 1024) ? 3 : 2;
    // Accumulator sets: the tensor core adds into the fp32 accumulator with truncation, so the error of a long K loop
    // grows with the number of accumulating MMAs.  Planar layers (all-positive post-ReLU inputs, up to 180 MMAs per
    // accumulator) alternate k-steps between two sets that the epilogue adds in fp32 round-to-nearest.
    static constexpr int ASETS = PLANAR ? 2 : 1;
    static constexpr int COLS_PER_SET = MT * N_ALL;
    static constexpr int COLS_PER_UNIT = ASETS * COLS_PER_SET;
    static constexpr int NBUF = (2 * COLS_PER_UNIT <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (NBUF * COLS_PER_UNIT <= 32) ? 32 : (NBUF * COLS_PER_UNIT <= 64) ? 64
                                   : (NBUF * COLS_PER_UNIT <= 128) ? 128 : (NBUF * COLS_PER_UNIT <= 256) ? 256 : 512;
    static constexpr int PLANES = PLANAR ? 1 : 3;
    static constexpr int N_STAGES_PER_UNIT = PLANES * NKS;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 256;     // + barriers / tmem base
    static_assert(COLS_PER_UNIT <= 512, "accumulators of one unit must fit TMEM");
    static_assert(2 * STAGE_BYTES + 256 <= 227 * 1024, "two stages must fit shared memory");
};

struct Params {
    const float* weight_tc;                     // [PLANES][NKS][9 taps][2 K-groups][2*COUT rows][16 bytes]
    int* status;                                // optional: set to 1 when an activation leaves the fp16 range (KIND_F16)
    ConvEpilogue ep;
    int in0_chunks;
    int D, H, W;
    int tiles_h, tiles_w, n_units;
};

template <class S>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Params p) {
    using C = S;
    using KT = S;
    constexpr int KIND = S::KIND, NKS = S::NKS, COUT = S::COUT, STAGES = S::STAGES;
    constexpr int N_STAGES_PER_UNIT = S::N_STAGES_PER_UNIT;
    constexpr int HALO_W = S::HALO_W, HALO_VOX = S::HALO_VOX, A_BYTES = S::A_BYTES;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES] TMA landed
    uint64_t* ready = bars + STAGES;        // [STAGES] split done
    uint64_t* empty = bars + 2 * STAGES;    // [STAGES] MMAs done reading
    uint64_t* acc_full = bars + 3 * STAGES; // [2]
    uint64_t* acc_empty = acc_full + 2;     // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ double s_red[4][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], SPLIT_THREADS); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    const int n_mine = (p.n_units > (int)blockIdx.x) ? (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto unit_origin = [&](int k, int& d, int& h0, int& w0) {
        const int u = blockIdx.x + k * gridDim.x;
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
                int d, h0, w0;
                unit_origin(k, d, h0, w0);
                for (int st = 0; st < N_STAGES_PER_UNIT; ++st, ++it) {
                    const int s = it % STAGES;
                    if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
                    const int dd = st / NKS, ks = st % NKS;
                    unsigned char* stage = smem + (size_t)s * C::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(KT::CHUNKS * S::KGROUP_BYTES + C::W_BYTES));
                    const int chunk = KT::CHUNKS * ks;
                    const int z = S::PLANAR ? d : d + dd - 1;
                    if (chunk < p.in0_chunks) tma_load_4d(stage, &map0, &full[s], 4 * (w0 - S::DIL), h0 - S::DIL, z, chunk);
                    else                      tma_load_4d(stage, &map1, &full[s], 4 * (w0 - S::DIL), h0 - S::DIL, z, chunk - p.in0_chunks);
                    bulk_load(stage + A_BYTES, p.weight_tc + (size_t)(dd * NKS + ks) * (C::W_BYTES / 4), (uint32_t)C::W_BYTES, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Warp-uniform control flow; one elected lane issues.  Descriptors are built once per stage and advanced by
        // compile-time constants (fully unrolled taps x M tiles) so that the single issuing thread spends a couple of
        // uniform-datapath instructions per MMA -- with per-MMA descriptor arithmetic the issue rate of that one
        // thread, not the tensor pipe, bounded the kernel (profiles/README.md, r01 -> r02).
        const uint32_t idesc_all = make_idesc(KT::FORMAT, C::N_ALL), idesc_hi = make_idesc(KT::FORMAT, COUT);
        const bool leader = elect_one();
        int it = 0;
        for (int k = 0; k < n_mine; ++k) {
            const int buf = k % C::NBUF;
            const int use = k / C::NBUF;                            // how many times this buffer was used before
            if (use > 0) mbar_wait(&acc_empty[buf], (uint32_t)((use - 1) & 1));
            tc_fence_after();
            const uint32_t acc0 = tmem_base + (uint32_t)(buf * C::COLS_PER_UNIT);
            for (int st = 0; st < N_STAGES_PER_UNIT; ++st, ++it) {
                const int s = it % STAGES;
                mbar_wait(&ready[s], (uint32_t)((it / STAGES) & 1));
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
                const uint64_t a_hi_desc = make_desc(a_hi, KT::A_LBO, HALO_W * 16);
                const uint64_t a_lo_desc = make_desc(a_hi + KT::A_LO_OFFSET, KT::A_LBO, HALO_W * 16);
                const uint64_t b_desc = make_desc(a_hi + A_BYTES, C::N_ALL * 16, 128);
                const uint32_t aset = (S::ASETS > 1) ? (uint32_t)(st % S::ASETS) : 0u;
                const uint32_t first = (st < S::ASETS) ? 0u : 1u;          // first MMA into this accumulator set
                const uint32_t acc_set = acc0 + aset * (uint32_t)S::COLS_PER_SET;
                if (leader) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                        for (int mt = 0; mt < S::MT; ++mt) {
                            // start-address field is in 16-byte units: one voxel (float4) per unit
                            const uint64_t a_off = (uint64_t)((tap / 3) * S::DIL * HALO_W + 8 * mt + (tap % 3) * S::DIL);
                            const uint64_t b_off = (uint64_t)(tap * (C::W_TAP_BYTES >> 4));
                            const uint32_t acc = acc_set + (uint32_t)(mt * C::N_ALL);
                            umma<KIND>(acc, a_hi_desc + a_off, b_desc + b_off, idesc_all, tap == 0 ? first : 1u);
                            umma<KIND>(acc + COUT, a_lo_desc + a_off, b_desc + b_off, idesc_hi, 1u);
                        }
                    }
                    umma_commit(&empty[s]);                              // stage s may be refilled once these MMAs retire
                    if (st == N_STAGES_PER_UNIT - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8) {
        // ===================== hi/lo splitter =====================
        const int t = tid - 256;
        int it = 0;
        for (int k = 0; k < n_mine; ++k) {
            for (int st = 0; st < N_STAGES_PER_UNIT; ++st, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
                unsigned char* area = smem + (size_t)s * C::STAGE_BYTES;
                if constexpr (KIND == KIND_TF32) {
                    uint4* hi = reinterpret_cast<uint4*>(area);
                    uint4* lo = reinterpret_cast<uint4*>(area + KT::A_LO_OFFSET);
                    for (int i = t; i < 2 * HALO_VOX; i += SPLIT_THREADS) {
                        const uint4 x = hi[i];
                        uint4 h, l;
                        h.x = x.x & TF32_MASK; h.y = x.y & TF32_MASK; h.z = x.z & TF32_MASK; h.w = x.w & TF32_MASK;
                        l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x)) & TF32_MASK;
                        l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y)) & TF32_MASK;
                        l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z)) & TF32_MASK;
                        l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w)) & TF32_MASK;
                        hi[i] = h;
                        lo[i] = l;
                    }
                } else {
                    float amax = 0.0f;
                    for (int i = t; i < 2 * HALO_VOX; i += SPLIT_THREADS) {
                        const int pair = i / HALO_VOX, v = i - pair * HALO_VOX;
                        float4* c0 = reinterpret_cast<float4*>(area + (size_t)pair * 2 * S::KGROUP_BYTES) + v;   // channels 8p..8p+3
                        float4* c1 = c0 + HALO_VOX;                                                                // channels 8p+4..8p+7
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
                        *reinterpret_cast<uint4*>(c0) = hv;          // x_hi K-group of this pair
                        *reinterpret_cast<uint4*>(c1) = lv;          // x_lo K-group of this pair
                    }
                    const bool bad = !(amax <= 65504.0f);            // Inf included; NaN propagates through fp16 as NaN, like the reference
                    if (bad && p.status) atomicOr(p.status, 1);
                }
                fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                mbar_arrive(&ready[s]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                 // row of the M tile = voxel (h = m / 8, w = m % 8)
        const int mh = m >> 3, mw = m & 7;
        double gs[2] = {0.0, 0.0}, gq[2] = {0.0, 0.0};
        const size_t vox = (size_t)p.D * p.H * p.W;
        for (int k = 0; k < n_mine; ++k) {
            const int buf = k % C::NBUF;
            const int use = k / C::NBUF;
            int d, h0, w0;
            unit_origin(k, d, h0, w0);
            mbar_wait(&acc_full[buf], (uint32_t)(use & 1));
            tc_fence_after();
            const uint32_t acc0 = tmem_base + (uint32_t)(buf * C::COLS_PER_UNIT) + ((uint32_t)(q * 32) << 16);
            const int h = h0 + mh;
#pragma unroll 1
            for (int mt = 0; mt < S::MT; ++mt) {
                const int w = w0 + 8 * mt + mw;
                const bool ok = (h < p.H) && (w < p.W);
                const size_t pos = ((size_t)d * p.H + h) * p.W + w;
                // truncation-bias compensation (common.cuh): the large-product accumulator(s) received one MMA per k-step for
                // every filter tap inside the volume, shared between ASETS accumulator sets
                const float comp = kTruncBiasPerMma * (float)(NKS * (S::PLANAR ? 1 : taps_inside(d, p.D, 1)) * taps_inside(h, p.H, S::DIL) * taps_inside(w, p.W, S::DIL)) / (float)S::ASETS;
#pragma unroll 1
                for (int c0 = 0; c0 < COUT; c0 += 16) {
                    float a[16], b[16];
                    tmem_ld16(acc0 + (uint32_t)(mt * C::N_ALL + c0), a);
                    tmem_ld16(acc0 + (uint32_t)(mt * C::N_ALL + COUT + c0), b);
                    tmem_ld_wait();
                    if constexpr (S::ASETS == 2) {
                        float a2[16], b2[16];
                        tmem_ld16(acc0 + (uint32_t)(S::COLS_PER_SET + mt * C::N_ALL + c0), a2);
                        tmem_ld16(acc0 + (uint32_t)(S::COLS_PER_SET + mt * C::N_ALL + COUT + c0), b2);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) { a[i] += a2[i]; b[i] += b2[i]; }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], comp, a[i]) + b[i];
                    float ts[2] = {0.f, 0.f}, tq[2] = {0.f, 0.f};
                    conv_epilogue_store16(p.ep, a, c0, ok, pos, vox, ts, tq);
                    gs[0] += (double)ts[0]; gq[0] += (double)tq[0];
                    gs[1] += (double)ts[1]; gq[1] += (double)tq[1];
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
        if (p.ep.gn_partials) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], o); gq[0] += __shfl_xor_sync(0xffffffffu, gq[0], o);
                gs[1] += __shfl_xor_sync(0xffffffffu, gs[1], o); gq[1] += __shfl_xor_sync(0xffffffffu, gq[1], o);
            }
            if (lane == 0) { s_red[q][0] = gs[0]; s_red[q][1] = gq[0]; s_red[q][2] = gs[1]; s_red[q][3] = gq[1]; }
            asm volatile("bar.sync 1, 128;" ::: "memory");            // the 4 epilogue warps only
            if (warp == 4 && lane == 0) {
                double* dst = p.ep.gn_partials + (size_t)blockIdx.x * 4;
                for (int j = 0; j < 4; ++j) dst[j] = ((s_red[0][j] + s_red[1][j]) + s_red[2][j]) + s_red[3][j];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <class S>
static int launch(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    const int tiles_h = (d->H + S::TILE_H - 1) / S::TILE_H, tiles_w = (d->W + S::TILE_W - 1) / S::TILE_W;
    const int n_units = d->D * tiles_h * tiles_w;
    const int grid = n_units < sm_count() ? n_units : sm_count();
    *n_ctas = grid;
    if (count_only) return ESTD_OK;
    ESTD_REQUIRE(d->weight_tc && aligned16(d->weight_tc), "estd_conv3d: tensor-core precision needs a 16-byte aligned weight_tc");
    ESTD_REQUIRE(d->in1_chunks == 0 || (d->in0_chunks % S::CHUNKS) == 0,
                 "estd_conv3d(tensor cores): first input segment must hold a multiple of %d chunks", S::CHUNKS);
    CUtensorMap map0, map1;
    int rc = make_vol4_tensor_map(&map0, d->in0, d->in0_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, S::CHUNKS);
    if (rc) return rc;
    if (d->in1_chunks > 0) rc = make_vol4_tensor_map(&map1, d->in1, d->in1_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, S::CHUNKS);
    else map1 = map0;
    if (rc) return rc;
    Params p;
    p.weight_tc = d->weight_tc;
    p.status = d->status;
    fill_epilogue(&p.ep, d);
    p.in0_chunks = d->in0_chunks;
    p.D = d->D; p.H = d->H; p.W = d->W;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w; p.n_units = n_units;
    auto kern = conv3d_tc_kernel<S>;
    static DeviceOnce attr_set;                  // the opt-in shared-memory limit is a PER-DEVICE function attribute
    if (!attr_set.done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "estd_conv3d(tensor cores): cannot reserve %zu B of shared memory: %s", S::SMEM, cudaGetErrorString(e));
        attr_set.set();
    }
    kern<<<grid, THREADS, S::SMEM, stream>>>(map0, map1, p);
    return check_launch("estd_conv3d(tensor cores)");
}

}  // namespace tc

int dispatch_tc(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    using namespace tc;
    const int cin_chunks = d->in0_chunks + d->in1_chunks;
    const int C = d->cout_pad;
    const int dil = d->dilation > 0 ? d->dilation : 1;
    ESTD_REQUIRE(!d->planar, "estd_conv3d: planar convolutions live in conv2d_tc.cu");
    ESTD_REQUIRE(dil == 1, "estd_conv3d: dilation is supported for planar convolutions only");
    if (d->precision == ESTD_PREC_3XTF32) {
        const int nks = (cin_chunks + 1) / 2;                     // 8 channels per stage
#define ESTD_TF32(NKS, COUT) if (nks == NKS && C == COUT) return launch<Shape<KIND_TF32, NKS, COUT, 4, 1, false>>(d, stream, count_only, n_ctas)
        ESTD_TF32(4, 32); ESTD_TF32(5, 48); ESTD_TF32(5, 32); ESTD_TF32(2, 16); ESTD_TF32(4, 16);
#undef ESTD_TF32
    } else {
        const int nks = (cin_chunks + 3) / 4;                     // 16 channels per stage
#define ESTD_F16(NKS, COUT) if (nks == NKS && C == COUT) return launch<Shape<KIND_F16, NKS, COUT, 4, 1, false>>(d, stream, count_only, n_ctas)
        ESTD_F16(2, 32); ESTD_F16(3, 48); ESTD_F16(3, 32); ESTD_F16(1, 16); ESTD_F16(2, 16);
#undef ESTD_F16
    }
    return fail(ESTD_EUNSUPPORTED, "estd_conv3d(tensor cores): no kernel for %d input chunks -> cout_pad %d (precision %d)",
                cin_chunks, d->cout_pad, d->precision);
}

}  // namespace estd
