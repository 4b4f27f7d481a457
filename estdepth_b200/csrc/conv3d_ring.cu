// K2, plane-ring schedule: 3x3x3 convolution on the tcgen05 tensor cores with the fp16 two-term split of conv3d_tc.cu
// (x = x_hi + x_lo, w * 2^k = w_hi + w_lo; x_hi w_hi + x_hi w_lo + x_lo w_hi accumulated in fp32 in TMEM), re-scheduled so
// that the depth taps ride in the N dimension of the MMA.
//
// Why: with N = Cout = 32 the output-stationary kernel reads a 4 KB A operand from shared memory for every 32 accumulator
// columns: the tensor core waits for operands (profiles/README.md).  Here the INPUT plane is stationary: one halo tile
// of input plane z is loaded and split once and multiplied against the weights of all three depth taps at once,
// N = 3*Cout -- the three column blocks are the accumulators of output planes z-1, z, z+1.  Per product that is one
// third of the A reads, one third of the TMA traffic and one third of the split work.  What remains is the operand
// fetch of an M=128 x N=96 x K=16 MMA (4 KB of A + 3 KB of B = 56 shared-memory wavefronts against 48 cycles of math).
//
//   unit      a column of 16 x 32 voxels (4 M tiles) walked along depth.  The accumulators of the three output planes
//             in flight live in a ring of 3 TMEM slots per M tile (slot = z mod 3; 4 x 3 x Cout columns).
//   weights   the slot <-> depth-tap assignment rotates with z mod 3, and B rows map 1:1 onto D columns, so the packed
//             weights come in the 3 rotations; the stage of input plane z streams rotation z mod 3 (55 KB per 16 channels).
//   MMA       per in-plane tap and M tile: A_hi x W_hi, A_hi x W_lo, A_lo x W_hi, each M = 128, N = 3*Cout, K = 16, all
//             accumulating (slots are zeroed by the epilogue after it has read them, so there is no "first" MMA).
//   epilogue  the column is two half tiles (M tiles 0,1 | 2,3).  After the MMAs of plane z, output plane z-1 is complete
//             and its slot is needed again by plane z+1 (for z+2).  The issuer commits half A, then issues half B; the
//             epilogue warps drain and zero half A's slot while the tensor core works on half B, and vice versa.
//             Within a half the issue order is tap > product > M tile so that consecutive MMAs alternate accumulators.
//             (A 4-slot ring gives the epilogue a whole plane of slack and 4-way interleaving, but two planes in four
//             then wrap around the ring and must be issued as N = 64 + N = 32: measured 207 us against this schedule.)
//   balance   the flat list of (column, plane) pairs is cut into one contiguous range per CTA; a range that starts or
//             ends inside a column pays one partial extra input plane (a single depth tap, N = Cout) on that side.
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "conv3d_common.cuh"
#include "tc_ptx.cuh"
#include "ring_epilogue.cuh"

namespace estd {
namespace ring {

using namespace tc;

constexpr int EPI_WARPS = 8, SPLIT_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32, SPLIT_THREADS = SPLIT_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS + SPLIT_THREADS;       // warps 0-3 control, 4-11 epilogue, 12-19 splitters
constexpr int FIRST_SPLIT_WARP = 4 + EPI_WARPS;

// NKS k-steps of 16 input channels; COUT padded output channels (multiple of 16); MT M tiles (16 x 8 voxels each) per column:
// 4 unless the 3*COUT*MT accumulator columns would not fit TMEM (COUT = 48 -> 2)
template <int NKS_, int COUT_, int MT_>
struct Shape {
    static constexpr int NKS = NKS_, COUT = COUT_, MT = MT_, MH = MT_ / 2;      // MH: M tiles per half tile
    static constexpr int TILE_H = 16, TILE_W = 8 * MT;
    static constexpr int HALO_H = TILE_H + 2, HALO_W = TILE_W + 2, HALO_VOX = HALO_H * HALO_W;
    static constexpr int KGROUP_BYTES = HALO_VOX * 16;            // one 16-byte K-group (8 x fp16) of the halo tile
    static constexpr int A_BYTES = 4 * KGROUP_BYTES;              // 16 channels: lands as 4 fp32 chunks, becomes hi|lo|hi|lo
    static constexpr int N3 = 3 * COUT;                           // weight rows of one rotation = accumulator columns of one M tile
    static constexpr int W_PART_BYTES = 2 * N3 * 16;              // [2 K-groups][N3 rows][16 B] of w_hi (or w_lo)
    static constexpr int W_TAP_BYTES = 2 * W_PART_BYTES;          // w_hi block, then w_lo block
    static constexpr int W_BYTES = 9 * W_TAP_BYTES;
    static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES + 127) / 128 * 128;
    static constexpr int STAGES = (3 * STAGE_BYTES + 2048 <= 227 * 1024) ? 3 : 2;
    static constexpr int COLS = MT * N3;
    static constexpr int TMEM_COLS = (COLS <= 128) ? 128 : (COLS <= 256) ? 256 : 512;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 256;
    static_assert(COLS <= 512, "ring accumulators must fit TMEM");
    static_assert(SMEM + 2048 <= 227 * 1024, "stages must fit shared memory");
    static_assert(COUT % 16 == 0 && COUT <= 48 && N3 <= 256 && (MT == 2 || MT == 4), "bad shape");
};


#ifdef ESTD_RING_TIMING
// experiment only (make EXTRA=-DESTD_RING_TIMING): cycles the MMA issuer spent waiting, per CTA: {ready, acc_empty, total, stages}
__device__ long long g_ring_timing[148 * 4];
#define RING_T0() const long long t_dbg0 = clock64()
#define RING_T1(slot) t_dbg[slot] += clock64() - t_dbg0
#else
#define RING_T0()
#define RING_T1(slot)
#endif

struct Params {
    const float* weight_ring;                   // [3 rotations][NKS][9 taps][hi,lo][2 K-groups][3*COUT rows][16 bytes]
    int* status;
    ConvEpilogue ep;
    int in0_chunks;
    int in0_split, in1_split;                   // the input segments are pre-split (vol4s): the splitter warps pass them through
    int D, H, W;
    int tiles_h, tiles_w;
    int total;                                  // columns * D  (flat (column, plane) index space)
};

// The contiguous piece [f0, f1) of the flat (column, plane) list owned by this CTA, walked segment by segment.
struct Segment { int col, z0, z1; };
__device__ __forceinline__ bool next_segment(int& f, int f1, int D, Segment& s) {
    if (f >= f1) return false;
    s.col = f / D;
    s.z0 = f - s.col * D;
    const int n = min(D - s.z0, f1 - f);
    s.z1 = s.z0 + n;
    f += n;
    return true;
}

template <class S>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_ring_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Params p) {
    constexpr int NKS = S::NKS, COUT = S::COUT, STAGES = S::STAGES, N3 = S::N3;
    constexpr int HALO_W = S::HALO_W, HALO_VOX = S::HALO_VOX, A_BYTES = S::A_BYTES;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * S::STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES] TMA landed
    uint64_t* ready = bars + STAGES;        // [STAGES] split done
    uint64_t* empty = bars + 2 * STAGES;    // [STAGES] MMAs done reading
    uint64_t* acc_full = bars + 3 * STAGES; // [2 halves] an output plane of this half tile is complete
    uint64_t* acc_empty = acc_full + 2;     // [2 halves] its slot has been read and zeroed
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ double s_red[EPI_WARPS][4];
    __shared__ __align__(16) float s_shift[COUT];                    // per-channel offset; the per-channel multiplier is folded into the weights

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    pdl_launch_dependents();                 // the next kernel of the stream may start its prologue as SMs free up
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], SPLIT_THREADS); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128 * S::MH); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, S::TMEM_COLS);
    if (warp == 3) for (int i = lane; i < COUT; i += 32) s_shift[i] = p.ep.shift[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();                              // the prologue above overlapped the previous kernel's tail; activations from here on

    // every ring slot starts at zero: all MMAs accumulate
    if (warp >= 4 && warp < FIRST_SPLIT_WARP) {
        const int e = warp - 4, q = e & 3, part = e >> 2;
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * (S::COLS / 2));
#pragma unroll
        for (int c = 0; c < S::COLS / 2; c += 16) tmem_st16_zero(t0 + (uint32_t)c);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int f_begin = (int)(((long long)p.total * blockIdx.x) / gridDim.x);
    const int f_end = (int)(((long long)p.total * (blockIdx.x + 1)) / gridDim.x);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0, f = f_begin;
            Segment sg;
            while (next_segment(f, f_end, p.D, sg)) {
                const int h0 = (sg.col / p.tiles_w) * S::TILE_H, w0 = (sg.col % p.tiles_w) * S::TILE_W;
                const int zin_hi = min(sg.z1, p.D - 1);
                for (int z = max(sg.z0 - 1, 0); z <= zin_hi; ++z) {
                    const int rot = z % 3;
                    for (int ks = 0; ks < NKS; ++ks, ++it) {
                        const int s = it % STAGES;
                        if (it >= STAGES) mbar_wait_polls(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
                        unsigned char* stage = smem + (size_t)s * S::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(A_BYTES + S::W_BYTES));
                        const int chunk = 4 * ks;
                        if (chunk < p.in0_chunks) tma_load_4d(stage, &map0, &full[s], 4 * (w0 - 1), h0 - 1, z, chunk);
                        else                      tma_load_4d(stage, &map1, &full[s], 4 * (w0 - 1), h0 - 1, z, chunk - p.in0_chunks);
                        bulk_load(stage + A_BYTES, p.weight_ring + (size_t)(rot * NKS + ks) * (S::W_BYTES / 4), (uint32_t)S::W_BYTES, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        int it = 0, f = f_begin;
        int n_sig = 0;                                           // hand-overs issued so far (the same for both halves)
        Segment sg;
#ifdef ESTD_RING_TIMING
        long long t_dbg[2] = {0, 0};
        const long long t_dbg_start = clock64();
#endif
        while (next_segment(f, f_end, p.D, sg)) {
            for (int z = max(sg.z0 - 1, 0); z <= sg.z1; ++z) {
                const bool real = z < p.D;                       // z == D: nothing to add, only the last plane to hand over
                const bool completes = (z - 1) >= sg.z0;         // output plane z-1 is finished after this input plane
                const uint32_t drained = (uint32_t)((n_sig - 1) & 1);   // parity of the previous hand-over's drain
                if (!real) {
                    for (int half = 0; half < 2; ++half) {
                        if (n_sig > 0) mbar_wait_polls(&acc_empty[half], drained);
                        if (leader) umma_commit(&acc_full[half]);
                    }
                    ++n_sig;
                    __syncwarp();
                    continue;
                }
                // active output planes -> ring slots -> runs of adjacent slots (only {0,2} needs two)
                const int o_lo = max(z - 1, sg.z0), o_hi = min(z + 1, sg.z1 - 1);
                uint32_t mask = 0;
                for (int o = o_lo; o <= o_hi; ++o) mask |= 1u << (o % 3);
                const int n_runs = (mask == 5u) ? 2 : 1;
                const int run0_first = (mask & 1u) ? 0 : (mask & 2u) ? 1 : 2;
                const int run0_n = (mask == 5u) ? 1 : __popc(mask);
                for (int ks = 0; ks < NKS; ++ks, ++it) {
                    const int s = it % STAGES;
                    { RING_T0(); mbar_wait_polls(&ready[s], (uint32_t)((it / STAGES) & 1)); RING_T1(0); }
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * S::STAGE_BYTES);
                    const uint64_t a_hi_desc = make_desc(a_hi, 2 * S::KGROUP_BYTES, HALO_W * 16);
                    const uint64_t a_lo_desc = make_desc(a_hi + S::KGROUP_BYTES, 2 * S::KGROUP_BYTES, HALO_W * 16);
                    const uint64_t w_hi_desc = make_desc(a_hi + A_BYTES, N3 * 16, 128);
                    const uint64_t w_lo_desc = make_desc(a_hi + A_BYTES + S::W_PART_BYTES, N3 * 16, 128);
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        if (ks == 0 && n_sig > 0) {
                            // the slot that starts a new output plane now was handed to the epilogue one plane ago
                            { RING_T0(); mbar_wait_polls(&acc_empty[half], drained); RING_T1(1); }
                            tc_fence_after();
                        }
                        if (leader) {
#pragma unroll 1
                            for (int r = 0; r < n_runs; ++r) {
                                const int first = (r == 0) ? run0_first : 2, count = (r == 0) ? run0_n : 1;
                                const uint32_t idesc = make_idesc(0u, count * COUT);
                                const uint32_t acc0 = tmem_base + (uint32_t)(half * S::MH * N3 + first * COUT);
                                const uint64_t wh = w_hi_desc + (uint64_t)(first * COUT), wl = w_lo_desc + (uint64_t)(first * COUT);   // rows = 16-byte units
                                const uint64_t a_base = (uint64_t)(half * S::MH * 8);            // MH M tiles x 8 voxels
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    const uint64_t b_off = (uint64_t)(tap * (S::W_TAP_BYTES >> 4));
#pragma unroll
                                    for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                                        for (int m2 = 0; m2 < S::MH; ++m2) {
                                            const uint64_t a_off = a_base + (uint64_t)((tap / 3) * HALO_W + 8 * m2 + (tap % 3));
                                            const uint32_t acc = acc0 + (uint32_t)(m2 * N3);
                                            umma<KIND_F16>(acc, (prod == 2 ? a_lo_desc : a_hi_desc) + a_off, (prod == 1 ? wl : wh) + b_off, idesc, 1u);
                                        }
                                    }
                                }
                            }
                            if (ks == NKS - 1 && completes) umma_commit(&acc_full[half]);
                        }
                        __syncwarp();
                    }
                    if (leader) umma_commit(&empty[s]);                  // stage s may be refilled once these MMAs retire
                    __syncwarp();
                }
                if (completes) ++n_sig;
            }
        }
#ifdef ESTD_RING_TIMING
        if (leader && blockIdx.x < 148) {
            g_ring_timing[blockIdx.x * 4 + 0] = t_dbg[0]; g_ring_timing[blockIdx.x * 4 + 1] = t_dbg[1];
            g_ring_timing[blockIdx.x * 4 + 2] = clock64() - t_dbg_start; g_ring_timing[blockIdx.x * 4 + 3] = it;
        }
#endif
    } else if (warp >= FIRST_SPLIT_WARP) {
        // ===================== hi/lo splitter =====================
        const int t = tid - FIRST_SPLIT_WARP * 32;
        int it = 0, f = f_begin;
        float amax = 0.0f;
        Segment sg;
        while (next_segment(f, f_end, p.D, sg)) {
            const int n_planes = min(sg.z1, p.D - 1) - max(sg.z0 - 1, 0) + 1;
            for (int st = 0; st < n_planes * NKS; ++st, ++it) {
                const int s = it % STAGES;
                if (warp == FIRST_SPLIT_WARP) mbar_wait_polls(&full[s], (uint32_t)((it / STAGES) & 1));
                named_barrier(2, SPLIT_THREADS);                 // one warp polls, the barrier releases the other seven
                unsigned char* area = smem + (size_t)s * S::STAGE_BYTES;
                const bool presplit = (4 * (st % NKS) < p.in0_chunks) ? (p.in0_split != 0) : (p.in1_split != 0);
#ifndef ESTD_EXP_NOSPLIT      // timing experiment only: skip the hi/lo split (operands are garbage)
                for (int i = t; !presplit && i < 2 * HALO_VOX; i += SPLIT_THREADS) {
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
#endif
                fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                mbar_arrive(&ready[s]);
            }
        }
        const bool bad = !(amax <= 65504.0f);        // Inf included; NaN propagates through fp16 as NaN, like the reference
        if (bad && p.status) atomicOr(p.status, 1);
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int e = warp - 4, q = e & 3, m2 = e >> 2;          // TMEM lane quarter; which M tile of each half tile
        const int m = q * 32 + lane;                             // row of the M tile = voxel (h = m / 8, w = m % 8)
        const int mh = m >> 3, mw = m & 7;
        double gs[2] = {0.0, 0.0}, gq[2] = {0.0, 0.0};
        const size_t vox = (size_t)p.D * p.H * p.W;
        const ConvEpilogue& ep = p.ep;
        const bool want_gn = ep.gn_partials != nullptr;
        const float mult = __ldg(ep.scale);                      // uniform: 2^-k of the fp16 weight scaling (pack_weight_ring)
        int n_seen = 0;
        int f = (m2 < S::MH) ? f_begin : f_end;                  // with 2 M tiles per column only the first 4 warps have work
        Segment sg;
        while (next_segment(f, f_end, p.D, sg)) {
            const int h0 = (sg.col / p.tiles_w) * S::TILE_H, w0 = (sg.col % p.tiles_w) * S::TILE_W;
            const int h = h0 + mh;
            for (int z = sg.z0; z < sg.z1; ++z, ++n_seen) {
                const int slot = z % 3;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int mt = S::MH * half + m2;
                    const int w = w0 + 8 * mt + mw;
                    const bool ok = (h < p.H) && (w < p.W);
                    const size_t pos = ((size_t)z * p.H + h) * p.W + w;
                    // truncation-bias compensation (common.cuh): this voxel's slot received 3 products x NKS k-steps for every
                    // filter tap that lies inside the volume
                    const float mult_v = mult * (1.0f + kTruncBiasPerMma * (float)(3 * NKS * taps_inside(z, p.D, 1) * taps_inside(h, p.H, 1) * taps_inside(w, p.W, 1)));
                    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * N3 + slot * COUT);
                    if (e == 0) mbar_wait_polls(&acc_full[half], (uint32_t)(n_seen & 1));
                    named_barrier(3, 128 * S::MH);
                    tc_fence_after();
                    ring_drain_slot<COUT, false, 0, 0, 1>(ep, s_shift, mult_v, 0.0f, t0, ok, pos, vox, want_gn, gs, gq, [&]() { mbar_arrive(&acc_empty[half]); });
                }
            }
        }
        if (ep.gn_partials) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], o); gq[0] += __shfl_xor_sync(0xffffffffu, gq[0], o);
                gs[1] += __shfl_xor_sync(0xffffffffu, gs[1], o); gq[1] += __shfl_xor_sync(0xffffffffu, gq[1], o);
            }
            if (lane == 0) { s_red[e][0] = gs[0]; s_red[e][1] = gq[0]; s_red[e][2] = gs[1]; s_red[e][3] = gq[1]; }
            named_barrier(1, EPI_THREADS);                            // the 8 epilogue warps only
            if (e == 0 && lane == 0) {
                double* dst = ep.gn_partials + (size_t)blockIdx.x * 4;
                for (int j = 0; j < 4; ++j) {
                    double acc = s_red[0][j];
                    for (int k = 1; k < EPI_WARPS; ++k) acc += s_red[k][j];
                    dst[j] = acc;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

template <class S>
static int launch(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    const int tiles_h = (d->H + S::TILE_H - 1) / S::TILE_H, tiles_w = (d->W + S::TILE_W - 1) / S::TILE_W;
    const long long total = (long long)tiles_h * tiles_w * d->D;
    ESTD_REQUIRE(total < (1ll << 30), "estd_conv3d(ring): volume too large");
    const int grid = total < sm_count() ? (int)total : sm_count();
    *n_ctas = grid;
    if (count_only) return ESTD_OK;
    ESTD_REQUIRE(d->weight_tc && aligned16(d->weight_tc), "estd_conv3d(ring): needs a 16-byte aligned ring weight packing in weight_tc");
    ESTD_REQUIRE(d->in1_chunks == 0 || (d->in0_chunks % 4) == 0, "estd_conv3d(ring): first input segment must hold a multiple of 4 chunks");
    CUtensorMap map0, map1;
    int rc = make_vol4_tensor_map(&map0, d->in0, d->in0_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, 4);
    if (rc) return rc;
    if (d->in1_chunks > 0) rc = make_vol4_tensor_map(&map1, d->in1, d->in1_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, 4);
    else map1 = map0;
    if (rc) return rc;
    Params p;
    p.weight_ring = d->weight_tc;
    p.status = d->status;
    fill_epilogue(&p.ep, d);
    p.in0_chunks = d->in0_chunks;
    p.in0_split = d->in0_split; p.in1_split = d->in1_split;
    p.D = d->D; p.H = d->H; p.W = d->W;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w; p.total = (int)total;
    auto kern = conv3d_ring_kernel<S>;
    static DeviceOnce attr_set;                  // the opt-in shared-memory limit is a PER-DEVICE function attribute
    if (!attr_set.done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "estd_conv3d(ring): cannot reserve %zu B of shared memory: %s", S::SMEM, cudaGetErrorString(e));
        attr_set.set();
    }
    {
        cudaError_t e = launch_pdl(kern, grid, THREADS, S::SMEM, stream, map0, map1, p);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "%s launch: %s", __FILE__, cudaGetErrorString(e));
    }
    return check_launch("estd_conv3d(ring)");
}

}  // namespace ring

#ifdef ESTD_RING_TIMING
extern "C" __attribute__((visibility("default"))) int estd_ring_timing(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, ring::g_ring_timing, sizeof(long long) * 148 * 4) == cudaSuccess ? 0 : -2;
}
#endif

int dispatch_ring(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    using namespace ring;
    const int cin_chunks = d->in0_chunks + d->in1_chunks;
    const int nks = (cin_chunks + 3) / 4;                         // 16 channels per stage
    ESTD_REQUIRE(!d->planar && (d->dilation == 0 || d->dilation == 1), "estd_conv3d(ring): 3x3x3, dilation 1 only");
#define ESTD_RING(NKS, COUT, MT) if (nks == NKS && d->cout_pad == COUT) return launch<Shape<NKS, COUT, MT>>(d, stream, count_only, n_ctas)
    ESTD_RING(2, 32, 4); ESTD_RING(3, 32, 4); ESTD_RING(1, 16, 4); ESTD_RING(2, 16, 4); ESTD_RING(3, 48, 2);
#undef ESTD_RING
    return fail(ESTD_EUNSUPPORTED, "estd_conv3d(ring): no kernel for %d input chunks -> cout_pad %d", cin_chunks, d->cout_pad);
}

}  // namespace estd
