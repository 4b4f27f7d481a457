// K2, plane-ring schedule on CTA PAIRS (tcgen05 cta_group::2): the same algorithm as conv3d_ring.cu -- input plane stationary,
// the three depth taps in the N dimension of the MMA, a ring of 3 TMEM slots per M tile, fp16 two-term split -- with two
// CTAs of a cluster (two SMs of a TPC) sharing every MMA: M = 256 (each CTA's own 128 voxels of its own column), N = 3*Cout,
// and each CTA holds only HALF of the weight rows.  Why: conv3d_ring.cu is bound by the 128 B/clk shared-memory pipe
// (profiles/README.md: 56 wavefronts of operand fetch per 48-cycle MMA, plus the weight fill and the hi/lo split); with the
// pair, the B fetch and the weight fill per SM halve: 32 + 12 = 44 wavefronts per MMA, i.e. the MMA math becomes the bound.
//
// Differences from the single-CTA kernel:
//   * a cluster walks a PAIR of adjacent columns (rank 0: column 2p, rank 1: column 2p+1) over the same plane range; the
//     flat (column pair, plane) list is cut into one contiguous range per cluster;
//   * every plane issues the full N = 3*Cout MMA; partial first/last planes of a range (depth taps that belong to a
//     neighbouring range or fall outside the volume) use a packed weight VARIANT in which those taps are zero
//     (packing.pack_weight_ring2: 7 live-tap masks x 3 rotations), so the B descriptor is the same in both CTAs;
//   * each CTA runs its own TMA producer (its halo tile + its half of the weight rows), splitter and epilogue warps on its own
//     shared memory / TMEM; only rank 0 issues MMAs.  "stage ready" and "slot drained" reach the issuer as ONE remote
//     mbarrier arrival per CTA (after a CTA-local named barrier); "stage free" and "plane complete" come back to both CTAs
//     through multicast tcgen05.commit.
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "conv3d_common.cuh"
#include "tc_ptx.cuh"
#include "ring_epilogue.cuh"

namespace estd {
namespace ring2 {

using namespace tc;

constexpr int EPI_WARPS = 8, SPLIT_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32, SPLIT_THREADS = SPLIT_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS + SPLIT_THREADS;       // warps 0-3 control, 4-11 epilogue, 12-19 splitters
constexpr int FIRST_SPLIT_WARP = 4 + EPI_WARPS;

// DUAL: two accumulators per ring slot (precision 3xf16r2d).  The tensor core adds into its fp32 accumulator with truncation;
// with all three products of the split in one accumulator every output takes 3 x 27 x NKS truncating adds at its full
// magnitude, and although the mean of that error is compensated (common.cuh) its spread is three times that of the
// large products alone.  DUAL keeps the small products (x_hi w_lo, x_lo w_hi) in a second accumulator, N3 columns further, which
// the epilogue adds in round-to-nearest: the error of a layer drops to that of the exact-fp32 kernel (profiles/trunc_probe.py),
// at twice the TMEM columns -- 2 M tiles per CTA instead of 4 for the 32-channel layers.
template <int NKS_, int COUT_, int MT_, bool DUAL_ = false>
struct Shape {
    static constexpr int NKS = NKS_, COUT = COUT_, MT = MT_, MH = MT_ / 2;      // MH: M tiles per half tile
    static constexpr bool DUAL = DUAL_;
    static constexpr int TILE_H = 16, TILE_W = 8 * MT;
    static constexpr int HALO_H = TILE_H + 2, HALO_W = TILE_W + 2, HALO_VOX = HALO_H * HALO_W;
    static constexpr int KGROUP_BYTES = HALO_VOX * 16;
    static constexpr int A_BYTES = 4 * KGROUP_BYTES;
    // N of the MMA = accumulator columns of one M tile: three slots of COUT columns each, rounded up to the MMA's granularity
    // (COUT = 33: 99 -> 112; the 13 extra columns belong to zero weight rows and are never read)
    static constexpr int N3 = (3 * COUT + 15) / 16 * 16;
    static constexpr int SHIFT_N = (COUT + 15) / 16 * 16;         // per-channel offsets kept in shared memory
    static constexpr int NH = N3 / 2;                             // weight rows held by one CTA of the pair
    static constexpr int W_PART_BYTES = 2 * NH * 16;              // [2 K-groups][NH rows][16 B] of w_hi (or w_lo)
    // DUAL: per tap a "merged" block [2 K-groups][N3 rows] -- rank 0 holds ALL of w_hi, rank 1 ALL of w_lo, so that ONE MMA with
    // N = 2*N3 and A = x_hi writes x_hi w_hi into the large accumulator and x_hi w_lo into the small one next to it -- and a
    // "third" block [2 K-groups][NH rows] with this rank's half of w_hi for A = x_lo.  Two MMAs per tap and M tile instead of
    // three: one fetch of the A tile (32 of the ~44 shared-memory wavefronts of an MMA) is saved.
    static constexpr int W_MERGED_BYTES = 2 * N3 * 16;
    static constexpr int W_TAP_BYTES = DUAL ? (W_MERGED_BYTES + W_PART_BYTES) : 2 * W_PART_BYTES;
    static_assert(!DUAL || 2 * N3 <= 256, "the merged MMA needs N = 2 * N3 <= 256");
    static constexpr int W_BYTES = 9 * W_TAP_BYTES;               // per CTA and stage
    static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES + 127) / 128 * 128;
    static constexpr int STAGES = (DUAL && 4 * STAGE_BYTES + 2048 <= 227 * 1024) ? 4 : (3 * STAGE_BYTES + 2048 <= 227 * 1024) ? 3 : 2;
    static constexpr int ACC_N = DUAL ? 2 * N3 : N3;             // accumulator columns of one M tile (large | small products)
    static constexpr int COLS = MT * ACC_N;
    static constexpr int TMEM_COLS = 512;                         // the pair allocates symmetrically: everything
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 256;
    static_assert(COLS <= 512, "ring accumulators must fit TMEM");
    static_assert(SMEM + 2048 <= 227 * 1024, "stages must fit shared memory");
    static_assert(COUT % 16 <= 1 && COUT <= 48 && N3 <= 256 && N3 % 16 == 0 && (MT == 2 || MT == 4), "bad shape");
};


#ifdef ESTD_RING_TIMING
__device__ long long g_ring2_timing[148 * 4];
__device__ long long g_ring2_epi[148 * 8];      // per CTA, epilogue warp 4: {cycles in tcgen05.ld + wait, cycles full->arrive, hand-overs, cycles waiting for acc_full}
#define RING_T0() const long long t_dbg0 = clock64()
#define RING_T1(slot) t_dbg[slot] += clock64() - t_dbg0
#else
#define RING_T0()
#define RING_T1(slot)
#endif

struct Params {
    const float* weight_ring2;                  // [7 live-tap masks][3 rotations][NKS][2 CTAs][9 taps][hi,lo][2 K-groups][NH rows][16 bytes]
                                                // DUAL: [...][9 taps]{[2 K-groups][N3 rows] merged, [2 K-groups][NH rows] third}[16 bytes]
    int* status;
    ConvEpilogue ep;
    int in0_chunks;
    int in0_split, in1_split;                   // the input segments are pre-split (vol4s): the splitter warps pass them through
    int D, H, W;
    int tiles_h, tiles_w;
    int total;                                  // column pairs * D  (flat (column pair, plane) index space)
};

// The contiguous piece [f0, f1) of the flat (column pair, plane) list owned by this cluster, walked segment by segment.
struct Segment { int pair, z0, z1; };
__device__ __forceinline__ bool next_segment(int& f, int f1, int D, Segment& s) {
    if (f >= f1) return false;
    s.pair = f / D;
    s.z0 = f - s.pair * D;
    const int n = min(D - s.z0, f1 - f);
    s.z1 = s.z0 + n;
    f += n;
    return true;
}
// depth taps of input plane z that land on output planes of [z0, z1): bit kd set <=> output z + 1 - kd is in the range
__device__ __forceinline__ int live_taps(int z, int z0, int z1) {
    return ((z + 1 < z1) ? 1 : 0) | ((z >= z0 && z < z1) ? 2 : 0) | ((z - 1 >= z0) ? 4 : 0);
}

template <class S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
conv3d_ring2_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Params p) {
    constexpr int NKS = S::NKS, COUT = S::COUT, STAGES = S::STAGES, N3 = S::N3;
    constexpr int HALO_W = S::HALO_W, HALO_VOX = S::HALO_VOX, A_BYTES = S::A_BYTES;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * S::STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES] own TMA landed
    uint64_t* ready = bars + STAGES;        // [STAGES] (rank 0's copy is used) both CTAs have split their stage
    uint64_t* empty = bars + 2 * STAGES;    // [STAGES] MMAs done reading (multicast commit)
    uint64_t* acc_full = bars + 3 * STAGES; // [2 halves] an output plane of this half tile is complete (multicast commit)
    uint64_t* acc_empty = acc_full + 2;     // [2 halves] (rank 0's copy is used) both CTAs have drained and zeroed the slot
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ double s_red[EPI_WARPS][4];
    __shared__ __align__(16) float s_shift[S::SHIFT_N];              // per-channel offset; the per-channel multiplier is folded into the weights

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();

    pdl_launch_dependents();                 // the next kernel of the stream may start its prologue as SMs free up
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 2); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2(tmem_base_smem, S::TMEM_COLS);
    if (warp == 3) for (int i = lane; i < S::SHIFT_N; i += 32) s_shift[i] = p.ep.shift[i];
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
    cluster_sync_all();                      // barriers initialised and accumulators zero in BOTH CTAs before anything is signalled
    tc_fence_after();

    const int n_clusters = (int)gridDim.x / 2, cluster = (int)blockIdx.x / 2;
    const int f_begin = (int)(((long long)p.total * cluster) / n_clusters);
    const int f_end = (int)(((long long)p.total * (cluster + 1)) / n_clusters);
    auto column_origin = [&](int pair, int& h0, int& w0) {
        const int col = 2 * pair + (int)rank;                    // may be one past the last column: everything out of bounds
        h0 = (col / p.tiles_w) * S::TILE_H; w0 = (col % p.tiles_w) * S::TILE_W;
    };

    if (warp == 0) {
        // ===================== TMA producer (each CTA: its halo tile, its half of the weight rows) =====================
        if (lane == 0) {
            int it = 0, f = f_begin;
            Segment sg;
            while (next_segment(f, f_end, p.D, sg)) {
                int h0, w0;
                column_origin(sg.pair, h0, w0);
                const int zin_hi = min(sg.z1, p.D - 1);
                for (int z = max(sg.z0 - 1, 0); z <= zin_hi; ++z) {
                    const int variant = live_taps(z, sg.z0, sg.z1) - 1, rot = z % 3;
                    for (int ks = 0; ks < NKS; ++ks, ++it) {
                        const int s = it % STAGES;
                        if (it >= STAGES) mbar_wait_polls(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
                        unsigned char* stage = smem + (size_t)s * S::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(A_BYTES + S::W_BYTES));
                        const int chunk = 4 * ks;
                        if (chunk < p.in0_chunks) tma_load_4d(stage, &map0, &full[s], 4 * (w0 - 1), h0 - 1, z, chunk);
                        else                      tma_load_4d(stage, &map1, &full[s], 4 * (w0 - 1), h0 - 1, z, chunk - p.in0_chunks);
                        const size_t widx = ((size_t)((variant * 3 + rot) * NKS + ks) * 2 + rank) * (S::W_BYTES / 4);
                        bulk_load(stage + A_BYTES, p.weight_ring2 + widx, (uint32_t)S::W_BYTES, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (rank 0 only) =====================
        if (rank == 0) {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc2(0u, N3);
            const uint32_t idesc_merged = make_idesc2(0u, 2 * N3);
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
                    const uint32_t drained = (uint32_t)((n_sig - 1) & 1);
                    if (!real) {
                        for (int half = 0; half < 2; ++half) {
                            if (n_sig > 0) mbar_wait_polls(&acc_empty[half], drained);
                            if (leader) umma_commit2(&acc_full[half], 3);
                        }
                        ++n_sig;
                        __syncwarp();
                        continue;
                    }
                    for (int ks = 0; ks < NKS; ++ks, ++it) {
                        const int s = it % STAGES;
                        { RING_T0(); mbar_wait_polls(&ready[s], (uint32_t)((it / STAGES) & 1)); RING_T1(0); }
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + (size_t)s * S::STAGE_BYTES);
                        const uint64_t a_hi_desc = make_desc(a_hi, 2 * S::KGROUP_BYTES, HALO_W * 16);
                        const uint64_t a_lo_desc = make_desc(a_hi + S::KGROUP_BYTES, 2 * S::KGROUP_BYTES, HALO_W * 16);
                        // single accumulator: w_hi / w_lo halves of this rank.  DUAL: "w_hi_desc" = the merged block (N3 rows per K-group:
                        // w_hi in rank 0, w_lo in rank 1), "w_lo_desc" = the third block (this rank's half of w_hi)
                        const uint64_t w_hi_desc = S::DUAL ? make_desc(a_hi + A_BYTES, N3 * 16, 128) : make_desc(a_hi + A_BYTES, S::NH * 16, 128);
                        const uint64_t w_lo_desc = S::DUAL ? make_desc(a_hi + A_BYTES + S::W_MERGED_BYTES, S::NH * 16, 128)
                                                           : make_desc(a_hi + A_BYTES + S::W_PART_BYTES, S::NH * 16, 128);
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {
                            if (ks == 0 && n_sig > 0) {
                                { RING_T0(); mbar_wait_polls(&acc_empty[half], drained); RING_T1(1); }
                                tc_fence_after();
                            }
                            if (leader) {
                                const uint32_t acc0 = tmem_base + (uint32_t)(half * S::MH * S::ACC_N);
                                const uint64_t a_base = (uint64_t)(half * S::MH * 8);
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    const uint64_t b_off = (uint64_t)(tap * (S::W_TAP_BYTES >> 4));
                                    if constexpr (S::DUAL) {
#pragma unroll
                                        for (int m2 = 0; m2 < S::MH; ++m2) {
                                            const uint64_t a_off = a_base + (uint64_t)((tap / 3) * HALO_W + 8 * m2 + (tap % 3));
                                            const uint32_t acc = acc0 + (uint32_t)(m2 * S::ACC_N);
                                            // x_hi [w_hi | w_lo] -> large accumulator | small accumulator (N = 2 * N3), then x_lo w_hi -> small
                                            umma2_f16(acc, a_hi_desc + a_off, w_hi_desc + b_off, idesc_merged, 1u);
                                            umma2_f16(acc + (uint32_t)N3, a_lo_desc + a_off, w_lo_desc + b_off, idesc, 1u);
                                        }
                                    } else {
#pragma unroll
                                        for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                                            for (int m2 = 0; m2 < S::MH; ++m2) {
                                                const uint64_t a_off = a_base + (uint64_t)((tap / 3) * HALO_W + 8 * m2 + (tap % 3));
                                                const uint32_t acc = acc0 + (uint32_t)(m2 * S::ACC_N);
                                                umma2_f16(acc, (prod == 2 ? a_lo_desc : a_hi_desc) + a_off, (prod == 1 ? w_lo_desc : w_hi_desc) + b_off, idesc, 1u);
                                            }
                                        }
                                    }
                                }
                                if (ks == NKS - 1 && completes) umma_commit2(&acc_full[half], 3);
                            }
                            __syncwarp();
                        }
                        if (leader) umma_commit2(&empty[s], 3);          // both CTAs may refill stage s once these MMAs retire
                        __syncwarp();
                    }
                    if (completes) ++n_sig;
                }
            }
#ifdef ESTD_RING_TIMING
            if (leader) {
                long long* o = g_ring2_timing + (blockIdx.x / 2) * 4;
                o[0] = t_dbg[0]; o[1] = t_dbg[1]; o[2] = clock64() - t_dbg_start; o[3] = it;
            }
#endif
        }
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
                named_barrier(2, SPLIT_THREADS);
                unsigned char* area = smem + (size_t)s * S::STAGE_BYTES;
                const bool presplit = (4 * (st % NKS) < p.in0_chunks) ? (p.in0_split != 0) : (p.in1_split != 0);
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
                fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                named_barrier(4, SPLIT_THREADS);     // the whole tile is split ...
                if (t == 0) mbar_arrive_remote(&ready[s], 0);        // ... one arrival per CTA on the issuer's barrier
            }
        }
        const bool bad = !(amax <= 65504.0f);
        if (bad && p.status) atomicOr(p.status, 1);
    } else if (warp >= 4) {
        // ===================== epilogue (own column, own TMEM) =====================
        // Each group of four epilogue warps (warps 4-7, warps 8-11) OWNS one half tile for good and drains its MH M tiles plane after
        // plane.  A drain is TMEM -> registers, zero, hand the slot back, then affine / activation / residual loads / stores; the
        // issuer needs the slot back by the time it has issued the OTHER half's MMAs of the next plane.  With both groups working
        // on the same half (the round-1 layout) the loads / stores of half 0 had to finish before half 1's drain could even
        // start: measured 239 us instead of 202 us for a 32->32 layer with a residual at 2 M tiles per CTA, and the 16-channel
        // layers (one stage per plane) ran epilogue-bound at 40 % tensor-pipe activity.
        const int e = warp - 4, q = e & 3, half = e >> 2;
        const int m = q * 32 + lane;
        const int mh = m >> 3, mw = m & 7;
        double gs[2] = {0.0, 0.0}, gq[2] = {0.0, 0.0};
        const size_t vox = (size_t)p.D * p.H * p.W;
        const ConvEpilogue& ep = p.ep;
        const bool want_gn = ep.gn_partials != nullptr;
        const float mult = __ldg(ep.scale);                      // uniform: 2^-k of the fp16 weight scaling (pack_weight_ring)
        const int bar_full = half ? 6 : 3, bar_back = half ? 7 : 5;   // named barriers of this group (128 threads)
        int n_seen = 0;
        int f = f_begin;
        Segment sg;
        while (next_segment(f, f_end, p.D, sg)) {
            int h0, w0;
            column_origin(sg.pair, h0, w0);
            const int h = h0 + mh;
            for (int z = sg.z0; z < sg.z1; ++z, ++n_seen) {
                const int slot = z % 3;
                // residuals of the NEXT plane -> L2, one plane period ahead of their use.  The drain cannot hold them in registers
                // (two accumulators per slot fill the budget) and consumes every load right after issuing it: one exposed latency per
                // 16-channel block and residual.  32->32 at cfg2 size with 0 / 1 / 2 residuals: 171 / 197 / 269 us before, 171 / 190 /
                // 247 us with the prefetch (profiles/ring_residual_r02.txt).  Staging the tile in shared memory by TMA instead was
                // measured and is no better (195 / 245 us): what is left is not latency but the residual's bytes competing with the
                // tensor core's operand fetches for the shared-memory / L1 data pipe.
                if (ep.res0 && z + 1 < sg.z1) {
#pragma unroll
                    for (int m2 = 0; m2 < S::MH; ++m2) {
                        const int w = w0 + 8 * (S::MH * half + m2) + mw;
                        if (h < p.H && w < p.W) {
                            const size_t posn = ((size_t)(z + 1) * p.H + h) * p.W + w;
#pragma unroll
                            for (int ch = 0; ch < (COUT + 3) / 4; ++ch) {
                                if (ch >= ep.out_chunks) break;
                                prefetch_l2(ep.res0 + ((size_t)ch * vox + posn) * 4);
                                if (ep.res1) prefetch_l2(ep.res1 + ((size_t)ch * vox + posn) * 4);
                            }
                        }
                    }
                }
                if (q == 0) mbar_wait_polls(&acc_full[half], (uint32_t)(n_seen & 1));
                named_barrier(bar_full, 128);
                tc_fence_after();
                auto hand_back = [&]() {
                    named_barrier(bar_back, 128);                                        // this CTA's slot is drained and zeroed ...
                    if (q == 0 && lane == 0) mbar_arrive_remote(&acc_empty[half], 0);    // ... one arrival per CTA on the issuer's barrier
                };
#pragma unroll
                for (int m2 = 0; m2 < S::MH; ++m2) {
                    const int mt = S::MH * half + m2;
                    const int w = w0 + 8 * mt + mw;
                    const bool ok = (h < p.H) && (w < p.W);
                    const size_t pos = ((size_t)z * p.H + h) * p.W + w;
                    // truncation-bias compensation (common.cuh): NKS k-steps for every filter tap that lies inside the volume, times
                    // the products that share the accumulator (3; DUAL: only x_hi w_hi)
                    const float comp = kTruncBiasPerMma * (float)((S::DUAL ? 1 : 3) * NKS * taps_inside(z, p.D, 1) * taps_inside(h, p.H, 1) * taps_inside(w, p.W, 1));
                    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * S::ACC_N + slot * COUT);
                    const float mult_v = S::DUAL ? mult : mult * (1.0f + comp);
                    if (m2 == S::MH - 1) ring_drain_slot<COUT, S::DUAL, N3, 0, 1>(ep, s_shift, mult_v, comp, t0, ok, pos, vox, want_gn, gs, gq, hand_back);
                    else                 ring_drain_slot<COUT, S::DUAL, N3, 0, 1>(ep, s_shift, mult_v, comp, t0, ok, pos, vox, want_gn, gs, gq, []() {});
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
    cluster_sync_all();                      // no CTA of the pair leaves while the other may still signal it / read its smem
    if (warp == 2) tmem_dealloc2(tmem_base, S::TMEM_COLS);
}

template <class S>
static int launch(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    const int tiles_h = (d->H + S::TILE_H - 1) / S::TILE_H, tiles_w = (d->W + S::TILE_W - 1) / S::TILE_W;
    const long long pairs = ((long long)tiles_h * tiles_w + 1) / 2;
    const long long total = pairs * d->D;
    ESTD_REQUIRE(total < (1ll << 30), "estd_conv3d(ring2): volume too large");
    const int max_clusters = sm_count() / 2;
    const int n_clusters = total < max_clusters ? (int)total : max_clusters;
    const int grid = 2 * n_clusters;
    *n_ctas = grid;
    if (count_only) return ESTD_OK;
    ESTD_REQUIRE(d->weight_tc && aligned16(d->weight_tc), "estd_conv3d(ring2): needs a 16-byte aligned ring2 weight packing in weight_tc");
    ESTD_REQUIRE(d->in1_chunks == 0 || (d->in0_chunks % 4) == 0, "estd_conv3d(ring2): first input segment must hold a multiple of 4 chunks");
    CUtensorMap map0, map1;
    int rc = make_vol4_tensor_map(&map0, d->in0, d->in0_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, 4);
    if (rc) return rc;
    if (d->in1_chunks > 0) rc = make_vol4_tensor_map(&map1, d->in1, d->in1_chunks, d->D, d->H, d->W, S::HALO_W * 4, S::HALO_H, 1, 4);
    else map1 = map0;
    if (rc) return rc;
    Params p;
    p.weight_ring2 = d->weight_tc;
    p.status = d->status;
    fill_epilogue(&p.ep, d);
    p.in0_chunks = d->in0_chunks;
    p.in0_split = d->in0_split; p.in1_split = d->in1_split;
    p.D = d->D; p.H = d->H; p.W = d->W;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w; p.total = (int)total;
    auto kern = conv3d_ring2_kernel<S>;
    static DeviceOnce attr_set;                  // the opt-in shared-memory limit is a PER-DEVICE function attribute
    if (!attr_set.done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "estd_conv3d(ring2): cannot reserve %zu B of shared memory: %s", S::SMEM, cudaGetErrorString(e));
        attr_set.set();
    }
    {
        cudaError_t e = launch_pdl(kern, grid, THREADS, S::SMEM, stream, map0, map1, p);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "%s launch: %s", __FILE__, cudaGetErrorString(e));
    }
    return check_launch("estd_conv3d(ring2)");
}

}  // namespace ring2

#ifdef ESTD_RING_TIMING
extern "C" __attribute__((visibility("default"))) int estd_ring2_timing(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, ring2::g_ring2_timing, sizeof(long long) * 148 * 4) == cudaSuccess ? 0 : -2;
}
extern "C" __attribute__((visibility("default"))) int estd_ring2_epi_timing(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, ring2::g_ring2_epi, sizeof(long long) * 148 * 8) == cudaSuccess ? 0 : -2;
}
#endif

int dispatch_ring2(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas) {
    using namespace ring2;
    const int cin_chunks = d->in0_chunks + d->in1_chunks;
    const int nks = (cin_chunks + 3) / 4;                         // 16 channels per stage
    ESTD_REQUIRE(!d->planar && (d->dilation == 0 || d->dilation == 1), "estd_conv3d(ring2): 3x3x3, dilation 1 only");
    if (d->precision == ESTD_PREC_3XF16_RING2D) {
        // two accumulators per slot: 2 * MT * N3 <= 512 columns
#define ESTD_RING2D(NKS, COUT, MT) if (nks == NKS && d->cout_pad == COUT) return launch<Shape<NKS, COUT, MT, true>>(d, stream, count_only, n_ctas)
        ESTD_RING2D(2, 32, 2); ESTD_RING2D(3, 32, 2); ESTD_RING2D(1, 16, 4); ESTD_RING2D(2, 16, 4); ESTD_RING2D(3, 33, 2);
#undef ESTD_RING2D
        return fail(ESTD_EUNSUPPORTED, "estd_conv3d(ring2, dual accumulators): no kernel for %d input chunks -> cout_pad %d", cin_chunks, d->cout_pad);
    }
#define ESTD_RING2(NKS, COUT, MT) if (nks == NKS && d->cout_pad == COUT) return launch<Shape<NKS, COUT, MT>>(d, stream, count_only, n_ctas)
    ESTD_RING2(2, 32, 4); ESTD_RING2(3, 32, 4); ESTD_RING2(1, 16, 4); ESTD_RING2(2, 16, 4); ESTD_RING2(3, 48, 2);
    ESTD_RING2(3, 33, 4);         // dres2 (36 -> 33 channels): 33-column slots, N = 112, so that 4 M tiles fit TMEM (448 columns)
#undef ESTD_RING2
    return fail(ESTD_EUNSUPPORTED, "estd_conv3d(ring2): no kernel for %d input chunks -> cout_pad %d", cin_chunks, d->cout_pad);
}

}  // namespace estd
