// Epilogue of the plane-ring kernels (conv3d_ring.cu, conv3d_ring2.cu): drain one ring slot of one M tile.
#pragma once
#include "common.cuh"
#include "conv3d_common.cuh"
#include "tc_ptx.cuh"

namespace estd {
namespace tc {

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one 32-bit column (the 33rd channel of a 33-channel slot)
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st1_zero(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(0u) : "memory");
}

// The accumulator slot at TMEM address t0 (COUT columns, this thread's lane = its voxel) is complete.  COUT is a multiple of 16,
// or a multiple of 16 plus ONE (the 33-channel layer dres2, hybrid_depth_decoder.py:90: its slots are 33 columns wide so that
// the three of them fit N = 112 instead of 3 x 48 = 144).
//   1. TMEM -> registers, zero the slot, hand it back to the MMA issuer (hand_back());
//   2. only then: * mult + shift -> activation -> residuals -> post_scale -> GroupNorm partial sums -> 16-byte stores.
// The order matters: the issuer needs the slot back within the time the tensor core spends on the other half tile, and every
// load/store of step 2 queues behind the tensor core's operand fetches on the saturated l1tex data pipe (measured: ~1900
// cycles per 16 channels, profiles/README.md) -- with the stores on the critical path the issuer waited a third of the time.
//
// DUAL (conv3d_ring2.cu, precision 3xf16r2d): the slot is a PAIR of accumulators -- the large products x_hi w_hi at t0, the small
// ones (x_hi w_lo + x_lo w_hi) SMALL_OFF columns further -- added here in fp32 round-to-nearest; only the large one carries a
// truncation bias worth compensating (`comp` = n_mma * kTruncBiasPerMma; the single-accumulator kernels fold it into `mult`).
// B0 / BSTEP: this thread drains the 16-channel blocks B0, B0 + BSTEP, ... of the slot (two warps can share one M tile); the
// extra 33rd channel belongs to the thread with B0 == 0.
template <int COUT, bool DUAL, int SMALL_OFF, int B0, int BSTEP, class HandBack>
__device__ __forceinline__ void ring_drain_slot(const ConvEpilogue& ep, const float* s_shift, float mult, float comp, uint32_t t0, bool ok,
                                                size_t pos, size_t vox, bool want_gn, double (&gs)[2], double (&gq)[2],
                                                HandBack&& hand_back) {
    constexpr int NB_ALL = COUT / 16;
    constexpr int NB = (NB_ALL - B0 + BSTEP - 1) / BSTEP;          // blocks of this thread
    constexpr bool XTRA = (COUT % 16) != 0 && B0 == 0;
    static_assert(COUT % 16 <= 1, "a slot holds a multiple of 16 channels, plus at most one");
    float acc[NB > 0 ? NB : 1][16];
    float sm[(DUAL && NB > 0) ? NB : 1][16];
    float accx = 0.0f, smx = 0.0f;
    // every TMEM load of the slot is issued before the one wait: the hand-back below is on the MMA issuer's critical path
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        tmem_ld16(t0 + (uint32_t)(16 * (B0 + i * BSTEP)), acc[i]);
        if constexpr (DUAL) tmem_ld16(t0 + (uint32_t)(SMALL_OFF + 16 * (B0 + i * BSTEP)), sm[i]);
    }
    if constexpr (XTRA) {
        accx = tmem_ld1(t0 + (uint32_t)(16 * NB_ALL));
        if constexpr (DUAL) smx = tmem_ld1(t0 + (uint32_t)(SMALL_OFF + 16 * NB_ALL));
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        tmem_st16_zero(t0 + (uint32_t)(16 * (B0 + i * BSTEP)));
        if constexpr (DUAL) tmem_st16_zero(t0 + (uint32_t)(SMALL_OFF + 16 * (B0 + i * BSTEP)));
    }
    if constexpr (XTRA) {
        tmem_st1_zero(t0 + (uint32_t)(16 * NB_ALL));
        if constexpr (DUAL) tmem_st1_zero(t0 + (uint32_t)(SMALL_OFF + 16 * NB_ALL));
    }
    tmem_st_wait();
    tc_fence_before();
    hand_back();

    if constexpr (DUAL) {
        // large-product accumulator: undo its truncation bias, then add the small products (held apart so that their 2 x n adds
        // do not truncate at the large sum's magnitude)
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[i][k] = fmaf(acc[i][k], comp, acc[i][k]) + sm[i][k];
        if constexpr (XTRA) accx = fmaf(accx, comp, accx) + smx;
    }
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int c0 = 16 * (B0 + i * BSTEP);
        float (&v)[16] = acc[i];                                   // finished in place: the kernel is at the register limit of its 640 threads
        // what the block needs from memory is requested in batches (offsets + first residual, then the second residual)
        float4 sh[4], r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[j] = *reinterpret_cast<const float4*>(s_shift + c0 + 4 * j);
        if (ep.res0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = (c0 >> 2) + j;
                r[j] = (ok && ch < ep.out_chunks) ? ldg4(ep.res0 + ((size_t)ch * vox + pos) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (ep.res_split) {                                    // vol4s: chunks (0,1) = hi / lo of channels c0..c0+7, (2,3) of c0+8..c0+15
                float t[16];
                join8(r[0], r[1], t); join8(r[2], r[3], t + 8);
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            const int act = (c < ep.act_split) ? ep.act_lo : ep.act_hi;
            v[4 * j + 0] = fmaf(v[4 * j + 0], mult, sh[j].x); v[4 * j + 1] = fmaf(v[4 * j + 1], mult, sh[j].y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], mult, sh[j].z); v[4 * j + 3] = fmaf(v[4 * j + 3], mult, sh[j].w);
            if (act == ESTD_ACT_RELU) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[4 * j + k] = fmaxf(v[4 * j + k], 0.0f);
            } else if (act == ESTD_ACT_TANH) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[4 * j + k] = tanhf(v[4 * j + k]);
            }
            v[4 * j + 0] += r[j].x; v[4 * j + 1] += r[j].y; v[4 * j + 2] += r[j].z; v[4 * j + 3] += r[j].w;
        }
        if (ep.res1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = (c0 >> 2) + j;
                r[j] = (ok && ch < ep.out_chunks) ? ldg4(ep.res1 + ((size_t)ch * vox + pos) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (ep.res_split) {
                float t[16];
                join8(r[0], r[1], t); join8(r[2], r[3], t + 8);
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[4 * j + 0] += r[j].x; v[4 * j + 1] += r[j].y; v[4 * j + 2] += r[j].z; v[4 * j + 3] += r[j].w; }
        }
        float ts[2] = {0.f, 0.f}, tq[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            const int ch = c >> 2;
            const int grp = (c < ep.act_split) ? 0 : 1;
            float s4 = 0.f, q4 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[4 * j + k] *= ep.post_scale;
                s4 += v[4 * j + k];
                q4 = fmaf(v[4 * j + k], v[4 * j + k], q4);
            }
            if (!ok || ch >= ep.out_chunks) continue;
            if (grp == 0) { ts[0] += s4; tq[0] += q4; } else { ts[1] += s4; tq[1] += q4; }
            if (ep.out_split || !ep.out0) continue;
            const size_t off = ((size_t)ch * vox + pos) * 4;
            float* dst = (ch < ep.out0_chunks) ? ep.out0 + off : ep.out1 + (off - (size_t)ep.out0_chunks * vox * 4);
            st4(dst, make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
        if (ep.out_split) {
            // vol4s: x_hi of channels c0..c0+7 -> chunk c0/4, x_lo -> chunk c0/4 + 1; c0+8..c0+15 -> the next two chunks
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int ch = (c0 >> 2) + 2 * g;
                if (!ok || ch >= ep.out_chunks) continue;
                uint4 hi, lo;
                split8(v + 8 * g, hi, lo, amax);
                float* dst = ep.out0 + ((size_t)ch * vox + pos) * 4;
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + vox * 4) = lo;
            }
        }
        if constexpr (COUT == 16) {
            if (ep.head_out && ok) {                               // fused 1x1x1 logit head over the 16 finished channels
                float logit = __ldg(ep.head_b);
#pragma unroll
                for (int k = 0; k < 16; ++k) logit = fmaf(__ldg(ep.head_w + k), v[k], logit);
                ep.head_out[pos] = logit;
            }
        }
        if (want_gn) {
            gs[0] += (double)ts[0]; gq[0] += (double)tq[0];
            gs[1] += (double)ts[1]; gq[1] += (double)tq[1];
        }
    }
    if constexpr (XTRA) {
        // the one channel beyond the 16-channel blocks: c = 16 * NB, first element of chunk c / 4 (the rest of that chunk is padding)
        constexpr int c = 16 * NB_ALL;
        const int ch = c >> 2;
        const int act = (c < ep.act_split) ? ep.act_lo : ep.act_hi;
        float v = fmaf(accx, mult, s_shift[c]);
        if (act == ESTD_ACT_RELU) v = fmaxf(v, 0.0f); else if (act == ESTD_ACT_TANH) v = tanhf(v);
        if (ok && ch < ep.out_chunks) {
            const size_t off = ((size_t)ch * vox + pos) * 4;
            if (ep.res0) {
                if (ep.res_split) { float t[8]; join8(ldg4(ep.res0 + off), ldg4(ep.res0 + off + vox * 4), t); v += t[0]; }
                else v += __ldg(ep.res0 + off);
            }
            if (ep.res1) {
                if (ep.res_split) { float t[8]; join8(ldg4(ep.res1 + off), ldg4(ep.res1 + off + vox * 4), t); v += t[0]; }
                else v += __ldg(ep.res1 + off);
            }
            v *= ep.post_scale;
            if (want_gn) { const int grp = (c < ep.act_split) ? 0 : 1; gs[grp] += (double)v; gq[grp] += (double)v * (double)v; }
            if (ep.out_split) {
                const float v8[8] = {v, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                uint4 hi, lo;
                split8(v8, hi, lo, amax);
                *reinterpret_cast<uint4*>(ep.out0 + off) = hi;
                *reinterpret_cast<uint4*>(ep.out0 + off + vox * 4) = lo;
            } else if (ep.out0) {
                float* dst = (ch < ep.out0_chunks) ? ep.out0 + off : ep.out1 + (off - (size_t)ep.out0_chunks * vox * 4);
                st4(dst, make_float4(v, 0.f, 0.f, 0.f));
            }
        }
    }
    if (ep.out_split && !(amax <= 65504.0f) && ep.status) atomicOr(ep.status, 1);
}

}  // namespace tc
}  // namespace estd
