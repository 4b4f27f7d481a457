// Pieces shared by the two implementations of K2 (conv3d.cu: exact fp32 SIMT; conv3d_tc.cu: 3xTF32 on tcgen05):
// mbarrier / TMA PTX wrappers, the vol4 tensor-map builder and the fused epilogue.
#pragma once
#include <cstdlib>
#include <utility>
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace estd {

// ---------------------------------------------------------------- TMA / mbarrier PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a TMA / MMA that never completes (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// Same bound without reading the clock in the loop, with a pause between polls: every poll is a shared-memory wavefront
// and the ring kernel's waiting warps were taking a fifth of the shared-memory pipe from the tensor core's operand fetch
// (profiles/README.md).  Roles with several warps let ONE warp poll and release the others through a named barrier.
__device__ __forceinline__ void mbar_wait_polls(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; !mbar_try_wait(bar, parity); ++i) {
        __nanosleep(40);
        if (i > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---------------------------------------------------------------- fused epilogue (shared semantics)
struct ConvEpilogue {
    const float* scale; const float* shift;
    const float* res0; const float* res1;
    float* out0; float* out1;
    double* gn_partials;
    int out0_chunks, out_chunks;
    int act_split, act_lo, act_hi;
    float post_scale;
    int res_split, out_split;                   // vol4s residuals / output (include/estdepth_b200.h)
    int* status;                                // fp16 range flag (out_split)
    const float* head_w; const float* head_b; float* head_out;      // fused 1x1x1 logit head
    int out_up2;                                                    // planar: nearest x2 up-sampled output
};

inline void fill_epilogue(ConvEpilogue* e, const estd_conv3d_desc* d) {
    e->scale = d->scale; e->shift = d->shift;
    e->res0 = d->res0; e->res1 = d->res1;
    e->out0 = d->out0; e->out1 = d->out1;
    e->gn_partials = d->gn_partials;
    e->out0_chunks = d->out0_chunks; e->out_chunks = d->out0_chunks + d->out1_chunks;
    e->act_split = d->act_split; e->act_lo = d->act_lo; e->act_hi = d->act_hi;
    e->post_scale = d->post_scale;
    e->res_split = d->res_split; e->out_split = d->out_split;
    e->status = d->status;
    e->head_w = d->head_w; e->head_b = d->head_b; e->head_out = d->head_out;
    e->out_up2 = d->out_up2;
}

// ---- vol4s helpers: 8 channels of one voxel = 16 B of x_hi (8 x fp16) in chunk 2g + 16 B of x_lo in chunk 2g+1 ----
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo, float& amax) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
        const __half2 hh = __floats2half2_rn(a, b);
        const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a - f.x, b - f.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// x_hi + x_lo of 8 channels (exact in fp32: 22 significant bits)
__device__ __forceinline__ void join8(const float4& hi4, const float4& lo4, float* out) {
    const uint32_t h[4] = {__float_as_uint(hi4.x), __float_as_uint(hi4.y), __float_as_uint(hi4.z), __float_as_uint(hi4.w)};
    const uint32_t l[4] = {__float_as_uint(lo4.x), __float_as_uint(lo4.y), __float_as_uint(lo4.z), __float_as_uint(lo4.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
        out[2 * i] = a.x + b.x; out[2 * i + 1] = a.y + b.y;
    }
}

// 16 consecutive output channels [c0, c0+16) of one voxel: affine -> activation -> residuals -> scale -> 4 x 16-byte stores.
// ts/tq accumulate sum / sum of squares of the stored values per GroupNorm group (0: c < act_split, 1: otherwise).
__device__ __forceinline__ void conv_epilogue_store16(const ConvEpilogue& e, const float (&acc)[16], int c0, bool ok, size_t pos,
                                                      size_t vox, float (&ts)[2], float (&tq)[2]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + 4 * j;
        const int ch = c >> 2;
        if (!ok || ch >= e.out_chunks) continue;
        const int act = (c < e.act_split) ? e.act_lo : e.act_hi;
        const int grp = (c < e.act_split) ? 0 : 1;
        const float4 sc = ldg4(e.scale + c), sh = ldg4(e.shift + c);
        float v[4];
        v[0] = apply_act(fmaf(acc[4 * j + 0], sc.x, sh.x), act);
        v[1] = apply_act(fmaf(acc[4 * j + 1], sc.y, sh.y), act);
        v[2] = apply_act(fmaf(acc[4 * j + 2], sc.z, sh.z), act);
        v[3] = apply_act(fmaf(acc[4 * j + 3], sc.w, sh.w), act);
        const size_t off = ((size_t)ch * vox + pos) * 4;
        if (e.res0) { const float4 q = ldg4(e.res0 + off); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
        if (e.res1) { const float4 q = ldg4(e.res1 + off); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] *= e.post_scale;
            ts[grp] += v[k];
            tq[grp] = fmaf(v[k], v[k], tq[grp]);
        }
        float* dst = (ch < e.out0_chunks) ? e.out0 + off : e.out1 + (off - (size_t)e.out0_chunks * vox * 4);
        st4(dst, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// ---------------------------------------------------------------- host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// vol4 tensor [chunks][D][H][W][4] seen by TMA as 4-D (x = 4W floats, H, D, chunks); box = (box_x floats, box_h, box_d, box_c).
// Out-of-bounds box elements (negative coordinates, beyond any extent incl. the chunk axis) are zero-filled.
inline int make_vol4_tensor_map(CUtensorMap* map, const float* base, int chunks, int D, int H, int W, int box_x, int box_h,
                                int box_d, int box_c) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ESTD_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)chunks};
    cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)box_x, (cuuint32_t)box_h, (cuuint32_t)box_d, (cuuint32_t)box_c};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(ESTD_ECUDA, "cuTensorMapEncodeTiled failed (%d) for vol4 [%d][%d][%d][%d][4]", (int)r, chunks, D, H, W);
    return ESTD_OK;
}

// ESTD_PDL=0 launches the tensor-core kernels without programmatic stream serialisation (default: on)
inline bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("ESTD_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

// <<<grid, threads, smem, stream>>> with the programmatic-stream-serialisation attribute (tc_ptx.cuh: pdl_wait)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int threads, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// A per-device "already done" flag (one bit per device ordinal, thread safe): function attributes such as the opt-in
// dynamic shared-memory limit belong to the (function, device) pair, so a process-wide bool would leave every device after
// the first without them.  Devices beyond ordinal 63 simply repeat the (cheap, idempotent) call.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    static int device() { int dev = 0; return cudaGetDevice(&dev) == cudaSuccess ? dev : 0; }
    bool done() const { const int d = device(); return d < 64 && ((mask.load(std::memory_order_acquire) >> d) & 1ull); }
    void set() { const int d = device(); if (d < 64) mask.fetch_or(1ull << d, std::memory_order_release); }
};

// SM count of the CURRENT device (cached per device ordinal)
inline int sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64) { const int c = cache[dev].load(std::memory_order_relaxed); if (c > 0) return c; }
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    if (dev >= 0 && dev < 64) cache[dev].store(v, std::memory_order_relaxed);
    return v;
}

// implemented in conv3d_tc.cu
int dispatch_tc(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas);
// implemented in conv3d_ring2.cu
int dispatch_ring2(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas);
// implemented in conv2d_tc.cu
int dispatch_planar(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas);
// implemented in conv3d_ring.cu
int dispatch_ring(const estd_conv3d_desc* d, cudaStream_t stream, bool count_only, int* n_ctas);

}  // namespace estd
