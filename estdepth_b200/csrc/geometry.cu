// Device-side camera algebra: one tiny kernel per (target, source) pair replaces the reference's chains of
// torch.inverse / matmul (each a cuSOLVER/cuBLAS launch plus a host-syncing error check; 60 per window,
// SURVEY.md section 2.4).  Inputs are fp32 device tensors; the algebra runs in fp64 and is rounded once.
#include "common.cuh"

namespace estd {

template <int N>
__device__ void invert(const double* a, double* inv) {
    double m[N][2 * N];
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) { m[r][c] = a[r * N + c]; m[r][N + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < N; ++col) {
        int piv = col;
        double best = fabs(m[col][col]);
        for (int r = col + 1; r < N; ++r) if (fabs(m[r][col]) > best) { best = fabs(m[r][col]); piv = r; }
        if (piv != col) for (int c = 0; c < 2 * N; ++c) { double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
        double d = 1.0 / m[col][col];
        for (int c = 0; c < 2 * N; ++c) m[col][c] *= d;
        for (int r = 0; r < N; ++r) {
            if (r == col) continue;
            double f = m[r][col];
            for (int c = 0; c < 2 * N; ++c) m[r][c] -= f * m[col][c];
        }
    }
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) inv[r * N + c] = m[r][N + c];
}

__device__ void matmul4(const double* a, const double* b, double* c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = s;
        }
}

// proj = [K * E[:3,:4] ; E[3,:]]   (model_hybrid.py:83-88)
__device__ void projection(const double* K, const double* E, double* P) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * E[k * 4 + j];
            P[i * 4 + j] = s;
        }
    for (int j = 0; j < 4; ++j) P[12 + j] = E[12 + j];
}

__device__ void homography_pair(const float* __restrict__ ref_pose, const float* __restrict__ src_pose,
                                const float* __restrict__ cam_intr, float* __restrict__ out12) {
    double Pr[16], Ps[16], K[9], Er[16], Es[16], R[16], S[16], Rinv[16], M[16];
    for (int i = 0; i < 16; ++i) { Pr[i] = ref_pose[i]; Ps[i] = src_pose[i]; }
    for (int i = 0; i < 9; ++i) K[i] = cam_intr[i];
    invert<4>(Pr, Er);
    invert<4>(Ps, Es);
    projection(K, Er, R);
    projection(K, Es, S);
    invert<4>(R, Rinv);
    matmul4(S, Rinv, M);                                     // homo_utils.py:469
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out12[i * 3 + j] = (float)M[i * 4 + j];
        out12[9 + i] = (float)M[i * 4 + 3];
    }
}

__global__ void homography_setup_kernel(const float* __restrict__ ref_pose, const float* __restrict__ src_pose,
                                        const float* __restrict__ cam_intr, float* __restrict__ out12) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    homography_pair(ref_pose, src_pose, cam_intr, out12);
}

// every (reference view, source view) pair of a window in ONE launch: thread i -> row i of the [n][12] table
struct HomographyBatch { int n; int ref[ESTD_MAX_GEOMETRY_PAIRS]; int src[ESTD_MAX_GEOMETRY_PAIRS]; };
__global__ void homography_table_kernel(const float* __restrict__ poses, const float* __restrict__ cam_intr,
                                        const HomographyBatch b, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.n) homography_pair(poses + 16 * b.ref[i], poses + 16 * b.src[i], cam_intr, out + 12 * i);
}

__global__ void homography_from_proj_kernel(const float* __restrict__ src_proj, const float* __restrict__ ref_proj,
                                            float* __restrict__ out12) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double S[16], R[16], Rinv[16], M[16];
    for (int i = 0; i < 16; ++i) { S[i] = src_proj[i]; R[i] = ref_proj[i]; }
    invert<4>(R, Rinv);
    matmul4(S, Rinv, M);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out12[i * 3 + j] = (float)M[i * 4 + j];
        out12[9 + i] = (float)M[i * 4 + 3];
    }
}

__device__ void volume_warp_pair(const float* __restrict__ pose_i, const float* __restrict__ pose_j,
                                 const float* __restrict__ cam_intr, float* __restrict__ out30) {
    double Pi[16], Pj[16], K[9], Kinv[9], Piinv[16], rel[16], Minv[16];
    for (int i = 0; i < 16; ++i) { Pi[i] = pose_i[i]; Pj[i] = pose_j[i]; }
    for (int i = 0; i < 9; ++i) K[i] = cam_intr[i];
    invert<3>(K, Kinv);                                      // homo_utils.py:51
    invert<4>(Pi, Piinv);
    matmul4(Pj, Piinv, rel);                                 // hybrid_depth_decoder.py:235 (quirk Q7)
    invert<4>(rel, Minv);                                    // homo_utils.py:258
    for (int i = 0; i < 9; ++i) out30[i] = (float)Kinv[i];
    for (int i = 0; i < 12; ++i) out30[9 + i] = (float)Minv[i];
    for (int i = 0; i < 9; ++i) out30[21 + i] = (float)K[i];
}

__global__ void volume_warp_setup_kernel(const float* __restrict__ pose_i, const float* __restrict__ pose_j,
                                         const float* __restrict__ cam_intr, float* __restrict__ out30) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    volume_warp_pair(pose_i, pose_j, cam_intr, out30);
}

// every (target, source) pair of an EST fusion step in ONE launch; the poses live in separate tensors (the window's targets,
// the memory's poses), so their addresses travel by value
struct VolumeWarpBatch { int n; const float* pose[ESTD_MAX_GEOMETRY_POSES]; int target[ESTD_MAX_GEOMETRY_PAIRS]; int source[ESTD_MAX_GEOMETRY_PAIRS]; };
__global__ void volume_warp_table_kernel(const float* __restrict__ cam_intr, const VolumeWarpBatch b, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.n) volume_warp_pair(b.pose[b.target[i]], b.pose[b.source[i]], cam_intr, out + 30 * i);
}

}  // namespace estd

extern "C" int estd_homography_table(const float* poses, int n_views, const float* cam_intr, const int* pairs, int n_pairs,
                                     float* out, void* stream) {
    ESTD_REQUIRE(poses && cam_intr && pairs && out, "estd_homography_table: null pointer");
    ESTD_REQUIRE(n_pairs >= 1 && n_pairs <= ESTD_MAX_GEOMETRY_PAIRS, "estd_homography_table: 1..%d pairs, got %d", ESTD_MAX_GEOMETRY_PAIRS, n_pairs);
    estd::HomographyBatch b;
    b.n = n_pairs;
    for (int i = 0; i < n_pairs; ++i) {
        b.ref[i] = pairs[2 * i]; b.src[i] = pairs[2 * i + 1];
        ESTD_REQUIRE(b.ref[i] >= 0 && b.ref[i] < n_views && b.src[i] >= 0 && b.src[i] < n_views, "estd_homography_table: view index out of range");
    }
    estd::homography_table_kernel<<<1, ESTD_MAX_GEOMETRY_PAIRS, 0, (cudaStream_t)stream>>>(poses, cam_intr, b, out);
    return estd::check_launch("estd_homography_table");
}

extern "C" int estd_volume_warp_table(const float* const* pose_ptrs, int n_poses, const float* cam_intr, const int* pairs,
                                      int n_pairs, float* out, void* stream) {
    ESTD_REQUIRE(pose_ptrs && cam_intr && pairs && out, "estd_volume_warp_table: null pointer");
    ESTD_REQUIRE(n_poses >= 1 && n_poses <= ESTD_MAX_GEOMETRY_POSES, "estd_volume_warp_table: 1..%d poses, got %d", ESTD_MAX_GEOMETRY_POSES, n_poses);
    ESTD_REQUIRE(n_pairs >= 1 && n_pairs <= ESTD_MAX_GEOMETRY_PAIRS, "estd_volume_warp_table: 1..%d pairs, got %d", ESTD_MAX_GEOMETRY_PAIRS, n_pairs);
    estd::VolumeWarpBatch b;
    b.n = n_pairs;
    for (int i = 0; i < n_poses; ++i) { ESTD_REQUIRE(pose_ptrs[i], "estd_volume_warp_table: null pose"); b.pose[i] = pose_ptrs[i]; }
    for (int i = 0; i < n_pairs; ++i) {
        b.target[i] = pairs[2 * i]; b.source[i] = pairs[2 * i + 1];
        ESTD_REQUIRE(b.target[i] >= 0 && b.target[i] < n_poses && b.source[i] >= 0 && b.source[i] < n_poses, "estd_volume_warp_table: pose index out of range");
    }
    estd::volume_warp_table_kernel<<<1, ESTD_MAX_GEOMETRY_PAIRS, 0, (cudaStream_t)stream>>>(cam_intr, b, out);
    return estd::check_launch("estd_volume_warp_table");
}

extern "C" int estd_homography_setup(const float* ref_pose, const float* src_pose, const float* cam_intr,
                                     float* out12, void* stream) {
    ESTD_REQUIRE(ref_pose && src_pose && cam_intr && out12, "estd_homography_setup: null pointer");
    estd::homography_setup_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ref_pose, src_pose, cam_intr, out12);
    return estd::check_launch("estd_homography_setup");
}

extern "C" int estd_homography_from_proj(const float* src_proj, const float* ref_proj, float* out12, void* stream) {
    ESTD_REQUIRE(src_proj && ref_proj && out12, "estd_homography_from_proj: null pointer");
    estd::homography_from_proj_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(src_proj, ref_proj, out12);
    return estd::check_launch("estd_homography_from_proj");
}

extern "C" int estd_volume_warp_setup(const float* pose_i, const float* pose_j, const float* cam_intr,
                                      float* out30, void* stream) {
    ESTD_REQUIRE(pose_i && pose_j && cam_intr && out30, "estd_volume_warp_setup: null pointer");
    estd::volume_warp_setup_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pose_i, pose_j, cam_intr, out30);
    return estd::check_launch("estd_volume_warp_setup");
}
