// Evaluation metrics next to the hot path: cal_mpjpe of evaluate.py:120-133 with the Procrustes alignment
// of utils.py:138-203, fused so that neither of the two 6890-vertex meshes is ever written to HBM:
//   rc_smpl_chain (kinematics.cu)  skinning transforms of the predicted and the ground-truth pose
//   rc_metrics_mesh_kernel         block per frame: skin every vertex of both meshes, accumulate the vertex distance (PVE) and the
//                                  regressed joints J_regressor[:14] @ vertices of both, deterministic block reduction
//   rc_metrics_joint_kernel        thread per frame: pelvis alignment, MPJPE, optional Procrustes (PA-MPJPE)
// HBM-bound on reading the skinning weights / regressor (0.66 MB + 0.39 MB per frame block, L2 resident): ~4 MFLOP per frame.
#include "rc_common.cuh"
#include "rc_rows.h"
#include "rc_model.cuh"

int rc_smpl_chain_launch(const rc_model* m, const float* pose, long long b, float* Rg, float* joint, float* Tskin, void* stream);

namespace {

constexpr int kMJ = 14;                 // joints used by cal_mpjpe (evaluate.py:123)
constexpr int kMT = 256;

__global__ void __launch_bounds__(kMT) rc_metrics_mesh_kernel(const float* __restrict__ Tp, const float* __restrict__ Tg,
                                                               const float* __restrict__ vrest, const float* __restrict__ W,
                                                               const float* __restrict__ jreg, int nv, float* __restrict__ pve,
                                                               float* __restrict__ kp) {
    __shared__ float sT[2][288];
    __shared__ float red[kMT / 32][2 * kMJ * 3 + 1];
    const long long f = blockIdx.x;
    for (int e = threadIdx.x; e < 288; e += kMT) { sT[0][e] = Tp[f * 288 + e]; sT[1][e] = Tg[f * 288 + e]; }
    __syncthreads();
    float acc[2 * kMJ * 3 + 1];
#pragma unroll
    for (int e = 0; e < 2 * kMJ * 3 + 1; ++e) acc[e] = 0.f;
    for (int v = threadIdx.x; v < nv; v += kMT) {
        float w[RC_NJ];
        const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)v * RC_NJ);
#pragma unroll
        for (int q = 0; q < 6; ++q) { const float4 t = __ldg(w4 + q); w[q * 4] = t.x; w[q * 4 + 1] = t.y; w[q * 4 + 2] = t.z; w[q * 4 + 3] = t.w; }
        const float x = vrest[v * 3], y = vrest[v * 3 + 1], z = vrest[v * 3 + 2];
        float p[2][3];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float T[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
            for (int j = 0; j < RC_NJ; ++j) {
#pragma unroll
                for (int e = 0; e < 12; ++e) T[e] = fmaf(w[j], sT[s][j * 12 + e], T[e]);
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) p[s][r] = T[r * 4] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3];
        }
        const float dx = p[1][0] - p[0][0], dy = p[1][1] - p[0][1], dz = p[1][2] - p[0][2];
        acc[2 * kMJ * 3] += sqrtf(dx * dx + dy * dy + dz * dz);
#pragma unroll
        for (int jr = 0; jr < kMJ; ++jr) {
            const float c = __ldg(jreg + (size_t)jr * nv + v);
            if (c != 0.f) {
#pragma unroll
                for (int r = 0; r < 3; ++r) { acc[jr * 3 + r] = fmaf(c, p[0][r], acc[jr * 3 + r]); acc[(kMJ + jr) * 3 + r] = fmaf(c, p[1][r], acc[(kMJ + jr) * 3 + r]); }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < 2 * kMJ * 3 + 1; ++e) {
        float vsum = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if (lane == 0) red[warp][e] = vsum;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * kMJ * 3 + 1; e += kMT) {
        float s = 0.f;
        for (int q = 0; q < kMT / 32; ++q) s += red[q][e];
        if (e == 2 * kMJ * 3) pve[f * 3] = s / (float)nv;          // column 1 of the [b,3] result
        else kp[f * (2 * kMJ * 3) + e] = s;
    }
}

// Horn's closed form: the optimal proper rotation is the dominant eigenvector (as a quaternion) of a symmetric 4x4 matrix built
// from the cross-covariance; equivalent to utils.py:160-170 (SVD with the det(R) = +1 fix-up).  Jacobi sweeps in double.
__device__ void best_rotation(const double K[9], double R[9]) {      // K = X1 X2^T ; R maximises trace(R K)
    // M = K^T here in Horn's notation Sxy = sum x1 y2
    const double Sxx = K[0], Sxy = K[1], Sxz = K[2], Syx = K[3], Syy = K[4], Syz = K[5], Szx = K[6], Szy = K[7], Szz = K[8];
    double N[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                      {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                      {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                      {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += N[p][q] * N[p][q];
        if (off < 1e-30) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (fabs(N[p][q]) < 1e-300) continue;
                const double th = (N[q][q] - N[p][p]) / (2.0 * N[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) { const double a = N[k][p], b = N[k][q]; N[k][p] = c * a - s * b; N[k][q] = s * a + c * b; }
                for (int k = 0; k < 4; ++k) { const double a = N[p][k], b = N[q][k]; N[p][k] = c * a - s * b; N[q][k] = s * a + c * b; }
                for (int k = 0; k < 4; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
            }
    }
    int best = 0;
    for (int k = 1; k < 4; ++k) if (N[k][k] > N[best][best]) best = k;
    const double w = V[0][best], x = V[1][best], y = V[2][best], z = V[3][best];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z);           R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);           R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);           R[7] = 2 * (y * z + w * x);           R[8] = w * w - x * x - y * y + z * z;
}

__global__ void __launch_bounds__(64) rc_metrics_joint_kernel(const float* __restrict__ kp, long long b, int with_pa, float* __restrict__ out) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= b) return;
    float P[kMJ][3], G[kMJ][3];
    for (int j = 0; j < kMJ; ++j)
        for (int r = 0; r < 3; ++r) { P[j][r] = kp[f * (2 * kMJ * 3) + j * 3 + r]; G[j][r] = kp[f * (2 * kMJ * 3) + (kMJ + j) * 3 + r]; }
    for (int j = kMJ - 1; j >= 0; --j)
        for (int r = 0; r < 3; ++r) { P[j][r] -= P[0][r]; G[j][r] -= G[0][r]; }      // pelvis alignment (evaluate.py:125-128)
    float e = 0.f;
    for (int j = 0; j < kMJ; ++j) {
        const float dx = G[j][0] - P[j][0], dy = G[j][1] - P[j][1], dz = G[j][2] - P[j][2];
        e += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    out[f * 3] = e / kMJ;
    if (!with_pa) { out[f * 3 + 2] = 0.f; return; }
    // utils.py:138-186: similarity transform of the prediction closest to the ground truth
    double m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
    for (int j = 0; j < kMJ; ++j) for (int r = 0; r < 3; ++r) { m1[r] += P[j][r]; m2[r] += G[j][r]; }
    for (int r = 0; r < 3; ++r) { m1[r] /= kMJ; m2[r] /= kMJ; }
    double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, var1 = 0;
    for (int j = 0; j < kMJ; ++j)
        for (int r = 0; r < 3; ++r) {
            const double a = P[j][r] - m1[r];
            var1 += a * a;
            for (int c = 0; c < 3; ++c) K[r * 3 + c] += a * (G[j][c] - m2[c]);
        }
    double R[9];
    best_rotation(K, R);
    // R maps X1 -> X2 : y = s R x + t ; trace(R K) with K = X1 X2^T
    double tr = 0;
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) tr += R[i * 3 + k] * K[k * 3 + i];
    const double s = tr / var1;
    double err = 0;
    for (int j = 0; j < kMJ; ++j) {
        double d2 = 0;
        for (int i = 0; i < 3; ++i) {
            double yv = m2[i];
            for (int k = 0; k < 3; ++k) yv += s * R[i * 3 + k] * (P[j][k] - m1[k]);
            const double d = yv - G[j][i];
            d2 += d * d;
        }
        err += sqrt(d2);
    }
    out[f * 3 + 2] = (float)(err / kMJ);
}

}  // namespace

extern "C" int rc_metrics_mpjpe(const rc_model* m, const float* d_jreg, int32_t nj_rows, const float* d_pose, const float* d_gt_pose,
                                int64_t b, int with_pa, float* d_out, void* stream) {
    RC_ARG(m && d_jreg && nj_rows >= kMJ && d_pose && d_gt_pose && d_out && b >= 0);
    if (b == 0) return RC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* scratch = nullptr;
    // [2][b,288] skinning transforms + [b,216] + [b,72] throw-away chain outputs + [b, 84] regressed joints
    const size_t n = (size_t)b * (2 * 288 + 216 + 72 + 2 * kMJ * 3);
    RC_CUDA(cudaMallocAsync(&scratch, n * sizeof(float), st));
    float* Tp = scratch; float* Tg = Tp + b * 288; float* Rg = Tg + b * 288; float* jt = Rg + b * 216; float* kp = jt + b * 72;
    int rc = rc_smpl_chain_launch(m, d_pose, b, Rg, jt, Tp, stream);
    if (rc == RC_OK) rc = rc_smpl_chain_launch(m, d_gt_pose, b, Rg, jt, Tg, stream);
    if (rc != RC_OK) { cudaFreeAsync(scratch, st); return rc; }
    RC_LAUNCH(rc_metrics_mesh_kernel, (unsigned)b, kMT, 0, stream, Tp, Tg, m->d_verts, m->d_skin_w, d_jreg, m->nv, d_out + 1, kp);
    RC_CHECK_LAUNCH();
    (void)nj_rows;
    RC_LAUNCH(rc_metrics_joint_kernel, rc_cdiv(b, 64), 64, 0, stream, kp, (long long)b, with_pa, d_out);
    RC_CHECK_LAUNCH();
    RC_CUDA(cudaFreeAsync(scratch, st));
    return RC_OK;
}
