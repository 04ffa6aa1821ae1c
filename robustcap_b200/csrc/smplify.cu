// SMPLify objective and its analytic gradient (net/smplify/losses.py:23-91 on top of temporal_smplify.py:153-165).
//
// The reference builds an autograd graph through batch_rodrigues -> 4x4 FK chain -> linear-blend skinning of all 6890
// vertices -> 33-point gather -> loss and back-propagates it ~26 times per sequence.  Here the same scalar function of
// (axis-angle pose [T,72], translation [T,3]) is evaluated with hand-derived derivatives, skinning only the 21 vertices
// that the 33 MediaPipe points read:
//   rc_gmm_kernel          block per frame, warp per mixture component: min_m(0.5 d^T P_m d - log w_m) and P_m* d
//   rc_smplify_fwd_kernel  thread per frame: Rodrigues, FK chain, key points, projection, per-frame loss terms
//   rc_smplify_bwd_kernel  thread per frame: d loss / d points (incl. the temporal L1 terms that couple t-1, t, t+1),
//                          back through skinning, the kinematic chain and Rodrigues
//   rc_sum_kernel          deterministic reduction of the per-frame losses
// The optimiser stays torch.optim.LBFGS (the third-party class the reference itself calls).
#include <vector>
#include "rc_common.cuh"
#include "rc_rows.h"
#include "rc_model.cuh"

namespace {

constexpr int NG = 8, ND = 69;
__constant__ int c_ji_mask[6] = {18, 19, 4, 5, 15, 0};          // config.py:101
__constant__ int c_angle_idx[4] = {52, 55, 9, 12};              // losses.py:20 (indices into pose[3:])
__constant__ float c_angle_sign[4] = {1.f, -1.f, -1.f, -1.f};

struct SmplifyConst {
    float means[NG][ND];
    float logw[NG];               // log(nll_weights)
};

}  // namespace

struct rc_smplify {
    const rc_model* model = nullptr;
    int T = 0;
    SmplifyConst* d_const = nullptr;
    float* d_psym = nullptr;      // [8,69,69] 0.5 (P + P^T)
    float *d_G = nullptr, *d_p = nullptr, *d_proj = nullptr, *d_lossf = nullptr, *d_prior = nullptr, *d_gprior = nullptr;
    int* d_argmin = nullptr;
    std::vector<void*> allocs;
};

namespace {

// ---- GMM prior (prior.py:164-179) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rc_gmm_kernel(const SmplifyConst* __restrict__ C, const float* __restrict__ Psym,
                                                      const float* __restrict__ aa, float* __restrict__ prior,
                                                      float* __restrict__ gprior) {
    __shared__ float d[NG][ND];
    __shared__ float pd[NG][ND];
    __shared__ float ll[NG];
    const int t = blockIdx.x, m = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = lane; j < ND; j += 32) d[m][j] = aa[(size_t)t * 72 + 3 + j] - C->means[m][j];
    __syncwarp();
    const float* P = Psym + (size_t)m * ND * ND;
    float q = 0.f;
    for (int i = lane; i < ND; i += 32) {
        float s = 0.f;
        for (int j = 0; j < ND; ++j) s = fmaf(P[i * ND + j], d[m][j], s);
        pd[m][i] = s;
        q = fmaf(s, d[m][i], q);
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) ll[m] = 0.5f * q - C->logw[m];
    __syncthreads();
    int best = 0;
    for (int k = 1; k < NG; ++k) if (ll[k] < ll[best]) best = k;          // torch.min: first minimum
    if (threadIdx.x == 0) prior[t] = ll[best];
    for (int j = threadIdx.x; j < ND; j += blockDim.x) gprior[(size_t)t * ND + j] = pd[best][j];
}

// ---- forward ---------------------------------------------------------------------------------------------------------
struct FwdArgs {
    const float *aa, *tran, *j2d, *conf, *camk, *ref3d, *imu_aa, *prior;
    float *G, *p, *proj, *lossf, *reproj;
    int T, rodrigues;     // 0: batch_rodrigues (temporal_smplify.py:25-59), 1: axis_angle_to_rotation_matrix (angular.py:221-233),
                          // 2: the pose pointer holds [T,24,3,3] rotation matrices (get_fitting_loss, value only)
};

__device__ __forceinline__ void keypoints_from_G(const RcModelConst& M, const float (*G)[12], const float* tran, float* p) {
    for (int k = 0; k < RC_NKP; ++k) {
        float o[3];
        if (M.kp_is_joint[k]) {
            const int j = M.kp_index[k];
            for (int r = 0; r < 3; ++r) o[r] = G[j][r * 4 + 3];
        } else {
            float Tv[12];
            for (int e = 0; e < 12; ++e) Tv[e] = 0.f;
            for (int j = 0; j < RC_NJ; ++j) {
                const float w = M.kp_w[k][j];
                if (w != 0.f) {
                    // skinning transform [R_j | t_j - R_j jrest_j]  (model.py:235)
                    const float tx = G[j][3] - (G[j][0] * M.jrest[j][0] + G[j][1] * M.jrest[j][1] + G[j][2] * M.jrest[j][2]);
                    const float ty = G[j][7] - (G[j][4] * M.jrest[j][0] + G[j][5] * M.jrest[j][1] + G[j][6] * M.jrest[j][2]);
                    const float tz = G[j][11] - (G[j][8] * M.jrest[j][0] + G[j][9] * M.jrest[j][1] + G[j][10] * M.jrest[j][2]);
                    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Tv[r * 4 + c] = fmaf(w, G[j][r * 4 + c], Tv[r * 4 + c]);
                    Tv[3] = fmaf(w, tx, Tv[3]); Tv[7] = fmaf(w, ty, Tv[7]); Tv[11] = fmaf(w, tz, Tv[11]);
                }
            }
            for (int r = 0; r < 3; ++r)
                o[r] = Tv[r * 4] * M.kp_rest[k][0] + Tv[r * 4 + 1] * M.kp_rest[k][1] + Tv[r * 4 + 2] * M.kp_rest[k][2] + Tv[r * 4 + 3];
        }
        for (int r = 0; r < 3; ++r) p[k * 3 + r] = o[r] + tran[r];
    }
}

__global__ void __launch_bounds__(32) rc_smplify_fwd_kernel(const RcModelConst* __restrict__ Mp, FwdArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.T) return;
    const RcModelConst& M = *Mp;
    float G[RC_NJ][12];
    for (int i = 0; i < RC_NJ; ++i) {
        float R[9], L[12];
        if (a.rodrigues == 0) rc_batch_rodrigues(a.aa + (size_t)t * 72 + i * 3, R);
        else if (a.rodrigues == 1) rc_aa_to_mat(a.aa + (size_t)t * 72 + i * 3, R);
        else { for (int e = 0; e < 9; ++e) R[e] = a.aa[(size_t)t * 216 + i * 9 + e]; }   // mode 2: `aa` holds rotation matrices
        for (int r = 0; r < 3; ++r) { L[r * 4] = R[r * 3]; L[r * 4 + 1] = R[r * 3 + 1]; L[r * 4 + 2] = R[r * 3 + 2]; L[r * 4 + 3] = M.bone[i][r]; }
        if (i == 0) for (int e = 0; e < 12; ++e) G[0][e] = L[e];
        else rc_rigid_mul(G[M.parent[i]], L, G[i]);
    }
    float tr[3] = {a.tran[t * 3], a.tran[t * 3 + 1], a.tran[t * 3 + 2]};
    float p[99];
    keypoints_from_G(M, G, tr, p);
    float* Gs = a.G + (size_t)t * 288;
    for (int i = 0; i < RC_NJ; ++i) for (int e = 0; e < 12; ++e) Gs[i * 12 + e] = G[i][e];
    float K[9];
    for (int e = 0; e < 9; ++e) K[e] = a.camk[e];
    float loss = 0.f;
    // re-projection (losses.py:36-46): conf^2 * sum_xy gmof(K (p / p_z) - j2d, 100)
    for (int k = 0; k < RC_NKP; ++k) {
        const float x = p[k * 3] / p[k * 3 + 2], y = p[k * 3 + 1] / p[k * 3 + 2], z = p[k * 3 + 2] / p[k * 3 + 2];
        const float u = K[0] * x + K[1] * y + K[2] * z, v = K[3] * x + K[4] * y + K[5] * z;
        a.p[(size_t)t * 99 + k * 3] = p[k * 3]; a.p[(size_t)t * 99 + k * 3 + 1] = p[k * 3 + 1]; a.p[(size_t)t * 99 + k * 3 + 2] = p[k * 3 + 2];
        a.proj[(size_t)t * 66 + k * 2] = u; a.proj[(size_t)t * 66 + k * 2 + 1] = v;
        const float c = a.conf[(size_t)t * 33 + k];
        const float ru = u - a.j2d[(size_t)t * 66 + k * 2], rv = v - a.j2d[(size_t)t * 66 + k * 2 + 1];
        const float e = (1e4f * ru * ru) / (1e4f + ru * ru) + (1e4f * rv * rv) / (1e4f + rv * rv);
        const float rl = c * c * e;
        if (a.reproj) a.reproj[(size_t)t * 33 + k] = rl;
        loss += rl;
    }
    // 3-D term (losses.py:32-34): sum_i |(p_i - p_0) - (ref_i - ref_0)|^2
    for (int k = 1; k < RC_NKP; ++k)
        for (int r = 0; r < 3; ++r) {
            const float d = (p[k * 3 + r] - p[r]) - (a.ref3d[(size_t)t * 99 + k * 3 + r] - a.ref3d[(size_t)t * 99 + r]);
            loss = fmaf(d, d, loss);
        }
    // GMM prior (0.1^2) and angle prior (15.2^2 * exp(+-x)^2)   (losses.py:49-54)
    if (a.rodrigues != 2) {
        loss = fmaf(0.01f, a.prior[t], loss);
        for (int q = 0; q < 4; ++q) {
            const float e = expf(a.aa[(size_t)t * 72 + 3 + c_angle_idx[q]] * c_angle_sign[q]);
            loss = fmaf(231.04f, e * e, loss);
        }
    }
    // IMU term (losses.py:39-40), value only: 0.5^2 |aa(imu_ori) - aa(R_glb[ji_mask])|^2, cv2.Rodrigues semantics.
    // losses.py:63 broadcasts the sum over frames onto every frame; the caller multiplies by T (see rc_smplify_loss_grad).
    float imu = 0.f;
    for (int s = 0; s < 6; ++s) {
        const int j = c_ji_mask[s];
        float R[9], v[3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = G[j][r * 4 + c];
        rc_mat_to_aa(R, v);
        for (int r = 0; r < 3; ++r) { const float d = a.imu_aa[(size_t)t * 18 + s * 3 + r] - v[r]; imu = fmaf(d, d, imu); }
    }
    a.lossf[t] = loss;
    a.lossf[a.T + t] = 0.25f * imu;
}

// ---- backward ----------------------------------------------------------------------------------------------------------
struct BwdArgs {
    const float *aa, *j2d, *conf, *camk, *ref3d, *G, *p, *proj, *gprior;
    float *lossf, *gaa, *gtran;
    int T;
};

__global__ void __launch_bounds__(32) rc_smplify_bwd_kernel(const RcModelConst* __restrict__ Mp, BwdArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.T) return;
    const RcModelConst& M = *Mp;
    const float* p = a.p + (size_t)t * 99;
    const float* pr = a.proj + (size_t)t * 66;
    float K[9];
    for (int e = 0; e < 9; ++e) K[e] = a.camk[e];
    float gp[99];
    float smooth = 0.f;
    float g0[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < RC_NKP; ++k) {
        const float c = a.conf[(size_t)t * 33 + k], c2 = c * c;
        // d gmof(r)/dr = 2 r s^4 / (s^2 + r^2)^2
        const float ru = pr[k * 2] - a.j2d[(size_t)t * 66 + k * 2], rv = pr[k * 2 + 1] - a.j2d[(size_t)t * 66 + k * 2 + 1];
        const float du = 1e4f + ru * ru, dv = 1e4f + rv * rv;
        float gu = c2 * 2.f * ru * 1e8f / (du * du), gv = c2 * 2.f * rv * 1e8f / (dv * dv);
        float g3[3] = {0.f, 0.f, 0.f};
        // temporal L1 terms (losses.py:66-84): frame t pulls towards t-1 with conf_t^2, frame t+1 pulls on t with conf_{t+1}^2
        if (t >= 1) {
            const float* pp = a.p + (size_t)(t - 1) * 99;
            const float* qq = a.proj + (size_t)(t - 1) * 66;
            for (int r = 0; r < 2; ++r) {
                const float d = pr[k * 2 + r] - qq[k * 2 + r];
                smooth = fmaf(1e-4f * c2, fabsf(d), smooth);
                const float s = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
                if (r == 0) gu = fmaf(1e-4f * c2, s, gu); else gv = fmaf(1e-4f * c2, s, gv);
            }
            for (int r = 0; r < 3; ++r) {
                const float d = p[k * 3 + r] - pp[k * 3 + r];
                smooth = fmaf(c2, fabsf(d), smooth);
                g3[r] += c2 * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
            }
        }
        if (t + 1 < a.T) {
            const float cn = a.conf[(size_t)(t + 1) * 33 + k], cn2 = cn * cn;
            const float* pn = a.p + (size_t)(t + 1) * 99;
            const float* qn = a.proj + (size_t)(t + 1) * 66;
            for (int r = 0; r < 2; ++r) {
                const float d = qn[k * 2 + r] - pr[k * 2 + r];
                const float s = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
                if (r == 0) gu = fmaf(-1e-4f * cn2, s, gu); else gv = fmaf(-1e-4f * cn2, s, gv);
            }
            for (int r = 0; r < 3; ++r) {
                const float d = pn[k * 3 + r] - p[k * 3 + r];
                g3[r] -= cn2 * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
            }
        }
        // (u, v) = K[:2] (x/z, y/z, z/z): d/dx = K?0 / z, d/dy = K?1 / z, d/dz = -(K?0 x + K?1 y) / z^2  (the z/z term is constant 1
        // for autograd too: d(z/z)/dz = 1/z - z/z^2 = 0)
        const float x = p[k * 3], y = p[k * 3 + 1], z = p[k * 3 + 2], iz = 1.f / z;
        const float gx = (gu * K[0] + gv * K[3]) * iz, gy = (gu * K[1] + gv * K[4]) * iz;
        const float gz = -(gu * (K[0] * x + K[1] * y) + gv * (K[3] * x + K[4] * y)) * iz * iz;
        g3[0] += gx; g3[1] += gy; g3[2] += gz;
        if (k >= 1) {
            for (int r = 0; r < 3; ++r) {
                const float d = (p[k * 3 + r] - p[r]) - (a.ref3d[(size_t)t * 99 + k * 3 + r] - a.ref3d[(size_t)t * 99 + r]);
                g3[r] = fmaf(2.f, d, g3[r]);
                g0[r] = fmaf(-2.f, d, g0[r]);
            }
        }
        gp[k * 3] = g3[0]; gp[k * 3 + 1] = g3[1]; gp[k * 3 + 2] = g3[2];
    }
    for (int r = 0; r < 3; ++r) gp[r] += g0[r];
    a.lossf[2 * a.T + t] = smooth;

    // points -> global joint transforms
    const float* G = a.G + (size_t)t * 288;
    float gR[RC_NJ][9], gt[RC_NJ][3];
    for (int j = 0; j < RC_NJ; ++j) { for (int e = 0; e < 9; ++e) gR[j][e] = 0.f; gt[j][0] = gt[j][1] = gt[j][2] = 0.f; }
    float gtr[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < RC_NKP; ++k) {
        const float* g = gp + k * 3;
        for (int r = 0; r < 3; ++r) gtr[r] += g[r];
        if (M.kp_is_joint[k]) {
            const int j = M.kp_index[k];
            for (int r = 0; r < 3; ++r) gt[j][r] += g[r];
        } else {
            for (int j = 0; j < RC_NJ; ++j) {
                const float w = M.kp_w[k][j];
                if (w != 0.f) {
                    // p += w (R_j (v - jrest_j) + t_j)
                    const float dx = M.kp_rest[k][0] - M.jrest[j][0], dy = M.kp_rest[k][1] - M.jrest[j][1], dz = M.kp_rest[k][2] - M.jrest[j][2];
                    for (int r = 0; r < 3; ++r) {
                        const float wg = w * g[r];
                        gR[j][r * 3] = fmaf(wg, dx, gR[j][r * 3]); gR[j][r * 3 + 1] = fmaf(wg, dy, gR[j][r * 3 + 1]); gR[j][r * 3 + 2] = fmaf(wg, dz, gR[j][r * 3 + 2]);
                        gt[j][r] += wg;
                    }
                }
            }
        }
    }
    // chain, leaves to root: R_j = R_p Rl_j, t_j = R_p b_j + t_p
    float gaa[72];
    for (int j = RC_NJ - 1; j >= 0; --j) {
        float gRl[9];
        float Rl[9];
        const int pj = M.parent[j];
        if (j > 0) {
            const float* Gp = G + pj * 12;
            const float* Gj = G + j * 12;
            // Rl_j = R_p^T R_j ; gRl = R_p^T gR_j ; gR_p += gR_j Rl_j^T + gt_j b_j^T ; gt_p += gt_j
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    Rl[r * 3 + c] = Gp[0 * 4 + r] * Gj[0 * 4 + c] + Gp[1 * 4 + r] * Gj[1 * 4 + c] + Gp[2 * 4 + r] * Gj[2 * 4 + c];
                    gRl[r * 3 + c] = Gp[0 * 4 + r] * gR[j][0 * 3 + c] + Gp[1 * 4 + r] * gR[j][1 * 3 + c] + Gp[2 * 4 + r] * gR[j][2 * 3 + c];
                }
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    const float v = gR[j][r * 3] * Rl[c * 3] + gR[j][r * 3 + 1] * Rl[c * 3 + 1] + gR[j][r * 3 + 2] * Rl[c * 3 + 2];
                    gR[pj][r * 3 + c] += v + gt[j][r] * M.bone[j][c];
                }
            for (int r = 0; r < 3; ++r) gt[pj][r] += gt[j][r];
        } else {
            for (int e = 0; e < 9; ++e) gRl[e] = gR[0][e];
        }
        // Rodrigues backward: R = I + sin(th) K(d) + (1 - cos(th)) K(d)^2, th = |v + 1e-8|, d = v / th
        const float* v = a.aa + (size_t)t * 72 + j * 3;
        const float e0 = v[0] + 1e-8f, e1 = v[1] + 1e-8f, e2 = v[2] + 1e-8f;
        const float th = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
        const float d0 = v[0] / th, d1 = v[1] / th, d2 = v[2] / th;
        const float sn = sinf(th), cs = cosf(th), om = 1.f - cs;
        const float Km[9] = {0.f, -d2, d1, d2, 0.f, -d0, -d1, d0, 0.f};
        float KK[9];
        rc_mat3_mul(Km, Km, KK);
        float gs = 0.f, go = 0.f;
        for (int e = 0; e < 9; ++e) { gs = fmaf(gRl[e], Km[e], gs); go = fmaf(gRl[e], KK[e], go); }
        const float gth = gs * cs + go * sn;
        // gK = sn gRl + om (gRl K^T + K^T gRl)
        float gK[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                float s1 = 0.f, s2 = 0.f;
                for (int q = 0; q < 3; ++q) { s1 = fmaf(gRl[r * 3 + q], Km[c * 3 + q], s1); s2 = fmaf(Km[q * 3 + r], gRl[q * 3 + c], s2); }
                gK[r * 3 + c] = sn * gRl[r * 3 + c] + om * (s1 + s2);
            }
        const float gd0 = gK[7] - gK[5], gd1 = gK[2] - gK[6], gd2 = gK[3] - gK[1];
        const float gdv = gd0 * v[0] + gd1 * v[1] + gd2 * v[2];
        const float ith = 1.f / th, ith3 = ith * ith * ith;
        gaa[j * 3 + 0] = gd0 * ith - gdv * e0 * ith3 + gth * e0 * ith;
        gaa[j * 3 + 1] = gd1 * ith - gdv * e1 * ith3 + gth * e1 * ith;
        gaa[j * 3 + 2] = gd2 * ith - gdv * e2 * ith3 + gth * e2 * ith;
    }
    // priors act directly on pose[3:]
    for (int q = 0; q < ND; ++q) gaa[3 + q] = fmaf(0.01f, a.gprior[(size_t)t * ND + q], gaa[3 + q]);
    for (int q = 0; q < 4; ++q) {
        const float sg = c_angle_sign[q];
        const float e = expf(a.aa[(size_t)t * 72 + 3 + c_angle_idx[q]] * sg);
        gaa[3 + c_angle_idx[q]] = fmaf(231.04f * 2.f * sg, e * e, gaa[3 + c_angle_idx[q]]);
    }
    for (int e = 0; e < 72; ++e) a.gaa[(size_t)t * 72 + e] = gaa[e];
    for (int r = 0; r < 3; ++r) a.gtran[(size_t)t * 3 + r] = gtr[r];
}

// loss = sum_t (frame + smooth) + T * sum_t imu    (losses.py:63 broadcast, see above); one block, fixed order
__global__ void __launch_bounds__(256) rc_smplify_sum_kernel(const float* __restrict__ lossf, int T, int with_smooth, float* __restrict__ out) {
    __shared__ double s1[256], s2[256];
    double a = 0.0, b = 0.0;
    for (int t = threadIdx.x; t < T; t += 256) {
        a += (double)lossf[t] + (with_smooth ? (double)lossf[2 * T + t] : 0.0);
        b += (double)lossf[T + t];
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(s1[0] + (double)T * s2[0]);
}

}  // namespace

extern "C" {

int rc_smplify_create(rc_smplify** out, const rc_model* model, const float* h_means, const float* h_precisions,
                      const float* h_log_nll_weights, int32_t T) {
    RC_ARG(out && model && h_means && h_precisions && h_log_nll_weights && T > 0);
    rc_smplify* s = new rc_smplify();
    s->model = model; s->T = T;
    SmplifyConst c;
    for (int m = 0; m < NG; ++m) {
        for (int j = 0; j < ND; ++j) c.means[m][j] = h_means[m * ND + j];
        c.logw[m] = h_log_nll_weights[m];
    }
    std::vector<float> psym((size_t)NG * ND * ND);
    for (int m = 0; m < NG; ++m)
        for (int i = 0; i < ND; ++i)
            for (int j = 0; j < ND; ++j)
                psym[((size_t)m * ND + i) * ND + j] = 0.5f * (h_precisions[((size_t)m * ND + i) * ND + j] + h_precisions[((size_t)m * ND + j) * ND + i]);
    auto alloc = [&](void** p, size_t bytes) -> bool {
        if (cudaMalloc(p, bytes) != cudaSuccess) return false;
        s->allocs.push_back(*p);
        return true;
    };
    bool ok = alloc((void**)&s->d_const, sizeof(SmplifyConst)) && alloc((void**)&s->d_psym, psym.size() * 4) &&
              alloc((void**)&s->d_G, (size_t)T * 288 * 4) && alloc((void**)&s->d_p, (size_t)T * 99 * 4) &&
              alloc((void**)&s->d_proj, (size_t)T * 66 * 4) && alloc((void**)&s->d_lossf, (size_t)T * 3 * 4) &&
              alloc((void**)&s->d_prior, (size_t)T * 4) && alloc((void**)&s->d_gprior, (size_t)T * ND * 4);
    if (ok) ok = cudaMemcpy(s->d_const, &c, sizeof(c), cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(s->d_psym, psym.data(), psym.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { rc_set_error("rc_smplify_create: %s", cudaGetErrorString(cudaGetLastError())); rc_smplify_destroy(s); return RC_ERR_CUDA; }
    *out = s;
    return RC_OK;
}

void rc_smplify_destroy(rc_smplify* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    delete s;
}

int rc_smplify_loss_grad(rc_smplify* s, const float* aa, const float* tran, const float* j2d, const float* conf, const float* camk,
                         const float* ref3d, const float* imu_aa, int rodrigues, float* loss, float* grad_aa, float* grad_tran,
                         float* reproj, void* stream) {
    RC_ARG(s && aa && tran && j2d && conf && camk && ref3d && imu_aa && (loss || reproj));
    RC_ARG((grad_aa == nullptr) == (grad_tran == nullptr));
    const int T = s->T;
    RC_ARG(rodrigues >= 0 && rodrigues <= 2 && (rodrigues != 2 || loss == nullptr));
    if (rodrigues != 2) {
        RC_LAUNCH(rc_gmm_kernel, T, 256, 0, stream, s->d_const, s->d_psym, aa, s->d_prior, s->d_gprior);
        RC_CHECK_LAUNCH();
    }
    FwdArgs f;
    f.aa = aa; f.tran = tran; f.j2d = j2d; f.conf = conf; f.camk = camk; f.ref3d = ref3d; f.imu_aa = imu_aa; f.prior = s->d_prior;
    f.G = s->d_G; f.p = s->d_p; f.proj = s->d_proj; f.lossf = s->d_lossf; f.reproj = reproj; f.T = T; f.rodrigues = rodrigues;
    RC_LAUNCH(rc_smplify_fwd_kernel, rc_cdiv(T, 32), 32, 0, stream, s->model->d_const, f);
    RC_CHECK_LAUNCH();
    if (grad_aa) {
        BwdArgs b;
        b.aa = aa; b.j2d = j2d; b.conf = conf; b.camk = camk; b.ref3d = ref3d; b.G = s->d_G; b.p = s->d_p; b.proj = s->d_proj;
        b.gprior = s->d_gprior; b.lossf = s->d_lossf; b.gaa = grad_aa; b.gtran = grad_tran; b.T = T;
        RC_LAUNCH(rc_smplify_bwd_kernel, rc_cdiv(T, 32), 32, 0, stream, s->model->d_const, b);
        RC_CHECK_LAUNCH();
    }
    if (loss) {
        RC_ARG(grad_aa != nullptr);     // the temporal terms of the loss value are produced by the backward kernel
        RC_LAUNCH(rc_smplify_sum_kernel, 1, 256, 0, stream, s->d_lossf, T, 1, loss);
        RC_CHECK_LAUNCH();
    }
    return RC_OK;
}

}  // extern "C"
