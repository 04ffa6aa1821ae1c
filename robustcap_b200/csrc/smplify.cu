// SMPLify objective and its analytic gradient (net/smplify/losses.py:23-91 on top of temporal_smplify.py:153-165).
//
// The reference builds an autograd graph through batch_rodrigues -> 4x4 FK chain -> linear-blend skinning of all 6890
// vertices -> 33-point gather -> loss and back-propagates it ~26 times per sequence.  Here the same scalar function of
// (axis-angle pose [T,72], translation [T,3]) is evaluated with hand-derived derivatives, skinning only the 21 vertices
// that the 33 MediaPipe points read:
//   rc_gmm_kernel          block per frame, warp per mixture component: min_m(0.5 d^T P_m d - log w_m) and P_m* d
//   rc_smplify_fwd_kernel  warp per frame (lanes = joints / key points / IMU sensors): Rodrigues, FK chain by tree level, key points,
//                          projection, per-frame loss terms; contiguous global accesses
//   rc_smplify_bwd_kernel  warp per frame: d loss / d points (incl. the temporal L1 terms that couple t-1, t, t+1), back through
//                          skinning (lane j gathers over the key points), the kinematic chain (parents gather from their children, level
//                          by level) and Rodrigues
//   rc_sum_kernel          deterministic reduction of the per-frame losses
//   rc_smplify_lbfgs_kernel  the optimiser: torch.optim.LBFGS.step + strong Wolfe on the device, the closure above inside
#include <algorithm>
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
    // scratch of rc_smplify_run (device-resident L-BFGS), sized for run_S sequences and run_hist history pairs
    float *r_vec = nullptr, *r_G = nullptr, *r_p = nullptr, *r_proj = nullptr, *r_lossf = nullptr, *r_prior = nullptr, *r_gprior = nullptr;
    int run_S = 0, run_hist = 0;
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
        for (int j = 0; j < ND; ++j) s = fmaf(P[j * ND + i], d[m][j], s);     // P is symmetric: column walk = coalesced across the lanes
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

// The same prior for one frame by ONE warp (components one after the other; per-lane arithmetic and reduction order identical to
// rc_gmm_kernel, so both give the same bits).  d / pd / best: 3 x 69 floats of shared memory private to the warp.
__device__ __forceinline__ void gmm_frame_warp(const SmplifyConst* __restrict__ C, const float* __restrict__ Psym, const float* aa_t,
                                               float* prior_t, float* gprior_t, float* d, float* pd, float* best_pd, int lane) {
    float best_ll = 0.f;
    for (int m = 0; m < NG; ++m) {
        for (int j = lane; j < ND; j += 32) d[j] = aa_t[3 + j] - C->means[m][j];
        __syncwarp();
        const float* P = Psym + (size_t)m * ND * ND;
        float q = 0.f;
        for (int i = lane; i < ND; i += 32) {
            float s = 0.f;
            for (int j = 0; j < ND; ++j) s = fmaf(P[j * ND + i], d[j], s);        // P is symmetric: column walk = coalesced across the lanes
            pd[i] = s;
            q = fmaf(s, d[i], q);
        }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float ll = 0.5f * q - C->logw[m];
        __syncwarp();
        if (m == 0 || ll < best_ll) {                                    // torch.min: first minimum
            best_ll = ll;
            for (int j = lane; j < ND; j += 32) best_pd[j] = pd[j];
        }
        __syncwarp();
    }
    if (lane == 0) prior_t[0] = best_ll;
    for (int j = lane; j < ND; j += 32) gprior_t[j] = best_pd[j];
}

// ---- forward ---------------------------------------------------------------------------------------------------------
struct FwdArgs {
    const float *aa, *tran, *j2d, *conf, *camk, *ref3d, *imu_aa, *prior;
    float *G, *p, *proj, *lossf, *reproj;
    int T, rodrigues;     // 0: batch_rodrigues (temporal_smplify.py:25-59), 1: axis_angle_to_rotation_matrix (angular.py:221-233),
                          // 2: the pose pointer holds [T,24,3,3] rotation matrices (get_fitting_loss, value only)
};

// ---- backward ----------------------------------------------------------------------------------------------------------
struct BwdArgs {
    const float *aa, *j2d, *conf, *camk, *ref3d, *G, *p, *proj, *gprior;
    float *lossf, *gaa, *gtran;
    int T;
};


// ---- warp-per-frame closure ---------------------------------------------------------------------------------------------------
// One warp evaluates one frame: lanes = joints (Rodrigues, chain by tree level, Rodrigues backward) / key points (skinning of the 21
// vertices, projection, loss terms, their gradients) / IMU sensors (the six cv2.Rodrigues log maps); the per-frame arrays live in
// shared memory, global reads and writes are contiguous across the warp.  Same formulas as the scalar derivation they replace.
struct SmpWarp {
    float L[RC_NJ][12];        // local transforms [R | bone]; backward: reused for gR (9) + gt (3) per joint
    float G[RC_NJ][12];        // global transforms
    float p[RC_NKP][3];        // key points
    float gp[RC_NKP][3];       // d loss / d key point
    float red[4];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// key point k from the global transforms (model.py:235-241 restricted to one vertex / joint)
__device__ __forceinline__ void keypoint_from_G(const RcModelConst& M, const float (*G)[12], const float* tran, int k, float* o) {
    if (M.kp_is_joint[k]) {
        const int j = M.kp_index[k];
        for (int r = 0; r < 3; ++r) o[r] = G[j][r * 4 + 3] + tran[r];
        return;
    }
    float Tv[12];
    for (int e = 0; e < 12; ++e) Tv[e] = 0.f;
    for (int j = 0; j < RC_NJ; ++j) {
        const float w = M.kp_w[k][j];
        if (w != 0.f) {
            const float tx = G[j][3] - (G[j][0] * M.jrest[j][0] + G[j][1] * M.jrest[j][1] + G[j][2] * M.jrest[j][2]);
            const float ty = G[j][7] - (G[j][4] * M.jrest[j][0] + G[j][5] * M.jrest[j][1] + G[j][6] * M.jrest[j][2]);
            const float tz = G[j][11] - (G[j][8] * M.jrest[j][0] + G[j][9] * M.jrest[j][1] + G[j][10] * M.jrest[j][2]);
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Tv[r * 4 + c] = fmaf(w, G[j][r * 4 + c], Tv[r * 4 + c]);
            Tv[3] = fmaf(w, tx, Tv[3]); Tv[7] = fmaf(w, ty, Tv[7]); Tv[11] = fmaf(w, tz, Tv[11]);
        }
    }
    for (int r = 0; r < 3; ++r)
        o[r] = Tv[r * 4] * M.kp_rest[k][0] + Tv[r * 4 + 1] * M.kp_rest[k][1] + Tv[r * 4 + 2] * M.kp_rest[k][2] + Tv[r * 4 + 3] + tran[r];
}

__device__ __forceinline__ void smplify_fwd_warp(const RcModelConst& M, const FwdArgs& a, int t, SmpWarp& S, int lane) {
    // local transforms, one joint per lane
    if (lane < RC_NJ) {
        float R[9];
        if (a.rodrigues == 0) rc_batch_rodrigues(a.aa + (size_t)t * 72 + lane * 3, R);
        else if (a.rodrigues == 1) rc_aa_to_mat(a.aa + (size_t)t * 72 + lane * 3, R);
        else { for (int e = 0; e < 9; ++e) R[e] = a.aa[(size_t)t * 216 + lane * 9 + e]; }
        for (int r = 0; r < 3; ++r) { S.L[lane][r * 4] = R[r * 3]; S.L[lane][r * 4 + 1] = R[r * 3 + 1]; S.L[lane][r * 4 + 2] = R[r * 3 + 2]; S.L[lane][r * 4 + 3] = M.bone[lane][r]; }
        if (lane == 0) for (int e = 0; e < 12; ++e) S.G[0][e] = S.L[0][e];
    }
    __syncwarp();
    for (int lev = 1; lev <= M.max_depth; ++lev) {                         // chain by tree level (spatial.py:224-249)
        if (lane < RC_NJ && M.depth[lane] == lev) rc_rigid_mul(S.G[M.parent[lane]], S.L[lane], S.G[lane]);
        __syncwarp();
    }
    const float tr[3] = {a.tran[t * 3], a.tran[t * 3 + 1], a.tran[t * 3 + 2]};
    for (int k = lane; k < RC_NKP; k += 32) keypoint_from_G(M, S.G, tr, k, S.p[k]);
    __syncwarp();
    for (int e = lane; e < RC_NJ * 12; e += 32) a.G[(size_t)t * 288 + e] = S.G[e / 12][e % 12];
    for (int e = lane; e < RC_NKP * 3; e += 32) a.p[(size_t)t * 99 + e] = S.p[e / 3][e % 3];
    float K[6];
    for (int e = 0; e < 6; ++e) K[e] = a.camk[e];
    float loss = 0.f;
    for (int k = lane; k < RC_NKP; k += 32) {
        const float* p = S.p[k];
        // re-projection (losses.py:36-46): conf^2 * sum_xy gmof(K (p / p_z) - j2d, 100)
        const float x = p[0] / p[2], y = p[1] / p[2], z = p[2] / p[2];
        const float u = K[0] * x + K[1] * y + K[2] * z, v = K[3] * x + K[4] * y + K[5] * z;
        a.proj[(size_t)t * 66 + k * 2] = u; a.proj[(size_t)t * 66 + k * 2 + 1] = v;
        const float c = a.conf[(size_t)t * 33 + k];
        const float ru = u - a.j2d[(size_t)t * 66 + k * 2], rv = v - a.j2d[(size_t)t * 66 + k * 2 + 1];
        const float e = (1e4f * ru * ru) / (1e4f + ru * ru) + (1e4f * rv * rv) / (1e4f + rv * rv);
        const float rl = c * c * e;
        if (a.reproj) a.reproj[(size_t)t * 33 + k] = rl;
        loss += rl;
        // 3-D term (losses.py:32-34): sum_i |(p_i - p_0) - (ref_i - ref_0)|^2
        if (k >= 1)
            for (int r = 0; r < 3; ++r) {
                const float d = (p[r] - S.p[0][r]) - (a.ref3d[(size_t)t * 99 + k * 3 + r] - a.ref3d[(size_t)t * 99 + r]);
                loss = fmaf(d, d, loss);
            }
    }
    // IMU term (losses.py:39-40), value only: 0.5^2 |aa(imu_ori) - aa(R_glb[ji_mask])|^2, cv2.Rodrigues semantics, one sensor per lane
    float imu = 0.f;
    if (lane < 6) {
        const int j = c_ji_mask[lane];
        float R[9], v[3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = S.G[j][r * 4 + c];
        rc_mat_to_aa(R, v);
        for (int r = 0; r < 3; ++r) { const float d = a.imu_aa[(size_t)t * 18 + lane * 3 + r] - v[r]; imu = fmaf(d, d, imu); }
    }
    // GMM prior (0.1^2) and angle prior (15.2^2 * exp(+-x)^2)   (losses.py:49-54), lanes 8..11
    if (a.rodrigues != 2 && lane >= 8 && lane < 12) {
        const int q = lane - 8;
        const float e = expf(a.aa[(size_t)t * 72 + 3 + c_angle_idx[q]] * c_angle_sign[q]);
        loss = fmaf(231.04f, e * e, loss);
    }
    loss = warp_sum(loss);
    imu = warp_sum(imu);
    if (lane == 0) {
        if (a.rodrigues != 2) loss = fmaf(0.01f, a.prior[t], loss);
        a.lossf[t] = loss;
        a.lossf[a.T + t] = 0.25f * imu;
    }
}

__device__ __forceinline__ void smplify_bwd_warp(const RcModelConst& M, const BwdArgs& a, int t, SmpWarp& S, int lane) {
    float K[6];
    for (int e = 0; e < 6; ++e) K[e] = a.camk[e];
    // this frame's transforms and points back into shared memory (coalesced)
    for (int e = lane; e < RC_NJ * 12; e += 32) S.G[e / 12][e % 12] = a.G[(size_t)t * 288 + e];
    for (int e = lane; e < RC_NKP * 3; e += 32) S.p[e / 3][e % 3] = a.p[(size_t)t * 99 + e];
    __syncwarp();
    float smooth = 0.f, g0[3] = {0.f, 0.f, 0.f}, gtr[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < RC_NKP; k += 32) {
        const float* p = S.p[k];
        const float* pr = a.proj + (size_t)t * 66 + k * 2;
        const float c = a.conf[(size_t)t * 33 + k], c2 = c * c;
        const float ru = pr[0] - a.j2d[(size_t)t * 66 + k * 2], rv = pr[1] - a.j2d[(size_t)t * 66 + k * 2 + 1];
        const float du = 1e4f + ru * ru, dv = 1e4f + rv * rv;
        float gu = c2 * 2.f * ru * 1e8f / (du * du), gv = c2 * 2.f * rv * 1e8f / (dv * dv);
        float g3[3] = {0.f, 0.f, 0.f};
        if (t >= 1) {                                                       // temporal L1 terms (losses.py:66-84)
            const float* pp = a.p + (size_t)(t - 1) * 99 + k * 3;
            const float* qq = a.proj + (size_t)(t - 1) * 66 + k * 2;
            for (int r = 0; r < 2; ++r) {
                const float d = pr[r] - qq[r];
                smooth = fmaf(1e-4f * c2, fabsf(d), smooth);
                const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
                if (r == 0) gu = fmaf(1e-4f * c2, sg, gu); else gv = fmaf(1e-4f * c2, sg, gv);
            }
            for (int r = 0; r < 3; ++r) {
                const float d = p[r] - pp[r];
                smooth = fmaf(c2, fabsf(d), smooth);
                g3[r] += c2 * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
            }
        }
        if (t + 1 < a.T) {
            const float cn = a.conf[(size_t)(t + 1) * 33 + k], cn2 = cn * cn;
            const float* pn = a.p + (size_t)(t + 1) * 99 + k * 3;
            const float* qn = a.proj + (size_t)(t + 1) * 66 + k * 2;
            for (int r = 0; r < 2; ++r) {
                const float d = qn[r] - pr[r];
                const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
                if (r == 0) gu = fmaf(-1e-4f * cn2, sg, gu); else gv = fmaf(-1e-4f * cn2, sg, gv);
            }
            for (int r = 0; r < 3; ++r) {
                const float d = pn[r] - p[r];
                g3[r] -= cn2 * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
            }
        }
        const float x = p[0], y = p[1], z = p[2], iz = 1.f / z;
        g3[0] += (gu * K[0] + gv * K[3]) * iz;
        g3[1] += (gu * K[1] + gv * K[4]) * iz;
        g3[2] += -(gu * (K[0] * x + K[1] * y) + gv * (K[3] * x + K[4] * y)) * iz * iz;
        if (k >= 1) {
            for (int r = 0; r < 3; ++r) {
                const float d = (p[r] - S.p[0][r]) - (a.ref3d[(size_t)t * 99 + k * 3 + r] - a.ref3d[(size_t)t * 99 + r]);
                g3[r] = fmaf(2.f, d, g3[r]);
                g0[r] = fmaf(-2.f, d, g0[r]);
            }
        }
        for (int r = 0; r < 3; ++r) S.gp[k][r] = g3[r];
    }
    smooth = warp_sum(smooth);
    for (int r = 0; r < 3; ++r) g0[r] = warp_sum(g0[r]);
    __syncwarp();
    if (lane == 0) { for (int r = 0; r < 3; ++r) S.gp[0][r] += g0[r]; a.lossf[2 * a.T + t] = smooth; }
    __syncwarp();
    for (int k = lane; k < RC_NKP; k += 32) for (int r = 0; r < 3; ++r) gtr[r] += S.gp[k][r];
    for (int r = 0; r < 3; ++r) gtr[r] = warp_sum(gtr[r]);
    // points -> global joint transforms: lane j gathers over the key points (S.L[j] = [gR (9) | gt (3)])
    if (lane < RC_NJ) {
        const int j = lane;
        float gR[9], gt[3] = {0.f, 0.f, 0.f};
        for (int e = 0; e < 9; ++e) gR[e] = 0.f;
        for (int k = 0; k < RC_NKP; ++k) {
            const float* g = S.gp[k];
            if (M.kp_is_joint[k]) {
                if (M.kp_index[k] == j) for (int r = 0; r < 3; ++r) gt[r] += g[r];
            } else {
                const float w = M.kp_w[k][j];
                if (w != 0.f) {
                    const float dx = M.kp_rest[k][0] - M.jrest[j][0], dy = M.kp_rest[k][1] - M.jrest[j][1], dz = M.kp_rest[k][2] - M.jrest[j][2];
                    for (int r = 0; r < 3; ++r) {
                        const float wg = w * g[r];
                        gR[r * 3] = fmaf(wg, dx, gR[r * 3]); gR[r * 3 + 1] = fmaf(wg, dy, gR[r * 3 + 1]); gR[r * 3 + 2] = fmaf(wg, dz, gR[r * 3 + 2]);
                        gt[r] += wg;
                    }
                }
            }
        }
        for (int e = 0; e < 9; ++e) S.L[j][e] = gR[e];
        for (int r = 0; r < 3; ++r) S.L[j][9 + r] = gt[r];
    }
    __syncwarp();
    // chain, leaves to root, one tree level at a time: a parent gathers from its children
    //   R_c = R_p Rl_c, t_c = R_p b_c + t_p  =>  gR_p += gR_c Rl_c^T + gt_c b_c^T, gt_p += gt_c
    for (int lev = M.max_depth; lev >= 1; --lev) {
        if (lane < RC_NJ && M.depth[lane] == lev - 1) {
            const int pj = lane;
            for (int c = 1; c < RC_NJ; ++c) {
                if (M.parent[c] != pj || M.depth[c] != lev) continue;
                const float* Gp = S.G[pj];
                const float* Gc = S.G[c];
                float Rl[9];
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) Rl[r * 3 + q] = Gp[0 * 4 + r] * Gc[0 * 4 + q] + Gp[1 * 4 + r] * Gc[1 * 4 + q] + Gp[2 * 4 + r] * Gc[2 * 4 + q];
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) {
                        const float v = S.L[c][r * 3] * Rl[q * 3] + S.L[c][r * 3 + 1] * Rl[q * 3 + 1] + S.L[c][r * 3 + 2] * Rl[q * 3 + 2];
                        S.L[pj][r * 3 + q] += v + S.L[c][9 + r] * M.bone[c][q];
                    }
                for (int r = 0; r < 3; ++r) S.L[pj][9 + r] += S.L[c][9 + r];
            }
        }
        __syncwarp();
    }
    // Rodrigues backward, one joint per lane: R = I + sin(th) K(d) + (1 - cos(th)) K(d)^2, th = |v + 1e-8|, d = v / th
    if (lane < RC_NJ) {
        const int j = lane;
        float gRl[9];
        if (j > 0) {
            const float* Gp = S.G[M.parent[j]];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) gRl[r * 3 + c] = Gp[0 * 4 + r] * S.L[j][0 * 3 + c] + Gp[1 * 4 + r] * S.L[j][1 * 3 + c] + Gp[2 * 4 + r] * S.L[j][2 * 3 + c];
        } else {
            for (int e = 0; e < 9; ++e) gRl[e] = S.L[0][e];
        }
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
        S.p[j][0] = gd0 * ith - gdv * e0 * ith3 + gth * e0 * ith;            // S.p is free now: gaa staging
        S.p[j][1] = gd1 * ith - gdv * e1 * ith3 + gth * e1 * ith;
        S.p[j][2] = gd2 * ith - gdv * e2 * ith3 + gth * e2 * ith;
    }
    __syncwarp();
    float* gaa = &S.p[0][0];                                               // [72]
    // priors act directly on pose[3:]
    for (int q = lane; q < ND; q += 32) gaa[3 + q] = fmaf(0.01f, a.gprior[(size_t)t * ND + q], gaa[3 + q]);
    __syncwarp();
    if (lane < 4) {
        const float sg = c_angle_sign[lane];
        const float e = expf(a.aa[(size_t)t * 72 + 3 + c_angle_idx[lane]] * sg);
        gaa[3 + c_angle_idx[lane]] = fmaf(231.04f * 2.f * sg, e * e, gaa[3 + c_angle_idx[lane]]);
    }
    __syncwarp();
    for (int e = lane; e < 72; e += 32) a.gaa[(size_t)t * 72 + e] = gaa[e];
    if (lane < 3) a.gtran[(size_t)t * 3 + lane] = gtr[lane];
    __syncwarp();
}

constexpr int kSmpWarps = 8;
__global__ void __launch_bounds__(kSmpWarps * 32) rc_smplify_fwd_kernel(const RcModelConst* __restrict__ Mp, FwdArgs a) {
    __shared__ SmpWarp S[kSmpWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kSmpWarps + warp;
    if (t >= a.T) return;
    smplify_fwd_warp(*Mp, a, t, S[warp], lane);
}
__global__ void __launch_bounds__(kSmpWarps * 32) rc_smplify_bwd_kernel(const RcModelConst* __restrict__ Mp, BwdArgs a) {
    __shared__ SmpWarp S[kSmpWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kSmpWarps + warp;
    if (t >= a.T) return;
    smplify_bwd_warp(*Mp, a, t, S[warp], lane);
}

// ---- device-resident L-BFGS (torch.optim.LBFGS.step with line_search_fn='strong_wolfe', temporal_smplify.py:151-166) ----------------
// ONE thread block per sequence runs the whole optimisation: closure evaluations (GMM prior -> forward -> backward -> loss sum, the
// functions above, one thread per frame / one warp per frame for the prior), the two-loop recursion, the strong-Wolfe bracketing /
// zoom with cubic interpolation, every reduction inside the block in a fixed order.  No host round trip, no launch per iteration;
// independent sequences run as independent blocks.  Scalar arithmetic is float32 like torch's 0-dim tensors (lbfgs.py); the
// expressions follow torch/optim/lbfgs.py line by line (_cubic_interpolate, _strong_wolfe, LBFGS.step; defaults tolerance_grad 1e-7,
// tolerance_change 1e-9, history_size 100, max_eval = max_iter * 5 // 4, first step t = min(1, 1 / |g|_1) * lr).
constexpr int kLbThreadsMax = 512;
enum { LB_X = 0, LB_XINIT, LB_D, LB_G, LB_PG, LB_GN, LB_GPREV, LB_BG0, LB_BG1, LB_Q, LB_NFIXED };   // then old_dirs[h], old_stps[h]

struct LbfgsArgs {
    const float *aa0, *tran0, *j2d, *conf, *camk, *ref3d, *imu_aa;
    float* vec;            // [S][LB_NFIXED + 2 hist][n]
    float *G, *p, *proj, *lossf, *prior, *gprior;     // per-sequence scratch of the closure
    float *aa_out, *tran_out, *stats;
    const SmplifyConst* C;
    const float* Psym;
    int T, n, max_iter, max_eval, hist, camk_stride;
    float lr;
};

struct LbShared {
    double red[kLbThreadsMax];
    double s1[256], s2[256];
    float gmm[kLbThreadsMax / 32][3][ND + 3];
    float bcast;
};

__device__ __forceinline__ double lb_block_sum(double v, LbShared& sh) {
    const int tid = threadIdx.x, nt = blockDim.x;
    sh.red[tid] = v;
    __syncthreads();
    for (int o = kLbThreadsMax / 2; o > 0; o >>= 1) {
        if (tid < o && tid + o < nt) sh.red[tid] += sh.red[tid + o];
        __syncthreads();
    }
    const double r = sh.red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ float lb_block_max(float v, LbShared& sh) {
    const int tid = threadIdx.x, nt = blockDim.x;
    sh.red[tid] = (double)v;
    __syncthreads();
    for (int o = kLbThreadsMax / 2; o > 0; o >>= 1) {
        if (tid < o && tid + o < nt) sh.red[tid] = fmax(sh.red[tid], sh.red[tid + o]);
        __syncthreads();
    }
    const float r = (float)sh.red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ float lb_dot(const float* a, const float* b, int n, LbShared& sh) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)a[i] * (double)b[i];
    return (float)lb_block_sum(acc, sh);
}
__device__ __forceinline__ float lb_absmax(const float* a, float scale, int n, LbShared& sh) {
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(a[i] * scale));
    return lb_block_max(m, sh);
}
__device__ __forceinline__ float lb_abssum(const float* a, int n, LbShared& sh) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)fabsf(a[i]);
    return (float)lb_block_sum(acc, sh);
}
__device__ __forceinline__ void lb_copy(float* dst, const float* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}
// dst = a + alpha * b
__device__ __forceinline__ void lb_axpy_to(float* dst, const float* a, float alpha, const float* b, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = fmaf(alpha, b[i], a[i]);
    __syncthreads();
}

// torch/optim/lbfgs.py:_cubic_interpolate, float32
__device__ __forceinline__ float lb_cubic(float x1, float f1, float g1, float x2, float f2, float g2, bool bounded, float lo, float hi) {
    if (!bounded) { lo = (x1 <= x2) ? x1 : x2; hi = (x1 <= x2) ? x2 : x1; }
    const float d1 = g1 + g2 - (float)(3.0 * ((double)f1 - (double)f2)) / (x1 - x2);
    const float d2sq = d1 * d1 - g1 * g2;
    if (d2sq >= 0.f) {
        const float d2 = sqrtf(d2sq);
        float mp;
        if (x1 <= x2) mp = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.f * d2));
        else mp = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.f * d2));
        return fminf(fmaxf(mp, lo), hi);
    }
    return (lo + hi) / 2.f;
}

// closure: loss and gradient at x (aa | tran) -> g; every thread returns the loss
__device__ float lb_eval(const RcModelConst& M, const LbfgsArgs& a, int seq, const float* x, float* g, LbShared& sh, SmpWarp* sw) {
    const int T = a.T, tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const size_t so = (size_t)seq * T;
    float* prior = a.prior + so;
    float* gprior = a.gprior + so * ND;
    for (int t = warp; t < T; t += nw)
        gmm_frame_warp(a.C, a.Psym, x + (size_t)t * 72, prior + t, gprior + (size_t)t * ND, sh.gmm[warp][0], sh.gmm[warp][1], sh.gmm[warp][2], lane);
    __syncthreads();
    FwdArgs f;
    f.aa = x; f.tran = x + (size_t)T * 72; f.j2d = a.j2d + so * 66; f.conf = a.conf + so * 33; f.camk = a.camk + (size_t)seq * a.camk_stride;
    f.ref3d = a.ref3d + so * 99; f.imu_aa = a.imu_aa + so * 18; f.prior = prior;
    f.G = a.G + so * 288; f.p = a.p + so * 99; f.proj = a.proj + so * 66; f.lossf = a.lossf + so * 3; f.reproj = nullptr; f.T = T; f.rodrigues = 0;
    for (int t = warp; t < T; t += nw) smplify_fwd_warp(M, f, t, sw[warp], lane);
    __syncthreads();
    BwdArgs b;
    b.aa = x; b.j2d = f.j2d; b.conf = f.conf; b.camk = f.camk; b.ref3d = f.ref3d; b.G = f.G; b.p = f.p; b.proj = f.proj; b.gprior = gprior;
    b.lossf = f.lossf; b.gaa = g; b.gtran = g + (size_t)T * 72; b.T = T;
    for (int t = warp; t < T; t += nw) smplify_bwd_warp(M, b, t, sw[warp], lane);
    __syncthreads();
    // loss = sum_t (frame + smooth) + T * sum_t imu: the reduction order of rc_smplify_sum_kernel (256 strided partial sums, tree)
    if (tid < 256) {
        double u = 0.0, v = 0.0;
        for (int t = tid; t < T; t += 256) { u += (double)f.lossf[t] + (double)f.lossf[2 * T + t]; v += (double)f.lossf[T + t]; }
        sh.s1[tid] = u; sh.s2[tid] = v;
    }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { sh.s1[tid] += sh.s1[tid + o]; sh.s2[tid] += sh.s2[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) sh.bcast = (float)(sh.s1[0] + (double)T * sh.s2[0]);
    __syncthreads();
    const float loss = sh.bcast;
    __syncthreads();
    return loss;
}

__global__ void __launch_bounds__(kLbThreadsMax) rc_smplify_lbfgs_kernel(const RcModelConst* __restrict__ Mp, const LbfgsArgs a) {
    extern __shared__ __align__(16) unsigned char lb_smem[];
    LbShared& sh = *reinterpret_cast<LbShared*>(lb_smem);
    SmpWarp* sw = reinterpret_cast<SmpWarp*>(lb_smem + ((sizeof(LbShared) + 15) / 16) * 16);
    const RcModelConst& M = *Mp;
    const int seq = blockIdx.x, n = a.n, T = a.T;
    const int nvec = LB_NFIXED + 2 * a.hist;
    float* V = a.vec + (size_t)seq * nvec * n;
    auto vec = [&](int i) { return V + (size_t)i * n; };
    auto old_dir = [&](int i) { return V + (size_t)(LB_NFIXED + i) * n; };
    auto old_stp = [&](int i) { return V + (size_t)(LB_NFIXED + a.hist + i) * n; };
    float* X = vec(LB_X);
    for (int i = threadIdx.x; i < T * 72; i += blockDim.x) X[i] = a.aa0[(size_t)seq * T * 72 + i];
    for (int i = threadIdx.x; i < T * 3; i += blockDim.x) X[T * 72 + i] = a.tran0[(size_t)seq * T * 3 + i];
    __syncthreads();
    const float tol_grad = 1e-7f, tol_change = 1e-9f, c1 = 1e-4f, c2 = 0.9f;
    float ro[100], al[100];
    int nh = 0;
    float H_diag = 1.f, t = 0.f;

    float loss = lb_eval(M, a, seq, X, vec(LB_G), sh, sw);
    const float first_loss = loss;
    int evals = 1, n_iter = 0;
    bool stop = lb_absmax(vec(LB_G), 1.f, n, sh) <= tol_grad;
    while (!stop && n_iter < a.max_iter) {
        ++n_iter;
        float* G = vec(LB_G);
        float* D = vec(LB_D);
        if (n_iter == 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) D[i] = -G[i];
            __syncthreads();
            nh = 0; H_diag = 1.f;
        } else {
            // memory update: y = g - prev_g, s = d * t
            float* Y = old_dir(nh < a.hist ? nh : a.hist - 1);
            float* S = old_stp(nh < a.hist ? nh : a.hist - 1);
            if (nh == a.hist) {                                       // shift the history by one (limited memory)
                for (int h = 0; h + 1 < a.hist; ++h) { lb_copy(old_dir(h), old_dir(h + 1), n); lb_copy(old_stp(h), old_stp(h + 1), n); ro[h] = ro[h + 1]; }
                nh = a.hist - 1;
            }
            const float* PG = vec(LB_PG);
            for (int i = threadIdx.x; i < n; i += blockDim.x) { Y[i] = G[i] - PG[i]; S[i] = D[i] * t; }
            __syncthreads();
            const float ys = lb_dot(Y, S, n, sh);
            if (ys > 1e-10f) {
                ro[nh] = 1.f / ys;
                H_diag = ys / lb_dot(Y, Y, n, sh);
                ++nh;
            }
            float* Q = vec(LB_Q);
            for (int i = threadIdx.x; i < n; i += blockDim.x) Q[i] = -G[i];
            __syncthreads();
            for (int h = nh - 1; h >= 0; --h) {
                al[h] = lb_dot(old_stp(h), Q, n, sh) * ro[h];
                lb_axpy_to(Q, Q, -al[h], old_dir(h), n);
            }
            for (int i = threadIdx.x; i < n; i += blockDim.x) D[i] = Q[i] * H_diag;
            __syncthreads();
            for (int h = 0; h < nh; ++h) {
                const float be = lb_dot(old_dir(h), D, n, sh) * ro[h];
                lb_axpy_to(D, D, al[h] - be, old_stp(h), n);
            }
        }
        lb_copy(vec(LB_PG), G, n);
        const float prev_loss = loss;
        if (n_iter == 1) t = fminf(1.f, 1.f / lb_abssum(G, n, sh)) * a.lr;
        else t = a.lr;
        const float gtd = lb_dot(G, D, n, sh);
        if (gtd > -tol_change) break;

        // ---- _strong_wolfe(obj_func, x_init, t, d, loss, flat_grad, gtd, max_ls = max_eval - current_evals) ----
        const int max_ls = a.max_eval - evals;
        lb_copy(vec(LB_XINIT), X, n);
        const float* XI = vec(LB_XINIT);
        const float f = loss;
        const float d_norm = lb_absmax(D, 1.f, n, sh);
        float* GN = vec(LB_GN);
        lb_axpy_to(X, XI, t, D, n);
        float f_new = lb_eval(M, a, seq, X, GN, sh, sw);
        int ls_evals = 1;
        float gtd_new = lb_dot(GN, D, n, sh);
        float t_prev = 0.f, f_prev = f, gtd_prev = gtd;
        lb_copy(vec(LB_GPREV), G, n);
        bool done = false;
        int ls_iter = 0, nb = 0;
        float br[2] = {0.f, 0.f}, brf[2] = {0.f, 0.f}, brgtd[2] = {0.f, 0.f};
        float* BG[2] = {vec(LB_BG0), vec(LB_BG1)};
        while (ls_iter < max_ls) {
            if (f_new > (f + c1 * t * gtd) || (ls_iter > 1 && f_new >= f_prev) || (!(fabsf(gtd_new) <= -c2 * gtd) && gtd_new >= 0.f)) {
                br[0] = t_prev; br[1] = t; brf[0] = f_prev; brf[1] = f_new; brgtd[0] = gtd_prev; brgtd[1] = gtd_new; nb = 2;
                lb_copy(BG[0], vec(LB_GPREV), n); lb_copy(BG[1], GN, n);
                break;
            }
            if (fabsf(gtd_new) <= -c2 * gtd) {
                br[0] = t; brf[0] = f_new; nb = 1; done = true;
                lb_copy(BG[0], GN, n);
                break;
            }
            const float min_step = t + 0.01f * (t - t_prev), max_step = t * 10.f, tmp = t;
            t = lb_cubic(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, true, min_step, max_step);
            t_prev = tmp; f_prev = f_new; gtd_prev = gtd_new;
            lb_copy(vec(LB_GPREV), GN, n);
            lb_axpy_to(X, XI, t, D, n);
            f_new = lb_eval(M, a, seq, X, GN, sh, sw);
            ++ls_evals;
            gtd_new = lb_dot(GN, D, n, sh);
            ++ls_iter;
        }
        if (nb == 0) {                                                // reached max_ls without a bracket
            br[0] = 0.f; br[1] = t; brf[0] = f; brf[1] = f_new; nb = 2;
            lb_copy(BG[0], G, n); lb_copy(BG[1], GN, n);
        }
        bool insuf = false;
        int low = (brf[0] <= brf[nb - 1]) ? 0 : 1, high = 1 - low;
        while (!done && ls_iter < max_ls) {
            if (fabsf(br[1] - br[0]) * d_norm < tol_change) break;
            t = lb_cubic(br[0], brf[0], brgtd[0], br[1], brf[1], brgtd[1], false, 0.f, 0.f);
            const float bmax = fmaxf(br[0], br[1]), bmin = fminf(br[0], br[1]);
            const float eps = 0.1f * (bmax - bmin);
            if (fminf(bmax - t, t - bmin) < eps) {
                if (insuf || t >= bmax || t <= bmin) {
                    t = (fabsf(t - bmax) < fabsf(t - bmin)) ? bmax - eps : bmin + eps;
                    insuf = false;
                } else insuf = true;
            } else insuf = false;
            lb_axpy_to(X, XI, t, D, n);
            f_new = lb_eval(M, a, seq, X, GN, sh, sw);
            ++ls_evals;
            gtd_new = lb_dot(GN, D, n, sh);
            ++ls_iter;
            if (f_new > (f + c1 * t * gtd) || f_new >= brf[low]) {
                br[high] = t; brf[high] = f_new; brgtd[high] = gtd_new;
                lb_copy(BG[high], GN, n);
                low = (brf[0] <= brf[1]) ? 0 : 1; high = 1 - low;
            } else {
                if (fabsf(gtd_new) <= -c2 * gtd) done = true;
                else if (gtd_new * (br[high] - br[low]) >= 0.f) {
                    br[high] = br[low]; brf[high] = brf[low]; brgtd[high] = brgtd[low];
                    lb_copy(BG[high], BG[low], n);
                }
                br[low] = t; brf[low] = f_new; brgtd[low] = gtd_new;
                lb_copy(BG[low], GN, n);
            }
        }
        t = br[low]; loss = brf[low];
        lb_copy(G, BG[low], n);
        lb_axpy_to(X, XI, t, D, n);                                   // _add_grad(t, d) from x_init
        const bool opt_cond = lb_absmax(G, 1.f, n, sh) <= tol_grad;
        evals += ls_evals;
        if (n_iter == a.max_iter || evals >= a.max_eval || opt_cond) break;
        if (lb_absmax(D, t, n, sh) <= tol_change) break;
        if (fabs((double)loss - (double)prev_loss) < 1e-9) break;
    }
    for (int i = threadIdx.x; i < T * 72; i += blockDim.x) a.aa_out[(size_t)seq * T * 72 + i] = X[i];
    for (int i = threadIdx.x; i < T * 3; i += blockDim.x) a.tran_out[(size_t)seq * T * 3 + i] = X[T * 72 + i];
    if (threadIdx.x == 0 && a.stats) {
        float* st = a.stats + (size_t)seq * 4;
        st[0] = first_loss; st[1] = loss; st[2] = (float)evals; st[3] = (float)n_iter;
    }
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

static void smplify_free_run(rc_smplify* s) {
    cudaFree(s->r_vec); cudaFree(s->r_G); cudaFree(s->r_p); cudaFree(s->r_proj); cudaFree(s->r_lossf); cudaFree(s->r_prior); cudaFree(s->r_gprior);
    s->r_vec = s->r_G = s->r_p = s->r_proj = s->r_lossf = s->r_prior = s->r_gprior = nullptr;
    s->run_S = s->run_hist = 0;
}

void rc_smplify_destroy(rc_smplify* s) {
    if (!s) return;
    smplify_free_run(s);
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
    RC_LAUNCH(rc_smplify_fwd_kernel, rc_cdiv(T, kSmpWarps), kSmpWarps * 32, 0, stream, s->model->d_const, f);
    RC_CHECK_LAUNCH();
    if (grad_aa) {
        BwdArgs b;
        b.aa = aa; b.j2d = j2d; b.conf = conf; b.camk = camk; b.ref3d = ref3d; b.G = s->d_G; b.p = s->d_p; b.proj = s->d_proj;
        b.gprior = s->d_gprior; b.lossf = s->d_lossf; b.gaa = grad_aa; b.gtran = grad_tran; b.T = T;
        RC_LAUNCH(rc_smplify_bwd_kernel, rc_cdiv(T, kSmpWarps), kSmpWarps * 32, 0, stream, s->model->d_const, b);
        RC_CHECK_LAUNCH();
    }
    if (loss) {
        RC_ARG(grad_aa != nullptr);     // the temporal terms of the loss value are produced by the backward kernel
        RC_LAUNCH(rc_smplify_sum_kernel, 1, 256, 0, stream, s->d_lossf, T, 1, loss);
        RC_CHECK_LAUNCH();
    }
    return RC_OK;
}

// TemporalSMPLify.__call__'s optimisation (temporal_smplify.py:139-166) for n_seq independent sequences of T frames each (T = the
// handle's batch size): parameters start at (aa_init [S,T,72], tran_init [S,T,3]) and the optimised values are written to
// (aa_out, tran_out); stats (optional) [S,4] = {first loss, final loss, closure evaluations, iterations}.  One thread block per
// sequence runs torch.optim.LBFGS.step (strong Wolfe) entirely on the device; nothing is copied to the host.
int rc_smplify_run(rc_smplify* s, int32_t n_seq, const float* aa_init, const float* tran_init, const float* j2d, const float* conf,
                   const float* camk, int32_t camk_per_seq, const float* ref3d, const float* imu_aa, int32_t max_iter, float lr,
                   float* aa_out, float* tran_out, float* stats, void* stream) {
    RC_ARG(s && n_seq > 0 && aa_init && tran_init && j2d && conf && camk && ref3d && imu_aa && aa_out && tran_out && max_iter > 0);
    const int T = s->T;
    const int hist = std::min(100, max_iter);
    const size_t n = (size_t)T * 75;
    if (s->run_S < n_seq || s->run_hist < hist) {
        smplify_free_run(s);
        const size_t S = (size_t)n_seq, nvec = (size_t)LB_NFIXED + 2 * (size_t)hist;
        bool ok = cudaMalloc(&s->r_vec, S * nvec * n * 4) == cudaSuccess && cudaMalloc(&s->r_G, S * T * 288 * 4) == cudaSuccess &&
                  cudaMalloc(&s->r_p, S * T * 99 * 4) == cudaSuccess && cudaMalloc(&s->r_proj, S * T * 66 * 4) == cudaSuccess &&
                  cudaMalloc(&s->r_lossf, S * T * 3 * 4) == cudaSuccess && cudaMalloc(&s->r_prior, S * T * 4) == cudaSuccess &&
                  cudaMalloc(&s->r_gprior, S * T * ND * 4) == cudaSuccess;
        if (!ok) { rc_set_error("rc_smplify_run: out of device memory for %d sequences x %d frames", n_seq, T); smplify_free_run(s); return RC_ERR_ALLOC; }
        s->run_S = n_seq; s->run_hist = hist;
    }
    LbfgsArgs a;
    a.aa0 = aa_init; a.tran0 = tran_init; a.j2d = j2d; a.conf = conf; a.camk = camk; a.ref3d = ref3d; a.imu_aa = imu_aa;
    a.vec = s->r_vec; a.G = s->r_G; a.p = s->r_p; a.proj = s->r_proj; a.lossf = s->r_lossf; a.prior = s->r_prior; a.gprior = s->r_gprior;
    a.aa_out = aa_out; a.tran_out = tran_out; a.stats = stats; a.C = s->d_const; a.Psym = s->d_psym;
    a.T = T; a.n = (int)n; a.max_iter = max_iter; a.max_eval = max_iter * 5 / 4; a.hist = s->run_hist; a.camk_stride = camk_per_seq ? 9 : 0;
    a.lr = lr;
    // one warp per frame inside the block: 16 warps walk the frames of the sequence
    const int threads = kLbThreadsMax;
    const size_t smem = ((sizeof(LbShared) + 15) / 16) * 16 + (size_t)(threads / 32) * sizeof(SmpWarp);
    static bool attr = false;
    if (!attr) { RC_CUDA(cudaFuncSetAttribute(rc_smplify_lbfgs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    RC_LAUNCH(rc_smplify_lbfgs_kernel, n_seq, threads, smem, stream, s->model->d_const, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

}  // extern "C"
