// Warp-cooperative versions of the per-stream frame logic (rc_rows.h): one warp per stream, lanes = joints / key points.
// Every per-element expression is the scalar one of rc_rows.h (same operation order), so the results are bit-identical to
// the one-thread-per-stream kernels (tests/test_gpu_parity.py::test_warp_rows_match_scalar); only the mapping of work to
// threads changes: dependent 23-step chains become 9 tree levels, loads/stores are coalesced across the warp.
#pragma once
#include "rc_rows.h"

// rc_normalise_kp (sig_mp.py:150-152 / 268-270) across a warp: lanes = key points (lane 0 also takes key point 32).  max / min do not
// depend on the order and every quotient / difference is the scalar function's, so the result is bit-identical to rc_normalise_kp.
__device__ __forceinline__ void rc_normalise_kp_warp(const float* kp, float* out, int lane) {
    float u = kp[lane * 3], v = kp[lane * 3 + 1];
    float umax = u, umin = u, vmax = v, vmin = v;
    if (lane == 0) {
        const float u2 = kp[32 * 3], v2 = kp[32 * 3 + 1];
        umax = fmaxf(umax, u2); umin = fminf(umin, u2); vmax = fmaxf(vmax, v2); vmin = fminf(vmin, v2);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, s)); umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, s));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s)); vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
    }
    const float sc = fmaxf(RC_SUB(umax, umin), RC_SUB(vmax, vmin));
    const float rx = RC_DIV(kp[23 * 3], sc), ry = RC_DIV(kp[23 * 3 + 1], sc);
    for (int i = lane; i < RC_NKP; i += 32) {
        float x = RC_DIV(kp[i * 3], sc), y = RC_DIV(kp[i * 3 + 1], sc);
        if (i != 23) { x = RC_SUB(x, rx); y = RC_SUB(y, ry); }
        out[i * 3] = x; out[i * 3 + 1] = y; out[i * 3 + 2] = kp[i * 3 + 2];
    }
}

struct RcKinWarpSmem {
    float y7[144];
    float G[RC_NJ][9];
    float pose[RC_NJ][9];
    float pb[RC_NJ][3];
    float jp[RC_NJ][3];
    float F[RC_NJ][12];
    float joint[RC_NJ][3];
    float kp[RC_NKP][3];
    float syn[RC_NKP * 3];
    float tran[4];
    RcRowState st;
};

// returns need_init (uniform across the warp)
__device__ __forceinline__ int rc_kin_warp(const RcNetCfg& cfg, const RcModelConst& M, RcKinWarpSmem& S, RcRowState* gst, int flags,
                                           const float* __restrict__ gy7, const float* y8, const float* vr, const float* pc,
                                           const float* rcr, float conf, const float* gravity, const float* first_tran,
                                           float* __restrict__ gpose, float* __restrict__ gtran, float* __restrict__ x4,
                                           float* __restrict__ x6, int lane, int* branch = nullptr) {
    for (int e = lane; e < 144; e += 32) S.y7[e] = gy7[e];
    {   // state: struct copy through 4-byte words, coalesced
        const int* src = reinterpret_cast<const int*>(gst);
        int* dst = reinterpret_cast<int*>(&S.st);
        for (int e = lane; e < (int)(sizeof(RcRowState) / 4); e += 32) dst[e] = src[e];
    }
    __syncwarp();
    if (lane < RC_NJ) rc_r6d_to_mat(S.y7 + lane * 6, S.G[lane]);                                   // :173
    __syncwarp();
    if (lane < RC_NJ) {
        if (lane == 0) { for (int e = 0; e < 9; ++e) S.pose[0][e] = rcr[e]; }                      // pose[0] = Rcr (:175)
        else rc_mat3_tmul(S.G[M.parent[lane]], S.G[lane], S.pose[lane]);                           // IK (:174)
        if (lane == 0) { S.pb[0][0] = S.pb[0][1] = S.pb[0][2] = 0.f; S.jp[0][0] = S.jp[0][1] = S.jp[0][2] = 0.f; }
        else rc_mat3_vec(S.G[M.parent[lane]], M.bone[lane], S.pb[lane]);                           // :133
    }
    __syncwarp();
    for (int lev = 1; lev <= M.max_depth; ++lev) {                                                 // bone_vector_to_joint_position (:135)
        if (lane < RC_NJ && M.depth[lane] == lev)
            for (int r = 0; r < 3; ++r) S.jp[lane][r] = RC_ADD(S.jp[M.parent[lane]][r], S.pb[lane][r]);
        __syncwarp();
    }
    int need_init = 0;
    if (lane == 0) {
        if ((flags & RC_F_GE) && S.st.first_reach) { S.st.first_reach = 0; need_init = 1; }        // :178-183
        float pfoot[6];
        for (int f = 0; f < 2; ++f)                                                                // fk(poseg)[10:12].mm(Rcr.t())
            for (int j = 0; j < 3; ++j)
                pfoot[f * 3 + j] = S.jp[10 + f][0] * rcr[j * 3 + 0] + S.jp[10 + f][1] * rcr[j * 3 + 1] + S.jp[10 + f][2] * rcr[j * 3 + 2];
        float tran[3];
        int br = 0;
        rc_tran_update(cfg, &S.st, flags, pfoot, y8, vr, pc, rcr, conf, gravity, first_tran, tran, &br);
        if (branch) *branch = br | (need_init ? RC_BR_INIT : 0);
        S.tran[0] = tran[0]; S.tran[1] = tran[1]; S.tran[2] = tran[2];
    }
    need_init = __shfl_sync(0xffffffffu, need_init, 0);
    __syncwarp();
    const bool do_fk = (flags & RC_F_DO_FK) != 0;
    const bool need = do_fk && (cfg.live || (flags & RC_F_LATE));
    if (need) {
        // SMPL FK by tree level (model.py:209-241), then the 33 key points (sig_mp.py:287-299); same arithmetic as rc_fk_keypoints
        float L[12];
        if (lane < RC_NJ)
            for (int r = 0; r < 3; ++r) {
                L[r * 4 + 0] = S.pose[lane][r * 3 + 0]; L[r * 4 + 1] = S.pose[lane][r * 3 + 1]; L[r * 4 + 2] = S.pose[lane][r * 3 + 2];
                L[r * 4 + 3] = M.bone[lane][r];
            }
        if (lane == 0) for (int e = 0; e < 12; ++e) S.F[0][e] = L[e];
        __syncwarp();
        for (int lev = 1; lev <= M.max_depth; ++lev) {
            if (lane < RC_NJ && M.depth[lane] == lev) rc_rigid_mul(S.F[M.parent[lane]], L, S.F[lane]);
            __syncwarp();
        }
        if (lane < RC_NJ) {
            for (int r = 0; r < 3; ++r) S.joint[lane][r] = RC_ADD(S.F[lane][r * 4 + 3], S.tran[r]);
            for (int r = 0; r < 3; ++r) {
                const float d = S.F[lane][r * 4 + 0] * M.jrest[lane][0] + S.F[lane][r * 4 + 1] * M.jrest[lane][1] + S.F[lane][r * 4 + 2] * M.jrest[lane][2];
                S.F[lane][r * 4 + 3] = RC_SUB(S.F[lane][r * 4 + 3], d);
            }
        }
        __syncwarp();
        for (int k = lane; k < RC_NKP; k += 32) {
            if (M.kp_is_joint[k]) {
                for (int r = 0; r < 3; ++r) S.kp[k][r] = S.joint[M.kp_index[k]][r];
            } else {
                float Tv[12];
                for (int e = 0; e < 12; ++e) Tv[e] = 0.f;
                for (int j = 0; j < RC_NJ; ++j) {
                    const float w = M.kp_w[k][j];
                    for (int e = 0; e < 12; ++e) Tv[e] += w * S.F[j][e];
                }
                for (int r = 0; r < 3; ++r) {
                    const float p = Tv[r * 4 + 0] * M.kp_rest[k][0] + Tv[r * 4 + 1] * M.kp_rest[k][1] + Tv[r * 4 + 2] * M.kp_rest[k][2] + Tv[r * 4 + 3];
                    S.kp[k][r] = RC_ADD(p, S.tran[r]);
                }
            }
        }
        __syncwarp();
        if (cfg.live) {
            for (int e = lane; e < RC_NKP * 3; e += 32) S.st.j_temp[e / 3][e % 3] = S.kp[e / 3][e % 3];
            if (lane == 0) S.st.vision_count = cfg.update_vision_freq;
        }
        if (flags & RC_F_LATE) {                                                                   // :263-271
            for (int k = lane; k < RC_NKP; k += 32)
                for (int r = 0; r < 3; ++r) S.syn[k * 3 + r] = RC_DIV(S.kp[k][r], S.kp[k][2]);
            __syncwarp();
            for (int e = lane; e < 99; e += 32) x6[72 + e] = S.syn[e];
            for (int e = lane; e < 69; e += 32) x6[171 + e] = RC_SUB(S.joint[1 + e / 3][e % 3], S.joint[0][e % 3]);
            rc_normalise_kp_warp(S.syn, x4 + 72, lane);
        }
    } else if (!do_fk) {
        if (lane == 0) S.st.vision_count -= 1;
    }
    __syncwarp();
    {   // write back state, pose, tran
        int* dst = reinterpret_cast<int*>(gst);
        const int* src = reinterpret_cast<const int*>(&S.st);
        for (int e = lane; e < (int)(sizeof(RcRowState) / 4); e += 32) dst[e] = src[e];
    }
    for (int e = lane; e < 216; e += 32) gpose[e] = S.pose[e / 9][e % 9];
    if (lane < 3) gtran[lane] = S.tran[lane];
    return need_init;
}

struct RcPrepWarpSmem {
    float kp[99], acc[18], ori[54], xr[72], kpn[99];
};

// returns the flag word (uniform across the warp)
__device__ __forceinline__ int rc_prep_warp(const RcNetCfg& cfg, int vision_count, RcPrepWarpSmem& S, const float* __restrict__ pj,
                                            const float* __restrict__ pa, const float* __restrict__ po, int in_flags,
                                            float* x2, float* x3, float* x4, float* x6, float* x7, float* rcr, float* conf,
                                            float* lerpw, int lane) {
    for (int e = lane; e < 99; e += 32) S.kp[e] = pj[e];
    if (lane < 18) S.acc[lane] = pa[lane];
    for (int e = lane; e < 54; e += 32) S.ori[e] = po[e];
    __syncwarp();
    const float* R = S.ori + 45;                                                                   // Rcr = oric[-1] (:139)
    int f = 0;
    if (lane == 0) {
        const float cf = rc_conf_mean(S.kp);
        float lw[2];
        f = rc_prep_flags(cfg, vision_count, cf, in_flags, lw);
        conf[0] = cf; lerpw[0] = lw[0]; lerpw[1] = lw[1];
    }
    f = __shfl_sync(0xffffffffu, f, 0);
    if (lane < 9) rcr[lane] = R[lane];
    if (lane < 6) rc_vec_mat3(S.acc + lane * 3, R, S.xr + lane * 3);                               // accr = accc @ Rcr  (:142)
    else if (lane < 12) rc_mat3_tmul(R, S.ori + (lane - 6) * 9, S.xr + 18 + (lane - 6) * 9);       // orir = Rcr^T @ oric (:143)
    __syncwarp();
    if (f & RC_F_HI) rc_normalise_kp_warp(S.kp, S.kpn, lane);                                     // :150-152
    __syncwarp();
    for (int e = lane; e < 72; e += 32) {
        const float v = S.xr[e];
        x2[e] = v; x3[e] = v; x7[e] = v;
        const float c = (e < 18) ? S.acc[e] : S.ori[e - 18];
        x4[e] = c; x6[e] = c;
    }
    if (lane < RC_K2 - 72) x2[72 + lane] = 0.f;
    if (lane < RC_K3 - 141) { x3[141 + lane] = 0.f; x7[141 + lane] = 0.f; }
    if (lane < RC_K4 - 171) x4[171 + lane] = 0.f;
    for (int e = lane; e < 99; e += 32) {
        x6[72 + e] = S.kp[e];                                                                      // raw key points for rnn6 (:156)
        if (f & RC_F_HI) x4[72 + e] = S.kpn[e];
    }
    return f;
}
