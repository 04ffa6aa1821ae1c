// Per-stream, per-frame scalar logic of Net.forward_online (net/sig_mp.py:113-274) around the six LSTM stacks:
//   rc_prep_row  — confidence, IMU change of frame, key-point normalisation, sub-net input assembly (:138-153)
//   rc_mid_row   — camera->root rotation of the vision joints and the confidence lerp (:154-167)
//   rc_kin_row   — 6D -> R, IK, foot FK, translation / contact / floor state machine, SMPL FK + 33 synthetic
//                  MediaPipe points, vision-updater inputs (:173-273)
// Written as scalar host/device functions: the CUDA kernels run one stream per thread; tests/host_harness.cpp
// compiles the same code for the CPU to check the branch logic without a GPU (test-only).
#pragma once
#include "rc_math.h"

#define RC_NJ 24
#define RC_NKP 33
#define RC_FLOOR_CAP 11

// padded input widths of the sub-nets (K padded to a multiple of 16, zero weights in the pad)
#define RC_K2 80     // 72
#define RC_K3 144    // 141
#define RC_K4 176    // 171
#define RC_K6 240    // 240
#define RC_K7 144    // 141 (rnn7 and rnn8 share the input)

enum RcFlag {
    RC_F_FIRST_FRAME = 1,      // caller's first_frame
    RC_F_FIRST_TRAN  = 2,      // caller passed first_tran
    RC_F_HI          = 4,      // c > lo or first_frame : rnn4 runs on the real key points   (:149)
    RC_F_GE          = 8,      // c >= hi                                                    (:159)
    RC_F_MID         = 16,     // lo < c < hi                                                (:162)
    RC_F_R6B         = 32,     // c > lo : rnn6 runs (second time on a first_frame)          (:161,165)
    RC_F_LATE        = 64,     // vision updater runs rnn6 + rnn4 on synthetic key points    (:264)
    RC_F_DO_FK       = 128,    // mesh FK this frame (always unless live)                    (:229-242)
    RC_F_ACTIVE      = 256,    // stream still has frames (ragged batches)
};

// bits of the optional per-frame branch log (debug / parity tests: which data-dependent decision a frame took, :185-225)
enum RcBranch {
    RC_BR_CONTACT = 1, RC_BR_ARGMAX = 2, RC_BR_SNAP = 4, RC_BR_LERP = 8, RC_BR_FLOOR_ADD = 16, RC_BR_FLOOR_P1 = 32,
    RC_BR_FLOOR_P0 = 64, RC_BR_INIT = 128,
};

struct RcModelConst {                 // SMPL constants the per-frame path needs (articulate/model.py:29-39)
    int parent[RC_NJ];                // parent[0] = -1
    int depth[RC_NJ];                 // tree depth of each joint (root 0); max_depth = deepest level
    int max_depth;
    float jrest[RC_NJ][3];            // zero-pose joints, root at origin (model.py:87)
    float bone[RC_NJ][3];             // bone[i] = -jrest[parent] + jrest[i]   (spatial.py:148-167)
    int kp_is_joint[RC_NKP];          // the 33 synthetic MediaPipe points (sig_mp.py:287-299)
    int kp_index[RC_NKP];
    float kp_rest[RC_NKP][3];         // rest vertex, root at origin
    float kp_w[RC_NKP][RC_NJ];        // skinning weights of that vertex
};

struct RcNetCfg {                     // class-level knobs of Net (sig_mp.py:27-45, 91-93)
    double conf_lo, conf_hi;          // conf_range
    double tran_filter;               // tran_filter_num
    float contact_thr;                // contact_threshold, compared in float32 like torch does
    float height_thr;                 // height_threhold
    float dist_thr;                   // distrance_threshold
    int use_flat_floor;
    int live;
    int update_vision_freq;
};

struct RcRowState {                   // mutable per-stream state other than the LSTM (h, c)   (sig_mp.py:85-90)
    float last_pfoot[6];
    float last_tran[3];
    float floor_y[RC_FLOOR_CAP][3];
    float j_temp[RC_NKP][3];          // live mode: cached synthetic points
    float joint_temp[RC_NJ][3];
    int has_last;                     // last_pfoot / last_tran are set
    int floor_n;
    int first_reach;
    int vision_count;                 // update_vision_count
};

// Net.reset_states (sig_mp.py:95-104).  update_vision_count is NOT among the fields the reference resets: it keeps counting
// across sequences, so it is only zeroed when the state is created.
RC_HD void rc_row_state_reset(RcRowState* s, bool created = false) {
    s->has_last = 0; s->floor_n = 0; s->first_reach = 1;
    if (created) s->vision_count = 0;
}

// torch's float32 mean over the 33 strided confidences: 4 interleaved partial sums, combined left to right
// (measured against torch 2.11 CPU, see tests/test_host_logic.py), then one division by 33.
RC_HD float rc_conf_mean(const float* j2dc) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < RC_NKP; ++i) acc[i & 3] = RC_ADD(acc[i & 3], j2dc[i * 3 + 2]);
    float s = RC_ADD(RC_ADD(RC_ADD(acc[0], acc[1]), acc[2]), acc[3]);
    return RC_DIV(s, 33.f);
}

// sig_mp.py:150-152 / 268-270.  kp: [33][3] in, out: [33][3]
RC_HD void rc_normalise_kp(const float* kp, float* out) {
    float umax = kp[0], umin = kp[0], vmax = kp[1], vmin = kp[1];
    for (int i = 1; i < RC_NKP; ++i) {
        umax = fmaxf(umax, kp[i * 3]); umin = fminf(umin, kp[i * 3]);
        vmax = fmaxf(vmax, kp[i * 3 + 1]); vmin = fminf(vmin, kp[i * 3 + 1]);
    }
    float sc = fmaxf(RC_SUB(umax, umin), RC_SUB(vmax, vmin));
    float rx = RC_DIV(kp[23 * 3], sc), ry = RC_DIV(kp[23 * 3 + 1], sc);
    for (int i = 0; i < RC_NKP; ++i) {
        float x = RC_DIV(kp[i * 3], sc), y = RC_DIV(kp[i * 3 + 1], sc);
        if (i != 23) { x = RC_SUB(x, rx); y = RC_SUB(y, ry); }
        out[i * 3] = x; out[i * 3 + 1] = y; out[i * 3 + 2] = kp[i * 3 + 2];
    }
}

// Branch flags of one frame from the mean confidence (sig_mp.py:149-167, 229-242, 264) + the lerp weights (:163-164).
RC_HD int rc_prep_flags(const RcNetCfg& cfg, int vision_count, float cf, int in_flags, float* lerpw) {
    const double c = (double)cf;                             // .item() -> python float     (:138)
    int f = in_flags & (RC_F_FIRST_FRAME | RC_F_FIRST_TRAN | RC_F_ACTIVE);
    const bool ff = (in_flags & RC_F_FIRST_FRAME) != 0;
    if (c > cfg.conf_lo || ff) f |= RC_F_HI;                                          // :149
    if (c >= cfg.conf_hi) f |= RC_F_GE;
    else if (c > cfg.conf_lo) f |= RC_F_MID;
    if (c > cfg.conf_lo) f |= RC_F_R6B;
    const bool do_fk = (!cfg.live) || vision_count == 0;                              // :229-242
    if (do_fk) f |= RC_F_DO_FK;
    if (c <= cfg.conf_lo && do_fk) f |= RC_F_LATE;                                    // :264
    const double k = (c - cfg.conf_lo) / (cfg.conf_hi - cfg.conf_lo);                 // :163
    lerpw[0] = (float)(1.0 - k); lerpw[1] = (float)k;
    return f;
}

// ---- prep ---------------------------------------------------------------------------------------------------
// Outputs (row pointers into the [B, K] sub-net input buffers): x2[80], x3[144], x4[176], x6[240], x7[144];
// rcr[9]; conf[0] = c; lerpw[2] = (float)(1-k), (float)k; returns the flag word.
RC_HD int rc_prep_row(const RcNetCfg& cfg, const RcRowState& st, const float* j2dc, const float* accc,
                      const float* oric, int in_flags, float* x2, float* x3, float* x4, float* x6, float* x7,
                      float* rcr, float* conf, float* lerpw) {
    float cf = rc_conf_mean(j2dc);
    const float* R = oric + 5 * 9;                           // Rcr = oric[-1]              (:139)
    for (int i = 0; i < 9; ++i) rcr[i] = R[i];
    float xr[72], *xc = x4;
    for (int i = 0; i < 6; ++i) rc_vec_mat3(accc + i * 3, R, xr + i * 3);            // accr = accc @ Rcr  (:142)
    for (int i = 0; i < 6; ++i) rc_mat3_tmul(R, oric + i * 9, xr + 18 + i * 9);      // orir = Rcr^T @ oric (:143)
    for (int i = 0; i < 72; ++i) { x2[i] = xr[i]; x3[i] = xr[i]; x7[i] = xr[i]; }
    for (int i = 72; i < RC_K2; ++i) x2[i] = 0.f;
    for (int i = 141; i < RC_K3; ++i) { x3[i] = 0.f; x7[i] = 0.f; }
    for (int i = 0; i < 18; ++i) { xc[i] = accc[i]; x6[i] = accc[i]; }
    for (int i = 0; i < 54; ++i) { xc[18 + i] = oric[i]; x6[18 + i] = oric[i]; }
    for (int i = 0; i < 99; ++i) x6[72 + i] = j2dc[i];                                // raw key points for rnn6 (:156)
    for (int i = 171; i < RC_K4; ++i) x4[i] = 0.f;

    const int f = rc_prep_flags(cfg, st.vision_count, cf, in_flags, lerpw);
    if (f & RC_F_HI) rc_normalise_kp(j2dc, x4 + 72);                                   // :150-152
    conf[0] = cf;
    return f;
}

// ---- mid ----------------------------------------------------------------------------------------------------
// x3 + 72 holds j3dr_i (rnn2 out), x6 + 171 holds j3dc (rnn4 out, camera frame).  Writes j3dr into x7 + 72.
RC_HD void rc_mid_joint(int flags, const float* rcr, const float* lerpw, const float* j3dr_i, const float* j3dc, float* j3dr) {
    float v[3];
    if (flags & (RC_F_GE | RC_F_MID)) rc_vec_mat3(j3dc, rcr, v);                          // j3dc.view(23,3).mm(Rcr) (:154)
    for (int j = 0; j < 3; ++j) {
        const float a = j3dr_i[j];
        if (flags & RC_F_GE) j3dr[j] = v[j];
        else if (flags & RC_F_MID) j3dr[j] = RC_ADD(RC_MUL(a, lerpw[0]), RC_MUL(v[j], lerpw[1]));     // lerp (:164)
        else j3dr[j] = a;
    }
}
RC_HD void rc_mid_row(int flags, const float* rcr, const float* lerpw, const float* j3dr_i, const float* j3dc,
                      float* j3dr) {
    for (int i = 0; i < 23; ++i) rc_mid_joint(flags, rcr, lerpw, j3dr_i + i * 3, j3dc + i * 3, j3dr + i * 3);
}

// ---- kin ----------------------------------------------------------------------------------------------------
// SMPL FK (model.py:209-241) restricted to what sync_mp3d reads: 24 joints + 21 skinned vertices.
// pose: local rotations [24][9]; out: joint[24][3] (with tran), kp[33][3] (with tran)
RC_HD void rc_fk_keypoints(const RcModelConst& M, const float* pose, const float* tran, float* joint, float* kp) {
    float G[RC_NJ][12];
    for (int i = 0; i < RC_NJ; ++i) {
        float L[12];
        for (int r = 0; r < 3; ++r) {
            L[r * 4 + 0] = pose[i * 9 + r * 3 + 0]; L[r * 4 + 1] = pose[i * 9 + r * 3 + 1];
            L[r * 4 + 2] = pose[i * 9 + r * 3 + 2]; L[r * 4 + 3] = M.bone[i][r];
        }
        if (i == 0) for (int e = 0; e < 12; ++e) G[0][e] = L[e];
        else rc_rigid_mul(G[M.parent[i]], L, G[i]);
    }
    for (int i = 0; i < RC_NJ; ++i)
        for (int r = 0; r < 3; ++r) joint[i * 3 + r] = RC_ADD(G[i][r * 4 + 3], tran[r]);
    // T[..., -1:] -= T @ [j; 0]   (model.py:235)
    for (int i = 0; i < RC_NJ; ++i)
        for (int r = 0; r < 3; ++r) {
            float d = G[i][r * 4 + 0] * M.jrest[i][0] + G[i][r * 4 + 1] * M.jrest[i][1] + G[i][r * 4 + 2] * M.jrest[i][2];
            G[i][r * 4 + 3] = RC_SUB(G[i][r * 4 + 3], d);
        }
    for (int k = 0; k < RC_NKP; ++k) {
        if (M.kp_is_joint[k]) {
            for (int r = 0; r < 3; ++r) kp[k * 3 + r] = joint[M.kp_index[k] * 3 + r];
        } else {
            float Tv[12];
            for (int e = 0; e < 12; ++e) Tv[e] = 0.f;
            for (int j = 0; j < RC_NJ; ++j) {
                float w = M.kp_w[k][j];
                for (int e = 0; e < 12; ++e) Tv[e] += w * G[j][e];
            }
            for (int r = 0; r < 3; ++r) {
                float p = Tv[r * 4 + 0] * M.kp_rest[k][0] + Tv[r * 4 + 1] * M.kp_rest[k][1] +
                          Tv[r * 4 + 2] * M.kp_rest[k][2] + Tv[r * 4 + 3];
                kp[k * 3 + r] = RC_ADD(p, tran[r]);
            }
        }
    }
}

RC_HD void rc_floor_point(const float* pf, const float* tran, const float* g, float* p) {   // dot(pfoot + tran, g) * g
    float a0 = RC_ADD(pf[0], tran[0]), a1 = RC_ADD(pf[1], tran[1]), a2 = RC_ADD(pf[2], tran[2]);
    float d = RC_ADD(RC_ADD(RC_MUL(a0, g[0]), RC_MUL(a1, g[1])), RC_MUL(a2, g[2]));
    p[0] = RC_MUL(d, g[0]); p[1] = RC_MUL(d, g[1]); p[2] = RC_MUL(d, g[2]);
}

// Translation / contact / floor state machine of one frame (sig_mp.py:185-227, 273): returns tran and updates the state.
RC_HD void rc_tran_update(const RcNetCfg& cfg, RcRowState* st, int flags, const float* pfoot, const float* y8, const float* vr,
                          const float* pc, const float* rcr, float conf, const float* gravity, const float* first_tran,
                          float* tran_out, int* branch = nullptr) {
    const double c = (double)conf;
    int br = 0;
    float ct0 = 1.f / (1.f + expf(-y8[0])), ct1 = 1.f / (1.f + expf(-y8[1]));           // sigmoid (:170)
    float cmax = fmaxf(ct0, ct1);
    int carg = (ct1 > ct0) ? 1 : 0;                                                      // argmax, first max wins
    float v[3];
    if (cmax < cfg.contact_thr || !st->has_last) {                                       // :187-188
        float rv[3];
        rc_mat3_vec(rcr, vr, rv);
        for (int r = 0; r < 3; ++r) v[r] = RC_DIV(RC_MUL(rv[r], 3.f), 60.f);             // * vel_scale / 60
    } else {
        for (int r = 0; r < 3; ++r) v[r] = RC_SUB(st->last_pfoot[carg * 3 + r], pfoot[carg * 3 + r]);   // :190
        br |= RC_BR_CONTACT | (carg ? RC_BR_ARGMAX : 0);
    }
    float tran[3];
    for (int r = 0; r < 3; ++r) tran[r] = st->has_last ? RC_ADD(st->last_tran[r], v[r]) : v[r];       // :191-194

    if (flags & RC_F_GE) {                                                               // :196-203
        double k = (c - cfg.conf_lo) / (cfg.conf_hi - cfg.conf_lo);
        if (k > 1) k = 1;
        float d[3] = {RC_SUB(pc[0], tran[0]), RC_SUB(pc[1], tran[1]), RC_SUB(pc[2], tran[2])};
        if (rc_norm3(d) > cfg.dist_thr || cfg.tran_filter > 1) {
            for (int r = 0; r < 3; ++r) tran[r] = pc[r];
            br |= RC_BR_SNAP;
        } else {
            br |= RC_BR_LERP;
            double w = cfg.tran_filter * k;
            float wa = (float)(1.0 - w), wb = (float)w;
            for (int r = 0; r < 3; ++r) tran[r] = RC_ADD(RC_MUL(tran[r], wa), RC_MUL(pc[r], wb));
        }
    }

    bool ff = (flags & RC_F_FIRST_FRAME) != 0, ft = (flags & RC_F_FIRST_TRAN) != 0;
    if (st->floor_n < RC_FLOOR_CAP && !ff && !ft && cmax > cfg.contact_thr && cfg.use_flat_floor && (flags & RC_F_GE)) {   // :208-214
        float p0[3], p1[3];
        rc_floor_point(pfoot, tran, gravity, p0);
        rc_floor_point(pfoot + 3, tran, gravity, p1);
        const float* p = (rc_norm3(p0) < rc_norm3(p1)) ? p1 : p0;
        for (int r = 0; r < 3; ++r) st->floor_y[st->floor_n][r] = p[r];
        st->floor_n += 1;
        br |= RC_BR_FLOOR_ADD;
    }
    if (cfg.use_flat_floor && st->floor_n > 10 && cmax > cfg.contact_thr) {             // :215-221
        float p0[3], p1[3], m[3], d0[3], d1[3];
        rc_floor_point(pfoot, tran, gravity, p0);
        rc_floor_point(pfoot + 3, tran, gravity, p1);
        for (int r = 0; r < 3; ++r) {                                                    // sum(floor_y[-6:]) / 6, python sum starts at 0
            float s = 0.f;
            for (int q = st->floor_n - 6; q < st->floor_n; ++q) s = RC_ADD(s, st->floor_y[q][r]);
            m[r] = RC_DIV(s, 6.f);
            d0[r] = RC_SUB(m[r], p0[r]); d1[r] = RC_SUB(m[r], p1[r]);
        }
        if (rc_norm3(p0) < rc_norm3(p1) && rc_norm3(d1) < cfg.height_thr) {
            for (int r = 0; r < 3; ++r) tran[r] = RC_ADD(tran[r], d1[r]);
            br |= RC_BR_FLOOR_P1;
        } else if (rc_norm3(d0) < cfg.height_thr) {
            for (int r = 0; r < 3; ++r) tran[r] = RC_ADD(tran[r], d0[r]);
            br |= RC_BR_FLOOR_P0;
        }
    }
    if (ft) { for (int r = 0; r < 3; ++r) tran[r] = first_tran[r]; }                     // :222-225
    else if (ff) { for (int r = 0; r < 3; ++r) tran[r] = pc[r]; }

    for (int e = 0; e < 6; ++e) st->last_pfoot[e] = pfoot[e];                            // :227
    for (int r = 0; r < 3; ++r) { st->last_tran[r] = tran[r]; tran_out[r] = tran[r]; }   // :273
    st->has_last = 1;
    if (branch) *branch = br;
}

// Inputs: y7[144] (6D global pose), y8[2] (contact logits), vr[3] (rnn3), pc[3] (rnn6 early result, valid when
// GE or FIRST_FRAME), rcr[9], conf, gravity[3], first_tran[3].
// Outputs: pose[216], tran[3]; when RC_F_LATE: x6/x4 rows rewritten with the synthetic key points;
// returns 1 when this stream must re-seed rnn2's state through init_net (:178-183) with j3dr (= x7 + 72).
RC_HD int rc_kin_row(const RcNetCfg& cfg, const RcModelConst& M, RcRowState* st, int flags, const float* y7,
                     const float* y8, const float* vr, const float* pc, const float* rcr, float conf,
                     const float* gravity, const float* first_tran, float* pose, float* tran_out, float* x4,
                     float* x6, int* branch = nullptr) {
    float G[RC_NJ][9];
    for (int i = 0; i < RC_NJ; ++i) rc_r6d_to_mat(y7 + i * 6, G[i]);                   // :173
    for (int e = 0; e < 9; ++e) pose[e] = rcr[e];                                        // pose[0] = Rcr (:175)
    for (int i = 1; i < RC_NJ; ++i) rc_mat3_tmul(G[M.parent[i]], G[i], pose + i * 9);    // IK (:174)

    int need_init = 0;
    if ((flags & RC_F_GE) && st->first_reach) { st->first_reach = 0; need_init = 1; }    // :178-183

    // foot FK from global rotations and rest bones (:131-135, :186)
    float jp[RC_NJ][3];
    jp[0][0] = jp[0][1] = jp[0][2] = 0.f;
    for (int i = 1; i < RC_NJ; ++i) {
        float pb[3];
        rc_mat3_vec(G[M.parent[i]], M.bone[i], pb);
        for (int r = 0; r < 3; ++r) jp[i][r] = RC_ADD(jp[M.parent[i]][r], pb[r]);
    }
    float pfoot[6];
    for (int f = 0; f < 2; ++f)                                                          // fk(poseg)[10:12].mm(Rcr.t())
        for (int j = 0; j < 3; ++j)
            pfoot[f * 3 + j] = jp[10 + f][0] * rcr[j * 3 + 0] + jp[10 + f][1] * rcr[j * 3 + 1] + jp[10 + f][2] * rcr[j * 3 + 2];

    float tran[3];
    int br = 0;
    rc_tran_update(cfg, st, flags, pfoot, y8, vr, pc, rcr, conf, gravity, first_tran, tran, &br);
    if (branch) *branch = br | (need_init ? RC_BR_INIT : 0);
    for (int r = 0; r < 3; ++r) tran_out[r] = tran[r];

    // :228-242 mesh FK -> synthetic key points; live mode recomputes every (freq+1)-th frame only
    if (flags & RC_F_DO_FK) {
        float joint[RC_NJ * 3], kp[RC_NKP * 3];
        bool need = cfg.live || (flags & RC_F_LATE);
        if (need) {
            rc_fk_keypoints(M, pose, tran, joint, kp);
            if (cfg.live) {
                for (int e = 0; e < RC_NKP * 3; ++e) st->j_temp[e / 3][e % 3] = kp[e];
                st->vision_count = cfg.update_vision_freq;
            }
        }
        if (flags & RC_F_LATE) {                                                         // :263-271
            float syn[RC_NKP * 3];
            for (int k = 0; k < RC_NKP; ++k)
                for (int r = 0; r < 3; ++r) syn[k * 3 + r] = RC_DIV(kp[k * 3 + r], kp[k * 3 + 2]);   // j / j[:, 2:]
            for (int e = 0; e < 99; ++e) x6[72 + e] = syn[e];
            for (int i = 1; i < RC_NJ; ++i)
                for (int r = 0; r < 3; ++r) x6[171 + (i - 1) * 3 + r] = RC_SUB(joint[i * 3 + r], joint[r]);   // joint[1:] - joint[:1]
            rc_normalise_kp(syn, x4 + 72);
        }
    } else {
        st->vision_count -= 1;
    }
    return need_init;
}
