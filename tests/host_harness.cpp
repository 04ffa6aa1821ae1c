// TEST-ONLY host build of the scalar per-stream logic (robustcap_b200/csrc/rc_rows.h, rc_math.h, rc_pack.h).
//
// There is no GPU in the build container, so the branch logic, the packed weight layout and the per-frame pass
// order of fusion.cu are exercised here on the CPU against the oracle (tests/test_host_logic.py).  This file is
// compiled by the test, lives under tests/, and is never loaded by the product: the shipped library has no CPU path.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../robustcap_b200/csrc/rc_rows.h"
#include "../robustcap_b200/csrc/rc_pack.h"

namespace {
const int kNetIn[6] = {72, 141, 171, 240, 141, 141};
const int kNetK1[6] = {RC_K2, RC_K3, RC_K4, RC_K6, RC_K7, RC_K7};
const int kNetH[6] = {512, 512, 1280, 1024, 512, 512};
const int kNetOut[6] = {69, 3, 69, 3, 144, 2};

struct HNet {
    int K1, H, out, out4;
    std::vector<float> W1, b1, WL[2], bL[2], W2, b2;
    std::vector<float> h[2], c[2];
};

struct Harness {
    RcModelConst M;
    RcNetCfg cfg;
    HNet nets[6];
    std::vector<float> Wi[3], bi[3];
    RcRowState st;
    float gravity[3];
    float sX4[RC_K4], sX6[RC_K6];   // persist between frames like the device buffers
};

// the semantics of RcLinear (rc_linear.cuh) on the packed layout, one row
void linear(const std::vector<float>& W, const std::vector<float>& b, int N, int K, const float* x, float* y, bool relu) {
    for (int n = 0; n < N; ++n) {
        float s = 0.f;
        const float* w = &W[(size_t)n * K];
        for (int k = 0; k < K; ++k) s = fmaf(w[k], x[k], s);
        s += b[n];
        y[n] = relu ? fmaxf(s, 0.f) : s;
    }
}
float sigm(float x) { return 1.f / (1.f + expf(-x)); }

void net_pass(HNet& n, const float* x, float* y) {
    std::vector<float> a1(n.H), xin(2 * n.H), hn(n.H);
    linear(n.W1, n.b1, n.H, n.K1, x, a1.data(), true);
    const float* in = a1.data();
    std::vector<float> hnew[2];
    for (int l = 0; l < 2; ++l) {
        memcpy(xin.data(), in, n.H * sizeof(float));
        memcpy(xin.data() + n.H, n.h[l].data(), n.H * sizeof(float));
        hnew[l].resize(n.H);
        for (int j = 0; j < n.H; ++j) {
            float g[4];
            for (int q = 0; q < 4; ++q) {
                float s = 0.f;
                const float* w = &n.WL[l][(size_t)(4 * j + q) * 2 * n.H];
                for (int k = 0; k < 2 * n.H; ++k) s = fmaf(w[k], xin[k], s);
                g[q] = s + n.bL[l][4 * j + q];
            }
            const float cn = fmaf(sigm(g[1]), n.c[l][j], sigm(g[0]) * tanhf(g[2]));
            n.c[l][j] = cn;
            hnew[l][j] = sigm(g[3]) * tanhf(cn);
        }
        in = hnew[l].data();
    }
    n.h[0] = hnew[0];
    n.h[1] = hnew[1];
    if (y) linear(n.W2, n.b2, n.out, n.H, hnew[1].data(), y, false);
}
}  // namespace

extern "C" {

void* hh_create(const float* joints, const float* verts, const float* skin_w, int nv, const int* parent, const int* mp_mask, int live) {
    Harness* h = new Harness();
    RcModelConst& C = h->M;
    for (int i = 0; i < RC_NJ; ++i) {
        C.parent[i] = i == 0 ? -1 : parent[i];
        for (int r = 0; r < 3; ++r) C.jrest[i][r] = joints[i * 3 + r];
    }
    for (int i = 0; i < RC_NJ; ++i)
        for (int r = 0; r < 3; ++r) C.bone[i][r] = i == 0 ? C.jrest[0][r] : (-C.jrest[C.parent[i]][r] + C.jrest[i][r]);
    C.max_depth = 0;
    for (int i = 0; i < RC_NJ; ++i) { C.depth[i] = i == 0 ? 0 : C.depth[C.parent[i]] + 1; if (C.depth[i] > C.max_depth) C.max_depth = C.depth[i]; }
    for (int k = 0; k < RC_NKP; ++k) {
        int j = -1;
        if (k >= 11 && k <= 16) j = 16 + (k - 11);
        else if (k == 23 || k == 24) j = 1 + (k - 23);
        else if (k == 25 || k == 26) j = 4 + (k - 25);
        else if (k == 27 || k == 28) j = 7 + (k - 27);
        C.kp_is_joint[k] = j >= 0;
        C.kp_index[k] = j >= 0 ? j : mp_mask[k];
        for (int r = 0; r < 3; ++r) C.kp_rest[k][r] = verts[mp_mask[k] * 3 + r];
        for (int q = 0; q < RC_NJ; ++q) C.kp_w[k][q] = skin_w[(size_t)mp_mask[k] * RC_NJ + q];
    }
    h->cfg.conf_lo = live ? 0.85 : 0.7; h->cfg.conf_hi = live ? 0.9 : 0.8; h->cfg.tran_filter = live ? 0.01 : 0.05;
    h->cfg.contact_thr = 0.7f; h->cfg.height_thr = 0.15f; h->cfg.dist_thr = 10.f; h->cfg.use_flat_floor = 1;
    h->cfg.live = live; h->cfg.update_vision_freq = 30;
    h->gravity[0] = -0.0029f; h->gravity[1] = 0.9980f; h->gravity[2] = -0.0273f;
    return h;
}
void hh_destroy(void* p) { delete (Harness*)p; }

// weights in state_dict order per net: w1,b1, (wih,whh,bih,bhh) x2, w2,b2
void hh_set_net(void* p, int ni, const float* w1, const float* b1, const float* wih0, const float* whh0, const float* bih0,
                const float* bhh0, const float* wih1, const float* whh1, const float* bih1, const float* bhh1,
                const float* w2, const float* b2) {
    Harness* h = (Harness*)p;
    HNet& n = h->nets[ni];
    n.K1 = kNetK1[ni]; n.H = kNetH[ni]; n.out = kNetOut[ni]; n.out4 = (n.out + 3) / 4 * 4;
    rc_pack_linear(w1, b1, n.H, kNetIn[ni], n.H, n.K1, n.W1, n.b1);
    rc_pack_lstm(wih0, whh0, bih0, bhh0, n.H, n.WL[0], n.bL[0]);
    rc_pack_lstm(wih1, whh1, bih1, bhh1, n.H, n.WL[1], n.bL[1]);
    rc_pack_linear(w2, b2, n.out, n.H, n.out4, n.H, n.W2, n.b2);
}
void hh_set_init(void* p, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2) {
    Harness* h = (Harness*)p;
    rc_pack_linear(w0, b0, 512, 69, 512, 80, h->Wi[0], h->bi[0]);
    rc_pack_linear(w1, b1, 1024, 512, 1024, 512, h->Wi[1], h->bi[1]);
    rc_pack_linear(w2, b2, 2048, 1024, 2048, 1024, h->Wi[2], h->bi[2]);
}
void hh_set_gravity(void* p, const float* g) { memcpy(((Harness*)p)->gravity, g, 12); }
void hh_reset(void* p) {
    Harness* h = (Harness*)p;
    for (int i = 0; i < 6; ++i)
        for (int l = 0; l < 2; ++l) { h->nets[i].h[l].assign(h->nets[i].H, 0.f); h->nets[i].c[l].assign(h->nets[i].H, 0.f); }
    rc_row_state_reset(&h->st, true);
    memset(h->sX4, 0, sizeof(h->sX4)); memset(h->sX6, 0, sizeof(h->sX6));
}
int hh_floor_n(void* p) { return ((Harness*)p)->st.floor_n; }

// one frame, same pass order as enqueue_step() in fusion.cu
void hh_step(void* p, const float* j2dc, const float* accc, const float* oric, int in_flags, const float* first_tran,
             float* pose, float* tran) {
    Harness* h = (Harness*)p;
    float X2[RC_K2], X3[RC_K3], X4[RC_K4], X6[RC_K6], X7[RC_K7], rcr[9], conf, lerpw[2];
    float Y3[4] = {0}, Y6[4] = {0}, Y7[144], Y8[4] = {0};
    memcpy(X4, h->sX4, sizeof(X4)); memcpy(X6, h->sX6, sizeof(X6));
    const int f = rc_prep_row(h->cfg, h->st, j2dc, accc, oric, in_flags | RC_F_ACTIVE, X2, X3, X4, X6, X7, rcr, &conf, lerpw);
    float y[144];
    net_pass(h->nets[0], X2, y); memcpy(X3 + 72, y, 69 * 4);
    net_pass(h->nets[1], X3, Y3);
    if (f & RC_F_HI) { net_pass(h->nets[2], X4, y); memcpy(X6 + 171, y, 69 * 4); }
    if (f & RC_F_FIRST_FRAME) net_pass(h->nets[3], X6, Y6);
    if (f & RC_F_R6B) net_pass(h->nets[3], X6, Y6);
    rc_mid_row(f, rcr, lerpw, X3 + 72, X6 + 171, X7 + 72);
    net_pass(h->nets[4], X7, Y7);
    net_pass(h->nets[5], X7, Y8);
    const int need_init = rc_kin_row(h->cfg, h->M, &h->st, f, Y7, Y8, Y3, Y6, rcr, conf, h->gravity, first_tran, pose, tran, X4, X6);
    if (need_init) {
        float xi[80] = {0}, i1[512], i2[1024], i3[2048];
        memcpy(xi, X7 + 72, 69 * 4);
        linear(h->Wi[0], h->bi[0], 512, 80, xi, i1, true);
        linear(h->Wi[1], h->bi[1], 1024, 512, i1, i2, true);
        linear(h->Wi[2], h->bi[2], 2048, 1024, i2, i3, false);
        HNet& n = h->nets[0];
        for (int u = 0; u < 512; ++u) { n.h[0][u] = i3[u]; n.h[1][u] = i3[512 + u]; n.c[0][u] = i3[1024 + u]; n.c[1][u] = i3[1536 + u]; }
    }
    if (f & RC_F_LATE) { net_pass(h->nets[3], X6, nullptr); net_pass(h->nets[2], X4, nullptr); }
    memcpy(h->sX4, X4, sizeof(X4)); memcpy(h->sX6, X6, sizeof(X6));
}

// element-wise checks of rc_math.h
void hh_r6d_to_mat(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_r6d_to_mat(x + i * 6, o + i * 9); }
void hh_aa_to_mat(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_aa_to_mat(x + i * 3, o + i * 9); }
void hh_batch_rodrigues(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_batch_rodrigues(x + i * 3, o + i * 9); }
void hh_mat_to_aa(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_mat_to_aa(x + i * 9, o + i * 3); }
void hh_quat_to_mat(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_quat_to_mat(x + i * 4, o + i * 9); }
void hh_quat_to_aa(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_quat_to_aa(x + i * 4, o + i * 3); }
void hh_aa_to_quat(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) rc_aa_to_quat(x + i * 3, o + i * 4); }
void hh_quat_mul(const float* a, const float* b, float* o, int n) { for (int i = 0; i < n; ++i) rc_quat_mul(a + i * 4, b + i * 4, o + i * 4); }
float hh_conf_mean(const float* kp) { return rc_conf_mean(kp); }
void hh_fk_keypoints(void* p, const float* pose, const float* tran, float* joint, float* kp) {
    rc_fk_keypoints(((Harness*)p)->M, pose, tran, joint, kp);
}
}
