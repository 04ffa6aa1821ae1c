// Single-stream (B = 1) frame of Net.forward_online (net/sig_mp.py:113-274) as ONE cooperative kernel with the weights TMA-staged
// into shared memory — the latency path of BASELINE configs[1].
//
// At B = 1 every weight byte (243 MB) is used exactly once per frame, so the frame is HBM-bound: 37 us at the measured copy
// bandwidth.  The multi-kernel path (CUDA graph of ~60 GEMV launches) and the first cooperative kernel (stream.cu) both sit at
// 120-150 us because each of their ~18 dependent steps pays launch / barrier latency with the memory pipe idle.  Here:
//
//   * one 288-thread CTA per SM for the whole frame; warp 8 is a PRODUCER that walks the frame's static list of weight slices
//     (this CTA's rows of every layer, in the order they are needed) and keeps a 5 x 40 KB shared-memory ring full with
//     cp.async.bulk copies (mbarrier complete_tx) — it never waits for data dependencies, only for free ring slots, so HBM keeps
//     streaming across the grid barriers;
//   * the 8 consumer warps compute GEMV rows out of shared memory (one hidden unit = the 4 gate rows of the gate-interleaved
//     packing per warp pass, 128-bit loads, warp-shuffle reduction) and apply the fused epilogue (bias, relu / LSTM cell update);
//   * the recurrent halves W_hh . h_{t-1} of all LSTM layers (half of all weight bytes) depend only on the previous frame: they are
//     scheduled BETWEEN a layer's "arrive" and the matching "wait" of the split-phase grid barrier, so the dependency latency of the
//     input halves W_ih . x is covered by useful streaming work;
//   * prep (:138-153) and the joint blend (:154-167) are recomputed by every CTA (a few hundred flops) instead of being broadcast
//     through another barrier; kin (:173-273) runs on CTA 0.
// Grid barriers per frame: 8 (+1 kin + 2 for the vision updater on occluded frames).  Hidden states are double-buffered (h / hn), the
// host swaps the pointers after every frame.  First frames with first_frame = True (double rnn6 pass) stay on the multi-launch path.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "rc_fusion.cuh"
#include "rc_rows_warp.cuh"
#include "rc_tc_dev.cuh"

namespace {

constexpr int kS2Cons = 8;                          // consumer warps
constexpr int kS2Threads = (kS2Cons + 1) * 32;
constexpr int kS2SlotBytes = 20480;                 // one hidden unit (4 gate rows) of the widest layer (K = 1280)
constexpr int kS2Slots = 8;
constexpr int kS2MaxUnits = 12;                     // units (4 rows) of one matrix owned by one CTA
constexpr int kS2MaxGroups = 24;
constexpr int kS2XMax = 1280;

enum { PK_LIN1 = 0, PK_LH0, PK_LX0, PK_LH1, PK_LX1, PK_LIN2, PK_KIN };
enum { PRE_NONE = 0, PRE_BLEND = 1 };

struct S2Mat { const float* W; const float* bias; int units; int K; };      // rows = 4 * units, row-major [rows, K]
struct S2Net {
    S2Mat lin1, lx[2], lh[2], lin2;
    float *h[2], *hn[2], *c[2], *a1;
    float* yout;
    int H, out;
};
struct S2Args {
    RcNetCfg cfg;
    const RcModelConst* M;
    RcRowState* row;
    StepIO io;
    int t;                                          // frame index inside io (sequence mode), 0 for the staged single frame
    S2Net net[NNETS];
    const float* Wi[3]; const float* bi[3];
    float *X4, *X6, *X7, *Y3, *Y6, *Y7, *Y8, *I1, *I2, *gravity;
    int* gflags;                                    // [0] need_init (written by CTA 0)
    unsigned* bar;                                  // [2][16] grid-barrier words, one set per frame parity
    int parity;
    int program;                                    // order of the frame's groups (s2_build)
    unsigned long long* ts;
};
struct S2Group { int kind, mask, wait_before, arrive_after, pre, late; };

struct S2Smem {
    RcPrepWarpSmem prep;
    RcKinWarpSmem kin;
    RcModelConst M;
    alignas(16) float x2[RC_K2];
    alignas(16) float x3[RC_K3];
    alignas(16) float x4[RC_K4];
    alignas(16) float x6[RC_K6];
    alignas(16) float x7[RC_K7];
    alignas(16) float xbuf[2][4][kS2XMax];           // input vectors of the (up to 4) matrices of a group, double-buffered by group
    alignas(16) float hp[NNETS][2][kS2MaxUnits][4];
    float rcr[12], lerpw[2], conf;
    S2Group groups[kS2MaxGroups];
    int ngroups, flags;
    alignas(8) unsigned long long full[kS2Slots];
    alignas(8) unsigned long long empty[kS2Slots];
};

__device__ __forceinline__ unsigned s2_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void s2_cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kS2Cons * 32) : "memory"); }
__device__ __forceinline__ void s2_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void s2_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// this CTA's units of a matrix: contiguous block of ceil(units / grid) units
__device__ __forceinline__ void s2_slice(const S2Mat& m, int& u0, int& nu) {
    const int per = (m.units + (int)gridDim.x - 1) / (int)gridDim.x;
    u0 = (int)blockIdx.x * per;
    nu = max(0, min(per, m.units - u0));
}
__device__ __forceinline__ const S2Mat& s2_mat(const S2Net& n, int kind) {
    switch (kind) {
        case PK_LIN1: return n.lin1;
        case PK_LH0: return n.lh[0];
        case PK_LX0: return n.lx[0];
        case PK_LH1: return n.lh[1];
        case PK_LX1: return n.lx[1];
        default: return n.lin2;
    }
}
__device__ __forceinline__ int s2_units_per_chunk(int K) { return max(1, kS2SlotBytes / (K * 16)); }

// The frame's program: the same list drives the producer (which slices to stream, in which order) and the consumers.
__device__ int s2_build(int f, S2Group* g, int program) {
    const bool hi = (f & RC_F_HI) != 0, r6b = (f & RC_F_R6B) != 0, late = (f & RC_F_LATE) != 0;
    const int p1 = (hi ? (1 << NET4) : 0) | (1 << NET2);
    const int p2 = (r6b ? (1 << NET6) : 0) | (1 << NET3) | (1 << NET7) | (1 << NET8);
    const int pl = (1 << NET4) | (1 << NET6);
    int n = 0;
    auto add = [&](int kind, int mask, int wait, int arrive, int pre, int lt) { g[n].kind = kind; g[n].mask = mask; g[n].wait_before = wait; g[n].arrive_after = arrive; g[n].pre = pre; g[n].late = lt; ++n; };
    // The recurrent halves (LH*, no dependency inside the frame) sit between an arrive and the matching wait of the grid barriers, so the
    // dependency latency is covered by streaming.  program 1 (default) spreads the second phase's halves — rnn6 apart from the three
    // H = 512 nets — over four barrier gaps instead of two; program 0 is the former order (A/B: RC_S2_PROGRAM).
    const int p2a = r6b ? (1 << NET6) : 0, p2b = (1 << NET3) | (1 << NET7) | (1 << NET8);
    add(PK_LIN1, p1, 0, 1, PRE_NONE, 0);
    add(PK_LH0, p1, 0, 0, PRE_NONE, 0);
    add(PK_LX0, p1, 1, 1, PRE_NONE, 0);
    add(PK_LH1, p1, 0, 0, PRE_NONE, 0);
    add(PK_LX1, p1, 1, 1, PRE_NONE, 0);
    if (program == 0 || !p2a) {
        add(PK_LH0, p2, 0, 0, PRE_NONE, 0);
        add(PK_LIN2, p1, 1, 1, PRE_NONE, 0);
        add(PK_LH1, p2, 0, 0, PRE_NONE, 0);
        add(PK_LIN1, p2, 1, 1, PRE_BLEND, 0);
        add(PK_LX0, p2, 1, 1, PRE_NONE, 0);
    } else {
        add(PK_LH0, p2a, 0, 0, PRE_NONE, 0);
        add(PK_LIN2, p1, 1, 1, PRE_NONE, 0);
        add(PK_LH0, p2b, 0, 0, PRE_NONE, 0);
        add(PK_LIN1, p2, 1, 1, PRE_BLEND, 0);
        add(PK_LH1, p2a, 0, 0, PRE_NONE, 0);
        add(PK_LX0, p2, 1, 1, PRE_NONE, 0);
        add(PK_LH1, p2b, 0, 0, PRE_NONE, 0);
    }
    add(PK_LX1, p2, 1, 1, PRE_NONE, 0);
    add(PK_LIN2, p2, 1, 1, PRE_NONE, 0);
    if (late) { add(PK_LH0, pl, 0, 0, PRE_NONE, 1); add(PK_LH1, pl, 0, 0, PRE_NONE, 1); }
    add(PK_KIN, 0, 1, late ? 1 : 0, PRE_NONE, 0);
    if (late) {
        add(PK_LIN1, pl, 1, 1, PRE_NONE, 1);
        add(PK_LX0, pl, 1, 1, PRE_NONE, 1);
        add(PK_LX1, pl, 1, 0, PRE_NONE, 1);
    }
    return n;
}

__global__ void __launch_bounds__(kS2Threads, 1) rc_stream2_kernel(const __grid_constant__ S2Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    S2Smem& S = *reinterpret_cast<S2Smem*>(ring + kS2Slots * kS2SlotBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const StepIO& io = a.io;
    unsigned* const bar = a.bar + a.parity * 16;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kS2Slots; ++s) { mbar_init(smem_u32(&S.full[s]), 1); mbar_init(smem_u32(&S.empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (blockIdx.x == 0) for (int i = 0; i < 16; ++i) a.bar[(a.parity ^ 1) * 16 + i] = 0u;   // the other parity's words: next frame
    }
    if (blockIdx.x == 0) {
        const int* src = reinterpret_cast<const int*>(a.M);
        int* dst = reinterpret_cast<int*>(&S.M);
        for (int e = threadIdx.x; e < (int)(sizeof(RcModelConst) / 4); e += kS2Threads) dst[e] = src[e];
    }
    // ---- prep, in every CTA (sig_mp.py:138-153) ----
    if (warp == 0) {
        int inflags = 0;
        if (io.row_flags && (io.first_mode == 1 || (io.first_mode == 2 && a.t == 0))) inflags = io.row_flags[0] & 3;
        if (!io.first_tran) inflags &= ~RC_F_FIRST_TRAN;
        inflags |= RC_F_ACTIVE;
        const int f = rc_prep_warp(a.cfg, a.row->vision_count, S.prep, io.j2dc + (long long)a.t * 99, io.accc + (long long)a.t * 18,
                                   io.oric + (long long)a.t * 54, inflags, S.x2, S.x3, S.x4, S.x6, S.x7, S.rcr, &S.conf, S.lerpw, lane);
        if (lane == 0) { S.flags = f; S.ngroups = s2_build(f, S.groups, a.program); }
    }
    __syncthreads();
    const int f = S.flags;
    const int ng = S.ngroups;

    if (warp == kS2Cons) {
        // ---- producer: stream this CTA's weight slices in program order ----
        if (lane == 0) {
            uint32_t cc = 0;
            for (int gi = 0; gi < ng; ++gi) {
                const S2Group g = S.groups[gi];
                if (g.kind == PK_KIN) continue;
                for (int ni = 0; ni < NNETS; ++ni) {
                    if (!(g.mask & (1 << ni))) continue;
                    const S2Mat& m = s2_mat(a.net[ni], g.kind);
                    int u0, nu;
                    s2_slice(m, u0, nu);
                    const int upc = s2_units_per_chunk(m.K);
                    for (int u = 0; u < nu; u += upc, ++cc) {
                        const int cu = min(upc, nu - u);
                        const int s = cc % kS2Slots;
                        mbar_wait(smem_u32(&S.empty[s]), ((cc / kS2Slots) & 1u) ^ 1u);
                        const uint32_t bytes = (uint32_t)cu * 4u * (uint32_t)m.K * 4u;
                        mbar_expect_tx(smem_u32(&S.full[s]), bytes);
                        s2_bulk_load(smem_u32(ring + (size_t)s * kS2SlotBytes), m.W + (size_t)(u0 + u) * 4 * m.K, bytes, smem_u32(&S.full[s]));
                    }
                }
            }
        }
        return;
    }

    // ---- consumers ----
    const int ctid = threadIdx.x;                       // 0 .. 255
    uint32_t cc = 0;
    unsigned arrived = 0, waited = 0;
    int nts = 0;
    auto stamp = [&]() { if (a.ts && blockIdx.x == 0 && ctid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.ts[nts] = t; } ++nts; };
    stamp();
    auto grid_arrive = [&]() {
        s2_cons_sync();                                 // every consumer's global writes happen-before the release below
        if (ctid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + arrived) : "memory");
        ++arrived;
    };
    auto grid_wait = [&]() {
        if (ctid == 0) {
            const long long t0 = clock64();
            while (s2_ld_acquire(bar + waited) < gridDim.x) { if (clock64() - t0 > 4000000000LL) __trap(); }
        }
        ++waited;
        s2_cons_sync();
    };
    bool ran[NNETS];
#pragma unroll
    for (int i = 0; i < NNETS; ++i) ran[i] = false;

    for (int gi = 0; gi < ng; ++gi) {
        const S2Group g = S.groups[gi];
        if (a.ts && blockIdx.x == 0 && ctid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.ts[4 + 3 * gi] = t; }
        if (g.wait_before) grid_wait();
        if (a.ts && blockIdx.x == 0 && ctid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.ts[5 + 3 * gi] = t; }
        if (g.kind == PK_KIN) {
            stamp();
            // ---- kin on CTA 0 (:173-273): pose / tran out, row state, vision-updater inputs, init_net flag ----
            if (blockIdx.x == 0) {
                if (warp == 0) {
                    float g3[3], ft[3] = {0.f, 0.f, 0.f}, y8[2], vr[3], pc[3];
                    for (int i = 0; i < 3; ++i) { g3[i] = (io.gravity ? io.gravity : a.gravity)[i]; vr[i] = __ldcg(a.Y3 + i); pc[i] = __ldcg(a.Y6 + i); }
                    y8[0] = __ldcg(a.Y8); y8[1] = __ldcg(a.Y8 + 1);
                    if (f & RC_F_FIRST_TRAN) for (int i = 0; i < 3; ++i) ft[i] = io.first_tran[i];
                    for (int e = lane; e < 144; e += 32) S.xbuf[0][0][e] = __ldcg(a.Y7 + e);
                    __syncwarp();
                    // late inputs: rows of X4 / X6 as prep left them (acc / ori part), kin overwrites the key-point part
                    for (int e = lane; e < RC_K4; e += 32) a.X4[e] = S.x4[e];
                    for (int e = lane; e < RC_K6; e += 32) a.X6[e] = S.x6[e];
                    __syncwarp();
                    const int need_init = rc_kin_warp(a.cfg, S.M, S.kin, a.row, f, S.xbuf[0][0], y8, vr, pc, S.rcr, S.conf, g3, ft,
                                                      io.pose + (long long)a.t * 216, io.tran + (long long)a.t * 3, a.X4, a.X6, lane,
                                                      io.branch ? io.branch + a.t : nullptr);
                    if (lane == 0) a.gflags[0] = need_init;
                    S.flags = need_init ? (f | (1 << 30)) : f;
                }
                s2_cons_sync();
                if (S.flags & (1 << 30)) {
                    // rnn2.init_net re-seed (:178-183), once per stream: 69 -> 512 -> 1024 -> 2048 by the 8 consumer warps of CTA 0
                    const float* xin = a.X7 + 72;
                    float* l1 = a.I1; float* l2 = a.I2;
                    const int dims[4] = {kInitK0, 512, 1024, 2048};
                    for (int e = ctid; e < kInitK0; e += kS2Cons * 32) S.xbuf[0][0][e] = (e < 69) ? __ldcg(xin + e) : 0.f;
                    s2_cons_sync();
                    for (int l = 0; l < 3; ++l) {
                        const float* W = a.Wi[l]; const float* bs = a.bi[l];
                        const int K = dims[l], N = dims[l + 1];
                        const float* x = (l == 0) ? S.xbuf[0][0] : (l == 1 ? l1 : l2);
                        for (int o = warp * 4; o < N; o += kS2Cons * 4) {
                            float acc[4] = {0.f, 0.f, 0.f, 0.f};
                            for (int k = lane * 4; k < K; k += 128) {
                                const float4 xv = (l == 0) ? *reinterpret_cast<const float4*>(x + k) : __ldcg(reinterpret_cast<const float4*>(x + k));
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)(o + r) * K + k));
                                    acc[r] = fmaf(w.x, xv.x, acc[r]); acc[r] = fmaf(w.y, xv.y, acc[r]); acc[r] = fmaf(w.z, xv.z, acc[r]); acc[r] = fmaf(w.w, xv.w, acc[r]);
                                }
                            }
#pragma unroll
                            for (int r = 0; r < 4; ++r)
#pragma unroll
                                for (int s = 16; s > 0; s >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
                            if (lane == 0) {
                                float v[4];
#pragma unroll
                                for (int r = 0; r < 4; ++r) { v[r] = acc[r] + bs[o + r]; if (l < 2) v[r] = fmaxf(v[r], 0.f); }
                                if (l == 0) for (int r = 0; r < 4; ++r) __stcg(l1 + o + r, v[r]);
                                else if (l == 1) for (int r = 0; r < 4; ++r) __stcg(l2 + o + r, v[r]);
                                else {                                   // [h0 | h1 | c0 | c1] -> the state the NEXT frame reads
                                    const S2Net& n2 = a.net[NET2];
                                    float* dst = (o < 512) ? n2.hn[0] : (o < 1024 ? n2.hn[1] : (o < 1536 ? n2.c[0] : n2.c[1]));
                                    for (int r = 0; r < 4; ++r) __stcg(dst + (o & 511) + r, v[r]);
                                }
                            }
                        }
                        __threadfence_block();
                        s2_cons_sync();
                    }
                }
            }
            stamp();
            if (g.arrive_after) grid_arrive();
            continue;
        }
        if (g.pre == PRE_BLEND) {
            // joint blend (:154-167) from rnn2's / rnn4's outputs, in every CTA; CTA 0 keeps the fp32 row for kin / init_net
            const float* y2 = a.net[NET2].yout;
            const float* y4 = a.net[NET4].yout;
            if (ctid < 69) { S.x3[72 + ctid] = __ldcg(y2 + ctid); if (f & RC_F_HI) S.x6[171 + ctid] = __ldcg(y4 + ctid); }
            s2_cons_sync();
            if (ctid < 23) {
                float out[3];
                rc_mid_joint(f, S.rcr, S.lerpw, S.x3 + 72 + ctid * 3, S.x6 + 171 + ctid * 3, out);
                for (int c = 0; c < 3; ++c) { S.x7[72 + ctid * 3 + c] = out[c]; if (blockIdx.x == 0) a.X7[72 + ctid * 3 + c] = out[c]; }
            }
            s2_cons_sync();
        }
        // ---- the input vectors of all matrices of the group into shared memory, one barrier for the group.  Double-buffered by
        // group: a warp that loads group g has passed the barrier of group g - 1, which every warp only reaches after its reads of
        // group g - 2 (the previous user of this buffer set).  (Tried: fetching the vectors of the next recurrent-half group into
        // registers before the current group's weights are consumed — the recurrent groups got ~0.9 us shorter, the groups carrying
        // the loads ~2.5 us longer; dropped.)
        {
            int k = 0;
            for (int ni = 0; ni < NNETS; ++ni) {
                if (!(g.mask & (1 << ni))) continue;
                const S2Net& N = a.net[ni];
                const S2Mat& m = s2_mat(N, g.kind);
                const float* xs = nullptr;     // shared-memory source
                const float* xg = nullptr;     // global source (produced by other CTAs: L2-coherent loads)
                switch (g.kind) {
                    case PK_LIN1:
                        if (g.late) xg = (ni == NET4) ? a.X4 : a.X6;
                        else xs = (ni == NET2) ? S.x2 : (ni == NET3) ? S.x3 : (ni == NET4) ? S.x4 : (ni == NET6) ? S.x6 : S.x7;
                        break;
                    case PK_LH0: xg = N.h[0]; break;
                    case PK_LH1: xg = N.h[1]; break;
                    case PK_LX0: xg = N.a1; break;
                    case PK_LX1: xg = N.hn[0]; break;
                    default: xg = N.hn[1]; break;
                }
                float* xb = S.xbuf[gi & 1][k++];
                for (int e = ctid * 4; e < m.K; e += kS2Cons * 128) {
                    const float4 v = xs ? *reinterpret_cast<const float4*>(xs + e) : __ldcg(reinterpret_cast<const float4*>(xg + e));
                    *reinterpret_cast<float4*>(xb + e) = v;
                }
            }
            s2_cons_sync();
        }
        int kx = 0;
        for (int ni = 0; ni < NNETS; ++ni) {
            if (!(g.mask & (1 << ni))) continue;
            const S2Net& N = a.net[ni];
            const S2Mat& m = s2_mat(N, g.kind);
            ran[ni] = true;
            const float* xb = S.xbuf[gi & 1][kx++];
            int u0, nu;
            s2_slice(m, u0, nu);
            const int upc = s2_units_per_chunk(m.K);
            const int layer = (g.kind == PK_LH1 || g.kind == PK_LX1) ? 1 : 0;
            const bool lstm_x = g.kind == PK_LX0 || g.kind == PK_LX1;
            // one warp per chunk (round robin): up to 8 chunks of a matrix are consumed concurrently
            for (int u = 0; u < nu; u += upc, ++cc) {
                if ((int)(cc % kS2Cons) != warp) continue;
                const int cu = min(upc, nu - u);
                const int s = cc % kS2Slots;
                float cprev[4] = {0.f, 0.f, 0.f, 0.f};
                if (lstm_x && lane < cu) cprev[0] = __ldcg(N.c[layer] + u0 + u + lane);      // cell state ahead of the weights
                mbar_wait(smem_u32(&S.full[s]), (cc / kS2Slots) & 1u);
                const float* Ws = reinterpret_cast<const float*>(ring + (size_t)s * kS2SlotBytes);
                for (int i = 0; i < cu; ++i) {
                    const float* w0 = Ws + (size_t)i * 4 * m.K;
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                    for (int k = lane * 4; k < m.K; k += 128) {
                        const float4 xv = *reinterpret_cast<const float4*>(xb + k);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float4 w = *reinterpret_cast<const float4*>(w0 + (size_t)r * m.K + k);
                            acc[r] = fmaf(w.x, xv.x, acc[r]); acc[r] = fmaf(w.y, xv.y, acc[r]); acc[r] = fmaf(w.z, xv.z, acc[r]); acc[r] = fmaf(w.w, xv.w, acc[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                    const float cp = __shfl_sync(0xffffffffu, cprev[0], i);
                    if (lane == 0) {
                        const int ul = u + i;                   // unit index inside this CTA's slice
                        const int j = u0 + ul;                  // global unit = hidden unit (LSTM) / group of 4 output rows (linear)
                        if (g.kind == PK_LH0 || g.kind == PK_LH1) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) S.hp[ni][layer][ul][r] = acc[r];
                        } else if (lstm_x) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(m.bias + j * 4));
                            const float* hp = S.hp[ni][layer][ul];
                            const float pi = acc[0] + hp[0] + b.x, pf = acc[1] + hp[1] + b.y, pg = acc[2] + hp[2] + b.z, po = acc[3] + hp[3] + b.w;
                            const float cn = fmaf(sigm(pf), cp, sigm(pi) * tanhf(pg));
                            const float hn = sigm(po) * tanhf(cn);
                            __stcg(N.c[layer] + j, cn);
                            __stcg(N.hn[layer] + j, hn);
                        } else if (g.kind == PK_LIN1) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(m.bias + j * 4));
                            __stcg(reinterpret_cast<float4*>(N.a1 + j * 4), make_float4(fmaxf(acc[0] + b.x, 0.f), fmaxf(acc[1] + b.y, 0.f), fmaxf(acc[2] + b.z, 0.f), fmaxf(acc[3] + b.w, 0.f)));
                        } else if (N.yout) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) { const int n = j * 4 + r; if (n < N.out) __stcg(N.yout + n, acc[r] + m.bias[n]); }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) s2_mbar_arrive(smem_u32(&S.empty[s]));
            }
        }
        if (a.ts && blockIdx.x == 0 && ctid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.ts[6 + 3 * gi] = t; }
        if (g.arrive_after) grid_arrive();
    }
    // sub-nets that did not run this frame (live mode between key-point refreshes): carry their state into the new buffers
    for (int ni = 0; ni < NNETS; ++ni) {
        if (ran[ni]) continue;
        const S2Net& N = a.net[ni];
        for (int l = 0; l < 2; ++l) {
            int u0, nu;
            s2_slice(N.lx[l], u0, nu);
            for (int e = u0 + ctid; e < u0 + nu; e += kS2Cons * 32) N.hn[l][e] = N.h[l][e];
        }
    }
    stamp();
}

__global__ void __launch_bounds__(256) s2_split_kernel(const float* __restrict__ W, float* __restrict__ Wx, float* __restrict__ Wh, int H) {
    const long long total = (long long)4 * H * H / 4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / (H / 4);
        const int k = (int)(e % (H / 4)) * 4;
        *reinterpret_cast<float4*>(Wx + row * H + k) = *reinterpret_cast<const float4*>(W + row * 2 * H + k);
        *reinterpret_cast<float4*>(Wh + row * H + k) = *reinterpret_cast<const float4*>(W + row * 2 * H + H + k);
    }
}

constexpr int kS2Smem = kS2Slots * kS2SlotBytes + (int)sizeof(S2Smem) + 256;

}  // namespace

bool rc_stream2_supported(const rc_state* s) {
    static const bool off = getenv("RC_STREAM2") && atoi(getenv("RC_STREAM2")) == 0;
    return !off && s && s->B == 1;
}

// One frame (index t of io; io strides are those of a single stream).  The caller guarantees that the frame has no first_frame flag.
int rc_stream2_frame(rc_state* s, const StepIO& io, int t, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    rc_net* n = const_cast<rc_net*>(s->net);
    if (!n->s2_ready) {
        // separate contiguous [4H, H] matrices for the input and the recurrent half of every LSTM layer (sliceable by rows for the
        // bulk copies), derived on the device from the gate-interleaved [4H, 2H] packing
        for (int i = 0; i < NNETS; ++i)
            for (int l = 0; l < 2; ++l) {
                const size_t H = (size_t)n->nets[i].H;
                float *wx = nullptr, *wh = nullptr;
                RC_CUDA(cudaMalloc(&wx, 4 * H * H * sizeof(float)));
                RC_CUDA(cudaMalloc(&wh, 4 * H * H * sizeof(float)));
                n->allocs.push_back(wx); n->allocs.push_back(wh);
                RC_LAUNCH(s2_split_kernel, 592, 256, 0, stream, (const float*)n->nets[i].WL[l], wx, wh, (int)H);
                n->s2_Wx[i][l] = wx; n->s2_Wh[i][l] = wh;
            }
        RC_CHECK_LAUNCH();
        RC_CUDA(cudaFuncSetAttribute(rc_stream2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS2Smem));
        n->s2_ready = true;
    }
    if (!s->s2_bar) {
        RC_CUDA(cudaMalloc(&s->s2_bar, 64 * sizeof(unsigned)));
        RC_CUDA(cudaMemsetAsync(s->s2_bar, 0, 64 * sizeof(unsigned), st));
        int dev = 0, sms = 0, per_sm = 0;
        RC_CUDA(cudaGetDevice(&dev));
        RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        RC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rc_stream2_kernel, kS2Threads, kS2Smem));
        if (per_sm < 1) { rc_set_error("rc_stream2_kernel does not fit on an SM"); return RC_ERR_CUDA; }
        s->s2_grid = sms;
        // cached graphs of the multi-launch path bake the h / hn pointers that this path swaps every frame
        if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
        if (s->on_graph) { cudaGraphExecDestroy(s->on_graph); s->on_graph = nullptr; }
    }
    S2Args a;
    memset(&a, 0, sizeof(a));
    a.cfg = n->cfg; a.M = n->model->d_const; a.row = s->rows; a.io = io; a.t = t;
    for (int i = 0; i < NNETS; ++i) {
        const NetDev& w = n->nets[i];
        NetBuf& nb = s->nb[i];
        S2Net& k = a.net[i];
        k.H = w.H; k.out = w.out;
        k.lin1 = S2Mat{w.W1, w.b1, w.H / 4, w.K1};
        for (int l = 0; l < 2; ++l) {
            k.lx[l] = S2Mat{n->s2_Wx[i][l], w.bL[l], w.H, w.H};
            k.lh[l] = S2Mat{n->s2_Wh[i][l], w.bL[l], w.H, w.H};
            k.h[l] = nb.h[l]; k.hn[l] = nb.hn[l]; k.c[l] = nb.c[l];
        }
        k.lin2 = S2Mat{w.W2, w.b2, w.out4 / 4, w.H};
        k.a1 = nb.a1;
    }
    a.net[NET2].yout = s->X3 + 72; a.net[NET3].yout = s->Y3; a.net[NET4].yout = s->X6 + 171;
    a.net[NET6].yout = s->Y6; a.net[NET7].yout = s->Y7; a.net[NET8].yout = s->Y8;
    for (int l = 0; l < 3; ++l) { a.Wi[l] = n->Wi[l]; a.bi[l] = n->bi[l]; }
    a.X4 = s->X4; a.X6 = s->X6; a.X7 = s->X7; a.Y3 = s->Y3; a.Y6 = s->Y6; a.Y7 = s->Y7; a.Y8 = s->Y8; a.I1 = s->I1; a.I2 = s->I2;
    a.gravity = s->gravity; a.gflags = (int*)(s->s2_bar + 48); a.bar = s->s2_bar; a.parity = s->s2_parity;
    static const int program = getenv("RC_S2_PROGRAM") ? atoi(getenv("RC_S2_PROGRAM")) : 1;
    a.program = program;
    static const bool want_ts = getenv("RC_STREAM_TS") != nullptr;
    if (want_ts && !s->s2_ts) { RC_CUDA(cudaMalloc(&s->s2_ts, (4 + 3 * kS2MaxGroups) * 8)); RC_CUDA(cudaMemsetAsync(s->s2_ts, 0, (4 + 3 * kS2MaxGroups) * 8, st)); }
    a.ts = want_ts ? s->s2_ts : nullptr;
    // One CTA per SM (228 KB of shared memory each) on an otherwise idle device: all CTAs are co-resident, which the grid barriers
    // need (every spin is bounded and traps instead of hanging).  A plain launch: the cooperative launch API costs ~10 us more per frame.
    RC_LAUNCH(rc_stream2_kernel, s->s2_grid, kS2Threads, kS2Smem, stream, a);
    RC_CHECK_LAUNCH();
    s->s2_parity ^= 1;
    for (int i = 0; i < NNETS; ++i)
        for (int l = 0; l < 2; ++l) std::swap(s->nb[i].h[l], s->nb[i].hn[l]);
    if (want_ts) {
        static int printed = 0;
        unsigned long long h[4 + 3 * kS2MaxGroups];
        RC_CUDA(cudaStreamSynchronize(st));
        RC_CUDA(cudaMemcpy(h, s->s2_ts, sizeof(h), cudaMemcpyDeviceToHost));
        if (printed++ % 50 == 10) {
            fprintf(stderr, "[stream2 kernel us] nets %.1f | kin %.1f | rest %.1f | total %.1f || per group (barrier wait + work):", (h[1] - h[0]) * 1e-3, (h[2] - h[1]) * 1e-3,
                    (h[3] - h[2]) * 1e-3, (h[3] - h[0]) * 1e-3);
            for (int gi = 0; gi < kS2MaxGroups && h[4 + 3 * gi]; ++gi)
                fprintf(stderr, " g%d %.1f+%.1f", gi, (h[5 + 3 * gi] - h[4 + 3 * gi]) * 1e-3, (h[6 + 3 * gi] - h[5 + 3 * gi]) * 1e-3);
            fprintf(stderr, "\n");
        }
        RC_CUDA(cudaMemsetAsync(s->s2_ts, 0, sizeof(h), st));
    }
    return RC_OK;
}
