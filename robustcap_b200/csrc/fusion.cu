// Net.forward_online / forward_offline for B independent streams (net/sig_mp.py:23-274).
//
// Per frame the host enqueues (no host<->device sync anywhere, unlike the reference's .item() at :138):
//   prep    — confidence, IMU change of frame, key-point normalisation, sub-net inputs, branch flags
//   lists   — compacts the branch flags into row lists (all / vision / rnn6-first-frame / rnn6 / late)
//   rnn2, rnn3 (all rows) ; rnn4, rnn6 (vision rows) ; mid (lerp) ; rnn7, rnn8 (all rows)
//   kin     — 6D->R, IK, foot FK, translation state machine, SMPL FK of the 33 key points, vision-updater inputs
//   init    — rnn2.init_net re-seed for rows that reached c >= hi for the first time (:178-183)
//   rnn6, rnn4 late pass on the synthetic key points (:263-271), outputs discarded (linear2 skipped)
// Each sub-net pass = linear1+relu -> fused LSTM layer 0 -> fused LSTM layer 1 -> commit(h) [-> linear2].
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include <string.h>
#include <stdlib.h>
#include "rc_common.cuh"
#include "rc_rows.h"
#include "rc_rows_warp.cuh"
#include "rc_model.cuh"
#include "rc_linear.cuh"
#include "rc_pack.h"
#include "rc_tc.cuh"
#include "rc_fusion.cuh"
#include "rc_seq.cuh"

namespace {


__global__ void __launch_bounds__(128) rc_prep_kernel(RcNetCfg cfg, const RcRowState* __restrict__ rows, StepIO io, int B,
                                                       float* X2, float* X3, float* X4, float* X6, float* X7,
                                                       float* rcr, float* conf, float* lerpw, int* flags) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int t = io.d_t ? *io.d_t : 0;
    int inflags = 0;
    if (io.row_flags && (io.first_mode == 1 || (io.first_mode == 2 && t == 0))) inflags = io.row_flags[b] & 3;
    if (!io.first_tran) inflags &= ~RC_F_FIRST_TRAN;
    const bool active = !io.lengths || t < io.lengths[b];
    if (!active) { flags[b] = 0; return; }
    inflags |= RC_F_ACTIVE;
    float kp[99], acc[18], ori[54];
    const float* pj = io.j2dc + b * io.sj + (long long)t * 99;
    const float* pa = io.accc + b * io.sa + (long long)t * 18;
    const float* po = io.oric + b * io.so + (long long)t * 54;
    for (int i = 0; i < 99; ++i) kp[i] = pj[i];
    for (int i = 0; i < 18; ++i) acc[i] = pa[i];
    for (int i = 0; i < 54; ++i) ori[i] = po[i];
    flags[b] = rc_prep_row(cfg, rows[b], kp, acc, ori, inflags, X2 + (size_t)b * RC_K2, X3 + (size_t)b * RC_K3,
                           X4 + (size_t)b * RC_K4, X6 + (size_t)b * RC_K6, X7 + (size_t)b * RC_K7, rcr + b * 9,
                           conf + b, lerpw + b * 2);
}

// ordered compaction of the flag predicates into the five row lists: one 1024-thread block, a ballot per warp and list, a scan of
// the 32 warp counts (the one-warp-per-list version took 7.8 us per frame, mostly serial rounds of dependent flag loads)
__global__ void __launch_bounds__(1024) rc_lists_kernel(const int* __restrict__ flags, int B, int* __restrict__ lists, int* __restrict__ counts) {
    rc_pdl_wait();
    rc_pdl_trigger();
    constexpr int NL = L_INIT;                                   // the lists built here: L_ALL .. L_LATE
    __shared__ int wcount[NL][32];
    __shared__ int running[NL];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NL) running[threadIdx.x] = 0;
    if (threadIdx.x == 0) counts[L_INIT] = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + threadIdx.x;
        const int f = (b < B) ? flags[b] : 0;
        const bool act = (f & RC_F_ACTIVE) != 0;
        bool p[NL];
        p[L_ALL] = act;
        p[L_HI] = act && (f & RC_F_HI);
        p[L_6A] = act && (f & RC_F_FIRST_FRAME);
        p[L_6B] = act && (f & RC_F_R6B);
        p[L_LATE] = act && (f & RC_F_LATE);
        unsigned m[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            m[l] = __ballot_sync(0xffffffffu, p[l]);
            if (lane == 0) wcount[l][warp] = __popc(m[l]);
        }
        __syncthreads();
        if (warp < NL) {                                         // exclusive scan of the 32 warp counts of list `warp`
            const int c = wcount[warp][lane];
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            wcount[warp][lane] = running[warp] + incl - c;
            __syncwarp();
            if (lane == 31) running[warp] += incl;
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < NL; ++l)
            if (p[l]) lists[l * B + wcount[l][warp] + __popc(m[l] & ((1u << lane) - 1u))] = b;
        __syncthreads();
    }
    if (threadIdx.x < NL) counts[threadIdx.x] = running[threadIdx.x];
}

__global__ void __launch_bounds__(128) rc_mid_kernel(const int* __restrict__ flags, int B, const float* __restrict__ rcr,
                                                      const float* __restrict__ lerpw, const float* X3, const float* X6, float* X7) {
    rc_pdl_wait();
    rc_pdl_trigger();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;           // one thread per (stream, joint)
    const int b = e / 23, i = e % 23;
    if (b >= B) return;
    const int f = flags[b];
    if (!(f & RC_F_ACTIVE)) return;
    float r[9];
    for (int q = 0; q < 9; ++q) r[q] = rcr[b * 9 + q];
    rc_mid_joint(f, r, lerpw + b * 2, X3 + (size_t)b * RC_K3 + 72 + i * 3, X6 + (size_t)b * RC_K6 + 171 + i * 3,
                 X7 + (size_t)b * RC_K7 + 72 + i * 3);
}

__global__ void __launch_bounds__(64) rc_kin_kernel(RcNetCfg cfg, const RcModelConst* __restrict__ M, RcRowState* rows,
                                                     const int* __restrict__ flags, int B, StepIO io,
                                                     const float* __restrict__ Y7, const float* __restrict__ Y8,
                                                     const float* __restrict__ Y3, const float* __restrict__ Y6,
                                                     const float* __restrict__ rcr, const float* __restrict__ conf,
                                                     const float* __restrict__ gravity_all, float* X4, float* X6,
                                                     const float* X7, float* XI, int* lists, int* counts) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int f = flags[b];
    if (!(f & RC_F_ACTIVE)) return;
    const int t = io.d_t ? *io.d_t : 0;
    float y7[144], r[9], g[3], ft[3] = {0.f, 0.f, 0.f}, pose[216], tran[3];
    for (int i = 0; i < 144; ++i) y7[i] = Y7[(size_t)b * 144 + i];
    for (int i = 0; i < 9; ++i) r[i] = rcr[b * 9 + i];
    const float* gp = io.gravity ? (io.gravity + (size_t)b * 3) : gravity_all;
    for (int i = 0; i < 3; ++i) g[i] = gp[i];
    if (f & RC_F_FIRST_TRAN) for (int i = 0; i < 3; ++i) ft[i] = io.first_tran[(size_t)b * 3 + i];
    RcRowState st = rows[b];
    const int need_init = rc_kin_row(cfg, *M, &st, f, y7, Y8 + b * 4, Y3 + b * 4, Y6 + b * 4, r, conf[b], g, ft, pose, tran,
                                     X4 + (size_t)b * RC_K4, X6 + (size_t)b * RC_K6, io.branch ? io.branch + b * io.sb + t : nullptr);
    rows[b] = st;
    float* po = io.pose + b * io.sp + (long long)t * 216;
    float* to = io.tran + b * io.st + (long long)t * 3;
    for (int i = 0; i < 216; ++i) po[i] = pose[i];
    for (int i = 0; i < 3; ++i) to[i] = tran[i];
    if (need_init) {
        for (int i = 0; i < 69; ++i) XI[(size_t)b * kInitK0 + i] = X7[(size_t)b * RC_K7 + 72 + i];
        for (int i = 69; i < kInitK0; ++i) XI[(size_t)b * kInitK0 + i] = 0.f;
        lists[L_INIT * B + atomicAdd(&counts[L_INIT], 1)] = b;
    }
}

// ---- warp-per-stream versions (default): same arithmetic, 4 streams per block -------------------------------------------
constexpr int kRowWarps = 4;

__global__ void __launch_bounds__(kRowWarps * 32) rc_prep_warp_kernel(RcNetCfg cfg, const RcRowState* __restrict__ rows, StepIO io, int B,
                                                                       float* X2, float* X3, float* X4, float* X6, float* X7,
                                                                       float* rcr, float* conf, float* lerpw, int* flags, int* lists, int* counts) {
    rc_pdl_wait();
    rc_pdl_trigger();
    __shared__ RcPrepWarpSmem S[kRowWarps];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRowWarps + w;
    if (b >= B) return;
    const int t = io.d_t ? *io.d_t : 0;
    int inflags = 0;
    if (io.row_flags && (io.first_mode == 1 || (io.first_mode == 2 && t == 0))) inflags = io.row_flags[b] & 3;
    if (!io.first_tran) inflags &= ~RC_F_FIRST_TRAN;
    const bool active = !io.lengths || t < io.lengths[b];
    if (!active) {
        if (lane == 0) { flags[b] = 0; if (B == 1 && lists) for (int l = 0; l < NLISTS; ++l) counts[l] = 0; }
        return;
    }
    inflags |= RC_F_ACTIVE;
    const int f = rc_prep_warp(cfg, rows[b].vision_count, S[w], io.j2dc + b * io.sj + (long long)t * 99,
                               io.accc + b * io.sa + (long long)t * 18, io.oric + b * io.so + (long long)t * 54, inflags,
                               X2 + (size_t)b * RC_K2, X3 + (size_t)b * RC_K3, X4 + (size_t)b * RC_K4, X6 + (size_t)b * RC_K6,
                               X7 + (size_t)b * RC_K7, rcr + b * 9, conf + b, lerpw + b * 2, lane);
    if (lane == 0) flags[b] = f;
    if (B == 1 && lane == 0 && lists) {        // single stream: the row lists are trivial, no separate compaction launch
        const int on[NLISTS] = {1, (f & RC_F_HI) != 0, (f & RC_F_FIRST_FRAME) != 0, (f & RC_F_R6B) != 0, (f & RC_F_LATE) != 0, 0};
        for (int l = 0; l < NLISTS; ++l) { lists[l] = 0; counts[l] = on[l]; }
    }
}

__global__ void __launch_bounds__(kRowWarps * 32) rc_kin_warp_kernel(RcNetCfg cfg, const RcModelConst* __restrict__ M, RcRowState* rows,
                                                                      const int* __restrict__ flags, int B, StepIO io,
                                                                      const float* __restrict__ Y7, const float* __restrict__ Y8,
                                                                      const float* __restrict__ Y3, const float* __restrict__ Y6,
                                                                      const float* __restrict__ rcr, const float* __restrict__ conf,
                                                                      const float* __restrict__ gravity_all, float* X4, float* X6,
                                                                      const float* X7, float* XI, int* lists, int* counts) {
    rc_pdl_wait();
    rc_pdl_trigger();
    __shared__ RcKinWarpSmem S[kRowWarps];
    __shared__ RcModelConst Ms;
    {   // SMPL constants once per block
        const int* src = reinterpret_cast<const int*>(M);
        int* dst = reinterpret_cast<int*>(&Ms);
        for (int e = threadIdx.x; e < (int)(sizeof(RcModelConst) / 4); e += blockDim.x) dst[e] = src[e];
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRowWarps + w;
    if (b >= B) return;
    const int f = flags[b];
    if (!(f & RC_F_ACTIVE)) return;
    const int t = io.d_t ? *io.d_t : 0;
    float r[9], g[3], ft[3] = {0.f, 0.f, 0.f}, y8[2], vr[3], pc[3];
    for (int i = 0; i < 9; ++i) r[i] = rcr[b * 9 + i];
    const float* gp = io.gravity ? (io.gravity + (size_t)b * 3) : gravity_all;
    for (int i = 0; i < 3; ++i) { g[i] = gp[i]; vr[i] = Y3[b * 4 + i]; pc[i] = Y6[b * 4 + i]; }
    y8[0] = Y8[b * 4]; y8[1] = Y8[b * 4 + 1];
    if (f & RC_F_FIRST_TRAN) for (int i = 0; i < 3; ++i) ft[i] = io.first_tran[(size_t)b * 3 + i];
    const int need_init = rc_kin_warp(cfg, Ms, S[w], rows + b, f, Y7 + (size_t)b * 144, y8, vr, pc, r, conf[b], g, ft,
                                      io.pose + b * io.sp + (long long)t * 216, io.tran + b * io.st + (long long)t * 3,
                                      X4 + (size_t)b * RC_K4, X6 + (size_t)b * RC_K6, lane, io.branch ? io.branch + b * io.sb + t : nullptr);
    if (need_init) {
        for (int i = lane; i < kInitK0; i += 32) XI[(size_t)b * kInitK0 + i] = (i < 69) ? X7[(size_t)b * RC_K7 + 72 + i] : 0.f;
        if (lane == 0) lists[L_INIT * B + atomicAdd(&counts[L_INIT], 1)] = b;
    }
}

// init_net output [h0 | h1 | c0 | c1] -> rnn2 state (sig_mp.py:182-183)
__global__ void __launch_bounds__(256) rc_init_scatter_kernel(const float* __restrict__ I3, const int* __restrict__ rows,
                                                               const int* __restrict__ count, float* h0, float* h1, float* c0, float* c1) {
    const int cnt = *count;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)cnt * 2048; e += (long long)gridDim.x * blockDim.x) {
        const int r = rows[e / 2048], q = (int)(e % 2048), u = q & 511;
        const float v = I3[(size_t)r * 2048 + q];
        float* dst = (q < 512) ? h0 : (q < 1024 ? h1 : (q < 1536 ? c0 : c1));
        dst[(size_t)r * 512 + u] = v;
    }
}

__global__ void rc_advance_kernel(int* d_t) { *d_t += 1; }
__global__ void rc_reset_rows_kernel(RcRowState* rows, int B, int created) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) rc_row_state_reset(&rows[b], created != 0);
}

// ---- host helpers ------------------------------------------------------------------------------------------------
template <class T>
int dev_alloc(std::vector<void*>& pool, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) { rc_set_error("cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); return RC_ERR_ALLOC; }
    pool.push_back(q);
    *p = (T*)q;
    return RC_OK;
}
#define RC_TRY(x) do { int rc_ = (x); if (rc_ != RC_OK) return rc_; } while (0)

int upload(std::vector<void*>& pool, float** d, const std::vector<float>& h) {
    RC_TRY(dev_alloc(pool, d, h.size()));
    RC_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return RC_OK;
}

int launch_linear(const RcLinear& a, int B, bool lstm, void* stream) {
    if (B <= 8) {
        const int njobs = a.Nw / 4;
        const int K = a.K1 + a.K2;
        const int ksplit = rc_gemv_ksplit(K);
        const int jpb = 8 / ksplit;
        const int grid = rc_cdiv(njobs, jpb);
#define RC_GEMV(MM)                                                                          \
        if (lstm) RC_LAUNCH((rc_gemv_kernel<MM, true>), grid, 256, 0, stream, a, ksplit);    \
        else RC_LAUNCH((rc_gemv_kernel<MM, false>), grid, 256, 0, stream, a, ksplit)
        if (B == 1) { RC_GEMV(1); }
        else if (B == 2) { RC_GEMV(2); }
        else if (B <= 4) { RC_GEMV(4); }
        else { RC_GEMV(8); }
#undef RC_GEMV
    } else {
        dim3 grid(rc_cdiv(a.Nw, kBN), rc_cdiv(B, kBM));
        if (lstm) RC_LAUNCH((rc_gemm_kernel<true>), grid, 256, 0, stream, a);
        else RC_LAUNCH((rc_gemm_kernel<false>), grid, 256, 0, stream, a);
    }
    RC_CHECK_LAUNCH();
    return RC_OK;
}

// one sub-net over the rows of list `li`: X [B, K1] -> (optional) Y
int net_pass(rc_state* s, int ni, int li, const float* X, float* Y, int ldy, void* stream, int lane = 0) {
    const NetDev& w = s->net->nets[ni];
    NetBuf& nb = s->nb[ni];
    const int B = s->B;
    const int* rows = s->lists + (size_t)li * B;
    const int* count = s->counts + li;
    RcLinear a;
    memset(&a, 0, sizeof(a));
    a.rows = rows; a.count = count;
    // linear1 + relu
    a.X = X; a.ldx = w.K1; a.X2 = nullptr; a.ldx2 = 0; a.K1 = w.K1; a.K2 = 0;
    a.W = w.W1; a.bias = w.b1; a.N = w.H; a.Nw = w.H; a.Y = nb.a1; a.ldy = w.H; a.relu = 1; a.H = w.H;
    const bool tc = s->net->gemm_mode >= 1 && s->tc_ready && B > 8;
    if (tc) {
        // tensor-core path: one gather/split pre-pass, then lin1 -> LSTM0 -> LSTM1 -> commit -> lin2, every epilogue writing the
        // next GEMM's A operand directly (split fp16 halves, compact row order)
        uint16_t **Ah = s->Ahi[lane], **Al = s->Alo[lane];
        const RcSplitSeg segs[3] = {
            {X, w.K1, w.K1, w.K1p, 0, w.K1p, Ah[0], Al[0]},                      // linear1 input
            {nb.h[0], w.H, w.H, w.H, w.H, 2 * w.H, Ah[1], Al[1]},                // h_prev of layer 0 -> columns [H, 2H) of buf 1
            {nb.h[1], w.H, w.H, w.H, w.H, 2 * w.H, Ah[2], Al[2]}};               // h_prev of layer 1 -> columns [H, 2H) of buf 2
        RC_TRY(rc_tc_split_pass(segs, 3, rows, count, B, stream));
        RC_TRY(rc_tc_linear(&s->mA0hi[lane][ni], &s->mA0lo[lane][ni], &w.mW1hi, &w.mW1lo, w.b1, nullptr, 0, w.H, w.K1p, 1, rows, count, B, stream,
                            Ah[1], Al[1], 2 * w.H));
        for (int l = 0; l < 2; ++l) {
            const bool prof = s->prof_on && ni == NET4;
            if (prof) {            // CUDA events right around the dominant GEMM launch, on the stream it is launched on
                if (s->prof_used + 2 > s->prof_ev.size()) {
                    for (int q = 0; q < 256; ++q) { cudaEvent_t e; RC_CUDA(cudaEventCreate(&e)); s->prof_ev.push_back(e); }
                }
                RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], (cudaStream_t)stream));
            }
            void* nh = (l == 0) ? (void*)Ah[2] : (Y ? (void*)Ah[1] : nullptr);
            void* nl = (l == 0) ? (void*)Al[2] : (Y ? (void*)Al[1] : nullptr);
            const int np = (l == 0) ? 2 * w.H : w.H;
            // The GEMM reads the split copies in the operand buffers, never the fp32 state, so the new hidden state goes straight
            // into h (no scratch + commit launch as in the fp32 paths, where other CTAs still read h_prev while h_new is produced).
            RC_TRY(rc_tc_lstm_layer(l == 0 ? &s->mA1hi[lane][ni] : &s->mA2hi[lane][ni], l == 0 ? &s->mA1lo[lane][ni] : &s->mA2lo[lane][ni],
                                    &w.mWhi[l], &w.mWlo[l], w.bL[l], nb.c[l], nb.h[l], w.H, rows, count, B, stream, nh, nl, np));
            if (prof) RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], (cudaStream_t)stream));
        }
        if (Y) RC_TRY(rc_tc_linear(&s->mA1Hhi[lane][ni], &s->mA1Hlo[lane][ni], &w.mW2hi, &w.mW2lo, w.b2, Y, ldy, w.out, w.H, 0, rows, count, B, stream));
        return RC_OK;
    }
    RC_TRY(launch_linear(a, B, false, stream));
    // LSTM layers
    for (int l = 0; l < 2; ++l) {
        a.X = (l == 0) ? nb.a1 : nb.hn[0]; a.ldx = w.H; a.X2 = nb.h[l]; a.ldx2 = w.H; a.K1 = w.H; a.K2 = w.H;
        a.W = w.WL[l]; a.bias = w.bL[l]; a.N = 4 * w.H; a.Nw = 4 * w.H; a.Y = nullptr; a.ldy = 0; a.relu = 0;
        a.C = nb.c[l]; a.Hout = nb.hn[l];
        const bool prof = s->prof_on && ni == NET4;
        if (prof) {
            if (s->prof_used + 2 > s->prof_ev.size()) {
                for (int q = 0; q < 256; ++q) { cudaEvent_t e; RC_CUDA(cudaEventCreate(&e)); s->prof_ev.push_back(e); }
            }
            RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], (cudaStream_t)stream));
        }
        RC_TRY(launch_linear(a, B, true, stream));
        if (prof) RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], (cudaStream_t)stream));
    }
    const bool fuse_commit = Y && B <= 8;          // the GEMV linear2 launch also commits h <- h_new (it only reads the scratch)
    if (!fuse_commit) {
        const long long work = (long long)B * (w.H / 4);
        const int grid = (int)std::min<long long>(rc_cdiv(work, 256), 1184);
        RC_LAUNCH(rc_commit_kernel, grid, 256, 0, stream, nb.hn[0], nb.hn[1], nb.h[0], nb.h[1], w.H, rows, count);
        RC_CHECK_LAUNCH();
    }
    if (Y) {
        a.X = nb.hn[1]; a.ldx = w.H; a.X2 = nullptr; a.ldx2 = 0; a.K1 = w.H; a.K2 = 0;
        a.W = w.W2; a.bias = w.b2; a.N = w.out; a.Nw = w.out4; a.Y = Y; a.ldy = ldy; a.relu = 0; a.C = nullptr; a.Hout = nullptr;
        if (fuse_commit) { a.commit_src[0] = nb.hn[0]; a.commit_src[1] = nb.hn[1]; a.commit_dst[0] = nb.h[0]; a.commit_dst[1] = nb.h[1]; a.commit_H = w.H; }
        RC_TRY(launch_linear(a, B, false, stream));
    }
    return RC_OK;
}

int init_pass(rc_state* s, void* stream) {
    const rc_net* n = s->net;
    const int B = s->B;
    RcLinear a;
    memset(&a, 0, sizeof(a));
    a.rows = s->lists + (size_t)L_INIT * B; a.count = s->counts + L_INIT;
    const float* xin[3] = {s->XI, s->I1, s->I2};
    float* yout[3] = {s->I1, s->I2, s->I3};
    const int kin[3] = {kInitK0, 512, 1024};
    for (int l = 0; l < 3; ++l) {
        a.X = xin[l]; a.ldx = kin[l]; a.K1 = kin[l]; a.K2 = 0; a.X2 = nullptr;
        a.W = n->Wi[l]; a.bias = n->bi[l]; a.N = kInitDims[l + 1]; a.Nw = kInitDims[l + 1];
        a.Y = yout[l]; a.ldy = kInitDims[l + 1]; a.relu = (l < 2); a.H = 0;
        RC_TRY(launch_linear(a, B, false, stream));
    }
    const NetBuf& nb = s->nb[NET2];
    RC_LAUNCH(rc_init_scatter_kernel, std::min(rc_cdiv((long long)B * 2048, 256), 1184), 256, 0, stream, s->I3,
              s->lists + (size_t)L_INIT * B, s->counts + L_INIT, nb.h[0], nb.h[1], nb.c[0], nb.c[1]);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

// ---- persistent grouped kernel (gemm mode 2): job lists per phase -------------------------------------------------------------
struct ChainSpec { int ni, li; float* Y; int ldy; };

int build_phase(rc_state* s, int ph, const ChainSpec* ch, int nch) {
    const rc_net* net = s->net;
    const int B = s->B;
    const long long Bpad = (long long)((B + 127) / 128) * 128;
    const int MT = (int)(Bpad / 128 + 1) / 2 * 2;          // even: the CTA-pair kernel works on pairs of row blocks
    RcPhDesc* d = new RcPhDesc();
    memset(d, 0, sizeof(*d));
    int idx[4][4];
    int nj = 0;
    for (int layer = 0; layer < 4; ++layer)                 // layer-major job order: every chain's linear1 first, linear2 last
        for (int c = 0; c < nch; ++c) idx[c][layer] = (layer == 3 && !ch[c].Y) ? -1 : nj++;
    int rc = RC_OK, max_tiles = 0, nseg = 0;
    if (nj > RC_PH_MAXJOBS || nch * 3 > RC_PH_MAXSEGS) { rc_set_error("build_phase: too many jobs"); rc = RC_ERR_ARG; }
    auto mk = [&](RcTensorMap* m, const void* base, int K) { if (rc == RC_OK) rc = rc_tc_make_map(m, base, Bpad, K, 128); };
    for (int c = 0; c < nch && rc == RC_OK; ++c) {
        const int ni = ch[c].ni;
        const NetDev& w = net->nets[ni];
        const NetBuf& nb = s->nb[ni];
        const int* rows = s->lists + (size_t)ch[c].li * B;
        const int* count = s->counts + ch[c].li;
        uint16_t** Ph = s->Phi[ni];
        uint16_t** Pl = s->Plo[ni];
        const int H = w.H;
        for (int layer = 0; layer < 4; ++layer) {
            if (idx[c][layer] < 0) continue;
            RcPhJob& j = d->job[idx[c][layer]];
            j.rows = rows; j.count = count; j.H = H;
            j.level = layer;
            j.dep = layer ? idx[c][layer - 1] : -1;
            if (layer == 0) {
                mk(&j.mAhi, Ph[0], w.K1p); mk(&j.mAlo, Pl[0], w.K1p);
                j.mWhi = w.mW1hi; j.mWlo = w.mW1lo; j.mWhi64 = w.mW1hi64; j.mWlo64 = w.mW1lo64; j.bias = w.b1; j.N = H; j.relu = 1; j.K = w.K1p;
                j.nAhi = Ph[1]; j.nAlo = Pl[1]; j.npitch = 2 * H; j.kind = 0; j.nt = H / RC_TC_BN;
                j.nt2 = (H + 255) / 256; j.nmma = H >= 256 ? 256 : (H + 15) / 16 * 16;
            } else if (layer == 1 || layer == 2) {
                const int l = layer - 1;
                mk(&j.mAhi, Ph[layer], 2 * H); mk(&j.mAlo, Pl[layer], 2 * H);
                j.mWhi = w.mWhi[l]; j.mWlo = w.mWlo[l]; j.mWhi64 = w.mWhi64[l]; j.mWlo64 = w.mWlo64[l]; j.bias = w.bL[l]; j.C = nb.c[l]; j.Hout = nb.h[l]; j.N = 4 * H; j.K = 2 * H;
                if (l == 0) { j.nAhi = Ph[2]; j.nAlo = Pl[2]; j.npitch = 2 * H; }
                else if (ch[c].Y) { j.nAhi = Ph[3]; j.nAlo = Pl[3]; j.npitch = H; }
                j.kind = 1; j.nt = 4 * H / RC_TC_BN;
                j.nt2 = 4 * H / 256; j.nmma = 256;
            } else {
                mk(&j.mAhi, Ph[3], H); mk(&j.mAlo, Pl[3], H);
                j.mWhi = w.mW2hi; j.mWlo = w.mW2lo; j.mWhi64 = w.mW2hi64; j.mWlo64 = w.mW2lo64; j.bias = w.b2; j.Y = ch[c].Y; j.ldy = ch[c].ldy; j.N = w.out; j.K = H;
                j.kind = 0; j.nt = w.outp / RC_TC_BN;
                j.nt2 = (w.out + 255) / 256; j.nmma = w.out >= 256 ? 256 : (w.out + 15) / 16 * 16;
            }
            max_tiles += MT * j.nt;
        }
        const float* X = ni == NET2 ? s->X2 : ni == NET3 ? s->X3 : ni == NET4 ? s->X4 : ni == NET6 ? s->X6 : s->X7;
        RcSplitSegM sx{X, w.K1, w.K1, w.K1p, 0, w.K1p, Ph[0], Pl[0], rows, count, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (ni == NET7 || ni == NET8) {        // the joint blend ("mid", :154-167) is computed by the pre-pass; rnn7's segment also stores it for kin
            sx.mid_flags = s->flags; sx.mid_rcr = s->rcr; sx.mid_lerpw = s->lerpw; sx.mid_x3 = s->X3; sx.mid_x6 = s->X6;
            sx.mid_out = (ni == NET7) ? s->X7 : nullptr;
        }
        s->ph_segs[ph][nseg++] = sx;
        s->ph_segs[ph][nseg++] = RcSplitSegM{nb.h[0], H, H, H, H, 2 * H, Ph[1], Pl[1], rows, count, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        s->ph_segs[ph][nseg++] = RcSplitSegM{nb.h[1], H, H, H, H, 2 * H, Ph[2], Pl[2], rows, count, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    }
    d->njobs = nj;
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->d_phase[ph], (size_t)1);
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->d_ctl[ph], (size_t)(1 + RC_PH_MAXJOBS * MT));
    if (rc == RC_OK && cudaMemcpy(s->d_phase[ph], d, sizeof(*d), cudaMemcpyHostToDevice) != cudaSuccess) { rc_set_error("build_phase: upload failed"); rc = RC_ERR_CUDA; }
    delete d;
    s->ph_nseg[ph] = nseg; s->ph_max_tiles[ph] = max_tiles; s->ph_MT = MT;
    return rc;
}

// The grouped kernel needs the maximum shared-memory carve-out (193 KB per CTA).  A kernel with another carve-out before or after
// it makes the SMs re-partition L1 / shared memory, which drains them; asking for the same carve-out in the small per-frame
// kernels of this path avoids that switch twice per grouped launch.
int prefer_max_carveout() {
    static bool done = false;
    if (done) return RC_OK;
    const int pct = cudaSharedmemCarveoutMaxShared;
    RC_CUDA(cudaFuncSetAttribute(rc_prep_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_lists_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_mid_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_kin_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_init_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_advance_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_gemm_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RC_CUDA(cudaFuncSetAttribute(rc_gemm_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    done = true;
    return RC_OK;
}

int build_phases(rc_state* s) {
    static const bool keep_carveout = getenv("RC_NO_CARVEOUT_HINT") != nullptr;      // A/B switch
    if (!keep_carveout) RC_TRY(prefer_max_carveout());
    const rc_net* net = s->net;
    const long long Bpad = (long long)((s->B + 127) / 128) * 128;
    for (int i = 0; i < NNETS; ++i) {
        const NetDev& w = net->nets[i];
        const size_t n[4] = {(size_t)Bpad * w.K1p, (size_t)Bpad * 2 * w.H, (size_t)Bpad * 2 * w.H, (size_t)Bpad * w.H};
        for (int q = 0; q < 4; ++q) {
            RC_TRY(dev_alloc(s->allocs, &s->Phi[i][q], n[q]));
            RC_TRY(dev_alloc(s->allocs, &s->Plo[i][q], n[q]));
            RC_CUDA(cudaMemset(s->Phi[i][q], 0, n[q] * 2));
            RC_CUDA(cudaMemset(s->Plo[i][q], 0, n[q] * 2));
        }
    }
    // within a layer the chain with the longest K goes first, so the short tiles fill the tail of the phase
    const ChainSpec p1[2] = {{NET4, L_HI, s->X6 + 171, RC_K6}, {NET2, L_ALL, s->X3 + 72, RC_K3}};
    const ChainSpec p6a[1] = {{NET6, L_6A, s->Y6, 4}};
    const ChainSpec p2[4] = {{NET6, L_6B, s->Y6, 4}, {NET3, L_ALL, s->Y3, 4}, {NET7, L_ALL, s->Y7, 144}, {NET8, L_ALL, s->Y8, 4}};
    const ChainSpec pl[2] = {{NET4, L_LATE, nullptr, 0}, {NET6, L_LATE, nullptr, 0}};
    RC_TRY(build_phase(s, PH_1, p1, 2));
    RC_TRY(build_phase(s, PH_6A, p6a, 1));
    RC_TRY(build_phase(s, PH_2, p2, 4));
    RC_TRY(build_phase(s, PH_LATE, pl, 2));
    s->ph_ready = true;
    return RC_OK;
}

// Debug timeline (RC_FRAME_TIMELINE=1, stream launches only): an event after every launch of the grouped path; rc_forward_sequence
// prints the mean interval per slot over the frames of the call.
struct FrameTimeline {
    bool on = getenv("RC_FRAME_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> tag;
    size_t used = 0;
    void mark(const char* t, void* stream) {
        if (!on) return;
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); tag.push_back(t); }
        tag[used] = t;
        cudaEventRecord(ev[used++], (cudaStream_t)stream);
    }
    void report(int slots_per_frame) {
        if (!on || used < 2) { used = 0; return; }
        cudaDeviceSynchronize();
        std::map<std::string, std::pair<double, int>> acc;
        std::vector<std::string> order;
        for (size_t i = 1; i < used; ++i) {
            static const int skip = getenv("RC_TL_SKIP") ? atoi(getenv("RC_TL_SKIP")) : 2;
            if ((int)i < skip * slots_per_frame) continue;              // skip the first frames (default two; the init_net re-seeds fall into the first ~16)
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            auto& a = acc[tag[i]];
            if (a.second == 0) order.push_back(tag[i]);
            a.first += ms; a.second += 1;
        }
        if (getenv("RC_TL_FRAMES")) {                                    // frame-by-frame totals (us), ten per line
            fprintf(stderr, "[frame totals, us]");
            for (size_t f = 0; (f + 1) * slots_per_frame < used; ++f) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[f * slots_per_frame], ev[(f + 1) * slots_per_frame]);
                fprintf(stderr, "%s%.0f", f % 10 ? " " : "\n  ", 1e3 * ms);
            }
            fprintf(stderr, "\n");
        }
        double tot = 0;
        for (auto& k : order) tot += acc[k].first / acc[k].second;
        fprintf(stderr, "[frame timeline, us per frame]");
        for (auto& k : order) fprintf(stderr, " %s %.1f |", k.c_str(), 1e3 * acc[k].first / acc[k].second);
        fprintf(stderr, " total %.1f\n", 1e3 * tot);
        used = 0;
    }
};
FrameTimeline g_tl;

int run_phase(rc_state* s, int ph, void* stream, int* advance = nullptr) {
    RC_TRY(rc_tc_split_multi(s->ph_segs[ph], s->ph_nseg[ph], s->B, s->d_ctl[ph], 1 + RC_PH_MAXJOBS * s->ph_MT, stream, advance));
    g_tl.mark(ph == PH_1 ? "split1" : ph == PH_2 ? "split2" : ph == PH_LATE ? "splitL" : "split6a", stream);
    if (s->prof_on) {                  // CUDA events right around the grouped GEMM launch, on the stream it is launched on
        if (s->prof_used + 2 > s->prof_ev.size()) {
            for (int q = 0; q < 256; ++q) { cudaEvent_t e; RC_CUDA(cudaEventCreate(&e)); s->prof_ev.push_back(e); }
        }
        RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], (cudaStream_t)stream));
    }
    // (Measured: in a steady frame the init_net launches of the side stream are empty and the join costs 2.6 us; leaving SMs free for
    // them during the updater phase, RC_LATE_RESERVE, does not pay.)
    static const int late_reserve = getenv("RC_LATE_RESERVE") ? atoi(getenv("RC_LATE_RESERVE")) : 0;
    RC_TRY(rc_tc_phase(s->d_phase[ph], s->d_ctl[ph], s->ph_MT, s->ph_max_tiles[ph], stream, s->d_trace[ph], ph == PH_LATE || ph == PH_6A ? 128 : 256,
                       ph == PH_LATE ? late_reserve : 0));
    if (s->prof_on) RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], (cudaStream_t)stream));
    g_tl.mark(ph == PH_1 ? "phase1" : ph == PH_2 ? "phase2" : ph == PH_LATE ? "phaseL" : "phase6a", stream);
    return RC_OK;
}

// prep + list compaction, shared by every path
int enqueue_prep(rc_state* s, const StepIO& io, void* stream) {
    const rc_net* n = s->net;
    const int B = s->B;
    static const bool scalar_rows = getenv("RC_SCALAR_ROWS") != nullptr;     // validation switch: one-thread-per-stream kernels
    if (scalar_rows)
        RC_LAUNCH(rc_prep_kernel, rc_cdiv(B, 128), 128, 0, stream, n->cfg, s->rows, io, B, s->X2, s->X3, s->X4, s->X6, s->X7,
                  s->rcr, s->conf, s->lerpw, s->flags);
    else
        RC_LAUNCH_PDL(rc_prep_warp_kernel, rc_cdiv(B, kRowWarps), kRowWarps * 32, 0, stream, n->cfg, s->rows, io, B, s->X2, s->X3, s->X4,
                  s->X6, s->X7, s->rcr, s->conf, s->lerpw, s->flags, s->lists, s->counts);
    RC_CHECK_LAUNCH();
    if (B > 1 || scalar_rows) {
        RC_LAUNCH_PDL(rc_lists_kernel, 1, 1024, 0, stream, (const int*)s->flags, B, s->lists, s->counts);
        RC_CHECK_LAUNCH();
    }
    return RC_OK;
}

// Grouped path (gemm mode 2), first part of a frame: prep -> lists -> [rnn4 + rnn2] -> mid -> [rnn6 on first-frame rows] ->
// [rnn6 + rnn3 + rnn7 + rnn8] -> kin (rnn7 / rnn8 only need the outputs of rnn2 / rnn4 (:169-170), so they share a launch with rnn3 / rnn6).
int grouped_head(rc_state* s, const StepIO& io, int any_first_frame, void* stream) {
    const rc_net* n = s->net;
    const int B = s->B;
    static const bool scalar_rows = getenv("RC_SCALAR_ROWS") != nullptr;
    RC_TRY(enqueue_prep(s, io, stream));
    g_tl.mark("prep+lists", stream);
    RC_TRY(run_phase(s, PH_1, stream));
    // the joint blend ("mid") is fused into the second pre-pass (RcSplitSegM::mid_*); RC_MID_KERNEL=1 keeps the separate launch
    static const bool mid_kernel = getenv("RC_MID_KERNEL") != nullptr;
    if (mid_kernel) {
        RC_LAUNCH_PDL(rc_mid_kernel, rc_cdiv((long long)B * 23, 128), 128, 0, stream, (const int*)s->flags, B, (const float*)s->rcr, (const float*)s->lerpw, (const float*)s->X3, (const float*)s->X6, s->X7);
        RC_CHECK_LAUNCH();
        g_tl.mark("mid", stream);
    }
    if (any_first_frame) RC_TRY(run_phase(s, PH_6A, stream));
    RC_TRY(run_phase(s, PH_2, stream));
    if (scalar_rows)
        RC_LAUNCH(rc_kin_kernel, rc_cdiv(B, 64), 64, 0, stream, n->cfg, n->model->d_const, s->rows, s->flags, B, io, s->Y7, s->Y8,
                  s->Y3, s->Y6, s->rcr, s->conf, s->gravity, s->X4, s->X6, s->X7, s->XI, s->lists, s->counts);
    else
        RC_LAUNCH_PDL(rc_kin_warp_kernel, rc_cdiv(B, kRowWarps), kRowWarps * 32, 0, stream, n->cfg, n->model->d_const, s->rows, s->flags, B,
                  io, s->Y7, s->Y8, s->Y3, s->Y6, s->rcr, s->conf, s->gravity, s->X4, s->X6, s->X7, s->XI, s->lists, s->counts);
    RC_CHECK_LAUNCH();
    g_tl.mark("kin", stream);
    return RC_OK;
}

// Second part: init_net (re-seeds rnn2, :178-183) on the side stream — its blocks (128 registers x 256 threads) cannot share an SM
// with the resident grouped kernel, so it effectively runs after the vision updater — and the updater itself, whose pre-pass also
// advances the frame cursor.
int grouped_tail(rc_state* s, bool advance, void* stream) {
    static const bool serial2 = getenv("RC_SERIAL") != nullptr;
    if (!serial2) { RC_CUDA(cudaEventRecord(s->ev_fork, (cudaStream_t)stream)); RC_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0)); }
    RC_TRY(init_pass(s, serial2 ? stream : (void*)s->side));
    if (!serial2) RC_CUDA(cudaEventRecord(s->ev_join, s->side));
    RC_TRY(run_phase(s, PH_LATE, stream, advance ? s->d_t : nullptr));
    if (!serial2) RC_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_join, 0));
    g_tl.mark("join", stream);
    return RC_OK;
}

// ---- time-chunked host transfers (rc_forward_sequence_host) ------------------------------------------------------------------------
int chunk_before(rc_state* s, int t, cudaStream_t st) {
    const rc_state::ChunkIO& c = s->cio;
    if (!c.chunk || t % c.chunk) return RC_OK;
    RC_CUDA(cudaStreamWaitEvent(st, c.ev_in[t / c.chunk], 0));          // the chunk's inputs have landed
    return RC_OK;
}
int chunk_after(rc_state* s, int t, cudaStream_t st) {
    const rc_state::ChunkIO& c = s->cio;
    if (!c.chunk || ((t + 1) % c.chunk && t + 1 != c.T)) return RC_OK;
    const int k = t / c.chunk, t0 = k * c.chunk, n = t + 1 - t0;
    const size_t B = (size_t)s->B, T = (size_t)c.T;
    RC_CUDA(cudaEventRecord(c.ev_out[k], st));
    RC_CUDA(cudaStreamWaitEvent(c.d2h, c.ev_out[k], 0));
    RC_CUDA(cudaMemcpy2DAsync(c.hp + (size_t)t0 * 216, T * 216 * sizeof(float), s->hp + (size_t)t0 * 216, T * 216 * sizeof(float),
                              (size_t)n * 216 * sizeof(float), B, cudaMemcpyDeviceToHost, c.d2h));
    RC_CUDA(cudaMemcpy2DAsync(c.ht + (size_t)t0 * 3, T * 3 * sizeof(float), s->ht + (size_t)t0 * 3, T * 3 * sizeof(float),
                              (size_t)n * 3 * sizeof(float), B, cudaMemcpyDeviceToHost, c.d2h));
    return RC_OK;
}

bool grouped_path(const rc_state* s) { return s->net->gemm_mode >= 2 && s->ph_ready && s->B > 8; }

int enqueue_step(rc_state* s, const StepIO& io, int any_first_frame, bool advance, void* stream) {
    const rc_net* n = s->net;
    const int B = s->B;
    static const bool scalar_rows = getenv("RC_SCALAR_ROWS") != nullptr;     // validation switch: one-thread-per-stream kernels
    if (grouped_path(s)) {
        RC_TRY(grouped_head(s, io, any_first_frame, stream));
        return grouped_tail(s, advance, stream);
    }
    RC_TRY(enqueue_prep(s, io, stream));
    // fork/join helpers: the side stream runs the chain that is independent of the main one (both inside the same graph when captured)
    static const bool serial = getenv("RC_SERIAL") != nullptr;               // validation switch: single stream
    cudaStream_t ms = (cudaStream_t)stream;
    void* side = serial ? stream : (void*)s->side;
    const int sl = serial ? 0 : 1;
    auto fork = [&]() -> int { if (serial) return RC_OK; RC_CUDA(cudaEventRecord(s->ev_fork, ms)); RC_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0)); return RC_OK; };
    auto join = [&]() -> int { if (serial) return RC_OK; RC_CUDA(cudaEventRecord(s->ev_join, s->side)); RC_CUDA(cudaStreamWaitEvent(ms, s->ev_join, 0)); return RC_OK; };
    RC_TRY(fork());
    RC_TRY(net_pass(s, NET2, L_ALL, s->X2, s->X3 + 72, RC_K3, stream, 0));              // j3dr_i            (:144)
    RC_TRY(net_pass(s, NET3, L_ALL, s->X3, s->Y3, 4, stream, 0));                      // vr                (:145)
    RC_TRY(net_pass(s, NET4, L_HI, s->X4, s->X6 + 171, RC_K6, side, sl));              // j3dc              (:153)
    if (any_first_frame) RC_TRY(net_pass(s, NET6, L_6A, s->X6, s->Y6, 4, side, sl));   // pc on first_frame (:156)
    RC_TRY(net_pass(s, NET6, L_6B, s->X6, s->Y6, 4, side, sl));                        // pc                (:161,165)
    RC_TRY(join());
    RC_LAUNCH(rc_mid_kernel, rc_cdiv((long long)B * 23, 128), 128, 0, stream, s->flags, B, s->rcr, s->lerpw, s->X3, s->X6, s->X7);
    RC_CHECK_LAUNCH();
    RC_TRY(fork());
    RC_TRY(net_pass(s, NET7, L_ALL, s->X7, s->Y7, 144, stream, 0));                    // poseg6d           (:169)
    RC_TRY(net_pass(s, NET8, L_ALL, s->X7, s->Y8, 4, side, sl));                       // contact logits    (:170)
    RC_TRY(join());
    if (scalar_rows)
        RC_LAUNCH(rc_kin_kernel, rc_cdiv(B, 64), 64, 0, stream, n->cfg, n->model->d_const, s->rows, s->flags, B, io, s->Y7, s->Y8,
                  s->Y3, s->Y6, s->rcr, s->conf, s->gravity, s->X4, s->X6, s->X7, s->XI, s->lists, s->counts);
    else
        RC_LAUNCH(rc_kin_warp_kernel, rc_cdiv(B, kRowWarps), kRowWarps * 32, 0, stream, n->cfg, n->model->d_const, s->rows, s->flags, B,
                  io, s->Y7, s->Y8, s->Y3, s->Y6, s->rcr, s->conf, s->gravity, s->X4, s->X6, s->X7, s->XI, s->lists, s->counts);
    RC_CHECK_LAUNCH();
    RC_TRY(fork());
    RC_TRY(net_pass(s, NET4, L_LATE, s->X4, nullptr, 0, side, sl));                    // vision updater    (:271)
    RC_TRY(init_pass(s, stream));                                                      // (:178-183)
    RC_TRY(net_pass(s, NET6, L_LATE, s->X6, nullptr, 0, stream, 0));                   //                   (:267)
    RC_TRY(join());
    if (advance) { RC_LAUNCH(rc_advance_kernel, 1, 1, 0, stream, s->d_t); RC_CHECK_LAUNCH(); }
    return RC_OK;
}

// ---- weight packing (host) ------------------------------------------------------------------------------------
const std::vector<float>* get_staged(rc_net* n, const std::string& key, size_t numel) {
    auto it = n->staging.find(key);
    if (it == n->staging.end()) { rc_set_error("rc_net_finalize: missing tensor %s", key.c_str()); return nullptr; }
    if (it->second.size() != numel) { rc_set_error("rc_net_finalize: %s has %zu elements, expected %zu", key.c_str(), it->second.size(), numel); return nullptr; }
    return &it->second;
}

int pack_linear(rc_net* n, const std::string& prefix, int out, int in, int out_pad, int in_pad, float** dW, float** db,
                int tc_out_pad = 0, int tc_in_pad = 0, uint16_t** dWhi = nullptr, uint16_t** dWlo = nullptr) {
    const std::vector<float>* w = get_staged(n, prefix + ".weight", (size_t)out * in);
    const std::vector<float>* b = get_staged(n, prefix + ".bias", (size_t)out);
    if (!w || !b) return RC_ERR_STATE;
    std::vector<float> pw, pb;
    rc_pack_linear(w->data(), b->data(), out, in, out_pad, in_pad, pw, pb);
    RC_TRY(upload(n->allocs, dW, pw));
    RC_TRY(upload(n->allocs, db, pb));
    n->weight_bytes += (int64_t)(pw.size() + pb.size()) * 4;
    if (dWhi) {                                   // tensor-core copy: rows padded to the tile width, K to 64, split fp16
        std::vector<float> tw, tb;
        std::vector<uint16_t> hi, lo;
        rc_pack_linear(w->data(), b->data(), out, in, tc_out_pad, tc_in_pad, tw, tb);
        rc_tc_split_host(tw.data(), tw.size(), hi, lo);
        RC_TRY(dev_alloc(n->allocs, dWhi, hi.size()));
        RC_TRY(dev_alloc(n->allocs, dWlo, lo.size()));
        RC_CUDA(cudaMemcpy(*dWhi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
        RC_CUDA(cudaMemcpy(*dWlo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    }
    return RC_OK;
}

int pack_lstm(rc_net* n, const std::string& prefix, int layer, int H, float** dW, float** db, uint16_t** dWhi, uint16_t** dWlo) {
    const std::string sfx = "_l" + std::to_string(layer);
    const std::vector<float>* wih = get_staged(n, prefix + ".rnn.weight_ih" + sfx, (size_t)4 * H * H);
    const std::vector<float>* whh = get_staged(n, prefix + ".rnn.weight_hh" + sfx, (size_t)4 * H * H);
    const std::vector<float>* bih = get_staged(n, prefix + ".rnn.bias_ih" + sfx, (size_t)4 * H);
    const std::vector<float>* bhh = get_staged(n, prefix + ".rnn.bias_hh" + sfx, (size_t)4 * H);
    if (!wih || !whh || !bih || !bhh) return RC_ERR_STATE;
    std::vector<float> pw, pb;
    rc_pack_lstm(wih->data(), whh->data(), bih->data(), bhh->data(), H, pw, pb);
    RC_TRY(upload(n->allocs, dW, pw));
    RC_TRY(upload(n->allocs, db, pb));
    n->weight_bytes += (int64_t)(pw.size() + pb.size()) * 4;
    if (dWhi) {
        std::vector<uint16_t> hi, lo;
        rc_tc_split_host(pw.data(), pw.size(), hi, lo);
        RC_TRY(dev_alloc(n->allocs, dWhi, hi.size()));
        RC_TRY(dev_alloc(n->allocs, dWlo, lo.size()));
        RC_CUDA(cudaMemcpy(*dWhi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
        RC_CUDA(cudaMemcpy(*dWlo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    }
    return RC_OK;
}

}  // namespace

extern "C" {

void rc_net_default_config(rc_net_config* c, int live) {
    if (!c) return;
    c->conf_lo = live ? 0.85 : 0.7;
    c->conf_hi = live ? 0.9 : 0.8;
    c->tran_filter_num = live ? 0.01 : 0.05;
    c->contact_threshold = 0.7f;
    c->height_threshold = 0.15f;
    c->distance_threshold = 10.f;
    c->use_flat_floor = 1;
    c->live = live ? 1 : 0;
    c->update_vision_freq = 30;
}

static void apply_cfg(rc_net* n, const rc_net_config* c) {
    n->cfg.conf_lo = c->conf_lo; n->cfg.conf_hi = c->conf_hi; n->cfg.tran_filter = c->tran_filter_num;
    n->cfg.contact_thr = c->contact_threshold; n->cfg.height_thr = c->height_threshold; n->cfg.dist_thr = c->distance_threshold;
    n->cfg.use_flat_floor = c->use_flat_floor; n->cfg.live = c->live; n->cfg.update_vision_freq = c->update_vision_freq;
}

int rc_net_create(rc_net** out, const rc_model* model, const rc_net_config* cfg) {
    RC_ARG(out && model);
    rc_net* n = new rc_net();
    n->model = model;
    rc_net_config def;
    rc_net_default_config(&def, 0);
    apply_cfg(n, cfg ? cfg : &def);
    if (getenv("RC_SEQ_AUTO_B")) n->seq_auto_B = atoi(getenv("RC_SEQ_AUTO_B"));      // A/B switches for bench runs
    if (getenv("RC_SEQ_WARM")) n->seq_warm = std::max(1, atoi(getenv("RC_SEQ_WARM")));
    *out = n;
    return RC_OK;
}

int rc_net_set_config(rc_net* n, const rc_net_config* cfg) {
    RC_ARG(n && cfg);
    apply_cfg(n, cfg);
    n->cfg_version += 1;        // cached CUDA graphs of every state of this net hold the old config by value: they re-capture
    return RC_OK;
}

void rc_net_destroy(rc_net* n) {
    if (!n) return;
    for (void* p : n->allocs) cudaFree(p);
    delete n;
}

int rc_net_set_tensor(rc_net* n, const char* key, const float* data, int64_t numel) {
    RC_ARG(n && key && data && numel > 0);
    if (n->finalized) { rc_set_error("rc_net_set_tensor after rc_net_finalize"); return RC_ERR_STATE; }
    n->staging[key].assign(data, data + numel);
    return RC_OK;
}

int rc_net_finalize(rc_net* n) {
    RC_ARG(n);
    if (n->finalized) return RC_OK;
    for (int i = 0; i < NNETS; ++i) {
        NetDev& d = n->nets[i];
        d.in = kNetIn[i]; d.K1 = kNetK1[i]; d.H = kNetH[i]; d.out = kNetOut[i]; d.out4 = (d.out + 3) / 4 * 4;
        const std::string p = "rnn" + std::to_string(kNetId[i]);
        d.K1p = (d.K1 + 63) / 64 * 64;
        d.outp = (d.out + RC_TC_BN - 1) / RC_TC_BN * RC_TC_BN;
        RC_TRY(pack_linear(n, p + ".linear1", d.H, d.in, d.H, d.K1, &d.W1, &d.b1, d.H, d.K1p, &d.W1hi, &d.W1lo));
        for (int l = 0; l < 2; ++l) RC_TRY(pack_lstm(n, p, l, d.H, &d.WL[l], &d.bL[l], &d.WLhi[l], &d.WLlo[l]));
        n->tc_ready = true;
        for (int l = 0; l < 2 && n->tc_ready; ++l) {
            if (rc_tc_make_map(&d.mWhi[l], d.WLhi[l], 4LL * d.H, 2 * d.H, RC_TC_BN) != RC_OK ||
                rc_tc_make_map(&d.mWlo[l], d.WLlo[l], 4LL * d.H, 2 * d.H, RC_TC_BN) != RC_OK ||
                rc_tc_make_map(&d.mWhi64[l], d.WLhi[l], 4LL * d.H, 2 * d.H, 64) != RC_OK ||
                rc_tc_make_map(&d.mWlo64[l], d.WLlo[l], 4LL * d.H, 2 * d.H, 64) != RC_OK) n->tc_ready = false;
        }
        RC_TRY(pack_linear(n, p + ".linear2", d.out, d.H, d.out4, d.H, &d.W2, &d.b2, d.outp, d.H, &d.W2hi, &d.W2lo));
        if (n->tc_ready) {
            if (rc_tc_make_map(&d.mW1hi, d.W1hi, d.H, d.K1p, RC_TC_BN) != RC_OK || rc_tc_make_map(&d.mW1lo, d.W1lo, d.H, d.K1p, RC_TC_BN) != RC_OK ||
                rc_tc_make_map(&d.mW2hi, d.W2hi, d.outp, d.H, RC_TC_BN) != RC_OK || rc_tc_make_map(&d.mW2lo, d.W2lo, d.outp, d.H, RC_TC_BN) != RC_OK ||
                rc_tc_make_map(&d.mW1hi64, d.W1hi, d.H, d.K1p, 64) != RC_OK || rc_tc_make_map(&d.mW1lo64, d.W1lo, d.H, d.K1p, 64) != RC_OK ||
                rc_tc_make_map(&d.mW2hi64, d.W2hi, d.outp, d.H, 64) != RC_OK || rc_tc_make_map(&d.mW2lo64, d.W2lo, d.outp, d.H, 64) != RC_OK)
                n->tc_ready = false;
        }
    }
    const int64_t per_frame = n->weight_bytes;
    const int kin[3] = {kInitK0, 512, 1024};
    for (int l = 0; l < 3; ++l)
        RC_TRY(pack_linear(n, "rnn2.init_net." + std::to_string(2 * l), kInitDims[l + 1], kInitDims[l], kInitDims[l + 1], kin[l], &n->Wi[l], &n->bi[l]));
    n->weight_bytes = per_frame;
    n->staging.clear();
    n->finalized = true;
    return RC_OK;
}

int64_t rc_net_weight_bytes(const rc_net* n) { return n ? n->weight_bytes : 0; }

int rc_net_set_seq_options(rc_net* n, int auto_max_streams, int warm_frames) {
    RC_ARG(n);
    if (auto_max_streams >= 0) n->seq_auto_B = auto_max_streams;
    if (warm_frames >= 1) n->seq_warm = warm_frames;
    return RC_OK;
}

int rc_net_set_gemm_mode(rc_net* n, int mode) {
    RC_ARG(n && mode >= 0 && mode <= 3);
    if (mode >= 1 && !n->tc_ready) { rc_set_error("tensor-core path unavailable (tensor maps could not be created)"); return RC_ERR_STATE; }
    n->gemm_mode = mode;
    return RC_OK;
}

int rc_state_create(rc_state** out, const rc_net* net, int32_t B) {
    RC_ARG(out && net && B > 0);
    if (!net->finalized) { rc_set_error("rc_state_create: net not finalized"); return RC_ERR_STATE; }
    rc_state* s = new rc_state();
    s->net = net; s->B = B;
    int rc = RC_OK;
    auto A = [&](float** p, size_t n) { if (rc == RC_OK) rc = dev_alloc(s->allocs, p, n); };
    for (int i = 0; i < NNETS; ++i) {
        const size_t n = (size_t)B * net->nets[i].H;
        for (int l = 0; l < 2; ++l) { A(&s->nb[i].h[l], n); A(&s->nb[i].c[l], n); A(&s->nb[i].hn[l], n); }
        A(&s->nb[i].a1, n);
    }
    A(&s->X2, (size_t)B * RC_K2); A(&s->X3, (size_t)B * RC_K3); A(&s->X4, (size_t)B * RC_K4); A(&s->X6, (size_t)B * RC_K6);
    A(&s->X7, (size_t)B * RC_K7); A(&s->XI, (size_t)B * kInitK0);
    A(&s->Y3, (size_t)B * 4); A(&s->Y6, (size_t)B * 4); A(&s->Y7, (size_t)B * 144); A(&s->Y8, (size_t)B * 4);
    A(&s->I1, (size_t)B * 512); A(&s->I2, (size_t)B * 1024); A(&s->I3, (size_t)B * 2048);
    A(&s->rcr, (size_t)B * 9); A(&s->conf, (size_t)B); A(&s->lerpw, (size_t)B * 2); A(&s->gravity, 4);
    if (rc == RC_OK && net->tc_ready && B > 8) {
        const long long Bpad = (long long)((B + 127) / 128) * 128;
        int Hmax = 0;
        for (int i = 0; i < NNETS; ++i) Hmax = std::max(Hmax, net->nets[i].H);
        s->tc_ready = true;
        for (int ln = 0; ln < 2 && rc == RC_OK; ++ln) {
            for (int q = 0; q < 3 && rc == RC_OK; ++q) {
                rc = dev_alloc(s->allocs, &s->Ahi[ln][q], (size_t)Bpad * 2 * Hmax);
                if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->Alo[ln][q], (size_t)Bpad * 2 * Hmax);
                if (rc == RC_OK) { cudaMemset(s->Ahi[ln][q], 0, (size_t)Bpad * 2 * Hmax * 2); cudaMemset(s->Alo[ln][q], 0, (size_t)Bpad * 2 * Hmax * 2); }
            }
            if (rc != RC_OK) break;
            for (int i = 0; i < NNETS && s->tc_ready; ++i) {
                const NetDev& d = net->nets[i];
                auto mk = [&](RcTensorMap* m, const void* base, int K) { return rc_tc_make_map(m, base, Bpad, K, 128) == RC_OK; };
                if (!mk(&s->mA0hi[ln][i], s->Ahi[ln][0], d.K1p) || !mk(&s->mA0lo[ln][i], s->Alo[ln][0], d.K1p) ||
                    !mk(&s->mA1hi[ln][i], s->Ahi[ln][1], 2 * d.H) || !mk(&s->mA1lo[ln][i], s->Alo[ln][1], 2 * d.H) ||
                    !mk(&s->mA1Hhi[ln][i], s->Ahi[ln][1], d.H) || !mk(&s->mA1Hlo[ln][i], s->Alo[ln][1], d.H) ||
                    !mk(&s->mA2hi[ln][i], s->Ahi[ln][2], 2 * d.H) || !mk(&s->mA2lo[ln][i], s->Alo[ln][2], 2 * d.H))
                    s->tc_ready = false;
            }
        }
    }
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->flags, (size_t)B);
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->lists, (size_t)NLISTS * B);
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->counts, (size_t)NLISTS);
    if (rc == RC_OK && s->tc_ready) rc = build_phases(s);
    if (rc == RC_OK) {
        if (cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess) { rc_set_error("stream/event creation failed"); rc = RC_ERR_CUDA; }
    }
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->d_t, (size_t)1);
    if (rc == RC_OK) rc = dev_alloc(s->allocs, &s->rows, (size_t)B);
    if (rc != RC_OK) { rc_state_destroy(s); return rc; }
    const float g[4] = {-0.0029f, 0.9980f, -0.0273f, 0.f};     // Net.gravityc default (sig_mp.py:36)
    cudaError_t e = cudaMemcpy(s->gravity, g, sizeof(g), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { rc_set_error("rc_state_create: %s", cudaGetErrorString(e)); rc_state_destroy(s); return RC_ERR_CUDA; }
    rc = rc_state_reset(s, nullptr);
    if (rc == RC_OK && cudaDeviceSynchronize() != cudaSuccess) rc = RC_ERR_CUDA;
    if (rc != RC_OK) { rc_state_destroy(s); return rc; }
    *out = s;
    return RC_OK;
}

void rc_state_destroy(rc_state* s) {
    if (!s) return;
    if (s->graph) cudaGraphExecDestroy(s->graph);
    if (s->on_graph) cudaGraphExecDestroy(s->on_graph);
    cudaFree(s->sk_bar); cudaFree(s->sk_rows); cudaFree(s->s2_bar); cudaFree(s->s2_ts);
    cudaFree(s->on_din); cudaFree(s->on_dout); cudaFreeHost(s->on_hin); cudaFreeHost(s->on_hout);
    rc_seq_destroy(s);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
    for (void* p : s->allocs) cudaFree(p);
    cudaFree(s->hj); cudaFree(s->ha); cudaFree(s->ho); cudaFree(s->hp); cudaFree(s->ht); cudaFree(s->hft);
    cudaFree(s->hlen); cudaFree(s->hfl);
    if (s->cio.h2d) cudaStreamDestroy(s->cio.h2d);
    if (s->cio.d2h) cudaStreamDestroy(s->cio.d2h);
    for (cudaEvent_t e : s->cio.ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : s->cio.ev_out) cudaEventDestroy(e);
    delete s;
}

int rc_state_reset(rc_state* s, void* stream) {
    RC_ARG(s);
    cudaStream_t st = (cudaStream_t)stream;
    const rc_net* n = s->net;
    for (int i = 0; i < NNETS; ++i) {
        const size_t bytes = (size_t)s->B * n->nets[i].H * sizeof(float);
        for (int l = 0; l < 2; ++l) {
            RC_CUDA(cudaMemsetAsync(s->nb[i].h[l], 0, bytes, st));
            RC_CUDA(cudaMemsetAsync(s->nb[i].c[l], 0, bytes, st));
            RC_CUDA(cudaMemsetAsync(s->nb[i].hn[l], 0, bytes, st));
        }
    }
    RC_CUDA(cudaMemsetAsync(s->Y3, 0, (size_t)s->B * 4 * sizeof(float), st));
    RC_CUDA(cudaMemsetAsync(s->Y6, 0, (size_t)s->B * 4 * sizeof(float), st));
    RC_CUDA(cudaMemsetAsync(s->X6, 0, (size_t)s->B * RC_K6 * sizeof(float), st));
    RC_CUDA(cudaMemsetAsync(s->X4, 0, (size_t)s->B * RC_K4 * sizeof(float), st));
    RC_CUDA(cudaMemsetAsync(s->d_t, 0, sizeof(int), st));
    RC_CUDA(cudaMemsetAsync(s->counts, 0, NLISTS * sizeof(int), st));
    RC_LAUNCH(rc_reset_rows_kernel, rc_cdiv(s->B, 128), 128, 0, stream, s->rows, s->B, s->fresh ? 1 : 0);
    s->fresh = false;
    RC_CHECK_LAUNCH();
    return RC_OK;
}

int rc_state_set_gravity(rc_state* s, const float* g, void* stream) {
    RC_ARG(s && g);
    RC_CUDA(cudaMemcpyAsync(s->gravity, g, 3 * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    RC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return RC_OK;
}

int rc_forward_step(rc_state* s, const float* j2dc, const float* accc, const float* oric, const float* gravity,
                    const float* first_tran, const int32_t* row_flags, int any_first_frame, float* pose, float* tran,
                    void* stream) {
    RC_ARG(s && j2dc && accc && oric && pose && tran);
    StepIO io;
    io.j2dc = j2dc; io.accc = accc; io.oric = oric; io.sj = 99; io.sa = 18; io.so = 54;
    io.gravity = gravity; io.first_tran = first_tran; io.row_flags = row_flags; io.lengths = nullptr;
    io.pose = pose; io.tran = tran; io.sp = 216; io.st = 3; io.d_t = nullptr; io.first_mode = 1;
    return enqueue_step(s, io, any_first_frame, false, stream);
}

// Layout of the single-frame staging buffer: j2dc[99] | accc[18] | oric[54] | first_tran[3] | flags (int) | pad
constexpr int kOnIn = 176, kOnOut = 220;

int rc_forward_online(rc_state* s, const float* j2dc, const float* accc, const float* oric, const float* first_tran,
                      int first_frame, int inputs_on_device, float* h_pose, float* h_tran, void* stream) {
    RC_ARG(s && s->B == 1 && j2dc && accc && oric && h_pose && h_tran);
    cudaStream_t st = (cudaStream_t)stream;
    if (!s->on_din) {
        RC_CUDA(cudaMalloc(&s->on_din, kOnIn * sizeof(float)));
        RC_CUDA(cudaMalloc(&s->on_dout, kOnOut * sizeof(float)));
        RC_CUDA(cudaMallocHost(&s->on_hin, kOnIn * sizeof(float)));
        RC_CUDA(cudaMallocHost(&s->on_hout, kOnOut * sizeof(float)));
    }
    int flags = (first_frame ? RC_ROW_FIRST_FRAME : 0) | (first_tran ? RC_ROW_FIRST_TRAN : 0);
    if (inputs_on_device) {
        RC_CUDA(cudaMemcpyAsync(s->on_din, j2dc, 99 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        RC_CUDA(cudaMemcpyAsync(s->on_din + 99, accc, 18 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        RC_CUDA(cudaMemcpyAsync(s->on_din + 117, oric, 54 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (first_tran) RC_CUDA(cudaMemcpyAsync(s->on_din + 171, first_tran, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        memcpy(s->on_hin + 174, &flags, sizeof(int));
        RC_CUDA(cudaMemcpyAsync(s->on_din + 174, s->on_hin + 174, sizeof(int), cudaMemcpyHostToDevice, st));
    } else {
        memcpy(s->on_hin, j2dc, 99 * sizeof(float));
        memcpy(s->on_hin + 99, accc, 18 * sizeof(float));
        memcpy(s->on_hin + 117, oric, 54 * sizeof(float));
        if (first_tran) memcpy(s->on_hin + 171, first_tran, 3 * sizeof(float));
        memcpy(s->on_hin + 174, &flags, sizeof(int));
        // (Tried: no copy — the single-launch kernel reading the frame straight out of the pinned host buffer, as it writes pose / tran into
        // it: 148 CTAs fetching the 700 bytes over PCIe at kernel start cost +27 us of p50, 168-172 vs 141-143 us.)
        RC_CUDA(cudaMemcpyAsync(s->on_din, s->on_hin, kOnIn * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    StepIO io;
    io.j2dc = s->on_din; io.accc = s->on_din + 99; io.oric = s->on_din + 117; io.sj = 99; io.sa = 18; io.so = 54;
    io.gravity = nullptr; io.first_tran = s->on_din + 171; io.row_flags = (const int*)(s->on_din + 174); io.lengths = nullptr;
    io.pose = s->on_dout; io.tran = s->on_dout + 216; io.sp = 216; io.st = 3; io.d_t = nullptr; io.first_mode = 1;
    bool direct_out = false;
    if (rc_stream_supported(s)) {
        RC_TRY(rc_stream_frame(s, io, first_frame, stream));    // the whole frame as one cooperative kernel (stream.cu)
    } else if (!first_frame && rc_stream2_supported(s)) {
        // one kernel, weights TMA-staged through shared memory (stream2.cu); kin writes pose / tran straight into the pinned host
        // buffer (unified addressing), which saves the device-to-host copy on the latency path
        io.pose = s->on_hout; io.tran = s->on_hout + 216;
        direct_out = true;
        RC_TRY(rc_stream2_frame(s, io, 0, stream));
    } else if (first_frame) {
        RC_TRY(enqueue_step(s, io, 1, false, stream));          // extra rnn6 pass (sig_mp.py:155-156): direct launches
    } else {
        if (!s->on_graph || s->on_graph_stream != stream || s->on_graph_cfg_version != s->net->cfg_version) {
            if (s->on_graph) { cudaGraphExecDestroy(s->on_graph); s->on_graph = nullptr; }
            cudaGraph_t g = nullptr;
            if (!s->cap_stream) RC_CUDA(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
            RC_CUDA(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
            const long long before = g_rc_launches.load();
            const int rc = enqueue_step(s, io, 0, false, (void*)s->cap_stream);
            s->on_graph_nodes = g_rc_launches.load() - before;
            g_rc_launches.fetch_sub(s->on_graph_nodes);
            cudaError_t e = cudaStreamEndCapture(s->cap_stream, &g);
            if (rc != RC_OK) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) { rc_set_error("graph capture: %s", cudaGetErrorString(e)); return RC_ERR_CUDA; }
            e = cudaGraphInstantiate(&s->on_graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { s->on_graph = nullptr; rc_set_error("graph instantiate: %s", cudaGetErrorString(e)); return RC_ERR_CUDA; }
            s->on_graph_stream = stream;
            s->on_graph_cfg_version = s->net->cfg_version;
        }
        RC_CUDA(cudaGraphLaunch(s->on_graph, st));
        g_rc_launches.fetch_add(s->on_graph_nodes);
    }
    if (!direct_out) RC_CUDA(cudaMemcpyAsync(s->on_hout, s->on_dout, kOnOut * sizeof(float), cudaMemcpyDeviceToHost, st));
    RC_CUDA(cudaStreamSynchronize(st));
    memcpy(h_pose, s->on_hout, 216 * sizeof(float));
    memcpy(h_tran, s->on_hout + 216, 3 * sizeof(float));
    return RC_OK;
}

int rc_forward_sequence(rc_state* s, int32_t T, const float* j2dc, const float* accc, const float* oric,
                        const int32_t* lengths, const float* gravity, const float* first_tran, const int32_t* row_flags,
                        int any_first_frame, float* pose, float* tran, int use_graph, void* stream) {
    RC_ARG(s && T >= 0);
    if (T == 0) return RC_OK;
    RC_ARG(j2dc && accc && oric && pose && tran);
    cudaStream_t st = (cudaStream_t)stream;
    RC_TRY(rc_state_reset(s, stream));
    StepIO io;
    io.j2dc = j2dc; io.accc = accc; io.oric = oric;
    io.sj = (long long)T * 99; io.sa = (long long)T * 18; io.so = (long long)T * 54;
    io.gravity = gravity; io.first_tran = first_tran; io.row_flags = row_flags; io.lengths = lengths;
    io.pose = pose; io.tran = tran; io.sp = (long long)T * 216; io.st = (long long)T * 3;
    io.d_t = s->d_t; io.first_mode = 2;
    io.branch = s->branch_log; io.sb = T;
    // (Tried: rotating the loop so that init_net of frame t overlaps prep + lists of frame t + 1 — its four launches take ~37 us on the
    // side stream even when the list is empty, longer than prep + lists, so nothing was gained: 525 vs 527 us per frame.)
    RC_TRY(chunk_before(s, 0, st));
    RC_TRY(enqueue_step(s, io, any_first_frame, true, stream));
    RC_TRY(chunk_after(s, 0, st));
    if (T == 1) return RC_OK;
    if (s->B == 1 && !lengths && rc_stream2_supported(s) && !rc_stream_supported(s)) {
        // a single stream: one TMA-staged cooperative kernel per frame (stream2.cu)
        for (int t = 1; t < T; ++t) RC_TRY(rc_stream2_frame(s, io, t, stream));
        return RC_OK;
    }
    // Sequence kernel: forced by gemm mode 3; chosen automatically in the default mode 2 for small shards (<= 128 streams = one
    // row block, e.g. the 8-GPU split of 1024 sequences), where a frame is bound by its dependency chain and the 13 launch
    // boundaries of the multi-launch path cost more than the masked (uncompacted) passes of the sequence kernel.
    const bool want_seq = s->net->gemm_mode == 3 || (s->net->gemm_mode == 2 && s->B <= s->net->seq_auto_B && T >= 3 * s->net->seq_warm);
    if (want_seq && grouped_path(s) && rc_seq_supported(s)) {
        // frames 1 .. T-1 in ONE launch: persistent sequence kernel (tile width 64 for small shards, where the frame is bound by the
        // depth of its dependency chain; 128 where it is bound by the tensor pipe)
        static const int bn_env = getenv("RC_SEQ_BN") ? atoi(getenv("RC_SEQ_BN")) : 0;
        const int bn = (bn_env == 64 || bn_env == 128) ? bn_env : (s->B <= 256 ? 64 : 128);
        // The first frames of a sequence are where most streams reach c >= hi for the first time and re-seed rnn2 through init_net
        // (:178-183); that MLP runs on the tensor cores in the multi-launch path but only as a (correct, slow) per-stream row job
        // inside the sequence kernel, so the sequence kernel takes over after a few frames.
        const int t0 = std::max(1, std::min(s->net->seq_warm, T - 1));
        for (int t = 1; t < t0; ++t) RC_TRY(enqueue_step(s, io, 0, true, stream));
        return rc_seq_run(s, io, T, t0, bn, stream);
    }
    if (!use_graph) {
        for (int t = 1; t < T; ++t) {
            RC_TRY(chunk_before(s, t, st));
            RC_TRY(enqueue_step(s, io, 0, true, stream));
            RC_TRY(chunk_after(s, t, st));
        }
        g_tl.report(11);
        return RC_OK;
    }
    std::vector<const void*> key = {j2dc, accc, oric, lengths, gravity, first_tran, row_flags, pose, tran,
                                    (const void*)(intptr_t)T, (const void*)stream, (const void*)(intptr_t)s->net->gemm_mode,
                                    (const void*)(intptr_t)s->net->cfg_version, (const void*)s->branch_log};
    if (!s->graph || key != s->graph_key) {
        if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
        cudaGraph_t g = nullptr;
        if (!s->cap_stream) RC_CUDA(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
        RC_CUDA(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
        const long long before = g_rc_launches.load();
        const int rc = enqueue_step(s, io, 0, true, (void*)s->cap_stream);
        s->graph_nodes = g_rc_launches.load() - before;
        g_rc_launches.fetch_sub(s->graph_nodes);          // captured, not launched
        cudaError_t e = cudaStreamEndCapture(s->cap_stream, &g);
        if (rc != RC_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) { rc_set_error("graph capture: %s", cudaGetErrorString(e)); return RC_ERR_CUDA; }
        e = cudaGraphInstantiate(&s->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { s->graph = nullptr; rc_set_error("graph instantiate: %s", cudaGetErrorString(e)); return RC_ERR_CUDA; }
        s->graph_key = key;
    }
    for (int t = 1; t < T; ++t) {
        RC_TRY(chunk_before(s, t, st));
        RC_CUDA(cudaGraphLaunch(s->graph, st));
        RC_TRY(chunk_after(s, t, st));
    }
    g_rc_launches.fetch_add(s->graph_nodes * (long long)(T - 1));
    return RC_OK;
}

int rc_forward_sequence_host(rc_state* s, int32_t T, const float* hj, const float* ha, const float* ho,
                             const int32_t* hlen, const float* hft, const int32_t* hfl, float* hp, float* ht,
                             int use_graph, void* stream) {
    RC_ARG(s && T > 0 && hj && ha && ho && hp && ht);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = (size_t)s->B;
    if (s->host_cap_T < T) {
        cudaFree(s->hj); cudaFree(s->ha); cudaFree(s->ho); cudaFree(s->hp); cudaFree(s->ht);
        s->hj = s->ha = s->ho = s->hp = s->ht = nullptr;
        s->host_cap_T = 0;
        if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
        RC_CUDA(cudaMalloc(&s->hj, B * T * 99 * sizeof(float)));
        RC_CUDA(cudaMalloc(&s->ha, B * T * 18 * sizeof(float)));
        RC_CUDA(cudaMalloc(&s->ho, B * T * 54 * sizeof(float)));
        RC_CUDA(cudaMalloc(&s->hp, B * T * 216 * sizeof(float)));
        RC_CUDA(cudaMalloc(&s->ht, B * T * 3 * sizeof(float)));
        if (!s->hft) RC_CUDA(cudaMalloc(&s->hft, B * 3 * sizeof(float)));
        if (!s->hlen) RC_CUDA(cudaMalloc(&s->hlen, B * sizeof(int)));
        if (!s->hfl) RC_CUDA(cudaMalloc(&s->hfl, B * sizeof(int)));
        s->host_cap_T = T;
    }
    // Multi-launch path (more than 128 streams): the frames are launched one by one, so the transfers are cut into chunks of frames —
    // inputs of chunk k + 1 upload and results of chunk k - 1 download while chunk k computes (two copy streams, strided 2-D copies
    // of the [B, T, ...] slabs).  The one-launch paths (sequence kernel, single stream) copy everything before / after.
    static const int chunk_env = getenv("RC_HOST_CHUNK") ? atoi(getenv("RC_HOST_CHUNK")) : 30;
    const bool want_seq = s->net->gemm_mode == 3 || (s->net->gemm_mode == 2 && s->B <= s->net->seq_auto_B && T >= 3 * s->net->seq_warm);
    const bool chunked = chunk_env > 0 && grouped_path(s) && !(want_seq && rc_seq_supported(s)) && T > chunk_env;
    rc_state::ChunkIO& c = s->cio;
    c.chunk = 0;
    if (hft) RC_CUDA(cudaMemcpyAsync(s->hft, hft, B * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (hlen) RC_CUDA(cudaMemcpyAsync(s->hlen, hlen, B * sizeof(int), cudaMemcpyHostToDevice, st));
    int any_ff = 0;
    if (hfl) {
        RC_CUDA(cudaMemcpyAsync(s->hfl, hfl, B * sizeof(int), cudaMemcpyHostToDevice, st));
        for (size_t b = 0; b < B; ++b) any_ff |= (hfl[b] & RC_ROW_FIRST_FRAME);
    }
    if (hlen) {        // frames beyond a sequence's length are never written: they must read back as zeros, not stale data
        RC_CUDA(cudaMemsetAsync(s->hp, 0, B * T * 216 * sizeof(float), st));
        RC_CUDA(cudaMemsetAsync(s->ht, 0, B * T * 3 * sizeof(float), st));
    }
    if (chunked) {
        if (!c.h2d) {
            RC_CUDA(cudaStreamCreateWithFlags(&c.h2d, cudaStreamNonBlocking));
            RC_CUDA(cudaStreamCreateWithFlags(&c.d2h, cudaStreamNonBlocking));
        }
        const int nchunk = (T + chunk_env - 1) / chunk_env;
        while ((int)c.ev_in.size() < nchunk) {
            cudaEvent_t a, b2;
            RC_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            RC_CUDA(cudaEventCreateWithFlags(&b2, cudaEventDisableTiming));
            c.ev_in.push_back(a); c.ev_out.push_back(b2);
        }
        // the copy streams start after what is already queued on the caller's stream (buffer reuse by a previous call)
        RC_CUDA(cudaEventRecord(c.ev_out[0], st));
        RC_CUDA(cudaStreamWaitEvent(c.h2d, c.ev_out[0], 0));
        RC_CUDA(cudaStreamWaitEvent(c.d2h, c.ev_out[0], 0));
        for (int k = 0; k < nchunk; ++k) {
            const size_t t0 = (size_t)k * chunk_env, n = std::min<size_t>(chunk_env, T - t0);
            RC_CUDA(cudaMemcpy2DAsync(s->hj + t0 * 99, (size_t)T * 99 * 4, hj + t0 * 99, (size_t)T * 99 * 4, n * 99 * 4, B, cudaMemcpyHostToDevice, c.h2d));
            RC_CUDA(cudaMemcpy2DAsync(s->ha + t0 * 18, (size_t)T * 18 * 4, ha + t0 * 18, (size_t)T * 18 * 4, n * 18 * 4, B, cudaMemcpyHostToDevice, c.h2d));
            RC_CUDA(cudaMemcpy2DAsync(s->ho + t0 * 54, (size_t)T * 54 * 4, ho + t0 * 54, (size_t)T * 54 * 4, n * 54 * 4, B, cudaMemcpyHostToDevice, c.h2d));
            RC_CUDA(cudaEventRecord(c.ev_in[k], c.h2d));
        }
        c.chunk = chunk_env; c.T = T; c.hp = hp; c.ht = ht;
    } else {
        RC_CUDA(cudaMemcpyAsync(s->hj, hj, B * T * 99 * sizeof(float), cudaMemcpyHostToDevice, st));
        RC_CUDA(cudaMemcpyAsync(s->ha, ha, B * T * 18 * sizeof(float), cudaMemcpyHostToDevice, st));
        RC_CUDA(cudaMemcpyAsync(s->ho, ho, B * T * 54 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    const int rc = rc_forward_sequence(s, T, s->hj, s->ha, s->ho, hlen ? s->hlen : nullptr, nullptr, hft ? s->hft : nullptr,
                                       hfl ? s->hfl : nullptr, any_ff, s->hp, s->ht, use_graph, stream);
    c.chunk = 0;
    if (rc != RC_OK) return rc;
    if (chunked) {
        RC_CUDA(cudaStreamSynchronize(c.d2h));
    } else {
        RC_CUDA(cudaMemcpyAsync(hp, s->hp, B * T * 216 * sizeof(float), cudaMemcpyDeviceToHost, st));
        RC_CUDA(cudaMemcpyAsync(ht, s->ht, B * T * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    RC_CUDA(cudaStreamSynchronize(st));
    return RC_OK;
}

// Test tap: run ONE fused LSTM layer (sub-net ni in 0..5 = rnn2,3,4,6,7,8; layer 0/1) on caller data for all B rows of the
// state: x [B,H], h_prev [B,H], c [B,H] (updated in place), h_out [B,H]; mode 0 = fp32 SIMT/GEMV, 1 = tcgen05.
int rc_state_debug_lstm(rc_state* s, int ni, int layer, int mode, const float* x, const float* hprev, float* c, float* hout, void* stream) {
    RC_ARG(s && ni >= 0 && ni < NNETS && (layer == 0 || layer == 1) && x && hprev && c && hout);
    const NetDev& w = s->net->nets[ni];
    const int B = s->B;
    std::vector<int> ident(B);
    for (int i = 0; i < B; ++i) ident[i] = i;
    RC_CUDA(cudaMemcpyAsync(s->lists + (size_t)L_ALL * B, ident.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    RC_CUDA(cudaMemcpyAsync(s->counts + L_ALL, &B, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    RC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const int* rows = s->lists + (size_t)L_ALL * B;
    const int* count = s->counts + L_ALL;
    if (mode == 1) {
        if (!(s->tc_ready && B > 8)) { rc_set_error("tensor-core path not available for this state"); return RC_ERR_STATE; }
        RC_TRY(rc_tc_split_rows(x, w.H, hprev, w.H, w.H, w.H, 2 * w.H, rows, count, B, s->Ahi[0][1], s->Alo[0][1], stream));
        RC_TRY(rc_tc_lstm_layer(&s->mA1hi[0][ni], &s->mA1lo[0][ni], &w.mWhi[layer], &w.mWlo[layer], w.bL[layer], c, hout, w.H, rows, count, B, stream));
    } else {
        RcLinear a;
        memset(&a, 0, sizeof(a));
        a.rows = rows; a.count = count;
        a.X = x; a.ldx = w.H; a.X2 = hprev; a.ldx2 = w.H; a.K1 = w.H; a.K2 = w.H;
        a.W = w.WL[layer]; a.bias = w.bL[layer]; a.N = 4 * w.H; a.Nw = 4 * w.H; a.C = c; a.Hout = hout; a.H = w.H;
        RC_TRY(launch_linear(a, B, true, stream));
    }
    return RC_OK;
}

// Debug / parity aid: `log` (device, int32 [B, T], caller-owned, or null to switch off) receives per frame the RcBranch bits of the
// data-dependent decisions of the translation / contact / floor logic (net/sig_mp.py:185-225) taken by rc_forward_sequence.
int rc_state_set_branch_log(rc_state* s, int32_t* log) {
    RC_ARG(s);
    s->branch_log = log;
    return RC_OK;
}

int rc_state_debug_seq_stats(rc_state* s, long long* out, int32_t max_jobs) {
    if (!s || !out) return 0;
    return rc_seq_stats(s, out, max_jobs);
}

int rc_profile_enable(rc_state* s, int on) {
    RC_ARG(s);
    s->prof_on = on;
    s->prof_used = 0;
    return RC_OK;
}

int rc_profile_collect(rc_state* s, double* total_ms, int64_t* launches, double* flop_per_row) {
    RC_ARG(s && total_ms && launches);
    RC_CUDA(cudaDeviceSynchronize());
    double tot = 0;
    for (size_t i = 0; i + 1 < s->prof_used; i += 2) {
        float ms = 0.f;
        RC_CUDA(cudaEventElapsedTime(&ms, s->prof_ev[i], s->prof_ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(s->prof_used / 2);
    if (flop_per_row) {
        if (s->net->gemm_mode >= 2 && s->ph_ready && s->B > 8) {
            // grouped kernel: the launches of one frame together run the whole LSTM stack of one stream once (SURVEY.md 8d:
            // 60 689 920 MAC = 121.4 MFLOP per stream-frame)
            double mac = 0;
            for (int i = 0; i < NNETS; ++i) {
                const NetDev& w = s->net->nets[i];
                mac += (double)w.in * w.H + 16.0 * w.H * w.H + (double)w.H * w.out;
            }
            *flop_per_row = 2.0 * mac;
        } else {                              // one stream-row through one rnn4 LSTM layer: 2 * 4H * 2H
            const double H = (double)s->net->nets[NET4].H;
            *flop_per_row = 2.0 * 4.0 * H * 2.0 * H;
        }
    }
    s->prof_used = 0;
    return RC_OK;
}

// Debug tap of the persistent grouped kernel: enable != 0 allocates per-phase tile traces (16 x int64 per tile, see rc_phase.cuh) that
// every later launch overwrites; phase in [0, 4) with out != null copies that phase's trace (up to max_tiles tiles) and returns in
// *ntiles the upper bound of tiles of the phase.
int rc_state_debug_phase_trace(rc_state* s, int enable, int phase, long long* out, int max_tiles, int* ntiles) {
    RC_ARG(s);
    if (!s->ph_ready) { rc_set_error("grouped kernel not available for this state"); return RC_ERR_STATE; }
    if (enable) {
        for (int ph = 0; ph < PH_COUNT; ++ph) {
            if (s->d_trace[ph]) continue;
            RC_TRY(dev_alloc(s->allocs, &s->d_trace[ph], (size_t)(s->ph_max_tiles[ph] + 1) * 16));
            RC_CUDA(cudaMemset(s->d_trace[ph], 0, (size_t)(s->ph_max_tiles[ph] + 1) * 128));
        }
        if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
    }
    if (out) {
        RC_ARG(phase >= 0 && phase < PH_COUNT && s->d_trace[phase]);
        RC_CUDA(cudaDeviceSynchronize());
        const int n = std::min(max_tiles, s->ph_max_tiles[phase] + 1);
        RC_CUDA(cudaMemcpy(out, s->d_trace[phase], (size_t)n * 128, cudaMemcpyDeviceToHost));
        if (ntiles) *ntiles = s->ph_max_tiles[phase];
    }
    return RC_OK;
}

int rc_state_debug_output(rc_state* s, int which, float* out, void* stream) {
    RC_ARG(s && out);
    const float* src; int ld, w;
    switch (which) {
        case 2: src = s->X3 + 72; ld = RC_K3; w = 69; break;
        case 3: src = s->Y3; ld = 4; w = 3; break;
        case 4: src = s->X6 + 171; ld = RC_K6; w = 69; break;
        case 6: src = s->Y6; ld = 4; w = 3; break;
        case 7: src = s->Y7; ld = 144; w = 144; break;
        case 8: src = s->Y8; ld = 4; w = 2; break;
        case 70: src = s->X7 + 72; ld = RC_K7; w = 69; break;                 // blended joints, multi-launch path
        case 71: src = rc_seq_debug_j3dr(s); ld = 72; w = 69; if (!src) { rc_set_error("sequence kernel has not run"); return RC_ERR_STATE; } break;
        default: rc_set_error("rc_state_debug_output: which=%d", which); return RC_ERR_ARG;
    }
    RC_CUDA(cudaMemcpy2DAsync(out, (size_t)w * 4, src, (size_t)ld * 4, (size_t)w * 4, (size_t)s->B, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return RC_OK;
}

}  // extern "C"
