// Persistent SEQUENCE kernel: frames t0 .. T-1 of Net.forward_offline (net/sig_mp.py:113-274 per frame; the frame loop of
// evaluate.py:75-85) for B independent streams in ONE launch.
//
// Why: with one launch per phase of the frame (phase_tc.cu) a frame costs 13 kernel boundaries plus the ramp / tail of three
// grouped launches — ~90 us of a 260 us frame at 128 streams per GPU (the strong-scaled 8-GPU shard of BASELINE configs[2]), where the
// frame is bound by the DEPTH of its dependency chain, not by throughput.  Here one CTA per SM stays resident for the whole
// sequence and pulls work items from one global, statically ordered queue:
//
//   GEMM tiles    128 rows x BN gate columns (BN = 64 for small shards, 128 otherwise) of linear1 / LSTM-0 / LSTM-1 / linear2 of the
//                 six sub-nets: TMA -> 3-stage smem ring -> tcgen05.mma (split-fp16, 3 MMAs per fp32-accurate product, accumulators
//                 in TMEM) -> 16 epilogue warps (bias, gates, cell update, split-fp16 operand of the NEXT GEMM), as in phase_tc.cu.
//   row jobs      8 streams each, run by the epilogue warps (one warp per stream): PREP (IMU change of frame, key-point
//                 normalisation, branch flags, :138-153), MID (vision / inertial joint blend, :154-167), KIN (6D -> R, IK, foot FK,
//                 translation / contact / floor state machine, SMPL FK of the 33 key points, vision-updater inputs, :173-273),
//                 INIT (rnn2.init_net re-seed, :178-183).
//
// Rows keep their position (stream b = operand row b, 128-row blocks), so nothing is gathered or compacted between frames: every
// producer writes the split-fp16 operand planes of its consumer directly (h_t of an LSTM layer goes both into the x-half of the next
// layer's operand and into the h_prev-half of its own operand for frame t+1, double-buffered by frame parity).  A sub-net pass
// that only applies to some rows (rnn4 / rnn6 on the vision rows, the vision updater on the occluded rows) computes the whole
// 128-row tile and masks the state update by the row's branch flags; jobs no row of a block needs are skipped from flags
// precomputed for all frames (they only depend on the inputs' confidences).
//
// Dependencies: one monotone counter per (job, 128-row block) counts finished tiles over the whole sequence; job X of frame f is
// complete when counter >= (f + 1) * tiles(X).  Every job lists (producer job, frame lag) pairs for its read-after-write AND
// write-after-read hazards; an LSTM tile streams the h_prev half of K first (its producers finished a frame ago) and waits for the
// producer of the x half in the middle of its main loop.  The queue is handed out in an order in which every tile only depends on
// earlier tiles, so the scheme cannot deadlock whatever the residency.
//
// Frame 0 (first_frame / first_tran specials, double rnn6 pass) runs through the multi-launch path; rc_seq_run converts its fp32
// LSTM state into the operand planes and takes over from frame 1.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "rc_common.cuh"
#include "rc_tc.cuh"
#include "rc_tc_dev.cuh"
#include "rc_rows.h"
#include "rc_rows_warp.cuh"
#include "rc_fusion.cuh"
#include "rc_seq.cuh"

namespace {

constexpr int kSqEpiWarps = 16;
constexpr int kSqThreads = 64 + kSqEpiWarps * 32;
constexpr int kSqQ = 4;
constexpr int kSqStages = 3;
constexpr int kSqRowTiles = 8;                  // row-job items per 128-row block
constexpr int kSqRowsPerTile = 128 / kSqRowTiles;   // = one stream per epilogue warp of a row CTA
constexpr int kSqFlDepth = 4;                   // frames of branch flags kept (the vision updater of frame f reads them late)
constexpr int SQ_MAXJOBS = 36;
constexpr int SQ_MAXDEP = 8;
constexpr int SQ_F_INIT = 1 << 16;              // block-summary bit: some row re-seeds rnn2 this frame
constexpr int kSqCtlPad = 16;                   // ctl[0] = queue head, counters start at ctl[16]

enum { SQ_LIN = 0, SQ_LSTM = 1, SQ_PREP = 2, SQ_MID = 3, SQ_KIN = 4, SQ_INIT = 5 };

struct SqDep { short job, lag; };

struct alignas(128) SqJob {
    RcTensorMap mA[2][2];          // [frame parity][hi, lo]
    RcTensorMap mW[2];             // hi, lo
    const float* bias;
    float* C;                      // LSTM cell state [B, H], in place
    float* Y;                      // linear: fp32 output rows (may be null)
    __half* o1h[2]; __half* o1l[2];  // split-fp16 output target 1, indexed by the frame parity of the job
    __half* o2h[2]; __half* o2l[2];  // target 2 (LSTM only: x-half of the next layer / linear2 input)
    int o1pitch, o1col, o2pitch, o2col;
    int ldy, kind, net, nt, K, N, H, relu;
    int skipmask, rowmask;
    int ndep1, ndep2, tile0;
    int xfirst;                    // LSTM: stream the x half of K first and wait for the h_prev producers (dep2) in the middle
    int fshift;                    // the queue hands the job out in frame f for its own frame f + fshift (deferred, non-critical passes: -1)
    SqDep dep1[SQ_MAXDEP], dep2[2];
};

struct SqShared {
    RcNetCfg cfg;
    StepIO io;
    const RcModelConst* model;
    RcRowState* rows;
    const float* gravity_all;
    int* fl[kSqFlDepth]; float* rcr[2]; float* conf[2]; float* lerpw[2];
    float *Y2n, *Y4n, *J3DR, *Y3, *Y6, *Y7, *Y8;
    int* initflag;
    __half *xa2[2][2], *xa3[2][2], *xa4[2][2], *xa6[2][2], *xa7[2][2];   // [parity][hi, lo]
    __half *xa4l[2], *xa6l[2];
    int k2p, k3p, k4p, k6p, k7p;
    const float* Wi[3]; const float* bi[3];
    float *I1, *I2;
    __half *r2h0[2][2], *r2h1[2][2];     // rnn2 h_prev halves: [parity][hi, lo] of the LSTM-0 / LSTM-1 operand
    float *r2c0, *r2c1;
    const int* bf;                       // [T][MB] block summaries of the branch flags
};

struct SqDesc {
    int njobs, MB, B, T, t0, F, tpf, npre;      // tpf: GEMM tiles per frame; npre: PREP items of the first frame (row queue)
    int rpf, row_ctas;                          // row-queue items per frame; CTAs (the last ones of the grid) that serve the row queue
    int row_job[4];                             // PREP, MID, KIN, INIT job indices (row-queue order within a frame)
    SqShared sh;
    SqJob job[SQ_MAXJOBS];
};

struct SqKinScratch { RcKinWarpSmem k; float y7[144]; float x4[RC_K4]; float x6[RC_K6]; };
struct SqPrepScratch { RcPrepWarpSmem p; float x2[RC_K2], x3[RC_K3], x4[RC_K4], x6[RC_K6], x7[RC_K7]; };
union alignas(16) SqRowScratch {
    SqKinScratch kin;
    SqPrepScratch prep;
    float init_x[kInitK0];
};

__device__ __forceinline__ void sq_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ int sq_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int sq_ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sq_red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void sq_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void sq_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kSqEpiWarps * 32) : "memory"); }

// wait until counter >= need (relaxed polls, one acquire at the end); returns the clocks spent waiting
__device__ __forceinline__ long long sq_wait(const int* flag, int need) {
    if (need <= 0) return 0;
    long long waited = 0;
    if (sq_ld_relaxed(flag) < need) {
        const long long t0 = clock64();
        while (sq_ld_relaxed(flag) < need) {
            __nanosleep(20);
            if (clock64() - t0 > 8000000000LL) __trap();          // a broken dependency table must trap, not hang the GPU
        }
        waited = clock64() - t0;
    }
    (void)sq_ld_acquire(flag);
    return waited;
}

__device__ __forceinline__ unsigned long long sq_gtime() {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    return g;
}
// debug timeline (block 0 only): per frame and job {first start, last finish, last x-half dependency met, last accumulator complete,
// last epilogue stores issued, -} in globaltimer ns
__device__ __forceinline__ void sq_trace(unsigned long long* trace, int njobs, int fj, int j, int m, int which) {
    if (!trace || m != 0) return;
    const unsigned long long g = sq_gtime();
    if (which == 0) atomicMin(trace + ((size_t)fj * njobs + j) * 6, g);
    else atomicMax(trace + ((size_t)fj * njobs + j) * 6 + which, g);
}

template <int CW>
__device__ __forceinline__ void sq_tmem_ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void sq_tmem_ld<32>(uint32_t taddr, uint32_t* v) { tc_ld32(taddr, v); }
template <>
__device__ __forceinline__ void sq_tmem_ld<16>(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

__device__ __forceinline__ void sq_split1(float x, __half* hi, __half* lo, size_t idx) {
    const __half h = __float2half_rn(x);
    hi[idx] = h;
    lo[idx] = __float2half_rn((x - __half2float(h)) * 2048.f);
}
// 4 consecutive values -> one 8-byte store per plane
__device__ __forceinline__ void sq_split4(const float* x, __half* hi, __half* lo, size_t idx) {
    __half2 h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const __half h0 = __float2half_rn(x[2 * i]), h1 = __float2half_rn(x[2 * i + 1]);
        h[i] = __halves2half2(h0, h1);
        l[i] = __halves2half2(__float2half_rn((x[2 * i] - __half2float(h0)) * 2048.f), __float2half_rn((x[2 * i + 1] - __half2float(h1)) * 2048.f));
    }
    *reinterpret_cast<uint2*>(hi + idx) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + idx) = *reinterpret_cast<const uint2*>(l);
}
template <int NV>
__device__ __forceinline__ void sq_split_vec(const float* x, __half* hi, __half* lo, size_t idx) {
    if constexpr (NV == 8) tc_store_split<8>(x, hi, lo, idx);
    else sq_split4(x, hi, lo, idx);
}

// one dense layer of init_net for one stream, one warp: y[o] = act(W[o, :] . x + b[o]); x, y in global memory (L2-coherent accesses)
__device__ __forceinline__ void sq_init_layer(const float* __restrict__ W, const float* __restrict__ bias, int K, int N, const float* x,
                                              bool x_shared, bool relu, float* y, int lane) {
    for (int o = 0; o < N; o += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane * 4; k < K; k += 128) {
            float4 xv;
            if (x_shared) xv = *reinterpret_cast<const float4*>(x + k);
            else xv = __ldcg(reinterpret_cast<const float4*>(x + k));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)(o + r) * K + k));
                acc[r] = fmaf(w.x, xv.x, acc[r]); acc[r] = fmaf(w.y, xv.y, acc[r]);
                acc[r] = fmaf(w.z, xv.z, acc[r]); acc[r] = fmaf(w.w, xv.w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
        if (lane == 0) {
            float4 v = make_float4(acc[0] + bias[o], acc[1] + bias[o + 1], acc[2] + bias[o + 2], acc[3] + bias[o + 3]);
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            __stcg(reinterpret_cast<float4*>(y + o), v);
        }
    }
}

// ---- row jobs: one warp per stream -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void sq_row_prep(const SqDesc* D, SqPrepScratch& S, int r, int fj, int lane) {
    const SqShared& sh = D->sh;
    const int p = fj & 1, tt = D->t0 + fj;
    const int b = r;
    bool active = b < D->B;
    if (active && sh.io.lengths) active = tt < sh.io.lengths[b];
    if (!active) { if (lane == 0) sh.fl[fj & (kSqFlDepth - 1)][r] = 0; return; }
    const int f = rc_prep_warp(sh.cfg, 0, S.p, sh.io.j2dc + b * sh.io.sj + (long long)tt * 99, sh.io.accc + b * sh.io.sa + (long long)tt * 18,
                               sh.io.oric + b * sh.io.so + (long long)tt * 54, RC_F_ACTIVE, S.x2, S.x3, S.x4, S.x6, S.x7,
                               sh.rcr[p] + (size_t)r * 9, sh.conf[p] + r, sh.lerpw[p] + (size_t)r * 2, lane);
    if (lane == 0) sh.fl[fj & (kSqFlDepth - 1)][r] = f;
    __syncwarp();
    for (int e = lane; e < RC_K2; e += 32) sq_split1(S.x2[e], sh.xa2[p][0], sh.xa2[p][1], (size_t)r * sh.k2p + e);
    for (int e = lane; e < 72; e += 32) {
        sq_split1(S.x3[e], sh.xa3[p][0], sh.xa3[p][1], (size_t)r * sh.k3p + e);
        sq_split1(S.x7[e], sh.xa7[p][0], sh.xa7[p][1], (size_t)r * sh.k7p + e);
    }
    const bool hi = (f & RC_F_HI) != 0;
    for (int e = lane; e < 171; e += 32) {
        sq_split1((hi || e < 72) ? S.x4[e] : 0.f, sh.xa4[p][0], sh.xa4[p][1], (size_t)r * sh.k4p + e);
        sq_split1(S.x6[e], sh.xa6[p][0], sh.xa6[p][1], (size_t)r * sh.k6p + e);
    }
}

__device__ __forceinline__ void sq_row_mid(const SqDesc* D, int r, int fj, int lane) {
    const SqShared& sh = D->sh;
    const int p = fj & 1;
    const int f = __ldcg(sh.fl[fj & (kSqFlDepth - 1)] + r);
    if (!(f & RC_F_ACTIVE)) return;
    // rnn2's / rnn4's outputs as operand columns of rnn3 (:145) and rnn6 (:156,161,165): coalesced across the warp
    for (int e = lane; e < 69; e += 32) {
        sq_split1(__ldcg(sh.Y2n + (size_t)r * 72 + e), sh.xa3[p][0], sh.xa3[p][1], (size_t)r * sh.k3p + 72 + e);
        if (f & RC_F_HI) sq_split1(__ldcg(sh.Y4n + (size_t)r * 72 + e), sh.xa6[p][0], sh.xa6[p][1], (size_t)r * sh.k6p + 171 + e);
    }
    if (lane >= 23) return;
    float rr[9], lw[2], y2[3], y4[3] = {0.f, 0.f, 0.f}, out[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) rr[q] = __ldcg(sh.rcr[p] + (size_t)r * 9 + q);
    lw[0] = __ldcg(sh.lerpw[p] + (size_t)r * 2); lw[1] = __ldcg(sh.lerpw[p] + (size_t)r * 2 + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) y2[c] = __ldcg(sh.Y2n + (size_t)r * 72 + lane * 3 + c);
    if (f & (RC_F_GE | RC_F_MID)) {
#pragma unroll
        for (int c = 0; c < 3; ++c) y4[c] = __ldcg(sh.Y4n + (size_t)r * 72 + lane * 3 + c);
    }
    rc_mid_joint(f, rr, lw, y2, y4, out);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        __stcg(sh.J3DR + (size_t)r * 72 + lane * 3 + c, out[c]);
        sq_split1(out[c], sh.xa7[p][0], sh.xa7[p][1], (size_t)r * sh.k7p + 72 + lane * 3 + c);
    }
}

__device__ __forceinline__ void sq_row_kin(const SqDesc* D, const RcModelConst& M, SqKinScratch& S, int r, int fj, int lane) {
    const SqShared& sh = D->sh;
    const int p = fj & 1, tt = D->t0 + fj;
    const int b = r;
    int f = 0;
    if (b < D->B) f = __ldcg(sh.fl[fj & (kSqFlDepth - 1)] + r);
    if (!(f & RC_F_ACTIVE)) { if (lane == 0) sh.initflag[r] = 0; return; }
    for (int e = lane; e < 144; e += 32) S.y7[e] = __ldcg(sh.Y7 + (size_t)b * 144 + e);
    float rr[9], g[3], ft[3] = {0.f, 0.f, 0.f}, y8[2], vr[3], pc[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) rr[q] = __ldcg(sh.rcr[p] + (size_t)r * 9 + q);
    const float* gp = sh.io.gravity ? (sh.io.gravity + (size_t)b * 3) : sh.gravity_all;
#pragma unroll
    for (int q = 0; q < 3; ++q) { g[q] = gp[q]; vr[q] = __ldcg(sh.Y3 + (size_t)b * 4 + q); pc[q] = __ldcg(sh.Y6 + (size_t)b * 4 + q); }
    y8[0] = __ldcg(sh.Y8 + (size_t)b * 4); y8[1] = __ldcg(sh.Y8 + (size_t)b * 4 + 1);
    const float conf = __ldcg(sh.conf[p] + r);
    __syncwarp();
    const int need = rc_kin_warp(sh.cfg, M, S.k, sh.rows + b, f, S.y7, y8, vr, pc, rr, conf, g, ft,
                                 sh.io.pose + b * sh.io.sp + (long long)tt * 216, sh.io.tran + b * sh.io.st + (long long)tt * 3, S.x4, S.x6, lane,
                                 sh.io.branch ? sh.io.branch + b * sh.io.sb + tt : nullptr);
    if (lane == 0) sh.initflag[r] = need;
    if (f & RC_F_LATE) {                                      // vision-updater inputs (:263-271) straight into the operand planes
        __syncwarp();
        const float* pa = sh.io.accc + b * sh.io.sa + (long long)tt * 18;
        const float* po = sh.io.oric + b * sh.io.so + (long long)tt * 54;
        for (int e = lane; e < 72; e += 32) {
            const float v = (e < 18) ? pa[e] : po[e - 18];
            sq_split1(v, sh.xa4l[0], sh.xa4l[1], (size_t)r * sh.k4p + e);
            sq_split1(v, sh.xa6l[0], sh.xa6l[1], (size_t)r * sh.k6p + e);
        }
        for (int e = 72 + lane; e < 171; e += 32) sq_split1(S.x4[e], sh.xa4l[0], sh.xa4l[1], (size_t)r * sh.k4p + e);
        for (int e = 72 + lane; e < 240; e += 32) sq_split1(S.x6[e], sh.xa6l[0], sh.xa6l[1], (size_t)r * sh.k6p + e);
    }
}

__device__ __forceinline__ void sq_row_init(const SqDesc* D, float* xs, int r, int fj, int lane) {
    const SqShared& sh = D->sh;
    if (r >= D->B || !__ldcg(sh.initflag + r)) return;
    const int b = r, pn = (fj + 1) & 1;
    for (int e = lane; e < kInitK0; e += 32) xs[e] = (e < 69) ? __ldcg(sh.J3DR + (size_t)r * 72 + e) : 0.f;
    __syncwarp();
    float* y1 = sh.I1 + (size_t)b * 512;
    float* y2 = sh.I2 + (size_t)b * 1024;
    sq_init_layer(sh.Wi[0], sh.bi[0], kInitK0, 512, xs, true, true, y1, lane);
    __syncwarp();
    sq_init_layer(sh.Wi[1], sh.bi[1], 512, 1024, y1, false, true, y2, lane);
    __syncwarp();
    // last layer: [h0 | h1 | c0 | c1] (sig_mp.py:182) -> rnn2's state of the NEXT frame
    const float* W = sh.Wi[2];
    const float* bias = sh.bi[2];
    for (int o = 0; o < 2048; o += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane * 4; k < 1024; k += 128) {
            const float4 xv = __ldcg(reinterpret_cast<const float4*>(y2 + k));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)(o + q) * 1024 + k));
                acc[q] = fmaf(w.x, xv.x, acc[q]); acc[q] = fmaf(w.y, xv.y, acc[q]);
                acc[q] = fmaf(w.z, xv.z, acc[q]); acc[q] = fmaf(w.w, xv.w, acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], s);
        if (lane == 0) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = acc[q] + bias[o + q];
            const int u = o & 511;
            if (o < 512) sq_split4(v, sh.r2h0[pn][0], sh.r2h0[pn][1], (size_t)r * 1024 + 512 + u);
            else if (o < 1024) sq_split4(v, sh.r2h1[pn][0], sh.r2h1[pn][1], (size_t)r * 1024 + 512 + u);
            else if (o < 1536) __stcg(reinterpret_cast<float4*>(sh.r2c0 + (size_t)b * 512 + u), make_float4(v[0], v[1], v[2], v[3]));
            else __stcg(reinterpret_cast<float4*>(sh.r2c1 + (size_t)b * 512 + u), make_float4(v[0], v[1], v[2], v[3]));
        }
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------------------------
template <int BN>
struct SqCfg {
    static constexpr int kABytes = kTcBM * kTcBK * 2;
    static constexpr int kWBytes = BN * kTcBK * 2;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;
    static constexpr int kTmemCols = 4 * BN;
    static constexpr int CW = BN / 4;                                   // accumulator columns per epilogue warp
    static constexpr int kRowBytes = kSqEpiWarps * (int)sizeof(SqRowScratch) + (int)sizeof(RcModelConst) + 64;   // row CTAs: scratch + SMPL constants
    static constexpr int kSmem = (kSqStages * kStageBytes > kRowBytes ? kSqStages * kStageBytes : kRowBytes) + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kSqThreads, 1)
rc_seq_kernel(const SqDesc* __restrict__ D, int* __restrict__ ctl, long long* __restrict__ stats, unsigned long long* __restrict__ trace) {
    using Cfg = SqCfg<BN>;
    constexpr int CW = Cfg::CW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kSqStages];
    __shared__ __align__(8) uint64_t bar_empty[kSqStages];
    __shared__ __align__(8) uint64_t bar_acc_full;
    __shared__ __align__(8) uint64_t bar_acc_free;
    __shared__ __align__(8) uint64_t tq_full[kSqQ];
    __shared__ __align__(8) uint64_t tq_empty[kSqQ];
    __shared__ int4 tq_tile[kSqQ];
    __shared__ int tile_start[SQ_MAXJOBS + 1];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int njobs = D->njobs, MB = D->MB;
    int* const cnt = ctl + kSqCtlPad;

    if ((int)blockIdx.x >= (int)gridDim.x - D->row_ctas) {
        // ---- row CTA: serves the row queue (PREP / MID / KIN / INIT items of 16 streams, one warp per stream) -------------------
        SqRowScratch* rscr = reinterpret_cast<SqRowScratch*>(smem);
        RcModelConst* Ms = reinterpret_cast<RcModelConst*>(smem + kSqEpiWarps * sizeof(SqRowScratch));
        {
            const int* src = reinterpret_cast<const int*>(D->sh.model);
            int* dst = reinterpret_cast<int*>(Ms);
            for (int e = threadIdx.x; e < (int)(sizeof(RcModelConst) / 4); e += blockDim.x) dst[e] = src[e];
        }
        __shared__ int4 item;
        const int rtotal = D->npre + D->F * D->rpf;
        const int per_job = kSqRowTiles * MB;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) {
                int4 it4 = make_int4(-1, 0, 0, 0);
                for (;;) {
                    const int g = atomicAdd(&ctl[1], 1);
                    if (g >= rtotal) break;
                    int j, fj, local;
                    if (g < D->npre) { j = D->row_job[0]; fj = 0; local = g; }
                    else {
                        const int gg = g - D->npre;
                        const int f = gg / D->rpf;
                        const int l2 = gg - f * D->rpf;
                        const int k = l2 / per_job;
                        local = l2 - k * per_job;
                        j = D->row_job[k];
                        fj = (k == 0) ? f + 1 : f;                  // the PREP items of frame f prepare frame f + 1
                    }
                    if (fj >= D->F) continue;
                    const SqJob& J = D->job[j];
                    const int m = local / J.nt, n = local - m * J.nt;
                    int* const mine = cnt + j * MB + m;
                    if ((__ldg(D->sh.bf + (size_t)(D->t0 + fj) * MB + m) & J.skipmask) == 0) {
                        sq_wait(mine, fj * J.nt);
                        sq_red_release_add(mine, 1);
                        if (stats) atomicAdd((unsigned long long*)stats + j * 4 + 3, 1ULL);
                        continue;
                    }
                    long long w1 = sq_wait(mine, fj * J.nt);
                    for (int d = 0; d < J.ndep1; ++d) {
                        const SqDep dp = J.dep1[d];
                        w1 += sq_wait(cnt + dp.job * MB + m, (fj - dp.lag + 1) * D->job[dp.job].nt);
                    }
                    if (stats) { atomicAdd((unsigned long long*)stats + j * 4 + 0, 1ULL); atomicAdd((unsigned long long*)stats + j * 4 + 1, (unsigned long long)w1); }
                    sq_trace(trace, njobs, fj, j, m, 0);
                    it4 = make_int4(j, m, n, fj);
                    break;
                }
                item = it4;
            }
            __syncthreads();
            const int4 t = item;
            if (t.x < 0) break;
            const int kind = D->job[t.x].kind;
            if (warp < kSqEpiWarps) {
                const int r = t.y * 128 + t.z * kSqRowsPerTile + warp;
                if (kind == SQ_PREP) sq_row_prep(D, rscr[warp].prep, r, t.w, lane);
                else if (kind == SQ_MID) sq_row_mid(D, r, t.w, lane);
                else if (kind == SQ_KIN) sq_row_kin(D, *Ms, rscr[warp].kin, r, t.w, lane);
                else sq_row_init(D, rscr[warp].init_x, r, t.w, lane);
                if (warp == 0 && lane == 0) sq_trace(trace, njobs, t.w, t.x, t.y, 3);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                sq_fence_proxy_async();
                __threadfence();
                sq_red_release_add(cnt + t.x * MB + t.y, 1);
                sq_trace(trace, njobs, t.w, t.x, t.y, 1);
            }
        }
        return;
    }

    for (int j = threadIdx.x; j <= njobs; j += blockDim.x) tile_start[j] = (j < njobs) ? D->job[j].tile0 : D->tpf;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSqStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc_full), 1);
        mbar_init(smem_u32(&bar_acc_free), kSqEpiWarps);
        for (int q = 0; q < kSqQ; ++q) { mbar_init(smem_u32(&tq_full[q]), 1); mbar_init(smem_u32(&tq_empty[q]), 1 + kSqEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int total = (D->F + 1) * D->tpf;             // one extra queue frame for the deferred passes of the last frame

    if (warp == 0) {
        if (lane == 0) {
            // ---- scheduler + TMA producer -------------------------------------------------------------------------------
            uint32_t it = 0;
            int q = 0;
            for (;;) {
                const int g = atomicAdd(&ctl[0], 1);
                int j = -1, fj = 0, local = 0;
                if (g < total) {
                    fj = g / D->tpf;
                    local = g - fj * D->tpf;
                    j = 0;
                    while (local >= tile_start[j + 1]) ++j;          // row jobs have no tiles in this queue (tile_start[j + 1] == tile_start[j])
                    local -= tile_start[j];
                }
                if (j < 0) {                                       // end of the queue: tell the consumers
                    const int slot = q % kSqQ;
                    mbar_wait(smem_u32(&tq_empty[slot]), (((uint32_t)q / kSqQ) & 1u) ^ 1u);
                    tq_tile[slot] = make_int4(-1, 0, 0, 0);
                    sq_mbar_arrive(smem_u32(&tq_full[slot]));
                    break;
                }
                const SqJob& J = D->job[j];
                fj += J.fshift;
                if (fj < 0 || fj >= D->F) continue;
                const int m = local / J.nt, n = local - m * J.nt;
                int* const mine = cnt + j * MB + m;
                const int tt = D->t0 + fj;
                if ((__ldg(D->sh.bf + (size_t)tt * MB + m) & J.skipmask) == 0) {
                    // no row of this block needs the job this frame: count it as done once its previous execution is complete
                    sq_wait(mine, fj * J.nt);
                    sq_red_release_add(mine, 1);
                    if (stats) atomicAdd((unsigned long long*)stats + j * 4 + 3, 1ULL);
                    continue;
                }
                long long w1 = 0, w2 = 0;
                w1 += sq_wait(mine, fj * J.nt);                     // the state / hazard dependencies first
                for (int d = 0; d < J.ndep1; ++d) {
                    const SqDep dp = J.dep1[d];
                    w1 += sq_wait(cnt + dp.job * MB + m, (fj - dp.lag + 1) * D->job[dp.job].nt);
                }
                sq_fence_proxy_async();
                sq_trace(trace, njobs, fj, j, m, 0);
                const int slot = q % kSqQ;
                mbar_wait(smem_u32(&tq_empty[slot]), (((uint32_t)q / kSqQ) & 1u) ^ 1u);
                tq_tile[slot] = make_int4(j, m, n, fj);
                sq_mbar_arrive(smem_u32(&tq_full[slot]));
                ++q;
                const int KB = J.K / kTcBK;
                const int KH = (J.kind == SQ_LSTM) ? KB / 2 : 0;   // K blocks of the h_prev half, streamed first
                const CUtensorMap* mAh = (const CUtensorMap*)&J.mA[fj & 1][0];
                const CUtensorMap* mAl = (const CUtensorMap*)&J.mA[fj & 1][1];
                const CUtensorMap* mWh = (const CUtensorMap*)&J.mW[0];
                const CUtensorMap* mWl = (const CUtensorMap*)&J.mW[1];
                for (int i = 0; i < KB; ++i, ++it) {
                    if (i == KH && J.ndep2) {
                        for (int d = 0; d < J.ndep2; ++d) {
                            const SqDep dp = J.dep2[d];
                            w2 += sq_wait(cnt + dp.job * MB + m, (fj - dp.lag + 1) * D->job[dp.job].nt);
                        }
                        sq_fence_proxy_async();
                        sq_trace(trace, njobs, fj, j, m, 2);
                    }
                    const int kb = J.xfirst ? i : ((i < KH) ? (KB - KH + i) : (i - KH));
                    const int s = it % kSqStages;
                    const uint32_t ph = (it / kSqStages) & 1u;
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_expect_tx(full, Cfg::kStageBytes);
                    const uint32_t base = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                    tma_load_2d(base, mAh, kb * kTcBK, m * kTcBM, full);
                    tma_load_2d(base + Cfg::kABytes, mAl, kb * kTcBK, m * kTcBM, full);
                    tma_load_2d(base + 2 * Cfg::kABytes, mWh, kb * kTcBK, n * BN, full);
                    tma_load_2d(base + 2 * Cfg::kABytes + Cfg::kWBytes, mWl, kb * kTcBK, n * BN, full);
                }
                if (stats) {
                    atomicAdd((unsigned long long*)stats + j * 4 + 0, 1ULL);
                    atomicAdd((unsigned long long*)stats + j * 4 + 1, (unsigned long long)w1);
                    atomicAdd((unsigned long long*)stats + j * 4 + 2, (unsigned long long)w2);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer ------------------------------------------------------------------------------------------------
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
            uint32_t it = 0;
            int gq = 0;
            for (int q = 0;; ++q) {
                const int slot = q % kSqQ;
                mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kSqQ) & 1u);
                const int4 t = tq_tile[slot];
                sq_mbar_arrive(smem_u32(&tq_empty[slot]));
                if (t.x < 0) break;
                const int KB = D->job[t.x].K / kTcBK;
                if (gq > 0) {
                    mbar_wait(smem_u32(&bar_acc_free), (uint32_t)(gq - 1) & 1u);
                    tc_fence_after();
                }
                const uint32_t d_corr = tmem_base + (uint32_t)(((3 * gq) & 3) * BN);
                const uint32_t d_m0 = tmem_base + (uint32_t)(((3 * gq + 1) & 3) * BN);
                const uint32_t d_m1 = tmem_base + (uint32_t)(((3 * gq + 2) & 3) * BN);
                int g = 0;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kSqStages;
                    const uint32_t ph = (it / kSqStages) & 1u;
                    mbar_wait(smem_u32(&bar_full[s]), ph);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                    const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + Cfg::kABytes);
                    const uint64_t dWhi = make_desc(base + 2 * Cfg::kABytes), dWlo = make_desc(base + 2 * Cfg::kABytes + Cfg::kWBytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; ++k, ++g) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        tc_mma_f16((g & 1) ? d_m1 : d_m0, dAhi + adv, dWhi + adv, idesc, g >= 2 ? 1u : 0u);
                        tc_mma_f16(d_corr, dAhi + adv, dWlo + adv, idesc, g ? 1u : 0u);
                        tc_mma_f16(d_corr, dAlo + adv, dWhi + adv, idesc, 1u);
                    }
                    tc_commit(smem_u32(&bar_empty[s]));
                }
                tc_commit(smem_u32(&bar_acc_full));
                ++gq;
            }
        }
    } else {
        // ---- epilogue warps --------------------------------------------------------------------------------------------------------
        const int ewarp = warp - 2;
        const int q4 = warp & 3, part = ewarp >> 2;             // a warp reads TMEM lanes 32 * (warp_id % 4) ..; part = its share of the columns
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const SqShared& sh = D->sh;
        int gq = 0;
        for (int q = 0;; ++q) {
            const int slot = q % kSqQ;
            mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kSqQ) & 1u);
            const int4 t = tq_tile[slot];
            __syncwarp();
            if (lane == 0) sq_mbar_arrive(smem_u32(&tq_empty[slot]));
            if (t.x < 0) break;
            const SqJob& J = D->job[t.x];
            const int m = t.y, fj = t.w;
            int* const mine = cnt + t.x * MB + m;
            const bool lstm = J.kind == SQ_LSTM;
            const int n0 = t.z * BN;
            const int r = m * 128 + q4 * 32 + lane;
            const int H = J.H;
            const int c0 = n0 + part * CW;                       // first accumulator column of this thread
            float cprev[CW / 4];
            const bool rowok = r < D->B;
            if (lstm && rowok && !J.xfirst) {                    // cell state of this thread's units, ahead of the accumulators
                const float* cp = J.C + (size_t)r * H + (c0 >> 2);
#pragma unroll
                for (int u = 0; u < CW / 4; u += 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(cp + u));
                    cprev[u] = v.x; cprev[u + 1] = v.y; cprev[u + 2] = v.z; cprev[u + 3] = v.w;
                }
            }
            const uint32_t b_corr = (uint32_t)(((3 * gq) & 3) * BN);
            const uint32_t b_m0 = (uint32_t)(((3 * gq + 1) & 3) * BN);
            const uint32_t b_m1 = (uint32_t)(((3 * gq + 2) & 3) * BN);
            mbar_wait(smem_u32(&bar_acc_full), (uint32_t)gq & 1u);
            tc_fence_after();
            ++gq;
            if (ewarp == 0 && lane == 0) sq_trace(trace, njobs, fj, t.x, m, 3);
            if (lstm && rowok && J.xfirst) {                     // x-first jobs: the state producers are only known to be done once the MMAs are
                const float* cp = J.C + (size_t)r * H + (c0 >> 2);
#pragma unroll
                for (int u = 0; u < CW / 4; u += 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(cp + u));
                    cprev[u] = v.x; cprev[u + 1] = v.y; cprev[u + 2] = v.z; cprev[u + 3] = v.w;
                }
            }
            float acc[CW];
            {
                uint32_t v0[CW], v1[CW];
                sq_tmem_ld<CW>(lane_base + b_m0 + (uint32_t)(part * CW), v0);
                sq_tmem_ld<CW>(lane_base + b_corr + (uint32_t)(part * CW), v1);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < CW; ++e) acc[e] = fmaf(__uint_as_float(v1[e]), 4.8828125e-4f, __uint_as_float(v0[e]));   // main-0 + corr * 2^-11
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) sq_mbar_arrive(smem_u32(&bar_acc_free));
            {
                uint32_t v0[CW];
                sq_tmem_ld<CW>(lane_base + b_m1 + (uint32_t)(part * CW), v0);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < CW; ++e) acc[e] += __uint_as_float(v0[e]);
            }
            const int rf = rowok ? __ldcg(sh.fl[fj & (kSqFlDepth - 1)] + r) : 0;
            const bool inpass = (rf & J.rowmask) != 0;
            const int p = fj & 1;
            if (lstm) {
                if (inpass) {
                    float cn[CW / 4], hn[CW / 4];
#pragma unroll
                    for (int u = 0; u < CW / 4; ++u) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(J.bias + c0 + u * 4));
                        cn[u] = fmaf(sigm(acc[u * 4 + 1] + b.y), cprev[u], sigm(acc[u * 4 + 0] + b.x) * tanhf(acc[u * 4 + 2] + b.z));
                        hn[u] = sigm(acc[u * 4 + 3] + b.w) * tanhf(cn[u]);
                    }
                    float* cw = J.C + (size_t)r * H + (c0 >> 2);
#pragma unroll
                    for (int u = 0; u < CW / 4; u += 4) __stcg(reinterpret_cast<float4*>(cw + u), make_float4(cn[u], cn[u + 1], cn[u + 2], cn[u + 3]));
                    sq_split_vec<CW / 4>(hn, J.o1h[p], J.o1l[p], (size_t)r * J.o1pitch + J.o1col + (c0 >> 2));
                    if (J.o2h[p]) sq_split_vec<CW / 4>(hn, J.o2h[p], J.o2l[p], (size_t)r * J.o2pitch + J.o2col + (c0 >> 2));
                }
            } else if (J.relu) {                                 // linear1: relu, only the split operand of LSTM-0 is wanted
#pragma unroll
                for (int e = 0; e < CW; e += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(J.bias + c0 + e));
                    acc[e] = fmaxf(acc[e] + b.x, 0.f); acc[e + 1] = fmaxf(acc[e + 1] + b.y, 0.f);
                    acc[e + 2] = fmaxf(acc[e + 2] + b.z, 0.f); acc[e + 3] = fmaxf(acc[e + 3] + b.w, 0.f);
                }
#pragma unroll
                for (int e = 0; e < CW; e += CW / 4)
                    sq_split_vec<CW / 4>(acc + e, J.o1h[p], J.o1l[p], (size_t)r * J.o1pitch + J.o1col + c0 + e);
            } else if (inpass) {                                 // linear2: few valid columns, fp32 rows for the row jobs (16-byte stores)
#pragma unroll
                for (int e = 0; e < CW; e += 4) {
                    if (c0 + e < J.ldy) {                           // ldy = valid columns rounded up to 4; bias and weight rows are zero-padded
                        const float4 b = __ldg(reinterpret_cast<const float4*>(J.bias + c0 + e));
                        __stcg(reinterpret_cast<float4*>(J.Y + (size_t)r * J.ldy + c0 + e),
                               make_float4(acc[e] + b.x, acc[e + 1] + b.y, acc[e + 2] + b.z, acc[e + 3] + b.w));
                    }
                }
            }
            // publish: every thread's stores -> barrier of the epilogue warps -> ONE thread's proxy + gpu-scope fences -> release increment
            if (ewarp == 0 && lane == 0) sq_trace(trace, njobs, fj, t.x, m, 4);
            sq_epi_bar();
            if (ewarp == 0 && lane == 0) {
                sq_fence_proxy_async();
                sq_red_release_add(mine, 1);                     // release at gpu scope, cumulative over the barrier
                sq_trace(trace, njobs, fj, t.x, m, 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
    }
}

// ---- pre-pass: branch flags of all frames (they only depend on the inputs' confidences), block summaries, init frames ---------------
__global__ void __launch_bounds__(256) sq_flags_kernel(RcNetCfg cfg, StepIO io, int B, int Bp, int T, int t0, int* __restrict__ pf) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)(T - t0) * Bp) return;
    const int t = t0 + (int)(e / Bp), b = (int)(e % Bp);
    int f = 0;
    if (b < B && (!io.lengths || t < io.lengths[b])) {
        float lw[2];
        const float cf = rc_conf_mean(io.j2dc + b * io.sj + (long long)t * 99);
        f = rc_prep_flags(cfg, 0, cf, RC_F_ACTIVE, lw);
    }
    pf[(size_t)t * Bp + b] = f;
}
// first frame with c >= hi re-seeds rnn2 (:178-183); whether it already happened in frames < t0 is in the row state
__global__ void __launch_bounds__(128) sq_init_scan_kernel(const RcRowState* __restrict__ rows, int B, int Bp, int T, int t0, int* __restrict__ pf) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (!rows[b].first_reach) return;
    for (int t = t0; t < T; ++t) {
        const int f = pf[(size_t)t * Bp + b];
        if ((f & RC_F_ACTIVE) && (f & RC_F_GE)) { pf[(size_t)t * Bp + b] = f | SQ_F_INIT; return; }
    }
}
__global__ void __launch_bounds__(128) sq_blk_kernel(const int* __restrict__ pf, int Bp, int MB, int T, int t0, int* __restrict__ bf) {
    const int t = t0 + blockIdx.x / MB, m = blockIdx.x % MB;
    int f = pf[(size_t)t * Bp + m * 128 + threadIdx.x];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, s);
    __shared__ int w[4];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = f;
    __syncthreads();
    if (threadIdx.x == 0) bf[(size_t)t * MB + m] = w[0] | w[1] | w[2] | w[3];
}

// fp32 LSTM state of the multi-launch path -> h_prev halves of the operand planes (parity 0 = first frame of the sequence kernel)
struct SqConvArgs {
    const float* src[2 * NNETS]; __half* hi[2 * NNETS]; __half* lo[2 * NNETS]; int H[2 * NNETS];
    int B;
};
__global__ void __launch_bounds__(256) sq_convert_kernel(const __grid_constant__ SqConvArgs a) {
    const int seg = blockIdx.y;
    const int H = a.H[seg];
    const long long total = (long long)a.B * (H / 4);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / (H / 4)), k = (int)(e % (H / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4*>(a.src[seg] + (size_t)r * H + k);
        const float x[4] = {v.x, v.y, v.z, v.w};
        sq_split4(x, a.hi[seg], a.lo[seg], (size_t)r * 2 * H + H + k);
    }
}

int sq_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------------------------------------
struct RcSeq {
    int MB = 0, Bp = 0;
    std::vector<void*> allocs;
    __half *A0[NNETS][2][2] = {}, *A0L[NNETS][2] = {}, *A1[NNETS][2][2] = {}, *A2[NNETS][2][2] = {}, *A3[NNETS][2][2] = {};
    int* fl[kSqFlDepth] = {}; float* rcr[2] = {}; float* conf[2] = {}; float* lerpw[2] = {};
    float *Y2n = nullptr, *Y4n = nullptr, *J3DR = nullptr;
    int* initflag = nullptr;
    int *pf = nullptr, *bf = nullptr;
    int cap_T = 0;
    SqDesc* h_desc[2] = {nullptr, nullptr};        // host copies, [0] BN = 64, [1] BN = 128
    SqDesc* d_desc = nullptr;
    int* ctl = nullptr;
    long long* stats = nullptr;
    unsigned long long* trace = nullptr;
    int trace_F = 0;
    int njobs = 0;
    bool stats_on = false;
};

namespace {

template <class T>
int sq_alloc(RcSeq* q, T** p, size_t n, bool zero = true) {
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) { rc_set_error("rc_seq: cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); return RC_ERR_ALLOC; }
    q->allocs.push_back(v);
    if (zero && cudaMemset(v, 0, (n ? n : 1) * sizeof(T)) != cudaSuccess) { rc_set_error("rc_seq: cudaMemset failed"); return RC_ERR_CUDA; }
    *p = (T*)v;
    return RC_OK;
}
#define SQ_TRY(x) do { int rc_ = (x); if (rc_ != RC_OK) return rc_; } while (0)

int sq_alloc_buffers(rc_state* s) {
    RcSeq* q = s->seq;
    const rc_net* net = s->net;
    const int MB = (s->B + 127) / 128;
    q->MB = MB; q->Bp = MB * 128;
    const size_t Bp = (size_t)q->Bp;
    for (int i = 0; i < NNETS; ++i) {
        const NetDev& w = net->nets[i];
        for (int p = 0; p < 2; ++p)
            for (int h = 0; h < 2; ++h) {
                if (i == NET8) q->A0[i][p][h] = q->A0[NET7][p][h];            // rnn7 and rnn8 share their input (:169-170)
                else SQ_TRY(sq_alloc(q, &q->A0[i][p][h], Bp * w.K1p));
                SQ_TRY(sq_alloc(q, &q->A1[i][p][h], Bp * 2 * w.H));
                SQ_TRY(sq_alloc(q, &q->A2[i][p][h], Bp * 2 * w.H));
                SQ_TRY(sq_alloc(q, &q->A3[i][p][h], Bp * w.H));
            }
        if (i == NET4 || i == NET6)
            for (int h = 0; h < 2; ++h) SQ_TRY(sq_alloc(q, &q->A0L[i][h], Bp * w.K1p));
    }
    for (int p = 0; p < kSqFlDepth; ++p) SQ_TRY(sq_alloc(q, &q->fl[p], Bp));
    for (int p = 0; p < 2; ++p) {
        SQ_TRY(sq_alloc(q, &q->rcr[p], Bp * 9));
        SQ_TRY(sq_alloc(q, &q->conf[p], Bp));
        SQ_TRY(sq_alloc(q, &q->lerpw[p], Bp * 2));
    }
    SQ_TRY(sq_alloc(q, &q->Y2n, Bp * 72));
    SQ_TRY(sq_alloc(q, &q->Y4n, Bp * 72));
    SQ_TRY(sq_alloc(q, &q->J3DR, Bp * 72));
    SQ_TRY(sq_alloc(q, &q->initflag, Bp));
    SQ_TRY(sq_alloc(q, &q->d_desc, 1));
    SQ_TRY(sq_alloc(q, &q->ctl, (size_t)kSqCtlPad + (size_t)SQ_MAXJOBS * MB));
    SQ_TRY(sq_alloc(q, &q->stats, (size_t)SQ_MAXJOBS * 4));
    return RC_OK;
}

// The static job table of one frame, in queue order.  See the header comment for the dependency rules; `lag` counts frames back.
int sq_build_jobs(rc_state* s, int bn, SqDesc* d) {
    RcSeq* q = s->seq;
    const rc_net* net = s->net;
    const int MB = q->MB;
    const long long Bp = q->Bp;
    memset(d, 0, sizeof(*d));
    int nj = 0;
    int jPREP, jMID, jKIN, jINIT, jL1[NNETS][2], jA[NNETS][2], jB[NNETS][2], jO[NNETS];
    for (int i = 0; i < NNETS; ++i) { jO[i] = -1; for (int l = 0; l < 2; ++l) jL1[i][l] = jA[i][l] = jB[i][l] = -1; }
    jPREP = nj++;
    const int p1[2] = {NET4, NET2}, p2[4] = {NET6, NET3, NET7, NET8}, pl[2] = {NET4, NET6};
    // Queue order of a frame f.  The vision updater of frame f - 1 only has its rnn4 linear1 / LSTM-0 on the critical cycle (rnn4's
    // LSTM-0 of frame f needs that state); its other passes are handed out later, between the passes of frame f that need them.
    for (int c = 0; c < 2; ++c) jL1[p1[c]][0] = nj++;
    for (int c = 0; c < 2; ++c) jA[p1[c]][0] = nj++;
    jA[NET6][1] = nj++;                                    // frame f - 1
    for (int c = 0; c < 2; ++c) jB[p1[c]][0] = nj++;
    jB[NET6][1] = nj++;                                    // frame f - 1
    for (int c = 0; c < 2; ++c) jO[p1[c]] = nj++;
    jMID = nj++;
    for (int c = 0; c < 4; ++c) jL1[p2[c]][0] = nj++;
    for (int c = 0; c < 4; ++c) jA[p2[c]][0] = nj++;
    for (int c = 0; c < 4; ++c) jB[p2[c]][0] = nj++;
    for (int c = 0; c < 4; ++c) jO[p2[c]] = nj++;
    jKIN = nj++;
    jINIT = nj++;
    for (int c = 0; c < 2; ++c) jL1[pl[c]][1] = nj++;
    jA[NET4][1] = nj++;
    jB[NET4][1] = nj++;                                    // rnn4's LSTM-1 of frame f + 1 waits for it: right behind its LSTM-0
    if (nj > SQ_MAXJOBS) { rc_set_error("rc_seq: job table overflow"); return RC_ERR_STATE; }
    d->njobs = nj;
    q->njobs = nj;

    int rc = RC_OK;
    auto mkA = [&](RcTensorMap* m, const void* base, int K) { if (rc == RC_OK) rc = rc_tc_make_map(m, base, Bp, K, 128); };
    auto dep1 = [&](int j, int pj, int lag) {
        if (pj < 0) return;
        SqJob& J = d->job[j];
        if (J.ndep1 >= SQ_MAXDEP) { rc_set_error("rc_seq: too many dependencies for job %d", j); rc = RC_ERR_STATE; return; }
        J.dep1[J.ndep1].job = (short)pj; J.dep1[J.ndep1].lag = (short)lag; J.ndep1++;
    };
    auto dep2 = [&](int j, int pj, int lag) {
        SqJob& J = d->job[j];
        J.dep2[J.ndep2].job = (short)pj; J.dep2[J.ndep2].lag = (short)lag; J.ndep2++;
    };
    auto skipmask_of = [&](int ni, int late) { return late ? RC_F_LATE : (ni == NET4 ? RC_F_HI : (ni == NET6 ? RC_F_R6B : RC_F_ACTIVE)); };

    // row jobs
    const int rowjobs[4] = {jPREP, jMID, jKIN, jINIT};
    const int rowkinds[4] = {SQ_PREP, SQ_MID, SQ_KIN, SQ_INIT};
    for (int c = 0; c < 4; ++c) {
        SqJob& J = d->job[rowjobs[c]];
        J.kind = rowkinds[c]; J.nt = kSqRowTiles; J.net = -1;
        J.skipmask = (rowkinds[c] == SQ_INIT) ? SQ_F_INIT : RC_F_ACTIVE;
        J.rowmask = RC_F_ACTIVE;
    }
    // GEMM jobs
    for (int i = 0; i < NNETS; ++i) {
        const NetDev& w = net->nets[i];
        const NetBuf& nb = s->nb[i];
        const int H = w.H;
        for (int late = 0; late < 2; ++late) {
            if (jL1[i][late] < 0) continue;
            const int mask = skipmask_of(i, late);
            {   // linear1 + relu -> x-half of the LSTM-0 operand
                SqJob& J = d->job[jL1[i][late]];
                J.kind = SQ_LIN; J.net = i; J.K = w.K1p; J.N = H; J.H = H; J.relu = 1; J.nt = H / bn;
                J.skipmask = mask; J.rowmask = mask; J.bias = w.b1;
                for (int p = 0; p < 2; ++p) {
                    const __half* base_h = late ? q->A0L[i][0] : q->A0[i][p][0];
                    const __half* base_l = late ? q->A0L[i][1] : q->A0[i][p][1];
                    mkA(&J.mA[p][0], base_h, w.K1p); mkA(&J.mA[p][1], base_l, w.K1p);
                    J.o1h[p] = q->A1[i][p][0]; J.o1l[p] = q->A1[i][p][1];
                }
                J.o1pitch = 2 * H; J.o1col = 0;
                J.mW[0] = (bn == 64) ? w.mW1hi64 : w.mW1hi; J.mW[1] = (bn == 64) ? w.mW1lo64 : w.mW1lo;
            }
            for (int l = 0; l < 2; ++l) {
                SqJob& J = d->job[l == 0 ? jA[i][late] : jB[i][late]];
                J.kind = SQ_LSTM; J.net = i; J.K = 2 * H; J.N = 4 * H; J.H = H; J.nt = 4 * H / bn;
                J.skipmask = mask; J.rowmask = mask; J.bias = w.bL[l]; J.C = nb.c[l];
                __half* (*A)[2][2] = (l == 0) ? q->A1 : q->A2;
                for (int p = 0; p < 2; ++p) {
                    mkA(&J.mA[p][0], A[i][p][0], 2 * H); mkA(&J.mA[p][1], A[i][p][1], 2 * H);
                    J.o1h[p] = A[i][p ^ 1][0]; J.o1l[p] = A[i][p ^ 1][1];           // h_prev half of the NEXT frame's operand
                    if (l == 0) { J.o2h[p] = q->A2[i][p][0]; J.o2l[p] = q->A2[i][p][1]; }
                    else if (!late) { J.o2h[p] = q->A3[i][p][0]; J.o2l[p] = q->A3[i][p][1]; }
                }
                J.o1pitch = 2 * H; J.o1col = H;
                J.o2pitch = (l == 0) ? 2 * H : H; J.o2col = 0;
                J.mW[0] = (bn == 64) ? w.mWhi64[l] : w.mWhi[l]; J.mW[1] = (bn == 64) ? w.mWlo64[l] : w.mWlo[l];
            }
        }
        {   // linear2
            SqJob& J = d->job[jO[i]];
            const int mask = skipmask_of(i, 0);
            J.kind = SQ_LIN; J.net = i; J.K = H; J.N = w.out; J.H = H; J.relu = 0; J.nt = (w.out + bn - 1) / bn;
            J.skipmask = mask; J.rowmask = mask; J.bias = w.b2;
            for (int p = 0; p < 2; ++p) { mkA(&J.mA[p][0], q->A3[i][p][0], H); mkA(&J.mA[p][1], q->A3[i][p][1], H); }
            J.mW[0] = (bn == 64) ? w.mW2hi64 : w.mW2hi; J.mW[1] = (bn == 64) ? w.mW2lo64 : w.mW2lo;
            switch (i) {
                case NET2: J.Y = q->Y2n; J.ldy = 72; break;         // MID turns both into operand columns (rnn3, rnn6, rnn7 / rnn8)
                case NET4: J.Y = q->Y4n; J.ldy = 72; break;
                case NET3: J.Y = s->Y3; J.ldy = 4; break;
                case NET6: J.Y = s->Y6; J.ldy = 4; break;
                case NET7: J.Y = s->Y7; J.ldy = 144; break;
                default: J.Y = s->Y8; J.ldy = 4; break;
            }
        }
    }
    if (rc != RC_OK) return rc;
    d->job[jA[NET6][1]].fshift = -1; d->job[jB[NET6][1]].fshift = -1;

    // ---- dependencies (job, frame lag); the "same job, previous frame" dependency is implicit for every job ----
    dep1(jPREP, jKIN, 2); dep1(jPREP, jB[NET4][1], kSqFlDepth); dep1(jPREP, jB[NET6][1], kSqFlDepth);
    for (int i = 0; i < NNETS; ++i) {
        // linear1 (main): its input columns, and the readers of the x-half it overwrites (two frames back)
        dep1(jL1[i][0], jPREP, 0);
        dep1(jL1[i][0], jA[i][0], 2); dep1(jL1[i][0], jA[i][1], 2);
        if (i == NET3 || i == NET6 || i == NET7 || i == NET8) dep1(jL1[i][0], jMID, 0);
        for (int late = 0; late < 2; ++late) {
            if (jA[i][late] < 0) continue;
            // LSTM-0: h_prev of every pass of the previous frame; readers of the planes it writes; then (second half) its linear1
            dep1(jA[i][late], jB[i][0], 2); dep1(jA[i][late], jB[i][1], 2);
            if (i == NET2) dep1(jA[i][late], jINIT, 1);
            static const bool xfirst_on = getenv("RC_SEQ_XFIRST") != nullptr;
            if (xfirst_on && i == NET4 && !late) {
                // (A/B switch, off: measured no gain — 232 vs 229 us per frame at 128 streams — because the tiles only get a CTA once
                // the vision updater's tiles have drained.)  rnn4's main pass: its linear1 only needs the frame's inputs and is done long
                // before the vision updater of the previous frame delivers the last h_prev rows -> x half first, the state dependency in
                // the middle of the main loop.
                d->job[jA[i][late]].xfirst = 1;
                dep1(jA[i][late], jL1[i][late], 0);
                dep2(jA[i][late], jA[i][late ^ 1], 1);
            } else {
                dep1(jA[i][late], jA[i][late ^ 1], 1);
                dep2(jA[i][late], jL1[i][late], 0);
            }
            // LSTM-1
            dep1(jB[i][late], jB[i][late ^ 1], 1);
            if (!late) dep1(jB[i][late], jO[i], 2);
            if (i == NET2) dep1(jB[i][late], jINIT, 1);
            dep2(jB[i][late], jA[i][late], 0);
        }
        // linear2: LSTM-1 of the frame; readers of its outputs in the previous frame(s)
        dep1(jO[i], jB[i][0], 0);
        if (i == NET2 || i == NET4) dep1(jO[i], jMID, 1);
        else dep1(jO[i], jKIN, 1);
    }
    dep1(jMID, jO[NET2], 0); dep1(jMID, jO[NET4], 0); dep1(jMID, jINIT, 1);
    dep1(jMID, jL1[NET7][0], 2); dep1(jMID, jL1[NET8][0], 2); dep1(jMID, jL1[NET3][0], 2); dep1(jMID, jL1[NET6][0], 2);
    dep1(jKIN, jO[NET6], 0); dep1(jKIN, jO[NET3], 0); dep1(jKIN, jO[NET7], 0); dep1(jKIN, jO[NET8], 0);
    dep1(jKIN, jL1[NET4][1], 1); dep1(jKIN, jL1[NET6][1], 1); dep1(jKIN, jINIT, 1);
    dep1(jINIT, jKIN, 0);
    for (int c = 0; c < 2; ++c) {       // vision updater: inputs from KIN; the main pass of the frame has read the shared x-half
        const int i = pl[c];
        dep1(jL1[i][1], jKIN, 0); dep1(jL1[i][1], jA[i][0], 0); dep1(jL1[i][1], jA[i][1], 2);
    }
    if (rc != RC_OK) return rc;

    int tile0 = 0;
    for (int j = 0; j < nj; ++j) { d->job[j].tile0 = tile0; if (d->job[j].kind < 2) tile0 += d->job[j].nt * MB; }   // GEMM queue
    d->tpf = tile0;
    d->npre = kSqRowTiles * MB;
    d->rpf = 4 * kSqRowTiles * MB;                                  // row queue of a frame: PREP(f + 1), MID, KIN, INIT
    d->row_job[0] = jPREP; d->row_job[1] = jMID; d->row_job[2] = jKIN; d->row_job[3] = jINIT;
    d->row_ctas = std::min(16, kSqRowTiles * MB);
    d->MB = MB; d->B = s->B;

    SqShared& sh = d->sh;
    sh.cfg = net->cfg;
    sh.model = net->model->d_const;
    sh.rows = s->rows;
    sh.gravity_all = s->gravity;
    for (int p = 0; p < 2; ++p) {
        sh.rcr[p] = q->rcr[p]; sh.conf[p] = q->conf[p]; sh.lerpw[p] = q->lerpw[p];
        for (int h = 0; h < 2; ++h) {
            sh.xa2[p][h] = q->A0[NET2][p][h]; sh.xa3[p][h] = q->A0[NET3][p][h]; sh.xa4[p][h] = q->A0[NET4][p][h];
            sh.xa6[p][h] = q->A0[NET6][p][h]; sh.xa7[p][h] = q->A0[NET7][p][h];
            sh.r2h0[p][h] = q->A1[NET2][p][h]; sh.r2h1[p][h] = q->A2[NET2][p][h];
        }
    }
    for (int p = 0; p < kSqFlDepth; ++p) sh.fl[p] = q->fl[p];
    for (int h = 0; h < 2; ++h) { sh.xa4l[h] = q->A0L[NET4][h]; sh.xa6l[h] = q->A0L[NET6][h]; }
    sh.k2p = net->nets[NET2].K1p; sh.k3p = net->nets[NET3].K1p; sh.k4p = net->nets[NET4].K1p; sh.k6p = net->nets[NET6].K1p; sh.k7p = net->nets[NET7].K1p;
    sh.Y2n = q->Y2n; sh.Y4n = q->Y4n; sh.J3DR = q->J3DR; sh.Y3 = s->Y3; sh.Y6 = s->Y6; sh.Y7 = s->Y7; sh.Y8 = s->Y8;
    sh.initflag = q->initflag;
    for (int l = 0; l < 3; ++l) { sh.Wi[l] = net->Wi[l]; sh.bi[l] = net->bi[l]; }
    sh.I1 = s->I1; sh.I2 = s->I2;
    sh.r2c0 = s->nb[NET2].c[0]; sh.r2c1 = s->nb[NET2].c[1];
    return RC_OK;
}

}  // namespace

bool rc_seq_supported(const rc_state* s) {
    return s && s->net->tc_ready && s->tc_ready && s->B > 8 && !s->net->cfg.live;
}

void rc_seq_destroy(rc_state* s) {
    if (!s || !s->seq) return;
    for (void* p : s->seq->allocs) cudaFree(p);
    delete s->seq->h_desc[0];
    delete s->seq->h_desc[1];
    delete s->seq;
    s->seq = nullptr;
}

const float* rc_seq_debug_j3dr(const rc_state* s) { return (s && s->seq) ? s->seq->J3DR : nullptr; }

int rc_seq_stats(rc_state* s, long long* out, int max_jobs) {
    if (!s || !s->seq || !out) return 0;
    const int n = std::min(max_jobs, s->seq->njobs);
    cudaDeviceSynchronize();
    cudaMemcpy(out, s->seq->stats, (size_t)n * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    if (s->seq->trace && s->seq->trace_F > 20) {
        // timeline of block 0, averaged over the frames: start / finish of every job relative to the finish of the previous frame's KIN
        const int F = s->seq->trace_F, nj = s->seq->njobs;
        std::vector<unsigned long long> tr((size_t)F * nj * 6);
        cudaMemcpy(tr.data(), s->seq->trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        const SqDesc* d = s->seq->h_desc[0] ? s->seq->h_desc[0] : s->seq->h_desc[1];
        int jkin = 0;
        for (int j = 0; j < nj; ++j) if (d->job[j].kind == SQ_KIN) jkin = j;
        fprintf(stderr, "[seq timeline, block 0, us after the previous frame's KIN finished: job kind/net | first start | last: x-half dep met, accumulators done, stores issued, published]\n");
        for (int j = 0; j < nj; ++j) {
            double a = 0, b[6] = {0, 0, 0, 0, 0, 0}; int c = 0;
            for (int f = 10; f < F - 2; ++f) {
                const unsigned long long ref = tr[((size_t)(f - 1) * nj + jkin) * 6 + 1];
                const unsigned long long* e = &tr[((size_t)f * nj + j) * 6];
                if (e[0] == ~0ULL || e[1] == 0ULL || ref == 0ULL) continue;
                a += ((double)e[0] - (double)ref) * 1e-3;
                for (int k = 1; k < 5; ++k) b[k] += e[k] ? ((double)e[k] - (double)ref) * 1e-3 : 0.0;
                ++c;
            }
            if (c) fprintf(stderr, "  job %2d kind %d net %2d nt %3d: %8.1f | %8.1f %8.1f %8.1f %8.1f  (%d frames)\n", j, d->job[j].kind, d->job[j].net, d->job[j].nt,
                           a / c, b[2] / c, b[3] / c, b[4] / c, b[1] / c, c);
        }
    }
    return n;
}

int rc_seq_run(rc_state* s, const StepIO& io, int T, int t0, int bn, void* stream) {
    RC_ARG(s && T > t0 && t0 >= 1 && (bn == 64 || bn == 128));
    if (!rc_seq_supported(s)) { rc_set_error("rc_seq_run: sequence kernel not available for this state"); return RC_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (!s->seq) {
        s->seq = new RcSeq();
        const int rc = sq_alloc_buffers(s);
        if (rc != RC_OK) { rc_seq_destroy(s); return rc; }
        static const bool stats_on = getenv("RC_SEQ_STATS") != nullptr;
        s->seq->stats_on = stats_on;
    }
    RcSeq* q = s->seq;
    const int bi = (bn == 64) ? 0 : 1;
    if (!q->h_desc[bi]) {
        q->h_desc[bi] = new SqDesc();
        const int rc = sq_build_jobs(s, bn, q->h_desc[bi]);
        if (rc != RC_OK) { delete q->h_desc[bi]; q->h_desc[bi] = nullptr; return rc; }
    }
    if (q->cap_T < T) {
        SQ_TRY(sq_alloc(q, &q->pf, (size_t)T * q->Bp, false));
        SQ_TRY(sq_alloc(q, &q->bf, (size_t)T * q->MB, false));
        q->cap_T = T;
    }
    SqDesc* d = q->h_desc[bi];
    d->T = T; d->t0 = t0; d->F = T - t0;
    d->sh.io = io;
    d->sh.cfg = s->net->cfg;
    d->sh.bf = q->bf;
    RC_CUDA(cudaMemcpyAsync(q->d_desc, d, sizeof(SqDesc), cudaMemcpyHostToDevice, st));
    RC_CUDA(cudaMemsetAsync(q->ctl, 0, ((size_t)kSqCtlPad + (size_t)SQ_MAXJOBS * q->MB) * sizeof(int), st));
    if (q->stats_on) RC_CUDA(cudaMemsetAsync(q->stats, 0, (size_t)SQ_MAXJOBS * 4 * sizeof(long long), st));
    // branch flags of every frame, init frames, block summaries
    const int F = T - t0;
    RC_LAUNCH(sq_flags_kernel, rc_cdiv((long long)F * q->Bp, 256), 256, 0, stream, s->net->cfg, io, s->B, q->Bp, T, t0, q->pf);
    RC_LAUNCH(sq_init_scan_kernel, rc_cdiv(s->B, 128), 128, 0, stream, (const RcRowState*)s->rows, s->B, q->Bp, T, t0, q->pf);
    RC_LAUNCH(sq_blk_kernel, F * q->MB, 128, 0, stream, (const int*)q->pf, q->Bp, q->MB, T, t0, q->bf);
    // fp32 LSTM state -> h_prev halves (parity 0)
    SqConvArgs ca;
    memset(&ca, 0, sizeof(ca));
    for (int i = 0; i < NNETS; ++i)
        for (int l = 0; l < 2; ++l) {
            const int k = i * 2 + l;
            ca.src[k] = s->nb[i].h[l];
            ca.hi[k] = (l == 0) ? q->A1[i][0][0] : q->A2[i][0][0];
            ca.lo[k] = (l == 0) ? q->A1[i][0][1] : q->A2[i][0][1];
            ca.H[k] = s->net->nets[i].H;
        }
    ca.B = s->B;
    RC_LAUNCH(sq_convert_kernel, dim3(std::max(1, std::min(rc_cdiv((long long)s->B * 320, 256), 64)), 2 * NNETS), 256, 0, stream, ca);
    RC_CHECK_LAUNCH();
    const long long total = (long long)F * d->tpf;
    if (total > 2000000000LL) { rc_set_error("rc_seq_run: %lld queue items exceed the 32-bit queue", total); return RC_ERR_ARG; }
    const int grid = sq_sm_count();                                 // one resident CTA per SM; the last row_ctas of them serve the row queue
    long long* stats = q->stats_on ? q->stats : nullptr;
    unsigned long long* trace = nullptr;
    if (q->stats_on) {
        if (q->trace_F < F) { SQ_TRY(sq_alloc(q, &q->trace, (size_t)F * SQ_MAXJOBS * 6, false)); q->trace_F = F; }
        std::vector<unsigned long long> init((size_t)F * q->njobs * 6, 0ULL);
        for (size_t e = 0; e < init.size(); e += 6) init[e] = ~0ULL;
        RC_CUDA(cudaMemcpy(q->trace, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        trace = q->trace;
    }
    if (s->prof_on) {
        if (s->prof_used + 2 > s->prof_ev.size())
            for (int e = 0; e < 16; ++e) { cudaEvent_t ev; RC_CUDA(cudaEventCreate(&ev)); s->prof_ev.push_back(ev); }
        RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], st));
    }
    if (bn == 64) {
        static bool attr = false;
        if (!attr) { RC_CUDA(cudaFuncSetAttribute(rc_seq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SqCfg<64>::kSmem)); attr = true; }
        RC_LAUNCH(rc_seq_kernel<64>, grid, kSqThreads, SqCfg<64>::kSmem, stream, (const SqDesc*)q->d_desc, q->ctl, stats, trace);
    } else {
        static bool attr = false;
        if (!attr) { RC_CUDA(cudaFuncSetAttribute(rc_seq_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SqCfg<128>::kSmem)); attr = true; }
        RC_LAUNCH(rc_seq_kernel<128>, grid, kSqThreads, SqCfg<128>::kSmem, stream, (const SqDesc*)q->d_desc, q->ctl, stats, trace);
    }
    RC_CHECK_LAUNCH();
    if (s->prof_on) RC_CUDA(cudaEventRecord(s->prof_ev[s->prof_used++], st));
    return RC_OK;
}
