// Single-stream (B = 1) frame as ONE cooperative kernel — the latency path of Net.forward_online (net/sig_mp.py:113-274).
//
// At B = 1 every weight byte (243 MB) is used once per frame: the frame is HBM-bound (37 us at the measured copy bandwidth) and
// the multi-kernel path spends most of its 131 us in ~60 dependent launches.  Here one persistent grid (one 512-thread CTA per
// SM) walks the whole frame: prep -> {rnn2 | rnn4} -> {rnn3 | rnn6} -> mid -> {rnn7 | rnn8} -> kin -> [init_net] -> late {rnn6 |
// rnn4}, with a grid-wide barrier (one atomic + acquire spin, ~1 us) where a layer needs the complete output of the previous
// one, and the two independent sub-nets of a group sharing the SMs.  Weights stream with 128-bit no-allocate loads, activations
// are read L2-coherently (ld.global.cg) because other CTAs of the same kernel produced them.  The arithmetic per output is the
// rc_gemv_vblock routine of the multi-kernel path (same K split, same reduction order), so both paths agree bit for bit
// (tests/test_gpu_parity.py::test_stream_kernel_matches_multi_kernel).
#include <stdlib.h>
#include <string.h>
#include "rc_fusion.cuh"
#include "rc_linear.cuh"
#include "rc_rows_warp.cuh"

namespace {

constexpr int kSkThreads = 512;                 // 2 virtual GEMV blocks of 8 warps
constexpr int kSkVb = kSkThreads / 256;

struct SkNet {
    RcLinear lin1, l0, l1, lin2;
    float *h0, *h1, *hn0, *hn1;
    int H;
};

struct SkArgs {
    RcNetCfg cfg;
    const RcModelConst* M;
    RcRowState* row;
    StepIO io;
    SkNet net[NNETS];
    RcLinear init[3];
    float *X2, *X3, *X4, *X6, *X7, *XI, *Y3, *Y6, *Y7, *Y8, *I3, *rcr, *conf, *lerpw, *gravity;
    int* flags;        // [0] frame flags, [1] need_init
    unsigned* bar;     // grid barrier counter, zeroed by the host before the launch
    unsigned long long* ts;   // optional phase timestamps (RC_STREAM_TS=1), CTA 0 thread 0
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void stamp(unsigned long long* ts, int& n) {
    if (ts && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); ts[n] = t; }
    ++n;
}

__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& epoch) {
    __syncthreads();                                            // every thread's writes happen-before thread 0's release below
    if (threadIdx.x == 0) {
        // release-increment, then acquire-spin: no full memory fences, no sleep (the barrier is on the critical path ~20x per frame)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        const unsigned target = (epoch + 1u) * gridDim.x;
        long long t0 = clock64();
        while (ld_acquire(bar) < target) {
            if (clock64() - t0 > 4000000000LL) __trap();       // a lost CTA must not hang the GPU
        }
    }
    ++epoch;
    __syncthreads();
}

// NI independent GEMV items (virtual-block indices vb0, vb0 + stride, ...) of one layer processed together by one virtual block:
// all weight loads of the NI items are issued before any is consumed, so HBM latency is paid once per batch instead of once per
// item.  Per item the arithmetic is exactly rc_gemv_vblock<1, LSTM, true> (same K split, FMA order, shuffle tree, partial order).
template <bool LSTM, int NI>
__device__ __forceinline__ void gemv_items(const RcLinear& a, int ksplit, int vb0, int stride, int nvb, int w, int lane,
                                           float (*part)[8][4], int bar) {
    const int jpb = 8 / ksplit, ks = w % ksplit, njobs = a.Nw >> 2, K = a.K1 + a.K2;
    int job[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int vb = vb0 + i * stride;
        job[i] = (vb < nvb) ? vb * jpb + w / ksplit : njobs;
    }
    float acc[NI][4];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
    for (int k = (ks * 32 + lane) * 4; k < K; k += ksplit * 128) {
        float4 wv[NI][4];
#pragma unroll
        for (int i = 0; i < NI; ++i)
            if (job[i] < njobs) {
#pragma unroll
                for (int r = 0; r < 4; ++r) wv[i][r] = rc_ldg_stream(a.W + (size_t)(job[i] * 4 + r) * K + k);
            }
        const float* xp = (k < a.K1) ? (a.X + k) : (a.X2 + (k - a.K1));
        const float4 xv = __ldcg(reinterpret_cast<const float4*>(xp));
#pragma unroll
        for (int i = 0; i < NI; ++i)
            if (job[i] < njobs) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[i][r] = fmaf(wv[i][r].x, xv.x, acc[i][r]);
                    acc[i][r] = fmaf(wv[i][r].y, xv.y, acc[i][r]);
                    acc[i][r] = fmaf(wv[i][r].z, xv.z, acc[i][r]);
                    acc[i][r] = fmaf(wv[i][r].w, xv.w, acc[i][r]);
                }
            }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float v = acc[i][r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[i][w][r] = v;
        }
    asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");
    if (ks == 0 && lane == 0) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            if (job[i] >= njobs) continue;
            float tot[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float sacc = 0.f;
                for (int q = 0; q < ksplit; ++q) sacc += part[i][w + q][r];
                tot[r] = sacc;
            }
            if (LSTM) {
                const float4 b = *reinterpret_cast<const float4*>(a.bias + job[i] * 4);
                float cn, hn;
                rc_lstm_cell(tot[0] + b.x, tot[1] + b.y, tot[2] + b.z, tot[3] + b.w, __ldcg(a.C + job[i]), cn, hn);
                a.C[job[i]] = cn;
                a.Hout[job[i]] = hn;
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int n = job[i] * 4 + r;
                    if (n < a.N) {
                        float y = tot[r] + a.bias[n];
                        if (a.relu) y = fmaxf(y, 0.f);
                        a.Y[n] = y;
                    }
                }
            }
        }
    }
    asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");                  // `part` is reused by the next batch
}

constexpr int kSkNI = 4;

template <bool LSTM>
__device__ __forceinline__ void run_one(const RcLinear* L, int first_slot, float (*part)[8][4]) {
    if (!L) return;
    const int v = threadIdx.x >> 8, w = (threadIdx.x >> 5) & 7, lane = threadIdx.x & 31;
    const int nslots = gridDim.x * kSkVb;
    int slot = blockIdx.x * kSkVb + v - first_slot;                          // rotate so that two layers start on different slots
    if (slot < 0) slot += nslots;
    const int ks = rc_gemv_ksplit(L->K1 + L->K2);
    const int jpb = 8 / ks;
    const int nvb = ((L->Nw >> 2) + jpb - 1) / jpb;
    for (int vb0 = slot; vb0 < nvb; vb0 += nslots * kSkNI)
        gemv_items<LSTM, kSkNI>(*L, ks, vb0, nslots, nvb, w, lane, part + v * kSkNI, v + 1);
}

// All CTAs share the virtual blocks of up to two independent layers (A first: the heavier one).
template <bool LSTM>
__device__ __forceinline__ void run_layers(const RcLinear* A, const RcLinear* B, float (*part)[8][4]) {
    run_one<LSTM>(A, 0, part);
    // B's items start where A's last partial round ended, so the tail slots of A and the head slots of B differ
    int off = 0;
    if (A) {
        const int ks = rc_gemv_ksplit(A->K1 + A->K2), jpb = 8 / ks;
        off = (((A->Nw >> 2) + jpb - 1) / jpb) % (gridDim.x * kSkVb);
    }
    run_one<LSTM>(B, off, part);
}

__device__ __forceinline__ void commit_h(const float* hn, float* h, int H) {     // h <- h_new, one warp
    const int lane = threadIdx.x & 31;
    for (int e = lane * 4; e < H; e += 128)
        *reinterpret_cast<float4*>(h + e) = __ldcg(reinterpret_cast<const float4*>(hn + e));
}

// linear1 -> LSTM0 -> LSTM1 -> (linear2) for one or two independent sub-nets; 4 (3 without linear2) grid barriers
__device__ __forceinline__ void run_group(const SkNet* A, const SkNet* B, bool lin2, unsigned* bar, unsigned& epoch, float (*part)[8][4]) {
    const bool last_warp = blockIdx.x == gridDim.x - 1 && (threadIdx.x >> 5) == kSkThreads / 32 - 1;
    run_layers<false>(A ? &A->lin1 : nullptr, B ? &B->lin1 : nullptr, part);
    grid_sync(bar, epoch);
    run_layers<true>(A ? &A->l0 : nullptr, B ? &B->l0 : nullptr, part);
    grid_sync(bar, epoch);
    if (last_warp) { if (A) commit_h(A->hn0, A->h0, A->H); if (B) commit_h(B->hn0, B->h0, B->H); }   // layer-0 h is no longer read
    run_layers<true>(A ? &A->l1 : nullptr, B ? &B->l1 : nullptr, part);
    grid_sync(bar, epoch);
    if (last_warp) { if (A) commit_h(A->hn1, A->h1, A->H); if (B) commit_h(B->hn1, B->h1, B->H); }
    if (lin2) {
        run_layers<false>(A ? &A->lin2 : nullptr, B ? &B->lin2 : nullptr, part);
        grid_sync(bar, epoch);
    }
}

__global__ void __launch_bounds__(kSkThreads, 1) rc_stream_kernel(const __grid_constant__ SkArgs a) {
    __shared__ float part[kSkVb * kSkNI][8][4];
    __shared__ RcPrepWarpSmem sprep;
    __shared__ RcKinWarpSmem skin;
    __shared__ RcModelConst Ms;
    unsigned epoch = 0;
    int nts = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stamp(a.ts, nts);
    const StepIO& io = a.io;

    // ---- prep (sig_mp.py:138-153) -----------------------------------------------------------------------------------
    if (blockIdx.x == 0) {
        const int* src = reinterpret_cast<const int*>(a.M);
        int* dst = reinterpret_cast<int*>(&Ms);
        for (int e = threadIdx.x; e < (int)(sizeof(RcModelConst) / 4); e += kSkThreads) dst[e] = src[e];
        if (warp == 0) {
            int inflags = 0;
            if (io.row_flags) inflags = __ldcg(io.row_flags) & 3;
            if (!io.first_tran) inflags &= ~RC_F_FIRST_TRAN;
            inflags |= RC_F_ACTIVE;
            const int f = rc_prep_warp(a.cfg, a.row->vision_count, sprep, io.j2dc, io.accc, io.oric, inflags, a.X2, a.X3, a.X4, a.X6, a.X7,
                                       a.rcr, a.conf, a.lerpw, lane);
            if (lane == 0) { a.flags[0] = f; a.flags[1] = 0; }
        }
    }
    grid_sync(a.bar, epoch);
    stamp(a.ts, nts);   // 1: after prep
    const int f = __ldcg(a.flags);
    const bool hi = (f & RC_F_HI) != 0, ff = (f & RC_F_FIRST_FRAME) != 0, r6b = (f & RC_F_R6B) != 0, late = (f & RC_F_LATE) != 0;

    // ---- inertial chain || vision chain (:144-165) ------------------------------------------------------------------------
    run_group(hi ? &a.net[NET4] : &a.net[NET2], hi ? &a.net[NET2] : nullptr, true, a.bar, epoch, part);
    if (ff) {                                                         // rnn6 runs on a first frame (:156), then again when c > lo (:161,165)
        run_group(&a.net[NET6], &a.net[NET3], true, a.bar, epoch, part);
        if (r6b) run_group(&a.net[NET6], nullptr, true, a.bar, epoch, part);
    } else {
        run_group(r6b ? &a.net[NET6] : &a.net[NET3], r6b ? &a.net[NET3] : nullptr, true, a.bar, epoch, part);
    }

    stamp(a.ts, nts);   // 2: after groups 1-2
    // ---- mid: camera->root rotation of the vision joints and the confidence lerp (:154-167) ------------------------------------
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float r[9], lw[2], ji[69], jc[69], out[69];
        for (int i = 0; i < 9; ++i) r[i] = __ldcg(a.rcr + i);
        lw[0] = __ldcg(a.lerpw); lw[1] = __ldcg(a.lerpw + 1);
        for (int i = 0; i < 69; ++i) { ji[i] = __ldcg(a.X3 + 72 + i); jc[i] = __ldcg(a.X6 + 171 + i); }
        rc_mid_row(f, r, lw, ji, jc, out);
        for (int i = 0; i < 69; ++i) a.X7[72 + i] = out[i];
    }
    grid_sync(a.bar, epoch);

    stamp(a.ts, nts);   // 3: after mid
    // ---- pose and contact nets (:169-170) -------------------------------------------------------------------------------
    run_group(&a.net[NET7], &a.net[NET8], true, a.bar, epoch, part);

    stamp(a.ts, nts);   // 4: after group 3
    // ---- kin (:173-273) ------------------------------------------------------------------------------------------------------
    if (blockIdx.x == 0 && warp == 0) {
        float r[9], g[3], ft[3] = {0.f, 0.f, 0.f}, y8[2], vr[3], pc[3];
        for (int i = 0; i < 9; ++i) r[i] = __ldcg(a.rcr + i);
        for (int i = 0; i < 3; ++i) { g[i] = __ldcg(a.gravity + i); vr[i] = __ldcg(a.Y3 + i); pc[i] = __ldcg(a.Y6 + i); }
        y8[0] = __ldcg(a.Y8); y8[1] = __ldcg(a.Y8 + 1);
        if (f & RC_F_FIRST_TRAN) for (int i = 0; i < 3; ++i) ft[i] = __ldcg(io.first_tran + i);
        const int need_init = rc_kin_warp(a.cfg, Ms, skin, a.row, f, a.Y7, y8, vr, pc, r, __ldcg(a.conf), g, ft, io.pose, io.tran, a.X4, a.X6, lane);
        if (need_init) {
            for (int i = lane; i < 80; i += 32) a.XI[i] = (i < 69) ? a.X7[72 + i] : 0.f;
            if (lane == 0) a.flags[1] = 1;
        }
    }
    grid_sync(a.bar, epoch);

    stamp(a.ts, nts);   // 5: after kin
    // ---- rnn2.init_net re-seed on the first c >= hi (:178-183) --------------------------------------------------------------
    if (__ldcg(a.flags + 1)) {
        for (int l = 0; l < 3; ++l) { run_layers<false>(&a.init[l], nullptr, part); grid_sync(a.bar, epoch); }
        if (blockIdx.x == 0) {
            const SkNet& n2 = a.net[NET2];
            for (int q = threadIdx.x; q < 2048; q += kSkThreads) {
                float* dst = (q < 512) ? n2.h0 : (q < 1024 ? n2.h1 : (q < 1536 ? a.net[NET2].l0.C : a.net[NET2].l1.C));
                dst[q & 511] = __ldcg(a.I3 + q);
            }
        }
    }
    // ---- vision updater keeps rnn6 / rnn4 warm on the synthetic key points (:263-271) -----------------------------------------
    stamp(a.ts, nts);   // 6: after init
    if (late) run_group(&a.net[NET4], &a.net[NET6], false, a.bar, epoch, part);
    stamp(a.ts, nts);   // 7: end
}

}  // namespace

bool rc_stream_supported(const rc_state* s) {
    // Opt-in (RC_STREAM_KERNEL=1): validated bit-exact against the multi-kernel path, but measured 150 us per frame vs 141 us for
    // the 2-stream CUDA-graph path — both are bound by the ~18-deep chain of dependent layers (6-8 us per phase), not by HBM.
    static const bool on = getenv("RC_STREAM_KERNEL") != nullptr;
    return on && s->B == 1;
}

int rc_stream_frame(rc_state* s, const StepIO& io, int any_first_frame, void* stream) {
    (void)any_first_frame;
    cudaStream_t st = (cudaStream_t)stream;
    const rc_net* n = s->net;
    static int grid = 0;
    if (!grid) {
        int dev = 0, sms = 0, per_sm = 0;
        RC_CUDA(cudaGetDevice(&dev));
        RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        RC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rc_stream_kernel, kSkThreads, 0));
        if (per_sm < 1) { rc_set_error("rc_stream_kernel does not fit on an SM"); return RC_ERR_CUDA; }
        grid = sms;
    }
    if (!s->sk_bar) {
        RC_CUDA(cudaMalloc(&s->sk_bar, 256));
        const int init[8] = {0, 1, 0, 0, 0, 0, 0, 0};  // rows = {0}, count = 1, flags
        RC_CUDA(cudaMalloc(&s->sk_rows, sizeof(init)));
        RC_CUDA(cudaMemcpy(s->sk_rows, init, sizeof(init), cudaMemcpyHostToDevice));
        s->flags_sk = s->sk_rows + 4;
    }
    SkArgs a;
    memset(&a, 0, sizeof(a));
    a.cfg = n->cfg; a.M = n->model->d_const; a.row = s->rows; a.io = io;
    const int* rows = s->sk_rows;
    const int* count = s->sk_rows + 1;
    for (int i = 0; i < NNETS; ++i) {
        const NetDev& w = n->nets[i];
        const NetBuf& nb = s->nb[i];
        SkNet& k = a.net[i];
        k.h0 = nb.h[0]; k.h1 = nb.h[1]; k.hn0 = nb.hn[0]; k.hn1 = nb.hn[1]; k.H = w.H;
        RcLinear base;
        memset(&base, 0, sizeof(base));
        base.rows = rows; base.count = count; base.H = w.H;
        k.lin1 = base; k.lin1.K1 = w.K1; k.lin1.ldx = w.K1; k.lin1.W = w.W1; k.lin1.bias = w.b1; k.lin1.N = w.H; k.lin1.Nw = w.H;
        k.lin1.Y = nb.a1; k.lin1.ldy = w.H; k.lin1.relu = 1;
        for (int l = 0; l < 2; ++l) {
            RcLinear& d = l ? k.l1 : k.l0;
            d = base;
            d.X = l ? nb.hn[0] : nb.a1; d.ldx = w.H; d.X2 = nb.h[l]; d.ldx2 = w.H; d.K1 = w.H; d.K2 = w.H;
            d.W = w.WL[l]; d.bias = w.bL[l]; d.N = 4 * w.H; d.Nw = 4 * w.H; d.C = nb.c[l]; d.Hout = nb.hn[l];
        }
        k.lin2 = base; k.lin2.X = nb.hn[1]; k.lin2.ldx = w.H; k.lin2.K1 = w.H; k.lin2.W = w.W2; k.lin2.bias = w.b2; k.lin2.N = w.out; k.lin2.Nw = w.out4;
    }
    a.net[NET2].lin1.X = s->X2; a.net[NET2].lin2.Y = s->X3 + 72; a.net[NET2].lin2.ldy = RC_K3;
    a.net[NET3].lin1.X = s->X3; a.net[NET3].lin2.Y = s->Y3; a.net[NET3].lin2.ldy = 4;
    a.net[NET4].lin1.X = s->X4; a.net[NET4].lin2.Y = s->X6 + 171; a.net[NET4].lin2.ldy = RC_K6;
    a.net[NET6].lin1.X = s->X6; a.net[NET6].lin2.Y = s->Y6; a.net[NET6].lin2.ldy = 4;
    a.net[NET7].lin1.X = s->X7; a.net[NET7].lin2.Y = s->Y7; a.net[NET7].lin2.ldy = 144;
    a.net[NET8].lin1.X = s->X7; a.net[NET8].lin2.Y = s->Y8; a.net[NET8].lin2.ldy = 4;
    const float* xin[3] = {s->XI, s->I1, s->I2};
    float* yout[3] = {s->I1, s->I2, s->I3};
    const int kin[3] = {kInitK0, 512, 1024};
    for (int l = 0; l < 3; ++l) {
        RcLinear& d = a.init[l];
        memset(&d, 0, sizeof(d));
        d.rows = rows; d.count = count;
        d.X = xin[l]; d.ldx = kin[l]; d.K1 = kin[l]; d.W = n->Wi[l]; d.bias = n->bi[l]; d.N = kInitDims[l + 1]; d.Nw = kInitDims[l + 1];
        d.Y = yout[l]; d.ldy = kInitDims[l + 1]; d.relu = (l < 2);
    }
    a.X2 = s->X2; a.X3 = s->X3; a.X4 = s->X4; a.X6 = s->X6; a.X7 = s->X7; a.XI = s->XI; a.Y3 = s->Y3; a.Y6 = s->Y6; a.Y7 = s->Y7; a.Y8 = s->Y8;
    a.I3 = s->I3; a.rcr = s->rcr; a.conf = s->conf; a.lerpw = s->lerpw; a.gravity = s->gravity; a.flags = s->flags_sk; a.bar = s->sk_bar;
    static const bool want_ts = getenv("RC_STREAM_TS") != nullptr;
    a.ts = want_ts ? reinterpret_cast<unsigned long long*>(s->sk_bar + 4) : nullptr;
    RC_CUDA(cudaMemsetAsync(s->sk_bar, 0, sizeof(unsigned), st));
    void* args[1] = {&a};
    RC_CUDA(cudaLaunchCooperativeKernel((const void*)rc_stream_kernel, dim3(grid), dim3(kSkThreads), args, 0, st));
    g_rc_launches.fetch_add(1, std::memory_order_relaxed);
    if (want_ts) {
        static int printed = 0;
        unsigned long long h[8];
        RC_CUDA(cudaStreamSynchronize(st));
        RC_CUDA(cudaMemcpy(h, s->sk_bar + 4, sizeof(h), cudaMemcpyDeviceToHost));
        if (printed++ % 50 == 10)
            fprintf(stderr, "[stream kernel us] prep %.1f | groups1-2 %.1f | mid %.1f | group3 %.1f | kin %.1f | init %.1f | late %.1f | total %.1f\n",
                    (h[1] - h[0]) * 1e-3, (h[2] - h[1]) * 1e-3, (h[3] - h[2]) * 1e-3, (h[4] - h[3]) * 1e-3, (h[5] - h[4]) * 1e-3,
                    (h[6] - h[5]) * 1e-3, (h[7] - h[6]) * 1e-3, (h[7] - h[0]) * 1e-3);
    }
    return RC_OK;
}
