// Persistent grouped tcgen05 kernel: every GEMM of one phase of the frame in ONE launch.
//
// The per-layer kernels of gemm_tc.cu pay, per launch, the launch latency, the TMEM allocation, the pipeline fill, a fully
// exposed epilogue (the 512 TMEM columns hold one tile) and the wave quantisation of 128..320 tiles on 148 SMs; measured, that
// leaves the tensor pipe 45 % busy inside the big layers and far less over the whole frame.  Here one CTA per SM stays resident
// for a whole phase (e.g. rnn2 + rnn4: 8 GEMMs, ~800 tiles) and pulls tiles from a global queue:
//
//   warp 0      tile scheduler + TMA producer: atomicAdd on the queue head, decode (job, row block, column tile), wait until
//               the producing job has finished that row block (acquire on a global counter), then stream the K blocks of the
//               tile through the 3-stage shared-memory ring.  It runs ahead of the MMAs by up to three stages, across tiles.
//   warp 1      MMA issuer (one thread): split-fp16 scheme of gemm_tc.cu (3 kind::f16 MMAs per K step).
//   warps 2-9   epilogue: TMEM -> registers -> bias / LSTM cell update -> c, h and the split fp16 operand of the NEXT layer,
//               then a release-increment of the row block's counter.
//
// TMEM plan: four 128-column buffers used as a ring; a tile takes three of them (corr, main-0, main-1; the two main
// accumulators alternate over the K steps to halve the length of the truncating accumulate chain).  Tile t+1 uses the buffer
// tile t left free as its corr and tile t's (corr, main-0) as its (main-0, main-1).  The epilogue therefore first drains corr and
// main-0 into registers (64 per thread) and hands them back — the MMAs of the next tile start while main-1 is still being
// read and the gate math runs.  The epilogue is off the critical path as long as it is shorter than a tile's main loop.
//
// Cross-CTA hand-over of an activation row block: the epilogue's st.global (generic proxy) must be visible to the consumer's
// TMA loads (async proxy): writer = stores, fence.proxy.async, __threadfence, CTA barrier, red.release.gpu; reader =
// ld.acquire.gpu spin, fence.proxy.async, TMA.  The queue is handed out in dependency order (a tile only depends on tiles with a
// smaller index, which are finished or held by a running CTA), so the scheme cannot deadlock whatever the residency.
#include <string.h>
#include <algorithm>
#include "rc_common.cuh"
#include "rc_tc.cuh"
#include "rc_tc_dev.cuh"
#include "rc_phase.cuh"

namespace {

constexpr int kPhThreads = 320;
constexpr int kPhStages = 3;
constexpr int kPhQ = 4;                  // depth of the tile-descriptor ring between the scheduler and its two consumers
constexpr int kPhBN = RC_TC_BN;
constexpr int kPhABytes = kTcBM * kTcBK * 2;
constexpr int kPhWBytes = kPhBN * kTcBK * 2;
constexpr int kPhStageBytes = 2 * kPhABytes + 2 * kPhWBytes;
constexpr int kPhSmem = kPhStages * kPhStageBytes + 1024;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kPhThreads, 1)
rc_tc_phase_kernel(const RcPhDesc* __restrict__ D, int* __restrict__ ctl, int MT, long long* __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kPhStages];
    __shared__ __align__(8) uint64_t bar_empty[kPhStages];
    __shared__ __align__(8) uint64_t bar_acc_full;
    __shared__ __align__(8) uint64_t bar_acc_free;
    __shared__ __align__(8) uint64_t tq_full[kPhQ];
    __shared__ __align__(8) uint64_t tq_empty[kPhQ];
    __shared__ int4 tq_tile[kPhQ];
    __shared__ int tile_start[RC_PH_MAXJOBS + 1];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int njobs = D->njobs;

    if (threadIdx.x == 0) {
        int acc = 0;
        for (int j = 0; j < njobs; ++j) {
            const int cnt = *D->job[j].count;
            tile_start[j] = acc;
            acc += ((cnt + kTcBM - 1) / kTcBM) * D->job[j].nt;
        }
        tile_start[njobs] = acc;
        for (int s = 0; s < kPhStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc_full), 1);
        mbar_init(smem_u32(&bar_acc_free), 8);                 // one arrival per epilogue warp
        for (int q = 0; q < kPhQ; ++q) { mbar_init(smem_u32(&tq_full[q]), 1); mbar_init(smem_u32(&tq_empty[q]), 9); }   // MMA thread + 8 epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int total = tile_start[njobs];
    if ((int)blockIdx.x >= total) return;                      // uniform; the remaining CTAs drain the whole queue
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                                    // K-block counter over all tiles of this CTA
            for (int q = 0;; ++q) {
                const int slot = q % kPhQ;
                mbar_wait(smem_u32(&tq_empty[slot]), (((uint32_t)q / kPhQ) & 1u) ^ 1u);
                const int t = atomicAdd(&ctl[0], 1);
                int j = -1, m = 0, n = 0;
                if (t < total) {
                    j = 0;
                    while (t >= tile_start[j + 1]) ++j;
                    const int local = t - tile_start[j], nt = D->job[j].nt;
                    m = local / nt;
                    n = local - m * nt;
                }
                tq_tile[slot] = make_int4(j, m, n, t);
                mbar_arrive(smem_u32(&tq_full[slot]));
                if (j < 0) break;
                const RcPhJob& J = D->job[j];
                if (trace) { trace[(size_t)t * 16 + 0] = ((long long)blockIdx.x << 32) | (unsigned)((j << 16) | (m << 8) | n); trace[(size_t)t * 16 + 1] = clock64(); }
                if (J.dep >= 0) {
                    const int need = D->job[J.dep].nt;
                    const int* flag = ctl + 1 + J.dep * MT + m;
                    if (ld_acquire_gpu(flag) < need) {
                        const long long t0 = clock64();
                        while (ld_acquire_gpu(flag) < need) {
                            __nanosleep(32);
                            if (clock64() - t0 > 4000000000LL) __trap();
                        }
                    }
                    fence_proxy_async_all();                    // the producer's generic-proxy stores before our async-proxy loads
                }
                if (trace) trace[(size_t)t * 16 + 2] = clock64();
                const int KB = J.K / kTcBK;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kPhStages;
                    const uint32_t ph = (it / kPhStages) & 1u;
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_expect_tx(full, kPhStageBytes);
                    const uint32_t base = smem_u32(smem + (size_t)s * kPhStageBytes);
                    tma_load_2d(base, (const CUtensorMap*)&J.mAhi, kb * kTcBK, m * kTcBM, full);
                    tma_load_2d(base + kPhABytes, (const CUtensorMap*)&J.mAlo, kb * kTcBK, m * kTcBM, full);
                    tma_load_2d(base + 2 * kPhABytes, (const CUtensorMap*)&J.mWhi, kb * kTcBK, n * kPhBN, full);
                    tma_load_2d(base + 2 * kPhABytes + kPhWBytes, (const CUtensorMap*)&J.mWlo, kb * kTcBK, n * kPhBN, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute UMMA::InstrDescriptor): c_format F32 = 1 at [4,6); a/b format F16 = 0; K-major both;
            // n_dim = N >> 3 at [17,23); m_dim = M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(kPhBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
            uint32_t it = 0;
            for (int q = 0;; ++q) {
                const int slot = q % kPhQ;
                mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
                const int4 t = tq_tile[slot];
                mbar_arrive(smem_u32(&tq_empty[slot]));
                if (t.x < 0) break;
                const int KB = D->job[t.x].K / kTcBK;
                if (q > 0) {                                    // the previous tile's corr / main-0 are in registers
                    mbar_wait(smem_u32(&bar_acc_free), (uint32_t)(q - 1) & 1u);
                    tc_fence_after();
                }
                const uint32_t d_corr = tmem_base + (uint32_t)(((3 * q) & 3) * kPhBN);
                const uint32_t d_m0 = tmem_base + (uint32_t)(((3 * q + 1) & 3) * kPhBN);
                const uint32_t d_m1 = tmem_base + (uint32_t)(((3 * q + 2) & 3) * kPhBN);
                int g = 0;
                if (trace) trace[(size_t)t.w * 16 + 3] = clock64();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kPhStages;
                    const uint32_t ph = (it / kPhStages) & 1u;
                    mbar_wait(smem_u32(&bar_full[s]), ph);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + (size_t)s * kPhStageBytes);
                    const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + kPhABytes);
                    const uint64_t dWhi = make_desc(base + 2 * kPhABytes), dWlo = make_desc(base + 2 * kPhABytes + kPhWBytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; ++k, ++g) {
                        const uint64_t adv = (uint64_t)(k * 2);      // 16 fp16 = 32 bytes = 2 x 16-byte units
                        tc_mma_f16((g & 1) ? d_m1 : d_m0, dAhi + adv, dWhi + adv, idesc, g >= 2 ? 1u : 0u);
                        tc_mma_f16(d_corr, dAhi + adv, dWlo + adv, idesc, g ? 1u : 0u);
                        tc_mma_f16(d_corr, dAlo + adv, dWhi + adv, idesc, 1u);
                    }
                    tc_commit(smem_u32(&bar_empty[s]));
                }
                tc_commit(smem_u32(&bar_acc_full));
                if (trace) trace[(size_t)t.w * 16 + 4] = clock64();
            }
        }
    } else {
        const int ewarp = warp - 2;
        const int q4 = warp & 3, half = ewarp >> 2;             // a warp may only read TMEM lanes 32 * (warp_id % 4) ..
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int q = 0;; ++q) {
            const int slot = q % kPhQ;
            mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
            const int4 t = tq_tile[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tq_empty[slot]));
            if (t.x < 0) break;
            const RcPhJob& J = D->job[t.x];
            const bool lstm = J.kind == 1;
            const int cnt = *J.count;
            const int m0 = t.y * kTcBM, n0 = t.z * kPhBN;
            const int mrow = m0 + q4 * 32 + lane;
            const int row = (mrow < cnt) ? J.rows[mrow] : -1;
            const int H = J.H;
            float4 cprev[2][2];
            if (lstm && row >= 0) {                             // cell state of this thread's 16 units, ahead of the accumulators
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const float* cp = J.C + (size_t)row * H + ((n0 + (half * 2 + cc) * 32) >> 2);
                    cprev[cc][0] = *reinterpret_cast<const float4*>(cp);
                    cprev[cc][1] = *reinterpret_cast<const float4*>(cp + 4);
                }
            }
            const uint32_t b_corr = (uint32_t)(((3 * q) & 3) * kPhBN);
            const uint32_t b_m0 = (uint32_t)(((3 * q + 1) & 3) * kPhBN);
            const uint32_t b_m1 = (uint32_t)(((3 * q + 2) & 3) * kPhBN);
            mbar_wait(smem_u32(&bar_acc_full), (uint32_t)q & 1u);
            tc_fence_after();
            if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 5] = clock64();
            // phase A: corr and main-0 of this thread's 64 columns into registers, then the two buffers go back to the MMA warp
            float acc[64];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const uint32_t col = (uint32_t)((half * 2 + cc) * 32);
                uint32_t v0[32], v1[32];
                tc_ld32(lane_base + b_m0 + col, v0);
                tc_ld32(lane_base + b_corr + col, v1);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) acc[cc * 32 + e] = fmaf(__uint_as_float(v1[e]), 4.8828125e-4f, __uint_as_float(v0[e]));   // main-0 + corr * 2^-11
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_free));
            if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 8] = clock64();
            // phase B: main-1, gate math, stores
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = half * 2 + cc;
                uint32_t v0[32];
                tc_ld32(lane_base + b_m1 + (uint32_t)(c * 32), v0);
                tc_ld_wait();
                float* a = acc + cc * 32;
#pragma unroll
                for (int e = 0; e < 32; ++e) a[e] += __uint_as_float(v0[e]);
                if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 9 + cc * 3] = clock64();
                if (row < 0) continue;
                const int nb = n0 + c * 32;
                if (lstm) {
                    const float cp[8] = {cprev[cc][0].x, cprev[cc][0].y, cprev[cc][0].z, cprev[cc][0].w,
                                         cprev[cc][1].x, cprev[cc][1].y, cprev[cc][1].z, cprev[cc][1].w};
                    float cn[8], hn[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(J.bias + nb + u * 4));
                        cn[u] = fmaf(sigm(a[u * 4 + 1] + b.y), cp[u], sigm(a[u * 4 + 0] + b.x) * tanhf(a[u * 4 + 2] + b.z));
                        hn[u] = sigm(a[u * 4 + 3] + b.w) * tanhf(cn[u]);
                    }
                    const size_t idx = (size_t)row * H + (nb >> 2);
                    if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 10 + cc * 3] = clock64();
                    *reinterpret_cast<float4*>(J.C + idx) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(J.C + idx + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                    *reinterpret_cast<float4*>(J.Hout + idx) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(J.Hout + idx + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                    if (J.nAhi) tc_store_split<8>(hn, (__half*)J.nAhi, (__half*)J.nAlo, (size_t)mrow * J.npitch + (nb >> 2));
                    if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 11 + cc * 3] = clock64();
                } else {
                    float* yrow = J.Y + (size_t)row * J.ldy;
                    const bool vec = ((J.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(J.Y) & 15) == 0) && (nb + 32 <= J.N);   // Y may be null when only the split copy is wanted
                    if (vec) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(J.bias + nb + e));
                            float4 y = make_float4(a[e] + b.x, a[e + 1] + b.y, a[e + 2] + b.z, a[e + 3] + b.w);
                            if (J.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                            if (J.Y) *reinterpret_cast<float4*>(yrow + nb + e) = y;
                            a[e] = y.x; a[e + 1] = y.y; a[e + 2] = y.z; a[e + 3] = y.w;
                        }
                        if (J.nAhi) {
#pragma unroll
                            for (int e = 0; e < 32; e += 8)
                                tc_store_split<8>(a + e, (__half*)J.nAhi, (__half*)J.nAlo, (size_t)mrow * J.npitch + nb + e);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int n = nb + e;
                            if (n < J.N) {
                                float y = a[e] + J.bias[n];
                                if (J.relu) y = fmaxf(y, 0.f);
                                yrow[n] = y;
                            }
                        }
                    }
                }
            }
            // publish the tile: every thread's stores -> CTA barrier -> ONE thread's proxy + gpu-scope fences (cumulative over
            // everything ordered before the barrier) -> release increment.  (A fence per thread cost ~20 k cycles per tile.)
            if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 6] = clock64();
            epi_bar_sync();
            if (ewarp == 0 && lane == 0) {
                fence_proxy_async_all();
                __threadfence();
                red_release_gpu_add(ctl + 1 + t.x * MT + t.y, 1);
                if (trace) trace[(size_t)t.w * 16 + 7] = clock64();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

struct SplitMultiArgs {
    RcSplitSegM seg[RC_PH_MAXSEGS];
    int nseg;
    int nzero;
    int* zero;
    int* advance;          // optional frame cursor to increment (sequence mode), nullptr otherwise
};

__global__ void __launch_bounds__(256) rc_split_multi_kernel(const __grid_constant__ SplitMultiArgs a) {
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int i = threadIdx.x; i < a.nzero; i += blockDim.x) a.zero[i] = 0;
        if (a.advance && threadIdx.x == 0) *a.advance += 1;
    }
    const RcSplitSegM& g = a.seg[blockIdx.y];
    const int cnt = *g.count;
    const int q4 = g.Kout >> 2;
    const long long total = (long long)cnt * q4;
    __half* const ghi = (__half*)g.hi;
    __half* const glo = (__half*)g.lo;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / q4), k = (int)(e % q4) * 4;
        const int r = g.rows[i];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < g.K) v = *reinterpret_cast<const float4*>(g.src + (size_t)r * g.ld + k);
        const float x[4] = {v.x, v.y, v.z, v.w};
        __half hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2half_rn(x[j]);
            lo[j] = __float2half_rn((x[j] - __half2float(hi[j])) * 2048.f);
        }
        __half2* ph = reinterpret_cast<__half2*>(ghi + (size_t)i * g.pitch + g.col0 + k);
        __half2* pl = reinterpret_cast<__half2*>(glo + (size_t)i * g.pitch + g.col0 + k);
        ph[0] = __halves2half2(hi[0], hi[1]); ph[1] = __halves2half2(hi[2], hi[3]);
        pl[0] = __halves2half2(lo[0], lo[1]); pl[1] = __halves2half2(lo[2], lo[3]);
    }
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace

int rc_tc_phase(const RcPhDesc* d_desc, int* d_ctl, int MT, int max_tiles, void* stream, long long* d_trace) {
    static bool attr_set = false;
    if (!attr_set) {
        RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPhSmem));
        attr_set = true;
    }
    const int grid = std::max(1, std::min(sm_count(), max_tiles));
    RC_LAUNCH(rc_tc_phase_kernel, grid, kPhThreads, kPhSmem, stream, d_desc, d_ctl, MT, d_trace);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

int rc_tc_split_multi(const RcSplitSegM* segs, int nseg, int B, int* zero, int nzero, void* stream, int* advance) {
    if (nseg < 1 || nseg > RC_PH_MAXSEGS) { rc_set_error("rc_tc_split_multi: %d segments", nseg); return RC_ERR_ARG; }
    SplitMultiArgs a;
    memset(&a, 0, sizeof(a));
    long long work = 0;
    for (int i = 0; i < nseg; ++i) {
        a.seg[i] = segs[i];
        work = std::max(work, (long long)B * (segs[i].Kout / 4));
    }
    a.nseg = nseg; a.zero = zero; a.nzero = nzero; a.advance = advance;
    dim3 grid((unsigned)std::max(1, (int)std::min<long long>(rc_cdiv(work, 256), 296)), (unsigned)nseg);
    RC_LAUNCH(rc_split_multi_kernel, grid, 256, 0, stream, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}
