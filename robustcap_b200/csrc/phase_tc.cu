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
//   warps 2-17  epilogue: TMEM -> registers -> bias / LSTM cell update -> c, h and the split fp16 operand of the NEXT layer,
//               then a release-increment of the row block's counter.
//
// TMEM plan: four 128-column buffers used as a ring; a tile takes three of them (corr, main-0, main-1; the two main
// accumulators alternate over the K steps to halve the length of the truncating accumulate chain).  Tile t+1 uses the buffer
// tile t left free as its corr and tile t's (corr, main-0) as its (main-0, main-1).  The epilogue therefore first drains corr and
// main-0 into registers (32 per thread) and hands them back — the MMAs of the next tile start while main-1 is still being
// read and the gate math runs.  The epilogue is off the critical path as long as it is shorter than a tile's main loop.
//
// Cross-CTA hand-over of an activation row block: the epilogue's st.global (generic proxy) must be visible to the consumer's
// TMA loads (async proxy): writer = stores, CTA barrier of the epilogue warps, then ONE thread's fence.proxy.async +
// __threadfence + red.release.gpu (cumulative over what the barrier ordered); reader = relaxed polls + one ld.acquire.gpu,
// fence.proxy.async, TMA.  An LSTM tile loads the h_prev half of its K range first (written before the launch) and only then
// waits for the producing layer.  The queue is handed out in dependency order (a tile only depends on tiles with a smaller
// index, which are finished or held by a running CTA), so the scheme cannot deadlock whatever the residency.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "rc_common.cuh"
#include "rc_tc.cuh"
#include "rc_tc_dev.cuh"
#include "rc_phase.cuh"
#include "rc_rows.h"

namespace {

constexpr int kPhEpiWarps = 16;           // 4 per scheduler: the gate math is latency bound, two warps per scheduler ran it at IPC 0.4
constexpr int kPhThreads = 64 + kPhEpiWarps * 32;
constexpr int kPhCPW = 4 / (kPhEpiWarps / 4);   // 32-column chunks per epilogue warp
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
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPhEpiWarps * 32) : "memory"); }

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

    if (warp == 0) {                                           // tile counts of all jobs: one lane per job (dependent global loads in parallel)
        int n = 0;
        if (lane < njobs) n = ((*D->job[lane].count + kTcBM - 1) / kTcBM) * D->job[lane].nt;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane < njobs) tile_start[lane] = incl - n;
        if (lane == njobs - 1) tile_start[njobs] = incl;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPhStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc_full), 1);
        mbar_init(smem_u32(&bar_acc_free), kPhEpiWarps);       // one arrival per epilogue warp
        for (int q = 0; q < kPhQ; ++q) { mbar_init(smem_u32(&tq_full[q]), 1); mbar_init(smem_u32(&tq_empty[q]), 1 + kPhEpiWarps); }   // MMA thread + epilogue warps
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
                // K order of an LSTM layer: the h_prev half first, the dependency wait in the middle (see the pair kernel)
                const int KB = J.K / kTcBK;
                const int KD = (J.kind == 1) ? KB / 2 : KB;
                for (int i = 0; i < KB; ++i, ++it) {
                    if (i == KB - KD) {
                        if (J.dep >= 0) {
                            const int need = D->job[J.dep].nt;
                            const int* flag = ctl + 1 + J.dep * MT + m;
                            if (ld_relaxed_gpu(flag) < need) {
                                const long long t0 = clock64();
                                while (ld_relaxed_gpu(flag) < need) {
                                    __nanosleep(32);
                                    if (clock64() - t0 > 4000000000LL) __trap();
                                }
                            }
                            (void)ld_acquire_gpu(flag);
                            fence_proxy_async_all();
                        }
                        if (trace) trace[(size_t)t * 16 + 2] = clock64();
                    }
                    const int kb = (i < KB - KD) ? KD + i : i - (KB - KD);
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
        const int q4 = warp & 3, part = ewarp >> 2;             // a warp may only read TMEM lanes 32 * (warp_id % 4) ..; part = its share of the columns
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
            float4 cprev[kPhCPW][2];
            if (lstm && row >= 0) {                             // cell state of this thread's units, ahead of the accumulators
#pragma unroll
                for (int cc = 0; cc < kPhCPW; ++cc) {
                    const float* cp = J.C + (size_t)row * H + ((n0 + (part * kPhCPW + cc) * 32) >> 2);
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
            // phase A: corr and main-0 of this thread's columns into registers, then the two buffers go back to the MMA warp
            float acc[kPhCPW * 32];
#pragma unroll
            for (int cc = 0; cc < kPhCPW; ++cc) {
                const uint32_t col = (uint32_t)((part * kPhCPW + cc) * 32);
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
            for (int cc = 0; cc < kPhCPW; ++cc) {
                const int c = part * kPhCPW + cc;
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

// ---- CTA-pair version (cta_group::2) -------------------------------------------------------------------------------------------
// Measured on the single-CTA kernel above: the main loop is bound by the operand traffic into the SMs (64 KB per K block and CTA:
// ~63 B/clk per SM alone = the SM's L2 port, ~42 B/clk when all 148 SMs stream = the L2 fabric), not by the tensor pipe (768 clk
// of MMAs per K block).  Here two CTAs of a TPC work on a 256-row x 128-column tile with ONE tcgen05.mma.cta_group::2 stream
// (M = 256) issued by the leader: each CTA loads its own 128 rows of A but only HALF of the W tile (64 of the 128 rows), i.e.
// 48 KB per K block and CTA, 25 % less traffic at every level and 4 ring stages instead of 3 in the same shared memory.
//   TMA:     both CTAs load into their own shared memory; every load completes on the LEADER's full barrier (.cta_group::2).
//   MMA:     leader only; tcgen05.commit with multicast mask 0b11 releases the stage / publishes the accumulators in both CTAs.
//   TMEM:    same four-buffer ring, in both CTAs (rows 0-127 of the tile in the leader, 128-255 in the peer).
//   queue:   the leader's scheduler thread grabs the tile and writes the descriptor into both CTAs' rings (DSMEM store +
//            cluster-scope mbarrier arrive); consumers of both CTAs release the slot on the leader's barrier.
//   epilogue / dependencies: per CTA exactly as above (each CTA owns one 128-row block).
constexpr int kPairStages = 4;
#ifndef RC_PAIR_PREFETCH
#define RC_PAIR_PREFETCH 0
#endif
constexpr int kPairPrefetch = RC_PAIR_PREFETCH;   // K blocks of L2 prefetch distance for the weight tiles (0 = off; 8 measured 3.5 % SLOWER)
constexpr int kPairWBytes = 64 * kTcBK * 2;
constexpr int kPairStageBytes = 2 * kPhABytes + 2 * kPairWBytes;      // 48 KB
constexpr int kPairSmem = kPairStages * kPairStageBytes + 1024;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// remote arrive that publishes a DSMEM store (the tile descriptor): cluster-scope release
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// remote arrive that only says "done reading" (slot / accumulator hand-back): default scope, no gpu-wide MEMBAR (as CUTLASS's ClusterBarrier)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Waits on barriers that the peer CTA / the async proxy signals use the plain (CTA-scope) try_wait, as CUTLASS does: what they guard
// lives in shared or tensor memory.  A cluster-scope acquire makes the compiler emit CCTL.IVALL (invalidate the whole L1) after
// every wait — measured at 43 % of all warp samples of this kernel.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, int4 v) {
    asm volatile("st.shared::cluster.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster) : "memory");
}
#ifdef RC_FAST_GATES      // timing experiment only (not fp32-accurate)
__device__ __forceinline__ float gate_sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }
#else
__device__ __forceinline__ float gate_sigm(float x) { return sigm(x); }
__device__ __forceinline__ float gate_tanh(float x) { return tanhf(x); }
#endif
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPhThreads, 1)
rc_tc_phase_pair_kernel(const RcPhDesc* __restrict__ D, int* __restrict__ ctl, int MT, long long* __restrict__ trace, int korder) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kPairStages];
    __shared__ __align__(8) uint64_t bar_empty[kPairStages];
    __shared__ __align__(8) uint64_t bar_acc_full;
    __shared__ __align__(8) uint64_t bar_acc_free;
    __shared__ __align__(8) uint64_t tq_full[kPhQ];
    __shared__ __align__(8) uint64_t tq_empty[kPhQ];
    __shared__ __align__(16) int4 tq_tile[kPhQ];
    __shared__ int tile_start[RC_PH_MAXJOBS + 1];
    __shared__ uint32_t tmem_base_s;

    rc_pdl_wait();                                                              // the pre-pass (operands, zeroed control block) has completed
    rc_pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    unsigned long long t_entry = 0;
    if (trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
    const bool leader = rank == 0;
    const int njobs = D->njobs;

    if (warp == 0) {                                           // 256-row tile counts of all jobs: one lane per job (dependent global loads in parallel)
        int n = 0;
        if (lane < njobs) n = ((*D->job[lane].count + 2 * kTcBM - 1) / (2 * kTcBM)) * D->job[lane].nt;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane < njobs) tile_start[lane] = incl - n;
        if (lane == njobs - 1) tile_start[njobs] = incl;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPairStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc_full), 1);
        mbar_init(smem_u32(&bar_acc_free), 2 * kPhEpiWarps);                    // leader: the epilogue warps of both CTAs
        for (int q = 0; q < kPhQ; ++q) {
            mbar_init(smem_u32(&tq_full[q]), 1);
            mbar_init(smem_u32(&tq_empty[q]), 2 + 2 * kPhEpiWarps);             // leader: MMA thread, peer producer, all epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int total = tile_start[njobs];
    long long* const tg = trace ? trace + (size_t)total * 16 : nullptr;        // (debug) row after the last tile: {~min entry, ~min first grab, max last publish, max exit} in globaltimer ns
    if ((int)(blockIdx.x >> 1) >= total) return;                               // uniform over the pair
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                                         // both CTAs' barriers exist before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int q = 0;; ++q) {
                const int slot = q % kPhQ;
                int4 tl;
                if (leader) {
                    mbar_wait_cluster(smem_u32(&tq_empty[slot]), (((uint32_t)q / kPhQ) & 1u) ^ 1u);
                    const int t = atomicAdd(&ctl[0], 1);
                    int j = -1, m = 0, n = 0;
                    if (t < total) {
                        j = 0;
                        while (t >= tile_start[j + 1]) ++j;
                        const int local = t - tile_start[j], nt = D->job[j].nt;
                        m = local / nt;
                        n = local - m * nt;
                    }
                    tl = make_int4(j, m, n, t);
                    tq_tile[slot] = tl;
                    st_cluster_v4(mapa_u32(smem_u32(&tq_tile[slot]), 1), tl);
                    mbar_arrive(smem_u32(&tq_full[slot]));
                    mbar_arrive_cluster(mapa_u32(smem_u32(&tq_full[slot]), 1));
                } else {
                    mbar_wait_cluster(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
                    tl = tq_tile[slot];
                    mbar_arrive_remote(mapa_u32(smem_u32(&tq_empty[slot]), 0));
                }
                const int j = tl.x, n = tl.z, t = tl.w;
                if (j < 0) break;
                const int mb = 2 * tl.y + (int)rank;                            // this CTA's 128-row block
                const RcPhJob& J = D->job[j];
                if (trace && leader) {
                    trace[(size_t)t * 16 + 0] = ((long long)blockIdx.x << 32) | (unsigned)((j << 16) | (tl.y << 8) | n); trace[(size_t)t * 16 + 1] = clock64();
                    if (q == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMax((unsigned long long*)tg + 1, ~g); }
                }
                // K order of an LSTM layer: the h_prev half of [x | h_prev] first — the split pre-pass wrote it before the launch —, then
                // the half the previous layer of this launch produces.  The dependency wait sits between the two, so the producing
                // layer's epilogue / publish latency hides behind half of this tile's main loop (ramp of a phase, vision updater).
                const int KB = J.K / kTcBK;
                const int KD = (J.kind == 1 && korder) ? KB / 2 : KB;  // K blocks that depend on the producing job
                for (int i = 0; i < KB; ++i, ++it) {
                    if (i == KB - KD) {
                        if (J.dep >= 0) {
                            const int need = D->job[J.dep].nt;
                            const int* flag = ctl + 1 + J.dep * MT + mb;
                            if (ld_relaxed_gpu(flag) < need) {         // relaxed polls: every acquire load invalidates the SM's L1
                                const long long t0 = clock64();
                                while (ld_relaxed_gpu(flag) < need) {
                                    __nanosleep(32);
                                    if (clock64() - t0 > 4000000000LL) __trap();
                                }
                            }
                            (void)ld_acquire_gpu(flag);
                            fence_proxy_async_all();
                        }
                        if (trace && leader) trace[(size_t)t * 16 + 2] = clock64();
                    }
                    const int kb = (i < KB - KD) ? KD + i : i - (KB - KD);
                    const int s = it % kPairStages;
                    const uint32_t ph = (it / kPairStages) & 1u;
                    if (kPairPrefetch > 0 && kb + kPairPrefetch < KB) {        // weights come from HBM on first touch: pull the tile's later K blocks into L2 early
                        tma_prefetch_2d((const CUtensorMap*)&J.mWhi64, (kb + kPairPrefetch) * kTcBK, n * kPhBN + (int)rank * 64);
                        tma_prefetch_2d((const CUtensorMap*)&J.mWlo64, (kb + kPairPrefetch) * kTcBK, n * kPhBN + (int)rank * 64);
                    }
                    mbar_wait_cluster(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[s]), 0);
                    if (leader) mbar_expect_tx(smem_u32(&bar_full[s]), 2 * kPairStageBytes);     // both CTAs' bytes land on the leader's barrier
                    const uint32_t base = smem_u32(smem + (size_t)s * kPairStageBytes);
                    tma_load_2d_pair(base, (const CUtensorMap*)&J.mAhi, kb * kTcBK, mb * kTcBM, full_leader);
                    tma_load_2d_pair(base + kPhABytes, (const CUtensorMap*)&J.mAlo, kb * kTcBK, mb * kTcBM, full_leader);
                    tma_load_2d_pair(base + 2 * kPhABytes, (const CUtensorMap*)&J.mWhi64, kb * kTcBK, n * kPhBN + (int)rank * 64, full_leader);
                    tma_load_2d_pair(base + 2 * kPhABytes + kPairWBytes, (const CUtensorMap*)&J.mWlo64, kb * kTcBK, n * kPhBN + (int)rank * 64, full_leader);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // M = 256 (two CTAs x 128 rows), N = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(kPhBN >> 3) << 17) | ((uint32_t)((2 * kTcBM) >> 4) << 24);
            uint32_t it = 0;
            for (int q = 0;; ++q) {
                const int slot = q % kPhQ;
                long long w0 = trace ? clock64() : 0, w_tq = 0, w_free = 0, w_full = 0;
                mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
                const int4 t = tq_tile[slot];
                mbar_arrive(smem_u32(&tq_empty[slot]));
                if (t.x < 0) break;
                if (trace) { const long long c = clock64(); w_tq = c - w0; w0 = c; }
                const int KB = D->job[t.x].K / kTcBK;
                if (q > 0) {
                    mbar_wait_cluster(smem_u32(&bar_acc_free), (uint32_t)(q - 1) & 1u);
                    tc_fence_after();
                }
                if (trace) w_free = clock64() - w0;
                const uint32_t d_corr = tmem_base + (uint32_t)(((3 * q) & 3) * kPhBN);
                const uint32_t d_m0 = tmem_base + (uint32_t)(((3 * q + 1) & 3) * kPhBN);
                const uint32_t d_m1 = tmem_base + (uint32_t)(((3 * q + 2) & 3) * kPhBN);
                int g = 0;
                if (trace) trace[(size_t)t.w * 16 + 3] = clock64();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kPairStages;
                    const uint32_t ph = (it / kPairStages) & 1u;
                    if (trace) w0 = clock64();
                    mbar_wait_cluster(smem_u32(&bar_full[s]), ph);
                    if (trace) w_full += clock64() - w0;
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + (size_t)s * kPairStageBytes);
                    const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + kPhABytes);
                    const uint64_t dWhi = make_desc(base + 2 * kPhABytes), dWlo = make_desc(base + 2 * kPhABytes + kPairWBytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; ++k, ++g) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        tc_mma_f16_pair((g & 1) ? d_m1 : d_m0, dAhi + adv, dWhi + adv, idesc, g >= 2 ? 1u : 0u);
                        tc_mma_f16_pair(d_corr, dAhi + adv, dWlo + adv, idesc, g ? 1u : 0u);
                        tc_mma_f16_pair(d_corr, dAlo + adv, dWhi + adv, idesc, 1u);
                    }
                    tc_commit_pair(smem_u32(&bar_empty[s]));
                }
                tc_commit_pair(smem_u32(&bar_acc_full));
                if (trace) { trace[(size_t)t.w * 16 + 4] = clock64(); trace[(size_t)t.w * 16 + 9] = w_tq; trace[(size_t)t.w * 16 + 10] = w_free; trace[(size_t)t.w * 16 + 11] = w_full; }
            }
        }
    } else {
        const int ewarp = warp - 2;
        const int q4 = warp & 3, part = ewarp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const bool tr_on = trace && leader && ewarp == 0 && lane == 0;
        for (int q = 0;; ++q) {
            const int slot = q % kPhQ;
            mbar_wait_cluster(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
            const int4 t = tq_tile[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tq_empty[slot]), 0));
            if (t.x < 0) break;
            const RcPhJob& J = D->job[t.x];
            const bool lstm = J.kind == 1;
            const int cnt = *J.count;
            const int mb = 2 * t.y + (int)rank;
            const int m0 = mb * kTcBM, n0 = t.z * kPhBN;
            const int mrow = m0 + q4 * 32 + lane;
            const int row = (mrow < cnt) ? J.rows[mrow] : -1;
            const int H = J.H;
            float4 cprev[kPhCPW][2];
            if (lstm && row >= 0) {
#pragma unroll
                for (int cc = 0; cc < kPhCPW; ++cc) {
                    const float* cp = J.C + (size_t)row * H + ((n0 + (part * kPhCPW + cc) * 32) >> 2);
                    cprev[cc][0] = *reinterpret_cast<const float4*>(cp);
                    cprev[cc][1] = *reinterpret_cast<const float4*>(cp + 4);
                }
            }
            const uint32_t b_corr = (uint32_t)(((3 * q) & 3) * kPhBN);
            const uint32_t b_m0 = (uint32_t)(((3 * q + 1) & 3) * kPhBN);
            const uint32_t b_m1 = (uint32_t)(((3 * q + 2) & 3) * kPhBN);
            mbar_wait_cluster(smem_u32(&bar_acc_full), (uint32_t)q & 1u);
            tc_fence_after();
            if (tr_on) trace[(size_t)t.w * 16 + 5] = clock64();
            float acc[kPhCPW * 32];
#pragma unroll
            for (int cc = 0; cc < kPhCPW; ++cc) {
                const uint32_t col = (uint32_t)((part * kPhCPW + cc) * 32);
                uint32_t v0[32], v1[32];
                tc_ld32(lane_base + b_m0 + col, v0);
                tc_ld32(lane_base + b_corr + col, v1);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) acc[cc * 32 + e] = fmaf(__uint_as_float(v1[e]), 4.8828125e-4f, __uint_as_float(v0[e]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_acc_free), 0));
            if (tr_on) trace[(size_t)t.w * 16 + 8] = clock64();
#pragma unroll
            for (int cc = 0; cc < kPhCPW; ++cc) {
                const int c = part * kPhCPW + cc;
                uint32_t v0[32];
                tc_ld32(lane_base + b_m1 + (uint32_t)(c * 32), v0);
                tc_ld_wait();
                float* a = acc + cc * 32;
#pragma unroll
                for (int e = 0; e < 32; ++e) a[e] += __uint_as_float(v0[e]);
                if (row < 0) continue;
                const int nb = n0 + c * 32;
                if (lstm) {
                    const float cp[8] = {cprev[cc][0].x, cprev[cc][0].y, cprev[cc][0].z, cprev[cc][0].w,
                                         cprev[cc][1].x, cprev[cc][1].y, cprev[cc][1].z, cprev[cc][1].w};
                    float cn[8], hn[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(J.bias + nb + u * 4));
                        cn[u] = fmaf(gate_sigm(a[u * 4 + 1] + b.y), cp[u], gate_sigm(a[u * 4 + 0] + b.x) * gate_tanh(a[u * 4 + 2] + b.z));
                        hn[u] = gate_sigm(a[u * 4 + 3] + b.w) * gate_tanh(cn[u]);
                    }
                    const size_t idx = (size_t)row * H + (nb >> 2);
                    *reinterpret_cast<float4*>(J.C + idx) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(J.C + idx + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                    *reinterpret_cast<float4*>(J.Hout + idx) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(J.Hout + idx + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                    if (J.nAhi) tc_store_split<8>(hn, (__half*)J.nAhi, (__half*)J.nAlo, (size_t)mrow * J.npitch + (nb >> 2));
                } else {
                    float* yrow = J.Y + (size_t)row * J.ldy;
                    const bool vec = ((J.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(J.Y) & 15) == 0) && (nb + 32 <= J.N);
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
            if (tr_on) trace[(size_t)t.w * 16 + 6] = clock64();
            epi_bar_sync();
            if (ewarp == 0 && lane == 0) {
                fence_proxy_async_all();
                __threadfence();
                red_release_gpu_add(ctl + 1 + t.x * MT + mb, 1);               // always, also for a row block beyond the list: consumers count tiles
                if (tr_on) { trace[(size_t)t.w * 16 + 7] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMax((unsigned long long*)tg + 2, g); }
            }
        }
    }
    if (tg && threadIdx.x == 0) {
        unsigned long long t_exit;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
        atomicMax((unsigned long long*)tg + 0, ~t_entry);                  // min via max of the complement (the row starts zeroed)
        atomicMax((unsigned long long*)tg + 3, t_exit);
    }
    tc_fence_before();
    cluster_sync_all();                                                         // the leader's MMAs read the peer's shared memory until the end
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
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
    rc_pdl_wait();
    rc_pdl_trigger();
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
        float x[4] = {v.x, v.y, v.z, v.w};
        if (g.mid_flags && k + 3 >= 72 && k < 141) {              // joint blend of the rnn7 / rnn8 input, same arithmetic as rc_mid_joint
            const int f = g.mid_flags[r];
            float rcr[9], lw[2];
#pragma unroll
            for (int q = 0; q < 9; ++q) rcr[q] = g.mid_rcr[r * 9 + q];
            lw[0] = g.mid_lerpw[r * 2]; lw[1] = g.mid_lerpw[r * 2 + 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = k + j;
                if (col < 72 || col >= 141) continue;
                const int jt = (col - 72) / 3, c = (col - 72) % 3;
                float out[3];
                rc_mid_joint(f, rcr, lw, g.mid_x3 + (size_t)r * RC_K3 + 72 + jt * 3, g.mid_x6 + (size_t)r * RC_K6 + 171 + jt * 3, out);
                x[j] = out[c];
                if (g.mid_out) g.mid_out[(size_t)r * RC_K7 + col] = x[j];
            }
        }
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
    static int pair = -1;
    // Default: the single-CTA kernel (128-row tiles).  RC_PH_PAIR=1 selects the CTA-pair kernel (cta_group::2, 256-row tiles): 25 % less
    // L2 -> SM traffic, same main-loop rate (both sit at the shared-memory port), but the row lists of a frame (~768 / ~256 streams) end
    // in a mostly empty 256-row tile more often, so it is ~2 % slower on the mixed-confidence workload (same-box A/B: 544 vs 535 us).
    if (pair < 0) { const char* e = getenv("RC_PH_PAIR"); pair = e ? atoi(e) : 0; }
    if (pair) {
        static bool attr_set = false;
        if (!attr_set) {
            RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
            attr_set = true;
        }
        const int pairs = std::max(1, std::min(sm_count() / 2, max_tiles));             // max_tiles bounds the 256-row tiles too
        static int korder = -1;
        if (korder < 0) { const char* e = getenv("RC_PH_KORDER"); korder = e ? atoi(e) : 1; }   // 0: dependency wait before the first K block (A/B switch)
        RC_LAUNCH_PDL(rc_tc_phase_pair_kernel, 2 * pairs, kPhThreads, kPairSmem, stream, d_desc, d_ctl, MT, d_trace, korder);
        RC_CHECK_LAUNCH();
        return RC_OK;
    }
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
    static bool attr_set = false;
    if (!attr_set) {      // same shared-memory carve-out as the grouped kernel that follows (no L1 / shared re-partition between them)
        if (!getenv("RC_NO_CARVEOUT_HINT"))
            RC_CUDA(cudaFuncSetAttribute(rc_split_multi_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    dim3 grid((unsigned)std::max(1, (int)std::min<long long>(rc_cdiv(work, 256), 296)), (unsigned)nseg);
    RC_LAUNCH_PDL(rc_split_multi_kernel, grid, 256, 0, stream, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}
