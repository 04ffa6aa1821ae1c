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
#include <type_traits>
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

// N256 = true (RC_PH_PAIR=3, opt-in experiment): the products A-hi x W-hi and A-hi x W-lo of a K step as ONE N = 256 MMA — the W-hi and
// W-lo tiles are adjacent in the stage, the accumulators [main | corr] adjacent in tensor memory — followed by A-lo x W-hi (N = 128) into
// corr: two MMA instructions per K step instead of three, A-hi read from shared memory once (36 instead of 40 KB of shared-memory traffic
// per K step).  A tile's accumulators are then 256 columns; the two 256-column slots are a ring of SEGMENTS (see the MMA warp).
// Measured (B = 1024, same box, alternating runs): with two segments per LSTM tile (h_prev half, x half) 540 / 528 us per frame against
// 555 / 570 us (K = 2560 tiles 44 k -> 39-44 k clk of MMA issue), but one stream of test_grouped_kernel_odd_row_blocks[130] lands 1.21e-4 rad
// from the float64 oracle (bound 1.2e-4; the three-buffer scheme with its two interleaved chains: <= 5.4e-5); with segments of <= 10 K
// blocks (below) that stream is at 1.11e-4 and the gain is gone (544 / 555 vs 545 us: twice the drains).  N256 = 2 (RC_PH_PAIR=4) keeps the
// default's K-step-interleaved chains (even steps -> slot 0, odd -> slot 1, both slots per tile): default-grade accuracy, +1 %.  Neither is
// the default (profiles/r03_pair256.md).
__device__ __forceinline__ int n256_segments(int KB) { return KB >= 32 ? 4 : (KB >= 16 ? 2 : 1); }

template <int N256>                                    // 0: three accumulators (default), 1: fused MMA + segment ring, 2: fused MMA + interleaved chains
__global__ void __launch_bounds__(kPhThreads, 1)
rc_tc_phase_kernel(const RcPhDesc* __restrict__ D, int* __restrict__ ctl, int MT, long long* __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kPhStages];
    __shared__ __align__(8) uint64_t bar_empty[kPhStages];
    __shared__ __align__(8) uint64_t bar_acc_full;
    __shared__ __align__(8) uint64_t bar_acc_free;
    __shared__ __align__(8) uint64_t bar2_full[2], bar2_free[2];      // N256: per accumulator slot
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
        for (int q = 0; q < 2; ++q) { mbar_init(smem_u32(&bar2_full[q]), 1); mbar_init(smem_u32(&bar2_free[q]), kPhEpiWarps); }
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
            uint32_t it = 0, gseg = 0;
            for (int q = 0;; ++q) {
                const int slot = q % kPhQ;
                mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
                const int4 t = tq_tile[slot];
                mbar_arrive(smem_u32(&tq_empty[slot]));
                if (t.x < 0) break;
                const int KB = D->job[t.x].K / kTcBK;
                if (N256 == 2) {
                    // fused MMA with the accumulate chains of the default scheme: even K steps into slot 0 ([main-0 | corr-0]), odd ones into
                    // slot 1 ([main-1 | corr-1]) — main is the same two K-step-interleaved chains, corr two instead of one.  Both slots
                    // belong to one tile; the next tile's first even step waits for slot 0 only, its first odd step for slot 1.
                    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
                    if (trace) trace[(size_t)t.w * 16 + 3] = clock64();
                    int g = 0;
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const int s = it % kPhStages;
                        const uint32_t ph = (it / kPhStages) & 1u;
                        mbar_wait(smem_u32(&bar_full[s]), ph);
                        tc_fence_after();
                        const uint32_t base = smem_u32(smem + (size_t)s * kPhStageBytes);
                        const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + kPhABytes);
                        const uint64_t dWhi = make_desc(base + 2 * kPhABytes);       // [W-hi | W-lo]: 256 rows of 128 bytes
#pragma unroll
                        for (int k = 0; k < kTcBK / 16; ++k, ++g) {
                            if (g < 2 && q > 0) {                   // the previous tile has left this slot
                                mbar_wait(smem_u32(&bar2_free[g]), (uint32_t)(q - 1) & 1u);
                                tc_fence_after();
                            }
                            const uint64_t adv = (uint64_t)(k * 2);
                            const uint32_t d_main = tmem_base + (uint32_t)((g & 1) * 256);
                            tc_mma_f16(d_main, dAhi + adv, dWhi + adv, idesc2, g >= 2 ? 1u : 0u);
                            tc_mma_f16(d_main + 128u, dAlo + adv, dWhi + adv, idesc, 1u);
                        }
                        tc_commit(smem_u32(&bar_empty[s]));
                    }
                    tc_commit(smem_u32(&bar_acc_full));
                    if (trace) trace[(size_t)t.w * 16 + 4] = clock64();
                    continue;
                }
                if (N256) {
                    // accumulator slots of 256 columns ([main | corr]) used as a ring of SEGMENTS of <= 10 K blocks in alternating slots: the
                    // epilogue drains a segment into its registers while the next one accumulates, so the truncating accumulate chains are at
                    // most 40 steps long (three-buffer scheme: two interleaved chains of up to 80) and the MMA stream has no bubble.
                    const int nseg = n256_segments(KB), KBs = KB / nseg;
                    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
                    if (trace) trace[(size_t)t.w * 16 + 3] = clock64();
                    for (int sg = 0; sg < nseg; ++sg, ++gseg) {
                        const int sl = (int)(gseg & 1u);
                        if (gseg > 1) {                         // the segment before the previous one has left this slot
                            mbar_wait(smem_u32(&bar2_free[sl]), ((gseg >> 1) - 1u) & 1u);
                            tc_fence_after();
                        }
                        const uint32_t d_main = tmem_base + (uint32_t)(sl * 256), d_cr = d_main + 128u;
                        for (int kb = 0; kb < KBs; ++kb, ++it) {
                            const int s = it % kPhStages;
                            const uint32_t ph = (it / kPhStages) & 1u;
                            mbar_wait(smem_u32(&bar_full[s]), ph);
                            tc_fence_after();
                            const uint32_t base = smem_u32(smem + (size_t)s * kPhStageBytes);
                            const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + kPhABytes);
                            const uint64_t dWhi = make_desc(base + 2 * kPhABytes);       // [W-hi | W-lo]: 256 rows of 128 bytes
#pragma unroll
                            for (int k = 0; k < kTcBK / 16; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);
                                tc_mma_f16(d_main, dAhi + adv, dWhi + adv, idesc2, (kb | k) ? 1u : 0u);
                                tc_mma_f16(d_cr, dAlo + adv, dWhi + adv, idesc, 1u);
                            }
                            tc_commit(smem_u32(&bar_empty[s]));
                        }
                        tc_commit(smem_u32(&bar2_full[sl]));
                    }
                    if (trace) trace[(size_t)t.w * 16 + 4] = clock64();
                    continue;
                }
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
        uint32_t eseg = 0;                                       // N256: accumulator segments consumed so far
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
            float acc[kPhCPW * 32];
            if (N256 == 2) {
                mbar_wait(smem_u32(&bar_acc_full), (uint32_t)q & 1u);
                tc_fence_after();
                if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 5] = clock64();
                const bool one = J.K / 16 < 2;                      // a single K step never touches slot 1
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (!(sl && one)) {
#pragma unroll
                        for (int cc = 0; cc < 2 * kPhCPW; ++cc) {
                            const uint32_t col = (uint32_t)(sl * 256 + part * kPhCPW * 32 + cc * 16);
                            uint32_t v0[16], v1[16];
                            tc_ld16(lane_base + col, v0);
                            tc_ld16(lane_base + col + 128u, v1);
                            tc_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float x = fmaf(__uint_as_float(v1[e]), 4.8828125e-4f, __uint_as_float(v0[e]));   // main-s + corr-s * 2^-11
                                acc[cc * 16 + e] = sl ? acc[cc * 16 + e] + x : x;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar2_free[sl]));
                }
            } else if (N256) {
                const int nseg = n256_segments(J.K / kTcBK);
                for (int sg = 0; sg < nseg; ++sg, ++eseg) {
                    const uint32_t sl = eseg & 1u;
                    mbar_wait(smem_u32(&bar2_full[sl]), (eseg >> 1) & 1u);
                    tc_fence_after();
                    if (sg == nseg - 1 && trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 5] = clock64();
#pragma unroll
                    for (int cc = 0; cc < 2 * kPhCPW; ++cc) {                   // 16 columns at a time: 32 accumulators + 32 in flight
                        const uint32_t col = sl * 256u + (uint32_t)(part * kPhCPW * 32 + cc * 16);
                        uint32_t v0[16], v1[16];
                        tc_ld16(lane_base + col, v0);
                        tc_ld16(lane_base + col + 128u, v1);
                        tc_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float x = fmaf(__uint_as_float(v1[e]), 4.8828125e-4f, __uint_as_float(v0[e]));   // main + corr * 2^-11 of this half
                            acc[cc * 16 + e] = sg ? acc[cc * 16 + e] + x : x;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar2_free[sl]));
                }
            } else {
            mbar_wait(smem_u32(&bar_acc_full), (uint32_t)q & 1u);
            tc_fence_after();
            if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 5] = clock64();
            // phase A: corr and main-0 of this thread's columns into registers, then the two buffers go back to the MMA warp
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
            }
            if (trace && ewarp == 0 && lane == 0) trace[(size_t)t.w * 16 + 8] = clock64();
            // phase B: main-1, gate math, stores
#pragma unroll
            for (int cc = 0; cc < kPhCPW; ++cc) {
                const int c = part * kPhCPW + cc;
                float* a = acc + cc * 32;
                if (!N256) {
                    uint32_t v0[32];
                    tc_ld32(lane_base + b_m1 + (uint32_t)(c * 32), v0);
                    tc_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) a[e] += __uint_as_float(v0[e]);
                }
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

// ---- CTA-pair version with 256 x 256 tiles (RC_PH_PAIR=2, opt-in) ------------------------------------------------------------------
// What bounds the two kernels above is the operand traffic INTO an SM per MMA cycle (measured: ~63 B/clk into one SM alone, ~42 B/clk
// when all 148 stream): a K block of 64 brings 256 (m + n) bytes of split operands for 768 m n / 128^2 MMA cycles, i.e. 85 B/clk for a
// 128 x 128 tile per CTA and 62 B/clk for the 256 x 128 pair tile.  A pair tile of 256 rows x 256 columns — per CTA its 128 rows of A and
// its HALF (128 rows) of the W tile — needs 42.7 B/clk and, in shared memory, 40 KB of TMA writes + MMA reads per 384 MMA cycles, so the
// main loop can run at the tensor pipe's rate (measured: 129 clk per N = 256 MMA of nominally 128).  Price: the accumulators of ONE
// tile fill the tensor memory (main 256 + corr 256 columns), so
//   * the two accumulators are handed over separately and the MMA order inside a K block is corr, corr, ..., then main: the epilogue
//     drains corr while the tile's last main MMAs run, and main while the NEXT tile's first corr MMAs run;
//   * the truncating accumulate chain of `main` is not split over two alternating buffers but in time: an LSTM tile with K >= 2048 hands
//     main to the epilogue at half K (the point where its K order switches from the h_prev half to the x half), the epilogue keeps the
//     partial sums in registers (64 per thread) and the second half starts a fresh chain — same chain lengths as above (for K = 1024
//     the hand-over would serialise the epilogue with the second half of the 12 k-clk main loop, so those tiles run one 64-step chain);
//   * per job the MMA N is the job's width rounded up to 16 (linear2: 16 .. 144 columns instead of a padded 128 / 256);
//   * epilogue: a lane owns one ROW of the tile, so plain stores touch 32 rows x 16 bytes per instruction; everything goes through a
//     per-warp staging buffer instead, consecutive lanes on consecutive 16-byte pieces of a row, and the previous cell state arrives
//     in that buffer by cp.async issued when the tile is picked up; the gate math is branch-free (sigm()) so the units interleave.
// Template TW = 128 runs 128-column tiles with the same code (32 columns per epilogue thread, 64-row W boxes, corr as two partial sums so
// that no two consecutive MMAs share an accumulator).  Measured (B = 1024, mixed): a cta_group::2 MMA costs ~123 clk whatever its N
// (80 .. 256), so narrow tiles halve the work per MMA slot — TW = 128 is slower everywhere and only used for the first-frame rnn6 pass
// and the vision updater, where 256-column tiles would leave half of the CTA pairs without a tile.
// Result of the round (profiles/r03_pair256.md): main loop 62-68 k clk per K = 2560 tile (61.4 k ideal), but 288 tiles of up to 68 k clk
// on 74 pairs pack badly (per-pair span median 210 k, slowest 295 k clk: 4 x 20 tiles per layer for 74 pairs, the second round and the
// LSTM-1 -> linear2 chain form the tail) and the 64-column epilogue spills (64 accumulators of 96 registers): 576 us per frame against
// 558 us of the 128 x 128 kernel on the same box.  It stays opt-in; larger batches per GPU (more row blocks) are where it would pay.
constexpr int kP2Stages = 3;
constexpr int kP2StageBytes = 2 * kPhABytes + 2 * kPhWBytes;          // 64 KB: A hi / lo (this CTA's 128 rows), W hi / lo (this CTA's half of the tile's columns)
constexpr int kP2Smem = kP2Stages * kP2StageBytes + 1024;
constexpr int kP2Levels = 4;
constexpr int kP2FlushKB = 32;                                        // LSTM tiles with K >= 2048 restart the main chain at half K (shorter ones: their epilogue would serialise with the second half)                                          // linear1, LSTM-0, LSTM-1, linear2

// Staging buffer of an epilogue warp: 32 rows x 4 pieces of 16 bytes; piece p of row r at slot r * 4 + (p ^ ((r >> 1) & 3)) —
// conflict-free both for "lane = row, fixed piece" and for "4 consecutive lanes = the 4 pieces of a row".
__device__ __forceinline__ uint32_t epi_slot(int r, int p) { return (uint32_t)(r * 4 + (p ^ ((r >> 1) & 3))) * 16u; }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 f4u(float a, float b, float c, float d) { return make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d)); }
// two floats -> packed fp16 pair (round to nearest even, x0 in the low half) and back, entirely in registers
__device__ __forceinline__ uint32_t pack_h2(float x0, float x1) {
    uint32_t u;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x1), "f"(x0));
    return u;
}
__device__ __forceinline__ void unpack_h2(uint32_t u, float& f0, float& f1) {
    asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(f0), "=f"(f1) : "r"(u));
}
// 8 floats -> one 16-byte piece of fp16 halves: hi, or lo * 2^11 (same values as tc_store_split)
template <bool LO>
__device__ __forceinline__ uint4 epi_split8(const float* x) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x0 = x[2 * i], x1 = x[2 * i + 1];
        const uint32_t h = pack_h2(x0, x1);
        if (LO) {
            float f0, f1;
            unpack_h2(h, f0, f1);
            w[i] = pack_h2((x0 - f0) * 2048.f, (x1 - f1) * 2048.f);
        } else {
            w[i] = h;
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
// The buffer holds 4 pieces of every row (written by the row's lane); the warp stores them with 4 consecutive lanes per row:
// instruction k covers rows 8 k .. 8 k + 7.  `off` = THIS lane's row segment (offset in 16-byte units, negative for a row without a
// stream), `base` = destination of the piece this lane serves (already + 16 * piece; null: the lane's piece is not stored).
__device__ __forceinline__ void epi_flush4(uint32_t buf, int lane, char* base, int off) {
    __syncwarp();
    uint4 x[4];
    int o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[k] = lds128(buf + epi_slot(k * 8 + (lane >> 2), lane & 3));
        o[k] = __shfl_sync(0xffffffffu, off, k * 8 + (lane >> 2));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base && o[k] >= 0) *reinterpret_cast<uint4*>(base + (size_t)o[k] * 16) = x[k];
    __syncwarp();
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <bool TR, int TW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPhThreads, 1)
rc_tc_phase_pair256_kernel(const RcPhDesc* __restrict__ D, int* __restrict__ ctl, int MT, long long* __restrict__ trace_arg, int mix) {
    long long* const trace = TR ? trace_arg : nullptr;            // the trace stamps compile away in the production instance
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kP2Stages];
    __shared__ __align__(8) uint64_t bar_empty[kP2Stages];
    __shared__ __align__(8) uint64_t bar_corr_full, bar_main_full;    // tcgen05.commit, multicast to both CTAs
    __shared__ __align__(8) uint64_t bar_corr_free, bar_main_free;    // leader: the epilogue warps of both CTAs
    __shared__ __align__(8) uint64_t tq_full[kPhQ];
    __shared__ __align__(8) uint64_t tq_empty[kPhQ];
    __shared__ __align__(16) int4 tq_tile[kPhQ];
    __shared__ int s_mb[RC_PH_MAXJOBS], s_nt[RC_PH_MAXJOBS];   // per job: row-block pairs, column tiles per row block
    __shared__ int lvl_tile0[kP2Levels + 1], lvl_job0[kP2Levels + 1];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) uint4 epi_buf[kPhEpiWarps][128];         // 2 KB per epilogue warp

    rc_pdl_wait();
    rc_pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    unsigned long long t_entry = 0;
    if (trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
    const bool leader = rank == 0;
    const int njobs = D->njobs;

    __shared__ int s_lv[RC_PH_MAXJOBS], s_n2[RC_PH_MAXJOBS], s_nm[RC_PH_MAXJOBS];
    if (warp == 0 && lane < njobs) {                           // one lane per job: the dependent global loads in parallel
        const RcPhJob& Jl = D->job[lane];
        s_mb[lane] = (*Jl.count + 2 * kTcBM - 1) / (2 * kTcBM);
        s_lv[lane] = Jl.level; s_n2[lane] = Jl.nt2; s_nm[lane] = Jl.nmma;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // tile plan (identical in every CTA): jobs are stored level by level
        int lv = 0, total = 0;
        lvl_job0[0] = 0;
        for (int j = 0; j <= njobs; ++j) {
            const int l = (j < njobs) ? s_lv[j] : kP2Levels;
            while (lv < l) {                                        // close level lv: jobs lvl_job0[lv] .. j - 1
                lvl_tile0[lv] = total;
                for (int q = lvl_job0[lv]; q < j; ++q) {
                    // every job in tiles of TW columns — the narrow ones (linear2) too: a 144-column job takes two 128-column tiles
                    s_nt[q] = s_nm[q] == 256 ? s_n2[q] * (256 / TW) : (s_nm[q] + TW - 1) / TW;
                    total += s_mb[q] * s_nt[q];
                }
                lvl_job0[++lv] = j;
            }
        }
        lvl_tile0[kP2Levels] = total;
        for (int s = 0; s < kP2Stages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_corr_full), 1);
        mbar_init(smem_u32(&bar_main_full), 1);
        mbar_init(smem_u32(&bar_corr_free), 2 * kPhEpiWarps);
        mbar_init(smem_u32(&bar_main_free), 2 * kPhEpiWarps);
        for (int q = 0; q < kPhQ; ++q) {
            mbar_init(smem_u32(&tq_full[q]), 1);
            mbar_init(smem_u32(&tq_empty[q]), 2 + 2 * kPhEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int total = lvl_tile0[kP2Levels];
    long long* const tg = trace ? trace + (size_t)total * 16 : nullptr;
    if ((int)(blockIdx.x >> 1) >= total) return;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
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
                    if (t < total) {                                // level -> row-block pair -> job -> column tile
                        int lv = 0;
                        while (t >= lvl_tile0[lv + 1]) ++lv;
                        int lt = t - lvl_tile0[lv];
                        const int j0 = lvl_job0[lv], j1 = lvl_job0[lv + 1];
                        if (mix) {                                  // row-block pair by row-block pair over all jobs of the level
                            for (m = 0;; ++m) {
                                int rowtiles = 0;
                                for (int c = j0; c < j1; ++c) rowtiles += (s_mb[c] > m) ? s_nt[c] : 0;
                                if (lt < rowtiles) break;
                                lt -= rowtiles;
                            }
                            for (j = j0; j < j1; ++j) {
                                const int c = (s_mb[j] > m) ? s_nt[j] : 0;
                                if (lt < c) break;
                                lt -= c;
                            }
                        } else {                                    // job by job (the chain with the longest K first: short tiles fill the tail)
                            for (j = j0; j < j1; ++j) {
                                const int c = s_mb[j] * s_nt[j];
                                if (lt < c) break;
                                lt -= c;
                            }
                            m = lt / s_nt[j];
                            lt -= m * s_nt[j];
                        }
                        n = lt;
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
                const int mb = 2 * tl.y + (int)rank;
                const RcPhJob& J = D->job[j];
                if (trace && leader) {
                    trace[(size_t)t * 16 + 0] = ((long long)blockIdx.x << 32) | (unsigned)((j << 16) | (tl.y << 8) | n); trace[(size_t)t * 16 + 1] = clock64();
                    if (q == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMax((unsigned long long*)tg + 1, ~g); }
                }
                const int nmma = J.nmma == 256 ? TW : min(TW, J.nmma - n * TW);  // MMA N of this tile
                const int wrow = n * TW + (int)rank * (nmma >> 1);              // this CTA's half of the tile's W rows (= output columns)
                const bool w64 = TW == 128 && J.nmma == 256;                    // 64-row W boxes (8 KB) instead of 128-row ones
                const uint32_t stage_tx = 2u * (2 * kPhABytes + (w64 ? kPhWBytes : 2 * kPhWBytes));
                const CUtensorMap* const mWh = (const CUtensorMap*)(w64 ? &J.mWhi64 : &J.mWhi);
                const CUtensorMap* const mWl = (const CUtensorMap*)(w64 ? &J.mWlo64 : &J.mWlo);
                const int KB = J.K / kTcBK;
                const int KD = (J.kind == 1) ? KB / 2 : KB;
                for (int i = 0; i < KB; ++i, ++it) {
                    if (i == KB - KD) {
                        if (J.dep >= 0) {
                            const int need = s_nt[J.dep];
                            const int* flag = ctl + 1 + J.dep * MT + mb;
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
                        if (trace && leader) trace[(size_t)t * 16 + 2] = clock64();
                    }
                    const int kb = (i < KB - KD) ? KD + i : i - (KB - KD);
                    const int s = it % kP2Stages;
                    const uint32_t ph = (it / kP2Stages) & 1u;
                    mbar_wait_cluster(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[s]), 0);
                    if (leader) mbar_expect_tx(smem_u32(&bar_full[s]), stage_tx);
                    const uint32_t base = smem_u32(smem + (size_t)s * kP2StageBytes);
                    tma_load_2d_pair(base, (const CUtensorMap*)&J.mAhi, kb * kTcBK, mb * kTcBM, full_leader);
                    tma_load_2d_pair(base + kPhABytes, (const CUtensorMap*)&J.mAlo, kb * kTcBK, mb * kTcBM, full_leader);
                    tma_load_2d_pair(base + 2 * kPhABytes, mWh, kb * kTcBK, wrow, full_leader);
                    tma_load_2d_pair(base + 2 * kPhABytes + kPhWBytes, mWl, kb * kTcBK, wrow, full_leader);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t d_main = tmem_base, d_corr = tmem_base + 256u;
            uint32_t it = 0, seg = 0;                                           // seg: `main` segments started so far
            for (int q = 0;; ++q) {
                const int slot = q % kPhQ;
                long long w0 = trace ? clock64() : 0, w_tq = 0, w_free = 0, w_full = 0;
                mbar_wait(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
                const int4 t = tq_tile[slot];
                mbar_arrive(smem_u32(&tq_empty[slot]));
                if (t.x < 0) break;
                if (trace) { const long long c = clock64(); w_tq = c - w0; }
                const RcPhJob& J = D->job[t.x];
                const int KB = J.K / kTcBK;
                const int half = (J.kind == 1 && KB >= kP2FlushKB) ? KB / 2 : KB;   // length of a `main` segment in K blocks
                const int nmma = J.nmma == 256 ? TW : min(TW, J.nmma - t.z * TW);
                const uint32_t idesc = (1u << 4) | ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)((2 * kTcBM) >> 4) << 24);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kP2Stages;
                    const uint32_t ph = (it / kP2Stages) & 1u;
                    if (trace) w0 = clock64();
                    mbar_wait_cluster(smem_u32(&bar_full[s]), ph);
                    if (trace) w_full += clock64() - w0;
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + (size_t)s * kP2StageBytes);
                    const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + kPhABytes);
                    const uint64_t dWhi = make_desc(base + 2 * kPhABytes), dWlo = make_desc(base + 2 * kPhABytes + kPhWBytes);
                    if (kb == 0) {                                              // the previous tile's corr has been drained
                        if (trace) w0 = clock64();
                        mbar_wait_cluster(smem_u32(&bar_corr_free), ((uint32_t)q & 1u) ^ 1u);
                        if (trace) { w_free += clock64() - w0; trace[(size_t)t.w * 16 + 3] = clock64(); }
                        tc_fence_after();
                    }
                    const bool seg_first = kb == 0 || kb == half;
                    if (TW == 256) {
                        // N = 256: an MMA takes as long as the accumulate latency, so the order is free: corr first, main last (the epilogue
                        // drains corr during the tile's last main MMAs and main during the next tile's first corr MMAs)
#pragma unroll
                        for (int k = 0; k < kTcBK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            tc_mma_f16_pair(d_corr, dAhi + adv, dWlo + adv, idesc, (kb | k) ? 1u : 0u);
                            tc_mma_f16_pair(d_corr, dAlo + adv, dWhi + adv, idesc, 1u);
                        }
                        if (kb == KB - 1) tc_commit_pair(smem_u32(&bar_corr_full));
                        if (seg_first) {                                        // the previous segment of main is in the epilogue's registers
                            if (trace) w0 = clock64();
                            mbar_wait_cluster(smem_u32(&bar_main_free), (seg & 1u) ^ 1u);
                            if (trace) w_free += clock64() - w0;
                            tc_fence_after();
                            ++seg;
                        }
#pragma unroll
                        for (int k = 0; k < kTcBK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            tc_mma_f16_pair(d_main, dAhi + adv, dWhi + adv, idesc, (seg_first && k == 0) ? 0u : 1u);
                        }
                    } else {
                        // N = 128: an MMA (64 clk) is shorter than the accumulate latency (measured: back-to-back MMAs on ONE accumulator issue
                        // every ~126 clk), so the three products of a K step go to three accumulators — corr is kept as two partial sums
                        // (A-hi x W-lo in columns 256.., A-lo x W-hi in columns 384..) which the epilogue adds.
                        if (seg_first) {
                            if (trace) w0 = clock64();
                            mbar_wait_cluster(smem_u32(&bar_main_free), (seg & 1u) ^ 1u);
                            if (trace) w_free += clock64() - w0;
                            tc_fence_after();
                            ++seg;
                        }
#pragma unroll
                        for (int k = 0; k < kTcBK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            tc_mma_f16_pair(d_corr, dAhi + adv, dWlo + adv, idesc, (kb | k) ? 1u : 0u);
                            tc_mma_f16_pair(d_main, dAhi + adv, dWhi + adv, idesc, (seg_first && k == 0) ? 0u : 1u);
                            tc_mma_f16_pair(d_corr + 128u, dAlo + adv, dWhi + adv, idesc, (kb | k) ? 1u : 0u);
                        }
                        if (kb == KB - 1) tc_commit_pair(smem_u32(&bar_corr_full));
                    }
                    tc_commit_pair(smem_u32(&bar_empty[s]));
                    if (kb == KB - 1 || kb == half - 1) tc_commit_pair(smem_u32(&bar_main_full));
                }
                if (trace) { trace[(size_t)t.w * 16 + 4] = clock64(); trace[(size_t)t.w * 16 + 9] = w_tq; trace[(size_t)t.w * 16 + 10] = w_free; trace[(size_t)t.w * 16 + 11] = w_full; }
            }
        }
    } else {
        const int ewarp = warp - 2;
        const int q4 = warp & 3, part = ewarp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const bool tr_on = trace && leader && ewarp == 0 && lane == 0;
        const uint32_t wbuf = smem_u32(&epi_buf[ewarp][0]);
        uint32_t seg = 0;
        for (int q = 0;; ++q) {
            const int slot = q % kPhQ;
            mbar_wait_cluster(smem_u32(&tq_full[slot]), ((uint32_t)q / kPhQ) & 1u);
            const int4 t = tq_tile[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tq_empty[slot]), 0));
            if (t.x < 0) break;
            const RcPhJob& J = D->job[t.x];
            const bool lstm = J.kind == 1;
            constexpr int tw = TW;
            const int mb = 2 * t.y + (int)rank;
            const int n0 = t.z * tw;
            const int mrow = mb * kTcBM + q4 * 32 + lane;
            const int row = (mrow < *J.count) ? J.rows[mrow] : -1;
            const int ncol = lstm ? tw : min(J.N - n0, tw);                     // valid columns of this tile (warp-uniform)
            const bool flush = lstm && J.K >= kP2FlushKB * kTcBK;
            // A 256-column tile gives every epilogue warp 64 columns (16 hidden units) of its 32 rows, a 128-column tile 32 columns: the
            // accumulators of the former fill two thirds of the 96 registers a thread of this 18-warp block can have.
            auto run_tile = [&](auto cpt_tag) {
                constexpr int CPT = decltype(cpt_tag)::value;
                constexpr int NCH = CPT / 32;                                   // 32-column chunks (8 hidden units each)
                const int c_lo = part * CPT;
                const bool active = c_lo < ncol;
                const int u0 = (n0 + c_lo) >> 2;
                if (lstm && active) {
                    const int off = row >= 0 ? (int)(((size_t)row * J.H + u0) >> 2) : -1;   // this row's segment in C / Hout, in 16-byte units
                    const char* const Cst = (const char*)J.C;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                               // previous cell state -> staging buffer (pieces 0 .. CPT / 16 - 1 of the row)
                        const int o = __shfl_sync(0xffffffffu, off, k * 8 + (lane >> 2));
                        if (o >= 0 && (lane & 3) < CPT / 16) cp_async16(wbuf + epi_slot(k * 8 + (lane >> 2), lane & 3), Cst + (size_t)o * 16 + (lane & 3) * 16);
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                float acc[CPT];
                if (flush) {                                                    // first half of K: main -> registers, then the second half restarts the chain
                    mbar_wait_cluster(smem_u32(&bar_main_full), seg & 1u);
                    ++seg;
                    tc_fence_after();
                    if (active) {
#pragma unroll
                        for (int cc = 0; cc < CPT / 8; ++cc) {
                            uint32_t v0[8];
                            tc_ld8(lane_base + (uint32_t)(c_lo + cc * 8), v0);
                            tc_ld_wait();
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[cc * 8 + e] = __uint_as_float(v0[e]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_main_free), 0));
                }
                mbar_wait_cluster(smem_u32(&bar_corr_full), (uint32_t)q & 1u);
                tc_fence_after();
                if (tr_on) trace[(size_t)t.w * 16 + 5] = clock64();
#pragma unroll
                for (int cc = 0; cc < CPT / 8; ++cc) {
                    const int c0 = c_lo + cc * 8;
                    if (c0 >= ncol) continue;
                    uint32_t v1[8];
                    tc_ld8(lane_base + 256u + (uint32_t)c0, v1);
                    if (TW == 128) {                                            // corr = the two partial sums
                        uint32_t v2[8];
                        tc_ld8(lane_base + 384u + (uint32_t)c0, v2);
                        tc_ld_wait();
#pragma unroll
                        for (int e = 0; e < 8; ++e) v1[e] = __float_as_uint(__uint_as_float(v1[e]) + __uint_as_float(v2[e]));
                    } else {
                        tc_ld_wait();
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = __uint_as_float(v1[e]);
                        acc[cc * 8 + e] = flush ? fmaf(x, 4.8828125e-4f, acc[cc * 8 + e]) : x * 4.8828125e-4f;   // main (first half) + corr * 2^-11
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_corr_free), 0));
                if (tr_on) trace[(size_t)t.w * 16 + 12] = clock64();
                mbar_wait_cluster(smem_u32(&bar_main_full), seg & 1u);
                ++seg;
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < CPT / 8; ++cc) {
                    const int c0 = c_lo + cc * 8;
                    if (c0 >= ncol) continue;
                    uint32_t v0[8];
                    tc_ld8(lane_base + (uint32_t)c0, v0);
                    tc_ld_wait();
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[cc * 8 + e] += __uint_as_float(v0[e]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_main_free), 0));
                if (tr_on) trace[(size_t)t.w * 16 + 8] = clock64();
                // gate math and stores: the tensor memory already belongs to the next tile
                if (lstm && active) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    if (tr_on) trace[(size_t)t.w * 16 + 13] = clock64();
                    float hn[CPT / 4];
                    const float* const bias = J.bias + n0 + c_lo;
                    char* const pC = (char*)J.C;                                // the job's fields in one batch of loads (every shared-memory asm below is a memory clobber)
                    char* const pH = (char*)J.Hout;
                    char* const pAhi = (char*)J.nAhi;
                    char* const pAlo = (char*)J.nAlo;
                    const int npitch = J.npitch, Hj = J.H;
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {                          // 8 units: the previous cell state out of the lane's own slots, the new one back into them
                        const uint32_t sa = wbuf + epi_slot(lane, 2 * cc), sb = wbuf + epi_slot(lane, 2 * cc + 1);
                        const uint4 ca = lds128(sa), cb = lds128(sb);
                        const float cp[8] = {__uint_as_float(ca.x), __uint_as_float(ca.y), __uint_as_float(ca.z), __uint_as_float(ca.w),
                                             __uint_as_float(cb.x), __uint_as_float(cb.y), __uint_as_float(cb.z), __uint_as_float(cb.w)};
                        const float* a = acc + cc * 32;
                        float cn[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + u * 4));
                            cn[u] = fmaf(sigm(a[u * 4 + 1] + b.y), cp[u], sigm(a[u * 4 + 0] + b.x) * gate_tanh(a[u * 4 + 2] + b.z));
                            hn[cc * 8 + u] = sigm(a[u * 4 + 3] + b.w) * gate_tanh(cn[u]);
                        }
                        sts128(sa, f4u(cn[0], cn[1], cn[2], cn[3]));
                        sts128(sb, f4u(cn[4], cn[5], cn[6], cn[7]));
                    }
                    if (tr_on) trace[(size_t)t.w * 16 + 14] = clock64();
                    const int off = row >= 0 ? (int)(((size_t)row * Hj + u0) >> 2) : -1;
                    const int offs = row >= 0 ? (int)(((size_t)mrow * npitch + u0) >> 3) : -1;   // the row's segment in the next operand (16-byte units)
                    if (CPT == 64) {
                        epi_flush4(wbuf, lane, pC + (lane & 3) * 16, off);
#pragma unroll
                        for (int p = 0; p < 4; ++p) sts128(wbuf + epi_slot(lane, p), f4u(hn[4 * p], hn[4 * p + 1], hn[4 * p + 2], hn[4 * p + 3]));
                        epi_flush4(wbuf, lane, pH + (lane & 3) * 16, off);
                        if (tr_on) trace[(size_t)t.w * 16 + 15] = clock64();
                        if (pAhi) {                                           // split operand of the next layer: hi (32 bytes per row) and lo in one pass
                            sts128(wbuf + epi_slot(lane, 0), epi_split8<false>(hn));
                            sts128(wbuf + epi_slot(lane, 1), epi_split8<false>(hn + 8 % (CPT / 4)));
                            sts128(wbuf + epi_slot(lane, 2), epi_split8<true>(hn));
                            sts128(wbuf + epi_slot(lane, 3), epi_split8<true>(hn + 8 % (CPT / 4)));
                            epi_flush4(wbuf, lane, ((lane & 2) ? pAlo : pAhi) + (lane & 1) * 16, offs);
                        }
                    } else {                                                    // 8 units: [c | h] of a row side by side in the buffer, one pass
                        sts128(wbuf + epi_slot(lane, 2), f4u(hn[0], hn[1], hn[2], hn[3]));
                        sts128(wbuf + epi_slot(lane, 3), f4u(hn[4], hn[5], hn[6], hn[7]));
                        epi_flush4(wbuf, lane, ((lane & 2) ? pH : pC) + (lane & 1) * 16, off);
                        if (tr_on) trace[(size_t)t.w * 16 + 15] = clock64();
                        if (pAhi) {
                            sts128(wbuf + epi_slot(lane, 0), epi_split8<false>(hn));
                            sts128(wbuf + epi_slot(lane, 1), epi_split8<true>(hn));
                            epi_flush4(wbuf, lane, (lane & 2) ? nullptr : ((lane & 1) ? pAlo : pAhi), offs);
                        }
                    }
                } else if (!lstm && active && J.nAhi && !J.Y && ncol - c_lo >= CPT) {   // linear1: bias + relu, only the split operand of the LSTM is written
                    const int offs = row >= 0 ? (int)(((size_t)mrow * J.npitch + n0 + c_lo) >> 3) : -1;
                    const int relu = J.relu;
                    const float* const bias = J.bias + n0 + c_lo;
#pragma unroll
                    for (int e = 0; e < CPT; e += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + e));
                        acc[e] += b.x; acc[e + 1] += b.y; acc[e + 2] += b.z; acc[e + 3] += b.w;
                        if (relu) { acc[e] = fmaxf(acc[e], 0.f); acc[e + 1] = fmaxf(acc[e + 1], 0.f); acc[e + 2] = fmaxf(acc[e + 2], 0.f); acc[e + 3] = fmaxf(acc[e + 3], 0.f); }
                    }
#pragma unroll
                    for (int h = 0; h < 2 * NCH; ++h) {                         // (32-column chunk, plane): 64 bytes per row at a time
                        const int cc = h >> 1;
#pragma unroll
                        for (int p = 0; p < 4; ++p) sts128(wbuf + epi_slot(lane, p), (h & 1) ? epi_split8<true>(acc + cc * 32 + p * 8) : epi_split8<false>(acc + cc * 32 + p * 8));
                        epi_flush4(wbuf, lane, (char*)((h & 1) ? J.nAlo : J.nAhi) + cc * 64 + (lane & 3) * 16, offs);
                    }
                } else if (!lstm) {
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        const int nb = n0 + c_lo + cc * 32;
                        if (nb - n0 >= ncol || row < 0) continue;
                        float* a = acc + cc * 32;
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
            };
            run_tile(std::integral_constant<int, TW / 4>());
            if (tr_on) trace[(size_t)t.w * 16 + 6] = clock64();
            epi_bar_sync();
            if (ewarp == 0 && lane == 0) {
                fence_proxy_async_all();
                __threadfence();
                red_release_gpu_add(ctl + 1 + t.x * MT + mb, 1);
                if (tr_on) { trace[(size_t)t.w * 16 + 7] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMax((unsigned long long*)tg + 2, g); }
            }
        }
    }
    if (tg && threadIdx.x == 0) {
        unsigned long long t_exit;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
        atomicMax((unsigned long long*)tg + 0, ~t_entry);
        atomicMax((unsigned long long*)tg + 3, t_exit);
    }
    tc_fence_before();
    cluster_sync_all();
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

int rc_tc_phase(const RcPhDesc* d_desc, int* d_ctl, int MT, int max_tiles, void* stream, long long* d_trace, int tile_width_hint, int reserve_sms) {
    static int pair = -1;
    // Default: the single-CTA kernel (128-row tiles).  RC_PH_PAIR=1 selects the CTA-pair kernel (cta_group::2, 256-row tiles): 25 % less
    // L2 -> SM traffic, same main-loop rate (both sit at the shared-memory port), but the row lists of a frame (~768 / ~256 streams) end
    // in a mostly empty 256-row tile more often, so it is ~2 % slower on the mixed-confidence workload (same-box A/B: 544 vs 535 us).
    if (pair < 0) { const char* e = getenv("RC_PH_PAIR"); pair = e ? atoi(e) : 0; }
    if (pair == 2) {
        // tile width of the wide jobs: 256 columns (best main loop), or 128 (every epilogue warp active on 32 columns: half the latency per
        // tile and twice the tiles — for phases with few rows).  RC_PH_TW = 128 / 256 forces one, default: by the phase's row bound.
        static int tw_env = -1;
        if (tw_env < 0) { const char* e = getenv("RC_PH_TW"); tw_env = e ? atoi(e) : 0; }
        const int tw = tw_env ? tw_env : tile_width_hint;
        static bool attr_set = false;
        if (!attr_set) {
            RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_pair256_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Smem));
            RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_pair256_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Smem));
            RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_pair256_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Smem));
            RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_pair256_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Smem));
            attr_set = true;
        }
        const int pairs = std::max(1, std::min((sm_count() - reserve_sms) / 2, max_tiles));
        static int mix = -1;
        if (mix < 0) { const char* e = getenv("RC_PH_MIX"); mix = e ? atoi(e) : 0; }   // tile order inside a layer level (A/B switch)
        if (tw == 128) {
            if (d_trace) RC_LAUNCH_PDL((rc_tc_phase_pair256_kernel<true, 128>), 2 * pairs, kPhThreads, kP2Smem, stream, d_desc, d_ctl, MT, d_trace, mix);
            else RC_LAUNCH_PDL((rc_tc_phase_pair256_kernel<false, 128>), 2 * pairs, kPhThreads, kP2Smem, stream, d_desc, d_ctl, MT, d_trace, mix);
        } else {
            if (d_trace) RC_LAUNCH_PDL((rc_tc_phase_pair256_kernel<true, 256>), 2 * pairs, kPhThreads, kP2Smem, stream, d_desc, d_ctl, MT, d_trace, mix);
            else RC_LAUNCH_PDL((rc_tc_phase_pair256_kernel<false, 256>), 2 * pairs, kPhThreads, kP2Smem, stream, d_desc, d_ctl, MT, d_trace, mix);
        }
        RC_CHECK_LAUNCH();
        return RC_OK;
    }
    if (pair == 1) {
        static bool attr_set = false;
        if (!attr_set) {
            RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
            attr_set = true;
        }
        const int pairs = std::max(1, std::min((sm_count() - reserve_sms) / 2, max_tiles));   // max_tiles bounds the 256-row tiles too
        static int korder = -1;
        if (korder < 0) { const char* e = getenv("RC_PH_KORDER"); korder = e ? atoi(e) : 1; }   // 0: dependency wait before the first K block (A/B switch)
        RC_LAUNCH_PDL(rc_tc_phase_pair_kernel, 2 * pairs, kPhThreads, kPairSmem, stream, d_desc, d_ctl, MT, d_trace, korder);
        RC_CHECK_LAUNCH();
        return RC_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPhSmem));
        RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPhSmem));
        RC_CUDA(cudaFuncSetAttribute(rc_tc_phase_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPhSmem));
        attr_set = true;
    }
    const int grid = std::max(1, std::min(sm_count() - reserve_sms, max_tiles));
    if (pair == 3) RC_LAUNCH(rc_tc_phase_kernel<1>, grid, kPhThreads, kPhSmem, stream, d_desc, d_ctl, MT, d_trace);
    else if (pair == 4) RC_LAUNCH(rc_tc_phase_kernel<2>, grid, kPhThreads, kPhSmem, stream, d_desc, d_ctl, MT, d_trace);
    else RC_LAUNCH(rc_tc_phase_kernel<0>, grid, kPhThreads, kPhSmem, stream, d_desc, d_ctl, MT, d_trace);
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
