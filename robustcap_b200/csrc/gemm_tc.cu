// Fused LSTM-layer GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
//   gates[rows, 4H] = [x | h_prev][rows, 2H] @ Wcat[4H, 2H]^T        (articulate/utils/torch/rnn.py:111 via sig_mp.py:128)
//
// The parity bar (1e-4 rad) rules out plain TF32/BF16/FP16 operands (weight rounding alone costs 1e-2 rad),
// so every fp32 operand is split into two fp16 halves with an exact power-of-two rescale of the low half,
//     x = hi + lo * 2^-11,   hi = fp16(x),   lo = fp16((x - hi) * 2^11)        (|x - hi - lo 2^-11| <= 2^-22 |x|)
// and the product is evaluated with three kind::f16 MMAs into two fp32 TMEM accumulators
//     main += Ahi Whi ;   corr += Ahi Wlo + Alo Whi ;   gates = main + 2^-11 corr
// (fp16 x fp16 products are exact in the fp32 accumulator; the dropped Alo Wlo term is 2^-22 relative).  This costs 3 MMAs
// at the fp16 rate — half the tensor time of the usual 3xTF32 scheme — and moves 4 bytes per operand element, the same as
// fp32.  Weights are split once on the host; activation rows are gathered through the row list and split by a small
// pre-pass (rc_split_rows_kernel), so the GEMM sees dense K-major operands that TMA can tile with the 128-byte swizzle.
//
// Kernel: one CTA per 128 x BN output tile, 6 warps: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected thread), warps 2-5 = epilogue (tcgen05.ld -> bias + LSTM cell update -> c (in place), h_new).  BK = 64 fp16
// (= one 128B swizzle atom), mbarrier full/empty ring between TMA and MMA, tcgen05.commit frees stages and signals the
// epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "rc_common.cuh"
#include "rc_tc.cuh"
#include "rc_tc_dev.cuh"

namespace {

struct TcArgs {
    const float* bias;     // LSTM: [4H] gate-interleaved (b_ih + b_hh); linear: [N]
    float* C;              // LSTM: [*, H] cell state, in place
    float* Hout;           // LSTM: [*, H] new hidden state
    float* Y; int ldy;     // linear: output rows
    int N, relu;           // linear: valid outputs, activation
    const int* rows;       // row list (stream indices); compact row i of the A operand belongs to stream rows[i]
    const int* count;
    int H, K;
    __half* nAhi; __half* nAlo; int npitch;   // optional: the outputs, split into fp16 halves, written as rows of the NEXT GEMM's A operand
    int dbg;               // timing experiments only: 1 = skip the MMAs, 2 = skip the TMA loads (results are garbage)
};

// Epilogue shared by the tcgen05 kernels: 8 warps (2 per TMEM lane quarter, each taking half of the BN columns).  Thread =
// one output row.  The cell state of the row is prefetched with 128-bit loads BEFORE the accumulators are waited for (the
// epilogue warps idle during the main loop), and c / h leave as 128-bit stores: a row's 8 units are one full 32-byte sector,
// instead of 32 scattered 4-byte accesses per warp instruction with the dependent-load latency exposed per unit (measured:
// the old per-unit load-compute-store loop made the kernel time insensitive to removing every MMA or every TMA load).
template <int BN, bool LSTM>
__device__ __forceinline__ void tc_epilogue(const TcArgs& a, uint32_t tmem_base, uint32_t bar_acc_addr, int m0, int n0, int cnt,
                                            int ewarp, int lane, int nmain) {
    constexpr int CH = BN / 32;                 // 32-column chunks in the tile
    constexpr int CPW = CH / 2;                 // chunks per warp
    const int q = (ewarp + 2) & 3, half = ewarp >> 2;   // a warp may only read TMEM lanes 32 * (warp_id % 4) ..; ewarp = warp_id - 2
    const int mrow = m0 + q * 32 + lane;
    const int row = (mrow < cnt) ? a.rows[mrow] : -1;
    float4 cprev[CPW][2];
    if (LSTM && row >= 0) {
#pragma unroll
        for (int cc = 0; cc < CPW; ++cc) {
            const float* cp = a.C + (size_t)row * a.H + ((n0 + (half * CPW + cc) * 32) >> 2);
            cprev[cc][0] = *reinterpret_cast<const float4*>(cp);
            cprev[cc][1] = *reinterpret_cast<const float4*>(cp + 4);
        }
    }
    mbar_wait(bar_acc_addr, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int cc = 0; cc < CPW; ++cc) {
        const int c = half * CPW + cc;
        uint32_t v0[32], v1[32];
        float acc[32];
        tc_ld32(lane_base + (uint32_t)(c * 32), v0);
        tc_ld32(lane_base + (uint32_t)(3 * BN + c * 32), v1);
        tc_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(v0[e]);
        if (nmain > 1) {
            tc_ld32(lane_base + (uint32_t)(BN + c * 32), v0);
            tc_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] += __uint_as_float(v0[e]);
        }
        if (nmain > 2) {
            tc_ld32(lane_base + (uint32_t)(2 * BN + c * 32), v0);
            tc_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] += __uint_as_float(v0[e]);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = fmaf(__uint_as_float(v1[e]), 4.8828125e-4f, acc[e]);    // + corr * 2^-11
        if (row < 0) continue;
        const int nb = n0 + c * 32;
        if (LSTM) {
            const float cp[8] = {cprev[cc][0].x, cprev[cc][0].y, cprev[cc][0].z, cprev[cc][0].w,
                                 cprev[cc][1].x, cprev[cc][1].y, cprev[cc][1].z, cprev[cc][1].w};
            float cn[8], hn[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + nb + u * 4));
                cn[u] = fmaf(sigm(acc[u * 4 + 1] + b.y), cp[u], sigm(acc[u * 4 + 0] + b.x) * tanhf(acc[u * 4 + 2] + b.z));
                hn[u] = sigm(acc[u * 4 + 3] + b.w) * tanhf(cn[u]);
            }
            const size_t idx = (size_t)row * a.H + (nb >> 2);
            *reinterpret_cast<float4*>(a.C + idx) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(a.C + idx + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
            *reinterpret_cast<float4*>(a.Hout + idx) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4*>(a.Hout + idx + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
            if (a.nAhi) tc_store_split<8>(hn, a.nAhi, a.nAlo, (size_t)mrow * a.npitch + (nb >> 2));
        } else {
            float* yrow = a.Y + (size_t)row * a.ldy;
            const bool vec = ((a.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) && (nb + 32 <= a.N);   // Y may be null when only the split copy is wanted
            if (vec) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + nb + e));
                    float4 y = make_float4(acc[e] + b.x, acc[e + 1] + b.y, acc[e + 2] + b.z, acc[e + 3] + b.w);
                    if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    if (a.Y) *reinterpret_cast<float4*>(yrow + nb + e) = y;
                    acc[e] = y.x; acc[e + 1] = y.y; acc[e + 2] = y.z; acc[e + 3] = y.w;
                }
                if (a.nAhi) {
#pragma unroll
                    for (int e = 0; e < 32; e += 8) tc_store_split<8>(acc + e, a.nAhi, a.nAlo, (size_t)mrow * a.npitch + nb + e);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int n = nb + e;
                    if (n < a.N) {
                        float y = acc[e] + a.bias[n];
                        if (a.relu) y = fmaxf(y, 0.f);
                        yrow[n] = y;
                    }
                }
            }
        }
    }
}

// TMEM plan (512 columns): three "main" accumulators used round-robin over the K steps + one "corr" accumulator.
// The tensor core adds into the fp32 accumulator with truncation, a drift that grows with the number of MMAs chained on
// one accumulator (measured: 3.6x the rms error of sequential fp32 FMAs at K = 2560); three independent chains cut it 3x
// and are summed in the epilogue with round-to-nearest adds.  corr holds values 2^11 times smaller, its drift is moot.
template <int BN, int STAGES, bool LSTM>
__global__ void __launch_bounds__(kTcThreads, 1)
rc_tc_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
             const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, TcArgs a) {
    constexpr int A_BYTES = kTcBM * kTcBK * 2;     // 16 KB
    constexpr int W_BYTES = BN * kTcBK * 2;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    constexpr int TM_COLS = 4 * BN;                // 3 x main + corr
    static_assert(TM_COLS == 512 || TM_COLS == 256, "TMEM allocation must be a power of two");
    const int cnt = *a.count;
    const int m0 = blockIdx.y * kTcBM;
    if (m0 >= cnt) return;
    const int n0 = blockIdx.x * BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = a.K / kTcBK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                const uint32_t full = smem_u32(&bar_full[s]);
                if (a.dbg == 2) { mbar_expect_tx(full, 0); continue; }
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                tma_load_2d(base, &tmAhi, kb * kTcBK, m0, full);
                tma_load_2d(base + A_BYTES, &tmAlo, kb * kTcBK, m0, full);
                tma_load_2d(base + 2 * A_BYTES, &tmWhi, kb * kTcBK, n0, full);
                tma_load_2d(base + 2 * A_BYTES + W_BYTES, &tmWlo, kb * kTcBK, n0, full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute UMMA::InstrDescriptor): c_format F32 = 1 at [4,6); a/b format F16 = 0; K-major both;
            // n_dim = N >> 3 at [17,23); m_dim = M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
            const uint32_t d_corr = tmem_base + 3 * BN;
            int g = 0;                                           // global K-step counter
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + A_BYTES);
                const uint64_t dWhi = make_desc(base + 2 * A_BYTES), dWlo = make_desc(base + 2 * A_BYTES + W_BYTES);
#pragma unroll
                for (int k = 0; k < kTcBK / 16; ++k, ++g) {
                    if (a.dbg == 1) continue;
                    const uint64_t adv = (uint64_t)(k * 2);      // 16 fp16 = 32 bytes = 2 x 16-byte units
                    const uint32_t d_main = tmem_base + (uint32_t)((g % 3) * BN);
                    tc_mma_f16(d_main, dAhi + adv, dWhi + adv, idesc, g >= 3 ? 1u : 0u);
                    tc_mma_f16(d_corr, dAhi + adv, dWlo + adv, idesc, g ? 1u : 0u);
                    tc_mma_f16(d_corr, dAlo + adv, dWhi + adv, idesc, 1u);
                }
                tc_commit(smem_u32(&bar_empty[s]));              // frees the stage once these MMAs have read it
            }
            tc_commit(smem_u32(&bar_acc));                       // accumulators complete
        }
    } else {
        const int nmain = (KB * (kTcBK / 16) >= 3) ? 3 : KB * (kTcBK / 16);   // accumulators that were written
        tc_epilogue<BN, LSTM>(a, tmem_base, smem_u32(&bar_acc), m0, n0, cnt, warp - 2, lane, nmain);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TM_COLS) : "memory");
    }
}

// 2 x 2 thread-block-cluster variant of the fused LSTM layer: the four CTAs of a cluster cover 2 N-tiles x 2 M-tiles; every
// operand tile is fetched from L2 once per cluster (each CTA loads a 64-row half of its A tile and of its W tile and TMA-
// multicasts it to the CTA that shares that tile), halving the L2->SM operand traffic that bounds the single-CTA kernel.
// Protocol: full[s] expects the whole 64 KB stage (halves arrive from two CTAs); empty[s] counts 3 arrivals — the
// tcgen05.commit of this CTA, of its row peer and of its column peer (multicast commit) — because those are the CTAs whose
// smem this CTA's producer overwrites; cluster-wide barriers fence start and exit.
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 2, 1) __launch_bounds__(kTcThreads, 1)
rc_tc_lstm_cluster_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
             const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, TcArgs a) {
    constexpr int A_BYTES = kTcBM * kTcBK * 2;     // 16 KB
    constexpr int W_BYTES = BN * kTcBK * 2;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    constexpr int TM_COLS = 4 * BN;                // 3 x main + corr
    static_assert(TM_COLS == 512 || TM_COLS == 256, "TMEM allocation must be a power of two");
    constexpr bool LSTM = true;
    static_assert(BN == 128, "half tiles are 64 rows");
    const int cnt = *a.count;
    const int m0 = blockIdx.y * kTcBM;
    if ((int)(blockIdx.y & ~1u) * kTcBM >= cnt) return;          // uniform over the cluster: both M tiles are empty
    const int n0 = blockIdx.x * BN;
    const uint32_t cx = blockIdx.x & 1u, cy = blockIdx.y & 1u;     // position inside the 2 x 2 cluster
    const uint32_t crank = cluster_ctarank();                      // == cx + 2 * cy (x fastest)
    const uint16_t row_mask = (uint16_t)(3u << (cy * 2));          // CTAs sharing this A tile (same M tile)
    const uint16_t col_mask = (uint16_t)(5u << cx);                // CTAs sharing this W tile (same N tile)
    const uint16_t commit_mask = (uint16_t)((1u << crank) | (1u << (crank ^ 1u)) | (1u << (crank ^ 2u)));

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = a.K / kTcBK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 3); }
        mbar_init(smem_u32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                            // every CTA's barriers exist before any multicast targets them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                // 64-row halves (8 KB each): A half `cx` to the row peers, W half `cy` to the column peers
                tma_load_2d_mc(base + cx * (A_BYTES / 2), &tmAhi, kb * kTcBK, m0 + (int)cx * 64, full, row_mask);
                tma_load_2d_mc(base + A_BYTES + cx * (A_BYTES / 2), &tmAlo, kb * kTcBK, m0 + (int)cx * 64, full, row_mask);
                tma_load_2d_mc(base + 2 * A_BYTES + cy * (W_BYTES / 2), &tmWhi, kb * kTcBK, n0 + (int)cy * 64, full, col_mask);
                tma_load_2d_mc(base + 2 * A_BYTES + W_BYTES + cy * (W_BYTES / 2), &tmWlo, kb * kTcBK, n0 + (int)cy * 64, full, col_mask);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute UMMA::InstrDescriptor): c_format F32 = 1 at [4,6); a/b format F16 = 0; K-major both;
            // n_dim = N >> 3 at [17,23); m_dim = M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
            const uint32_t d_corr = tmem_base + 3 * BN;
            int g = 0;                                           // global K-step counter
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint64_t dAhi = make_desc(base), dAlo = make_desc(base + A_BYTES);
                const uint64_t dWhi = make_desc(base + 2 * A_BYTES), dWlo = make_desc(base + 2 * A_BYTES + W_BYTES);
#pragma unroll
                for (int k = 0; k < kTcBK / 16; ++k, ++g) {
                    const uint64_t adv = (uint64_t)(k * 2);      // 16 fp16 = 32 bytes = 2 x 16-byte units
                    const uint32_t d_main = tmem_base + (uint32_t)((g % 3) * BN);
                    tc_mma_f16(d_main, dAhi + adv, dWhi + adv, idesc, g >= 3 ? 1u : 0u);
                    tc_mma_f16(d_corr, dAhi + adv, dWlo + adv, idesc, g ? 1u : 0u);
                    tc_mma_f16(d_corr, dAlo + adv, dWhi + adv, idesc, 1u);
                }
                tc_commit_mc(smem_u32(&bar_empty[s]), commit_mask);  // frees the stage here and in the two CTAs that write into it
            }
            tc_commit(smem_u32(&bar_acc));                       // accumulators complete
        }
    } else {
        const int nmain = (KB * (kTcBK / 16) >= 3) ? 3 : KB * (kTcBK / 16);   // accumulators that were written
        tc_epilogue<BN, LSTM>(a, tmem_base, smem_u32(&bar_acc), m0, n0, cnt, warp - 2, lane, nmain);
    }
    tc_fence_before();
    cluster_sync_all();                                            // no CTA leaves while a peer can still write its smem / barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TM_COLS) : "memory");
    }
}

// Gather the rows of a list from the two K segments, split every fp32 into (hi, lo) fp16 halves, write dense [*, K] rows.
__global__ void __launch_bounds__(256) rc_split_rows_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ X2, int ldx2,
                                                             int K1, int K2, int Kout, const int* __restrict__ rows,
                                                             const int* __restrict__ count, __half* __restrict__ Ahi, __half* __restrict__ Alo) {
    const int cnt = *count;
    const int K = K1 + K2, q4 = Kout >> 2;
    const long long total = (long long)cnt * q4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / q4), k = (int)(e % q4) * 4;
        const int r = rows[i];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K1) v = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + k);
        else if (k < K) v = *reinterpret_cast<const float4*>(X2 + (size_t)r * ldx2 + (k - K1));
        const float x[4] = {v.x, v.y, v.z, v.w};
        __half hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2half_rn(x[j]);
            lo[j] = __float2half_rn((x[j] - __half2float(hi[j])) * 2048.f);
        }
        __half2* ph = reinterpret_cast<__half2*>(Ahi + (size_t)i * Kout + k);
        __half2* pl = reinterpret_cast<__half2*>(Alo + (size_t)i * Kout + k);
        ph[0] = __halves2half2(hi[0], hi[1]); ph[1] = __halves2half2(hi[2], hi[3]);
        pl[0] = __halves2half2(lo[0], lo[1]); pl[1] = __halves2half2(lo[2], lo[3]);
    }
}

// One pre-pass per sub-net pass: every operand that does not come out of a GEMM epilogue is gathered / split here in one launch
// (linear1 input X -> A0, h_prev of layer 0 -> second half of A1's rows, h_prev of layer 1 -> second half of A2's rows).
struct SplitSeg { const float* src; int ld; int K; int Kout; int col0; int pitch; __half* hi; __half* lo; };
struct SplitPassArgs { SplitSeg seg[3]; int nseg; const int* rows; const int* count; };

__global__ void __launch_bounds__(256) rc_split_pass_kernel(SplitPassArgs a) {
    const int cnt = *a.count;
    for (int sidx = 0; sidx < a.nseg; ++sidx) {
        const SplitSeg g = a.seg[sidx];
        const int q4 = g.Kout >> 2;
        const long long total = (long long)cnt * q4;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
            const int i = (int)(e / q4), k = (int)(e % q4) * 4;
            const int r = a.rows[i];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < g.K) v = *reinterpret_cast<const float4*>(g.src + (size_t)r * g.ld + k);
            const float x[4] = {v.x, v.y, v.z, v.w};
            __half hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi[j] = __float2half_rn(x[j]);
                lo[j] = __float2half_rn((x[j] - __half2float(hi[j])) * 2048.f);
            }
            __half2* ph = reinterpret_cast<__half2*>(g.hi + (size_t)i * g.pitch + g.col0 + k);
            __half2* pl = reinterpret_cast<__half2*>(g.lo + (size_t)i * g.pitch + g.col0 + k);
            ph[0] = __halves2half2(hi[0], hi[1]); ph[1] = __halves2half2(hi[2], hi[3]);
            pl[0] = __halves2half2(lo[0], lo[1]); pl[1] = __halves2half2(lo[2], lo[3]);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

int rc_tc_make_map(RcTensorMap* out, const void* base, long long rows, int K, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { rc_set_error("cuTensorMapEncodeTiled entry point not available"); return RC_ERR_CUDA; }
    static_assert(sizeof(RcTensorMap) == sizeof(CUtensorMap), "tensor map size");
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rc_set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%d box=%d", (int)r, rows, K, box_rows); return RC_ERR_CUDA; }
    return RC_OK;
}

// fp32 [n] -> (hi, lo) fp16 halves on the host (weights, once at rc_net_finalize)
void rc_tc_split_host(const float* w, size_t n, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo) {
    hi.resize(n);
    lo.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const __half h = __float2half_rn(w[i]);
        const __half l = __float2half_rn((w[i] - __half2float(h)) * 2048.f);
        hi[i] = *reinterpret_cast<const uint16_t*>(&h);
        lo[i] = *reinterpret_cast<const uint16_t*>(&l);
    }
}

int rc_tc_split_pass(const RcSplitSeg* segs, int nseg, const int* rows, const int* count, int B, void* stream) {
    SplitPassArgs a;
    memset(&a, 0, sizeof(a));
    long long work = 0;
    for (int i = 0; i < nseg && i < 3; ++i) {
        a.seg[i].src = segs[i].src; a.seg[i].ld = segs[i].ld; a.seg[i].K = segs[i].K; a.seg[i].Kout = segs[i].Kout; a.seg[i].col0 = segs[i].col0;
        a.seg[i].pitch = segs[i].pitch; a.seg[i].hi = (__half*)segs[i].hi; a.seg[i].lo = (__half*)segs[i].lo;
        work = std::max(work, (long long)B * (segs[i].Kout / 4));
    }
    a.nseg = nseg; a.rows = rows; a.count = count;
    const int grid = (int)std::min<long long>(rc_cdiv(work, 256), 148 * 8);
    RC_LAUNCH(rc_split_pass_kernel, grid, 256, 0, stream, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

int rc_tc_split_rows(const float* X, int ldx, const float* X2, int ldx2, int K1, int K2, int Kout, const int* rows, const int* count,
                     int B, void* Ahi, void* Alo, void* stream) {
    const long long work = (long long)B * (Kout / 4);
    const int grid = (int)std::min<long long>(rc_cdiv(work, 256), 148 * 8);
    RC_LAUNCH(rc_split_rows_kernel, grid, 256, 0, stream, X, ldx, X2, ldx2, K1, K2, Kout, rows, count, (__half*)Ahi, (__half*)Alo);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

template <int BN, int STAGES, bool LSTM>
int launch_tc(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo, const TcArgs& a,
              int n_total, int B, void* stream) {
    constexpr int SMEM = STAGES * (2 * kTcBM * kTcBK * 2 + 2 * BN * kTcBK * 2) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        RC_CUDA(cudaFuncSetAttribute(rc_tc_kernel<BN, STAGES, LSTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    dim3 grid(rc_cdiv(n_total, BN), rc_cdiv(B, kTcBM));
    RC_LAUNCH((rc_tc_kernel<BN, STAGES, LSTM>), grid, kTcThreads, SMEM, stream, *(const CUtensorMap*)mAhi, *(const CUtensorMap*)mAlo,
              *(const CUtensorMap*)mWhi, *(const CUtensorMap*)mWlo, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

static int tc_stages() {
    static int stages = 0;
    if (!stages) {
        const char* e = getenv("RC_TC_STAGES");       // tuning knob (2 or 3 ring stages of 64 KB)
        stages = (e && atoi(e) == 2) ? 2 : 3;
    }
    return stages;
}

int rc_tc_lstm_layer_cluster(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo,
                             const float* bias, float* C, float* Hout, int H, const int* rows, const int* count, int B, void* stream) {
    // tensor maps here have 64-row boxes (half tiles)
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.bias = bias; a.C = C; a.Hout = Hout; a.rows = rows; a.count = count; a.H = H; a.K = 2 * H;
    constexpr int BN = RC_TC_BN, STAGES = 3;
    constexpr int SMEM = STAGES * (2 * kTcBM * kTcBK * 2 + 2 * BN * kTcBK * 2) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        RC_CUDA(cudaFuncSetAttribute(rc_tc_lstm_cluster_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    const int mt = rc_cdiv(B, kTcBM);
    dim3 grid(4 * H / BN, (mt + 1) / 2 * 2);            // both grid dimensions are multiples of the 2 x 2 cluster
    RC_LAUNCH((rc_tc_lstm_cluster_kernel<BN, STAGES>), grid, kTcThreads, SMEM, stream, *(const CUtensorMap*)mAhi, *(const CUtensorMap*)mAlo,
              *(const CUtensorMap*)mWhi, *(const CUtensorMap*)mWlo, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

int rc_tc_lstm_layer(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo,
                     const float* bias, float* C, float* Hout, int H, const int* rows, const int* count, int B, void* stream,
                     void* nAhi, void* nAlo, int npitch) {
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.bias = bias; a.C = C; a.Hout = Hout; a.rows = rows; a.count = count; a.H = H; a.K = 2 * H;
    a.nAhi = (__half*)nAhi; a.nAlo = (__half*)nAlo; a.npitch = npitch;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("RC_TC_DBG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
    if (tc_stages() == 2) return launch_tc<RC_TC_BN, 2, true>(mAhi, mAlo, mWhi, mWlo, a, 4 * H, B, stream);
    return launch_tc<RC_TC_BN, 3, true>(mAhi, mAlo, mWhi, mWlo, a, 4 * H, B, stream);
}

int rc_tc_linear(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo,
                 const float* bias, float* Y, int ldy, int N, int K, int relu, const int* rows, const int* count, int B, void* stream,
                 void* nAhi, void* nAlo, int npitch) {
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.bias = bias; a.Y = Y; a.ldy = ldy; a.N = N; a.relu = relu; a.rows = rows; a.count = count; a.K = K;
    a.nAhi = (__half*)nAhi; a.nAlo = (__half*)nAlo; a.npitch = npitch;
    return launch_tc<RC_TC_BN, 3, false>(mAhi, mAlo, mWhi, mWlo, a, N, B, stream);
}
