// Dense layers of the six LSTM stacks (articulate/utils/torch/rnn.py:111-114 as driven by net/sig_mp.py:126-129).
//
// One descriptor (RcLinear) covers `linear1+relu`, `linear2`, the init_net layers and the fused LSTM layer
//   gates = [x | h_prev] @ Wcat^T + (b_ih + b_hh);  c' = s(f) c + s(i) tanh(g);  h' = s(o) tanh(c')
// whose weights are packed gate-interleaved ([4H, 2H], row 4j+g = gate g of hidden unit j) so one thread owns the
// four gates of a unit and the cell update is fused into the epilogue — gate pre-activations never touch HBM.
//
// Two kernels:
//   rc_gemv_kernel  B <= 8 streams: weight-streaming GEMV, HBM-bound (every weight byte is read exactly once per
//                   frame; a warp streams 4 weight rows with 128-bit no-allocate loads, K split across warps).
//   rc_gemm_kernel  larger batches: 128x128x16 shared-memory tiled fp32 SIMT GEMM, register double buffering,
//                   8x8 micro-tiles (fp32 FFMA keeps the 1e-4 rad parity bar; a 3xTF32 tcgen05
//                   variant is the planned replacement).
// Rows are addressed through an index list + device-side count, so one launch serves any subset of streams
// (confidence branches of sig_mp.py:149-167, 264) without a host round trip.
#pragma once
#include "rc_common.cuh"

struct RcLinear {
    const float* X;  int ldx;      // first K segment, rows of length ldx
    const float* X2; int ldx2;     // second K segment (LSTM: h_prev) or nullptr
    int K1, K2;                    // segment lengths, multiples of 16
    const float* W;                // [Nw, K1+K2] row-major
    const float* bias;             // [Nw]
    int N, Nw;                     // valid outputs, allocated rows (multiple of 4)
    float* Y; int ldy;             // plain layer output
    float* C;                      // LSTM cell state [*, H] (updated in place)
    float* Hout;                   // LSTM new hidden state [*, H]
    int H;
    const int* rows;               // row list
    const int* count;              // device count of valid list entries
    int relu;
    // optional (GEMV linear2 launch): h <- h_new of both LSTM layers for the rows of the list, so no separate commit launch
    const float* commit_src[2]; float* commit_dst[2]; int commit_H;
};

__device__ __forceinline__ float4 rc_ldg_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float rc_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void rc_lstm_cell(float pi, float pf, float pg, float po, float c_prev, float& c_new, float& h_new) {
    const float ig = rc_sigmoid(pi), fg = rc_sigmoid(pf), gg = tanhf(pg), og = rc_sigmoid(po);
    c_new = fmaf(fg, c_prev, ig * gg);
    h_new = og * tanhf(c_new);
}

// ---------------------------------------------------------------------------------------------------------------
// GEMV path: up to M streams.  Block = 8 warps; `ksplit` warps share one job (= 4 consecutive weight rows = one
// hidden unit for the LSTM layers) by splitting K.
// K-split policy shared by the GEMV launches and the single-stream cooperative kernel (identical arithmetic in both).
__host__ __device__ __forceinline__ int rc_gemv_ksplit(int K) { return K >= 1024 ? 8 : (K >= 512 ? 4 : 1); }

// One "virtual block" (8 warps, `warp` in 0..7) of the GEMV: jobs [vb * jpb, (vb + 1) * jpb).  `part` is 8 x 4M floats of shared
// memory private to the virtual block; `bar` is the named barrier the 8 warps synchronise on (0 = __syncthreads of a 256-thread
// block).  COHERENT selects L2-coherent activation loads (ld.global.cg) for use inside a kernel whose other CTAs produced X.
template <int M, bool LSTM, bool COHERENT>
__device__ __forceinline__ void rc_gemv_vblock(const RcLinear& a, int ksplit, int vb, int warp, int lane, float (*part)[4 * M], int cnt, int bar) {
    const int jpb = 8 / ksplit;
    const int job = vb * jpb + warp / ksplit;
    const int ks = warp % ksplit;
    const int njobs = a.Nw >> 2;
    const int K = a.K1 + a.K2;
    int rowid[M];
#pragma unroll
    for (int m = 0; m < M; ++m) rowid[m] = (m < cnt) ? a.rows[m] : -1;
    float acc[4][M];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
    // the finalising lane fetches its cell state now, so the load overlaps the weight stream instead of following it
    float cprev = 0.f;
    const bool fin = (ks == 0 && job < njobs && lane < cnt);
    if (LSTM && fin) {
        const size_t cidx = (size_t)a.rows[lane] * a.H + job;
        cprev = COHERENT ? __ldcg(a.C + cidx) : a.C[cidx];
    }

    if (job < njobs) {
        const float* w0 = a.W + (size_t)(job * 4) * K;
#pragma unroll 2
        for (int k = (ks * 32 + lane) * 4; k < K; k += ksplit * 128) {
            float4 wv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) wv[r] = rc_ldg_stream(w0 + (size_t)r * K + k);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rowid[m] >= 0) {
                    const float* xp = (k < a.K1) ? (a.X + (size_t)rowid[m] * a.ldx + k)
                                                 : (a.X2 + (size_t)rowid[m] * a.ldx2 + (k - a.K1));
                    xv = COHERENT ? __ldcg(reinterpret_cast<const float4*>(xp)) : __ldg(reinterpret_cast<const float4*>(xp));
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[r][m] = fmaf(wv[r].x, xv.x, acc[r][m]);
                    acc[r][m] = fmaf(wv[r].y, xv.y, acc[r][m]);
                    acc[r][m] = fmaf(wv[r].z, xv.z, acc[r][m]);
                    acc[r][m] = fmaf(wv[r].w, xv.w, acc[r][m]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float v = acc[r][m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[r][m] = v;
        }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int m = 0; m < M; ++m) part[warp][r * M + m] = acc[r][m];
    }
    if (bar == 0) __syncthreads(); else asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");
    if (!(ks != 0 || job >= njobs || lane >= cnt)) {
        const int m = lane;
        const int row = a.rows[m];
        float tot[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = 0.f;
            for (int q = 0; q < ksplit; ++q) s += part[warp + q][r * M + m];
            tot[r] = s;
        }
        if (LSTM) {
            const float4 b = *reinterpret_cast<const float4*>(a.bias + job * 4);
            const size_t idx = (size_t)row * a.H + job;
            float cn, hn;
            rc_lstm_cell(tot[0] + b.x, tot[1] + b.y, tot[2] + b.z, tot[3] + b.w, cprev, cn, hn);
            a.C[idx] = cn;
            a.Hout[idx] = hn;
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int n = job * 4 + r;
                if (n < a.N) {
                    float y = tot[r] + a.bias[n];
                    if (a.relu) y = fmaxf(y, 0.f);
                    a.Y[(size_t)row * a.ldy + n] = y;
                }
            }
        }
    }
}

template <int M, bool LSTM>
__global__ void __launch_bounds__(256) rc_gemv_kernel(RcLinear a, int ksplit) {
    __shared__ float part[8][4 * M];
    const int cnt = min(*a.count, M);
    if (cnt == 0) return;
    rc_gemv_vblock<M, LSTM, false>(a, ksplit, blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, part, cnt, 0);
    if (!LSTM && a.commit_H) {
        const int q4 = a.commit_H >> 2;
        for (int e = blockIdx.x * 256 + threadIdx.x; e < cnt * 2 * q4; e += gridDim.x * 256) {
            const int m = e / (2 * q4), l = (e / q4) & 1, k = (e % q4) * 4;
            const size_t o = (size_t)a.rows[m] * a.commit_H + k;
            *reinterpret_cast<float4*>(a.commit_dst[l] + o) = *reinterpret_cast<const float4*>(a.commit_src[l] + o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tiled GEMM path.
constexpr int kBM = 128, kBN = 128, kBK = 16, kLd = 132;

template <bool LSTM>
__global__ void __launch_bounds__(256, 2) rc_gemm_kernel(RcLinear a) {
    __shared__ __align__(16) float As[2][kBK][kLd];
    __shared__ __align__(16) float Bs[2][kBK][kLd];
    __shared__ int rowid[kBM];
    const int cnt = *a.count;
    const int m0 = blockIdx.y * kBM;
    if (m0 >= cnt) return;
    const int n0 = blockIdx.x * kBN;
    const int tid = threadIdx.x;
    if (tid < kBM) rowid[tid] = (m0 + tid < cnt) ? a.rows[m0 + tid] : -1;
    __syncthreads();
    const int K = a.K1 + a.K2, KT = K / kBK;
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int ra0 = rowid[lr], ra1 = rowid[lr + 64];
    const int nb0 = n0 + lr, nb1 = n0 + lr + 64;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pa0, pa1, pb0, pb1;

    auto gload = [&](int kt) {
        const int k = kt * kBK + lk;
        const float* xb;
        int ld;
        if (k < a.K1) { xb = a.X + k; ld = a.ldx; } else { xb = a.X2 + (k - a.K1); ld = a.ldx2; }
        pa0 = ra0 >= 0 ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)ra0 * ld)) : zero4;
        pa1 = ra1 >= 0 ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)ra1 * ld)) : zero4;
        pb0 = nb0 < a.Nw ? __ldg(reinterpret_cast<const float4*>(a.W + (size_t)nb0 * K + k)) : zero4;
        pb1 = nb1 < a.Nw ? __ldg(reinterpret_cast<const float4*>(a.W + (size_t)nb1 * K + k)) : zero4;
    };
    auto sstore = [&](int buf) {
        As[buf][lk + 0][lr] = pa0.x; As[buf][lk + 1][lr] = pa0.y; As[buf][lk + 2][lr] = pa0.z; As[buf][lk + 3][lr] = pa0.w;
        As[buf][lk + 0][lr + 64] = pa1.x; As[buf][lk + 1][lr + 64] = pa1.y; As[buf][lk + 2][lr + 64] = pa1.z; As[buf][lk + 3][lr + 64] = pa1.w;
        Bs[buf][lk + 0][lr] = pb0.x; Bs[buf][lk + 1][lr] = pb0.y; Bs[buf][lk + 2][lr] = pb0.z; Bs[buf][lk + 3][lr] = pb0.w;
        Bs[buf][lk + 0][lr + 64] = pb1.x; Bs[buf][lk + 1][lr + 64] = pb1.y; Bs[buf][lk + 2][lr + 64] = pb1.z; Bs[buf][lk + 3][lr + 64] = pb1.w;
    };

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < KT) gload(kt + 1);
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < KT) sstore(cur ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ml = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
        const int row = rowid[ml];
        if (row < 0) continue;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int n = n0 + g * 64 + tx * 4;
            if (LSTM) {
                if (n < a.N) {
                    const float4 b = *reinterpret_cast<const float4*>(a.bias + n);
                    const int unit = n >> 2;
                    const size_t idx = (size_t)row * a.H + unit;
                    float cn, hn;
                    rc_lstm_cell(acc[i][g * 4 + 0] + b.x, acc[i][g * 4 + 1] + b.y, acc[i][g * 4 + 2] + b.z,
                                 acc[i][g * 4 + 3] + b.w, a.C[idx], cn, hn);
                    a.C[idx] = cn;
                    a.Hout[idx] = hn;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < a.N) {
                        float y = acc[i][g * 4 + j] + a.bias[n + j];
                        if (a.relu) y = fmaxf(y, 0.f);
                        a.Y[(size_t)row * a.ldy + n + j] = y;
                    }
                }
            }
        }
    }
}

// h <- h_new for the rows of a list, both LSTM layers (the GEMM reads all of h_prev while it writes h_new, so the
// new hidden state is committed by a separate launch).
static __global__ void __launch_bounds__(256) rc_commit_kernel(const float* __restrict__ hn0, const float* __restrict__ hn1,
                                                         float* __restrict__ h0, float* __restrict__ h1, int H,
                                                         const int* __restrict__ rows, const int* __restrict__ count) {
    const int cnt = *count;
    const int q4 = H >> 2;
    const long long total = (long long)cnt * q4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = rows[e / q4];
        const size_t o = (size_t)r * H + (size_t)(e % q4) * 4;
        *reinterpret_cast<float4*>(h0 + o) = *reinterpret_cast<const float4*>(hn0 + o);
        *reinterpret_cast<float4*>(h1 + o) = *reinterpret_cast<const float4*>(hn1 + o);
    }
}
