// Interface of the tensor-core (tcgen05) LSTM-layer GEMM (gemm_tc.cu) used by fusion.cu.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

#define RC_TC_BN 128                  // output-tile width (gate columns) = 32 hidden units

struct alignas(64) RcTensorMap { unsigned char opaque[128]; };   // == CUtensorMap

// 2-D fp16 [rows, K] row-major tensor map, box = 64 (K) x box_rows, 128-byte swizzle
int rc_tc_make_map(RcTensorMap* out, const void* base, long long rows, int K, int box_rows);
// fp32 -> (hi, lo) fp16 halves with lo pre-scaled by 2^11
void rc_tc_split_host(const float* w, size_t n, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo);
// gather list rows from [X | X2], split to fp16 halves, write dense [*, Kout] (zero beyond K1+K2; Kout multiple of 64)
int rc_tc_split_rows(const float* X, int ldx, const float* X2, int ldx2, int K1, int K2, int Kout, const int* rows, const int* count,
                     int B, void* Ahi, void* Alo, void* stream);
// plain linear layer Y = act(A W^T + b) on the tensor cores (W rows padded to a multiple of RC_TC_BN, K to 64)
// (nAhi, nAlo, npitch): when not null the outputs are ALSO written, split into fp16 halves, at [compact row * npitch + column] —
// i.e. directly as rows of the next GEMM's A operand, so no separate split pass is needed; Y may then be null.
int rc_tc_linear(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo,
                 const float* bias, float* Y, int ldy, int N, int K, int relu, const int* rows, const int* count, int B, void* stream,
                 void* nAhi = nullptr, void* nAlo = nullptr, int npitch = 0);
// one gather + split launch for up to 3 operand segments of a sub-net pass
struct RcSplitSeg { const float* src; int ld; int K; int Kout; int col0; int pitch; void* hi; void* lo; };
int rc_tc_split_pass(const RcSplitSeg* segs, int nseg, const int* rows, const int* count, int B, void* stream);
// same on 2 x 2 thread-block clusters with TMA multicast; the four tensor maps must have 64-row boxes
int rc_tc_lstm_layer_cluster(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo,
                             const float* bias, float* C, float* Hout, int H, const int* rows, const int* count, int B, void* stream);
// fused LSTM layer on the tensor cores over the rows of a list
int rc_tc_lstm_layer(const RcTensorMap* mAhi, const RcTensorMap* mAlo, const RcTensorMap* mWhi, const RcTensorMap* mWlo,
                     const float* bias, float* C, float* Hout, int H, const int* rows, const int* count, int B, void* stream,
                     void* nAhi = nullptr, void* nAlo = nullptr, int npitch = 0);
