// Interface of the persistent grouped tcgen05 kernel (phase_tc.cu): ONE launch executes every GEMM of a "phase" of the frame
// (the linear1 -> LSTM-0 -> LSTM-1 -> linear2 chains of up to four sub-nets of net/sig_mp.py:126-129 that have no mutual
// dependencies), tile by tile, with the layer-to-layer dependencies tracked per 128-row block in global memory.
#pragma once
#include "rc_tc.cuh"

constexpr int RC_PH_MAXJOBS = 16;
constexpr int RC_PH_MAXSEGS = 12;

// One GEMM of a phase.  kind 0: Y = act(A W^T + b) (linear1 / linear2); kind 1: fused LSTM layer (gates -> c, h).
// A is the compact-row split-fp16 operand buffer of the chain ([Bpad, K], 128-row TMA boxes), W the split weights
// ([N padded to 128, K]).  (nAhi, nAlo, npitch): when set, the outputs are also written as fp16 halves into the NEXT job's
// A operand.  dep: index of the job (of the same chain) whose 128-row block must be complete before this job may load it.
struct alignas(64) RcPhJob {
    RcTensorMap mAhi, mAlo, mWhi, mWlo;
    RcTensorMap mWhi64, mWlo64;   // the same weights with 64-row boxes (CTA-pair kernel: each CTA loads half of a W tile)
    const float* bias;
    float* C;              // LSTM: cell state [*, H], in place
    float* Hout;           // LSTM: hidden state [*, H]
    float* Y;              // linear: output rows (may be null)
    const int* rows;       // row list of the chain
    const int* count;
    void* nAhi;
    void* nAlo;
    int ldy, N, relu, H, K, npitch;
    int kind, nt, dep;
    int nt2, nmma;         // 256 x 256 CTA-pair kernel: 256-column tiles per row-block pair; N of its MMAs (multiple of 16, <= 256)
    int level;             // 0 linear1, 1 LSTM-0, 2 LSTM-1, 3 linear2 (jobs are stored level by level)
};
struct alignas(64) RcPhDesc {
    int njobs;
    int pad_[15];
    RcPhJob job[RC_PH_MAXJOBS];
};

// d_ctl: [1 + RC_PH_MAXJOBS * MT] ints, all zero at launch (ctl[0] = tile scheduler, then per job and 128-row block the number
// of finished tiles); MT = row blocks per job at most; max_tiles bounds the grid.
// d_trace (debug, normally null): per tile 16 x int64 {cta<<32|job<<16|m<<8|n, t_grab, t_dep, t_mma0, t_commit, t_epi0, t_stored, t_done,
// t_handed_back, then per 32-column chunk (LSTM jobs): t_tmem_read, t_math, t_stores_issued, ...} (clock64 of the SM)
// tile_width_hint (256 x 256 CTA-pair kernel only): 256 or 128 columns per tile for the wide jobs
// reserve_sms: SMs the persistent grid leaves free (for kernels of a concurrent stream: the resident CTAs take a whole SM each)
int rc_tc_phase(const RcPhDesc* d_desc, int* d_ctl, int MT, int max_tiles, void* stream, long long* d_trace = nullptr, int tile_width_hint = 256,
                int reserve_sms = 0);

// Gather + split pre-pass for all chains of a phase in one launch: segment i copies rows rows_i[0..*count_i) of src
// ([*, ld], first K columns valid, zero up to Kout) into (hi, lo)[compact row * pitch + col0 ...]; also zeroes `zero[0..nzero)`.
struct RcSplitSegM {
    const float* src; int ld, K, Kout, col0, pitch; void* hi; void* lo; const int* rows; const int* count;
    // optional "mid" fusion (net/sig_mp.py:154-167): columns [72, 141) of the rnn7 / rnn8 input are the vision / inertial joint blend,
    // computed here from the row's flags, Rcr, lerp weights and the rnn2 / rnn4 outputs instead of being read from src;
    // mid_out (one of the two segments) also stores them to the fp32 input row for the kin kernel.
    const int* mid_flags; const float* mid_rcr; const float* mid_lerpw; const float* mid_x3; const float* mid_x6; float* mid_out;
};
// `advance` (optional): an int the launch increments once (the sequence-mode frame cursor).
int rc_tc_split_multi(const RcSplitSegM* segs, int nseg, int B, int* zero, int nzero, void* stream, int* advance = nullptr);
