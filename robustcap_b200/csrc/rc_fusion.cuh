// Shared definitions of the fusion pipeline (fusion.cu: batched multi-kernel path; stream.cu: single-stream cooperative kernel).
#pragma once
#include <map>
#include <string>
#include <vector>
#include "rc_common.cuh"
#include "rc_rows.h"
#include "rc_model.cuh"
#include "rc_tc.cuh"
#include "rc_phase.cuh"

enum { NET2 = 0, NET3, NET4, NET6, NET7, NET8, NNETS };
static const int kNetId[NNETS] = {2, 3, 4, 6, 7, 8};
static const int kNetIn[NNETS] = {72, 141, 171, 240, 141, 141};
static const int kNetK1[NNETS] = {RC_K2, RC_K3, RC_K4, RC_K6, RC_K7, RC_K7};
static const int kNetH[NNETS] = {512, 512, 1280, 1024, 512, 512};
static const int kNetOut[NNETS] = {69, 3, 69, 3, 144, 2};
static const int kInitDims[4] = {69, 512, 1024, 2048};
constexpr int kInitK0 = 80;

enum { L_ALL = 0, L_HI, L_6A, L_6B, L_LATE, L_INIT, NLISTS };
// phases of a frame for the persistent grouped kernel: rnn4 + rnn2 | rnn6 on first-frame rows | rnn6 + rnn3 + rnn7 + rnn8 |
// vision updater (rnn4 + rnn6 on the low-confidence rows)
enum { PH_1 = 0, PH_6A, PH_2, PH_LATE, PH_COUNT };

struct NetDev {
    int in = 0, K1 = 0, H = 0, out = 0, out4 = 0;
    float *W1 = nullptr, *b1 = nullptr, *WL[2] = {nullptr, nullptr}, *bL[2] = {nullptr, nullptr}, *W2 = nullptr, *b2 = nullptr;
    uint16_t *WLhi[2] = {nullptr, nullptr}, *WLlo[2] = {nullptr, nullptr};   // split-fp16 copies for the tensor-core path
    RcTensorMap mWhi[2], mWlo[2];
    RcTensorMap mWhi64[2], mWlo64[2];                                        // 64-row boxes for the cluster-multicast kernel
    int K1p = 0, outp = 0;                                                   // linear1 K padded to 64, linear2 rows padded to RC_TC_BN
    uint16_t *W1hi = nullptr, *W1lo = nullptr, *W2hi = nullptr, *W2lo = nullptr;
    RcTensorMap mW1hi, mW1lo, mW2hi, mW2lo;
    RcTensorMap mW1hi64, mW1lo64, mW2hi64, mW2lo64;                          // 64-row boxes (CTA-pair kernel)
};
struct NetBuf {
    float *h[2] = {nullptr, nullptr}, *c[2] = {nullptr, nullptr}, *hn[2] = {nullptr, nullptr}, *a1 = nullptr;
};

struct rc_net {
    const rc_model* model = nullptr;
    RcNetCfg cfg;
    std::map<std::string, std::vector<float>> staging;
    bool finalized = false;
    NetDev nets[NNETS];
    float *Wi[3] = {nullptr, nullptr, nullptr}, *bi[3] = {nullptr, nullptr, nullptr};   // init_net
    int64_t weight_bytes = 0;
    std::vector<void*> allocs;
    int gemm_mode = 2;          // 0 = fp32 SIMT tiles, 1 = tcgen05 split-fp16, one launch per layer, 2 = tcgen05 persistent
                                // grouped kernel, one launch per phase of the frame (batches > 8 streams), 3 = persistent SEQUENCE
                                // kernel: frames 1 .. T-1 of rc_forward_sequence in one launch (seq_tc.cu), mode 2 elsewhere
    bool tc_ready = false;
    // sequence kernel (seq_tc.cu): in mode 2 it is chosen automatically for batches of at most seq_auto_B streams; the first
    // seq_warm frames of a sequence (where most streams re-seed rnn2 through init_net) always go through the multi-launch path
    // single-stream TMA-staged kernel (stream2.cu): LSTM weights split into contiguous input / recurrent halves [4H, H]
    bool s2_ready = false;
    float* s2_Wx[NNETS][2] = {};
    float* s2_Wh[NNETS][2] = {};
    int seq_auto_B = 128;
    int seq_warm = 16;
    int cfg_version = 0;        // bumped by rc_net_set_config: captured CUDA graphs bake the config into kernel arguments
};

struct RcSeq;                          // persistent sequence kernel (seq_tc.cu)

struct rc_state {
    const rc_net* net = nullptr;
    RcSeq* seq = nullptr;
    int B = 0;
    bool fresh = true;                 // first reset after creation also zeroes the fields reset_states leaves alone
    std::vector<void*> allocs;
    NetBuf nb[NNETS];
    float *X2 = nullptr, *X3 = nullptr, *X4 = nullptr, *X6 = nullptr, *X7 = nullptr, *XI = nullptr;
    float *Y3 = nullptr, *Y6 = nullptr, *Y7 = nullptr, *Y8 = nullptr, *Ydump = nullptr;
    float *I1 = nullptr, *I2 = nullptr, *I3 = nullptr;
    float *rcr = nullptr, *conf = nullptr, *lerpw = nullptr, *gravity = nullptr;
    // split activations (tensor-core path), one set per concurrent lane: [lane][Bpad, 2*Hmax]
    // Three operand buffers per lane: buf 0 = linear1 input [Bpad, K1p]; buf 1 = LSTM-0 input [Bpad, 2H], later linear2 input
    // [Bpad, H]; buf 2 = LSTM-1 input [Bpad, 2H].  Each GEMM epilogue writes its (split) output straight into the next buffer.
    uint16_t *Ahi[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}, *Alo[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    RcTensorMap mA0hi[2][NNETS], mA0lo[2][NNETS];  // buf 0 as [Bpad, K1p]
    RcTensorMap mA1hi[2][NNETS], mA1lo[2][NNETS];  // buf 1 as [Bpad, 2H]
    RcTensorMap mA1Hhi[2][NNETS], mA1Hlo[2][NNETS];// buf 1 as [Bpad, H]
    RcTensorMap mA2hi[2][NNETS], mA2lo[2][NNETS];  // buf 2 as [Bpad, 2H]
    // independent sub-net chains (rnn2->rnn3 || rnn4->rnn6, rnn7 || rnn8, late rnn6 || late rnn4) run on two streams
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool tc_ready = false;
    // persistent grouped kernel (phase_tc.cu): per-net operand buffers (0 = linear1 input [Bpad,K1p], 1 = LSTM-0 input [Bpad,2H],
    // 2 = LSTM-1 input [Bpad,2H], 3 = linear2 input [Bpad,H]), one job list + control block + split-segment list per phase
    uint16_t *Phi[NNETS][4] = {}, *Plo[NNETS][4] = {};
    RcPhDesc* d_phase[PH_COUNT] = {};
    int* d_ctl[PH_COUNT] = {};
    int ph_max_tiles[PH_COUNT] = {};
    RcSplitSegM ph_segs[PH_COUNT][RC_PH_MAXSEGS];
    int ph_nseg[PH_COUNT] = {};
    int ph_MT = 0;
    long long* d_trace[PH_COUNT] = {};   // debug tile traces (rc_state_debug_phase_trace)
    bool ph_ready = false;
    int* flags = nullptr;
    int* lists = nullptr;      // [NLISTS][B]
    int* counts = nullptr;     // [NLISTS]
    int* d_t = nullptr;        // device frame cursor (sequence mode)
    int* branch_log = nullptr; // caller-owned device buffer [B, T] for the branch log, or null
    RcRowState* rows = nullptr;
    // cached CUDA graph of one steady-state frame
    cudaGraphExec_t graph = nullptr;
    std::vector<const void*> graph_key;
    cudaStream_t cap_stream = nullptr;   // capture happens here (the legacy default stream cannot be captured)
    long long graph_nodes = 0;           // kernel nodes in the captured frame (launch accounting)
    // optional CUDA-event timing of the dominant kernel (rnn4's fused LSTM layers), see rc_profile_*
    int prof_on = 0;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    // low-latency single-frame path (rc_forward_online): fixed staging buffers so the frame is one cached CUDA graph
    float *on_din = nullptr, *on_dout = nullptr, *on_hin = nullptr, *on_hout = nullptr;   // device / pinned host
    cudaGraphExec_t on_graph = nullptr;
    void* on_graph_stream = nullptr;
    long long on_graph_nodes = 0;
    int on_graph_cfg_version = -1;
    // single-stream TMA-staged cooperative kernel (stream2.cu)
    unsigned* s2_bar = nullptr;
    int s2_grid = 0, s2_parity = 0;
    unsigned long long* s2_ts = nullptr;
    // single-stream cooperative kernel (stream.cu)
    unsigned* sk_bar = nullptr;
    int* sk_rows = nullptr;            // [0] = row 0, [1] = count 1, [4..5] = frame flags / need_init
    int* flags_sk = nullptr;
    // rc_forward_sequence_host with time-chunked transfers (multi-launch path): the frame loop waits for the chunk's inputs and
    // signals finished chunks, the copies run on two extra streams
    struct ChunkIO {
        int chunk = 0, T = 0;                        // frames per chunk (0 = off)
        cudaStream_t h2d = nullptr, d2h = nullptr;
        std::vector<cudaEvent_t> ev_in, ev_out;
        float *hp = nullptr, *ht = nullptr;          // host destinations
    } cio;
    // staging buffers of rc_forward_sequence_host
    float *hj = nullptr, *ha = nullptr, *ho = nullptr, *hp = nullptr, *ht = nullptr, *hft = nullptr;
    int *hlen = nullptr, *hfl = nullptr;
    int64_t host_cap_T = 0;
};


struct StepIO {
    const float *j2dc, *accc, *oric;       // bases
    long long sj, sa, so;                  // per-stream strides (floats)
    const float* gravity;                  // [B,3] or nullptr
    const float* first_tran;               // [B,3] or nullptr
    const int* row_flags;                  // [B] or nullptr
    const int* lengths;                    // [B] or nullptr
    float *pose, *tran;                    // bases
    long long sp, st;                      // per-stream strides
    const int* d_t;                        // frame cursor or nullptr (t = 0)
    int first_mode;                        // 0 never, 1 always, 2 only at t == 0
    int* branch = nullptr;                 // optional debug log [B, T] of RcBranch bits (rc_state_set_branch_log)
    long long sb = 0;                      // its per-stream stride
};

// stream.cu: one frame of ONE stream as a single cooperative kernel (all layers, grid-wide barriers instead of ~60 launches)
int rc_stream_frame(rc_state* s, const StepIO& io, int any_first_frame, void* stream);
bool rc_stream_supported(const rc_state* s);
// stream2.cu: the same frame with the weights TMA-staged through shared memory and the recurrent halves off the dependency chain
bool rc_stream2_supported(const rc_state* s);
int rc_stream2_frame(rc_state* s, const StepIO& io, int t, void* stream);
