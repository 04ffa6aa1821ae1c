// Interface of the persistent SEQUENCE kernel (seq_tc.cu): frames t0 .. T-1 of Net.forward_offline (net/sig_mp.py:113-274 per
// frame, evaluate.py:75-85 over the frames) in ONE launch — every GEMM of the six LSTM stacks as tcgen05 tiles AND the per-frame
// row logic (prep / joint blend / kinematics + translation state machine / init_net) as jobs of the same dependency queue.
#pragma once

struct rc_state;
struct StepIO;

// true when the sequence kernel can run this state (tensor-core path built, non-live config, B > 8)
bool rc_seq_supported(const rc_state* s);
// Runs frames t0 .. T-1 (t0 >= 1: frame 0 with its first_frame / first_tran specials has gone through the multi-launch path, whose
// fp32 LSTM state this call converts into the split operand planes).  bn = tile width (64 or 128 gate columns).
int rc_seq_run(rc_state* s, const StepIO& io, int T, int t0, int bn, void* stream);
void rc_seq_destroy(rc_state* s);
// per-job scheduler statistics of the last run (debug): out[j*4 + {tiles, dep1 wait clk, dep2 wait clk, skipped}], returns the number of jobs
int rc_seq_stats(rc_state* s, long long* out, int max_jobs);
// debug: the blended joints (input columns 72..140 of rnn7 / rnn8) of the last frame the sequence kernel ran, [Bpad, 72] fp32
const float* rc_seq_debug_j3dr(const rc_state* s);
