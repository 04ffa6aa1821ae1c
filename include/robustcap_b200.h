/* robustcap_b200 — C ABI of the B200-native RobustCap hot path (fusion LSTM stack + SMPL kinematics + SMPLify).
 *
 * The reference (shaohua-pan/RobustCap) is pure Python and has no FFI layer; its "plugin boundary" for this path
 * is the Python import surface consumed by evaluate.py:1-17 and live_server.py:5-9.  This header is the C-ABI
 * that a reference maintainer binds (ctypes) to replace the arithmetic behind that surface.
 * Every entry point cites the reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - Pointers named d_* are DEVICE pointers (float32 unless noted), h_* are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls only enqueue work unless
 *     documented otherwise; the caller synchronises.
 *   - Matrices are row-major; rotation matrices [..,3,3]; quaternions wxyz.
 *   - Return value: 0 on success, negative rc_status on failure; rc_last_error() returns a message
 *     (thread-local).  No entry point falls back to a CPU computation: without a usable CUDA device every
 *     compute call returns RC_ERR_CUDA.
 *   - Handles are thread-compatible (one thread at a time per handle); no ownership of caller buffers is taken.
 */
#ifndef ROBUSTCAP_B200_H
#define ROBUSTCAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rc_status {
    RC_OK = 0,
    RC_ERR_ARG = -1,      /* bad argument (null pointer, size mismatch, unknown tensor name) */
    RC_ERR_CUDA = -2,     /* CUDA runtime error, incl. no device */
    RC_ERR_STATE = -3,    /* call order violated (e.g. net not finalised) */
    RC_ERR_ALLOC = -4
} rc_status;

typedef struct rc_model rc_model;   /* SMPL constants on the device          (articulate/model.py:17-40)   */
typedef struct rc_net rc_net;       /* packed weights of the six LSTM stacks (net/sig_mp.py:47-93)         */
typedef struct rc_state rc_state;   /* per-stream recurrent + tracker state for B streams (sig_mp.py:85-104) */

const char* rc_version(void);
const char* rc_last_error(void);
/* Number of this library's kernels launched by the calling process so far (bench.py "gpu_launches"). */
int64_t rc_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Rotation representations — articulate/math/angular.py (N independent items, contiguous).
 * ------------------------------------------------------------------------------------------------------- */
int rc_r6d_to_rotmat(const float* d_r6d, float* d_R, int64_t n, void* stream);        /* angular.py:249-264 */
int rc_rotmat_to_r6d(const float* d_R, float* d_r6d, int64_t n, void* stream);        /* angular.py:267-274 */
int rc_axis_angle_to_rotmat(const float* d_aa, float* d_R, int64_t n, void* stream);  /* angular.py:221-233 */
int rc_rotmat_to_axis_angle(const float* d_R, float* d_aa, int64_t n, void* stream);  /* angular.py:236-246 (cv2.Rodrigues semantics) */
int rc_batch_rodrigues(const float* d_aa, float* d_R, int64_t n, void* stream);       /* net/smplify/temporal_smplify.py:25-59 */
int rc_quat_to_rotmat(const float* d_q, float* d_R, int64_t n, void* stream);         /* angular.py:306-318 */
int rc_quat_to_axis_angle(const float* d_q, float* d_aa, int64_t n, void* stream);    /* angular.py:277-290 */
int rc_axis_angle_to_quat(const float* d_aa, float* d_q, int64_t n, void* stream);    /* angular.py:293-303 */
int rc_quat_product(const float* d_q1, const float* d_q2, float* d_q, int64_t n, void* stream); /* angular.py:79-93 */

/* ---------------------------------------------------------------------------------------------------------
 * Kinematic tree — articulate/math/spatial.py.  `h_parent` is a HOST int32[nj] table, parent[0] < 0,
 * parent[i] < i; nj <= 64.  b = number of frames.
 * ------------------------------------------------------------------------------------------------------- */
int rc_tree_fk_R(const float* d_Rl, float* d_Rg, const int32_t* h_parent, int nj, int64_t b, void* stream);   /* spatial.py:170-194 */
int rc_tree_ik_R(const float* d_Rg, float* d_Rl, const int32_t* h_parent, int nj, int64_t b, void* stream);   /* spatial.py:197-221 */
int rc_tree_fk_T(const float* d_Tl, float* d_Tg, const int32_t* h_parent, int nj, int64_t b, void* stream);   /* spatial.py:224-249, [b,nj,4,4] */
int rc_tree_ik_T(const float* d_Tg, float* d_Tl, const int32_t* h_parent, int nj, int64_t b, void* stream);   /* spatial.py:252-277 */
int rc_tree_bone_to_joint(const float* d_bone, float* d_joint, const int32_t* h_parent, int nj, int64_t b, void* stream); /* spatial.py:126-145 */
int rc_tree_joint_to_bone(const float* d_joint, float* d_bone, const int32_t* h_parent, int nj, int64_t b, void* stream); /* spatial.py:148-167 */

/* ---------------------------------------------------------------------------------------------------------
 * SMPL body model — articulate/model.py:17-241.
 * rc_model_create copies the (host) constants: zero-pose joints J[24,3] and vertices v[nv,3] (both already
 * root-centred as get_zero_pose_joint_and_vertex returns them, model.py:87), skinning weights [nv,24], parent
 * table, and the 33-entry MediaPipe vertex table (config.py:99).
 * ------------------------------------------------------------------------------------------------------- */
int rc_model_create(rc_model** out, const float* h_joints, const float* h_verts, const float* h_skin_w,
                    int32_t nv, const int32_t* h_parent, const int32_t* h_mp_mask);
void rc_model_destroy(rc_model* m);
/* ParametricModel.forward_kinematics (model.py:209-241).
 *   d_pose [b,24,3,3] local rotations; d_tran [b,3] or NULL; d_joints_rest/d_verts_rest: per-frame zero-pose
 *   joints [b,24,3] / vertices [b,nv,3] for shaped bodies, or NULL for the model's mean shape.
 *   Outputs: d_Rg [b,24,3,3], d_joint [b,24,3]; d_vert [b,nv,3] when not NULL (calc_mesh=True). */
int rc_model_forward_kinematics(const rc_model* m, const float* d_pose, const float* d_tran,
                                const float* d_joints_rest, const float* d_verts_rest, int64_t b,
                                float* d_Rg, float* d_joint, float* d_vert, void* stream);
/* The 33 synthetic MediaPipe points only (net/sig_mp.py:287-299 on top of model.py:209-241): d_kp [b,33,3].
 * Skins just the 21 vertices sync_mp3d reads instead of all 6890. */
int rc_model_keypoints(const rc_model* m, const float* d_pose, const float* d_tran, int64_t b,
                       float* d_joint, float* d_kp, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fusion network — net/sig_mp.py:23-274 (class Net) on top of articulate/utils/torch/rnn.py:92-219.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct rc_net_config {      /* Net class attributes, sig_mp.py:27-45 and :91-93 */
    double conf_lo, conf_hi;        /* conf_range            (0.7, 0.8); live (0.85, 0.9) */
    double tran_filter_num;         /* tran_filter_num       0.05; live 0.01              */
    float contact_threshold;        /* 0.7  */
    float height_threshold;         /* height_threhold 0.15 */
    float distance_threshold;       /* distrance_threshold 10 */
    int32_t use_flat_floor;         /* 1 */
    int32_t live;                   /* 0 */
    int32_t update_vision_freq;     /* 30 */
} rc_net_config;

void rc_net_default_config(rc_net_config* cfg, int live);
int rc_net_create(rc_net** out, const rc_model* model, const rc_net_config* cfg);
void rc_net_destroy(rc_net* n);
int rc_net_set_config(rc_net* n, const rc_net_config* cfg);
/* Net.load_state_dict: one call per state_dict entry (key set of SURVEY.md §5, e.g. "rnn4.rnn.weight_hh_l1",
 * "rnn2.init_net.2.bias"); h_data is HOST float32, numel must match.  Then rc_net_finalize packs the tensors
 * into the device layout (gate-interleaved [4H, 2H] LSTM blocks, K padded to 16) and frees the staging copy. */
int rc_net_set_tensor(rc_net* n, const char* key, const float* h_data, int64_t numel);
int rc_net_finalize(rc_net* n);
int64_t rc_net_weight_bytes(const rc_net* n);     /* bytes of packed per-frame weights resident in HBM */
/* Batched (B > 8) back end: 3 = persistent SEQUENCE kernel (seq_tc.cu): after a few warm-up frames all remaining frames of
 * rc_forward_sequence run in ONE launch — GEMM tiles (tcgen05) and the per-frame row logic (prep / joint blend / kinematics +
 * translation state machine / init_net) as items of one dependency queue, resident CTAs, no kernel boundary per frame;
 * 2 (default) = persistent grouped tcgen05 kernel, one launch per phase of the frame (phase_tc.cu), which hands small batches
 * (<= 128 streams, see rc_net_set_seq_options) to the sequence kernel; 1 = tcgen05, one launch per layer (gemm_tc.cu); 0 = fp32 SIMT
 * tiles.  All tcgen05 paths use split-fp16 operands with fp32-level accuracy.  B <= 8 always uses the weight-streaming GEMV kernels. */
int rc_net_set_gemm_mode(rc_net* n, int mode);
/* Sequence-kernel policy: in gemm mode 2, batches of at most auto_max_streams streams (default 128 = one row block; 0 = never) use
 * the sequence kernel; the first warm_frames frames (default 16) of every sequence go through the multi-launch path, which runs the
 * init_net re-seeds (net/sig_mp.py:178-183, mostly in the first frames) on the tensor cores.  Negative / zero values keep the setting. */
int rc_net_set_seq_options(rc_net* n, int32_t auto_max_streams, int32_t warm_frames);

int rc_state_create(rc_state** out, const rc_net* net, int32_t b);
void rc_state_destroy(rc_state* s);
int rc_state_reset(rc_state* s, void* stream);                                   /* Net.reset_states, sig_mp.py:95-104 */
int rc_state_set_gravity(rc_state* s, const float* h_gravity3, void* stream);   /* Net.gravityc, one value for all streams */

/* flag bits for d_row_flags */
#define RC_ROW_FIRST_FRAME 1   /* forward_online(first_frame=True) */
#define RC_ROW_FIRST_TRAN 2    /* forward_online(first_tran=...)   */

/* Net.forward_online for B streams, one frame (sig_mp.py:113-274).
 *   d_j2dc [b,33,3], d_accc [b,6,3], d_oric [b,6,3,3]; d_first_tran [b,3] or NULL; d_row_flags int32[b] or NULL;
 *   d_gravity [b,3] per-stream gravity or NULL (use rc_state_set_gravity value);
 *   any_first_frame: host hint, non-zero when some row has RC_ROW_FIRST_FRAME (adds the extra rnn6 pass of
 *   sig_mp.py:155-156).  Outputs d_pose [b,24,3,3], d_tran [b,3]. */
int rc_forward_step(rc_state* s, const float* d_j2dc, const float* d_accc, const float* d_oric,
                    const float* d_gravity, const float* d_first_tran, const int32_t* d_row_flags,
                    int any_first_frame, float* d_pose, float* d_tran, void* stream);

/* Net.forward_online for ONE stream (state created with b = 1), lowest latency: inputs ([33,3], [6,3], [6,3,3], optional
 * first_tran [3]) are HOST pointers (inputs_on_device = 0; packed into one pinned buffer, one H2D copy) or DEVICE pointers
 * (inputs_on_device = 1); the frame runs as one cached CUDA graph; results are copied back to the HOST pointers h_pose
 * [24,3,3] and h_tran [3] and the stream is synchronised — exactly the reference's contract of returning CPU tensors
 * (sig_mp.py:274). */
int rc_forward_online(rc_state* s, const float* j2dc, const float* accc, const float* oric, const float* first_tran,
                      int first_frame, int inputs_on_device, float* h_pose, float* h_tran, void* stream);

/* forward_offline := reset_states + forward_online over t (evaluate.py:75-85,93), batched over B sequences.
 *   d_j2dc [b,T,33,3], d_accc [b,T,6,3], d_oric [b,T,6,3,3]; d_lengths int32[b] or NULL (ragged batch: rows
 *   stop advancing after their length; outputs beyond it are left untouched);
 *   first-frame arguments apply to t = 0.  Outputs d_pose [b,T,24,3,3], d_tran [b,T,3].
 *   use_graph != 0 replays one captured CUDA graph per frame for t >= 1. */
int rc_forward_sequence(rc_state* s, int32_t T, const float* d_j2dc, const float* d_accc, const float* d_oric,
                        const int32_t* d_lengths, const float* d_gravity, const float* d_first_tran,
                        const int32_t* d_row_flags, int any_first_frame, float* d_pose, float* d_tran,
                        int use_graph, void* stream);

/* Same with HOST buffers (pinned recommended): copies inputs H2D, runs, copies results D2H and synchronises
 * the stream — the end-to-end call of the plugin. */
int rc_forward_sequence_host(rc_state* s, int32_t T, const float* h_j2dc, const float* h_accc, const float* h_oric,
                             const int32_t* h_lengths, const float* h_first_tran, const int32_t* h_row_flags,
                             float* h_pose, float* h_tran, int use_graph, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * SMPLify objective — net/smplify/losses.py:23-91 evaluated as in temporal_smplify.py:153-165, with the analytic
 * gradient that replaces the reference's autograd pass.  The optimiser (torch.optim.LBFGS, a third-party class the
 * reference calls at temporal_smplify.py:151) stays on the host side.
 *   rc_smplify_create: GMM prior constants as prior.py:113-143 builds them: means [8,69], precisions [8,69,69],
 *                      log(nll_weights) [8] (HOST float32); t = frames per sequence.
 *   rc_smplify_loss_grad: d_pose [t,72] axis-angle (rodrigues 0 = batch_rodrigues, temporal_smplify.py:25-59;
 *                      1 = axis_angle_to_rotation_matrix, angular.py:221-233) or [t,24,3,3] matrices (rodrigues 2,
 *                      value-only, get_fitting_loss temporal_smplify.py:198-220); d_tran [t,3]; d_j2d [t,33,2] pixels;
 *                      d_conf [t,33]; d_cam_k [3,3]; d_ref3d [t,33,3] (the detached initial points); d_imu_aa [t,6,3]
 *                      (axis-angles of imu_ori).  Outputs: d_loss [1] (output='sum'), d_grad_pose [t,72], d_grad_tran
 *                      [t,3] (both or neither), d_reproj [t,33] (output='reprojection') or NULL.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct rc_smplify rc_smplify;
int rc_smplify_create(rc_smplify** out, const rc_model* model, const float* h_means, const float* h_precisions,
                      const float* h_log_nll_weights, int32_t t);
void rc_smplify_destroy(rc_smplify* s);
int rc_smplify_loss_grad(rc_smplify* s, const float* d_pose, const float* d_tran, const float* d_j2d, const float* d_conf,
                         const float* d_cam_k, const float* d_ref3d, const float* d_imu_aa, int rodrigues, float* d_loss,
                         float* d_grad_pose, float* d_grad_tran, float* d_reproj, void* stream);
/* The optimisation of TemporalSMPLify.__call__ (net/smplify/temporal_smplify.py:139-166: torch.optim.LBFGS(max_iter, lr,
 * line_search_fn='strong_wolfe').step(closure), one step) for n_seq independent sequences of T frames (T = the handle's batch size),
 * entirely on the device: ONE thread block per sequence evaluates the closure (objective + analytic gradient, the kernels behind
 * rc_smplify_loss_grad) and runs torch/optim/lbfgs.py's two-loop recursion and strong-Wolfe line search (_strong_wolfe,
 * _cubic_interpolate) with every reduction inside the block — no launch and no host round trip per iteration.
 * d_aa_init [S,T,72] / d_tran_init [S,T,3]: start (axis-angle from cv2.Rodrigues semantics, :106); d_j2d [S,T,33,2] pixel key points,
 * d_conf [S,T,33] (ignored joints already zeroed, :148), d_camk [9] (camk_per_seq = 0) or [S,9]; d_ref3d [S,T,33,3] (:108-109);
 * d_imu_aa [S,T,18].  Outputs d_aa_out / d_tran_out (may alias nothing of the inputs); d_stats (optional) [S,4] =
 * {first loss, final loss, closure evaluations, iterations}. */
int rc_smplify_run(rc_smplify* s, int32_t n_seq, const float* d_aa_init, const float* d_tran_init, const float* d_j2d, const float* d_conf,
                   const float* d_camk, int32_t camk_per_seq, const float* d_ref3d, const float* d_imu_aa, int32_t max_iter, float lr,
                   float* d_aa_out, float* d_tran_out, float* d_stats, void* stream);


/* ---------------------------------------------------------------------------------------------------------
 * Evaluation metrics next to the hot path — evaluate.py:120-133 (cal_mpjpe) with utils.py:138-203 (Procrustes).
 *   d_jreg [nj_rows >= 14, nv] dense joint regressor (J_regressor_h36m), d_pose / d_gt_pose [b,24,3,3] local rotations.
 *   d_out [b,3] per frame: mean distance of the first 14 regressed joints after pelvis alignment, mean vertex distance,
 *   Procrustes-aligned mean joint distance (0 when with_pa == 0).  The caller averages over frames like the reference.
 * ------------------------------------------------------------------------------------------------------- */
int rc_metrics_mpjpe(const rc_model* m, const float* d_jreg, int32_t nj_rows, const float* d_pose, const float* d_gt_pose,
                     int64_t b, int with_pa, float* d_out, void* stream);

/* Test tap: one fused LSTM layer of sub-net `ni` (0..5 = rnn2,3,4,6,7,8), layer 0/1, on caller-provided device data for
 * all b rows of the state: d_x [b,H], d_hprev [b,H], d_c [b,H] (in place), d_hout [b,H]; mode as rc_net_set_gemm_mode. */
int rc_state_debug_lstm(rc_state* s, int ni, int layer, int mode, const float* d_x, const float* d_hprev, float* d_c,
                        float* d_hout, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Callers / data formats on either side of the hot path.
 * ------------------------------------------------------------------------------------------------------- */
/* evaluate.py:38-52, 68-73 — dataset rows -> network inputs, packed on the device.  Row b (one sequence seen by one camera)
 * has row_off[b+1] - row_off[b] frames; its key points d_j2d [sum, 33, 3] are (u, v, confidence) with u, v in [0, 1]
 * (scaled by img_w, img_h to pixels); the world-frame IMU arrays d_imu_acc [*, 6, 3] / d_imu_ori [*, 6, 3, 3] are indexed through
 * d_src[b] (source sequence) and d_seq_off (int64 frame offsets), so the cameras of a sequence share them.  d_cam_T [B,4,4] is
 * Tcw, d_cam_K [B,3,3] the intrinsics.  Outputs: d_j2dc [B,Tmax,33,3] = (K^-1 [u w, v h, 1])_xy with the confidence,
 * d_accc = R acc, d_oric = R ori (zeros / identity beyond the row's length), d_gravity [B,3] = R [0,-1,0], d_lengths int32 [B]. */
int rc_pack_inputs(int32_t b, int32_t t_max, const int32_t* d_src, const int64_t* d_seq_off, const int64_t* d_row_off,
                   const float* d_j2d, const float* d_imu_acc, const float* d_imu_ori, const float* d_cam_T, const float* d_cam_K,
                   float img_w, float img_h, float* d_j2dc, float* d_accc, float* d_oric, float* d_gravity, int32_t* d_lengths,
                   void* stream);
/* preprocess.py:22-33, 290-302 — synthetic IMU readings of a pose sequence: mesh FK restricted to the n_imu vertices h_vi
 * (config.vi_mask), accelerations by _syn_acc (second differences at 60 fps, smoothed over +-smooth_n frames; smooth_n = 2 in the
 * reference), orientations = global rotations of the joints h_ji (config.ji_mask).  d_pose [n,24,3,3] local rotations, d_tran [n,3]
 * or NULL, d_joints_rest [n,24,3] / d_imu_vrest [n_imu,3] zero-pose joints / IMU vertices of a shaped body or NULL (mean shape).
 * Outputs d_acc [n,n_imu,3], d_ori [n,n_imu,3,3]; optional d_joint [n,24,3], d_vimu [n,n_imu,3]. */
int rc_synthesize_imu(const rc_model* m, const float* d_pose, const float* d_tran, const float* d_joints_rest,
                      const float* d_imu_vrest, const int32_t* h_vi, const int32_t* h_ji, int32_t n_imu, int32_t smooth_n, int64_t n,
                      float* d_acc, float* d_ori, float* d_joint, float* d_vimu, void* stream);

/* Live wire formats (host code): the UDP datagram of live_detector.py:57-61 "uv(99)#ori(54)#acc(18)#RCM(9)" (comma-separated
 * decimals, parsed like live_server.py:42-45: float() per token, then float32; h_rcm may be NULL) and the Unity TCP message of
 * live_server.py:55-59 "%g,...(72 axis-angle values)#%g,%g,%g$".  rc_live_format_pose returns the number of bytes written
 * (>= 0) or a negative rc_status. */
int rc_live_parse_frame(const char* h_text, int32_t len, float* h_uv, float* h_ori, float* h_acc, float* h_rcm);
int rc_live_format_pose(const float* h_pose_aa, const float* h_tran, char* h_out, int32_t cap);
/* The sensor server's binary datagram (live_demo_sync.py:262-268, get_from_udp): float32[8 n] = t[n] | q[n,4] (wxyz) | a[n,3]. */
int rc_live_parse_imu_packet(const void* h_data, int32_t nbytes, int32_t n_imu, float* h_t, float* h_q, float* h_a);

/* CUDA-event timing of the dominant kernel (the fused LSTM layers of rnn4: [rows, 2H] x [2H, 4H], H = 1280) for
 * bench.py's roofline: enable, run (non-graph launches), collect.  collect synchronises the device, returns the
 * summed duration of the recorded launches, their number, and the algorithmic FLOPs one stream-row costs in one
 * such launch (2 * 4H * 2H). */
int rc_profile_enable(rc_state* s, int on);
int rc_profile_collect(rc_state* s, double* total_ms, int64_t* launches, double* flop_per_row);

/* Debug / test taps: copy a sub-net output of the last step to the host: which in {2,3,4,6,7,8}; out has
 * b * width floats (width = 69,3,69,3,144,2). */
int rc_state_debug_output(rc_state* s, int which, float* h_out, void* stream);

/* Debug / parity aid: `d_log` (device, int32 [B, T], caller-owned; NULL switches it off) receives, for every frame the following
 * rc_forward_sequence calls process, which data-dependent decisions the translation / contact / floor logic took
 * (net/sig_mp.py:185-225): bit 1 contact branch (:190), 2 contact argmax, 4 tran snapped to pc (:200-201), 8 tran lerped to pc
 * (:203), 16 floor sample stored (:208-214), 32 / 64 floor snap through the far / near foot (:217-221), 128 rnn2 re-seeded by
 * init_net (:178-183).  Parity tests report the first frame where the CUDA path and the oracle disagree and which bit flipped. */
int rc_state_set_branch_log(rc_state* s, int32_t* d_log);

/* Debug tap of the persistent sequence kernel (gemm mode 3; environment RC_SEQ_STATS=1 switches the counters on): per job of the
 * frame's work queue, h_out[4 j + {0: work items run, 1: clocks spent waiting for state / hazard dependencies, 2: clocks waiting for
 * the producer of the x half of an LSTM operand, 3: items skipped because no stream of the block needed them}] of the last
 * rc_forward_sequence call; returns the number of jobs written (<= max_jobs). */
int rc_state_debug_seq_stats(rc_state* s, long long* h_out, int32_t max_jobs);

/* Debug tap of the persistent grouped GEMM kernel (gemm mode 2): enable != 0 makes every later launch record, per tile, 16 int64
 * {cta<<32|job<<16|row block<<8|column tile, clock64 at: grab, dependency met, first MMA, last commit, accumulators seen by the
 * epilogue, outputs stored, tile published, then finer epilogue stamps}; with h_out != NULL the trace of `phase` (0 rnn4+rnn2, 1 rnn6 on
 * first-frame rows, 2 rnn6+rnn3+rnn7+rnn8, 3 vision updater) of the last frame is copied (at most max_tiles tiles) and
 * *ntiles receives the phase's tile bound. */
int rc_state_debug_phase_trace(rc_state* s, int enable, int phase, long long* h_out, int max_tiles, int* ntiles);

#ifdef __cplusplus
}
#endif
#endif /* ROBUSTCAP_B200_H */
