// SURVEY.md §8(f) rows 2 and 3 — the callers / data formats on either side of the hot path.
//
//   rc_pack_inputs       evaluate.py:38-52, 68-73: per (sequence, camera) row, world-frame IMU data and image-plane key points ->
//                        camera-frame, K^-1-normalised network inputs, packed [B, Tmax, ...] + lengths on the device (replaces the
//                        RNNDataset lists and the per-sequence host loop).
//   rc_synthesize_imu    preprocess.py:22-33, 290-302: SMPL pose -> mesh FK of the six IMU vertices -> second-difference
//                        accelerations (_syn_acc) and the global orientations of the six IMU joints.
//
// Both are HBM-bound element work (no tensor cores): coalesced reads, one pass.
#include "rc_common.cuh"
#include "rc_rows.h"
#include "rc_model.cuh"

int rc_smpl_chain_launch_ex(const rc_model* m, const float* pose, const float* tran, const float* jrest_b, long long b, float* Rg,
                            float* joint, float* Tskin, void* stream);

namespace {

constexpr int kPackFrames = 4;      // frames per block

// 3x3 inverse by the adjugate (fp32).  torch.inverse (LU) differs by rounding only; the parity tests bound it.
__device__ __forceinline__ void inv3(const float* k, float* o) {
    const float c00 = k[4] * k[8] - k[5] * k[7], c01 = k[5] * k[6] - k[3] * k[8], c02 = k[3] * k[7] - k[4] * k[6];
    const float det = k[0] * c00 + k[1] * c01 + k[2] * c02;
    const float r = 1.f / det;
    o[0] = c00 * r; o[1] = (k[2] * k[7] - k[1] * k[8]) * r; o[2] = (k[1] * k[5] - k[2] * k[4]) * r;
    o[3] = c01 * r; o[4] = (k[0] * k[8] - k[2] * k[6]) * r; o[5] = (k[2] * k[3] - k[0] * k[5]) * r;
    o[6] = c02 * r; o[7] = (k[1] * k[6] - k[0] * k[7]) * r; o[8] = (k[0] * k[4] - k[1] * k[3]) * r;
}

struct PackArgs {
    int B, Tmax;
    const int* src;              // [B] source sequence of the row (world-frame IMU arrays are shared by the cameras of a sequence)
    const long long* seq_off;    // [S+1] frame offsets of the sequences in imu_acc / imu_ori
    const long long* row_off;    // [B+1] frame offsets of the rows in j2d
    const float *j2d, *acc, *ori, *cam_T, *cam_K;
    float img_w, img_h;
    float *j2dc, *accc, *oric, *gravity;
    int* lengths;
};

__global__ void __launch_bounds__(192) rc_pack_inputs_kernel(PackArgs a) {
    const int b = blockIdx.y;
    const long long r0 = a.row_off[b];
    const int len = (int)(a.row_off[b + 1] - r0);
    const long long s0 = a.seq_off[a.src[b]];
    __shared__ float R[9], Ki[9];
    if (threadIdx.x == 0) {
        const float* T = a.cam_T + (size_t)b * 16;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = T[r * 4 + c];
        inv3(a.cam_K + (size_t)b * 9, Ki);
        if (blockIdx.x == 0) {
            a.lengths[b] = len < a.Tmax ? len : a.Tmax;
            for (int r = 0; r < 3; ++r) a.gravity[b * 3 + r] = R[r * 3 + 0] * 0.f + R[r * 3 + 1] * -1.f + R[r * 3 + 2] * 0.f;   // Tcw.mm([0,-1,0]) (evaluate.py:73)
        }
    }
    __syncthreads();
    const int t0 = blockIdx.x * kPackFrames;
    for (int e = threadIdx.x; e < kPackFrames * 171; e += blockDim.x) {
        const int t = t0 + e / 171, q = e % 171;
        if (t >= a.Tmax) break;
        const bool live = t < len;
        const size_t ob = (size_t)b * a.Tmax + t;
        if (q < 99) {                                            // key points: pixels -> z = 1 plane, confidence kept (evaluate.py:45-49, 70-72)
            const int k = q / 3, c = q % 3;
            float v = 0.f;
            if (live) {
                const float* p = a.j2d + ((size_t)(r0 + t) * 33 + k) * 3;
                if (c == 2) v = p[2];
                else {
                    const float x = p[0] * a.img_w, y = p[1] * a.img_h;
                    v = Ki[c * 3 + 0] * x + Ki[c * 3 + 1] * y + Ki[c * 3 + 2] * 1.f;
                }
            }
            a.j2dc[ob * 99 + q] = v;
        } else if (q < 117) {                                    // accc = Tcw[:3,:3] acc (the homogeneous 0 drops the translation, :44)
            const int i = (q - 99) / 3, r = (q - 99) % 3;
            float v = 0.f;
            if (live) {
                const float* p = a.acc + ((size_t)(s0 + t) * 6 + i) * 3;
                v = R[r * 3 + 0] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2];
            }
            a.accc[ob * 18 + (q - 99)] = v;
        } else {                                                 // oric = Tcw[:3,:3] ori (:43)
            const int i = (q - 117) / 9, r = ((q - 117) % 9) / 3, c = (q - 117) % 3;
            float v = (r == c) ? 1.f : 0.f;                      // identity beyond the sequence
            if (live) {
                const float* p = a.ori + ((size_t)(s0 + t) * 6 + i) * 9;
                v = R[r * 3 + 0] * p[c] + R[r * 3 + 1] * p[3 + c] + R[r * 3 + 2] * p[6 + c];
            }
            a.oric[ob * 54 + (q - 117)] = v;
        }
    }
}

// the IMU vertices of every frame: v = (sum_j w_j T'_j) [rest; 1] + tran   (model.py:236-241 restricted to vi_mask)
__global__ void __launch_bounds__(128) rc_imu_vertex_kernel(const float* __restrict__ Tskin, const float* __restrict__ W,
                                                             const float* __restrict__ vrest, const float* __restrict__ tran,
                                                             const int* __restrict__ vid, int nimu, long long n, float* __restrict__ v) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * nimu) return;
    const long long f = e / nimu;
    const int k = (int)(e % nimu);
    const float* w = W + (size_t)vid[k] * RC_NJ;
    float T[12];
    for (int q = 0; q < 12; ++q) T[q] = 0.f;
    for (int j = 0; j < RC_NJ; ++j) {
        const float wj = w[j];
        for (int q = 0; q < 12; ++q) T[q] = fmaf(wj, Tskin[f * 288 + j * 12 + q], T[q]);
    }
    const float x = vrest[k * 3], y = vrest[k * 3 + 1], z = vrest[k * 3 + 2];
    for (int r = 0; r < 3; ++r) {
        float o = T[r * 4] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3];
        if (tran) o = RC_ADD(o, tran[f * 3 + r]);
        v[e * 3 + r] = o;
    }
}

// _syn_acc (preprocess.py:22-33): second differences at 60 fps, smoothed over +-smooth_n frames away from the ends
__global__ void __launch_bounds__(128) rc_syn_acc_kernel(const float* __restrict__ v, long long n, int width, int smooth_n,
                                                          float* __restrict__ acc) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * width) return;
    const long long t = e / width;
    const int c = (int)(e % width);
    float a = 0.f;
    if (t >= 1 && t <= n - 2) a = RC_MUL(RC_SUB(RC_ADD(v[(t - 1) * width + c], v[(t + 1) * width + c]), RC_MUL(2.f, v[t * width + c])), 3600.f);
    if (smooth_n / 2 != 0 && t >= smooth_n && t < n - smooth_n) {
        const float d = RC_SUB(RC_ADD(v[(t - smooth_n) * width + c], v[(t + smooth_n) * width + c]), RC_MUL(2.f, v[t * width + c]));
        a = RC_DIV(RC_MUL(d, 3600.f), (float)(smooth_n * smooth_n));
    }
    acc[e] = a;
}

__global__ void __launch_bounds__(128) rc_gather_ori_kernel(const float* __restrict__ Rg, const int* __restrict__ jid, int nimu, long long n,
                                                             float* __restrict__ ori) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * nimu * 9) return;
    const long long f = e / (nimu * 9);
    const int k = (int)(e / 9 % nimu), q = (int)(e % 9);
    ori[e] = Rg[(f * RC_NJ + jid[k]) * 9 + q];
}

}  // namespace

extern "C" {

int rc_pack_inputs(int32_t B, int32_t Tmax, const int32_t* d_src, const int64_t* d_seq_off, const int64_t* d_row_off,
                   const float* d_j2d, const float* d_imu_acc, const float* d_imu_ori, const float* d_cam_T, const float* d_cam_K,
                   float img_w, float img_h, float* d_j2dc, float* d_accc, float* d_oric, float* d_gravity, int32_t* d_lengths,
                   void* stream) {
    RC_ARG(B >= 0 && Tmax >= 0);
    if (B == 0 || Tmax == 0) return RC_OK;
    RC_ARG(d_src && d_seq_off && d_row_off && d_j2d && d_imu_acc && d_imu_ori && d_cam_T && d_cam_K);
    RC_ARG(d_j2dc && d_accc && d_oric && d_gravity && d_lengths);
    PackArgs a;
    a.B = B; a.Tmax = Tmax; a.src = d_src; a.seq_off = (const long long*)d_seq_off; a.row_off = (const long long*)d_row_off;
    a.j2d = d_j2d; a.acc = d_imu_acc; a.ori = d_imu_ori; a.cam_T = d_cam_T; a.cam_K = d_cam_K; a.img_w = img_w; a.img_h = img_h;
    a.j2dc = d_j2dc; a.accc = d_accc; a.oric = d_oric; a.gravity = d_gravity; a.lengths = d_lengths;
    dim3 grid(rc_cdiv(Tmax, kPackFrames), B);
    RC_LAUNCH(rc_pack_inputs_kernel, grid, 192, 0, stream, a);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

int rc_synthesize_imu(const rc_model* m, const float* d_pose, const float* d_tran, const float* d_joints_rest,
                      const float* d_imu_vrest, const int32_t* h_vi, const int32_t* h_ji, int32_t n_imu, int32_t smooth_n, int64_t n,
                      float* d_acc, float* d_ori, float* d_joint, float* d_vimu, void* stream) {
    RC_ARG(m && n >= 0 && n_imu > 0 && n_imu <= 32 && smooth_n >= 0 && h_vi && h_ji);
    if (n == 0) return RC_OK;
    RC_ARG(d_pose && d_acc && d_ori);
    for (int k = 0; k < n_imu; ++k) RC_ARG(h_vi[k] >= 0 && h_vi[k] < m->nv && h_ji[k] >= 0 && h_ji[k] < RC_NJ);
    cudaStream_t st = (cudaStream_t)stream;
    float *Rg = nullptr, *jt = nullptr, *Ts = nullptr, *v = nullptr, *vrest = nullptr;
    int* ids = nullptr;
    RC_CUDA(cudaMallocAsync(&Rg, (size_t)n * 216 * sizeof(float), st));
    RC_CUDA(cudaMallocAsync(&jt, (size_t)n * 72 * sizeof(float), st));
    RC_CUDA(cudaMallocAsync(&Ts, (size_t)n * 288 * sizeof(float), st));
    RC_CUDA(cudaMallocAsync(&v, (size_t)n * n_imu * 3 * sizeof(float), st));
    RC_CUDA(cudaMallocAsync(&vrest, (size_t)n_imu * 3 * sizeof(float), st));
    RC_CUDA(cudaMallocAsync(&ids, (size_t)2 * n_imu * sizeof(int), st));
    int h_ids[64];
    for (int k = 0; k < n_imu; ++k) { h_ids[k] = h_vi[k]; h_ids[n_imu + k] = h_ji[k]; }
    RC_CUDA(cudaMemcpyAsync(ids, h_ids, (size_t)2 * n_imu * sizeof(int), cudaMemcpyHostToDevice, st));
    if (d_imu_vrest) RC_CUDA(cudaMemcpyAsync(vrest, d_imu_vrest, (size_t)n_imu * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else for (int k = 0; k < n_imu; ++k)
        RC_CUDA(cudaMemcpyAsync(vrest + k * 3, m->d_verts + (size_t)h_vi[k] * 3, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    RC_CUDA(cudaStreamSynchronize(st));          // h_ids is a stack buffer
    int rc = rc_smpl_chain_launch_ex(m, d_pose, d_tran, d_joints_rest, n, Rg, jt, Ts, stream);
    if (rc == RC_OK) {
        RC_LAUNCH(rc_imu_vertex_kernel, rc_cdiv(n * n_imu, 128), 128, 0, stream, Ts, m->d_skin_w, vrest, d_tran, ids, n_imu, (long long)n, v);
        RC_LAUNCH(rc_syn_acc_kernel, rc_cdiv(n * n_imu * 3, 128), 128, 0, stream, v, (long long)n, n_imu * 3, smooth_n, d_acc);
        RC_LAUNCH(rc_gather_ori_kernel, rc_cdiv(n * n_imu * 9, 128), 128, 0, stream, Rg, ids + n_imu, n_imu, (long long)n, d_ori);
        if (cudaGetLastError() != cudaSuccess) { rc_set_error("rc_synthesize_imu: launch failed"); rc = RC_ERR_CUDA; }
    }
    if (rc == RC_OK && d_joint) cudaMemcpyAsync(d_joint, jt, (size_t)n * 72 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (rc == RC_OK && d_vimu) cudaMemcpyAsync(d_vimu, v, (size_t)n * n_imu * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaFreeAsync(Rg, st); cudaFreeAsync(jt, st); cudaFreeAsync(Ts, st); cudaFreeAsync(v, st); cudaFreeAsync(vrest, st); cudaFreeAsync(ids, st);
    return rc;
}

}  // extern "C"

// ---- live wire formats (host code, no CUDA) ---------------------------------------------------------------------
// live_detector.py:57-61 sends one UDP datagram per frame: four '#'-separated groups of comma-separated decimal numbers
// "uv(99)#ori(54)#acc(18)#RCM(9)"; live_server.py:42-45 parses it with float() per token; live_server.py:55-59 answers Unity with
// "%g,...(72 axis-angle)#%g,%g,%g$".
#include <stdlib.h>
#include <string.h>

namespace {
// parses up to `want` comma-separated numbers from [p, end); returns the count, -1 on a malformed token
int parse_group(const char* p, const char* end, float* out, int want) {
    int n = 0;
    while (p < end) {
        char* q = nullptr;
        const double v = strtod(p, &q);          // float(token) then .float(): double parse, one rounding to float32
        if (q == p) return -1;
        if (n < want) out[n] = (float)v;
        ++n;
        p = q;
        while (p < end && (*p == ' ' || *p == '\t')) ++p;
        if (p < end) {
            if (*p != ',') return -1;
            ++p;
        }
    }
    return n;
}
}  // namespace

extern "C" {

int rc_live_parse_frame(const char* h_text, int32_t len, float* h_uv, float* h_ori, float* h_acc, float* h_rcm) {
    RC_ARG(h_text && len > 0 && h_uv && h_ori && h_acc);
    std::string buf(h_text, h_text + len);       // strtod needs a terminated buffer
    const char* p = buf.c_str();
    const char* end = p + buf.size();
    float* dst[4] = {h_uv, h_ori, h_acc, h_rcm};
    const int want[4] = {99, 54, 18, 9};
    float scratch[9];
    for (int gidx = 0; gidx < 4; ++gidx) {
        const char* sep = (const char*)memchr(p, '#', (size_t)(end - p));
        const char* ge = sep ? sep : end;
        if ((gidx < 3) != (sep != nullptr)) { rc_set_error("rc_live_parse_frame: expected 4 '#'-separated groups"); return RC_ERR_ARG; }
        const int n = parse_group(p, ge, dst[gidx] ? dst[gidx] : scratch, want[gidx]);
        if (n != want[gidx]) { rc_set_error("rc_live_parse_frame: group %d has %d numbers, expected %d", gidx, n, want[gidx]); return RC_ERR_ARG; }
        p = ge + 1;
    }
    return RC_OK;
}

// live_demo_sync.py:262-268 (get_from_udp): the sensor server's binary datagram is float32[8 N] = t[N] | q[N,4] (wxyz) | a[N,3]
int rc_live_parse_imu_packet(const void* h_data, int32_t nbytes, int32_t n_imu, float* h_t, float* h_q, float* h_a) {
    RC_ARG(h_data && n_imu > 0 && h_t && h_q && h_a);
    if (nbytes != 32 * n_imu) { rc_set_error("rc_live_parse_imu_packet: %d bytes, expected %d for %d sensors", nbytes, 32 * n_imu, n_imu); return RC_ERR_ARG; }
    const unsigned char* p = (const unsigned char*)h_data;          // the datagram need not be aligned
    memcpy(h_t, p, (size_t)n_imu * 4);
    memcpy(h_q, p + (size_t)n_imu * 4, (size_t)n_imu * 16);
    memcpy(h_a, p + (size_t)n_imu * 20, (size_t)n_imu * 12);
    return RC_OK;
}

int rc_live_format_pose(const float* h_pose_aa, const float* h_tran, char* h_out, int32_t cap) {
    RC_ARG(h_pose_aa && h_tran && h_out && cap > 0);
    int pos = 0;
    for (int i = 0; i < 75; ++i) {
        const float v = i < 72 ? h_pose_aa[i] : h_tran[i - 72];
        const int w = snprintf(h_out + pos, (size_t)(cap - pos), "%g%s", (double)v, i == 71 ? "#" : (i == 74 ? "$" : ","));
        if (w < 0 || pos + w >= cap) { rc_set_error("rc_live_format_pose: buffer of %d bytes too small", cap); return RC_ERR_ARG; }
        pos += w;
    }
    return pos;                                   // bytes written (no terminator counted)
}

}  // extern "C"
