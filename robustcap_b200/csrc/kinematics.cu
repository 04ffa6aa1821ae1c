// Kinematic-tree kernels (articulate/math/spatial.py) and the SMPL body model (articulate/model.py:209-241).
// All HBM-bound: frames are staged through shared memory so global traffic is coalesced; one thread walks one
// frame's tree (depth <= 9 for SMPL, 23 dependent 3x3 / 3x4 products), the mesh is skinned one vertex per thread.
#include <vector>
#include "rc_common.cuh"
#include "rc_rows.h"
#include "rc_model.cuh"

namespace {

constexpr int kMaxJ = 64;
struct Parents { int nj; int p[kMaxJ]; };

enum TreeOp { FK_R, IK_R, FK_T, IK_T, BONE2JOINT, JOINT2BONE };

template <int OP> struct Elem { static constexpr int E = (OP == FK_R || OP == IK_R) ? 9 : ((OP == FK_T || OP == IK_T) ? 16 : 3); };

__device__ __forceinline__ void mat4_mul(const float* a, const float* b, float* c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            c[i * 4 + j] = a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j] + a[i * 4 + 2] * b[2 * 4 + j] + a[i * 4 + 3] * b[3 * 4 + j];
}
// spatial.py:90-101
__device__ __forceinline__ void mat4_rigid_inverse(const float* T, float* o) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) o[i * 4 + j] = T[j * 4 + i];
        o[i * 4 + 3] = -(T[0 * 4 + i] * T[0 * 4 + 3] + T[1 * 4 + i] * T[1 * 4 + 3] + T[2 * 4 + i] * T[2 * 4 + 3]);
    }
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

template <int OP>
__global__ void __launch_bounds__(32) rc_tree_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      Parents P, long long b) {
    constexpr int E = Elem<OP>::E;
    extern __shared__ float sm[];
    const int W = P.nj * E, S = W | 1;
    float* sin_ = sm;
    float* sout = sm + 32 * S;
    const long long base = (long long)blockIdx.x * 32;
    const int cnt = (int)min(32LL, b - base);
    for (int e = threadIdx.x; e < cnt * W; e += 32) sin_[(e / W) * S + e % W] = in[base * W + e];
    __syncwarp();
    if ((int)threadIdx.x < cnt) {
        const float* x = sin_ + threadIdx.x * S;
        float* y = sout + threadIdx.x * S;
        for (int e = 0; e < E; ++e) y[e] = x[e];
        for (int i = 1; i < P.nj; ++i) {
            const int p = P.p[i];
            if (OP == FK_R) rc_mat3_mul(y + p * 9, x + i * 9, y + i * 9);
            else if (OP == IK_R) rc_mat3_tmul(x + p * 9, x + i * 9, y + i * 9);
            else if (OP == FK_T) mat4_mul(y + p * 16, x + i * 16, y + i * 16);
            else if (OP == IK_T) { float inv[16]; mat4_rigid_inverse(x + p * 16, inv); mat4_mul(inv, x + i * 16, y + i * 16); }
            else if (OP == BONE2JOINT) { for (int r = 0; r < 3; ++r) y[i * 3 + r] = RC_ADD(y[p * 3 + r], x[i * 3 + r]); }
            else { for (int r = 0; r < 3; ++r) y[i * 3 + r] = RC_ADD(-x[p * 3 + r], x[i * 3 + r]); }
        }
    }
    __syncwarp();
    for (int e = threadIdx.x; e < cnt * W; e += 32) out[base * W + e] = sout[(e / W) * S + e % W];
}

template <int OP>
int run_tree(const float* in, float* out, const int32_t* parent, int nj, int64_t b, void* stream) {
    RC_ARG(nj >= 1 && nj <= kMaxJ && parent != nullptr && b >= 0);
    if (b == 0) return RC_OK;
    RC_ARG(in != nullptr && out != nullptr);
    Parents P;
    P.nj = nj;
    for (int i = 0; i < nj; ++i) {
        P.p[i] = parent[i];
        RC_ARG(i == 0 || (parent[i] >= 0 && parent[i] < i));
    }
    const int S = (nj * Elem<OP>::E) | 1;
    const size_t smem = (size_t)2 * 32 * S * sizeof(float);
    RC_CUDA(cudaFuncSetAttribute(rc_tree_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RC_LAUNCH((rc_tree_kernel<OP>), rc_cdiv(b, 32), 32, smem, stream, in, out, P, (long long)b);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

// ---- SMPL forward kinematics -----------------------------------------------------------------------------------
// One thread per frame walks the tree; writes global rotations, joints (+tran) and, for the mesh, the skinning
// transforms T'_j = [R_j | t_j - R_j jrest_j] (model.py:235) to a [b,24,12] scratch.
constexpr int kChainFrames = 16;   // frames per block (shared memory: 16 * (217 + 289) floats = 32 KB)
__global__ void __launch_bounds__(32) rc_smpl_chain_kernel(const RcModelConst* __restrict__ M,
                                                            const float* __restrict__ pose, const float* __restrict__ tran,
                                                            const float* __restrict__ jrest_b, long long b,
                                                            float* __restrict__ Rg, float* __restrict__ joint,
                                                            float* __restrict__ Tskin) {
    __shared__ float sp[kChainFrames * 217];
    __shared__ float sg[kChainFrames * 289];
    const long long base = (long long)blockIdx.x * kChainFrames;
    const int cnt = (int)min((long long)kChainFrames, b - base);
    for (int e = threadIdx.x; e < cnt * 216; e += 32) sp[(e / 216) * 217 + e % 216] = pose[base * 216 + e];
    __syncwarp();
    if ((int)threadIdx.x < cnt) {
        const long long f = base + threadIdx.x;
        const float* P = sp + threadIdx.x * 217;
        float* G = sg + threadIdx.x * 289;
        float jr[RC_NJ][3];
        for (int i = 0; i < RC_NJ; ++i)
            for (int r = 0; r < 3; ++r) jr[i][r] = jrest_b ? jrest_b[f * 72 + i * 3 + r] : M->jrest[i][r];
        for (int i = 0; i < RC_NJ; ++i) {
            const int p = M->parent[i];
            float L[12];
            for (int r = 0; r < 3; ++r) {
                L[r * 4 + 0] = P[i * 9 + r * 3 + 0]; L[r * 4 + 1] = P[i * 9 + r * 3 + 1]; L[r * 4 + 2] = P[i * 9 + r * 3 + 2];
                L[r * 4 + 3] = (i == 0) ? jr[0][r] : RC_ADD(-jr[p][r], jr[i][r]);       // bone vector (spatial.py:148-167)
            }
            if (i == 0) for (int e = 0; e < 12; ++e) G[e] = L[e];
            else rc_rigid_mul(G + p * 12, L, G + i * 12);
        }
        for (int i = 0; i < RC_NJ; ++i)
            for (int r = 0; r < 3; ++r) {
                const float t = G[i * 12 + r * 4 + 3];
                sp[threadIdx.x * 217 + i * 3 + r] = tran ? RC_ADD(t, tran[f * 3 + r]) : t;   // reuse sp for the joints
                const float d = G[i * 12 + r * 4 + 0] * jr[i][0] + G[i * 12 + r * 4 + 1] * jr[i][1] + G[i * 12 + r * 4 + 2] * jr[i][2];
                G[i * 12 + r * 4 + 3] = RC_SUB(t, d);
            }
    }
    __syncwarp();
    for (int e = threadIdx.x; e < cnt * 72; e += 32) joint[base * 72 + e] = sp[(e / 72) * 217 + e % 72];
    for (int e = threadIdx.x; e < cnt * 216; e += 32) {
        const int f = e / 216, q = e % 216, i = q / 9, r = (q % 9) / 3, c = q % 3;
        Rg[base * 216 + e] = sg[f * 289 + i * 12 + r * 4 + c];
    }
    if (Tskin)
        for (int e = threadIdx.x; e < cnt * 288; e += 32) Tskin[base * 288 + e] = sg[(e / 288) * 289 + e % 288];
}

// Linear blend skinning (model.py:236-241): v' = (sum_j w_vj T'_j) [v; 1] + tran.  One vertex per thread, the
// 24 weights stay in registers while the block sweeps kFrames frames whose transforms sit in shared memory.
constexpr int kLbsFrames = 8;
__global__ void __launch_bounds__(128) rc_lbs_kernel(const float* __restrict__ Tskin, const float* __restrict__ vrest,
                                                      const float* __restrict__ vrest_b, const float* __restrict__ W,
                                                      const float* __restrict__ tran, int nv, long long b,
                                                      float* __restrict__ vert) {
    __shared__ float sT[kLbsFrames * 288];
    const long long f0 = (long long)blockIdx.y * kLbsFrames;
    const int nf = (int)min((long long)kLbsFrames, b - f0);
    for (int e = threadIdx.x; e < nf * 288; e += 128) sT[e] = Tskin[f0 * 288 + e];
    __syncthreads();
    const int v = blockIdx.x * 128 + threadIdx.x;
    if (v >= nv) return;
    float w[RC_NJ];
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)v * RC_NJ);
#pragma unroll
    for (int q = 0; q < 6; ++q) { float4 t = __ldg(w4 + q); w[q * 4] = t.x; w[q * 4 + 1] = t.y; w[q * 4 + 2] = t.z; w[q * 4 + 3] = t.w; }
    float x = 0.f, y = 0.f, z = 0.f;
    if (!vrest_b) { x = vrest[v * 3]; y = vrest[v * 3 + 1]; z = vrest[v * 3 + 2]; }
    for (int f = 0; f < nf; ++f) {
        if (vrest_b) { const float* p = vrest_b + ((f0 + f) * nv + v) * 3; x = p[0]; y = p[1]; z = p[2]; }
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
        for (int j = 0; j < RC_NJ; ++j) {
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = fmaf(w[j], sT[f * 288 + j * 12 + e], T[e]);
        }
        float o[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            o[r] = T[r * 4] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3];
            if (tran) o[r] = RC_ADD(o[r], tran[(f0 + f) * 3 + r]);
        }
        float* dst = vert + ((f0 + f) * nv + v) * 3;
        dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
}

__global__ void __launch_bounds__(64) rc_keypoints_kernel(const RcModelConst* __restrict__ M, const float* __restrict__ pose,
                                                           const float* __restrict__ tran, long long b,
                                                           float* __restrict__ joint, float* __restrict__ kp) {
    const long long f = (long long)blockIdx.x * 64 + threadIdx.x;
    if (f >= b) return;
    float P[216], t[3] = {0.f, 0.f, 0.f}, J[72], K[99];
    for (int e = 0; e < 216; ++e) P[e] = pose[f * 216 + e];
    if (tran) for (int r = 0; r < 3; ++r) t[r] = tran[f * 3 + r];
    rc_fk_keypoints(*M, P, t, J, K);
    if (joint) for (int e = 0; e < 72; ++e) joint[f * 72 + e] = J[e];
    for (int e = 0; e < 99; ++e) kp[f * 99 + e] = K[e];
}

}  // namespace

int rc_smpl_chain_launch(const rc_model* m, const float* pose, long long b, float* Rg, float* joint, float* Tskin, void* stream) {
    RC_LAUNCH(rc_smpl_chain_kernel, rc_cdiv(b, kChainFrames), 32, 0, stream, m->d_const, pose, (const float*)nullptr, (const float*)nullptr, b, Rg, joint, Tskin);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

int rc_smpl_chain_launch_ex(const rc_model* m, const float* pose, const float* tran, const float* jrest_b, long long b, float* Rg,
                            float* joint, float* Tskin, void* stream) {
    RC_LAUNCH(rc_smpl_chain_kernel, rc_cdiv(b, kChainFrames), 32, 0, stream, m->d_const, pose, tran, jrest_b, b, Rg, joint, Tskin);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

extern "C" {

int rc_tree_fk_R(const float* i, float* o, const int32_t* p, int nj, int64_t b, void* s) { return run_tree<FK_R>(i, o, p, nj, b, s); }
int rc_tree_ik_R(const float* i, float* o, const int32_t* p, int nj, int64_t b, void* s) { return run_tree<IK_R>(i, o, p, nj, b, s); }
int rc_tree_fk_T(const float* i, float* o, const int32_t* p, int nj, int64_t b, void* s) { return run_tree<FK_T>(i, o, p, nj, b, s); }
int rc_tree_ik_T(const float* i, float* o, const int32_t* p, int nj, int64_t b, void* s) { return run_tree<IK_T>(i, o, p, nj, b, s); }
int rc_tree_bone_to_joint(const float* i, float* o, const int32_t* p, int nj, int64_t b, void* s) { return run_tree<BONE2JOINT>(i, o, p, nj, b, s); }
int rc_tree_joint_to_bone(const float* i, float* o, const int32_t* p, int nj, int64_t b, void* s) { return run_tree<JOINT2BONE>(i, o, p, nj, b, s); }

int rc_model_create(rc_model** out, const float* joints, const float* verts, const float* skin_w, int32_t nv,
                    const int32_t* parent, const int32_t* mp_mask) {
    RC_ARG(out && joints && verts && skin_w && parent && mp_mask && nv > 0);
    rc_model* m = new rc_model();
    m->nv = nv;
    RcModelConst& C = m->host;
    for (int i = 0; i < RC_NJ; ++i) {
        C.parent[i] = (i == 0) ? -1 : parent[i];
        if (i > 0 && (parent[i] < 0 || parent[i] >= i)) { delete m; rc_set_error("parent[%d]=%d invalid", i, parent[i]); return RC_ERR_ARG; }
        for (int r = 0; r < 3; ++r) C.jrest[i][r] = joints[i * 3 + r];
    }
    for (int i = 0; i < RC_NJ; ++i)
        for (int r = 0; r < 3; ++r) C.bone[i][r] = (i == 0) ? C.jrest[0][r] : (-C.jrest[C.parent[i]][r] + C.jrest[i][r]);
    C.max_depth = 0;
    for (int i = 0; i < RC_NJ; ++i) {
        C.depth[i] = (i == 0) ? 0 : C.depth[C.parent[i]] + 1;
        if (C.depth[i] > C.max_depth) C.max_depth = C.depth[i];
    }
    // net/sig_mp.py:287-299: rows 11-16 <- joints 16-21, 23-24 <- 1-2, 25-26 <- 4-5, 27-28 <- 7-8
    for (int k = 0; k < RC_NKP; ++k) {
        int j = -1;
        if (k >= 11 && k <= 16) j = 16 + (k - 11);
        else if (k == 23 || k == 24) j = 1 + (k - 23);
        else if (k == 25 || k == 26) j = 4 + (k - 25);
        else if (k == 27 || k == 28) j = 7 + (k - 27);
        C.kp_is_joint[k] = j >= 0;
        C.kp_index[k] = j >= 0 ? j : mp_mask[k];
        const int v = mp_mask[k];
        if (v < 0 || v >= nv) { delete m; rc_set_error("mp_mask[%d]=%d out of range", k, v); return RC_ERR_ARG; }
        for (int r = 0; r < 3; ++r) C.kp_rest[k][r] = verts[v * 3 + r];
        for (int q = 0; q < RC_NJ; ++q) C.kp_w[k][q] = skin_w[(size_t)v * RC_NJ + q];
    }
    cudaError_t e = cudaMalloc(&m->d_const, sizeof(RcModelConst));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_const, &C, sizeof(RcModelConst), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_verts, (size_t)nv * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_verts, verts, (size_t)nv * 3 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_skin_w, (size_t)nv * RC_NJ * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_skin_w, skin_w, (size_t)nv * RC_NJ * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        rc_set_error("rc_model_create: %s", cudaGetErrorString(e));
        rc_model_destroy(m);
        return RC_ERR_CUDA;
    }
    *out = m;
    return RC_OK;
}

void rc_model_destroy(rc_model* m) {
    if (!m) return;
    cudaFree(m->d_const); cudaFree(m->d_verts); cudaFree(m->d_skin_w);
    delete m;
}

int rc_model_forward_kinematics(const rc_model* m, const float* pose, const float* tran, const float* jrest_b,
                                const float* vrest_b, int64_t b, float* Rg, float* joint, float* vert, void* stream) {
    RC_ARG(m && b >= 0);
    if (b == 0) return RC_OK;
    RC_ARG(pose && Rg && joint);
    float* Tskin = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (vert) RC_CUDA(cudaMallocAsync(&Tskin, (size_t)b * 288 * sizeof(float), st));
    RC_LAUNCH(rc_smpl_chain_kernel, rc_cdiv(b, kChainFrames), 32, 0, stream, m->d_const, pose, tran, jrest_b, (long long)b, Rg, joint, Tskin);
    RC_CHECK_LAUNCH();
    if (vert) {
        dim3 grid(rc_cdiv(m->nv, 128), rc_cdiv(b, kLbsFrames));
        RC_LAUNCH(rc_lbs_kernel, grid, 128, 0, stream, Tskin, m->d_verts, vrest_b, m->d_skin_w, tran, m->nv, (long long)b, vert);
        RC_CHECK_LAUNCH();
        RC_CUDA(cudaFreeAsync(Tskin, st));
    }
    return RC_OK;
}

int rc_model_keypoints(const rc_model* m, const float* pose, const float* tran, int64_t b, float* joint, float* kp, void* stream) {
    RC_ARG(m && b >= 0);
    if (b == 0) return RC_OK;
    RC_ARG(pose && kp);
    RC_LAUNCH(rc_keypoints_kernel, rc_cdiv(b, 64), 64, 0, stream, m->d_const, pose, tran, (long long)b, joint, kp);
    RC_CHECK_LAUNCH();
    return RC_OK;
}

}  // extern "C"
