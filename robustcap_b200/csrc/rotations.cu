// Rotation-representation conversion kernels (articulate/math/angular.py) + library-wide error plumbing.
// HBM-bound streaming kernels: each block stages 128 items through shared memory so that global loads and
// stores are fully coalesced although one thread converts one item.
#include <stdlib.h>
#include <stdarg.h>
#include "rc_common.cuh"
#include "rc_math.h"

std::atomic<long long> g_rc_launches{0};
bool rc_pdl_enabled() {
    static const bool on = getenv("RC_NO_PDL") == nullptr;
    return on;
}
static thread_local char g_err[512] = "";

void rc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* rc_version(void) { return "robustcap_b200 0.1 (sm_100a)"; }
extern "C" const char* rc_last_error(void) { return g_err; }
extern "C" int64_t rc_launch_count(void) { return (int64_t)g_rc_launches.load(); }

namespace {

constexpr int kItems = 128;   // items (= threads) per block

template <int IN, int IN2, int OUT, class Op>
__global__ void __launch_bounds__(kItems) rc_map_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         float* __restrict__ o, long long n, Op op) {
    constexpr int SI = IN | 1, SI2 = (IN2 > 0 ? IN2 : 1) | 1, SO = OUT | 1;   // odd strides: no bank conflicts
    __shared__ float sa[kItems * SI];
    __shared__ float sb[kItems * SI2];
    __shared__ float so[kItems * SO];
    const long long base = (long long)blockIdx.x * kItems;
    const int cnt = (int)min((long long)kItems, n - base);
    for (int e = threadIdx.x; e < cnt * IN; e += kItems) sa[(e / IN) * SI + e % IN] = a[base * IN + e];
    if constexpr (IN2 > 0) {
        for (int e = threadIdx.x; e < cnt * IN2; e += kItems) sb[(e / IN2) * SI2 + e % IN2] = b[base * IN2 + e];
    }
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
        float x[IN], y[IN2 > 0 ? IN2 : 1], r[OUT];
#pragma unroll
        for (int i = 0; i < IN; ++i) x[i] = sa[threadIdx.x * SI + i];
#pragma unroll
        for (int i = 0; i < IN2; ++i) y[i] = sb[threadIdx.x * SI2 + i];
        op(x, y, r);
#pragma unroll
        for (int i = 0; i < OUT; ++i) so[threadIdx.x * SO + i] = r[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * OUT; e += kItems) o[base * OUT + e] = so[(e / OUT) * SO + e % OUT];
}

struct OpR6dToMat { __device__ void operator()(const float* x, const float*, float* r) const { rc_r6d_to_mat(x, r); } };
struct OpMatToR6d { __device__ void operator()(const float* x, const float*, float* r) const { rc_mat_to_r6d(x, r); } };
struct OpAaToMat { __device__ void operator()(const float* x, const float*, float* r) const { rc_aa_to_mat(x, r); } };
struct OpMatToAa { __device__ void operator()(const float* x, const float*, float* r) const { rc_mat_to_aa(x, r); } };
struct OpRodrigues { __device__ void operator()(const float* x, const float*, float* r) const { rc_batch_rodrigues(x, r); } };
struct OpQuatToMat { __device__ void operator()(const float* x, const float*, float* r) const { rc_quat_to_mat(x, r); } };
struct OpQuatToAa { __device__ void operator()(const float* x, const float*, float* r) const { rc_quat_to_aa(x, r); } };
struct OpAaToQuat { __device__ void operator()(const float* x, const float*, float* r) const { rc_aa_to_quat(x, r); } };
struct OpQuatMul { __device__ void operator()(const float* x, const float* y, float* r) const { rc_quat_mul(x, y, r); } };

template <int IN, int IN2, int OUT, class Op>
int run_map(const float* a, const float* b, float* o, int64_t n, void* stream) {
    RC_ARG(n >= 0);
    if (n == 0) return RC_OK;
    RC_ARG(a != nullptr && o != nullptr && (IN2 == 0 || b != nullptr));
    RC_LAUNCH((rc_map_kernel<IN, IN2, OUT, Op>), rc_cdiv(n, kItems), kItems, 0, stream, a, b, o, (long long)n, Op());
    RC_CHECK_LAUNCH();
    return RC_OK;
}

}  // namespace

extern "C" {
int rc_r6d_to_rotmat(const float* i, float* o, int64_t n, void* s) { return run_map<6, 0, 9, OpR6dToMat>(i, nullptr, o, n, s); }
int rc_rotmat_to_r6d(const float* i, float* o, int64_t n, void* s) { return run_map<9, 0, 6, OpMatToR6d>(i, nullptr, o, n, s); }
int rc_axis_angle_to_rotmat(const float* i, float* o, int64_t n, void* s) { return run_map<3, 0, 9, OpAaToMat>(i, nullptr, o, n, s); }
int rc_rotmat_to_axis_angle(const float* i, float* o, int64_t n, void* s) { return run_map<9, 0, 3, OpMatToAa>(i, nullptr, o, n, s); }
int rc_batch_rodrigues(const float* i, float* o, int64_t n, void* s) { return run_map<3, 0, 9, OpRodrigues>(i, nullptr, o, n, s); }
int rc_quat_to_rotmat(const float* i, float* o, int64_t n, void* s) { return run_map<4, 0, 9, OpQuatToMat>(i, nullptr, o, n, s); }
int rc_quat_to_axis_angle(const float* i, float* o, int64_t n, void* s) { return run_map<4, 0, 3, OpQuatToAa>(i, nullptr, o, n, s); }
int rc_axis_angle_to_quat(const float* i, float* o, int64_t n, void* s) { return run_map<3, 0, 4, OpAaToQuat>(i, nullptr, o, n, s); }
int rc_quat_product(const float* a, const float* b, float* o, int64_t n, void* s) { return run_map<4, 4, 4, OpQuatMul>(a, b, o, n, s); }
}
