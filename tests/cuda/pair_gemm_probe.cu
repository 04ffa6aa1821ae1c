// GPU probe (not part of the product): minimal cta_group::2 GEMM — one CTA pair computes C[256,128] = A[256,K] W[128,K]^T with
// fp16 operands and fp32 accumulation in TMEM.  Validates, on the box, the 2-SM pieces the pair version of the grouped kernel
// relies on: tcgen05.alloc.cta_group::2 in both CTAs, TMA loads of the peer completing on the LEADER's mbarrier
// (cp.async.bulk.tensor ... .cta_group::2), tcgen05.mma.cta_group::2 with M = 256 (A rows and W rows split across the two CTAs'
// shared memory), tcgen05.commit with a cluster multicast mask, TMEM read-back in both CTAs.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tests/_build/pair_gemm_probe tests/cuda/pair_gemm_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../robustcap_b200/csrc/rc_tc_dev.cuh"

namespace {

constexpr int BK = 64, STAGES = 2;
constexpr int A_BYTES = 128 * BK * 2, W_BYTES = 64 * BK * 2, STAGE_BYTES = A_BYTES + W_BYTES;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, float* C, int K) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int KB = K / BK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[s]), 0);
                if (rank == 0) mbar_expect_tx(smem_u32(&bar_full[s]), 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                tma_load_2d_pair(base, &tmA, kb * BK, (int)rank * 128, full_leader);
                tma_load_2d_pair(base + A_BYTES, &tmW, kb * BK, (int)rank * 64, full_leader);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint64_t dA = make_desc(base), dW = make_desc(base + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)(k * 2);
                    tc_mma_f16_pair(tmem_base, dA + adv, dW + adv, idesc, (kb | k) ? 1u : 0u);
                }
                tc_commit_pair(smem_u32(&bar_empty[s]), 3);
            }
            tc_commit_pair(smem_u32(&bar_acc), 3);
        }
    } else {
        const int q = warp & 3;                                   // warps 2..5 -> TMEM lane quarters 2,3,0,1
        mbar_wait(smem_u32(&bar_acc), 0);
        tc_fence_after();
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row = (int)rank * 128 + q * 32 + lane;
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tc_ld32(lane_base + (uint32_t)(c * 32), v);
            tc_ld_wait();
            for (int e = 0; e < 32; ++e) C[(size_t)row * 128 + c * 32 + e] = __uint_as_float(v[e]);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* out, const void* base, long long rows, int K, int box_rows) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return 1;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return ((EncodeTiledFn)p)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

}  // namespace

int main() {
    const int M = 256, N = 128, K = 256;
    std::vector<__half> hA((size_t)M * K), hW((size_t)N * K);
    std::vector<float> fA((size_t)M * K), fW((size_t)N * K);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fA[i] = __half2float(hA[i]); }
    for (size_t i = 0; i < hW.size(); ++i) { hW[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fW[i] = __half2float(hW[i]); }
    __half *dA, *dW;
    float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&dC, (size_t)M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)M * N * 4);
    CUtensorMap mA, mW;
    if (make_map(&mA, dA, M, K, 128) || make_map(&mW, dW, N, K, 64)) { printf("tensor map failed\n"); return 2; }
    const int smem = STAGES * STAGE_BYTES + 1024;
    cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_gemm_kernel<<<2, 192, smem>>>(mA, mW, dC, K);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<float> hC((size_t)M * N);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad_r = -1, bad_c = -1;
    for (int r = 0; r < M; ++r)
        for (int c = 0; c < N; ++c) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)fA[(size_t)r * K + k] * fW[(size_t)c * K + k];
            const double d = fabs(acc - hC[(size_t)r * N + c]);
            if (!(d <= worst)) { worst = d; bad_r = r; bad_c = c; }
        }
    printf("pair gemm: max |C - ref| = %.3e at (%d, %d)  C[0,0]=%f C[128,64]=%f C[255,127]=%f\n", worst, bad_r, bad_c, hC[0], hC[128 * N + 64], hC[255 * N + 127]);
    // quadrant errors help to see which half (rows = CTA, columns = W half) is wrong
    for (int qr = 0; qr < 2; ++qr)
        for (int qc = 0; qc < 2; ++qc) {
            double w = 0;
            for (int r = qr * 128; r < qr * 128 + 128; ++r)
                for (int c = qc * 64; c < qc * 64 + 64; ++c) {
                    double acc = 0;
                    for (int k = 0; k < K; ++k) acc += (double)fA[(size_t)r * K + k] * fW[(size_t)c * K + k];
                    const double d = fabs(acc - hC[(size_t)r * N + c]);
                    if (!(d <= w)) w = d;
                }
            printf("  rows %3d.. cols %3d..: %.3e\n", qr * 128, qc * 64, w);
        }
    return worst < 1e-2 ? 0 : 1;
}
