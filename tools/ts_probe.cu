// Probe: tcgen05.cp (shared -> TMEM) staging of the A operand and TS-mode tcgen05.mma (A from TMEM).
//  1. semantics: does cp.128x256b of the no-swizzle K-major A tile give the operand a TS-mode MMA expects?
//  2. cost: cycles per cp, per TS MMA (N = 16/32/48), and whether the two overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../audio_sheet_retrieval_b200/csrc/common.cuh"
using namespace asr;

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

constexpr int A_LBO = 3072;   // stride between the two 8-element K chunks of the A tile
constexpr int B_LBO = 1024;

// value of A[r][k] for the semantic check (exact in bf16)
__host__ __device__ inline float a_val(int r, int k) { return (float)(((r * 3 + k * 5) % 17) - 8); }

// mode 0: semantic check.  mode 1: TS MMAs only.  2: cps only.  3: per step 3 cps then 9 TS MMAs.  4: SS MMAs (9 per step)
__global__ void probe(long long *cycles, float *out, int mode, int n, int steps, int issuers) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    __nv_bfloat16 *sa = reinterpret_cast<__nv_bfloat16 *>(smem);
    __nv_bfloat16 *sb = reinterpret_cast<__nv_bfloat16 *>(smem + 48 * 1024);
    for (int i = tid; i < 32 * 1024; i += blockDim.x) sa[i] = __float2bfloat16(0.f);
    __syncthreads();
    for (int i = tid; i < 128 * 16; i += blockDim.x) {     // A tile in the canonical no-swizzle K-major layout
        const int r = i >> 4, k = i & 15;
        sa[((k >> 3) * A_LBO + (r >> 3) * 128 + (r & 7) * 16) / 2 + (k & 7)] = __float2bfloat16(a_val(r, k));
    }
    for (int i = tid; i < 64 * 16; i += blockDim.x) {      // B[nn][k] = (nn % 16 == k): D[:, nn] = A[:, nn % 16]
        const int nn = i >> 4, k = i & 15;
        sb[((k >> 3) * B_LBO + (nn >> 3) * 128 + (nn & 7) * 16) / 2 + (k & 7)] = __float2bfloat16((nn & 15) == k ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(&bar, mode == 0 ? 1 : issuers); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tptr;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    const uint32_t idesc = umma_idesc_bf16(n);
    const uint64_t ad0 = umma_desc(a0, A_LBO, 128);
    const uint64_t bd0 = umma_desc(b0, B_LBO, 128);
    long long t0 = clock64();
    if (mode == 0) {
        if (tid == 0) {
            mma_ss(base, ad0, bd0, idesc, 0);                 // D0 = cols 0..n-1
            cp_128x256b(base + 256, ad0);                     // A -> TMEM cols 256..263
            mma_ts(base + 64, base + 256, bd0, idesc, 0);     // D1 = cols 64..64+n-1
            tc_commit(&bar);
        }
    } else if ((tid & 31) == 0 && warp < issuers) {
        const uint32_t dbase = base + warp * 64;              // 64 accumulator columns per issuer (n <= 48 -> overlap is harmless)
        const uint32_t abase = base + 256 + warp * 32;        // 3 x 8 columns of staged A per issuer
        cp_128x256b(abase, ad0);
        cp_128x256b(abase + 8, ad0 + 1);
        cp_128x256b(abase + 16, ad0 + 2);
#pragma unroll 1
        for (int s = 0; s < steps; ++s) {
            const uint32_t acc = s > 0;
            const uint64_t ad = ad0 + (uint64_t)(s & 31) * 8;  // a different 128-row window every step
            if (mode == 2 || mode == 3) {
                cp_128x256b(abase, ad);
                cp_128x256b(abase + 8, ad + 1);
                cp_128x256b(abase + 16, ad + 2);
            }
            if (mode == 1 || mode == 3) {
#pragma unroll
                for (int t = 0; t < 9; ++t) mma_ts(dbase, abase + (t % 3) * 8, bd0 + (t & 1) * 8, idesc, acc | (t > 0));
            }
            if (mode == 4) {
#pragma unroll
                for (int t = 0; t < 9; ++t) mma_ss(dbase, ad + t, bd0 + (t & 1) * 8, idesc, acc | (t > 0));
            }
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (tid == 0) cycles[0] = clock64() - t0;
    if (mode == 0) {
        float v[16];
        tmem_ld16(base + ((uint32_t)(warp * 32) << 16), v);
        for (int j = 0; j < 16; ++j) out[tid * 32 + j] = v[j];
        tmem_ld16(base + 64 + ((uint32_t)(warp * 32) << 16), v);
        for (int j = 0; j < 16; ++j) out[tid * 32 + 16 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

int main() {
    long long *dc, hc;
    float *dout;
    static float h[128 * 32];
    cudaMalloc(&dc, 8);
    cudaMalloc(&dout, sizeof(h));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe<<<1, 128, 64 * 1024>>>(dc, dout, 0, 16, 1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("semantic check: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    int bad_ss = 0, bad_ts = 0;
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 16; ++k) {
            if (h[r * 32 + k] != a_val(r, k)) ++bad_ss;
            if (h[r * 32 + 16 + k] != a_val(r, k)) ++bad_ts;
        }
    printf("semantic check: SS mismatches %d, cp+TS mismatches %d of 2048\n", bad_ss, bad_ts);
    if (bad_ts) {
        for (int r : {0, 1, 8, 9, 32, 127}) {
            printf("  row %3d TS:", r);
            for (int k = 0; k < 16; ++k) printf(" %g", h[r * 32 + 16 + k]);
            printf("\n          want:");
            for (int k = 0; k < 16; ++k) printf(" %g", a_val(r, k));
            printf("\n");
        }
    }
    const int steps = 1024;
    const char *names[] = {"", "9 TS MMA / step        ", "3 cp / step            ", "3 cp + 9 TS MMA / step ", "9 SS MMA / step        "};
    for (int n : {16, 32, 48})
        for (int issuers : {1, 2, 4})
            for (int mode = 1; mode <= 4; ++mode) {
                for (int it = 0; it < 2; ++it) {
                    probe<<<1, 128, 64 * 1024>>>(dc, dout, mode, n, steps, issuers);
                    e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
                printf("N=%2d issuers %d %s: %8.1f cycles/step\n", n, issuers, names[mode], (double)hc / ((double)steps * issuers));
            }
    return 0;
}
