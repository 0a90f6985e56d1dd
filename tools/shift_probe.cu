// Probe the semantics of tcgen05.shift.cta_group::1.down on sm_100a: fill a TMEM region with a known
// pattern, shift, read back, print what moved.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../audio_sheet_retrieval_b200/csrc/common.cuh"
using namespace asr;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void probe(float *out, int lane_off, int col_off, int n_shifts) {
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tptr, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tptr;
    const uint32_t mine = base + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < 4; ++g) {
        float v[16];
        for (int j = 0; j < 16; ++j) v[j] = (float)((warp * 32 + lane) * 100 + g * 16 + j);
        tmem_st16(mine + g * 16, v);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (tid == 0) {
        for (int s = 0; s < n_shifts; ++s)
            asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(base + ((uint32_t)lane_off << 16) + (uint32_t)col_off) : "memory");
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int g = 0; g < 4; ++g) {
        float v[16];
        tmem_ld16(mine + g * 16, v);
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + g * 16 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 64); }
}

int main() {
    float *d, h[128 * 64];
    cudaMalloc(&d, sizeof(h));
    int cfgs[][3] = {{0, 0, 1}, {0, 8, 1}, {32, 0, 1}, {0, 0, 2}, {0, 16, 1}};
    for (auto &c : cfgs) {
        probe<<<1, 128>>>(d, c[0], c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        printf("== lane_off %d col_off %d shifts %d : %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        // report, per column, which lanes changed and by how many rows
        int first_changed_col = -1, last_changed_col = -1;
        for (int col = 0; col < 64; ++col) {
            int changed = 0;
            for (int l = 0; l < 128; ++l) if (h[l * 64 + col] != (float)(l * 100 + col)) ++changed;
            if (changed) { if (first_changed_col < 0) first_changed_col = col; last_changed_col = col; }
        }
        printf("   columns changed: %d..%d\n", first_changed_col, last_changed_col);
        if (first_changed_col >= 0) {
            int col = first_changed_col;
            printf("   col %d lanes: ", col);
            for (int l = 0; l < 128; ++l) {
                float v = h[l * 64 + col];
                int src = (int)((v - col) / 100.0f + 0.5f);
                if (src != l) printf("[%d<-%d] ", l, src);
            }
            printf("\n");
        }
    }
    return 0;
}
