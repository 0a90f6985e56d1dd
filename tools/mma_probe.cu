// Probe: how many cycles does one SS-mode tcgen05.mma (M=128, K=16, bf16) occupy for small N, and does the
// A collector buffer (.collector::a::fill/use/lastuse) remove the shared-memory re-read of the A tile?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe tools/mma_probe.cu && gpurun_out/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../audio_sheet_retrieval_b200/csrc/common.cuh"
using namespace asr;

template <int OP>   // 0 none, 1 fill, 2 use, 3 lastuse
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (OP == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (OP == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (OP == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (OP == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: a different (shifted) A per MMA; 1: same A, no qualifier; 2: triples fill/use/lastuse, A changes per triple;
// 3: same as 2 but the three MMAs of a triple write three different accumulators (the conv use case);
// 4: distinct A, three accumulators round robin (baseline for 3)
template <int MODE>
__global__ void probe(long long *cycles, float *out, int n, int triples, int issuers) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // A region: 48 KB of bf16 ones; B region after it: 16 KB of bf16 ones
    for (int i = tid; i < (64 * 1024) / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(smem)[i] = 0x3F80;
    if (tid == 0) { mbar_init(&bar, issuers); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tptr;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    const uint32_t idesc = umma_idesc_bf16(n);
    long long t0 = clock64(), t1 = 0;
    if ((tid & 31) == 0 && warp < issuers) {
        // K=16 = two 8-element chunks; chunk stride (LBO) 3072 B, 8-row-group stride (SBO) 128 B
        const uint64_t ad0 = umma_desc(a0 + warp * 8192, 3072, 128);
        const uint64_t bd0 = umma_desc(b0, 2048, 128);
        const uint32_t base = tptr + warp * 128;
        const uint32_t dstep = (MODE == 3 || MODE == 4) ? 32u : 0u;
#pragma unroll 1
        for (int tr = 0; tr < triples; ++tr) {
            const uint64_t sh = (uint64_t)(tr & 63);                 // descriptor address field is in 16-byte units
            const uint32_t acc = tr > 0;
            if (MODE == 0 || MODE == 4) {
                mma<0>(base, ad0 + ((3 * sh) & 63), bd0, idesc, acc);
                mma<0>(base + dstep, ad0 + ((3 * sh + 1) & 63), bd0 + 256, idesc, acc);
                mma<0>(base + 2 * dstep, ad0 + ((3 * sh + 2) & 63), bd0 + 512, idesc, acc);
            } else if (MODE == 1) {
                mma<0>(base, ad0, bd0, idesc, acc);
                mma<0>(base, ad0, bd0 + 256, idesc, acc);
                mma<0>(base, ad0, bd0 + 512, idesc, acc);
            } else {
                mma<1>(base, ad0 + sh, bd0, idesc, acc);
                mma<2>(base + dstep, ad0 + sh, bd0 + 256, idesc, acc);
                mma<3>(base + 2 * dstep, ad0 + sh, bd0 + 512, idesc, acc);
            }
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    float v[16];
    tmem_ld16(base + ((uint32_t)(warp * 32) << 16), v);
    out[tid] = v[0];
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

template <int MODE>
static int run(long long *dc, float *dout, int n, int triples, int issuers, const char *name) {
    long long hc;
    float hout[128];
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int it = 0; it < 2; ++it) {
        probe<MODE><<<1, 128, 64 * 1024>>>(dc, dout, n, triples, issuers);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d n %d: %s\n", MODE, n, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost);
    printf("N=%3d issuers %d %s : %7.2f cycles/MMA   (D[0,0]=%g D[127,0]=%g)\n", n, issuers, name, (double)hc / (3.0 * triples * issuers), hout[0], hout[127]);
    return 0;
}

int main() {
    long long *dc;
    float *dout;
    cudaMalloc(&dc, 8);
    cudaMalloc(&dout, 128 * 4);
    const int triples = 2048;
    for (int n : {16, 32, 48, 96})
        for (int issuers : {1, 2, 4}) {
            if (run<0>(dc, dout, n, triples, issuers, "distinct A          ")) return 1;
            if (run<1>(dc, dout, n, triples, issuers, "same A, no qualifier")) return 1;
            if (run<2>(dc, dout, n, triples, issuers, "fill/use/lastuse    ")) return 1;
            if (run<3>(dc, dout, n, triples, issuers, "fill/use/lastuse 3 D")) return 1;
            if (run<4>(dc, dout, n, triples, issuers, "distinct A, 3 D     ")) return 1;
        }
    return 0;
}
