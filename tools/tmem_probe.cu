// Probe: TMEM read (tcgen05.ld) and write (tcgen05.st) bandwidth per SM, alone and while SS-mode MMAs run.
// The pre-filter top-k kernel and the conv epilogues are modelled on "TMEM read-out = 64 B/cycle/SM"; this pins it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tmem_probe tools/tmem_probe.cu && tools/_bin/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../audio_sheet_retrieval_b200/csrc/common.cuh"
using namespace asr;

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void st_x32(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void st_x8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: ld x32, one wait per load; 1: ld x32, two loads in flight; 2: st x32; 3: st x8 (an A-operand row group);
// mma_n > 0: warp `rw` streams SS MMAs of that N into columns 384.. meanwhile (and the reported MMA rate is for it)
template <int MODE>
__global__ void probe(long long *cycles, uint32_t *out, int rw, int iters, int mma_n, int mma_iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 * 1024) / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(smem)[i] = 0x3F80;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tptr;
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = tid + i;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < rw) {
        const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 64u;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0) {
                ld_x32(taddr, r);
                tmem_ld_wait();
                acc += r[0] ^ r[31];
            } else if (MODE == 1) {
                ld_x32(taddr, r);
                ld_x32(taddr + 32, r + 32);
                tmem_ld_wait();
                acc += r[0] ^ r[63];
            } else if (MODE == 2) {
                st_x32(taddr, r);
                st_wait();
            } else {
                st_x8(taddr, r);
                st_x8(taddr + 8, r + 8);
                st_x8(taddr + 16, r + 16);
                st_x8(taddr + 24, r + 24);
                st_wait();
            }
        }
        const long long t1 = clock64();
        if ((tid & 31) == 0) cycles[1 + warp] = t1 - t0;
    } else if (warp == rw && mma_n > 0) {
        if ((tid & 31) == 0) {
            const uint64_t ad0 = umma_desc(smem_u32(smem), 3072, 128);
            const uint64_t bd0 = umma_desc(smem_u32(smem + 48 * 1024), 2048, 128);
            const uint32_t idesc = umma_idesc_bf16(mma_n);
#pragma unroll 1
            for (int i = 0; i < mma_iters; ++i) mma_ss(base + 384, ad0 + (uint64_t)(i & 63), bd0, idesc, i > 0);
            tc_commit(&bar);
        }
        mbar_wait(&bar, 0);
        tc_fence_after();
        if ((tid & 31) == 0) cycles[0] = clock64() - t0;
    }
    out[tid] = acc + r[5];
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

template <int MODE>
static void run(long long *dc, uint32_t *dout, int rw, int mma_n, const char *name, int bytes_per_iter_per_warp) {
    const int iters = 4096, mma_iters = 8192;
    long long hc[17];
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaMemset(dc, 0, sizeof(hc));
    for (int it = 0; it < 2; ++it) {
        probe<MODE><<<1, 32 * (rw + 1), 64 * 1024>>>(dc, dout, rw, iters, mma_n, mma_iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < rw; ++w) mx = hc[1 + w] > mx ? hc[1 + w] : mx;
    printf("%-26s warps %2d  mma N=%3d : %8.1f B/cycle/SM (%6.1f cycles per warp-iteration)", name, rw, mma_n,
           (double)bytes_per_iter_per_warp * iters * rw / (double)mx, (double)mx / iters);
    if (mma_n > 0) printf("   MMA: %6.1f cycles each", (double)hc[0] / mma_iters);
    printf("\n");
}

int main() {
    long long *dc;
    uint32_t *dout;
    cudaMalloc(&dc, 17 * 8);
    cudaMalloc(&dout, 1024 * 4);
    for (int rw : {1, 4, 8, 16}) {
        run<0>(dc, dout, rw, 0, "ld x32, wait each", 32 * 32 * 4);
        run<1>(dc, dout, rw, 0, "ld 2 x32 in flight", 2 * 32 * 32 * 4);
        run<2>(dc, dout, rw, 0, "st x32", 32 * 32 * 4);
        run<3>(dc, dout, rw, 0, "st 4 x8", 32 * 32 * 4);
    }
    for (int n : {16, 48, 256}) {
        run<1>(dc, dout, 0, n, "MMA alone", 0);
        run<1>(dc, dout, 4, n, "ld 2 x32 + MMA", 2 * 32 * 32 * 4);
        run<1>(dc, dout, 8, n, "ld 2 x32 + MMA", 2 * 32 * 32 * 4);
        run<2>(dc, dout, 4, n, "st x32 + MMA", 32 * 32 * 4);
    }
    return 0;
}
