// Shared helpers: error plumbing, launch counter, PTX wrappers (mbarrier, TMA bulk
// copies, tcgen05/TMEM).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/asr_b200.h"

namespace asr {

void set_error(const std::string &msg);
extern int64_t g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

#define ASR_CHECK_ARG(cond, msg)                                                     \
    do {                                                                             \
        if (!(cond)) {                                                               \
            asr::set_error(std::string(__func__) + ": " + (msg));                    \
            return ASR_ERR_ARG;                                                      \
        }                                                                            \
    } while (0)

#define ASR_CUDA(expr)                                                               \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            asr::set_error(std::string(__func__) + ": " #expr " -> " +               \
                           cudaGetErrorString(_e));                                  \
            return ASR_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

#define ASR_LAUNCH_CHECK()                                                           \
    do {                                                                             \
        asr::count_launch();                                                         \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) {                                                     \
            asr::set_error(std::string(__func__) + ": kernel launch -> " +           \
                           cudaGetErrorString(_e));                                  \
            return ASR_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

int ensure_device();   // returns ASR_OK or ASR_ERR_CUDA (no device / not sm_100)
int sm_count();
// per-device one-time state (kernel attributes, scratch) is kept in arrays indexed by the device ordinal
constexpr int ASR_MAX_DEVICES = 64;
inline int current_device() {
    int d = 0;
    return cudaGetDevice(&d) == cudaSuccess ? d : -1;
}
// Handles are bound to the device that was current at create.  The CUDA runtime's current device is per host thread
// (a new thread starts on device 0), so every entry point that takes a handle switches to the handle's device for the
// duration of the call and restores the caller's device afterwards.
// The previous device is only restored when this process already has a primary context there: since CUDA 12
// cudaSetDevice CREATES the context, and a fresh host thread "was" on device 0 without ever having used it -- restoring
// that would build a context on GPU 0 in every rank of a multi-GPU job (measured: 0.5-1 s and a few hundred MB, once
// per process, in the middle of the first call from a worker thread).
bool device_context_active(int dev);   // abi.cu (cuDevicePrimaryCtxGetState)
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        int cur = current_device();
        if (dev >= 0 && cur != dev) {
            prev = cur;
            cudaSetDevice(dev);
        }
    }
    ~DeviceGuard() {
        if (prev >= 0 && device_context_active(prev)) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------
// device-side PTX wrappers
// ------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // no suspend-time hint: a hint compiles to
        "selp.b32 %0, 1, 0, p;\n\t}"                                   // NANOSLEEP.SYNCS and adds wake-up latency
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box (that would be a strike); after ~4 s of
// wall time the kernel traps and the host sees a launch failure.
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (globaltimer_ns() - t0 > 4000000000ull) {
            printf("asr: mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// Relaxed wait for roles that run ahead of the critical path (a producer waiting for a free stage): try_wait with a
// suspend-time hint parks the warp in hardware instead of spinning, which gives its issue slots to the warps that work
// (at the price of some wake-up latency).  Same bounded-wait guarantee.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    for (;;) {
#pragma unroll 1
        for (int spin = 0; spin < 256; ++spin) {         // the clock is only read every 256 tries: the loop is 3 instructions
            uint32_t ok;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.b32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
                : "memory");
            if (ok) return;
        }
        if (globaltimer_ns() - t0 > 4000000000ull) {
            printf("asr: mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// the same with a caller-supplied tag in the timeout message (which role waited, at which step)
__device__ __forceinline__ void mbar_wait_tag(uint64_t *bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (globaltimer_ns() - t0 > 2000000000ull) {
            if ((threadIdx.x & 31) == 0)
                printf("asr: mbarrier timeout block %d warp %d parity %u tag %d\n", blockIdx.x, threadIdx.x >> 5, parity, tag);
            __trap();
        }
    }
}

// TMA bulk copy global -> shared (1-D), completion on an mbarrier.  size % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA tiled 2-D tensor load global -> shared.
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ bool elect_one() {   // one lane of a fully converged warp
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, predicated on `leader` (non-zero in exactly one lane): lets a fully converged warp run the
// issue loop with warp-uniform operands.  Descriptors are passed as their low words; the high
// word (SBO = 128 B, version 1, no swizzle) is the constant UMMA_DESC_HI.
constexpr uint32_t UMMA_DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ void tc_mma_bf16_pred(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
// The same issued by ONE elected lane of a fully converged warp, with the election INSIDE the asm block: ptxas then
// sees a tcgen05.mma guarded by elect.sync's own predicate and emits a single UTCHMMA on uniform registers (a handful
// of UIADD3 / UMOV per MMA when the operands derive from kernel parameters, blockIdx and loop counters).  With a
// predicate that went through a general register (tc_mma_bf16_pred) it must assume any subset of lanes and wraps every
// MMA in an elect / broadcast / branch sequence of ~17 instructions.
__device__ __forceinline__ void tc_mma_bf16_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_commit_pred(uint64_t *bar, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)), "r"(leader)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// two 8-column groups at taddr and taddr + stride in one round trip (thread = lane = row)
__device__ __forceinline__ void tmem_ld8x2(uint32_t taddr, uint32_t stride, float *v) {
    uint32_t r[16];
#pragma unroll
    for (int g = 0; g < 2; ++g)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[8 * g]), "=r"(r[8 * g + 1]), "=r"(r[8 * g + 2]), "=r"(r[8 * g + 3]), "=r"(r[8 * g + 4]),
                       "=r"(r[8 * g + 5]), "=r"(r[8 * g + 6]), "=r"(r[8 * g + 7])
                     : "r"(taddr + (uint32_t)g * stride)
                     : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// four 8-column groups at taddr + {0, 1, 2, 3} * stride in one round trip (thread = lane = row)
__device__ __forceinline__ void tmem_ld8x4(uint32_t taddr, uint32_t stride, float *v) {
    uint32_t r[32];
#pragma unroll
    for (int g = 0; g < 4; ++g)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[8 * g]), "=r"(r[8 * g + 1]), "=r"(r[8 * g + 2]), "=r"(r[8 * g + 3]), "=r"(r[8 * g + 4]),
                       "=r"(r[8 * g + 5]), "=r"(r[8 * g + 6]), "=r"(r[8 * g + 7])
                     : "r"(taddr + (uint32_t)g * stride)
                     : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 12 consecutive columns (8 + 4) in one round trip: the 12 channels of one layer-0 output row
__device__ __forceinline__ void tmem_ld12(uint32_t taddr, float *v) {
    uint32_t r[12];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11])
                 : "r"(taddr + 8u)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = __uint_as_float(r[i]);
}

// the same without the wait (valid after tmem_ld_wait())
__device__ __forceinline__ void tmem_ld12_issue(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11])
                 : "r"(taddr + 8u)
                 : "memory");
}

// shared-memory progress counters (producer: red.release after its stores; consumer: ld.acquire poll)
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_shared_add(uint32_t *p, uint32_t v) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 64 consecutive fp32 columns -> 64 registers per thread (one round trip instead of four)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float *v) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait: the registers are valid only after tmem_ld_wait().  Lets the scan of one 64-column
// chunk overlap the TMEM round trip of the next (issue order: wait(c), issue(c + 1), compute(c), wait(c + 1), ...).
__device__ __forceinline__ void tmem_ld64_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 columns without the wait (valid after tmem_ld_wait())
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// volatile shared-memory accesses for the hand-rolled queues (no caching in registers, no reordering by the compiler)
__device__ __forceinline__ uint32_t lds_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long lds_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_volatile_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle ("interleaved" core matrices of
// 8 rows x 16 bytes):  element (row r, k) lives at
//     start + (r/8)*sbo + (r%8)*16 + (k/8)*lbo + (k%8)*2      (bf16)
// (cute::UMMA::SmemDescriptor, version 1 = Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor for kind::f16: bf16 A/B (K-major both), fp32 D, M = 128.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ELU with the raw MUFU.EX2 (ex2.approx.ftz): 2^(x*log2e) - 1 for x <= 0.  Without -use_fast_math
// __expf expands to a denormal-safe sequence of ~9 predicated instructions per value; flushing
// to zero is exact enough here (exp(x) < 2^-126 => elu = -1).
__device__ __forceinline__ float ex2_approx(float t) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e;
}
__device__ __forceinline__ float elu_f(float x) {
    const float e = ex2_approx(x * 1.4426950408889634f);
    return x > 0.f ? x : (e - 1.0f);
}

#endif  // __CUDACC__
}  // namespace asr
