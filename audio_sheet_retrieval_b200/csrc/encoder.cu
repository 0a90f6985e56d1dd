// The two encoder branches: prepare + 8 x (conv3x3 + BN + ELU [+ 2x2 max-pool]) + 1x1 conv + BN +
// global mean + CCA projection + length norm.
//
// Replaces the compiled Theano graph of (reference, paths relative to its root)
//   audio_sheet_retrieval/models/mutopia_ccal_cont_rsz.py:54-147,170-190
//   audio_sheet_retrieval/models/mutopia_ccal_cont.py:54-147,170-190
//   audio_sheet_retrieval/models/lasagne_extensions/layers/cca.py:184-203 (deterministic) and :39-40
//   audio_sheet_retrieval/retrieval_wrapper.py:33-38,47-77 ; audio_sheet_retrieval/refine_cca.py:86-89
//
// Product path (ASR_PATH_TCGEN05)
//   layer 0 (Cin = 1) is a banded-Toeplitz GEMM on tcgen05 fused with `prepare` (l0_tc_kernel; the
//   CUDA-core l0_conv_kernel remains for other channel counts) and writes bf16 activations in the
//   "P8" layout below; layers 1..7 are implicit GEMMs on tcgen05 with fp32 accumulators in TMEM --
//   conv3x3_rows_kernel (four output rows stacked along N, pooling in registers) for the wide pooled
//   layers, conv3x3_tc_kernel (raster tiles) for the rest; BatchNorm is folded into the bf16 weights
//   + an fp32 bias, ELU and the 2x2 max-pool run in the epilogue; the head stays in fp32.
//
// P8 activation layout (bf16):  [sample][channel chunk of 8][(H+2) x (W+2) padded positions][8]
//   - one padded position of one chunk = 16 bytes = one row of a UMMA "core matrix"; a run of
//     positions is therefore directly a K-major, non-swizzled A operand (SBO = 128 B between
//     8-row groups, LBO = plane stride between the two 8-channel halves of a K = 16 step);
//   - a band of image rows is contiguous per plane, so it is fetched with ONE TMA bulk copy per
//     plane, and the 9 filter taps are 9 *shifted descriptors* over the same shared-memory tile
//     (shift = (dy*(W+2) + dx) * 16 bytes): the im2col matrix is never materialised;
//   - the GEMM M dimension enumerates padded positions of the band in raster order, so the two
//     border columns per row produce garbage rows that the epilogue drops (2/(W+2) waste);
//   - borders are zero (written once at create) = the convolution's zero padding.
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"

namespace asr {

typedef __nv_bfloat16 bf16;

struct LayerGeom {
    int cin, cout, cinp, coutp;
    int H, W;        // conv spatial size (input = output, pad 1)
    int pool;        // 2x2 max-pool after this layer
    int Ho, Wo;      // spatial size of the stored output
};

// launch plan of one tcgen05 conv layer
struct ConvPlan {
    int TH, bands, MT, n_stages, slot_cols, n_slots, tmem_cols;
    int sps;            // smem plane stride (bytes)
    int stage_bytes;    // 16 (front guard) + KC*sps + tail slack
    int wbytes, staging_bytes, smem_bytes;
    int off_bias, off_stage, off_staging, off_bar;
    unsigned wp_magic, wo_magic;
    int Hc;
    int rows, JT;       // rows != 0: conv3x3_rows_kernel (4 output rows stacked along N)
    int SPT, SP, RP;
};

struct ConvParams {
    const bf16 *in;
    bf16 *out;
    const bf16 *wblob;   // [9][KC][NP][8] bf16 followed by NP fp32 biases
    int n_samples;
    int H, W, Wp, Hp, KC, NP, NCH, TH, bands, MT, pool, cout;
    int Ho, Wo, Wpo;
    long long in_plane, in_sample, out_plane, out_sample;   // bytes
    int sps, stage_bytes, n_stages, slot_cols, n_slots, tmem_cols, wbytes;
    int off_bias, off_stage, off_staging, off_bar;
    unsigned wp_magic;   // ceil(2^32 / Wp): o / Wp == __umulhi(o, wp_magic) for every o the kernel sees
    unsigned wo_magic;   // same for the pooled width Wo (0 when Wo == 1)
    int Hc;              // conv rows that are needed (H, or 2*Ho for a pooled layer)
    int JT;              // row-stacked kernel: 128-pixel tiles across one image row
    int coop;            // row-stacked kernel: all epilogue groups drain every tile together
    int SPT, SP, RP;     // row-stacked kernel: samples per tile row, per-sample pitch (even), smem row pitch (positions)
    unsigned sp_magic;   // ceil(2^32 / SP)
    int KCL;             // input chunks that hold real channels: the others are never loaded (zeroed in smem once)
    int NCHR;            // output chunks that hold real channels: the others are never stored (buffers are pre-zeroed)
    int dbg;             // diagnostics only (env ASR_CONV_DEBUG): 1 = epilogue releases slots without draining, 2 = no MMAs
    int pair;            // raster kernel: K pairing of a half-used last chunk (KCL == KC - 1), see the MMA issuers
};

constexpr int N_MMA_WARPS = 4;           // tile tc is issued by MMA warp (tc & 3) and drained by epilogue group (tc & 3)
constexpr int N_EPI_GROUPS = 4;
constexpr int EPI_WARP0 = 1 + N_MMA_WARPS;
constexpr int EPI_THREADS = N_EPI_GROUPS * 128;
constexpr int CONV_THREADS = 32 * EPI_WARP0 + EPI_THREADS;   // warp 0 TMA, warps 1..4 MMA, warps 5..20 epilogue
constexpr int TAIL_SLACK = 2080;
constexpr int MAX_SLOTS = 8;
constexpr int MAX_STAGES = 6;          // input ring of the raster kernel (2 for bands that are big, more for small ones)
constexpr int SMEM_LIMIT = 227 * 1024;

// --------------------------------------------------------------------------------------
// tcgen05 implicit-GEMM convolution (layers 1..7)
// --------------------------------------------------------------------------------------
// 128-position tiles of the band that starts at conv row y0 (the last band of a sample may be short)
__device__ __forceinline__ int band_tiles(const ConvParams &p, int y0) {
    return (min(p.TH, p.Hc - y0) * p.Wp + 127) >> 7;
}

template <int KPAIRS>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv3x3_tc_kernel(const ConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // warp index through a lane-0 broadcast: the compiler then knows the role branches are warp-uniform and keeps the
    // MMA issuers' descriptor arithmetic on the uniform datapath (see tc_mma_bf16_elect)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    uint8_t *w_sm = smem;
    float *bias_sm = reinterpret_cast<float *>(smem + p.off_bias);
    uint8_t *stage_sm = smem + p.off_stage;
    uint8_t *staging_sm = smem + p.off_staging;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *w_full = bars;                                        // 1
    uint64_t *in_full = bars + 1;                                   // [MAX_STAGES]
    uint64_t *in_empty = bars + 1 + MAX_STAGES;                     // [MAX_STAGES]
    uint64_t *acc_full = bars + 1 + 2 * MAX_STAGES;                 // [MAX_SLOTS]
    uint64_t *acc_empty = bars + 1 + 2 * MAX_STAGES + MAX_SLOTS;    // [MAX_SLOTS]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 1 + 2 * MAX_STAGES + 2 * MAX_SLOTS);

    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], N_MMA_WARPS); }
        for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    if (p.KCL < p.KC) {     // planes of all-padding input chunks: zero once, never loaded
        for (int st = 0; st < p.n_stages; ++st) {
            uint4 *z = reinterpret_cast<uint4 *>(smem + p.off_stage + (size_t)st * p.stage_bytes + 16 + (size_t)p.KCL * p.sps);
            const int n16 = (p.stage_bytes - 16 - p.KCL * p.sps) / 16;
            for (int i = tid; i < n16; i += CONV_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_items = p.n_samples * p.bands;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)(p.wbytes + p.NP * 4));
            tma_bulk_g2s(w_sm, p.wblob, (uint32_t)p.wbytes, w_full);
            tma_bulk_g2s(bias_sm, reinterpret_cast<const uint8_t *>(p.wblob) + p.wbytes, (uint32_t)(p.NP * 4), w_full);
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int n = item / p.bands, y0 = (item % p.bands) * p.TH;
                const int s = it % p.n_stages;
                const uint32_t ph = (uint32_t)((it / p.n_stages) & 1);
                mbar_wait(&in_empty[s], ph ^ 1u);
                const int rows_in = min(p.TH + 2, p.Hp - y0);
                const uint32_t bytes = (uint32_t)(rows_in * p.Wp * 16);
                mbar_expect_tx(&in_full[s], bytes * (uint32_t)p.KCL);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.in) + (long long)n * p.in_sample +
                                     (long long)y0 * p.Wp * 16;
                uint8_t *dst = stage_sm + (size_t)s * p.stage_bytes + 16;
                for (int kc = 0; kc < p.KCL; ++kc)
                    tma_bulk_g2s(dst + (size_t)kc * p.sps, src + (long long)kc * p.in_plane, bytes, &in_full[s]);
            }
        }
    } else if (warp < EPI_WARP0) {
        // ================= MMA issuers (4 warps, tile tc belongs to warp tc & 3) =================
        const uint32_t my = (uint32_t)(warp - 1);
        // The whole warp walks the (uniform, fully unrolled) issue sequence so that descriptors live in
        // uniform registers; the tcgen05 instructions themselves are issued by one elected lane (elect.sync inside
        // the asm block).  Descriptors advance by adding to their low word (start-address field, 16-byte units).
        const bool no_mma = (p.dbg & 2) != 0;
        const uint32_t idesc = umma_idesc_bf16(p.NP);
        const uint32_t a_lbo = ((uint32_t)p.sps >> 4) << 16;                   // LBO = smem plane stride
        const uint32_t w_lo = ((smem_u32(w_sm) & 0x3FFFFu) >> 4) | ((uint32_t)p.NP << 16);   // LBO = NP*16 B
        const uint32_t kstep_a = (uint32_t)(2 * p.sps) >> 4;
        const uint32_t kstep_b = (uint32_t)(2 * p.NP);
        const uint32_t slot_mask = (uint32_t)p.n_slots - 1u;                   // n_slots is a power of two
        const uint32_t slot_shift = (uint32_t)__ffs(p.n_slots) - 1u;
        const uint32_t wp = (uint32_t)p.Wp;
        const uint32_t tap_stride = (uint32_t)(p.KC * p.NP);                   // 16-byte units between the taps of the blob
        const bool pair = KPAIRS >= 2 && p.pair != 0;
        mbar_wait(w_full, 0);
        int it = 0;
        uint32_t tc0 = 0;      // running tile counter (same sequence in the MMA and the epilogue warps)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int s = it % p.n_stages;
            const uint32_t ph = (uint32_t)((it / p.n_stages) & 1);
            const int mtb = band_tiles(p, (item % p.bands) * p.TH);
            mbar_wait(&in_full[s], ph);
            tc_fence_after();
            // descriptor of padded position -1 of the band (tap dy=-1, dx=-1 of output position 0)
            const uint32_t band_lo = (((smem_u32(stage_sm + (size_t)s * p.stage_bytes + 16) & 0x3FFFFu) >> 4) | a_lbo) - 1u;
            // my tiles of this band: tc0 + mt with (tc0 + mt) & 3 == my
            for (int mt = (int)((my - tc0) & (uint32_t)(N_MMA_WARPS - 1)); mt < mtb; mt += N_MMA_WARPS) {
                const uint32_t tc = tc0 + (uint32_t)mt;
                const uint32_t tile_lo = band_lo + (uint32_t)mt * 128u;
                const uint32_t slot = tc & slot_mask;
                const uint32_t sph = (tc >> slot_shift) & 1u;
                mbar_wait(&acc_empty[slot], sph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + slot * (uint32_t)p.slot_cols;
                uint32_t b_lo = w_lo;
                if (!no_mma && pair) {
                    // K pairing: the last chunk with real channels is only half of a K = 16 step (24 -> 32 channels).
                    // Pair it with ITSELF one position further (A: LBO = 16 B) and the weights of tap (dy, 0) with
                    // those of (dy, 1) (B: LBO = the tap stride of the blob); tap (dy, 2) pairs with the all-zero
                    // padding chunk of the blob.  5 instead of 6 MMAs per dy; the padding plane is never read.
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t row_lo = tile_lo + (uint32_t)dy * wp;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            uint32_t a_lo = row_lo + (uint32_t)dx;
                            uint32_t bb = b_lo + (uint32_t)dx * tap_stride;
#pragma unroll
                            for (int kp = 0; kp < KPAIRS - 1; ++kp) {
                                tc_mma_bf16_elect(d_tmem, a_lo, bb, UMMA_DESC_HI, idesc, (dy | dx | kp) != 0 ? 1u : 0u);
                                a_lo += kstep_a;
                                bb += kstep_b;
                            }
                        }
                        // descriptors of the last real chunk: A with LBO = 1 position, B with LBO = tap stride
                        const uint32_t a2 = ((row_lo + (uint32_t)(KPAIRS - 1) * kstep_a) & 0xFFFFu) | (1u << 16);
                        const uint32_t b2 = (b_lo + (uint32_t)(KPAIRS - 1) * kstep_b) & 0xFFFFu;
                        tc_mma_bf16_elect(d_tmem, a2, b2 | (tap_stride << 16), UMMA_DESC_HI, idesc, 1u);
                        tc_mma_bf16_elect(d_tmem, a2 + 2u, (b2 + 2u * tap_stride) | ((uint32_t)p.NP << 16), UMMA_DESC_HI, idesc, 1u);
                        b_lo += 3u * tap_stride;
                    }
                } else if (!no_mma) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        uint32_t a_lo = tile_lo + (uint32_t)(t / 3) * wp + (uint32_t)(t % 3);
#pragma unroll
                        for (int kp = 0; kp < KPAIRS; ++kp) {
                            tc_mma_bf16_elect(d_tmem, a_lo, b_lo, UMMA_DESC_HI, idesc, (t | kp) != 0 ? 1u : 0u);
                            a_lo += kstep_a;
                            b_lo += kstep_b;
                        }
                    }
                }
                tc_commit_elect(&acc_full[slot]);
            }
            tc_commit_elect(&in_empty[s]);
            tc0 += (uint32_t)mtb;
        }
    } else {
        // ================= epilogue: TMEM -> bias + ELU -> bf16 -> (pool) -> global =================
        const int ew = warp - EPI_WARP0;         // 0..15
        const uint32_t grp = (uint32_t)(ew >> 2); // tile tc is drained by group tc & 3
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int etid = tid - 32 * EPI_WARP0;   // 0..511
        mbar_wait(w_full, 0);                    // bias visible
        int it = 0;
        uint32_t tc0 = 0;
        const int n_groups16 = (p.cout + 3) >> 2;     // 4-channel groups that hold real channels
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int n = item / p.bands, y0 = (item % p.bands) * p.TH;
            const int mtb = band_tiles(p, y0);
            uint8_t *out_n = reinterpret_cast<uint8_t *>(p.out) + (long long)n * p.out_sample;
            for (int mt = (int)((grp - tc0) & (uint32_t)(N_EPI_GROUPS - 1)); mt < mtb; mt += N_EPI_GROUPS) {
                const uint32_t tc = tc0 + (uint32_t)mt;
                const uint32_t slot = tc & ((uint32_t)p.n_slots - 1u);
                const uint32_t sph = (tc >> ((uint32_t)__ffs(p.n_slots) - 1u)) & 1u;
                mbar_wait(&acc_full[slot], sph);
                tc_fence_after();
                const int o = mt * 128 + quarter * 32 + lane;       // padded raster position in the band
                const uint32_t taddr = tmem_base + slot * (uint32_t)p.slot_cols + ((uint32_t)(quarter * 32) << 16);
                if (p.pool) {
                    // ELU and the bf16 rounding are monotone, so max-pool commutes with them: stage the raw
                    // accumulators (fp16: 2^-12 relative, far below the bf16 output rounding) and apply
                    // bias + ELU after the 2x2 max, on a quarter of the values.
                    const bool in_band = o < p.TH * p.Wp;
                    for (int ng = 0; ng < ((p.dbg & 1) ? 0 : p.NP / 16); ++ng) {
                        float v[16];
                        tmem_ld16(taddr + (uint32_t)(ng * 16), v);
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                            pk[j] = *reinterpret_cast<uint32_t *>(&h);
                        }
                        if (in_band) {
                            uint4 *d0 = reinterpret_cast<uint4 *>(staging_sm + ((size_t)(2 * ng) * p.TH * p.Wp + o) * 16);
                            uint4 *d1 = reinterpret_cast<uint4 *>(staging_sm + ((size_t)(2 * ng + 1) * p.TH * p.Wp + o) * 16);
                            *d0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            *d1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                    }
                } else {
                    const int r = (int)__umulhi((unsigned)o, p.wp_magic), c = o - r * p.Wp;
                    const int y = y0 + r;
                    const bool valid = r < p.TH && c >= 1 && c <= p.W && y < p.H;
                    for (int ng = 0; ng < ((p.dbg & 1) ? 0 : p.NP / 16); ++ng) {
                        float v[16];
                        tmem_ld16(taddr + (uint32_t)(ng * 16), v);
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (ng * 4 + j < n_groups16) {     // warp-uniform: padded channels stay exactly zero
                                const float4 b4 = *reinterpret_cast<const float4 *>(&bias_sm[ng * 16 + 4 * j]);
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_f(v[4 * j] + b4.x), elu_f(v[4 * j + 1] + b4.y));
                                __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_f(v[4 * j + 2] + b4.z), elu_f(v[4 * j + 3] + b4.w));
                                pk[2 * j] = *reinterpret_cast<uint32_t *>(&h0);
                                pk[2 * j + 1] = *reinterpret_cast<uint32_t *>(&h1);
                            } else {
                                pk[2 * j] = 0u;
                                pk[2 * j + 1] = 0u;
                            }
                        }
                        if (valid) {
                            const long long pos = ((long long)(y + 1) * p.Wp + c) * 16;
                            if (2 * ng < p.NCHR)
                                *reinterpret_cast<uint4 *>(out_n + (long long)(2 * ng) * p.out_plane + pos) =
                                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            if (2 * ng + 1 < p.NCHR)
                                *reinterpret_cast<uint4 *>(out_n + (long long)(2 * ng + 1) * p.out_plane + pos) =
                                    make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[slot]);
            }
            tc0 += (uint32_t)mtb;
            if (p.pool && !(p.dbg & 1)) {
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // band staged
                const int per_ch = (p.TH / 2) * p.Wo;
                const int yo0 = y0 / 2;
                // one pass over (chunk, pooled position): the small bands of the deep layers have far fewer pooled
                // positions per chunk than epilogue threads, so a loop over the chunks would run them one after the other
                for (int idx = etid; idx < p.NCHR * per_ch; idx += EPI_THREADS) {
                    const int ch = idx / per_ch, e = idx - ch * per_ch;
                    const int nreal = min(8, p.cout - 8 * ch);       // real channels of this chunk (<= 0: padding only)
                    const uint8_t *sch = staging_sm + (size_t)ch * p.TH * p.Wp * 16 + 16;
                    uint8_t *och = out_n + (long long)ch * p.out_plane + 16;
                    const int pr = p.wo_magic ? (int)__umulhi((unsigned)e, p.wo_magic) : e;
                    const int pc = e - pr * p.Wo;
                    if (yo0 + pr < p.Ho) {
                        uint4 o4 = make_uint4(0u, 0u, 0u, 0u);
                        if (nreal > 0) {
                            const uint8_t *b = sch + ((size_t)(2 * pr) * p.Wp + 2 * pc) * 16;
                            const uint4 q00 = *reinterpret_cast<const uint4 *>(b);
                            const uint4 q01 = *reinterpret_cast<const uint4 *>(b + 16);
                            const uint4 q10 = *reinterpret_cast<const uint4 *>(b + (size_t)p.Wp * 16);
                            const uint4 q11 = *reinterpret_cast<const uint4 *>(b + (size_t)p.Wp * 16 + 16);
                            const uint32_t *a0 = &q00.x, *a1 = &q01.x, *a2 = &q10.x, *a3 = &q11.x;
                            uint32_t *oo = &o4.x;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (2 * j < nreal) {          // channel counts are multiples of 4 ... or 2
                                    const __half2 m0 = __hmax2(*reinterpret_cast<const __half2 *>(&a0[j]),
                                                               *reinterpret_cast<const __half2 *>(&a1[j]));
                                    const __half2 m1 = __hmax2(*reinterpret_cast<const __half2 *>(&a2[j]),
                                                               *reinterpret_cast<const __half2 *>(&a3[j]));
                                    const float2 f = __half22float2(__hmax2(m0, m1));
                                    const float2 b2 = *reinterpret_cast<const float2 *>(&bias_sm[ch * 8 + 2 * j]);
                                    const float r0 = elu_f(f.x + b2.x);
                                    const float r1 = (2 * j + 1 < nreal) ? elu_f(f.y + b2.y) : 0.f;
                                    __nv_bfloat162 h = __floats2bfloat162_rn(r0, r1);
                                    oo[j] = *reinterpret_cast<uint32_t *>(&h);
                                }
                            }
                        }
                        *reinterpret_cast<uint4 *>(och + ((long long)(yo0 + pr + 1) * p.Wpo + pc) * 16) = o4;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // staging free again
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// --------------------------------------------------------------------------------------
// Row-stacked variant for the wide layers (W >= ~100, Cout <= 32): an MMA tile is 128 consecutive
// pixels of ONE image row, and RS_R = 4 consecutive output rows are stacked along N.  Input row
// r0 + v (v = 0..5) feeds output rows r0 + v - dy, so one A view is multiplied with the weights of
// up to three taps at once (N = 1..3 x Cout instead of Cout): 18 instead of 36 A-tile reads per
// 4 x 128 outputs -- the shared-memory operand fetch is what bounds the SS-mode MMA at small N.
// The weight blob holds, per (dx, K chunk), the N-major block list [W(dy=2), W(dy=1), W(dy=0), 0];
// view v uses a window of it.  2x2 max-pooling happens in registers (rows are column groups of
// the same TMEM lane, the horizontal neighbour is the adjacent lane): no staging band, no barrier.
// --------------------------------------------------------------------------------------
constexpr int RS_R = 4;

__device__ __forceinline__ int rows_band_tiles(const ConvParams &p, int y0) {
    return ((min(p.TH, p.Hc - y0) + RS_R - 1) >> 2) * p.JT;
}

template <int KPAIRS>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv3x3_rows_kernel(const ConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (see conv3x3_tc_kernel)
    uint8_t *w_sm = smem;
    float *bias_sm = reinterpret_cast<float *>(smem + p.off_bias);
    uint8_t *stage_sm = smem + p.off_stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *w_full = bars;                       // 1
    uint64_t *in_full = bars + 1;                  // [2]
    uint64_t *in_empty = bars + 3;                 // [2]
    uint64_t *acc_full = bars + 5;                 // [MAX_SLOTS]
    uint64_t *acc_empty = bars + 5 + MAX_SLOTS;    // [MAX_SLOTS]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 5 + 2 * MAX_SLOTS);

    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], N_MMA_WARPS); }
        for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], p.coop ? EPI_THREADS / 32 : 4); }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    if (p.KCL < p.KC) {     // planes of all-padding input chunks: zero once, never loaded
        for (int st = 0; st < p.n_stages; ++st) {
            uint4 *z = reinterpret_cast<uint4 *>(smem + p.off_stage + (size_t)st * p.stage_bytes + 16 + (size_t)p.KCL * p.sps);
            const int n16 = (p.stage_bytes - 16 - p.KCL * p.sps) / 16;
            for (int i = tid; i < n16; i += CONV_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // narrow images: SPT samples side by side in one tile row (smem row = [sample 0 | sample 1 | ...], pitch SP each)
    const int n_sgroups = (p.n_samples + p.SPT - 1) / p.SPT;
    const int n_items = n_sgroups * p.bands;

    if (warp == 0 && p.SPT > 1) {
        // ================= TMA producer, gather form: one bulk copy per (input row, sample, chunk) =================
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)(p.wbytes + p.NP * 4));
            tma_bulk_g2s(w_sm, p.wblob, (uint32_t)p.wbytes, w_full);
            tma_bulk_g2s(bias_sm, reinterpret_cast<const uint8_t *>(p.wblob) + p.wbytes, (uint32_t)(p.NP * 4), w_full);
        }
        const uint32_t row_bytes = (uint32_t)(p.Wp * 16);
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int sg = item / p.bands, y0 = (item % p.bands) * p.TH;
            const int s = it % p.n_stages;
            const uint32_t ph = (uint32_t)((it / p.n_stages) & 1);
            const int rows_in = min(p.TH + 2, p.Hp - y0);
            const int ns = min(p.SPT, p.n_samples - sg * p.SPT);
            const int per_chunk = rows_in * ns;
            const int total = per_chunk * p.KCL;
            if (lane == 0) {
                mbar_wait(&in_empty[s], ph ^ 1u);
                mbar_expect_tx(&in_full[s], (uint32_t)total * row_bytes);
            }
            __syncwarp();
            uint8_t *dst0 = stage_sm + (size_t)s * p.stage_bytes + 16;
            for (int i = lane; i < total; i += 32) {
                const int kc = i / per_chunk, rem = i - kc * per_chunk;
                const int r = rem / ns, u = rem - r * ns;
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.in) + (long long)(sg * p.SPT + u) * p.in_sample +
                                     (long long)kc * p.in_plane + (long long)(y0 + r) * row_bytes;
                tma_bulk_g2s(dst0 + (size_t)kc * p.sps + ((size_t)r * p.RP + (size_t)u * p.SP) * 16, src, row_bytes, &in_full[s]);
            }
        }
    } else if (warp == 0) {
        // ================= TMA producer (same bands as conv3x3_tc_kernel) =================
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)(p.wbytes + p.NP * 4));
            tma_bulk_g2s(w_sm, p.wblob, (uint32_t)p.wbytes, w_full);
            tma_bulk_g2s(bias_sm, reinterpret_cast<const uint8_t *>(p.wblob) + p.wbytes, (uint32_t)(p.NP * 4), w_full);
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int n = item / p.bands, y0 = (item % p.bands) * p.TH;
                const int s = it % p.n_stages;
                const uint32_t ph = (uint32_t)((it / p.n_stages) & 1);
                mbar_wait(&in_empty[s], ph ^ 1u);
                const int rows_in = min(p.TH + 2, p.Hp - y0);
                const uint32_t bytes = (uint32_t)(rows_in * p.Wp * 16);
                mbar_expect_tx(&in_full[s], bytes * (uint32_t)p.KCL);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.in) + (long long)n * p.in_sample +
                                     (long long)y0 * p.Wp * 16;
                uint8_t *dst = stage_sm + (size_t)s * p.stage_bytes + 16;
                for (int kc = 0; kc < p.KCL; ++kc)
                    tma_bulk_g2s(dst + (size_t)kc * p.sps, src + (long long)kc * p.in_plane, bytes, &in_full[s]);
            }
        }
    } else if (warp < EPI_WARP0) {
        // ================= MMA issuers =================
        const uint32_t my = (uint32_t)(warp - 1);
        const bool no_mma = (p.dbg & 2) != 0;
        const uint32_t np = (uint32_t)p.NP;
        uint32_t idesc[RS_R + 1];
#pragma unroll
        for (int k = 1; k <= RS_R; ++k) idesc[k] = umma_idesc_bf16((int)np * k);
        const uint32_t a_lbo = ((uint32_t)p.sps >> 4) << 16;
        const uint32_t w_lo = ((smem_u32(w_sm) & 0x3FFFFu) >> 4) | ((RS_R * np) << 16);   // LBO = 4*NP*16 B
        const uint32_t kstep_a = (uint32_t)(2 * p.sps) >> 4;
        const uint32_t slot_mask = (uint32_t)p.n_slots - 1u;
        const uint32_t slot_shift = (uint32_t)__ffs(p.n_slots) - 1u;
        const uint32_t wp = (uint32_t)p.RP;          // smem row pitch in positions
        const uint32_t jt = (uint32_t)p.JT;
        mbar_wait(w_full, 0);
        int it = 0;
        uint32_t tc0 = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int s = it % p.n_stages;
            const uint32_t ph = (uint32_t)((it / p.n_stages) & 1);
            const int mtb = rows_band_tiles(p, (item % p.bands) * p.TH);
            mbar_wait(&in_full[s], ph);
            tc_fence_after();
            // padded position (row 0, col 0) of the band; lane l of tile (rg, j) is padded column 1 + 128 j + l
            const uint32_t band_lo = ((smem_u32(stage_sm + (size_t)s * p.stage_bytes + 16) & 0x3FFFFu) >> 4) | a_lbo;
            for (int mt = (int)((my - tc0) & (uint32_t)(N_MMA_WARPS - 1)); mt < mtb; mt += N_MMA_WARPS) {
                const uint32_t tc = tc0 + (uint32_t)mt;
                const uint32_t rg = (uint32_t)mt / jt, j = (uint32_t)mt - rg * jt;
                const uint32_t tile_lo = band_lo + RS_R * rg * wp + 128u * j;
                const uint32_t slot = tc & slot_mask;
                const uint32_t sph = (tc >> slot_shift) & 1u;
                mbar_wait(&acc_empty[slot], sph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + slot * (uint32_t)p.slot_cols;
                // input row v of the group: block list window [zs, zs + nb) -> accumulator blocks [db, db + nb)
                //   v : 2 0 1 3 4 5   (v = 2 first: with a 4th, all-zero block it initialises every column)
                if (!no_mma) {
#pragma unroll
                for (int i = 0; i < RS_R + 2; ++i) {
                    constexpr int order[6] = {2, 0, 1, 3, 4, 5};
                    const int v = order[i];
                    const int zs = v < 2 ? 2 - v : 0;
                    const int db = v > 2 ? v - 2 : 0;
                    const int nb = v < 2 ? v + 1 : (v > 3 ? 6 - v : 3);
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        uint32_t a_lo = tile_lo + (uint32_t)v * wp + (uint32_t)dx;
                        uint32_t b_lo = w_lo + (uint32_t)(dx * 2 * KPAIRS) * (RS_R * np) + (uint32_t)zs * np;
#pragma unroll
                        for (int kp = 0; kp < KPAIRS; ++kp) {
                            const bool first = (i == 0 && dx == 0 && kp == 0);
                            tc_mma_bf16_elect(d_tmem + (uint32_t)db * np, a_lo, b_lo, UMMA_DESC_HI, idesc[first ? RS_R : nb],
                                              first ? 0u : 1u);
                            a_lo += kstep_a;
                            b_lo += 2u * RS_R * np;
                        }
                    }
                }
                }
                tc_commit_elect(&acc_full[slot]);
            }
            tc_commit_elect(&in_empty[s]);
            tc0 += (uint32_t)mtb;
        }
    } else {
        // ================= epilogue =================
        // coop (few accumulator slots, NP = 32): all four groups drain every tile together -- group g takes
        // output row g (no pool) or pooled row g >> 1 and every other channel chunk (pool) -- so a slot is
        // back with the MMA warps after a quarter of the per-tile work.  Otherwise tile tc belongs to
        // group tc & 3, which drains all of it.
        const int ew = warp - EPI_WARP0;
        const int grp = ew >> 2;
        const int quarter = warp & 3;
        const bool coop = p.coop != 0;
        mbar_wait(w_full, 0);
        int it = 0;
        uint32_t tc0 = 0;
        const int n_groups16 = (p.cout + 3) >> 2;
        const int odd = lane & 1;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int sg = item / p.bands, y0 = (item % p.bands) * p.TH;
            const int mtb = rows_band_tiles(p, y0);
            const int mt0 = coop ? 0 : (int)(((uint32_t)grp - tc0) & (uint32_t)(N_EPI_GROUPS - 1));
            for (int mt = mt0; mt < mtb; mt += coop ? 1 : N_EPI_GROUPS) {
                const uint32_t tc = tc0 + (uint32_t)mt;
                const uint32_t slot = tc & ((uint32_t)p.n_slots - 1u);
                const uint32_t sph = (tc >> ((uint32_t)__ffs(p.n_slots) - 1u)) & 1u;
                mbar_wait(&acc_full[slot], sph);
                tc_fence_after();
                const int rg = mt / p.JT, j = mt - rg * p.JT;
                // the lane's output position in the tile row: sample u of the group, padded column c
                const int pos_row = 1 + 128 * j + quarter * 32 + lane;
                const int u = p.SPT > 1 ? (int)__umulhi((unsigned)pos_row, p.sp_magic) : 0;
                const int c = pos_row - u * p.SP;
                const int n = sg * p.SPT + u;
                const int y = y0 + RS_R * rg;                        // first of the four conv rows
                const bool valid = c >= 1 && c <= p.W && u < p.SPT && n < p.n_samples;
                uint8_t *out_n = reinterpret_cast<uint8_t *>(p.out) + (long long)min(n, p.n_samples - 1) * p.out_sample;
                const uint32_t taddr = tmem_base + slot * (uint32_t)p.slot_cols + ((uint32_t)(quarter * 32) << 16);
                if (p.dbg & 1) {
                } else if (p.pool && coop) {
                    // rows (2 pr, 2 pr + 1) pool vertically inside the thread; lanes (2k, 2k+1) are one pooled
                    // column: the even lane finishes channels 0-3 of the chunk, the odd lane channels 4-7.
                    const int pr = grp >> 1;
                    const int yo = (y >> 1) + pr;
                    const long long opos = ((long long)(yo + 1) * p.Wpo + ((c - 1) >> 1) + 1) * 16 + 8 * odd;
                    for (int h = grp & 1; h < p.NCHR; h += 2) {
                        float v[16];
                        tmem_ld8x2(taddr + (uint32_t)(2 * pr * p.NP + h * 8), (uint32_t)p.NP, v);
                        float m[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float lo = fmaxf(v[k], v[8 + k]), hi = fmaxf(v[4 + k], v[12 + k]);
                            const float other = __shfl_xor_sync(0xffffffffu, odd ? lo : hi, 1);
                            m[k] = fmaxf(odd ? hi : lo, other);
                        }
                        const float4 b4 = *reinterpret_cast<const float4 *>(&bias_sm[h * 8 + 4 * odd]);
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_f(m[0] + b4.x), elu_f(m[1] + b4.y));
                        __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_f(m[2] + b4.z), elu_f(m[3] + b4.w));
                        const bool real = h * 2 + odd < n_groups16;       // padded channels stay exactly zero
                        const uint2 o2 = make_uint2(real ? *reinterpret_cast<uint32_t *>(&h0) : 0u,
                                                    real ? *reinterpret_cast<uint32_t *>(&h1) : 0u);
                        if (valid && yo < p.Ho && ((c - 1) >> 1) < p.Wo)
                            *reinterpret_cast<uint2 *>(out_n + (long long)h * p.out_plane + opos) = o2;
                    }
                } else if (p.pool) {
                    // rows (0,1) and (2,3) pool vertically inside the thread; lanes (2k, 2k+1) are one pooled
                    // column: the even lane finishes pooled row 0, the odd lane pooled row 1.
                    const int yo = (y >> 1) + odd;
                    const long long opos = ((long long)(yo + 1) * p.Wpo + ((c - 1) >> 1) + 1) * 16;
                    for (int h = 0; h < p.NCHR; ++h) {
                        float v[32];
                        tmem_ld8x4(taddr + (uint32_t)(h * 8), (uint32_t)p.NP, v);
                        float m[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float top = fmaxf(v[k], v[8 + k]), bot = fmaxf(v[16 + k], v[24 + k]);
                            const float other = __shfl_xor_sync(0xffffffffu, odd ? top : bot, 1);
                            m[k] = fmaxf(odd ? bot : top, other);
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int q4 = 0; q4 < 2; ++q4) {
                            if (h * 2 + q4 < n_groups16) {     // warp-uniform: padded channels stay exactly zero
                                const float4 b4 = *reinterpret_cast<const float4 *>(&bias_sm[h * 8 + 4 * q4]);
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_f(m[4 * q4] + b4.x), elu_f(m[4 * q4 + 1] + b4.y));
                                __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_f(m[4 * q4 + 2] + b4.z), elu_f(m[4 * q4 + 3] + b4.w));
                                pk[2 * q4] = *reinterpret_cast<uint32_t *>(&h0);
                                pk[2 * q4 + 1] = *reinterpret_cast<uint32_t *>(&h1);
                            } else {
                                pk[2 * q4] = 0u;
                                pk[2 * q4 + 1] = 0u;
                            }
                        }
                        if (valid && yo < p.Ho && ((c - 1) >> 1) < p.Wo)
                            *reinterpret_cast<uint4 *>(out_n + (long long)h * p.out_plane + opos) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                } else {
#pragma unroll 1
                    for (int jr = coop ? grp : 0; jr < (coop ? grp + 1 : RS_R); ++jr) {
                        const long long pos = ((long long)(y + jr + 1) * p.Wp + c) * 16;
                        for (int ng = 0; ng < p.NP / 16; ++ng) {
                            float v[16];
                            tmem_ld16(taddr + (uint32_t)(jr * p.NP + ng * 16), v);
                            uint32_t pk[8];
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) {
                                if (ng * 4 + q4 < n_groups16) {
                                    const float4 b4 = *reinterpret_cast<const float4 *>(&bias_sm[ng * 16 + 4 * q4]);
                                    __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_f(v[4 * q4] + b4.x), elu_f(v[4 * q4 + 1] + b4.y));
                                    __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_f(v[4 * q4 + 2] + b4.z), elu_f(v[4 * q4 + 3] + b4.w));
                                    pk[2 * q4] = *reinterpret_cast<uint32_t *>(&h0);
                                    pk[2 * q4 + 1] = *reinterpret_cast<uint32_t *>(&h1);
                                } else {
                                    pk[2 * q4] = 0u;
                                    pk[2 * q4 + 1] = 0u;
                                }
                            }
                            if (valid) {
                                if (2 * ng < p.NCHR)
                                    *reinterpret_cast<uint4 *>(out_n + (long long)(2 * ng) * p.out_plane + pos) =
                                        make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                if (2 * ng + 1 < p.NCHR)
                                    *reinterpret_cast<uint4 *>(out_n + (long long)(2 * ng + 1) * p.out_plane + pos) =
                                        make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[slot]);
            }
            tc0 += (uint32_t)mtb;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// --------------------------------------------------------------------------------------
// layer 0: prepare + conv3x3 (Cin = 1) + BN + ELU on CUDA cores -> P8 bf16
// --------------------------------------------------------------------------------------
struct L0Params {
    const void *x;          // (n,1,Hin,Win) f32 or u8
    int x_u8, prepare;      // ASR_PREP_*
    int Hin, Win, H, W, Wp, Hp;
    int C, NCH;             // real channels, output chunks
    const float *w;         // [C][9] folded (scale included), then [C] bias
    bf16 *out;
    long long out_plane, out_sample;
    int n;
};

__device__ __forceinline__ float l0_fetch(const L0Params &p, const uint8_t *xu, const float *xf, int y, int x) {
    if (y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.f;
    if (p.prepare == ASR_PREP_SCALE_HALF) {
        int yy = 2 * y, xx = 2 * x;
        float a, b, c, d;
        if (p.x_u8) {
            a = xu[yy * p.Win + xx]; b = xu[yy * p.Win + xx + 1];
            c = xu[(yy + 1) * p.Win + xx]; d = xu[(yy + 1) * p.Win + xx + 1];
        } else {
            a = xf[yy * p.Win + xx]; b = xf[yy * p.Win + xx + 1];
            c = xf[(yy + 1) * p.Win + xx]; d = xf[(yy + 1) * p.Win + xx + 1];
        }
        const float s = 1.0f / 255.0f;
        return ((a * s + b * s) + (c * s + d * s)) * 0.25f;
    }
    float v = p.x_u8 ? (float)xu[y * p.Win + x] : xf[y * p.Win + x];
    return p.prepare == ASR_PREP_SCALE ? v / 255.0f : v;
}

constexpr int L0_MAXC = 32;
constexpr int L0_TW = 32, L0_TROWS = 8, L0_PIX = 4;   // 32 x 8 threads, each computes 4 vertically adjacent pixels
constexpr int L0_TH = L0_TROWS * L0_PIX;                // 32 x 32 pixels per block

// Folded weights travel in the kernel parameter block (constant bank): with the channel loops fully
// unrolled every FFMA takes its weight straight from c[0][..], no load instruction.
struct L0Weights {
    float w[L0_MAXC * 9];
    float b[L0_MAXC];
};

template <int C>   // C > 0: compile-time channel count; C == 0: run-time p.C
__global__ void __launch_bounds__(L0_TW * L0_TROWS) l0_conv_kernel(const L0Params p, const L0Weights wt) {
    __shared__ float tile[L0_TH + 2][L0_TW + 2];
    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * L0_TW, y0 = blockIdx.y * L0_TH;
    const size_t in_off = (size_t)n * p.Hin * p.Win;
    const uint8_t *xu = reinterpret_cast<const uint8_t *>(p.x) + in_off;
    const float *xf = reinterpret_cast<const float *>(p.x) + in_off;
    for (int i = tid; i < (L0_TH + 2) * (L0_TW + 2); i += L0_TW * L0_TROWS) {
        const int ty = i / (L0_TW + 2), tx = i - ty * (L0_TW + 2);
        tile[ty][tx] = l0_fetch(p, xu, xf, y0 + ty - 1, x0 + tx - 1);
    }
    __syncthreads();
    const int tr = tid / L0_TW, tx = tid - tr * L0_TW;
    const int x = x0 + tx;
    if (x >= p.W) return;
    constexpr int NCH_T = C > 0 ? (C + 15) / 16 * 2 : 0;
    const int nch = C > 0 ? NCH_T : p.NCH;
    const int cc = C > 0 ? C : p.C;
    // sliding 3-row window over the thread's 4 pixels
    float v[9];
#pragma unroll
    for (int t = 0; t < 6; ++t) v[3 + t] = tile[tr * L0_PIX + t / 3][tx + t % 3];
#pragma unroll
    for (int py = 0; py < L0_PIX; ++py) {
        const int ty = tr * L0_PIX + py, y = y0 + ty;
#pragma unroll
        for (int t = 0; t < 6; ++t) v[t] = v[t + 3];
#pragma unroll
        for (int t = 0; t < 3; ++t) v[6 + t] = tile[ty + 2][tx + t];
        if (y < p.H) {
            uint8_t *out_px = reinterpret_cast<uint8_t *>(p.out) + (long long)n * p.out_sample +
                              ((long long)(y + 1) * p.Wp + x + 1) * 16;
#pragma unroll
            for (int ch = 0; ch < (C > 0 ? NCH_T : 4); ++ch) {
                if (ch < nch) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float r[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = ch * 8 + 2 * j + h;
                            float acc = 0.f;
                            if (c < cc) {
                                acc = wt.b[c];
#pragma unroll
                                for (int t = 0; t < 9; ++t) acc = fmaf(v[t], wt.w[c * 9 + t], acc);
                                acc = elu_f(acc);
                            }
                            r[h] = acc;
                        }
                        __nv_bfloat162 hh = __floats2bfloat162_rn(r[0], r[1]);
                        pk[j] = *reinterpret_cast<uint32_t *>(&hh);
                    }
                    *reinterpret_cast<uint4 *>(out_px + (long long)ch * p.out_plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------
// layer 0 on the tensor cores.  With Cin = 1 the 3x3 convolution is a banded (Toeplitz) product:
//   D[g, (r, co)] = sum_dx sum_k A[g + dx, k] * B_dx[k, (r, co)],   B_dx[k, (r, co)] = w[co][dy = k - r][dx]
// g runs over the padded columns of a batch of samples (raster over (sample, column): A row = one
// pixel column, garbage at the two pad columns as in the raster conv kernel), k over 16 input rows,
// r over the R0 output rows they support.  A holds, per pixel column, 8 vertically adjacent pixels
// in 16 bytes -- the canonical K-major core-matrix row -- so the dx shift is again a descriptor
// shift and no im2col is materialised.  Pixels and weights are split into bf16 hi + lo parts and
// three products (hi*hi, lo*hi, hi*lo) are accumulated: ~2^-16 relative, i.e. fp32-grade for the
// bf16 activations that follow, at an MMA cost that stays negligible next to the ELU epilogue.
// One persistent CTA per SM: 8 converter warps (global -> hi/lo A tiles, one tile each in flight), one MMA
// warp (9 MMAs per tile into one of two TMEM slots), 16 drain warps (warp = TMEM lane quarter x
// row subset: bias + ELU + bf16 + P8 store).
// --------------------------------------------------------------------------------------
constexpr int L0T_PROD_WARPS = 8;                                   // converter warps; warp w converts every 8th tile
constexpr int L0T_EPI_WARP0 = L0T_PROD_WARPS + 1;                    // warp 8 issues the MMAs
constexpr int L0T_EPI_WARPS = 16;
constexpr int L0T_THREADS = 32 * (L0T_EPI_WARP0 + L0T_EPI_WARPS);    // 800
constexpr int L0T_AROWS = 136;      // 130 pixel columns of a tile + slack
constexpr int L0T_NBUF = L0T_PROD_WARPS;   // one A tile (8.5 KB) per converter warp
constexpr int L0T_SLOT_COLS = 256;

struct L0TcParams {
    L0Params b;                 // input description (x, dtype, prepare, sizes), output pointers
    const uint8_t *blob;        // [dx][part: hi, lo][K chunk][NPAD][8] bf16, then C fp32 biases
    int int_pixels;             // u8 input with prepare = x/255: A = integer pixels (hi part only), blob carries w/255
    int R0, NPAD;               // output rows per tile, padded N = R0 * C rounded up to 16
    int TY;                     // row tiles per sample
    int GT;                     // 128-column tiles over n * Wp raster columns
    unsigned wp_magic;
    int dbg;                    // diagnostics (env ASR_L0_DEBUG): 1 = no drain work, 2 = no input loads, 4 = no MMAs
};

// Converter, phase 1: issue the loads of 8 vertically adjacent pixels (image rows yb .. yb+7) of raster
// column g.  Addresses are clamped so that every load is legal and none depends on a branch; `active`
// predicates the loads of a thread that has no item.  raw[k][e]: e = 0 for a plain pixel, 0..3 for
// the 2x2 box of ASR_PREP_SCALE_HALF.  Returns the validity mask of the column.
template <int NE>
__device__ __forceinline__ bool l0t_load8(const L0TcParams &q, long long g, int yb, bool active, float (*raw)[NE]) {
    const L0Params &p = q.b;
    const long long gc = g < 0 ? 0 : g;
    const unsigned s = __umulhi((unsigned)gc, q.wp_magic);
    const int xp = (int)((unsigned)gc - s * (unsigned)p.Wp);
    const int x = xp - 1;
    const bool col_ok = active && g >= 0 && (int)s < p.n && x >= 0 && x < p.W;
    const size_t in_off = (size_t)min((int)s, p.n - 1) * p.Hin * p.Win;
    const uint8_t *xu = reinterpret_cast<const uint8_t *>(p.x) + in_off;
    const float *xf = reinterpret_cast<const float *>(p.x) + in_off;
    const int xc = min(max(x, 0), p.W - 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int yc = min(max(yb + k, 0), p.H - 1);
#pragma unroll
        for (int e = 0; e < NE; ++e) raw[k][e] = 0.f;
        if (active) {
            if (NE == 4) {
                const int o00 = 2 * yc * p.Win + 2 * xc;
                if (p.x_u8) {
                    raw[k][0] = xu[o00]; raw[k][1] = xu[o00 + 1];
                    raw[k][NE - 2] = xu[o00 + p.Win]; raw[k][NE - 1] = xu[o00 + p.Win + 1];
                } else {
                    raw[k][0] = xf[o00]; raw[k][1] = xf[o00 + 1];
                    raw[k][NE - 2] = xf[o00 + p.Win]; raw[k][NE - 1] = xf[o00 + p.Win + 1];
                }
            } else {
                raw[k][0] = p.x_u8 ? (float)xu[yc * p.Win + xc] : xf[yc * p.Win + xc];
            }
        }
    }
    return col_ok;
}
// Converter, phase 2: `prepare` with the operation order of l0_fetch, then the border mask.  u8 / 255
// comes from a 256-entry table of the correctly rounded quotients (same bits, no division sequence).
template <int NE>
__device__ __forceinline__ void l0t_finish8(const L0Params &p, const float *lut, bool col_ok, int yb, const float (*raw)[NE],
                                            float *v) {
    if (NE == 4) {
        const float sc = 1.0f / 255.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v[k] = ((raw[k][0] * sc + raw[k][1] * sc) + (raw[k][NE - 2] * sc + raw[k][NE - 1] * sc)) * 0.25f;
    } else if (p.prepare == ASR_PREP_SCALE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = p.x_u8 ? lut[(int)raw[k][0]] : raw[k][0] / 255.0f;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = raw[k][0];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (col_ok && yb + k >= 0 && yb + k < p.H) ? v[k] : 0.f;
}

// bf16 hi + lo split of 8 values -> the 16-byte K-major rows of the two A parts
__device__ __forceinline__ void l0t_store8(uint8_t *a_buf, int c, int i, const float *vin) {
    // K element 15 (image row y0 + 14) is outside every tap window (R0 <= 12): it carries the constant 1
    // that multiplies the bias row of the weight matrix, so the MMA adds the bias.
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = vin[k];
    if (c == 1) v[7] = 1.0f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
        hi[k] = *reinterpret_cast<const uint32_t *>(&h);
        lo[k] = *reinterpret_cast<const uint32_t *>(&l);
    }
    *reinterpret_cast<uint4 *>(a_buf + ((0 * 2 + c) * L0T_AROWS + i) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(a_buf + ((1 * 2 + c) * L0T_AROWS + i) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// One converter warp builds whole A tiles (every L0T_PROD_WARPS-th tile of the CTA) in its own buffer:
// 260 items (K chunk c, tile column i) = 9 per lane, in batches of three whose 24 loads are in flight
// together (one item at a time for the 4-loads-per-pixel box filter); lanes walk along an image row.  Eight warps keep eight tiles in flight.
template <int NE>
__device__ __forceinline__ void l0t_convert(const L0TcParams &q, uint8_t *a_buf, const float *lut, uint64_t *a_ready,
                                            uint64_t *a_free, int n_tiles, int cw) {
    const L0Params &p = q.b;
    const int lane = threadIdx.x & 31;
    const bool off = (q.dbg & 2) != 0;
    int k = cw;
    for (int t = blockIdx.x + cw * (int)gridDim.x; t < n_tiles; t += L0T_PROD_WARPS * (int)gridDim.x, k += L0T_PROD_WARPS) {
        const int gt = t / q.TY, ty = t - gt * q.TY;
        const long long g0 = (long long)gt * 128 - 1;
        const int y0 = ty * q.R0 - 1;
        mbar_wait(a_free, (uint32_t)(((k / L0T_PROD_WARPS) & 1) ^ 1));
        constexpr int U = NE == 1 ? 3 : 1;          // items whose loads are in flight together
#pragma unroll 1
        for (int batch = 0; batch < 9 / U; ++batch) {
            float raw[U][8][NE];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int item = lane + 32 * (U * batch + u);
                const int c = item >= 130 ? 1 : 0, i = item - 130 * c;
                ok[u] = l0t_load8<NE>(q, g0 + i, y0 + 8 * c, item < 260 && !off, raw[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int item = lane + 32 * (U * batch + u);
                const int c = item >= 130 ? 1 : 0, i = item - 130 * c;
                float v[8];
                l0t_finish8<NE>(p, lut, ok[u], y0 + 8 * c, raw[u], v);
                if (item < 260) l0t_store8(a_buf, c, i, v);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
    }
}

// The same converter specialised for u8 pixels with prepare = x / 255 (MODE 1): the integer pixel is exact
// in bf16, so A has no lo part and 1/255 is folded into the weight matrix (blob_u8); the row clamps and
// masks are dropped for tiles that do not touch the top or bottom image border.  (MODE 2, fp32 pixels as
// they are, measured slower than the generic loop.)
template <int MODE>
__device__ __forceinline__ void l0t_convert_fast(const L0TcParams &q, uint8_t *a_buf, const float *lut, uint64_t *a_ready,
                                                 uint64_t *a_free, int n_tiles, int cw) {
    const L0Params &p = q.b;
    const int lane = threadIdx.x & 31;
    const bool off = (q.dbg & 2) != 0;
    const uint8_t *xu = reinterpret_cast<const uint8_t *>(p.x);
    const float *xf = reinterpret_cast<const float *>(p.x);
    const unsigned img = (unsigned)(p.Hin * p.Win);
    int k = cw;
    for (int t = blockIdx.x + cw * (int)gridDim.x; t < n_tiles; t += L0T_PROD_WARPS * (int)gridDim.x, k += L0T_PROD_WARPS) {
        const int gt = t / q.TY, ty = t - gt * q.TY;
        const int g0 = gt * 128 - 1;
        const int y0 = ty * q.R0 - 1;
        const bool interior = y0 >= 0 && y0 + 15 < p.H;     // warp-uniform
        mbar_wait(a_free, (uint32_t)(((k / L0T_PROD_WARPS) & 1) ^ 1));
#pragma unroll 1
        for (int batch = 0; batch < 3; ++batch) {
            uint32_t raw[3][8];
            bool ok[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int item = lane + 32 * (3 * batch + u);
                const int c = item >= 130 ? 1 : 0, i = item - 130 * c;
                const int g = g0 + i;
                const unsigned gc = (unsigned)max(g, 0);
                const unsigned s = __umulhi(gc, q.wp_magic);
                const int x = (int)(gc - s * (unsigned)p.Wp) - 1;
                ok[u] = item < 260 && !off && g >= 0 && (int)s < p.n && x >= 0 && x < p.W;
                const unsigned base = min(s, (unsigned)p.n - 1u) * img + (unsigned)min(max(x, 0), p.W - 1);
                const int yb = y0 + 8 * c;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int yc = interior ? yb + kk : min(max(yb + kk, 0), p.H - 1);
                    const unsigned o = base + (unsigned)(yc * p.Win);
                    raw[u][kk] = 0u;
                    if (ok[u]) raw[u][kk] = MODE == 1 ? (uint32_t)xu[o] : __float_as_uint(xf[o]);
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int item = lane + 32 * (3 * batch + u);
                const int c = item >= 130 ? 1 : 0, i = item - 130 * c;
                const int yb = y0 + 8 * c;
                float v[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    // MODE 1: the integer pixel itself (exact in bf16; 1/255 lives in the weight matrix)
                    float f = MODE == 1 ? (float)raw[u][kk] : __uint_as_float(raw[u][kk]);
                    const bool row_ok = interior || (yb + kk >= 0 && yb + kk < p.H);
                    v[kk] = (ok[u] && row_ok) ? f : 0.f;
                }
                if (MODE == 1) {
                    if (c == 1) v[7] = 1.0f;             // the constant-one K row (bias)
                    uint32_t hi[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * kk], v[2 * kk + 1]);
                        hi[kk] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    if (item < 260)
                        *reinterpret_cast<uint4 *>(a_buf + (c * L0T_AROWS + i) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                } else if (item < 260) {
                    l0t_store8(a_buf, c, i, v);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
    }
}

template <int C>
__global__ void __launch_bounds__(L0T_THREADS, 1) l0_tc_kernel(const L0TcParams q) {
    extern __shared__ __align__(128) uint8_t smem[];
    const L0Params &p = q.b;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (see conv3x3_tc_kernel)
    const int bbytes = 3 * 2 * 2 * q.NPAD * 16;
    constexpr int ABUF = 2 * 2 * L0T_AROWS * 16;                    // [part][K chunk][L0T_AROWS][8] bf16
    uint8_t *b_sm = smem;
    uint8_t *a_sm = smem + bbytes;
    float *bias_sm = reinterpret_cast<float *>(a_sm + L0T_NBUF * ABUF);
    float *lut_sm = bias_sm + 32;                   // v / 255.0f for v = 0..255
    uint64_t *bars = reinterpret_cast<uint64_t *>(lut_sm + 256);
    uint64_t *a_ready = bars;                       // [NBUF] converters -> MMA warp
    uint64_t *a_free = bars + L0T_NBUF;             // [NBUF] MMAs done reading the buffer
    uint64_t *acc_full = bars + 2 * L0T_NBUF;       // [2]
    uint64_t *acc_empty = bars + 2 * L0T_NBUF + 2;  // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * L0T_NBUF + 4);

    for (int i = tid; i < bbytes / 16; i += L0T_THREADS)
        reinterpret_cast<uint4 *>(b_sm)[i] = reinterpret_cast<const uint4 *>(q.blob)[i];
    if (tid < 32) bias_sm[tid] = tid < C ? reinterpret_cast<const float *>(q.blob + bbytes)[tid] : 0.f;
    if (tid >= 64 && tid < 64 + 256) lut_sm[tid - 64] = (float)(tid - 64) / 255.0f;
    if (tid == 0) {
        for (int s = 0; s < L0T_NBUF; ++s) { mbar_init(&a_ready[s], 1); mbar_init(&a_free[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], L0T_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == L0T_PROD_WARPS) { tmem_alloc(tmem_ptr, 2 * L0T_SLOT_COLS); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // weight matrices visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_tiles = q.GT * q.TY;

    if (warp < L0T_PROD_WARPS) {
        // ================= converters: A row i <-> raster column g0 - 1 + i; K element k <-> image row y0 - 1 + k
        if (q.int_pixels)
            l0t_convert_fast<1>(q, a_sm + warp * ABUF, lut_sm, &a_ready[warp], &a_free[warp], n_tiles, warp);
        else if (p.prepare == ASR_PREP_SCALE_HALF)
            l0t_convert<4>(q, a_sm + warp * ABUF, lut_sm, &a_ready[warp], &a_free[warp], n_tiles, warp);
        else
            l0t_convert<1>(q, a_sm + warp * ABUF, lut_sm, &a_ready[warp], &a_free[warp], n_tiles, warp);
    } else if (warp == L0T_PROD_WARPS) {
        // ================= MMA issuer: per dx (A_hi, B_hi), (A_lo, B_hi), (A_hi, B_lo); whole warp, one elected lane issues
        {
            const uint32_t idesc = umma_idesc_bf16(q.NPAD);
            const uint32_t b_part = (uint32_t)(2 * q.NPAD);                                              // 16-byte units
            const uint32_t b_lo0 = ((smem_u32(b_sm) & 0x3FFFFu) >> 4) | ((uint32_t)q.NPAD << 16);       // LBO = NPAD * 16 B
            const uint32_t a_base = ((smem_u32(a_sm) & 0x3FFFFu) >> 4) | ((uint32_t)L0T_AROWS << 16);    // LBO = 136 * 16 B
            const bool two_parts = !q.int_pixels, no_mma = (q.dbg & 4) != 0;
            int k = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
                const int buf = k % L0T_NBUF, slot = k & 1;
                mbar_wait(&a_ready[buf], (uint32_t)((k / L0T_NBUF) & 1));
                mbar_wait(&acc_empty[slot], (uint32_t)(((k >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t a0 = a_base + (uint32_t)buf * (ABUF / 16);
                const uint32_t d = tmem_base + (uint32_t)slot * L0T_SLOT_COLS;
                if (!no_mma) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const uint32_t ah = a0 + (uint32_t)dx, al = a0 + (ABUF / 32) + (uint32_t)dx;
                        const uint32_t bh = b_lo0 + (uint32_t)(dx * 2) * b_part, bl = b_lo0 + (uint32_t)(dx * 2 + 1) * b_part;
                        tc_mma_bf16_elect(d, ah, bh, UMMA_DESC_HI, idesc, dx > 0 ? 1u : 0u);
                        if (two_parts) tc_mma_bf16_elect(d, al, bh, UMMA_DESC_HI, idesc, 1u);
                        tc_mma_bf16_elect(d, ah, bl, UMMA_DESC_HI, idesc, 1u);
                    }
                }
                tc_commit_elect(&acc_full[slot]);
                tc_commit_elect(&a_free[buf]);
            }
        }
    } else {
        // ================= drain: lane = raster column; group g takes rows g, g + 4, ...
        const int ew = warp - L0T_EPI_WARP0;
        const int quarter = warp & 3, grp = ew >> 2;
        int k = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
            const int slot = k & 1;
            const int gt = t / q.TY, ty = t - gt * q.TY;
            const int y0 = ty * q.R0;
            const unsigned g = (unsigned)gt * 128u + (unsigned)(quarter * 32 + lane);
            const unsigned s = __umulhi(g, q.wp_magic);
            const int xp = (int)(g - s * (unsigned)p.Wp);
            const bool valid = (int)s < p.n && xp >= 1 && xp <= p.W;
            uint8_t *out_px = reinterpret_cast<uint8_t *>(p.out) + (long long)s * p.out_sample + ((long long)(y0 + 1) * p.Wp + xp) * 16;
            const uint32_t taddr = tmem_base + (uint32_t)slot * L0T_SLOT_COLS + ((uint32_t)(quarter * 32) << 16);
            const int nrows = (q.dbg & 1) ? 0 : min(q.R0, p.H - y0);
            mbar_wait(&acc_full[slot], (uint32_t)((k >> 1) & 1));
            tc_fence_after();
            for (int r = grp; r < nrows; r += L0T_EPI_WARPS / 4) {
                uint8_t *o = out_px + (long long)r * p.Wp * 16;
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)(r * C + c0), v);
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int ch = c0 + 2 * j;
                        const float r0 = ch < C ? elu_f(v[2 * j]) : 0.f;          // bias already added by the MMA
                        const float r1 = ch + 1 < C ? elu_f(v[2 * j + 1]) : 0.f;
                        const __nv_bfloat162 h = __floats2bfloat162_rn(r0, r1);
                        pk[j] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    if (valid) {
                        *reinterpret_cast<uint4 *>(o + (long long)(c0 / 8) * p.out_plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (c0 + 8 < C)
                            *reinterpret_cast<uint4 *>(o + (long long)(c0 / 8 + 1) * p.out_plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == L0T_PROD_WARPS) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * L0T_SLOT_COLS); }
}

#include "encoder_fused.cuh"
#include "encoder_fused23.cuh"

// --------------------------------------------------------------------------------------
// head: spatial mean -> 1x1 conv + BN (folded) -> latent -> (latent - mean) . P -> unit norm
// --------------------------------------------------------------------------------------
struct HeadParams {
    const void *in;            // P8 bf16 (tc path) or NCHW fp32 (fp32 path)
    int is_p8;
    int C, H, W, Wp;           // channels, interior size
    long long plane, sample;   // bytes (P8) ; for NCHW: plane = H*W*4, sample = C*plane
    const float *A;            // [32][C] folded 1x1 weights, then [32] bias
    const float *cca_mean;     // [32]
    const float *cca_proj;     // [32][32] (in, out)
    float *codes, *latents;    // may be NULL
};

__global__ void __launch_bounds__(128) head_kernel(const HeadParams p) {
    __shared__ float m[128];
    __shared__ float lat[32];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv = 1.0f / (float)(p.H * p.W);
    if (p.is_p8) {
        // warp w sums channel chunks w, w+4, ...: lanes stride over positions with 16-byte loads
        const int n_chunks = (p.C + 7) >> 3;
        const int npos = p.H * p.W;
        for (int ch = warp; ch < n_chunks; ch += 4) {
            const uint8_t *base = reinterpret_cast<const uint8_t *>(p.in) + (long long)n * p.sample + (long long)ch * p.plane;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int i = lane; i < npos; i += 32) {
                const int y = i / p.W, x = i - y * p.W;
                const uint4 v = *reinterpret_cast<const uint4 *>(base + ((long long)(y + 1) * p.Wp + x + 1) * 16);
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w4[j]));
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s = acc[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && ch * 8 + j < p.C) m[ch * 8 + j] = s * inv;
            }
        }
    } else {
        for (int c = tid; c < p.C; c += blockDim.x) {
            const float *base = reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(p.in) +
                                                                (long long)n * p.sample + (long long)c * p.plane);
            float s = 0.f;
            for (int i = 0; i < p.H * p.W; ++i) s += base[i];
            m[c] = s * inv;
        }
    }
    __syncthreads();
    // latent_j = sum_c A[j][c] m[c] + b[j]: four lanes per output, then a 4-lane reduce
    {
        const int j = tid >> 2, part = tid & 3;
        float s = 0.f;
        for (int c = part; c < p.C; c += 4) s = fmaf(p.A[j * p.C + c], m[c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
            s += p.A[32 * p.C + j];
            lat[j] = s;
            if (p.latents) p.latents[(size_t)n * 32 + j] = s;
        }
    }
    __syncthreads();
    if (tid < 32 && p.codes) {
        float s = 0.f;
        for (int j = 0; j < 32; ++j) s = fmaf(lat[j] - p.cca_mean[j], p.cca_proj[j * 32 + tid], s);
        float ss = s * s;
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        p.codes[(size_t)n * 32 + tid] = s / sqrtf(ss);
    }
}

// --------------------------------------------------------------------------------------
// fp32 CUDA-core path (on-device debug reference; NCHW fp32, unfolded BN exactly as the oracle)
// --------------------------------------------------------------------------------------
__global__ void ref_prepare_kernel(const void *x, int x_u8, int prepare, int Hin, int Win, int H, int W, int n, float *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * H * W) return;
    const int s = (int)(i / ((size_t)H * W));
    const int pix = (int)(i - (size_t)s * H * W);
    L0Params p;
    p.x_u8 = x_u8; p.prepare = prepare; p.Hin = Hin; p.Win = Win; p.H = H; p.W = W;
    const size_t in_off = (size_t)s * Hin * Win;
    out[i] = l0_fetch(p, reinterpret_cast<const uint8_t *>(x) + in_off, reinterpret_cast<const float *>(x) + in_off,
                      pix / W, pix % W);
}

// thread per output element; w (cout,cin,3,3) already flipped if requested
__global__ void ref_conv3x3_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bn,
                                   int n, int cin, int cout, int H, int W, int apply_elu, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * cout * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), co = (int)((i / ((size_t)W * H)) % cout);
    const int s = (int)(i / ((size_t)W * H * cout));
    float acc = 0.f;
    for (int ci = 0; ci < cin; ++ci) {
        const float *ip = in + ((size_t)s * cin + ci) * H * W;
        const float *wp = w + ((size_t)co * cin + ci) * 9;
        for (int t = 0; t < 9; ++t) {
            int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc = fmaf(ip[yy * W + xx], wp[t], acc);
        }
    }
    // bn: [beta | gamma | mean | inv_std] each cout
    float r = (acc - bn[2 * cout + co]) * (bn[cout + co] * bn[3 * cout + co]) + bn[co];
    if (apply_elu) r = r > 0.f ? r : expm1f(r);
    out[i] = r;
}

__global__ void ref_pool_kernel(const float *__restrict__ in, int nc, int H, int W, float *__restrict__ out) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nc * Ho * Wo) return;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
    const size_t c = i / ((size_t)Wo * Ho);
    const float *ip = in + c * H * W + (size_t)(2 * y) * W + 2 * x;
    out[i] = fmaxf(fmaxf(ip[0], ip[1]), fmaxf(ip[W], ip[W + 1]));
}

// --------------------------------------------------------------------------------------
// window extraction: out[i] = src[r0 : r0+win_h, starts[i] : starts[i]+win_w]
// (audio_sheet_retrieval/audio_sheet_server.py:216-223, 260-271, 421-428, 465-477)
// --------------------------------------------------------------------------------------
template <typename T>
__global__ void extract_windows_kernel(const T *__restrict__ src, int src_w, const int *__restrict__ starts, int r0,
                                       int win_h, int win_w, T *__restrict__ out) {
    const int i = blockIdx.y;
    const int start = starts[i];
    T *o = out + (size_t)i * win_h * win_w;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < win_h * win_w; e += gridDim.x * blockDim.x) {
        const int y = e / win_w, x = e - y * win_w;
        o[e] = src[(size_t)(r0 + y) * src_w + start + x];
    }
}

}  // namespace asr

using namespace asr;

// --------------------------------------------------------------------------------------
// handle
// --------------------------------------------------------------------------------------
struct asr_encoder {
    asr_encoder_desc d;
    int max_batch;
    int device = -1;                // the handle is bound to the device that was current at create
    int H0, W0;                     // prepared input size
    LayerGeom g[8];
    int head_c, head_h, head_w;
    double flops;
    // tcgen05 path
    float *l0_w = nullptr;          // [C0*9 + C0] (device copy, unused by the product path)
    L0Weights l0_host;              // folded layer-0 weights, passed by value at launch
    uint8_t *l0_blob = nullptr;     // tensor-core layer 0: banded hi/lo weight matrices + biases
    uint8_t *l0_blob_u8 = nullptr;  // same with w/255 in the pixel rows (u8 input, prepare = x/255: A = integer pixels)
    int l0_R0 = 0, l0_NPAD = 0;
    unsigned l0_wp_magic = 0;
    // layers 0 + 1 fused (l01_fused_kernel): Toeplitz blobs with 8 output rows per tile, launch constants
    uint8_t *f01_blob = nullptr, *f01_blob_u8 = nullptr;
    bool f01_ok = false;
    int fuse_mask = 0;              // bit 0: layers 0 + 1 run fused, bit 1: layers 2 + 3 (asr_encoder_set_fusion)
    F01Params f01;                  // geometry / shared-memory layout, filled at create
    int f01_smem = 0, f01_smem_int = 0;
    // layers 2 + 3 fused (l23_fused_kernel): layer 2's weights in the row-stacked blob format, launch constants
    uint8_t *f23_blobA = nullptr;
    bool f23_ok = false;
    F23Params f23;
    int f23_smem = 0;
    bf16 *wblob[8] = {nullptr};     // layers 1..7
    ConvPlan plan[8];
    bf16 *act[8] = {nullptr};       // P8 activations (output of layer l)
    long long act_plane[8], act_sample[8];
    // head (shared by both paths)
    float *head_A = nullptr;        // [32*C + 32]
    float *cca_mean = nullptr, *cca_proj = nullptr;
    // fp32 path
    float *ref_w[8] = {nullptr}, *ref_bn[8] = {nullptr};
    float *ref_in = nullptr, *ref_act[8] = {nullptr}, *ref_tmp = nullptr;
    int last_path = -1;
    int64_t l01_chunk = 0;          // samples per layer-0/layer-1 sub-chunk (0 = whole batch)
    // optional per-group device timing (bench.py roofline): events around [layer 0], [layers 1..7], [head]
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;     // pairs (start, stop)
    std::vector<int> ev_cat;              // category of pair i: 0 layer 0, 1 tcgen05 conv, 2 head
    size_t ev_used = 0;
    // host-buffer entry
    cudaStream_t s_copy = nullptr, s_comp = nullptr;   // host-buffer entry: shared per device (see host_streams) unless owned
    bool streams_owned = false;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_final = nullptr;     // blocking-sync event: the host thread sleeps until the call's results are back
    void *dev_in[2] = {nullptr, nullptr};
    void *pin_in[2] = {nullptr, nullptr};
    void *pin_out[4] = {nullptr, nullptr, nullptr, nullptr};   // pinned result slots: codes [0,1], latents [2,3]
    float *dev_codes = nullptr, *dev_lat = nullptr;
    int64_t out_cap = 0;
};

static bool plan_conv(const LayerGeom &g, ConvPlan &pl) {
    pl = ConvPlan();
    const int Wp = g.W + 2, KC = g.cinp / 8, NP = g.coutp, NCH = NP / 8;
    pl.wbytes = 9 * KC * NP * 16;
    pl.slot_cols = NP <= 16 ? 16 : (NP <= 32 ? 32 : (NP <= 64 ? 64 : 128));
    pl.n_slots = std::min(MAX_SLOTS, 512 / pl.slot_cols);
    int cols = pl.n_slots * pl.slot_cols;
    pl.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    // rows of conv output that are needed: all of them, or the 2*Ho rows the pool reads
    const int Hc = g.pool ? 2 * g.Ho : g.H;
    const int step = g.pool ? 2 : 1;
    // Pick the band height: fewest 128-position tiles per sample (the MMA work), then fewest halo
    // re-reads; two input stages whenever some band height fits with two.
    for (int ns = 2; ns >= 1; --ns) {
        int best = 0;
        long long best_cost = 0;
        for (int th = step; th <= std::min(Hc, 64); th += step) {
            long long sps = (long long)(th + 2) * Wp * 16;
            long long stage = 16 + KC * sps + TAIL_SLACK;
            long long staging = g.pool ? (long long)NCH * th * Wp * 16 : 0;
            long long tot = pl.wbytes + NP * 4 + 128 + ns * stage + 128 + staging + 256 + 256;
            if (sps / 16 >= 16384 || tot > SMEM_LIMIT) continue;
            long long tiles = 0, rows_in = 0;
            for (int y0 = 0; y0 < Hc; y0 += th) {
                tiles += (std::min(th, Hc - y0) * Wp + 127) / 128;
                rows_in += std::min(th, Hc - y0) + 2;
            }
            // Tiles first, halo as tie-break.  Pooled layers synchronise their epilogue warps per band, so
            // for them fewer, taller bands win over a slightly lower tile count (measured on layer 3).
            const long long cost = g.pool ? -th : tiles * 1000 + rows_in * 1000 / (Hc + 2);
            if (!best || cost <= best_cost) { best = th; best_cost = cost; }
        }
        if (best) {
            pl.TH = best;
            pl.n_stages = ns;
            break;
        }
        if (ns == 1) return false;
    }
    {   // Small bands (the narrow deep layers: a whole 22x27 or 13x7 image is 1-5 tiles) finish faster than a bulk copy
        // arrives, so two stages leave the MMA warps waiting for data: use the shared memory that is left for a deeper
        // ring (ASR_CONV_STAGES caps it; 2 = the previous behaviour).
        static const int cap = getenv("ASR_CONV_STAGES") ? std::max(1, std::min(MAX_STAGES, atoi(getenv("ASR_CONV_STAGES")))) : MAX_STAGES;
        const long long sps = (long long)(pl.TH + 2) * Wp * 16;
        const long long stage = ((16 + KC * sps + TAIL_SLACK) + 127) / 128 * 128;
        const long long staging = g.pool ? (long long)NCH * pl.TH * Wp * 16 : 0;
        while (pl.n_stages >= 2 && pl.n_stages < cap &&
               pl.wbytes + NP * 4 + 128 + (pl.n_stages + 1) * stage + 128 + staging + 256 + 256 <= SMEM_LIMIT)
            ++pl.n_stages;
    }
    pl.bands = (Hc + pl.TH - 1) / pl.TH;
    pl.MT = (pl.TH * Wp + 127) / 128;
    pl.sps = (pl.TH + 2) * Wp * 16;
    pl.stage_bytes = ((16 + KC * pl.sps + TAIL_SLACK) + 127) / 128 * 128;
    pl.staging_bytes = g.pool ? NCH * pl.TH * Wp * 16 : 0;
    int off = pl.wbytes;
    pl.off_bias = off; off += NP * 4; off = (off + 127) / 128 * 128;
    pl.off_stage = off; off += pl.n_stages * pl.stage_bytes;
    pl.off_staging = off; off += pl.staging_bytes; off = (off + 127) / 128 * 128;
    pl.off_bar = off; off += 256;
    pl.smem_bytes = off;
    pl.wp_magic = (unsigned)((0x100000000ull + (unsigned)Wp - 1) / (unsigned)Wp);
    for (unsigned o = 0; o < (unsigned)pl.MT * 128u; ++o)       // fast division must be exact for every o the kernel sees
        if ((unsigned)(((unsigned long long)o * pl.wp_magic) >> 32) != o / (unsigned)Wp) return false;
    pl.Hc = Hc;
    pl.wo_magic = 0;
    if (g.pool && g.Wo > 1) {
        pl.wo_magic = (unsigned)((0x100000000ull + (unsigned)g.Wo - 1) / (unsigned)g.Wo);
        for (unsigned e = 0; e < (unsigned)(pl.TH / 2 * g.Wo) + EPI_THREADS; ++e)
            if ((unsigned)(((unsigned long long)e * pl.wo_magic) >> 32) != e / (unsigned)g.Wo) return false;
    }
    return off <= SMEM_LIMIT;
}

// Launch plan of the row-stacked kernel; false when the layer does not qualify.
static bool plan_rows(const LayerGeom &g, ConvPlan &pl) {
    static const int enabled = getenv("ASR_CONV_ROWS") ? atoi(getenv("ASR_CONV_ROWS")) : 1;
    const int Wp = g.W + 2, KC = g.cinp / 8, NP = g.coutp;
    const int Hc = g.pool ? 2 * g.Ho : g.H;
    // Pooled layers only: without the in-register pooling the variant merely ties with the raster kernel
    // (measured on layer 2: 184-209 us vs 180 us per 1024 samples); ASR_CONV_ROWS=2 forces it for those too.
    if (!enabled || (!g.pool && enabled < 2)) return false;
    if (NP > 48 || Hc < RS_R || (Hc & 1) || KC / 2 > 3 || (KC & 1)) return false;
    pl = ConvPlan();
    // wide images: one sample per tile row, JT tiles across it, rows contiguous as in global memory;
    // narrow images: SPT samples side by side (pitch SP, even so that pooling pairs stay lane-aligned)
    const int SP = (Wp + 1) & ~1;
    if (g.W >= 90 && (!g.pool || !(g.W & 1))) {
        pl.SPT = 1; pl.SP = Wp; pl.RP = Wp; pl.JT = (g.W + 127) / 128;
        if (g.W * 10 < pl.JT * 128 * 7) return false;             // < 70 % of the MMA rows would be real pixels
    } else {
        pl.SPT = 128 / SP; pl.SP = SP; pl.RP = pl.SPT * SP; pl.JT = 1;
        // Measured on B200 (per 1024 samples): 40x50 48->48: 137 vs 142 us raster; 20x25: 61 vs 42; 92x42 12->12:
        // 83 vs 81; 23x10: 72 vs 39 -- the many small bulk copies and the 4-row bands (1.5x halo) eat the MMA
        // saving, so the side-by-side form stays off unless ASR_CONV_ROWS_MULTI=1.
        static const int multi = getenv("ASR_CONV_ROWS_MULTI") ? atoi(getenv("ASR_CONV_ROWS_MULTI")) : 0;
        if (!multi || pl.SPT < 2 || pl.SPT * g.W * 10 < 128 * 6) return false;
    }
    pl.rows = 1; pl.Hc = Hc;
    pl.wbytes = 3 * KC * RS_R * NP * 16;
    pl.slot_cols = RS_R * NP <= 64 ? 64 : (RS_R * NP <= 128 ? 128 : 256);
    pl.n_slots = std::min(MAX_SLOTS, 512 / pl.slot_cols);
    pl.tmem_cols = 512;
    pl.TH = 0;
    const int Hc4 = (Hc + RS_R - 1) / RS_R * RS_R;
    for (int th = RS_R; th <= std::min(Hc4, 64); th += RS_R) {
        long long sps = (long long)(th + 2) * pl.RP * 16;
        long long stage = 16 + KC * sps + TAIL_SLACK;
        long long tot = pl.wbytes + NP * 4 + 128 + 2 * stage + 128 + 256 + 256;
        if (sps / 16 >= 16384 || tot > SMEM_LIMIT) continue;
        pl.TH = th;
    }
    if (!pl.TH) return false;
    pl.n_stages = 2;
    pl.bands = (Hc + pl.TH - 1) / pl.TH;
    pl.MT = pl.TH / RS_R * pl.JT;
    pl.sps = (pl.TH + 2) * pl.RP * 16;
    pl.stage_bytes = ((16 + KC * pl.sps + TAIL_SLACK) + 127) / 128 * 128;
    pl.staging_bytes = 0;
    int off = pl.wbytes;
    pl.off_bias = off; off += NP * 4; off = (off + 127) / 128 * 128;
    pl.off_stage = off; off += pl.n_stages * pl.stage_bytes;
    pl.off_staging = off;
    pl.off_bar = off; off += 256;
    pl.smem_bytes = off;
    pl.wp_magic = pl.wo_magic = 0;
    // the last view of the last tile reads up to (TH + 1) * RP + 128 * JT + 2 positions
    if ((long long)((pl.TH + 1) * pl.RP + 128 * pl.JT + 2) * 16 > (long long)pl.sps + TAIL_SLACK) return false;
    return off <= SMEM_LIMIT;
}

static int pad16(int c) { return (c + 15) / 16 * 16; }

static void mark(asr_encoder *e, cudaStream_t st, int cat, bool start) {
    if (!e->timing) return;
    if (e->ev_used == e->ev_pool.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->ev_pool.push_back(ev);
        e->ev_cat.push_back(0);
    }
    if (start) e->ev_cat[e->ev_used / 2] = cat;
    e->ev_cat.resize(std::max(e->ev_cat.size(), e->ev_pool.size()));
    cudaEventRecord(e->ev_pool[e->ev_used++], st);
}

extern "C" {

int asr_encoder_destroy(asr_encoder_t *e) {
    if (!e) return ASR_OK;
    cudaFree(e->l0_w); cudaFree(e->l0_blob); cudaFree(e->l0_blob_u8); cudaFree(e->f01_blob); cudaFree(e->f01_blob_u8); cudaFree(e->f23_blobA);
    for (int l = 0; l < 8; ++l) {
        cudaFree(e->wblob[l]); cudaFree(e->act[l]); cudaFree(e->ref_w[l]); cudaFree(e->ref_bn[l]); cudaFree(e->ref_act[l]);
    }
    cudaFree(e->head_A); cudaFree(e->cca_mean); cudaFree(e->cca_proj); cudaFree(e->ref_in); cudaFree(e->ref_tmp);
    for (int b = 0; b < 2; ++b) {
        cudaFree(e->dev_in[b]);
        if (e->pin_in[b]) cudaFreeHost(e->pin_in[b]);
        if (e->pin_out[b]) cudaFreeHost(e->pin_out[b]);
        if (e->pin_out[2 + b]) cudaFreeHost(e->pin_out[2 + b]);
        if (e->ev_copied[b]) cudaEventDestroy(e->ev_copied[b]);
        if (e->ev_done[b]) cudaEventDestroy(e->ev_done[b]);
    }
    cudaFree(e->dev_codes); cudaFree(e->dev_lat);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->ev_final) cudaEventDestroy(e->ev_final);
    if (e->streams_owned && e->s_copy) cudaStreamDestroy(e->s_copy);
    if (e->streams_owned && e->s_comp) cudaStreamDestroy(e->s_comp);
    delete e;
    return ASR_OK;
}

int asr_encoder_set_cca(asr_encoder_t *e, const float *cca_mean, const float *cca_proj) {
    ASR_CHECK_ARG(e && cca_mean && cca_proj, "NULL argument");
    ASR_CUDA(cudaMemcpy(e->cca_mean, cca_mean, 32 * 4, cudaMemcpyHostToDevice));
    ASR_CUDA(cudaMemcpy(e->cca_proj, cca_proj, 1024 * 4, cudaMemcpyHostToDevice));
    return ASR_OK;
}

static int ensure_act(asr_encoder *e, int l);

int asr_encoder_create(asr_encoder_t **out, const asr_encoder_desc *d, int max_batch) {
    ASR_CHECK_ARG(out && d, "NULL argument");
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(max_batch >= 1, "max_batch < 1");
    ASR_CHECK_ARG(d->channels[8] == ASR_DIM, "last layer must have 32 channels");
    ASR_CHECK_ARG(d->prepare >= 0 && d->prepare <= 2, "bad prepare mode");
    for (int l = 0; l < 9; ++l)
        ASR_CHECK_ARG(d->W[l] && d->beta[l] && d->gamma[l] && d->mean[l] && d->inv_std[l], "NULL layer parameter");
    ASR_CHECK_ARG(d->cca_mean && d->cca_proj, "NULL CCA parameter");
    ASR_CHECK_ARG(d->channels[0] <= L0_MAXC, "layer 0 supports at most 32 channels");
    asr_encoder *e = new asr_encoder();
    e->d = *d;
    e->max_batch = max_batch;
    e->device = current_device();
    e->H0 = d->prepare == ASR_PREP_SCALE_HALF ? d->in_h / 2 : d->in_h;
    e->W0 = d->prepare == ASR_PREP_SCALE_HALF ? d->in_w / 2 : d->in_w;
    int H = e->H0, W = e->W0, cin = 1;
    e->flops = 0;
    for (int l = 0; l < 8; ++l) {
        LayerGeom &g = e->g[l];
        g.cin = cin; g.cout = d->channels[l]; g.cinp = pad16(cin); g.coutp = pad16(g.cout);
        g.H = H; g.W = W; g.pool = (l & 1);
        g.Ho = g.pool ? H / 2 : H; g.Wo = g.pool ? W / 2 : W;
        e->flops += 2.0 * 9.0 * cin * g.cout * H * W;
        H = g.Ho; W = g.Wo; cin = g.cout;
        if (H < 1 || W < 1) { delete e; set_error("input too small for four 2x2 pools"); return ASR_ERR_ARG; }
        const char *bad = g.coutp > 128 ? "at most 128 channels per layer"
                          : !(l == 0 || g.cinp == 16 || g.cinp == 32 || g.cinp == 48 || g.cinp == 64 || g.cinp == 96 || g.cinp == 128)
                              ? "input channels must pad to 16, 32, 48, 64, 96 or 128" : nullptr;
        if (bad) { delete e; set_error(std::string("asr_encoder_create: ") + bad); return ASR_ERR_ARG; }
    }
    e->head_c = cin; e->head_h = H; e->head_w = W;
    e->flops += 2.0 * cin * 32 * H * W;
    if (e->head_c > 128) { delete e; set_error("asr_encoder_create: head supports at most 128 input channels"); return ASR_ERR_ARG; }

    const size_t B = (size_t)max_batch;
#define E_CUDA(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            set_error(std::string("asr_encoder_create: " #expr " -> ") + cudaGetErrorString(_e));      \
            asr_encoder_destroy(e);                                                                    \
            return ASR_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

    // ---- fold BN, pack weights ----
    std::vector<float> scale, bias;
    for (int l = 0; l < 8; ++l) {
        const LayerGeom &g = e->g[l];
        scale.assign(g.cout, 0.f); bias.assign(g.cout, 0.f);
        for (int c = 0; c < g.cout; ++c) {
            scale[c] = d->gamma[l][c] * d->inv_std[l][c];
            bias[c] = d->beta[l][c] - d->mean[l][c] * scale[c];
        }
        // fp32 path: raw (optionally flipped) weights + bn vectors
        std::vector<float> wr((size_t)g.cout * g.cin * 9);
        for (int co = 0; co < g.cout; ++co)
            for (int ci = 0; ci < g.cin; ++ci)
                for (int t = 0; t < 9; ++t)
                    wr[((size_t)co * g.cin + ci) * 9 + t] = d->W[l][((size_t)co * g.cin + ci) * 9 + (d->flip_filters ? 8 - t : t)];
        std::vector<float> bn(4 * (size_t)g.cout);
        for (int c = 0; c < g.cout; ++c) {
            bn[c] = d->beta[l][c]; bn[g.cout + c] = d->gamma[l][c];
            bn[2 * g.cout + c] = d->mean[l][c]; bn[3 * g.cout + c] = d->inv_std[l][c];
        }
        E_CUDA(cudaMalloc(&e->ref_w[l], wr.size() * 4));
        E_CUDA(cudaMemcpy(e->ref_w[l], wr.data(), wr.size() * 4, cudaMemcpyHostToDevice));
        E_CUDA(cudaMalloc(&e->ref_bn[l], bn.size() * 4));
        E_CUDA(cudaMemcpy(e->ref_bn[l], bn.data(), bn.size() * 4, cudaMemcpyHostToDevice));
        if (l == 0) {
            std::vector<float> w0((size_t)g.cout * 10);
            for (int c = 0; c < g.cout; ++c) {
                for (int t = 0; t < 9; ++t) w0[c * 9 + t] = wr[(size_t)c * 9 + t] * scale[c];
                w0[(size_t)g.cout * 9 + c] = bias[c];
            }
            memset(&e->l0_host, 0, sizeof(e->l0_host));
            for (int c = 0; c < g.cout; ++c) {
                for (int t = 0; t < 9; ++t) e->l0_host.w[c * 9 + t] = w0[(size_t)c * 9 + t];
                e->l0_host.b[c] = w0[(size_t)g.cout * 9 + c];
            }
            E_CUDA(cudaMalloc(&e->l0_w, w0.size() * 4));
            E_CUDA(cudaMemcpy(e->l0_w, w0.data(), w0.size() * 4, cudaMemcpyHostToDevice));
            static const int l0_tc = getenv("ASR_L0_TC") ? atoi(getenv("ASR_L0_TC")) : 1;
            if (l0_tc && (g.cout == 12 || g.cout == 24)) {
                // banded weight matrices of l0_tc_kernel: B_dx[k][(r, co)] = w[co][dy = k - r][dx], split hi + lo
                const int C = g.cout, R0 = C == 12 ? 12 : 8, NPAD = R0 * C;
                const size_t bbytes = (size_t)3 * 2 * 2 * NPAD * 16;
                auto make_blob = [&](int R0b, double pixel_scale) {
                    const int NPADb = R0b * C;
                    const size_t bb_bytes = (size_t)3 * 2 * 2 * NPADb * 16;
                    std::vector<uint8_t> blob(bb_bytes + C * 4, 0);
                    bf16 *wb = reinterpret_cast<bf16 *>(blob.data());
                    auto put = [&](int dx, int k, int nn, float w) {
                        const bf16 hi = __float2bfloat16_rn(w);
                        const bf16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
                        wb[((((size_t)dx * 2 + 0) * 2 + k / 8) * NPADb + nn) * 8 + (k & 7)] = hi;
                        wb[((((size_t)dx * 2 + 1) * 2 + k / 8) * NPADb + nn) * 8 + (k & 7)] = lo;
                    };
                    for (int dx = 0; dx < 3; ++dx)
                        for (int r = 0; r < R0b; ++r)
                            for (int dy = 0; dy < 3; ++dy)
                                for (int co = 0; co < C; ++co)
                                    put(dx, r + dy, r * C + co, (float)((double)w0[(size_t)co * 9 + dy * 3 + dx] * pixel_scale));
                    for (int r = 0; r < R0b; ++r)           // K row 15 of dx = 0: the bias (the A tile holds 1 there)
                        for (int co = 0; co < C; ++co) put(0, 15, r * C + co, w0[(size_t)C * 9 + co]);
                    float *bb = reinterpret_cast<float *>(blob.data() + bb_bytes);
                    for (int co = 0; co < C; ++co) bb[co] = w0[(size_t)C * 9 + co];
                    return blob;
                };
                const std::vector<uint8_t> blob = make_blob(R0, 1.0);
                const std::vector<uint8_t> blob_u8 = make_blob(R0, 1.0 / 255.0);
                (void)bbytes;
                if (C == F_C0 && d->prepare != ASR_PREP_SCALE_HALF) {     // blobs of the fused layer-0 + layer-1 kernel (8 rows per tile)
                    const std::vector<uint8_t> fb = make_blob(F_R0, 1.0);
                    E_CUDA(cudaMalloc(&e->f01_blob, fb.size()));
                    E_CUDA(cudaMemcpy(e->f01_blob, fb.data(), fb.size(), cudaMemcpyHostToDevice));
                    if (d->prepare == ASR_PREP_SCALE) {
                        const std::vector<uint8_t> fb8 = make_blob(F_R0, 1.0 / 255.0);
                        E_CUDA(cudaMalloc(&e->f01_blob_u8, fb8.size()));
                        E_CUDA(cudaMemcpy(e->f01_blob_u8, fb8.data(), fb8.size(), cudaMemcpyHostToDevice));
                    }
                }
                const unsigned Wp = (unsigned)g.W + 2;
                const unsigned magic = (unsigned)((0x100000000ull + Wp - 1) / Wp);
                bool exact = true;
                for (unsigned gg = 0; gg < (unsigned)max_batch * Wp + 512u && exact; ++gg)
                    exact = (unsigned)(((unsigned long long)gg * magic) >> 32) == gg / Wp;
                if (exact) {
                    E_CUDA(cudaMalloc(&e->l0_blob, blob.size()));
                    E_CUDA(cudaMemcpy(e->l0_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
                    if (d->prepare == ASR_PREP_SCALE) {
                        E_CUDA(cudaMalloc(&e->l0_blob_u8, blob_u8.size()));
                        E_CUDA(cudaMemcpy(e->l0_blob_u8, blob_u8.data(), blob_u8.size(), cudaMemcpyHostToDevice));
                    }
                    e->l0_R0 = R0; e->l0_NPAD = NPAD; e->l0_wp_magic = magic;
                }
            }
        } else {
            if (!plan_rows(g, e->plan[l]) && !plan_conv(g, e->plan[l])) {
                set_error("asr_encoder_create: layer " + std::to_string(l) + " does not fit in shared memory");
                asr_encoder_destroy(e);
                return ASR_ERR_UNSUPPORTED;
            }
            if (getenv("ASR_DEBUG_PLAN")) {
                const ConvPlan &pl = e->plan[l];
                fprintf(stderr, "[asr] layer %d: %dx%d cin %d->%d (pad %d->%d) pool %d | %s TH %d bands %d MT %d stages %d slots %d x %d cols "
                        "smem %d B (w %d, stage %d, staging %d)\n", l, g.H, g.W, g.cin, g.cout, g.cinp, g.coutp, g.pool,
                        pl.rows ? (pl.SPT > 1 ? "rows-multi" : "rows") : "raster", pl.TH,
                        pl.bands, pl.MT, pl.n_stages, pl.n_slots, pl.slot_cols, pl.smem_bytes, pl.wbytes, pl.stage_bytes,
                        pl.staging_bytes);
            }
            const int KC = g.cinp / 8, NP = g.coutp;
            auto pack_blob = [&](bool rows_format) {
                const size_t wbytes = rows_format ? (size_t)3 * KC * RS_R * NP * 16 : (size_t)9 * KC * NP * 16;
                std::vector<uint8_t> blob(wbytes + NP * 4, 0);
                bf16 *wb = reinterpret_cast<bf16 *>(blob.data());
                for (int t = 0; t < 9; ++t)
                    for (int ci = 0; ci < g.cin; ++ci)
                        for (int co = 0; co < g.cout; ++co) {
                            float v = wr[((size_t)co * g.cin + ci) * 9 + t] * scale[co];
                            size_t idx;
                            if (rows_format) {          // [dx][K chunk][block: dy = 2, 1, 0, zeros][NP][8]
                                const int dy = t / 3, dx = t % 3;
                                idx = ((((size_t)dx * KC + ci / 8) * RS_R + (2 - dy)) * NP + co) * 8 + (ci & 7);
                            } else {                    // [tap][K chunk][NP][8]
                                idx = (((size_t)t * KC + ci / 8) * NP + co) * 8 + (ci & 7);
                            }
                            wb[idx] = __float2bfloat16_rn(v);
                        }
                float *bb = reinterpret_cast<float *>(blob.data() + wbytes);
                for (int co = 0; co < g.cout; ++co) bb[co] = bias[co];
                return blob;
            };
            const std::vector<uint8_t> blob = pack_blob(e->plan[l].rows != 0);
            if (blob.size() != (size_t)e->plan[l].wbytes + NP * 4) {
                set_error("asr_encoder_create: weight blob size does not match the launch plan");
                asr_encoder_destroy(e);
                return ASR_ERR_UNSUPPORTED;
            }
            E_CUDA(cudaMalloc(&e->wblob[l], blob.size()));
            E_CUDA(cudaMemcpy(e->wblob[l], blob.data(), blob.size(), cudaMemcpyHostToDevice));
            if (l == 2 && !g.pool && !e->plan[l].rows) {      // layer A of the fused layer-2 + layer-3 kernel is row-stacked
                const std::vector<uint8_t> rb = pack_blob(true);
                E_CUDA(cudaMalloc(&e->f23_blobA, rb.size()));
                E_CUDA(cudaMemcpy(e->f23_blobA, rb.data(), rb.size(), cudaMemcpyHostToDevice));
            }
        }
        // activations of both paths
        if (l == 0) {
            // Optional experiment (off by default): run layers 0 and 1 over sub-chunks that reuse one L2-sized
            // region of the layer-0 buffer.  Measured on B200 it LOSES (479k -> 409k pairs/s at 128 samples,
            // worse below): the small launches cost more than the HBM round trip they save.
            long long c = getenv("ASR_L01_CHUNK") ? atoll(getenv("ASR_L01_CHUNK")) : 0;
            e->l01_chunk = (c <= 0 || c >= max_batch) ? 0 : std::max<long long>(16, c / 16 * 16);
        }
        e->act_plane[l] = (long long)(g.Ho + 2) * (g.Wo + 2) * 16;
        e->act_sample[l] = e->act_plane[l] * (g.coutp / 8);
        if (l != 0 && l != 2) {   // buffers of layers 0 and 2: ensure_act (first unfused call; never while they run fused)
            E_CUDA(cudaMalloc(&e->act[l], (size_t)e->act_sample[l] * B));
            E_CUDA(cudaMemset(e->act[l], 0, (size_t)e->act_sample[l] * B));
        }
    }
    {   // layers 0 + 1 fused: eligibility, geometry and shared-memory layout of l01_fused_kernel
        const LayerGeom &g0 = e->g[0], &g1 = e->g[1];
        const ConvPlan &p1 = e->plan[1];
        F01Params &f = e->f01;
        memset(&f, 0, sizeof(f));
        const int Wp = g0.W + 2;
        bool ok = e->f01_blob && e->l0_blob && g0.cout == F_C0 && g0.W >= 126 && g0.H % F_R0 == 0 && g1.pool &&
                  g1.cinp == 16 && g1.coutp == F_NP1 && p1.rows && p1.SPT == 1 && p1.Hc == g0.H && g0.H % RS_R == 0;
        if (ok) {
            f.H = g0.H; f.W = g0.W; f.Wp = Wp;
            f.NB = g0.H / F_R0; f.NG = g0.H / RS_R; f.T0 = (f.NB * Wp + 127) / 128; f.JT = p1.JT;
            f.wp_magic = (unsigned)((0x100000000ull + (unsigned)Wp - 1) / (unsigned)Wp);
            for (unsigned gg = 0; gg < (unsigned)f.T0 * 128u + 256u && ok; ++gg)
                ok = (unsigned)(((unsigned long long)gg * f.wp_magic) >> 32) == gg / (unsigned)Wp;
            f.ring_plane = (F_ZROW + 1) * Wp * 16;
            f.Ho = g1.Ho; f.Wo = g1.Wo; f.Wpo = g1.Wo + 2; f.cout1 = g1.cout;
            f.out_plane = e->act_plane[1]; f.out_sample = e->act_sample[1];
            auto layout = [&](int abuf) {
                int off = 3 * 2 * 2 * F_NPAD0 * 16;
                f.off_w1 = off; off += F_W1BYTES + F_NP1 * 4; off = (off + 127) / 128 * 128;
                f.off_a = off; off += F_CW * abuf; off = (off + 127) / 128 * 128;
                f.off_ring = off; off += 2 * f.ring_plane + F_TAIL; off = (off + 127) / 128 * 128;
                f.off_lut = off; off += 256 * 4;
                f.off_bar = off; off += 256;
                return off;
            };
            e->f01_smem_int = layout(2 * F_AROWS * 16);          // integer pixels: hi part only
            e->f01_smem = layout(2 * 2 * F_AROWS * 16);          // hi + lo A parts (fp32 pixels); f keeps these offsets
            // a view of the last tile of a row may read up to 128 JT + 2 positions of a ring row
            ok = ok && e->f01_smem <= SMEM_LIMIT && (128 * f.JT + 2 - Wp) * 16 <= F_TAIL;
        }
        e->f01_ok = ok;
        static const int fuse_env = getenv("ASR_FUSE01") ? atoi(getenv("ASR_FUSE01")) : 1;
        e->fuse_mask = (ok && fuse_env) ? 1 : 0;
        if (getenv("ASR_DEBUG_PLAN"))
            fprintf(stderr, "[asr] layers 0+1 fused: %s (NB %d NG %d T0 %d JT %d smem %d / %d B)\n", ok ? "available" : "no", f.NB,
                    f.NG, f.T0, f.JT, e->f01_smem, e->f01_smem_int);
    }
    {   // layers 2 + 3 fused: eligibility, geometry and shared-memory layout of l23_fused_kernel
        const LayerGeom &ga = e->g[2], &gb = e->g[3];
        const ConvPlan &pb = e->plan[3];
        F23Params &f = e->f23;
        memset(&f, 0, sizeof(f));
        const int Wp = ga.W + 2;
        bool ok = e->f23_blobA && !ga.pool && gb.pool && ga.W == gb.W && ga.H == gb.H && ga.W >= 90 && ga.W <= 128 &&
                  !(ga.W & 1) && ga.H % G_TH == 0 && ga.cinp == 16 && ga.coutp == 32 && gb.cinp == 32 && gb.coutp == 32 &&
                  pb.rows && pb.SPT == 1 && pb.JT == 1 && pb.Hc == ga.H &&
                  (ga.cin + 7) / 8 == ga.cinp / 8;     // every input plane of layer A is loaded (none left uninitialised)
        if (ok) {
            f.H = ga.H; f.W = ga.W; f.Wp = Wp; f.NG = ga.H / RS_R; f.bands = ga.H / G_TH;
            f.KCLA = (ga.cin + 7) / 8; f.NPA = ga.coutp; f.NPB = gb.coutp; f.coutA = ga.cout; f.coutB = gb.cout;
            f.Ho = gb.Ho; f.Wo = gb.Wo; f.Wpo = gb.Wo + 2;
            f.in_plane = e->act_plane[1]; f.in_sample = e->act_sample[1];
            f.out_plane = e->act_plane[3]; f.out_sample = e->act_sample[3];
            const int KCA = ga.cinp / 8, KCB = gb.cinp / 8;
            f.wbytesA = 3 * KCA * RS_R * f.NPA * 16; f.wbytesB = 3 * KCB * RS_R * f.NPB * 16;
            f.sps = (G_TH + 2) * Wp * 16;
            f.stage_bytes = ((16 + KCA * f.sps + TAIL_SLACK) + 127) / 128 * 128;
            f.ring_plane = (G_ZROW + 1) * Wp * 16;
            int off = f.wbytesA + f.NPA * 4; off = (off + 127) / 128 * 128;
            f.off_wB = off; off += f.wbytesB + f.NPB * 4; off = (off + 127) / 128 * 128;
            f.off_stage = off;
            // K pairing (G_PAIR): the all-padding last chunk of layer B's input is never read, so it gets no ring plane
            f.ring_planes = (G_PAIR && gb.cin <= 8 * (KCB - 1)) ? KCB - 1 : KCB;
            static const int st_env = getenv("ASR_F23_STAGES") ? atoi(getenv("ASR_F23_STAGES")) : 3;
            const int ring_bytes = f.ring_planes * f.ring_plane + G_TAIL;
            f.n_stages = (st_env >= 3 && off + 3 * f.stage_bytes + ring_bytes + 128 + 256 <= SMEM_LIMIT) ? 3 : 2;
            off += f.n_stages * f.stage_bytes;
            f.off_ring = off; off += ring_bytes; off = (off + 127) / 128 * 128;
            f.off_bar = off; off += 256;
            e->f23_smem = off;
            ok = off <= SMEM_LIMIT && f.wbytesB == pb.wbytes && (128 + 3 - Wp) * 16 <= G_TAIL && f.ring_plane / 16 < 16384 &&
                 (!G_PAIR || f.ring_planes == KCB - 1) &&
                 (long long)((G_TH + 1) * Wp + 128 + 2) * 16 <= (long long)f.sps + TAIL_SLACK;
        }
        e->f23_ok = ok;
        static const int fuse_env = getenv("ASR_FUSE23") ? atoi(getenv("ASR_FUSE23")) : 1;
        if (ok && fuse_env) e->fuse_mask |= 2;
        if (getenv("ASR_DEBUG_PLAN"))
            fprintf(stderr, "[asr] layers 2+3 fused: %s (NG %d bands %d stages %d smem %d B)\n", ok ? "available" : "no", f.NG, f.bands,
                    f.n_stages, e->f23_smem);
    }
    {   // head: A[j][c] = W8[j][c] * scale8[j]; b[j] = beta8 - mean8*scale8
        const int C = e->head_c;
        std::vector<float> A((size_t)32 * C + 32);
        for (int j = 0; j < 32; ++j) {
            float s = d->gamma[8][j] * d->inv_std[8][j];
            for (int c = 0; c < C; ++c) A[(size_t)j * C + c] = d->W[8][(size_t)j * C + c] * s;
            A[(size_t)32 * C + j] = d->beta[8][j] - d->mean[8][j] * s;
        }
        E_CUDA(cudaMalloc(&e->head_A, A.size() * 4));
        E_CUDA(cudaMemcpy(e->head_A, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        E_CUDA(cudaMalloc(&e->cca_mean, 32 * 4));
        E_CUDA(cudaMalloc(&e->cca_proj, 1024 * 4));
        E_CUDA(cudaMemcpy(e->cca_mean, d->cca_mean, 32 * 4, cudaMemcpyHostToDevice));
        E_CUDA(cudaMemcpy(e->cca_proj, d->cca_proj, 1024 * 4, cudaMemcpyHostToDevice));
    }
    static bool attr_done[ASR_MAX_DEVICES] = {false};     // cudaFuncSetAttribute is per device
    const int attr_dev = std::max(0, std::min(current_device(), ASR_MAX_DEVICES - 1));
    if (!attr_done[attr_dev]) {
        E_CUDA(cudaFuncSetAttribute(l0_tc_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(l0_tc_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(l01_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(l01_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA((cudaFuncSetAttribute(l01_fused_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)));
        E_CUDA((cudaFuncSetAttribute(l23_fused_kernel<1, 2, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)));
        E_CUDA(cudaFuncSetAttribute(conv3x3_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        E_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_done[attr_dev] = true;
    }
#undef E_CUDA
    // Activation buffers of layers whose output stays on the SM in the default configuration are only allocated when a
    // caller switches that fusion off (asr_encoder_set_fusion); everything the default path needs exists from here on,
    // so asr_encoder_embed never allocates.
    for (int l = 0; l <= 2; l += 2)
        if (!(e->fuse_mask & (l == 0 ? 1 : 2)) && ensure_act(e, l) != ASR_OK) {
            asr_encoder_destroy(e);
            return ASR_ERR_CUDA;
        }
    *out = e;
    return ASR_OK;
}

double asr_encoder_flops_per_sample(const asr_encoder_t *e) { return e ? e->flops : 0.0; }

static int ensure_act(asr_encoder *e, int l) {
    if (e->act[l]) return ASR_OK;
    ASR_CUDA(cudaMalloc(&e->act[l], (size_t)e->act_sample[l] * e->max_batch));
    ASR_CUDA(cudaMemset(e->act[l], 0, (size_t)e->act_sample[l] * e->max_batch));
    return ASR_OK;
}

static int ensure_ref_buffers(asr_encoder *e) {
    if (e->ref_in) return ASR_OK;
    const size_t B = (size_t)e->max_batch;
    ASR_CUDA(cudaMalloc(&e->ref_in, B * e->H0 * e->W0 * 4));
    size_t tmp = 0;
    for (int l = 0; l < 8; ++l) {
        const LayerGeom &g = e->g[l];
        ASR_CUDA(cudaMalloc(&e->ref_act[l], B * g.cout * g.Ho * g.Wo * 4));
        if (g.pool) tmp = std::max(tmp, B * g.cout * g.H * g.W * 4);
    }
    ASR_CUDA(cudaMalloc(&e->ref_tmp, tmp));
    return ASR_OK;
}

int asr_encoder_embed(asr_encoder_t *e, const void *x_dev, int x_dtype, int64_t n, float *codes_dev, float *latents_dev,
                      int path, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(e && x_dev, "NULL argument");
    DeviceGuard guard(e->device);
    ASR_CHECK_ARG(n >= 0 && n <= e->max_batch, "n exceeds max_batch");
    ASR_CHECK_ARG(x_dtype == ASR_IN_F32 || x_dtype == ASR_IN_U8, "bad x_dtype");
    ASR_CHECK_ARG(path == ASR_PATH_TCGEN05 || path == ASR_PATH_FP32, "bad path");
    if (n == 0) return ASR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const asr_encoder_desc &d = e->d;
    HeadParams hp;
    hp.C = e->head_c; hp.H = e->head_h; hp.W = e->head_w; hp.Wp = e->head_w + 2;
    hp.A = e->head_A; hp.cca_mean = e->cca_mean; hp.cca_proj = e->cca_proj;
    hp.codes = codes_dev; hp.latents = latents_dev;
    e->last_path = path;
    if (path == ASR_PATH_TCGEN05) {
        auto launch_l0 = [&](int64_t n0, int64_t nn, bf16 *out) -> int {
            const LayerGeom &g = e->g[0];
            L0Params p;
            const size_t esz = x_dtype == ASR_IN_U8 ? 1 : 4;
            p.x = reinterpret_cast<const uint8_t *>(x_dev) + (size_t)n0 * d.in_h * d.in_w * esz;
            p.x_u8 = x_dtype == ASR_IN_U8; p.prepare = d.prepare;
            p.Hin = d.in_h; p.Win = d.in_w; p.H = g.H; p.W = g.W; p.Wp = g.W + 2; p.Hp = g.H + 2;
            p.C = g.cout; p.NCH = g.coutp / 8; p.w = e->l0_w; p.out = out;
            p.out_plane = e->act_plane[0]; p.out_sample = e->act_sample[0]; p.n = (int)nn;
            if (e->l0_blob) {
                L0TcParams q;
                q.int_pixels = (p.x_u8 && p.prepare == ASR_PREP_SCALE && e->l0_blob_u8) ? 1 : 0;
                q.b = p; q.blob = q.int_pixels ? e->l0_blob_u8 : e->l0_blob; q.R0 = e->l0_R0; q.NPAD = e->l0_NPAD;
                q.TY = (g.H + q.R0 - 1) / q.R0;
                q.GT = (int)(((long long)nn * p.Wp + 127) / 128);
                q.wp_magic = e->l0_wp_magic;
                static const int l0_dbg = getenv("ASR_L0_DEBUG") ? atoi(getenv("ASR_L0_DEBUG")) : 0;
                q.dbg = l0_dbg;
                const long long tiles = (long long)q.GT * q.TY;
                const int grid_tc = (int)std::min<long long>(tiles, sm_count());
                const int smem_tc = 3 * 2 * 2 * q.NPAD * 16 + L0T_NBUF * 2 * 2 * L0T_AROWS * 16 + 32 * 4 + 256 * 4 + 256;
                if (g.cout == 12) l0_tc_kernel<12><<<grid_tc, L0T_THREADS, smem_tc, st>>>(q);
                else l0_tc_kernel<24><<<grid_tc, L0T_THREADS, smem_tc, st>>>(q);
                ASR_LAUNCH_CHECK();
                return ASR_OK;
            }
            dim3 grid((g.W + L0_TW - 1) / L0_TW, (g.H + L0_TH - 1) / L0_TH, (unsigned)nn);
            if (g.cout == 12) l0_conv_kernel<12><<<grid, L0_TW * L0_TROWS, 0, st>>>(p, e->l0_host);
            else if (g.cout == 24) l0_conv_kernel<24><<<grid, L0_TW * L0_TROWS, 0, st>>>(p, e->l0_host);
            else l0_conv_kernel<0><<<grid, L0_TW * L0_TROWS, 0, st>>>(p, e->l0_host);
            ASR_LAUNCH_CHECK();
            return ASR_OK;
        };
        auto launch_conv = [&](int l, const bf16 *in, bf16 *out, int64_t nn) -> int {
            const LayerGeom &g = e->g[l];
            const ConvPlan &pl = e->plan[l];
            ConvParams p;
            p.in = in; p.out = out; p.wblob = e->wblob[l]; p.n_samples = (int)nn;
            p.H = g.H; p.W = g.W; p.Wp = g.W + 2; p.Hp = g.H + 2; p.KC = g.cinp / 8; p.NP = g.coutp; p.NCH = g.coutp / 8;
            p.KCL = (g.cin + 7) / 8; p.NCHR = (g.cout + 7) / 8;
            p.TH = pl.TH; p.bands = pl.bands; p.MT = pl.MT; p.pool = g.pool; p.cout = g.cout;
            p.Ho = g.Ho; p.Wo = g.Wo; p.Wpo = g.Wo + 2;
            p.in_plane = e->act_plane[l - 1]; p.in_sample = e->act_sample[l - 1];
            p.out_plane = e->act_plane[l]; p.out_sample = e->act_sample[l];
            p.sps = pl.sps; p.stage_bytes = pl.stage_bytes; p.n_stages = pl.n_stages; p.slot_cols = pl.slot_cols;
            p.n_slots = pl.n_slots; p.tmem_cols = pl.tmem_cols; p.wbytes = pl.wbytes;
            p.off_bias = pl.off_bias; p.off_stage = pl.off_stage; p.off_staging = pl.off_staging; p.off_bar = pl.off_bar;
            p.wp_magic = pl.wp_magic; p.wo_magic = pl.wo_magic; p.Hc = pl.Hc; p.JT = pl.JT;
            p.coop = pl.rows && pl.n_slots < 8;
            p.SPT = pl.rows ? pl.SPT : 1; p.SP = pl.SP; p.RP = pl.RP;
            p.sp_magic = pl.rows && pl.SPT > 1 ? (unsigned)((0x100000000ull + (unsigned)pl.SP - 1) / (unsigned)pl.SP) : 0u;
            static const int conv_dbg = getenv("ASR_CONV_DEBUG") ? atoi(getenv("ASR_CONV_DEBUG")) : 0;
            p.dbg = conv_dbg;
            static const int pair_env = getenv("ASR_CONV_PAIR") ? atoi(getenv("ASR_CONV_PAIR")) : 1;
            // descriptor low words carry the 14-bit start address in bits 0-13: the masks in the kernel keep 16 bits,
            // which is exact because shared-memory addresses stay below 2^18 bytes
            p.pair = (pair_env && !pl.rows && p.KC >= 4 && p.KCL == p.KC - 1) ? 1 : 0;
            const int items = (int)nn * pl.bands;
            const int grid = std::min(items, sm_count());
            if (pl.rows) {
                const int sgroups = ((int)nn + pl.SPT - 1) / pl.SPT;
                const int rgrid = std::min(sgroups * pl.bands, sm_count());
                if (p.KC / 2 == 1) conv3x3_rows_kernel<1><<<rgrid, CONV_THREADS, pl.smem_bytes, st>>>(p);
                else if (p.KC / 2 == 2) conv3x3_rows_kernel<2><<<rgrid, CONV_THREADS, pl.smem_bytes, st>>>(p);
                else conv3x3_rows_kernel<3><<<rgrid, CONV_THREADS, pl.smem_bytes, st>>>(p);
                ASR_LAUNCH_CHECK();
                return ASR_OK;
            }
            switch (p.KC / 2) {
                case 1: conv3x3_tc_kernel<1><<<grid, CONV_THREADS, pl.smem_bytes, st>>>(p); break;
                case 2: conv3x3_tc_kernel<2><<<grid, CONV_THREADS, pl.smem_bytes, st>>>(p); break;
                case 3: conv3x3_tc_kernel<3><<<grid, CONV_THREADS, pl.smem_bytes, st>>>(p); break;
                case 4: conv3x3_tc_kernel<4><<<grid, CONV_THREADS, pl.smem_bytes, st>>>(p); break;
                case 6: conv3x3_tc_kernel<6><<<grid, CONV_THREADS, pl.smem_bytes, st>>>(p); break;
                case 8: conv3x3_tc_kernel<8><<<grid, CONV_THREADS, pl.smem_bytes, st>>>(p); break;
                default: set_error("asr_encoder_embed: unsupported input channel count"); return ASR_ERR_UNSUPPORTED;
            }
            ASR_LAUNCH_CHECK();
            return ASR_OK;
        };
        auto launch_fused01 = [&]() -> int {
            F01Params f = e->f01;
            f.x = x_dev; f.x_u8 = x_dtype == ASR_IN_U8; f.prepare = d.prepare; f.n = (int)n;
            f.int_pixels = (f.x_u8 && d.prepare == ASR_PREP_SCALE && e->f01_blob_u8) ? 1 : 0;
            f.blob0 = f.int_pixels ? e->f01_blob_u8 : e->f01_blob;
            f.wblob1 = reinterpret_cast<const uint8_t *>(e->wblob[1]);
            f.out = e->act[1];
            f.abuf = (f.int_pixels ? 2 : 4) * F_AROWS * 16;
            int smem = e->f01_smem;
            if (f.int_pixels) {       // same order of regions, smaller A buffers
                const int delta = F_CW * 2 * F_AROWS * 16;
                f.off_ring -= delta; f.off_lut -= delta; f.off_bar -= delta;
                smem = e->f01_smem_int;
            }
            static const int f01_dbg = getenv("ASR_F01_DEBUG") ? atoi(getenv("ASR_F01_DEBUG")) : 0;
            f.dbg = f01_dbg;
            const int grid = (int)std::min<int64_t>(n, sm_count());
            static const int variant = getenv("ASR_F01_VARIANT") ? atoi(getenv("ASR_F01_VARIANT")) : 0;
            if (variant == 1) l01_fused_kernel<1><<<grid, 32 * (F_DRAIN_WARP0 + 4 * (1 + F_EG)), smem, st>>>(f);
            else if (variant == 2) l01_fused_kernel<4, 2><<<grid, 32 * (F_DRAIN_WARP0 + 4 * (4 + 2)), smem, st>>>(f);
            else l01_fused_kernel<2><<<grid, 32 * (F_DRAIN_WARP0 + 4 * (2 + F_EG)), smem, st>>>(f);
            ASR_LAUNCH_CHECK();
            return ASR_OK;
        };
        auto launch_fused23 = [&]() -> int {
            F23Params f = e->f23;
            f.in = e->act[1]; f.out = e->act[3]; f.n = (int)n;
            f.wblobA = e->f23_blobA; f.wblobB = reinterpret_cast<const uint8_t *>(e->wblob[3]);
            static const int f23_dbg = getenv("ASR_F23_DEBUG") ? atoi(getenv("ASR_F23_DEBUG")) : 0;
            f.dbg = f23_dbg;
            const int grid = (int)std::min<int64_t>(n, sm_count());
            l23_fused_kernel<1, 2, 32, 32><<<grid, G_THREADS, e->f23_smem, st>>>(f);
            ASR_LAUNCH_CHECK();
            return ASR_OK;
        };
        const bool fused01 = (e->fuse_mask & 1) != 0;
        const bool fused23 = (e->fuse_mask & 2) != 0;
        if (!fused23 && (rc = ensure_act(e, 2))) return rc;
        if (fused01) {
            mark(e, st, 1, true);
            if ((rc = launch_fused01())) return rc;
            mark(e, st, 1, false);
        } else if ((rc = ensure_act(e, 0))) {
            return rc;
        }
        // Layers 0 and 1 can run over sub-chunks that reuse one region of the layer-0 buffer (see create).
        const int64_t sub = e->l01_chunk > 0 ? e->l01_chunk : n;
        for (int64_t n0 = 0; n0 < n && !fused01; n0 += sub) {
            const int64_t nn = std::min<int64_t>(sub, n - n0);
            mark(e, st, 0, true);
            if ((rc = launch_l0(n0, nn, e->act[0]))) return rc;
            mark(e, st, 0, false);
            mark(e, st, 1, true);
            if ((rc = launch_conv(1, e->act[0], reinterpret_cast<bf16 *>(reinterpret_cast<uint8_t *>(e->act[1]) +
                                                                           (size_t)n0 * e->act_sample[1]), nn)))
                return rc;
            mark(e, st, 1, false);
        }
        mark(e, st, 1, true);
        for (int l = 2; l < 8; ++l) {
            if (l == 2 && fused23) {
                if ((rc = launch_fused23())) return rc;
                ++l;
                continue;
            }
            if ((rc = launch_conv(l, e->act[l - 1], e->act[l], n))) return rc;
        }
        mark(e, st, 1, false);
        mark(e, st, 2, true);
        hp.in = e->act[7]; hp.is_p8 = 1; hp.plane = e->act_plane[7]; hp.sample = e->act_sample[7];
    } else {
        rc = ensure_ref_buffers(e);
        if (rc) return rc;
        const size_t tot0 = (size_t)n * e->H0 * e->W0;
        ref_prepare_kernel<<<(unsigned)((tot0 + 255) / 256), 256, 0, st>>>(x_dev, x_dtype == ASR_IN_U8, d.prepare, d.in_h,
                                                                          d.in_w, e->H0, e->W0, (int)n, e->ref_in);
        ASR_LAUNCH_CHECK();
        const float *cur = e->ref_in;
        for (int l = 0; l < 8; ++l) {
            const LayerGeom &g = e->g[l];
            const size_t tot = (size_t)n * g.cout * g.H * g.W;
            float *dst = g.pool ? e->ref_tmp : e->ref_act[l];
            ref_conv3x3_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cur, e->ref_w[l], e->ref_bn[l], (int)n, g.cin,
                                                                             g.cout, g.H, g.W, 1, dst);
            ASR_LAUNCH_CHECK();
            if (g.pool) {
                const size_t totp = (size_t)n * g.cout * g.Ho * g.Wo;
                ref_pool_kernel<<<(unsigned)((totp + 255) / 256), 256, 0, st>>>(dst, (int)n * g.cout, g.H, g.W, e->ref_act[l]);
                ASR_LAUNCH_CHECK();
            }
            cur = e->ref_act[l];
        }
        hp.in = e->ref_act[7]; hp.is_p8 = 0; hp.plane = (long long)e->head_h * e->head_w * 4;
        hp.sample = hp.plane * e->head_c;
    }
    head_kernel<<<(unsigned)n, 128, 0, st>>>(hp);
    ASR_LAUNCH_CHECK();
    if (path == ASR_PATH_TCGEN05) mark(e, st, 2, false);
    return ASR_OK;
}

int asr_encoder_set_timing(asr_encoder_t *e, int enable) {
    ASR_CHECK_ARG(e != nullptr, "NULL handle");
    e->timing = enable != 0;
    e->ev_used = 0;
    return ASR_OK;
}

int asr_encoder_set_fusion(asr_encoder_t *e, int mask) {
    ASR_CHECK_ARG(e != nullptr, "NULL handle");
    e->fuse_mask = ((mask & 1) && e->f01_ok ? 1 : 0) | ((mask & 2) && e->f23_ok ? 2 : 0);
    return ASR_OK;
}

int asr_encoder_get_fusion(const asr_encoder_t *e) { return e ? e->fuse_mask : 0; }

int asr_encoder_get_timing(asr_encoder_t *e, double *ms_layer0, double *ms_conv_tc, double *ms_head, int64_t *n_calls) {
    ASR_CHECK_ARG(e != nullptr, "NULL handle");
    double a = 0, b = 0, c = 0;
    const size_t pairs = e->ev_used / 2;
    size_t calls = 0;
    if (pairs) ASR_CUDA(cudaEventSynchronize(e->ev_pool[e->ev_used - 1]));
    for (size_t i = 0; i < pairs; ++i) {
        float t = 0;
        ASR_CUDA(cudaEventElapsedTime(&t, e->ev_pool[2 * i], e->ev_pool[2 * i + 1]));
        if (e->ev_cat[i] == 0) a += t; else if (e->ev_cat[i] == 1) b += t; else { c += t; ++calls; }
    }
    if (ms_layer0) *ms_layer0 = a;
    if (ms_conv_tc) *ms_conv_tc = b;
    if (ms_head) *ms_head = c;
    if (n_calls) *n_calls = (int64_t)calls;
    e->ev_used = 0;
    return ASR_OK;
}

int asr_encoder_debug_activation(asr_encoder_t *e, int layer, int path, int64_t n, float *out_host, int *c, int *h, int *w) {
    ASR_CHECK_ARG(e && layer >= 0 && layer < 8 && n >= 1 && n <= e->max_batch, "bad argument");
    ASR_CHECK_ARG(!(layer == 0 && path == ASR_PATH_TCGEN05 && e->l01_chunk > 0 && n > e->l01_chunk),
                  "layer-0 activations are only kept for the last sub-chunk");
    ASR_CHECK_ARG(!(layer == 0 && path == ASR_PATH_TCGEN05 && out_host && (e->fuse_mask & 1)),
                  "layer-0 activations never reach memory while layers 0 + 1 run fused (asr_encoder_set_fusion(enc, 0))");
    ASR_CHECK_ARG(!(layer == 2 && path == ASR_PATH_TCGEN05 && out_host && ((e->fuse_mask & 2) || !e->act[2])),
                  "layer-2 activations never reach memory while layers 2 + 3 run fused (asr_encoder_set_fusion(enc, 0))");
    const LayerGeom &g = e->g[layer];
    if (c) *c = g.cout;
    if (h) *h = g.Ho;
    if (w) *w = g.Wo;
    if (!out_host) return ASR_OK;
    ASR_CUDA(cudaDeviceSynchronize());
    if (path == ASR_PATH_FP32) {
        ASR_CHECK_ARG(e->ref_act[layer] != nullptr, "fp32 path has not run");
        ASR_CUDA(cudaMemcpy(out_host, e->ref_act[layer], (size_t)n * g.cout * g.Ho * g.Wo * 4, cudaMemcpyDeviceToHost));
        return ASR_OK;
    }
    std::vector<uint16_t> raw((size_t)e->act_sample[layer] / 2 * n);
    ASR_CUDA(cudaMemcpy(raw.data(), e->act[layer], raw.size() * 2, cudaMemcpyDeviceToHost));
    const int Wp = g.Wo + 2;
    for (int64_t s = 0; s < n; ++s)
        for (int ch = 0; ch < g.cout; ++ch)
            for (int y = 0; y < g.Ho; ++y)
                for (int x = 0; x < g.Wo; ++x) {
                    size_t idx = (size_t)s * (e->act_sample[layer] / 2) + (size_t)(ch >> 3) * (e->act_plane[layer] / 2) +
                                 ((size_t)(y + 1) * Wp + x + 1) * 8 + (ch & 7);
                    uint32_t bits = (uint32_t)raw[idx] << 16;
                    float f;
                    memcpy(&f, &bits, 4);
                    out_host[(((size_t)s * g.cout + ch) * g.Ho + y) * g.Wo + x] = f;
                }
    return ASR_OK;
}

int asr_extract_windows(const void *src_dev, int dtype, int src_h, int src_w, const int32_t *starts_dev, int n, int r0,
                        int win_h, int win_w, void *out_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(src_dev && starts_dev && out_dev, "NULL buffer");
    ASR_CHECK_ARG(dtype == ASR_IN_F32 || dtype == ASR_IN_U8, "bad dtype");
    ASR_CHECK_ARG(r0 >= 0 && win_h >= 1 && win_w >= 1 && r0 + win_h <= src_h && win_w <= src_w, "window does not fit");
    if (n == 0) return ASR_OK;
    dim3 grid((win_h * win_w + 255) / 256, (unsigned)n);
    if (dtype == ASR_IN_U8)
        extract_windows_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint8_t *>(src_dev), src_w, starts_dev, r0, win_h, win_w, reinterpret_cast<uint8_t *>(out_dev));
    else
        extract_windows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float *>(src_dev), src_w, starts_dev, r0, win_h, win_w, reinterpret_cast<float *>(out_dev));
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

// Streams of the host-buffer entry.  Default: every handle owns a copy stream and a kernel stream, so the two branches of
// a pair, called from two host threads, overlap not only one's copies with the other's kernels but also the tails of one's
// persistent kernels with the start of the other's (measured: 771 k pairs/s end to end against 733 k with the shared
// pair below).  ASR_HOST_SHARED_STREAMS=1 makes all handles of a device enqueue into ONE pair of streams instead (kernels
// of different handles then never overlap; no deadlock: each thread enqueues "copy k, then kernels k" in order and both
// streams are FIFO, so every dependency points backwards in enqueue order).
static cudaStream_t g_host_copy[ASR_MAX_DEVICES] = {nullptr}, g_host_comp[ASR_MAX_DEVICES] = {nullptr};
static std::mutex g_host_stream_mutex;
static int host_streams(int dev, cudaStream_t *copy, cudaStream_t *comp) {
    std::lock_guard<std::mutex> lock(g_host_stream_mutex);
    dev = std::max(0, std::min(dev, ASR_MAX_DEVICES - 1));
    if (!g_host_copy[dev]) {
        ASR_CUDA(cudaStreamCreateWithFlags(&g_host_copy[dev], cudaStreamNonBlocking));
        ASR_CUDA(cudaStreamCreateWithFlags(&g_host_comp[dev], cudaStreamNonBlocking));
    }
    *copy = g_host_copy[dev];
    *comp = g_host_comp[dev];
    return ASR_OK;
}

int asr_encoder_embed_host(asr_encoder_t *e, const void *x_host, int x_dtype, int64_t n, float *codes_host,
                           float *latents_host, int path) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(e && x_host, "NULL argument");
    DeviceGuard guard(e->device);
    ASR_CHECK_ARG(x_dtype == ASR_IN_F32 || x_dtype == ASR_IN_U8, "bad x_dtype");
    if (n == 0) return ASR_OK;
    const size_t esz = x_dtype == ASR_IN_U8 ? 1 : 4;
    const size_t sample_bytes = (size_t)e->d.in_h * e->d.in_w * esz;
    const size_t max_bytes = (size_t)e->d.in_h * e->d.in_w * 4 * e->max_batch;
    if (!e->s_copy) {
        static const int own_streams = getenv("ASR_HOST_SHARED_STREAMS") ? !atoi(getenv("ASR_HOST_SHARED_STREAMS")) : 1;
        if (own_streams) {
            ASR_CUDA(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
            ASR_CUDA(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
            e->streams_owned = true;
        } else if ((rc = host_streams(e->device, &e->s_copy, &e->s_comp))) {
            return rc;
        }
        ASR_CUDA(cudaEventCreateWithFlags(&e->ev_final, cudaEventDisableTiming | cudaEventBlockingSync));
        for (int b = 0; b < 2; ++b) {
            ASR_CUDA(cudaEventCreateWithFlags(&e->ev_copied[b], cudaEventDisableTiming));
            ASR_CUDA(cudaEventCreateWithFlags(&e->ev_done[b], cudaEventDisableTiming));
            ASR_CUDA(cudaMalloc(&e->dev_in[b], max_bytes));
        }
        // results leave the device chunk by chunk (two slots of max_batch rows): nothing here depends on n, so no call
        // ever allocates again (a cudaFree / cudaMalloc in a later, larger call used to stall the first big call of a
        // process for 0.05-0.7 s: cudaFree waits for every stream, including the other branch's whole pipeline)
        ASR_CUDA(cudaMalloc(&e->dev_codes, (size_t)2 * e->max_batch * 32 * 4));
        ASR_CUDA(cudaMalloc(&e->dev_lat, (size_t)2 * e->max_batch * 32 * 4));
    }
    cudaPointerAttributes pa;
    bool pinned = cudaPointerGetAttributes(&pa, x_host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (getenv("ASR_DEBUG_HOST"))
        fprintf(stderr, "[asr] embed_host: n %lld, input %s (attr type %d), chunk %d\n", (long long)n,
                pinned ? "pinned" : "PAGEABLE (staged through pinned buffers)", (int)pa.type, e->max_batch);
    if (!pinned && !e->pin_in[0])
        for (int b = 0; b < 2; ++b) ASR_CUDA(cudaMallocHost(&e->pin_in[b], max_bytes));
    auto is_pinned = [](const void *p) {
        cudaPointerAttributes a;
        const bool r = p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost;
        cudaGetLastError();
        return r;
    };
    const bool stage_c = codes_host && !is_pinned(codes_host), stage_l = latents_host && !is_pinned(latents_host);
    if ((stage_c || stage_l) && !e->pin_out[0])
        for (int b = 0; b < 4; ++b) ASR_CUDA(cudaMallocHost(&e->pin_out[b], (size_t)e->max_batch * 128));
    int64_t slot_s0[2] = {0, 0}, slot_nb[2] = {0, 0};
    // Chunk sizes ramp up (512, 1024, ... max_batch): the first host->device copy is the only one that no
    // compute hides, so it is kept short; every later copy overlaps the chunk before it.
    int64_t ci = 0, cur = std::min<int64_t>(e->max_batch, 512), nb = 0;
    for (int64_t s0 = 0; s0 < n; s0 += nb, ++ci, cur = std::min<int64_t>(2 * cur, e->max_batch)) {
        const int b = (int)(ci & 1);
        nb = std::min<int64_t>(cur, n - s0);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(x_host) + (size_t)s0 * sample_bytes;
        if (ci >= 2) ASR_CUDA(cudaStreamWaitEvent(e->s_copy, e->ev_done[b], 0));
        if (!pinned) {
            if (ci >= 2) ASR_CUDA(cudaEventSynchronize(e->ev_copied[b]));
            memcpy(e->pin_in[b], src, (size_t)nb * sample_bytes);
            src = reinterpret_cast<const uint8_t *>(e->pin_in[b]);
        }
        ASR_CUDA(cudaMemcpyAsync(e->dev_in[b], src, (size_t)nb * sample_bytes, cudaMemcpyHostToDevice, e->s_copy));
        ASR_CUDA(cudaEventRecord(e->ev_copied[b], e->s_copy));
        ASR_CUDA(cudaStreamWaitEvent(e->s_comp, e->ev_copied[b], 0));
        float *dc = e->dev_codes + (size_t)b * e->max_batch * 32, *dl = e->dev_lat + (size_t)b * e->max_batch * 32;
        // results of the chunk that used this slot two chunks ago: pageable destinations are filled from the pinned slot
        // once that chunk is done (a device->pageable copy would block the host and with it the whole pipeline)
        if (ci >= 2 && (stage_c || stage_l)) {
            ASR_CUDA(cudaEventSynchronize(e->ev_done[b]));
            if (stage_c) memcpy(codes_host + slot_s0[b] * 32, e->pin_out[b], (size_t)slot_nb[b] * 128);
            if (stage_l) memcpy(latents_host + slot_s0[b] * 32, e->pin_out[2 + b], (size_t)slot_nb[b] * 128);
        }
        rc = asr_encoder_embed(e, e->dev_in[b], x_dtype, nb, dc, dl, path, e->s_comp);
        if (rc) return rc;
        if (codes_host)
            ASR_CUDA(cudaMemcpyAsync(stage_c ? e->pin_out[b] : codes_host + s0 * 32, dc, (size_t)nb * 128, cudaMemcpyDeviceToHost, e->s_comp));
        if (latents_host)
            ASR_CUDA(cudaMemcpyAsync(stage_l ? e->pin_out[2 + b] : latents_host + s0 * 32, dl, (size_t)nb * 128, cudaMemcpyDeviceToHost,
                                     e->s_comp));
        ASR_CUDA(cudaEventRecord(e->ev_done[b], e->s_comp));
        slot_s0[b] = s0;
        slot_nb[b] = nb;
    }
    for (int64_t c = std::max<int64_t>(0, ci - 2); c < ci && (stage_c || stage_l); ++c) {       // the last one or two chunks
        const int b = (int)(c & 1);
        ASR_CUDA(cudaEventSynchronize(e->ev_done[b]));
        if (stage_c) memcpy(codes_host + slot_s0[b] * 32, e->pin_out[b], (size_t)slot_nb[b] * 128);
        if (stage_l) memcpy(latents_host + slot_s0[b] * 32, e->pin_out[2 + b], (size_t)slot_nb[b] * 128);
    }
    // Wait on a blocking-sync event rather than cudaStreamSynchronize: the latter spins, and a spinning thread per
    // branch and rank starves the launch threads of the other ranks on a box with fewer cores than busy threads.
    static const int host_spin = getenv("ASR_HOST_SPIN") ? atoi(getenv("ASR_HOST_SPIN")) : 0;
    if (host_spin) {
        ASR_CUDA(cudaStreamSynchronize(e->s_comp));
    } else {
        ASR_CUDA(cudaEventRecord(e->ev_final, e->s_comp));
        ASR_CUDA(cudaEventSynchronize(e->ev_final));
    }
    return ASR_OK;
}

}  // extern "C"
