// Layer 2 + layer 3 of the sheet branch in ONE persistent kernel (included by encoder.cu, namespace asr).
//
// Replaces, for the full-resolution model (audio_sheet_retrieval/models/mutopia_ccal_cont.py:80-84):
//   conv3x3 12->24 + BN + ELU  ->  conv3x3 24->24 + BN + ELU -> 2x2 max-pool      (80 x 100 pixels)
// Unfused, layer 2 writes 24 channels x 80 x 100 bf16 per sample to HBM and layer 3 reads them back: 3.2 of the
// 11.2 GB a 4096-pair chunk moves, and layer 2 alone runs at the DRAM floor.  Here the layer-2 output only ever exists
// as a ring of 16 image rows in shared memory:
//   TMA          bands of 8 (+2 halo) layer-1 rows, 3-stage ring (one bulk copy per 8-channel plane)
//   MMA A        tcgen05, row-stacked (4 output rows along N, one 128-pixel tile per image row) -> TMEM
//   drain        16 warps, warp = (row, lane quarter) of every tile: TMEM -> bias + ELU -> bf16 -> ring block of 4 rows
//                (P8 layout: the A operand of layer 3 as it is); one TMEM round trip per tile and warp
//   MMA B        two issuers (row groups of one parity each, own accumulator slot), row-stacked; A views are
//                descriptors into the ring; the half-used last K step of the 24 input channels is paired (G_PAIR)
//   epilogue     8 warps: TMEM -> 2x2 max in registers -> bias + ELU -> bf16 -> global (pooled P8 activations)
// A tile is exactly one ring block, so "block written" / "block free" are plain mbarriers (16 drain warps arrive /
// the two MMA-B warps commit).  Lane l of a tile is padded column 1 + l; W + 2 <= 130 so one tile spans a row.
// Every accumulator slot and every ring block is waited on by each of its consumers once per use, in order, so no
// waiter is ever two mbarrier phases behind (parity waits cannot tell those apart).
// Measured on B200 (profiles/r2_conv_ncu_summary.md): 0.90 ms per 4096 samples against 1.26 ms for the two separate
// launches; operand fetch from shared memory at 82 % of peak, tensor pipe 57 % active.
#pragma once

constexpr int G_RB = 4;                       // ring blocks of RS_R rows
constexpr int G_ZROW = G_RB * RS_R;           // index of the all-zero ring row
constexpr int G_TH = 8;                       // layer-A rows per input band
constexpr int G_SLOT = 128;                   // TMEM columns per accumulator slot (4 rows x 32 channels), 2 + 2 slots
constexpr int G_TAIL = 2176;
#ifndef ASR_F23_PAIR
#define ASR_F23_PAIR 1
#endif
#ifndef ASR_F23_LANE0_WAIT
#define ASR_F23_LANE0_WAIT 1
#endif
constexpr bool G_LANE0_WAIT = ASR_F23_LANE0_WAIT != 0;    // drain / epilogue warps wait on their mbarriers with one lane
constexpr int G_PAIR = ASR_F23_PAIR;          // layer B: 5 instead of 6 MMAs per (input row, 3 taps) -- see f23_mma_b
constexpr int G_DRAIN_WARP0 = 4, G_DRAIN_WARPS = 16, G_EPI_WARP0 = G_DRAIN_WARP0 + G_DRAIN_WARPS, G_EPI_WARPS = 8;
constexpr int G_THREADS = 32 * (G_EPI_WARP0 + G_EPI_WARPS);

struct F23Params {
    const bf16 *in;             // output of the layer before A (P8, padded, zero borders)
    bf16 *out;                  // pooled output of layer B
    const uint8_t *wblobA;      // rows-kernel blobs: [dx][K chunk][4 blocks][NP][8] bf16, then NP fp32 biases
    const uint8_t *wblobB;
    int n, H, W, Wp, NG, bands;
    int KCLA;                   // input chunks of layer A that hold real channels
    int NPA, NPB, coutA, coutB;
    int Ho, Wo, Wpo;
    long long in_plane, in_sample, out_plane, out_sample;
    int sps, stage_bytes, ring_plane, ring_planes, wbytesA, wbytesB;
    int off_wB, off_stage, off_ring, off_bar;
    int n_stages;               // input ring depth (2 or 3)
    int dbg;                    // diagnostics (env ASR_F23_DEBUG, bits): 1 no epilogue work, 2 no layer-B MMAs, 4 no drain work, 8 no layer-A MMAs
};

// MMA-B issuer of parity PW: row groups rg = PW, PW + 2, ... of every sample, accumulator slot PW.
template <int PW, int KPB, int NPB>
__device__ __forceinline__ void f23_mma_b(const F23Params &p, uint8_t *smem, int n_it, uint64_t *mid_full, uint64_t *mid_free,
                                          uint64_t *accB_full, uint64_t *accB_empty) {
    const bool no_mma = (p.dbg & 2) != 0;
    constexpr uint32_t np = (uint32_t)NPB;
    const uint32_t a_lbo = ((uint32_t)p.ring_plane >> 4) << 16;
    const uint32_t ring_lo = ((smem_u32(smem + p.off_ring) & 0x3FFFFu) >> 4) | a_lbo;
    const uint32_t w_lo = ((smem_u32(smem + p.off_wB) & 0x3FFFFu) >> 4) | ((RS_R * np) << 16);     // LBO = 4 * NP * 16 B
    const uint32_t kstep_a = (uint32_t)(2 * p.ring_plane) >> 4;
    const uint32_t wp = (uint32_t)p.Wp;
    const uint32_t d_tmem = 2u * G_SLOT + (uint32_t)PW * G_SLOT;
    // K pairing of the last, half-used chunk (see below): needs the chunk count to be 2 KPB with the last one padding
    constexpr bool PAIR = G_PAIR != 0;
    constexpr uint32_t KCL = 2 * KPB - 2;                               // the last chunk with real channels
    constexpr uint32_t dx_stride = 2 * KPB * RS_R * np;                 // 16-byte units between the dx blocks of the blob
    const uint32_t c2_lo = ((smem_u32(smem + p.off_ring + KCL * p.ring_plane) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w16 = (smem_u32(smem + p.off_wB) & 0x3FFFFu) >> 4;
    const uint32_t wpair_lo = (w16 + KCL * RS_R * np) | (dx_stride << 16);
    const uint32_t wlast_lo = (w16 + 2u * dx_stride + KCL * RS_R * np) | ((RS_R * np) << 16);
    uint32_t u = 0;                                   // this warp's running tile count
    for (int it = 0; it < n_it; ++it) {
        const int B0 = it * p.NG;
        for (int rg = PW; rg < p.NG; rg += 2, ++u) {
            // ring blocks this group reads for the first time (this warp): rg and rg + 1 (the very first group of
            // parity 1 also block 0; block rg - 1 was waited for as "rg + 1" of the group before)
            for (int jb = (rg == 1 ? 0 : rg); jb <= min(rg + 1, p.NG - 1); ++jb)
                mbar_wait_tag(&mid_full[(B0 + jb) % G_RB], (uint32_t)(((B0 + jb) / G_RB) & 1), 4000000 + (int)u);
            fence_proxy_async();
            uint32_t vrow[RS_R + 2];                  // ring position (16-byte units) of input row 4 rg - 1 + v, column 0
#pragma unroll
            for (int v = 0; v < RS_R + 2; ++v) {
                const int m = RS_R * rg - 1 + v;
                const int rr = (m < 0 || m >= p.H) ? G_ZROW : ((B0 + (m >> 2)) % G_RB) * RS_R + (m & 3);
                vrow[v] = (uint32_t)rr * wp;
            }
            mbar_wait_tag(&accB_empty[PW], ((u & 1u) ^ 1u), 5000000 + (int)u);
            tc_fence_after();
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
                        uint32_t a_lo = ring_lo + vrow[v] + (uint32_t)dx;
                        uint32_t b_lo = w_lo + (uint32_t)(dx * 2 * KPB) * (RS_R * np) + (uint32_t)zs * np;
#pragma unroll
                        for (int kp = 0; kp < (PAIR ? KPB - 1 : KPB); ++kp) {
                            const bool first = (i == 0 && dx == 0 && kp == 0);
                            tc_mma_bf16_elect(d_tmem + (uint32_t)db * np, a_lo, b_lo, UMMA_DESC_HI,
                                              umma_idesc_bf16(NPB * (first ? RS_R : nb)), first ? 0u : 1u);
                            a_lo += kstep_a;
                            b_lo += 2u * RS_R * np;
                        }
                    }
                    if (PAIR) {
                        // the half-used last K step (channels 16..23 + 8 of padding): pair the chunk with ITSELF one
                        // position further (A: LBO = 16 B) and the weights of tap dx with those of dx + 1 (B: LBO =
                        // the dx stride of the blob): K = [chunk @ dx, chunk @ dx + 1].  The third tap pairs with the
                        // all-zero padding chunk of the blob (whatever finite values the A side reads there count 0).
                        const uint32_t a2 = c2_lo + vrow[v];
                        tc_mma_bf16_elect(d_tmem + (uint32_t)db * np, a2, wpair_lo + (uint32_t)zs * np, UMMA_DESC_HI,
                                          umma_idesc_bf16(NPB * nb), 1u);
                        tc_mma_bf16_elect(d_tmem + (uint32_t)db * np, a2 + 2u, wlast_lo + (uint32_t)zs * np, UMMA_DESC_HI,
                                          umma_idesc_bf16(NPB * nb), 1u);
                    }
                }
            }
            tc_commit_elect(&accB_full[PW]);
            // free the ring blocks that the next group of this parity (rg + 2: blocks rg + 1 ..) no longer reads;
            // every block is read by groups of both parities, so each block use gets exactly two arrivals
            if (rg >= 1) tc_commit_elect(&mid_free[(B0 + rg - 1) % G_RB]);
            tc_commit_elect(&mid_free[(B0 + rg) % G_RB]);
            if (rg + 2 >= p.NG && rg + 1 < p.NG) tc_commit_elect(&mid_free[(B0 + rg + 1) % G_RB]);
        }
    }
}

template <int KPA, int KPB, int NPA, int NPB>
__global__ void __launch_bounds__(G_THREADS, 1) l23_fused_kernel(const F23Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (see conv3x3_tc_kernel)
    uint8_t *wA_sm = smem;
    float *biasA_sm = reinterpret_cast<float *>(smem + p.wbytesA);
    uint8_t *wB_sm = smem + p.off_wB;
    float *biasB_sm = reinterpret_cast<float *>(wB_sm + p.wbytesB);
    uint8_t *stage_sm = smem + p.off_stage;
    uint8_t *ring = smem + p.off_ring;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *w_full = bars;              // 1
    uint64_t *in_full = bars + 1;         // [3]
    uint64_t *in_empty = bars + 4;        // [3]
    uint64_t *accA_full = bars + 7;       // [2]
    uint64_t *accA_empty = bars + 9;      // [2]
    uint64_t *mid_full = bars + 11;       // [G_RB]
    uint64_t *mid_free = bars + 15;       // [G_RB]
    uint64_t *accB_full = bars + 19;      // [2]
    uint64_t *accB_empty = bars + 21;     // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 23);

    for (int i = tid; i < (p.ring_planes * p.ring_plane + G_TAIL) / 16; i += G_THREADS)     // borders, zero row, padding planes, tail
        reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < 3; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&accA_full[s], 1); mbar_init(&accA_empty[s], G_DRAIN_WARPS);
            mbar_init(&accB_full[s], 1); mbar_init(&accB_empty[s], G_EPI_WARPS);
        }
        for (int s = 0; s < G_RB; ++s) { mbar_init(&mid_full[s], G_DRAIN_WARPS); mbar_init(&mid_free[s], 2); }
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (tmem_base != 0u) __trap();        // one CTA per SM owns all 512 columns: slot addresses below are absolute
    const int n_it = (int)blockIdx.x < p.n ? (p.n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int tiles_per_band = G_TH / RS_R;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)(p.wbytesA + NPA * 4 + p.wbytesB + NPB * 4));
            tma_bulk_g2s(wA_sm, p.wblobA, (uint32_t)(p.wbytesA + NPA * 4), w_full);
            tma_bulk_g2s(wB_sm, p.wblobB, (uint32_t)(p.wbytesB + NPB * 4), w_full);
            const uint32_t bytes = (uint32_t)((G_TH + 2) * p.Wp * 16);
            int ii = 0;
            for (int it = 0; it < n_it; ++it) {
                const long long n = (long long)blockIdx.x + (long long)it * gridDim.x;
                for (int b = 0; b < p.bands; ++b, ++ii) {
                    const int s = ii % p.n_stages;
                    mbar_wait_tag(&in_empty[s], (uint32_t)(((ii / p.n_stages) & 1) ^ 1), 1000000 + ii);
                    mbar_expect_tx(&in_full[s], bytes * (uint32_t)p.KCLA);
                    const uint8_t *src = reinterpret_cast<const uint8_t *>(p.in) + n * p.in_sample + (long long)(b * G_TH) * p.Wp * 16;
                    uint8_t *dst = stage_sm + (size_t)s * p.stage_bytes + 16;
                    for (int kc = 0; kc < p.KCLA; ++kc)
                        tma_bulk_g2s(dst + (size_t)kc * p.sps, src + (long long)kc * p.in_plane, bytes, &in_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA A (row-stacked, as conv3x3_rows_kernel with one tile per image row) =================
        const bool no_mma = (p.dbg & 8) != 0;
        constexpr uint32_t np = (uint32_t)NPA;
        const uint32_t a_lbo = ((uint32_t)p.sps >> 4) << 16;
        const uint32_t w_lo = ((smem_u32(wA_sm) & 0x3FFFFu) >> 4) | ((RS_R * np) << 16);
        const uint32_t kstep_a = (uint32_t)(2 * p.sps) >> 4;
        const uint32_t wp = (uint32_t)p.Wp;
        mbar_wait_tag(w_full, 0, 2000000);
        const int n_items = n_it * p.bands;
        uint32_t k = 0;                               // running tile count = it * NG + rg
        for (int ii = 0; ii < n_items; ++ii) {
            const int s = ii % p.n_stages;
            mbar_wait_tag(&in_full[s], (uint32_t)((ii / p.n_stages) & 1), 2100000 + ii);
            tc_fence_after();
            const uint32_t band_lo = ((smem_u32(stage_sm + (size_t)s * p.stage_bytes + 16) & 0x3FFFFu) >> 4) | a_lbo;
            for (int rgl = 0; rgl < tiles_per_band; ++rgl, ++k) {
                const uint32_t slot = k & 1u;
                mbar_wait_tag(&accA_empty[slot], (((k >> 1) & 1u) ^ 1u), 2200000 + (int)k);
                tc_fence_after();
                const uint32_t d_tmem = slot * G_SLOT;
                const uint32_t tile_lo = band_lo + (uint32_t)(RS_R * rgl) * wp;
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
                            uint32_t b_lo = w_lo + (uint32_t)(dx * 2 * KPA) * (RS_R * np) + (uint32_t)zs * np;
#pragma unroll
                            for (int kp = 0; kp < KPA; ++kp) {
                                const bool first = (i == 0 && dx == 0 && kp == 0);
                                tc_mma_bf16_elect(d_tmem + (uint32_t)db * np, a_lo, b_lo, UMMA_DESC_HI,
                                                  umma_idesc_bf16(NPA * (first ? RS_R : nb)), first ? 0u : 1u);
                                a_lo += kstep_a;
                                b_lo += 2u * RS_R * np;
                            }
                        }
                    }
                }
                tc_commit_elect(&accA_full[slot]);
            }
            tc_commit_elect(&in_empty[s]);
        }
    } else if (warp == 2) {
        mbar_wait_tag(w_full, 0, 3000000);
        f23_mma_b<0, KPB, NPB>(p, smem, n_it, mid_full, mid_free, accB_full, accB_empty);
    } else if (warp == 3) {
        mbar_wait_tag(w_full, 0, 3000001);
        f23_mma_b<1, KPB, NPB>(p, smem, n_it, mid_full, mid_free, accB_full, accB_empty);
    } else if (warp < G_EPI_WARP0) {
        // ================= drain: TMEM -> bias + ELU -> bf16 -> ring block (all 16 warps on every tile: group g = row g) =================
        // One TMEM round trip per tile and warp (both 16-column loads of the row in flight): the drain is a latency
        // chain per warp, so rows are spread over four groups instead of looping inside a warp.
        static_assert(NPA == 32, "the drain reads one 32-column row per warp");
        const int q = warp & 3, jr = (warp - G_DRAIN_WARP0) >> 2;
        const int c = 1 + 32 * q + lane;              // padded column of this lane
        const bool col_real = c <= p.W;
        const int nchr = (p.coutA + 7) >> 3;
        const int n_groups16 = (p.coutA + 3) >> 2;
        mbar_wait_tag(w_full, 0, 6000000);            // bias visible
        const int total = n_it * p.NG;
        for (int k = 0; k < total; ++k) {
            const uint32_t slot = (uint32_t)k & 1u;
            const int blk = k % G_RB;
            if (G_LANE0_WAIT) {       // one polling lane per warp: 32 lanes on one mbarrier count as 32 shared-memory wavefronts
                if (lane == 0) {
                    mbar_wait_tag(&mid_free[blk], (uint32_t)(((k / G_RB) & 1) ^ 1), 6100000 + k);
                    mbar_wait_tag(&accA_full[slot], (uint32_t)((k >> 1) & 1), 6200000 + k);
                }
                __syncwarp();
            } else {
                mbar_wait_tag(&mid_free[blk], (uint32_t)(((k / G_RB) & 1) ^ 1), 6100000 + k);
                mbar_wait_tag(&accA_full[slot], (uint32_t)((k >> 1) & 1), 6200000 + k);
            }
            tc_fence_after();
            const uint32_t taddr = slot * G_SLOT + ((uint32_t)(q * 32) << 16) + (uint32_t)(jr * NPA);
            if (!(p.dbg & 4)) {
                uint8_t *dst = ring + ((size_t)(blk * RS_R + jr) * p.Wp + c) * 16;
                uint32_t r[32];
                tmem_ld16_issue(taddr, r);
                tmem_ld16_issue(taddr + 16u, r + 16);
                tmem_ld_wait();
#pragma unroll
                for (int ng = 0; ng < 2; ++ng) {
                    uint32_t pk[8];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        if (ng * 4 + q4 < n_groups16) {       // warp-uniform: padded channels stay exactly zero
                            const float4 b4 = *reinterpret_cast<const float4 *>(&biasA_sm[ng * 16 + 4 * q4]);
                            const float *v = reinterpret_cast<const float *>(r) + ng * 16 + 4 * q4;
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_f(v[0] + b4.x), elu_f(v[1] + b4.y));
                            __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_f(v[2] + b4.z), elu_f(v[3] + b4.w));
                            pk[2 * q4] = *reinterpret_cast<uint32_t *>(&h0);
                            pk[2 * q4 + 1] = *reinterpret_cast<uint32_t *>(&h1);
                        } else {
                            pk[2 * q4] = 0u;
                            pk[2 * q4 + 1] = 0u;
                        }
                    }
                    if (col_real) {
                        if (2 * ng < nchr)
                            *reinterpret_cast<uint4 *>(dst + (size_t)(2 * ng) * p.ring_plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (2 * ng + 1 < nchr)
                            *reinterpret_cast<uint4 *>(dst + (size_t)(2 * ng + 1) * p.ring_plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&accA_empty[slot]);
                mbar_arrive(&mid_full[blk]);
            }
        }
    } else {
        // ================= epilogue: TMEM -> 2x2 max -> bias + ELU -> bf16 -> global (all 8 warps on every tile) =================
        // group g takes pooled row g; rows (2 pr, 2 pr + 1) pool vertically inside the thread, lanes (2k, 2k+1) are
        // one pooled column: the even lane finishes channels 0-3 of a chunk, the odd lane 4-7
        const int q = warp & 3, grp = (warp - G_EPI_WARP0) >> 2;
        const int odd = lane & 1;
        const int c = 1 + 32 * q + lane;
        const bool valid = c <= p.W;
        const int nchr = (p.coutB + 7) >> 3;
        const int n_groups16 = (p.coutB + 3) >> 2;
        const int pr = grp;
        mbar_wait_tag(w_full, 0, 7000000);
        const int total = n_it * p.NG;
        for (int u = 0; u < total; ++u) {
            const int it = u / p.NG, rg = u - it * p.NG;
            const uint32_t slot = (uint32_t)rg & 1u;                 // NG is even: parity of u = parity of rg
            const long long n = (long long)blockIdx.x + (long long)it * gridDim.x;
            if (G_LANE0_WAIT) {
                if (lane == 0) mbar_wait_tag(&accB_full[slot], (uint32_t)((u >> 1) & 1), 7100000 + u);
                __syncwarp();
            } else {
                mbar_wait_tag(&accB_full[slot], (uint32_t)((u >> 1) & 1), 7100000 + u);
            }
            tc_fence_after();
            uint8_t *out_n = reinterpret_cast<uint8_t *>(p.out) + n * p.out_sample;
            const uint32_t taddr = 2u * G_SLOT + slot * G_SLOT + ((uint32_t)(q * 32) << 16);
            const int yo = 2 * rg + pr;
            const long long opos = ((long long)(yo + 1) * p.Wpo + ((c - 1) >> 1) + 1) * 16 + 8 * odd;
            for (int h = 0; h < ((p.dbg & 1) ? 0 : nchr); ++h) {
                float v[16];
                tmem_ld8x2(taddr + (uint32_t)(2 * pr * NPB + h * 8), (uint32_t)NPB, v);
                float m[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float lo = fmaxf(v[k], v[8 + k]), hi = fmaxf(v[4 + k], v[12 + k]);
                    const float other = __shfl_xor_sync(0xffffffffu, odd ? lo : hi, 1);
                    m[k] = fmaxf(odd ? hi : lo, other);
                }
                const float4 b4 = *reinterpret_cast<const float4 *>(&biasB_sm[h * 8 + 4 * odd]);
                __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_f(m[0] + b4.x), elu_f(m[1] + b4.y));
                __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_f(m[2] + b4.z), elu_f(m[3] + b4.w));
                const bool real = h * 2 + odd < n_groups16;       // padded channels stay exactly zero
                const uint2 o2 = make_uint2(real ? *reinterpret_cast<uint32_t *>(&h0) : 0u,
                                            real ? *reinterpret_cast<uint32_t *>(&h1) : 0u);
                if (valid && yo < p.Ho && ((c - 1) >> 1) < p.Wo)
                    *reinterpret_cast<uint2 *>(out_n + (long long)h * p.out_plane + opos) = o2;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accB_empty[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}
