// Layer 0 + layer 1 of the sheet branch in ONE persistent kernel (included by encoder.cu, namespace asr).
//
// Replaces, for the full-resolution model (audio_sheet_retrieval/models/mutopia_ccal_cont.py:75-79):
//   prepare (x / 255) -> conv3x3 1->12 + BN + ELU -> conv3x3 12->12 + BN + ELU -> 2x2 max-pool
// The unfused path writes the 160x200x16 bf16 output of layer 0 to HBM (1 MB per sample) and reads it
// back in layer 1; here it never leaves the SM:
//   converters   global pixels -> K-major A tiles of the banded-Toeplitz GEMM (as l0_tc_kernel)
//   MMA 0        tcgen05: 8 output rows x 12 channels (N = 96) per 128 raster positions -> TMEM
//   drain        TMEM -> ELU -> bf16 -> a RING of layer-0 output rows in shared memory, P8 layout
//                (3 blocks of 8 rows + one all-zero row that stands for the rows outside the image)
//   MMA 1        tcgen05, row-stacked as conv3x3_rows_kernel: its A views are descriptors into the ring
//   epilogue     TMEM -> 2x2 max in registers -> bias + ELU -> bf16 -> global (pooled P8 activations)
// The layer-0 GEMM rasterises over (row block, padded column) of a sample, so every A row is a real
// position (no 128-column tile padding); a tile may straddle two row blocks.
// Flow control: ring block "written" = a shared-memory counter of finished (position, row) units
// (release-add by the drain warps, acquire-poll by the MMA-1 warps; a tile straddles blocks, so
// arrival counts per block are not fixed and an mbarrier does not fit); ring block "free" = an mbarrier
// that the two MMA-1 warps commit to after the last row group of their parity that reads the block.
#pragma once

#ifndef ASR_F01_R0
#define ASR_F01_R0 8
#endif
constexpr int F_R0 = ASR_F01_R0;              // layer-0 output rows per tile = rows per ring block (4 or 8)
constexpr int F_R0_SHIFT = F_R0 == 8 ? 3 : 2;
constexpr int F_RB = 24 / F_R0;               // ring blocks (24 rows: finer blocks = more of the ring usable as slack)
constexpr int F_ZROW = F_R0 * F_RB;           // index of the all-zero ring row
constexpr int F_CW = 4;                       // converter warps = A buffers
constexpr int F_AROWS = 136;                  // 130 raster positions of a tile + slack
constexpr int F_C0 = 12;                      // layer-0 channels
constexpr int F_NPAD0 = F_R0 * F_C0;          // 96
constexpr int F_SLOT0 = F_R0 == 8 ? 128 : 64, F_SLOT1 = 64;    // TMEM columns per accumulator slot (layer 0: 2 slots, layer 1: 4)
constexpr int F_NP1 = 16;                     // padded layer-1 channels
constexpr int F_W1BYTES = 3 * 2 * RS_R * F_NP1 * 16;
constexpr int F_TAIL = 2176;
constexpr int F_MMA0_WARP = F_CW, F_MMA1_WARP0 = F_CW + 1, F_DRAIN_WARP0 = 8;

struct F01Params {
    const void *x;              // (n, 1, H, W) u8 or f32
    int x_u8, prepare, int_pixels;
    int H, W, Wp, n;
    const uint8_t *blob0;       // [dx][part hi, lo][K chunk][96][8] bf16
    const uint8_t *wblob1;      // rows-kernel blob of layer 1: [dx][K chunk][4 blocks][16][8] bf16, then 16 fp32 biases
    bf16 *out;
    long long out_plane, out_sample;
    int Ho, Wo, Wpo, cout1;
    int NB, NG, T0, JT;         // ring blocks / layer-1 row groups / layer-0 tiles per sample, 128-pixel tiles per row
    unsigned wp_magic;          // ceil(2^32 / Wp), exact for every g < T0 * 128 + 128
    int ring_plane;             // (F_ZROW + 1) * Wp * 16 bytes per 8-channel chunk
    int abuf;                   // bytes per A buffer
    int off_w1, off_a, off_ring, off_lut, off_bar;
    int dbg;                    // diagnostics (env ASR_F01_DEBUG, bits): 1 no epilogue work, 2 no layer-1 MMAs, 4 no drain work,
                                // 8 no layer-0 MMAs, 16 no converter work (results are garbage; for timing the roles apart)
};

// One converter warp builds whole A tiles: 260 items (K chunk c, tile row i) = 9 per lane in batches of three.
// A row i <-> raster position g = 128 t - 1 + i = rbl * Wp + xp; K element kk <-> image row 8 rbl - 1 + kk
// (kk = 0 .. F_R0 + 1 are used by the output rows of the block, kk = 15 carries the constant 1 of the bias row).
template <bool INT_PIXELS>
__device__ __forceinline__ void f01_convert_tile(const F01Params &p, uint8_t *a_buf, const float *lut, long long sample_off, int t) {
    const int lane = threadIdx.x & 31;
    const uint8_t *xu = reinterpret_cast<const uint8_t *>(p.x) + sample_off;
    const float *xf = reinterpret_cast<const float *>(p.x) + sample_off;
    const int g0 = 128 * t - 1;
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
            const unsigned rbl = __umulhi(gc, p.wp_magic);
            const int x = (int)(gc - rbl * (unsigned)p.Wp) - 1;
            ok[u] = item < 260 && g >= 0 && (int)rbl < p.NB && x >= 0 && x < p.W;
            const int yb = (int)rbl * F_R0 - 1 + 8 * c;
            const unsigned base = (unsigned)min(max(x, 0), p.W - 1);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int y = yb + kk;
                const bool in = ok[u] && y >= 0 && y < p.H && 8 * c + kk < F_R0 + 2;     // K rows the block's taps reach
                const unsigned o = base + (unsigned)(min(max(y, 0), p.H - 1) * p.W);
                raw[u][kk] = 0u;
                if (in) raw[u][kk] = p.x_u8 ? (uint32_t)xu[o] : __float_as_uint(xf[o]);
            }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int item = lane + 32 * (3 * batch + u);
            const int c = item >= 130 ? 1 : 0, i = item - 130 * c;
            float v[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                if (INT_PIXELS) {
                    v[kk] = (float)raw[u][kk];            // exact in bf16; 1/255 lives in the weight matrix
                } else {
                    float f = p.x_u8 ? (float)raw[u][kk] : __uint_as_float(raw[u][kk]);
                    if (p.prepare == ASR_PREP_SCALE) f = p.x_u8 ? lut[(int)raw[u][kk]] : f / 255.0f;
                    v[kk] = f;                             // raw = 0 where the position is outside the image
                }
            }
            if (c == 1) v[7] = 1.0f;
            if (item < 260) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * kk], v[2 * kk + 1]);
                    hi[kk] = *reinterpret_cast<const uint32_t *>(&h);
                    if (!INT_PIXELS) {
                        const float2 hf = __bfloat1622float2(h);
                        const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * kk] - hf.x, v[2 * kk + 1] - hf.y);
                        lo[kk] = *reinterpret_cast<const uint32_t *>(&l);
                    }
                }
                *reinterpret_cast<uint4 *>(a_buf + (c * F_AROWS + i) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                if (!INT_PIXELS)
                    *reinterpret_cast<uint4 *>(a_buf + ((2 + c) * F_AROWS + i) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

__device__ __forceinline__ void f01_wait_count(const uint32_t *cnt, uint32_t target) {
    if (ld_acquire_shared(cnt) >= target) return;
    const uint64_t t0 = globaltimer_ns();
    while (ld_acquire_shared(cnt) < target) {
        if (globaltimer_ns() - t0 > 2000000000ull) {
            if ((threadIdx.x & 31) == 0)
                printf("asr: ring counter timeout block %d warp %d target %u have %u\n", blockIdx.x, threadIdx.x >> 5, target,
                       ld_acquire_shared(cnt));
            __trap();
        }
    }
}

// MMA-1 issuer of parity W: row groups rg = W, W + 2, ..., both 128-pixel tiles of a group.  Everything that feeds a
// tcgen05 operand is computed from kernel parameters, blockIdx and loop counters only (W is a template constant, the
// TMEM base of a 512-column allocation is column 0 -- checked), so the compiler keeps descriptors in UNIFORM registers
// and an MMA costs a few uniform instructions instead of an elect / broadcast sequence per operand.
template <int W>
__device__ __forceinline__ void f01_mma1(const F01Params &p, uint8_t *smem, int n_it, uint32_t block_units, uint32_t *mid_cnt,
                                         uint64_t *acc1_full, uint64_t *acc1_empty, uint64_t *mid_free, uint32_t tmem_base) {
    if (tmem_base != 0u) __trap();                         // one CTA per SM owns all 512 columns: the allocation starts at 0
    const bool no_mma = (p.dbg & 2) != 0;
    const uint32_t a_lbo = ((uint32_t)p.ring_plane >> 4) << 16;
    const uint32_t w_lo = ((smem_u32(smem + p.off_w1) & 0x3FFFFu) >> 4) | ((uint32_t)(RS_R * F_NP1) << 16);   // LBO = 4 * 16 * 16 B
    const uint32_t ring_lo = ((smem_u32(smem + p.off_ring) & 0x3FFFFu) >> 4) | a_lbo;
    const uint32_t wp = (uint32_t)p.Wp;
    uint32_t u = 0;                                     // this warp's running tile count
    for (int it = 0; it < n_it; ++it) {
        const int B0 = it * p.NB;
        for (int rg = W; rg < p.NG; rg += 2) {
            const int j_lo = rg ? (4 * rg - 1) >> F_R0_SHIFT : 0;
            const int j_hi = min(p.NB - 1, (4 * rg + 4) >> F_R0_SHIFT);
            for (int jb = j_lo; jb <= j_hi; ++jb)
                f01_wait_count(&mid_cnt[(B0 + jb) % F_RB], (uint32_t)((B0 + jb) / F_RB + 1) * block_units);
            fence_proxy_async();
            uint32_t vrow[RS_R + 2];                    // ring position (16-byte units) of input row 4 rg - 1 + v, column 0
#pragma unroll
            for (int v = 0; v < RS_R + 2; ++v) {
                const int m = 4 * rg - 1 + v;
                const int rr = (m < 0 || m >= p.H) ? F_ZROW : ((B0 + (m >> F_R0_SHIFT)) % F_RB) * F_R0 + (m & (F_R0 - 1));
                vrow[v] = (uint32_t)rr * wp;
            }
            for (int j = 0; j < p.JT; ++j, ++u) {
                const uint32_t slot = 2u * (uint32_t)W + (u & 1u), sph = (u >> 1) & 1u;
                mbar_wait_tag(&acc1_empty[slot], sph ^ 1u, 4000000 + (int)u);
                tc_fence_after();
                const uint32_t d_tmem = 2 * F_SLOT0 + slot * F_SLOT1;
                const uint32_t tile_lo = ring_lo + 128u * (uint32_t)j;
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
                        const bool first = (i == 0 && dx == 0);
                        const uint32_t a_lo = tile_lo + vrow[v] + (uint32_t)dx;
                        const uint32_t b_lo = w_lo + (uint32_t)(dx * 2) * (RS_R * F_NP1) + (uint32_t)zs * F_NP1;
                        tc_mma_bf16_elect(d_tmem + (uint32_t)db * F_NP1, a_lo, b_lo, UMMA_DESC_HI,
                                          umma_idesc_bf16(F_NP1 * (first ? RS_R : nb)), first ? 0u : 1u);
                    }
                }
                }
                tc_commit_elect(&acc1_full[slot]);
            }
            // free the ring blocks for which this was the last row group of this warp's parity: those that the next
            // group of the same parity (rg + 2, rows 4 rg + 7 ...) no longer reads.  Every block is read by groups of
            // both parities, so each block use gets exactly two arrivals.
            const int next_lo = (4 * rg + 7) >> F_R0_SHIFT;
            for (int jb = j_lo; jb <= j_hi; ++jb)
                if (rg + 2 >= p.NG || jb < next_lo) tc_commit_elect(&mid_free[(B0 + jb) % F_RB]);
        }
    }
}


// DG drain groups (layer 0) and four epilogue groups (layer 1), four warps each.  Every layer-1 accumulator slot has
// exactly one producer and one consumer -- slot 2 w + (u & 1) belongs to MMA-1 warp w (u = its running tile count)
// and is drained by epilogue group 2 w + (u & 1) -- so nobody can run two mbarrier phases ahead of a slot (with
// shared slots a wait on "parity p" is satisfied by the phase before the previous one).
// EG = 2: one epilogue group per MMA-1 warp, which then drains both of that warp's slots alternately (still one
// consumer per slot, tiles in order); the freed warps go to the drain (DG = 4: two rows per warp, both TMEM loads of a
// tile in one round trip -- the drain is a latency chain per warp).
constexpr int F_EG = 4;
template <int DG, int EG = F_EG>
__global__ void __launch_bounds__(32 * (F_DRAIN_WARP0 + 4 * (DG + EG)), 1) l01_fused_kernel(const F01Params p) {
    static_assert(EG == 4 || EG == 2, "two MMA-1 warps: four or two epilogue groups");
    constexpr int NTHREADS = 32 * (F_DRAIN_WARP0 + 4 * (DG + EG));
    constexpr int EPI_WARP0_F = F_DRAIN_WARP0 + 4 * DG;
    extern __shared__ __align__(128) uint8_t smem[];
    // The warp index goes through a lane-0 broadcast so that the compiler KNOWS it is warp-uniform: the role branches
    // below are then uniform control flow, and the MMA issuers' descriptor arithmetic stays on the uniform datapath
    // (with `tid >> 5` every role body counts as divergent code and each tcgen05.mma gets an elect / R2UR sequence).
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    uint8_t *b0_sm = smem;
    uint8_t *w1_sm = smem + p.off_w1;
    float *bias1_sm = reinterpret_cast<float *>(w1_sm + F_W1BYTES);
    uint8_t *a_sm = smem + p.off_a;
    uint8_t *ring = smem + p.off_ring;
    float *lut_sm = reinterpret_cast<float *>(smem + p.off_lut);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *a_ready = bars;             // [F_CW]  converters -> MMA 0
    uint64_t *a_free = bars + 4;          // [F_CW]  MMA 0 done reading the A buffer
    uint64_t *acc0_full = bars + 8;       // [2]
    uint64_t *acc0_empty = bars + 10;     // [2]
    uint64_t *mid_free = bars + 12;       // [F_RB <= 6]  MMA 1 done reading the ring block
    uint64_t *acc1_full = bars + 18;      // [4]
    uint64_t *acc1_empty = bars + 22;     // [4]
    uint32_t *mid_cnt = reinterpret_cast<uint32_t *>(bars + 26);   // [F_RB] finished (position, row) units, monotonic
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 29);

    constexpr int B0BYTES = 3 * 2 * 2 * F_NPAD0 * 16;
    for (int i = tid; i < B0BYTES / 16; i += NTHREADS)
        reinterpret_cast<uint4 *>(b0_sm)[i] = reinterpret_cast<const uint4 *>(p.blob0)[i];
    for (int i = tid; i < (F_W1BYTES + F_NP1 * 4) / 16; i += NTHREADS)
        reinterpret_cast<uint4 *>(w1_sm)[i] = reinterpret_cast<const uint4 *>(p.wblob1)[i];
    for (int i = tid; i < (2 * p.ring_plane + F_TAIL) / 16; i += NTHREADS)      // zero row, tail slack (and a defined start)
        reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 256) lut_sm[tid] = (float)tid / 255.0f;
    if (tid == 0) {
        for (int s = 0; s < F_CW; ++s) { mbar_init(&a_ready[s], 1); mbar_init(&a_free[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc0_full[s], 1); mbar_init(&acc0_empty[s], 4 * DG); }
        for (int s = 0; s < F_RB; ++s) { mbar_init(&mid_free[s], 2); mid_cnt[s] = 0u; }
        for (int s = 0; s < 4; ++s) { mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], 4); }
        mbar_fence_init();
    }
    if (warp == F_MMA0_WARP) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_it = (int)blockIdx.x < p.n ? (p.n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total0 = n_it * p.T0;                       // layer-0 tiles of this CTA
    const uint32_t block_units = (uint32_t)(p.Wp * F_R0);

    if (warp < F_CW) {
        // ================= converters =================
        uint8_t *a_buf = a_sm + warp * p.abuf;
        for (int k = warp; k < total0; k += F_CW) {
            const int it = k / p.T0, t = k - it * p.T0;
            const long long n = (long long)blockIdx.x + (long long)it * gridDim.x;
            mbar_wait_tag(&a_free[warp], (uint32_t)((((k / F_CW) & 1)) ^ 1), 1000000 + k);
            if (p.dbg & 16) {
            } else if (p.int_pixels) f01_convert_tile<true>(p, a_buf, lut_sm, n * p.H * p.W, t);
            else f01_convert_tile<false>(p, a_buf, lut_sm, n * p.H * p.W, t);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[warp]);
        }
    } else if (warp == F_MMA0_WARP) {
        // ================= MMA 0: per dx (A_hi, B_hi) [, (A_lo, B_hi)], (A_hi, B_lo); whole warp, one elected lane issues =================
        if (tmem_base != 0u) __trap();
        const uint32_t idesc = umma_idesc_bf16(F_NPAD0);
        const uint32_t b_part = (2 * F_NPAD0 * 16) >> 4;                                    // 16-byte units
        const uint32_t b_lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | ((uint32_t)F_NPAD0 << 16);     // LBO = 96 * 16 B
        const uint32_t a_base = ((smem_u32(smem + p.off_a) & 0x3FFFFu) >> 4) | ((uint32_t)F_AROWS << 16);   // LBO = 136 * 16 B
        const uint32_t abuf16 = (uint32_t)p.abuf >> 4;
        const bool two_parts = !p.int_pixels, no_mma0 = (p.dbg & 8) != 0;
        for (int k = 0; k < total0; ++k) {
            const int buf = k & (F_CW - 1), slot = k & 1;
            mbar_wait_tag(&a_ready[buf], (uint32_t)((k / F_CW) & 1), 2000000 + k);
            mbar_wait_tag(&acc0_empty[slot], (uint32_t)(((k >> 1) & 1) ^ 1), 3000000 + k);
            tc_fence_after();
            const uint32_t a0 = a_base + (uint32_t)buf * abuf16;
            const uint32_t d = (uint32_t)slot * F_SLOT0;
            if (!no_mma0) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t ah = a0 + (uint32_t)dx, al = a0 + 2 * F_AROWS + (uint32_t)dx;
                    const uint32_t bh = b_lo0 + (uint32_t)(dx * 2) * b_part, bl = b_lo0 + (uint32_t)(dx * 2 + 1) * b_part;
                    tc_mma_bf16_elect(d, ah, bh, UMMA_DESC_HI, idesc, dx > 0 ? 1u : 0u);
                    if (two_parts) tc_mma_bf16_elect(d, al, bh, UMMA_DESC_HI, idesc, 1u);
                    tc_mma_bf16_elect(d, ah, bl, UMMA_DESC_HI, idesc, 1u);
                }
            }
            tc_commit_elect(&acc0_full[slot]);
            tc_commit_elect(&a_free[buf]);
        }
    } else if (warp == F_MMA1_WARP0) {
        f01_mma1<0>(p, smem, n_it, block_units, mid_cnt, acc1_full, acc1_empty, mid_free, tmem_base);
    } else if (warp == F_MMA1_WARP0 + 1) {
        f01_mma1<1>(p, smem, n_it, block_units, mid_cnt, acc1_full, acc1_empty, mid_free, tmem_base);
    } else if (warp >= F_DRAIN_WARP0 && warp < EPI_WARP0_F) {
        // ================= drain: TMEM -> ELU -> bf16 -> ring (lane = raster position, group g: rows g, g + DG, ..) =================
        const int q = warp & 3, grp = (warp - F_DRAIN_WARP0) >> 2;
        const int rows_mine = (F_R0 - grp + DG - 1) / DG;
        for (int k = 0; k < total0; ++k) {
            const int it = k / p.T0, t = k - it * p.T0, slot = k & 1;
            const int B0 = it * p.NB;
            const unsigned gw0 = (unsigned)(128 * t + 32 * q);
            const int rb_first = (int)__umulhi(gw0, p.wp_magic), rb_last = (int)__umulhi(gw0 + 31u, p.wp_magic);
            const unsigned g = gw0 + (unsigned)lane;
            const int rbl = (int)__umulhi(g, p.wp_magic);
            const int xp = (int)(g - (unsigned)rbl * (unsigned)p.Wp);
            if (rb_first < p.NB) mbar_wait_tag(&mid_free[(B0 + rb_first) % F_RB], (uint32_t)((((B0 + rb_first) / F_RB) & 1) ^ 1), 5000000 + k);
            if (rb_last != rb_first && rb_last < p.NB)
                mbar_wait_tag(&mid_free[(B0 + rb_last) % F_RB], (uint32_t)((((B0 + rb_last) / F_RB) & 1) ^ 1), 6000000 + k);
            mbar_wait_tag(&acc0_full[slot], (uint32_t)((k >> 1) & 1), 7000000 + k);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)slot * F_SLOT0 + ((uint32_t)(q * 32) << 16);
            const bool col_real = xp >= 1 && xp <= p.W;
            uint8_t *dst = ring + ((size_t)(((B0 + rbl) % F_RB) * F_R0) * p.Wp + xp) * 16;
            if (2 * DG == F_R0 && !(p.dbg & 4)) {
                // two rows per warp (grp, grp + DG): both loads in flight, one wait
                uint32_t raw[2][12];
                tmem_ld12_issue(taddr + (uint32_t)(grp * F_C0), raw[0]);
                tmem_ld12_issue(taddr + (uint32_t)((grp + DG) * F_C0), raw[1]);
                tmem_ld_wait();
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r = grp + h2 * DG;
                    uint32_t pk[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(elu_f(__uint_as_float(raw[h2][2 * j])),
                                                                       elu_f(__uint_as_float(raw[h2][2 * j + 1])));
                        pk[j] = col_real ? *reinterpret_cast<const uint32_t *>(&h) : 0u;
                    }
                    if (rbl < p.NB && rbl * F_R0 + r < p.H) {
                        *reinterpret_cast<uint4 *>(dst + (size_t)r * p.Wp * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4 *>(dst + (size_t)r * p.Wp * 16 + p.ring_plane) = make_uint4(pk[4], pk[5], 0u, 0u);
                    }
                }
            } else {
#pragma unroll 1
            for (int r = grp; r < ((p.dbg & 4) ? 0 : F_R0); r += DG) {
                float v[12];
                tmem_ld12(taddr + (uint32_t)(r * F_C0), v);
                uint32_t pk[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(elu_f(v[2 * j]), elu_f(v[2 * j + 1]));   // bias added by the MMA
                    pk[j] = col_real ? *reinterpret_cast<const uint32_t *>(&h) : 0u;
                }
                if (rbl < p.NB && rbl * F_R0 + r < p.H) {
                    *reinterpret_cast<uint4 *>(dst + (size_t)r * p.Wp * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4 *>(dst + (size_t)r * p.Wp * 16 + p.ring_plane) = make_uint4(pk[4], pk[5], 0u, 0u);
                }
            }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&acc0_empty[slot]);
                if (rb_first < p.NB) {
                    const int n1 = min(32, (rb_first + 1) * p.Wp - (int)gw0);
                    red_release_shared_add(&mid_cnt[(B0 + rb_first) % F_RB], (uint32_t)(n1 * rows_mine));
                    if (rb_last != rb_first && rb_last < p.NB)
                        red_release_shared_add(&mid_cnt[(B0 + rb_last) % F_RB], (uint32_t)((32 - n1) * rows_mine));
                }
            }
        }
    } else if (warp >= EPI_WARP0_F) {
        // ================= epilogue: TMEM -> 2x2 max -> bias + ELU -> bf16 -> global =================
        const int q = warp & 3, grp = (warp - EPI_WARP0_F) >> 2;
        const int odd = lane & 1;
        const int w = EG == 4 ? grp >> 1 : grp;                   // the MMA-1 warp whose tiles this group drains
        const int ngw = (p.NG - w + 1) / 2;                       // row groups per sample of that warp
        const int total1 = n_it * ngw * p.JT;                     // its tiles
        const int n_groups16 = (p.cout1 + 3) >> 2;
        const int nchr = (p.cout1 + 7) >> 3;
        for (int u = (EG == 4 ? grp & 1 : 0); u < total1; u += (EG == 4 ? 2 : 1)) {
            const uint32_t slot = EG == 4 ? (uint32_t)grp : (uint32_t)(2 * w + (u & 1));
            const uint32_t sph = ((uint32_t)u >> 1) & 1u;
            const int G = u / p.JT, j = u - G * p.JT;
            const int it = G / ngw, rg = w + 2 * (G - it * ngw);
            const long long n = (long long)blockIdx.x + (long long)it * gridDim.x;
            mbar_wait_tag(&acc1_full[slot], sph, 8000000 + u);
            tc_fence_after();
            const int c = 1 + 128 * j + q * 32 + lane;            // padded column of this lane
            const bool valid = c <= p.W;
            uint8_t *out_n = reinterpret_cast<uint8_t *>(p.out) + n * p.out_sample;
            const uint32_t taddr = tmem_base + 2 * F_SLOT0 + slot * F_SLOT1 + ((uint32_t)(q * 32) << 16);
            // rows (0,1) and (2,3) pool vertically inside the thread; lanes (2k, 2k+1) are one pooled column:
            // the even lane finishes pooled row 0, the odd lane pooled row 1
            const int yo = 2 * rg + odd;
            const long long opos = ((long long)(yo + 1) * p.Wpo + ((c - 1) >> 1) + 1) * 16;
            for (int h = 0; h < ((p.dbg & 1) ? 0 : nchr); ++h) {
                float v[32];
                tmem_ld8x4(taddr + (uint32_t)(h * 8), (uint32_t)F_NP1, v);
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
                        const float4 b4 = *reinterpret_cast<const float4 *>(&bias1_sm[h * 8 + 4 * q4]);
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc1_empty[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == F_MMA0_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}
