// Fused L2-normalise + 32-dim dot + per-query top-k over a resident embedding DB, the
// rank-of-target variant for eval_retrieval, the candidate-list merge and the piece vote.
//
// Replaces (reference, paths relative to its root):
//   audio_sheet_retrieval/audio_sheet_server.py:530-563  cdist(DB, q, 'cosine') + argsort[:n]
//   audio_sheet_retrieval/audio_sheet_server.py:230-240  per-window loop + np.unique vote
//   audio_sheet_retrieval/utils/train_dcca_pool.py:39-74 N x N cdist + per-row argsort
//
// Data movement: the DB (n,32) fp32 row-major is streamed once per query tile with 2-D TMA
// tensor loads (128-byte rows, SWIZZLE_128B so that "one thread = one row" reads are
// bank-conflict free) through a 3-stage mbarrier ring; scores follow the pinned-order fp32
// definition of oracle/search.py (separately rounded * and +, k = 0..31) so the returned
// indices are bit-exact; candidates live in registers / shared memory; the distance matrix
// is never written anywhere.
#include <cuda.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace asr {

constexpr int TK_THREADS = 256;
constexpr int TK_ROWS = 256;                 // DB rows per stage = 32 KB
constexpr int TK_QCAP = 512;                 // candidate queue entries per query
constexpr int TK_STAGE_BYTES = TK_ROWS * 128;
constexpr size_t TK_SCRATCH_BYTES = 64u << 20;
constexpr int TK_QT_MAX = 16;
constexpr int TK_STAGES = 3;                 // used by the rank kernel

// Per query-tile state.  QT queries are scored per pass over the DB slice; the smaller
// instantiations fit several CTAs per SM (the Q = 1 streaming case is latency-bound per warp:
// the pinned summation order is one dependent chain per row, so occupancy hides it).
template <int QT, int STAGES>
struct TkSmem {
    float stage[STAGES][TK_ROWS * 32];       // 1024-byte aligned (SWIZZLE_128B)
    float q[QT][32];
    float qs[QT][TK_QCAP];                   // candidate queue: score
    uint32_t qi[QT][TK_QCAP];                //                  local row
    float ls[2][QT][ASR_MAX_K];              // sorted lists (double buffered for the rank merge)
    uint32_t li[2][QT][ASR_MAX_K];
    int qcount[QT];
    int llen[QT];
    int lbuf[QT];
    float thr_s[QT];
    uint32_t thr_i[QT];
    uint64_t full[STAGES];
    int flush_flag[3];
};

__device__ __forceinline__ bool beats(float sa, uint32_t ia, float sb, uint32_t ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

// Pinned-order helpers (oracle/search.py): no FMA contraction anywhere.
__device__ __forceinline__ void normalise32(float *d) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) ss = __fadd_rn(ss, __fmul_rn(d[k], d[k]));
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(ss));
#pragma unroll
    for (int k = 0; k < 32; ++k) d[k] = __fmul_rn(d[k], inv);
}
__device__ __forceinline__ float score32(const float *q, const float *d) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = __fadd_rn(acc, __fmul_rn(q[k], d[k]));
    return (acc != acc) ? -CUDART_INF_F : acc;
}
// four queries at once: four independent dependent-add chains give the scheduler ILP
__device__ __forceinline__ void score32x4(const float (*q)[32], const float *d, float *out) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        const float4 q0 = *reinterpret_cast<const float4 *>(&q[0][4 * kb]);
        const float4 q1 = *reinterpret_cast<const float4 *>(&q[1][4 * kb]);
        const float4 q2 = *reinterpret_cast<const float4 *>(&q[2][4 * kb]);
        const float4 q3 = *reinterpret_cast<const float4 *>(&q[3][4 * kb]);
        const float d0 = d[4 * kb], d1 = d[4 * kb + 1], d2 = d[4 * kb + 2], d3 = d[4 * kb + 3];
        a0 = __fadd_rn(a0, __fmul_rn(q0.x, d0)); a1 = __fadd_rn(a1, __fmul_rn(q1.x, d0));
        a2 = __fadd_rn(a2, __fmul_rn(q2.x, d0)); a3 = __fadd_rn(a3, __fmul_rn(q3.x, d0));
        a0 = __fadd_rn(a0, __fmul_rn(q0.y, d1)); a1 = __fadd_rn(a1, __fmul_rn(q1.y, d1));
        a2 = __fadd_rn(a2, __fmul_rn(q2.y, d1)); a3 = __fadd_rn(a3, __fmul_rn(q3.y, d1));
        a0 = __fadd_rn(a0, __fmul_rn(q0.z, d2)); a1 = __fadd_rn(a1, __fmul_rn(q1.z, d2));
        a2 = __fadd_rn(a2, __fmul_rn(q2.z, d2)); a3 = __fadd_rn(a3, __fmul_rn(q3.z, d2));
        a0 = __fadd_rn(a0, __fmul_rn(q0.w, d3)); a1 = __fadd_rn(a1, __fmul_rn(q1.w, d3));
        a2 = __fadd_rn(a2, __fmul_rn(q2.w, d3)); a3 = __fadd_rn(a3, __fmul_rn(q3.w, d3));
    }
    out[0] = (a0 != a0) ? -CUDART_INF_F : a0;
    out[1] = (a1 != a1) ? -CUDART_INF_F : a1;
    out[2] = (a2 != a2) ? -CUDART_INF_F : a2;
    out[3] = (a3 != a3) ? -CUDART_INF_F : a3;
}

// One warp merges up to 32 candidates (one per lane; dead lanes carry alive=false) into a sorted
// list by rank counting: final position = number of elements that beat you.  List entries are
// held in registers (ceil(len/32) rounds) so every shuffle is warp-uniform.  I = index type.
template <typename I>
__device__ __forceinline__ bool beats_t(float sa, I ia, float sb, I ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}
template <typename I>
__device__ void warp_merge_batch(const float *ls, const I *li, float *ns, I *ni, int len, int k, float cs, I ci,
                                 bool alive, unsigned alive_mask, I sentinel, int lane) {
    int pos = 0;
    for (int j = 0; j < len; ++j) pos += beats_t<I>(ls[j], li[j], cs, ci) ? 1 : 0;   // broadcast reads
    float es[ASR_MAX_K / 32];
    I ei[ASR_MAX_K / 32];
    int shift[ASR_MAX_K / 32];
#pragma unroll
    for (int r = 0; r < ASR_MAX_K / 32; ++r) {
        int j = r * 32 + lane;
        bool v = j < len;
        es[r] = v ? ls[j] : -CUDART_INF_F;
        ei[r] = v ? li[j] : sentinel;
        shift[r] = 0;
    }
    for (int l = 0; l < 32; ++l) {
        if (!((alive_mask >> l) & 1u)) continue;          // warp-uniform
        float os = __shfl_sync(0xffffffffu, cs, l);
        I oi = __shfl_sync(0xffffffffu, ci, l);
        if (l != lane) pos += beats_t<I>(os, oi, cs, ci) ? 1 : 0;
#pragma unroll
        for (int r = 0; r < ASR_MAX_K / 32; ++r) shift[r] += beats_t<I>(os, oi, es[r], ei[r]) ? 1 : 0;
    }
#pragma unroll
    for (int r = 0; r < ASR_MAX_K / 32; ++r) {
        int j = r * 32 + lane;
        if (j < len) {
            int np = j + shift[r];
            if (np < k) { ns[np] = es[r]; ni[np] = ei[r]; }
        }
    }
    if (alive && pos < k) { ns[pos] = cs; ni[pos] = ci; }
    __syncwarp();
}

// Drain the candidate queues into the sorted lists.  Warp w owns queries w, w+8, ...
template <int QT, int STAGES>
__device__ void flush_queues(TkSmem<QT, STAGES> &sm, int nq_tile, int k, int warp, int lane) {
    for (int q = warp; q < nq_tile; q += TK_THREADS / 32) {
        int n = sm.qcount[q];
        n = n < TK_QCAP ? n : TK_QCAP;
        for (int base = 0; base < n; base += 32) {
            int c = base + lane;
            bool alive = c < n;
            float cs = alive ? sm.qs[q][c] : -CUDART_INF_F;
            uint32_t ci = alive ? sm.qi[q][c] : 0xffffffffu;
            const int len = sm.llen[q], cur = sm.lbuf[q];
            if (alive && len == k) alive = beats(cs, ci, sm.thr_s[q], sm.thr_i[q]);
            const unsigned mask = __ballot_sync(0xffffffffu, alive);
            if (mask == 0u) continue;
            warp_merge_batch<uint32_t>(sm.ls[cur][q], sm.li[cur][q], sm.ls[cur ^ 1][q], sm.li[cur ^ 1][q], len, k, cs, ci,
                                       alive, mask, 0xffffffffu, lane);
            if (lane == 0) {
                int nl = len + __popc(mask);
                nl = nl < k ? nl : k;
                sm.llen[q] = nl;
                sm.lbuf[q] = cur ^ 1;
                if (nl == k) { sm.thr_s[q] = sm.ls[cur ^ 1][q][k - 1]; sm.thr_i[q] = sm.li[cur ^ 1][q][k - 1]; }
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) sm.qcount[q] = 0;
    }
}

template <int QT, int STAGES>
__device__ __forceinline__ void push_candidate(TkSmem<QT, STAGES> &sm, int qq, float s, uint32_t row, int k, bool valid,
                                               int &overflow) {
    bool pass = valid && (sm.llen[qq] < k || beats(s, row, sm.thr_s[qq], sm.thr_i[qq]));
    if (pass) {
        int p = atomicAdd(&sm.qcount[qq], 1);
        if (p < TK_QCAP) { sm.qs[qq][p] = s; sm.qi[qq][p] = row; }
        if (p >= TK_QCAP - TK_ROWS) overflow = 1;
    }
}

// grid: (n_slices, n_qgroups).  part_s/part_i: (nq, n_slices, k).
template <int QT, int STAGES, int MINB>
__global__ void __launch_bounds__(TK_THREADS, MINB)
topk_stream_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_db, int tiles_per_slice, const float *__restrict__ q,
                   int nq, int k, int normalise, float *__restrict__ part_s, uint32_t *__restrict__ part_i) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    typedef TkSmem<QT, STAGES> Smem;
    // No integer round-trip on the pointer: the compiler must keep the shared address space (LDS/STS,
    // not generic LD/ST).  Dynamic shared memory starts at offset 0 of the CTA window (no static
    // __shared__ in this kernel), which satisfies SWIZZLE_128B's 1024-byte alignment; checked below.
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (smem_u32(smem_raw) & 1023u) __trap();
    const int slice = blockIdx.x, n_slices = gridDim.x;
    const int64_t n_tiles = (n_db + TK_ROWS - 1) / TK_ROWS;
    const int64_t tile0 = (int64_t)slice * tiles_per_slice;
    int64_t my_tiles = n_tiles - tile0;
    if (my_tiles > tiles_per_slice) my_tiles = tiles_per_slice;
    if (my_tiles < 0) my_tiles = 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    const int n_qt = (nq + QT - 1) / QT;
    int64_t it_global = 0;   // running tile counter across query tiles (barrier phases continue)
    for (int qt = blockIdx.y; qt < n_qt; qt += gridDim.y) {
        const int q0 = qt * QT;
        const int nq_tile = min(QT, nq - q0);
        // stage the query tile (normalised with the pinned definition); unused rows are zero
        if (tid < QT) {
            float v[32];
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) v[kk] = tid < nq_tile ? q[(int64_t)(q0 + tid) * 32 + kk] : 0.f;
            if ((normalise & 1) && tid < nq_tile) normalise32(v);
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) sm.q[tid][kk] = v[kk];
            sm.qcount[tid] = 0;
            sm.llen[tid] = 0;
            sm.lbuf[tid] = 0;
            sm.thr_s[tid] = -CUDART_INF_F;
            sm.thr_i[tid] = 0xffffffffu;
        }
        if (tid == 0) {
            sm.flush_flag[0] = sm.flush_flag[1] = sm.flush_flag[2] = 0;
            for (int s = 0; s < STAGES && s < my_tiles; ++s) {     // prologue: fill the ring
                int slot = (int)((it_global + s) % STAGES);
                mbar_expect_tx(&sm.full[slot], TK_STAGE_BYTES);
                tma_load_2d(sm.stage[slot], &tmap, 0, (int)((tile0 + s) * TK_ROWS), &sm.full[slot]);
            }
        }
        __syncthreads();

        for (int64_t it = 0; it < my_tiles; ++it, ++it_global) {
            const int slot = (int)(it_global % STAGES);
            const uint32_t parity = (uint32_t)((it_global / STAGES) & 1);
            mbar_wait(&sm.full[slot], parity);
            // my row: chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4)   (SWIZZLE_128B)
            float d[32];
            const float4 *rowp = reinterpret_cast<const float4 *>(sm.stage[slot] + tid * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = rowp[c ^ (tid & 7)];
                d[4 * c + 0] = v.x; d[4 * c + 1] = v.y; d[4 * c + 2] = v.z; d[4 * c + 3] = v.w;
            }
            // One barrier per tile: (a) every thread holds its row, so the slot can be refilled;
            // (b) the previous tile's candidate pushes are complete, so its overflow flag is final.
            __syncthreads();
            if (tid == 0 && it + STAGES < my_tiles) {
                mbar_expect_tx(&sm.full[slot], TK_STAGE_BYTES);
                tma_load_2d(sm.stage[slot], &tmap, 0, (int)((tile0 + it + STAGES) * TK_ROWS), &sm.full[slot]);
            }
            // flags rotate over three slots: written during tile it, read here at it+1, cleared at it+2
            if (it > 0 && sm.flush_flag[(it - 1) % 3]) {
                flush_queues(sm, nq_tile, k, warp, lane);
                __syncthreads();
            }
            if (tid == 0 && it > 1) sm.flush_flag[(it - 2) % 3] = 0;
            const int64_t row = (tile0 + it) * TK_ROWS + tid;
            const bool valid = row < n_db;
            if (normalise & 2) normalise32(d);       // bit 1 clear: the rows come from the pinned-normalised copy
            int overflow = 0;
            if (QT == 1) {
                float s = score32(sm.q[0], d);
                push_candidate(sm, 0, s, (uint32_t)row, k, valid, overflow);
            } else {
#pragma unroll 1
                for (int qq = 0; qq < nq_tile; qq += 4) {
                    float s[4];
                    score32x4(&sm.q[qq], d, s);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (qq + j < nq_tile) push_candidate(sm, qq + j, s[j], (uint32_t)row, k, valid, overflow);
                }
            }
            if (overflow) sm.flush_flag[it % 3] = 1;
        }
        __syncthreads();
        flush_queues(sm, nq_tile, k, warp, lane);
        __syncthreads();
        // write this slice's lists
        for (int e = tid; e < nq_tile * k; e += TK_THREADS) {
            int qq = e / k, j = e % k;
            int cur = sm.lbuf[qq];
            bool v = j < sm.llen[qq];
            size_t o = ((size_t)(q0 + qq) * n_slices + slice) * k + j;
            part_s[o] = v ? sm.ls[cur][qq][j] : -CUDART_INF_F;
            part_i[o] = v ? sm.li[cur][qq][j] : 0xffffffffu;
        }
        __syncthreads();
    }
}

// Merge n_lists sorted lists (each k long) per query into the final top-k.  One CTA of four
// warps per query: warp w folds lists w, w+4, ... into its own list, warp 0 folds the four.
// Input either (uint32 local rows + idx_base) or int64 global indices.
// MG_WARPS warps per query fold the per-slice lists (few lists per query, or 64-bit indices after the
// multi-GPU all-gather); see topk_merge_select_kernel for the many-lists case.
template <int MG_WARPS>
__global__ void __launch_bounds__(MG_WARPS * 32)
topk_merge_kernel(const float *__restrict__ in_s, const uint32_t *__restrict__ in_i32, const int64_t *__restrict__ in_i64,
                  int64_t idx_base, int n_lists, int k, float *__restrict__ out_s, int64_t *__restrict__ out_i,
                  int64_t q_stride, int64_t list_stride_s, int64_t list_stride_i) {
    // element (query qi, list l, position pos) sits at qi * q_stride + l * list_stride + pos: (nq, n_lists, k) arrays
    // have q_stride = n_lists * k, list_stride = k; the rank-major result of an all-gather has q_stride = k and
    // list_stride = one rank's chunk (scores and indices may have different chunk strides)
    extern __shared__ __align__(16) uint8_t mg_raw[];
    // [warp][2][k] int64 indices, then [warp][2][k] float scores
    int64_t *li_base = reinterpret_cast<int64_t *>(mg_raw);
    float *ls_base = reinterpret_cast<float *>(mg_raw + (size_t)MG_WARPS * 2 * k * sizeof(int64_t));
    auto LS = [&](int w, int c) { return ls_base + ((size_t)w * 2 + c) * k; };
    auto LI = [&](int w, int c) { return li_base + ((size_t)w * 2 + c) * k; };
    __shared__ int wlen[MG_WARPS], wcur[MG_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t qi = blockIdx.x;
    const size_t base = (size_t)qi * q_stride;
    int len = 0, cur = 0;
    float thr_s = -CUDART_INF_F;
    int64_t thr_i = INT64_MAX;
    auto fold = [&](float cs, int64_t ci, bool alive, int w) {
        if (alive && len == k) alive = beats_t<int64_t>(cs, ci, thr_s, thr_i);
        const unsigned mask = __ballot_sync(0xffffffffu, alive);
        if (mask == 0u) return;
        warp_merge_batch<int64_t>(LS(w, cur), LI(w, cur), LS(w, cur ^ 1), LI(w, cur ^ 1), len, k, cs, ci, alive, mask,
                                  (int64_t)INT64_MAX, lane);
        len = min(k, len + __popc(mask));
        cur ^= 1;
        if (len == k) { thr_s = LS(w, cur)[k - 1]; thr_i = LI(w, cur)[k - 1]; }
        __syncwarp();
    };
    // phase 1: each warp folds its lists; candidates are taken position-major across lists so the
    // threshold rises quickly (position 0 of every list first)
    const int my_lists = (n_lists - warp + MG_WARPS - 1) / MG_WARPS;
    const int total = my_lists * k;
    for (int b = 0; b < total; b += 32) {
        int c = b + lane;
        bool alive = c < total;
        float cs = -CUDART_INF_F;
        int64_t ci = INT64_MAX;
        if (alive) {
            int pos = c / my_lists, l = warp + (c % my_lists) * MG_WARPS;
            cs = in_s[base + (size_t)l * list_stride_s + pos];
            if (in_i32) {
                uint32_t r = in_i32[base + (size_t)l * list_stride_i + pos];
                alive = r != 0xffffffffu;
                ci = (int64_t)r + idx_base;
            } else {
                ci = in_i64[base + (size_t)l * list_stride_i + pos];
                alive = ci >= 0;
            }
            if (!alive) ci = INT64_MAX;
        }
        fold(cs, ci, alive, warp);
    }
    if (lane == 0) { wlen[warp] = len; wcur[warp] = cur; }
    __syncthreads();
    // phase 2: warp 0 folds the other warps' lists into its own
    if (warp == 0) {
        for (int w = 1; w < MG_WARPS; ++w) {
            const int n = wlen[w], c2 = wcur[w];
            for (int b = 0; b < n; b += 32) {
                int c = b + lane;
                bool alive = c < n;
                float cs = alive ? LS(w, c2)[c] : -CUDART_INF_F;
                int64_t ci = alive ? LI(w, c2)[c] : INT64_MAX;
                fold(cs, ci, alive, 0);
            }
        }
        for (int j = lane; j < k; j += 32) {
            bool v = j < len;
            out_s[(size_t)qi * k + j] = v ? LS(0, cur)[j] : -CUDART_INF_F;
            out_i[(size_t)qi * k + j] = v ? LI(0, cur)[j] : -1;
        }
    }
}

// Merge for few queries and MANY lists (the streaming-server case: one query, one list per CTA of the stream
// kernel = 444 lists): the fold above is a chain of dependent global loads (62 us measured at 444 x 25).  Here
// every candidate becomes a unique 64-bit key (order-preserving score bits | ~row: larger = better, i.e. score
// descending then row ascending), each thread keeps its keys in registers, an 8-pass byte-wise radix select finds
// the k-th largest key exactly, and the k survivors are ranked by counting.  Deterministic, no overflow case.
constexpr int MS_THREADS = 1024;                  // MS_KPT keys per thread: 12 (<= 12288 candidates per query) or 16 (<= 16384)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
template <int MS_KPT>
__global__ void __launch_bounds__(MS_THREADS)
topk_merge_select_kernel(const float *__restrict__ in_s, const uint32_t *__restrict__ in_i32, int64_t idx_base, int n_cand,
                         int k, float *__restrict__ out_s, int64_t *__restrict__ out_i) {
    __shared__ unsigned hist[256];
    __shared__ unsigned long long sel[ASR_MAX_K];
    __shared__ unsigned long long prefix_sm;
    __shared__ int krem_sm, nsel, nvalid_sm;
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * n_cand;
    unsigned long long key[MS_KPT];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < MS_KPT; ++j) {
        const int c = tid + j * MS_THREADS;
        key[j] = 0ull;                                   // 0 = "no candidate" (a real key has ~row != 0 or score bits != 0)
        // both loads unconditional (clamped) so that all 48 of a thread are in flight together
        const size_t o = base + (size_t)min(c, n_cand - 1);
        const uint32_t r = in_i32[o];
        const float sc = in_s[o];
        if (c < n_cand && r != 0xffffffffu) { key[j] = ((unsigned long long)f2ord(sc) << 32) | (unsigned long long)(~r); ++mine; }
    }
    if (tid == 0) { nvalid_sm = 0; nsel = 0; prefix_sm = 0ull; }
    __syncthreads();
    if (mine) atomicAdd(&nvalid_sm, mine);
    __syncthreads();
    const int k_eff = min(k, nvalid_sm);
    if (tid == 0) krem_sm = k_eff;
    // radix select of the k_eff-th largest key, most significant byte first
#pragma unroll
    for (int pass = 7; pass >= 0; --pass) {             // unrolled: the 64-bit shifts become register selects
        if (k_eff == 0) break;
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const unsigned long long prefix = prefix_sm;
#pragma unroll
        for (int j = 0; j < MS_KPT; ++j) {
            const unsigned long long kk = key[j];
            const bool match = kk != 0ull && (pass == 7 || (kk >> (8 * (pass + 1))) == prefix);
            const unsigned bin = match ? ((unsigned)(kk >> (8 * pass)) & 255u) : 256u;
            if (pass >= 5) {
                // leading bytes of cosine scores are nearly all equal: aggregate per warp, plain atomics would serialise
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (match && (int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
            } else if (match) {
                // by now only the keys near the threshold still match the prefix, spread over many bins
                atomicAdd(&hist[bin], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {      // warp 0 walks the histogram from the top: lane l owns bins 255 - 8 l ... 248 - 8 l
            unsigned h[8], tot = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) { h[i] = hist[255 - 8 * tid - i]; tot += h[i]; }
            unsigned inc = tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
                if (tid >= off) inc += t;
            }
            const unsigned exc = inc - tot, rem = (unsigned)krem_sm;
            __syncwarp();
            if (exc < rem && rem <= inc) {            // the wanted key lies in this lane's bins (exactly one lane)
                unsigned r = rem - exc;
                int i = 0;
                for (; i < 7; ++i) {
                    if (h[i] >= r) break;
                    r -= h[i];
                }
                krem_sm = (int)r;
                prefix_sm = (prefix << 8) | (unsigned long long)(255 - 8 * tid - i);
            }
        }
        __syncthreads();
    }
    const unsigned long long T = prefix_sm;             // the k_eff-th largest key (keys are unique)
    if (k_eff > 0) {
#pragma unroll
        for (int j = 0; j < MS_KPT; ++j) {
            if (key[j] != 0ull && key[j] >= T) sel[atomicAdd(&nsel, 1)] = key[j];
        }
    }
    __syncthreads();
    for (int j = tid; j < k; j += MS_THREADS) {
        if (j < k_eff) {
            const unsigned long long mk = sel[j];
            int rank = 0;
            for (int t = 0; t < k_eff; ++t) rank += sel[t] > mk ? 1 : 0;
            out_s[(size_t)blockIdx.x * k + rank] = ord2f((uint32_t)(mk >> 32));
            out_i[(size_t)blockIdx.x * k + rank] = (int64_t)(~(uint32_t)mk) + idx_base;
        } else {
            out_s[(size_t)blockIdx.x * k + j] = -CUDART_INF_F;
            out_i[(size_t)blockIdx.x * k + j] = -1;
        }
    }
}

// few queries over many per-slice lists: radix-select merge; otherwise the 4-warp fold
static void launch_merge(const float *in_s, const uint32_t *in_i32, const int64_t *in_i64, int64_t idx_base, int n_lists, int k,
                         float *out_s, int64_t *out_i, int64_t nq, cudaStream_t st) {
    const int64_t n_cand = (int64_t)n_lists * k;
    if (in_i32 && nq <= 8 && n_lists >= 64 && n_cand <= (int64_t)MS_THREADS * 12) {
        topk_merge_select_kernel<12><<<(unsigned)nq, MS_THREADS, 0, st>>>(in_s, in_i32, idx_base, (int)n_cand, k, out_s, out_i);
    } else if (in_i32 && nq <= 256 && n_lists >= 64 && n_cand <= (int64_t)MS_THREADS * 16) {
        // few queries over very many lists (the pre-filter's lists per column range x slices): the fold below is a
        // chain of dependent passes per list, the radix select is one pass over the candidates
        topk_merge_select_kernel<16><<<(unsigned)nq, MS_THREADS, 0, st>>>(in_s, in_i32, idx_base, (int)n_cand, k, out_s, out_i);
    } else {
        const size_t smem = (size_t)4 * 2 * k * (sizeof(int64_t) + sizeof(float));
        topk_merge_kernel<4><<<(unsigned)nq, 4 * 32, smem, st>>>(in_s, in_i32, in_i64, idx_base, n_lists, k, out_s, out_i,
                                                                 (int64_t)n_lists * k, k, k);
    }
}

// best correct item over the shards of an eval_retrieval query set: max score, ties to the smaller global index
// (the torch.stack / max / where of a host-side merge, as one kernel).  ts_all / ti_all: list l of query i at
// l * stride + i.
__global__ void rank_target_merge_kernel(const float *__restrict__ ts_all, const int64_t *__restrict__ ti_all, int n_lists,
                                         int64_t nq, int64_t stride_s, int64_t stride_i, float *__restrict__ ts,
                                         int64_t *__restrict__ ti) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float best = -CUDART_INF_F;
    int64_t bi = -1;
    for (int l = 0; l < n_lists; ++l) {
        const float sc = ts_all[(size_t)l * stride_s + i];
        const int64_t ix = ti_all[(size_t)l * stride_i + i];
        if (ix < 0) continue;                          // this shard holds none of the query's correct items
        if (bi < 0 || sc > best || (sc == best && ix < bi)) { best = sc; bi = ix; }
    }
    ts[i] = best;
    ti[i] = bi;
}

// ---- rank of target ---------------------------------------------------------------
// phase 0: best correct item per query among this shard's rows (thread per query).
__global__ void rank_target_kernel(const float *__restrict__ db, int64_t n_db, int64_t idx_base, const float *__restrict__ q,
                                   int nq, int64_t q_base, int64_t kg, int64_t hg, int normalise,
                                   float *__restrict__ tscore, int64_t *__restrict__ tidx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float qv[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) qv[kk] = q[(int64_t)i * 32 + kk];
    if (normalise) normalise32(qv);
    const int64_t g = (q_base + i) / hg;
    float best = -CUDART_INF_F;
    int64_t bj = -1;
    for (int64_t j = g * kg; j < (g + 1) * kg; ++j) {
        int64_t r = j - idx_base;
        if (r < 0 || r >= n_db) continue;
        float d[32];
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) d[kk] = db[r * 32 + kk];
        if (normalise) normalise32(d);
        float s = score32(qv, d);
        if (bj < 0 || s > best) { best = s; bj = j; }
    }
    tscore[i] = best;
    tidx[i] = bj;
}

struct RkSmem {
    float stage[TK_STAGES][TK_ROWS * 32];
    float q[TK_QT_MAX][32];
    float ts[TK_QT_MAX];
    int64_t ti[TK_QT_MAX];
    unsigned long long cnt[TK_QT_MAX];
    uint64_t full[TK_STAGES];
};

// phase 1: count rows ranked before the target.  grid (n_slices, n_qgroups).
__global__ void __launch_bounds__(TK_THREADS, 1)
rank_count_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_db, int64_t idx_base, int tiles_per_slice,
                  const float *__restrict__ q, int nq, int normalise, const float *__restrict__ tscore,
                  const int64_t *__restrict__ tidx, unsigned long long *__restrict__ better) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    RkSmem &sm = *reinterpret_cast<RkSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    if (smem_u32(smem_raw) & 1023u) __trap();
    const int slice = blockIdx.x;
    const int64_t n_tiles = (n_db + TK_ROWS - 1) / TK_ROWS;
    const int64_t tile0 = (int64_t)slice * tiles_per_slice;
    int64_t my_tiles = n_tiles - tile0;
    if (my_tiles > tiles_per_slice) my_tiles = tiles_per_slice;
    if (my_tiles < 0) my_tiles = 0;
    if (tid == 0) {
        for (int s = 0; s < TK_STAGES; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();
    const int n_qt = (nq + TK_QT_MAX - 1) / TK_QT_MAX;
    int64_t it_global = 0;
    for (int qt = blockIdx.y; qt < n_qt; qt += gridDim.y) {
        const int q0 = qt * TK_QT_MAX;
        const int nq_tile = min(TK_QT_MAX, nq - q0);
        if (tid < nq_tile) {
            float v[32];
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) v[kk] = q[(int64_t)(q0 + tid) * 32 + kk];
            if (normalise) normalise32(v);
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) sm.q[tid][kk] = v[kk];
            sm.ts[tid] = tscore[q0 + tid];
            sm.ti[tid] = tidx[q0 + tid];
            sm.cnt[tid] = 0ull;
        }
        if (tid == 0) {
            for (int s = 0; s < TK_STAGES && s < my_tiles; ++s) {
                int slot = (int)((it_global + s) % TK_STAGES);
                mbar_expect_tx(&sm.full[slot], TK_STAGE_BYTES);
                tma_load_2d(sm.stage[slot], &tmap, 0, (int)((tile0 + s) * TK_ROWS), &sm.full[slot]);
            }
        }
        __syncthreads();
        unsigned cnt[TK_QT_MAX];
#pragma unroll
        for (int qq = 0; qq < TK_QT_MAX; ++qq) cnt[qq] = 0;
        for (int64_t it = 0; it < my_tiles; ++it, ++it_global) {
            const int slot = (int)(it_global % TK_STAGES);
            const uint32_t parity = (uint32_t)((it_global / TK_STAGES) & 1);
            mbar_wait(&sm.full[slot], parity);
            float d[32];
            const float4 *rowp = reinterpret_cast<const float4 *>(sm.stage[slot] + tid * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = rowp[c ^ (tid & 7)];
                d[4 * c + 0] = v.x; d[4 * c + 1] = v.y; d[4 * c + 2] = v.z; d[4 * c + 3] = v.w;
            }
            __syncthreads();
            if (tid == 0 && it + TK_STAGES < my_tiles) {
                mbar_expect_tx(&sm.full[slot], TK_STAGE_BYTES);
                tma_load_2d(sm.stage[slot], &tmap, 0, (int)((tile0 + it + TK_STAGES) * TK_ROWS), &sm.full[slot]);
            }
            const int64_t row = (tile0 + it) * TK_ROWS + tid;
            const bool valid = row < n_db;
            const int64_t grow = row + idx_base;
            if (normalise) normalise32(d);
#pragma unroll
            for (int qq = 0; qq < TK_QT_MAX; ++qq) {
                if (qq < nq_tile) {
                    float s = score32(sm.q[qq], d);
                    float t = sm.ts[qq];
                    bool b = valid && ((s > t) || (s == t && grow < sm.ti[qq]));
                    cnt[qq] += b ? 1u : 0u;
                }
            }
        }
#pragma unroll
        for (int qq = 0; qq < TK_QT_MAX; ++qq) {
            unsigned c = cnt[qq];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0 && qq < nq_tile && c) atomicAdd(&sm.cnt[qq], (unsigned long long)c);
        }
        __syncthreads();
        if (tid < nq_tile && sm.cnt[tid]) atomicAdd(&better[q0 + tid], sm.cnt[tid]);
        __syncthreads();
    }
}

// ---- tensor-core pre-filter + exact re-scoring ------------------------------------------
// From a handful of queries on the pinned-order scoring (64 dependent fp32 instructions per (row, query)) is
// FP32-issue bound.  This path scores a 128-query x 256-row tile with four tcgen05 tf32 MMAs
// (fp32 accumulators in TMEM), which is only an APPROXIMATION s~ of the pinned score s, with
// |s~ - s| <= 2^-9 for unit vectors (each tf32 operand carries <= 2^-10 relative error and
// sum |q_k d_k| <= 1).  Exactness is kept by filtering, not by trusting s~: a row is re-scored with
// the pinned fp32 order iff s~ >= tau_q - eps, where tau_q is any LOWER BOUND of the query's final k-th best EXACT
// score (the list's current k-th entry, the warm-up floor, the bound the slices share) and eps = 2^-8.  Every true
// top-k row satisfies s >= tau_final >= tau_q, hence
// s~ >= tau_q - eps, so it is always re-scored; lists, thresholds and the final order use exact
// scores only.  Results are therefore identical to topk_stream_kernel / the oracle.
// Layout: TMEM lane = query.  SCANNER warps read the score tile out of TMEM and only FILTER it against the query's
// current threshold; what passes goes, as (row, query) entries, into shared-memory rings that OWNER warps consume:
// exact re-scoring and insertion into the query's top-k list (one list per query and work item: a sorted part plus an
// append buffer that is folded in by a bitonic merge).
constexpr int TC_QM = 128;
constexpr int TC_ROWS = 256;
constexpr int TC_STAGES = 3;
constexpr int TC_KMAX = 32;
constexpr int TC_SCAN_WARPS = 16;    // four per TMEM lane quarter: each scans one 64-column chunk of every tile
constexpr int TC_OWN_WARPS = 8;      // list owners: query qi of the tile belongs to owner qi % TC_OWN_WARPS
constexpr int TC_SCAN_WARP0 = 2, TC_OWN_WARP0 = 2 + TC_SCAN_WARPS;   // warp 0 TMA, warp 1 MMA
constexpr int TC_THREADS = 32 * (2 + TC_SCAN_WARPS + TC_OWN_WARPS);   // 832

constexpr int TC_LSTRIDE = 2 * TC_KMAX + 1;   // per query: sorted list [0, 32) + append buffer [32, 64); odd stride spreads the lists over the banks
constexpr int TC_RING = 256;              // candidate ring entries per owner warp (power of two)

struct TcSmem {
    float stage[TC_STAGES][TC_ROWS * 32];     // normalised DB rows, SWIZZLE_128B (written by TMA)
    float q[TC_QM * 32];                      // normalised queries, same swizzle (written by threads)
    unsigned long long lists[TC_QM * TC_LSTRIDE];   // per query: sorted top-k list + unsorted buffer of accepted candidates, 64-bit keys
    uint32_t fmin_key[TC_QM];                 // warm-up floor: minimum over the query's scanner threads (atomicMin)
    uint32_t pub[TC_QM];                      // the list has published its jpub-th best at least once
    uint32_t bcnt[TC_QM];                     // entries appended to the buffer since its last compaction (may overshoot 32)
    unsigned long long floor_key[TC_QM];      // threshold floor of the list (warm-up pass; atomicMax by the scanners)
    unsigned long long ring[TC_OWN_WARPS][TC_RING];   // candidates: tag (24) | query (8) | row (32); tag = lap + 1 marks a written slot
    uint32_t tau_key[TC_QM];                  // fkey of the approximate-score filter threshold of the query (atomicMax)
    uint32_t resv[TC_OWN_WARPS];              // ring positions handed out to producers (monotonic)
    uint32_t head[TC_OWN_WARPS];              // ring positions consumed by the owner (monotonic; producers wait on it when full)
    uint32_t scan_done;                       // scanner warps that finished pushing an item (monotonic)
    uint64_t full[TC_STAGES], empty[TC_STAGES], tfull[2], tempty[2];
    uint32_t tmem_ptr;
    unsigned item;
};

__global__ void normalise_rows_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[32];
    const float4 *src = reinterpret_cast<const float4 *>(x + i * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 t = src[c];
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
    normalise32(v);
    float4 *dst = reinterpret_cast<float4 *>(out + i * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// K-major, SWIZZLE_128B operand descriptor (128-byte rows, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;            // SBO
    d |= (uint64_t)1 << 46;                       // version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// order-preserving map float -> uint32 (and back); NaN never reaches it (scores map NaN to -inf)
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ bool key_gt(uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl) {
    return (ah > bh) || (ah == bh && al > bl);
}

// exact pinned-order score of query lane `ql` against row `r` of a swizzled stage
__device__ __forceinline__ float tc_exact_score(const float *qsm, const float *stage, int ql, int r) {
    float acc = 0.f;
    const float4 *qp = reinterpret_cast<const float4 *>(qsm + ql * 32);
    const float4 *dp = reinterpret_cast<const float4 *>(stage + r * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 a = qp[c ^ (ql & 7)];
        const float4 b = dp[c ^ (r & 7)];
        acc = __fadd_rn(acc, __fmul_rn(a.x, b.x));
        acc = __fadd_rn(acc, __fmul_rn(a.y, b.y));
        acc = __fadd_rn(acc, __fmul_rn(a.z, b.z));
        acc = __fadd_rn(acc, __fmul_rn(a.w, b.w));
    }
    return (acc != acc) ? -CUDART_INF_F : acc;
}

// exact pinned-order score of query row `ql` (swizzled smem tile) against a normalised DB row in global memory
__device__ __forceinline__ float tc_exact_score_g(const float *qsm, const float *__restrict__ drow, int ql) {
    float acc = 0.f;
    const float4 *qp = reinterpret_cast<const float4 *>(qsm + ql * 32);
    const float4 *dp = reinterpret_cast<const float4 *>(drow);
    float4 b[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) b[c] = __ldg(dp + c);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 a = qp[c ^ (ql & 7)];
        acc = __fadd_rn(acc, __fmul_rn(a.x, b[c].x));
        acc = __fadd_rn(acc, __fmul_rn(a.y, b[c].y));
        acc = __fadd_rn(acc, __fmul_rn(a.z, b[c].z));
        acc = __fadd_rn(acc, __fmul_rn(a.w, b[c].w));
    }
    return (acc != acc) ? -CUDART_INF_F : acc;
}

// Bounds shared between the lists of a query.  A query has one list per DB slice (S slices = S work items, usually on
// S different SMs at the same time), and every list starts cold: a row enters a list with probability ~ k / (rows the
// list has seen), so S cold lists re-score ~ S k ln(rows / (S k)) rows where one list over all rows would re-score
// k ln(rows / k).  What helps is a bound on the query's FINAL k-th best that uses all lists at once:
//   * every list publishes its j-th best exact score, j = ceil(k / min(S, k)), in slots[q][slice] (single writer);
//   * if m lists have published, m j >= k, there are k distinct rows scoring at least the SMALLEST of those m values, so
//     the m-th largest published value is a lower bound of the final k-th best.  S <= 8: j = ceil(k / S) and the
//     minimum over all S slots (tc_bound_small, one lane per list).  S > 8: the m-th largest of S values by bisection
//     on the key bits (tc_bound_large, one warp per list; for S >= k that is the k-th largest of the lists' BEST rows,
//     which is within a few ranks of the true k-th best of everything seen so far).
// The bound goes into gkey[q] (atomicMax; the scanners read it one tile ahead) and into the query's tau_key.
// A row scoring strictly below the bound cannot be in the query's top-k.
constexpr int TC_SLOT_SMALL = 8;

__device__ unsigned long long *g_tc_stats_dev = nullptr;
__device__ __forceinline__ void tc_apply_bound(TcSmem &sm, unsigned *__restrict__ gkey, int q0, int qi, unsigned b, float eps) {
    if (g_tc_stats_dev) atomicAdd(g_tc_stats_dev + 10, 1ull);
    if (b == 0u) return;                                     // not enough lists have published yet
    const unsigned old = atomicMax(gkey + q0 + qi, b);
    if (g_tc_stats_dev) { atomicAdd(g_tc_stats_dev + 11, 1ull); if (b > old) atomicAdd(g_tc_stats_dev + 12, 1ull); }
    atomicMax(&sm.tau_key[qi], fkey(fkey_inv(b) - eps));
}

// S > 8: the warp finds the m-th largest of the S slot values of query qi (16 bisection steps on the high key bits:
// the result is rounded DOWN to a multiple of 2^16 key units, still a valid bound)
__device__ __forceinline__ unsigned tc_bound_large(const unsigned *__restrict__ slots_q, int S, int m, int lane) {
    unsigned v[8];                                           // S <= 256
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = lane + 32 * i < S ? __ldcg(slots_q + lane + 32 * i) : 0u;
    unsigned lo = 0u;                                        // invariant: at least m values are >= lo
#pragma unroll 1
    for (int bit = 31; bit >= 16; --bit) {
        const unsigned mid = lo | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) c += v[i] >= mid ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= m) lo = mid;
    }
    return lo;
}

// Descending bitonic sort of one 64-bit key per lane, and the merge step of two descending runs.
__device__ __forceinline__ unsigned long long tc_cmpx(unsigned long long x, int stride, bool take_max) {
    const unsigned long long p = __shfl_xor_sync(0xffffffffu, x, stride);
    return (x > p) == take_max ? x : p;
}
__device__ __forceinline__ unsigned long long tc_sort32_desc(unsigned long long x, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
            x = tc_cmpx(x, stride, ((lane & stride) == 0) == ((lane & size) == 0));
    return x;
}
// top 32 of two descending runs (one key per lane each), descending
__device__ __forceinline__ unsigned long long tc_merge32_desc(unsigned long long y, unsigned long long x, int lane) {
    const unsigned long long xr = __shfl_sync(0xffffffffu, x, 31 - lane);
    unsigned long long z = y > xr ? y : xr;                  // bitonic: holds the 32 largest of the 64
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) z = tc_cmpx(z, stride, (lane & stride) == 0);
    return z;
}

// Compaction of query qc's list (whole warp): sort the append buffer, merge it with the sorted list, keep the best k,
// empty the buffer, raise the query's thresholds and publish the entries the other slices use (see above).
// ~200 instructions for up to 32 new entries, whatever their positions -- against ~60 per candidate for an insertion
// of one candidate at a time and ~300 per round for lane-serial shifting (the cost of this kernel is its instruction
// count: the owner warps get about one issue slot in ten).
__device__ __forceinline__ void tc_compact(TcSmem &sm, int qc, int k, float eps, int lane, unsigned *__restrict__ gkey,
                                           unsigned *__restrict__ slots, int S, int spad, int slice, int jpub, int q0) {
    unsigned long long *l = sm.lists + qc * TC_LSTRIDE;
    const int n = min((int)lds_volatile_u32(&sm.bcnt[qc]), 32);
    if (n == 0) return;                                      // warp-uniform
    const unsigned long long x = tc_sort32_desc(lane < n ? l[32 + lane] : 0ull, lane);
    const unsigned long long z = tc_merge32_desc(lane < k ? l[lane] : 0ull, x, lane);
    if (lane < k) l[lane] = z;
    const unsigned long long kth = __shfl_sync(0xffffffffu, z, k - 1), pj = __shfl_sync(0xffffffffu, z, jpub - 1);
    if (lane == 0) {
        sts_volatile_u32(&sm.bcnt[qc], 0u);
        const unsigned long long fl = lds_volatile_u64(&sm.floor_key[qc]);
        atomicMax(&sm.tau_key[qc], fkey(fkey_inv((uint32_t)((kth > fl ? kth : fl) >> 32)) - eps));
        if ((uint32_t)kth != 0u) atomicMax(gkey + q0 + qc, (unsigned)(kth >> 32));     // a real k-th entry: the list's own bound
        if (S > 1 && (uint32_t)pj != 0u) {
            __stcg(slots + (size_t)(q0 + qc) * spad + slice, (unsigned)(pj >> 32));
            sts_volatile_u32(&sm.pub[qc], 1u);
            if (g_tc_stats_dev) atomicAdd(g_tc_stats_dev + 13, 1ull);
        }
    }
    __syncwarp();
}

// Owner warp: up to 32 ring entries, ONE PER LANE: exact score (the row comes back from L2, where the TMA load of its
// tile just put it); what beats the list's current k-th entry, floor and shared bound is APPENDED to the list's buffer
// (one shared-memory atomic); a buffer that is full is compacted (tc_compact) and the lanes it turned away try again.
// Every list has exactly one owner warp, so list accesses need no locking.
__device__ __forceinline__ void tc_own_drain(TcSmem &sm, const float *__restrict__ rows, unsigned long long ent, bool have, int k,
                                             float eps, int lane, unsigned *__restrict__ gkey, unsigned *__restrict__ slots, int S,
                                             int spad, int slice, int jpub, int q0, unsigned long long *stats, uint32_t bcap) {
    const long long ts0 = stats ? clock64() : 0;
    int qi = 0;
    unsigned long long key = 0ull;
    bool todo = false;
    if (have) {
        const uint32_t row = (uint32_t)ent;
        qi = (int)((ent >> 32) & 0xffu);
        unsigned gbound = __ldcg(gkey + q0 + qi);                // in flight together with the row
        uint4 sa = make_uint4(~0u, ~0u, ~0u, ~0u), sb = sa;
        if (S > 1 && S <= TC_SLOT_SMALL) {
            sa = __ldcg(reinterpret_cast<const uint4 *>(slots + (size_t)(q0 + qi) * spad));
            sb = __ldcg(reinterpret_cast<const uint4 *>(slots + (size_t)(q0 + qi) * spad) + 1);
        }
        const float sc = tc_exact_score_g(sm.q, rows + (size_t)row * 32, qi);
        key = ((unsigned long long)fkey(sc) << 32) | (unsigned long long)(~row);
        if (S > 1 && S <= TC_SLOT_SMALL) {                       // S <= 8: the minimum over the S slots is a bound
            const unsigned v[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
            unsigned b = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < 8; ++i) b = i < S ? min(b, v[i]) : b;
            if (b > gbound) { tc_apply_bound(sm, gkey, q0, qi, b, eps); gbound = b; }
        }
        const unsigned long long kth = sm.lists[qi * TC_LSTRIDE + k - 1], fl = lds_volatile_u64(&sm.floor_key[qi]);
        todo = key > (kth > fl ? kth : fl) && fkey(sc) >= gbound;
    }
    if (stats) {
        const unsigned bl = __ballot_sync(0xffffffffu, todo);
        if (lane == 0) { atomicAdd(stats + 7, (unsigned long long)(clock64() - ts0)); atomicAdd(stats + 8, (unsigned long long)__popc(bl)); }
    }
    for (;;) {
        bool full = false, early = false;
        if (todo) {
            const uint32_t pos = atomicAdd(&sm.bcnt[qi], 1u);
            if (pos < (bcap & 0xffu)) {
                sm.lists[qi * TC_LSTRIDE + 32 + pos] = key;
                todo = false;
                // a young list is compacted as soon as it can publish: the bound the slices share (see above) is worth
                // most at the start of an item, when every list is cold
                early = S > 1 && pos + 1u >= (uint32_t)max(jpub, 4) && lds_volatile_u32(&sm.pub[qi]) == 0u;
            } else {
                full = true;
            }
        }
        __syncwarp();
        unsigned fm = __ballot_sync(0xffffffffu, full || early);
        if (!fm) break;
        while (fm) {                                             // every distinct full (or young) list once
            const int qc = __shfl_sync(0xffffffffu, qi, __ffs(fm) - 1);
            fm &= ~__ballot_sync(0xffffffffu, (full || early) && qi == qc);
            tc_compact(sm, qc, k, eps, lane, gkey, slots, S, spad, slice, jpub, q0);
        }
        if (!__any_sync(0xffffffffu, full)) break;
        // a lane that was turned away re-checks its key against the raised k-th entry before it tries again
        if (todo) todo = key > sm.lists[qi * TC_LSTRIDE + k - 1];
    }
}

// Persistent 1-D grid; work items (query tile, DB slice) are handed out by an atomic counter so that
// any (nq, n_db) shape fills all SMs.  qn = normalised queries (nq,32); tmap over the normalised DB rows.
//
// Roles.  A score tile is 128 TMEM lanes (queries) x 256 columns (DB rows).  TMEM read-out is fast (measured ~1 KB per
// cycle and SM, tools/tmem_probe.cu); what bounds this kernel is the dependent-instruction latency of the warps that
// look at the scores, so there are many of them and they do as little as possible:
//   16 scanner warps  lane quarter x 64-column chunk.  Per tile: load the chunk, max-reduce it in 16-column blocks,
//                     compare with the query's threshold; the (rare) passing columns are pushed as (row, query)
//                     entries into the ring of the query's owner (space reserved with one shared-memory atomic per
//                     lane that has hits).  The accumulator slot is released as soon as the chunk is in registers.
//   8 owner warps     pop up to 32 entries, re-score them exactly (one per lane) and insert them into the sorted
//                     top-k lists, ONE list per query and item (tc_own_drain); thresholds go back to the scanners
//                     through sm.tau_key.  At item end the owner writes its lists to the item's output slot.
// For <= 64 (<= 32) queries the query tile is replicated over the lane quarters (rep = 2, 4) and the replicas split
// each chunk's columns, so that all scanner warps stay busy.
__global__ void __launch_bounds__(TC_THREADS, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_db, int tiles_per_slice, int n_slices,
               const float *__restrict__ rows, const float *__restrict__ qn, int nq, int k, float eps, int rep,
               float *__restrict__ part_s, uint32_t *__restrict__ part_i, unsigned *__restrict__ work_counter,
               unsigned *__restrict__ gkey, unsigned *__restrict__ slots, float *dbg, unsigned long long *stats, uint32_t bcap, int ring_lim, int warm_max) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform: uniform role branches
    if (smem_u32(smem_raw) & 1023u) __trap();
    const int64_t n_tiles = (n_db + TC_ROWS - 1) / TC_ROWS;
    const unsigned long long NEG_KEY = (unsigned long long)0x007fffffu << 32;      // (fkey(-inf), row 0xffffffff)

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.tfull[s], 1); mbar_init(&sm.tempty[s], TC_SCAN_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
        sm.scan_done = 0u;
    }
    for (int i = tid; i < TC_OWN_WARPS * TC_RING; i += TC_THREADS) (&sm.ring[0][0])[i] = 0ull;     // tag 0 = never written
    if (tid < TC_OWN_WARPS) { sm.resv[tid] = 0u; sm.head[tid] = 0u; }
    if (warp == 1) { tmem_alloc(&sm.tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_ptr;
    // kind::tf32: a/b format 2 (TF32), fp32 accumulate, K-major both, M = 128, N = 256
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_ROWS >> 3) << 17) | ((uint32_t)(TC_QM >> 4) << 24);

    const int nqt = TC_QM / rep;                     // distinct queries per tile (128, 64 or 32)
    const int ncols = 64 / rep;                      // columns per scanner thread and tile (64, 32 or 16)
    const int n_qt = (nq + nqt - 1) / nqt;
    const unsigned n_items = (unsigned)n_qt * (unsigned)n_slices;
    int64_t itg = 0;                 // running tile counter (barrier phases continue across work items)
    uint32_t item_seq = 0;           // work items this CTA has finished
    uint32_t ring_h = 0;             // owner warps: consumed ring positions
    for (;;) {
        if (tid == 0) sm.item = atomicAdd(work_counter, 1u);
        __syncthreads();
        const unsigned item = sm.item;
        if (item >= n_items) break;
        // The slice varies fastest: all lists of a query tile run at the same time (on neighbouring SMs), so the bound
        // they share (tc_own_drain) is fed by every slice from the first tiles on; a slice is still streamed by
        // gridDim / n_slices CTAs at once, which is what gives the L2 reuse.
        const int qt = (int)(item / (unsigned)n_slices), slice = (int)(item % (unsigned)n_slices);
        const int64_t tile0 = (int64_t)slice * tiles_per_slice;
        int64_t my_tiles = n_tiles - tile0;
        if (my_tiles > tiles_per_slice) my_tiles = tiles_per_slice;
        if (my_tiles < 0) my_tiles = 0;
        const int q0 = qt * nqt;
        // Warm-up pass: the first T tiles of the item are scored TWICE.  The first time (virtual tiles 0 .. T-1) the
        // scanners only collect, per thread, the largest approximate score of each tile's chunk; from those group
        // maxima they derive a floor for the query's k-th best (below) before a single candidate is emitted.  The
        // second time round (virtual tiles T ...) is the normal scan from tile 0 on.  T tiles of MMA + scan buy a
        // floor at the ~k / (256 T) quantile instead of starting at -inf: without it the first tiles of an item
        // flood the owners with candidates, and while those are worked off the scanners run on stale thresholds.
        int T = 0;
        {
            int64_t full_tiles = n_db / TC_ROWS - tile0;            // complete tiles at the start of the item
            if (full_tiles > my_tiles) full_tiles = my_tiles;
            if (full_tiles >= 8 && warm_max > 0) {
                int64_t t = my_tiles / 8;
                t = t < 8 ? 8 : (t > warm_max ? warm_max : t);
                T = (int)(t > full_tiles ? full_tiles : t);
            }
        }
        const int64_t v_tiles = my_tiles + T;                       // virtual tiles of the item
        const uint64_t t_item = globaltimer_ns();
        // stage the query tile (rep copies), swizzled like a SWIZZLE_128B TMA load; rows beyond nq are zero
        for (int i = tid; i < TC_QM * 8; i += TC_THREADS) {
            const int row = i >> 3, c = i & 7;
            const int qi = row & (nqt - 1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q0 + qi < nq) v = *reinterpret_cast<const float4 *>(qn + (int64_t)(q0 + qi) * 32 + c * 4);
            *reinterpret_cast<float4 *>(sm.q + row * 32 + ((c ^ (row & 7)) << 2)) = v;
        }
        // The sorted top-k list of every query lives in shared memory as 64-bit keys (hi = order-preserving map of the
        // exact score, lo = ~row: one unsigned compare is exactly "score desc, index asc"; empty slots hold the
        // smallest key, (-inf, row 0xffffffff)).
        for (int i = tid; i < nqt * TC_LSTRIDE; i += TC_THREADS) sm.lists[i] = NEG_KEY;
        if (tid < TC_QM) { sm.floor_key[tid] = NEG_KEY; sm.tau_key[tid] = 0x007fffffu; sm.bcnt[tid] = 0u; sm.pub[tid] = 0u; sm.fmin_key[tid] = 0xffffffffu; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
        __syncthreads();

        if (warp == 0) {
            if (lane == 0) {
                for (int64_t v = 0; v < v_tiles; ++v) {
                    const int64_t g = itg + v, it = v < T ? v : v - T;
                    const int s = (int)(g % TC_STAGES);
                    mbar_wait_relaxed(&sm.empty[s], (uint32_t)(((g / TC_STAGES) & 1) ^ 1));
                    mbar_expect_tx(&sm.full[s], TC_ROWS * 128);
                    tma_load_2d(sm.stage[s], &tmap, 0, (int)((tile0 + it) * TC_ROWS), &sm.full[s]);
                }
            }
        } else if (warp == 1) {
            const uint32_t leader = elect_one() ? 1u : 0u;
            const uint32_t qaddr = smem_u32(sm.q);
            for (int64_t v = 0; v < v_tiles; ++v) {
                const int64_t g = itg + v;
                const int s = (int)(g % TC_STAGES);
                const uint32_t slot = (uint32_t)(g & 1);
                mbar_wait_relaxed(&sm.full[s], (uint32_t)((g / TC_STAGES) & 1));
                mbar_wait_relaxed(&sm.tempty[slot], (uint32_t)(((g >> 1) & 1) ^ 1));
                tc_fence_after();
                if (leader) {
                    const uint32_t baddr = smem_u32(sm.stage[s]);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_tf32(tmem_base + slot * 256u, umma_desc_sw128(qaddr + kk * 32), umma_desc_sw128(baddr + kk * 32),
                                    idesc, kk ? 1u : 0u);
                    tc_commit(&sm.tfull[slot]);
                    tc_commit(&sm.empty[s]);       // the stage is only the MMA's B operand: candidates are re-read from L2
                }
                __syncwarp();
            }
        } else if (warp < TC_OWN_WARP0) {
            // ================= scanners =================
            const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
            const int cidx = (warp - TC_SCAN_WARP0) >> 2;       // 64-column chunk of the tile
            const int ql = quarter * 32 + lane;                 // TMEM lane = row of the (replicated) query tile
            const int qi = ql & (nqt - 1);                      // query within the tile
            const int col0 = 64 * cidx + ncols * (ql / nqt);    // the replicas of a query split the chunk
            const bool qvalid = q0 + qi < nq;
            const bool warp_valid = __any_sync(0xffffffffu, qvalid);
            const int own = qi & (TC_OWN_WARPS - 1);
            // the best bound anybody on this GPU has published for this query: an L2 read (1000+ cycles while the TMA stream
            // keeps the L2 busy), so it is issued one tile ahead and never waited for
            unsigned gk_next = qvalid ? __ldcg(gkey + q0 + qi) : 0u;
            if (T > 0) {
                // ---- pass A (virtual tiles 0 .. T-1): group maxima only, then the floor ----
                float top[8];                                    // this thread's largest per-tile maxima, descending
#pragma unroll
                for (int j = 0; j < 8; ++j) top[j] = -CUDART_INF_F;
                for (int64_t v = 0; v < T; ++v) {
                    const int64_t g = itg + v;
                    const uint32_t slot = (uint32_t)(g & 1);
                    mbar_wait_relaxed(&sm.tfull[slot], (uint32_t)((g >> 1) & 1));
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + slot * 256u + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0;
                    // Pass A.  A query's scores of one tile sit in P = 4 rep scanner threads (ncols columns each).  Every
                    // thread keeps the 8 largest of its per-tile maxima; after T tiles its jfl-th largest, jfl =
                    // ceil(k / P), has jfl DISTINCT rows at or above it, so the minimum `lo` over the P threads has
                    // P jfl >= k rows at or above it, whose exact scores are >= lo - eps: the final k-th best of the
                    // list is >= lo - eps.  (All tiles of the pass are complete: no zero-filled rows are counted.)
                    if (warp_valid) {
                        float gm = -CUDART_INF_F;
#pragma unroll 1
                        for (int u = 0; u < ncols; u += 32) {
                            uint32_t vv[32];
                            tmem_ld16_issue(taddr + (uint32_t)u, vv);
                            if (ncols >= 32) tmem_ld16_issue(taddr + (uint32_t)(u + 16), vv + 16);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) gm = fmaxf(gm, __uint_as_float(vv[j]));      // fmaxf drops NaNs (zero rows)
                            if (ncols >= 32) {
#pragma unroll
                                for (int j = 16; j < 32; ++j) gm = fmaxf(gm, __uint_as_float(vv[j]));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {                // insertion into the descending array, branch-free
                            const float hi = fmaxf(top[j], gm), lo2 = fminf(top[j], gm);
                            top[j] = hi;
                            gm = lo2;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.tempty[slot]);
                    if (v == T - 1) {
                        const int jfl = (k + 4 * rep - 1) / (4 * rep);
                        float mine = top[0];
#pragma unroll
                        for (int j = 1; j < 8; ++j) mine = j == jfl - 1 ? top[j] : mine;
                        if (jfl > T) mine = -CUDART_INF_F;           // fewer groups than needed (cannot happen: T >= 8 >= jfl)
                        if (qvalid) atomicMin(&sm.fmin_key[qi], mine > -CUDART_INF_F ? fkey(mine) : 0u);   // 0: no floor
                        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_SCAN_WARPS) : "memory");
                        const uint32_t fm = sm.fmin_key[qi];
                        if (qvalid && fm != 0u && fm != 0xffffffffu && (ql / nqt) == 0 && cidx == 0) {   // one thread per query
                            const float floor_s = fkey_inv(fm) - eps;
                            atomicMax(&sm.floor_key[qi], (unsigned long long)fkey(floor_s) << 32);   // ties with the floor itself still enter
                            atomicMax(&sm.tau_key[qi], fkey(floor_s - eps));
                        }
                        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_SCAN_WARPS) : "memory");     // floors visible before pass B
                    }
                }
            }
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int64_t g = itg + T + it;
                const uint32_t slot = (uint32_t)(g & 1);
                const unsigned gk = gk_next;
                if (qvalid) gk_next = __ldcg(gkey + q0 + qi);
                const long long tw0 = stats ? clock64() : 0;
                mbar_wait_relaxed(&sm.tfull[slot], (uint32_t)((g >> 1) & 1));
                const long long tsc0 = stats ? clock64() : 0;
                if (stats && lane == 0) atomicAdd(stats + 4, (unsigned long long)(tsc0 - tw0));
                tc_fence_after();
                const uint32_t taddr = tmem_base + slot * 256u + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0;
                const int64_t row0 = (tile0 + it) * TC_ROWS;
                uint32_t hit0 = 0u, hit1 = 0u;
                if (warp_valid) {
                    const float tau = (qvalid && !(bcap & 0x100u)) ? fmaxf(fkey_inv(lds_volatile_u32(&sm.tau_key[qi])), fkey_inv(gk) - eps) : CUDART_INF_F;   // bcap bit 8: diagnosis, nothing passes
                    // Scan 32 columns per TMEM round trip (16 when the chunk is split four ways).  NaN approximations
                    // (zero DB rows) need no special case: they can only matter while tau is -inf, and then
                    // !(x < tau) is true for every x.
#pragma unroll 1
                    for (int u = 0; u < ncols; u += 32) {
                        uint32_t v[32];
                        tmem_ld16_issue(taddr + (uint32_t)u, v);
                        if (ncols >= 32) tmem_ld16_issue(taddr + (uint32_t)(u + 16), v + 16);
                        tmem_ld_wait();
                        if (dbg && slice == 0 && qt == 0 && it == 0) {      // debug: dump the approximate scores of tile 0 (pass B)
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) dbg[qi * TC_ROWS + col0 + u + j] = __uint_as_float(v[j]);
                        }
                        // 16-column blocks: max-reduce, and only where some lane of the warp has a block maximum that
                        // passes (one vote per block) are the block's 16 columns compared one by one
                        float m0 = __uint_as_float(v[0]), m1 = -CUDART_INF_F;
#pragma unroll
                        for (int j = 1; j < 16; ++j) m0 = fmaxf(m0, __uint_as_float(v[j]));
                        if (ncols >= 32) {
                            m1 = __uint_as_float(v[16]);
#pragma unroll
                            for (int j = 17; j < 32; ++j) m1 = fmaxf(m1, __uint_as_float(v[j]));
                        }
                        uint32_t mask = 0u;
                        if (__any_sync(0xffffffffu, qvalid && !(m0 < tau))) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) mask |= !(__uint_as_float(v[j]) < tau) ? (1u << j) : 0u;
                        }
                        if (ncols >= 32 && __any_sync(0xffffffffu, qvalid && !(m1 < tau))) {
#pragma unroll
                            for (int j = 16; j < 32; ++j) mask |= !(__uint_as_float(v[j]) < tau) ? (1u << j) : 0u;
                        }
                        if (!qvalid) mask = 0u;
                        if (u == 0) hit0 = mask; else hit1 = mask;
                    }
                }
                // the accumulator slot is free again as soon as every scanner warp has its chunk in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tempty[slot]);
                if (stats && lane == 0) atomicAdd(stats + 9, (unsigned long long)(clock64() - tsc0));
                // push what passed: rows beyond the DB (zero-filled by TMA) are dropped here
                if (row0 + TC_ROWS > n_db) {
                    const int64_t l64 = n_db - row0 - col0;       // valid columns of this thread's range
                    const int left = l64 < 0 ? 0 : (l64 > 64 ? 64 : (int)l64);
                    if (left < 32) hit0 &= (1u << left) - 1u;
                    if (left < 64) hit1 &= left <= 32 ? 0u : ((1u << (left - 32)) - 1u);
                }
                const int cnt = __popc(hit0) + __popc(hit1);
                const long long tp0 = stats ? clock64() : 0;
                if (__any_sync(0xffffffffu, cnt != 0)) {
                    // Reserve ring space with one atomic per lane that has hits, then write once the whole range fits.
                    // The wait is a CONVERGED loop (one shared-memory load per warp and poll): lanes that spin on their
                    // own flood the LSU and starve the owners.  A lane whose range fits writes at once, whatever the
                    // other lanes of the warp are waiting for -- the lane at a ring's head always fits, so the owners
                    // always make progress.
                    uint32_t pos = 0u;
                    if (cnt) pos = atomicAdd(&sm.resv[own], (uint32_t)cnt);
                    const uint32_t need = pos + (uint32_t)cnt;
                    bool pending = cnt != 0;
                    for (;;) {
                        if (pending && (int32_t)(need - lds_volatile_u32(&sm.head[own])) <= ring_lim) {
                            const unsigned long long qbits = (unsigned long long)qi << 32;
                            const uint32_t rbase = (uint32_t)(row0 + col0);
#pragma unroll 1
                            for (int half = 0; half < 2; ++half) {
                                uint32_t m = half ? hit1 : hit0;
                                while (m) {
                                    const int j = __ffs(m) - 1;
                                    m &= m - 1;
                                    const unsigned long long tag = (unsigned long long)(((pos / TC_RING) + 1u) & 0xffffffu) << 40;
                                    sts_volatile_u64(&sm.ring[own][pos & (TC_RING - 1)], tag | qbits | (unsigned long long)(rbase + 32 * half + j));
                                    ++pos;
                                }
                            }
                            pending = false;
                        }
                        if (!__any_sync(0xffffffffu, pending)) break;
                        __nanosleep(100);
                        if (globaltimer_ns() - t_item > 4000000000ull) {      // bounded wait: a protocol bug must not hang the GPU
                            if (lane == 0) printf("asr: pre-filter ring wait timeout, block %d warp %d\n", blockIdx.x, warp);
                            __trap();
                        }
                    }
                }
                __syncwarp();
                if (stats && lane == 0) atomicAdd(stats + 3, (unsigned long long)(clock64() - tp0));
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd(&sm.scan_done, 1u);
        } else {
            // ================= owners =================
            const int o = warp - TC_OWN_WARP0;
            const uint32_t done_target = (uint32_t)TC_SCAN_WARPS * (item_seq + 1u);
            const int S = n_slices, spad = (n_slices + 7) & ~7;
            const int jpub = (k + min(S, k) - 1) / min(S, k);          // which entry of its list every slice publishes
            const int m_need = (k + jpub - 1) / jpub;                  // published lists needed for a bound
            uint32_t drains = 0u, next_refresh = 1u;
            bool flush = false;
            for (;;) {
                // Cheap poll by ONE lane: how much has been reserved beyond what this warp has consumed, and whether the
                // scanners are done with the item.  Partial batches wait (a drain costs thousands of cycles whatever its
                // size), with exponential back-off: polling used to be a fifth of the SM's issued instructions.
                int st8;
                for (uint32_t idle_ns = 100u;;) {
                    int st = 0;
                    if (lane == 0) {
                        const uint32_t rv = lds_volatile_u32(&sm.resv[o]);
                        const uint32_t sd = lds_volatile_u32(&sm.scan_done);
                        st = ((uint32_t)(rv - ring_h) >= 32u ? 4 : 0) | (sd == done_target ? 1 : 0) | (rv == ring_h ? 2 : 0);
                    }
                    st8 = __shfl_sync(0xffffffffu, st, 0);
                    if (st8 & 5) break;                                // a full batch, or the item is over
                    __nanosleep(idle_ns);
                    idle_ns = min(idle_ns * 2u, 4000u);
                    if (idle_ns == 4000u && globaltimer_ns() - t_item > 4000000000ull) {       // bounded wait (see mbar_wait)
                        if (lane == 0) printf("asr: pre-filter owner timeout, block %d warp %d\n", blockIdx.x, warp);
                        __trap();
                    }
                }
                if ((st8 & 3) == 3) break;                             // scanners done, everything consumed
                const uint32_t pos = ring_h + (uint32_t)lane;
                const unsigned long long ent = lds_volatile_u64(&sm.ring[o][pos & (TC_RING - 1)]);
                const bool valid = (uint32_t)(ent >> 40) == (((pos / TC_RING) + 1u) & 0xffffffu);
                const unsigned bal = __ballot_sync(0xffffffffu, valid);
                const int n = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;
                if (n == 0 || (n < 32 && !(st8 & 1)) || (n < 32 && !flush)) {   // reserved but not written yet / last entries: look again
                    flush = (st8 & 1) != 0;
                    __nanosleep(100);
                    continue;
                }
                const long long td0 = stats ? clock64() : 0;
                tc_own_drain(sm, rows, ent, lane < n, k, eps, lane, gkey, slots, S, spad, slice, jpub, q0, stats, bcap);
                if (stats && lane == 0) {
                    atomicAdd(stats + 0, (unsigned long long)n);
                    atomicAdd(stats + 1, 1ull);
                    atomicAdd(stats + 2, (unsigned long long)(clock64() - td0));
                }
                ring_h += (uint32_t)n;
                if (lane == 0) sts_volatile_u32(&sm.head[o], ring_h);
                // S > 8: after drains 1, 2, 3, 5, 8, 12, ... of the item, recompute the bound of every list of this owner from
                // what all slices have published so far (the bound improves with the logarithm of the rows seen)
                if (S > TC_SLOT_SMALL && ++drains == next_refresh) {
                    next_refresh += (next_refresh + 1u) / 2u;            // drains 1, 2, 3, 5, 8, 12, 18, 27, ...
                    for (int qi = o; qi < nqt; qi += TC_OWN_WARPS) {
                        if (q0 + qi >= nq) break;
                        const unsigned b = tc_bound_large(slots + (size_t)(q0 + qi) * spad, S, m_need, lane);
                        if (lane == 0) tc_apply_bound(sm, gkey, q0, qi, b, eps);
                    }
                    __syncwarp();
                }
            }
            // item end: fold the remaining buffers in; this owner's lists go straight to the item's output slot
            __syncwarp();
            for (int qi = o; qi < nqt; qi += TC_OWN_WARPS) {
                tc_compact(sm, qi, k, eps, lane, gkey, slots, S, spad, slice, jpub, q0);
                if (q0 + qi < nq && lane < k) {
                    const unsigned long long mine = sm.lists[qi * TC_LSTRIDE + lane];
                    const size_t oo = ((size_t)(q0 + qi) * n_slices + slice) * k + lane;
                    const bool real = mine != NEG_KEY;
                    part_s[oo] = real ? fkey_inv((uint32_t)(mine >> 32)) : -CUDART_INF_F;
                    part_i[oo] = real ? ~(uint32_t)mine : 0xffffffffu;
                }
            }
        }
        if (stats && tid == 0) { atomicAdd(stats + 5, (unsigned long long)v_tiles); atomicAdd(stats + 6, 1ull); }
        itg += v_tiles;
        ++item_seq;
        __syncthreads();     // every role is done with sm.q / the lists before the next query tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- vote -------------------------------------------------------------------------
constexpr int VOTE_MAX = 8192;
constexpr int VOTE_THREADS = 256;

template <typename T, bool DESC>
__device__ void bitonic_sort_smem(T *a, int n, int tid, int nthreads) {   // n power of two
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < n / 2; i += nthreads) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                T x = a[lo], y = a[hi];
                bool sw = DESC ? (up ? (x < y) : (x > y)) : (up ? (x > y) : (x < y));
                if (sw) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

// one CTA per recording
__global__ void __launch_bounds__(VOTE_THREADS)
vote_kernel(const int64_t *__restrict__ cand, const int32_t *__restrict__ row_ids, int64_t n_rows, int m, int m_pow2,
            int top_k, int32_t *__restrict__ out_ids, int32_t *__restrict__ out_counts) {
    extern __shared__ unsigned long long vsm[];     // m_pow2 entries
    const int tid = threadIdx.x;
    const int rec = blockIdx.x;
    // keys = piece id (ascending sort); empty = ~0
    for (int i = tid; i < m_pow2; i += VOTE_THREADS) {
        unsigned long long key = ~0ull;
        if (i < m) {
            int64_t r = cand[(size_t)rec * m + i];
            if (r >= 0 && r < n_rows) key = (unsigned long long)(uint32_t)row_ids[r];
        }
        vsm[i] = key;
    }
    bitonic_sort_smem<unsigned long long, false>(vsm, m_pow2, tid, VOTE_THREADS);
    // run lengths: at each run start, count = (next run start) - i.  Two passes through registers.
    unsigned long long newkey[VOTE_MAX / VOTE_THREADS];
    int nn = 0;
    for (int i = tid; i < m_pow2; i += VOTE_THREADS, ++nn) {
        unsigned long long key = vsm[i];
        unsigned long long out = 0ull;
        if (key != ~0ull && (i == 0 || vsm[i - 1] != key)) {
            int j = i + 1;
            while (j < m_pow2 && vsm[j] == key) ++j;
            out = ((unsigned long long)(j - i) << 32) | (key & 0xffffffffull);
        }
        newkey[nn] = out;
    }
    __syncthreads();
    nn = 0;
    for (int i = tid; i < m_pow2; i += VOTE_THREADS, ++nn) vsm[i] = newkey[nn];
    bitonic_sort_smem<unsigned long long, true>(vsm, m_pow2, tid, VOTE_THREADS);
    for (int j = tid; j < top_k; j += VOTE_THREADS) {
        unsigned long long key = j < m_pow2 ? vsm[j] : 0ull;
        bool v = key != 0ull;
        out_ids[(size_t)rec * top_k + j] = v ? (int32_t)(key & 0xffffffffull) : -1;
        out_counts[(size_t)rec * top_k + j] = v ? (int32_t)(key >> 32) : 0;
    }
}

// ---- host side --------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

}  // namespace asr

struct asr_db {
    const float *codes;
    int64_t n;
    int64_t idx_base;
    int device;             // the handle is bound to the device that was current at create
    int flags;
    CUtensorMap tmap;       // over the caller's rows as they are
    void *scratch;          // TK_SCRATCH_BYTES: per-slice partial lists + the work counter of the pre-filter
    int sms;
    // cosine queries stream rows normalised with the pinned definition: an owned copy (default), the caller's own
    // rows normalised in place (ASR_DB_NORMALISE_IN_PLACE), or nothing (ASR_DB_NO_COSINE_COPY: rows are normalised
    // in-kernel by the exact path).  Built at create; published only when complete.
    const float *rows_n = nullptr;
    float *codes_n = nullptr;
    CUtensorMap tmap_n;
    float *qn = nullptr;    // normalised queries of the current call (pre-filter path), qn_cap rows, sized at create
    int64_t qn_cap = 0;
    unsigned *slots = nullptr;   // pre-filter path: per (query, DB slice) published list entries (see tc_own_drain), slots_cap words
    int64_t slots_cap = 0;
};

using namespace asr;

extern "C" {

static int encode_rows_map(CUtensorMap *tm, const float *rows, int64_t n, int box_rows) {
    cuuint64_t gdim[2] = {32, (cuuint64_t)n};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(rows), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return ASR_ERR_CUDA;
    }
    return ASR_OK;
}

int asr_db_create_ex(asr_db_t **out, const float *codes_dev, int64_t n, int64_t idx_base, int flags, int64_t max_queries) {
    ASR_CHECK_ARG(out != nullptr, "out is NULL");
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(codes_dev != nullptr && n > 0, "empty database");
    ASR_CHECK_ARG((reinterpret_cast<uintptr_t>(codes_dev) & 127) == 0, "codes_dev must be 128-byte aligned");
    ASR_CHECK_ARG(n < ((int64_t)1 << 31), "at most 2^31-1 rows per shard");
    ASR_CHECK_ARG(!((flags & ASR_DB_NORMALISE_IN_PLACE) && (flags & ASR_DB_NO_COSINE_COPY)), "contradictory flags");
    ASR_CHECK_ARG(max_queries >= 0, "max_queries < 0");
    if (!get_encode_fn()) { set_error("cuTensorMapEncodeTiled not available from the driver"); return ASR_ERR_CUDA; }
    static_assert(TK_ROWS == TC_ROWS, "one tensor-map box for both kernels");
    asr_db *db = new asr_db();
    db->codes = codes_dev;
    db->n = n;
    db->idx_base = idx_base;
    db->sms = sm_count();
    db->device = current_device();
    db->flags = flags;
    db->scratch = nullptr;
#define DB_FAIL(code)                 \
    do {                              \
        asr_db_destroy(db);           \
        return (code);                \
    } while (0)
    if ((rc = encode_rows_map(&db->tmap, codes_dev, n, TK_ROWS))) DB_FAIL(rc);
    if (cudaMalloc(&db->scratch, TK_SCRATCH_BYTES) != cudaSuccess) {
        set_error("cudaMalloc of the top-k scratch failed");
        DB_FAIL(ASR_ERR_CUDA);
    }
    static bool attr_done[ASR_MAX_DEVICES] = {false};     // cudaFuncSetAttribute is per device
    const int attr_dev = std::max(0, std::min(db->device, ASR_MAX_DEVICES - 1));
    if (!attr_done[attr_dev]) {
        cudaError_t e1 = cudaFuncSetAttribute(topk_stream_kernel<1, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sizeof(TkSmem<1, 2>) + 1024);
        cudaError_t e2 = cudaFuncSetAttribute(topk_stream_kernel<4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sizeof(TkSmem<4, 2>) + 1024);
        cudaError_t e3 = cudaFuncSetAttribute(topk_stream_kernel<16, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sizeof(TkSmem<16, 3>) + 1024);
        cudaError_t e4 = cudaFuncSetAttribute(rank_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sizeof(RkSmem) + 1024);
        cudaError_t e5 = cudaFuncSetAttribute(topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem));
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess) {
            set_error("asr_db_create: cudaFuncSetAttribute failed");
            DB_FAIL(ASR_ERR_CUDA);
        }
        attr_done[attr_dev] = true;
    }
    // rows normalised with the pinned definition (the same function the kernels apply per row, so results are
    // bit-identical to normalising in-kernel): built here, synchronously, and published only when complete
    if (!(flags & ASR_DB_NO_COSINE_COPY)) {
        float *dst = nullptr;
        if (flags & ASR_DB_NORMALISE_IN_PLACE) {
            dst = const_cast<float *>(codes_dev);
        } else if (cudaMalloc(&db->codes_n, (size_t)n * 128) != cudaSuccess) {
            cudaGetLastError();
            set_error("asr_db_create: cudaMalloc of the normalised copy (" + std::to_string((size_t)n * 128) +
                      " bytes) failed; ASR_DB_NORMALISE_IN_PLACE avoids the copy");
            DB_FAIL(ASR_ERR_CUDA);
        } else {
            dst = db->codes_n;
        }
        normalise_rows_kernel<<<(unsigned)((n + 255) / 256), 256>>>(codes_dev, n, dst);
        count_launch();
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            set_error("asr_db_create: normalising the rows failed");
            DB_FAIL(ASR_ERR_CUDA);
        }
        if ((rc = encode_rows_map(&db->tmap_n, dst, n, TC_ROWS))) DB_FAIL(rc);
        db->rows_n = dst;
    }
    db->qn_cap = std::max<int64_t>(max_queries, TC_QM);
    if (cudaMalloc(&db->qn, (size_t)db->qn_cap * 128) != cudaSuccess) {
        set_error("asr_db_create: cudaMalloc of the query workspace failed");
        DB_FAIL(ASR_ERR_CUDA);
    }
    db->slots_cap = db->qn_cap * 8 + 262144;
    if (cudaMalloc(&db->slots, (size_t)db->slots_cap * 4) != cudaSuccess) {
        set_error("asr_db_create: cudaMalloc of the bound-sharing workspace failed");
        DB_FAIL(ASR_ERR_CUDA);
    }
#undef DB_FAIL
    *out = db;
    return ASR_OK;
}

int asr_db_create(asr_db_t **out, const float *codes_dev, int64_t n, int64_t idx_base) {
    return asr_db_create_ex(out, codes_dev, n, idx_base, 0, 16384);
}

int asr_db_destroy(asr_db_t *db) {
    if (!db) return ASR_OK;
    cudaFree(db->scratch);
    cudaFree(db->codes_n);
    cudaFree(db->qn);
    cudaFree(db->slots);
    delete db;
    return ASR_OK;
}

// qt = queries per pass of the chosen kernel, ctas_per_sm = its occupancy
static void plan_grid(const asr_db *db, int64_t nq, int qt, int ctas_per_sm, int *n_slices, int *tiles_per_slice,
                      int *n_qgroups) {
    const int64_t n_tiles = (db->n + TK_ROWS - 1) / TK_ROWS;
    const int64_t n_qt = (nq + qt - 1) / qt;
    const int64_t slots = (int64_t)db->sms * ctas_per_sm;
    // few query tiles: slice the DB over all CTA slots.  Many query tiles: keep slices long
    // (>= 64k rows) so per-slice list maintenance stays negligible and spread queries instead.
    int64_t want = slots;
    if (n_qt >= 2 * slots) want = std::max<int64_t>(1, std::min<int64_t>(slots, db->n / 65536));
    else if (n_qt > 1) want = std::max<int64_t>(1, slots / n_qt);
    int64_t tps = (n_tiles + want - 1) / want;
    if (tps < 1) tps = 1;
    int64_t ns = (n_tiles + tps - 1) / tps;
    int64_t qg = std::min<int64_t>(n_qt, std::max<int64_t>(1, (2 * slots + ns - 1) / ns));
    *n_slices = (int)ns;
    *tiles_per_slice = (int)tps;
    *n_qgroups = (int)qg;
}

static float *g_tc_dbg = nullptr;   // set by asr_debug_tc_scores
static unsigned long long *g_tc_stats = nullptr;   // ASR_TC_STATS=1: device counters of the pre-filter kernel (diagnosis only)

static int topk_tc(asr_db *db, const float *q_dev, int64_t nq, int k, float *out_score_dev, int64_t *out_idx_dev,
                   cudaStream_t st) {
    if (nq > db->qn_cap) {      // more queries than the workspace sized at create: pass over them in chunks (no allocation here)
        for (int64_t q0 = 0; q0 < nq; q0 += db->qn_cap) {
            const int64_t nqc = std::min<int64_t>(db->qn_cap, nq - q0);
            int rcc = topk_tc(db, q_dev + q0 * 32, nqc, k, out_score_dev + q0 * k, out_idx_dev + q0 * k, st);
            if (rcc) return rcc;
        }
        return ASR_OK;
    }
    normalise_rows_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(q_dev, nq, db->qn);
    ASR_LAUNCH_CHECK();
    if (!g_tc_stats && getenv("ASR_TC_STATS")) {
        ASR_CUDA(cudaMalloc(&g_tc_stats, 128));
        ASR_CUDA(cudaMemset(g_tc_stats, 0, 128));
        ASR_CUDA(cudaMemcpyToSymbol(g_tc_stats_dev, &g_tc_stats, sizeof(g_tc_stats)));
    }
    // Replicate few queries over the TMEM lane quarters (the replicas split the columns) so that all scanner warps
    // have work: <= 32 queries four times, <= 64 twice.
    int rep = nq <= 32 ? 4 : (nq <= 64 ? 2 : 1);
    if (getenv("ASR_TC_REP")) { const int r = atoi(getenv("ASR_TC_REP")); rep = r >= 4 ? 4 : (r >= 2 ? 2 : 1); }
    const int nqt = TC_QM / rep, L = 1;           // one list per (query, slice)
    const int64_t n_tiles = (db->n + TC_ROWS - 1) / TC_ROWS;
    const int64_t n_qt_all = (nq + nqt - 1) / nqt;
    // Slices: the work items (query tile, slice) are handed out dynamically, but they are all about equally long, so the
    // kernel runs in rounds of `sms` items.  Choose the slice count S that minimises rounds x (tiles per slice + the
    // fixed cost of an item).  Measured, 10k queries x 1e6 rows (79 query tiles, 3907 DB tiles), S = 2 / 3 / 4 / 5 / 7:
    // 5.52 / 3.89 / 4.54 / 3.75 / 3.81 ms = rounds x (tiles per slice x 1.33 us + 0.22 ms): an item costs as much as
    // 165 tiles on top of its own (warm-up pass + list warm-up).
    const int min_tiles = getenv("ASR_TC_MIN_TILES") ? atoi(getenv("ASR_TC_MIN_TILES")) : 24;
    const int64_t s_max = std::max<int64_t>(1, std::min<int64_t>(256, n_tiles / min_tiles));   // <= 256: tc_bound_large
    int64_t best_s = 1;
    double best_cost = 1e300;
    for (int64_t sc = 1; sc <= s_max; ++sc) {
        const int64_t tps_c = (n_tiles + sc - 1) / sc, ns_c = (n_tiles + tps_c - 1) / tps_c;
        const int64_t rounds = (n_qt_all * ns_c + db->sms - 1) / db->sms;
        if (ns_c > TC_SLOT_SMALL && rounds > 1) continue;     // more than 8 slices share their bound only while ALL lists of a query run at once
        const double cost = (double)rounds * ((double)tps_c + 165.0);
        if (cost < best_cost * 0.995) { best_cost = cost; best_s = sc; }
    }
    if (getenv("ASR_TC_SLICES")) best_s = std::max<int64_t>(1, std::min<int64_t>(s_max, atoi(getenv("ASR_TC_SLICES"))));
    const int tps = (int)((n_tiles + best_s - 1) / best_s);
    const int n_slices = (int)((n_tiles + tps - 1) / tps);
    const int64_t per_q = (int64_t)n_slices * L * k * 8;
    const int spad = (n_slices + 7) & ~7;
    unsigned *slots = db->slots;
    int64_t q_chunk = std::max<int64_t>(nqt, (int64_t)((TK_SCRATCH_BYTES - 256 - (db->qn_cap + 64) * 4) / per_q) / nqt * nqt);
    q_chunk = std::max<int64_t>(nqt, std::min<int64_t>(q_chunk, db->slots_cap / spad / nqt * nqt));
    ASR_CHECK_ARG((int64_t)nqt * spad <= db->slots_cap, "bound-sharing workspace too small");
    unsigned *counter = reinterpret_cast<unsigned *>(reinterpret_cast<uint8_t *>(db->scratch) + TK_SCRATCH_BYTES - 256);
    // per-query bound shared by all lists of the query (see tc_drain): the last qn_cap words of the scratch before the counter
    unsigned *gkey = counter - ((db->qn_cap + 63) / 64 * 64);
    ASR_CUDA(cudaMemsetAsync(gkey, 0, (size_t)nq * sizeof(unsigned), st));        // 0 = below fkey(-inf): no bound yet
    for (int64_t q0 = 0; q0 < nq; q0 += q_chunk) {
        const int64_t nqc = std::min<int64_t>(q_chunk, nq - q0);
        const int64_t n_qt = (nqc + nqt - 1) / nqt;
        float *ps = reinterpret_cast<float *>(db->scratch);
        uint32_t *pi = reinterpret_cast<uint32_t *>(ps + (size_t)nqc * n_slices * L * k);
        ASR_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
        ASR_CUDA(cudaMemsetAsync(slots, 0, (size_t)(n_qt * nqt) * spad * 4, st));      // nothing published yet
        const int grid = (int)std::min<int64_t>(db->sms, n_qt * n_slices);
        topk_tc_kernel<<<grid, TC_THREADS, sizeof(TcSmem), st>>>(db->tmap_n, db->n, tps, n_slices, db->rows_n, db->qn + q0 * 32,
                                                                 (int)nqc, k, 0.00390625f, rep, ps, pi, counter, gkey + q0, slots, g_tc_dbg, g_tc_stats,
                                                                 (uint32_t)std::min(32, std::max(1, getenv("ASR_TC_BCAP") ? atoi(getenv("ASR_TC_BCAP")) : 32)) | (getenv("ASR_TC_NOHITS") ? 0x100u : 0u),
                                                                 std::min(TC_RING, std::max(64, getenv("ASR_TC_RING") ? atoi(getenv("ASR_TC_RING")) : TC_RING)),
                                                                 getenv("ASR_TC_WARM") ? atoi(getenv("ASR_TC_WARM")) : 32);
        ASR_LAUNCH_CHECK();
        launch_merge(ps, pi, nullptr, db->idx_base, n_slices * L, k, out_score_dev + q0 * k, out_idx_dev + q0 * k, nqc, st);
        ASR_LAUNCH_CHECK();
        if (g_tc_stats) {
            unsigned long long h[16];
            ASR_CUDA(cudaStreamSynchronize(st));
            ASR_CUDA(cudaMemcpy(h, g_tc_stats, sizeof(h), cudaMemcpyDeviceToHost));
            ASR_CUDA(cudaMemset(g_tc_stats, 0, sizeof(h)));
            fprintf(stderr, "[asr] tc stats: nq %lld rows %lld slices %d tps %d rep %d | items %llu tiles %llu | candidates %llu drains %llu "
                    "(%.1f per drain, %.0f cycles per drain: scoring %.0f, then %.1f insertions) | scanner cycles per tile and warp: wait-for-tile %.0f, scan %.0f, push %.0f | bounds: %llu computed, %llu non-zero, %llu raised gkey; %llu publications\n",
                    (long long)nqc, (long long)db->n, n_slices, tps, rep, h[6], h[5], h[0], h[1], h[1] ? (double)h[0] / h[1] : 0.0,
                    h[1] ? (double)h[2] / h[1] : 0.0, h[1] ? (double)h[7] / h[1] : 0.0, h[1] ? (double)h[8] / h[1] : 0.0, h[5] ? (double)h[4] / (h[5] * TC_SCAN_WARPS) : 0.0, h[5] ? (double)h[9] / (h[5] * TC_SCAN_WARPS) : 0.0,
                    h[5] ? (double)h[3] / (h[5] * TC_SCAN_WARPS) : 0.0, h[10], h[11], h[12], h[13]);
        }
    }
    return ASR_OK;
}

int asr_topk(asr_db_t *db, const float *q_dev, int64_t nq, int k, int normalise, float *out_score_dev,
             int64_t *out_idx_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(db != nullptr, "db is NULL");
    ASR_CHECK_ARG(k >= 1 && k <= ASR_MAX_K, "k must be in [1, ASR_MAX_K]");
    ASR_CHECK_ARG(nq >= 0, "nq < 0");
    if (nq == 0) return ASR_OK;
    ASR_CHECK_ARG(q_dev && out_score_dev && out_idx_dev, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const char *force = getenv("ASR_TOPK_PATH");       // "exact" | "tc" (tests exercise both)
    // measured crossovers (k = 25, ms, exact / pre-filter; profiles/r2_topk_summary.md): 1e5 rows: the exact kernel wins at
    // every query count (Q = 256: 0.34 / 0.61); 1e6 rows: the pre-filter from Q ~ 6 (Q = 8: 0.137 / 0.116, Q = 100:
    // 0.91 / 0.39); 1e7 rows: Q = 8 0.56 / 0.38, Q = 16 1.19 / 0.42; 1e8 rows: Q = 4 2.31 / 2.60, Q = 8 4.5 / 2.8.
    // Many queries over a short shard (10k x 125k rows) are the pre-filter's as well: 1e8 scores and more.
    DeviceGuard guard(db->device);
    ASR_CHECK_ARG(normalise || !(db->flags & ASR_DB_NORMALISE_IN_PLACE),
                  "the raw rows are gone: this DB was created with ASR_DB_NORMALISE_IN_PLACE");
    static const int tc_min_q = getenv("ASR_TC_MIN_Q") ? atoi(getenv("ASR_TC_MIN_Q")) : 6;
    static const double tc_min_scores = getenv("ASR_TC_MIN_SCORES") ? atof(getenv("ASR_TC_MIN_SCORES")) : 1.0e8;
    static const double tc_min_rows = getenv("ASR_TC_MIN_ROWS") ? atof(getenv("ASR_TC_MIN_ROWS")) : 4.0e5;
    const bool want_tc = force ? (strcmp(force, "tc") == 0)
                               : (nq >= tc_min_q && ((double)db->n >= tc_min_rows || (double)nq * (double)db->n >= tc_min_scores));
    if (want_tc && normalise && k <= TC_KMAX && db->rows_n) return topk_tc(db, q_dev, nq, k, out_score_dev, out_idx_dev, st);
    // cosine queries stream the pinned-normalised rows (kernel flag bit 0: normalise queries, bit 1: normalise rows
    // in-kernel -- only for handles created without them)
    const bool pre = normalise && db->rows_n;
    const CUtensorMap &tm = pre ? db->tmap_n : db->tmap;
    const int nflag = normalise ? (pre ? 1 : 3) : 0;
    const int qt = nq <= 2 ? 1 : (nq <= 8 ? 4 : 16);
    const int occ = qt == 1 ? 3 : (qt == 4 ? 2 : 1);
    int n_slices, tps, qg0;
    plan_grid(db, nq, qt, occ, &n_slices, &tps, &qg0);
    // queries per launch bounded by the scratch: (score f32 + row u32) per (query, slice, k)
    const int64_t per_q = (int64_t)n_slices * k * 8;
    const int64_t q_chunk = std::max<int64_t>(qt, (int64_t)(TK_SCRATCH_BYTES / per_q) / qt * qt);
    for (int64_t q0 = 0; q0 < nq; q0 += q_chunk) {
        const int64_t nqc = std::min<int64_t>(q_chunk, nq - q0);
        const int64_t n_qt = (nqc + qt - 1) / qt;
        const int qg = (int)std::min<int64_t>(n_qt, std::max<int64_t>(1, (2 * (int64_t)db->sms * occ + n_slices - 1) / n_slices));
        float *ps = reinterpret_cast<float *>(db->scratch);
        uint32_t *pi = reinterpret_cast<uint32_t *>(ps + (size_t)nqc * n_slices * k);
        dim3 grid(n_slices, qg);
        if (qt == 1)
            topk_stream_kernel<1, 2, 3><<<grid, TK_THREADS, sizeof(TkSmem<1, 2>) + 1024, st>>>(
                tm, db->n, tps, q_dev + q0 * 32, (int)nqc, k, nflag, ps, pi);
        else if (qt == 4)
            topk_stream_kernel<4, 2, 2><<<grid, TK_THREADS, sizeof(TkSmem<4, 2>) + 1024, st>>>(
                tm, db->n, tps, q_dev + q0 * 32, (int)nqc, k, nflag, ps, pi);
        else
            topk_stream_kernel<16, 3, 1><<<grid, TK_THREADS, sizeof(TkSmem<16, 3>) + 1024, st>>>(
                tm, db->n, tps, q_dev + q0 * 32, (int)nqc, k, nflag, ps, pi);
        ASR_LAUNCH_CHECK();
        launch_merge(ps, pi, nullptr, db->idx_base, n_slices, k, out_score_dev + q0 * k, out_idx_dev + q0 * k, nqc, st);
        ASR_LAUNCH_CHECK();
    }
    return ASR_OK;
}

// debug/validation: run the tensor-core path once and return the APPROXIMATE scores of the first
// 128 queries against the first 256 DB rows (out_host: 128 x 256 floats)
int asr_debug_tc_scores(asr_db_t *db, const float *q_dev, int64_t nq, float *out_host) {
    float *d = nullptr;
    ASR_CUDA(cudaMalloc(&d, TC_QM * TC_ROWS * 4));
    ASR_CUDA(cudaMemset(d, 0, TC_QM * TC_ROWS * 4));
    float *s = nullptr; int64_t *i = nullptr;
    ASR_CUDA(cudaMalloc(&s, (size_t)nq * 4)); ASR_CUDA(cudaMalloc(&i, (size_t)nq * 8));
    g_tc_dbg = d;
    int rc = topk_tc(db, q_dev, nq, 1, s, i, 0);
    g_tc_dbg = nullptr;
    if (rc) return rc;
    ASR_CUDA(cudaDeviceSynchronize());
    ASR_CUDA(cudaMemcpy(out_host, d, TC_QM * TC_ROWS * 4, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(s); cudaFree(i);
    return ASR_OK;
}

int asr_topk_merge(const float *score_dev, const int64_t *idx_dev, int64_t nq, int n_lists, int k, float *out_score_dev,
                   int64_t *out_idx_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(k >= 1 && k <= ASR_MAX_K && n_lists >= 1, "bad k / n_lists");
    if (nq == 0) return ASR_OK;
    ASR_CHECK_ARG(score_dev && idx_dev && out_score_dev && out_idx_dev, "NULL buffer");
    launch_merge(score_dev, nullptr, idx_dev, 0, n_lists, k, out_score_dev, out_idx_dev, nq, (cudaStream_t)stream);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

int asr_topk_merge_gathered(const void *gathered_dev, int64_t chunk_bytes, int64_t idx_offset_bytes, int64_t nq, int n_lists,
                            int k, float *out_score_dev, int64_t *out_idx_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(k >= 1 && k <= ASR_MAX_K && n_lists >= 1, "bad k / n_lists");
    if (nq == 0) return ASR_OK;
    ASR_CHECK_ARG(gathered_dev && out_score_dev && out_idx_dev, "NULL buffer");
    ASR_CHECK_ARG(chunk_bytes % 8 == 0 && idx_offset_bytes % 8 == 0 && idx_offset_bytes >= nq * k * 4 &&
                      chunk_bytes >= idx_offset_bytes + nq * k * 8,
                  "chunk layout: [scores (nq,k) f32 | pad to 8 | indices (nq,k) i64], chunk_bytes % 8 == 0");
    const uint8_t *g = reinterpret_cast<const uint8_t *>(gathered_dev);
    const size_t smem = (size_t)4 * 2 * k * (sizeof(int64_t) + sizeof(float));
    topk_merge_kernel<4><<<(unsigned)nq, 4 * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float *>(g), nullptr, reinterpret_cast<const int64_t *>(g + idx_offset_bytes), 0, n_lists, k,
        out_score_dev, out_idx_dev, (int64_t)k, chunk_bytes / 4, chunk_bytes / 8);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

int asr_rank_target_merge(const void *gathered_dev, int64_t chunk_bytes, int64_t idx_offset_bytes, int n_lists, int64_t nq,
                          float *tscore_dev, int64_t *tidx_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(n_lists >= 1, "n_lists < 1");
    if (nq == 0) return ASR_OK;
    ASR_CHECK_ARG(gathered_dev && tscore_dev && tidx_dev, "NULL buffer");
    ASR_CHECK_ARG(chunk_bytes % 8 == 0 && idx_offset_bytes % 8 == 0 && idx_offset_bytes >= nq * 4 &&
                      chunk_bytes >= idx_offset_bytes + nq * 8,
                  "chunk layout: [scores (nq) f32 | pad to 8 | indices (nq) i64], chunk_bytes % 8 == 0");
    const uint8_t *g = reinterpret_cast<const uint8_t *>(gathered_dev);
    rank_target_merge_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float *>(g), reinterpret_cast<const int64_t *>(g + idx_offset_bytes), n_lists, nq,
        chunk_bytes / 4, chunk_bytes / 8, tscore_dev, tidx_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

int asr_rank_of_target(asr_db_t *db, const float *q_dev, int64_t nq, int64_t q_base, int64_t kg, int64_t hg,
                       int normalise, int phase, float *tscore_dev, int64_t *tidx_dev, int64_t *better_dev,
                       void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(db != nullptr && q_dev && tscore_dev && tidx_dev, "NULL argument");
    ASR_CHECK_ARG(kg >= 1 && hg >= 1, "kg, hg must be >= 1");
    ASR_CHECK_ARG(!(db->flags & ASR_DB_NORMALISE_IN_PLACE), "not available for ASR_DB_NORMALISE_IN_PLACE handles");
    DeviceGuard guard(db->device);
    if (nq == 0) return ASR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (phase == 0) {
        rank_target_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(db->codes, db->n, db->idx_base, q_dev, (int)nq,
                                                                        q_base, kg, hg, normalise, tscore_dev, tidx_dev);
        ASR_LAUNCH_CHECK();
        return ASR_OK;
    }
    ASR_CHECK_ARG(better_dev != nullptr, "better_dev is NULL");
    int ns, tps, qg;
    plan_grid(db, nq, TK_QT_MAX, 1, &ns, &tps, &qg);
    dim3 grid(ns, qg);
    rank_count_kernel<<<grid, TK_THREADS, sizeof(RkSmem) + 1024, st>>>(db->tmap, db->n, db->idx_base, tps, q_dev, (int)nq,
                                                                       normalise, tscore_dev, tidx_dev,
                                                                       reinterpret_cast<unsigned long long *>(better_dev));
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

int asr_vote(const int64_t *cand_idx_dev, const int32_t *row_ids_dev, int64_t n_rows, int n_rec, int m, int top_k,
             int32_t *out_ids_dev, int32_t *out_counts_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(cand_idx_dev && row_ids_dev && out_ids_dev && out_counts_dev, "NULL buffer");
    ASR_CHECK_ARG(m >= 1 && m <= VOTE_MAX, "m must be in [1, 8192]");
    ASR_CHECK_ARG(top_k >= 1, "top_k < 1");
    if (n_rec == 0) return ASR_OK;
    int p2 = 32;
    while (p2 < m) p2 <<= 1;
    static bool attr_done[ASR_MAX_DEVICES] = {false};     // cudaFuncSetAttribute is per device
    const int attr_dev = std::max(0, std::min(current_device(), ASR_MAX_DEVICES - 1));
    if (!attr_done[attr_dev]) {
        ASR_CUDA(cudaFuncSetAttribute(vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VOTE_MAX * 8));
        attr_done[attr_dev] = true;
    }
    vote_kernel<<<n_rec, VOTE_THREADS, p2 * 8, (cudaStream_t)stream>>>(cand_idx_dev, row_ids_dev, n_rows, m, p2, top_k,
                                                                       out_ids_dev, out_counts_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

}  // extern "C"
