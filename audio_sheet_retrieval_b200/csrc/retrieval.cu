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
// For many queries the pinned-order scoring (64 dependent fp32 instructions per (row, query)) is
// FP32-issue bound.  This path scores a 128-query x 256-row tile with four tcgen05 tf32 MMAs
// (fp32 accumulators in TMEM), which is only an APPROXIMATION s~ of the pinned score s, with
// |s~ - s| <= 2^-9 for unit vectors (each tf32 operand carries <= 2^-10 relative error and
// sum |q_k d_k| <= 1).  Exactness is kept by filtering, not by trusting s~: a row is re-scored with
// the pinned fp32 order iff s~ >= tau_q - eps, where tau_q is the query's current k-th best EXACT
// score and eps = 2^-8.  Every true top-k row satisfies s >= tau_final >= tau_q, hence
// s~ >= tau_q - eps, so it is always re-scored; lists, thresholds and the final order use exact
// scores only.  Results are therefore identical to topk_stream_kernel / the oracle.
// Layout: thread = query (= TMEM lane); each thread keeps its query's sorted top-k list in registers.
constexpr int TC_QM = 128;
constexpr int TC_ROWS = 256;
constexpr int TC_STAGES = 3;
constexpr int TC_KMAX = 32;
constexpr int TC_EPI_WARPS = 8;      // two per TMEM lane quarter (column halves)
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

constexpr int TC_LSTRIDE = TC_KMAX + 1;   // list stride in keys: odd, so that the same position of different lists hits different banks
constexpr int TC_QCAP = 64;               // candidate queue entries per epilogue warp (drained 32 at a time)

struct TcSmem {
    float stage[TC_STAGES][TC_ROWS * 32];     // normalised DB rows, SWIZZLE_128B (written by TMA)
    float q[TC_QM * 32];                      // normalised queries, same swizzle (written by threads)
    unsigned short cand[16][TC_QM];           // pending-candidate masks of the current tile: [16-column group][TMEM lane]
    unsigned long long lists[2 * TC_QM * TC_LSTRIDE];   // sorted top-k list of every epilogue thread (column half, TMEM lane) as 64-bit keys
    unsigned long long floor_key[2 * TC_QM];  // threshold floor of the list (first-tile bisection)
    float tau[2 * TC_QM];                     // approximate-score filter threshold of the list
    uint32_t cq_row[TC_EPI_WARPS][TC_QCAP];   // candidate queues: DB row ...
    unsigned char cq_lane[TC_EPI_WARPS][TC_QCAP];   // ... and the lane (within the warp) whose list it is for
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

// Drain `n` (<= 32) queued candidates of one epilogue warp, ONE PER LANE: exact score (the row comes back from L2,
// where the TMA load of its tile just put it), then insertion into the owner lane's sorted list in shared memory.
// Lanes whose candidates belong to the same list take turns (lowest lane first); lists of different lanes are
// independent, so a round inserts up to 32 candidates at once.
// gkey[q]: fkey of a LOWER BOUND of query q's final k-th best exact score, shared by every list of the query on this
// GPU (other DB slices, column halves, lane copies): each list publishes its own k-th best with atomicMax as soon as it
// holds k real entries.  A row scoring strictly below the bound cannot be in the query's top-k, so every list filters
// with the best bound anybody has found -- the per-item threshold warm-up (k ln(rows / k) insertions per list) is paid
// about once per query instead of once per list.
__device__ __forceinline__ void tc_drain(TcSmem &sm, const float *__restrict__ rows, int n, int ew, int quarter, int lbase, int k,
                                         float eps, int lane, unsigned *__restrict__ gkey, int q0, int nqt_mask) {
    const bool have = lane < n;
    int owner = 64 + lane;                                   // distinct dummy owners for idle lanes
    unsigned long long key = 0ull;
    bool todo = false;
    if (have) {
        owner = sm.cq_lane[ew][lane];
        const uint32_t row = sm.cq_row[ew][lane];
        const int ql = quarter * 32 + owner;
        const float sc = tc_exact_score_g(sm.q, rows + (size_t)row * 32, ql);
        key = ((unsigned long long)fkey(sc) << 32) | (unsigned long long)(~row);
        const unsigned long long kth = sm.lists[(lbase + owner) * TC_LSTRIDE + k - 1], fl = sm.floor_key[lbase + owner];
        todo = key > (kth > fl ? kth : fl) && fkey(sc) >= __ldcg(gkey + q0 + (ql & nqt_mask));
    }
    while (__any_sync(0xffffffffu, todo)) {
        const unsigned peers = __match_any_sync(0xffffffffu, todo ? owner : 64 + lane);
        if (todo && (int)(__ffs(peers) - 1) == lane) {
            unsigned long long *l = sm.lists + (lbase + owner) * TC_LSTRIDE;
            if (key > l[k - 1]) {                            // an earlier round may have raised the k-th entry
                int j = k - 1;
                for (; j > 0; --j) {
                    const unsigned long long up = l[j - 1];
                    if (up > key) break;
                    l[j] = up;
                }
                l[j] = key;
                const unsigned long long kth = l[k - 1], fl = sm.floor_key[lbase + owner];
                sm.tau[lbase + owner] = fkey_inv((uint32_t)((kth > fl ? kth : fl) >> 32)) - eps;
                if ((uint32_t)kth != 0u)                     // a real k-th entry (empty slots have lo = 0): publish the bound
                    atomicMax(gkey + q0 + ((quarter * 32 + owner) & nqt_mask), (unsigned)(kth >> 32));
            }
            todo = false;
        }
        __syncwarp();
    }
}

// Persistent 1-D grid; work items (query tile, DB slice) are handed out by an atomic counter so that
// any (nq, n_db) shape fills all SMs.  qn = normalised queries (nq,32); tmap over the normalised DB rows.
//
// Epilogue layout.  A score tile is 128 TMEM lanes x 256 columns (DB rows).  Reading it out of TMEM and scanning it
// bounds this kernel, and with one warp per lane quarter that is a latency-bound chain.  EIGHT epilogue warps share a
// tile: the two warps of a lane quarter scan one half of the columns each, and for <= 64 queries the query tile is
// REPLICATED over the lane quarters (rep = 2) so that no lane idles.  Every (TMEM lane, column half) thread owns a
// sorted top-k list in SHARED memory; rows that pass the approximate filter go through a per-warp candidate queue and
// are re-scored exactly / inserted 32 at a time, one candidate per lane, whichever list it belongs to (the old
// thread-owns-its-list-in-registers form spent 2/3 of its instructions in a lock-step loop with ~1 active lane).
// At item end the L = 2 * rep lists of a query are merged by rank counting.
__global__ void __launch_bounds__(TC_THREADS, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_db, int tiles_per_slice, int n_slices,
               const float *__restrict__ rows, const float *__restrict__ qn, int nq, int k, float eps, int rep,
               float *__restrict__ part_s, uint32_t *__restrict__ part_i, unsigned *__restrict__ work_counter,
               unsigned *__restrict__ gkey, float *dbg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform: uniform role branches
    if (smem_u32(smem_raw) & 1023u) __trap();
    const int64_t n_tiles = (n_db + TC_ROWS - 1) / TC_ROWS;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.tfull[s], 1); mbar_init(&sm.tempty[s], TC_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    if (warp == 1) { tmem_alloc(&sm.tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_ptr;
    // kind::tf32: a/b format 2 (TF32), fp32 accumulate, K-major both, M = 128, N = 256
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_ROWS >> 3) << 17) | ((uint32_t)(TC_QM >> 4) << 24);

    const int nqt = TC_QM / rep;                     // distinct queries per tile
    const int L = 2 * rep;                           // lists per (query, item): rep lane copies x 2 column halves
    const int ncols = TC_ROWS / L;                   // columns per epilogue thread: 128 or 64
    const int n_qt = (nq + nqt - 1) / nqt;
    const unsigned n_items = (unsigned)n_qt * (unsigned)n_slices;
    int64_t itg = 0;                 // running tile counter (barrier phases continue across work items)
    for (;;) {
        if (tid == 0) sm.item = atomicAdd(work_counter, 1u);
        __syncthreads();
        const unsigned item = sm.item;
        if (item >= n_items) break;
        // consecutive items share the DB slice (L2 reuse across CTAs), query tile varies fastest
        const int slice = (int)(item / (unsigned)n_qt), qt = (int)(item % (unsigned)n_qt);
        const int64_t tile0 = (int64_t)slice * tiles_per_slice;
        int64_t my_tiles = n_tiles - tile0;
        if (my_tiles > tiles_per_slice) my_tiles = tiles_per_slice;
        if (my_tiles < 0) my_tiles = 0;
        const int q0 = qt * nqt;
        // stage the query tile (rep copies), swizzled like a SWIZZLE_128B TMA load; rows beyond nq are zero
        for (int i = tid; i < TC_QM * 8; i += TC_THREADS) {
            const int row = i >> 3, c = i & 7;
            const int qi = row & (nqt - 1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q0 + qi < nq) v = *reinterpret_cast<const float4 *>(qn + (int64_t)(q0 + qi) * 32 + c * 4);
            *reinterpret_cast<float4 *>(sm.q + row * 32 + ((c ^ (row & 7)) << 2)) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
        __syncthreads();

        if (warp == 0) {
            if (lane == 0) {
                for (int64_t it = 0; it < my_tiles; ++it) {
                    const int64_t g = itg + it;
                    const int s = (int)(g % TC_STAGES);
                    mbar_wait(&sm.empty[s], (uint32_t)(((g / TC_STAGES) & 1) ^ 1));
                    mbar_expect_tx(&sm.full[s], TC_ROWS * 128);
                    tma_load_2d(sm.stage[s], &tmap, 0, (int)((tile0 + it) * TC_ROWS), &sm.full[s]);
                }
            }
        } else if (warp == 1) {
            const uint32_t leader = elect_one() ? 1u : 0u;
            const uint32_t qaddr = smem_u32(sm.q);
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int64_t g = itg + it;
                const int s = (int)(g % TC_STAGES);
                const uint32_t slot = (uint32_t)(g & 1);
                mbar_wait(&sm.full[s], (uint32_t)((g / TC_STAGES) & 1));
                mbar_wait(&sm.tempty[slot], (uint32_t)(((g >> 1) & 1) ^ 1));
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
        } else {
            const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
            const int ql = quarter * 32 + lane;                 // TMEM lane = row of the (replicated) query tile
            const int qi = ql & (nqt - 1);                      // query within the tile
            const int hsel = (warp - 2) >> 2;                   // which of the two warps of this lane quarter
            const int sub = (ql / nqt) * 2 + hsel;              // which of the L column ranges of that query
            const int li = hsel * TC_QM + ql;                   // this thread's list
            const int col0 = sub * ncols;
            const bool qvalid = q0 + qi < nq;
            const int ew = warp - 2;
            // The sorted top-k list of every lane lives in SHARED memory as 64-bit keys (hi = order-preserving map of
            // the exact score, lo = ~row: one unsigned compare is exactly "score desc, index asc"; empty slots hold the
            // smallest key, (-inf, row 0xffffffff)).  Pass 1 scans the approximate scores against the lane's tau;
            // what passes goes into the warp's candidate queue and is re-scored / inserted 32 at a time (tc_drain).
            const unsigned long long NEG_KEY = (unsigned long long)0x007fffffu << 32;      // (fkey(-inf), 0)
            for (int j = 0; j < k; ++j) sm.lists[li * TC_LSTRIDE + j] = NEG_KEY;
            sm.floor_key[li] = NEG_KEY;
            sm.tau[li] = -CUDART_INF_F;
            __syncwarp();
            int qcount = 0;                                      // warp-uniform
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int64_t g = itg + it;
                const uint32_t slot = (uint32_t)(g & 1);
                mbar_wait(&sm.tfull[slot], (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + slot * 256u + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0;
                const int64_t row0 = (tile0 + it) * TC_ROWS;
                if (it == 0 && row0 + TC_ROWS <= n_db && ncols >= k) {
                    // Threshold floor from the first (full) tile: bisect for a value `lo` such that at least k of this
                    // lane's approximate scores are >= lo.  Those k rows have exact scores >= lo - eps, so the final
                    // k-th best of this list is >= lo - eps: a valid floor that removes most of the warm-up.
                    float lo = -2.0f, hi = 2.0f;
#pragma unroll 1
                    for (int iter = 0; iter < 12; ++iter) {
                        const float mid = 0.5f * (lo + hi);
                        int cnt = 0;
#pragma unroll 1
                        for (int c64 = 0; c64 < ncols; c64 += 64) {
                            float v[64];
                            tmem_ld64(taddr + (uint32_t)c64, v);
#pragma unroll
                            for (int j = 0; j < 64; ++j) cnt += v[j] >= mid ? 1 : 0;
                        }
                        if (cnt >= k) lo = mid; else hi = mid;
                    }
                    if (lo > -2.0f) {          // found (queries of all-NaN / degenerate tiles keep -inf)
                        const float floor_s = lo - eps;
                        sm.floor_key[li] = (unsigned long long)fkey(floor_s) << 32;   // ties with the floor itself still enter
                        sm.tau[li] = floor_s - eps;
                    }
                }
                // the best bound anybody on this GPU has published for this query (read from L2 once per tile)
                const float gtau = qvalid ? fkey_inv(__ldcg(gkey + q0 + qi)) - eps : -CUDART_INF_F;
                const float tau = fmaxf(sm.tau[li], gtau);
                // Pass 1: scan this lane's approximate scores, 64 columns per TMEM round trip, and remember which
                // columns pass the filter as 16-bit masks.  NaN approximations (zero DB rows) need no special case:
                // they can only matter while tau is -inf, and then !(x < tau) is true for every x.
                unsigned pend = 0u;
                // software pipeline over the 64-column chunks: the TMEM round trip of chunk c + 1 runs under the scan of c
                auto scan = [&](const uint32_t *vr, int c64) {
                    if (dbg && slice == 0 && qt == 0 && it == 0) {      // debug: dump the approximate scores of tile 0
#pragma unroll
                        for (int j = 0; j < 64; ++j) dbg[qi * TC_ROWS + col0 + c64 + j] = __uint_as_float(vr[j]);
                    }
                    float m[4];
#pragma unroll
                    for (int sb = 0; sb < 4; ++sb) {
                        m[sb] = __uint_as_float(vr[sb * 16]);
#pragma unroll
                        for (int j = 1; j < 16; ++j) m[sb] = fmaxf(m[sb], __uint_as_float(vr[sb * 16 + j]));
                    }
                    const float mm = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
                    if (qvalid && !(mm < tau)) {
#pragma unroll
                        for (int sb = 0; sb < 4; ++sb) {
                            if (!(m[sb] < tau)) {
                                unsigned mask = 0u;        // static indexing keeps the chunk in registers
#pragma unroll
                                for (int j = 0; j < 16; ++j) mask |= !(__uint_as_float(vr[sb * 16 + j]) < tau) ? (1u << j) : 0u;
                                const int g16 = (col0 + c64) / 16 + sb;
                                sm.cand[g16][ql] = (unsigned short)mask;
                                pend |= 1u << ((c64 / 16) + sb);          // bit = 16-column group within this lane's range
                            }
                        }
                    }
                };
                {
                    uint32_t va[64], vb[64];
                    tmem_ld64_issue(taddr, va);
#pragma unroll 1
                    for (int c64 = 0; c64 < ncols; c64 += 128) {
                        tmem_ld_wait();
                        if (c64 + 64 < ncols) tmem_ld64_issue(taddr + (uint32_t)(c64 + 64), vb);
                        scan(va, c64);
                        if (c64 + 64 < ncols) {
                            tmem_ld_wait();
                            if (c64 + 128 < ncols) tmem_ld64_issue(taddr + (uint32_t)(c64 + 128), va);
                            scan(vb, c64 + 64);
                        }
                    }
                }
                // the accumulator slot is free again as soon as every epilogue warp has scanned it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tempty[slot]);
                // Pass 2: move the candidates into the warp's queue (one per lane per round, a handful of instructions),
                // draining 32 of them whenever that many are waiting.
                while (__any_sync(0xffffffffu, pend != 0u)) {
                    bool have = false;
                    uint32_t row = 0u;
                    if (pend) {
                        const int gq = __ffs(pend) - 1;
                        const int g16 = col0 / 16 + gq;
                        unsigned mask = sm.cand[g16][ql];
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        sm.cand[g16][ql] = (unsigned short)mask;
                        if (!mask) pend &= ~(1u << gq);
                        const int64_t r = row0 + g16 * 16 + j;
                        have = r < n_db;
                        row = (uint32_t)r;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, have);
                    if (have) {
                        const int pos = qcount + __popc(bal & ((1u << lane) - 1u));
                        sm.cq_row[ew][pos] = row;
                        sm.cq_lane[ew][pos] = (unsigned char)lane;
                    }
                    qcount += __popc(bal);
                    __syncwarp();
                    if (qcount >= 32) {
                        tc_drain(sm, rows, 32, ew, quarter, hsel * TC_QM + quarter * 32, k, eps, lane, gkey, q0, nqt - 1);
                        qcount -= 32;
                        const uint32_t mr = sm.cq_row[ew][32 + lane];          // move the remainder (< 32 entries) to the front
                        const unsigned char ml = sm.cq_lane[ew][32 + lane];
                        __syncwarp();
                        if (lane < qcount) { sm.cq_row[ew][lane] = mr; sm.cq_lane[ew][lane] = ml; }
                        __syncwarp();
                    }
                }
                // thresholds only ever rise, so a drain may be deferred; it must happen before the lists are read
                if (qcount && (it + 1 == my_tiles || (it & 7) == 7)) {
                    tc_drain(sm, rows, qcount, ew, quarter, hsel * TC_QM + quarter * 32, k, eps, lane, gkey, q0, nqt - 1);
                    qcount = 0;
                }
            }
            // Item end: the rep lanes that served the same query merge their sorted lists by RANK COUNTING -- the merged
            // position of an entry is its own index plus the number of entries of the other lists that beat it (binary
            // search; keys are unique, so positions are too) -- and every lane writes its survivors straight to the
            // item's single output list.
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            if (qvalid) {
                const size_t o = ((size_t)(q0 + qi) * n_slices + slice) * k;
                const unsigned long long *ml = sm.lists + li * TC_LSTRIDE;
                for (int j = 0; j < k; ++j) {
                    const unsigned long long mine = ml[j];
                    int rank = j;
                    for (int t = 0; t < L; ++t) {
                        const int other = (t & 1) * TC_QM + qi + (t >> 1) * nqt;
                        if (other == li) continue;
                        const unsigned long long *ol = sm.lists + other * TC_LSTRIDE;
                        int lo = 0, hi = k;                   // number of entries of `ol` (sorted descending) greater than mine
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (ol[mid] > mine) lo = mid + 1; else hi = mid;
                        }
                        rank += lo;
                    }
                    if (rank < k) {
                        const bool real = mine != NEG_KEY;
                        part_s[o + rank] = real ? fkey_inv((uint32_t)(mine >> 32)) : -CUDART_INF_F;
                        part_i[o + rank] = real ? ~(uint32_t)mine : 0xffffffffu;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");   // lists[] free for the next item
        }
        itg += my_tiles;
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
    // Replicate few queries over the TMEM lane quarters so that all eight epilogue warps share a score tile.
    // (measured, 1e7 rows: 32 queries 1.01 ms replicated vs 1.07 ms not; 64 queries 1.40 vs 1.12 ms)
    const int rep = getenv("ASR_TC_REP") ? std::min(2, std::max(1, atoi(getenv("ASR_TC_REP")))) : (nq <= 32 ? 2 : 1);
    const int nqt = TC_QM / rep, L = 1;           // one list per (query, slice): the column ranges are merged in-kernel
    // Work items = (query tile, DB slice), ~2 per SM (1 when there are only a few query tiles: every slice adds a list
    // per query to the merge).  Since the lists of a query share their threshold bound (gkey), short items are cheap:
    // 10k queries x 125k rows (one shard of an 8-GPU DB) 2.94 -> 1.79 ms, 10k x 1M 7.0 -> 5.2 ms with the sharing.
    const int64_t n_tiles = (db->n + TC_ROWS - 1) / TC_ROWS;
    const int64_t n_qt_all = (nq + nqt - 1) / nqt;
    const int items_per_sm = getenv("ASR_TC_ITEMS_PER_SM") ? atoi(getenv("ASR_TC_ITEMS_PER_SM")) : (n_qt_all >= 8 ? 2 : 1);
    const int min_tiles = getenv("ASR_TC_MIN_TILES") ? atoi(getenv("ASR_TC_MIN_TILES")) : (n_qt_all >= 8 ? 64 : 32);
    int64_t want = std::max<int64_t>(1, (items_per_sm * (int64_t)db->sms + n_qt_all - 1) / n_qt_all);
    want = std::min<int64_t>(want, std::max<int64_t>(1, n_tiles / min_tiles));
    const int tps = (int)((n_tiles + want - 1) / want);
    const int n_slices = (int)((n_tiles + tps - 1) / tps);
    const int64_t per_q = (int64_t)n_slices * L * k * 8;
    const int64_t q_chunk = std::max<int64_t>(nqt, (int64_t)((TK_SCRATCH_BYTES - 256 - (db->qn_cap + 64) * 4) / per_q) / nqt * nqt);
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
        const int grid = (int)std::min<int64_t>(db->sms, n_qt * n_slices);
        topk_tc_kernel<<<grid, TC_THREADS, sizeof(TcSmem), st>>>(db->tmap_n, db->n, tps, n_slices, db->rows_n, db->qn + q0 * 32,
                                                                 (int)nqc, k, 0.00390625f, rep, ps, pi, counter, gkey + q0, g_tc_dbg);
        ASR_LAUNCH_CHECK();
        launch_merge(ps, pi, nullptr, db->idx_base, n_slices * L, k, out_score_dev + q0 * k, out_idx_dev + q0 * k, nqc, st);
        ASR_LAUNCH_CHECK();
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
    // measured crossovers (k = 25): the exact QT=16 kernel wins up to ~24 queries, and for small problems
    // whatever the query count: exact ~ 0.05 ms + 7.3 ms per 1e9 scores, pre-filter ~ 1.15 ms + 0.55 ms per 1e9
    // (profiles/r1_configs_3_5.json: its work items are at least 128 tiles long) -> equal at 1.6e8 scores
    DeviceGuard guard(db->device);
    ASR_CHECK_ARG(normalise || !(db->flags & ASR_DB_NORMALISE_IN_PLACE),
                  "the raw rows are gone: this DB was created with ASR_DB_NORMALISE_IN_PLACE");
    static const int tc_min_q = getenv("ASR_TC_MIN_Q") ? atoi(getenv("ASR_TC_MIN_Q")) : 12;
    static const double tc_min_scores = getenv("ASR_TC_MIN_SCORES") ? atof(getenv("ASR_TC_MIN_SCORES")) : 1.0e8;
    const bool want_tc = force ? (strcmp(force, "tc") == 0) : (nq >= tc_min_q && (double)nq * (double)db->n >= tc_min_scores);
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
