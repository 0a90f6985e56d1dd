// Audio <-> sheet alignment: dense cosine-distance matrix (fp64) + dynamic time warping.
//
// Replaces (reference, paths relative to its root):
//   audio_sheet_retrieval/utils/alignment.py:149   cdist(img_codes, spec_codes, 'cosine')
//   audio_sheet_retrieval/utils/dtw_by_dist.py:6-34   accumulated-cost matrix (pure-Python O(r*c) double loop)
//   audio_sheet_retrieval/utils/dtw_by_dist.py:69-83  traceback (first minimum of diag / up / left)
//
// The accumulated-cost recurrence only couples a cell to its three predecessors, so the matrix is
// filled anti-diagonal by anti-diagonal by one 1024-thread CTA (a grid-wide barrier per diagonal
// would cost more than the work); all arithmetic is fp64 like the reference's NumPy arrays, and
// every cell performs exactly the reference's operations (one min of three, one add), so the
// matrix and the path are identical to the reference's for the same distance matrix.
#include <math_constants.h>

#include "common.cuh"

namespace asr {

// dist[i][j] = 1 - <a_i, b_j> / sqrt(|a_i|^2 |b_j|^2)   (sequential fp64 sums, clipped like SciPy)
__global__ void cosine_dist_kernel(const float *__restrict__ a, int r, const float *__restrict__ b, int c,
                                   double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= c) return;
    double uv = 0.0, uu = 0.0, vv = 0.0;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const double u = (double)a[(size_t)i * 32 + k], v = (double)b[(size_t)j * 32 + k];
        uv = __dadd_rn(uv, __dmul_rn(u, v));
        uu = __dadd_rn(uu, __dmul_rn(u, u));
        vv = __dadd_rn(vv, __dmul_rn(v, v));
    }
    double cosine = uv / (sqrt(uu) * sqrt(vv));
    if (fabs(cosine) > 1.0) cosine = copysign(1.0, cosine);
    out[(size_t)i * c + j] = 1.0 - cosine;
}

constexpr int DTW_THREADS = 1024;

// acc (r,c) <- accumulated cost.  Borders as in dtw_by_dist.py:20-22 (D0[0,0] = 0, first row/col = inf).
__global__ void __launch_bounds__(DTW_THREADS) dtw_accumulate_kernel(const double *__restrict__ dist, int r, int c,
                                                                    double *__restrict__ acc) {
    const double INF = CUDART_INF;
    for (int d = 0; d < r + c - 1; ++d) {
        const int i_lo = d - c + 1 > 0 ? d - c + 1 : 0;
        const int i_hi = d < r - 1 ? d : r - 1;
        for (int i = i_lo + (int)threadIdx.x; i <= i_hi; i += DTW_THREADS) {
            const int j = d - i;
            const double diag = (i > 0 && j > 0) ? acc[(size_t)(i - 1) * c + j - 1] : ((i == 0 && j == 0) ? 0.0 : INF);
            const double up = i > 0 ? acc[(size_t)(i - 1) * c + j] : INF;
            const double left = j > 0 ? acc[(size_t)i * c + j - 1] : INF;
            acc[(size_t)i * c + j] = dist[(size_t)i * c + j] + fmin(fmin(diag, up), left);
        }
        __syncthreads();
    }
}

// path (i,j) pairs from (0,0) to (r-1,c-1); tie-break = first minimum of (diag, up, left)
__global__ void dtw_traceback_kernel(const double *__restrict__ acc, int r, int c, int *__restrict__ path_i,
                                     int *__restrict__ path_j, int *__restrict__ path_len) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double INF = CUDART_INF;
    const int cap = r + c - 1;
    int i = r - 1, j = c - 1, n = 0;
    path_i[cap - 1] = i; path_j[cap - 1] = j; n = 1;
    while (i > 0 || j > 0) {
        const double diag = (i > 0 && j > 0) ? acc[(size_t)(i - 1) * c + j - 1] : INF;
        const double up = i > 0 ? acc[(size_t)(i - 1) * c + j] : INF;
        const double left = j > 0 ? acc[(size_t)i * c + j - 1] : INF;
        if (diag <= up && diag <= left) { --i; --j; }
        else if (up <= left) { --i; }
        else { --j; }
        path_i[cap - 1 - n] = i; path_j[cap - 1 - n] = j; ++n;
    }
    for (int t = 0; t < n; ++t) { path_i[t] = path_i[cap - n + t]; path_j[t] = path_j[cap - n + t]; }
    *path_len = n;
}

}  // namespace asr

using namespace asr;

extern "C" {

int asr_cosine_distances(const float *a_dev, int r, const float *b_dev, int c, double *out_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(a_dev && b_dev && out_dev && r >= 1 && c >= 1, "bad argument");
    dim3 grid((c + 127) / 128, r);
    cosine_dist_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a_dev, r, b_dev, c, out_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

int asr_dtw(const double *dist_dev, int r, int c, double *acc_dev, int32_t *path_i_dev, int32_t *path_j_dev,
            int32_t *path_len_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(dist_dev && acc_dev && path_i_dev && path_j_dev && path_len_dev, "NULL buffer");
    ASR_CHECK_ARG(r >= 1 && c >= 1, "empty matrix");
    dtw_accumulate_kernel<<<1, DTW_THREADS, 0, (cudaStream_t)stream>>>(dist_dev, r, c, acc_dev);
    ASR_LAUNCH_CHECK();
    dtw_traceback_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc_dev, r, c, path_i_dev, path_j_dev, path_len_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

}  // extern "C"
