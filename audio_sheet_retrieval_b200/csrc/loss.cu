// Training objective of the reference (SURVEY 8f "next" row 4, first slice): the pairwise contrastive
// cosine hinge loss and its gradient with respect to both code matrices.
//
// Replaces (reference, paths relative to its root):
//   audio_sheet_retrieval/models/objectives.py:30-69   get_contrastive_cos_loss(weight, gamma, symmetric)
//     D = lv1 . lv2^T, d = diag(D);  L_ij = clip(gamma - d_i + D_ij, 0, 1000) for j != i;
//     loss = weight * (mean(L) [+ the same with the roles of the views swapped, i.e. on D^T]).
//   The gradient is what Theano's autodiff derives from that graph (clip passes the gradient on
//   [0, 1000], both ends included).
//
// One CTA per row i keeps row i and column i of D in shared memory (the batch is ~100 x 32: the
// n x n matrix is never written), reduces its hinge terms in a fixed order and accumulates its two
// gradient rows; a second launch adds the n row sums in order.  Everything is deterministic.
#include "common.cuh"

namespace asr {

constexpr int LOSS_THREADS = 128;
constexpr int LOSS_DIM = 32;

__device__ __forceinline__ float dot32(const float *__restrict__ a, const float *__restrict__ b) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LOSS_DIM; ++k) s = fmaf(a[k], b[k], s);
    return s;
}

// smem: row[n] = D_ij, col[n] = D_ji, diag[n] = d_j, then the coefficient vectors c1[n], c2[n]
__global__ void __launch_bounds__(LOSS_THREADS) contrastive_rows_kernel(const float *__restrict__ lv1,
                                                                        const float *__restrict__ lv2, int n, float gamma,
                                                                        int symmetric, float scale, double *__restrict__ row_loss,
                                                                        float *__restrict__ g1, float *__restrict__ g2) {
    extern __shared__ float sm[];
    float *row = sm, *col = sm + n, *diag = sm + 2 * n, *c1 = sm + 3 * n, *c2 = sm + 4 * n;
    __shared__ double red[LOSS_THREADS];
    __shared__ float gii_sm;
    const int i = blockIdx.x, tid = threadIdx.x;
    const float *a_i = lv1 + (size_t)i * LOSS_DIM, *b_i = lv2 + (size_t)i * LOSS_DIM;
    for (int j = tid; j < n; j += LOSS_THREADS) {
        const float *a_j = lv1 + (size_t)j * LOSS_DIM, *b_j = lv2 + (size_t)j * LOSS_DIM;
        row[j] = dot32(a_i, b_j);
        col[j] = dot32(a_j, b_i);
        diag[j] = dot32(a_j, b_j);
    }
    __syncthreads();
    const float d_i = diag[i];
    // hinge terms of row i (direction 1 on D, direction 2 on D^T) and the dL/dD coefficients
    double part = 0.0;
    float act_sum = 0.f;
    for (int j = tid; j < n; j += LOSS_THREADS) {
        float k1 = 0.f, k2 = 0.f;
        if (j != i) {
            const float x1 = gamma - d_i + row[j];                 // L_ij, direction 1
            const float a_ij = (x1 >= 0.f && x1 <= 1000.f) ? 1.f : 0.f;
            part += (double)fminf(fmaxf(x1, 0.f), 1000.f);
            const float y1 = gamma - diag[j] + col[j];             // L_ji, direction 1 (row j, column i)
            const float a_ji = (y1 >= 0.f && y1 <= 1000.f) ? 1.f : 0.f;
            float s_ij = 0.f, s_ji = 0.f;
            if (symmetric) {
                const float x2 = gamma - d_i + col[j];             // L'_ij = clip(gamma - d_i + D_ji), direction 2
                s_ij = (x2 >= 0.f && x2 <= 1000.f) ? 1.f : 0.f;
                part += (double)fminf(fmaxf(x2, 0.f), 1000.f);
                const float y2 = gamma - diag[j] + row[j];         // L'_ji = clip(gamma - d_j + D_ij)
                s_ji = (y2 >= 0.f && y2 <= 1000.f) ? 1.f : 0.f;
            }
            k1 = a_ij + s_ji;            // dL/dD_ij
            k2 = a_ji + s_ij;            // dL/dD_ji
            act_sum += a_ij + s_ij;      // -dL/dD_ii collects the active terms of row i in both directions
        }
        c1[j] = k1;
        c2[j] = k2;
    }
    red[tid] = part;
    __syncthreads();
    for (int s = LOSS_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    if (tid == 0) row_loss[i] = red[0];
    __syncthreads();
    red[tid] = (double)act_sum;
    __syncthreads();
    for (int s = LOSS_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    if (tid == 0) gii_sm = -(float)red[0];
    __syncthreads();
    if (g1 == nullptr && g2 == nullptr) return;
    // grad1[i] = scale * (sum_j c1[j] lv2[j] + G_ii lv2[i]);  grad2[i] = scale * (sum_j c2[j] lv1[j] + G_ii lv1[i])
    if (tid < 2 * LOSS_DIM) {
        const int t = tid & (LOSS_DIM - 1);
        const bool second = tid >= LOSS_DIM;
        const float *src = second ? lv1 : lv2;
        const float *cf = second ? c2 : c1;
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc = fmaf(cf[j], src[(size_t)j * LOSS_DIM + t], acc);
        acc = fmaf(gii_sm, src[(size_t)i * LOSS_DIM + t], acc);
        float *dst = second ? g2 : g1;
        if (dst) dst[(size_t)i * LOSS_DIM + t] = scale * acc;
    }
}

__global__ void contrastive_sum_kernel(const double *__restrict__ row_loss, int n, double scale, float *__restrict__ loss) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += row_loss[i];
    *loss = (float)(s * scale);
}

}  // namespace asr

using namespace asr;

extern "C" {

int asr_contrastive_loss(const float *lv1_dev, const float *lv2_dev, int64_t n, float weight, float gamma, int symmetric,
                         double *scratch_dev, float *loss_dev, float *grad1_dev, float *grad2_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(lv1_dev && lv2_dev && scratch_dev && loss_dev, "NULL buffer");
    ASR_CHECK_ARG(n >= 2 && n <= 8192, "batch size must be in [2, 8192]");
    const double denom = (double)n * (double)(n - 1);
    const size_t smem = (size_t)5 * n * sizeof(float);
    static bool attr_done[ASR_MAX_DEVICES] = {false};     // cudaFuncSetAttribute is per device
    const int attr_dev = std::max(0, std::min(current_device(), ASR_MAX_DEVICES - 1));
    if (!attr_done[attr_dev]) {
        ASR_CUDA(cudaFuncSetAttribute(contrastive_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 8192 * 4));
        attr_done[attr_dev] = true;
    }
    contrastive_rows_kernel<<<(unsigned)n, LOSS_THREADS, smem, (cudaStream_t)stream>>>(
        lv1_dev, lv2_dev, (int)n, gamma, symmetric ? 1 : 0, (float)((double)weight / denom), scratch_dev, grad1_dev, grad2_dev);
    ASR_LAUNCH_CHECK();
    contrastive_sum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scratch_dev, (int)n, (double)weight / denom, loss_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

}  // extern "C"
