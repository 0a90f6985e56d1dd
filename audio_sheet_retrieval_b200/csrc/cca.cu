// CCA sufficient statistics (one fused Gram pass, fp64 accumulation) and the 32x32 solve
// (inverse square roots by symmetric Jacobi, SVD by one-sided Jacobi) in a single CTA.
//
// Replaces (reference, paths relative to its root):
//   audio_sheet_retrieval/utils/cca.py:25-53      means + three covariances (+ r*I)
//   audio_sheet_retrieval/utils/cca.py:199-211    inv(sqrtm(S)), T, svd(T), U = S11^-1/2 U', V = S22^-1/2 V'
//   audio_sheet_retrieval/models/lasagne_extensions/layers/cca.py:117-175  batch statistics,
//       eigh-based inverse square roots, eigh(TT'+rT I), eigh(T'T+rT I), sign fix   (mode 1)
//
// Multi-GPU: every rank accumulates its rows into the same 3136-value layout; one NCCL
// all-reduce(sum) of that buffer precedes asr_cca_solve (done by the Python host).
#include <math_constants.h>

#include "common.cuh"

namespace asr {

constexpr int CC_THREADS = 256;
constexpr int CC_ROWS = 128;
constexpr int CC_PART = 64 + 64 * 64;   // per-CTA partial: sum z (64) | Gram (64x64)
constexpr int CC_MAX_CTAS = 296;

// partial[(cta)][CC_PART].  z = [x - shift1, y - shift2]
__global__ void __launch_bounds__(CC_THREADS)
cca_gram_kernel(const float *__restrict__ h1, const float *__restrict__ h2, int64_t n, const float *__restrict__ shift1,
                const float *__restrict__ shift2, double *__restrict__ partial) {
    __shared__ __align__(16) float z[CC_ROWS][64];
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    double colsum = 0.0;
    const int64_t n_tiles = (n + CC_ROWS - 1) / CC_ROWS;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * CC_ROWS;
        __syncthreads();
        // coalesced load: 128 rows x 32 floats per view
        for (int e = tid; e < CC_ROWS * 32; e += CC_THREADS) {
            int r = e >> 5, c = e & 31;
            int64_t gr = r0 + r;
            float s1 = shift1 ? shift1[c] : 0.f, s2 = shift2 ? shift2[c] : 0.f;
            z[r][c] = gr < n ? h1[gr * 32 + c] - s1 : 0.f;
            z[r][32 + c] = gr < n ? h2[gr * 32 + c] - s2 : 0.f;
        }
        __syncthreads();
        const int rows = (int)min((int64_t)CC_ROWS, n - r0);
        for (int r = 0; r < rows; ++r) {
            float4 a4 = *reinterpret_cast<const float4 *>(&z[r][4 * ti]);
            float4 b4 = *reinterpret_cast<const float4 *>(&z[r][4 * tj]);
            double a[4] = {a4.x, a4.y, a4.z, a4.w};
            double b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (tid < 64)
            for (int r = 0; r < rows; ++r) colsum += (double)z[r][tid];
    }
    double *p = partial + (size_t)blockIdx.x * CC_PART;
    if (tid < 64) p[tid] = colsum;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[64 + (4 * ti + i) * 64 + 4 * tj + j] = acc[i][j];
}

// fixed-order reduction of the per-CTA partials, added to sums (ASR_CCA_NSUMS layout)
__global__ void cca_reduce_kernel(const double *__restrict__ partial, int n_part, double *__restrict__ sums, double n_rows) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o == ASR_CCA_NSUMS && n_rows > 0.0) sums[o] += n_rows;     // counted layout: the row count travels with the sums
    if (o >= ASR_CCA_NSUMS) return;
    int src;
    if (o < 64) src = o;                                      // sum x | sum y
    else {
        int m = (o - 64) >> 10, e = (o - 64) & 1023, i = e >> 5, j = e & 31;
        int gi = (m == 1) ? 32 + i : i;                       // xx: (i,j)  yy: (32+i,32+j)  xy: (i,32+j)
        int gj = (m == 0) ? j : 32 + j;
        src = 64 + gi * 64 + gj;
    }
    double s = 0.0;
    for (int c = 0; c < n_part; ++c) s += partial[(size_t)c * CC_PART + src];
    sums[o] += s;
}

// ---------------------------------------------------------------------------------
// single-CTA solve, 512 threads = 16 warps = 16 disjoint Jacobi pairs per step
// ---------------------------------------------------------------------------------
constexpr int SV_THREADS = 512;
typedef double Mat[32][33];     // padded rows

__device__ __forceinline__ void rr_pair(int step, int i, int &p, int &q) {   // round-robin tournament, n = 32
    if (i == 0) { p = 31; q = step % 31; }
    else { p = (step + i) % 31; q = (step + 31 - i) % 31; }
    if (p > q) { int t = p; p = q; q = t; }
}

// Symmetric eigen-decomposition A = V diag(w) V^T by cyclic two-sided Jacobi (A destroyed).
__device__ void jacobi_eigh(Mat A, Mat V, double *w, double *cs /*[16][2]*/, int tid) {
    const int pi = tid >> 5, r = tid & 31;
    for (int e = tid; e < 1024; e += SV_THREADS) V[e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
    __syncthreads();
    for (int sweep = 0; sweep < 30; ++sweep) {
        // convergence: off-diagonal mass relative to the diagonal
        __shared__ double off_s, diag_s;
        if (tid == 0) { off_s = 0.0; diag_s = 0.0; }
        __syncthreads();
        if (tid < 32) {
            double off = 0.0;
            for (int j = 0; j < 32; ++j) if (j != tid) off += A[tid][j] * A[tid][j];
            double dg = A[tid][tid] * A[tid][tid];
            for (int o = 16; o > 0; o >>= 1) {
                off += __shfl_xor_sync(0xffffffffu, off, o);
                dg += __shfl_xor_sync(0xffffffffu, dg, o);
            }
            if (tid == 0) { off_s = off; diag_s = dg; }
        }
        __syncthreads();
        if (off_s <= 1e-30 * diag_s || off_s == 0.0) break;
        for (int step = 0; step < 31; ++step) {
            int p, q;
            rr_pair(step, pi, p, q);
            if (r == 0) {
                double app = A[p][p], aqq = A[q][q], apq = A[p][q];
                double c = 1.0, s = 0.0;
                if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * sqrt(fabs(app * aqq))) {
                    double tau = (aqq - app) / (2.0 * apq);
                    double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                cs[2 * pi] = c; cs[2 * pi + 1] = s;
            }
            __syncthreads();
            const double c = cs[2 * pi], s = cs[2 * pi + 1];
            {   // columns: A <- A J, V <- V J
                double ap = A[r][p], aq = A[r][q];
                A[r][p] = c * ap - s * aq; A[r][q] = s * ap + c * aq;
                double vp = V[r][p], vq = V[r][q];
                V[r][p] = c * vp - s * vq; V[r][q] = s * vp + c * vq;
            }
            __syncthreads();
            {   // rows: A <- J^T A
                double ap = A[p][r], aq = A[q][r];
                A[p][r] = c * ap - s * aq; A[q][r] = s * ap + c * aq;
            }
            __syncthreads();
        }
    }
    if (tid < 32) w[tid] = A[tid][tid];
    __syncthreads();
}

// One-sided Jacobi SVD: G (in: T, out: U' * diag(sigma) columns), Vr accumulates right vectors.
__device__ void jacobi_svd(Mat G, Mat Vr, int tid) {
    const int pi = tid >> 5, r = tid & 31;
    for (int e = tid; e < 1024; e += SV_THREADS) Vr[e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
    __syncthreads();
    __shared__ int rotated;
    for (int sweep = 0; sweep < 40; ++sweep) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int step = 0; step < 31; ++step) {
            int p, q;
            rr_pair(step, pi, p, q);
            double gp = G[r][p], gq = G[r][q];
            double al = gp * gp, be = gq * gq, ga = gp * gq;
            for (int o = 16; o > 0; o >>= 1) {
                al += __shfl_xor_sync(0xffffffffu, al, o);
                be += __shfl_xor_sync(0xffffffffu, be, o);
                ga += __shfl_xor_sync(0xffffffffu, ga, o);
            }
            if (fabs(ga) > 1e-15 * sqrt(al * be) && fabs(ga) > 1e-300) {
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                G[r][p] = c * gp - s * gq; G[r][q] = s * gp + c * gq;
                double vp = Vr[r][p], vq = Vr[r][q];
                Vr[r][p] = c * vp - s * vq; Vr[r][q] = s * vp + c * vq;
                if (r == 0) rotated = 1;
            }
            __syncthreads();
        }
        if (!rotated) break;     // read after the barrier above; uniform
        __syncthreads();
    }
}

// C = A * B (32x32), optional transposes
__device__ void mm32(Mat C, Mat A, bool ta, Mat B, bool tb, int tid) {
    for (int e = tid; e < 1024; e += SV_THREADS) {
        int i = e >> 5, j = e & 31;
        double s = 0.0;
        for (int k = 0; k < 32; ++k) s += (ta ? A[k][i] : A[i][k]) * (tb ? B[j][k] : B[k][j]);
        C[i][j] = s;
    }
    __syncthreads();
}

// Si = V diag(w^-1/2) V^T
__device__ void inv_sqrt_from_eig(Mat Si, Mat V, const double *w, int tid) {
    for (int e = tid; e < 1024; e += SV_THREADS) {
        int i = e >> 5, j = e & 31;
        double s = 0.0;
        for (int k = 0; k < 32; ++k) s += V[i][k] * V[j][k] / sqrt(w[k]);
        Si[i][j] = s;
    }
    __syncthreads();
}

struct SolveSmem {
    Mat S11, S22, S12, W0, W1, S11i, S22i, T, G, Vr;
    double w[32], w2[32], mx[32], my[32], cs[32];
    int perm[32], perm2[32];
};

__global__ void __launch_bounds__(SV_THREADS, 1)
cca_solve_kernel(const double *__restrict__ sums, double n_total, const float *__restrict__ shift1,
                 const float *__restrict__ shift2, double r1, double r2, double rT, int mode, double *__restrict__ m1,
                 double *__restrict__ m2, double *__restrict__ U, double *__restrict__ V, double *__restrict__ sigma) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SolveSmem &sm = *reinterpret_cast<SolveSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const double n = n_total > 0.0 ? n_total : sums[ASR_CCA_NSUMS];     // counted layout: read the (all-reduced) row count
    if (tid < 32) {
        sm.mx[tid] = sums[tid] / n;
        sm.my[tid] = sums[32 + tid] / n;
        m1[tid] = sm.mx[tid] + (shift1 ? (double)shift1[tid] : 0.0);
        m2[tid] = sm.my[tid] + (shift2 ? (double)shift2[tid] : 0.0);
    }
    __syncthreads();
    for (int e = tid; e < 1024; e += SV_THREADS) {
        int i = e >> 5, j = e & 31;
        double d = (i == j) ? 1.0 : 0.0;
        sm.S11[i][j] = (sums[64 + e] - n * sm.mx[i] * sm.mx[j]) / (n - 1.0) + r1 * d;
        sm.S22[i][j] = (sums[64 + 1024 + e] - n * sm.my[i] * sm.my[j]) / (n - 1.0) + r2 * d;
        sm.S12[i][j] = (sums[64 + 2048 + e] - n * sm.mx[i] * sm.my[j]) / (n - 1.0);
    }
    __syncthreads();
    // S11^-1/2, S22^-1/2
    for (int e = tid; e < 1024; e += SV_THREADS) sm.W0[e >> 5][e & 31] = sm.S11[e >> 5][e & 31];
    __syncthreads();
    jacobi_eigh(sm.W0, sm.W1, sm.w, sm.cs, tid);
    inv_sqrt_from_eig(sm.S11i, sm.W1, sm.w, tid);
    for (int e = tid; e < 1024; e += SV_THREADS) sm.W0[e >> 5][e & 31] = sm.S22[e >> 5][e & 31];
    __syncthreads();
    jacobi_eigh(sm.W0, sm.W1, sm.w, sm.cs, tid);
    inv_sqrt_from_eig(sm.S22i, sm.W1, sm.w, tid);
    // T = S11i * S12 * S22i
    mm32(sm.W0, sm.S11i, false, sm.S12, false, tid);
    mm32(sm.T, sm.W0, false, sm.S22i, false, tid);

    if (mode == 0) {
        // ---- CCA.fit('svd'): T = U' diag(sigma) V'^T, sigma descending ----
        for (int e = tid; e < 1024; e += SV_THREADS) sm.G[e >> 5][e & 31] = sm.T[e >> 5][e & 31];
        __syncthreads();
        jacobi_svd(sm.G, sm.Vr, tid);
        if (tid < 32) {
            double s = 0.0;
            for (int rr = 0; rr < 32; ++rr) s += sm.G[rr][tid] * sm.G[rr][tid];
            sm.w[tid] = sqrt(s);
        }
        __syncthreads();
        if (tid < 32) {   // rank by (sigma desc, index asc)
            int rk = 0;
            for (int j = 0; j < 32; ++j) rk += (sm.w[j] > sm.w[tid] || (sm.w[j] == sm.w[tid] && j < tid)) ? 1 : 0;
            sm.perm[rk] = tid;
        }
        __syncthreads();
        // W0 = U' (normalised, permuted), W1 = V' (permuted)
        for (int e = tid; e < 1024; e += SV_THREADS) {
            int i = e >> 5, j = e & 31, src = sm.perm[j];
            double sg = sm.w[src];
            sm.W0[i][j] = sg > 1e-300 ? sm.G[i][src] / sg : 0.0;
            sm.W1[i][j] = sm.Vr[i][src];
        }
        __syncthreads();
        if (tid < 32) sigma[tid] = sm.w[sm.perm[tid]];
        mm32(sm.G, sm.S11i, false, sm.W0, false, tid);     // U = S11i U'
        mm32(sm.Vr, sm.S22i, false, sm.W1, false, tid);    // V = S22i V'
        for (int e = tid; e < 1024; e += SV_THREADS) { U[e] = sm.G[e >> 5][e & 31]; V[e] = sm.Vr[e >> 5][e & 31]; }
    } else {
        // ---- CCALayer train forward: eigh(TT'+rT I) -> E, eigh(T'T+rT I) -> F, ascending ----
        mm32(sm.W0, sm.T, false, sm.T, true, tid);          // T T^T
        for (int e = tid; e < 1024; e += SV_THREADS) if ((e >> 5) == (e & 31)) sm.W0[e >> 5][e & 31] += rT;
        __syncthreads();
        jacobi_eigh(sm.W0, sm.G, sm.w, sm.cs, tid);         // G = E (unsorted)
        mm32(sm.W0, sm.T, true, sm.T, false, tid);          // T^T T
        for (int e = tid; e < 1024; e += SV_THREADS) if ((e >> 5) == (e & 31)) sm.W0[e >> 5][e & 31] += rT;
        __syncthreads();
        jacobi_eigh(sm.W0, sm.Vr, sm.w2, sm.cs, tid);       // Vr = F (unsorted)
        if (tid < 32) {   // ascending, ties by index
            int rk = 0, rk2 = 0;
            for (int j = 0; j < 32; ++j) {
                rk += (sm.w[j] < sm.w[tid] || (sm.w[j] == sm.w[tid] && j < tid)) ? 1 : 0;
                rk2 += (sm.w2[j] < sm.w2[tid] || (sm.w2[j] == sm.w2[tid] && j < tid)) ? 1 : 0;
            }
            sm.perm[rk] = tid;
            sm.perm2[rk2] = tid;
        }
        __syncthreads();
        for (int e = tid; e < 1024; e += SV_THREADS) {
            int i = e >> 5, j = e & 31;
            sm.W0[i][j] = sm.G[i][sm.perm[j]];
            sm.W1[i][j] = sm.Vr[i][sm.perm2[j]];
        }
        __syncthreads();
        if (tid < 32) {
            double e1 = sm.w[sm.perm[tid]];
            e1 = fmin(fmax(e1, 1e-7), 1.0);
            sigma[tid] = sqrt(e1);                           // corr (cca.py:163-166)
        }
        mm32(sm.G, sm.S11i, false, sm.W0, false, tid);      // U = S11si E
        mm32(sm.Vr, sm.S22i, false, sm.W1, false, tid);     // V = S22si F
        // sign fix: s = sgn(diag(U^T S12 V)); U *= s
        mm32(sm.W0, sm.S12, false, sm.Vr, false, tid);      // S12 V
        if (tid < 32) {
            double d = 0.0;
            for (int rr = 0; rr < 32; ++rr) d += sm.G[rr][tid] * sm.W0[rr][tid];
            sm.cs[tid] = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0);
        }
        __syncthreads();
        for (int e = tid; e < 1024; e += SV_THREADS) {
            U[e] = sm.G[e >> 5][e & 31] * sm.cs[e & 31];
            V[e] = sm.Vr[e >> 5][e & 31];
        }
    }
}

// ---------------------------------------------------------------------------------
// CCALayer training backward (layers/cca.py:91-203 differentiated; the reference gets it from Theano's autodiff:
// EighGrad for the four eigendecompositions, zero gradient through sgn and outside the clip range).  ALPHA = 1.
// Forward, with Xc = H1 - mean, Yc = H2 - mean (m x 32):
//   S11 = Xc'Xc/(m-1) + r1 I, S22, S12 = Xc'Yc/(m-1);  S11si = S11^-1/2 (eigh d1, A1), S22si (d2, A2)
//   T = S11si S12 S22si;  (E1, E) = eigh(TT' + rT I), (E2, F) = eigh(T'T + rT I), ascending
//   U = S11si E diag(s), V = S22si F, s = sgn(diag(E' S11si S12 V));  lv1 = Xc U, lv2 = Yc V, corr = sqrt(clip(E1))
// Backward for G1 = dL/dlv1, G2 = dL/dlv2, g_corr = dL/dcorr, everything 32 x 32 in fp64 in ONE CTA:
//   dU = Xc'G1, dV = Yc'G2 (from two more Gram passes);  dU0 = dU diag(s)
//   dS11si = dU0 E', dS22si = dV F', dE = S11si dU0, dF = S22si dV
//   dM1 = sym(E (K1 o (E'dE) + diag(dE1)) E'), K_ij = 1/(lam_j - lam_i);  dM2 likewise from (E2, F, dF)
//   dT = 2 dM1 T + 2 T dM2;  dS11si += dT S22si S12';  dS22si += S12' S11si dT;  dS12 = S11si dT S22si
//   dS11 = A1 (Phi1 o (A1' sym(dS11si) A1)) A1', Phi_ij = (f(d_i) - f(d_j))/(d_i - d_j), f = x^-1/2;  dS22 likewise
//   dXc = G1 U' + (2 Xc dS11 + Yc dS12')/(m-1), dYc = G2 V' + (2 Yc dS22 + Xc dS12)/(m-1), minus their column means
// The kernel leaves [U | V | R11 = 2 dS11/(m-1) | R22 | R12 = dS12/(m-1) | mean1 | mean2 | mean G1 | mean G2] for the
// row pass (cca_backward_rows_kernel).  oracle/cca.py: cca_layer_train_backward is the same chain in NumPy, pinned by
// central differences of the forward.
// ---------------------------------------------------------------------------------
constexpr int BW_WORDS = 5 * 1024 + 4 * 32;

struct BackSmem {
    Mat S12, A1, A2, S11i, S22i, T, E, F, U0, V, W0, W1, W2, dS11i, dS22i, dM, dT;
    double d1[32], d2[32], e1[32], e2[32], sg[32], mx[32], my[32], gx[32], gy[32], cs[32], wtmp[32];
    int perm[32];
};

// sort eigenpairs ascending (ties by index): Q <- Q[:, perm], w <- w[perm]; W is scratch
__device__ void sort_eig(Mat Q, double *w, Mat W, double *wtmp, int *perm, int tid) {
    if (tid < 32) {
        int rk = 0;
        for (int j = 0; j < 32; ++j) rk += (w[j] < w[tid] || (w[j] == w[tid] && j < tid)) ? 1 : 0;
        perm[rk] = tid;
    }
    __syncthreads();
    for (int e = tid; e < 1024; e += SV_THREADS) W[e >> 5][e & 31] = Q[e >> 5][perm[e & 31]];
    if (tid < 32) wtmp[tid] = w[perm[tid]];
    __syncthreads();
    for (int e = tid; e < 1024; e += SV_THREADS) Q[e >> 5][e & 31] = W[e >> 5][e & 31];
    if (tid < 32) w[tid] = wtmp[tid];
    __syncthreads();
}

// out = sym(Q (K o (Q' Qbar) + diag(lam_bar)) Q'), K_ij = 1 / (lam_j - lam_i); W0, W1 scratch
__device__ void eigvec_grad(Mat out, Mat Q, const double *lam, Mat Qbar, const double *lam_bar, Mat W0, Mat W1, int tid) {
    mm32(W0, Q, true, Qbar, false, tid);
    for (int e = tid; e < 1024; e += SV_THREADS) {
        const int i = e >> 5, j = e & 31;
        const double diff = lam[j] - lam[i];
        double v = (i != j && diff != 0.0) ? W0[i][j] / diff : 0.0;
        if (i == j && lam_bar) v += lam_bar[i];
        W0[i][j] = v;
    }
    __syncthreads();
    mm32(W1, Q, false, W0, false, tid);
    mm32(W0, W1, false, Q, true, tid);
    for (int e = tid; e < 1024; e += SV_THREADS) out[e >> 5][e & 31] = 0.5 * (W0[e >> 5][e & 31] + W0[e & 31][e >> 5]);
    __syncthreads();
}

// out = Q (Phi o (Q' sym(G) Q)) Q' for f(x) = x^-1/2 (Daleckii-Krein); W0, W1 scratch
__device__ void inv_sqrt_grad(Mat out, Mat Q, const double *lam, Mat G, Mat W0, Mat W1, int tid) {
    for (int e = tid; e < 1024; e += SV_THREADS) W0[e >> 5][e & 31] = 0.5 * (G[e >> 5][e & 31] + G[e & 31][e >> 5]);
    __syncthreads();
    mm32(W1, Q, true, W0, false, tid);
    mm32(W0, W1, false, Q, false, tid);
    double lmax = 0.0;
    for (int i = 0; i < 32; ++i) lmax = fmax(lmax, lam[i]);
    for (int e = tid; e < 1024; e += SV_THREADS) {
        const int i = e >> 5, j = e & 31;
        const double diff = lam[i] - lam[j];
        double phi;
        if (fabs(diff) > 1e-12 * lmax) phi = (1.0 / sqrt(lam[i]) - 1.0 / sqrt(lam[j])) / diff;
        else { const double a = 0.5 * (lam[i] + lam[j]); phi = -0.5 / (a * sqrt(a)); }
        W0[i][j] *= phi;
    }
    __syncthreads();
    mm32(W1, Q, false, W0, false, tid);
    mm32(out, W1, false, Q, true, tid);
}

__global__ void __launch_bounds__(SV_THREADS, 1)
cca_backward_kernel(const double *__restrict__ sums, const double *__restrict__ sums_g1, const double *__restrict__ sums_g2,
                    double n, double r1, double r2, double rT, const double *__restrict__ g_corr, double *__restrict__ bw) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    BackSmem &sm = *reinterpret_cast<BackSmem *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid < 32) {
        sm.mx[tid] = sums[tid] / n;
        sm.my[tid] = sums[32 + tid] / n;
        sm.gx[tid] = sums_g1[32 + tid] / n;            // accumulate(H1, G1): "y" is G1
        sm.gy[tid] = sums_g2[32 + tid] / n;
    }
    __syncthreads();
    // ---- forward (as cca_solve_kernel mode 1, keeping the eigen-systems) ----
    for (int e = tid; e < 1024; e += SV_THREADS) {
        const int i = e >> 5, j = e & 31;
        const double dg = (i == j) ? 1.0 : 0.0;
        sm.W0[i][j] = (sums[64 + e] - n * sm.mx[i] * sm.mx[j]) / (n - 1.0) + r1 * dg;
        sm.W1[i][j] = (sums[64 + 1024 + e] - n * sm.my[i] * sm.my[j]) / (n - 1.0) + r2 * dg;
        sm.S12[i][j] = (sums[64 + 2048 + e] - n * sm.mx[i] * sm.my[j]) / (n - 1.0);
    }
    __syncthreads();
    jacobi_eigh(sm.W0, sm.A1, sm.d1, sm.cs, tid);
    inv_sqrt_from_eig(sm.S11i, sm.A1, sm.d1, tid);
    jacobi_eigh(sm.W1, sm.A2, sm.d2, sm.cs, tid);
    inv_sqrt_from_eig(sm.S22i, sm.A2, sm.d2, tid);
    mm32(sm.W0, sm.S11i, false, sm.S12, false, tid);
    mm32(sm.T, sm.W0, false, sm.S22i, false, tid);
    mm32(sm.W0, sm.T, false, sm.T, true, tid);
    for (int e = tid; e < 32; e += SV_THREADS) sm.W0[e][e] += rT;
    __syncthreads();
    jacobi_eigh(sm.W0, sm.E, sm.e1, sm.cs, tid);
    sort_eig(sm.E, sm.e1, sm.W0, sm.wtmp, sm.perm, tid);
    mm32(sm.W0, sm.T, true, sm.T, false, tid);
    for (int e = tid; e < 32; e += SV_THREADS) sm.W0[e][e] += rT;
    __syncthreads();
    jacobi_eigh(sm.W0, sm.F, sm.e2, sm.cs, tid);
    sort_eig(sm.F, sm.e2, sm.W0, sm.wtmp, sm.perm, tid);
    mm32(sm.U0, sm.S11i, false, sm.E, false, tid);
    mm32(sm.V, sm.S22i, false, sm.F, false, tid);
    mm32(sm.W0, sm.S12, false, sm.V, false, tid);
    if (tid < 32) {
        double d = 0.0;
        for (int rr = 0; rr < 32; ++rr) d += sm.U0[rr][tid] * sm.W0[rr][tid];
        sm.sg[tid] = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0);
    }
    __syncthreads();
    // ---- backward ----
    // W0 = dU0 = (Xc'G1) diag(s), W1 = dV = Yc'G2   (raw moments minus n mean mean')
    for (int e = tid; e < 1024; e += SV_THREADS) {
        const int i = e >> 5, j = e & 31;
        sm.W0[i][j] = (sums_g1[64 + 2048 + e] - n * sm.mx[i] * sm.gx[j]) * sm.sg[j];
        sm.W1[i][j] = sums_g2[64 + 2048 + e] - n * sm.my[i] * sm.gy[j];
    }
    __syncthreads();
    mm32(sm.dS11i, sm.W0, false, sm.E, true, tid);            // dU0 E'
    mm32(sm.dS22i, sm.W1, false, sm.F, true, tid);            // dV F'
    mm32(sm.W2, sm.S11i, false, sm.W0, false, tid);           // dE = S11si dU0
    if (tid < 32) {                                           // dL/dE1 from dL/dcorr, corr = sqrt(clip(E1, 1e-7, 1))
        const double ev = sm.e1[tid];
        sm.wtmp[tid] = (g_corr && ev > 1e-7 && ev < 1.0) ? g_corr[tid] * 0.5 / sqrt(ev) : 0.0;
    }
    __syncthreads();
    eigvec_grad(sm.dM, sm.E, sm.e1, sm.W2, sm.wtmp, sm.W0, sm.dT, tid);          // dM1 (W0, dT scratch; W1 = dV still needed)
    mm32(sm.dT, sm.dM, false, sm.T, false, tid);              // dM1 T
    mm32(sm.W2, sm.S22i, false, sm.W1, false, tid);           // dF = S22si dV
    eigvec_grad(sm.dM, sm.F, sm.e2, sm.W2, nullptr, sm.W0, sm.W1, tid);          // dM2
    mm32(sm.W0, sm.T, false, sm.dM, false, tid);              // T dM2
    for (int e = tid; e < 1024; e += SV_THREADS) sm.dT[e >> 5][e & 31] = 2.0 * (sm.dT[e >> 5][e & 31] + sm.W0[e >> 5][e & 31]);
    __syncthreads();
    mm32(sm.W0, sm.dT, false, sm.S22i, false, tid);           // dT S22si
    mm32(sm.W1, sm.W0, false, sm.S12, true, tid);             // dT S22si S12'
    for (int e = tid; e < 1024; e += SV_THREADS) sm.dS11i[e >> 5][e & 31] += sm.W1[e >> 5][e & 31];
    __syncthreads();
    mm32(sm.W1, sm.S11i, false, sm.W0, false, tid);           // dS12 = S11si dT S22si  -> W1
    mm32(sm.W0, sm.S11i, false, sm.dT, false, tid);           // S11si dT
    mm32(sm.W2, sm.S12, true, sm.W0, false, tid);             // S12' S11si dT
    for (int e = tid; e < 1024; e += SV_THREADS) sm.dS22i[e >> 5][e & 31] += sm.W2[e >> 5][e & 31];
    __syncthreads();
    // outputs: U | V | R11 | R22 | R12 | means
    for (int e = tid; e < 1024; e += SV_THREADS) {
        bw[e] = sm.U0[e >> 5][e & 31] * sm.sg[e & 31];
        bw[1024 + e] = sm.V[e >> 5][e & 31];
        bw[4096 + e] = sm.W1[e >> 5][e & 31] / (n - 1.0);
    }
    __syncthreads();
    inv_sqrt_grad(sm.dM, sm.A1, sm.d1, sm.dS11i, sm.W0, sm.W2, tid);             // dS11
    for (int e = tid; e < 1024; e += SV_THREADS) bw[2048 + e] = 2.0 * sm.dM[e >> 5][e & 31] / (n - 1.0);
    __syncthreads();
    inv_sqrt_grad(sm.dM, sm.A2, sm.d2, sm.dS22i, sm.W0, sm.W2, tid);             // dS22
    for (int e = tid; e < 1024; e += SV_THREADS) bw[3072 + e] = 2.0 * sm.dM[e >> 5][e & 31] / (n - 1.0);
    if (tid < 32) {
        bw[5120 + tid] = sm.mx[tid];
        bw[5152 + tid] = sm.my[tid];
        bw[5184 + tid] = sm.gx[tid];
        bw[5216 + tid] = sm.gy[tid];
    }
}

// dH1[i] = (G1[i] - mean G1) U' + Xc[i] R11 + Yc[i] R12',  dH2[i] = (G2[i] - mean G2) V' + Yc[i] R22 + Xc[i] R12
__global__ void __launch_bounds__(256)
cca_backward_rows_kernel(const float *__restrict__ h1, const float *__restrict__ h2, const float *__restrict__ g1,
                         const float *__restrict__ g2, int64_t n, const double *__restrict__ bw, float *__restrict__ dh1,
                         float *__restrict__ dh2) {
    __shared__ double M[5][32][33];
    __shared__ double mean[4][32];
    for (int e = threadIdx.x; e < 5 * 1024; e += 256) M[e >> 10][(e >> 5) & 31][e & 31] = bw[e];
    if (threadIdx.x < 128) mean[threadIdx.x >> 5][threadIdx.x & 31] = bw[5120 + threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * 8) {
        const double x = (double)h1[row * 32 + lane] - mean[0][lane], y = (double)h2[row * 32 + lane] - mean[1][lane];
        const double a = (double)g1[row * 32 + lane] - mean[2][lane], b = (double)g2[row * 32 + lane] - mean[3][lane];
        double o1 = 0.0, o2 = 0.0;
        for (int k = 0; k < 32; ++k) {
            const double xk = __shfl_sync(0xffffffffu, x, k), yk = __shfl_sync(0xffffffffu, y, k);
            const double ak = __shfl_sync(0xffffffffu, a, k), bk = __shfl_sync(0xffffffffu, b, k);
            o1 += ak * M[0][lane][k] + xk * M[2][k][lane] + yk * M[4][lane][k];      // a U' + x R11 + y R12'
            o2 += bk * M[1][lane][k] + yk * M[3][k][lane] + xk * M[4][k][lane];      // b V' + y R22 + x R12
        }
        dh1[row * 32 + lane] = (float)o1;
        dh2[row * 32 + lane] = (float)o2;
    }
}

static double *g_partial[ASR_MAX_DEVICES] = {nullptr};     // per device (allocated on first use)

}  // namespace asr

using namespace asr;

extern "C" {

static int cca_accumulate(const float *h1_dev, const float *h2_dev, int64_t n, const float *shift1_dev,
                          const float *shift2_dev, double *sums_dev, bool counted, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(h1_dev && h2_dev && sums_dev, "NULL buffer");
    ASR_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return ASR_OK;
    const int dev = current_device();
    ASR_CHECK_ARG(dev >= 0 && dev < ASR_MAX_DEVICES, "device ordinal out of range");
    if (!g_partial[dev]) ASR_CUDA(cudaMalloc(&g_partial[dev], sizeof(double) * CC_PART * CC_MAX_CTAS));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n_tiles = (n + CC_ROWS - 1) / CC_ROWS;
    int grid = (int)std::min<int64_t>(n_tiles, std::min(CC_MAX_CTAS, 2 * sm_count()));
    cca_gram_kernel<<<grid, CC_THREADS, 0, st>>>(h1_dev, h2_dev, n, shift1_dev, shift2_dev, g_partial[dev]);
    ASR_LAUNCH_CHECK();
    cca_reduce_kernel<<<(ASR_CCA_NSUMS + 1 + 127) / 128, 128, 0, st>>>(g_partial[dev], grid, sums_dev, counted ? (double)n : 0.0);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

int asr_cca_accumulate(const float *h1_dev, const float *h2_dev, int64_t n, const float *shift1_dev,
                       const float *shift2_dev, double *sums_dev, void *stream) {
    return cca_accumulate(h1_dev, h2_dev, n, shift1_dev, shift2_dev, sums_dev, false, stream);
}

int asr_cca_accumulate_counted(const float *h1_dev, const float *h2_dev, int64_t n, const float *shift1_dev,
                               const float *shift2_dev, double *sums_dev, void *stream) {
    return cca_accumulate(h1_dev, h2_dev, n, shift1_dev, shift2_dev, sums_dev, true, stream);
}

int asr_cca_solve(const double *sums_dev, int64_t n_total, const float *shift1_dev, const float *shift2_dev, double r1,
                  double r2, double rT, int mode, double *m1_dev, double *m2_dev, double *U_dev, double *V_dev,
                  double *sigma_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(sums_dev && m1_dev && m2_dev && U_dev && V_dev && sigma_dev, "NULL buffer");
    ASR_CHECK_ARG(n_total >= 2 || n_total == ASR_CCA_COUNT_ON_DEVICE, "need at least 2 samples");
    ASR_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (svd) or 1 (layer)");
    static bool attr_done[ASR_MAX_DEVICES] = {false};
    const int dev = current_device();
    ASR_CHECK_ARG(dev >= 0 && dev < ASR_MAX_DEVICES, "device ordinal out of range");
    if (!attr_done[dev]) {
        ASR_CUDA(cudaFuncSetAttribute(cca_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(SolveSmem)));
        attr_done[dev] = true;
    }
    cca_solve_kernel<<<1, SV_THREADS, sizeof(SolveSmem), (cudaStream_t)stream>>>(
        sums_dev, (double)n_total, shift1_dev, shift2_dev, r1, r2, rT, mode, m1_dev, m2_dev, U_dev, V_dev, sigma_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}


static double *g_bw[ASR_MAX_DEVICES] = {nullptr};       // per device: three sums buffers + the backward kernel's outputs

int asr_cca_layer_backward(const float *h1_dev, const float *h2_dev, const float *g1_dev, const float *g2_dev, int64_t n,
                           double r1, double r2, double rT, const double *g_corr_dev, float *dh1_dev, float *dh2_dev,
                           void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(h1_dev && h2_dev && g1_dev && g2_dev && dh1_dev && dh2_dev, "NULL buffer");
    ASR_CHECK_ARG(n >= 2, "need at least 2 samples");
    const int dev = current_device();
    ASR_CHECK_ARG(dev >= 0 && dev < ASR_MAX_DEVICES, "device ordinal out of range");
    constexpr int NS = ASR_CCA_NSUMS + 8;
    if (!g_bw[dev]) ASR_CUDA(cudaMalloc(&g_bw[dev], sizeof(double) * (3 * NS + BW_WORDS)));
    static bool attr_done[ASR_MAX_DEVICES] = {false};
    if (!attr_done[dev]) {
        ASR_CUDA(cudaFuncSetAttribute(cca_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BackSmem)));
        attr_done[dev] = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *s0 = g_bw[dev], *s1 = s0 + NS, *s2 = s1 + NS, *bw = s2 + NS;
    ASR_CUDA(cudaMemsetAsync(s0, 0, sizeof(double) * 3 * NS, st));
    if ((rc = cca_accumulate(h1_dev, h2_dev, n, nullptr, nullptr, s0, false, stream))) return rc;
    if ((rc = cca_accumulate(h1_dev, g1_dev, n, nullptr, nullptr, s1, false, stream))) return rc;     // H1'G1, sum G1
    if ((rc = cca_accumulate(h2_dev, g2_dev, n, nullptr, nullptr, s2, false, stream))) return rc;     // H2'G2, sum G2
    cca_backward_kernel<<<1, SV_THREADS, sizeof(BackSmem), st>>>(s0, s1, s2, (double)n, r1, r2, rT, g_corr_dev, bw);
    ASR_LAUNCH_CHECK();
    const int grid = (int)std::min<int64_t>((n + 7) / 8, 4 * (int64_t)sm_count());
    cca_backward_rows_kernel<<<grid, 256, 0, st>>>(h1_dev, h2_dev, g1_dev, g2_dev, n, bw, dh1_dev, dh2_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

}  // extern "C"
