/* asr_b200.h - C ABI of the B200-native audio-sheet retrieval hot path.
 *
 * The reference (CPJKU/audio_sheet_retrieval) is pure Python on Theano/Lasagne and
 * has NO FFI of its own; these entry points sit *behind* its Python interface
 * (RetrievalWrapper / eval_retrieval / CCA / AudioSheetServer).  Every entry point
 * cites the reference interface it replaces (paths relative to the reference
 * root; asr/ = audio_sheet_retrieval/).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; asr_last_error() returns a
 *     thread-local message for the last failing call.
 *   - `*_dev` pointers are CUDA device pointers owned by the caller; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  No call
 *     synchronises the host unless its name ends in `_host`.
 *   - opaque handles own only what they allocate at create time (weights,
 *     activation arena, scratch, the DB's normalised rows and query workspace);
 *     the only first-use allocations are the staging buffers of the `_host`
 *     entry and the layer-0 buffer of an encoder whose fusion is switched off.
 *     Nothing is allocated in steady state.
 *   - handles are bound to the device that was current at create and are not
 *     thread-safe: one stream at a time per handle.
 *   - sm_100a only.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with ASR_ERR_CUDA.
 */
#ifndef ASR_B200_H
#define ASR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_OK 0
#define ASR_ERR_ARG (-1)
#define ASR_ERR_CUDA (-2)
#define ASR_ERR_UNSUPPORTED (-3)

#define ASR_DIM 32        /* DIM_LATENT, asr/models/mutopia_ccal_cont.py:36 */
#define ASR_N_LAYERS 9    /* 8 conv3x3+BN+ELU and the 1x1 conv+BN head per view */
#define ASR_MAX_K 128     /* largest n_candidates / top-k supported by asr_topk */

/* input element types for the encoders */
#define ASR_IN_F32 0      /* float32, the dtype the reference pools hand out (data_pools.py:203-228) */
#define ASR_IN_U8 1       /* uint8 sheet images (what the float32 arrays hold: 0..255) */

/* prepare modes (model.prepare) */
#define ASR_PREP_NONE 0     /* view 2 is never prepared (retrieval_wrapper.py:74-77) */
#define ASR_PREP_SCALE 1    /* x/255                     asr/models/mutopia_ccal_cont.py:170-190 */
#define ASR_PREP_SCALE_HALF 2 /* x/255 then resize to half (2x2 box mean)  ..._rsz.py:170-190 */

/* which kernels run the 3x3 layers */
#define ASR_PATH_TCGEN05 0  /* bf16 implicit GEMM on tcgen05/TMEM, TMA-fed (product path) */
#define ASR_PATH_FP32 1     /* fp32 CUDA-core kernels (on-device debug reference) */

const char *asr_last_error(void);
int asr_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t asr_launch_count(void);

/* ------------------------------------------------------------------------- *
 * Encoders.  Replaces the compiled Theano functions of one branch:
 *   asr/retrieval_wrapper.py:33-38 (compute_v1_latent / compute_v2_latent),
 *   graph asr/models/mutopia_ccal_cont{,_rsz}.py:64-145,
 *   CCALayer deterministic branch + LengthNormLayer (layers/cca.py:184-203, 39-40),
 *   and the pre-CCA latent functions of asr/refine_cca.py:86-89.
 * ------------------------------------------------------------------------- */
typedef struct asr_encoder asr_encoder_t;

typedef struct {
    int in_h, in_w;                 /* raw input size, e.g. 160x200 (view 1) or 92x42 (view 2) */
    int prepare;                    /* ASR_PREP_* */
    int flip_filters;               /* 0 = cross-correlation (cuDNN layer, shipped weights), 1 = true convolution */
    int channels[ASR_N_LAYERS];     /* output channels per layer; last must be ASR_DIM */
    /* host pointers, Lasagne get_all_param_values order (pickle layout, SURVEY.md 8b face 3) */
    const float *W[ASR_N_LAYERS];       /* (cout, cin, k, k), k = 3 for layers 0..7, 1 for layer 8 */
    const float *beta[ASR_N_LAYERS];
    const float *gamma[ASR_N_LAYERS];
    const float *mean[ASR_N_LAYERS];
    const float *inv_std[ASR_N_LAYERS];
    const float *cca_mean;          /* (32,)   mean1 or mean2 */
    const float *cca_proj;          /* (32,32) U or V, row-major (in, out) */
} asr_encoder_desc;

/* max_batch = largest number of samples one asr_encoder_embed call may carry */
int asr_encoder_create(asr_encoder_t **out, const asr_encoder_desc *desc, int max_batch);
int asr_encoder_destroy(asr_encoder_t *enc);
/* overwrite the CCA projection after a refit (asr/refine_cca.py:104-107) */
int asr_encoder_set_cca(asr_encoder_t *enc, const float *cca_mean, const float *cca_proj);

/* x_dev: (n, 1, in_h, in_w) of x_dtype.  codes_dev (n,32): unit-norm CCA codes (may be NULL).
 * latents_dev (n,32): pre-CCA latents (may be NULL).  n <= max_batch. */
int asr_encoder_embed(asr_encoder_t *enc, const void *x_dev, int x_dtype, int64_t n,
                      float *codes_dev, float *latents_dev, int path, void *stream);
/* Host buffers in, host buffers out (what RetrievalWrapper.compute_view_k does,
 * asr/retrieval_wrapper.py:47-77): chunks of max_batch, copy/compute overlap; pinned inputs / outputs are used where
 * they lie, pageable ones go through pinned staging slots.  Everything it needs (two input slots and two result slots of
 * max_batch rows, two streams, events) is allocated by the FIRST call of a handle and never again, whatever n.  Concurrent
 * calls on DIFFERENT handles from several host threads (the two branches of a pair) overlap each other's copies and kernels.
 * Blocks (on a blocking-sync event, not a spin) until the results are in
 * codes_host / latents_host. */
int asr_encoder_embed_host(asr_encoder_t *enc, const void *x_host, int x_dtype, int64_t n,
                           float *codes_host, float *latents_host, int path);
/* Cut n windows out of one unrolled sheet image / spectrogram resident on the device:
 * out[i] = src[r0 : r0+win_h, starts[i] : starts[i]+win_w]  (the slicing loops of
 * asr/audio_sheet_server.py:216-223, 260-271, 421-428, 465-477).  src (src_h, src_w) and out
 * (n, win_h, win_w) have element type `dtype` (ASR_IN_F32 / ASR_IN_U8); every start must satisfy
 * 0 <= start <= src_w - win_w (not checked on the device). */
int asr_extract_windows(const void *src_dev, int dtype, int src_h, int src_w, const int32_t *starts_dev, int n, int r0,
                        int win_h, int win_w, void *out_dev, void *stream);
/* debug: copy layer `layer`'s activation (after ELU/pool) of the last embed call to host as
 * NCHW float32 (n, c, h, w); returns c,h,w through the out params. */
int asr_encoder_debug_activation(asr_encoder_t *enc, int layer, int path, int64_t n, float *out_host,
                                 int *c, int *h, int *w);
/* Device timing for the roofline report: when enabled, every asr_encoder_embed on the tcgen05 path
 * records CUDA events on its stream around [layer 0] [layers 1..7 = the tcgen05 kernel] [head];
 * get_timing synchronises on the last event, returns the summed milliseconds and resets. */
int asr_encoder_set_timing(asr_encoder_t *enc, int enable);
int asr_encoder_get_timing(asr_encoder_t *enc, double *ms_layer0, double *ms_conv_tc, double *ms_head, int64_t *n_calls);
/* Kernel fusion switches (bit mask).  Bit 0: layers 0 + 1 run as ONE kernel whose layer-0 output stays in shared
 * memory (default on where the geometry qualifies: 12 first-layer channels, width >= 126, no box filter -- the
 * sheet branch of asr/models/mutopia_ccal_cont.py:75-79).  Switching it off restores one launch per layer (and
 * makes layer 0 visible to asr_encoder_debug_activation).  Bit 1: layers 2 + 3 run as ONE kernel whose layer-2
 * output only exists as a ring of image rows in shared memory (default on where the geometry qualifies: 12 -> 24
 * -> 24 channels, width 90..128, height a multiple of 8 -- the sheet branch of the same model, :80-84).
 * get_fusion returns the mask in effect. */
int asr_encoder_set_fusion(asr_encoder_t *enc, int mask);
int asr_encoder_get_fusion(const asr_encoder_t *enc);
/* algorithmic FLOPs per sample of this branch (2*MACs of the nine convolutions) */
double asr_encoder_flops_per_sample(const asr_encoder_t *enc);

/* ------------------------------------------------------------------------- *
 * Retrieval.  Replaces cdist(...,'cosine') + argsort of
 *   asr/audio_sheet_server.py:530-563 (_retrieve_*_ids), the loop at :230-234,
 *   and asr/utils/train_dcca_pool.py:39-74 (eval_retrieval ranking).
 * Scores follow the pinned-order fp32 definition in oracle/search.py; order is
 * (score desc, index asc); the distance matrix is never written to memory.
 * ------------------------------------------------------------------------- */
typedef struct asr_db asr_db_t;

/* codes_dev: (n, 32) float32 row-major, 128-byte aligned, caller-owned, kept alive and NOT modified while the
 * handle exists.  idx_base: global index of row 0 (this shard's offset in a sharded DB).
 * Everything the handle ever needs is allocated HERE (asr_topk / asr_rank_of_target never allocate):
 *   - 64 MB of scratch for per-slice partial lists,
 *   - a query workspace of max_queries rows (more queries per call are processed in chunks of that size) and
 *     (8 * max_queries + 262144) words in which the DB slices of a query share their threshold bounds,
 *   - the rows normalised with the pinned definition, which cosine queries stream:
 *       flags = 0                        an owned copy, n * 128 bytes (a 1e8-row DB: 12.8 GB + 12.8 GB)
 *       ASR_DB_NORMALISE_IN_PLACE        the caller donates its buffer: rows are overwritten with their normalised
 *                                        form, no second copy; only cosine (normalise != 0) queries afterwards
 *       ASR_DB_NO_COSINE_COPY            nothing: cosine queries normalise rows in-kernel (exact kernel only,
 *                                        the tensor-core pre-filter needs the normalised rows)
 * The create call runs one kernel on the default stream and synchronises; the handle is bound to the current
 * device.  One stream at a time per handle: its scratch is shared by consecutive calls (calls on the SAME stream
 * serialise naturally; concurrent use from two streams needs two handles).  Entry points may be called from any
 * host thread: they switch to the handle's device for the call and leave the thread on a device the process already
 * uses (a fresh thread's implicit device 0 is not touched, so no context is ever created as a side effect). */
#define ASR_DB_NORMALISE_IN_PLACE 1
#define ASR_DB_NO_COSINE_COPY 2
int asr_db_create_ex(asr_db_t **out, const float *codes_dev, int64_t n, int64_t idx_base, int flags, int64_t max_queries);
/* = asr_db_create_ex(out, codes_dev, n, idx_base, 0, 16384) */
int asr_db_create(asr_db_t **out, const float *codes_dev, int64_t n, int64_t idx_base);
int asr_db_destroy(asr_db_t *db);

/* q_dev (nq,32).  out_score_dev (nq,k) float32, out_idx_dev (nq,k) int64; slots beyond the
 * DB size hold (-inf, -1).  normalise != 0 L2-normalises queries and DB rows in-kernel. */
int asr_topk(asr_db_t *db, const float *q_dev, int64_t nq, int k, int normalise,
             float *out_score_dev, int64_t *out_idx_dev, void *stream);
/* validation hook: approximate (tf32 tensor-core) scores of the first 128 queries x first 256 DB rows as the
 * pre-filter of asr_topk sees them (out_host: 128*256 floats, row = query) */
int asr_debug_tc_scores(asr_db_t *db, const float *q_dev, int64_t nq, float *out_host);
/* merge n_lists candidate lists per query, laid out (nq, n_lists*k) - the step after the
 * all-gather of per-GPU top-k. */
int asr_topk_merge(const float *score_dev, const int64_t *idx_dev, int64_t nq, int n_lists, int k,
                   float *out_score_dev, int64_t *out_idx_dev, void *stream);
/* The same merge reading the result of ONE all-gather directly.  Every rank lays its local result out as one chunk
 * of chunk_bytes: [scores (nq,k) f32 | pad to 8 bytes | indices (nq,k) i64 at idx_offset_bytes] (asr_topk can write
 * straight into such a chunk); gathered_dev holds n_lists chunks back to back (rank-major), no transpose needed. */
int asr_topk_merge_gathered(const void *gathered_dev, int64_t chunk_bytes, int64_t idx_offset_bytes, int64_t nq,
                            int n_lists, int k, float *out_score_dev, int64_t *out_idx_dev, void *stream);
/* eval_retrieval over a sharded DB: after phase 0 of asr_rank_of_target every rank holds (tscore, tidx) of its
 * shard's best correct item (-inf / -1 if it owns none); all-gather chunks [tscore (nq) f32 | pad | tidx (nq) i64]
 * and this picks, per query, the best score (ties: smaller global index) -- the value every rank needs for phase 1. */
int asr_rank_target_merge(const void *gathered_dev, int64_t chunk_bytes, int64_t idx_offset_bytes, int n_lists,
                          int64_t nq, float *tscore_dev, int64_t *tidx_dev, void *stream);
/* eval_retrieval ranking: query i's correct items are DB rows j with j/kg == (q_base+i)/hg
 * (global j).  tscore_dev (nq) in/out: pass 1 (phase=0) computes the best correct score/index
 * on the shard that owns them (others leave -inf); after a max-allreduce, phase=1 adds to
 * better_dev (nq, int64) the number of local rows ranked before the best correct item. */
int asr_rank_of_target(asr_db_t *db, const float *q_dev, int64_t nq, int64_t q_base, int64_t kg,
                       int64_t hg, int normalise, int phase, float *tscore_dev, int64_t *tidx_dev,
                       int64_t *better_dev, void *stream);
/* per-recording vote (asr/audio_sheet_server.py:237-251): cand_idx_dev (n_rec, m) DB row indices
 * (-1 = empty), row_ids_dev maps DB row -> piece id.  out_ids/out_counts (n_rec, top_k), filled with
 * -1/0 beyond the number of distinct pieces.  Order: count desc, then piece id desc. */
int asr_vote(const int64_t *cand_idx_dev, const int32_t *row_ids_dev, int64_t n_rows, int n_rec, int m,
             int top_k, int32_t *out_ids_dev, int32_t *out_counts_dev, void *stream);

/* ------------------------------------------------------------------------- *
 * CCA statistics and solve.  Replaces asr/utils/cca.py:25-53,199-211 (CCA.fit 'svd')
 * and the forward of layers/cca.py:91-182.
 * ------------------------------------------------------------------------- */
#define ASR_CCA_NSUMS 3136   /* sum x (32) | sum y (32) | sum x x^T | sum y y^T | sum x y^T (3 x 1024) */

/* adds this shard's sufficient statistics of (h1,h2) (n,32) to sums_dev (fp64, ASR_CCA_NSUMS);
 * shift1/shift2 (32, fp32, device, may be NULL) are subtracted from every row first. */
int asr_cca_accumulate(const float *h1_dev, const float *h2_dev, int64_t n, const float *shift1_dev,
                       const float *shift2_dev, double *sums_dev, void *stream);
/* Counted layout for row-sharded fits: sums_dev has ASR_CCA_NSUMS + 1 entries and the call also adds n to the
 * last one, so ONE all-reduce of the buffer carries everything (asr/refine_cca.py:94-101 sharded over ranks);
 * asr_cca_solve then reads the total row count from the device (n_total = ASR_CCA_COUNT_ON_DEVICE) -- no host
 * synchronisation between accumulation, all-reduce and solve.  Use the same shift (or NULL) on every rank:
 * the solve removes it exactly (raw-moment correction in fp64), so no pass over the data for the means is needed. */
#define ASR_CCA_COUNT_ON_DEVICE (-1)
int asr_cca_accumulate_counted(const float *h1_dev, const float *h2_dev, int64_t n, const float *shift1_dev,
                               const float *shift2_dev, double *sums_dev, void *stream);
/* 'svd' fit from (all-reduced) sums.  mode 0 = CCA.fit('svd') (sigma descending);
 * mode 1 = CCALayer train forward (eigh of TT'+rT, ascending, sign fix) from BATCH statistics only, i.e. the
 * reference layer with ALPHA = 1.0 (asr/models/mutopia_ccal_cont.py:48; layers/cca.py:96-141 blends batch and
 * running statistics with alpha -- other alphas are not provided).
 * Outputs (device): m1,m2 (32) U,V (32x32 row-major) as fp64, sigma (32) fp64. */
int asr_cca_solve(const double *sums_dev, int64_t n_total, const float *shift1_dev, const float *shift2_dev,
                  double r1, double r2, double rT, int mode, double *m1_dev, double *m2_dev,
                  double *U_dev, double *V_dev, double *sigma_dev, void *stream);

/* CCALayer training backward (asr/models/lasagne_extensions/layers/cca.py:91-203 differentiated -- the reference gets
 * it from Theano's automatic differentiation: EighGrad through the four eigendecompositions, zero gradient through sgn
 * and outside the clip range of corr).  Batch statistics only (ALPHA = 1, as asr_cca_solve mode 1).
 * h1, h2 (n,32): the layer's inputs; g1 = dL/d lv1_cca, g2 = dL/d lv2_cca_fixed (n,32), e.g. from asr_contrastive_loss;
 * g_corr_dev: dL/d corr (32 doubles: the layer's own loss term -wl * mean(corr) gives -wl/32 each) or NULL.
 * Writes dL/dh1, dL/dh2 (n,32) float32.  Three Gram passes, one single-CTA fp64 kernel for the 32 x 32 chain, one row
 * pass; workspace allocated per device at first use. */
int asr_cca_layer_backward(const float *h1_dev, const float *h2_dev, const float *g1_dev, const float *g2_dev, int64_t n,
                           double r1, double r2, double rT, const double *g_corr_dev, float *dh1_dev, float *dh2_dev,
                           void *stream);

/* ------------------------------------------------------------------------- *
 * Spectrogram front-end (SURVEY 8f "next" row 4).  Replaces the madmom processor chain of
 * tutorials/Embedding Tutorial.ipynb cell 28 (and of the microphone stream, asr/audio_sheet_server.py:44-60):
 *   FramedSignalProcessor(frame_size, fps, origin='future') -> Hann window -> |FFT| (frame_size / 2 bins)
 *   -> FilteredSpectrogramProcessor(filterbank) -> LogarithmicSpectrogramProcessor (log10(1 + x)).
 * audio_dev: mono float32 samples in [-1, 1) on the device (int16 / 32768).  Frame i starts at
 * int(i * sample_rate / fps); there are ceil(n_samples / (sample_rate / fps)) frames (asr_spectrogram_num_frames),
 * samples beyond the end are zeros.  filterbank_dev: (frame_size / 2, n_bands) float32 row-major; band_lo/band_hi:
 * per band the half-open range of bins with non-zero weight.  out_dev: (n_bands, n_frames) float32 -- the layout
 * `processor.process(path).T` has in the reference, ready for asr_extract_windows.  frame_size: power of two <= 4096.
 * ------------------------------------------------------------------------- */
int asr_spectrogram_num_frames(int64_t n_samples, int sample_rate, double fps);
int asr_log_spectrogram(const float *audio_dev, int64_t n_samples, int sample_rate, int frame_size, double fps,
                        const float *filterbank_dev, const int32_t *band_lo_dev, const int32_t *band_hi_dev, int n_bands,
                        float *out_dev, void *stream);

/* ------------------------------------------------------------------------- *
 * Alignment (SURVEY 8f "next" row 3).  Replaces cdist(img_codes, spec_codes, 'cosine')
 * (asr/utils/alignment.py:149) and dtw_by_dist (asr/utils/dtw_by_dist.py:6-34, 69-83).
 * ------------------------------------------------------------------------- */
/* out_dev (r,c) float64 cosine distances between a (r,32) and b (c,32) float32 */
int asr_cosine_distances(const float *a_dev, int r, const float *b_dev, int c, double *out_dev, void *stream);
/* dist_dev (r,c) float64 -> acc_dev (r,c) accumulated cost; warp path (i,j) from (0,0) to (r-1,c-1) in
 * path_i_dev / path_j_dev (capacity r+c-1 each), its length in path_len_dev.  Ties: diag, then up, then left. */
int asr_dtw(const double *dist_dev, int r, int c, double *acc_dev, int32_t *path_i_dev, int32_t *path_j_dev,
            int32_t *path_len_dev, void *stream);

/* ------------------------------------------------------------------------- *
 * Training objective (SURVEY 8f "next" row 4, first slice).  Replaces the Theano graph of
 * get_contrastive_cos_loss(weight, gamma, symmetric) (asr/models/objectives.py:30-69) and the
 * gradient Theano derives from it.
 * ------------------------------------------------------------------------- */
/* lv1_dev, lv2_dev (n,32) float32 codes of matching pairs (row i of one view belongs to row i of the other).
 * loss_dev: 1 float.  grad1_dev / grad2_dev: (n,32) float32 d loss / d lv1, d loss / d lv2, or NULL.
 * scratch_dev: n doubles.  2 <= n <= 8192.  Deterministic (fixed reduction order). */
int asr_contrastive_loss(const float *lv1_dev, const float *lv2_dev, int64_t n, float weight, float gamma, int symmetric,
                         double *scratch_dev, float *loss_dev, float *grad1_dev, float *grad2_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H */
