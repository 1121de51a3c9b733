/*
 * ttlda.h — C ABI of libttlda.so: the sm_100a kernels behind TuoTuo's batch variational-inference LDA hot path.
 *
 * This is the drop-in boundary.  The reference exposes its only native code as the CPython
 * extension `tuotuo.cutils` (tuotuo/cutils.pyx:11-93, built by setup.py:18-27); everything else on the path is
 * NumPy/SciPy inside `tuotuo/lda_model.py`.  Each entry point below names the reference interface it replaces.
 *
 * Conventions (every function):
 *   - plain C types only; all array arguments are DEVICE pointers owned by the caller (e.g. torch
 *     `Tensor.data_ptr()`), `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - asynchronous on `stream`; never allocates, frees or synchronises; scratch space is caller-provided and
 *     sized by the matching `*_workspace_bytes` query;
 *   - returns 0 (TTLDA_OK) or a negative TTLDA_E* code; `ttlda_last_error()` gives the message (thread-local);
 *   - all floating point is IEEE double ("f64"), all results are bit-reproducible run to run unless a function
 *     says otherwise; indices are bit-exact.
 *
 * Device layouts ("word-major", chosen for coalesced K-vector gathers — DESIGN.md §3):
 *   ld            leading dimension of every K-vector table: K rounded up to a multiple of 4 (32-byte rows)
 *   theta tables  gamma, exp_elogtheta           [D][ld]   row d = document d        (reference: (M,K))
 *   beta tables   lambda, exp_elogbeta, sstats   [V][ld]   row w = vocabulary id w   (reference: (K,V) — transposed)
 *   padding columns k in [K, ld) hold 0 in exp_* tables and sstats and are ignored elsewhere.
 *   CSR corpus    rowptr int64[D+1], colidx int32[nnz] ascending within a row, counts int32[nnz]
 *   CSC mirror    colptr int64[V+1], docidx int32[nnz] ascending within a column, ccounts int32[nnz]
 */
#ifndef TTLDA_H
#define TTLDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTLDA_OK 0
#define TTLDA_EINVAL (-1)   /* bad argument (NULL pointer, K out of range, ld not a multiple of 4, ...) */
#define TTLDA_ECUDA (-2)    /* a CUDA runtime call / kernel launch failed; see ttlda_last_error() */
#define TTLDA_EWORKSPACE (-3) /* workspace too small */
#define TTLDA_EARCH (-4)    /* device is not sm_100 */

#define TTLDA_MAX_K 1024 /* topics per model supported by the register-resident E-step */
#define TTLDA_NPART 8    /* doubles per partial-sum record written by the reduction kernels */

/* library ------------------------------------------------------------------------------------------------ */
int ttlda_version(void);                 /* (major<<16)|(minor<<8)|patch */
const char* ttlda_last_error(void);      /* message of the last failing call on this thread ("" if none) */
int ttlda_check_device(int device);      /* TTLDA_OK iff `device` is compute capability 10.x */
int64_t ttlda_ld_for(int64_t K);         /* the leading dimension the kernels require for K topics */

/*
 * Row-wise Dirichlet expectation with the reference's AS-103 digamma.
 * Replaces: tuotuo.cutils._dirichlet_expectation_2d (tuotuo/cutils.pyx:41-68) and psi (:75-93); with
 * out_exp != NULL also the np.exp(...) applied to its result at tuotuo/lda_model.py:232-233,264.
 *   a [rows][ld] (first `cols` entries of each row are used);  out_elog / out_exp [rows][ld], either may be NULL.
 */
int ttlda_dirichlet_expectation_f64(const double* a, int64_t rows, int64_t cols, int64_t ld,
                                    double* out_elog, double* out_exp, void* stream);

/*
 * In-place variant of the single-sample form.
 * Replaces: tuotuo.cutils._dirichlet_expectation_1d (tuotuo/cutils.pyx:11-38):
 *   doc_topic[i] += prior;  out[i] = exp(psi~(doc_topic[i]) - psi~(sum_i doc_topic[i])).
 */
int ttlda_dirichlet_expectation_1d_f64(double* doc_topic, double prior, double* out, int64_t size, void* stream);

/*
 * Bag-of-words: token-id stream -> CSR doc-term matrix (sorted unique word ids per document + counts).
 * Replaces: the count matrix built by get_vocab_from_docs / get_np_wct (tuotuo/utils.py:177-206) and exposed as
 * Document.doc_vocab_count_array (tuotuo/document.py:33-36); the rows equal np.nonzero(X[d,:]) / X[d, ids].
 *   token_ids int32[T] (vocabulary ids, 0 <= id < V), doc_offsets int64[D+1] (token ranges per document).
 *   Outputs: rowptr int64[D+1], colidx/counts int32[>= T] (first *nnz entries valid), nnz_out int64[1] (device).
 */
int64_t ttlda_bow_to_csr_workspace_bytes(int64_t T, int64_t D, int64_t V);
int ttlda_bow_to_csr(const int32_t* token_ids, const int64_t* doc_offsets, int64_t T, int64_t D, int64_t V,
                     int64_t* rowptr, int32_t* colidx, int32_t* counts, int64_t* nnz_out,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* CSR -> CSC mirror (stable: documents ascending within each word).  No reference counterpart: the dense
 * matrix of tuotuo/document.py:36 is addressable both ways; the CSC copy gives the M-step the same access.
 * With nblocks > 1 the mirror is DOCUMENT-BLOCKED: documents are cut into nblocks blocks of block_docs documents
 * and "virtual word" b*V + w lists the postings of word w inside block b (colptr has nblocks*V + 1 entries), so
 * that the statistics pass gathers exp(E[log theta]) rows from one L2-sized window of the D x K table at a time.
 * csr2csc (uint32[nnz], may be NULL): position in the mirror of every CSR position — lets the E-step leave the
 * per-nonzero ratio c/N for the statistics pass in mirror order (ttlda_estep_f64 `ratio_csc`). */
int64_t ttlda_csr_to_csc_workspace_bytes(int64_t nnz, int64_t D, int64_t V, int64_t nblocks);
int ttlda_csr_to_csc(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts,
                     int64_t D, int64_t V, int64_t nnz, int64_t nblocks, int64_t block_docs,
                     int64_t* colptr, int32_t* docidx, int32_t* ccounts, uint32_t* csr2csc,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* Unsigned 8- or 16-bit items -> int32 (src_bytes_per_item = 1 | 2).  Lets a corpus cross PCIe in a compact form
 * (uint16 word ids when V <= 65536, uint8/uint16 counts: 3 instead of 8 bytes per nonzero) and be widened into the
 * colidx / counts arrays on the device. */
int ttlda_widen_i32(const void* src, int64_t src_bytes_per_item, int32_t* dst, int64_t n, int64_t bound,
                    int32_t* flags, void* stream);
/* bound > 0 (word ids: bound = V): items >= bound set TTLDA_BAD_COLIDX in *flags and are stored as 0. */

/* Validation of a caller-supplied CSR corpus before it drives device addressing (the reference indexes a dense
 * M x V matrix, tuotuo/lda_model.py:226-227, where a bad index is a NumPy IndexError; here it would be an
 * out-of-bounds gather).  Checks: rowptr[0] == pos_base, rowptr monotone, rowptr[D] == pos_base + nnz; 0 <= colidx < V;
 * counts >= 0.  Violations OR the TTLDA_BAD_* bits into flags[0] (device int32, zeroed by the caller).  sanitize != 0
 * also rewrites the offending entries (rowptr clamped into range, colidx/counts set to 0) so that kernels already
 * queued behind this call stay in bounds; the corpus must still be rejected.  rowptr or colidx may be NULL (skipped);
 * with rowptr == NULL this validates a token-id stream against V. */
#define TTLDA_BAD_ROWPTR 1
#define TTLDA_BAD_COLIDX 2
#define TTLDA_BAD_COUNT 4
int ttlda_validate_csr(int64_t* rowptr, int64_t D, int64_t pos_base, int32_t* colidx, int32_t* counts, int64_t nnz,
                       int64_t V, int sanitize, int32_t* flags, void* stream);

/* One document block of a document-blocked mirror, built on its own (streaming ingestion: block b can be mirrored,
 * E-stepped and accumulated while block b+1 is still crossing PCIe).  The block holds documents
 * [doc_base, doc_base + D_block) whose entries occupy positions [pos_base, pos_base + nnz_block) of the corpus-wide
 * CSR arrays AND of the corpus-wide mirror arrays (blocks are contiguous in both).  rowptr_block = &rowptr[doc_base]
 * (D_block + 1 absolute offsets); colidx/counts/docidx/ccounts/csr2csc are the corpus-wide array bases;
 * colptr_block = &colptr[b * V] (V + 1 absolute offsets are written).  The result is identical to the block's slice
 * of ttlda_csr_to_csc(nblocks, block_docs = D_block).  ccounts may be NULL (the mirrored counts are only read by
 * ttlda_sstats_csc_f64 without ratio_csc).  Workspace: ttlda_csr_to_csc_workspace_bytes(nnz_block,..,1). */
int ttlda_csr_to_csc_block(const int64_t* rowptr_block, const int32_t* colidx, const int32_t* counts,
                           int64_t D_block, int64_t V, int64_t doc_base, int64_t pos_base, int64_t nnz_block,
                           int64_t* colptr_block, int32_t* docidx, int32_t* ccounts, uint32_t* csr2csc,
                           void* workspace, int64_t workspace_bytes, void* stream);

/*
 * theta refresh: exp_elogtheta = exp(dirichlet_expectation(gamma)) and the theta-prior ELBO term.
 * Replaces: _dirichlet_expectation_2d(self._gamma_) at tuotuo/lda_model.py:116,256 (+ np.exp at :232) and
 * expec_real_dist_minus_suro_dist(Elogtheta, alpha, gamma) (:25-46, call site :142-148).
 *   alpha_vec double[K] on device (a scalar prior is broadcast by the caller); alpha_sum = np.sum(alpha) as the
 *   reference evaluates it (the scalar itself for a scalar prior — SURVEY.md §8 a7).
 *   partial_out double[TTLDA_NPART] (device): [0] = theta-prior term, rest 0.  exp_elogtheta may be NULL.
 */
int64_t ttlda_reduce_workspace_bytes(void);
int ttlda_theta_refresh_f64(const double* gamma, int64_t D, int64_t K, int64_t ld,
                            const double* alpha_vec, double alpha_sum,
                            double* exp_elogtheta, double* partial_out,
                            void* workspace, int64_t workspace_bytes, void* stream);

/*
 * E-step, document-major pass (one warp per document, K topics in registers).
 * Replaces: LDASmoothed.e_step (tuotuo/lda_model.py:197-270) for gamma, plus — fused — the per-token ELBO term
 * of approx_elbo (:125-139) evaluated at the *incoming* state and the theta-prior term (:142-148) and theta
 * refresh (:256) at the *outgoing* state.  Per document d with words w, counts c:
 *      N_dw = sum_k eT_in[d,k]*eB[w,k] (+1e-100 in the normaliser, :242);  gamma[d,k] = alpha_k + eT_in[d,k] *
 *      sum_w (c_dw/N_dw) eB[w,k] (:248);  eT_out[d,:] = exp(dirichlet_expectation(gamma[d,:])).
 * `max_iter`/`tol` restate the reference's while-loop (:236): because exp(Elogtheta_d) is frozen inside it, the
 * update is idempotent and the loop exits on its second pass; the kernel evaluates the update once and reports
 * `iters_out[0]` = sum over documents of the passes the reference would have made, `iters_out[1]` = documents
 * that hit the cap (`it > max_iter`, warned about at :258-259).
 *   exp_elogtheta_out may be NULL: then only gamma_out is written and partial_out[1] = 0 — run
 *   ttlda_theta_refresh_f64 on gamma_out afterwards (same results; the refresh's digamma/log-gamma/exp chains then
 *   run as a separate fully parallel pass instead of on the critical path of the per-document warps).
 *   gamma_in: the previous gamma [D][ld]; read only for the loop statistics (may be NULL iff iters_out is NULL,
 *   may alias gamma_out).  gamma_out / exp_elogtheta_out are written; exp_elogtheta_out must not alias the input
 *   (ping-pong), which lets a caller run the pass speculatively and discard it.
 *   csr2csc / ratio_csc (both or neither): if given, ratio_csc[csr2csc[p]] = c_p / N_p is stored for every nonzero p
 *   (one double per nonzero — not phi): the statistics pass then needs no dot product, reduction or division
 *   (ttlda_sstats_csc_f64 `ratio_csc`).
 *   partial_out double[TTLDA_NPART]: [0] token term at the incoming state, [1] theta-prior term at the outgoing
 *   state, [2] sum of counts.
 */
int ttlda_estep_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts,
                    int64_t D, int64_t K, int64_t V, int64_t ld,
                    const double* alpha_vec, double alpha_sum,
                    const double* exp_elogtheta_in, const double* exp_elogbeta,
                    const double* gamma_in, double* gamma_out, double* exp_elogtheta_out,
                    const uint32_t* csr2csc, double* ratio_csc,
                    int64_t max_iter, double tol, int64_t* iters_out, double* partial_out,
                    void* workspace, int64_t workspace_bytes, void* stream);

/*
 * The E-step with a shared-memory cache of the most frequent words' exp(E[log beta_w]) rows (north_star (2): "the
 * needed columns of exp(E[log beta]) staged in shared memory").  Same update as ttlda_estep_f64 with
 * exp_elogtheta_out == NULL (gamma only; run ttlda_theta_refresh_f64 afterwards) — tuotuo/lda_model.py:242,248 and the
 * token term of :125-139 — on an "E-step stream" of the corpus instead of the plain CSR arrays:
 *   ttlda_estep_hot_rows(ld)   how many rows of ld doubles the cache holds (0: this row width runs ttlda_estep_f64);
 *   ttlda_build_estream        slot_of_word int32[V] (cache slot of a word, -1 = not cached; the caller picks the H most
 *                              frequent words) + CSR (+ csr2csc) -> e_idx / e_cnt / e_pos: within every document the
 *                              cached words first (stable), e_idx = slot | 0x80000000 for them, the word id otherwise;
 *                              counts and mirror positions permuted alike.  Integer work, bit-exact, once per corpus.
 *   ttlda_estep_hot_f64        hot_words int32[H] = the word in each slot.  e_pos / ratio_csc: both or neither, as
 *                              csr2csc / ratio_csc of ttlda_estep_f64.  gamma differs from ttlda_estep_f64's only by the
 *                              order in which a document's nonzeros are summed (~1 ulp); bit-reproducible run to run.
 */
int64_t ttlda_estep_hot_rows(int64_t ld);
int ttlda_build_estream(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, const uint32_t* csr2csc,
                        int64_t D, int64_t V, const int32_t* slot_of_word,
                        int32_t* e_idx, int32_t* e_cnt, uint32_t* e_pos, void* stream);
int ttlda_estep_hot_f64(const int64_t* rowptr, const int32_t* e_idx, const int32_t* e_cnt, const uint32_t* e_pos,
                        const int32_t* hot_words, int64_t H,
                        int64_t D, int64_t K, int64_t V, int64_t ld, const double* alpha_vec,
                        const double* exp_elogtheta_in, const double* exp_elogbeta,
                        const double* gamma_in, double* gamma_out, double* ratio_csc,
                        int64_t max_iter, double tol, int64_t* iters_out, double* partial_out,
                        void* workspace, int64_t workspace_bytes, void* stream);

/*
 * OPT-IN E-step with a true per-document fixed point (not the reference's behaviour — SURVEY.md §8f row 1):
 * exp(E[log theta_d]) is refreshed after every gamma_d update inside the reference's loop
 * (`while l1 > tol and it <= max_iter`, tuotuo/lda_model.py:236) instead of being frozen, as in Blei/Hoffman and
 * scikit-learn's _update_doc_distribution.  Same arguments as ttlda_estep_f64 (all required, outputs must not alias
 * inputs).  iters_out[0] = total inner iterations, iters_out[1] = documents that hit the cap.  The statistics pass
 * must then be given exp_elogtheta_out (the refreshed table), which re-evaluates the normaliser with the final theta.
 */
int ttlda_estep_fixed_point_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts,
                                int64_t D, int64_t K, int64_t V, int64_t ld,
                                const double* alpha_vec, double alpha_sum,
                                const double* exp_elogtheta_in, const double* exp_elogbeta,
                                const double* gamma_in, double* gamma_out, double* exp_elogtheta_out,
                                int64_t max_iter, double tol, int64_t* iters_out, double* partial_out,
                                void* workspace, int64_t workspace_bytes, void* stream);

/*
 * M-step sufficient statistics, word-major pass over the CSC mirror (no atomics, no phi in HBM):
 *      sstats[w,k] = sum_{d: X_dw>0} eT[d,k] * c_dw / (sum_k' eT[d,k'] eB[w,k'] + 1e-100)
 * Replaces: lambda_update[:, ids] += np.outer(eT, cts/phi_norm) at tuotuo/lda_model.py:262 (the trailing
 * `*= exp(Elogbeta)` of :264 is applied by ttlda_mstep_f64, after the cross-GPU sum — it is linear).
 * colptr/docidx/ccounts: the (possibly document-blocked, nblocks >= 1) mirror built by ttlda_csr_to_csc; per-block
 * partial statistics live in the workspace and are folded in block order; sstats [V][ld] is the total.
 * ratio_csc (double[nnz] in mirror order, may be NULL): the c/N ratios left by ttlda_estep_f64 for the SAME
 * exp_elogtheta / exp_elogbeta; when given the pass is a pure gather-accumulate S[w,:] += ratio * eT[d,:]
 * (exp_elogbeta is then unused and may be NULL), otherwise the normaliser is recomputed from the two tables.
 */
int64_t ttlda_sstats_workspace_bytes(int64_t nnz, int64_t V, int64_t ld, int64_t nblocks);
int ttlda_sstats_csc_f64(const int64_t* colptr, const int32_t* docidx, const int32_t* ccounts,
                         int64_t D, int64_t V, int64_t K, int64_t ld, int64_t nnz, int64_t nblocks,
                         const double* exp_elogtheta, const double* exp_elogbeta, const double* ratio_csc,
                         double* sstats,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* Topic-windowed form of the ratio pass, for tables whose rows are long (K >~ 256): one walk of the whole (blocked)
 * mirror per window of kw columns (kw a multiple of 16, <= 256), gathering only exp_elogtheta[d][k0 .. k0+kw) — so the
 * L2 working set of a document block is block_docs * kw * 8 bytes and the per-block partials are nblocks * V * kw
 * doubles, re-used by every window.  ratio_csc is required (ttlda_estep_f64 leaves it); same (K,V) statistics as
 * ttlda_sstats_csc_f64 up to summation order (~1 ulp), bit-reproducible run to run.  Replaces tuotuo/lda_model.py:262
 * like the unwindowed pass. */
int64_t ttlda_sstats_windowed_workspace_bytes(int64_t nnz, int64_t V, int64_t kw, int64_t nblocks);
int ttlda_sstats_csc_windowed_f64(const int64_t* colptr, const int32_t* docidx,
                                  int64_t D, int64_t V, int64_t K, int64_t ld, int64_t nnz, int64_t nblocks, int64_t kw,
                                  const double* exp_elogtheta, const double* ratio_csc, double* sstats,
                                  void* workspace, int64_t workspace_bytes, void* stream);

/* The same pass over ONE document block of a blocked mirror (built by ttlda_csr_to_csc_block or a slice of
 * ttlda_csr_to_csc): colptr_block = &colptr[b * V] (absolute offsets), docidx / ccounts / ratio_csc / exp_elogtheta
 * are the corpus-wide bases, sstats_block [V][ld] receives the block's partial statistics;
 * ttlda_sstats_fold_f64 then sums the nblocks partials [nblocks][V][ld] in block order into sstats [V][ld] —
 * bit-identical to the one-call form.  accumulate != 0: sstats_block is added to instead of overwritten — calling the
 * blocks in order on ONE [V][ld] table (accumulate = 0 for the first) gives the same bits without any per-block
 * buffers: the form for tables whose nblocks * V * ld partials would not fit (K*V and K*D both >> L2).
 * Workspace: ttlda_sstats_workspace_bytes(nnz_block, V, ld, 1). */
int ttlda_sstats_csc_block_f64(const int64_t* colptr_block, const int32_t* docidx, const int32_t* ccounts,
                               int64_t D, int64_t V, int64_t K, int64_t ld, int64_t pos_base, int64_t nnz_block,
                               const double* exp_elogtheta, const double* exp_elogbeta, const double* ratio_csc,
                               double* sstats_block, int accumulate,
                               void* workspace, int64_t workspace_bytes, void* stream);
int ttlda_sstats_fold_f64(const double* sstats_blocks, int64_t nblocks, int64_t V, int64_t ld, double* sstats,
                          void* stream);

/* One-pass alternative to ttlda_sstats_csc_f64 (the north-star's literal formulation): document-major scatter of
 * eT[d,k]*c/N into sstats[w,k] with red.global.add.f64.  Same result up to summation order (NOT bit-reproducible);
 * bounded by REDG throughput, kept for comparison.  sstats [V][ld] is zeroed by the call. */
int ttlda_sstats_atomic_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts,
                            int64_t D, int64_t K, int64_t V, int64_t ld,
                            const double* exp_elogtheta, const double* exp_elogbeta, double* sstats, void* stream);

/*
 * M-step: lambda = eta + sstats * exp_elogbeta;  exp_elogbeta <- exp(dirichlet_expectation(lambda));
 * beta-prior ELBO term at the new lambda.
 * Replaces: lambda_update *= np.exp(expec_log_beta) (tuotuo/lda_model.py:264), LDASmoothed.m_step (:273-283)
 * and expec_real_dist_minus_suro_dist(Elogbeta, eta, lambda) (:157-163).
 *   lambda [V][ld] out; exp_elogbeta [V][ld] in/out; elogbeta_out may be NULL; colsum_out double[ld] = sum_w lambda[w,k].
 *   partial_out double[TTLDA_NPART]: [0] = beta-prior term.
 */
int64_t ttlda_mstep_workspace_bytes(int64_t V, int64_t ld);
int ttlda_mstep_f64(const double* sstats, int64_t V, int64_t K, int64_t ld, double eta,
                    double* exp_elogbeta, double* lambda, double* elogbeta_out, double* colsum_out,
                    double* partial_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Online (stochastic) M-step — Hoffman, Blei & Bach, "Online Learning for LDA" (2010), eq. 7-9; the reference only
 * lists it as roadmap (readme.md:131-136), its `sampling=True` has no update rule:
 *      lambda <- (1 - rho) * lambda + rho * (eta + scale * sstats * exp_elogbeta),   scale = D_population / |minibatch|
 * then the same exp_elogbeta refresh and beta-prior record as ttlda_mstep_f64 (which is rho = scale = 1). */
int ttlda_mstep_online_f64(const double* sstats, int64_t V, int64_t K, int64_t ld, double eta, double scale,
                           double rho, double* exp_elogbeta, double* lambda, double* elogbeta_out,
                           double* colsum_out, double* partial_out,
                           void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Sharded M-step for document-sharded multi-GPU runs: rank r owns rows [w0, w0 + n) of
 * lambda; what the reference does in `lambda_update *= exp(Elogbeta)` / m_step (tuotuo/lda_model.py:264,273-283) over
 * the whole (K,V) matrix is done once, split over the ranks, instead of on every rank after an all-reduce.
 *   phase A  ttlda_mstep_slice_lambda_f64: S[w] = sum_i src_tables[i][src_row0 + (w - w0)] in index order — host array of
 *            `nsrc` DEVICE pointers: every rank's [V][ld] statistics table mapped through peer memory (nsrc = world,
 *            src_row0 = w0: the reduce-scatter is fused into the kernel as plain NVLink loads), or one buffer a library
 *            reduce-scatter filled (nsrc = 1, src_row0 = 0);  lambda[w] = eta + S[w] * exp_elogbeta[w] for the slice;
 *            colsum_part_out double[ld] = column sums of the slice.
 *   (the caller exchanges the ranks' colsum parts: `colsum_parts` double[nparts][ld], rank order)
 *   phase B  ttlda_mstep_slice_beta_f64: colsum = sum of the parts in rank order (identical bits on every rank);
 *            exp_elogbeta rows of the slice = exp(psi~(lambda) - psi~(colsum)), stored into each of the `ndst` tables of
 *            dst_tables (host array of device pointers: this rank's table, or all ranks' through peer memory — the
 *            all-gather fused into the kernel as NVLink stores); partial_out[0] = the slice's share of the beta-prior
 *            term (+ the colsum-only part iff add_colsum_term, which exactly one rank passes).
 * src_multicast / dst_multicast (may be NULL): the NVSwitch MULTICAST address of the same tables (NVLink SHARP).  When
 * given, phase A reads the cross-rank sum with one multimem.ld_reduce.add.f64 per element — the reduction happens in the
 * switch and this rank's link carries its slice once instead of once per peer — and phase B replicates every row into
 * all ranks' tables with one multimem.st; src_tables / dst_tables are then only checked.  The in-switch summation order
 * is the hardware's (each lambda row is still computed by exactly one rank, so the ranks cannot disagree).
 * Workspace for both: ttlda_mstep_workspace_bytes.  lambda / exp_elogbeta are this rank's full [V][ld] tables.
 */
int ttlda_mstep_slice_lambda_f64(const double* const* src_tables, int64_t nsrc, const double* src_multicast,
                                 int64_t src_row0, int64_t w0, int64_t n,
                                 int64_t K, int64_t ld, double eta, const double* exp_elogbeta, double* lambda,
                                 double* colsum_part_out, void* workspace, int64_t workspace_bytes, void* stream);
int ttlda_mstep_slice_beta_f64(const double* colsum_parts, int64_t nparts, int64_t w0, int64_t n, int64_t K, int64_t ld,
                               double eta, const double* lambda, double* const* dst_tables, int64_t ndst,
                               double* dst_multicast, int add_colsum_term, double* colsum_out, double* partial_out,
                               void* workspace, int64_t workspace_bytes, void* stream);

/* beta refresh from an existing lambda (initial state, or after a hyper-parameter update):
 * exp_elogbeta = exp(dirichlet_expectation over the vocabulary of lambda[:,k]) and the beta-prior term.
 * Replaces: _dirichlet_expectation_2d(self._lambda_) at tuotuo/lda_model.py:120 and the term at :157-163. */
int ttlda_beta_refresh_f64(const double* lambda, int64_t V, int64_t K, int64_t ld, double eta,
                           double* exp_elogbeta, double* elogbeta_out, double* colsum_out,
                           double* partial_out, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Token term of the ELBO: sum_d sum_w c_dw * log(sum_k eT[d,k]*eB[w,k])  (= logsumexp_k(Elogtheta+Elogbeta)).
 * Replaces: the double loop over documents and words of approx_elbo, tuotuo/lda_model.py:125-139.
 *   partial_out double[TTLDA_NPART]: [0] token term, [2] sum of counts.
 */
int ttlda_elbo_tokens_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts,
                          int64_t D, int64_t K, int64_t V, int64_t ld,
                          const double* exp_elogtheta, const double* exp_elogbeta, double* partial_out,
                          void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Newton statistics for the hyper-parameter updates, with an EXACT digamma (the reference calls
 * scipy.special.psi here, not AS-103): out[0] = sum_ij psi(x_ij), out[1] = sum_i psi(sum_j x_ij) over the
 * first `cols` entries of each of `rows` rows of x [rows][ld].
 * Replaces: psi(self._gamma_) / psi(np.sum(self._gamma_,1)) at tuotuo/lda_model.py:449-462 and the lambda
 * analogue at :522-528 (for the word-major lambda pass cols=K and use out[0] with ttlda_psi_f64 on colsum).
 */
int ttlda_hyper_stats_f64(const double* x, int64_t rows, int64_t cols, int64_t ld, double* partial_out,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* Per-column form: out[k] = sum_r (psi(x_rk) - psi(sum_j x_rj)), k < cols — the gradient term of the reference's
 * non-exchangeable update_alpha, (phsi_gamma - psi(rowsum)).sum(axis=0) at tuotuo/lda_model.py:452-456. */
int64_t ttlda_hyper_colstats_workspace_bytes(int64_t rows, int64_t cols);
int ttlda_hyper_colstats_f64(const double* x, int64_t rows, int64_t cols, int64_t ld, double* out,
                             void* workspace, int64_t workspace_bytes, void* stream);
int ttlda_psi_f64(const double* x, int64_t n, double* out, void* stream); /* exact digamma, elementwise, x > 0 */

/* Measurement aid (bench.py `roofline.peak`): the row-gather rate this device sustains for `n` rows table[idx[i]][0..ld)
 * read with the library's own traversal (same lane groups, batching and register double-buffering as the E-step and
 * the statistics pass) and no arithmetic — the ceiling of the np.dot / np.outer gathers of tuotuo/lda_model.py:242,
 * 248,262 on this table and this index stream, whether they are served by L2 or by HBM.  0 <= idx[i] < rows is the
 * caller's responsibility; sink double[1] is never written in practice.  Time it with events around the call. */
int ttlda_probe_gather_f64(const double* table, int64_t rows, int64_t ld, const int32_t* idx, int64_t n,
                           double* sink, void* stream);

/* layout helpers between the reference's (K,V) topic-major matrices and the word-major device tables */
int ttlda_transpose_pad_f64(const double* src, int64_t rows, int64_t cols, double* dst, int64_t dst_ld,
                            void* stream); /* dst[c][r] = src[r][c]; dst is [cols][dst_ld], padding zeroed */
int ttlda_transpose_unpad_f64(const double* src, int64_t rows, int64_t cols, int64_t src_ld, double* dst,
                              void* stream); /* dst[c][r] = src[r][c]; src is [rows][src_ld], dst is [cols][rows] */

#ifdef __cplusplus
}
#endif
#endif /* TTLDA_H */
