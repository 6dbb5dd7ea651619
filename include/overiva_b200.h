/*
 * overiva_b200 -- C ABI of the B200-native OverIVA demixing loop (liboveriva_b200.so).
 *
 * The reference (onolab-tmu/overiva) is pure Python/NumPy and has no FFI layer: its drop-in boundary
 * is the Python call signature overiva()/auxiva_pca()/ogive() (overiva.py:28-38, auxiva_pca.py:30,
 * ive.py:33-45).  The Python host layer in overiva_b200/ keeps those signatures and drives this
 * library through ctypes.  Every entry point below replaces one NumPy/LAPACK step of the reference;
 * the file:line of the step it replaces is given with each declaration (paths relative to the
 * reference repository).
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types.  All data pointers are DEVICE pointers unless the
 *    name says host.  `stream` is a cudaStream_t passed as void* (NULL = default stream).  All calls
 *    are asynchronous on that stream and return 0 (OIVA_OK) or a negative OIVA_ERR_* code (a failing
 *    CUDA runtime call is OIVA_ERR_CUDA); oiva_last_error() gives a thread-local message with the
 *    details (for OIVA_ERR_CUDA: the call, the cudaError_t value and its string).  Return values are
 *    never positive except where a declaration says it returns a status word.
 *  - complex numbers are interleaved (re, im).  "c128" = two doubles, "c64" = two floats.  `dtype`
 *    (OIVA_C128 / OIVA_C64) is the storage type of X and Y only; covariances, demixing matrices and all
 *    on-chip arithmetic are fp64 in both modes.
 *  - shapes: B mixtures, T frames, F bins, M channels (1..16), K sources (1..M); R = B*F "rows"
 *    (row = b*F + f); NG = ceil(F/32) groups of 32 bins per mixture, G = B*NG; NE = M(M+1)/2.
 *  - grouped layouts (lane <-> bin, see csrc/common.cuh):
 *        Xg[gi][t][c][l]   complex<dtype>   samples,   gi = b*NG + f/32, l = f%32 (zero for f >= F)
 *        Vg[gi][k][e][l]   c128             covariance entry e = i(i+1)/2 + j (i >= j) of source k
 *    per-frame buffers (phi, r2) have the padded pitch Tp = oiva_frame_pitch(T).
 *  - numerical failure (singular pivot / non-finite value; the reference raises
 *    numpy.linalg.LinAlgError from overiva.py:98,182) is reported through a device status word that the
 *    caller reads back when convenient: bit 0 = singular pivot, bit 1 = non-finite result.
 *    Status words are PER MIXTURE: every `int* status` argument points to n_batch ints (word b belongs to
 *    mixture b), so that one failing mixture of a batch can be told from the others -- the reference's
 *    sweep records NaN for the failing mixture only and carries on (overiva_sim.py:334-350).
 */
#ifndef OVERIVA_B200_H
#define OVERIVA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OIVA_OK 0
#define OIVA_ERR_INVALID (-1)
#define OIVA_ERR_NOMEM (-2)
#define OIVA_ERR_STATE (-3)
#define OIVA_ERR_CUDA (-4) /* a CUDA runtime call failed; oiva_last_error() names it */
#define OIVA_ERR_UNSUPPORTED (-5) /* the shape is outside what this entry point covers (callers fall back) */

#define OIVA_C128 0
#define OIVA_C64 1

#define OIVA_MODEL_LAPLACE 0 /* r = 2 sqrt(sum_f |y|^2)          overiva.py:152-153 */
#define OIVA_MODEL_GAUSS 1   /* r = sum_f |y|^2 / F               overiva.py:154-155 */
#define OIVA_MODEL_NONE 2    /* any other string: r stays 0, as in the reference */
#define OIVA_MODEL_OGIVE_LAPLACE 3 /* r = sqrt(sum_f |y|^2 / F), no gamma rescale  ive.py:204-205 */
#define OIVA_MODEL_OGIVE_GAUSS 4   /* r = sum_f |y|^2 / F, no gamma rescale        ive.py:207-208 */

#define OIVA_INIT_EYE 0 /* overiva.py:111-114 */
#define OIVA_INIT_EIG 1 /* overiva.py:103-109 */
#define OIVA_INIT_W0 2  /* overiva.py:116-117 */

#define OIVA_STATUS_SINGULAR 1
#define OIVA_STATUS_NONFINITE 2
#define OIVA_STATUS_STALLED 4 /* a hand-over inside the single-launch loop timed out (never expected): results invalid */

int oiva_version(void);
const char* oiva_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Layout helpers (host-side arithmetic only, no GPU needed).
 * ---------------------------------------------------------------------------------------------- */
int oiva_bin_groups(int n_freq);  /* NG = ceil(F/32) */
int oiva_frame_pitch(int n_frames); /* Tp: T rounded up to a multiple of 32 */
size_t oiva_grouped_bytes(int n_batch, int n_frames, int n_freq, int n_chan, int dtype);   /* bytes of Xg */
size_t oiva_grouped_cov_bytes(int n_batch, int n_freq, int n_chan, int n_src);             /* bytes of Vg */

/* ------------------------------------------------------------------------------------------------
 * Kernels (one per step of the reference loop).
 * ---------------------------------------------------------------------------------------------- */

/* X (B,T,F,M) interleaved complex -> grouped samples Xg.   replaces: overiva.py:131-132 (swapaxes+copy) */
int oiva_relayout(const void* X, void* Xg, int n_batch, int n_frames, int n_freq, int n_chan, int dtype,
                  void* stream);

/* Vg[gi][k] = (1/T) sum_t phi[b][k][t] x x^H  (Hermitian: lower triangle stored, grouped layout).
 * phi: (B,K,Tp) inverse source-model weights, or NULL with n_src = 1 for the plain covariance.
 * replaces: overiva.py:179 (all K sources in one pass over X) and overiva.py:87 / ive.py:97 /
 * auxiva_pca.py:71 (phi == NULL). */
int oiva_weighted_cov(const void* Xg, const double* phi, void* Vg, int n_batch, int n_frames, int n_freq,
                      int n_chan, int n_src, int dtype, void* stream);
/* Same, with caller-provided scratch (oiva_weighted_cov_scratch_bytes(); may be 0 = not needed).  Inputs with few bin
 * groups (a single mixture) split the frames of a group over several thread teams to fill the GPU: with scratch every
 * split writes its partial sum to its own slot and the slots are added in a fixed order (bit-reproducible results);
 * without it (oiva_weighted_cov) the partial sums are combined with fp64 atomics, whose order is not fixed. */
size_t oiva_weighted_cov_scratch_bytes(int n_batch, int n_frames, int n_freq, int n_chan, int n_src);
int oiva_weighted_cov_ws(const void* Xg, const double* phi, void* Vg, void* scratch, size_t scratch_bytes,
                         int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream);

/* Weights per (source, frame, BIN): winv grouped [gi][k][Tp][32] float64.  V_k[f] = (1/T) sum_t winv_k(f,t) x x^H, the
 * auxiliary variable of ILRMA (source model r_k(f,t), winv = 1/r).  complex128 samples, n_chan <= 8 (else
 * OIVA_ERR_UNSUPPORTED); scratch as for oiva_weighted_cov_ws.
 * replaces: the covariance inside pyroomacoustics.bss.ilrma (called at overiva_oneshot.py:331-339, overiva_sim.py:309-311). */
int oiva_weighted_cov_binwise(const void* Xg, const double* winv, void* Vg, void* scratch, size_t scratch_bytes,
                              int n_batch, int n_frames, int n_freq, int n_chan, int n_src, void* stream);

/* oiva_relayout + the unweighted oiva_weighted_cov in ONE pass over X (read X once, write Xg and Cg): M <= 8, and for
 * complex64 an even F*M (16-byte aligned rows); oiva_relayout_cov_supported() tells.  Cg: grouped lower-triangle
 * covariance (oiva_unpack_cov gives the full matrices), bit-identical to the two-kernel path.  scratch as for
 * oiva_weighted_cov_ws with n_src = 1 (may be NULL: no frame splitting).
 * replaces: overiva.py:131-132 and overiva.py:87 together. */
int oiva_relayout_cov_supported(int n_freq, int n_chan, int dtype);
int oiva_relayout_cov(const void* X, void* Xg, void* Cg, void* scratch, size_t scratch_bytes, int n_batch,
                      int n_frames, int n_freq, int n_chan, int dtype, void* stream);

/* per-bin arrays of n_elems c128 each: row-major (R, n_elems) <-> grouped [gi][n_elems][32] (padded bins: 0).
 * Inside the loop the demixing matrices live in the grouped form (coalesced for lane <-> bin kernels). */
int oiva_group_rows(const void* rows, void* grouped, int n_batch, int n_freq, int n_elems, void* stream);
int oiva_ungroup_rows(const void* grouped, void* rows, int n_batch, int n_freq, int n_elems, void* stream);

/* Vg (grouped lower triangles) -> V (R,K,M,M) c128 full Hermitian matrices, row-major. */
int oiva_unpack_cov(const void* Vg, void* V, int n_batch, int n_freq, int n_chan, int n_src, void* stream);

/* r2part[b][g][k][t] = sum_{f in group g} |w_k(f)^H x(f,t)|^2 : (B,NG,K,Tp), padding frames written as 0.
 * W: per-bin M x w_cols c128 matrices, columns :K used (w_cols = M for W_hat, K for plain filters), either
 * row-major (R,M,w_cols) (w_grouped = 0) or grouped [gi][M*w_cols][32] (w_grouped = 1, what the plan keeps).
 * replaces: overiva.py:140 (demix) + the norm over frequency at overiva.py:152-155 / ive.py:204-208;
 * Y is never materialised inside the loop. */
int oiva_demix_power(const void* Xg, const void* W, int w_cols, int w_grouped, double* r2part, int n_batch,
                     int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream);

/* oiva_demix_power that also writes the per-bin powers |y_k(f,t)|^2: Pfull grouped [gi][k][Tp][32] float64 (padding
 * frames and padded bins zero).  replaces: the demix + np.power(abs(Y), 2) of pyroomacoustics.bss.ilrma. */
int oiva_demix_power_full(const void* Xg, const void* W, int w_cols, int w_grouped, double* r2part, double* Pfull,
                          int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream);

/* r2[b][k][t] = sum_chunks r2part[b][chunk][k][t] (fixed order, deterministic).  Used on its own by the
 * frequency-sharded driver, which all-reduces r2 across ranks before oiva_source_model(n_chunks=1). */
int oiva_sum_partials(const double* r2part, int n_chunks, double* r2, int n_batch, int n_frames, int n_src,
                      void* stream);

/* source model + scale normalisation + clamp + inverse: overiva.py:152-173 (ive.py:204-214 for the
 * OGIVE models).  r2part: (B,n_chunks,K,Tp).  phi (B,K,Tp) <- 1/max(r/gamma, 1e-15); wscale (B,K) (may be
 * NULL) <- 1/gamma (laplace), 1/sqrt(gamma) (gauss), 1 (others): the factor W must be MULTIPLIED by
 * (overiva.py:161-167).  n_freq_total is the F the gauss model divides by (the full F when bins are
 * sharded across GPUs). */
int oiva_source_model(const double* r2part, int n_chunks, double* phi, double* wscale, int n_batch,
                      int n_frames, int n_src, int n_freq_total, int model, void* stream);

/* oiva_source_model with n_batch * n_src zeroed words of device memory (left zeroed on return; NULL = none): where the
 * frames are processed in parallel (long mixtures, or few (mixture, source) pairs) the last CTA of a pair then finishes it,
 * one launch instead of two. */
int oiva_source_model_ws(const double* r2part, int n_chunks, double* phi, double* wscale, unsigned* counters, int n_batch,
                         int n_frames, int n_src, int n_freq_total, int model, void* stream);

/* One sweep over the K sources, per bin, in order:  W[:, :K] *= wscale;  for s: w_s = (What^H V_s)^-1 e_s;
 * w_s /= sqrt(w_s^H V_s w_s);  J = (W^H C E1)^-1 (W^H C E2).   Wg: the W_hat matrices in the GROUPED layout
 * [gi][M*M][32] (oiva_group_rows of the (R,M,M) array), updated in place; Vg grouped; C (R,M,M) c128 full,
 * Cg = the same covariance in the grouped lower-triangle layout (may be NULL).
 * With Cg given and M <= 6 the sweep runs one THREAD per bin (lane <-> bin): for K < M the update is
 * w = V^-1 q with q = W_hat^-H e_s from a K x K pivoted solve and V factorised by Cholesky (same result as
 * the reference's (W_hat^H V)^-1 e_s, ~3x fewer operations); for K = M an in-register LU with partial
 * pivoting.  Otherwise a group of next_pow2(M) lanes owns a bin (one matrix row per lane, Gauss-Jordan).
 * replaces: overiva.py:161-167 (W rescale), :181-182 (zgemm + zgesv), :185-186, :189-190 (:96-98). */
int oiva_ip_update(void* Wg, const void* Vg, const void* C, const void* Cg, const double* wscale, int* status,
                   int n_batch, int n_freq, int n_chan, int n_src, void* stream);

/* oiva_weighted_cov + oiva_ip_update in ONE kernel: the pass over X that accumulates V_1..V_K of a bin ends, per bin
 * group, in the thread-per-bin sweep of that group with the covariances still in registers (Vg is never written or
 * re-read).  Wg is bit-identical to the two-call sequence.  Covered: M <= 8 with all K lower triangles in one lane's
 * registers and K <= 3 ((M <= 5, K <= min(M, 3)), (6, K <= 2), (7 | 8, 1)) and n_batch * ceil(F/32) >= 4096 bin groups (no
 * frame splitting); otherwise OIVA_ERR_UNSUPPORTED is returned, nothing is launched and no error text is set -- run the
 * two calls instead.  Arguments as for those two calls.
 * replaces: overiva.py:161-167 and :176-190 (the whole source loop of an epoch). */
int oiva_cov_ip_update(const void* Xg, const double* phi, void* Wg, const void* Cg, const double* wscale, int* status,
                       int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream);

/* Build What (R,M,M): W from eye / eigenvectors / W0, then J and the -I block.
 * evecs: (R,M,M) from oiva_eigh (ascending; used when mode == OIVA_INIT_EIG: w_k = conj(v_{M-K+k})).
 * W0: (R,M,K) c128 (mode == OIVA_INIT_W0).  status: n_rows / rows_per_mixture words (row r reports to word
 * r / rows_per_mixture).                                      replaces: overiva.py:89-123 */
int oiva_init_demix(void* What, const void* C, const void* W0, const void* evecs, int mode, int* status,
                    int n_rows, int rows_per_mixture, int n_chan, int n_src, void* stream);

/* The same for the identity / W0 initialisations, written straight into the GROUPED layout Wg[gi][M*M][32] by one
 * thread per bin from the grouped covariance Cg (W0: (R,M,K) c128 or NULL = identity; status: n_batch words).
 * OIVA_ERR_UNSUPPORTED (nothing launched, no error text) outside the thread-per-bin shapes (M <= 6, and M = 7, 8 with
 * K <= 4): use oiva_init_demix + oiva_group_rows there. */
int oiva_init_demix_grouped(void* Wg, const void* Cg, const void* W0, int* status, int n_batch, int n_freq, int n_chan,
                            int n_src, void* stream);

/* Hermitian eigendecomposition per row (cyclic Jacobi, fp64): evals (R,M) ascending, evecs (R,M,M) with
 * eigenvectors in columns.  lapack_phase != 0 rotates each eigenvector so that its largest-magnitude
 * component is real positive (what zgeev, i.e. np.linalg.eig, returns).  status as for oiva_init_demix.
 * replaces: overiva.py:106 / ive.py:110 (np.linalg.eig) and auxiva_pca.py:75 (np.linalg.eigh). */
int oiva_eigh(const void* C, double* evals, void* evecs, int* status, int n_rows, int rows_per_mixture,
              int n_chan, int lapack_phase, void* stream);

/* W: (R,M,w_cols) c128.  Weff (R,M,K) c128 = W[:, :, k] * z_k with z_k = (w_k^H C e_0)/(w_k^H C w_k)
 * (1 if the denominator is 0) when proj_back != 0, else a plain copy of the W columns.
 * replaces: pyroomacoustics.bss.projection_back as called at overiva.py:197-199 (no pass over Y needed:
 * sum_t conj(x_0) y_k = T w_k^H C e_0 and sum_t |y_k|^2 = T w_k^H C w_k). */
int oiva_projback_filters(const void* W, int w_cols, const void* C, void* Weff, int n_rows, int n_chan,
                          int n_src, int proj_back, void* stream);

/* Y (B,T,F,K) interleaved complex (dtype) = Weff^H x, Weff (R,M,K) c128.   replaces: overiva.py:192-199 */
int oiva_demix_output(const void* Xg, const void* Weff, void* Y, int n_batch, int n_frames, int n_freq,
                      int n_chan, int n_src, int dtype, void* stream);

/* The same straight from the loop's grouped state: Wg = the grouped W_hat [gi][M*M][32] (columns :K are the filters).
 * With Cg (grouped lower-triangle input covariance) the projection-back scales z_k of oiva_projback_filters are computed
 * per bin by a small kernel into zscratch (n_batch * ceil(F/32) * n_src * 32 c128 of device memory, required then; same
 * arithmetic and order) and folded into the filters by the output kernel; Cg == NULL: no projection back.  Two launches,
 * no row-major copies, for overiva.py:192-199. */
int oiva_demix_output_grouped(const void* Xg, const void* Wg, const void* Cg, void* zscratch, void* Y, int n_batch,
                              int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream);

/* Xr = grouped samples (K channels) of E_K^H x with E_K (R,M,K) c128 -- the PCA projection of
 * auxiva_pca.py:79-81 written directly in the layout the loop streams. */
int oiva_project_rows(const void* Xg, const void* E, void* Xr, int n_batch, int n_frames, int n_freq,
                      int n_chan, int n_src, int dtype, void* stream);

/* Wout (R,M,K) = E (R,M,Kr) @ Wr (R,Kr,Kr)[:, :, :K]  (auxiva_pca: filters on the original channels) */
int oiva_compose_filters(const void* E, const void* Wr, void* Wout, int n_rows, int n_chan, int n_red,
                         int n_src, void* stream);

/* OGIVE per-bin update (ive.py:132-140, 216-241): from V (R,M,M) full (oiva_unpack_cov of the K=1 weighted
 * covariance) and the state (w, a, lambda_a), one masked w-step / a-step with the orthogonal constraints;
 * delta_max[0] <- max_f ||delta_f|| via an atomic max on its bit pattern (caller zeroes it).
 * do_a: (R,) uint8 mask (1 = a-step bin). */
int oiva_ogive_update(void* w, void* a, double* lambda_a, const void* V, const void* C, const void* Cinv,
                      const uint8_t* do_a, double step_size, double* delta_max, int n_rows, int n_chan,
                      void* stream);
/* The same update for epoch `epoch` with the stopping rule of ive.py:238-241 evaluated on the device: delta_hist is a
 * zero-initialised array of at least epoch + 1 doubles; delta_hist[epoch] receives max_f ||delta_f||, and once
 * delta_hist[epoch - 1] < tol (the epoch at which the reference leaves its loop) the update is a no-op that carries
 * the value forward.  The host may therefore look at delta_hist every few epochs instead of after each one and still
 * ends in the state the reference stops in. */
int oiva_ogive_update_gated(void* w, void* a, double* lambda_a, const void* V, const void* C, const void* Cinv,
                            const uint8_t* do_a, double step_size, double* delta_hist, int epoch, double tol,
                            int n_rows, int n_chan, void* stream);
/* n_epochs epochs of the loop in one call (oiva_demix_power -> oiva_source_model -> oiva_weighted_cov_ws (K = 1) ->
 * oiva_unpack_cov -> oiva_ogive_update_gated with epoch = epoch0 + e); work arrays as for those calls.   ive.py:191-241 */
int oiva_ogive_iterate(const void* Xg, void* w, void* a, double* lambda_a, double* r2part, double* phi, void* Vg,
                       void* cov_scratch, size_t cov_scratch_bytes, void* V, const void* C, const void* Cinv,
                       const uint8_t* do_a, double step_size, double* delta_hist, int epoch0, int n_epochs, double tol,
                       int n_frames, int n_freq, int n_chan, int model, int dtype, void* stream);
/* OGIVE set-up: Cinv = C^-1 (ive.py:98), cnorm (R,) = ||C||_F (ive.py:99); a from w (ive.py:132-135,168);
 * switching criterion masks (ive.py:142-161). */
int oiva_ogive_setup(const void* C, void* Cinv, double* cnorm, int* status, int n_rows, int n_chan, void* stream);
int oiva_ogive_a_from_w(const void* w, void* a, const void* C, int n_rows, int n_chan, void* stream);
int oiva_ogive_switching(const void* a, const void* C, const double* cnorm, uint8_t* do_a, int n_rows,
                         int n_chan, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ILRMA: the NMF source model around the shared demix / covariance / sweep kernels (csrc/ilrma.cu).
 * replaces: pyroomacoustics.bss.ilrma(X, n_iter, n_components, proj_back, callback) as called at
 * overiva_oneshot.py:331-339 and overiva_sim.py:309-311 (third-party, absent from the reference tree; restated from the
 * published algorithm in oracle/ilrma_oracle.py -- parity unpinned).  Determined (n_src == n_chan), n_chan <= 8.
 * Arrays (device, float64): Pg, iRg grouped [gi][k][Tp][32] (powers |y|^2 and 1 / r_k(f,t)); Tg grouped [gi][k][L][32]
 * (NMF bases); Vn (B, K, Tp, L) (NMF activations); Vpart oiva_ilrma_vpart_bytes(); lam (B, K).  L = n_components <= 8.
 * One epoch = oiva_ilrma_nmf -> oiva_weighted_cov_binwise(iRg) -> oiva_ip_update -> oiva_demix_power_full ->
 * oiva_ilrma_rescale.
 * ---------------------------------------------------------------------------------------------- */
size_t oiva_ilrma_vpart_bytes(int n_batch, int n_frames, int n_freq, int n_src, int n_comp);
/* T0 (B,K,F,L), V0 (B,K,T,L) row-major device arrays -> Tg, Vn and iRg = 1 / (T V^T) */
int oiva_ilrma_set_model(const double* T0, const double* V0, double* Tg, double* Vn, double* iRg, int n_batch, int n_frames,
                         int n_freq, int n_src, int n_comp, void* stream);
int oiva_ilrma_get_model(const double* Tg, const double* Vn, double* Tout, double* Vout, int n_batch, int n_frames,
                         int n_freq, int n_src, int n_comp, void* stream);
/* the multiplicative updates of all K sources: T *= sqrt((P r^-2) V / (r^-1 V)), clamp at eps, r = T V^T;
 * V *= sqrt((P r^-2)^T T / (r^-1)^T T), clamp, r = T V^T; iRg <- 1 / r */
int oiva_ilrma_nmf(const double* Pg, double* iRg, double* Tg, double* Vn, double* Vpart, int n_batch, int n_frames,
                   int n_freq, int n_src, int n_comp, double eps, void* stream);
/* lam[b][k] = 1 / sqrt(mean_{f,t} P_k) from the statistic partials r2part (B,NG,K,Tp); then w_k *= lam (scale_w != 0),
 * P_k, T_k *= lam^2, iR_k /= lam^2 */
int oiva_ilrma_rescale(const double* r2part, double* lam, void* Wg, double* Pg, double* iRg, double* Tg, int n_batch,
                       int n_frames, int n_freq, int n_chan, int n_src, int n_comp, int scale_w, void* stream);
/* n_epochs whole epochs (nmf -> weighted_cov_binwise -> ip_update -> demix_power_full -> rescale) in one call; C / Cg:
 * the input covariance, row-major (R,M,M) and grouped, as for oiva_ip_update; scratch as for oiva_weighted_cov_ws */
int oiva_ilrma_iterate(const void* Xg, void* Wg, void* Vg, const void* C, const void* Cg, double* r2part, void* scratch,
                       size_t scratch_bytes, double* Pg, double* iRg, double* Tg, double* Vn, double* Vpart, double* lam,
                       int* status, int n_batch, int n_frames, int n_freq, int n_chan, int n_comp, double eps, int n_epochs,
                       void* stream);
/* Zg [gi][K][32] c128 <- lam[b][k] (invert != 0: 1 / lam): a real per-source scale for oiva_demix_output_scaled */
int oiva_ilrma_fill_scale(const double* lam, void* Zg, int n_batch, int n_freq, int n_src, int invert, void* stream);
/* oiva_demix_output_grouped with caller-supplied per-bin scales Zg [gi][K][32] c128 (NULL: none): Y = (w_k z_k)^H x */
int oiva_demix_output_scaled(const void* Xg, const void* Wg, const void* Zg, void* Y, int n_batch, int n_frames, int n_freq,
                             int n_chan, int n_src, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * STFT analysis / synthesis on the device (the step either side of the loop in the reference's drivers).
 * replaces: pra.transform.analysis(mix.T, 4096, 2048, win=win_a)   overiva_oneshot.py:293-295, overiva_sim.py:206-207
 *           pra.transform.synthesis(Y, 4096, 2048, win=win_s)      overiva_oneshot.py:371-379, overiva_sim.py:213-218
 * (pyroomacoustics is third-party and absent from the reference tree: X[t] = rfft(win * frame_t), no scaling;
 * synthesis = overlap-add of win * irfft(Y[t]).)  frame_len: a power of two in 8..8192; F = frame_len/2 + 1.
 * ---------------------------------------------------------------------------------------------- */
/* tw (oiva_stft_twiddle_bytes(frame_len) bytes) <- the twiddle factors both transforms need, one compact block
 * per FFT pass (csrc/stft.cu) */
size_t oiva_stft_twiddle_bytes(int frame_len);
int oiva_stft_twiddles(void* tw, int frame_len, void* stream);
/* frames of length frame_len every hop samples over pad_front zeros + the signal + pad_back zeros */
int oiva_stft_num_frames(long long n_samples, int frame_len, int hop, long long pad_front, long long pad_back);
/* x: real audio, sample (b, n, c) at x[b*stride_b + n*stride_n + c*stride_c] (elements; fp64, or fp32 if x_f32);
 * frame t covers samples t*hop - pad_front + [0, frame_len), samples outside [0, n_samples) read as zero.
 * win: (frame_len) fp64 or NULL.  out: grouped != 0 -> grouped samples Xg (what oiva_plan_samples() holds, then
 * oiva_plan_adopt_samples); grouped == 0 -> X (B,T,F,M) interleaved complex.  dtype: storage of out. */
int oiva_stft_analysis(const void* x, int x_f32, long long stride_b, long long stride_n, long long stride_c,
                       long long n_samples, long long pad_front, const double* win, const void* tw, void* out,
                       int grouped, int n_batch, int n_frames, int n_chan, int frame_len, int hop, int dtype,
                       void* stream);
/* Y (B,T,F,K) complex -> y (B, (T-1)*hop + frame_len, K) real (fp64, or fp32 if y_f32): overlap-add of
 * win * irfft(Y[b,t,:,k]) in ascending frame order (deterministic).  scratch: oiva_stft_scratch_bytes() bytes. */
size_t oiva_stft_scratch_bytes(int n_batch, int n_frames, int n_src, int frame_len);
int oiva_stft_synthesis(const void* Y, const double* win, const void* tw, void* scratch, void* y, int y_f32,
                        int n_batch, int n_frames, int n_src, int frame_len, int hop, int dtype, void* stream);

/* out (R,R) fp64, R = a_rows + b_rows <= 64: Gram matrix of the real signals a[r*a_row_stride + n*a_sample_stride]
 * (rows 0..a_rows-1) and b[...] (the following b_rows rows; b may be NULL with b_rows = 0) over n_samples samples
 * -- the one pass over the audio an SDR / SIR evaluation needs.  Deterministic (two fixed-order passes).
 * scratch: oiva_gram_scratch_bytes(R, n_samples) bytes.
 * replaces: the inner products inside mir_eval.separation.bss_eval_sources as called by the drivers'
 * convergence_callback (overiva_oneshot.py:263-284, overiva_sim.py:210-232). */
size_t oiva_gram_scratch_bytes(int n_rows, long long n_samples);
int oiva_gram(const double* a, long long a_row_stride, long long a_sample_stride, int a_rows, const double* b,
              long long b_row_stride, long long b_sample_stride, int b_rows, long long n_samples, void* scratch,
              double* out, void* stream);

/* Cross-correlations at lags 0 .. flen-1 (flen <= 1024): out (a_rows, a_rows + b_rows, flen) float64 with
 * out[i][j][m] = sum_n s_i[n] s_j[n + m], s_i a row of a, s_j a row of a (j < a_rows) or of b; signals as for oiva_gram
 * (element strides in doubles).  With a = references and b = estimates these are all the inner products
 * mir_eval.separation.bss_eval_sources (the reference callback's metric, overiva_oneshot.py:263-284: 512-tap distortion
 * filters) needs: the block-Toeplitz normal matrix and its right-hand sides.  Deterministic (fixed-order tile sums).
 * scratch: oiva_xcorr_scratch_bytes(...) bytes. */
size_t oiva_xcorr_scratch_bytes(int a_rows, int b_rows, long long n_samples, int flen);
int oiva_xcorr(const double* a, long long a_row_stride, long long a_sample_stride, int a_rows, const double* b,
               long long b_row_stride, long long b_sample_stride, int b_rows, long long n_samples, int flen, void* scratch,
               double* out, void* stream);

/* n_iter epochs of the loop (overiva.py:138-190) in ONE persistent cooperative launch, for inputs short enough that
 * every bin group's samples fit the shared memory of the SMs (one short mixture: BASELINE configs 1-2; M <= 8): the grid
 * stays resident, the samples are read from HBM once and kept in shared memory for all epochs, the epochs are separated
 * by grid barriers instead of kernel boundaries (csrc/resident.cuh).  Same arguments as the separate kernels: Xg grouped
 * samples, Wg grouped W_hat (in/out), Cg grouped input covariance, r2part (B,NG,K,Tp) and rbuf (B,K,Tp) scratch,
 * scratch = oiva_weighted_cov_scratch_bytes() bytes of partial-covariance slots, sync = oiva_loop_resident_sync_bytes()
 * bytes (zeroed by the call), status = n_batch words.  Returns OIVA_ERR_UNSUPPORTED (nothing launched) when the shape
 * does not fit; callers then run the kernel-per-step loop.
 * replaces: overiva.py:138-190 (the whole epoch loop). */
size_t oiva_loop_resident_sync_bytes(int n_batch, int n_freq);
int oiva_loop_resident(const void* Xg, void* Wg, const void* Cg, double* r2part, double* rbuf, void* scratch,
                       size_t scratch_bytes, void* sync, int* status, int n_batch, int n_frames, int n_freq,
                       int n_freq_total, int n_chan, int n_src, int model, int dtype, int n_iter, void* stream);

/* fp64 throughput of this GPU in TFLOP/s, measured: kind 0 = plain DFMA, kind 1 = the fp64 tensor-core instruction
 * (DMMA, mma.sync.m8n8k4).  Synchronous (best of `reps` timed launches on `stream`).  The denominator of the
 * FMA-bound covariance shapes (M = 16, K = 4) and the DMMA-vs-DFMA evaluation of SURVEY.md 7.1(12). */
int oiva_fp64_peak(int kind, int iters, int reps, double* tflops, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Plan: the whole overiva() call on device pointers (what the Python entry points use).
 * The plan owns no device memory: the caller provides one workspace block.
 * ---------------------------------------------------------------------------------------------- */
typedef struct oiva_plan oiva_plan_t;

typedef struct oiva_plan_desc {
    int n_batch;      /* B independent mixtures                                        */
    int n_frames;     /* T                                                             */
    int n_freq;       /* F bins held by this plan (a shard of the full F when sharded) */
    int n_freq_total; /* F of the whole mixture (gauss model divisor); 0 => n_freq     */
    int n_chan;       /* M                                                             */
    int n_src;        /* K                                                             */
    int model;        /* OIVA_MODEL_*                                                  */
    int dtype;        /* OIVA_C128 / OIVA_C64: storage of X and Y                      */
    int flags;        /* reserved, 0                                                   */
} oiva_plan_desc;

int oiva_plan_create(oiva_plan_t** plan, const oiva_plan_desc* desc);
void oiva_plan_destroy(oiva_plan_t* plan);
size_t oiva_plan_workspace_bytes(const oiva_plan_t* plan);
int oiva_plan_bind(oiva_plan_t* plan, void* workspace, size_t bytes);

/* relayout X and compute the input covariance C                    overiva.py:87,131-132 */
int oiva_plan_load(oiva_plan_t* plan, const void* X, void* stream);
/* alternative to oiva_plan_load: the caller has written grouped samples into oiva_plan_samples() (e.g.
 * with oiva_project_rows); computes their covariance and marks the plan loaded. */
int oiva_plan_adopt_samples(oiva_plan_t* plan, void* stream);
/* initialise What (mode OIVA_INIT_*; W0 (B,F,M,K) c128 or NULL)    overiva.py:89-123 */
int oiva_plan_init(oiva_plan_t* plan, int mode, const void* W0, void* stream);
/* n_iter epochs of the loop                                        overiva.py:138-190 */
int oiva_plan_iterate(oiva_plan_t* plan, int n_iter, void* stream);
/* the two halves of one epoch, for the frequency-sharded driver: power() leaves the partial
 * statistic summed over this plan's bins in r2 (oiva_plan_r2()); the caller all-reduces it across
 * ranks; update() then runs source model + weighted covariance + IP sweep from r2. */
int oiva_plan_power(oiva_plan_t* plan, void* stream);
int oiva_plan_update(oiva_plan_t* plan, void* stream);
double* oiva_plan_r2(oiva_plan_t* plan);   /* (B,K,Tp) doubles */
size_t oiva_plan_r2_elems(const oiva_plan_t* plan);
/* final demix (+ projection back) into Y (B,T,F,K)                 overiva.py:192-199 */
int oiva_plan_output(oiva_plan_t* plan, int proj_back, void* Y, void* stream);
/* copy the filters W (B,F,M,K) c128 (contiguous) out of What       overiva.py:201-202 */
int oiva_plan_filters(oiva_plan_t* plan, void* W, void* stream);
/* oiva_plan_load + _init + _iterate + _output (+ _filters if W != NULL) in one call (one foreign-function round trip
 * instead of five: for one short mixture the host language's call overhead is comparable to the GPU work) */
int oiva_plan_run(oiva_plan_t* plan, const void* X, int init_mode, const void* W0, int n_iter, int proj_back, void* Y,
                  void* W, void* stream);
/* device pointers into the workspace (for tests and wrappers) */
void* oiva_plan_what(oiva_plan_t* plan);    /* (R,M,M) c128 row-major copy of W_hat: written by oiva_plan_filters and by
                                               an oiva_plan_init that takes the row-major path (eig; shapes outside
                                               the thread-per-bin kernels); the loop itself works on the grouped array
                                               (oiva_plan_array(plan, 0)) */
void* oiva_plan_cov(oiva_plan_t* plan);     /* (R,M,M) c128, full: valid after oiva_plan_load / oiva_plan_adopt_samples
                                               (oiva_plan_run produces it only when something on its path reads it) */
void* oiva_plan_samples(oiva_plan_t* plan); /* Xg */
/* the status words: n_batch ints, one per mixture.  The CALLER zeroes them after oiva_plan_bind (the workspace is
 * uninitialised memory; oiva_plan_reset_status does it); they then accumulate (atomic OR) over everything the plan
 * runs, across loads, until the caller zeroes them again */
int* oiva_plan_status_ptr(oiva_plan_t* plan);
int oiva_plan_reset_status(oiva_plan_t* plan, void* stream);
/* synchronises the stream and returns the OR of all mixtures' status words (0 = fine, > 0 = OIVA_STATUS_* bits,
 * < 0 = OIVA_ERR_*); status_host (may be NULL) receives the n_batch per-mixture words */
int oiva_plan_status(oiva_plan_t* plan, void* stream);
int oiva_plan_status_vector(oiva_plan_t* plan, int* status_host, void* stream);
/* number of kernels launched by this plan since creation (bench.py's gpu_launches) */
long long oiva_plan_launch_count(const oiva_plan_t* plan);
/* device pointers into the bound workspace for callers that sequence kernels themselves on the plan's arrays (ILRMA):
 * which = 0 grouped W_hat (G, M*M, 32) c128 | 1 grouped input covariance (G, NE, 32) | 2 grouped weighted covariances
 * (G, K, NE, 32) | 3 statistic partials r2part (B, NG, K, Tp) f64 | 4 frame-split scratch of the covariance kernel
 * (oiva_plan_scratch_bytes bytes; NULL if none) | 5 scratch for output scales (G, K, 32) c128 */
void* oiva_plan_array(oiva_plan_t* plan, int which);
size_t oiva_plan_scratch_bytes(const oiva_plan_t* plan);

/* Optional per-kernel timing for the roofline report: when enabled, CUDA events are recorded on the
 * launching stream around every launch of the three loop kernels.  oiva_plan_read_timing() (call it after
 * synchronising the stream) returns the summed milliseconds and launch counts for
 * [0] weighted covariance, [1] demix power, [2] IP solve, and resets the record. */
int oiva_plan_enable_timing(oiva_plan_t* plan, int enable);
int oiva_plan_read_timing(oiva_plan_t* plan, double* ms3, long long* counts3);

/* Convenience: the complete call with HOST buffers (pinned or pageable): H2D of X (B,T,F,M), the loop,
 * D2H of Y (B,T,F,K) [and W (B,F,M,K) c128 if W_host != NULL].  Allocates and frees its own device memory.
 * Returns 0, a negative OIVA_ERR_* code, or (> 0) the OR of the mixtures' status words on numerical failure;
 * status_host (may be NULL) receives the n_batch per-mixture words. */
int oiva_overiva_host(const void* X_host, void* Y_host, void* W_host, const void* W0_host,
                      const oiva_plan_desc* desc, int n_iter, int proj_back, int init_mode, int* status_host);

#ifdef __cplusplus
}
#endif
#endif /* OVERIVA_B200_H */
