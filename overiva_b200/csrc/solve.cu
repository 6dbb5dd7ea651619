// Per-bin kernels: IP sweep, demixing-matrix initialisation, projection-back scaling, Hermitian eigh.
// (include/overiva_b200.h: oiva_ip_update, oiva_init_demix, oiva_projback_filters, oiva_eigh,
//  oiva_compose_filters)
#include <stdlib.h>

#include "solve.cuh"

namespace oiva {

constexpr int SOLVE_WARPS = 4;

// ---------------------------------------------------------------------------------------------------
// IP sweep over the K sources of every bin (overiva.py:161-167 W rescale, :176-190 source loop)
// ---------------------------------------------------------------------------------------------------
template <int M>
__host__ __device__ constexpr int ip_warps() { return M > 8 ? 2 : 4; }

// V arrives in the grouped lower-triangle layout Vg[gi][k][e][l] written by the covariance kernel.
template <int M>
__global__ void __launch_bounds__(ip_warps<M>() * 32) k_ip_update(cplx* __restrict__ Wg, const cplx* __restrict__ Vg,
                                                                  const cplx* __restrict__ C,
                                                                  const double* __restrict__ wscale, int* status,
                                                                  long long R, int F, int NG, int K) {
    constexpr int G = Grp<M>::G, BINS = Grp<M>::BINS, WARPS = ip_warps<M>(), NE = oiva_tri(M);
    __shared__ cplx sWall[WARPS * BINS * M * M];
    __shared__ cplx sVall[WARPS * BINS * M * M];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;
    const long long row = ((long long)blockIdx.x * WARPS + warp) * BINS + grp;
    const bool row_ok = row < R;
    const long long rowc = row_ok ? row : R - 1;
    cplx* sW = sWall + (size_t)(warp * BINS + grp) * M * M;
    cplx* Vs = sVall + (size_t)(warp * BINS + grp) * M * M;
    const bool rv = gl < M;
    const long long bmix = rowc / F;
    const int fbin = (int)(rowc - bmix * F);
    const cplx* Vbin = Vg + ((size_t)(bmix * NG + fbin / OIVA_GROUP) * K * NE) * OIVA_GROUP + fbin % OIVA_GROUP;
    // this bin's W_hat inside the grouped array Wg[gi][M*M][32]
    cplx* Wbin = Wg + ((size_t)(bmix * NG + fbin / OIVA_GROUP) * M * M) * OIVA_GROUP + fbin % OIVA_GROUP;

    if (rv) {
#pragma unroll
        for (int c = 0; c < M; ++c) sW[gl * M + c] = Wbin[(size_t)(gl * M + c) * OIVA_GROUP];
        if (wscale) {
            const long long b = rowc / F;
            for (int c = 0; c < K; ++c) sW[gl * M + c] = cscale(sW[gl * M + c], wscale[b * K + c]);
        }
    }
    __syncwarp();

    int singular = 0;
    for (int s = 0; s < K; ++s) {
        // unpack V_s (Hermitian, lower triangle stored) into the bin's shared tile
        for (int idx = gl; idx < M * M; idx += G) {
            const int j = idx / M, c = idx - j * M;
            const int hi = j >= c ? j : c, lo = j >= c ? c : j;
            cplx v = ld_nc_c(Vbin + ((size_t)s * NE + hi * (hi + 1) / 2 + lo) * OIVA_GROUP);
            if (j < c) v.y = -v.y;
            Vs[idx] = v;
        }
        __syncwarp();
        // row gl of What^H V_s, augmented with e_s
        cplx A[M + 1];
#pragma unroll
        for (int c = 0; c <= M; ++c) A[c] = cmake(0.0, 0.0);
        if (rv) {
            for (int j = 0; j < M; ++j) {
                const cplx a = sW[j * M + gl];
#pragma unroll
                for (int c = 0; c < M; ++c) cfmac(A[c], a, Vs[j * M + c]);
            }
            if (gl == s) A[M] = cmake(1.0, 0.0);
        }
        const int col = gauss_jordan<M, G>(A, M, gl, lane, rv, singular);
        __syncwarp();
        if (col >= 0) sW[col * M + s] = A[M];
        __syncwarp();
        // normalise: w_s /= sqrt(w_s^H V_s w_s)
        cplx wi = cmake(0.0, 0.0), u = cmake(0.0, 0.0);
        if (rv) {
            wi = sW[gl * M + s];
#pragma unroll
            for (int j = 0; j < M; ++j) cfma(u, Vs[gl * M + j], sW[j * M + s]);
        }
        const cplx d = group_sum<G>(cmulc(wi, u));
        const cplx inv = crecip(csqrt_(d));
        __syncwarp();
        if (rv) sW[gl * M + s] = cmul(wi, inv);
        __syncwarp();
        if (K < M) update_background<M, G>(sW, C + (size_t)rowc * M * M, K, gl, lane, singular);
    }

    bool bad = false;
    if (rv) {
#pragma unroll
        for (int c = 0; c < M; ++c) {
            const cplx v = sW[gl * M + c];
            if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
            if (row_ok) Wbin[(size_t)(gl * M + c) * OIVA_GROUP] = v;
        }
    }
    if (row_ok && (singular || bad))
        atomicOr(status + bmix, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
}

// ---------------------------------------------------------------------------------------------------
// initial W_hat (overiva.py:89-123)
// ---------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(SOLVE_WARPS * 32) k_init_demix(cplx* __restrict__ What, const cplx* __restrict__ C,
                                                                 const cplx* __restrict__ W0,
                                                                 const cplx* __restrict__ evecs, int mode, int* status,
                                                                 long long R, int rows_per_mixture, int K) {
    constexpr int G = Grp<M>::G, BINS = Grp<M>::BINS;
    __shared__ cplx sWall[SOLVE_WARPS * BINS * M * M];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;
    const long long row = ((long long)blockIdx.x * SOLVE_WARPS + warp) * BINS + grp;
    const bool row_ok = row < R;
    const long long rowc = row_ok ? row : R - 1;
    cplx* sW = sWall + (size_t)(warp * BINS + grp) * M * M;
    const bool rv = gl < M;
    if (rv) {
#pragma unroll
        for (int c = 0; c < M; ++c) {
            cplx v = cmake(0.0, 0.0);
            if (c < K) {
                if (mode == OIVA_INIT_W0)
                    v = W0[((size_t)rowc * M + gl) * K + c];
                else if (mode == OIVA_INIT_EIG)
                    v = cconj(evecs[((size_t)rowc * M + gl) * M + (M - K + c)]);
                else
                    v = cmake(gl == c ? 1.0 : 0.0, 0.0);
            }
            sW[gl * M + c] = v;
        }
    }
    __syncwarp();
    int singular = 0;
    if (K < M) {
        update_background<M, G>(sW, C + (size_t)rowc * M * M, K, gl, lane, singular);
        if (rv && gl >= K) sW[gl * M + gl] = cmake(-1.0, 0.0);
        __syncwarp();
    }
    if (row_ok && rv) {
#pragma unroll
        for (int c = 0; c < M; ++c) What[row * M * M + gl * M + c] = sW[gl * M + c];
    }
    if (row_ok && singular) atomicOr(status + row / rows_per_mixture, OIVA_STATUS_SINGULAR);
}

// ---------------------------------------------------------------------------------------------------
// projection back folded into the filters: Weff[:, k] = w_k * z_k, z_k = (w_k^H C e_0)/(w_k^H C w_k)
// (pyroomacoustics.bss.projection_back as used at overiva.py:197-199); one thread per (row, k)
// ---------------------------------------------------------------------------------------------------
__global__ void k_projback_filters(const cplx* __restrict__ What, int wc, const cplx* __restrict__ C,
                                   cplx* __restrict__ Weff, long long R, int M, int K, int proj_back) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * K) return;
    const long long row = i / K;
    const int k = (int)(i - row * K);
    const cplx* W = What + (size_t)row * M * wc;
    const cplx* Cr = C + (size_t)row * M * M;
    cplx z = cmake(1.0, 0.0);
    if (proj_back) {
        cplx num = cmake(0.0, 0.0);
        double den = 0.0;
        for (int a = 0; a < M; ++a) {
            const cplx wa = W[a * wc + k];
            cfmac(num, wa, Cr[a * M + 0]);
            cplx cw = cmake(0.0, 0.0);
            for (int b = 0; b < M; ++b) cfma(cw, Cr[a * M + b], W[b * wc + k]);
            den += wa.x * cw.x + wa.y * cw.y;  // Re(conj(w_a) (C w)_a)
        }
        if (den > 0.0) z = cmake(num.x / den, num.y / den);
    }
    for (int a = 0; a < M; ++a) Weff[((size_t)row * M + a) * K + k] = cmul(W[a * wc + k], z);
}

// Wout (R,M,K) = E (R,M,Kr) @ Wr (R,Kr,Kr)[:, :, :K]  -- auxiva_pca: full-rank filters from the PCA basis
__global__ void k_compose_filters(const cplx* __restrict__ E, const cplx* __restrict__ Wr, cplx* __restrict__ Wout,
                                  long long R, int M, int Kr, int K) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * M * K) return;
    const long long row = i / (M * K);
    const int rem = (int)(i - row * M * K);
    const int m = rem / K, k = rem - m * K;
    cplx acc = cmake(0.0, 0.0);
    for (int j = 0; j < Kr; ++j) cfma(acc, E[((size_t)row * M + m) * Kr + j], Wr[((size_t)row * Kr + j) * Kr + k]);
    Wout[i] = acc;
}

// ---------------------------------------------------------------------------------------------------
// Hermitian eigendecomposition: cyclic Jacobi, one warp per matrix, matrices in shared memory
// ---------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(SOLVE_WARPS * 32) k_eigh(const cplx* __restrict__ C, double* __restrict__ evals,
                                                           cplx* __restrict__ evecs, int* status, long long R,
                                                           int rows_per_mixture, int lapack_phase) {
    __shared__ cplx sA[SOLVE_WARPS][M * M];
    __shared__ cplx sQ[SOLVE_WARPS][M * M];
    __shared__ int sPerm[SOLVE_WARPS][M];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * SOLVE_WARPS + warp;
    if (row >= R) return;  // whole warp exits together
    cplx* A = sA[warp];
    cplx* Q = sQ[warp];
    for (int i = lane; i < M * M; i += 32) {
        A[i] = C[(size_t)row * M * M + i];
        Q[i] = cmake((i / M == i % M) ? 1.0 : 0.0, 0.0);
    }
    __syncwarp();
    // Parallel-ordered Jacobi: a sweep is Me - 1 rounds (Me = M rounded up to even) of a round-robin tournament; the
    // Me / 2 index pairs of a round are disjoint, so their rotations are computed from the same A and applied together --
    // first to the columns (A <- A U, Q <- Q U), then to the rows (A <- U^H A) -- by all 32 lanes.  (The first version
    // applied one rotation at a time with M of 32 lanes busy: 28 dependent steps per sweep at M = 8 instead of 7.)
    const double eps = 1.1e-16;
    constexpr int Me = M + (M & 1), NP = Me / 2;
    __shared__ double sRot[SOLVE_WARPS][NP > 0 ? NP : 1][4];  // c, s, Re ph, Im ph  (c = 2: pair idle this round)
    __shared__ int sPair[SOLVE_WARPS][NP > 0 ? NP : 1][2];
    for (int sweep = 0; sweep < 40 && M > 1; ++sweep) {
        int rotated = 0;
        for (int round = 0; round < Me - 1; ++round) {
            if (lane < NP) {
                int p, q;
                if (lane == 0) {
                    p = Me - 1;
                    q = round;
                } else {
                    p = (round + lane) % (Me - 1);
                    q = (round - lane + (Me - 1)) % (Me - 1);
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                double c = 2.0, sn = 0.0, phx = 1.0, phy = 0.0;
                if (q < M) {  // (odd M: the pair with the padding index sits out)
                    const cplx apq = A[p * M + q];
                    const double app = A[p * M + p].x, aqq = A[q * M + q].x;
                    const double g = hypot(apq.x, apq.y);
                    if (g > eps * sqrt(fabs(app * aqq)) && !(g < 1e-300)) {
                        phx = apq.x / g;
                        phy = apq.y / g;
                        const double tau = (aqq - app) / (2.0 * g);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        sn = t * c;
                    }
                }
                sRot[warp][lane][0] = c;
                sRot[warp][lane][1] = sn;
                sRot[warp][lane][2] = phx;
                sRot[warp][lane][3] = phy;
                sPair[warp][lane][0] = p;
                sPair[warp][lane][1] = q;
            }
            __syncwarp();
            bool any = false;
            // columns p, q of A and Q, one (pair, row) per lane and pass
            for (int it = lane; it < NP * M; it += 32) {
                const int pr = it / M, r = it - pr * M;
                const double c = sRot[warp][pr][0];
                if (c > 1.5) continue;
                any = true;
                const double sn = sRot[warp][pr][1];
                const cplx ph = cmake(sRot[warp][pr][2], sRot[warp][pr][3]);
                const int p = sPair[warp][pr][0], q = sPair[warp][pr][1];
                const cplx sphc = cmake(sn * ph.x, -sn * ph.y);  // s conj(ph)
                const cplx cphc = cmake(c * ph.x, -c * ph.y);    // c conj(ph)
                const cplx ap = A[r * M + p], aq = A[r * M + q];
                A[r * M + p] = csub(cscale(ap, c), cmul(sphc, aq));
                A[r * M + q] = cadd(cscale(ap, sn), cmul(cphc, aq));
                const cplx qp = Q[r * M + p], qq = Q[r * M + q];
                Q[r * M + p] = csub(cscale(qp, c), cmul(sphc, qq));
                Q[r * M + q] = cadd(cscale(qp, sn), cmul(cphc, qq));
            }
            __syncwarp();
            // rows p, q of A
            for (int it = lane; it < NP * M; it += 32) {
                const int pr = it / M, r = it - pr * M;
                const double c = sRot[warp][pr][0];
                if (c > 1.5) continue;
                const double sn = sRot[warp][pr][1];
                const cplx ph = cmake(sRot[warp][pr][2], sRot[warp][pr][3]);
                const int p = sPair[warp][pr][0], q = sPair[warp][pr][1];
                const cplx sph = cmake(sn * ph.x, sn * ph.y), cph = cmake(c * ph.x, c * ph.y);
                const cplx ap = A[p * M + r], aq = A[q * M + r];
                A[p * M + r] = csub(cscale(ap, c), cmul(sph, aq));
                A[q * M + r] = cadd(cscale(ap, sn), cmul(cph, aq));
            }
            __syncwarp();
            if (lane < NP && sRot[warp][lane][0] < 1.5) {
                const int p = sPair[warp][lane][0], q = sPair[warp][lane][1];
                A[p * M + q] = cmake(0.0, 0.0);
                A[q * M + p] = cmake(0.0, 0.0);
                A[p * M + p].y = 0.0;
                A[q * M + q].y = 0.0;
            }
            __syncwarp();
            if (__any_sync(0xffffffffu, any)) ++rotated;
        }
        if (rotated == 0) break;
    }
    // ascending order of the eigenvalues (stable selection by rank), one lane per eigenvalue
    if (lane < M) {
        const double li = A[lane * M + lane].x;
        int rank = 0;
        for (int j = 0; j < M; ++j) {
            const double lj = A[j * M + j].x;
            if (lj < li || (lj == li && j < lane)) ++rank;
        }
        sPerm[warp][rank] = lane;
        if (!isfinite(li)) atomicOr(status + row / rows_per_mixture, OIVA_STATUS_NONFINITE);
    }
    __syncwarp();
    if (lane < M) {
        const int k = lane;  // output column
        const int src = sPerm[warp][k];
        evals[(size_t)row * M + k] = A[src * M + src].x;
        // normalise + phase convention
        double nrm2 = 0.0, best = -1.0;
        int ib = 0;
        for (int i = 0; i < M; ++i) {
            const cplx v = Q[i * M + src];
            const double m2 = v.x * v.x + v.y * v.y;
            nrm2 += m2;
            if (m2 > best) {
                best = m2;
                ib = i;
            }
        }
        const double inrm = 1.0 / sqrt(nrm2);
        cplx rot = cmake(inrm, 0.0);
        if (lapack_phase) {
            const cplx vb = Q[ib * M + src];
            const double mb = sqrt(best);
            rot = cmake(vb.x / mb * inrm, -vb.y / mb * inrm);  // conj(v_b)/|v_b| / ||v||
        }
        for (int i = 0; i < M; ++i) {
            cplx v = cmul(Q[i * M + src], rot);
            if (lapack_phase && i == ib) v.y = 0.0;
            evecs[((size_t)row * M + i) * M + k] = v;
        }
    }
}

}  // namespace oiva

using namespace oiva;

template <int M>
static int bins_per_cta() {
    return SOLVE_WARPS * Grp<M>::BINS;
}

namespace oiva {
int ip_update_tpb(int M, int K, cplx* What, const cplx* Vg, const cplx* Cg, const double* wscale, int* status, int F,
                  int NG, long long G, cudaStream_t st);
int ip_update_pair(int M, int K, cplx* Wg, const cplx* Vg, const double* wscale, int* status, int F, int NG, long long G,
                   cudaStream_t st);
}

extern "C" int oiva_ip_update(void* What, const void* V, const void* C, const void* Cg, const double* wscale,
                              int* status, int n_batch, int n_freq, int n_chan, int n_src, void* stream) {
    // What: the demixing matrices in the GROUPED layout Wg[gi][M*M][32] (oiva_group_rows of the (R,M,M) array)
    OIVA_REQUIRE(What && V && C && status, "oiva_ip_update: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_freq > 0 && n_src >= 1 && n_src <= n_chan, "oiva_ip_update: bad shape");
    const long long R = (long long)n_batch * n_freq;
    cudaStream_t st = (cudaStream_t)stream;
    const char* env = getenv("OIVA_SOLVER_ROWOWNER");
    const bool force_rowowner = env && *env && *env != '0';
    if (Cg && n_chan <= 8 && !force_rowowner) {  // thread-per-bin sweep (solve_tpb.cuh) where instantiated
        const int NG = oiva_bin_groups(n_freq);
        int rc = ip_update_tpb(n_chan, n_src, (cplx*)What, (const cplx*)V, (const cplx*)Cg, wscale, status, n_freq, NG,
                               (long long)n_batch * NG, st);
        if (rc != OIVA_ERR_INVALID) return rc;
        // determined sweep of 7 / 8 channels: two lanes per bin (solve_pair.cuh)
        rc = ip_update_pair(n_chan, n_src, (cplx*)What, (const cplx*)V, wscale, status, n_freq, NG, (long long)n_batch * NG, st);
        if (rc != OIVA_ERR_INVALID) return rc;
    }
    OIVA_DISPATCH_M(n_chan, {
        const int per = ip_warps<M_>() * Grp<M_>::BINS;
        k_ip_update<M_><<<(unsigned)((R + per - 1) / per), ip_warps<M_>() * 32, 0, st>>>(
            (cplx*)What, (const cplx*)V, (const cplx*)C, wscale, status, R, n_freq, oiva_bin_groups(n_freq), n_src);
    });
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

namespace oiva {
int init_demix_tpb(int M, int K, cplx* Wg, const cplx* Cg, const cplx* W0, int* status, int F, int NG, long long G,
                   cudaStream_t st);
}

// initial W_hat (identity or W0) written straight into the grouped layout by one thread per bin; OIVA_ERR_UNSUPPORTED
// (nothing launched, no error text) for the shapes the thread-per-bin kernels do not cover
extern "C" int oiva_init_demix_grouped(void* Wg, const void* Cg, const void* W0, int* status, int n_batch, int n_freq,
                                       int n_chan, int n_src, void* stream) {
    OIVA_REQUIRE(Wg && Cg && status, "oiva_init_demix_grouped: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_freq > 0 && n_src >= 1 && n_src <= n_chan, "oiva_init_demix_grouped: bad shape");
    if (n_chan > 8) return OIVA_ERR_UNSUPPORTED;
    const int NG = oiva_bin_groups(n_freq);
    return init_demix_tpb(n_chan, n_src, (cplx*)Wg, (const cplx*)Cg, (const cplx*)W0, status, n_freq, NG,
                          (long long)n_batch * NG, (cudaStream_t)stream);
}

extern "C" int oiva_init_demix(void* What, const void* C, const void* W0, const void* evecs, int mode, int* status,
                               int n_rows, int rows_per_mixture, int n_chan, int n_src, void* stream) {
    OIVA_REQUIRE(What && C && status, "oiva_init_demix: null pointer");
    OIVA_REQUIRE(n_rows > 0 && rows_per_mixture > 0 && n_src >= 1 && n_src <= n_chan, "oiva_init_demix: bad shape");
    OIVA_REQUIRE(mode != OIVA_INIT_W0 || W0, "oiva_init_demix: W0 missing");
    OIVA_REQUIRE(mode != OIVA_INIT_EIG || evecs, "oiva_init_demix: eigenvectors missing");
    const long long R = n_rows;
    cudaStream_t st = (cudaStream_t)stream;
    OIVA_DISPATCH_M(n_chan, {
        const int per = bins_per_cta<M_>();
        k_init_demix<M_><<<(unsigned)((R + per - 1) / per), SOLVE_WARPS * 32, 0, st>>>(
            (cplx*)What, (const cplx*)C, (const cplx*)W0, (const cplx*)evecs, mode, status, R, rows_per_mixture, n_src);
    });
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_projback_filters(const void* What, int w_cols, const void* C, void* Weff, int n_rows, int n_chan,
                                     int n_src, int proj_back, void* stream) {
    OIVA_REQUIRE(What && C && Weff, "oiva_projback_filters: null pointer");
    OIVA_REQUIRE(w_cols >= n_src, "oiva_projback_filters: w_cols %d < n_src %d", w_cols, n_src);
    OIVA_REQUIRE(n_rows > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M && n_src >= 1 && n_src <= n_chan,
                 "oiva_projback_filters: bad shape");
    const long long n = (long long)n_rows * n_src;
    k_projback_filters<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const cplx*)What, w_cols, (const cplx*)C, (cplx*)Weff, n_rows, n_chan, n_src, proj_back);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_compose_filters(const void* E, const void* Wr, void* Wout, int n_rows, int n_chan, int n_red,
                                    int n_src, void* stream) {
    OIVA_REQUIRE(E && Wr && Wout, "oiva_compose_filters: null pointer");
    OIVA_REQUIRE(n_rows > 0 && n_chan >= 1 && n_red >= 1 && n_src >= 1 && n_src <= n_red,
                 "oiva_compose_filters: bad shape");
    const long long n = (long long)n_rows * n_chan * n_src;
    k_compose_filters<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const cplx*)E, (const cplx*)Wr, (cplx*)Wout, n_rows, n_chan, n_red, n_src);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_eigh(const void* C, double* evals, void* evecs, int* status, int n_rows, int rows_per_mixture,
                         int n_chan, int lapack_phase, void* stream) {
    OIVA_REQUIRE(C && evals && evecs && status, "oiva_eigh: null pointer");
    OIVA_REQUIRE(n_rows > 0 && rows_per_mixture > 0, "oiva_eigh: bad shape");
    const long long R = n_rows;
    cudaStream_t st = (cudaStream_t)stream;
    OIVA_DISPATCH_M(n_chan, {
        k_eigh<M_><<<(unsigned)((R + SOLVE_WARPS - 1) / SOLVE_WARPS), SOLVE_WARPS * 32, 0, st>>>(
            (const cplx*)C, evals, (cplx*)evecs, status, R, rows_per_mixture, lapack_phase);
    });
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}
