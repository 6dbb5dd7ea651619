// OGIVE (orthogonally constrained gradient IVE, K = 1) per-bin kernels: ive.py:93-161 and :216-241.
// The heavy statistic x_psi = (sum_t x conj(y)/r) / (sum_t |y|^2/r) equals V w / (w^H V w) with the
// weighted covariance V of the shared streaming kernel (cov.cuh, K = 1), so only M x M work is left
// per bin and iteration; one thread per bin with runtime M (<= 16) is plenty here.
#include "common.cuh"

namespace oiva {

__device__ __forceinline__ void matvec(cplx* out, const cplx* __restrict__ A, const cplx* x, int M) {
    for (int i = 0; i < M; ++i) {
        cplx acc = cmake(0.0, 0.0);
        for (int j = 0; j < M; ++j) cfma(acc, A[i * M + j], x[j]);
        out[i] = acc;
    }
}
__device__ __forceinline__ cplx vdot(const cplx* a, const cplx* b, int M) {  // a^H b
    cplx acc = cmake(0.0, 0.0);
    for (int i = 0; i < M; ++i) cfmac(acc, a[i], b[i]);
    return acc;
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// Cinv = C^-1 by Gauss-Jordan with partial pivoting (ive.py:98), cnorm = ||C||_F (ive.py:99)
__global__ void k_ogive_setup(const cplx* __restrict__ C, cplx* __restrict__ Cinv, double* __restrict__ cnorm,
                              int* status, long long R, int M) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R) return;
    cplx a[OIVA_MAX_M * OIVA_MAX_M], inv[OIVA_MAX_M * OIVA_MAX_M];
    double fro = 0.0;
    for (int i = 0; i < M * M; ++i) {
        a[i] = C[(size_t)row * M * M + i];
        fro += a[i].x * a[i].x + a[i].y * a[i].y;
        inv[i] = cmake((i / M == i % M) ? 1.0 : 0.0, 0.0);
    }
    cnorm[row] = sqrt(fro);
    bool singular = false;
    for (int c = 0; c < M; ++c) {
        int piv = c;
        double best = -1.0;
        for (int r = c; r < M; ++r) {
            double m = fabs(a[r * M + c].x) + fabs(a[r * M + c].y);
            if (m > best) {
                best = m;
                piv = r;
            }
        }
        if (!(best > 0.0)) singular = true;
        if (piv != c) {
            for (int j = 0; j < M; ++j) {
                cplx t = a[c * M + j];
                a[c * M + j] = a[piv * M + j];
                a[piv * M + j] = t;
                t = inv[c * M + j];
                inv[c * M + j] = inv[piv * M + j];
                inv[piv * M + j] = t;
            }
        }
        const cplx rinv = crecip(a[c * M + c]);
        for (int j = 0; j < M; ++j) {
            a[c * M + j] = cmul(a[c * M + j], rinv);
            inv[c * M + j] = cmul(inv[c * M + j], rinv);
        }
        for (int r = 0; r < M; ++r) {
            if (r == c) continue;
            const cplx f = a[r * M + c];
            for (int j = 0; j < M; ++j) {
                cfms(a[r * M + j], f, a[c * M + j]);
                cfms(inv[r * M + j], f, inv[c * M + j]);
            }
        }
    }
    for (int i = 0; i < M * M; ++i) Cinv[(size_t)row * M * M + i] = inv[i];
    if (singular) atomicOr(status, OIVA_STATUS_SINGULAR);
}

// a = C w / Re(w^H C w)   (ive.py:132-135), all rows
__global__ void k_ogive_a_from_w(const cplx* __restrict__ w, cplx* __restrict__ a, const cplx* __restrict__ C,
                                 long long R, int M) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R) return;
    cplx wv[OIVA_MAX_M], v[OIVA_MAX_M];
    for (int i = 0; i < M; ++i) wv[i] = w[(size_t)row * M + i];
    matvec(v, C + (size_t)row * M * M, wv, M);
    const double lam = 1.0 / vdot(wv, v, M).x;
    for (int i = 0; i < M; ++i) a[(size_t)row * M + i] = cscale(v[i], lam);
}

// switching criterion (ive.py:142-161): do_a = kappa >= 0.1
__global__ void k_ogive_switching(const cplx* __restrict__ a, const cplx* __restrict__ C,
                                  const double* __restrict__ cnorm, uint8_t* __restrict__ do_a, long long R, int M) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R) return;
    const cplx* Cr = C + (size_t)row * M * M;
    cplx an[OIVA_MAX_M], bn[OIVA_MAX_M];
    const cplx a0inv = crecip(a[(size_t)row * M]);
    for (int i = 0; i < M; ++i) an[i] = cmul(a[(size_t)row * M + i], a0inv);
    matvec(bn, Cr, an, M);
    const cplx lmb = bn[0];
    const cplx linv = crecip(lmb);
    double d1 = 0.0, nb2 = 0.0;
    for (int i = 0; i < M; ++i) {
        bn[i] = cmul(bn[i], linv);
        const cplx d = csub(an[i], bn[i]);
        d1 += d.x * d.x + d.y * d.y;
        nb2 += bn[i].x * bn[i].x + bn[i].y * bn[i].y;
    }
    const double p1 = sqrt(d1) / cnorm[row];
    double d2 = 0.0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
            // Cbb = lmb * b_i conj(b_j) / ||b||^2
            cplx bb = cmul(bn[i], cconj(bn[j]));
            cplx cbb = cscale(cmul(lmb, bb), 1.0 / nb2);
            const cplx d = csub(Cr[i * M + j], cbb);
            d2 += d.x * d.x + d.y * d.y;
        }
    const double kappa = p1 * sqrt(d2) / sqrt((double)M);
    do_a[row] = kappa >= 0.1 ? 1 : 0;
}

// one OGIVE iteration for a bin (ive.py:216-241), V = weighted covariance of the current extraction
// gate (may be NULL): max_f ||delta_f|| of the PREVIOUS epoch.  Once it is below tol the reference has left its loop
// (ive.py:238-241), so the update becomes a no-op that only carries the value forward: the host can then check for
// convergence every few epochs instead of synchronising after each one, and still ends in the state of the epoch at
// which the reference stops.
__global__ void k_ogive_update(cplx* __restrict__ w, cplx* __restrict__ a, double* __restrict__ lambda_a,
                               const cplx* __restrict__ V, const cplx* __restrict__ C, const cplx* __restrict__ Cinv,
                               const uint8_t* __restrict__ do_a, double step, double* delta_max, const double* gate,
                               double tol, long long R, int M) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gate != nullptr && *gate < tol) {
        if (row == 0) *delta_max = *gate;
        return;
    }
    if (row >= R) return;
    cplx wv[OIVA_MAX_M], av[OIVA_MAX_M], xpsi[OIVA_MAX_M], tmp[OIVA_MAX_M];
    for (int i = 0; i < M; ++i) {
        wv[i] = w[(size_t)row * M + i];
        av[i] = a[(size_t)row * M + i];
    }
    matvec(xpsi, V + (size_t)row * M * M, wv, M);
    const cplx zinv = crecip(vdot(wv, xpsi, M));
    for (int i = 0; i < M; ++i) xpsi[i] = cmul(xpsi[i], zinv);
    const bool astep = do_a[row] != 0;
    double dn = 0.0;
    if (!astep) {  // w-step then a <- C w / (w^H C w)
        for (int i = 0; i < M; ++i) {
            const cplx d = csub(av[i], xpsi[i]);
            dn += d.x * d.x + d.y * d.y;
            wv[i] = cadd(wv[i], cscale(d, step));
        }
        matvec(tmp, C + (size_t)row * M * M, wv, M);
        const double lam = 1.0 / vdot(wv, tmp, M).x;
        for (int i = 0; i < M; ++i) av[i] = cscale(tmp[i], lam);
    } else {  // a-step
        matvec(tmp, Cinv + (size_t)row * M * M, xpsi, M);
        const double la = lambda_a[row];
        for (int i = 0; i < M; ++i) {
            const cplx d = csub(wv[i], cscale(tmp[i], la));
            dn += d.x * d.x + d.y * d.y;
            av[i] = cadd(av[i], cscale(d, step));
        }
    }
    // lambda_a is refreshed for every bin (ive.py:139), w only for a-step bins (ive.py:140)
    matvec(tmp, Cinv + (size_t)row * M * M, av, M);
    const double la = 1.0 / vdot(av, tmp, M).x;
    lambda_a[row] = la;
    if (astep)
        for (int i = 0; i < M; ++i) wv[i] = cscale(tmp[i], la);
    for (int i = 0; i < M; ++i) {
        w[(size_t)row * M + i] = wv[i];
        a[(size_t)row * M + i] = av[i];
    }
    atomic_max_nonneg(delta_max, sqrt(dn));
}

}  // namespace oiva

using namespace oiva;

extern "C" int oiva_ogive_setup(const void* C, void* Cinv, double* cnorm, int* status, int n_rows, int n_chan,
                                void* stream) {
    OIVA_REQUIRE(C && Cinv && cnorm && status && n_rows > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M,
                 "oiva_ogive_setup: bad arguments");
    k_ogive_setup<<<(n_rows + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const cplx*)C, (cplx*)Cinv, cnorm, status,
                                                                        n_rows, n_chan);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ogive_a_from_w(const void* w, void* a, const void* C, int n_rows, int n_chan, void* stream) {
    OIVA_REQUIRE(w && a && C && n_rows > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M, "oiva_ogive_a_from_w: bad arguments");
    k_ogive_a_from_w<<<(n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const cplx*)w, (cplx*)a, (const cplx*)C,
                                                                              n_rows, n_chan);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ogive_switching(const void* a, const void* C, const double* cnorm, uint8_t* do_a, int n_rows,
                                    int n_chan, void* stream) {
    OIVA_REQUIRE(a && C && cnorm && do_a && n_rows > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M,
                 "oiva_ogive_switching: bad arguments");
    k_ogive_switching<<<(n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const cplx*)a, (const cplx*)C, cnorm,
                                                                               do_a, n_rows, n_chan);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ogive_update(void* w, void* a, double* lambda_a, const void* V, const void* C, const void* Cinv,
                                 const uint8_t* do_a, double step_size, double* delta_max, int n_rows, int n_chan,
                                 void* stream) {
    OIVA_REQUIRE(w && a && lambda_a && V && C && Cinv && do_a && delta_max && n_rows > 0 && n_chan >= 1 &&
                     n_chan <= OIVA_MAX_M,
                 "oiva_ogive_update: bad arguments");
    k_ogive_update<<<(n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        (cplx*)w, (cplx*)a, lambda_a, (const cplx*)V, (const cplx*)C, (const cplx*)Cinv, do_a, step_size, delta_max,
        nullptr, 0.0, n_rows, n_chan);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ogive_update_gated(void* w, void* a, double* lambda_a, const void* V, const void* C,
                                       const void* Cinv, const uint8_t* do_a, double step_size, double* delta_hist,
                                       int epoch, double tol, int n_rows, int n_chan, void* stream) {
    OIVA_REQUIRE(w && a && lambda_a && V && C && Cinv && do_a && delta_hist && epoch >= 0 && n_rows > 0 &&
                     n_chan >= 1 && n_chan <= OIVA_MAX_M,
                 "oiva_ogive_update_gated: bad arguments");
    k_ogive_update<<<(n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        (cplx*)w, (cplx*)a, lambda_a, (const cplx*)V, (const cplx*)C, (const cplx*)Cinv, do_a, step_size,
        delta_hist + epoch, epoch > 0 ? delta_hist + epoch - 1 : nullptr, tol, n_rows, n_chan);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// n_epochs epochs of the OGIVE loop in one library call (ive.py:191-241 without the callback / switching steps, which
// the caller places between blocks): statistic -> source model -> weighted covariance (K = 1) -> full matrices ->
// gated update.  The Python loop paid five ctypes round trips per epoch, of the order of the GPU work for one mixture.
// Xg: grouped samples; w (F, M) c128 row-major filters; r2part (NG, Tp), phi (Tp), Vg, cov scratch, V (F, M, M): work
// arrays as for the individual calls; epoch0: index of the first epoch (row of delta_hist).
extern "C" int oiva_ogive_iterate(const void* Xg, void* w, void* a, double* lambda_a, double* r2part, double* phi, void* Vg,
                                  void* cov_scratch, size_t cov_scratch_bytes, void* V, const void* C, const void* Cinv,
                                  const uint8_t* do_a, double step_size, double* delta_hist, int epoch0, int n_epochs,
                                  double tol, int n_frames, int n_freq, int n_chan, int model, int dtype, void* stream) {
    OIVA_REQUIRE(Xg && w && a && lambda_a && r2part && phi && Vg && V && C && Cinv && do_a && delta_hist,
                 "oiva_ogive_iterate: null pointer");
    const int NG = oiva_bin_groups(n_freq);
    for (int e = 0; e < n_epochs; ++e) {
        int rc = oiva_demix_power(Xg, w, 1, 0, r2part, 1, n_frames, n_freq, n_chan, 1, dtype, stream);
        if (rc) return rc;
        rc = oiva_source_model(r2part, NG, phi, nullptr, 1, n_frames, 1, n_freq, model, stream);
        if (rc) return rc;
        rc = oiva_weighted_cov_ws(Xg, phi, Vg, cov_scratch, cov_scratch_bytes, 1, n_frames, n_freq, n_chan, 1, dtype, stream);
        if (rc) return rc;
        rc = oiva_unpack_cov(Vg, V, 1, n_freq, n_chan, 1, stream);
        if (rc) return rc;
        rc = oiva_ogive_update_gated(w, a, lambda_a, V, C, Cinv, do_a, step_size, delta_hist, epoch0 + e, tol, n_freq, n_chan,
                                     stream);
        if (rc) return rc;
    }
    return OIVA_OK;
}
