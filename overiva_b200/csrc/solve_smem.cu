// Thread-per-bin IP sweep with the per-bin matrices in SHARED memory -- an experiment for the shapes whose matrices do
// not fit one thread's registers (solve_tpb.cuh stops at M = 6, and at M = 7, 8 with K <= 4): M = 7..16 with any K, in
// particular the determined case K = M of AuxIVA (overiva_oneshot.py:301-309, overiva_sim.py:295-298).  Correct, but
// measured slower than the row-owner kernel: opt-in only (see ip_update_smem below).
//
// One warp owns a group of 32 bins, lane <-> bin as everywhere else; every per-bin array lives in the warp's shared
// memory as [entry][lane], so an access is 32 consecutive complex numbers (conflict-free LDS.128 / STS.128) and a
// PER-LANE row index -- partial pivoting picks a different row in every bin -- costs nothing: the bank is decided by
// the lane, not by the row.  All loops have run-time bounds (one code path for every (M, K)), the inner loops are
// blocked by 4 so that independent shared-memory loads are in flight together.
//
// Per source s (reference: overiva.py:176-190), with the identity (W_hat^H V)^-1 e_s = V^-1 (W_hat^-H e_s) that the
// register version already uses:
//   q = (W_hat^H)^-1 e_s      Gaussian elimination with partial pivoting (izamax rule) on the M x M system
//   V = L L^H                 in-place Cholesky on the lower triangle;  y = L^-1 q;  w^H V w = |y|^2;  w = L^-H y
//   w_s = w / sqrt(|y|^2)                                                                      overiva.py:185-186
//   J = (W^H C E1)^-1 (W^H C E2)   K x K pivoted elimination with M-K right-hand sides         overiva.py:96-98
#include <stdlib.h>

#include "common.cuh"

namespace oiva {

namespace {

struct LaneMat {  // a per-lane array in the [entry][lane] layout
    cplx* p;
    __device__ __forceinline__ cplx& operator[](int i) const { return p[i * OIVA_GROUP]; }
};

// Solve A x = b (n x n, row-major in `A`, right-hand sides: nrhs columns of `B`, row-major n x nrhs) in place by
// Gaussian elimination with partial pivoting; on return B holds the solution.  Every lane works on its own system.
__device__ __forceinline__ void lane_solve(LaneMat A, LaneMat B, int n, int nrhs, bool& singular) {
    for (int c = 0; c < n; ++c) {
        // pivot: largest |re| + |im| in column c among rows c..n-1, first one wins (LAPACK izamax)
        cplx v = A[c * n + c];
        double best = fabs(v.x) + fabs(v.y);
        int piv = c;
        for (int r = c + 1; r < n; ++r) {
            v = A[r * n + c];
            const double m = fabs(v.x) + fabs(v.y);
            if (m > best) {
                best = m;
                piv = r;
            }
        }
        if (!(best > 0.0)) singular = true;
        if (piv != c) {  // per-lane row swap (lanes that do not swap skip it: divergence inside a short loop only)
            for (int k = c; k < n; ++k) {
                const cplx t = A[c * n + k];
                A[c * n + k] = A[piv * n + k];
                A[piv * n + k] = t;
            }
            for (int k = 0; k < nrhs; ++k) {
                const cplx t = B[c * nrhs + k];
                B[c * nrhs + k] = B[piv * nrhs + k];
                B[piv * nrhs + k] = t;
            }
        }
        const cplx rinv = crecip(A[c * n + c]);
        for (int k = c + 1; k < n; ++k) A[c * n + k] = cmul(A[c * n + k], rinv);
        for (int k = 0; k < nrhs; ++k) B[c * nrhs + k] = cmul(B[c * nrhs + k], rinv);
        for (int r = c + 1; r < n; ++r) {
            const cplx f = A[r * n + c];
            int k = c + 1;
            for (; k + 4 <= n; k += 4) {  // 4 independent updates in flight
                cplx a0 = A[r * n + k], a1 = A[r * n + k + 1], a2 = A[r * n + k + 2], a3 = A[r * n + k + 3];
                const cplx p0 = A[c * n + k], p1 = A[c * n + k + 1], p2 = A[c * n + k + 2], p3 = A[c * n + k + 3];
                cfms(a0, f, p0);
                cfms(a1, f, p1);
                cfms(a2, f, p2);
                cfms(a3, f, p3);
                A[r * n + k] = a0;
                A[r * n + k + 1] = a1;
                A[r * n + k + 2] = a2;
                A[r * n + k + 3] = a3;
            }
            for (; k < n; ++k) {
                cplx a = A[r * n + k];
                cfms(a, f, A[c * n + k]);
                A[r * n + k] = a;
            }
            for (k = 0; k < nrhs; ++k) {
                cplx b = B[r * nrhs + k];
                cfms(b, f, B[c * nrhs + k]);
                B[r * nrhs + k] = b;
            }
        }
    }
    // back substitution with the unit-diagonal U
    for (int c = n - 1; c > 0; --c)
        for (int r = 0; r < c; ++r) {
            const cplx u = A[r * n + c];
            for (int k = 0; k < nrhs; ++k) {
                cplx b = B[r * nrhs + k];
                cfms(b, u, B[c * nrhs + k]);
                B[r * nrhs + k] = b;
            }
        }
}

// acc -= a * conj(b)
__device__ __forceinline__ void cfms_bconj(cplx& acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.y, b.x, acc.y);
    acc.y = fma(a.x, b.y, acc.y);
}

// grid: one warp per CTA, CTA <-> group gi.  Dynamic shared memory: (M*M + M + tri(M) + M) * 32 complex.
__global__ void __launch_bounds__(32) k_ip_update_smem(cplx* __restrict__ Wg, const cplx* __restrict__ Vg,
                                                       const cplx* __restrict__ Cg, const double* __restrict__ wscale,
                                                       int* status, int F, int NG, int M, int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const long long gi = blockIdx.x;
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + lane;
    const bool valid = f < F;
    const int NE = oiva_tri(M);
    cplx* sm = reinterpret_cast<cplx*>(smem_raw) + lane;
    const LaneMat A = {sm};                                          // M x M  (later K x K for the J system)
    const LaneMat Bv = {sm + (size_t)M * M * OIVA_GROUP};            // M      (later K x (M-K), inside A's tail)
    const LaneMat Lm = {sm + (size_t)(M * M + M) * OIVA_GROUP};      // tri(M): V_s, then its Cholesky factor
    const LaneMat q = {sm + (size_t)(M * M + M + NE) * OIVA_GROUP};  // M
    const LaneMat Wm = {Wg + (size_t)gi * M * M * OIVA_GROUP + lane};
    const cplx* Cgrp = Cg + (size_t)gi * NE * OIVA_GROUP + lane;
    bool singular = false;
    if (valid) {
        if (wscale) {  // W /= gamma (laplace) or sqrt(gamma) (gauss)                          overiva.py:161-167
            for (int k = 0; k < K; ++k) {
                const double sc = wscale[b * K + k];
                for (int j = 0; j < M; ++j) Wm[j * M + k] = cscale(Wm[j * M + k], sc);
            }
        }
        for (int s = 0; s < K; ++s) {
            // q = (W_hat^H)^-1 e_s
            for (int i = 0; i < M; ++i) {
                for (int c = 0; c < M; ++c) A[i * M + c] = cconj(Wm[c * M + i]);
                Bv[i] = cmake(i == s ? 1.0 : 0.0, 0.0);
            }
            lane_solve(A, Bv, M, 1, singular);
            for (int i = 0; i < M; ++i) q[i] = Bv[i];
            // Cholesky of V_s (lower triangle, e = i(i+1)/2 + j)
            const cplx* Vs = Vg + ((size_t)gi * K + s) * NE * OIVA_GROUP + lane;
            for (int e = 0; e < NE; ++e) Lm[e] = ld_nc_c(Vs + (size_t)e * OIVA_GROUP);
            for (int j = 0; j < M; ++j) {
                const int jj = j * (j + 1) / 2;
                double djj = Lm[jj + j].x;
                for (int k = 0; k < j; ++k) {
                    const cplx l = Lm[jj + k];
                    djj = fma(-l.x, l.x, fma(-l.y, l.y, djj));
                }
                if (!(djj > 0.0)) singular = true;
                const double dinv = 1.0 / sqrt(djj);
                Lm[jj + j] = cmake(dinv, 0.0);  // the diagonal slot keeps 1 / L_jj
                for (int i = j + 1; i < M; ++i) {
                    const int ii = i * (i + 1) / 2;
                    cplx v = Lm[ii + j];
                    for (int k = 0; k < j; ++k) cfms_bconj(v, Lm[ii + k], Lm[jj + k]);  // v -= L[i][k] conj(L[j][k])
                    Lm[ii + j] = cscale(v, dinv);
                }
            }
            // forward: L y = q;  |y|^2 = w^H V w
            double den = 0.0;
            for (int i = 0; i < M; ++i) {
                const int ii = i * (i + 1) / 2;
                cplx v = q[i];
                for (int k = 0; k < i; ++k) cfms(v, Lm[ii + k], q[k]);
                v = cscale(v, Lm[ii + i].x);
                q[i] = v;
                den = fma(v.x, v.x, fma(v.y, v.y, den));
            }
            // backward: L^H w = y
            for (int i = M - 1; i >= 0; --i) {
                cplx v = q[i];
                for (int k = i + 1; k < M; ++k) {
                    const cplx l = Lm[k * (k + 1) / 2 + i];  // L^H[i][k] = conj(L[k][i])
                    const cplx wk = q[k];
                    v.x = fma(-l.x, wk.x, fma(-l.y, wk.y, v.x));
                    v.y = fma(-l.x, wk.y, fma(l.y, wk.x, v.y));
                }
                q[i] = cscale(v, Lm[i * (i + 1) / 2 + i].x);
            }
            const double inv = 1.0 / sqrt(den);
            for (int i = 0; i < M; ++i) Wm[i * M + s] = cscale(q[i], inv);
            // J = (W^H C E1)^-1 (W^H C E2)                                                     overiva.py:189-190
            if (K < M) {
                const int R = M - K;
                const LaneMat T1 = A;                                       // K x K
                const LaneMat T2 = {A.p + (size_t)K * K * OIVA_GROUP};      // K x R  (K*K + K*R = K*M <= M*M)
                for (int i = 0; i < K; ++i) {
                    for (int c = 0; c < M; ++c) {
                        cplx acc = cmake(0.0, 0.0);
                        for (int j = 0; j < M; ++j) {
                            const int hi = j >= c ? j : c, lo = j >= c ? c : j;
                            cplx cv = ld_nc_c(Cgrp + (size_t)(hi * (hi + 1) / 2 + lo) * OIVA_GROUP);
                            if (j < c) cv.y = -cv.y;  // C[j][c] from the stored lower triangle
                            cfmac(acc, Wm[j * M + i], cv);
                        }
                        if (c < K) T1[i * K + c] = acc;
                        else T2[i * R + (c - K)] = acc;
                    }
                }
                lane_solve(T1, T2, K, R, singular);
                for (int r = 0; r < K; ++r)
                    for (int c = 0; c < R; ++c) Wm[r * M + K + c] = T2[r * R + c];
            }
        }
        bool bad = false;
        for (int j = 0; j < M; ++j)
            for (int c = 0; c < M; ++c)
                if (c < K || j < K) {
                    const cplx v = Wm[j * M + c];
                    if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
                }
        if (singular || bad)
            atomicOr(status, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
    }
}

}  // namespace

// returns OIVA_ERR_INVALID (without setting an error) when disabled
int ip_update_smem(int M, int K, cplx* Wg, const cplx* Vg, const cplx* Cg, const double* wscale, int* status, int F,
                   int NG, long long G, cudaStream_t st) {
    // OPT-IN (OIVA_SOLVER_SMEM=1): measured on B200 it is 2-9x SLOWER than the row-owner kernel of solve.cu
    // (256 mixtures, M = K = 8: 10.3 vs 5.4 ms per sweep; M = K = 7: 5.7 vs 2.6; M = 8, K = 5: 14.5 vs 5.1; config 5,
    // M = 16, K = 4: 1.07 vs 0.13 ms) -- one warp per CTA walking run-time loops over shared memory is a long
    // dependent chain with nothing to overlap it, while the row-owner kernel spreads a bin over 8-16 lanes.  Kept as
    // a tested alternative; read per call so that tests can switch it on.
    const char* v = getenv("OIVA_SOLVER_SMEM");
    const bool enabled = v && *v && *v != '0';
    if (!enabled || G > 0x7fffffffll) return OIVA_ERR_INVALID;
    const size_t smem = (size_t)(M * M + M + oiva_tri(M) + M) * OIVA_GROUP * sizeof(cplx);
    if (smem > 220 * 1024) return OIVA_ERR_INVALID;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(k_ip_update_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    k_ip_update_smem<<<(unsigned)G, 32, smem, st>>>(Wg, Vg, Cg, wscale, status, F, NG, M, K);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

}  // namespace oiva
