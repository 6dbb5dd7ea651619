// Thread-per-bin IP sweep (M <= 6): the lane <-> bin mapping of the streaming kernels carried into the
// per-bin linear algebra.  One thread owns one frequency bin and does the whole sweep for it in registers:
// no shuffles, no shared memory, no divergence (pivoting is done with selects), every lane busy, and the
// grouped covariances Vg / Cg are read as 512-byte coalesced warp loads.
//
// Per source s (reference: overiva.py:176-190):
//   A = W_hat^H V_s, built row by row from the structure W_hat^H = [W^H ; (J^H, -I)]  (K*M^2 + (M-K)*K*M complex
//       FMAs instead of M^3);  LU with partial pivoting (LAPACK izamax rule: max |re|+|im|, first wins) on
//       [A | e_s];  back substitution;  w_s /= sqrt(w_s^H V_s w_s);  J = (W^H C E1)^-1 (W^H C E2).
#pragma once
#include "cov.cuh"  // static_for
#include "common.cuh"

namespace oiva {

// Hermitian matrix stored as its lower triangle in the grouped layout: element (i, j)
__device__ __forceinline__ cplx herm_load(const cplx* __restrict__ base, int i, int j) {
    const int hi = i >= j ? i : j, lo = i >= j ? j : i;
    cplx v = ld_nc_c(base + (size_t)(hi * (hi + 1) / 2 + lo) * OIVA_GROUP);
    if (i < j) v.y = -v.y;
    return v;
}

__device__ __forceinline__ void cswap_if(bool c, cplx& a, cplx& b) {
    const cplx ta = a, tb = b;
    a.x = c ? tb.x : ta.x;
    a.y = c ? tb.y : ta.y;
    b.x = c ? ta.x : tb.x;
    b.y = c ? ta.y : tb.y;
}

// In-register LU with partial pivoting of the N x N system [A | rhs] (rhs: NR columns), then back
// substitution: on return rhs holds A^-1 rhs.  `singular` is set on a zero / NaN pivot.
template <int N, int NR>
__device__ __forceinline__ void lu_solve(cplx (&A)[N][N], cplx (&rhs)[N][NR], bool& singular) {
    static_for<N>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        // pivot search in column c, rows c..N-1
        double best = fabs(A[c][c].x) + fabs(A[c][c].y);
        int p = c;
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            const double m = fabs(A[r][c].x) + fabs(A[r][c].y);
            if (m > best) {
                best = m;
                p = r;
            }
        }
        if (!(best > 0.0)) singular = true;
        // bring the pivot row to position c (select-based conditional swaps: no divergence)
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            const bool sw = (p == r);
#pragma unroll
            for (int c2 = c; c2 < N; ++c2) cswap_if(sw, A[c][c2], A[r][c2]);
#pragma unroll
            for (int q = 0; q < NR; ++q) cswap_if(sw, rhs[c][q], rhs[r][q]);
        }
        const cplx rinv = crecip(A[c][c]);
        // normalise the pivot row (U gets a unit diagonal), eliminate below
#pragma unroll
        for (int c2 = c + 1; c2 < N; ++c2) A[c][c2] = cmul(A[c][c2], rinv);
#pragma unroll
        for (int q = 0; q < NR; ++q) rhs[c][q] = cmul(rhs[c][q], rinv);
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            const cplx f = A[r][c];
#pragma unroll
            for (int c2 = c + 1; c2 < N; ++c2) cfms(A[r][c2], f, A[c][c2]);
#pragma unroll
            for (int q = 0; q < NR; ++q) cfms(rhs[r][q], f, rhs[c][q]);
        }
    });
    // back substitution with the unit-diagonal U
    static_for<N>([&](auto cc) {
        constexpr int c = N - 1 - decltype(cc)::value;
#pragma unroll
        for (int r = 0; r < c; ++r) {
#pragma unroll
            for (int q = 0; q < NR; ++q) cfms(rhs[r][q], A[r][c], rhs[c][q]);
        }
    });
}

// OverIVA background refresh for one bin held by one thread: J = (W^H C E1)^-1 (W^H C E2)      overiva.py:96-98
template <int M, int K>
__device__ __forceinline__ void background_tpb(cplx* __restrict__ Wm, const cplx* __restrict__ Cb, bool& singular) {
    if constexpr (K < M) {
        cplx T1[K][K], T2[K][M - K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int c = 0; c < K; ++c) T1[i][c] = cmake(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < M - K; ++c) T2[i][c] = cmake(0.0, 0.0);
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
            cplx crow[M];
#pragma unroll
            for (int c = 0; c < M; ++c) crow[c] = herm_load(Cb, j, c);
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const cplx a = Wm[j * M + i];
#pragma unroll
                for (int c = 0; c < K; ++c) cfmac(T1[i][c], a, crow[c]);
#pragma unroll
                for (int c = 0; c < M - K; ++c) cfmac(T2[i][c], a, crow[K + c]);
            }
        }
        lu_solve<K, M - K>(T1, T2, singular);
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int c = 0; c < M - K; ++c) Wm[r * M + K + c] = T2[r][c];
    }
}

// grid: ceil(G*32 / 128) CTAs of 128 threads; thread <-> (group gi, lane = bin % 32)
template <int M, int K>
__global__ void __launch_bounds__(128) k_ip_update_tpb(cplx* __restrict__ What, const cplx* __restrict__ Vg,
                                                       const cplx* __restrict__ Cg, const double* __restrict__ wscale,
                                                       int* status, int F, int NG, long long G) {
    constexpr int NE = oiva_tri(M);
    const long long gi = ((long long)blockIdx.x * 128 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gi >= G) return;
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + lane;
    if (f >= F) return;
    cplx* Wm = What + ((size_t)b * F + f) * M * M;  // this bin's W_hat, row-major (read back after writes: no __ldg)
    const cplx* Vb = Vg + (size_t)gi * K * NE * OIVA_GROUP + lane;
    const cplx* Cb = Cg + (size_t)gi * NE * OIVA_GROUP + lane;
    bool singular = false;

    if (wscale) {  // W /= gamma (laplace) or sqrt(gamma) (gauss)                       overiva.py:161-167
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double sc = wscale[b * K + k];
#pragma unroll
            for (int j = 0; j < M; ++j) Wm[j * M + k] = cscale(Wm[j * M + k], sc);
        }
    }

#pragma unroll 1
    for (int s = 0; s < K; ++s) {
        const cplx* Vs = Vb + (size_t)s * NE * OIVA_GROUP;
        // A = W_hat^H V_s, rhs = e_s
        cplx A[M][M], rhs[M][1];
#pragma unroll
        for (int i = 0; i < M; ++i) {
#pragma unroll
            for (int c = 0; c < M; ++c) A[i][c] = cmake(0.0, 0.0);
            rhs[i][0] = cmake(i == s ? 1.0 : 0.0, 0.0);
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
            cplx vrow[M];
#pragma unroll
            for (int c = 0; c < M; ++c) vrow[c] = herm_load(Vs, j, c);
            if (j < K) {
                // rows i < K: conj(W[j][i]);  rows i >= K: conj(J[j][i-K]) -- both are W_hat[j][i]
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const cplx a = Wm[j * M + i];
#pragma unroll
                    for (int c = 0; c < M; ++c) cfmac(A[i][c], a, vrow[c]);
                }
            } else {
                // W_hat[j][i] for j >= K: W[j][i] for i < K, -delta(i, j) otherwise
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const cplx a = Wm[j * M + i];
#pragma unroll
                    for (int c = 0; c < M; ++c) cfmac(A[i][c], a, vrow[c]);
                }
#pragma unroll
                for (int c = 0; c < M; ++c) A[j][c] = csub(A[j][c], vrow[c]);
            }
        }
        lu_solve<M, 1>(A, rhs, singular);
        // normalise: w /= sqrt(w^H V_s w)                                               overiva.py:185-186
        cplx d = cmake(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            cplx u = cmake(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < M; ++j) cfma(u, herm_load(Vs, i, j), rhs[j][0]);
            cfmac(d, rhs[i][0], u);
        }
        const cplx inv = crecip(csqrt_(d));
#pragma unroll
        for (int i = 0; i < M; ++i) Wm[i * M + s] = cmul(rhs[i][0], inv);
        background_tpb<M, K>(Wm, Cb, singular);
    }

    bool bad = false;
#pragma unroll
    for (int i = 0; i < M * M; ++i) {
        const cplx v = Wm[i];
        if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
    }
    if (singular || bad) atomicOr(status, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
}

}  // namespace oiva
