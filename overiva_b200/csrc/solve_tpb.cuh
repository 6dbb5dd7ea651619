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

// covariance loads: NC = true reads global memory through the read-only path (the batched sweep kernel); NC = false is
// a plain load that also works on shared memory (the persistent single-mixture loop keeps C and V there)
template <bool NC>
__device__ __forceinline__ cplx cov_ld(const cplx* p) {
    if constexpr (NC) return ld_nc_c(p);
    else return *p;
}
// Hermitian matrix stored as its lower triangle in the grouped layout: element (i, j)
template <bool NC = true>
__device__ __forceinline__ cplx herm_load(const cplx* __restrict__ base, int i, int j) {
    const int hi = i >= j ? i : j, lo = i >= j ? j : i;
    cplx v = cov_ld<NC>(base + (size_t)(hi * (hi + 1) / 2 + lo) * OIVA_GROUP);
    if (i < j) v.y = -v.y;
    return v;
}

// this lane's W_hat inside the grouped array Wg[gi][M*M][32]: element idx of the row-major M x M matrix
struct WLane {
    cplx* base;
    __device__ __forceinline__ cplx& operator[](int idx) const { return base[idx * OIVA_GROUP]; }
};

// acc -= conj(a) * b
__device__ __forceinline__ void cfms_conj(cplx& acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
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
        const cplx rinv = crecip_fast(A[c][c]);
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
template <int M, int K, bool NC = true>
__device__ __forceinline__ void background_tpb(WLane Wm, const cplx* sC, bool& singular) {
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
            for (int c = 0; c < M; ++c) crow[c] = herm_load<NC>(sC, j, c);
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

// ---- one IP update, determined case (K == M): w_s = (W^H V_s)^-1 e_s by LU with partial pivoting ----------
// V_s is read through a getter vget(i, j) (full Hermitian element): the grouped array in global / shared memory, or a
// register copy of the lower triangle (the fused covariance + sweep kernel, cov_sweep.cuh).
template <bool NC>
struct HermFromMemory {
    const cplx* base;
    __device__ __forceinline__ cplx operator()(int i, int j) const { return herm_load<NC>(base, i, j); }
};
template <int M>
struct HermFromRegs {
    const cplx (&tri)[oiva_tri(M)];
    __device__ __forceinline__ cplx operator()(int i, int j) const {
        const int hi = i >= j ? i : j, lo = i >= j ? j : i;
        cplx v = tri[hi * (hi + 1) / 2 + lo];
        if (i < j) v.y = -v.y;
        return v;
    }
};
template <int M, typename VGet>
__device__ __forceinline__ void ip_source_full_v(WLane Wm, const VGet& vget, int s, bool& singular) {
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
        for (int c = 0; c < M; ++c) vrow[c] = vget(j, c);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const cplx a = Wm[j * M + i];
#pragma unroll
            for (int c = 0; c < M; ++c) cfmac(A[i][c], a, vrow[c]);
        }
    }
    lu_solve<M, 1>(A, rhs, singular);
    // normalise: w /= sqrt(w^H V_s w)                                                   overiva.py:185-186
    cplx d = cmake(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        cplx u = cmake(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < M; ++j) cfma(u, vget(i, j), rhs[j][0]);
        cfmac(d, rhs[i][0], u);
    }
    const cplx inv = crsqrt_pos(d);
#pragma unroll
    for (int i = 0; i < M; ++i) Wm[i * M + s] = cmul(rhs[i][0], inv);
}
template <int M, bool NC = true>
__device__ __forceinline__ void ip_source_full(WLane Wm, const cplx* sV, int s, bool& singular) {
    ip_source_full_v<M>(Wm, HermFromMemory<NC>{sV}, s, singular);
}

// ---- Cholesky pieces shared by the overdetermined update and the tracked-inverse determined sweep ----------------
// V = L L^H in place on the packed lower triangle (row-major, e = i(i+1)/2 + j); dinv[j] = 1 / L[j][j].  (static_for: the
// triple loop must be unrolled completely so that Lm is indexed with constants and stays in registers -- with
// "#pragma unroll" alone ptxas kept a 336-byte local-memory copy of it at M = 6)
// (LArr: any array of at least oiva_tri(M) complex numbers indexed with constants)
template <int M, typename LArr>
__device__ __forceinline__ void chol_factor(LArr& Lm, double (&dinv)[M], bool& singular) {
    static_for<M>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        double djj = Lm[j * (j + 1) / 2 + j].x;
        static_for<j>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const cplx l = Lm[j * (j + 1) / 2 + k];
            djj = fma(-l.x, l.x, fma(-l.y, l.y, djj));
        });
        if (!(djj > 0.0)) singular = true;  // not positive definite (or NaN)
        dinv[j] = rsqrt(djj);  // 1 / L[j][j]: one MUFU.RSQ64H + 4 fp64 operations where sqrt and a division are ~3x that,
                               // on the critical path of every pivot (1 ulp, like the two roundings it replaces)
        static_for<M - 1 - j>([&](auto ic) {
            constexpr int i = j + 1 + decltype(ic)::value;
            cplx v = Lm[i * (i + 1) / 2 + j];
            static_for<j>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                // v -= L[i][k] conj(L[j][k])
                const cplx a = Lm[i * (i + 1) / 2 + k], bb = Lm[j * (j + 1) / 2 + k];
                v.x = fma(-a.x, bb.x, fma(-a.y, bb.y, v.x));
                v.y = fma(-a.y, bb.x, fma(a.x, bb.y, v.y));
            });
            Lm[i * (i + 1) / 2 + j] = cscale(v, dinv[j]);
        });
    });
}
// q <- V^-1 q / sqrt(q^H V^-1 q) from the factor: L y = q, |y|^2 (= w^H V w, real: overiva.py:185-186), L^H w = y
template <int M, typename LArr>
__device__ __forceinline__ void chol_solve_normalise(const LArr& Lm, const double (&dinv)[M], cplx (&q)[M]) {
    static_for<M>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        cplx v = q[i];
        static_for<i>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            cfms(v, Lm[i * (i + 1) / 2 + k], q[k]);
        });
        q[i] = cscale(v, dinv[i]);
    });
    // normalisation: w^H V w = w^H L L^H w = |L^H w|^2 = |y|^2 -- available before the back substitution
    double den = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) den = fma(q[i].x, q[i].x, fma(q[i].y, q[i].y, den));
    static_for<M>([&](auto iic) {
        constexpr int i = M - 1 - decltype(iic)::value;
        cplx v = q[i];
        static_for<M - 1 - i>([&](auto kc) {
            constexpr int k = i + 1 + decltype(kc)::value;
            cfms_conj(v, Lm[k * (k + 1) / 2 + i], q[k]);  // L^H[i][k] = conj(L[k][i])
        });
        q[i] = cscale(v, dinv[i]);
    });
    const double inv = rsqrt(den);
#pragma unroll
    for (int i = 0; i < M; ++i) q[i] = cscale(q[i], inv);
}

// A <- A^-1 in place (registers, row-major flat array Af[i M + j]): Gauss-Jordan with partial pivoting (zgetrf's rule: largest |re| + |im| in the column, the
// lower row wins a tie), the row swaps undone on the columns at the end.  For the tracked-inverse determined sweep below.
template <int M>
__device__ __forceinline__ void invert_inplace(cplx (&Af)[M * M], bool& singular) {
    int perm[M];
    static_for<M>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        double best = fabs(Af[(c) * M + (c)].x) + fabs(Af[(c) * M + (c)].y);
        int p = c;
        static_for<M - 1 - c>([&](auto ic) {
            constexpr int i = c + 1 + decltype(ic)::value;
            const double m = fabs(Af[(i) * M + (c)].x) + fabs(Af[(i) * M + (c)].y);
            if (m > best) {
                best = m;
                p = i;
            }
        });
        if (!(best > 0.0)) singular = true;
        perm[c] = p;
        static_for<M - 1 - c>([&](auto ic) {  // rows c <-> p
            constexpr int i = c + 1 + decltype(ic)::value;
            const bool sw = p == i;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const cplx t = Af[(i) * M + (j)], u = Af[(c) * M + (j)];
                Af[(i) * M + (j)].x = sw ? u.x : t.x;
                Af[(i) * M + (j)].y = sw ? u.y : t.y;
                Af[(c) * M + (j)].x = sw ? t.x : u.x;
                Af[(c) * M + (j)].y = sw ? t.y : u.y;
            }
        });
        const cplx piv = crecip_fast(Af[(c) * M + (c)]);
        Af[(c) * M + (c)] = cmake(1.0, 0.0);
#pragma unroll
        for (int j = 0; j < M; ++j) Af[(c) * M + (j)] = cmul(Af[(c) * M + (j)], piv);
        static_for<M>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (i != c) {
                const cplx f = Af[(i) * M + (c)];
                Af[(i) * M + (c)] = cmake(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < M; ++j) cfms(Af[(i) * M + (j)], f, Af[(c) * M + (j)]);
            }
        });
    });
    static_for<M>([&](auto cc) {  // (P W)^-1 = W^-1 P^T: the swaps on the columns, last one first
        constexpr int c = M - 1 - decltype(cc)::value;
        const int p = perm[c];
        static_for<M - 1 - c>([&](auto jc) {
            constexpr int j = c + 1 + decltype(jc)::value;
            const bool sw = p == j;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const cplx t = Af[(i) * M + (j)], u = Af[(i) * M + (c)];
                Af[(i) * M + (j)].x = sw ? u.x : t.x;
                Af[(i) * M + (j)].y = sw ? u.y : t.y;
                Af[(i) * M + (c)].x = sw ? t.x : u.x;
                Af[(i) * M + (c)].y = sw ? t.y : u.y;
            }
        });
    });
}

// ---- one IP update, overdetermined case (K < M) ------------------------------------------------------------
// (W_hat^H V)^-1 e_s = V^-1 q with q = W_hat^-H e_s.  Because W_hat = [W | (J; -I)], q follows from a K x K
// system:  q1 = (W1^H + W2^H J^H)^-1 e_s,  q2 = J^H q1   (W1 / W2: top K / bottom M-K rows of W);  then V w = q is
// solved by Cholesky (V is Hermitian positive definite: no pivoting, no row swaps), and the normalisation needs
// only w^H V w = w^H q.  ~3x fewer operations than forming W_hat^H V and factorising it, same result.
// (Lm: the lower triangle of V_s, row-major packed, e = i(i+1)/2 + j; overwritten by its Cholesky factor)
template <int M, int K>
__device__ __forceinline__ void ip_source_reduced_tri(WLane Wm, cplx (&Lm)[oiva_tri(M)], int s, bool& singular) {
    constexpr int R = M - K;
    cplx q[M];
    {
        cplx Bm[K][K], e[K][1];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            e[i][0] = cmake(i == s ? 1.0 : 0.0, 0.0);
#pragma unroll
            for (int c = 0; c < K; ++c) Bm[i][c] = cconj(Wm[c * M + i]);  // W1^H
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx jc[K];
#pragma unroll
            for (int c = 0; c < K; ++c) jc[c] = cconj(Wm[c * M + K + r]);  // conj(J[c][r])
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const cplx w2 = Wm[(K + r) * M + i];  // W2[r][i]
#pragma unroll
                for (int c = 0; c < K; ++c) cfmac(Bm[i][c], w2, jc[c]);  // += conj(W2[r][i]) conj(J[c][r])
            }
        }
        lu_solve<K, 1>(Bm, e, singular);
#pragma unroll
        for (int i = 0; i < K; ++i) q[i] = e[i][0];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx acc = cmake(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < K; ++c) cfmac(acc, Wm[c * M + K + r], q[c]);  // conj(J[c][r]) q1[c]
            q[K + r] = acc;
        }
    }
    double dinv[M];
    chol_factor<M>(Lm, dinv, singular);
    chol_solve_normalise<M>(Lm, dinv, q);
#pragma unroll
    for (int i = 0; i < M; ++i) Wm[i * M + s] = q[i];
}
template <int M, int K, bool NC = true>
__device__ __forceinline__ void ip_source_reduced(WLane Wm, const cplx* sV, int s, bool& singular) {
    cplx Lm[oiva_tri(M)];
#pragma unroll
    for (int e = 0; e < oiva_tri(M); ++e) Lm[e] = cov_ld<NC>(sV + (size_t)e * OIVA_GROUP);
    ip_source_reduced_tri<M, K>(Wm, Lm, s, singular);
}

constexpr int TPB_WARPS = 4;

// The sweep of one bin by its thread, in steps: W rescale, then per source the IP update + J refresh, then the
// finiteness check.  Cl / Vs point at this lane's element 0 of the grouped lower triangles (Cg[gi][e][lane],
// Vg[gi][s][e][lane]); NC as for cov_ld.  Shared by the batched sweep kernel below and by the persistent
// single-mixture loop (resident.cuh), which runs exactly the same arithmetic.
template <int M, int K>
__device__ __forceinline__ void ip_sweep_rescale(WLane Wm, const double* wscale_b) {
    // W /= gamma (laplace) or sqrt(gamma) (gauss)                                   overiva.py:161-167
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double sc = wscale_b[k];
#pragma unroll
        for (int j = 0; j < M; ++j) Wm[j * M + k] = cscale(Wm[j * M + k], sc);
    }
}
template <int M, int K, bool NC>
__device__ __forceinline__ void ip_sweep_source(WLane Wm, const cplx* Vs, const cplx* Cl, int s, bool& singular) {
    if constexpr (K < M) {
        ip_source_reduced<M, K, NC>(Wm, Vs, s, singular);
        background_tpb<M, K, NC>(Wm, Cl, singular);
    } else {
        ip_source_full<M, NC>(Wm, Vs, s, singular);
    }
}
template <int M, int K>
__device__ __forceinline__ bool ip_sweep_nonfinite(WLane Wm) {
    // only the entries the sweep writes can go bad: the K filter columns and the K x (M-K) block J (the rest of
    // W_hat is the constant [0; -I]) -- 20 of 36 entries at M = 6, K = 2, i.e. 0.27 GB less DRAM read per sweep
    bool bad = false;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int c = 0; c < M; ++c)
            if (c < K || j < K) {
                const cplx v = Wm[j * M + c];
                if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
            }
    return bad;
}
template <int M, int K>
__device__ __forceinline__ void ip_sweep_bin(WLane Wm, const cplx* Vl, const cplx* Cl, const double* wscale_b,
                                             bool& singular, bool& bad) {
    constexpr int NE = oiva_tri(M);
    if (wscale_b) ip_sweep_rescale<M, K>(Wm, wscale_b);
#pragma unroll 1
    for (int s = 0; s < K; ++s) ip_sweep_source<M, K, true>(Wm, Vl + (size_t)s * NE * OIVA_GROUP, Cl, s, singular);
    bad = ip_sweep_nonfinite<M, K>(Wm);
}

// grid: ceil(G / 4) CTAs of 4 warps; warp <-> group gi, lane <-> bin; the grouped covariances Cg[gi] and Vg[gi][s]
// are read straight from global memory as 512-byte warp segments (staging them in shared memory by bulk TMA was
// measured slower: profiles/r01_notes).  status: one word per mixture.
template <int M, int K>
__global__ void __launch_bounds__(TPB_WARPS * 32) k_ip_update_tpb(cplx* __restrict__ Wg, const cplx* __restrict__ Vg,
                                                                  const cplx* __restrict__ Cg,
                                                                  const double* __restrict__ wscale, int* status,
                                                                  int F, int NG, long long G) {
    constexpr int NE = oiva_tri(M);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * TPB_WARPS + warp;
    if (gi >= G) return;  // whole warp
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + lane;
    if (f >= F) return;  // padded lanes of the last group
    const WLane Wm = {Wg + (size_t)gi * M * M * OIVA_GROUP + lane};  // this bin's W_hat
    bool singular = false, bad = false;
    ip_sweep_bin<M, K>(Wm, Vg + (size_t)gi * K * NE * OIVA_GROUP + lane, Cg + (size_t)gi * NE * OIVA_GROUP + lane,
                       wscale ? wscale + b * K : nullptr, singular, bad);
    if (singular || bad)
        atomicOr(status + b, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
}

// Initial W_hat written straight into the grouped layout, one thread per bin (overiva.py:89-123): the K filter columns
// from the identity (W0 == nullptr) or from W0 (R, M, K) row-major, J from the K x K solve of background_tpb, the -I
// block, zeros on the padded bins of a mixture's last group.  One launch instead of k_init_demix + the regrouping of
// its row-major result.  (init_eig keeps the row-major path: its eigenvectors come from k_eigh.)
template <int M, int K>
__global__ void __launch_bounds__(TPB_WARPS * 32) k_init_grouped(cplx* __restrict__ Wg, const cplx* __restrict__ Cg,
                                                                 const cplx* __restrict__ W0, int* status, int F, int NG,
                                                                 long long G) {
    constexpr int NE = oiva_tri(M);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * TPB_WARPS + warp;
    if (gi >= G) return;  // whole warp
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + lane;
    const bool ok = f < F;
    const WLane Wm = {Wg + (size_t)gi * M * M * OIVA_GROUP + lane};
    const cplx* w0 = W0 ? W0 + ((size_t)b * F + (ok ? f : 0)) * M * K : nullptr;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int c = 0; c < M; ++c) {
            cplx v = cmake(0.0, 0.0);
            if (ok) {
                if (c < K) v = w0 ? w0[j * K + c] : cmake(j == c ? 1.0 : 0.0, 0.0);
                else if (j == c) v = cmake(-1.0, 0.0);  // (c >= K implies K < M: the [K:, K:] = -I block)
            }
            Wm[j * M + c] = v;
        }
    if (!ok) return;
    bool singular = false;
    background_tpb<M, K, true>(Wm, Cg + (size_t)gi * NE * OIVA_GROUP + lane, singular);
    if (singular) atomicOr(status + b, OIVA_STATUS_SINGULAR);
}

}  // namespace oiva
