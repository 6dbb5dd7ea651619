// Determined IP sweep (K = M) for 7 and 8 channels with TWO lanes per bin.
//
// The thread-per-bin sweep (solve_tpb.cuh) keeps the whole M x M system of a bin in one thread's registers; at M = 8 that
// is 64 complex numbers plus the right-hand side and does not fit (the row-owner kernel of solve.cu took over: one row
// per lane, every pivot step a chain of shuffles and shared-memory round trips -- ncu: 6.7 short-scoreboard stalls per
// issued instruction, fp64 pipe 26 % busy, 5.4 ms per 256 mixtures, 40 % of that shape's epoch).  Here lanes l and l + 16
// of a warp share a bin: each keeps FOUR rows of [W_hat^H V_s | e_s] in registers (32 complex), a warp covers 16 bins (half
// a bin group), and all accesses to the grouped arrays stay coalesced (16 consecutive bins = 256-byte segments).  Per
// source (reference: overiva.py:181-186): the rows are built straight from global memory; Gauss-Jordan with partial
// pivoting over the 8 rows -- the local candidate of each lane by selects, ONE exchange with the partner lane per pivot
// (magnitude, then the normalised pivot row), elimination in the lane's other rows; the solution component of a row is
// written to W_hat by the lane that holds it; the normalisation w^H V w is summed as (rows 0-3) + (rows 4-7).  7 channels
// run as an 8 x 8 system with an identity row / column.  The per-source body (PairSweep) is shared with the resident
// single-mixture loop, which runs it on shared memory for every determined shape with M >= 3.  Same pivot rule as LAPACK's zgetrf (largest |re| + |im|, the
// lower row index wins a tie), so the result agrees with the other sweep kernels to rounding.
#pragma once
#include "solve_tpb.cuh"

namespace oiva {

constexpr int PAIR_WARPS = 4;

// One source of the determined sweep for the bin shared by lanes l and l ^ 16 of a warp (h = which half of the rows this
// lane holds).  Wm / Vs: the bin's W_hat and the lower triangle of V_s in a grouped array (global memory, NC = true, or
// shared memory, NC = false: the resident single-mixture loop).  M odd runs as an (M + 1) x (M + 1) system with an
// identity row / column.  Every lane of the warp must call it (shuffles); `ok` = false lanes write nothing.
template <int M, bool NC>
struct PairSweep {
    static constexpr int RH = (M + 1) / 2, N = 2 * RH;

    __device__ static __forceinline__ void source(WLane Wm, const cplx* Vs, int s, int h, bool ok, bool& singular) {
        // rows r = RH h + i of A = W_hat^H V_s (A[r][c] = sum_j conj(W[j][r]) V[j][c]), right-hand side e_s
        cplx A[RH][N], rhs[RH];
#pragma unroll
        for (int i = 0; i < RH; ++i) {
#pragma unroll
            for (int c = 0; c < N; ++c) A[i][c] = cmake(0.0, 0.0);
            const int r = h * RH + i;
            rhs[i] = cmake(r == s ? 1.0 : 0.0, 0.0);
            if (M < N && r == N - 1) A[i][N - 1] = cmake(1.0, 0.0);  // odd M: identity row / column N - 1
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
            cplx vrow[M];
#pragma unroll
            for (int c = 0; c < M; ++c) vrow[c] = herm_load<NC>(Vs, j, c);
#pragma unroll
            for (int i = 0; i < RH; ++i) {
                const int r = h * RH + i;
                if (r < M) {
                    const cplx a = Wm[j * M + r];
#pragma unroll
                    for (int c = 0; c < M; ++c) cfmac(A[i][c], a, vrow[c]);
                }
            }
        }
        // Gauss-Jordan with partial pivoting over the N rows of the pair
        bool used[RH];
        int col_of[RH];
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            used[i] = false;
            col_of[i] = 0;
        }
        static_for<N>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            double best = -1.0;
            int bi = 0;
#pragma unroll
            for (int i = 0; i < RH; ++i) {
                const double m = used[i] ? -1.0 : fabs(A[i][c].x) + fabs(A[i][c].y);
                if (m > best) {
                    best = m;
                    bi = i;
                }
            }
            const double pbest = __shfl_xor_sync(0xffffffffu, best, 16);
            const bool mine = best > pbest || (best == pbest && h == 0);  // (a tie goes to the lower row index)
            if (!((mine ? best : pbest) > 0.0)) singular = true;          // zero or NaN pivot column
            // this lane's candidate row (columns c..N-1 and the right-hand side), normalised by its pivot element
            cplx cand[N + 1];
#pragma unroll
            for (int col = c; col < N; ++col) {
                cplx v = A[0][col];
#pragma unroll
                for (int i = 1; i < RH; ++i) {
                    v.x = bi == i ? A[i][col].x : v.x;
                    v.y = bi == i ? A[i][col].y : v.y;
                }
                cand[col] = v;
            }
            {
                cplx v = rhs[0];
#pragma unroll
                for (int i = 1; i < RH; ++i) {
                    v.x = bi == i ? rhs[i].x : v.x;
                    v.y = bi == i ? rhs[i].y : v.y;
                }
                cand[N] = v;
            }
            // (the fast reciprocal only on shared memory: in the batched kernel, at the register limit with M = 8, it
            // lets ptxas overlap the pivots and spill 2 KB per thread -- sweep 3.2 -> 4.6 ms per 256 mixtures)
            cplx inv;
            if constexpr (NC) inv = crecip(cand[c]);
            else inv = crecip_fast(cand[c]);
#pragma unroll
            for (int col = c + 1; col <= N; ++col) cand[col] = cmul(cand[col], inv);
            // the winner's row reaches both lanes
            cplx prow[N + 1];
#pragma unroll
            for (int col = c + 1; col <= N; ++col) {
                const cplx other = shfl_xor_c(cand[col], 16);
                prow[col].x = mine ? cand[col].x : other.x;
                prow[col].y = mine ? cand[col].y : other.y;
            }
#pragma unroll
            for (int i = 0; i < RH; ++i) {
                const bool is_p = mine && bi == i;
                const cplx f = A[i][c];
#pragma unroll
                for (int col = c + 1; col < N; ++col) {
                    cplx v = A[i][col];
                    cfms(v, f, prow[col]);
                    A[i][col].x = is_p ? prow[col].x : v.x;
                    A[i][col].y = is_p ? prow[col].y : v.y;
                }
                cplx v = rhs[i];
                cfms(v, f, prow[N]);
                rhs[i].x = is_p ? prow[N].x : v.x;
                rhs[i].y = is_p ? prow[N].y : v.y;
                used[i] = used[i] || is_p;
                col_of[i] = is_p ? c : col_of[i];
            }
        });
        // the row that pivoted column c holds component c of w_s = (W_hat^H V_s)^-1 e_s
        if (ok) {
#pragma unroll
            for (int i = 0; i < RH; ++i)
                if (col_of[i] < M) Wm[col_of[i] * M + s] = rhs[i];
        }
        __syncwarp();
        // normalise: w_s /= sqrt(w_s^H V_s w_s)                                                 overiva.py:185-186
        cplx w[M];
#pragma unroll
        for (int j = 0; j < M; ++j) w[j] = Wm[j * M + s];
        cplx dpart = cmake(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < M; ++r) {
            if (r / RH == h) {
                cplx u = cmake(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < M; ++j) cfma(u, herm_load<NC>(Vs, r, j), w[j]);
                cfmac(dpart, w[r], u);
            }
        }
        const cplx dother = shfl_xor_c(dpart, 16);
        const cplx d = h == 0 ? cadd(dpart, dother) : cadd(dother, dpart);  // (first half of the rows) + (second half)
        const cplx inv = crsqrt_pos(d);
        __syncwarp();  // (every lane has read the un-normalised column before anybody overwrites it)
        if (ok) {
#pragma unroll
            for (int r = 0; r < M; ++r)
                if (r / RH == h) Wm[r * M + s] = cmul(w[r], inv);
        }
        __syncwarp();
    }
};

template <int M>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_ip_update_pair(cplx* __restrict__ Wg, const cplx* __restrict__ Vg,
                                                                    const double* __restrict__ wscale, int* status, int F,
                                                                    int NG, long long G) {
    static_assert(M == 7 || M == 8, "instantiated for the shapes the thread-per-bin sweep does not cover");
    constexpr int RH = PairSweep<M, true>::RH, NE = oiva_tri(M), K = M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long hw = (long long)blockIdx.x * PAIR_WARPS + warp;  // half-group
    if (hw >= 2 * G) return;                                         // whole warp
    const long long gi = hw >> 1;
    const int h = lane >> 4;                                 // which half of the rows this lane holds
    const int l32 = (int)(hw & 1) * 16 + (lane & 15);        // the bin's lane position inside its group
    const long long b = gi / NG;
    const bool ok = (int)(gi - b * NG) * OIVA_GROUP + l32 < F;  // (padded bins run along -- the shuffles need every lane --
                                                                 //  on the zeros stored there, and write nothing)
    const WLane Wm = {Wg + (size_t)gi * M * M * OIVA_GROUP + l32};
    const cplx* Vl = Vg + (size_t)gi * K * NE * OIVA_GROUP + l32;

    if (wscale && ok) {  // W /= gamma (laplace) or sqrt(gamma) (gauss): this lane's rows      overiva.py:161-167
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            const int j = h * RH + i;
            if (j < M) {
#pragma unroll
                for (int k = 0; k < K; ++k) Wm[j * M + k] = cscale(Wm[j * M + k], wscale[b * K + k]);
            }
        }
    }
    __syncwarp();
    bool singular = false;
#pragma unroll 1
    for (int s = 0; s < K; ++s) PairSweep<M, true>::source(Wm, Vl + (size_t)s * NE * OIVA_GROUP, s, h, ok, singular);
    if (ok) {
        bool bad = false;
#pragma unroll
        for (int r = 0; r < M; ++r)
            if (r / RH == h) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cplx v = Wm[r * M + k];
                    if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
                }
            }
        if (singular || bad)
            atomicOr(status + b, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
    }
}

}  // namespace oiva
