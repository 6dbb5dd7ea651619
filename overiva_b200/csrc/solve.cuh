// Per-bin small-matrix linear algebra in fp64 registers ("row-owner" layout).
//
// A group of G = next_pow2(M) lanes owns one frequency bin; lane i of the group holds row i of the
// working matrix in registers (compile-time M => static register indexing), so a warp works on 32/G
// bins at once.  Pivot search, pivot-row broadcast and dot products are xor/idx shuffles inside the
// group; there is no shared-memory traffic for the factorisation itself and no __syncthreads.
//
// Replaces (reference): the stacked zgemm + zgesv of overiva.py:181-182, the normalisation :185-186 and
// update_J_from_orth_const :96-98, all of which run per bin on (M x M) complex128 matrices.
#pragma once
#include "common.cuh"

namespace oiva {

template <int M>
struct Grp {
    static constexpr int G = (M <= 1) ? 1 : (M <= 2) ? 2 : (M <= 4) ? 4 : (M <= 8) ? 8 : 16;
    static constexpr int BINS = 32 / G;  // bins per warp
};

// sum over the lanes of a group
template <int G>
__device__ __forceinline__ cplx group_sum(cplx v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
    }
    return v;
}
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// Gauss-Jordan elimination with partial pivoting (pivot = largest |re|+|im| among the rows not yet used,
// ties to the lowest row -- LAPACK's izamax rule) on the augmented rows A[0..M] held one per lane.
// Columns [0, n) are eliminated; columns [n, M] are right-hand sides.  Lanes with gl >= nrows hold zero
// rows and never pivot.  On return the lane that pivoted column c holds row c of [I | A^-1 B]; its column
// index is returned (-1 for lanes that did not pivot).  `singular` is set when a pivot is zero / NaN.
template <int M, int G>
__device__ __forceinline__ int gauss_jordan(cplx (&A)[M + 1], int n, int gl, int lane, bool row_valid,
                                            int& singular) {
    int mycol = -1;
#pragma unroll
    for (int c = 0; c < M; ++c) {
        if (c < n) {
            double bm = (row_valid && mycol < 0) ? fabs(A[c].x) + fabs(A[c].y) : -1.0;
            int bl = gl;
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) {
                double om = __shfl_xor_sync(0xffffffffu, bm, off);
                int ol = __shfl_xor_sync(0xffffffffu, bl, off);
                if (om > bm || (om == bm && ol < bl)) {
                    bm = om;
                    bl = ol;
                }
            }
            if (!(bm > 0.0)) singular = 1;
            const int src = (lane & ~(G - 1)) + bl;
            const cplx rinv = crecip(shfl_c(A[c], src));
            const bool isp = (gl == bl);
            const cplx f = A[c];
#pragma unroll
            for (int c2 = c + 1; c2 <= M; ++c2) {
                cplx pr = cmul(shfl_c(A[c2], src), rinv);
                if (isp)
                    A[c2] = pr;
                else
                    cfms(A[c2], f, pr);
            }
            A[c] = isp ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
            if (isp) mycol = c;
        }
    }
    return mycol;
}

// OverIVA's background refresh for one bin: tmp = W^H C (K x M), J = tmp[:, :K]^-1 tmp[:, K:], written to
// sW[r][K..M) for r < K.  sW: M x M row-major in shared memory (the reference's W_hat[f]); C in global.
template <int M, int G>
__device__ __forceinline__ void update_background(cplx* sW, const cplx* __restrict__ C, int K, int gl, int lane,
                                                  int& singular) {
    cplx tmp[M + 1];
#pragma unroll
    for (int c = 0; c <= M; ++c) tmp[c] = cmake(0.0, 0.0);
    const bool valid = gl < K;
    if (valid) {
        for (int j = 0; j < M; ++j) {
            const cplx a = sW[j * M + gl];
#pragma unroll
            for (int c = 0; c < M; ++c) cfmac(tmp[c], a, ld_nc_c(&C[j * M + c]));
        }
    }
    const int r = gauss_jordan<M, G>(tmp, K, gl, lane, valid, singular);
    __syncwarp();
    if (r >= 0) {
#pragma unroll
        for (int c = 0; c < M; ++c)
            if (c >= K) sW[r * M + c] = tmp[c];
    }
    __syncwarp();
}

}  // namespace oiva
