// Fused "IP sweep + statistic of the next epoch" kernel (M <= 6, K <= 4): solve(it) -> power(it+1) per bin group.
//
// The only ordering constraint between the IP sweep of epoch `it` and the demix-power pass of epoch `it+1` is
// per bin, and both kernels map lane <-> bin.  So one warp first runs the thread-per-bin sweep for its 32 bins
// (latency-bound: long dependent chains, ~140 registers) and then, with the fresh w_k still in registers, streams
// the group's frames for the statistic (bandwidth-bound: lots of idle issue slots).  With a dozen warps per SM in
// different phases the sweep's latency hides behind the streaming of the others, W_hat makes no round trip through
// memory between the two steps, and one launch per epoch disappears.
// (reference steps: overiva.py:161-167,176-190 then overiva.py:140,152-155 of the following epoch)
#pragma once
#include "solve_tpb.cuh"
#include "stream.cuh"

namespace oiva {

struct FusedParams {
    cplx* Wg;             // [G][M*M][32]
    const cplx* Vg;       // [G][K][NE][32]
    const cplx* Cg;       // [G][NE][32]
    const double* wscale; // (B, K) or null
    int* status;
    const void* Xg;
    double* r2part;       // (B, NG, K, Tp)
    GroupLayout L;
    long long G;
};

constexpr int FUSED_WARPS = 4;

template <typename ST, int M, int K>
__global__ void __launch_bounds__(FUSED_WARPS * 32) k_solve_power(const FusedParams p) {
    typedef typename StoreC<ST>::type XC;
    constexpr int NE = oiva_tri(M);
    const GroupLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * FUSED_WARPS + warp;
    if (gi >= p.G) return;  // whole warp
    const long long b = gi / L.NG;
    const int f = (int)(gi - b * L.NG) * OIVA_GROUP + lane;
    const bool valid = f < L.F;
    const WLane Wm = {p.Wg + (size_t)gi * M * M * OIVA_GROUP + lane};

    // ---- phase 1: IP sweep of this lane's bin ---------------------------------------------------------
    if (valid) {
        const cplx* Vb = p.Vg + (size_t)gi * K * NE * OIVA_GROUP + lane;
        const cplx* Cb = p.Cg + (size_t)gi * NE * OIVA_GROUP + lane;
        bool singular = false;
        if (p.wscale) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double sc = p.wscale[b * K + k];
#pragma unroll
                for (int j = 0; j < M; ++j) Wm[j * M + k] = cscale(Wm[j * M + k], sc);
            }
        }
#pragma unroll 1
        for (int s = 0; s < K; ++s) {
            const cplx* sV = Vb + (size_t)s * NE * OIVA_GROUP;
            if constexpr (K < M) ip_source_reduced<M, K, false>(Wm, sV, s, singular);
            else ip_source_full<M, false>(Wm, sV, s, singular);
            background_tpb<M, K, false>(Wm, Cb, singular);
        }
        bool bad = false;
#pragma unroll
        for (int i = 0; i < M * M; ++i) {
            const cplx v = Wm[i];
            if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
        }
        if (singular || bad)
            atomicOr(p.status, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
    }

    // ---- phase 2: statistic of the next epoch with the updated filters --------------------------------
    cplx w[M][K];
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) w[c][k] = valid ? Wm[c * M + k] : cmake(0.0, 0.0);
    const int Tp = L.frame_pitch();
    const XC* xg = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems();
    for (int t0 = 0; t0 < Tp; t0 += POWER_FB) {
        double v[K][POWER_FB];
#pragma unroll
        for (int j = 0; j < POWER_FB; ++j) {
            cplx x[M], y[K];
            const int t = t0 + j;
            if (t < L.T) {
#pragma unroll
                for (int c = 0; c < M; ++c) x[c] = ldg_x(xg + ((size_t)t * M + c) * OIVA_GROUP + lane);
                demix_frame<M, K>(y, x, w);
#pragma unroll
                for (int k = 0; k < K; ++k) v[k][j] = fma(y[k].x, y[k].x, y[k].y * y[k].y);
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) v[k][j] = 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int lvl = 0; lvl < 3; ++lvl) {
                const int H = POWER_FB >> (lvl + 1);
                const int off = 16 >> lvl;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int n = 0; n < H; ++n) {
                    const double lo = v[k][n], hi = v[k][n + H];
                    const double send = up ? lo : hi;
                    const double keep = up ? hi : lo;
                    v[k][n] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            double s = v[k][0];
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if ((lane & 3) == 0) p.r2part[((size_t)gi * K + k) * Tp + t0 + (lane >> 2)] = s;
        }
    }
}

}  // namespace oiva
