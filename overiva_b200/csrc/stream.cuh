// Streaming demix kernels (lane <-> frame over the planar rows):
//   k_demix_power : r2part[b][chunk][k][t] = sum_{f in chunk} |w_k(f)^H x(f,t)|^2     (overiva.py:140,152-155)
//   k_demix_output: Y[b][t][f][k] = weff_k(f)^H x(f,t)                                 (overiva.py:192-199)
//   k_project_rows: planar rows of the K-channel signal E_K^H x                        (auxiva_pca.py:79-81)
// Every element of X is read exactly once, by one lane, as part of a contiguous 256-byte warp segment, so
// these kernels read global memory directly (no shared-memory staging: there is no reuse to capture).
#pragma once
#include "common.cuh"

namespace oiva {

constexpr int STREAM_WARPS = 4;

struct StreamParams {
    const void* Xp;
    const cplx* W;   // filters: element (row, c, k) at W[row*w_row + c*w_c + k]
    long long w_row;
    int w_c;
    RowLayout L;
    int F, K, k0;
    int NCH, NBC;     // chunks per mixture, bins per chunk (power kernel)
    double* r2part;   // (B, NCH, K, Tp)
    void* Y;          // (B, T, F, K) interleaved complex ST
    void* Xr;         // planar rows with K channels
    RowLayout Lr;     // layout of Xr
    int NBF;          // bins per CTA (output kernel)
};

// load the 2M plane values of one frame
template <typename ST, int M>
__device__ __forceinline__ void load_frame(double (&xr)[M], double (&xi)[M], const ST* __restrict__ tile, int pitch,
                                           int tloc) {
#pragma unroll
    for (int c = 0; c < M; ++c) {
        xr[c] = ld_nc(tile + (size_t)(2 * c) * pitch + tloc);
        xi[c] = ld_nc(tile + (size_t)(2 * c + 1) * pitch + tloc);
    }
}

// y_k = sum_c conj(w[c][k]) x_c for KC sources; sources >= kmax are skipped (y = 0)
template <int M, int KC>
__device__ __forceinline__ void demix_frame(double (&yr)[KC], double (&yi)[KC], const double (&xr)[M],
                                            const double (&xi)[M], const cplx* __restrict__ Wrow, int w_c, int k0,
                                            int kmax) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        yr[k] = 0.0;
        yi[k] = 0.0;
        if (k0 + k < kmax) {
#pragma unroll
            for (int c = 0; c < M; ++c) {
                const cplx w = ld_nc_c(Wrow + (size_t)c * w_c + k0 + k);
                // conj(w) * x = (wr xr + wi xi) + i (wr xi - wi xr)
                yr[k] = fma(w.x, xr[c], yr[k]);
                yr[k] = fma(w.y, xi[c], yr[k]);
                yi[k] = fma(w.x, xi[c], yi[k]);
                yi[k] = fma(-w.y, xr[c], yi[k]);
            }
        }
    }
}

// grid (NCH, ceil(slots/4), B); warp <-> 32-frame slot, loop over the bins of the chunk
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(STREAM_WARPS * 32) k_demix_power(const StreamParams p) {
    const RowLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, ch = blockIdx.x;
    const int slot = blockIdx.y * STREAM_WARPS + warp;
    const int t = slot * 32 + lane;
    const int Tp = L.frame_pitch();
    if (slot * 32 >= Tp) return;
    const int ti = t / L.TT, tloc = t - ti * L.TT;
    const bool valid = t < L.T;
    const int pitch = L.pitch(ti < L.nT ? ti : L.nT - 1);
    const int f0 = ch * p.NBC;
    const int f1 = min(p.F, f0 + p.NBC);
    const ST* Xp = reinterpret_cast<const ST*>(p.Xp);
    const size_t row_elems = L.row_elems();
    double acc[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.0;
    if (valid) {
#pragma unroll 2
        for (int f = f0; f < f1; ++f) {
            const size_t row = (size_t)b * p.F + f;
            double xr[M], xi[M], yr[KC], yi[KC];
            load_frame<ST, M>(xr, xi, Xp + row * row_elems + L.tile_off(ti), pitch, tloc);
            demix_frame<M, KC>(yr, yi, xr, xi, p.W + row * p.w_row, p.w_c, p.k0, p.K);
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[k] = fma(yr[k], yr[k], fma(yi[k], yi[k], acc[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k)
        if (p.k0 + k < p.K) p.r2part[(((size_t)b * p.NCH + ch) * p.K + p.k0 + k) * Tp + t] = acc[k];
}

// grid (ceil(F/NBF), slots, B); CTA <-> NBF bins x one 32-frame slot; shared-memory transpose so that the
// (T,F,K)-ordered output is written in contiguous NBF*K-element runs
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(STREAM_WARPS * 32) k_demix_output(const StreamParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename std::conditional<sizeof(ST) == 8, double2, float2>::type OutC;
    OutC* tileo = reinterpret_cast<OutC*>(smem_raw);  // [32][NBF*K + 1]
    const RowLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z;
    const int f0 = blockIdx.x * p.NBF;
    const int nb = min(p.NBF, p.F - f0);
    const int slot = blockIdx.y;
    const int t = slot * 32 + lane;
    const int ti = t / L.TT, tloc = t - ti * L.TT;
    const bool valid = t < L.T;
    const int pitch = L.pitch(ti < L.nT ? ti : L.nT - 1);
    const int opitch = p.NBF * p.K + 1;
    const ST* Xp = reinterpret_cast<const ST*>(p.Xp);
    const size_t row_elems = L.row_elems();
    for (int k0 = 0; k0 < p.K; k0 += KC) {
        for (int fb = warp; fb < nb; fb += STREAM_WARPS) {
            const size_t row = (size_t)b * p.F + f0 + fb;
            double xr[M], xi[M], yr[KC], yi[KC];
            if (valid) {
                load_frame<ST, M>(xr, xi, Xp + row * row_elems + L.tile_off(ti), pitch, tloc);
                demix_frame<M, KC>(yr, yi, xr, xi, p.W + row * p.w_row, p.w_c, k0, p.K);
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k0 + k < p.K) {
                        OutC o;
                        o.x = (ST)yr[k];
                        o.y = (ST)yi[k];
                        tileo[lane * opitch + fb * p.K + k0 + k] = o;
                    }
            }
        }
    }
    __syncthreads();
    OutC* Y = reinterpret_cast<OutC*>(p.Y);
    const int run = nb * p.K;
    for (int i = threadIdx.x; i < 32 * run; i += STREAM_WARPS * 32) {
        const int tl = i / run, j = i - tl * run;
        const int tt = slot * 32 + tl;
        if (tt < L.T) Y[(((size_t)b * L.T + tt) * p.F + f0) * p.K + j] = tileo[tl * opitch + j];
    }
}

// grid (ceil(R/4), slots): warp <-> (row, slot); writes the K-channel planar row
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(STREAM_WARPS * 32) k_project_rows(const StreamParams p, long long R) {
    const RowLayout& L = p.L;
    const RowLayout& Lr = p.Lr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * STREAM_WARPS + warp;
    if (row >= R) return;
    const int slot = blockIdx.y;
    const int t = slot * 32 + lane;
    const int ti = t / L.TT, tloc = t - ti * L.TT;
    const bool valid = t < L.T;
    const int pitch = L.pitch(ti < L.nT ? ti : L.nT - 1);
    const ST* Xp = reinterpret_cast<const ST*>(p.Xp);
    ST* Xr = reinterpret_cast<ST*>(p.Xr);
    // Lr has the same T and hence the same tiling only if TT matches; address through Lr explicitly
    const int tir = t / Lr.TT, tlocr = t - tir * Lr.TT;
    if (tir >= Lr.nT) return;
    const int pitchr = Lr.pitch(tir);
    if (tlocr >= pitchr) return;
    double xr[M], xi[M];
#pragma unroll
    for (int c = 0; c < M; ++c) xr[c] = xi[c] = 0.0;
    if (valid) load_frame<ST, M>(xr, xi, Xp + (size_t)row * L.row_elems() + L.tile_off(ti), pitch, tloc);
    for (int k0 = 0; k0 < p.K; k0 += KC) {
        double yr[KC], yi[KC];
        demix_frame<M, KC>(yr, yi, xr, xi, p.W + row * p.w_row, p.w_c, k0, p.K);
#pragma unroll
        for (int k = 0; k < KC; ++k)
            if (k0 + k < p.K) {
                ST* dst = Xr + (size_t)row * Lr.row_elems() + Lr.tile_off(tir);
                dst[(size_t)(2 * (k0 + k)) * pitchr + tlocr] = (ST)yr[k];
                dst[(size_t)(2 * (k0 + k) + 1) * pitchr + tlocr] = (ST)yi[k];
            }
    }
}

}  // namespace oiva
