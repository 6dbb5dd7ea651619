// Streaming demix kernels over the grouped layout (lane <-> frequency bin):
//   k_demix_power : r2part[b][g][k][t] = sum_{32 bins of group g} |w_k(f)^H x(f,t)|^2   (overiva.py:140,152-155)
//   k_demix_output: Y[b][t][f][k] = weff_k(f)^H x(f,t)                                   (overiva.py:192-199)
//   k_project_rows: grouped samples of the K-channel signal E_K^H x                      (auxiva_pca.py:79-81)
// Each lane keeps ITS bin's demixing vectors in registers for the whole frame range, every element of X is
// read exactly once as part of a 512-byte warp segment, so these kernels read global memory directly (no
// shared-memory staging: there is no reuse to capture), and the (T,F,K)-ordered output is written as
// 32*K consecutive complex numbers per warp and frame.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace oiva {

constexpr int POWER_FB = 8;  // frames reduced together in the power kernel
// one warp per chunk of KC sources, K <= 16: the block size bound of the one-warp-per-chunk kernels.  (A blanket bound of
// 512 threads capped them at 128 registers: the KC = 3 and 4 variants with M >= 4 spilled their filters.)
__host__ __device__ constexpr int stream_max_threads(int KC) { return 32 * ((OIVA_MAX_M + KC - 1) / KC); }

// Sum over the 32 lanes (bins) of NF per-lane values (NF = 8, 4 or 2 frames): log2(NF) transposing exchanges (each halves
// the live values), then plain xor sums.  Whatever NF, a frame's 32 values are added in the same order (partners at lane
// distance 16, 8, 4, 2, 1), so the result does not depend on how the frames are blocked.  Returns the sum for frame
// lane / (32 / NF), valid on the lanes with lane % (32 / NF) == 0.
template <int NF>
__device__ __forceinline__ double lane_sum_frames(double (&v)[NF], int lane) {
    int off = 16;
#pragma unroll
    for (int H = NF / 2; H >= 1; H >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int n = 0; n < H; ++n) {
            const double lo = v[n], hi = v[n + H];
            const double send = up ? lo : hi;
            const double keep = up ? hi : lo;
            v[n] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    double s = v[0];
#pragma unroll
    for (int o = 16 / NF; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

struct StreamParams {
    const void* Xg;
    const cplx* W;   // filters: element (c, k) of bin (b, f): row-major W[(b*F+f)*w_row + c*w_c + k], or, when
    long long w_row; //          w_grouped, W[((gi*w_row) + c*w_c + k)*32 + lane]  (w_row = M * w_c in both cases)
    int w_c;
    int w_grouped;
    GroupLayout L;
    int K;
    int nsplit;       // frame splits per group (grid.y)
    double* r2part;   // (B, NG, K, Tp)
    void* Y;          // (B, T, F, K) interleaved complex ST
    void* Xr;         // grouped samples with K channels
    double* Pfull;    // power kernels: |y_k(f, t)|^2 per bin, grouped [gi][k][Tp][32] (ILRMA's spectrogram model), or nullptr
    const cplx* Zg;   // output kernels: projection-back scales z [gi][K][32] from k_projback_z (stream.cu), or nullptr.
                      // (Computing z inside the output kernels was tried: the M x M products next to the filter
                      // registers spill at the 128-register cap of these kernels from M = 5 on -- 66 ms instead of 3 ms
                      // for the M = K = 8 output.)
};

// w_k <- w_k z_k with z from the grouped array of k_projback_z
template <int M, int KC>
__device__ __forceinline__ void projback_apply(cplx (&w)[M][KC], const cplx* __restrict__ Zl, int k0, int K) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (k0 + k < K) {
            const cplx z = ld_nc_c(Zl + (size_t)(k0 + k) * OIVA_GROUP);
#pragma unroll
            for (int a = 0; a < M; ++a) w[a][k] = cmul(w[a][k], z);
        }
    }
}

// this lane's demixing vectors: w[c][k] for KC sources starting at k0 (zero beyond K / beyond F)
template <int M, int KC>
__device__ __forceinline__ void load_filters(cplx (&w)[M][KC], const StreamParams& p, long long row, long long gi,
                                             int lane, bool bin_ok, int k0) {
    const cplx* base = p.w_grouped ? p.W + (size_t)gi * p.w_row * OIVA_GROUP + lane : p.W + row * p.w_row;
    const int es = p.w_grouped ? OIVA_GROUP : 1;
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            w[c][k] = cmake(0.0, 0.0);
            if (bin_ok && k0 + k < p.K) w[c][k] = ld_nc_c(base + ((size_t)c * p.w_c + k0 + k) * es);
        }
}

// y_k = sum_c conj(w[c][k]) x_c
template <int M, int KC>
__device__ __forceinline__ void demix_frame(cplx (&y)[KC], const cplx (&x)[M], const cplx (&w)[M][KC]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        y[k] = cmake(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < M; ++c) cfmac(y[k], w[c][k], x[c]);
    }
}

// grid (G, nsplit), block = 32 * ceil(K/KC): warp w handles sources [w*KC, w*KC+KC) of group blockIdx.x over the
// frame blocks of its split.  After |y|^2 the 32 lanes (bins) of POWER_FB frames are summed with a transposing
// butterfly (each exchange halves the live values), the group's partial statistic goes to r2part.
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(stream_max_threads(KC)) k_demix_power(const StreamParams p) {
    typedef typename StoreC<ST>::type XC;
    const GroupLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gi = blockIdx.x;
    const long long b = gi / L.NG;
    const int g = (int)(gi - b * L.NG);
    const int f = g * OIVA_GROUP + lane;
    const bool bin_ok = f < L.F;
    const int k0 = warp * KC;
    const int Tp = L.frame_pitch();
    cplx w[M][KC];
    load_filters<M, KC>(w, p, b * L.F + (bin_ok ? f : 0), gi, lane, bin_ok, k0);
    const XC* xg = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems();
    const int nblk = Tp / POWER_FB;  // blocks over the padded frame range: the padding is written as zeros
    const int blk0 = (int)((long long)nblk * blockIdx.y / p.nsplit);
    const int blk1 = (int)((long long)nblk * (blockIdx.y + 1) / p.nsplit);
    for (int blk = blk0; blk < blk1; ++blk) {
        const int t0 = blk * POWER_FB;
        double v[KC][POWER_FB];
#pragma unroll
        for (int j = 0; j < POWER_FB; ++j) {
            cplx x[M], y[KC];
            const int t = t0 + j;
            if (t < L.T) {
#pragma unroll
                for (int c = 0; c < M; ++c) x[c] = ldg_x(xg + ((size_t)t * M + c) * OIVA_GROUP + lane);
                demix_frame<M, KC>(y, x, w);
#pragma unroll
                for (int k = 0; k < KC; ++k) v[k][j] = fma(y[k].x, y[k].x, y[k].y * y[k].y);
            } else {
#pragma unroll
                for (int k = 0; k < KC; ++k) v[k][j] = 0.0;
            }
        }
        if (p.Pfull) {  // (padding frames are written as zeros, padded bins hold zeros already)
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (k0 + k < p.K) {
#pragma unroll
                    for (int j = 0; j < POWER_FB; ++j)
                        p.Pfull[(((size_t)gi * p.K + k0 + k) * Tp + t0 + j) * OIVA_GROUP + lane] = v[k][j];
                }
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            // 8 -> 4 -> 2 -> 1 values across lane bits 4, 3, 2; then plain sums across bits 1, 0
            const double s = lane_sum_frames<POWER_FB>(v[k], lane);
            // lanes with (lane & 3) == 0 hold the sum over the 32 bins for frame t0 + (lane >> 2)
            if ((lane & 3) == 0 && k0 + k < p.K) p.r2part[((size_t)gi * p.K + k0 + k) * Tp + t0 + (lane >> 2)] = s;
        }
    }
}

// Same kernels for several warps per group: instead of every warp re-reading X from global memory, the CTA stages
// each block of POWER_FB frames ONCE in shared memory -- a contiguous piece of the grouped layout, i.e. one 1-D
// bulk-TMA copy into a 2-stage ring (full / empty mbarriers as in cov.cuh) -- and all warps read it from there
// (conflict-free LDS.128).  Warp = (source chunk of KC, sub-block of FBW frames): blockDim = 32 * ceil(K/KC) * (8/FBW),
// so that few source chunks still give 8 warps per CTA (config 5: M = 16, K = 4 -> 2 chunks x 4 sub-blocks).
// OUTPUT = false: the statistic (k_demix_power's result, same arithmetic); OUTPUT = true: the demixed samples Y
// (k_demix_output's result).  Measured: DESIGN.md.
// grid (G, nsplit); dynamic smem = 128 + 2 * POWER_FB * M * 32 * sizeof(XC)
// (M >= 13 runs at most 12 warps -- <= 8 source chunks of 2, or <= 3 chunks x 4 sub-blocks -- and needs ~170 registers;
// otherwise the block size follows from the launcher's rule: >= 8 chunks x 1 sub-block, 4-7 chunks x 2, 1-3 chunks x 4)
template <typename ST, int M, int KC, int FBW, bool OUTPUT>
__global__ void __launch_bounds__(M >= 13 ? 384 : (FBW == 8 ? stream_max_threads(KC) : (FBW == 4 ? 448 : 384)))
    k_demix_staged(const StreamParams p) {
    typedef typename StoreC<ST>::type XC;
    constexpr int NSUB = POWER_FB / FBW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr size_t stage_bytes = (size_t)POWER_FB * M * OIVA_GROUP * sizeof(XC);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + 2;
    unsigned char* stage0 = smem_raw + 128;
    const GroupLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], nwarps);
        mbar_init(&empty[1], nwarps);
        mbar_fence_init();
    }
    __syncthreads();
    const long long gi = blockIdx.x;
    const long long b = gi / L.NG;
    const int g = (int)(gi - b * L.NG);
    const int f = g * OIVA_GROUP + lane;
    const bool bin_ok = f < L.F;
    const int k0 = (warp / NSUB) * KC;
    const int j0 = (warp % NSUB) * FBW;  // this warp's frames inside every block
    const int Tp = L.frame_pitch();
    cplx w[M][KC];
    load_filters<M, KC>(w, p, b * L.F + (bin_ok ? f : 0), gi, lane, bin_ok, k0);
    if constexpr (OUTPUT) {
        if (p.Zg) projback_apply<M, KC>(w, p.Zg + (size_t)gi * p.K * OIVA_GROUP + lane, k0, p.K);
    }
    const XC* xg = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems();
    XC* Y = reinterpret_cast<XC*>(p.Y);
    const int nblk = Tp / POWER_FB;
    const int blk0 = (int)((long long)nblk * blockIdx.y / p.nsplit);
    const int blk1 = (int)((long long)nblk * (blockIdx.y + 1) / p.nsplit);
    // producer: thread 0 keeps one block in flight ahead of the consumers
    auto issue = [&](int blk, int stage, int use) {
        const int t0 = blk * POWER_FB;
        const int nfr = min(POWER_FB, L.T - t0);
        if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1);
        if (nfr > 0) {
            const uint32_t bytes = (uint32_t)((size_t)nfr * M * OIVA_GROUP * sizeof(XC));
            mbar_arrive_expect_tx(&full[stage], bytes);
            tma_load_1d(stage0 + (size_t)stage * stage_bytes, xg + (size_t)t0 * M * OIVA_GROUP, bytes, &full[stage]);
        } else {
            mbar_arrive(&full[stage]);  // a block made of padding frames only: nothing to copy
        }
    };
    if (threadIdx.x == 0 && blk0 < blk1) issue(blk0, 0, 0);
    int it = 0;
    for (int blk = blk0; blk < blk1; ++blk, ++it) {
        const int stage = it & 1, use = it >> 1;
        if (threadIdx.x == 0 && blk + 1 < blk1) issue(blk + 1, (it + 1) & 1, (it + 1) >> 1);
        mbar_wait(&full[stage], use & 1);
        const XC* xs = reinterpret_cast<const XC*>(stage0 + (size_t)stage * stage_bytes);
        const int t0 = blk * POWER_FB + j0;
        double v[KC][FBW];
#pragma unroll
        for (int j = 0; j < FBW; ++j) {
            cplx x[M], y[KC];
            if (t0 + j < L.T) {
#pragma unroll
                for (int c = 0; c < M; ++c) x[c] = widen(xs[((size_t)(j0 + j) * M + c) * OIVA_GROUP + lane]);
                demix_frame<M, KC>(y, x, w);
                if constexpr (OUTPUT) {
                    if (bin_ok) {
                        XC* dst = Y + (((size_t)b * L.T + t0 + j) * L.F + f) * p.K + k0;
#pragma unroll
                        for (int k = 0; k < KC; ++k)
                            if (k0 + k < p.K) narrow(dst[k], y[k]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KC; ++k) v[k][j] = fma(y[k].x, y[k].x, y[k].y * y[k].y);
                }
            } else if constexpr (!OUTPUT) {
#pragma unroll
                for (int k = 0; k < KC; ++k) v[k][j] = 0.0;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if constexpr (!OUTPUT) {
            if (p.Pfull) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k0 + k < p.K) {
#pragma unroll
                        for (int j = 0; j < FBW; ++j)
                            p.Pfull[(((size_t)gi * p.K + k0 + k) * Tp + t0 + j) * OIVA_GROUP + lane] = v[k][j];
                    }
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const double sres = lane_sum_frames<FBW>(v[k], lane);
                constexpr int LPF = OIVA_GROUP / FBW;  // lanes per frame after the exchanges
                if ((lane % LPF) == 0 && k0 + k < p.K) p.r2part[((size_t)gi * p.K + k0 + k) * Tp + t0 + lane / LPF] = sres;
            }
        }
    }
}

// grid (G, nsplit), block = 32 * ceil(K/KC)
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(stream_max_threads(KC)) k_demix_output(const StreamParams p) {
    typedef typename StoreC<ST>::type XC;
    const GroupLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gi = blockIdx.x;
    const long long b = gi / L.NG;
    const int g = (int)(gi - b * L.NG);
    const int f = g * OIVA_GROUP + lane;
    const bool bin_ok = f < L.F;
    const int k0 = warp * KC;
    cplx w[M][KC];
    load_filters<M, KC>(w, p, b * L.F + (bin_ok ? f : 0), gi, lane, bin_ok, k0);
    if (p.Zg) projback_apply<M, KC>(w, p.Zg + (size_t)gi * p.K * OIVA_GROUP + lane, k0, p.K);
    const XC* xg = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems();
    XC* Y = reinterpret_cast<XC*>(p.Y);
    const int t_begin = (int)((long long)L.T * blockIdx.y / p.nsplit);
    const int t_end = (int)((long long)L.T * (blockIdx.y + 1) / p.nsplit);
#pragma unroll 2
    for (int t = t_begin; t < t_end; ++t) {
        cplx x[M], y[KC];
#pragma unroll
        for (int c = 0; c < M; ++c) x[c] = ldg_x(xg + ((size_t)t * M + c) * OIVA_GROUP + lane);
        demix_frame<M, KC>(y, x, w);
        if (bin_ok) {
            XC* dst = Y + (((size_t)b * L.T + t) * L.F + f) * p.K + k0;
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (k0 + k < p.K) narrow(dst[k], y[k]);
        }
    }
}

// grid (G, nsplit), block = 32 * ceil(K/KC): writes Xr[gi][t][k][lane]
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(stream_max_threads(KC)) k_project_rows(const StreamParams p) {
    typedef typename StoreC<ST>::type XC;
    const GroupLayout& L = p.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gi = blockIdx.x;
    const long long b = gi / L.NG;
    const int g = (int)(gi - b * L.NG);
    const int f = g * OIVA_GROUP + lane;
    const bool bin_ok = f < L.F;
    const int k0 = warp * KC;
    cplx w[M][KC];
    load_filters<M, KC>(w, p, b * L.F + (bin_ok ? f : 0), gi, lane, bin_ok, k0);
    const XC* xg = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems();
    XC* xr = reinterpret_cast<XC*>(p.Xr) + (size_t)gi * L.T * p.K * OIVA_GROUP;
    const int t_begin = (int)((long long)L.T * blockIdx.y / p.nsplit);
    const int t_end = (int)((long long)L.T * (blockIdx.y + 1) / p.nsplit);
#pragma unroll 2
    for (int t = t_begin; t < t_end; ++t) {
        cplx x[M], y[KC];
#pragma unroll
        for (int c = 0; c < M; ++c) x[c] = ldg_x(xg + ((size_t)t * M + c) * OIVA_GROUP + lane);
        demix_frame<M, KC>(y, x, w);  // zero filters on padded bins => zero samples
#pragma unroll
        for (int k = 0; k < KC; ++k)
            if (k0 + k < p.K) narrow(xr[((size_t)t * p.K + k0 + k) * OIVA_GROUP + lane], y[k]);
    }
}

}  // namespace oiva
