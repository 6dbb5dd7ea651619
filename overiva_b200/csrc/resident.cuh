// The whole epoch loop of a SHORT input in ONE persistent cooperative launch (configs 1-2 of BASELINE.json: a single
// 15 s mixture; the call the reference times at overiva_oneshot.py:298-368).
//
// For one short mixture the streaming kernels are latency-bound: 65 bin groups cannot fill 148 SMs, and every epoch is
// a chain of 5 dependent kernels (demix-power -> source model -> covariance -> partial sums -> IP sweep) whose launch
// gaps cost more than their work (29 us per epoch at config 1, of which the HBM floor is 4 us).  Here the grid stays
// resident for all n_iter epochs (overiva.py:138-190):
//   * every CTA owns one SLICE (bin group gi, frame range) and keeps its samples in SHARED MEMORY for the whole loop
//     (one bulk-TMA copy at start; config 1: 15 MB over 130 SMs) -- X is never read from L2 / HBM again;
//   * epoch = statistic of the slice (lane <-> bin, the same butterfly as k_demix_power) -> grid barrier -> the sum
//     over the bin groups in k_source_model's order ((k, t) pairs dealt to the CTAs, r exchanged through L2 and a
//     second barrier) -> gamma, phi and the W scale for the slice's frames -> weighted covariance of the slice from shared memory (per-warp partial sums to an L2-resident scratch)
//     -> the LAST CTA of a bin group to arrive adds the partial sums in a fixed order and runs the group's IP sweep
//     with C, V_s and W_hat staged in shared memory so that the dependent chain never waits for L2 (K < M: thread per
//     bin, exactly the arithmetic of k_ip_update_tpb; K = M: two lanes per bin, the arithmetic of k_ip_update_pair),
//     then releases the group's epoch flag; the other CTAs of the group spin on it.
//   All cross-CTA traffic (statistic partials, covariance partials, W_hat) is a few hundred KB per epoch and stays in L2.
// Results are deterministic (fixed summation orders everywhere) and agree with the multi-kernel path to rounding
// (same statistic arithmetic; the covariance sums frames in different sub-ranges).
#pragma once
#include "cov.cuh"
#include "solve.cuh"
#include "solve_pair.cuh"
#include "solve_tpb.cuh"
#include "stream.cuh"

namespace oiva {

constexpr int RES_WARPS = 8;
constexpr int RES_THREADS = RES_WARPS * 32;
constexpr int RES_TRACE_POINTS = 14;  // stamps per CTA and epoch when ResidentParams::trace is set
constexpr int RES_TAG_MAX_GROUPS = 160;    // bin groups per mixture the tagged statistic sum holds in registers (5 per lane)
constexpr unsigned RES_SPIN_LIMIT = 1u << 21;  // polls (~1 us each) before a waiter gives up and reports OIVA_STATUS_STALLED
constexpr int RES_SYNC_HEADER = 8;  // sync[0]: grid-barrier counter; [8 + gi]: arrivals of group gi; [8 + G + gi]: its flag

struct ResidentParams {
    const void* Xg;   // grouped samples
    cplx* Wg;         // grouped W_hat [gi][M*M][32], updated in place
    const cplx* Cg;   // grouped input covariance [gi][NE][32]
    double* r2part;   // (G, K, Tp) per-group statistic
    double* rbuf;     // (B, K, Tp) r = model(sum over groups)
    cplx* Vpart;      // (NSLOT, G, K, NE, 32) partial covariances, NSLOT = SG * FW
    unsigned* sync;   // RES_SYNC_HEADER + 2 G words, zeroed before the launch
    int* status;      // one word per mixture
    GroupLayout L;
    long long G;
    int B, SG, n_iter, model, F_total;
    int slice_cap;    // frames a slice can hold (= max over slices)
    int v_bufs;       // 1 or 2 shared-memory buffers for the reduced V_s
    int cluster;      // 1: the SG CTAs of a bin group form a thread-block cluster and hand over through cluster barriers
    long long* trace; // null, or (grid, n_iter, RES_TRACE_POINTS) clock64 stamps of thread 0 (OIVA_RES_TRACE, profiling only)
    int tracked;      // 1: determined sweep with the inverse of W_hat carried along (one Cholesky per source, no LU)
    int tagged;       // 1: no grid barrier inside the epoch -- the statistic words carry the epoch's parity in their sign bit
    int poll;         // how waiters spin: 0 acquire loads, 1 relaxed loads + one fence, 2 relaxed loads + one acquire load
    double invT;
};

// how the covariance accumulators of a slice are spread over the 8 warps: P entry parts x FW frame ranges
template <int M, int K>
struct ResCfg {
    static constexpr int NE = oiva_tri(M);
    static constexpr int P = (NE * K <= 72) ? 1 : (NE * K <= 144) ? 2 : (NE * K <= 288) ? 4 : 8;
    static constexpr int FW = RES_WARPS / P;
    static constexpr int NEP = (NE + P - 1) / P;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// spin until *p >= target. An acquire load invalidates L1 on every iteration; the relaxed modes pay that once.
__device__ __forceinline__ void spin_until_ge(const unsigned* p, unsigned target, int mode) {
    if (mode == 0) {
        while (ld_acquire_u32(p) < target) {
        }
    } else {
        while (ld_relaxed_u32(p) < target) {
        }
        if (mode == 1) fence_acquire_gpu();
        else (void)ld_acquire_u32(p);
    }
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_acq_rel_add_u32(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
#define OIVA_RES_STAMP(n)                                                                                    \
    do {                                                                                                     \
        if (p.trace) {                                                                                       \
            __syncthreads();                                                                                 \
            if (threadIdx.x == 0) p.trace[((size_t)blockIdx.x * p.n_iter + epoch) * RES_TRACE_POINTS + (n)] = clock64(); \
        }                                                                                                    \
    } while (0)
// ---- statistic words that announce themselves ----------------------------------------------------------------------
// r2part and r are sums of squares / norms: never negative, so the sign bit is free.  A producer stores the value of epoch
// e with sign bit e & 1 (one 8-byte relaxed store: value and flag cannot be seen apart); a consumer polls the word itself
// until the sign is the epoch's and strips it.  That replaces "store, release, count, spin on the counter, load" -- a grid
// barrier, ~2700 cycles twice per epoch -- by the store and the load.  Nobody can lap anybody: a producer reaches epoch
// e + 1 only through phase (3) of epoch e, which needs every r of its mixture, each published only after its summing warp
// has read every r2part word it covers.
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// (In PTX on purpose: written as C++ bit operations the compiler recognises fabs(), emits abs.f64 -- an arithmetic DADD in
// SASS, which turns a NaN into the canonical NaN, sign bit SET -- and a NaN statistic of a singular mixture would carry
// the wrong parity for ever.  The integer instructions keep the NaN a NaN and make its sign ours.)
__device__ __forceinline__ double tag_word(double v, unsigned parity) {
    double r;
    asm volatile(
        "{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tand.b32 hi, hi, 0x7fffffff;\n\tor.b32 hi, hi, %2;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "=d"(r)
        : "d"(v), "r"(parity << 31));
    return r;
}
__device__ __forceinline__ bool tag_is(double v, unsigned parity) {
    return ((unsigned)__double2hiint(v) >> 31) == parity;
}
__device__ __forceinline__ void named_barrier(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
// all threads of the cluster; release/acquire at cluster scope orders the partial sums and W (st.cg / ld.cg, L2) between
// the CTAs of a bin group without a round trip through a flag in global memory
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the complex number at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ cplx ld_dsmem_c(const cplx* local, unsigned rank) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t ra;
    cplx v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ra) : "memory");
    return v;
}
// all CTAs of the (cooperative, co-resident) grid; `target` = gridDim.x * (number of barriers so far).
// One release-add and an acquire spin by thread 0 between two CTA barriers: the CTA barrier orders the other threads'
// writes before thread 0's release and thread 0's acquire before their later reads (causality order is cumulative), so
// no separate __threadfence() is needed (the first version had two per barrier: MEMBAR.SC.GPU at 3.3 us per barrier).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target, int mode) {
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_add_u32(counter, 1u);
        spin_until_ge(counter, target, mode);
    }
    __syncthreads();
}

// weighted covariance of frames [f0, f1) of the slice for the entries of one part, all K sources; same products and
// accumulation as CovPart::accumulate / finish (cov.cuh)
template <typename ST, int M, int K, int P, int PART>
struct ResCovPart {
    typedef typename StoreC<ST>::type XC;
    static constexpr int NE = oiva_tri(M), NEP = (NE + P - 1) / P;
    __device__ static __forceinline__ void run(const XC* __restrict__ sX, const double* __restrict__ sPhi, int pitch, int f0,
                                               int f1, cplx* __restrict__ dst, double invT, int lane) {
        cplx acc[NEP][K];
#pragma unroll
        for (int n = 0; n < NEP; ++n)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[n][k] = cmake(0.0, 0.0);
#pragma unroll 2
        for (int fr = f0; fr < f1; ++fr) {
            cplx x[M];
            double w[K];
#pragma unroll
            for (int c = 0; c < M; ++c) x[c] = widen(sX[((size_t)fr * M + c) * OIVA_GROUP + lane]);
#pragma unroll
            for (int k = 0; k < K; ++k) w[k] = sPhi[k * pitch + fr];
            static_for<NEP>([&](auto nc) {
                constexpr int n = decltype(nc)::value;
                constexpr int e = PART + n * P;
                if constexpr (e < NE) {
                    constexpr int i = ent_row(e), j = ent_col(e);
                    if constexpr (i == j) {
                        const double pr = fma(x[i].x, x[i].x, x[i].y * x[i].y);
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[n][k].x = fma(w[k], pr, acc[n][k].x);
                    } else {
                        const double pr = fma(x[i].x, x[j].x, x[i].y * x[j].y);
                        const double pi = fma(x[i].y, x[j].x, -(x[i].x * x[j].y));
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            acc[n][k].x = fma(w[k], pr, acc[n][k].x);
                            acc[n][k].y = fma(w[k], pi, acc[n][k].y);
                        }
                    }
                }
            });
        }
        static_for<NEP>([&](auto nc) {
            constexpr int n = decltype(nc)::value;
            constexpr int e = PART + n * P;
            if constexpr (e < NE) {
                constexpr bool diag = ent_row(e) == ent_col(e);
#pragma unroll
                for (int k = 0; k < K; ++k)
                    __stcg(dst + ((size_t)k * NE + e) * OIVA_GROUP + lane,
                           cmake(acc[n][k].x * invT, diag ? 0.0 : acc[n][k].y * invT));
            }
        });
    }
};
template <typename ST, int M, int K, int P, int PART = 0>
struct ResCovDispatch {
    typedef typename StoreC<ST>::type XC;
    __device__ static __forceinline__ void run(int part, const XC* sX, const double* sPhi, int pitch, int f0, int f1, cplx* dst,
                                               double invT, int lane) {
        if (part == PART) ResCovPart<ST, M, K, P, PART>::run(sX, sPhi, pitch, f0, f1, dst, invT, lane);
        else if constexpr (PART + 1 < P) ResCovDispatch<ST, M, K, P, PART + 1>::run(part, sX, sPhi, pitch, f0, f1, dst, invT, lane);
    }
};

// shared-memory carve-up (host and device agree through this one function); offsets in bytes from the dynamic base
struct ResSmem {
    size_t x, phi, misc, c, v, w, total;
};
__host__ __device__ inline ResSmem res_smem_layout(int M, int K, int slice_cap, int v_bufs, int elem_bytes) {
    ResSmem s;
    const size_t mat = (size_t)oiva_tri(M) * OIVA_GROUP * sizeof(cplx);
    size_t o = 128;  // [0]: mbarrier of the slice load
    s.x = o;    o += (((size_t)slice_cap * M * OIVA_GROUP * elem_bytes) + 127) / 128 * 128;
    s.phi = o;  o += (((size_t)K * slice_cap * sizeof(double)) + 127) / 128 * 128;
    s.misc = o; o += 256;  // gamma[8], wscale[8], flags
    s.c = o;    o += mat;
    s.v = o;    o += (size_t)v_bufs * mat;
    s.w = o;    o += (size_t)M * M * OIVA_GROUP * sizeof(cplx);
    s.total = o;
    return s;
}

// grid = G * SG CTAs of 256 threads (cooperative launch); CTA c owns slice s = c % SG of bin group gi = c / SG
template <typename ST, int M, int K, bool TRACKED = false>
__global__ void __launch_bounds__(RES_THREADS, 1) k_loop_resident(const ResidentParams p) {
    static_assert(!TRACKED || (K == M && M >= 3), "the tracked-inverse sweep is the determined one");
    typedef typename StoreC<ST>::type XC;
    typedef ResCfg<M, K> RC;
    constexpr int NE = RC::NE;
    constexpr uint32_t MAT_ELEMS = NE * OIVA_GROUP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ResSmem lay = res_smem_layout(M, K, p.slice_cap, p.v_bufs, (int)sizeof(XC));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    XC* sX = reinterpret_cast<XC*>(smem_raw + lay.x);
    double* sPhi = reinterpret_cast<double*>(smem_raw + lay.phi);
    double* sGamma = reinterpret_cast<double*>(smem_raw + lay.misc);
    double* sWs = sGamma + 8;
    int* sFlag = reinterpret_cast<int*>(sWs + 8);
    cplx* sC = reinterpret_cast<cplx*>(smem_raw + lay.c);
    cplx* sV = reinterpret_cast<cplx*>(smem_raw + lay.v);
    cplx* sW = reinterpret_cast<cplx*>(smem_raw + lay.w);

    const GroupLayout& L = p.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long gi = blockIdx.x / p.SG;
    const int sl = (int)(blockIdx.x - gi * p.SG);
    const long long b = gi / L.NG;
    const int g = (int)(gi - b * L.NG);
    const int T = L.T, Tp = L.frame_pitch();
    const int t0 = (int)((long long)T * sl / p.SG), t1 = (int)((long long)T * (sl + 1) / p.SG);
    const int nfr = t1 - t0;
    const int pitch = p.slice_cap;
    const bool bin_ok = g * OIVA_GROUP + lane < L.F;
    unsigned* bar_counter = p.sync;
    unsigned* arrive = p.sync + RES_SYNC_HEADER + gi;
    unsigned* flag = p.sync + RES_SYNC_HEADER + p.G + gi;
    cplx* Wgrp = p.Wg + (size_t)gi * M * M * OIVA_GROUP;
    const size_t grp_cov = (size_t)K * NE * OIVA_GROUP;

    // ---- one-time: the slice's samples (bulk TMA) and the group's input covariance into shared memory ---------------
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const XC* src = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems() + (size_t)t0 * L.frame_elems();
        const size_t bytes = (size_t)nfr * L.frame_elems() * sizeof(XC);
        mbar_arrive_expect_tx(bar, (uint32_t)bytes);
        const size_t piece = 32768;
        for (size_t o = 0; o < bytes; o += piece)
            tma_load_1d(reinterpret_cast<unsigned char*>(sX) + o, reinterpret_cast<const unsigned char*>(src) + o,
                        (uint32_t)(bytes - o < piece ? bytes - o : piece), bar);
    }
    for (uint32_t i = tid; i < MAT_ELEMS; i += RES_THREADS) sC[i] = p.Cg[(size_t)gi * MAT_ELEMS + i];
    __syncthreads();
    mbar_wait(bar, 0);

    unsigned n_bar = 0;
    bool stalled = false;
    // the (k, t) pairs of the statistic sum are dealt to the warps of the grid, CTA-first so that a short list spreads over
    // all SMs; lanes per pair: 8 (four pairs per warp), or the whole warp when a mixture has more than 16 bin groups (one
    // L2 latency instead of three for the 65 groups of a 2049-bin mixture)
    const long long n_pairs = (long long)p.B * K * T;
    const int lp = L.NG > 16 ? 32 : 8, ppw = 32 / lp;
    const int pair_sub = lane / lp, pair_cs = lane % lp;
    const long long pair_q0 = ((long long)warp * gridDim.x + blockIdx.x) * ppw + pair_sub;
    const long long pair_step = (long long)gridDim.x * RES_WARPS * ppw;
    if (p.tagged) {  // every word this CTA will produce starts as "epoch -1" (sign set), then ONE grid barrier per launch
        double* r2g = p.r2part + (size_t)gi * K * Tp;
        for (int i = tid; i < K * nfr; i += RES_THREADS) {
            const int k = i / nfr, fr = i - k * nfr;
            st_relaxed_f64(r2g + (size_t)k * Tp + t0 + fr, -1.0);
        }
        if (pair_cs == 0)
            for (long long q = pair_q0; q < n_pairs; q += pair_step) {
                const long long pb = q / ((long long)K * T);
                const int rem = (int)(q - pb * K * T);
                st_relaxed_f64(p.rbuf + ((size_t)pb * K + rem / T) * Tp + rem % T, -1.0);
            }
        grid_barrier(bar_counter, (++n_bar) * gridDim.x, p.poll);
    }
#pragma unroll 1
    for (int epoch = 0; epoch < p.n_iter; ++epoch) {
        const unsigned par = (unsigned)epoch & 1u;
        OIVA_RES_STAMP(0);
        // ---- (0) this epoch's W_hat in shared memory.  In a cluster it never leaves the chip between epochs: the first CTA
        //      sweeps in place and keeps it, the others copy it out of that CTA's shared memory after the hand-over
        //      barrier; global memory gets the last epoch's only.
        const bool w_via_l2 = !p.cluster && p.SG > 1;  // (flag hand-over: the sweeping CTA changes from epoch to epoch)
        if (w_via_l2 || epoch == 0) {
            for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS) sW[i] = __ldcg(Wgrp + i);
        } else if (p.cluster && sl != 0) {
            for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS) sW[i] = ld_dsmem_c(sW + i, 0u);
        }
        __syncthreads();
        // ---- (1) statistic of the slice: r2part[gi][k][t] = sum over the 32 bins |w_k^H x|^2      overiva.py:140,152-155
        {
            cplx w[M][K];
#pragma unroll
            for (int c = 0; c < M; ++c)
#pragma unroll
                for (int k = 0; k < K; ++k) w[c][k] = sW[(c * M + k) * OIVA_GROUP + lane];
            double* r2g = p.r2part + (size_t)gi * K * Tp;
            for (int blk = warp; blk * POWER_FB < nfr; blk += RES_WARPS) {
                const int fb = blk * POWER_FB;
                double v[K][POWER_FB];
#pragma unroll
                for (int j = 0; j < POWER_FB; ++j) {
                    if (fb + j < nfr) {
                        cplx x[M], y[K];
#pragma unroll
                        for (int c = 0; c < M; ++c) x[c] = widen(sX[((size_t)(fb + j) * M + c) * OIVA_GROUP + lane]);
                        demix_frame<M, K>(y, x, w);
#pragma unroll
                        for (int k = 0; k < K; ++k) v[k][j] = fma(y[k].x, y[k].x, y[k].y * y[k].y);
                    } else {
#pragma unroll
                        for (int k = 0; k < K; ++k) v[k][j] = 0.0;
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {  // the transposing butterfly of k_demix_power (stream.cuh)
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
                    const int j = lane >> 2;
                    if ((lane & 3) == 0 && fb + j < nfr) {
                        if (p.tagged) st_relaxed_f64(r2g + (size_t)k * Tp + t0 + fb + j, tag_word(s, par));
                        else __stcg(r2g + (size_t)k * Tp + t0 + fb + j, s);
                    }
                }
            }
        }
        OIVA_RES_STAMP(1);
        if (!p.tagged) grid_barrier(bar_counter, (++n_bar) * gridDim.x, p.poll);
        OIVA_RES_STAMP(2);

        // ---- (2) r[b][k][t] = model(sum over the bin groups): interleaved slices of the groups per lane, then a fixed
        //      tree -- deterministic; k_source_model's order when a mixture has at most 16 groups   overiva.py:152-155
        auto model_fn = [&](double s) {
            switch (p.model) {
                case OIVA_MODEL_LAPLACE: return 2.0 * sqrt(s);
                case OIVA_MODEL_GAUSS: return s / (double)p.F_total;
                default: return 0.0;
            }
        };
        {
            // the (k, t) pairs are dealt round-robin to the CTAs, r goes through rbuf (L2) and a second grid barrier.
            // (Every CTA summing the partials itself instead -- one grid barrier less -- measured no faster, twice: all
            // frames per CTA, 8 of every 32-byte sector used: 0.47 vs 0.41 ms per 20 epochs of config 1; only the CTA's own
            // frames, coalesced, gamma sums exchanged through the cluster's shared memory: 5400 cycles for the sums where
            // this phase and its barrier take 2800 + 2700, and 17400 at config 2 -- 130 CTAs pull the same lines out of L2.)
            for (long long q = pair_q0 - pair_sub; q < n_pairs; q += pair_step) {  // (whole warps: the shuffles below)
                const long long qs = q + pair_sub;
                const bool ok = qs < n_pairs;
                const long long qq = ok ? qs : 0;
                const long long pb = qq / ((long long)K * T);
                const int rem = (int)(qq - pb * K * T);
                const int k = rem / T, t = rem - k * T;
                const double* src = p.r2part + ((size_t)pb * L.NG * K + k) * Tp + t;
                double sum = 0.0;
                if (p.tagged) {
                    constexpr int NV = RES_TAG_MAX_GROUPS / 32;
                    double v[NV];
                    unsigned spins = 0;
                    for (;;) {  // all of the lane's words in flight together, again until each carries this epoch's sign
                        bool all = true;
#pragma unroll
                        for (int j = 0; j < NV; ++j) {
                            const int ch = pair_cs + j * lp;
                            if (ok && ch < L.NG) {
                                v[j] = ld_relaxed_f64(src + (size_t)ch * K * Tp);
                                all = all && tag_is(v[j], par);
                            } else {
                                v[j] = 0.0;
                            }
                        }
                        if (all || stalled) break;
                        if (++spins > RES_SPIN_LIMIT) stalled = true;
                    }
#pragma unroll
                    for (int j = 0; j < NV; ++j) sum += fabs(v[j]);
                    __syncwarp();
                } else if (ok) {
#pragma unroll 4
                    for (int ch = pair_cs; ch < L.NG; ch += lp) sum += __ldcg(src + (size_t)ch * K * Tp);
                }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                if (lp == 32) {
                    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                }
                if (ok && pair_cs == 0) {
                    double* dst = p.rbuf + ((size_t)pb * K + k) * Tp + t;
                    if (p.tagged) st_relaxed_f64(dst, tag_word(model_fn(sum), par));
                    else __stcg(dst, model_fn(sum));
                }
            }
            OIVA_RES_STAMP(3);
            if (!p.tagged) grid_barrier(bar_counter, (++n_bar) * gridDim.x, p.poll);
            OIVA_RES_STAMP(4);
        }

        // ---- (3) gamma = mean_t r, phi = 1 / max(r / gamma, 1e-15) for the slice's frames, W scale   overiva.py:158-173
        const double* rglob = p.rbuf + (size_t)b * K * Tp;
        // (tagged words: warp k waits for row k here, everybody else reads it after the CTA barrier below)
        auto r_at = [&](int k, int t) {
            return p.tagged ? fabs(ld_relaxed_f64(rglob + (size_t)k * Tp + t)) : __ldcg(rglob + (size_t)k * Tp + t);
        };
        if (warp < K) {
            double lsum = 0.0;
            for (int tt = 0; tt < Tp; tt += 128) {  // 4 loads in flight, added in ascending order (+0.0 beyond T)
                double v[4];
                unsigned spins = 0;
                for (;;) {
                    bool all = true;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int t = tt + 32 * j + lane;
                        if (t < T) {
                            if (p.tagged) {
                                v[j] = ld_relaxed_f64(rglob + (size_t)warp * Tp + t);
                                all = all && tag_is(v[j], par);
                                v[j] = fabs(v[j]);
                            } else {
                                v[j] = __ldcg(rglob + (size_t)warp * Tp + t);
                            }
                        } else {
                            v[j] = 0.0;
                        }
                    }
                    if (all || stalled) break;
                    if (++spins > RES_SPIN_LIMIT) stalled = true;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) lsum += v[j];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
            if (lane == 0) {
                const double gamma = lsum / (double)T;
                sGamma[warp] = gamma;
                sWs[warp] = p.model == OIVA_MODEL_LAPLACE ? 1.0 / gamma : (p.model == OIVA_MODEL_GAUSS ? 1.0 / sqrt(gamma) : 1.0);
            }
        }
        __syncthreads();
        for (int i = tid; i < K * nfr; i += RES_THREADS) {
            const int k = i / nfr, fr = i - k * nfr;
            double r = r_at(k, t0 + fr) / sGamma[k];  // 0/0 -> NaN as in numpy
            if (r < 1e-15) r = 1e-15;                // NaN stays NaN
            sPhi[k * pitch + fr] = 1.0 / r;
        }
        __syncthreads();
        OIVA_RES_STAMP(5);

        // ---- (4) weighted covariance of the slice, per-warp partial sums to the L2 scratch            overiva.py:179
        {
            const int part = warp % RC::P, fw = warp / RC::P;
            const int f0 = (int)((long long)nfr * fw / RC::FW), f1 = (int)((long long)nfr * (fw + 1) / RC::FW);
            const int slot = sl * RC::FW + fw;
            cplx* dst = p.Vpart + ((size_t)slot * p.G + gi) * grp_cov;
            ResCovDispatch<ST, M, K, RC::P>::run(part, sX, sPhi, pitch, f0, f1, dst, p.invT, lane);
        }
        bool owner;
        OIVA_RES_STAMP(6);
        if (p.cluster) {  // (the first CTA of the cluster sums and sweeps)
            cluster_barrier();
            owner = sl == 0;
        } else {
            __syncthreads();
            if (tid == 0) {  // (acq_rel: publishes this CTA's partial sums, and the last arriver sees everybody else's)
                const unsigned old = atom_acq_rel_add_u32(arrive, 1u);
                sFlag[0] = (old == (unsigned)(p.SG * (epoch + 1) - 1)) ? 1 : 0;
            }
            __syncthreads();
            owner = sFlag[0] != 0;
        }

        OIVA_RES_STAMP(7);
        if (owner) {
            // ---- (5) last CTA of the group: fixed-order sum of the partial covariances + the IP sweep   overiva.py:176-190
            const int n_slots = p.SG * RC::FW;
            // (the slots are added in ascending order, as k_cov_sum_partials does; the loads of 8 slots are issued
            // before the first add -- one L2 latency per 8 slots instead of one per slot: ncu showed the other CTAs of
            // the group waiting 30 % of the epoch for this loop when it was a plain load-add chain)
            // Two elements per thread and round: 16 loads in flight, so the 320 elements of config 1 (16 slots) take
            // two L2 latencies instead of four and the 672 of config 2 (8 slots) two instead of three.
            auto reduce_source = [&](int s, cplx* out, int first_thread, int n_threads) {
                const size_t slot_stride = (size_t)p.G * grp_cov;
                for (uint32_t i = tid - first_thread; i < MAT_ELEMS; i += 2 * n_threads) {
                    const uint32_t i2 = i + n_threads;
                    const bool two = i2 < MAT_ELEMS;
                    const cplx* src = p.Vpart + (size_t)gi * grp_cov + (size_t)s * MAT_ELEMS + i;
                    const cplx* src2 = src + (two ? n_threads : 0);
                    cplx acc = cmake(0.0, 0.0), acc2 = cmake(0.0, 0.0);
                    for (int sp0 = 0; sp0 < n_slots; sp0 += 8) {
                        cplx v[8], v2[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const bool in = sp0 + j < n_slots;
                            v[j] = in ? __ldcg(src + (size_t)(sp0 + j) * slot_stride) : cmake(0.0, 0.0);
                            v2[j] = in && two ? __ldcg(src2 + (size_t)(sp0 + j) * slot_stride) : cmake(0.0, 0.0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (sp0 == 0 && j == 0) {  // (the first slot starts the sum: no 0.0 + x, which would turn -0.0 into +0.0)
                                acc = v[0];
                                acc2 = v2[0];
                            } else if (sp0 + j < n_slots) {
                                acc.x += v[j].x;
                                acc.y += v[j].y;
                                acc2.x += v2[j].x;
                                acc2.y += v2[j].y;
                            }
                        }
                    }
                    out[i] = acc;
                    if (two) out[i2] = acc2;
                }
            };
            if constexpr (!TRACKED) {
                reduce_source(0, sV, 0, RES_THREADS);  // (sW holds this epoch's W_hat since phase (0))
                __syncthreads();
            }
            OIVA_RES_STAMP(10);
            if constexpr (K == M && M >= 3) {
                // determined case: the in-thread LU of the thread-per-bin sweep is one long dependent chain per source
                // (config 2: ~6 us x 6 sources).  Two lanes per bin instead (solve_pair.cuh: half of the rows each in
                // registers, one exchange with the partner lane per pivot): warps 0 and 1 cover the 32 bins.  (Measured
                // equal to the lane-group Gauss-Jordan over all 8 warps this loop used first -- config 2: 1.258 vs 1.257 ms
                // per 20 epochs: fewer shuffles per pivot, three times the row-building work per lane -- and kept because
                // it shares its code with the batched kernel.)
                for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS)  // W *= wscale   overiva.py:161-167
                    if (g * OIVA_GROUP + (int)(i % OIVA_GROUP) < L.F)  // (padded bins keep their W: nothing renormalises it)
                        sW[i] = cscale(sW[i], sWs[(i / OIVA_GROUP) % M]);
                __syncthreads();
                bool singular = false;
                const int pl = (warp & 1) * 16 + (lane & 15), ph = lane >> 4;  // (warps 0, 1: bin and row half)
                const bool pok = g * OIVA_GROUP + pl < L.F;
                if constexpr (TRACKED) {
                    // w_s = (W^H V_s)^-1 e_s = V_s^-1 (W^-H e_s): W^-H e_s is the conjugate of row s of A = W^-1, and
                    // A follows the new column s of W by a rank-one update (A u = A w - e_s since A W = I):
                    //   v = A w,  A[s] /= v_s,  A[i] -= v_i A[s]  (i != s).
                    // So a source costs one Cholesky of V_s and O(M^2) instead of forming W^H V_s and an LU with
                    // pivoting; A is recomputed from W at the start of every epoch's sweep (no drift beyond M updates).
                    // Warp 1 keeps A in registers, warp 0 factors V_s while warp 1 is still updating A for the previous
                    // source; they exchange q and w through the (otherwise unused) shared-memory copy of C.
                    // One register array serves both roles -- warp 1: A (M x M, loop-carried); warp 0: the Cholesky
                    // factor of V_s in its first NE entries -- so the allocator does not keep A alive next to warp 0's
                    // working set (separate arrays spilled 1.7 KB per thread); every barrier is executed at ONE program
                    // point by all of its participants (barriers 1 and 2: warps 0 and 1; __syncthreads: the block).
                    // Needs both V buffers: the launcher falls back to the pair sweep otherwise.
                    static_assert(NE <= M * M, "the factor fits the shared register array");
                    cplx* mail = sC;  // [2 M][32]: q, then w
                    cplx st[M * M];
                    double dinv[M];
                    auto extract_q = [&](int s) {  // warp 1: q = conj(A[s][:]) into the mailbox
#pragma unroll
                        for (int j = 0; j < M; ++j) {
                            cplx a = st[j];
#pragma unroll
                            for (int i = 1; i < M; ++i) {
                                a.x = s == i ? st[i * M + j].x : a.x;
                                a.y = s == i ? st[i * M + j].y : a.y;
                            }
                            mail[j * OIVA_GROUP + lane] = cconj(a);
                        }
                    };
                    auto factor = [&](int s) {  // warp 0: Cholesky of V_s (buffer s & 1) in registers
                        const cplx* cur = sV + (size_t)(s & 1) * MAT_ELEMS;
#pragma unroll
                        for (int e = 0; e < NE; ++e) st[e] = cur[e * OIVA_GROUP + lane];
                        chol_factor<M>(st, dinv, singular);
                    };
                    if (warp == 1) {
#pragma unroll
                        for (int i = 0; i < M * M; ++i) st[i] = sW[i * OIVA_GROUP + lane];  // A = W (row-major j M + k)
                        invert_inplace<M>(st, singular);  // (while warps 2-7 add the partial sums of source 0)
                    } else {
#pragma unroll
                        for (int i = 0; i < M * M; ++i) st[i] = cmake(0.0, 0.0);
                        if (warp >= 2) reduce_source(0, sV, 64, RES_THREADS - 64);
                    }
#pragma unroll
                    for (int i = 0; i < M; ++i) dinv[i] = 0.0;
                    __syncthreads();
                    // Software pipeline over the sources: while warp 1 folds w_s into A and extracts q_{s+1}, warp 0
                    // already factors V_{s+1} and the other warps add the partial sums of source s + 2 (into the buffer
                    // V_s was factored from).  Per source the chain is solve -> max(update, Cholesky) instead of
                    // Cholesky -> solve -> update.
                    if (warp == 1) extract_q(0);
                    else if (warp == 0) factor(0);
                    else if (1 < K) reduce_source(1, sV + MAT_ELEMS, 64, RES_THREADS - 64);
#pragma unroll 1
                    for (int s = 0; s < K; ++s) {
                        if (warp < 2) {
                            named_barrier(1, 64);  // q_s is in the mailbox, V_s is factored
                            if (warp == 0) {
                                cplx q[M];
#pragma unroll
                                for (int j = 0; j < M; ++j) q[j] = mail[j * OIVA_GROUP + lane];
                                chol_solve_normalise<M>(st, dinv, q);
#pragma unroll
                                for (int j = 0; j < M; ++j) {
                                    if (bin_ok) sW[(j * M + s) * OIVA_GROUP + lane] = q[j];
                                    mail[(M + j) * OIVA_GROUP + lane] = q[j];
                                }
                            }
                            named_barrier(2, 64);  // w_s is in the mailbox
                        }
                        __syncthreads();  // V_{s+1} is summed; V_s's buffer is free
                        if (warp == 1) {
                            cplx v[M];  // v = A w (w read from the mailbox one component at a time)
#pragma unroll
                            for (int i = 0; i < M; ++i) v[i] = cmake(0.0, 0.0);
#pragma unroll
                            for (int j = 0; j < M; ++j) {
                                const cplx wj = mail[(M + j) * OIVA_GROUP + lane];
#pragma unroll
                                for (int i = 0; i < M; ++i) cfma(v[i], st[i * M + j], wj);
                            }
                            cplx vs = v[0];
#pragma unroll
                            for (int i = 1; i < M; ++i) {
                                vs.x = s == i ? v[i].x : vs.x;
                                vs.y = s == i ? v[i].y : vs.y;
                            }
                            if (!(fabs(vs.x) + fabs(vs.y) > 0.0)) singular = true;
                            const cplx vinv = crecip_fast(vs);
#pragma unroll
                            for (int j = 0; j < M; ++j) {  // column j of A: A[s][j] /= v_s, A[i][j] -= v_i A[s][j]
                                cplx a = st[j];
#pragma unroll
                                for (int i = 1; i < M; ++i) {
                                    a.x = s == i ? st[i * M + j].x : a.x;
                                    a.y = s == i ? st[i * M + j].y : a.y;
                                }
                                const cplx asj = cmul(a, vinv);
#pragma unroll
                                for (int i = 0; i < M; ++i) {
                                    cplx t = st[i * M + j];
                                    cfms(t, v[i], asj);
                                    st[i * M + j].x = s == i ? asj.x : t.x;
                                    st[i * M + j].y = s == i ? asj.y : t.y;
                                }
                            }
                            if (s + 1 < K) extract_q(s + 1);
                        } else if (warp == 0) {
                            if (s + 1 < K) factor(s + 1);
                        } else if (s + 2 < K) {
                            reduce_source(s + 2, sV + (size_t)(s & 1) * MAT_ELEMS, 64, RES_THREADS - 64);
                        }
                    }
                    __syncthreads();
                    if (warp < 2 && singular && bin_ok) atomicOr(p.status + b, OIVA_STATUS_SINGULAR);
                } else {
#pragma unroll 1
                for (int s = 0; s < K; ++s) {
                    // (with a second V buffer the six idle warps sum the next source's partial covariances meanwhile:
                    // config 2 spent 3200 of every 10900 cycles per source in that sum, profiles/r02_res_trace.jsonl)
                    const cplx* cur = sV + (size_t)(p.v_bufs == 2 ? (s & 1) : 0) * MAT_ELEMS;
                    if (warp < 2) PairSweep<M, false>::source(WLane{sW + pl}, cur + pl, s, ph, pok, singular);
                    else if (p.v_bufs == 2 && s + 1 < K)
                        reduce_source(s + 1, sV + (size_t)((s + 1) & 1) * MAT_ELEMS, 64, RES_THREADS - 64);
                    __syncthreads();
                    if (s < 3) OIVA_RES_STAMP(11 + s);
                    if (p.v_bufs == 1 && s + 1 < K) {
                        reduce_source(s + 1, sV, 0, RES_THREADS);
                        __syncthreads();
                    }
                }
                if (warp < 2 && singular && pok) atomicOr(p.status + b, OIVA_STATUS_SINGULAR);
                }
                bool bad = false;
                for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS)
                    if (g * OIVA_GROUP + (int)(i % OIVA_GROUP) < L.F && (!isfinite(sW[i].x) || !isfinite(sW[i].y))) bad = true;
                if (bad) atomicOr(p.status + b, OIVA_STATUS_NONFINITE);
            } else {
                bool singular = false;
                const WLane Wm = {sW + lane};
#pragma unroll 1
                for (int s = 0; s < K; ++s) {
                    cplx* cur = sV + (size_t)(p.v_bufs == 2 ? (s & 1) : 0) * MAT_ELEMS;
                    if (warp == 0) {
                        if (bin_ok) {
                            if (s == 0) ip_sweep_rescale<M, K>(Wm, sWs);
                            ip_sweep_source<M, K, false>(Wm, cur + lane, sC + lane, s, singular);
                        }
                    } else if (p.v_bufs == 2 && s + 1 < K) {
                        reduce_source(s + 1, sV + (size_t)((s + 1) & 1) * MAT_ELEMS, 32, RES_THREADS - 32);
                    }
                    __syncthreads();
                    if (p.v_bufs == 1 && s + 1 < K) {
                        reduce_source(s + 1, sV, 0, RES_THREADS);
                        __syncthreads();
                    }
                    if (s < 3) OIVA_RES_STAMP(11 + s);
                }
                if (warp == 0 && bin_ok) {
                    const bool bad = ip_sweep_nonfinite<M, K>(Wm);
                    if (singular || bad)
                        atomicOr(p.status + b, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
                }
            }
            __syncthreads();
            if (w_via_l2 || epoch == p.n_iter - 1)
                for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS) __stcg(Wgrp + i, sW[i]);
            if (!p.cluster) {
                __syncthreads();
                if (tid == 0) st_release_u32(flag, (unsigned)(epoch + 1));
            }
        } else if (!p.cluster) {
            if (tid == 0) spin_until_ge(flag, (unsigned)(epoch + 1), p.poll);
            __syncthreads();
        }
        OIVA_RES_STAMP(8);
        if (p.cluster) cluster_barrier();
        OIVA_RES_STAMP(9);
    }
    if (stalled) atomicOr(p.status + b, OIVA_STATUS_STALLED);
}

}  // namespace oiva
