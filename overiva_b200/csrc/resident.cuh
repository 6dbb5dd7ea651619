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
//     over the bin groups in k_source_model's order (few frames: by every CTA for its own mixture; many frames: pairs
//     dealt to the CTAs, r exchanged through L2 and a second barrier) -> gamma, phi and the W scale for the slice's
//     frames -> weighted covariance of the slice from shared memory (per-warp partial sums to an L2-resident scratch)
//     -> the LAST CTA of a bin group to arrive adds the partial sums in a fixed order and runs the group's IP sweep
//     with C, V_s and W_hat staged in shared memory so that the dependent chain never waits for L2 (K < M: thread per
//     bin, exactly the arithmetic of k_ip_update_tpb; K = M: a lane group per bin over all 8 warps, the arithmetic of
//     k_ip_update), then releases the group's epoch flag; the other CTAs of the group spin on it.
//   All cross-CTA traffic (statistic partials, covariance partials, W_hat) is a few hundred KB per epoch and stays in L2.
// Results are deterministic (fixed summation orders everywhere) and agree with the multi-kernel path to rounding
// (same statistic arithmetic; the covariance sums frames in different sub-ranges).
//
// STREAM variant (a mixture too long for shared memory, still few bin groups -- config 3: 60 s, M = 8, 122 MB): the same
// loop, but a slice is streamed through a ring of 16-frame stages by bulk TMA in both passes of every epoch.  The
// chunk sequence of a slice is known in advance (chunks 0..NC-1, twice per epoch), so the ring runs CONTINUOUSLY across
// the phases and epochs: while the CTAs sit in a grid barrier or wait for a sweep, the first stages of the next pass
// are already landing.
#pragma once
#include "cov.cuh"
#include "solve.cuh"
#include "solve_tpb.cuh"
#include "stream.cuh"

namespace oiva {

constexpr int RES_WARPS = 8;
constexpr int RES_THREADS = RES_WARPS * 32;
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
    int n_stages;     // STREAM kernels: stages of the sample ring (RES_CH frames each); resident kernels: 0
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
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all CTAs of the (cooperative, co-resident) grid; `target` = gridDim.x * (number of barriers so far)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// weighted covariance accumulators of one warp: the entries of one part, all K sources; same products and accumulation
// as CovPart::accumulate / finish (cov.cuh).  add(): frames [f0, f1) of a block of samples xs ([frame][M][32]) with the
// weights ph[k * pitch + frame].
template <typename ST, int M, int K, int P, int PART>
struct ResCovAcc {
    typedef typename StoreC<ST>::type XC;
    static constexpr int NE = oiva_tri(M), NEP = (NE + P - 1) / P;
    cplx acc[NEP][K];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int n = 0; n < NEP; ++n)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[n][k] = cmake(0.0, 0.0);
    }
    __device__ __forceinline__ void add(const XC* __restrict__ xs, const double* __restrict__ ph, int pitch, int f0, int f1,
                                        int lane) {
#pragma unroll 2
        for (int fr = f0; fr < f1; ++fr) {
            cplx x[M];
            double w[K];
#pragma unroll
            for (int c = 0; c < M; ++c) x[c] = widen(xs[((size_t)fr * M + c) * OIVA_GROUP + lane]);
#pragma unroll
            for (int k = 0; k < K; ++k) w[k] = ph[k * pitch + fr];
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
    }
    __device__ __forceinline__ void store(cplx* __restrict__ dst, double invT, int lane) const {
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

// STREAM kernels: the CTA's ring of sample stages.  One producer (thread 0) that never blocks and is polled whenever
// its warp waits for data or releases a stage; every warp consumes every chunk (wait -> use -> release).
constexpr int RES_CH = 16;  // frames per stage: 2 per warp in the statistic pass
template <typename XC>
struct ResRing {
    uint64_t* full;
    uint64_t* empty;
    unsigned char* stage0;
    const unsigned char* src;  // the slice in global memory
    size_t stage_bytes, frame_bytes;
    int n_stages, n_chunks, nfr;
    int cstage, cphase;                // consumer cursor (all threads, in step)
    long long pq, ptotal;              // producer cursor (thread 0)
    int pstage, puse, pc;
    __device__ __forceinline__ void try_issue() {
        while (pq < ptotal) {
            if (puse > 0 && !mbar_test(&empty[pstage], (puse - 1) & 1)) return;
            const int n = min(RES_CH, nfr - pc * RES_CH);
            const uint32_t bytes = (uint32_t)((size_t)n * frame_bytes);
            mbar_arrive_expect_tx(&full[pstage], bytes);
            tma_load_1d(stage0 + (size_t)pstage * stage_bytes, src + (size_t)pc * RES_CH * frame_bytes, bytes, &full[pstage]);
            ++pq;
            if (++pc == n_chunks) pc = 0;
            if (++pstage == n_stages) {
                pstage = 0;
                ++puse;
            }
        }
    }
    __device__ __forceinline__ const XC* wait() {
        if (threadIdx.x == 0) {
            while (!mbar_test(&full[cstage], cphase)) try_issue();
        }
        __syncwarp();
        mbar_wait(&full[cstage], cphase);
        return reinterpret_cast<const XC*>(stage0 + (size_t)cstage * stage_bytes);
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[cstage]);
        if (++cstage == n_stages) {
            cstage = 0;
            cphase ^= 1;
        }
        if (threadIdx.x == 0) try_issue();
    }
};

// shared-memory carve-up (host and device agree through this one function); offsets in bytes from the dynamic base
struct ResSmem {
    size_t x, phi, misc, c, v, w, total;
};
// n_stages > 0 (STREAM kernels): the sample region holds n_stages stages of RES_CH frames instead of the whole slice
__host__ __device__ inline ResSmem res_smem_layout(int M, int K, int slice_cap, int v_bufs, int elem_bytes, int n_stages = 0) {
    ResSmem s;
    const size_t mat = (size_t)oiva_tri(M) * OIVA_GROUP * sizeof(cplx);
    size_t o = 128;  // [0]: mbarrier of the slice load
    s.x = o;    o += (((size_t)(n_stages > 0 ? n_stages * 16 : slice_cap) * M * OIVA_GROUP * elem_bytes) + 127) / 128 * 128;
    s.phi = o;  o += (((size_t)K * slice_cap * sizeof(double)) + 127) / 128 * 128;
    s.misc = o; o += 256;  // gamma[8], wscale[8], flags
    s.c = o;    o += mat;
    s.v = o;    o += (size_t)v_bufs * mat;
    s.w = o;    o += (size_t)M * M * OIVA_GROUP * sizeof(cplx);
    s.total = o;
    return s;
}

// grid = G * SG CTAs of 256 threads (cooperative launch); CTA c owns slice s = c % SG of bin group gi = c / SG
template <typename ST, int M, int K, bool STREAM>
__global__ void __launch_bounds__(RES_THREADS, 1) k_loop_resident(const ResidentParams p) {
    typedef typename StoreC<ST>::type XC;
    typedef ResCfg<M, K> RC;
    constexpr int NE = RC::NE;
    constexpr uint32_t MAT_ELEMS = NE * OIVA_GROUP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ResSmem lay = res_smem_layout(M, K, p.slice_cap, p.v_bufs, (int)sizeof(XC), STREAM ? p.n_stages : 0);
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

    // ---- one-time: the slice's samples (bulk TMA; STREAM: the ring and its first stages) and the group's input covariance
    const XC* slice_src = reinterpret_cast<const XC*>(p.Xg) + (size_t)gi * L.group_elems() + (size_t)t0 * L.frame_elems();
    ResRing<XC> ring;
    if constexpr (STREAM) {
        ring.full = bar + 1;
        ring.empty = bar + 1 + p.n_stages;
        ring.stage0 = reinterpret_cast<unsigned char*>(sX);
        ring.src = reinterpret_cast<const unsigned char*>(slice_src);
        ring.frame_bytes = L.frame_elems() * sizeof(XC);
        ring.stage_bytes = (size_t)RES_CH * ring.frame_bytes;
        ring.n_stages = p.n_stages;
        ring.n_chunks = (nfr + RES_CH - 1) / RES_CH;
        ring.nfr = nfr;
        ring.cstage = 0;
        ring.cphase = 0;
        ring.pq = 0;
        ring.ptotal = 2ll * p.n_iter * ring.n_chunks;  // every epoch walks the slice twice (statistic, covariance)
        ring.pstage = 0;
        ring.puse = 0;
        ring.pc = 0;
        if (tid == 0) {
            for (int s2 = 0; s2 < p.n_stages; ++s2) {
                mbar_init(&ring.full[s2], 1);
                mbar_init(&ring.empty[s2], RES_WARPS);
            }
            mbar_fence_init();
        }
    } else if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const size_t bytes = (size_t)nfr * L.frame_elems() * sizeof(XC);
        mbar_arrive_expect_tx(bar, (uint32_t)bytes);
        const size_t piece = 32768;
        for (size_t o = 0; o < bytes; o += piece)
            tma_load_1d(reinterpret_cast<unsigned char*>(sX) + o, reinterpret_cast<const unsigned char*>(slice_src) + o,
                        (uint32_t)(bytes - o < piece ? bytes - o : piece), bar);
    }
    for (uint32_t i = tid; i < MAT_ELEMS; i += RES_THREADS) sC[i] = p.Cg[(size_t)gi * MAT_ELEMS + i];
    __syncthreads();
    if constexpr (STREAM) {
        if (tid == 0) ring.try_issue();
    } else {
        mbar_wait(bar, 0);
    }

    unsigned n_bar = 0;
#pragma unroll 1
    for (int epoch = 0; epoch < p.n_iter; ++epoch) {
        // ---- (1) statistic of the slice: r2part[gi][k][t] = sum over the 32 bins |w_k^H x|^2      overiva.py:140,152-155
        {
            cplx w[M][K];
#pragma unroll
            for (int c = 0; c < M; ++c)
#pragma unroll
                for (int k = 0; k < K; ++k) w[c][k] = __ldcg(Wgrp + (size_t)(c * M + k) * OIVA_GROUP + lane);
            double* r2g = p.r2part + (size_t)gi * K * Tp;
            if constexpr (STREAM) {
                constexpr int FBW = RES_CH / RES_WARPS;  // frames per warp and chunk
                constexpr int LPF = OIVA_GROUP / FBW;
                for (int c = 0; c < ring.n_chunks; ++c) {
                    const XC* xs = ring.wait();
                    const int nfc = min(RES_CH, nfr - c * RES_CH);
                    const int j0 = warp * FBW;
                    double v[K][FBW];
#pragma unroll
                    for (int j = 0; j < FBW; ++j) {
                        if (j0 + j < nfc) {
                            cplx x[M], y[K];
#pragma unroll
                            for (int ch = 0; ch < M; ++ch) x[ch] = widen(xs[((size_t)(j0 + j) * M + ch) * OIVA_GROUP + lane]);
                            demix_frame<M, K>(y, x, w);
#pragma unroll
                            for (int k = 0; k < K; ++k) v[k][j] = fma(y[k].x, y[k].x, y[k].y * y[k].y);
                        } else {
#pragma unroll
                            for (int k = 0; k < K; ++k) v[k][j] = 0.0;
                        }
                    }
                    ring.release();
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double sres = lane_sum_frames<FBW>(v[k], lane);
                        const int j = j0 + lane / LPF;
                        if ((lane % LPF) == 0 && j < nfc) __stcg(r2g + (size_t)k * Tp + t0 + c * RES_CH + j, sres);
                    }
                }
            } else {
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
                    for (int k = 0; k < K; ++k) {  // the lane sum of k_demix_power (stream.cuh)
                        const double sres = lane_sum_frames<POWER_FB>(v[k], lane);
                        const int j = lane >> 2;
                        if ((lane & 3) == 0 && fb + j < nfr) __stcg(r2g + (size_t)k * Tp + t0 + fb + j, sres);
                    }
                }
            }
        }
        grid_barrier(bar_counter, (++n_bar) * gridDim.x);

        // ---- (2) r[b][k][t] = model(sum over the bin groups); the summation order is k_source_model's (8 interleaved
        //      slices, then a fixed tree)                                                           overiva.py:152-155
        auto model_fn = [&](double s) {
            switch (p.model) {
                case OIVA_MODEL_LAPLACE: return 2.0 * sqrt(s);
                case OIVA_MODEL_GAUSS: return s / (double)p.F_total;
                default: return 0.0;
            }
        };
        {
            // the (k, t) pairs are dealt round-robin to the CTAs, r goes through rbuf (L2) and a second grid barrier.
            // (Every CTA summing all partials of its mixture itself -- one barrier less -- measured slower at config 1:
            // 0.47 vs 0.41 ms per 20 epochs; the 130 CTAs re-read K * T * NG partials each, 8 of every 32-byte sector.)
            const long long n_pairs = (long long)p.B * K * T;
            const int sub = lane >> 3, cs = lane & 7;
            for (long long q0 = ((long long)blockIdx.x * RES_WARPS + warp) * 4; q0 < n_pairs;
                 q0 += (long long)gridDim.x * RES_WARPS * 4) {
                const long long q = q0 + sub;
                const bool ok = q < n_pairs;
                const long long qq = ok ? q : 0;
                const long long pb = qq / ((long long)K * T);
                const int rem = (int)(qq - pb * K * T);
                const int k = rem / T, t = rem - k * T;
                double sum = 0.0;
                if (ok) {
                    const double* src = p.r2part + ((size_t)pb * L.NG * K + k) * Tp + t;
#pragma unroll 4
                    for (int ch = cs; ch < L.NG; ch += 8) sum += __ldcg(src + (size_t)ch * K * Tp);
                }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                if (ok && cs == 0) __stcg(p.rbuf + ((size_t)pb * K + k) * Tp + t, model_fn(sum));
            }
            grid_barrier(bar_counter, (++n_bar) * gridDim.x);
        }

        // ---- (3) gamma = mean_t r, phi = 1 / max(r / gamma, 1e-15) for the slice's frames, W scale   overiva.py:158-173
        const double* rglob = p.rbuf + (size_t)b * K * Tp;
        auto r_at = [&](int k, int t) { return __ldcg(rglob + (size_t)k * Tp + t); };
        if (warp < K) {
            double lsum = 0.0;
            for (int tt = 0; tt < Tp; tt += 128) {  // 4 loads in flight, added in ascending order (+0.0 beyond T)
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = tt + 32 * j + lane;
                    v[j] = t < T ? r_at(warp, t) : 0.0;
                }
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

        // ---- (4) weighted covariance of the slice, per-warp partial sums to the L2 scratch            overiva.py:179
        {
            const int part = warp % RC::P, fw = warp / RC::P;
            const int slot = sl * RC::FW + fw;
            cplx* dst = p.Vpart + ((size_t)slot * p.G + gi) * grp_cov;
            static_for<RC::P>([&](auto pc_) {
                constexpr int PART = decltype(pc_)::value;
                if (part == PART) {
                    ResCovAcc<ST, M, K, RC::P, PART> acc;
                    acc.zero();
                    if constexpr (STREAM) {
                        for (int c = 0; c < ring.n_chunks; ++c) {
                            const XC* xs = ring.wait();
                            const int nfc = min(RES_CH, nfr - c * RES_CH);
                            const int f0 = min(nfc, RES_CH * fw / RC::FW), f1 = min(nfc, RES_CH * (fw + 1) / RC::FW);
                            acc.add(xs, sPhi + c * RES_CH, pitch, f0, f1, lane);
                            ring.release();
                        }
                    } else {
                        const int f0 = (int)((long long)nfr * fw / RC::FW), f1 = (int)((long long)nfr * (fw + 1) / RC::FW);
                        acc.add(sX, sPhi, pitch, f0, f1, lane);
                    }
                    acc.store(dst, p.invT, lane);
                }
            });
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned old = atomicAdd(arrive, 1u);
            sFlag[0] = (old == (unsigned)(p.SG * (epoch + 1) - 1)) ? 1 : 0;
        }
        __syncthreads();

        if (sFlag[0]) {
            // ---- (5) last CTA of the group: fixed-order sum of the partial covariances + the IP sweep   overiva.py:176-190
            __threadfence();
            const int n_slots = p.SG * RC::FW;
            // (the slots are added in ascending order, as k_cov_sum_partials does; the loads of 8 slots are issued
            // before the first add -- one L2 latency per 8 slots instead of one per slot: ncu showed the other CTAs of
            // the group waiting 30 % of the epoch for this loop when it was a plain load-add chain)
            auto reduce_source = [&](int s, cplx* out, int first_thread, int n_threads) {
                for (uint32_t i = tid - first_thread; i < MAT_ELEMS; i += n_threads) {
                    const cplx* src = p.Vpart + (size_t)gi * grp_cov + (size_t)s * MAT_ELEMS + i;
                    const size_t slot_stride = (size_t)p.G * grp_cov;
                    cplx acc = cmake(0.0, 0.0);
                    for (int sp0 = 0; sp0 < n_slots; sp0 += 8) {
                        cplx v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            v[j] = sp0 + j < n_slots ? __ldcg(src + (size_t)(sp0 + j) * slot_stride) : cmake(0.0, 0.0);
                        if (sp0 == 0) {
                            acc = v[0];
#pragma unroll
                            for (int j = 1; j < 8; ++j)
                                if (j < n_slots) {
                                    acc.x += v[j].x;
                                    acc.y += v[j].y;
                                }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (sp0 + j < n_slots) {
                                    acc.x += v[j].x;
                                    acc.y += v[j].y;
                                }
                        }
                    }
                    out[i] = acc;
                }
            };
            for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS) sW[i] = __ldcg(Wgrp + i);
            reduce_source(0, sV, 0, RES_THREADS);
            __syncthreads();
            if constexpr (K == M && M >= 3) {
                // determined case: the in-thread LU of the thread-per-bin sweep is one long dependent chain per source
                // (config 2: ~6 us x 6 sources with 31 warps of the GPU idle).  Here all 8 warps take part: a group
                // of G = next_pow2(M) lanes owns a bin, one row of [W_hat^H V_s | e_s] per lane, Gauss-Jordan with
                // shuffle pivoting (solve.cuh, the arithmetic of k_ip_update) -- ~M times shorter chains.
                constexpr int GL = Grp<M>::G, BPW = Grp<M>::BINS;
                int singular = 0;
                for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS)  // W *= wscale   overiva.py:161-167
                    sW[i] = cscale(sW[i], sWs[(i / OIVA_GROUP) % M]);
                __syncthreads();
#pragma unroll 1
                for (int s = 0; s < K; ++s) {
                    for (int bin0 = warp * BPW; bin0 < OIVA_GROUP; bin0 += RES_WARPS * BPW) {
                        const int l = bin0 + lane / GL, gl = lane % GL;
                        const bool rv = gl < M;
                        const bool lv = g * OIVA_GROUP + l < L.F;  // padded bins keep their zero W_hat (nothing is written)
                        auto Wat = [&](int r, int c) -> cplx& { return sW[(size_t)(r * M + c) * OIVA_GROUP + l]; };
                        auto Vat = [&](int r, int c) { return herm_load<false>(sV + l, r, c); };
                        cplx A[M + 1];
#pragma unroll
                        for (int c = 0; c <= M; ++c) A[c] = cmake(0.0, 0.0);
                        if (rv) {
                            for (int j = 0; j < M; ++j) {
                                const cplx a = Wat(j, gl);
#pragma unroll
                                for (int c = 0; c < M; ++c) cfmac(A[c], a, Vat(j, c));
                            }
                            if (gl == s) A[M] = cmake(1.0, 0.0);
                        }
                        const int col = gauss_jordan<M, GL>(A, M, gl, lane, rv, singular);
                        __syncwarp();
                        if (col >= 0 && lv) Wat(col, s) = A[M];
                        __syncwarp();
                        cplx wi = cmake(0.0, 0.0), u = cmake(0.0, 0.0);  // w_s /= sqrt(w_s^H V_s w_s)    overiva.py:185-186
                        if (rv) {
                            wi = Wat(gl, s);
#pragma unroll
                            for (int j = 0; j < M; ++j) cfma(u, Vat(gl, j), Wat(j, s));
                        }
                        const cplx d = group_sum<GL>(cmulc(wi, u));
                        const cplx inv = crecip(csqrt_(d));
                        __syncwarp();
                        if (rv && lv) Wat(gl, s) = cmul(wi, inv);
                        __syncwarp();
                        if (singular && lv) atomicOr(p.status + b, OIVA_STATUS_SINGULAR);
                        singular = 0;
                    }
                    __syncthreads();
                    if (s + 1 < K) {
                        reduce_source(s + 1, sV, 0, RES_THREADS);
                        __syncthreads();
                    }
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
                }
                if (warp == 0 && bin_ok) {
                    const bool bad = ip_sweep_nonfinite<M, K>(Wm);
                    if (singular || bad)
                        atomicOr(p.status + b, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
                }
            }
            __syncthreads();
            for (uint32_t i = tid; i < (uint32_t)(M * M * OIVA_GROUP); i += RES_THREADS) __stcg(Wgrp + i, sW[i]);
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                st_release_u32(flag, (unsigned)(epoch + 1));
            }
        } else {
            if (tid == 0) {
                while (ld_acquire_u32(flag) < (unsigned)(epoch + 1)) {
                }
                __threadfence();
            }
            __syncthreads();
        }
    }
}

}  // namespace oiva
