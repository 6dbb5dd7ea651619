// C entry points of the streaming kernels and the source-model finalisation
// (include/overiva_b200.h: oiva_demix_power, oiva_sum_partials, oiva_source_model, oiva_demix_output,
//  oiva_project_rows).
#include "stream.cuh"

namespace oiva {

#define OIVA_DECL(M) int stream_launch_m##M(int kind, int dtype, const StreamParams& p, long long G, cudaStream_t st);
OIVA_DECL(1) OIVA_DECL(2) OIVA_DECL(3) OIVA_DECL(4) OIVA_DECL(5) OIVA_DECL(6) OIVA_DECL(7) OIVA_DECL(8)
OIVA_DECL(9) OIVA_DECL(10) OIVA_DECL(11) OIVA_DECL(12) OIVA_DECL(13) OIVA_DECL(14) OIVA_DECL(15) OIVA_DECL(16)
#undef OIVA_DECL

enum { KIND_POWER = 0, KIND_OUTPUT = 1, KIND_PROJECT = 2 };

static int stream_launch(int M, int kind, int dtype, const StreamParams& p, long long G, cudaStream_t st) {
    switch (M) {
#define OIVA_CASE(M_) case M_: return stream_launch_m##M_(kind, dtype, p, G, st);
        OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
        OIVA_CASE(9) OIVA_CASE(10) OIVA_CASE(11) OIVA_CASE(12) OIVA_CASE(13) OIVA_CASE(14) OIVA_CASE(15) OIVA_CASE(16)
#undef OIVA_CASE
    }
    oiva_set_error("n_chan=%d unsupported (1..16)", M);
    return OIVA_ERR_INVALID;
}

// r2[b][k][t] = sum_ch r2part[b][ch][k][t]   (one thread per (b,k,t), fixed order => deterministic)
__global__ void k_sum_partials(const double* __restrict__ part, double* __restrict__ r2, int NCH, int K, int Tp,
                               long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long b = i / ((long long)K * Tp);
    const long long rem = i - b * K * Tp;
    double s = 0.0;
    for (int ch = 0; ch < NCH; ++ch) s += part[((size_t)b * NCH + ch) * K * Tp + rem];
    r2[i] = s;
}

// one CTA per (mixture, source): statistic -> r -> gamma -> phi, wscale        (overiva.py:152-173)
// 256 threads = 32 frame lanes x 8 slices of the partial sums: a thread adds every 8th partial of its frames (8
// independent loads in flight per thread instead of a serial chain over all NCH partials -- this kernel sits on
// the critical path of every epoch and is pure latency), the slices are combined through shared memory in a fixed
// order (deterministic).
__global__ void __launch_bounds__(256) k_source_model(const double* __restrict__ part, int NCH, double* __restrict__ phi,
                                                      double* __restrict__ wscale, int T, int Tp, int K, int F_total,
                                                      int model) {
    __shared__ double slice[8][32];
    __shared__ double red[8];
    const int b = blockIdx.x / K, k = blockIdx.x - b * K;
    const int tl = threadIdx.x & 31, cs = threadIdx.x >> 5;
    double* ph = phi + ((size_t)b * K + k) * Tp;
    const double* pb = part + ((size_t)b * NCH * K + k) * Tp;
    double lsum = 0.0;
    for (int t0 = 0; t0 < Tp; t0 += 32) {
        const int t = t0 + tl;  // Tp is a multiple of 32
        double s = 0.0;
        if (t < T) {
#pragma unroll 4
            for (int ch = cs; ch < NCH; ch += 8) s += pb[(size_t)ch * K * Tp + t];
        }
        slice[cs][tl] = s;
        __syncthreads();
        if (cs == 0) {
            double r = 0.0;
            if (t < T) {
                s = ((slice[0][tl] + slice[1][tl]) + (slice[2][tl] + slice[3][tl])) +
                    ((slice[4][tl] + slice[5][tl]) + (slice[6][tl] + slice[7][tl]));
                switch (model) {
                    case OIVA_MODEL_LAPLACE: r = 2.0 * sqrt(s); break;
                    case OIVA_MODEL_GAUSS: r = s / (double)F_total; break;
                    case OIVA_MODEL_OGIVE_LAPLACE: r = sqrt(s) / sqrt((double)F_total); break;
                    case OIVA_MODEL_OGIVE_GAUSS: r = s / (double)F_total; break;
                    default: r = 0.0; break;
                }
                lsum += r;
            }
            ph[t] = r;
        }
        __syncthreads();
    }
    // gamma = mean_t r: warp 0 holds the per-lane sums
    if (cs == 0) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
        if (tl == 0) red[0] = lsum;
    }
    __syncthreads();
    const bool rescale = (model == OIVA_MODEL_LAPLACE || model == OIVA_MODEL_GAUSS || model == OIVA_MODEL_NONE);
    const double gamma = rescale ? red[0] / (double)T : 1.0;
    for (int t = threadIdx.x; t < Tp; t += 256) {
        double r = 0.0;
        if (t < T) {
            r = ph[t] / gamma;              // 0/0 -> NaN exactly as the reference's r /= gamma
            if (r < 1e-15) r = 1e-15;       // NaN compares false and stays NaN, as in numpy
            r = 1.0 / r;
        }
        ph[t] = r;
    }
    if (threadIdx.x == 0 && wscale) {
        double w = 1.0;
        if (model == OIVA_MODEL_LAPLACE) w = 1.0 / gamma;
        else if (model == OIVA_MODEL_GAUSS) w = 1.0 / sqrt(gamma);
        wscale[(size_t)b * K + k] = w;
    }
}

// Long mixtures (config 5: T = 14061): the single kernel above walks the frames of one (mixture, source) with ONE CTA,
// 32 frames at a time (measured 0.4 ms per epoch at config 5 -- 4 CTAs on a 148-SM GPU).  Two kernels instead, same
// arithmetic and summation orders: (1) the sum over the partials and the model function for chunks of frames in
// parallel (r is parked in phi), (2) per (mixture, source) gamma -- lane tl adds r[tl], r[tl+32], ... in ascending
// order, then the xor tree, exactly as above -- followed by phi = 1 / max(r / gamma, 1e-15) in place.
// (ph: the r values of one (mixture, source), possibly written by other CTAs of the same launch: read through L2)
__device__ __forceinline__ void source_finish_body(double* ph, double* __restrict__ wscale, int b, int k, int T, int Tp,
                                                   int K, int model, double* red) {
    if (threadIdx.x < 32) {
        double lsum = 0.0;
#pragma unroll 8
        for (int t = threadIdx.x; t < T; t += 32) lsum += __ldcg(ph + t);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
        if (threadIdx.x == 0) red[0] = lsum;
    }
    __syncthreads();
    const bool rescale = (model == OIVA_MODEL_LAPLACE || model == OIVA_MODEL_GAUSS || model == OIVA_MODEL_NONE);
    const double gamma = rescale ? red[0] / (double)T : 1.0;
    for (int t = threadIdx.x; t < Tp; t += 256) {
        double r = 0.0;
        if (t < T) {
            r = __ldcg(ph + t) / gamma;
            if (r < 1e-15) r = 1e-15;
            r = 1.0 / r;
        }
        ph[t] = r;
    }
    if (threadIdx.x == 0 && wscale) {
        double w = 1.0;
        if (model == OIVA_MODEL_LAPLACE) w = 1.0 / gamma;
        else if (model == OIVA_MODEL_GAUSS) w = 1.0 / sqrt(gamma);
        wscale[(size_t)b * K + k] = w;
    }
}
__global__ void __launch_bounds__(256) k_source_finish(double* __restrict__ phi, double* __restrict__ wscale, int T, int Tp,
                                                       int K, int model) {
    __shared__ double red[1];
    const int b = blockIdx.x / K, k = blockIdx.x - b * K;
    source_finish_body(phi + ((size_t)b * K + k) * Tp, wscale, b, k, T, Tp, K, model, red);
}
// (sm_chunk: frames per CTA of the first kernel, a multiple of 32)
// counters != nullptr (one zeroed word per (mixture, source), left at zero again): the LAST CTA of a (mixture, source) to
// finish its frames also runs the finishing step -- one launch for the whole source model.
__global__ void __launch_bounds__(256) k_source_r(const double* __restrict__ part, int NCH, double* __restrict__ phi, int T,
                                                  int Tp, int K, int F_total, int model, int sm_chunk,
                                                  unsigned* counters, double* __restrict__ wscale) {
    __shared__ double slice[8][32];
    __shared__ double red[1];
    __shared__ int is_last;
    const int b = blockIdx.x / K, k = blockIdx.x - b * K;
    const int tl = threadIdx.x & 31, cs = threadIdx.x >> 5;
    double* ph = phi + ((size_t)b * K + k) * Tp;
    const double* pb = part + ((size_t)b * NCH * K + k) * Tp;
    const int t_end = min(Tp, (int)(blockIdx.y + 1) * sm_chunk);
    for (int t0 = blockIdx.y * sm_chunk; t0 < t_end; t0 += 32) {
        const int t = t0 + tl;
        double s = 0.0;
        if (t < T) {
#pragma unroll 4
            for (int ch = cs; ch < NCH; ch += 8) s += pb[(size_t)ch * K * Tp + t];
        }
        slice[cs][tl] = s;
        __syncthreads();
        if (cs == 0) {
            double r = 0.0;
            if (t < T) {
                s = ((slice[0][tl] + slice[1][tl]) + (slice[2][tl] + slice[3][tl])) +
                    ((slice[4][tl] + slice[5][tl]) + (slice[6][tl] + slice[7][tl]));
                switch (model) {
                    case OIVA_MODEL_LAPLACE: r = 2.0 * sqrt(s); break;
                    case OIVA_MODEL_GAUSS: r = s / (double)F_total; break;
                    case OIVA_MODEL_OGIVE_LAPLACE: r = sqrt(s) / sqrt((double)F_total); break;
                    case OIVA_MODEL_OGIVE_GAUSS: r = s / (double)F_total; break;
                    default: r = 0.0; break;
                }
            }
            ph[t] = r;
        }
        __syncthreads();
    }
    if (!counters) return;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned old = atomicAdd(counters + blockIdx.x, 1u);
        is_last = old == gridDim.y - 1;
        if (is_last) {
            counters[blockIdx.x] = 0;  // ready for the next epoch
            __threadfence();
        }
    }
    __syncthreads();
    if (is_last) source_finish_body(ph, wscale, b, k, T, Tp, K, model, red);
}

// projection-back scales from the grouped state, any M: z[gi][k][lane] with exactly the arithmetic of k_projback_filters
// (solve.cu).  One thread per (bin, source) -- blockIdx.y = k; the source's column of W_hat and one row of C at a time in
// registers, all of a row's loads in flight together.  (The first version walked a, b with one dependent load per step: 25 us
// for the 2049 bins of config 3, a quarter of its output step.)
__global__ void __launch_bounds__(128) k_projback_z(const cplx* __restrict__ Wg, const cplx* __restrict__ Cg,
                                                    cplx* __restrict__ Zg, long long G, int M, int K) {
    const int lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int k = blockIdx.y;
    if (gi >= G) return;
    const cplx* Wl = Wg + (size_t)gi * M * M * OIVA_GROUP + lane;
    const cplx* Cl = Cg + (size_t)gi * (M * (M + 1) / 2) * OIVA_GROUP + lane;
    cplx wk[OIVA_MAX_M];
#pragma unroll
    for (int b = 0; b < OIVA_MAX_M; ++b) wk[b] = b < M ? Wl[(size_t)(b * M + k) * OIVA_GROUP] : cmake(0.0, 0.0);
    cplx num = cmake(0.0, 0.0);
    double den = 0.0;
    for (int a = 0; a < M; ++a) {
        cplx crow[OIVA_MAX_M];  // C[a][b]
#pragma unroll
        for (int b = 0; b < OIVA_MAX_M; ++b) {
            if (b < M) {
                const int hi = a >= b ? a : b, lo = a >= b ? b : a;
                cplx v = ld_nc_c(Cl + (size_t)(hi * (hi + 1) / 2 + lo) * OIVA_GROUP);
                if (a < b) v.y = -v.y;
                crow[b] = v;
            } else {
                crow[b] = cmake(0.0, 0.0);
            }
        }
        const cplx wa = Wl[(size_t)(a * M + k) * OIVA_GROUP];
        cfmac(num, wa, crow[0]);
        cplx cw = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < OIVA_MAX_M; ++b)
            if (b < M) cfma(cw, crow[b], wk[b]);
        den += wa.x * cw.x + wa.y * cw.y;
    }
    cplx z = cmake(1.0, 0.0);
    if (den > 0.0) z = cmake(num.x / den, num.y / den);
    Zg[((size_t)gi * K + k) * OIVA_GROUP + lane] = z;
}

}  // namespace oiva

using namespace oiva;

static int check_dims(const char* who, int B, int T, int F, int M, int K) {
    OIVA_REQUIRE(B > 0 && T > 0 && F > 0 && M >= 1 && M <= OIVA_MAX_M && K >= 1 && K <= OIVA_MAX_M,
                 "%s: bad shape B=%d T=%d F=%d M=%d K=%d", who, B, T, F, M, K);
    return OIVA_OK;
}

static StreamParams make_params(const void* Xg, const void* W, int w_cols, int T, int F, int M, int K,
                                int w_grouped = 0) {
    StreamParams p = {};
    p.Xg = Xg;
    p.W = (const cplx*)W;
    p.w_row = (long long)M * w_cols;
    p.w_c = w_cols;
    p.w_grouped = w_grouped;
    p.L = oiva_make_layout(T, F, M);
    p.K = K;
    p.nsplit = 1;
    return p;
}

extern "C" int oiva_demix_power(const void* Xg, const void* W, int w_cols, int w_grouped, double* r2part, int n_batch,
                                int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xg && W && r2part, "oiva_demix_power: null pointer");
    OIVA_REQUIRE(w_cols >= n_src, "oiva_demix_power: w_cols %d < n_src %d", w_cols, n_src);
    int rc = check_dims("oiva_demix_power", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    OIVA_REQUIRE(n_src <= n_chan, "oiva_demix_power: n_src > n_chan");
    StreamParams p = make_params(Xg, W, w_cols, n_frames, n_freq, n_chan, n_src, w_grouped);
    p.r2part = r2part;
    return stream_launch(n_chan, KIND_POWER, dtype, p, (long long)n_batch * p.L.NG, (cudaStream_t)stream);
}

// the statistic AND the per-bin powers |y_k(f, t)|^2 (Pfull: grouped [gi][k][Tp][32] float64) in one pass: ILRMA's
// spectrogram model needs every bin's power, its scale normalisation the sums (r2part).  W as for oiva_demix_power.
extern "C" int oiva_demix_power_full(const void* Xg, const void* W, int w_cols, int w_grouped, double* r2part, double* Pfull,
                                     int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xg && W && r2part && Pfull, "oiva_demix_power_full: null pointer");
    OIVA_REQUIRE(w_cols >= n_src, "oiva_demix_power_full: w_cols %d < n_src %d", w_cols, n_src);
    int rc = check_dims("oiva_demix_power_full", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    OIVA_REQUIRE(n_src <= n_chan, "oiva_demix_power_full: n_src > n_chan");
    StreamParams p = make_params(Xg, W, w_cols, n_frames, n_freq, n_chan, n_src, w_grouped);
    p.r2part = r2part;
    p.Pfull = Pfull;
    return stream_launch(n_chan, KIND_POWER, dtype, p, (long long)n_batch * p.L.NG, (cudaStream_t)stream);
}

extern "C" int oiva_sum_partials(const double* r2part, int n_chunks, double* r2, int n_batch, int n_frames, int n_src,
                                 void* stream) {
    OIVA_REQUIRE(r2part && r2 && n_chunks >= 1, "oiva_sum_partials: bad arguments");
    const int Tp = oiva_frame_pitch(n_frames);
    const long long n = (long long)n_batch * n_src * Tp;
    k_sum_partials<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r2part, r2, n_chunks, n_src, Tp, n);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_source_model_ws(const double* r2part, int n_chunks, double* phi, double* wscale, unsigned* counters,
                                    int n_batch, int n_frames, int n_src, int n_freq_total, int model, void* stream);

extern "C" int oiva_source_model(const double* r2part, int n_chunks, double* phi, double* wscale, int n_batch,
                                 int n_frames, int n_src, int n_freq_total, int model, void* stream) {
    return oiva_source_model_ws(r2part, n_chunks, phi, wscale, nullptr, n_batch, n_frames, n_src, n_freq_total, model,
                                stream);
}

// counters: n_batch * n_src zeroed words of device memory (left zeroed), or NULL.  With counters the frame-parallel form
// is one launch (the last CTA of a (mixture, source) finishes it) instead of two.
extern "C" int oiva_source_model_ws(const double* r2part, int n_chunks, double* phi, double* wscale, unsigned* counters,
                                    int n_batch, int n_frames, int n_src, int n_freq_total, int model, void* stream) {
    OIVA_REQUIRE(r2part && phi && n_chunks >= 1, "oiva_source_model: bad arguments");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_src >= 1 && n_freq_total > 0, "oiva_source_model: bad shape");
    const int Tp = oiva_frame_pitch(n_frames);
    cudaStream_t st = (cudaStream_t)stream;
    // Long mixtures, and few (mixture, source) pairs with more than a couple of 32-frame blocks (one short mixture: 2-6
    // CTAs walking their frames block after block, each block a dependent round of loads over the partials -- ~20 us per
    // epoch at config 3): the frames in parallel, then one small finishing kernel.  Same sums in the same order.
    const bool few_pairs = (long long)n_batch * n_src < 64 && n_frames > 64;
    if (n_frames > 2048 || few_pairs) {
        const int sm_chunk = n_frames > 2048 ? 256 : 32;
        const int chunks = (Tp + sm_chunk - 1) / sm_chunk;
        OIVA_REQUIRE(chunks <= 65535, "oiva_source_model: too many frames");
        k_source_r<<<dim3((unsigned)(n_batch * n_src), (unsigned)chunks), 256, 0, st>>>(
            r2part, n_chunks, phi, n_frames, Tp, n_src, n_freq_total, model, sm_chunk, counters, wscale);
        OIVA_LAUNCH_CHECK();
        if (!counters) {
            k_source_finish<<<(unsigned)(n_batch * n_src), 256, 0, st>>>(phi, wscale, n_frames, Tp, n_src, model);
            OIVA_LAUNCH_CHECK();
        }
        return OIVA_OK;
    }
    k_source_model<<<(unsigned)(n_batch * n_src), 256, 0, st>>>(r2part, n_chunks, phi, wscale, n_frames, Tp, n_src,
                                                                 n_freq_total, model);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_demix_output(const void* Xg, const void* Weff, void* Y, int n_batch, int n_frames, int n_freq,
                                 int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xg && Weff && Y, "oiva_demix_output: null pointer");
    int rc = check_dims("oiva_demix_output", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    StreamParams p = make_params(Xg, Weff, n_src, n_frames, n_freq, n_chan, n_src);
    p.Y = Y;
    return stream_launch(n_chan, KIND_OUTPUT, dtype, p, (long long)n_batch * p.L.NG, (cudaStream_t)stream);
}

// the final demix straight from the loop's grouped state: Y = (w_k z_k)^H x with the projection-back scales computed per
// bin from the grouped covariance (Cg != NULL) -- two launches, no row-major copies (was ungroup + projback + demix)
extern "C" int oiva_demix_output_grouped(const void* Xg, const void* Wg, const void* Cg, void* zscratch, void* Y,
                                         int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype,
                                         void* stream) {
    OIVA_REQUIRE(Xg && Wg && Y, "oiva_demix_output_grouped: null pointer");
    int rc = check_dims("oiva_demix_output_grouped", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    OIVA_REQUIRE(n_src <= n_chan, "oiva_demix_output_grouped: n_src > n_chan");
    StreamParams p = make_params(Xg, Wg, n_chan, n_frames, n_freq, n_chan, n_src, 1);
    p.Y = Y;
    if (Cg) {  // the scales in a small kernel of their own (thread per bin on the grouped arrays)
        OIVA_REQUIRE(zscratch, "oiva_demix_output_grouped: projection back needs zscratch");
        const long long G = (long long)n_batch * p.L.NG;
        k_projback_z<<<dim3((unsigned)((G + 3) / 4), (unsigned)n_src), 128, 0, (cudaStream_t)stream>>>((const cplx*)Wg, (const cplx*)Cg,
                                                                                  (cplx*)zscratch, G, n_chan, n_src);
        OIVA_LAUNCH_CHECK();
        p.Zg = (const cplx*)zscratch;
    }
    return stream_launch(n_chan, KIND_OUTPUT, dtype, p, (long long)n_batch * p.L.NG, (cudaStream_t)stream);
}

// oiva_demix_output_grouped with caller-supplied per-bin scales Zg [gi][K][32] c128 (NULL: none) instead of the
// projection-back scales: Y = (w_k z_k)^H x
extern "C" int oiva_demix_output_scaled(const void* Xg, const void* Wg, const void* Zg, void* Y, int n_batch, int n_frames,
                                        int n_freq, int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xg && Wg && Y, "oiva_demix_output_scaled: null pointer");
    int rc = check_dims("oiva_demix_output_scaled", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    OIVA_REQUIRE(n_src <= n_chan, "oiva_demix_output_scaled: n_src > n_chan");
    StreamParams p = make_params(Xg, Wg, n_chan, n_frames, n_freq, n_chan, n_src, 1);
    p.Y = Y;
    p.Zg = (const cplx*)Zg;
    return stream_launch(n_chan, KIND_OUTPUT, dtype, p, (long long)n_batch * p.L.NG, (cudaStream_t)stream);
}

extern "C" int oiva_project_rows(const void* Xg, const void* E, void* Xr, int n_batch, int n_frames, int n_freq,
                                 int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xg && E && Xr, "oiva_project_rows: null pointer");
    int rc = check_dims("oiva_project_rows", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    StreamParams p = make_params(Xg, E, n_src, n_frames, n_freq, n_chan, n_src);
    p.Xr = Xr;
    return stream_launch(n_chan, KIND_PROJECT, dtype, p, (long long)n_batch * p.L.NG, (cudaStream_t)stream);
}
