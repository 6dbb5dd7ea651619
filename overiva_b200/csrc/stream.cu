// C entry points of the streaming kernels and the source-model finalisation
// (include/overiva_b200.h: oiva_demix_power, oiva_sum_partials, oiva_source_model, oiva_demix_output,
//  oiva_project_rows).
#include "stream.cuh"

namespace oiva {

#define OIVA_DECL(M)                                                                                \
    int power_launch_m##M(int dtype, const StreamParams& p, int n_batch, cudaStream_t st);          \
    int output_launch_m##M(int dtype, const StreamParams& p, int n_batch, cudaStream_t st);         \
    int project_launch_m##M(int dtype, const StreamParams& p, long long R, cudaStream_t st);
OIVA_DECL(1) OIVA_DECL(2) OIVA_DECL(3) OIVA_DECL(4) OIVA_DECL(5) OIVA_DECL(6) OIVA_DECL(7) OIVA_DECL(8)
OIVA_DECL(9) OIVA_DECL(10) OIVA_DECL(11) OIVA_DECL(12) OIVA_DECL(13) OIVA_DECL(14) OIVA_DECL(15) OIVA_DECL(16)
#undef OIVA_DECL

#define OIVA_M_SWITCH(FN, M, ...)                        \
    switch (M) {                                         \
        case 1: return FN##1(__VA_ARGS__);               \
        case 2: return FN##2(__VA_ARGS__);               \
        case 3: return FN##3(__VA_ARGS__);               \
        case 4: return FN##4(__VA_ARGS__);               \
        case 5: return FN##5(__VA_ARGS__);               \
        case 6: return FN##6(__VA_ARGS__);               \
        case 7: return FN##7(__VA_ARGS__);               \
        case 8: return FN##8(__VA_ARGS__);               \
        case 9: return FN##9(__VA_ARGS__);               \
        case 10: return FN##10(__VA_ARGS__);             \
        case 11: return FN##11(__VA_ARGS__);             \
        case 12: return FN##12(__VA_ARGS__);             \
        case 13: return FN##13(__VA_ARGS__);             \
        case 14: return FN##14(__VA_ARGS__);             \
        case 15: return FN##15(__VA_ARGS__);             \
        case 16: return FN##16(__VA_ARGS__);             \
        default: break;                                  \
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
__global__ void __launch_bounds__(256) k_source_model(const double* __restrict__ part, int NCH, double* __restrict__ phi,
                                                      double* __restrict__ wscale, int T, int Tp, int K, int F_total,
                                                      int model) {
    __shared__ double red[256];
    const int b = blockIdx.x / K, k = blockIdx.x - b * K;
    double* ph = phi + ((size_t)b * K + k) * Tp;
    double lsum = 0.0;
    for (int t = threadIdx.x; t < Tp; t += 256) {
        double r = 0.0;
        if (t < T) {
            double s = 0.0;
            for (int ch = 0; ch < NCH; ++ch) s += part[(((size_t)b * NCH + ch) * K + k) * Tp + t];
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
    red[threadIdx.x] = lsum;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
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

}  // namespace oiva

using namespace oiva;

static int check_dims(const char* who, int B, int T, int F, int M, int K) {
    OIVA_REQUIRE(B > 0 && T > 0 && F > 0 && M >= 1 && M <= OIVA_MAX_M && K >= 1 && K <= OIVA_MAX_M,
                 "%s: bad shape B=%d T=%d F=%d M=%d K=%d", who, B, T, F, M, K);
    OIVA_REQUIRE(B <= 65535, "%s: n_batch %d > 65535", who, B);
    return OIVA_OK;
}

extern "C" int oiva_demix_power(const void* Xp, const void* What, int w_cols, double* r2part, int n_chunks,
                                int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype,
                                void* stream) {
    OIVA_REQUIRE(Xp && What && r2part, "oiva_demix_power: null pointer");
    OIVA_REQUIRE(w_cols >= n_src, "oiva_demix_power: w_cols %d < n_src %d", w_cols, n_src);
    int rc = check_dims("oiva_demix_power", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    OIVA_REQUIRE(n_src <= n_chan && n_chunks >= 1 && n_chunks <= n_freq, "oiva_demix_power: bad K / chunk count");
    StreamParams p = {};
    p.Xp = Xp;
    p.W = (const cplx*)What;
    p.w_row = (long long)n_chan * w_cols;
    p.w_c = w_cols;
    p.L = oiva_make_layout(n_frames, n_chan, dtype);
    p.F = n_freq;
    p.K = n_src;
    p.NCH = n_chunks;
    p.NBC = oiva_div_up(n_freq, n_chunks);
    OIVA_REQUIRE(oiva_div_up(n_freq, p.NBC) <= n_chunks, "oiva_demix_power: chunking mismatch");
    p.r2part = r2part;
    // chunks that own no bins still write zeros (grid.x == n_chunks)
    OIVA_M_SWITCH(power_launch_m, n_chan, dtype, p, n_batch, (cudaStream_t)stream)
    return OIVA_ERR_INVALID;
}

extern "C" int oiva_sum_partials(const double* r2part, int n_chunks, double* r2, int n_batch, int n_frames, int n_chan,
                                 int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(r2part && r2 && n_chunks >= 1, "oiva_sum_partials: bad arguments");
    const int Tp = oiva_frame_pitch(n_frames, n_chan, dtype);
    const long long n = (long long)n_batch * n_src * Tp;
    k_sum_partials<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r2part, r2, n_chunks, n_src, Tp, n);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_source_model(const double* r2part, int n_chunks, double* phi, double* wscale, int n_batch,
                                 int n_frames, int n_chan, int n_src, int n_freq_total, int model, int dtype,
                                 void* stream) {
    OIVA_REQUIRE(r2part && phi && n_chunks >= 1, "oiva_source_model: bad arguments");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_src >= 1 && n_freq_total > 0, "oiva_source_model: bad shape");
    const int Tp = oiva_frame_pitch(n_frames, n_chan, dtype);
    k_source_model<<<(unsigned)(n_batch * n_src), 256, 0, (cudaStream_t)stream>>>(r2part, n_chunks, phi, wscale,
                                                                                   n_frames, Tp, n_src, n_freq_total,
                                                                                   model);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_demix_output(const void* Xp, const void* Weff, void* Y, int n_batch, int n_frames, int n_freq,
                                 int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xp && Weff && Y, "oiva_demix_output: null pointer");
    int rc = check_dims("oiva_demix_output", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    StreamParams p = {};
    p.Xp = Xp;
    p.W = (const cplx*)Weff;
    p.w_row = (long long)n_chan * n_src;
    p.w_c = n_src;
    p.L = oiva_make_layout(n_frames, n_chan, dtype);
    p.F = n_freq;
    p.K = n_src;
    p.Y = Y;
    p.NBF = 64 / n_src;
    if (p.NBF < 4) p.NBF = 4;
    OIVA_REQUIRE(oiva_div_up(n_frames, 32) <= 65535, "oiva_demix_output: too many frames");
    OIVA_M_SWITCH(output_launch_m, n_chan, dtype, p, n_batch, (cudaStream_t)stream)
    return OIVA_ERR_INVALID;
}

extern "C" int oiva_project_rows(const void* Xp, const void* E, void* Xr, int n_batch, int n_frames, int n_freq,
                                 int n_chan, int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Xp && E && Xr, "oiva_project_rows: null pointer");
    int rc = check_dims("oiva_project_rows", n_batch, n_frames, n_freq, n_chan, n_src);
    if (rc) return rc;
    StreamParams p = {};
    p.Xp = Xp;
    p.W = (const cplx*)E;
    p.w_row = (long long)n_chan * n_src;
    p.w_c = n_src;
    p.L = oiva_make_layout(n_frames, n_chan, dtype);
    p.Lr = oiva_make_layout(n_frames, n_src, dtype);
    p.F = n_freq;
    p.K = n_src;
    p.Xr = Xr;
    const long long R = (long long)n_batch * n_freq;
    OIVA_M_SWITCH(project_launch_m, n_chan, dtype, p, R, (cudaStream_t)stream)
    return OIVA_ERR_INVALID;
}
