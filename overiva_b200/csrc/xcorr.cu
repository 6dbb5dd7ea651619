// Cross-correlations at lags 0 .. flen-1 between real signals on the device: the only pass over the audio that BSS Eval
// with time-invariant distortion FILTERS needs (mir_eval.separation.bss_eval_sources, the metric of the reference's
// convergence callback, overiva_oneshot.py:263-284, overiva_sim.py:210-232: flen = 512).  Every quantity of that metric
// is a function of   c_ij[m] = sum_n s_i[n] s_j[n + m]   between the references (the block-Toeplitz normal matrix of
// the projection) and between references and estimates (its right-hand sides), plus the energies of the estimates --
// (K + J) K flen numbers go to the host instead of the signals (overiva_b200/monitor.py).
#include "common.cuh"

namespace oiva {

constexpr int XC_TN = 2048;      // samples of s_i per CTA
constexpr int XC_THREADS = 256;
constexpr int XC_MAX_LAGS = 1024;

struct XcorrParams {
    const double* a;  // a_rows signals: the "left" signals s_i (and the first a_rows "right" signals)
    long long a_rs, a_ss;
    int a_rows;
    const double* b;  // b_rows further "right" signals
    long long b_rs, b_ss;
    int b_rows;
    long long N;
    int flen, tiles;
    double* part;  // (a_rows, a_rows + b_rows, tiles, flen)
};

// grid (tiles, a_rows + b_rows, a_rows): partial sums over one tile of n for one ordered pair (i, j), all lags
__global__ void __launch_bounds__(XC_THREADS) k_xcorr_partial(const XcorrParams p) {
    extern __shared__ double sm[];
    double* si = sm;                 // [XC_TN]
    double* sj = sm + XC_TN;         // [XC_TN + flen]
    const int tile = blockIdx.x, j = blockIdx.y, i = blockIdx.z;
    const long long n0 = (long long)tile * XC_TN;
    const double* xi = p.a + (long long)i * p.a_rs;
    const double* xj = j < p.a_rows ? p.a + (long long)j * p.a_rs : p.b + (long long)(j - p.a_rows) * p.b_rs;
    const long long xj_ss = j < p.a_rows ? p.a_ss : p.b_ss;
    for (int n = threadIdx.x; n < XC_TN; n += XC_THREADS) si[n] = n0 + n < p.N ? xi[(n0 + n) * p.a_ss] : 0.0;
    for (int n = threadIdx.x; n < XC_TN + p.flen; n += XC_THREADS) sj[n] = n0 + n < p.N ? xj[(n0 + n) * xj_ss] : 0.0;
    __syncthreads();
    double acc[XC_MAX_LAGS / XC_THREADS];
#pragma unroll
    for (int q = 0; q < XC_MAX_LAGS / XC_THREADS; ++q) acc[q] = 0.0;
    const int nq = (p.flen + XC_THREADS - 1) / XC_THREADS;
    const int nmax = (int)(p.N - n0 < XC_TN ? p.N - n0 : XC_TN);
    for (int n = 0; n < nmax; ++n) {
        const double x = si[n];
#pragma unroll
        for (int q = 0; q < XC_MAX_LAGS / XC_THREADS; ++q)
            if (q < nq) {
                const int m = threadIdx.x + q * XC_THREADS;
                if (m < p.flen) acc[q] = fma(x, sj[n + m], acc[q]);
            }
    }
    double* out = p.part + (((size_t)i * (p.a_rows + p.b_rows) + j) * p.tiles + tile) * p.flen;
#pragma unroll
    for (int q = 0; q < XC_MAX_LAGS / XC_THREADS; ++q) {
        const int m = threadIdx.x + q * XC_THREADS;
        if (q < nq && m < p.flen) out[m] = acc[q];
    }
}

// out[pair][m] = sum over the tiles in ascending order (deterministic)
__global__ void k_xcorr_sum(const double* __restrict__ part, double* __restrict__ out, long long n_out, int tiles, int flen) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const long long pair = i / flen;
    const int m = (int)(i - pair * flen);
    double s = 0.0;
    for (int t = 0; t < tiles; ++t) s += part[((size_t)pair * tiles + t) * flen + m];
    out[i] = s;
}

}  // namespace oiva

using namespace oiva;

extern "C" size_t oiva_xcorr_scratch_bytes(int a_rows, int b_rows, long long n_samples, int flen) {
    if (a_rows < 1 || b_rows < 0 || n_samples < 1 || flen < 1) return 0;
    const long long tiles = (n_samples + XC_TN - 1) / XC_TN;
    return (size_t)a_rows * (a_rows + b_rows) * tiles * flen * sizeof(double);
}

extern "C" int oiva_xcorr(const double* a, long long a_row_stride, long long a_sample_stride, int a_rows, const double* b,
                          long long b_row_stride, long long b_sample_stride, int b_rows, long long n_samples, int flen,
                          void* scratch, double* out, void* stream) {
    OIVA_REQUIRE(a && scratch && out && a_rows >= 1 && b_rows >= 0 && (b || b_rows == 0), "oiva_xcorr: bad arguments");
    OIVA_REQUIRE(n_samples >= 1 && flen >= 1 && flen <= XC_MAX_LAGS, "oiva_xcorr: flen=%d not in 1..%d", flen, XC_MAX_LAGS);
    OIVA_REQUIRE(a_rows + b_rows <= 65535 && a_rows <= 65535, "oiva_xcorr: too many signals");
    cudaStream_t st = (cudaStream_t)stream;
    XcorrParams p;
    p.a = a; p.a_rs = a_row_stride; p.a_ss = a_sample_stride; p.a_rows = a_rows;
    p.b = b; p.b_rs = b_row_stride; p.b_ss = b_sample_stride; p.b_rows = b_rows;
    p.N = n_samples;
    p.flen = flen;
    p.tiles = (int)((n_samples + XC_TN - 1) / XC_TN);
    p.part = (double*)scratch;
    const size_t smem = (size_t)(2 * XC_TN + flen) * sizeof(double);
    OIVA_SET_MAX_SMEM_ONCE(k_xcorr_partial, 2 * XC_TN * sizeof(double) + XC_MAX_LAGS * sizeof(double));
    k_xcorr_partial<<<dim3((unsigned)p.tiles, (unsigned)(a_rows + b_rows), (unsigned)a_rows), XC_THREADS, smem, st>>>(p);
    OIVA_LAUNCH_CHECK();
    const long long n_out = (long long)a_rows * (a_rows + b_rows) * flen;
    k_xcorr_sum<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(p.part, out, n_out, p.tiles, flen);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}
