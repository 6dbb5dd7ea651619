// STFT analysis / synthesis feeding and consuming the demixing loop on the device.
//
// The reference's drivers call pyroomacoustics for this (third-party, not in the reference tree):
//   X = pra.transform.analysis(mix.T, framesize, framesize // 2, win=win_a)      overiva_oneshot.py:293-295,
//                                                                                 overiva_sim.py:206-207
//   y = pra.transform.synthesis(Y, framesize, framesize // 2, win=win_s)          overiva_oneshot.py:371-379,
//                                                                                 overiva_sim.py:213-218
// with framesize 4096, hop 2048, a Hann analysis window and the matched synthesis window
// (overiva_oneshot.py:156-158).  Here one CTA transforms one frame of one channel entirely in shared memory
// (real FFT of length L as a complex FFT of length L/2 on the even/odd samples, fp64, radix-2 with table
// twiddles) and writes the L/2+1 bins EITHER in the caller's (B,T,F,M) order OR directly in the grouped layout
// Xg[gi][t][c][lane] the loop streams -- so "audio in" needs neither a (T,F,M) array in HBM nor the relayout pass.
// Synthesis is the inverse: one CTA per (mixture, frame, source) -> windowed time frame, then a gather
// overlap-add (fixed summation order: deterministic).
#include "common.cuh"

namespace {

constexpr int STFT_THREADS = 256;

// tw[q] = exp(-2 pi i q / L), q < L/2
__global__ void k_twiddles(cplx* __restrict__ tw, int L) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < L / 2) {
        double s, c;
        sincospi(-2.0 * (double)q / (double)L, &s, &c);
        tw[q] = cmake(c, s);
    }
}

// In-place radix-2 decimation-in-time FFT of n2 = 2^lg points held in shared memory in BIT-REVERSED order.
// tw[2 q] = exp(-2 pi i q / n2); INVERSE conjugates the twiddles (no scaling).
template <bool INVERSE>
__device__ __forceinline__ void fft_smem(cplx* z, int lg, const cplx* __restrict__ tw) {
    const int n2 = 1 << lg;
    for (int s = 0; s < lg; ++s) {
        __syncthreads();
        const int half = 1 << s;
        for (int j = threadIdx.x; j < n2 / 2; j += STFT_THREADS) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> s) << (s + 1)) + pos;
            const int i1 = i0 + half;
            cplx w = __ldg(&tw[(size_t)(pos << (lg - 1 - s)) * 2]);
            if (INVERSE) w.y = -w.y;
            const cplx a = z[i0];
            const cplx b = cmul(z[i1], w);
            z[i0] = cadd(a, b);
            z[i1] = csub(a, b);
        }
    }
    __syncthreads();
}

struct AnalysisParams {
    const void* x;        // audio samples, element (b, n, c) at x[b*sb + n*sn + c*sc]
    long long sb, sn, sc; // strides in elements
    long long N;          // samples per channel
    long long first;      // frame t covers samples first + t*hop + [0, L); samples outside [0, N) read as 0
    const double* win;    // (L) analysis window or NULL
    const cplx* tw;       // (L/2)
    void* out;
    int B, T, M, L, lg, hop, F, NG;
    int grouped;          // 1: Xg[gi][t][c][lane];  0: X (B,T,F,M)
};

template <typename AT, typename ST>
__global__ void __launch_bounds__(STFT_THREADS) k_stft_analysis(const AnalysisParams p) {
    typedef typename StoreC<ST>::type XC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* z = reinterpret_cast<cplx*>(smem_raw);
    const int c = blockIdx.x % p.M;
    const int t = blockIdx.x / p.M;
    const int b = blockIdx.y;
    const int n2 = p.L / 2, lg = p.lg;
    const AT* x = reinterpret_cast<const AT*>(p.x) + (size_t)b * p.sb + (size_t)c * p.sc;
    const long long n0 = p.first + (long long)t * p.hop;
    for (int n = threadIdx.x; n < n2; n += STFT_THREADS) {
        const long long s0 = n0 + 2 * n, s1 = s0 + 1;
        double v0 = (s0 >= 0 && s0 < p.N) ? (double)x[s0 * p.sn] : 0.0;
        double v1 = (s1 >= 0 && s1 < p.N) ? (double)x[s1 * p.sn] : 0.0;
        if (p.win) {
            v0 *= __ldg(&p.win[2 * n]);
            v1 *= __ldg(&p.win[2 * n + 1]);
        }
        z[__brev((unsigned)n) >> (32 - lg)] = cmake(v0, v1);
    }
    fft_smem<false>(z, lg, p.tw);
    // X[k] = E[k] + exp(-2 pi i k / L) O[k],  E = (Z[k] + conj Z[n2-k]) / 2,  O = -i (Z[k] - conj Z[n2-k]) / 2
    XC* out = reinterpret_cast<XC*>(p.out);
    const int kmax = p.grouped ? p.NG * OIVA_GROUP : p.F;
    for (int k = threadIdx.x; k < kmax; k += STFT_THREADS) {
        cplx X = cmake(0.0, 0.0);
        if (k < p.F) {
            const cplx zk = z[k & (n2 - 1)];
            const cplx zm = cconj(z[(n2 - k) & (n2 - 1)]);
            const cplx e = cscale(cadd(zk, zm), 0.5);
            const cplx d = cscale(csub(zk, zm), 0.5);
            const cplx o = cmake(d.y, -d.x);  // -i d
            const cplx w = k < n2 ? __ldg(&p.tw[k]) : cmake(-1.0, 0.0);
            X = cadd(e, cmul(w, o));
            if (k == 0 || k == n2) X.y = 0.0;  // exactly real for real input
        }
        XC v;
        narrow(v, X);
        if (p.grouped) {
            const size_t gi = (size_t)b * p.NG + (k >> 5);
            out[((gi * p.T + t) * p.M + c) * OIVA_GROUP + (k & 31)] = v;
        } else {
            out[(((size_t)b * p.T + t) * p.F + k) * p.M + c] = v;
        }
    }
}

struct SynthesisParams {
    const void* Y;      // (B,T,F,K) complex
    const double* win;  // (L) synthesis window or NULL
    const cplx* tw;
    double* frames;     // (B,T,K,L) windowed time frames
    int B, T, K, L, lg, F;
};

template <typename ST>
__global__ void __launch_bounds__(STFT_THREADS) k_stft_frames(const SynthesisParams p) {
    typedef typename StoreC<ST>::type XC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* z = reinterpret_cast<cplx*>(smem_raw);
    const int k = blockIdx.x % p.K;
    const int t = blockIdx.x / p.K;
    const int b = blockIdx.y;
    const int n2 = p.L / 2, lg = p.lg;
    const XC* Y = reinterpret_cast<const XC*>(p.Y) + ((size_t)b * p.T + t) * p.F * p.K + k;
    // Z[q] = E[q] + i O[q],  E = (X[q] + conj X[n2-q]) / 2,  O = exp(+2 pi i q / L) (X[q] - conj X[n2-q]) / 2
    for (int q = threadIdx.x; q < n2; q += STFT_THREADS) {
        cplx xq = widen(Y[(size_t)q * p.K]);
        cplx xm = widen(Y[(size_t)(n2 - q) * p.K]);
        if (q == 0) {  // numpy.fft.irfft ignores the imaginary parts of the DC and Nyquist bins
            xq.y = 0.0;
            xm.y = 0.0;
        }
        xm.y = -xm.y;
        const cplx e = cscale(cadd(xq, xm), 0.5);
        const cplx d = cscale(csub(xq, xm), 0.5);
        cplx w = __ldg(&p.tw[q]);
        w.y = -w.y;
        const cplx o = cmul(w, d);
        z[__brev((unsigned)q) >> (32 - lg)] = cmake(e.x - o.y, e.y + o.x);  // e + i o
    }
    fft_smem<true>(z, lg, p.tw);
    const double scale = 1.0 / (double)n2;
    double* fr = p.frames + (((size_t)b * p.T + t) * p.K + k) * p.L;
    for (int n = threadIdx.x; n < n2; n += STFT_THREADS) {
        const cplx v = z[n];
        double a = v.x * scale, c = v.y * scale;
        if (p.win) {
            a *= __ldg(&p.win[2 * n]);
            c *= __ldg(&p.win[2 * n + 1]);
        }
        reinterpret_cast<double2*>(fr)[n] = make_double2(a, c);
    }
}

// y[b][n][k] = sum over the frames t covering sample n of frames[b][t][k][n - t*hop], t ascending
template <typename AT>
__global__ void k_overlap_add(const double* __restrict__ frames, AT* __restrict__ y, int B, int T, int K, int L, int hop,
                              long long n_out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_b = n_out * K;
    if (idx >= per_b) return;
    const int b = blockIdx.y;
    const long long n = idx / K;
    const int k = (int)(idx - n * K);
    long long t_hi = n / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    long long t_lo = (n - L + hop) / hop;  // smallest t with n - t*hop < L  (ceil((n - L + 1) / hop))
    if (n - L + 1 <= 0) t_lo = 0;
    double acc = 0.0;
    for (long long t = t_lo; t <= t_hi; ++t)
        acc += frames[(((size_t)b * T + t) * K + k) * L + (n - t * hop)];
    y[(size_t)b * per_b + idx] = (AT)acc;
}

int log2_exact(int v) {
    int lg = 0;
    while ((1 << lg) < v) ++lg;
    return (1 << lg) == v ? lg : -1;
}

}  // namespace

extern "C" int oiva_stft_twiddles(void* tw, int frame_len, void* stream) {
    OIVA_REQUIRE(tw, "oiva_stft_twiddles: null pointer");
    OIVA_REQUIRE(log2_exact(frame_len) >= 3 && frame_len <= 8192, "oiva_stft_twiddles: frame_len=%d must be a power of two in 8..8192",
                 frame_len);
    k_twiddles<<<oiva_div_up(frame_len / 2, 256), 256, 0, (cudaStream_t)stream>>>((cplx*)tw, frame_len);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_stft_num_frames(long long n_samples, int frame_len, int hop, long long pad_front, long long pad_back) {
    if (frame_len <= 0 || hop <= 0) return 0;
    const long long n = n_samples + pad_front + pad_back;
    if (n < frame_len) return 0;
    const long long t = (n - frame_len) / hop + 1;
    return t > 0x7fffffff ? 0 : (int)t;
}

extern "C" int oiva_stft_analysis(const void* x, int x_f32, long long stride_b, long long stride_n, long long stride_c,
                                  long long n_samples, long long pad_front, const double* win, const void* tw, void* out,
                                  int grouped, int n_batch, int n_frames, int n_chan, int frame_len, int hop, int dtype,
                                  void* stream) {
    OIVA_REQUIRE(x && tw && out, "oiva_stft_analysis: null pointer");
    const int lg = log2_exact(frame_len);
    OIVA_REQUIRE(lg >= 3 && frame_len <= 8192, "oiva_stft_analysis: frame_len=%d must be a power of two in 8..8192", frame_len);
    OIVA_REQUIRE(hop >= 1 && hop <= frame_len, "oiva_stft_analysis: hop=%d not in 1..frame_len", hop);
    OIVA_REQUIRE(n_batch > 0 && n_batch <= 65535 && n_frames > 0 && n_chan > 0 && n_chan <= OIVA_MAX_M,
                 "oiva_stft_analysis: bad shape B=%d T=%d M=%d", n_batch, n_frames, n_chan);
    OIVA_REQUIRE(dtype == OIVA_C128 || dtype == OIVA_C64, "oiva_stft_analysis: bad dtype %d", dtype);
    AnalysisParams p;
    p.x = x;
    p.sb = stride_b;
    p.sn = stride_n;
    p.sc = stride_c;
    p.N = n_samples;
    p.first = -pad_front;
    p.win = win;
    p.tw = (const cplx*)tw;
    p.out = out;
    p.B = n_batch;
    p.T = n_frames;
    p.M = n_chan;
    p.L = frame_len;
    p.lg = lg - 1;
    p.hop = hop;
    p.F = frame_len / 2 + 1;
    p.NG = oiva_bin_groups(p.F);
    p.grouped = grouped;
    const size_t smem = (size_t)(frame_len / 2) * sizeof(cplx);
    dim3 grid((unsigned)((long long)n_frames * n_chan), n_batch);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_A(AT, ST)                                                                                           \
    do {                                                                                                           \
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(k_stft_analysis<AT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             64 * 1024));                                                          \
        k_stft_analysis<AT, ST><<<grid, STFT_THREADS, smem, st>>>(p);                                              \
    } while (0)
    if (x_f32) {
        if (dtype == OIVA_C64) LAUNCH_A(float, float); else LAUNCH_A(float, double);
    } else {
        if (dtype == OIVA_C64) LAUNCH_A(double, float); else LAUNCH_A(double, double);
    }
#undef LAUNCH_A
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" size_t oiva_stft_scratch_bytes(int n_batch, int n_frames, int n_src, int frame_len) {
    if (n_batch <= 0 || n_frames <= 0 || n_src <= 0 || frame_len <= 0) return 0;
    return (size_t)n_batch * n_frames * n_src * frame_len * sizeof(double);
}

extern "C" int oiva_stft_synthesis(const void* Y, const double* win, const void* tw, void* scratch, void* y, int y_f32,
                                   int n_batch, int n_frames, int n_src, int frame_len, int hop, int dtype,
                                   void* stream) {
    OIVA_REQUIRE(Y && tw && scratch && y, "oiva_stft_synthesis: null pointer");
    const int lg = log2_exact(frame_len);
    OIVA_REQUIRE(lg >= 3 && frame_len <= 8192, "oiva_stft_synthesis: frame_len=%d must be a power of two in 8..8192", frame_len);
    OIVA_REQUIRE(hop >= 1 && hop <= frame_len, "oiva_stft_synthesis: hop=%d not in 1..frame_len", hop);
    OIVA_REQUIRE(n_batch > 0 && n_batch <= 65535 && n_frames > 0 && n_src > 0, "oiva_stft_synthesis: bad shape B=%d T=%d K=%d",
                 n_batch, n_frames, n_src);
    OIVA_REQUIRE(dtype == OIVA_C128 || dtype == OIVA_C64, "oiva_stft_synthesis: bad dtype %d", dtype);
    SynthesisParams p;
    p.Y = Y;
    p.win = win;
    p.tw = (const cplx*)tw;
    p.frames = (double*)scratch;
    p.B = n_batch;
    p.T = n_frames;
    p.K = n_src;
    p.L = frame_len;
    p.lg = lg - 1;
    p.F = frame_len / 2 + 1;
    const size_t smem = (size_t)(frame_len / 2) * sizeof(cplx);
    dim3 grid((unsigned)((long long)n_frames * n_src), n_batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OIVA_C64) {
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(k_stft_frames<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        k_stft_frames<float><<<grid, STFT_THREADS, smem, st>>>(p);
    } else {
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(k_stft_frames<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        k_stft_frames<double><<<grid, STFT_THREADS, smem, st>>>(p);
    }
    OIVA_LAUNCH_CHECK();
    const long long n_out = (long long)(n_frames - 1) * hop + frame_len;
    dim3 g2((unsigned)oiva_div_up(n_out * n_src, 256), n_batch);
    if (y_f32)
        k_overlap_add<float><<<g2, 256, 0, st>>>(p.frames, (float*)y, n_batch, n_frames, n_src, frame_len, hop, n_out);
    else
        k_overlap_add<double><<<g2, 256, 0, st>>>(p.frames, (double*)y, n_batch, n_frames, n_src, frame_len, hop, n_out);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Gram matrix of a few long real signals: the only pass over the audio that an SDR / SIR evaluation needs
// (convergence monitoring without copying the separated audio to the host every 10
// epochs; reference: mir_eval.separation.bss_eval_sources inside convergence_callback,
// overiva_oneshot.py:263-284, overiva_sim.py:210-232).  Rows 0..Ra-1 come from `a`, rows Ra..Ra+Rb-1 from `b`
// (arbitrary row / sample strides, so slices of (N, K) channel-last audio are read in place).
// Two passes, fixed summation order (deterministic): per-chunk partial dot products, then their sum.
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_CHUNK = 8192;  // samples per CTA

struct GramParams {
    const double* a;
    long long a_rs, a_ss;
    int Ra;
    const double* b;
    long long b_rs, b_ss;
    int Rb;
    long long n;
    int n_chunks;
    double* partials;  // (R(R+1)/2, n_chunks)
    double* out;       // (R, R)
};

__device__ __forceinline__ const double* gram_row(const GramParams& p, int r, long long& ss) {
    if (r < p.Ra) {
        ss = p.a_ss;
        return p.a + (size_t)r * p.a_rs;
    }
    ss = p.b_ss;
    return p.b + (size_t)(r - p.Ra) * p.b_rs;
}

// grid (n_chunks, n_pairs): pair e = i(i+1)/2 + j, i >= j
__global__ void __launch_bounds__(GRAM_THREADS) k_gram_partial(const GramParams p) {
    __shared__ double red[GRAM_THREADS / 32];
    int i = 0;
    const int e = blockIdx.y;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    const int j = e - i * (i + 1) / 2;
    long long si, sj;
    const double* ri = gram_row(p, i, si);
    const double* rj = gram_row(p, j, sj);
    const long long n0 = (long long)blockIdx.x * GRAM_CHUNK;
    const long long n1 = min(p.n, n0 + GRAM_CHUNK);
    double acc = 0.0;
    for (long long n = n0 + threadIdx.x; n < n1; n += GRAM_THREADS) acc = fma(ri[n * si], rj[n * sj], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < GRAM_THREADS / 32; ++w) s += red[w];
        p.partials[(size_t)e * p.n_chunks + blockIdx.x] = s;
    }
}

// one warp per pair: fixed-order strided sum of the chunk partials, butterfly, symmetric write
__global__ void k_gram_final(const GramParams p) {
    const int e = blockIdx.x;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    const int j = e - i * (i + 1) / 2;
    double acc = 0.0;
    for (int c = threadIdx.x; c < p.n_chunks; c += 32) acc += p.partials[(size_t)e * p.n_chunks + c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) {
        const int R = p.Ra + p.Rb;
        p.out[i * R + j] = acc;
        p.out[j * R + i] = acc;
    }
}
}  // namespace

extern "C" size_t oiva_gram_scratch_bytes(int n_rows, long long n_samples) {
    if (n_rows <= 0 || n_samples <= 0) return 0;
    return (size_t)(n_rows * (n_rows + 1) / 2) * (size_t)((n_samples + GRAM_CHUNK - 1) / GRAM_CHUNK) * sizeof(double);
}

extern "C" int oiva_gram(const double* a, long long a_row_stride, long long a_sample_stride, int a_rows, const double* b,
                         long long b_row_stride, long long b_sample_stride, int b_rows, long long n_samples,
                         void* scratch, double* out, void* stream) {
    OIVA_REQUIRE(a && scratch && out && (b || b_rows == 0), "oiva_gram: null pointer");
    OIVA_REQUIRE(a_rows >= 1 && b_rows >= 0 && a_rows + b_rows <= 64, "oiva_gram: %d + %d rows not in 1..64", a_rows, b_rows);
    OIVA_REQUIRE(n_samples >= 1 && n_samples < (1ll << 40), "oiva_gram: bad n_samples");
    GramParams p;
    p.a = a;
    p.a_rs = a_row_stride;
    p.a_ss = a_sample_stride;
    p.Ra = a_rows;
    p.b = b;
    p.b_rs = b_row_stride;
    p.b_ss = b_sample_stride;
    p.Rb = b_rows;
    p.n = n_samples;
    p.n_chunks = (int)((n_samples + GRAM_CHUNK - 1) / GRAM_CHUNK);
    p.partials = (double*)scratch;
    p.out = out;
    const int R = a_rows + b_rows, pairs = R * (R + 1) / 2;
    cudaStream_t st = (cudaStream_t)stream;
    k_gram_partial<<<dim3(p.n_chunks, pairs), GRAM_THREADS, 0, st>>>(p);
    OIVA_LAUNCH_CHECK();
    k_gram_final<<<pairs, 32, 0, st>>>(p);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}
