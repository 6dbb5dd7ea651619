// STFT analysis / synthesis feeding and consuming the demixing loop on the device.
//
// The reference's drivers call pyroomacoustics for this (third-party, not in the reference tree):
//   X = pra.transform.analysis(mix.T, framesize, framesize // 2, win=win_a)      overiva_oneshot.py:293-295,
//                                                                                 overiva_sim.py:206-207
//   y = pra.transform.synthesis(Y, framesize, framesize // 2, win=win_s)          overiva_oneshot.py:371-379,
//                                                                                 overiva_sim.py:213-218
// with framesize 4096, hop 2048, a Hann analysis window and the matched synthesis window
// (overiva_oneshot.py:156-158).
//
// One CTA transforms one frame of TWO channels at once: z = x_c + i x_{c+1} goes through ONE complex FFT of length L
// held in shared memory and the two real spectra are separated afterwards (X_c = (Z[k] + conj Z[L-k]) / 2,
// X_{c+1} = -i (Z[k] - conj Z[L-k]) / 2).  The FFT is a Stockham autosort transform in fp64: radix-8 passes (plus one
// radix-2 / radix-4 pass when log2 L is not a multiple of 3), each thread holding its butterflies in registers
// between the read and the write of a pass so a single buffer suffices; indices are padded by i/8 so that every
// quarter-warp of a 16-byte access hits 32 distinct banks in all passes; the three twiddles w, w^2, w^4 of a
// butterfly come from a table, the others are their products.  Bins are written EITHER in the caller's (B,T,F,M)
// order OR directly in the grouped layout Xg[gi][t][c][lane] the loop streams -- "audio in" then needs neither a
// (T,F,M) array in HBM nor the relayout pass.  Synthesis is the inverse (two sources per complex FFT, inverse by
// conjugation), followed by a gather overlap-add in ascending frame order (deterministic).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int FFT_PTS = 16;  // points per thread: blockDim = L / 16 (at least one warp)

__host__ __device__ inline int stft_threads(int L) { return L / FFT_PTS < 32 ? 32 : L / FFT_PTS; }
__host__ __device__ inline size_t stft_smem_bytes(int L) { return (size_t)(L + L / 8 + 1) * sizeof(cplx); }
__device__ __forceinline__ int pidx(int i) { return i + (i >> 3); }

// Twiddle table of the complex FFT of N = 2^lg points, one compact block per Stockham pass after the first:
// the pass with Ns = 8^p (p >= 1) and radix R reads, for butterfly phase k < Ns, the record
//     tw[3 * ((Ns - 8) / 7 + k) + {0, 1, 2}] = w, w^2, w^4,   w = exp(-2 pi i k / (Ns R)),
// so that consecutive lanes read consecutive 48-byte records (coalesced / conflict-free) instead of gathering
// from one long table with a stride that changes per pass.
__host__ __device__ inline int twiddle_records(int lg) {  // sum of Ns over the passes after the first
    int n = 0, Ns = 8;
    const int passes = lg / 3 + (lg % 3 ? 1 : 0);
    for (int p = 1; p < passes; ++p) {
        n += Ns;
        Ns *= 8;
    }
    return n;
}
__global__ void k_twiddles(cplx* __restrict__ tw, int lg, int n_records) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_records) return;
    int Ns = 8, base = 0, p = 1;
    while (i >= base + Ns) {
        base += Ns;
        Ns *= 8;
        ++p;
    }
    const int k = i - base;
    const int R = p < lg / 3 ? 8 : (lg % 3 == 0 ? 8 : (lg % 3 == 1 ? 2 : 4));
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        double sn, cs;
        sincospi(-2.0 * (double)(k << e) / (double)(Ns * R), &sn, &cs);
        tw[3 * i + e] = cmake(cs, sn);
    }
}

// natural-order forward DFTs of 2 / 4 / 8 points in registers
__device__ __forceinline__ cplx mul_mi(cplx a) { return cmake(a.y, -a.x); }  // -i a
__device__ __forceinline__ void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
    const cplx s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));
    a0 = cadd(s02, s13);
    a1 = cadd(d02, d13);
    a2 = csub(s02, s13);
    a3 = csub(d02, d13);
}
template <int R>
__device__ __forceinline__ void dft_r(cplx (&v)[R]) {
    if constexpr (R == 2) {
        const cplx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if constexpr (R == 4) {
        dft4(v[0], v[1], v[2], v[3]);
    } else {
        cplx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
        dft4(e0, e1, e2, e3);
        dft4(o0, o1, o2, o3);
        const double h = 0.70710678118654752440;
        o1 = cmake(h * (o1.x + o1.y), h * (o1.y - o1.x));    // * (1 - i)/sqrt2
        o2 = mul_mi(o2);                                     // * -i
        o3 = cmake(h * (o3.y - o3.x), -h * (o3.x + o3.y));   // * (-1 - i)/sqrt2
        v[0] = cadd(e0, o0);
        v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1);
        v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2);
        v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3);
        v[7] = csub(e3, o3);
    }
}

// One Stockham pass of radix R over the N points in z (padded indexing); Ns = product of the radices done so far.
// tw: the per-pass twiddle records (k_twiddles).  Every thread reads all its butterflies, then (barrier) writes them.
template <int R>
__device__ __forceinline__ void stockham_pass(cplx* z, int N, int Ns, const cplx* __restrict__ tw) {
    constexpr int NB = FFT_PTS / R;
    cplx v[NB][R];
    const int nbf = N / R;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        if (j < nbf) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[i][r] = z[pidx(j + r * nbf)];
            const int k = j & (Ns - 1);
            if (k != 0) {
                const cplx* rec = tw + 3 * ((Ns - 8) / 7 + k);
                const cplx w1 = __ldg(rec);
                v[i][1] = cmul(v[i][1], w1);
                if constexpr (R >= 4) {
                    const cplx w2 = __ldg(rec + 1);
                    const cplx w3 = cmul(w1, w2);
                    v[i][2] = cmul(v[i][2], w2);
                    v[i][3] = cmul(v[i][3], w3);
                    if constexpr (R == 8) {
                        const cplx w4 = __ldg(rec + 2);
                        v[i][4] = cmul(v[i][4], w4);
                        v[i][5] = cmul(v[i][5], cmul(w4, w1));
                        v[i][6] = cmul(v[i][6], cmul(w4, w2));
                        v[i][7] = cmul(v[i][7], cmul(w4, w3));
                    }
                }
            }
            dft_r<R>(v[i]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        if (j < nbf) {
            const int k = j & (Ns - 1);
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) z[pidx(j0 + r * Ns)] = v[i][r];
        }
    }
    __syncthreads();
}

// forward complex FFT of N = 2^lg points in z (natural order in and out); the caller has synchronised after filling z
__device__ __forceinline__ void fft_forward(cplx* z, int lg, const cplx* __restrict__ tw) {
    const int N = 1 << lg;
    int Ns = 1;
    for (int p = 0; p < lg / 3; ++p) {
        stockham_pass<8>(z, N, Ns, tw);
        Ns *= 8;
    }
    if (lg % 3 == 1) stockham_pass<2>(z, N, Ns, tw);
    if (lg % 3 == 2) stockham_pass<4>(z, N, Ns, tw);
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

// pair (x_c0[s], x_c0+1[s]) of interleaved channel-last audio as one aligned 16-byte (fp64) / 8-byte (fp32) load:
// half the sectors of two scalar loads at a stride of M elements
__device__ __forceinline__ void load_pair(const double* p, double& a, double& b) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void load_pair(const float* p, double& a, double& b) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    a = (double)v.x;
    b = (double)v.y;
}

// grid (T * ceil(M/2), B).  MAXT: 256 (frame lengths <= 4096, three CTAs per SM) or 512 (8192).
// VEC: channel-last audio (stride_c == 1) with an even channel count and an aligned base.
template <typename AT, typename ST, int MAXT, bool VEC, int MINB>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? MINB : 1) k_stft_analysis(const AnalysisParams p) {
    typedef typename StoreC<ST>::type XC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* z = reinterpret_cast<cplx*>(smem_raw);
    const int MP = (p.M + 1) / 2;
    const int c0 = 2 * (blockIdx.x % MP);
    const int t = blockIdx.x / MP;
    const int b = blockIdx.y;
    const bool two = c0 + 1 < p.M;
    const AT* xa = reinterpret_cast<const AT*>(p.x) + (size_t)b * p.sb + (size_t)c0 * p.sc;
    const AT* xb = xa + (two ? p.sc : 0);
    const long long n0 = p.first + (long long)t * p.hop;
    // all FFT_PTS samples of a thread are requested before the first one is used (blockDim >= L / FFT_PTS)
    double va[FFT_PTS], vb[FFT_PTS];
#pragma unroll
    for (int i = 0; i < FFT_PTS; ++i) {
        const int n = threadIdx.x + i * blockDim.x;
        const long long s = n0 + n;
        va[i] = 0.0;
        vb[i] = 0.0;
        if (n < p.L && s >= 0 && s < p.N) {
            if constexpr (VEC) {
                load_pair(xa + s * p.sn, va[i], vb[i]);
            } else {
                va[i] = (double)xa[s * p.sn];
                if (two) vb[i] = (double)xb[s * p.sn];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < FFT_PTS; ++i) {
        const int n = threadIdx.x + i * blockDim.x;
        if (n < p.L) {
            const double w = p.win ? __ldg(&p.win[n]) : 1.0;
            z[pidx(n)] = cmake(va[i] * w, vb[i] * w);
        }
    }
    __syncthreads();
    fft_forward(z, p.lg, p.tw);
    XC* out = reinterpret_cast<XC*>(p.out);
    const int kmax = p.grouped ? p.NG * OIVA_GROUP : p.F;
    for (int k = threadIdx.x; k < kmax; k += blockDim.x) {
        cplx Xa = cmake(0.0, 0.0), Xb = cmake(0.0, 0.0);
        if (k < p.F) {
            const cplx zk = z[pidx(k)];
            const cplx zm = cconj(z[pidx((p.L - k) & (p.L - 1))]);
            Xa = cscale(cadd(zk, zm), 0.5);
            Xb = mul_mi(cscale(csub(zk, zm), 0.5));
        }
        XC va, vb;
        narrow(va, Xa);
        narrow(vb, Xb);
        if (p.grouped) {
            const size_t gi = (size_t)b * p.NG + (k >> 5);
            XC* dst = out + ((gi * p.T + t) * p.M + c0) * OIVA_GROUP + (k & 31);
            dst[0] = va;
            if (two) dst[OIVA_GROUP] = vb;
        } else {
            XC* dst = out + (((size_t)b * p.T + t) * p.F + k) * p.M + c0;
            dst[0] = va;
            if (two) dst[1] = vb;
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

// grid (T * ceil(K/2), B): irfft of two sources through one complex FFT (inverse = conj o forward o conj)
template <typename ST, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? MINB : 1) k_stft_frames(const SynthesisParams p) {
    typedef typename StoreC<ST>::type XC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* z = reinterpret_cast<cplx*>(smem_raw);
    const int KP = (p.K + 1) / 2;
    const int k0 = 2 * (blockIdx.x % KP);
    const int t = blockIdx.x / KP;
    const int b = blockIdx.y;
    const bool two = k0 + 1 < p.K;
    const int half = p.L / 2;
    const XC* Y = reinterpret_cast<const XC*>(p.Y) + ((size_t)b * p.T + t) * p.F * p.K + k0;
    // Z[q] = Xa[q] + i Xb[q] (q <= L/2),  Z[L-q] = conj Xa[q] + i conj Xb[q];  the buffer receives conj Z
    constexpr int NQ = FFT_PTS / 2 + 1;  // ceil((L/2 + 1) / blockDim) for blockDim >= L / FFT_PTS
    cplx ya[NQ], yb[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int q = threadIdx.x + i * blockDim.x;
        ya[i] = cmake(0.0, 0.0);
        yb[i] = cmake(0.0, 0.0);
        if (q <= half) {
            ya[i] = widen(Y[(size_t)q * p.K]);
            if (two) yb[i] = widen(Y[(size_t)q * p.K + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int q = threadIdx.x + i * blockDim.x;
        if (q <= half) {
            cplx xa = ya[i], xb = yb[i];
            if (q == 0 || q == half) {  // numpy.fft.irfft ignores the imaginary parts of the DC and Nyquist bins
                xa.y = 0.0;
                xb.y = 0.0;
            }
            z[pidx(q)] = cmake(xa.x - xb.y, -(xa.y + xb.x));
            if (q != 0 && q != half) z[pidx(p.L - q)] = cmake(xa.x + xb.y, -(xb.x - xa.y));
        }
    }
    __syncthreads();
    fft_forward(z, p.lg, p.tw);
    const double scale = 1.0 / (double)p.L;
    double* fa = p.frames + (((size_t)b * p.T + t) * p.K + k0) * p.L;
    double* fb = fa + p.L;
    for (int n = threadIdx.x; n < p.L; n += blockDim.x) {
        const cplx v = z[pidx(n)];
        const double w = p.win ? scale * __ldg(&p.win[n]) : scale;
        fa[n] = v.x * w;
        if (two) fb[n] = -v.y * w;
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

extern "C" size_t oiva_stft_twiddle_bytes(int frame_len) {
    const int lg = log2_exact(frame_len);
    if (lg < 3 || frame_len > 8192) return 0;
    const size_t n = 3 * (size_t)twiddle_records(lg) * sizeof(cplx);
    return n ? n : sizeof(cplx);
}

extern "C" int oiva_stft_twiddles(void* tw, int frame_len, void* stream) {
    OIVA_REQUIRE(tw, "oiva_stft_twiddles: null pointer");
    const int lg = log2_exact(frame_len);
    OIVA_REQUIRE(lg >= 3 && frame_len <= 8192, "oiva_stft_twiddles: frame_len=%d must be a power of two in 8..8192",
                 frame_len);
    const int n = twiddle_records(lg);
    if (n > 0) {
        k_twiddles<<<oiva_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>((cplx*)tw, lg, n);
        OIVA_LAUNCH_CHECK();
    }
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
    p.lg = lg;
    p.hop = hop;
    p.F = frame_len / 2 + 1;
    p.NG = oiva_bin_groups(p.F);
    p.grouped = grouped;
    const size_t smem = stft_smem_bytes(frame_len);
    dim3 grid((unsigned)((long long)n_frames * ((n_chan + 1) / 2)), n_batch);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = stride_c == 1 && (n_chan % 2) == 0 && (stride_n % 2) == 0 && (stride_b % 2) == 0 &&
                     ((uintptr_t)x % (x_f32 ? 8 : 16)) == 0;
    // resident CTAs per SM for frame lengths <= 4096: 2 leaves ~90 KB of L1 for the twiddle / window tables,
    // 3 hides more barrier latency but shrinks L1 to its minimum (OIVA_STFT_CTAS=2|3)
    static const int ctas = [] {
        const char* v = getenv("OIVA_STFT_CTAS");
        return (v && *v == '3') ? 3 : 2;
    }();
#define LAUNCH_A3(AT, ST, MAXT, VEC, MINB)                                                                           \
    do {                                                                                                             \
        auto kern = k_stft_analysis<AT, ST, MAXT, VEC, MINB>;                                                        \
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                                             (int)stft_smem_bytes(MAXT * FFT_PTS)));                                 \
        if (MINB == 3)                                                                                               \
            OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));        \
        kern<<<grid, stft_threads(frame_len), smem, st>>>(p);                                                        \
    } while (0)
#define LAUNCH_A2(AT, ST, MAXT, VEC)                                                                                 \
    do {                                                                                                             \
        if (MAXT == 256 && ctas == 3) LAUNCH_A3(AT, ST, MAXT, VEC, 3); else LAUNCH_A3(AT, ST, MAXT, VEC, 2);         \
    } while (0)
#define LAUNCH_A(AT, ST)                                                                                             \
    do {                                                                                                             \
        if (frame_len > 4096) { if (vec) LAUNCH_A2(AT, ST, 512, true); else LAUNCH_A2(AT, ST, 512, false); }         \
        else { if (vec) LAUNCH_A2(AT, ST, 256, true); else LAUNCH_A2(AT, ST, 256, false); }                          \
    } while (0)
    if (x_f32) {
        if (dtype == OIVA_C64) LAUNCH_A(float, float); else LAUNCH_A(float, double);
    } else {
        if (dtype == OIVA_C64) LAUNCH_A(double, float); else LAUNCH_A(double, double);
    }
#undef LAUNCH_A
#undef LAUNCH_A2
#undef LAUNCH_A3
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
    p.lg = lg;
    p.F = frame_len / 2 + 1;
    const size_t smem = stft_smem_bytes(frame_len);
    dim3 grid((unsigned)((long long)n_frames * ((n_src + 1) / 2)), n_batch);
    cudaStream_t st = (cudaStream_t)stream;
    static const int ctas = [] {
        const char* v = getenv("OIVA_STFT_CTAS");
        return (v && *v == '3') ? 3 : 2;
    }();
#define LAUNCH_S2(ST, MAXT, MINB)                                                                            \
    do {                                                                                                     \
        auto kern = k_stft_frames<ST, MAXT, MINB>;                                                           \
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                             (int)stft_smem_bytes(MAXT * FFT_PTS)));                         \
        if (MINB == 3)                                                                                       \
            OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
        kern<<<grid, stft_threads(frame_len), smem, st>>>(p);                                                \
    } while (0)
#define LAUNCH_S(ST, MAXT)                                                                                   \
    do {                                                                                                     \
        if (MAXT == 256 && ctas == 3) LAUNCH_S2(ST, MAXT, 3); else LAUNCH_S2(ST, MAXT, 2);                   \
    } while (0)
    if (dtype == OIVA_C64) {
        if (frame_len > 4096) LAUNCH_S(float, 512); else LAUNCH_S(float, 256);
    } else {
        if (frame_len > 4096) LAUNCH_S(double, 512); else LAUNCH_S(double, 256);
    }
#undef LAUNCH_S
#undef LAUNCH_S2
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
