// ILRMA (independent low-rank matrix analysis): the NMF source model around the shared demix / covariance / IP-sweep
// kernels.  Replaces pyroomacoustics.bss.ilrma as called by the reference drivers (overiva_oneshot.py:331-339,
// overiva_sim.py:309-311; third-party, absent from the reference tree: restated from the published algorithm --
// Kitamura et al., "Determined blind source separation unifying independent vector analysis and nonnegative matrix
// factorization", 2016 -- in oracle/ilrma_oracle.py, parity unpinned).
//
// Source k has the spectrogram model r_k(f, t) = sum_l T_k(f, l) V_k(t, l) (L = n_components).  Per epoch, with
// P_k(f, t) = |y_k(f, t)|^2 of the previous demix:
//     T_k(f, l) *= sqrt( sum_t P r^-2 V_k(t, l) / sum_t r^-1 V_k(t, l) ),  clamp, r = T V^T
//     V_k(t, l) *= sqrt( sum_f P r^-2 T_k(f, l) / sum_f r^-1 T_k(f, l) ),  clamp, r = T V^T
//     V_k[f] = (1/T) sum_t x x^H / r_k(f, t)   (oiva_weighted_cov_binwise),  w_k = (W^H V_k)^-1 e_k, normalised   (oiva_ip_update)
// then y = W^H x, P = |y|^2, lambda_k = 1 / sqrt(mean P_k), and w_k *= lambda_k, P_k, r_k, T_k *= lambda_k^2.
//
// Layout: everything per bin is grouped like the samples (lane <-> bin): Pg, iRg [gi][k][Tp][32] (iR = 1 / r, what the
// covariance kernel consumes; zero on the padded bins of a mixture's last group), Tg [gi][k][L][32]; the activations
// Vn [b][k][Tp][L] are small and plain.  The only cross-bin step is the V update (sums over f): per-group partial sums
// through a warp butterfly, then a fixed-order sum over the groups (deterministic).
#include "common.cuh"

namespace oiva {

constexpr int ILRMA_MAX_L = 8;

// warp = (group gi, source k): T update for the lane's bin, then r and iR for every frame of the bin
__global__ void __launch_bounds__(128) k_ilrma_update_T(const double* __restrict__ Pg, double* __restrict__ iRg,
                                                        double* __restrict__ Tg, const double* __restrict__ Vn, long long GK,
                                                        int NG, int K, int T, int Tp, int F, int L, double eps) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= GK) return;
    const long long gi = w / K;
    const int k = (int)(w - gi * K);
    const long long b = gi / NG;
    const bool bin_ok = (int)(gi - b * NG) * OIVA_GROUP + lane < F;
    const double* P = Pg + (size_t)w * Tp * OIVA_GROUP + lane;
    double* iR = iRg + (size_t)w * Tp * OIVA_GROUP + lane;
    double* Tl = Tg + (size_t)w * L * OIVA_GROUP + lane;
    const double* V = Vn + ((size_t)b * K + k) * Tp * L;
    double num[ILRMA_MAX_L], den[ILRMA_MAX_L], tv[ILRMA_MAX_L];
#pragma unroll
    for (int l = 0; l < ILRMA_MAX_L; ++l) num[l] = den[l] = 0.0;
    // (unrolled: the loads of several frames are in flight at once -- one mixture has ~100 warps here, pure latency)
#pragma unroll 8
    for (int t = 0; t < T; ++t) {
        const double p = P[(size_t)t * OIVA_GROUP], ir = iR[(size_t)t * OIVA_GROUP];
        const double a = p * ir * ir;
#pragma unroll
        for (int l = 0; l < ILRMA_MAX_L; ++l)
            if (l < L) {
                const double v = V[(size_t)t * L + l];
                num[l] = fma(a, v, num[l]);
                den[l] = fma(ir, v, den[l]);
            }
    }
#pragma unroll
    for (int l = 0; l < ILRMA_MAX_L; ++l)
        if (l < L) {
            double x = Tl[(size_t)l * OIVA_GROUP] * sqrt(num[l] / den[l]);
            if (x < eps) x = eps;  // (NaN stays NaN, as numpy's T[T < eps] = eps)
            tv[l] = bin_ok ? x : 0.0;
            Tl[(size_t)l * OIVA_GROUP] = tv[l];
        }
    // (unrolled: the loads of several frames are in flight at once -- one mixture has ~100 warps here, pure latency)
#pragma unroll 8
    for (int t = 0; t < T; ++t) {
        double r = 0.0;
#pragma unroll
        for (int l = 0; l < ILRMA_MAX_L; ++l)
            if (l < L) r = fma(tv[l], V[(size_t)t * L + l], r);
        iR[(size_t)t * OIVA_GROUP] = bin_ok ? 1.0 / r : 0.0;
    }
}

// warp = (group gi, source k): per frame the sums over the group's 32 bins of P r^-2 T_l and r^-1 T_l -> Vpart[w][t][2L]
__global__ void __launch_bounds__(128) k_ilrma_V_partial(const double* __restrict__ Pg, const double* __restrict__ iRg,
                                                         const double* __restrict__ Tg, double* __restrict__ Vpart,
                                                         long long GK, int T, int Tp, int L) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= GK) return;
    const double* P = Pg + (size_t)w * Tp * OIVA_GROUP + lane;
    const double* iR = iRg + (size_t)w * Tp * OIVA_GROUP + lane;
    double tv[ILRMA_MAX_L];
#pragma unroll
    for (int l = 0; l < ILRMA_MAX_L; ++l) tv[l] = l < L ? Tg[((size_t)w * L + l) * OIVA_GROUP + lane] : 0.0;
    double* out = Vpart + (size_t)w * Tp * 2 * L;
    // (unrolled: the loads of several frames are in flight at once -- one mixture has ~100 warps here, pure latency)
#pragma unroll 8
    for (int t = 0; t < T; ++t) {
        const double p = P[(size_t)t * OIVA_GROUP], ir = iR[(size_t)t * OIVA_GROUP];
        const double a = p * ir * ir;
#pragma unroll
        for (int l = 0; l < ILRMA_MAX_L; ++l)
            if (l < L) {
                double x = a * tv[l], y = ir * tv[l];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    x += __shfl_xor_sync(0xffffffffu, x, off);
                    y += __shfl_xor_sync(0xffffffffu, y, off);
                }
                if (lane == 0) {
                    out[((size_t)t * L + l) * 2] = x;
                    out[((size_t)t * L + l) * 2 + 1] = y;
                }
            }
    }
}

// thread = (mixture b, source k, frame t, component l): fixed-order sum over the NG groups, V update, clamp
__global__ void k_ilrma_V_finish(const double* __restrict__ Vpart, double* __restrict__ Vn, long long n, int NG, int K, int T,
                                 int Tp, int L, double eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = (int)(i % L);
    const long long r1 = i / L;
    const int t = (int)(r1 % T);
    const long long bk = r1 / T;
    const long long b = bk / K;
    const int k = (int)(bk - b * K);
    double num = 0.0, den = 0.0;
    for (int g = 0; g < NG; ++g) {
        const double* src = Vpart + ((((size_t)b * NG + g) * K + k) * Tp + t) * 2 * L + 2 * l;
        num += src[0];
        den += src[1];
    }
    double* v = Vn + (((size_t)b * K + k) * Tp + t) * L + l;
    double x = *v * sqrt(num / den);
    if (x < eps) x = eps;
    *v = x;
}

// warp = (group gi, source k): iR = 1 / (T V^T) for every frame of the lane's bin
__global__ void __launch_bounds__(128) k_ilrma_model(double* __restrict__ iRg, const double* __restrict__ Tg,
                                                     const double* __restrict__ Vn, long long GK, int NG, int K, int T, int Tp,
                                                     int F, int L) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= GK) return;
    const long long gi = w / K;
    const int k = (int)(w - gi * K);
    const long long b = gi / NG;
    const bool bin_ok = (int)(gi - b * NG) * OIVA_GROUP + lane < F;
    double* iR = iRg + (size_t)w * Tp * OIVA_GROUP + lane;
    const double* V = Vn + ((size_t)b * K + k) * Tp * L;
    double tv[ILRMA_MAX_L];
#pragma unroll
    for (int l = 0; l < ILRMA_MAX_L; ++l) tv[l] = l < L ? Tg[((size_t)w * L + l) * OIVA_GROUP + lane] : 0.0;
#pragma unroll 8
    for (int t = 0; t < Tp; ++t) {
        double r = 0.0;
#pragma unroll
        for (int l = 0; l < ILRMA_MAX_L; ++l)
            if (l < L) r = fma(tv[l], V[(size_t)t * L + l], r);
        iR[(size_t)t * OIVA_GROUP] = (bin_ok && t < T) ? 1.0 / r : 0.0;
    }
}

// T0 (B, K, F, L) row-major -> Tg [gi][k][L][32] (zero on padded bins); thread per element of Tg
__global__ void k_ilrma_group_T(const double* __restrict__ T0, double* __restrict__ Tg, long long n, int NG, int K, int F,
                                int L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lane = (int)(i % OIVA_GROUP);
    long long r = i / OIVA_GROUP;
    const int l = (int)(r % L);
    r /= L;
    const int k = (int)(r % K);
    const long long gi = r / K;
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + lane;
    Tg[i] = f < F ? T0[(((size_t)b * K + k) * F + f) * L + l] : 0.0;
}
// Tg -> (B, K, F, L) row-major
__global__ void k_ilrma_ungroup_T(const double* __restrict__ Tg, double* __restrict__ Tout, long long n, int NG, int K, int F,
                                  int L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lane = (int)(i % OIVA_GROUP);
    long long r = i / OIVA_GROUP;
    const int l = (int)(r % L);
    r /= L;
    const int k = (int)(r % K);
    const long long gi = r / K;
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + lane;
    if (f < F) Tout[(((size_t)b * K + k) * F + f) * L + l] = Tg[i];
}

// one CTA per (mixture, source): lambda = 1 / sqrt( sum_{g, t} r2part / (F T) ), sums in a fixed order
__global__ void __launch_bounds__(256) k_ilrma_lambda(const double* __restrict__ r2part, double* __restrict__ lam, int NG,
                                                      int K, int T, int Tp, int F) {
    __shared__ double red[256];
    const int b = blockIdx.x / K, k = blockIdx.x - b * K;
    double s = 0.0;
    for (int t = threadIdx.x; t < T; t += 256) {
        double st = 0.0;
        for (int g = 0; g < NG; ++g) st += r2part[(((size_t)b * NG + g) * K + k) * Tp + t];
        s += st;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) lam[blockIdx.x] = 1.0 / sqrt(red[0] / ((double)F * (double)T));
}

// warp = (group gi, source k): w_k *= lambda (scale_w), P_k, T_k *= lambda^2, iR_k /= lambda^2
__global__ void __launch_bounds__(128) k_ilrma_rescale(const double* __restrict__ lam, cplx* __restrict__ Wg,
                                                       double* __restrict__ Pg, double* __restrict__ iRg,
                                                       double* __restrict__ Tg, long long GK, int NG, int K, int M, int T,
                                                       int Tp, int L, int scale_w) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= GK) return;
    const long long gi = w / K;
    const int k = (int)(w - gi * K);
    const long long b = gi / NG;
    const double la = lam[b * K + k], la2 = la * la, ila2 = 1.0 / la2;
    if (scale_w) {
        cplx* Wl = Wg + (size_t)gi * M * M * OIVA_GROUP + lane;
        for (int c = 0; c < M; ++c) {
            cplx v = Wl[(size_t)(c * M + k) * OIVA_GROUP];
            Wl[(size_t)(c * M + k) * OIVA_GROUP] = cscale(v, la);
        }
    }
    double* P = Pg + (size_t)w * Tp * OIVA_GROUP + lane;
    double* iR = iRg + (size_t)w * Tp * OIVA_GROUP + lane;
    // (unrolled: the loads of several frames are in flight at once -- one mixture has ~100 warps here, pure latency)
#pragma unroll 8
    for (int t = 0; t < T; ++t) {
        P[(size_t)t * OIVA_GROUP] *= la2;
        iR[(size_t)t * OIVA_GROUP] *= ila2;
    }
    for (int l = 0; l < L; ++l) Tg[((size_t)w * L + l) * OIVA_GROUP + lane] *= la2;
}

// Zg[gi][k][lane] = (scale[b][k], 0): a real per-source scale for the output kernel (the last demix of the reference
// predates the last lambda normalisation of W: y = (w_k / lambda_k)^H x)
__global__ void k_ilrma_fill_scale(const double* __restrict__ lam, cplx* __restrict__ Zg, long long n, int NG, int K,
                                   int invert) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = i / OIVA_GROUP;
    const int k = (int)(r % K);
    const long long b = (r / K) / NG;
    const double la = lam[b * K + k];
    Zg[i] = cmake(invert ? 1.0 / la : la, 0.0);
}

}  // namespace oiva

using namespace oiva;

static int ilrma_check(const char* who, int B, int T, int F, int K, int L) {
    OIVA_REQUIRE(B > 0 && T > 0 && F > 0 && K >= 1 && K <= OIVA_MAX_M, "%s: bad shape B=%d T=%d F=%d K=%d", who, B, T, F, K);
    OIVA_REQUIRE(L >= 1 && L <= ILRMA_MAX_L, "%s: n_components=%d not in 1..%d", who, L, ILRMA_MAX_L);
    return OIVA_OK;
}

extern "C" size_t oiva_ilrma_vpart_bytes(int n_batch, int n_frames, int n_freq, int n_src, int n_comp) {
    return (size_t)n_batch * oiva_bin_groups(n_freq) * n_src * oiva_frame_pitch(n_frames) * 2 * n_comp * sizeof(double);
}

extern "C" int oiva_ilrma_set_model(const double* T0, const double* V0, double* Tg, double* Vn, double* iRg, int n_batch,
                                    int n_frames, int n_freq, int n_src, int n_comp, void* stream) {
    OIVA_REQUIRE(T0 && V0 && Tg && Vn && iRg, "oiva_ilrma_set_model: null pointer");
    int rc = ilrma_check("oiva_ilrma_set_model", n_batch, n_frames, n_freq, n_src, n_comp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int NG = oiva_bin_groups(n_freq), Tp = oiva_frame_pitch(n_frames);
    const long long GK = (long long)n_batch * NG * n_src;
    const long long nT = GK * n_comp * OIVA_GROUP;
    k_ilrma_group_T<<<(unsigned)((nT + 255) / 256), 256, 0, st>>>(T0, Tg, nT, NG, n_src, n_freq, n_comp);
    OIVA_LAUNCH_CHECK();
    // V0 (B, K, T, L) -> Vn (B, K, Tp, L), padding frames zero
    OIVA_CUDA_CHECK(cudaMemsetAsync(Vn, 0, (size_t)n_batch * n_src * Tp * n_comp * sizeof(double), st));
    OIVA_CUDA_CHECK(cudaMemcpy2DAsync(Vn, (size_t)Tp * n_comp * sizeof(double), V0, (size_t)n_frames * n_comp * sizeof(double),
                                      (size_t)n_frames * n_comp * sizeof(double), (size_t)n_batch * n_src,
                                      cudaMemcpyDeviceToDevice, st));
    k_ilrma_model<<<(unsigned)((GK + 3) / 4), 128, 0, st>>>(iRg, Tg, Vn, GK, NG, n_src, n_frames, Tp, n_freq, n_comp);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ilrma_get_model(const double* Tg, const double* Vn, double* Tout, double* Vout, int n_batch, int n_frames,
                                    int n_freq, int n_src, int n_comp, void* stream) {
    OIVA_REQUIRE(Tg && Vn && Tout && Vout, "oiva_ilrma_get_model: null pointer");
    int rc = ilrma_check("oiva_ilrma_get_model", n_batch, n_frames, n_freq, n_src, n_comp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int NG = oiva_bin_groups(n_freq), Tp = oiva_frame_pitch(n_frames);
    const long long nT = (long long)n_batch * NG * n_src * n_comp * OIVA_GROUP;
    k_ilrma_ungroup_T<<<(unsigned)((nT + 255) / 256), 256, 0, st>>>(Tg, Tout, nT, NG, n_src, n_freq, n_comp);
    OIVA_LAUNCH_CHECK();
    OIVA_CUDA_CHECK(cudaMemcpy2DAsync(Vout, (size_t)n_frames * n_comp * sizeof(double), Vn, (size_t)Tp * n_comp * sizeof(double),
                                      (size_t)n_frames * n_comp * sizeof(double), (size_t)n_batch * n_src,
                                      cudaMemcpyDeviceToDevice, st));
    return OIVA_OK;
}

extern "C" int oiva_ilrma_nmf(const double* Pg, double* iRg, double* Tg, double* Vn, double* Vpart, int n_batch, int n_frames,
                              int n_freq, int n_src, int n_comp, double eps, void* stream) {
    OIVA_REQUIRE(Pg && iRg && Tg && Vn && Vpart, "oiva_ilrma_nmf: null pointer");
    int rc = ilrma_check("oiva_ilrma_nmf", n_batch, n_frames, n_freq, n_src, n_comp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int NG = oiva_bin_groups(n_freq), Tp = oiva_frame_pitch(n_frames);
    const long long GK = (long long)n_batch * NG * n_src;
    const unsigned gw = (unsigned)((GK + 3) / 4);
    k_ilrma_update_T<<<gw, 128, 0, st>>>(Pg, iRg, Tg, Vn, GK, NG, n_src, n_frames, Tp, n_freq, n_comp, eps);
    OIVA_LAUNCH_CHECK();
    k_ilrma_V_partial<<<gw, 128, 0, st>>>(Pg, iRg, Tg, Vpart, GK, n_frames, Tp, n_comp);
    OIVA_LAUNCH_CHECK();
    const long long nV = (long long)n_batch * n_src * n_frames * n_comp;
    k_ilrma_V_finish<<<(unsigned)((nV + 127) / 128), 128, 0, st>>>(Vpart, Vn, nV, NG, n_src, n_frames, Tp, n_comp, eps);
    OIVA_LAUNCH_CHECK();
    k_ilrma_model<<<gw, 128, 0, st>>>(iRg, Tg, Vn, GK, NG, n_src, n_frames, Tp, n_freq, n_comp);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ilrma_rescale(const double* r2part, double* lam, void* Wg, double* Pg, double* iRg, double* Tg,
                                  int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int n_comp, int scale_w,
                                  void* stream) {
    OIVA_REQUIRE(r2part && lam && Wg && Pg && iRg && Tg, "oiva_ilrma_rescale: null pointer");
    int rc = ilrma_check("oiva_ilrma_rescale", n_batch, n_frames, n_freq, n_src, n_comp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int NG = oiva_bin_groups(n_freq), Tp = oiva_frame_pitch(n_frames);
    const long long GK = (long long)n_batch * NG * n_src;
    k_ilrma_lambda<<<(unsigned)(n_batch * n_src), 256, 0, st>>>(r2part, lam, NG, n_src, n_frames, Tp, n_freq);
    OIVA_LAUNCH_CHECK();
    k_ilrma_rescale<<<(unsigned)((GK + 3) / 4), 128, 0, st>>>(lam, (cplx*)Wg, Pg, iRg, Tg, GK, NG, n_src, n_chan, n_frames,
                                                               Tp, n_comp, scale_w);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_ilrma_fill_scale(const double* lam, void* Zg, int n_batch, int n_freq, int n_src, int invert,
                                     void* stream) {
    OIVA_REQUIRE(lam && Zg, "oiva_ilrma_fill_scale: null pointer");
    const int NG = oiva_bin_groups(n_freq);
    const long long n = (long long)n_batch * NG * n_src * OIVA_GROUP;
    k_ilrma_fill_scale<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lam, (cplx*)Zg, n, NG, n_src, invert);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// n_epochs whole epochs in one library call (the Python loop cost ~7 ctypes round trips per epoch, of the order of the
// GPU work itself for one mixture): NMF updates -> covariance with per-bin weights -> determined IP sweep -> demix with
// per-bin powers -> scale normalisation.  Pointers as for the individual calls; C (row-major) / Cg as for oiva_ip_update.
extern "C" int oiva_ilrma_iterate(const void* Xg, void* Wg, void* Vg, const void* C, const void* Cg, double* r2part,
                                  void* scratch, size_t scratch_bytes, double* Pg, double* iRg, double* Tg, double* Vn,
                                  double* Vpart, double* lam, int* status, int n_batch, int n_frames, int n_freq, int n_chan,
                                  int n_comp, double eps, int n_epochs, void* stream) {
    OIVA_REQUIRE(Xg && Wg && Vg && C && Cg && r2part && Pg && iRg && Tg && Vn && Vpart && lam && status,
                 "oiva_ilrma_iterate: null pointer");
    const int K = n_chan;
    for (int e = 0; e < n_epochs; ++e) {
        int rc = oiva_ilrma_nmf(Pg, iRg, Tg, Vn, Vpart, n_batch, n_frames, n_freq, K, n_comp, eps, stream);
        if (rc) return rc;
        rc = oiva_weighted_cov_binwise(Xg, iRg, Vg, scratch, scratch_bytes, n_batch, n_frames, n_freq, n_chan, K, stream);
        if (rc) return rc;
        rc = oiva_ip_update(Wg, Vg, C, Cg, nullptr, status, n_batch, n_freq, n_chan, K, stream);
        if (rc) return rc;
        rc = oiva_demix_power_full(Xg, Wg, n_chan, 1, r2part, Pg, n_batch, n_frames, n_freq, n_chan, K, OIVA_C128, stream);
        if (rc) return rc;
        rc = oiva_ilrma_rescale(r2part, lam, Wg, Pg, iRg, Tg, n_batch, n_frames, n_freq, n_chan, K, n_comp, 1, stream);
        if (rc) return rc;
    }
    return OIVA_OK;
}
