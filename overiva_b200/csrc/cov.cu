// C entry point of the weighted-covariance kernel (include/overiva_b200.h: oiva_weighted_cov).
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "cov.cuh"
#include "relayout_cov.cuh"

namespace oiva {
#define OIVA_DECL(M)                                                                                 \
    int cov_launch_m##M(int dtype, int KC, const CovParams& p, cudaStream_t st, int* nsplit_out); \
    int cov_max_kc_m##M();                                                                                     \
    int cov_launch_tiled_m##M(int dtype, const CovParams& p, cudaStream_t st, int* nsplit_out);                \
    int cov_launch_wbin_m##M(int KC, const CovParams& p, cudaStream_t st, int* nsplit_out);                   \
    int cov_sweep_launch_m##M(int dtype, int K, const CovParams& p, cudaStream_t st);                          \
    int relayout_cov_launch_m##M(int dtype, RelayoutCovParams p, int max_split, cudaStream_t st, int* nsplit_out);
OIVA_DECL(1) OIVA_DECL(2) OIVA_DECL(3) OIVA_DECL(4) OIVA_DECL(5) OIVA_DECL(6) OIVA_DECL(7) OIVA_DECL(8)
OIVA_DECL(9) OIVA_DECL(10) OIVA_DECL(11) OIVA_DECL(12) OIVA_DECL(13) OIVA_DECL(14) OIVA_DECL(15) OIVA_DECL(16)
#undef OIVA_DECL

// phi == NULL: a device buffer of ones (grown on demand, per device) stands in for the weights.  The buffer is
// created under a mutex and filled SYNCHRONOUSLY (blocking cudaMemcpy from a host vector) before its pointer is
// published, so any stream / thread that sees it sees ones; growing it must not happen under stream capture
// (plans size it once at creation through oiva_reserve_ones).
static std::mutex g_ones_mu;
static double* g_ones[64] = {nullptr};
static size_t g_ones_n[64] = {0};
static int get_ones(size_t n, const double** out) {
    int dev = 0;
    OIVA_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(g_ones_mu);
    if (g_ones_n[dev] < n) {
        // a previous, smaller buffer may still be in use by queued kernels: keep it alive (tiny leak by design)
        size_t cap = n < 4096 ? 4096 : n;
        double* q = nullptr;
        OIVA_CUDA_CHECK(cudaMalloc(&q, cap * sizeof(double)));
        std::vector<double> host(cap, 1.0);
        OIVA_CUDA_CHECK(cudaMemcpy(q, host.data(), cap * sizeof(double), cudaMemcpyHostToDevice));
        g_ones[dev] = q;
        g_ones_n[dev] = cap;
    }
    *out = g_ones[dev];
    return OIVA_OK;
}

// Vg[i] = sum_sp Vpart[sp][i], sp ascending: the deterministic combination of the frame-split partial sums
__global__ void k_cov_sum_partials(const cplx* __restrict__ part, cplx* __restrict__ Vg, int nsplit, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx s = part[i];
    for (int sp = 1; sp < nsplit; ++sp) {
        const cplx v = part[(size_t)sp * n + i];
        s.x += v.x;
        s.y += v.y;
    }
    Vg[i] = s;
}

static int pick_chunk(int rem, int max_kc) {
    static const int sizes[] = {1, 2, 3, 4, 6, 8};
    int best = 1;
    for (int s : sizes) {
        if (s > max_kc) break;
        if (s >= rem) return s;  // smallest chunk that covers the remainder in one pass
        best = s;
    }
    return best;  // largest usable chunk; more passes follow
}
}  // namespace oiva

// make sure the ones buffer of the current device covers n_frames (called at plan creation, outside any capture)
int oiva_reserve_ones(int n_frames) {
    const double* ones = nullptr;
    return oiva::get_ones((size_t)oiva_frame_pitch(n_frames), &ones);
}

extern "C" size_t oiva_weighted_cov_scratch_bytes(int n_batch, int n_frames, int n_freq, int n_chan, int n_src) {
    // frame splitting only happens when there are few bin groups; 64 slots or 512 MiB, whichever is smaller, keeps
    // every SM busy for the shapes that need it (the split count is clamped to the slots available)
    const size_t vg = oiva_grouped_cov_bytes(n_batch, n_freq, n_chan, n_src);
    if (vg == 0 || n_frames <= 0) return 0;
    const long long G = (long long)n_batch * oiva_bin_groups(n_freq);
    if (G >= 4096) return 0;  // enough groups for every team of every SM: never split
    size_t slots = 64;
    const size_t cap = (size_t)512 << 20;
    if (slots * vg > cap) slots = cap / vg;
    return slots < 2 ? 0 : slots * vg;
}

static int weighted_cov_impl(const void* Xg, const double* phi, bool wbin, void* Vg, void* scratch, size_t scratch_bytes,
                             int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream);

extern "C" int oiva_weighted_cov_ws(const void* Xg, const double* phi, void* Vg, void* scratch, size_t scratch_bytes,
                                    int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype,
                                    void* stream) {
    return weighted_cov_impl(Xg, phi, false, Vg, scratch, scratch_bytes, n_batch, n_frames, n_freq, n_chan, n_src, dtype,
                             stream);
}

// Weights per (source, frame, BIN): winv is a grouped array [gi][k][Tp][32] of float64 (Tp = oiva_frame_pitch(T); the
// padding frames are never read).  complex128 samples, M <= 8.  V_k[f] = (1/T) sum_t winv_k(f, t) x(f, t) x(f, t)^H --
// the auxiliary variable of ILRMA (pyroomacoustics.bss.ilrma as called at overiva_oneshot.py:331-339), whose source
// model r_k(f, t) is a low-rank spectrogram instead of overiva's r_k(t).
extern "C" int oiva_weighted_cov_binwise(const void* Xg, const double* winv, void* Vg, void* scratch, size_t scratch_bytes,
                                         int n_batch, int n_frames, int n_freq, int n_chan, int n_src, void* stream) {
    OIVA_REQUIRE(winv, "oiva_weighted_cov_binwise: null weights");
    if (n_chan > 8) {
        oiva_set_error("oiva_weighted_cov_binwise: implemented for n_chan <= 8 (got %d)", n_chan);
        return OIVA_ERR_UNSUPPORTED;
    }
    return weighted_cov_impl(Xg, winv, true, Vg, scratch, scratch_bytes, n_batch, n_frames, n_freq, n_chan, n_src,
                             OIVA_C128, stream);
}

static int weighted_cov_impl(const void* Xg, const double* phi, bool wbin, void* Vg, void* scratch, size_t scratch_bytes,
                             int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype, void* stream) {
    using namespace oiva;
    OIVA_REQUIRE(Xg && Vg, "oiva_weighted_cov: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M && n_src >= 1,
                 "oiva_weighted_cov: bad shape B=%d T=%d F=%d M=%d K=%d", n_batch, n_frames, n_freq, n_chan, n_src);
    OIVA_REQUIRE(phi || n_src == 1, "oiva_weighted_cov: phi == NULL needs n_src == 1");
    cudaStream_t st = (cudaStream_t)stream;
    CovParams p;
    p.Xg = Xg;
    p.Vg = (cplx*)Vg;
    p.L = oiva_make_layout(n_frames, n_freq, n_chan);
    p.G = (long long)n_batch * p.L.NG;
    p.NGphi = p.L.NG;
    p.K = n_src;
    p.invT = 1.0 / (double)n_frames;
    p.nsplit = 0;  // chosen by the first pass
    p.stages = 4;
    const size_t vg_bytes = oiva_grouped_cov_bytes(n_batch, n_freq, n_chan, n_src);
    p.Vpart = nullptr;
    p.max_split = 1;
    p.Wg = nullptr;
    p.Cg = nullptr;
    p.wscale = nullptr;
    p.status = nullptr;
    if (scratch && scratch_bytes >= 2 * vg_bytes) {
        p.Vpart = (cplx*)scratch;
        const size_t slots = scratch_bytes / vg_bytes;
        p.max_split = slots > 4096 ? 4096 : (int)slots;
    }
    if (phi) {
        p.phi = phi;
    } else {
        const double* ones = nullptr;
        int rc = get_ones((size_t)p.L.frame_pitch(), &ones);
        if (rc) return rc;
        p.phi = ones;
        OIVA_REQUIRE(p.G < (1ll << 31), "oiva_weighted_cov: too many groups");
        p.NGphi = (int)p.G;  // every group maps to "mixture 0": the ones buffer holds one (K=1, Tp) block
    }
    int max_kc = 1;
    switch (n_chan) {
#define OIVA_CASE(M) case M: max_kc = cov_max_kc_m##M(); break;
        OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
        OIVA_CASE(9) OIVA_CASE(10) OIVA_CASE(11) OIVA_CASE(12) OIVA_CASE(13) OIVA_CASE(14) OIVA_CASE(15) OIVA_CASE(16)
#undef OIVA_CASE
    }
    // many channels and >= 3 sources left: the tiled kernel takes 4 sources per pass (one pass over X for config 5's
    // K = 4; cov.cuh).  It combines frame splits through scratch slots only.
    static const bool no_tiled = [] {
        const char* v = getenv("OIVA_COV_NO_TILED");
        return v && *v && *v != '0';
    }();
    int k0 = 0;
    while (k0 < n_src) {
        const bool tiled =
            !wbin && !no_tiled && n_chan >= 9 && n_src - k0 >= 3 && (p.Vpart || p.nsplit == 1 || p.nsplit <= 0);
        int KC = tiled ? 4 : pick_chunk(n_src - k0, max_kc);
        p.k0 = k0;
        int rc = OIVA_ERR_INVALID;
        int nsplit_used = 0;
        if (tiled) {
            CovParams q = p;
            if (!q.Vpart) q.nsplit = 1;  // no scratch: no frame splits
            switch (n_chan) {
#define OIVA_CASE(M) case M: rc = cov_launch_tiled_m##M(dtype, q, st, &nsplit_used); break;
                OIVA_CASE(9) OIVA_CASE(10) OIVA_CASE(11) OIVA_CASE(12) OIVA_CASE(13) OIVA_CASE(14) OIVA_CASE(15)
                OIVA_CASE(16)
#undef OIVA_CASE
            }
        } else if (wbin) {
            switch (n_chan) {
#define OIVA_CASE(M) case M: rc = cov_launch_wbin_m##M(KC, p, st, &nsplit_used); break;
                OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
#undef OIVA_CASE
            }
        } else {
            switch (n_chan) {
#define OIVA_CASE(M) case M: rc = cov_launch_m##M(dtype, KC, p, st, &nsplit_used); break;
                OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
                OIVA_CASE(9) OIVA_CASE(10) OIVA_CASE(11) OIVA_CASE(12) OIVA_CASE(13) OIVA_CASE(14) OIVA_CASE(15)
                OIVA_CASE(16)
#undef OIVA_CASE
            }
        }
        if (rc) return rc;
        p.nsplit = nsplit_used;  // later passes keep the first pass's split
        k0 += KC;
    }
    if (p.Vpart && p.nsplit > 1) {
        const size_t n = vg_bytes / sizeof(cplx);
        k_cov_sum_partials<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.Vpart, p.Vg, p.nsplit, n);
        OIVA_LAUNCH_CHECK();
    }
    return OIVA_OK;
}

// One pass over X that ends, per bin group, in the IP sweep of that group (cov_sweep.cuh).  Covers the shapes whose K
// lower triangles fit one lane's registers and inputs with enough bin groups that no group is split over teams;
// everything else returns OIVA_ERR_UNSUPPORTED (no error text) and the caller runs oiva_weighted_cov + oiva_ip_update.
extern "C" int oiva_cov_ip_update(const void* Xg, const double* phi, void* Wg, const void* Cg, const double* wscale,
                                  int* status, int n_batch, int n_frames, int n_freq, int n_chan, int n_src, int dtype,
                                  void* stream) {
    using namespace oiva;
    OIVA_REQUIRE(Xg && phi && Wg && Cg && status, "oiva_cov_ip_update: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M && n_src >= 1 &&
                     n_src <= n_chan,
                 "oiva_cov_ip_update: bad shape B=%d T=%d F=%d M=%d K=%d", n_batch, n_frames, n_freq, n_chan, n_src);
    OIVA_REQUIRE(dtype == OIVA_C64 || dtype == OIVA_C128, "oiva_cov_ip_update: bad dtype %d", dtype);
    CovParams p;
    p.Xg = Xg;
    p.phi = phi;
    p.Vg = nullptr;
    p.Vpart = nullptr;
    p.L = oiva_make_layout(n_frames, n_freq, n_chan);
    p.G = (long long)n_batch * p.L.NG;
    p.NGphi = p.L.NG;
    p.K = n_src;
    p.k0 = 0;
    p.nsplit = 1;
    p.max_split = 1;
    p.stages = 2;
    p.invT = 1.0 / (double)n_frames;
    p.Wg = (cplx*)Wg;
    p.Cg = (const cplx*)Cg;
    p.wscale = wscale;
    p.status = status;
    if (n_chan > 8 || p.G < 4096) return OIVA_ERR_UNSUPPORTED;  // (few groups: the frames of a group are split)
    cudaStream_t st = (cudaStream_t)stream;
    switch (n_chan) {
#define OIVA_CASE(M) case M: return cov_sweep_launch_m##M(dtype, n_src, p, st);
        OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
#undef OIVA_CASE
    }
    return OIVA_ERR_UNSUPPORTED;
}

extern "C" int oiva_weighted_cov(const void* Xg, const double* phi, void* Vg, int n_batch, int n_frames, int n_freq,
                                 int n_chan, int n_src, int dtype, void* stream) {
    return oiva_weighted_cov_ws(Xg, phi, Vg, nullptr, 0, n_batch, n_frames, n_freq, n_chan, n_src, dtype, stream);
}

extern "C" int oiva_relayout_cov_supported(int n_freq, int n_chan, int dtype) {
    if (n_chan < 1 || n_chan > 8 || n_freq < 1) return 0;
    // the per-frame bulk copies need 16-byte aligned rows: always true for complex128; for complex64 (8-byte
    // elements) only when F * M is even
    if (dtype == OIVA_C64 && (((long long)n_freq * n_chan) & 1)) return 0;
    return dtype == OIVA_C64 || dtype == OIVA_C128;
}

extern "C" int oiva_relayout_cov(const void* X, void* Xg, void* Cg, void* scratch, size_t scratch_bytes, int n_batch,
                                 int n_frames, int n_freq, int n_chan, int dtype, void* stream) {
    using namespace oiva;
    OIVA_REQUIRE(X && Xg && Cg, "oiva_relayout_cov: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0, "oiva_relayout_cov: bad shape B=%d T=%d F=%d", n_batch,
                 n_frames, n_freq);
    OIVA_REQUIRE(oiva_relayout_cov_supported(n_freq, n_chan, dtype), "oiva_relayout_cov: unsupported (M=%d, F=%d, dtype=%d)",
                 n_chan, n_freq, dtype);
    OIVA_REQUIRE(((uintptr_t)X & 15) == 0, "oiva_relayout_cov: X must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    RelayoutCovParams p;
    p.X = X;
    p.Xg = Xg;
    p.Cg = (cplx*)Cg;
    p.L = oiva_make_layout(n_frames, n_freq, n_chan);
    p.G = (long long)n_batch * p.L.NG;
    p.invT = 1.0 / (double)n_frames;
    p.nsplit = 1;
    p.stages = 2;
    const size_t cg_bytes = oiva_grouped_cov_bytes(n_batch, n_freq, n_chan, 1);
    int max_split = 1;
    p.Cpart = nullptr;
    if (scratch && scratch_bytes >= 2 * cg_bytes) {
        p.Cpart = (cplx*)scratch;
        const size_t slots = scratch_bytes / cg_bytes;
        max_split = slots > 4096 ? 4096 : (int)slots;
    }
    int rc = OIVA_ERR_INVALID, nsplit = 1;
    switch (n_chan) {
#define OIVA_CASE(M) case M: rc = relayout_cov_launch_m##M(dtype, p, max_split, st, &nsplit); break;
        OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
#undef OIVA_CASE
    }
    if (rc) return rc;
    if (nsplit > 1) {
        const size_t n = cg_bytes / sizeof(cplx);
        k_cov_sum_partials<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.Cpart, p.Cg, nsplit, n);
        OIVA_LAUNCH_CHECK();
    }
    return OIVA_OK;
}
