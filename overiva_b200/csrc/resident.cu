// C entry point of the persistent single-launch epoch loop (include/overiva_b200.h: oiva_loop_resident).
#include <stdio.h>
#include <stdlib.h>

#include "resident.cuh"

namespace oiva {
#define OIVA_DECL(M)                                                                                              \
    int resident_launch_m##M(int dtype, int K, const ResidentParams& p, unsigned grid, size_t smem, cudaStream_t st); \
    int resident_fw_m##M(int K);
OIVA_DECL(1) OIVA_DECL(2) OIVA_DECL(3) OIVA_DECL(4) OIVA_DECL(5) OIVA_DECL(6) OIVA_DECL(7) OIVA_DECL(8)
#undef OIVA_DECL

static int resident_fw(int M, int K) {
    switch (M) {
#define OIVA_CASE(M_) case M_: return resident_fw_m##M_(K);
        OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
#undef OIVA_CASE
    }
    return 0;
}

constexpr size_t RES_MAX_SMEM = 232448;  // 227 KB: the per-CTA maximum of sm_100
constexpr int RES_MAX_SG = 8;            // slices per bin group (keeps the partial-sum slots <= 64)

struct ResidentChoice {
    int SG, slice_cap, v_bufs, fw;
    size_t smem;
};

// slices per group / shared-memory carve-up for a shape, or false when the samples cannot stay resident
static bool resident_choose(int B, int T, int F, int M, int K, int dtype, int max_ctas, size_t scratch_bytes,
                            ResidentChoice* out) {
    if (M < 1 || M > 8 || K < 1 || K > M || T < 1) return false;
    const long long G = (long long)B * oiva_bin_groups(F);
    if (G < 1 || G > max_ctas) return false;
    const int fw = resident_fw(M, K);
    if (fw < 1) return false;
    const size_t vg = oiva_grouped_cov_bytes(B, F, M, K);
    int SG = (int)(max_ctas / G);
    if (SG > RES_MAX_SG) SG = RES_MAX_SG;
    if (SG > T) SG = T;
    const int esz = dtype == OIVA_C64 ? 8 : 16;
    for (; SG >= 1; --SG) {
        if ((size_t)SG * fw * vg > scratch_bytes) continue;
        const int cap = (T + SG - 1) / SG;
        for (int vb = 2; vb >= 1; --vb) {  // (two V buffers: the next source's partial sums are added during a sweep)
            const ResSmem lay = res_smem_layout(M, K, cap, vb, esz);
            if (lay.total <= RES_MAX_SMEM) {
                out->SG = SG;
                out->slice_cap = cap;
                out->v_bufs = vb;
                out->fw = fw;
                out->smem = lay.total;
                return true;
            }
        }
        // a finer slicing needs more CTAs than the GPU has: only coarser ones remain, which need even more shared memory
        return false;
    }
    return false;
}
}  // namespace oiva

extern "C" size_t oiva_loop_resident_sync_bytes(int n_batch, int n_freq) {
    return sizeof(unsigned) * (oiva::RES_SYNC_HEADER + 2 * (size_t)n_batch * oiva_bin_groups(n_freq));
}

extern "C" int oiva_loop_resident(const void* Xg, void* Wg, const void* Cg, double* r2part, double* rbuf, void* scratch,
                                  size_t scratch_bytes, void* sync, int* status, int n_batch, int n_frames, int n_freq,
                                  int n_freq_total, int n_chan, int n_src, int model, int dtype, int n_iter, void* stream) {
    using namespace oiva;
    OIVA_REQUIRE(Xg && Wg && Cg && r2part && rbuf && scratch && sync && status, "oiva_loop_resident: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0 && n_chan >= 1 && n_src >= 1 && n_src <= n_chan,
                 "oiva_loop_resident: bad shape");
    OIVA_REQUIRE(model == OIVA_MODEL_LAPLACE || model == OIVA_MODEL_GAUSS || model == OIVA_MODEL_NONE,
                 "oiva_loop_resident: model %d not supported", model);
    if (n_iter <= 0) return OIVA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148, coop = 0;
    OIVA_CUDA_CHECK(cudaGetDevice(&dev));
    OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    ResidentChoice ch;
    if (!coop || !resident_choose(n_batch, n_frames, n_freq, n_chan, n_src, dtype, sms, scratch_bytes, &ch)) {
        oiva_set_error("oiva_loop_resident: shape B=%d T=%d F=%d M=%d K=%d does not fit the resident loop", n_batch,
                       n_frames, n_freq, n_chan, n_src);
        return OIVA_ERR_UNSUPPORTED;
    }
    ResidentParams p;
    p.Xg = Xg;
    p.Wg = (cplx*)Wg;
    p.Cg = (const cplx*)Cg;
    p.r2part = r2part;
    p.rbuf = rbuf;
    p.Vpart = (cplx*)scratch;
    p.sync = (unsigned*)sync;
    p.status = status;
    p.L = oiva_make_layout(n_frames, n_freq, n_chan);
    p.G = (long long)n_batch * p.L.NG;
    p.B = n_batch;
    p.SG = ch.SG;
    p.n_iter = n_iter;
    p.model = model;
    p.F_total = n_freq_total > 0 ? n_freq_total : n_freq;
    p.slice_cap = ch.slice_cap;
    p.v_bufs = ch.v_bufs;
    {
        const char* v = getenv("OIVA_RES_POLL");  // (read per call: profiles compare the modes in one process)
        p.poll = v ? atoi(v) : 2;
        v = getenv("OIVA_RES_TRACKED");
        p.tracked = v ? atoi(v) : 1;
        v = getenv("OIVA_RES_TAGGED");
        p.tagged = (v ? atoi(v) : 1) && p.L.NG <= RES_TAG_MAX_GROUPS;
        v = getenv("OIVA_RES_CLUSTER");
        p.cluster = v ? atoi(v) : 1;
    }
    p.invT = 1.0 / (double)n_frames;
    OIVA_CUDA_CHECK(cudaMemsetAsync(sync, 0, oiva_loop_resident_sync_bytes(n_batch, n_freq), st));
    const unsigned grid = (unsigned)(p.G * ch.SG);
    p.trace = nullptr;
    if (const char* path = getenv("OIVA_RES_TRACE")) {
        // profiling only: one synchronous launch with clock64 stamps of every CTA's phases, written to `path` as text
        // (cta epoch stamp0..stamp9, in SM cycles); the stamps add CTA barriers, so the traced launch is a little slower
        const size_t n = (size_t)grid * n_iter * RES_TRACE_POINTS;
        OIVA_CUDA_CHECK(cudaMalloc(&p.trace, n * sizeof(long long)));
        OIVA_CUDA_CHECK(cudaMemsetAsync(p.trace, 0, n * sizeof(long long), st));
        int rc = OIVA_ERR_UNSUPPORTED;
        switch (n_chan) {
#define OIVA_CASE(M_) case M_: rc = resident_launch_m##M_(dtype, n_src, p, grid, ch.smem, st); break;
            OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
#undef OIVA_CASE
        }
        if (rc == OIVA_OK && cudaStreamSynchronize(st) == cudaSuccess) {
            long long* h = (long long*)malloc(n * sizeof(long long));
            if (h && cudaMemcpy(h, p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
                if (FILE* f = fopen(path, "w")) {
                    for (size_t i = 0; i < (size_t)grid * n_iter; ++i) {
                        fprintf(f, "%zu %zu", i / n_iter, i % n_iter);
                        for (int j = 0; j < RES_TRACE_POINTS; ++j) fprintf(f, " %lld", h[i * RES_TRACE_POINTS + j]);
                        fprintf(f, "\n");
                    }
                    fclose(f);
                }
            }
            free(h);
        }
        cudaFree(p.trace);
        return rc;
    }
    switch (n_chan) {
#define OIVA_CASE(M_) case M_: return resident_launch_m##M_(dtype, n_src, p, grid, ch.smem, st);
        OIVA_CASE(1) OIVA_CASE(2) OIVA_CASE(3) OIVA_CASE(4) OIVA_CASE(5) OIVA_CASE(6) OIVA_CASE(7) OIVA_CASE(8)
#undef OIVA_CASE
    }
    return OIVA_ERR_UNSUPPORTED;
}
