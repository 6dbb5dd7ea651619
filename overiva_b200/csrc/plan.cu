// The whole overiva() call on device pointers: a plan carves one caller-provided workspace and sequences
// the kernels (include/overiva_b200.h, "Plan" section).  Mirrors overiva.py:80-204 step by step.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

int oiva_reserve_ones(int n_frames);  // cov.cu

// optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launching stream
// around the launches of the three loop kernels; read back after the caller has synchronised
enum { TK_COV = 0, TK_POWER = 1, TK_SOLVE = 2, TK_N = 3 };
struct TimedSpan {
    cudaEvent_t a, b;
    int kind;
};

struct oiva_plan {
    oiva_plan_desc d;
    int n_freq_total;
    long long R, G;
    GroupLayout L;
    int Tp, NG;
    int es;  // bytes per real element of X / Y
    // workspace offsets
    size_t off_xg, off_c, off_cg, off_what, off_wg, off_vg, off_weff, off_r2part, off_r2, off_phi, off_wscale, off_evals, off_status, off_covws, off_sync, off_smcount;
    int resident;  // the persistent single-launch loop: 0 = not tried yet, 1 = in use, -1 = shape does not fit
    size_t covws_bytes;
    size_t ws_bytes;
    unsigned char* ws;
    long long launches;
    bool loaded, inited;
    bool c_full;  // the row-major copy C (R,M,M) of the input covariance is up to date (oiva_plan_run skips it when
                  // nothing on its path reads it: the thread-per-bin kernels work on the grouped Cg)
    bool timing;
    std::vector<TimedSpan>* spans;
    std::vector<cudaEvent_t>* pool;
    // the n_iter epochs of oiva_plan_iterate captured once as a CUDA graph (launch-bound shapes: 4 launches/epoch)
    cudaGraphExec_t graph_exec;
    int graph_n_iter;
    long long graph_launches;  // kernel launches one replay stands for
};

static cudaEvent_t plan_event(oiva_plan* p) {
    cudaEvent_t e = nullptr;
    if (!p->pool->empty()) {
        e = p->pool->back();
        p->pool->pop_back();
    } else if (cudaEventCreate(&e) != cudaSuccess) {
        cudaGetLastError();
        e = nullptr;
    }
    return e;
}
struct SpanGuard {  // records an event pair around a launch sequence when timing is on
    oiva_plan* p;
    cudaStream_t st;
    TimedSpan sp;
    bool on;
    SpanGuard(oiva_plan* p_, int kind, void* stream) : p(p_), st((cudaStream_t)stream), on(p_->timing) {
        if (on) {
            sp.kind = kind;
            sp.a = plan_event(p);
            sp.b = plan_event(p);
            if (!sp.a || !sp.b) {  // no event to be had: this span simply goes untimed
                if (sp.a) p->pool->push_back(sp.a);
                if (sp.b) p->pool->push_back(sp.b);
                on = false;
                return;
            }
            cudaEventRecord(sp.a, st);
        }
    }
    void cancel() {  // nothing was launched: hand the events back
        if (on) {
            p->pool->push_back(sp.a);
            p->pool->push_back(sp.b);
            on = false;
        }
    }
    ~SpanGuard() {
        if (on) {
            cudaEventRecord(sp.b, st);
            p->spans->push_back(sp);
        }
    }
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

extern "C" int oiva_plan_create(oiva_plan_t** out, const oiva_plan_desc* desc) {
    OIVA_REQUIRE(out && desc, "oiva_plan_create: null pointer");
    const oiva_plan_desc& d = *desc;
    OIVA_REQUIRE(d.n_batch > 0 && d.n_batch <= 65535 && d.n_frames > 0 && d.n_freq > 0,
                 "oiva_plan_create: bad shape B=%d T=%d F=%d", d.n_batch, d.n_frames, d.n_freq);
    OIVA_REQUIRE(d.n_chan >= 1 && d.n_chan <= OIVA_MAX_M, "oiva_plan_create: n_chan=%d not in 1..16", d.n_chan);
    OIVA_REQUIRE(d.n_src >= 1 && d.n_src <= d.n_chan, "oiva_plan_create: n_src=%d not in 1..n_chan=%d", d.n_src,
                 d.n_chan);
    OIVA_REQUIRE(d.dtype == OIVA_C128 || d.dtype == OIVA_C64, "oiva_plan_create: bad dtype %d", d.dtype);
    OIVA_REQUIRE(d.model >= OIVA_MODEL_LAPLACE && d.model <= OIVA_MODEL_OGIVE_GAUSS, "oiva_plan_create: bad model %d",
                 d.model);
    OIVA_REQUIRE((long long)d.n_batch * d.n_freq < (1ll << 31), "oiva_plan_create: too many rows");
    oiva_plan* p = (oiva_plan*)calloc(1, sizeof(oiva_plan));
    if (!p) {
        oiva_set_error("oiva_plan_create: out of host memory");
        return OIVA_ERR_NOMEM;
    }
    p->d = d;
    p->n_freq_total = d.n_freq_total > 0 ? d.n_freq_total : d.n_freq;
    p->R = (long long)d.n_batch * d.n_freq;
    p->L = oiva_make_layout(d.n_frames, d.n_freq, d.n_chan);
    p->NG = p->L.NG;
    p->G = (long long)d.n_batch * p->NG;
    p->Tp = p->L.frame_pitch();
    p->es = d.dtype == OIVA_C64 ? 4 : 8;
    const size_t M = d.n_chan, K = d.n_src, R = (size_t)p->R, B = d.n_batch, G = (size_t)p->G;
    size_t o = 0;
    p->off_xg = o;      o += align_up(oiva_grouped_bytes(d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.dtype));
    p->off_c = o;       o += align_up(R * M * M * 16);
    p->off_cg = o;      o += align_up(G * oiva_tri((int)M) * OIVA_GROUP * 16);
    p->off_what = o;    o += align_up(R * M * M * 16);
    p->off_wg = o;      o += align_up(G * M * M * OIVA_GROUP * 16);  // the loop's W_hat, grouped
    // grouped covariances of the K sources; also scratch for the eigenvectors at init time
    p->off_vg = o;      o += align_up(max_sz(G * K * oiva_tri((int)M) * OIVA_GROUP * 16, R * M * M * 16));
    p->off_weff = o;    o += align_up(max_sz(R * M * K, G * K * OIVA_GROUP) * 16);  // (also the z scratch of the output)
    p->off_r2part = o;  o += align_up(B * p->NG * K * p->Tp * 8);
    p->off_r2 = o;      o += align_up(B * K * p->Tp * 8);
    p->off_phi = o;     o += align_up(B * K * p->Tp * 8);
    p->off_wscale = o;  o += align_up(B * K * 8);
    p->off_evals = o;   o += align_up(R * M * 8);
    p->off_status = o;  o += align_up(sizeof(int) * (B + 4));  // one status word per mixture
    // per-split partial covariances of inputs with few bin groups (deterministic frame-split accumulation)
    p->covws_bytes = oiva_weighted_cov_scratch_bytes(d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.n_src);
    p->off_covws = o;   o += align_up(p->covws_bytes);
    p->off_sync = o;    o += align_up(oiva_loop_resident_sync_bytes(d.n_batch, d.n_freq));
    p->off_smcount = o; o += align_up(sizeof(unsigned) * B * K);  // completion counters of the frame-parallel source model
    p->ws_bytes = o;
    p->spans = new std::vector<TimedSpan>();
    p->pool = new std::vector<cudaEvent_t>();
    *out = p;
    // the "ones" weights of the unweighted input covariance: sized here so that no allocation happens mid-stream
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) == cudaSuccess && dev_count > 0) {
        int rc = oiva_reserve_ones(d.n_frames);
        if (rc) {
            oiva_plan_destroy(p);
            *out = nullptr;
            return rc;
        }
    } else {
        cudaGetLastError();
    }
    return OIVA_OK;
}

extern "C" void oiva_plan_destroy(oiva_plan_t* plan) {
    if (!plan) return;
    if (plan->spans) {
        for (auto& s : *plan->spans) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        delete plan->spans;
    }
    if (plan->pool) {
        for (auto e : *plan->pool) cudaEventDestroy(e);
        delete plan->pool;
    }
    if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
    free(plan);
}

extern "C" int oiva_plan_enable_timing(oiva_plan_t* plan, int enable) {
    OIVA_REQUIRE(plan != nullptr, "oiva_plan_enable_timing: null plan");
    plan->timing = enable != 0;
    return OIVA_OK;
}

extern "C" int oiva_plan_read_timing(oiva_plan_t* plan, double* ms, long long* counts) {
    OIVA_REQUIRE(plan && ms && counts, "oiva_plan_read_timing: null pointer");
    for (int k = 0; k < TK_N; ++k) {
        ms[k] = 0.0;
        counts[k] = 0;
    }
    for (auto& s : *plan->spans) {
        float t = 0.f;
        OIVA_CUDA_CHECK(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.kind] += t;
        counts[s.kind] += 1;
        plan->pool->push_back(s.a);
        plan->pool->push_back(s.b);
    }
    plan->spans->clear();
    return OIVA_OK;
}

extern "C" size_t oiva_plan_workspace_bytes(const oiva_plan_t* plan) { return plan ? plan->ws_bytes : 0; }

extern "C" int oiva_plan_bind(oiva_plan_t* plan, void* workspace, size_t bytes) {
    OIVA_REQUIRE(plan && workspace, "oiva_plan_bind: null pointer");
    OIVA_REQUIRE(bytes >= plan->ws_bytes, "oiva_plan_bind: workspace %zu < %zu bytes", bytes, plan->ws_bytes);
    OIVA_REQUIRE(((uintptr_t)workspace & 255) == 0, "oiva_plan_bind: workspace must be 256-byte aligned");
    if (plan->graph_exec && plan->ws != (unsigned char*)workspace) {  // captured pointers are stale
        cudaGraphExecDestroy(plan->graph_exec);
        plan->graph_exec = nullptr;
    }
    plan->ws = (unsigned char*)workspace;
    plan->loaded = plan->inited = false;
    return OIVA_OK;
}

#define PLAN_READY(p, who)                                                        \
    OIVA_REQUIRE((p) != nullptr, who ": null plan");                              \
    if (!(p)->ws) {                                                               \
        oiva_set_error(who ": no workspace bound");                               \
        return OIVA_ERR_STATE;                                                    \
    }

extern "C" void* oiva_plan_what(oiva_plan_t* p) { return (p && p->ws) ? p->ws + p->off_what : nullptr; }
extern "C" void* oiva_plan_cov(oiva_plan_t* p) { return (p && p->ws) ? p->ws + p->off_c : nullptr; }
extern "C" void* oiva_plan_samples(oiva_plan_t* p) { return (p && p->ws) ? p->ws + p->off_xg : nullptr; }
extern "C" int* oiva_plan_status_ptr(oiva_plan_t* p) { return (p && p->ws) ? (int*)(p->ws + p->off_status) : nullptr; }
extern "C" double* oiva_plan_r2(oiva_plan_t* p) { return (p && p->ws) ? (double*)(p->ws + p->off_r2) : nullptr; }
extern "C" size_t oiva_plan_r2_elems(const oiva_plan_t* p) {
    return p ? (size_t)p->d.n_batch * p->d.n_src * p->Tp : 0;
}
// device pointers into the bound workspace for callers that sequence kernels themselves on the plan's arrays (ILRMA):
// which = 0 grouped W_hat (G, M*M, 32) c128 | 1 grouped input covariance Cg (G, NE, 32) | 2 grouped weighted covariances
// Vg (G, K, NE, 32) | 3 statistic partials r2part (B, NG, K, Tp) f64 | 4 frame-split scratch of the covariance kernel
// (oiva_plan_scratch_bytes) | 5 scratch of the output scales (G, K, 32) c128
extern "C" void* oiva_plan_array(oiva_plan_t* p, int which) {
    if (!p || !p->ws) return nullptr;
    switch (which) {
        case 0: return p->ws + p->off_wg;
        case 1: return p->ws + p->off_cg;
        case 2: return p->ws + p->off_vg;
        case 3: return p->ws + p->off_r2part;
        case 4: return p->covws_bytes ? p->ws + p->off_covws : nullptr;
        case 5: return p->ws + p->off_weff;
    }
    return nullptr;
}
extern "C" size_t oiva_plan_scratch_bytes(const oiva_plan_t* p) { return p ? p->covws_bytes : 0; }

extern "C" long long oiva_plan_launch_count(const oiva_plan_t* p) { return p ? p->launches : 0; }

// input covariance C = (1/T) sum_t x x^H (overiva.py:87): grouped accumulation, then full row-major matrices
static bool plan_force_rowowner() {  // (read per call: the tests run both solver families in one process)
    const char* env = getenv("OIVA_SOLVER_ROWOWNER");
    return env && *env && *env != '0';
}
// the shapes whose init / sweep / output run thread-per-bin on the grouped arrays only (solve_tpb.cu's instantiations)
static bool plan_tpb_shape(const oiva_plan_t* p) {
    const int M = p->d.n_chan, K = p->d.n_src;
    return !plan_force_rowowner() && (M <= 6 || (M <= 8 && K <= 4));
}

// full row-major matrices C (R,M,M) from the grouped lower triangles (needed by the eigendecomposition, the row-owner
// kernels and callers of oiva_plan_cov)
static int plan_full_cov(oiva_plan_t* p, void* stream) {
    if (p->c_full) return OIVA_OK;
    const oiva_plan_desc& d = p->d;
    int rc = oiva_unpack_cov(p->ws + p->off_cg, p->ws + p->off_c, d.n_batch, d.n_freq, d.n_chan, 1, stream);
    if (rc) return rc;
    p->launches += 1;
    p->c_full = true;
    return OIVA_OK;
}

static int plan_input_cov(oiva_plan_t* p, void* stream) {
    const oiva_plan_desc& d = p->d;
    int rc = oiva_weighted_cov_ws(p->ws + p->off_xg, nullptr, p->ws + p->off_cg, p->covws_bytes ? p->ws + p->off_covws : nullptr,
                                  p->covws_bytes, d.n_batch, d.n_frames, d.n_freq, d.n_chan, 1, d.dtype, stream);
    if (rc) return rc;
    p->launches += 1;
    p->c_full = false;
    return plan_full_cov(p, stream);
}

static int plan_load_impl(oiva_plan_t* p, const void* X, void* stream, bool full_c);

extern "C" int oiva_plan_load(oiva_plan_t* p, const void* X, void* stream) {
    return plan_load_impl(p, X, stream, true);
}

static int plan_load_impl(oiva_plan_t* p, const void* X, void* stream, bool full_c) {
    PLAN_READY(p, "oiva_plan_load");
    OIVA_REQUIRE(X, "oiva_plan_load: null X");
    const oiva_plan_desc& d = p->d;
    int rc;
    // (the source model's completion counters must start at zero; they return to zero after every epoch)
    OIVA_CUDA_CHECK(cudaMemsetAsync(p->ws + p->off_smcount, 0, sizeof(unsigned) * (size_t)d.n_batch * d.n_src,
                                    (cudaStream_t)stream));
    static const bool no_fuse = [] {
        const char* v = getenv("OIVA_NO_RELAYOUT_COV");
        return v && *v && *v != '0';
    }();
    if (!no_fuse && oiva_relayout_cov_supported(d.n_freq, d.n_chan, d.dtype) && (((uintptr_t)X) & 15) == 0) {
        // one pass: grouped samples + grouped covariance, then the full row-major matrices        overiva.py:87,131-132
        rc = oiva_relayout_cov(X, p->ws + p->off_xg, p->ws + p->off_cg, p->covws_bytes ? p->ws + p->off_covws : nullptr,
                               p->covws_bytes, d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.dtype, stream);
        if (rc) return rc;
        p->launches += 1;
        p->c_full = false;
        if (full_c) {
            rc = plan_full_cov(p, stream);
            if (rc) return rc;
        }
    } else {
        rc = oiva_relayout(X, p->ws + p->off_xg, d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.dtype, stream);
        if (rc) return rc;
        p->launches += 1;
        rc = plan_input_cov(p, stream);
        if (rc) return rc;
    }
    p->loaded = true;
    p->inited = false;
    return OIVA_OK;
}

extern "C" int oiva_plan_adopt_samples(oiva_plan_t* p, void* stream) {
    PLAN_READY(p, "oiva_plan_adopt_samples");
    OIVA_CUDA_CHECK(cudaMemsetAsync(p->ws + p->off_smcount, 0, sizeof(unsigned) * (size_t)p->d.n_batch * p->d.n_src,
                                    (cudaStream_t)stream));
    int rc = plan_input_cov(p, stream);
    if (rc) return rc;
    p->loaded = true;
    p->inited = false;
    return OIVA_OK;
}

extern "C" int oiva_plan_init(oiva_plan_t* p, int mode, const void* W0, void* stream) {
    PLAN_READY(p, "oiva_plan_init");
    if (!p->loaded) {
        oiva_set_error("oiva_plan_init: call oiva_plan_load first");
        return OIVA_ERR_STATE;
    }
    const oiva_plan_desc& d = p->d;
    int* status = (int*)(p->ws + p->off_status);
    const void* evecs = nullptr;
    if (mode != OIVA_INIT_EIG && !plan_force_rowowner()) {
        // identity / W0: one thread-per-bin kernel writes the grouped W_hat directly (solve_tpb.cuh)
        OIVA_REQUIRE(mode != OIVA_INIT_W0 || W0, "oiva_plan_init: W0 missing");
        const int rc = oiva_init_demix_grouped(p->ws + p->off_wg, p->ws + p->off_cg, mode == OIVA_INIT_W0 ? W0 : nullptr,
                                               status, d.n_batch, d.n_freq, d.n_chan, d.n_src, stream);
        if (rc == OIVA_OK) {
            p->launches += 1;
            p->inited = true;
            return OIVA_OK;
        }
        if (rc != OIVA_ERR_UNSUPPORTED) return rc;
    }
    {
        const int rc = plan_full_cov(p, stream);  // (everything below reads the row-major C)
        if (rc) return rc;
    }
    if (mode == OIVA_INIT_EIG) {
        // principal eigenvectors of C with np.linalg.eig's phase convention            overiva.py:103-109
        int rc = oiva_eigh(p->ws + p->off_c, (double*)(p->ws + p->off_evals), p->ws + p->off_vg, status, (int)p->R,
                           d.n_freq, d.n_chan, 1, stream);
        if (rc) return rc;
        evecs = p->ws + p->off_vg;
        p->launches += 1;
    }
    int rc = oiva_init_demix(p->ws + p->off_what, p->ws + p->off_c, W0, evecs, mode, status, (int)p->R, d.n_freq,
                             d.n_chan, d.n_src, stream);
    if (rc) return rc;
    rc = oiva_group_rows(p->ws + p->off_what, p->ws + p->off_wg, d.n_batch, d.n_freq, d.n_chan * d.n_chan, stream);
    if (rc) return rc;
    p->launches += 2;
    p->inited = true;
    return OIVA_OK;
}

// row-major copy of the loop's grouped W_hat (for projection back / filters / callers)
static int plan_sync_what(oiva_plan_t* p, void* stream) {
    const oiva_plan_desc& d = p->d;
    int rc = oiva_ungroup_rows(p->ws + p->off_wg, p->ws + p->off_what, d.n_batch, d.n_freq, d.n_chan * d.n_chan, stream);
    if (rc) return rc;
    p->launches += 1;
    return OIVA_OK;
}

static int plan_power_partials(oiva_plan_t* p, void* stream) {
    const oiva_plan_desc& d = p->d;
    SpanGuard g(p, TK_POWER, stream);
    int rc = oiva_demix_power(p->ws + p->off_xg, p->ws + p->off_wg, d.n_chan, 1, (double*)(p->ws + p->off_r2part),
                              d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.n_src, d.dtype, stream);
    if (rc) return rc;
    p->launches += 1;
    return OIVA_OK;
}

static int plan_update_from(oiva_plan_t* p, const double* r2src, int nch, void* stream) {
    const oiva_plan_desc& d = p->d;
    double* phi = (double*)(p->ws + p->off_phi);
    double* wscale = (double*)(p->ws + p->off_wscale);
    int rc = oiva_source_model_ws(r2src, nch, phi, wscale, (unsigned*)(p->ws + p->off_smcount), d.n_batch, d.n_frames,
                                  d.n_src, p->n_freq_total, d.model, stream);
    if (rc) return rc;
    {
        // one kernel for covariance + sweep where the covariances of a bin fit its lane's registers (cov_sweep.cuh)
        const char* off = getenv("OIVA_NO_COV_SWEEP");  // (read per call: the tests compare both paths in one process)
        if (!(off && *off && *off != '0')) {
            SpanGuard g(p, TK_COV, stream);
            rc = oiva_cov_ip_update(p->ws + p->off_xg, phi, p->ws + p->off_wg, p->ws + p->off_cg, wscale,
                                    (int*)(p->ws + p->off_status), d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.n_src,
                                    d.dtype, stream);
            if (rc == OIVA_OK) {
                p->launches += 2;  // source model + the fused kernel
                return rc;
            }
            g.cancel();
            if (rc != OIVA_ERR_UNSUPPORTED) return rc;
        }
    }
    {
        SpanGuard g(p, TK_COV, stream);
        rc = oiva_weighted_cov_ws(p->ws + p->off_xg, phi, p->ws + p->off_vg, p->covws_bytes ? p->ws + p->off_covws : nullptr,
                                  p->covws_bytes, d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.n_src, d.dtype, stream);
    }
    if (rc) return rc;
    if (!plan_tpb_shape(p)) {
        rc = plan_full_cov(p, stream);  // the row-owner sweep reads the row-major C
        if (rc) return rc;
    }
    {
        SpanGuard g(p, TK_SOLVE, stream);
        rc = oiva_ip_update(p->ws + p->off_wg, p->ws + p->off_vg, p->ws + p->off_c, p->ws + p->off_cg, wscale,
                            (int*)(p->ws + p->off_status), d.n_batch, d.n_freq, d.n_chan, d.n_src, stream);
    }
    if (rc) return rc;
    p->launches += 3;
    return OIVA_OK;
}

#define PLAN_INITED(p, who)                                                   \
    PLAN_READY(p, who);                                                       \
    if (!(p)->inited) {                                                       \
        oiva_set_error(who ": call oiva_plan_load and oiva_plan_init first"); \
        return OIVA_ERR_STATE;                                                \
    }

static int plan_iterate_eager(oiva_plan_t* p, int n_iter, void* stream) {
    for (int it = 0; it < n_iter; ++it) {
        int rc = plan_power_partials(p, stream);
        if (rc) return rc;
        rc = plan_update_from(p, (const double*)(p->ws + p->off_r2part), p->NG, stream);
        if (rc) return rc;
    }
    return OIVA_OK;
}

// Small problems are launch-bound (config 1: ~15 us kernels, 4 per epoch): the whole n_iter loop is captured once
// per (plan, n_iter) as a CUDA graph and replayed with a single launch.  Large batches (>= 4096 bin groups) run
// eagerly: their kernels are milliseconds long and the per-kernel event timing of bench.py needs real launches.
static bool plan_use_graph(const oiva_plan_t* p, int n_iter) {
    static const bool disabled = [] {
        const char* v = getenv("OIVA_NO_GRAPH");
        return v && *v && *v != '0';
    }();
    // (inputs of hundreds of MB run kernels of milliseconds: nothing to gain, and their plans are too large for the
    // Python-side plan cache, so the capture would be paid on every call)
    const size_t xg_bytes = oiva_grouped_bytes(p->d.n_batch, p->d.n_frames, p->d.n_freq, p->d.n_chan, p->d.dtype);
    return !disabled && !p->timing && n_iter >= 2 && p->G < 4096 && xg_bytes < ((size_t)256 << 20);
}

// One short mixture (few bin groups, samples that fit the shared memory of the SMs): the whole loop in ONE persistent
// cooperative launch (resident.cuh) instead of 5 kernels per epoch.  Returns OIVA_ERR_UNSUPPORTED when the shape does not
// fit (remembered: later calls go straight to the kernel-per-step loop).
static int plan_iterate_resident(oiva_plan_t* p, int n_iter, void* stream) {
    const char* off = getenv("OIVA_NO_RESIDENT");  // (read per call: the tests run both loops in one process)
    const bool disabled = off && *off && *off != '0';
    const oiva_plan_desc& d = p->d;
    if (disabled || p->timing || p->resident < 0 || !p->covws_bytes || d.n_chan > 8 ||
        !(d.model == OIVA_MODEL_LAPLACE || d.model == OIVA_MODEL_GAUSS || d.model == OIVA_MODEL_NONE))
        return OIVA_ERR_UNSUPPORTED;
    int rc = oiva_loop_resident(p->ws + p->off_xg, p->ws + p->off_wg, p->ws + p->off_cg, (double*)(p->ws + p->off_r2part),
                                (double*)(p->ws + p->off_r2), p->ws + p->off_covws, p->covws_bytes, p->ws + p->off_sync,
                                (int*)(p->ws + p->off_status), d.n_batch, d.n_frames, d.n_freq, p->n_freq_total, d.n_chan,
                                d.n_src, d.model, d.dtype, n_iter, stream);
    if (rc == OIVA_ERR_UNSUPPORTED) {
        p->resident = -1;
        return rc;
    }
    if (rc) return rc;
    p->resident = 1;
    p->launches += 1;
    return OIVA_OK;
}

extern "C" int oiva_plan_iterate(oiva_plan_t* p, int n_iter, void* stream) {
    PLAN_INITED(p, "oiva_plan_iterate");
    if (n_iter <= 0) return OIVA_OK;
    if (!plan_tpb_shape(p)) {  // (before any graph capture: the row-owner sweep reads the row-major C)
        const int rc = plan_full_cov(p, stream);
        if (rc) return rc;
    }
    {
        const int rc = plan_iterate_resident(p, n_iter, stream);
        if (rc != OIVA_ERR_UNSUPPORTED) return rc;
    }
    if (!plan_use_graph(p, n_iter)) return plan_iterate_eager(p, n_iter, stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->graph_exec || p->graph_n_iter != n_iter) {
        if (p->graph_exec) {
            cudaGraphExecDestroy(p->graph_exec);
            p->graph_exec = nullptr;
        }
        const long long l0 = p->launches;
        cudaGraph_t graph = nullptr;
        cudaStream_t cap = nullptr;
        OIVA_CUDA_CHECK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        // capture on a private stream: nothing executes, the launchers' one-time attribute / occupancy calls are legal
        // during capture, and the caller's stream is left untouched if capture fails
        cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        int rc = OIVA_OK;
        if (e == cudaSuccess) {
            rc = plan_iterate_eager(p, n_iter, cap);
            e = cudaStreamEndCapture(cap, &graph);
        }
        p->graph_launches = p->launches - l0;
        p->launches = l0;
        if (rc == OIVA_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(cap);
        if (rc != OIVA_OK || e != cudaSuccess || !p->graph_exec) {
            cudaGetLastError();  // clear; fall back to eager launches
            p->graph_exec = nullptr;
            return plan_iterate_eager(p, n_iter, stream);
        }
        p->graph_n_iter = n_iter;
    }
    OIVA_CUDA_CHECK(cudaGraphLaunch(p->graph_exec, st));
    p->launches += p->graph_launches;
    return OIVA_OK;
}

extern "C" int oiva_plan_power(oiva_plan_t* p, void* stream) {
    PLAN_INITED(p, "oiva_plan_power");
    int rc = plan_power_partials(p, stream);
    if (rc) return rc;
    rc = oiva_sum_partials((const double*)(p->ws + p->off_r2part), p->NG, (double*)(p->ws + p->off_r2), p->d.n_batch,
                           p->d.n_frames, p->d.n_src, stream);
    if (rc) return rc;
    p->launches += 1;
    return OIVA_OK;
}

extern "C" int oiva_plan_update(oiva_plan_t* p, void* stream) {
    PLAN_INITED(p, "oiva_plan_update");
    return plan_update_from(p, (const double*)(p->ws + p->off_r2), 1, stream);
}

extern "C" int oiva_plan_filters(oiva_plan_t* p, void* W, void* stream);

extern "C" int oiva_plan_output(oiva_plan_t* p, int proj_back, void* Y, void* stream) {
    PLAN_INITED(p, "oiva_plan_output");
    OIVA_REQUIRE(Y, "oiva_plan_output: null Y");
    const oiva_plan_desc& d = p->d;
    // the filters come from the grouped W_hat, the projection-back scales from the grouped C (off_weff is their scratch)
    int rc = oiva_demix_output_grouped(p->ws + p->off_xg, p->ws + p->off_wg, proj_back ? p->ws + p->off_cg : nullptr,
                                       p->ws + p->off_weff, Y, d.n_batch, d.n_frames, d.n_freq, d.n_chan, d.n_src, d.dtype, stream);
    if (rc) return rc;
    p->launches += proj_back ? 2 : 1;
    return OIVA_OK;
}

// load + init + iterate + output (+ filters) in one call: for one short mixture the per-call overhead of the host
// language (five ctypes round trips from Python) is of the order of the GPU work itself
extern "C" int oiva_plan_run(oiva_plan_t* p, const void* X, int init_mode, const void* W0, int n_iter, int proj_back,
                             void* Y, void* W, void* stream) {
    // (the row-major copy of C is only produced when something on the path reads it)
    int rc = plan_load_impl(p, X, stream, false);
    if (rc) return rc;
    rc = oiva_plan_init(p, init_mode, W0, stream);
    if (rc) return rc;
    rc = oiva_plan_iterate(p, n_iter, stream);
    if (rc) return rc;
    rc = oiva_plan_output(p, proj_back, Y, stream);
    if (rc) return rc;
    return W ? oiva_plan_filters(p, W, stream) : OIVA_OK;
}

extern "C" int oiva_plan_filters(oiva_plan_t* p, void* W, void* stream) {
    PLAN_INITED(p, "oiva_plan_filters");
    OIVA_REQUIRE(W, "oiva_plan_filters: null W");
    const oiva_plan_desc& d = p->d;
    // plain copy of the W columns: proj_back = 0
    int rc = plan_sync_what(p, stream);
    if (rc) return rc;
    rc = oiva_projback_filters(p->ws + p->off_what, d.n_chan, p->ws + p->off_c, W, (int)p->R, d.n_chan, d.n_src, 0,
                                   stream);
    if (rc) return rc;
    p->launches += 1;
    return OIVA_OK;
}

extern "C" int oiva_plan_reset_status(oiva_plan_t* p, void* stream) {
    PLAN_READY(p, "oiva_plan_reset_status");
    OIVA_CUDA_CHECK(cudaMemsetAsync(p->ws + p->off_status, 0, sizeof(int) * ((size_t)p->d.n_batch + 4),
                                    (cudaStream_t)stream));
    return OIVA_OK;
}

extern "C" int oiva_plan_status_vector(oiva_plan_t* p, int* status_host, void* stream) {
    PLAN_READY(p, "oiva_plan_status");
    const int B = p->d.n_batch;
    std::vector<int> tmp;
    int* h = status_host;
    if (!h) {
        tmp.resize(B);
        h = tmp.data();
    }
    OIVA_CUDA_CHECK(cudaMemcpyAsync(h, p->ws + p->off_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream));
    OIVA_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    int all = 0;
    for (int b = 0; b < B; ++b) all |= h[b];
    return all & (OIVA_STATUS_SINGULAR | OIVA_STATUS_NONFINITE);
}

extern "C" int oiva_plan_status(oiva_plan_t* p, void* stream) { return oiva_plan_status_vector(p, nullptr, stream); }

extern "C" int oiva_overiva_host(const void* X_host, void* Y_host, void* W_host, const void* W0_host,
                                 const oiva_plan_desc* desc, int n_iter, int proj_back, int init_mode,
                                 int* status_host) {
    OIVA_REQUIRE(X_host && Y_host && desc, "oiva_overiva_host: null pointer");
    OIVA_REQUIRE(init_mode != OIVA_INIT_W0 || W0_host, "oiva_overiva_host: W0 missing");
    oiva_plan_t* p = nullptr;
    int rc = oiva_plan_create(&p, desc);
    if (rc) return rc;
    const oiva_plan_desc& d = p->d;
    const size_t ce = d.dtype == OIVA_C64 ? 8 : 16;
    const size_t xbytes = (size_t)d.n_batch * d.n_frames * d.n_freq * d.n_chan * ce;
    const size_t ybytes = (size_t)d.n_batch * d.n_frames * d.n_freq * d.n_src * ce;
    const size_t wbytes = (size_t)p->R * d.n_chan * d.n_src * 16;
    unsigned char *dX = nullptr, *dY = nullptr, *dW = nullptr, *ws = nullptr;
    cudaStream_t st = nullptr;
    int status = OIVA_OK;
#define HOST_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            oiva_set_error("oiva_overiva_host: %s -> CUDA error %d (%s)", #expr, (int)_e,           \
                           cudaGetErrorString(_e));                                                 \
            status = OIVA_ERR_CUDA;                                                                 \
            goto done;                                                                              \
        }                                                                                           \
    } while (0)
#define HOST_RC(expr)        \
    do {                     \
        status = (expr);     \
        if (status) goto done; \
    } while (0)
    HOST_TRY(cudaStreamCreate(&st));
    HOST_TRY(cudaMalloc(&dX, xbytes));
    HOST_TRY(cudaMalloc(&dY, ybytes));
    HOST_TRY(cudaMalloc(&dW, wbytes));
    HOST_TRY(cudaMalloc(&ws, p->ws_bytes));
    HOST_RC(oiva_plan_bind(p, ws, p->ws_bytes));
    HOST_RC(oiva_plan_reset_status(p, st));
    HOST_TRY(cudaMemcpyAsync(dX, X_host, xbytes, cudaMemcpyHostToDevice, st));
    if (init_mode == OIVA_INIT_W0) HOST_TRY(cudaMemcpyAsync(dW, W0_host, wbytes, cudaMemcpyHostToDevice, st));
    HOST_RC(oiva_plan_load(p, dX, st));
    HOST_RC(oiva_plan_init(p, init_mode, init_mode == OIVA_INIT_W0 ? dW : nullptr, st));
    HOST_RC(oiva_plan_iterate(p, n_iter, st));
    HOST_RC(oiva_plan_output(p, proj_back, dY, st));
    HOST_TRY(cudaMemcpyAsync(Y_host, dY, ybytes, cudaMemcpyDeviceToHost, st));
    if (W_host) {
        HOST_RC(oiva_plan_filters(p, dW, st));
        HOST_TRY(cudaMemcpyAsync(W_host, dW, wbytes, cudaMemcpyDeviceToHost, st));
    }
    status = oiva_plan_status_vector(p, status_host, st);  // synchronises
done:
    if (st) cudaStreamSynchronize(st);
    cudaFree(dX);
    cudaFree(dY);
    cudaFree(dW);
    cudaFree(ws);
    if (st) cudaStreamDestroy(st);
    oiva_plan_destroy(p);
    return status;
#undef HOST_TRY
#undef HOST_RC
}
