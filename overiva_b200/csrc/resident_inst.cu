// Instantiates the persistent single-launch loop (resident.cuh) for ONE channel count (-DOIVA_M=<1..8>), K = 1..M.
#include "resident.cuh"

#ifndef OIVA_M
#error "compile with -DOIVA_M=<1..8>"
#endif

namespace oiva {

#define OIVA_CAT2(a, b) a##b
#define OIVA_CAT(a, b) OIVA_CAT2(a, b)

template <typename ST, int K>
static int launch_res(const ResidentParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    constexpr int M = OIVA_M;
    // the sweep runs one thread per bin: the shapes k_ip_update_tpb is instantiated for (solve_tpb.cu) -- the determined
    // LU of 7 or 8 channels does not fit one thread's registers
    if constexpr (K > M || (M >= 7 && K > 4)) {
        return OIVA_ERR_UNSUPPORTED;
    } else {
        // (the tracked-inverse determined sweep is a kernel of its own: inside the one kernel its registers spilled and
        // slowed the other phases -- config 2 loop 1.12 -> 1.23 ms with the switch off)
        constexpr bool CAN_TRACK = K == M && M >= 3;
        auto kern = k_loop_resident<ST, M, K, false>;
        if constexpr (CAN_TRACK) {
            if (p.tracked && p.v_bufs == 2) kern = k_loop_resident<ST, M, K, true>;
        }
        // (two kernels share this call site: set the attribute per launch -- it is cheap -- instead of once per device)
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        int dev = 0, sms = 0, occ = 0;
        OIVA_CUDA_CHECK(cudaGetDevice(&dev));
        OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        OIVA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RES_THREADS, smem));
        if ((long long)occ * sms < (long long)grid) {
            oiva_set_error("oiva_loop_resident: %u CTAs cannot be co-resident (%d x %d)", grid, occ, sms);
            return OIVA_ERR_UNSUPPORTED;
        }
        ResidentParams q = p;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(RES_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attrs[2];
        attrs[0].id = cudaLaunchAttributeCooperative;
        attrs[0].val.cooperative = 1;
        attrs[1].id = cudaLaunchAttributeClusterDimension;
        attrs[1].val.clusterDim.x = (unsigned)p.SG;
        attrs[1].val.clusterDim.y = 1;
        attrs[1].val.clusterDim.z = 1;
        cfg.attrs = attrs;
        if (q.cluster && p.SG > 1) {  // the slices of a bin group as one cluster, when that many clusters are co-resident
            cfg.numAttrs = 2;
            int n_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) != cudaSuccess ||
                (long long)n_clusters * p.SG < (long long)grid) {
                (void)cudaGetLastError();
                q.cluster = 0;
            }
        } else {
            q.cluster = 0;
        }
        if (q.cluster) {
            cfg.numAttrs = 2;
            if (cudaLaunchKernelEx(&cfg, kern, q) == cudaSuccess) return OIVA_OK;
            (void)cudaGetLastError();  // (cooperative + cluster refused: the flag hand-over needs neither)
            q.cluster = 0;
        }
        cfg.numAttrs = 1;
        OIVA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, q));
        return OIVA_OK;
    }
}

int OIVA_CAT(resident_launch_m, OIVA_M)(int dtype, int K, const ResidentParams& p, unsigned grid, size_t smem,
                                        cudaStream_t st) {
#define OIVA_RES_CASE(K_)                                                                        \
    case K_:                                                                                     \
        return dtype == OIVA_C64 ? launch_res<float, K_>(p, grid, smem, st) : launch_res<double, K_>(p, grid, smem, st);
    switch (K) {
        OIVA_RES_CASE(1) OIVA_RES_CASE(2) OIVA_RES_CASE(3) OIVA_RES_CASE(4)
        OIVA_RES_CASE(5) OIVA_RES_CASE(6) OIVA_RES_CASE(7) OIVA_RES_CASE(8)
    }
    return OIVA_ERR_INVALID;
}

// covariance frame ranges per CTA (ResCfg<M, K>::FW) for the host-side sizing of the partial-sum slots
int OIVA_CAT(resident_fw_m, OIVA_M)(int K) {
    constexpr int M = OIVA_M;
    switch (K) {
#define OIVA_FW_CASE(K_)                                               \
    case K_:                                                           \
        if constexpr (K_ <= M && !(M >= 7 && K_ > 4)) return ResCfg<M, (K_ <= M ? K_ : 1)>::FW; \
        else return 0;
        OIVA_FW_CASE(1) OIVA_FW_CASE(2) OIVA_FW_CASE(3) OIVA_FW_CASE(4)
        OIVA_FW_CASE(5) OIVA_FW_CASE(6) OIVA_FW_CASE(7) OIVA_FW_CASE(8)
    }
    return 0;
}

}  // namespace oiva
