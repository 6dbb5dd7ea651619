/* Plain-C caller of the C ABI (include/overiva_b200.h): reads X (B,T,F,M) complex128 from a raw file, runs
 * oiva_overiva_host, writes Y (B,T,F,K) and W (B,F,M,K).  Built and run by tests/test_c_abi.py.
 *   host_call <in.bin> <out.bin> B T F M K n_iter model init proj_back */
#include <stdio.h>
#include <stdlib.h>

#include "overiva_b200.h"

int main(int argc, char** argv) {
    if (argc != 12) {
        fprintf(stderr, "usage: host_call in out B T F M K n_iter model init proj_back\n");
        return 2;
    }
    const int B = atoi(argv[3]), T = atoi(argv[4]), F = atoi(argv[5]), M = atoi(argv[6]), K = atoi(argv[7]);
    const int n_iter = atoi(argv[8]), model = atoi(argv[9]), init = atoi(argv[10]), proj_back = atoi(argv[11]);
    const size_t nx = (size_t)B * T * F * M * 2, ny = (size_t)B * T * F * K * 2, nw = (size_t)B * F * M * K * 2;
    double* X = (double*)malloc(nx * sizeof(double));
    double* Y = (double*)malloc(ny * sizeof(double));
    double* W = (double*)malloc(nw * sizeof(double));
    if (!X || !Y || !W) return 3;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(X, sizeof(double), nx, f) != nx) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 4;
    }
    fclose(f);
    oiva_plan_desc d = {B, T, F, 0, M, K, model, OIVA_C128, 0};
    int* per_mixture = (int*)calloc((size_t)B, sizeof(int));
    const int status = oiva_overiva_host(X, Y, W, NULL, &d, n_iter, proj_back, init, per_mixture);
    if (status < 0) {
        fprintf(stderr, "oiva_overiva_host failed (%d): %s\n", status, oiva_last_error());
        return 5;
    }
    if (status & OIVA_STATUS_SINGULAR) {
        for (int b = 0; b < B; ++b)
            if (per_mixture[b] & OIVA_STATUS_SINGULAR) fprintf(stderr, "mixture %d: singular\n", b);
        return 6;
    }
    f = fopen(argv[2], "wb");
    if (!f || fwrite(Y, sizeof(double), ny, f) != ny || fwrite(W, sizeof(double), nw, f) != nw) return 7;
    fclose(f);
    printf("version %d ok\n", oiva_version());
    free(X);
    free(Y);
    free(W);
    free(per_mixture);
    return 0;
}
