#include "model_loader.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static long file_size(FILE *f) {
    long cur = ftell(f);
    if (fseek(f, 0, SEEK_END) != 0) return -1;
    long sz = ftell(f);
    fseek(f, cur, SEEK_SET);
    return sz;
}

int mcr_load_host_model(const char *mu_path, const char *sigma_path, McrHostModel *out) {
    memset(out, 0, sizeof(*out));
    FILE *fm = fopen(mu_path, "rb");
    if (!fm) { fprintf(stderr, "Error: Cannot open %s\n", mu_path); return -1; }
    int32_t N = 0;
    if (fread(&N, sizeof(int32_t), 1, fm) != 1 || N <= 0 || N > 1000000) {
        fprintf(stderr, "Error: Invalid dimension header in %s\n", mu_path);
        fclose(fm);
        return -1;
    }
    if (file_size(fm) != (long)(sizeof(int32_t) + (size_t)N * sizeof(double))) {
        fprintf(stderr, "Error: %s has the wrong length for N=%d\n", mu_path, N);
        fclose(fm);
        return -1;
    }
    out->mu = (double *)malloc((size_t)N * sizeof(double));
    if (!out->mu || fread(out->mu, sizeof(double), (size_t)N, fm) != (size_t)N) {
        fprintf(stderr, "Error: Failed to read data from %s\n", mu_path);
        fclose(fm);
        mcr_free_host_model(out);
        return -1;
    }
    fclose(fm);

    FILE *fs = fopen(sigma_path, "rb");
    if (!fs) { fprintf(stderr, "Error: Cannot open %s\n", sigma_path); mcr_free_host_model(out); return -1; }
    int32_t dims[2] = {0, 0};
    if (fread(dims, sizeof(int32_t), 2, fs) != 2 || dims[0] != N || dims[1] != N) {
        fprintf(stderr, "Error: %s is %dx%d but %s has N=%d\n", sigma_path, dims[0], dims[1], mu_path, N);
        fclose(fs);
        mcr_free_host_model(out);
        return -1;
    }
    const size_t nn = (size_t)N * (size_t)N;
    if (file_size(fs) != (long)(2 * sizeof(int32_t) + nn * sizeof(double))) {
        fprintf(stderr, "Error: %s has the wrong length for a %dx%d matrix\n", sigma_path, N, N);
        fclose(fs);
        mcr_free_host_model(out);
        return -1;
    }
    out->Sigma = (double *)malloc(nn * sizeof(double));
    if (!out->Sigma || fread(out->Sigma, sizeof(double), nn, fs) != nn) {
        fprintf(stderr, "Error: Failed to read data from %s\n", sigma_path);
        fclose(fs);
        mcr_free_host_model(out);
        return -1;
    }
    fclose(fs);
    out->N = N;
    return 0;
}

void mcr_free_host_model(McrHostModel *m) {
    if (!m) return;
    free(m->mu);
    free(m->Sigma);
    memset(m, 0, sizeof(*m));
}
