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

static void *default_alloc(size_t bytes) { return malloc(bytes); }
static void default_free(void *p) { free(p); }

int mcr_load_host_model_ex(const char *mu_path, const char *sigma_path, const McrHostAllocator *alloc, McrHostModel *out) {
    memset(out, 0, sizeof(*out));
    out->alloc.alloc = (alloc && alloc->alloc) ? alloc->alloc : default_alloc;
    out->alloc.free = (alloc && alloc->free) ? alloc->free : default_free;
    FILE *fm = NULL, *fs = NULL;
    int rc = -1;
    /* every header / length / dimension check comes before the first allocation or payload read */
    fm = fopen(mu_path, "rb");
    if (!fm) { fprintf(stderr, "Error: Cannot open %s\n", mu_path); goto done; }
    int32_t N = 0;
    if (fread(&N, sizeof(int32_t), 1, fm) != 1 || N <= 0 || N > 1000000) {
        fprintf(stderr, "Error: Invalid dimension header in %s\n", mu_path);
        goto done;
    }
    if (file_size(fm) != (long)(sizeof(int32_t) + (size_t)N * sizeof(double))) {
        fprintf(stderr, "Error: %s has the wrong length for N=%d\n", mu_path, N);
        goto done;
    }
    fs = fopen(sigma_path, "rb");
    if (!fs) { fprintf(stderr, "Error: Cannot open %s\n", sigma_path); goto done; }
    int32_t dims[2] = {0, 0};
    if (fread(dims, sizeof(int32_t), 2, fs) != 2 || dims[0] != N || dims[1] != N) {
        fprintf(stderr, "Error: %s is %dx%d but %s has N=%d\n", sigma_path, dims[0], dims[1], mu_path, N);
        goto done;
    }
    const size_t nn = (size_t)N * (size_t)N;
    if (file_size(fs) != (long)(2 * sizeof(int32_t) + nn * sizeof(double))) {
        fprintf(stderr, "Error: %s has the wrong length for a %dx%d matrix\n", sigma_path, N, N);
        goto done;
    }
    /* payload straight into the caller's (pinned) staging buffers */
    out->mu = (double *)out->alloc.alloc((size_t)N * sizeof(double));
    out->Sigma = (double *)out->alloc.alloc(nn * sizeof(double));
    if (!out->mu || !out->Sigma) { fprintf(stderr, "Error: cannot allocate host staging for N=%d\n", N); goto done; }
    if (fread(out->mu, sizeof(double), (size_t)N, fm) != (size_t)N) { fprintf(stderr, "Error: Failed to read data from %s\n", mu_path); goto done; }
    if (fread(out->Sigma, sizeof(double), nn, fs) != nn) { fprintf(stderr, "Error: Failed to read data from %s\n", sigma_path); goto done; }
    out->N = N;
    rc = 0;
done:
    if (fm) fclose(fm);
    if (fs) fclose(fs);
    if (rc != 0) mcr_free_host_model(out);
    return rc;
}

int mcr_load_host_model(const char *mu_path, const char *sigma_path, McrHostModel *out) {
    return mcr_load_host_model_ex(mu_path, sigma_path, NULL, out);
}

void mcr_free_host_model(McrHostModel *m) {
    if (!m) return;
    void (*fr)(void *) = m->alloc.free ? m->alloc.free : default_free;
    if (m->mu) fr(m->mu);
    if (m->Sigma) fr(m->Sigma);
    memset(m, 0, sizeof(*m));
}
