#include "cholesky_upload.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcr_b200.h"

int mcr_transform_mode_from_env(void) {
    const char *m = getenv("MCR_TRANSFORM");
    if (!m || !*m || strcmp(m, "lower") == 0) return MCR_TRANSFORM_LOWER;
    if (strcmp(m, "gsl_full") == 0) return MCR_TRANSFORM_GSL_FULL;
    fprintf(stderr, "Warning: unknown MCR_TRANSFORM='%s' (use lower|gsl_full); using lower\n", m);
    return MCR_TRANSFORM_LOWER;
}

int mcr_prepare_model(const gsl_vector *mu, const gsl_matrix *Sigma, int mode, double **mu_out, double **A_out) {
    *mu_out = NULL;
    *A_out = NULL;
    if (!mu || !Sigma || Sigma->size1 != Sigma->size2 || Sigma->size1 != mu->size || mu->size == 0) {
        fprintf(stderr, "Error: mu/Sigma dimensions are inconsistent\n");
        return -1;
    }
    const size_t N = mu->size;
    double *m = (double *)malloc(N * sizeof(double));
    double *A = (double *)malloc(N * N * sizeof(double));
    if (!m || !A) {
        fprintf(stderr, "Error: Failed to allocate host model buffers\n");
        free(m); free(A);
        return -1;
    }
    for (size_t i = 0; i < N; i++) m[i] = mu->data[i * mu->stride];
    /* Sigma->tda is the row pitch; mcr_build_transform reads rows with that leading dimension.
     * MCR_DEVICE_CHOLESKY=1: the same factorisation on the GPU (mcr_device_transform: same summation order, bitwise the same
     * factor; worth it from a few hundred assets up). */
    int rc;
    const char *dev = getenv("MCR_DEVICE_CHOLESKY");
    if (dev && *dev && strcmp(dev, "0") != 0) {
        mcr_engine *e = NULL;
        const char *fd = getenv("MCR_FIRST_DEVICE");
        rc = mcr_create(fd && *fd ? atoi(fd) : 0, &e);
        if (rc == 0) rc = mcr_device_transform(e, Sigma->data, (int)N, (int)Sigma->tda, mode, A);
        mcr_destroy(e);
    } else {
        rc = mcr_build_transform(Sigma->data, (int)N, (int)Sigma->tda, mode, A);
    }
    if (rc != 0) {
        fprintf(stderr, "Error: Cholesky decomposition failed (%s)\n", mcr_last_error());
        free(m); free(A);
        return -1;
    }
    *mu_out = m;
    *A_out = A;
    return 0;
}
