/*
 * Loader for the preprocessing binaries (data/mu.bin, data/sigma.bin; written by
 * preprocessing/save_binary.py:23-45, read in the reference by src/utilities/load_binary.c) into
 * plain host buffers ready for upload — no GSL containers in between.  Hardened where the reference
 * is not: file length is checked against the header, and mu/sigma dimensions must agree
 * (the reference trusts mu->size, main_runner.c:130).
 */
#ifndef MCR_MODEL_LOADER_H
#define MCR_MODEL_LOADER_H

typedef struct {
    int N;
    double *mu;     /* N */
    double *Sigma;  /* N x N row-major */
} McrHostModel;

/* 0 on success; -1 on failure with a message on stderr.  Free with mcr_free_host_model. */
int mcr_load_host_model(const char *mu_path, const char *sigma_path, McrHostModel *out);
void mcr_free_host_model(McrHostModel *m);

#endif
