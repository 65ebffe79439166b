/*
 * Loader for the preprocessing binaries (data/mu.bin, data/sigma.bin; written by
 * preprocessing/save_binary.py:23-45, read in the reference by src/utilities/load_binary.c:13-36,56-87) into
 * plain host buffers ready for upload — no GSL containers in between.  Hardened where the reference
 * is not: file length is checked against the header, and mu/sigma dimensions must agree
 * (the reference trusts mu->size, main_runner.c:130); every check runs before anything is allocated.
 * The payload is read straight into buffers from a caller-supplied allocator, so a GPU caller stages it in
 * pinned memory (mcr_host_alloc_pinned of include/mcr_b200.h) and the H2D upload needs no bounce copy.
 */
#ifndef MCR_MODEL_LOADER_H
#define MCR_MODEL_LOADER_H

#include <stddef.h>

typedef struct {
    void *(*alloc)(size_t bytes); /* NULL = malloc */
    void (*free)(void *p);        /* NULL = free */
} McrHostAllocator;

typedef struct {
    int N;
    double *mu;     /* N */
    double *Sigma;  /* N x N row-major */
    McrHostAllocator alloc; /* what the buffers came from (used by mcr_free_host_model) */
} McrHostModel;

/* 0 on success; -1 on failure with a message on stderr.  Free with mcr_free_host_model. */
int mcr_load_host_model(const char *mu_path, const char *sigma_path, McrHostModel *out);
/* same, with the staging buffers taken from `alloc` (NULL = malloc / free) */
int mcr_load_host_model_ex(const char *mu_path, const char *sigma_path, const McrHostAllocator *alloc, McrHostModel *out);
void mcr_free_host_model(McrHostModel *m);

#endif
