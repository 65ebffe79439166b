/*
 * GSL-subset compatibility layer (build-time stand-in, written from scratch).
 *
 * GSL 2.7 is a dependency of the reference (javidsegura/gpu-montecarlo-risk:
 * Makefile:8 `-lgsl -lgslcblas`, src/model_interface.h:4-6) but is not
 * installed in this image and cannot be fetched (no network).  This header
 * declares ONLY the surface the reference's hot path touches
 * with struct layouts identical to GSL's public ones, so that
 *   - the reference's unmodified C sources compile here (oracle/_ref), and
 *   - the B200 plugin built against it is ABI-compatible with a host program
 *     built against the real GSL headers (gsl_vector / gsl_matrix field order
 *     and types are GSL's).
 * With real GSL installed, drop `-I third_party_compat` and link `-lgsl -lgslcblas`.
 *
 * Arithmetic that matters for parity (MT19937, polar Gaussian, Cholesky
 * decomp1, row-major dgemv) is restated from the published GSL 2.7 algorithms;
 * see gslcompat.c for what is pinned by known-answer tests and what is not.
 */
#ifndef MCR_GSL_COMPAT_CORE_H
#define MCR_GSL_COMPAT_CORE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- block / vector / matrix (layouts as in gsl_block_double.h,
 *      gsl_vector_double.h, gsl_matrix_double.h) ---- */
typedef struct { size_t size; double *data; } gsl_block;

typedef struct {
    size_t size;
    size_t stride;
    double *data;
    gsl_block *block;
    int owner;
} gsl_vector;

typedef struct {
    size_t size1;
    size_t size2;
    size_t tda;
    double *data;
    gsl_block *block;
    int owner;
} gsl_matrix;

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
double gsl_vector_get(const gsl_vector *v, size_t i);
void gsl_vector_set(gsl_vector *v, size_t i, double x);
int gsl_vector_add(gsl_vector *a, const gsl_vector *b);
int gsl_vector_scale(gsl_vector *a, double x);
int gsl_vector_memcpy(gsl_vector *dest, const gsl_vector *src);

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j);
void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x);
int gsl_matrix_memcpy(gsl_matrix *dest, const gsl_matrix *src);

/* ---- rng ---- */
typedef struct {
    const char *name;
    unsigned long int max;
    unsigned long int min;
    size_t size;
    void (*set)(void *state, unsigned long int seed);
    unsigned long int (*get)(void *state);
    double (*get_double)(void *state);
} gsl_rng_type;

typedef struct {
    const gsl_rng_type *type;
    void *state;
} gsl_rng;

extern const gsl_rng_type *gsl_rng_mt19937;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
void gsl_rng_free(gsl_rng *r);
unsigned long int gsl_rng_get(const gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
double gsl_rng_uniform_pos(const gsl_rng *r);

/* ---- randist ---- */
double gsl_ran_gaussian(const gsl_rng *r, double sigma);

/* ---- cblas enums / blas / linalg ---- */
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE_t;

int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A,
                   const gsl_vector *X, double beta, gsl_vector *Y);
int gsl_linalg_cholesky_decomp1(gsl_matrix *A);

#define GSL_SUCCESS 0
#define GSL_EDOM 1
#define GSL_EBADLEN 19
#define GSL_ENOTSQR 20

#ifdef __cplusplus
}
#endif
#endif
