/*
 * GSL-subset compatibility implementation (from scratch; see gsl/gsl_compat_core.h).
 *
 * What is pinned and what is not (also stated in DESIGN.md):
 *  - MT19937 (gsl_rng_mt19937): the 2002 Matsumoto/Nishimura generator with
 *    init_genrand seeding, seed 0 -> 4357 as GSL's rng/mt.c does.  PINNED by
 *    known-answer vectors generated with numpy's legacy MT19937 seeding
 *    (tests/golden/mt19937_kat.json) and by the classic "seed 4357, 1000th
 *    output = 1186927261" value.
 *  - gsl_rng_uniform_pos / gsl_ran_gaussian: GSL's polar (Marsaglia) method,
 *    restated from the published algorithm (randist/gauss.c): x,y = -1+2u,
 *    accept 0 < r2 <= 1, return sigma*y*sqrt(-2 ln r2 / r2); no cached second
 *    variate.  UNPINNED against a real libgsl (none is reachable here).
 *  - gsl_blas_dgemv (row-major, NoTrans): gslcblas' reference loop,
 *    y = beta*y; temp = sum_j x[j]*A[i][j] (sequential, unfused); y[i] += alpha*temp.
 *  - gsl_linalg_cholesky_decomp1: GSL >= 2.2 behaviour: the strict upper
 *    triangle keeps the ORIGINAL matrix entries (copied there "for later rcond
 *    calculation"), the lower triangle receives L.  N <= 24 uses the unblocked
 *    left-looking (dgemv-based) order; larger N uses GSL 2.6+'s recursive
 *    Level-3 split.  UNPINNED bit-for-bit, but any backward-stable Cholesky
 *    agrees to ~1e-16 relative.
 */
#include "gsl/gsl_compat_core.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ vector */
static gsl_block *block_alloc(size_t n, int zero) {
    gsl_block *b = (gsl_block *)malloc(sizeof(gsl_block));
    if (!b) return NULL;
    b->size = n;
    b->data = (double *)(zero ? calloc(n ? n : 1, sizeof(double)) : malloc((n ? n : 1) * sizeof(double)));
    if (!b->data) { free(b); return NULL; }
    return b;
}

static gsl_vector *vector_new(size_t n, int zero) {
    gsl_vector *v = (gsl_vector *)malloc(sizeof(gsl_vector));
    if (!v) return NULL;
    v->block = block_alloc(n, zero);
    if (!v->block) { free(v); return NULL; }
    v->size = n; v->stride = 1; v->data = v->block->data; v->owner = 1;
    return v;
}
gsl_vector *gsl_vector_alloc(size_t n) { return vector_new(n, 0); }
gsl_vector *gsl_vector_calloc(size_t n) { return vector_new(n, 1); }
void gsl_vector_free(gsl_vector *v) {
    if (!v) return;
    if (v->owner && v->block) { free(v->block->data); free(v->block); }
    free(v);
}
double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i * v->stride] = x; }
int gsl_vector_add(gsl_vector *a, const gsl_vector *b) {
    if (a->size != b->size) return GSL_EBADLEN;
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] += b->data[i * b->stride];
    return GSL_SUCCESS;
}
int gsl_vector_scale(gsl_vector *a, double x) {
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] *= x;
    return GSL_SUCCESS;
}
int gsl_vector_memcpy(gsl_vector *d, const gsl_vector *s) {
    if (d->size != s->size) return GSL_EBADLEN;
    for (size_t i = 0; i < s->size; i++) d->data[i * d->stride] = s->data[i * s->stride];
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ matrix */
static gsl_matrix *matrix_new(size_t n1, size_t n2, int zero) {
    gsl_matrix *m = (gsl_matrix *)malloc(sizeof(gsl_matrix));
    if (!m) return NULL;
    m->block = block_alloc(n1 * n2, zero);
    if (!m->block) { free(m); return NULL; }
    m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = m->block->data; m->owner = 1;
    return m;
}
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) { return matrix_new(n1, n2, 0); }
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2) { return matrix_new(n1, n2, 1); }
void gsl_matrix_free(gsl_matrix *m) {
    if (!m) return;
    if (m->owner && m->block) { free(m->block->data); free(m->block); }
    free(m);
}
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
int gsl_matrix_memcpy(gsl_matrix *d, const gsl_matrix *s) {
    if (d->size1 != s->size1 || d->size2 != s->size2) return GSL_EBADLEN;
    for (size_t i = 0; i < s->size1; i++)
        for (size_t j = 0; j < s->size2; j++) d->data[i * d->tda + j] = s->data[i * s->tda + j];
    return GSL_SUCCESS;
}

/* --------------------------------------------------------------------- rng */
#define MT_N 624
#define MT_M 397
typedef struct { unsigned long mt[MT_N]; int mti; } mt_state_t;

static void mt_set(void *vstate, unsigned long int s) {
    mt_state_t *st = (mt_state_t *)vstate;
    if (s == 0) s = 4357; /* GSL's default seed */
    st->mt[0] = s & 0xffffffffUL;
    for (int i = 1; i < MT_N; i++) {
        st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i);
        st->mt[i] &= 0xffffffffUL;
    }
    st->mti = MT_N;
}

static unsigned long int mt_get(void *vstate) {
    mt_state_t *st = (mt_state_t *)vstate;
    unsigned long *const mt = st->mt;
    const unsigned long UPPER = 0x80000000UL, LOWER = 0x7fffffffUL;
#define MT_MAGIC(y) (((y) & 0x1UL) ? 0x9908b0dfUL : 0UL)
    if (st->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ MT_MAGIC(y);
        }
        for (; kk < MT_N - 1; kk++) {
            unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ MT_MAGIC(y);
        }
        {
            unsigned long y = (mt[MT_N - 1] & UPPER) | (mt[0] & LOWER);
            mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ MT_MAGIC(y);
        }
        st->mti = 0;
    }
#undef MT_MAGIC
    unsigned long k = mt[st->mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680UL;
    k ^= (k << 15) & 0xefc60000UL;
    k ^= (k >> 18);
    return k & 0xffffffffUL;
}

static double mt_get_double(void *vstate) { return (double)mt_get(vstate) / 4294967296.0; }

static const gsl_rng_type mt_type = {
    "mt19937", 0xffffffffUL, 0, sizeof(mt_state_t), &mt_set, &mt_get, &mt_get_double};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
    gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
    if (!r) return NULL;
    r->state = calloc(1, T->size);
    if (!r->state) { free(r); return NULL; }
    r->type = T;
    T->set(r->state, 0); /* gsl_rng_default_seed == 0 */
    return r;
}
void gsl_rng_set(const gsl_rng *r, unsigned long int seed) { r->type->set(r->state, seed); }
void gsl_rng_free(gsl_rng *r) { if (r) { free(r->state); free(r); } }
unsigned long int gsl_rng_get(const gsl_rng *r) { return r->type->get(r->state); }
double gsl_rng_uniform(const gsl_rng *r) { return r->type->get_double(r->state); }
double gsl_rng_uniform_pos(const gsl_rng *r) {
    double x;
    do { x = r->type->get_double(r->state); } while (x == 0);
    return x;
}

/* ----------------------------------------------------------------- randist */
double gsl_ran_gaussian(const gsl_rng *r, const double sigma) {
    double x, y, r2;
    do {
        x = -1 + 2 * gsl_rng_uniform_pos(r);
        y = -1 + 2 * gsl_rng_uniform_pos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

/* -------------------------------------------------------------------- blas */
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A,
                   const gsl_vector *X, double beta, gsl_vector *Y) {
    const size_t M = A->size1, N = A->size2;
    const size_t lenX = (TransA == CblasNoTrans) ? N : M;
    const size_t lenY = (TransA == CblasNoTrans) ? M : N;
    if (X->size != lenX || Y->size != lenY) return GSL_EBADLEN;
    if (beta == 0.0) {
        for (size_t i = 0; i < lenY; i++) Y->data[i * Y->stride] = 0.0;
    } else if (beta != 1.0) {
        for (size_t i = 0; i < lenY; i++) Y->data[i * Y->stride] *= beta;
    }
    if (alpha == 0.0) return GSL_SUCCESS;
    if (TransA == CblasNoTrans) {
        for (size_t i = 0; i < lenY; i++) {
            double temp = 0.0;
            for (size_t j = 0; j < lenX; j++) temp += X->data[j * X->stride] * A->data[A->tda * i + j];
            Y->data[i * Y->stride] += alpha * temp;
        }
    } else {
        for (size_t j = 0; j < lenX; j++) {
            const double temp = alpha * X->data[j * X->stride];
            if (temp != 0.0)
                for (size_t i = 0; i < lenY; i++) Y->data[i * Y->stride] += temp * A->data[A->tda * j + i];
        }
    }
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ linalg */
/* views on a row-major parent: (data pointer, rows, cols, tda) */
typedef struct { double *d; size_t n1, n2, tda; } mview;

static int chol_unblocked(mview A) {
    /* left-looking, column j: v = A(j:n,j) - A(j:n,0:j) * A(j,0:j)^T ; scale by 1/sqrt(ajj) */
    const size_t N = A.n1;
    for (size_t j = 0; j < N; j++) {
        if (j > 0) {
            for (size_t i = j; i < N; i++) {
                double temp = 0.0;
                for (size_t k = 0; k < j; k++) temp += A.d[j * A.tda + k] * A.d[i * A.tda + k];
                A.d[i * A.tda + j] += -1.0 * temp;
            }
        }
        double ajj = A.d[j * A.tda + j];
        if (!(ajj > 0.0)) return GSL_EDOM;
        ajj = sqrt(ajj);
        const double inv = 1.0 / ajj;
        for (size_t i = j; i < N; i++) A.d[i * A.tda + j] *= inv;
    }
    return GSL_SUCCESS;
}

static int chol_recursive(mview A) {
    const size_t N = A.n1;
    if (N <= 24) return chol_unblocked(A);
    const size_t N1 = (N >= 16) ? ((N + 8) / 16) * 8 : N / 2;
    const size_t N2 = N - N1;
    mview A11 = {A.d, N1, N1, A.tda};
    mview A21 = {A.d + N1 * A.tda, N2, N1, A.tda};
    mview A22 = {A.d + N1 * A.tda + N1, N2, N2, A.tda};
    int status = chol_recursive(A11);
    if (status) return status;
    /* A21 := A21 * L11^{-T}  (dtrsm Right, Lower, Trans, NonUnit) */
    for (size_t i = 0; i < N2; i++) {
        for (size_t j = 0; j < N1; j++) {
            A21.d[i * A.tda + j] /= A11.d[j * A.tda + j];
            const double Bij = A21.d[i * A.tda + j];
            for (size_t k = j + 1; k < N1; k++) A21.d[i * A.tda + k] -= A11.d[k * A.tda + j] * Bij;
        }
    }
    /* A22 := A22 - L21 L21^T on the lower triangle (dsyrk Lower, NoTrans) */
    for (size_t i = 0; i < N2; i++) {
        for (size_t j = 0; j <= i; j++) {
            double temp = 0.0;
            for (size_t k = 0; k < N1; k++) temp += A21.d[i * A.tda + k] * A21.d[j * A.tda + k];
            A22.d[i * A.tda + j] += -1.0 * temp;
        }
    }
    return chol_recursive(A22);
}

int gsl_linalg_cholesky_decomp1(gsl_matrix *A) {
    const size_t N = A->size1;
    if (N != A->size2) {
        fprintf(stderr, "gsl: cholesky.c: ERROR: Cholesky decomposition requires square matrix\n");
        return GSL_ENOTSQR;
    }
    /* keep the original matrix in the strict upper triangle (GSL: "for later rcond calculation") */
    for (size_t i = 0; i < N; i++)
        for (size_t j = 0; j < i; j++) A->data[j * A->tda + i] = A->data[i * A->tda + j];
    mview V = {A->data, N, N, A->tda};
    int status = chol_recursive(V);
    if (status) fprintf(stderr, "gsl: cholesky.c: ERROR: matrix is not positive definite\n");
    return status;
}
