/*
 * Transform-matrix construction for the B200 model: replaces the per-model
 *   gsl_matrix_memcpy(L, Sigma); gsl_linalg_cholesky_decomp1(L);
 * of the reference (src/02-C-serial/monte_carlo_serial.c:36-37, src/03-openMP/monte_carlo_omp.c:47-48,
 * src/05-GPU/monte_carlos_cuda.cu:84-108) with one fp64 host factorisation whose result is handed to
 * the device library as a dense row-major matrix.
 */
#ifndef MCR_CHOLESKY_UPLOAD_H
#define MCR_CHOLESKY_UPLOAD_H

#include <gsl/gsl_matrix.h>
#include <gsl/gsl_vector.h>

/* which "L" to apply (see include/mcr_b200.h MCR_TRANSFORM_*): read from env MCR_TRANSFORM
 * ("lower" = default, what the reference CUDA model computes; "gsl_full" = what C-serial/OpenMP compute). */
int mcr_transform_mode_from_env(void);

/* Flatten mu (respecting stride) and build A from Sigma (respecting tda).  Returns malloc'd buffers
 * the caller frees.  0 on success, -1 on failure (message on stderr, as the reference models do). */
int mcr_prepare_model(const gsl_vector *mu, const gsl_matrix *Sigma, int mode, double **mu_out, double **A_out);

#endif
