/* sep_compat.c -- small host helpers the example programs call: allocators (reference
 * source/separray.c:16-213) and sep_dot (source/seputil.c:393-403).  The samplers live in sep_sampler.c. */
#include "sep_host.h"

double *sep_vector(size_t length)
{
    double *p = calloc(length ? length : 1, sizeof(double));     /* zero-filled like the reference */
    if (!p) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    return p;
}

int *sep_vector_int(size_t length)
{
    int *p = calloc(length ? length : 1, sizeof(int));
    if (!p) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    return p;
}

double **sep_matrix(size_t nrow, size_t ncol)
{
    double **m = malloc(sizeof(double *) * (nrow ? nrow : 1));
    if (!m) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    for (size_t r = 0; r < nrow; r++) m[r] = sep_vector(ncol);
    return m;
}

void sep_free_matrix(double **ptr, size_t nrow)
{
    for (size_t r = 0; r < nrow; r++) free(ptr[r]);
    free(ptr);
}

void sep_matrix_set(double **a, size_t nrow, size_t ncol, double value)
{
    for (size_t r = 0; r < nrow; r++)
        for (size_t c = 0; c < ncol; c++) a[r][c] = value;
}

float ***sep_tensor_float(size_t nx, size_t ny, size_t nz)
{
    float ***t = malloc(sizeof(float **) * (nx ? nx : 1));
    if (!t) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    for (size_t i = 0; i < nx; i++) {
        t[i] = malloc(sizeof(float *) * (ny ? ny : 1));
        if (!t[i]) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
        for (size_t j = 0; j < ny; j++) {
            t[i][j] = calloc(nz ? nz : 1, sizeof(float));
            if (!t[i][j]) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
        }
    }
    return t;
}

void sep_free_tensor_float(float ***ptr, size_t nx, size_t ny)
{
    for (size_t i = 0; i < nx; i++) {
        for (size_t j = 0; j < ny; j++) free(ptr[i][j]);
        free(ptr[i]);
    }
    free(ptr);
}

double sep_dot(double *a, double *b, int length)
{
    double s = 0.0;
    for (int n = 0; n < length; n++) s += a[n] * b[n];
    return s;
}

void sep_vector_set(double *vec, size_t length, double value)
{
    for (size_t n = 0; n < length; n++) vec[n] = value;
}
