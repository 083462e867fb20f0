/*
 * lapack_shim.c -- TEST INFRASTRUCTURE. Maps the three Fortran LAPACK symbols the reference Griffon links
 * against (blas_lapack_kernels.h:84,103,149: dgetrf_, dgetrs_, dgeev_) onto the `scipy_`-prefixed LP64 symbols
 * exported by SciPy's bundled OpenBLAS (site-packages/scipy.libs/libscipy_openblas-*.so). The reference's own
 * build links system blas/lapack (setup.py:112); this image has none, SciPy's is the LAPACK that is present.
 */
extern void scipy_dgetrf_(const int *m, const int *n, double *a, const int *lda, int *ipiv, int *info);
extern void scipy_dgetrs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *lda,
                          const int *ipiv, double *b, const int *ldb, int *info);
extern void scipy_dgeev_(const char *jobvl, const char *jobvr, const int *n, double *a, const int *lda, double *wr,
                         double *wi, double *vl, const int *ldvl, double *vr, const int *ldvr, double *work,
                         int *lwork, int *info);

void dgetrf_(const int *m, const int *n, double *a, const int *lda, int *ipiv, int *info)
{
  scipy_dgetrf_(m, n, a, lda, ipiv, info);
}
void dgetrs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *lda, const int *ipiv,
             double *b, const int *ldb, int *info)
{
  scipy_dgetrs_(trans, n, nrhs, a, lda, ipiv, b, ldb, info);
}
void dgeev_(const char *jobvl, const char *jobvr, const int *n, double *a, const int *lda, double *wr, double *wi,
            double *vl, const int *ldvl, double *vr, const int *ldvr, double *work, int *lwork, int *info)
{
  scipy_dgeev_(jobvl, jobvr, n, a, lda, wr, wi, vl, ldvl, vr, ldvr, work, lwork, info);
}
