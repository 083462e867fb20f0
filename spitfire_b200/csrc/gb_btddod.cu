// gb_btddod.cu -- batched block-tridiagonal (diagonal off-diagonal blocks) factor / solve / matvec on sm_100a.
//
// Replaces griffon::btddod::btddod_full_factorize / _full_solve / _full_matvec / _scale_and_add_diagonal
// (btddod_matrix_kernels.cpp:19-165, 447-465) and the LAPACK dgetrf/dgetrs calls inside them
// (blas_lapack_kernels.h:84-124). Storage is the reference's BTDDOD layout: num_blocks column-major bs x bs diagonal
// blocks, then (num_blocks-1)*bs sub-diagonal and (num_blocks-1)*bs super-diagonal scalars.
//
// One CTA owns one system (one flamelet); the block recurrence D_i <- D_i - diag(sub) D_{i-1}^{-1} diag(sup) is
// sequential in i, the work inside a block is spread over the CTA with the current block and the inverse being
// built resident in shared memory. Partial pivoting follows dgetrf: first row of maximum modulus, 1-based pivots,
// row interchanges applied to the whole block; the pivot search is a warp-shuffle arg-max.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <string>

#include "../../include/griffon_b200.h"
#include "gb_mech.h"

namespace gb
{
extern std::atomic<long> g_btddod_launches;
std::atomic<long> g_btddod_launches{0};

#define LD(bs) ((bs) | 1) // odd leading dimension in shared memory: conflict-free row and column walks

// in-place LU with partial pivoting of the bs x bs block A (shared, leading dimension ld). piv: 1-based, shared int.
__device__ void lu_factor_smem(double *A, int bs, int ld, int *piv)
{
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < bs; ++k)
  {
    if (tid < 32)
    { // warp 0: arg-max of |A[k..bs-1, k]|, ties -> smallest row (idamax)
      double best = -1.;
      int bi = k;
      for (int i = k + tid; i < bs; i += 32)
      {
        const double v = fabs(A[i + k * ld]);
        if (v > best)
        {
          best = v;
          bi = i;
        }
      }
      for (int off = 16; off > 0; off >>= 1)
      {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi))
        {
          best = ov;
          bi = oi;
        }
      }
      if (tid == 0)
        piv[k] = bi + 1;
    }
    __syncthreads();
    const int p = piv[k] - 1;
    if (p != k)
      for (int j = tid; j < bs; j += nt)
      {
        const double t = A[k + j * ld];
        A[k + j * ld] = A[p + j * ld];
        A[p + j * ld] = t;
      }
    __syncthreads();
    const double pivot = A[k + k * ld];
    if (pivot != 0.)
    {
      const double rp = 1. / pivot; // dgetf2 scales by the reciprocal
      // each thread scales its own multipliers on the fly and applies the rank-1 update of its entries
      const int m = bs - k - 1;
      for (int e = tid; e < m * m; e += nt)
      {
        const int i = k + 1 + e % m, j = k + 1 + e / m;
        A[i + j * ld] -= (A[i + k * ld] * rp) * A[k + j * ld];
      }
      __syncthreads();
      for (int i = k + 1 + tid; i < bs; i += nt)
        A[i + k * ld] *= rp;
    }
    __syncthreads();
  }
}

// X <- A^{-1} given the LU factors and pivots (dgetrs with the identity as right-hand side,
// btddod_matrix_kernels.cpp:48-53): X = U^{-1} L^{-1} P
__device__ void lu_inverse_smem(const double *A, int bs, int ld, const int *piv, double *X)
{
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < bs * bs; e += nt)
  {
    const int i = e % bs, j = e / bs;
    X[i + j * ld] = (i == j) ? 1. : 0.;
  }
  __syncthreads();
  // apply the row interchanges to the identity (dlaswp): sequential in k, parallel over columns
  for (int j = tid; j < bs; j += nt)
    for (int k = 0; k < bs; ++k)
    {
      const int p = piv[k] - 1;
      if (p != k)
      {
        const double t = X[k + j * ld];
        X[k + j * ld] = X[p + j * ld];
        X[p + j * ld] = t;
      }
    }
  __syncthreads();
  // forward substitution with unit lower L
  for (int k = 0; k < bs - 1; ++k)
  {
    const int m = bs - k - 1;
    for (int e = tid; e < m * bs; e += nt)
    {
      const int i = k + 1 + e % m, j = e / m;
      X[i + j * ld] -= A[i + k * ld] * X[k + j * ld];
    }
    __syncthreads();
  }
  // backward substitution with U
  for (int k = bs - 1; k >= 0; --k)
  {
    const double ukk = A[k + k * ld];
    for (int j = tid; j < bs; j += nt)
      X[k + j * ld] = X[k + j * ld] / ukk;
    __syncthreads();
    for (int e = tid; e < k * bs; e += nt)
    {
      const int i = e % k, j = e / k;
      X[i + j * ld] -= A[i + k * ld] * X[k + j * ld];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_btddod_factorize(int nsys, double *d_factors, int nb, int bs,
                                                          double *l_values, int *pivots)
{
  extern __shared__ double sm[];
  const int ld = LD(bs);
  double *A = sm;            // current diagonal block / its LU factors
  double *X = A + ld * bs;   // inverse of the previous block -> L_i
  int *spiv = (int *)(X + ld * bs);
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t nb2 = (size_t)bs * bs;
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));
  for (int sys = blockIdx.x; sys < nsys; sys += gridDim.x)
  {
    double *D = d_factors + (size_t)sys * mat_stride;
    double *Lv = l_values + (size_t)sys * nb * nb2;
    int *piv = pivots + (size_t)sys * nb * bs;
    const double *sub = D + (size_t)nb * nb2;
    const double *sup = sub + (size_t)(nb - 1) * bs;
    __syncthreads();
    for (int e = tid; e < bs * bs; e += nt)
      A[e % bs + (e / bs) * ld] = D[e];
    __syncthreads();
    for (int i = 0; i < nb; ++i)
    {
      lu_factor_smem(A, bs, ld, spiv);
      for (int e = tid; e < bs * bs; e += nt)
        D[(size_t)i * nb2 + e] = A[e % bs + (e / bs) * ld];
      for (int k = tid; k < bs; k += nt)
        piv[(size_t)i * bs + k] = spiv[k];
      if (i == nb - 1)
        break;
      lu_inverse_smem(A, bs, ld, spiv, X);
      // L_{i+1} = diag(sub_i) * D_i^{-1} (row scaling, :55-63); D_{i+1} -= L_{i+1} * diag(sup_i) (:65-75)
      const double *subi = sub + (size_t)i * bs, *supi = sup + (size_t)i * bs;
      const double *Dn = D + (size_t)(i + 1) * nb2;
      double *Ln = Lv + (size_t)(i + 1) * nb2;
      for (int e = tid; e < bs * bs; e += nt)
      {
        const int r = e % bs, c = e / bs;
        const double l = X[r + c * ld] * subi[r];
        Ln[e] = l;
        A[r + c * ld] = Dn[e] + (-supi[c]) * l;
      }
      __syncthreads();
    }
    if (nb > 0)
      for (int e = tid; e < bs * bs; e += nt)
        Lv[e] = 0.; // block 0 of l_values is never referenced by the reference; define it
  }
}

// forward: y_i = b_i - L_i y_{i-1} (:95-103); back: x_i = D_i^{-1} (y_i - sup_i o x_{i+1}) (:105-118)
__global__ void __launch_bounds__(128) k_btddod_solve(int nsys, const double *d_factors, const double *l_values,
                                                      const int *pivots, const double *rhs, int nb, int bs,
                                                      double *solution)
{
  extern __shared__ double sm[];
  const int ld = LD(bs);
  double *A = sm;          // LU factors of the current block
  double *v = A + ld * bs; // working vector [bs]
  double *vp = v + bs;     // previous y / next x [bs]
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t nb2 = (size_t)bs * bs;
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));
  for (int sys = blockIdx.x; sys < nsys; sys += gridDim.x)
  {
    const double *D = d_factors + (size_t)sys * mat_stride;
    const double *Lv = l_values + (size_t)sys * nb * nb2;
    const int *piv = pivots + (size_t)sys * nb * bs;
    const double *sup = D + (size_t)nb * nb2 + (size_t)(nb - 1) * bs;
    const double *b = rhs + (size_t)sys * nb * bs;
    double *x = solution + (size_t)sys * nb * bs; // y is stored in x during the forward sweep
    __syncthreads();
    for (int j = tid; j < bs; j += nt)
    {
      vp[j] = b[j];
      x[j] = b[j];
    }
    __syncthreads();
    for (int i = 1; i < nb; ++i)
    {
      const double *L = Lv + (size_t)i * nb2;
      for (int j = tid; j < bs; j += nt)
      {
        double yj = b[(size_t)i * bs + j];
        for (int c = 0; c < bs; ++c)
          yj = yj + L[(size_t)c * bs + j] * (-1. * vp[c]); // matrix_vector_multiply order, blas_lapack_kernels.h:60-77
        v[j] = yj;
      }
      __syncthreads();
      for (int j = tid; j < bs; j += nt)
      {
        vp[j] = v[j];
        x[(size_t)i * bs + j] = v[j];
      }
      __syncthreads();
    }
    for (int i = nb - 1; i >= 0; --i)
    {
      // load factors, form the right-hand side
      for (int e = tid; e < bs * bs; e += nt)
        A[e % bs + (e / bs) * ld] = D[(size_t)i * nb2 + e];
      for (int j = tid; j < bs; j += nt)
      {
        const double yj = x[(size_t)i * bs + j];
        v[j] = (i == nb - 1) ? yj : yj - sup[(size_t)i * bs + j] * vp[j];
      }
      __syncthreads();
      if (tid < 32)
      { // dgetrs on one vector by one warp: row interchanges, unit-lower forward, upper backward
        if (tid == 0)
          for (int k = 0; k < bs; ++k)
          {
            const int p = piv[(size_t)i * bs + k] - 1;
            if (p != k)
            {
              const double t = v[k];
              v[k] = v[p];
              v[p] = t;
            }
          }
        __syncwarp();
        for (int k = 0; k < bs - 1; ++k)
        {
          const double vk = v[k];
          for (int r = k + 1 + tid; r < bs; r += 32)
            v[r] -= A[r + k * ld] * vk;
          __syncwarp();
        }
        for (int k = bs - 1; k >= 0; --k)
        {
          if (tid == 0)
            v[k] = v[k] / A[k + k * ld];
          __syncwarp();
          const double vk = v[k];
          for (int r = tid; r < k; r += 32)
            v[r] -= A[r + k * ld] * vk;
          __syncwarp();
        }
      }
      __syncthreads();
      for (int j = tid; j < bs; j += nt)
      {
        vp[j] = v[j];
        x[(size_t)i * bs + j] = v[j];
      }
      __syncthreads();
    }
  }
}

__global__ void k_btddod_matvec(int nsys, const double *matrix, const double *vec, int nb, int bs, double *out)
{
  // one thread per output row: block-diagonal matvec in the reference's column order, then the off-diagonals
  const size_t nb2 = (size_t)bs * bs;
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));
  const size_t total = (size_t)nsys * nb * bs;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
  {
    const int sys = (int)(t / ((size_t)nb * bs));
    const int rem = (int)(t - (size_t)sys * nb * bs);
    const int iz = rem / bs, j = rem - iz * bs;
    const double *A = matrix + (size_t)sys * mat_stride;
    const double *x = vec + (size_t)sys * nb * bs;
    const double *sub = A + (size_t)nb * nb2, *sup = sub + (size_t)(nb - 1) * bs;
    double y = 0.;
    for (int c = 0; c < bs; ++c)
      y = y + A[(size_t)iz * nb2 + (size_t)c * bs + j] * (1. * x[iz * bs + c]);
    if (iz > 0 && iz < nb - 1)
      y += sup[iz * bs + j] * x[(iz + 1) * bs + j] + sub[(iz - 1) * bs + j] * x[(iz - 1) * bs + j];
    else if (iz == nb - 1 && nb > 1)
      y += sub[(iz - 1) * bs + j] * x[(iz - 1) * bs + j];
    else if (iz == 0 && nb > 1)
      y += sup[j] * x[bs + j];
    out[t] = y;
  }
}

__global__ void k_btddod_scale_add_diag(int nsys, double *matrix, double ms, const double *diag, double ds, int nb,
                                        int bs)
{
  const size_t nb2 = (size_t)bs * bs;
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));
  const size_t total = (size_t)nsys * mat_stride;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
  {
    const size_t sys = t / mat_stride, e = t - sys * mat_stride;
    double v = matrix[t] * ms;
    if (e < (size_t)nb * nb2)
    {
      const size_t iz = e / nb2, w = e - iz * nb2;
      const int r = (int)(w % bs), c = (int)(w / bs);
      if (r == c)
        v += ds * diag[sys * nb * bs + iz * bs + r];
    }
    matrix[t] = v;
  }
}

} // namespace gb

using namespace gb;

namespace
{
int bt_fail(cudaError_t e, const char *what)
{
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  cudaGetLastError();
  return GB_ERR_CUDA;
}
int bt_check(int n, int nb, int bs)
{
  if (n < 0 || nb < 1 || bs < 1)
  {
    set_error("bad btddod dimensions");
    return GB_ERR_ARG;
  }
  const size_t need = sizeof(double) * 2 * (size_t)LD(bs) * bs + sizeof(int) * bs + 64;
  if (need > 227 * 1024)
  {
    set_error("block size too large for the in-shared-memory block LU");
    return GB_ERR_UNSUPPORTED;
  }
  return GB_OK;
}
int sm_count_bt()
{
  int dev = 0, n = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
#define BCK(call)                     \
  do                                  \
  {                                   \
    cudaError_t e__ = (call);         \
    if (e__ != cudaSuccess)           \
      return bt_fail(e__, #call);     \
  } while (0)
} // namespace

extern "C"
{
  int gb_btddod_full_factorize_batch(int n, double *d_factors, int nb, int bs, double *l_values, int *pivots,
                                     void *stream)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t smem = sizeof(double) * 2 * (size_t)LD(bs) * bs + sizeof(int) * bs + 64;
    BCK(cudaFuncSetAttribute(k_btddod_factorize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_btddod_factorize<<<n, 256, smem, (cudaStream_t)stream>>>(n, d_factors, nb, bs, l_values, pivots);
    ++g_btddod_launches;
    BCK(cudaGetLastError());
    return GB_OK;
  }

  int gb_btddod_full_solve_batch(int n, const double *d_factors, const double *l_values, const int *pivots,
                                 const double *rhs, int nb, int bs, double *solution, void *stream)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t smem = sizeof(double) * ((size_t)LD(bs) * bs + 2 * bs) + 64;
    BCK(cudaFuncSetAttribute(k_btddod_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_btddod_solve<<<n, 128, smem, (cudaStream_t)stream>>>(n, d_factors, l_values, pivots, rhs, nb, bs, solution);
    ++g_btddod_launches;
    BCK(cudaGetLastError());
    return GB_OK;
  }

  int gb_btddod_full_matvec_batch(int n, const double *matrix, const double *vec, int nb, int bs, double *out,
                                  void *stream)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t total = (size_t)n * nb * bs;
    const int threads = 128;
    const int grid = (int)std::min<size_t>((total + threads - 1) / threads, (size_t)sm_count_bt() * 16);
    k_btddod_matvec<<<grid, threads, 0, (cudaStream_t)stream>>>(n, matrix, vec, nb, bs, out);
    ++g_btddod_launches;
    BCK(cudaGetLastError());
    return GB_OK;
  }

  int gb_btddod_scale_and_add_diagonal_batch(int n, double *matrix, double ms, const double *diag, double ds, int nb,
                                             int bs, void *stream)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t total = (size_t)n * bs * ((size_t)nb * bs + 2 * (nb - 1));
    const int threads = 256;
    const int grid = (int)std::min<size_t>((total + threads - 1) / threads, (size_t)sm_count_bt() * 16);
    k_btddod_scale_add_diag<<<grid, threads, 0, (cudaStream_t)stream>>>(n, matrix, ms, diag, ds, nb, bs);
    ++g_btddod_launches;
    BCK(cudaGetLastError());
    return GB_OK;
  }

  // ---- host-pointer variants (single-system calls of the reference API, griffon.pyx:1081-1113) -------------------
  static int bt_stage(void **d, const void *h, size_t bytes)
  {
    BCK(cudaMalloc(d, bytes ? bytes : 8));
    if (h && bytes)
      BCK(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
    return GB_OK;
  }

  int gb_btddod_full_factorize_host(int n, double *d_factors, int nb, int bs, double *l_values, int *pivots)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t mb = sizeof(double) * (size_t)n * bs * ((size_t)nb * bs + 2 * (nb - 1));
    const size_t lb = sizeof(double) * (size_t)n * nb * bs * bs, pb = sizeof(int) * (size_t)n * nb * bs;
    void *dm = nullptr, *dl = nullptr, *dp = nullptr;
    if ((rc = bt_stage(&dm, d_factors, mb)) || (rc = bt_stage(&dl, nullptr, lb)) || (rc = bt_stage(&dp, nullptr, pb)))
      return rc;
    rc = gb_btddod_full_factorize_batch(n, (double *)dm, nb, bs, (double *)dl, (int *)dp, nullptr);
    if (rc == GB_OK)
    {
      BCK(cudaMemcpy(d_factors, dm, mb, cudaMemcpyDeviceToHost));
      BCK(cudaMemcpy(l_values, dl, lb, cudaMemcpyDeviceToHost));
      BCK(cudaMemcpy(pivots, dp, pb, cudaMemcpyDeviceToHost));
    }
    cudaFree(dm), cudaFree(dl), cudaFree(dp);
    return rc;
  }

  int gb_btddod_full_solve_host(int n, const double *d_factors, const double *l_values, const int *pivots,
                                const double *rhs, int nb, int bs, double *solution)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t mb = sizeof(double) * (size_t)n * bs * ((size_t)nb * bs + 2 * (nb - 1));
    const size_t lb = sizeof(double) * (size_t)n * nb * bs * bs, pb = sizeof(int) * (size_t)n * nb * bs;
    const size_t vb = sizeof(double) * (size_t)n * nb * bs;
    void *dm = nullptr, *dl = nullptr, *dp = nullptr, *dr = nullptr, *dx = nullptr;
    if ((rc = bt_stage(&dm, d_factors, mb)) || (rc = bt_stage(&dl, l_values, lb)) || (rc = bt_stage(&dp, pivots, pb)) ||
        (rc = bt_stage(&dr, rhs, vb)) || (rc = bt_stage(&dx, nullptr, vb)))
      return rc;
    rc = gb_btddod_full_solve_batch(n, (double *)dm, (double *)dl, (int *)dp, (double *)dr, nb, bs, (double *)dx,
                                    nullptr);
    if (rc == GB_OK)
      BCK(cudaMemcpy(solution, dx, vb, cudaMemcpyDeviceToHost));
    cudaFree(dm), cudaFree(dl), cudaFree(dp), cudaFree(dr), cudaFree(dx);
    return rc;
  }

  int gb_btddod_full_matvec_host(int n, const double *matrix, const double *vec, int nb, int bs, double *out)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t mb = sizeof(double) * (size_t)n * bs * ((size_t)nb * bs + 2 * (nb - 1));
    const size_t vb = sizeof(double) * (size_t)n * nb * bs;
    void *dm = nullptr, *dv = nullptr, *dout = nullptr;
    if ((rc = bt_stage(&dm, matrix, mb)) || (rc = bt_stage(&dv, vec, vb)) || (rc = bt_stage(&dout, nullptr, vb)))
      return rc;
    rc = gb_btddod_full_matvec_batch(n, (double *)dm, (double *)dv, nb, bs, (double *)dout, nullptr);
    if (rc == GB_OK)
      BCK(cudaMemcpy(out, dout, vb, cudaMemcpyDeviceToHost));
    cudaFree(dm), cudaFree(dv), cudaFree(dout);
    return rc;
  }

  int gb_btddod_scale_and_add_diagonal_host(int n, double *matrix, double ms, const double *diag, double ds, int nb,
                                            int bs)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t mb = sizeof(double) * (size_t)n * bs * ((size_t)nb * bs + 2 * (nb - 1));
    const size_t vb = sizeof(double) * (size_t)n * nb * bs;
    void *dm = nullptr, *dv = nullptr;
    if ((rc = bt_stage(&dm, matrix, mb)) || (rc = bt_stage(&dv, diag, vb)))
      return rc;
    rc = gb_btddod_scale_and_add_diagonal_batch(n, (double *)dm, ms, (double *)dv, ds, nb, bs, nullptr);
    if (rc == GB_OK)
      BCK(cudaMemcpy(matrix, dm, mb, cudaMemcpyDeviceToHost));
    cudaFree(dm), cudaFree(dv);
    return rc;
  }
}
