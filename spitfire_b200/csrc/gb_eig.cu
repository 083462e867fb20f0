// Largest real part of the eigenvalues of many small dense matrices (the explosive-mode bound of the pseudo-transient
// flamelet solver).
//
// Replaces griffon::lapack::eigenvalues (blas_lapack_kernels.h:157-180: LAPACK dgeev, eigenvalues only) as it is used by
// flamelet_jacobian (flamelet_kernels.cpp:1329-1341): per grid point only max_i Re(lambda_i) of the ns x ns chemical
// block survives, so nothing but that number leaves the kernel.
//
// One warp per matrix, the matrix resident in shared memory (ns = 53: 22.5 KB, nine matrices per SM). The stages are the
// ones dgeev runs for a matrix of this size (below LAPACK's multishift crossover):
//   1. balancing by powers of two (norm-reducing diagonal similarity, EISPACK "balanc" without the permutation search),
//   2. Householder reduction to upper Hessenberg form,
//   3. Francis double-shift QR on the Hessenberg matrix with deflation, eigenvalues only (EISPACK "hqr").
// The 32 lanes split rows (stage 2 from the left, stage 3 row update) or columns (from the right, column update);
// the leading dimension is odd so that both directions are free of bank conflicts for 8-byte accesses. The scalar
// recurrences (shifts, reflectors, deflation tests) are computed redundantly by all lanes from broadcast shared-memory
// reads, which keeps every branch warp-uniform.
// Since the spectrum of the transpose is the same, the column-major block is read as if it were row-major (coalesced).
#include "gb_kernels.cuh"

#include <algorithm>
#include <cmath>

namespace gb
{

__device__ __forceinline__ double eig_warp_sum(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ double eig_warp_max(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

#define EA(i, j) a[(i) * ld + (j)]

// all lanes return the same value
__device__ double warp_max_real_eigenvalue(double *a, double *vec, const int n, const int ld, const int lane)
{
  // ---- 1. balance -----------------------------------------------------------------------------------------------------
  {
    bool done = false;
    for (int sweep = 0; sweep < 16 && !done; ++sweep)
    {
      done = true;
      for (int i = 0; i < n; ++i)
      {
        double c = 0., r = 0.;
        for (int j = lane; j < n; j += 32)
          if (j != i)
          {
            c += fabs(EA(j, i));
            r += fabs(EA(i, j));
          }
        c = eig_warp_sum(c);
        r = eig_warp_sum(r);
        if (c != 0. && r != 0. && isfinite(c) && isfinite(r))
        {
          double g = r * 0.5, f = 1.;
          const double s = c + r;
          while (c < g)
          {
            f *= 2.;
            c *= 4.;
          }
          g = r * 2.;
          while (c > g)
          {
            f *= 0.5;
            c *= 0.25;
          }
          if ((c + r) / f < 0.95 * s)
          {
            done = false;
            g = 1. / f;
            for (int j = lane; j < n; j += 32)
              EA(i, j) *= g;
            __syncwarp();
            for (int j = lane; j < n; j += 32)
              EA(j, i) *= f;
            __syncwarp();
          }
        }
      }
    }
  }
  // ---- 2. Householder reduction to Hessenberg form ----------------------------------------------------------------------
  for (int m = 1; m < n - 1; ++m)
  {
    double sc = 0.;
    for (int i = m + lane; i < n; i += 32)
      sc += fabs(EA(i, m - 1));
    sc = eig_warp_sum(sc);
    if (sc == 0. || !isfinite(sc))
      continue;
    double h = 0.;
    const double isc = 1. / sc;
    for (int i = m + lane; i < n; i += 32)
    {
      const double u = EA(i, m - 1) * isc;
      vec[i] = u;
      h += u * u;
    }
    h = eig_warp_sum(h);
    __syncwarp();
    const double um = vec[m];
    const double g = um >= 0. ? -sqrt(h) : sqrt(h);
    h -= um * g;
    __syncwarp();
    if (lane == 0)
      vec[m] = um - g;
    __syncwarp();
    const double ih = 1. / h;
    for (int j = m + lane; j < n; j += 32)
    { // (I - u u^T / h) A
      double f = 0.;
      for (int i = m; i < n; ++i)
        f += vec[i] * EA(i, j);
      f *= ih;
      for (int i = m; i < n; ++i)
        EA(i, j) -= f * vec[i];
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32)
    { // A (I - u u^T / h)
      double f = 0.;
      for (int j = m; j < n; ++j)
        f += vec[j] * EA(i, j);
      f *= ih;
      for (int j = m; j < n; ++j)
        EA(i, j) -= f * vec[j];
    }
    __syncwarp();
    if (lane == 0)
      EA(m, m - 1) = sc * g;
    for (int i = m + 1 + lane; i < n; i += 32)
      EA(i, m - 1) = 0.;
    __syncwarp();
  }
  // ---- 3. double-shift QR, eigenvalues only -------------------------------------------------------------------------------
  double anorm = 0.;
  for (int i = lane; i < n; i += 32)
    for (int j = (i > 0 ? i - 1 : 0); j < n; ++j)
      anorm += fabs(EA(i, j));
  anorm = eig_warp_sum(anorm);
  double best = -INFINITY, t = 0.;
  int nn = n - 1;
  bool fail = !isfinite(anorm);
  while (nn >= 0 && !fail)
  {
    int its = 0, l;
    do
    {
      // a negligible subdiagonal element splits the matrix
      for (l = nn; l >= 1; --l)
      {
        double s = fabs(EA(l - 1, l - 1)) + fabs(EA(l, l));
        if (s == 0.)
          s = anorm;
        if (fabs(EA(l, l - 1)) + s == s)
        { // (uniform: every lane evaluates the same test on the same values)
          __syncwarp();
          if (lane == 0)
            EA(l, l - 1) = 0.;
          __syncwarp();
          break;
        }
      }
      double x = EA(nn, nn);
      if (l == nn)
      { // one real eigenvalue
        best = fmax(best, x + t);
        --nn;
      }
      else
      {
        double y = EA(nn - 1, nn - 1), w = EA(nn, nn - 1) * EA(nn - 1, nn);
        if (l == nn - 1)
        { // a 2 x 2 block: a real pair or a complex conjugate pair
          const double p = 0.5 * (y - x), q = p * p + w;
          double z = sqrt(fabs(q));
          x += t;
          if (q >= 0.)
          {
            z = p + copysign(z, p);
            double r1 = x + z, r2 = r1;
            if (z != 0.)
              r2 = x - w / z;
            best = fmax(best, fmax(r1, r2));
          }
          else
            best = fmax(best, x + p);
          nn -= 2;
        }
        else
        {
          if (its == 60)
          {
            fail = true;
            break;
          }
          if (its > 0 && its % 10 == 0)
          { // exceptional shift
            t += x;
            __syncwarp();
            for (int i = lane; i <= nn; i += 32)
              EA(i, i) -= x;
            __syncwarp();
            const double s = fabs(EA(nn, nn - 1)) + fabs(EA(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          // two consecutive small subdiagonal elements let the sweep start further down
          int m;
          double p = 0., q = 0., r = 0., z;
          for (m = nn - 2; m >= l; --m)
          {
            z = EA(m, m);
            r = x - z;
            double s = y - z;
            p = (r * s - w) / EA(m + 1, m) + EA(m, m + 1);
            q = EA(m + 1, m + 1) - z - r - s;
            r = EA(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            const double is = 1. / s;
            p *= is;
            q *= is;
            r *= is;
            if (m == l)
              break;
            const double u = fabs(EA(m, m - 1)) * (fabs(q) + fabs(r));
            const double v = fabs(p) * (fabs(EA(m - 1, m - 1)) + fabs(z) + fabs(EA(m + 1, m + 1)));
            if (u + v == v)
              break;
          }
          __syncwarp();
          for (int i = m + 2 + lane; i <= nn; i += 32)
          {
            EA(i, i - 2) = 0.;
            if (i != m + 2)
              EA(i, i - 3) = 0.;
          }
          __syncwarp();
          // the bulge chase on rows l..nn and columns m..nn
          for (int k = m; k <= nn - 1; ++k)
          {
            if (k != m)
            {
              p = EA(k, k - 1);
              q = EA(k + 1, k - 1);
              r = (k != nn - 1) ? EA(k + 2, k - 1) : 0.;
              x = fabs(p) + fabs(q) + fabs(r);
              if (x != 0.)
              {
                const double ix = 1. / x;
                p *= ix;
                q *= ix;
                r *= ix;
              }
            }
            const double s = copysign(sqrt(p * p + q * q + r * r), p);
            if (s != 0.)
            {
              __syncwarp(); // every lane has read column k-1
              if (lane == 0)
              {
                if (k == m)
                {
                  if (l != m)
                    EA(k, k - 1) = -EA(k, k - 1);
                }
                else
                  EA(k, k - 1) = -s * x;
              }
              p += s;
              const double is = 1. / s, ip = 1. / p;
              x = p * is;
              y = q * is;
              z = r * is;
              q *= ip;
              r *= ip;
              const bool three = k != nn - 1;
              for (int j = k + lane; j <= nn; j += 32)
              { // rows k..k+2
                double pp = EA(k, j) + q * EA(k + 1, j);
                if (three)
                {
                  pp += r * EA(k + 2, j);
                  EA(k + 2, j) -= pp * z;
                }
                EA(k + 1, j) -= pp * y;
                EA(k, j) -= pp * x;
              }
              __syncwarp();
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l + lane; i <= mmin; i += 32)
              { // columns k..k+2
                double pp = x * EA(i, k) + y * EA(i, k + 1);
                if (three)
                {
                  pp += z * EA(i, k + 2);
                  EA(i, k + 2) -= pp * r;
                }
                EA(i, k + 1) -= pp * q;
                EA(i, k) -= pp;
              }
              __syncwarp();
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  if (fail)
  { // no convergence (or non-finite input): Gershgorin discs of what is left, a safe upper bound
    double b = -INFINITY;
    for (int i = lane; i <= nn; i += 32)
    {
      double rad = 0.;
      for (int j = 0; j <= nn; ++j)
        if (j != i)
          rad += fabs(EA(i, j));
      b = fmax(b, EA(i, i) + t + rad);
    }
    b = eig_warp_max(b);
    best = isfinite(anorm) ? fmax(best, b) : NAN;
  }
  return best;
}

// blocks: matrix b = (f, iz) at base + f*stride_f + iz*n*n; out[b] = max Re(lambda)
__global__ void __launch_bounds__(32) k_block_max_real_eig(int nblocks, const double *__restrict__ base, long stride_f,
                                                           int per_f, int n, double *__restrict__ out)
{
  extern __shared__ __align__(16) double eig_smem[];
  const int ld = n | 1, lane = threadIdx.x;
  double *a = eig_smem, *vec = eig_smem + (size_t)n * ld;
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x)
  {
    const double *src = base + (size_t)(b / per_f) * stride_f + (size_t)(b % per_f) * n * n;
    __syncwarp();
    for (int e = lane; e < n * n; e += 32)
      a[(e / n) * ld + e % n] = src[e];
    __syncwarp();
    const double v = warp_max_real_eigenvalue(a, vec, n, ld, lane);
    if (lane == 0)
      out[b] = v;
  }
}

// out_expeig[(b)*n + q] = max(maxre[b] - diffterm, 0), flamelet_kernels.cpp:1332-1340
__global__ void k_expand_expeig(int nblocks, int n, const double *__restrict__ maxre, double diffterm,
                                double *__restrict__ out)
{
  const size_t tot = (size_t)nblocks * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x)
    out[e] = fmax(maxre[e / n] - diffterm, 0.);
}

cudaError_t launch_block_max_real_eig(int nblocks, const double *base, long stride_f, int per_f, int n, double *out,
                                      cudaStream_t s)
{
  if (nblocks <= 0)
    return cudaSuccess;
  if (n == 0)
    return cudaErrorInvalidValue;
  const size_t smem = sizeof(double) * ((size_t)n * (n | 1) + n + 2);
  if (smem > (size_t)227 * 1024)
    return cudaErrorInvalidConfiguration;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr)
  {
    cudaError_t e = cudaFuncSetAttribute(k_block_max_real_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
      return e;
    attr = smem;
  }
  k_block_max_real_eig<<<nblocks, 32, smem, s>>>(nblocks, base, stride_f, per_f, n, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_expand_expeig(int nblocks, int n, const double *maxre, double diffterm, double *out, cudaStream_t s)
{
  if (nblocks <= 0)
    return cudaSuccess;
  const size_t tot = (size_t)nblocks * n;
  const int grid = (int)std::min<size_t>((tot + 255) / 256, 148 * 8);
  k_expand_expeig<<<grid, 256, 0, s>>>(nblocks, n, maxre, diffterm, out);
  count_launch();
  return cudaGetLastError();
}

} // namespace gb
