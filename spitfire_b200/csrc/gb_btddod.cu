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
#include <cstdlib>
#include <string>

#include "../../include/griffon_b200.h"
#include "gb_kernels.cuh"
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

// ------------------------------------------------------------------------------------------------------------------
// k_btddod_factorize: one CTA per system, Q lanes per block column, rows interleaved over the lanes, the lanes' rows of
// their column RESIDENT IN REGISTERS for the whole factorisation of a block.
//
// LU (dgetf2 semantics: first row of maximum modulus, reciprocal scaling, A[i,j] -= l[i]*A[k,j]) with ONE block-wide
// barrier per elimination step: the group of column j applies every step's row interchange (two shuffles) and update
// to its own registers, reading only the step's multiplier column from shared memory; the group of column k+1 finds
// the next pivot right after its update (three warp reductions on the bit pattern of |v|), interchanges, scales and
// publishes its finished column. Interchanges of already published columns are applied in shared memory by their own
// group. The inverse of the factorised block (dgetrs on the identity, btddod_matrix_kernels.cpp:48-53) needs no barrier
// at all: every group solves for its own column in (the same) registers, the substitution value travelling through
// the group by shuffle.
// Measured per GRI block (tools: -DGB_JAC_TIMELINE, tools/dev/dev_bt.py): LU 99 k cycles (1.9 k per step: update 0.5-0.7 k,
// pivot search + publication 0.7-0.9 k, barrier 0.4 k), inverse 51 k, rest 5 k. A barrier-free dataflow variant (column
// groups spin on a "column k published" flag) was tried and is 7 % SLOWER: the step's own dependent chain, not the
// barrier, is the limit, and spinning groups take issue slots from it.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

#ifdef GB_JAC_TIMELINE
__device__ long long g_bt_timeline[8];
__device__ long long g_bt_timeline2[8];
#define BT_MARK(k)                                   \
  if (blockIdx.x == 0 && tid == 0)                   \
  {                                                  \
    const long long c__ = clock64();                 \
    g_bt_timeline[k] += c__ - bt_prev;               \
    bt_prev = c__;                                   \
  }
#else
#define BT_MARK(k)
#endif

template <int Q, int RPL> // Q lanes per column (a power of two <= 16), rows per lane: bs <= Q*RPL
__global__ void __launch_bounds__(1024, 1)
    k_btddod_factorize(int nsys, double *d_factors, int nb, int bs, double *l_values, int *pivots, double *dinv)
{
  extern __shared__ double sm[];
  const int ld = LD(bs);
  double *A = sm;            // current diagonal block / its LU factors, column-major, leading dimension ld
  double *Dn = A + ld * bs;  // next diagonal block, staged with cp.async while this one is factorised
  double *srp = Dn + ld * bs; // [bs] reciprocal of the pivot of every step (0 if the pivot is zero)
  int *spiv = (int *)(srp + bs + (bs & 1)); // [bs] pivot rows (0-based) of the current block
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const int q = tid % Q, col = tid / Q;
  const bool active = col < bs;
  const int qbase = lane & ~(Q - 1);
  const unsigned int mask = (Q == 32 ? 0xffffffffu : ((1u << Q) - 1u)) << qbase; // the lanes of a column branch together
  const size_t nb2 = (size_t)bs * bs;
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));

  // This lane's rows (q, q+Q, ...) of its column live in registers for the whole factorisation of a block: the only
  // shared-memory traffic of an elimination step is the multiplier column of the step (read) and, once per column, its
  // publication. Register index m = row / Q is warp-uniform, so the dynamic accesses are short select chains.
  double a[RPL] = {};
  auto pick = [&](int m) {
    double v = a[0];
#pragma unroll
    for (int j = 1; j < RPL; ++j)
      if (m == j)
        v = a[j];
    return v;
  };
  auto put = [&](int m, double v) {
#pragma unroll
    for (int j = 0; j < RPL; ++j)
      if (m == j)
        a[j] = v;
  };
  auto row_value = [&](int row) { // value of row `row` of this group's column, on every lane of the group
    return __shfl_sync(mask, pick(row / Q), qbase + (row & (Q - 1)));
  };
  auto swap_reg_rows = [&](int k, int p, double vk, double vp) { // rows k and p of the column (values already known)
    if (q == (k & (Q - 1)))
      put(k / Q, vp);
    if (q == (p & (Q - 1)))
      put(p / Q, vk);
  };
  // Executed by the group of column `from` once its column has received the updates of steps < from: idamax over
  // rows >= from (first row of maximum modulus), the row interchange and the reciprocal scaling of the multipliers
  // (dgetf2), then the finished column (U above and on the diagonal, multipliers below) and the pivot are published.
  auto pivot_and_publish = [&](int from) {
    double best = 0., bv;
    int bi = 0x7fffffff;
    bool has = false;
#pragma unroll
    for (int m = 0; m < RPL; ++m)
    {
      const int r = q + Q * m;
      if (r >= from && r < bs)
      {
        const double v = fabs(a[m]);
        if (!has || v > best)
        {
          best = v;
          bi = r;
          has = true;
        }
      }
    }
    // |v| >= 0 orders like its bit pattern: maximum of the high words, then of the low words among those, then the
    // lowest row among the ties -- three warp reductions instead of log2(Q) rounds of three shuffles
    const unsigned long long key = has ? (unsigned long long)__double_as_longlong(best) : 0ull;
    const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
    const unsigned int mh = __reduce_max_sync(mask, hi);
    const bool c1 = has && hi == mh;
    const unsigned int ml = __reduce_max_sync(mask, c1 ? lo : 0u);
    const bool c2 = c1 && lo == ml;
    bi = (int)__reduce_min_sync(mask, c2 ? (unsigned int)bi : 0x7fffffffu);
    bv = row_value(bi);
    const double rp = bv != 0. ? 1. / bv : 0.; // dgetf2 scales by the reciprocal; a zero pivot skips the step
    if (bi != from)
    {
      const double vk = row_value(from);
      swap_reg_rows(from, bi, vk, bv);
    }
    if (rp != 0.)
    {
#pragma unroll
      for (int m = 0; m < RPL; ++m)
      {
        const int r = q + Q * m;
        if (r > from && r < bs)
          a[m] *= rp;
      }
    }
#pragma unroll
    for (int m = 0; m < RPL; ++m)
    {
      const int r = q + Q * m;
      if (r < bs)
        A[r + col * ld] = a[m];
    }
    if (q == 0)
      spiv[from] = bi;
  };
  auto swap_rows = [&](int k, int p) { // rows k and p of this group's published column (shared memory)
    if (p != k && q == 0)
    {
      const double vk = A[k + col * ld], vp = A[p + col * ld];
      A[k + col * ld] = vp;
      A[p + col * ld] = vk;
    }
  };

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
    for (int e = tid; e < bs * bs; e += nt)
      Lv[e] = 0.; // block 0 of l_values is never referenced by the reference; define it
    __syncthreads();
#pragma unroll
    for (int m = 0; m < RPL; ++m)
    {
      const int r = q + Q * m;
      a[m] = (active && r < bs) ? A[r + col * ld] : 0.;
    }
    if (active && col == 0)
      pivot_and_publish(0);
    __syncthreads();
#ifdef GB_JAC_TIMELINE
    long long bt_prev = clock64();
#endif
    for (int i = 0; i < nb; ++i)
    {
      BT_MARK(0)
      if (i + 1 < nb)
      { // stage the next diagonal block
        const double *Dnext = D + (size_t)(i + 1) * nb2;
        for (int e = tid; e < bs * bs; e += nt)
          cp_async8(Dn + e % bs + (e / bs) * ld, Dnext + e);
        cp_async_commit();
      }
      // ---- LU with partial pivoting, one barrier per step ------------------------------------------------------------
      for (int k = 0; k < bs; ++k)
      {
#ifdef GB_JAC_TIMELINE
        long long c0 = clock64(), c1 = 0, c2 = 0, c3 = 0;
#endif
        if (active)
        {
          const int p = spiv[k];
#ifdef GB_JAC_TIMELINE
          if (p >= 0)
            c1 = clock64();
#endif
          if (col > k)
          { // step k on this column: interchange, then A[r, col] -= l[r] * A[k, col] with the published multipliers
            const double vk = row_value(k);
            double akj = vk;
            if (p != k)
            {
              akj = row_value(p);
              swap_reg_rows(k, p, vk, akj);
            }
#pragma unroll
            for (int m = 0; m < RPL; ++m)
            {
              const int r = q + Q * m;
              if (r > k && r < bs)
                a[m] -= A[r + k * ld] * akj;
            }
#ifdef GB_JAC_TIMELINE
            if (a[0] != 1.2345e300)
              c2 = clock64();
#endif
            if (col == k + 1)
            {
              pivot_and_publish(k + 1);
#ifdef GB_JAC_TIMELINE
              c3 = clock64();
#endif
            }
          }
          else if (col < k)
            swap_rows(k, p); // dlaswp on the multipliers to the left
        }
        __syncthreads();
#ifdef GB_JAC_TIMELINE
        if (blockIdx.x == 0 && active && col == k + 1 && q == 0)
        {
          if (*((volatile int *)spiv) == -12345)
            printf("x");
          g_bt_timeline2[0] += c1 - c0;
          g_bt_timeline2[1] += c2 - c1;
          g_bt_timeline2[2] += c3 - c2;
          g_bt_timeline2[3] += clock64() - c3;
        }
#endif
      }
      BT_MARK(1)
      // (column bs-1 has no multipliers; column bs-2 was finalised in step bs-1)
      for (int e = tid; e < bs * bs; e += nt)
        D[(size_t)i * nb2 + e] = A[e % bs + (e / bs) * ld];
      for (int k = tid; k < bs; k += nt)
        piv[(size_t)i * bs + k] = spiv[k] + 1;
      if (i == nb - 1 && dinv == nullptr)
        break;
      BT_MARK(2)
      // ---- column `col` of the inverse, in registers --------------------------------------------------------------------
      double(&x)[RPL] = a; // the column registers are free between the publication of the LU column and the update below
      if (active)
      {
        int pos = col; // where the 1 of e_col ends up after the row interchanges (dlaswp on the identity)
        for (int k = 0; k < bs; ++k)
        {
          const int p = spiv[k];
          if (pos == k)
            pos = p;
          else if (pos == p)
            pos = k;
        }
        double rdiag[RPL]; // reciprocals of this lane's diagonal entries of U (the triangular solve multiplies by them)
#pragma unroll
        for (int m = 0; m < RPL; ++m)
        {
          const int r = q + Q * m;
          x[m] = (r == pos) ? 1. : 0.;
          rdiag[m] = r < bs ? 1. / A[r + r * ld] : 0.;
        }
        // forward substitution with the unit lower factor
#pragma unroll
        for (int m = 0; m < RPL; ++m)
        {
          for (int qq = 0; qq < Q; ++qq)
          {
            const int k = Q * m + qq;
            if (k >= bs - 1)
              break;
            const double xk = __shfl_sync(mask, x[m], qbase + qq);
            if (xk != 0.)
            {
#pragma unroll
              for (int m2 = m; m2 < RPL; ++m2)
              {
                const int r = q + Q * m2;
                if (r > k && r < bs)
                  x[m2] -= A[r + k * ld] * xk;
              }
            }
          }
        }
        // backward substitution with the upper factor
#pragma unroll
        for (int m = RPL - 1; m >= 0; --m)
        {
          for (int qq = Q - 1; qq >= 0; --qq)
          {
            const int k = Q * m + qq;
            if (k >= bs)
              continue;
            if (q == qq)
              x[m] = x[m] * rdiag[m];
            const double xk = __shfl_sync(mask, x[m], qbase + qq);
#pragma unroll
            for (int m2 = 0; m2 <= m; ++m2)
            {
              const int r = q + Q * m2;
              if (r < k)
                x[m2] -= A[r + k * ld] * xk;
            }
          }
        }
      }
      BT_MARK(3)
      if (dinv != nullptr && active)
      { // extension: keep the explicit inverse for the matvec-only back sweep of k_btddod_solve_inv
        double *Xo = dinv + ((size_t)sys * nb + i) * nb2 + (size_t)col * bs;
#pragma unroll
        for (int m = 0; m < RPL; ++m)
        {
          const int r = q + Q * m;
          if (r < bs)
            Xo[r] = x[m];
        }
      }
      if (i == nb - 1)
        break;
      cp_async_wait_all();
      __syncthreads(); // every group is done reading the factors; the staged block has landed
      BT_MARK(4)
      // ---- L_{i+1} = diag(sub_i) * D_i^{-1} (:55-63); D_{i+1} -= L_{i+1} * diag(sup_i) (:65-75) ----------------------
      if (active)
      {
        const double *subi = sub + (size_t)i * bs;
        const double msup = -sup[(size_t)i * bs + col];
        double *Ln = Lv + (size_t)(i + 1) * nb2 + (size_t)col * bs;
#pragma unroll
        for (int m = 0; m < RPL; ++m)
        {
          const int r = q + Q * m;
          if (r < bs)
          {
            const double l = x[m] * subi[r];
            Ln[r] = l;
            a[m] = Dn[r + col * ld] + msup * l;
          }
        }
        if (col == 0)
          pivot_and_publish(0);
      }
      __syncthreads();
      BT_MARK(5)
    }
  }
}
#ifdef GB_JAC_TIMELINE
int debug_bt_timeline(long long *out)
{
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out + 8, g_bt_timeline2, sizeof(long long) * 8) != cudaSuccess)
    return -3;
  return cudaMemcpyFromSymbol(out, g_bt_timeline, sizeof(long long) * 8) == cudaSuccess ? 0 : -3;
}
#endif

// ------------------------------------------------------------------------------------------------------------------
// k_btddod_solve: forward y_i = b_i - L_i y_{i-1} (:95-103); back x_i = D_i^{-1} (y_i - sup_i o x_{i+1}) (:105-118).
// One CTA of two warps per system. The L and LU blocks stream through a double buffer in shared memory (cp.async one
// block ahead); the row interchanges of all blocks are turned into gather permutations once, in parallel over the
// blocks; warp 0 does the triangular solves with the vector in registers (rows lane and lane+32), the substitution
// value travelling by shuffle.
// ------------------------------------------------------------------------------------------------------------------
template <int RPW> // rows per lane of the solving warp: bs <= 32*RPW
__global__ void __launch_bounds__(64) k_btddod_solve(int nsys, const double *d_factors, const double *l_values,
                                                      const int *pivots, const double *rhs, int nb, int bs,
                                                      double *solution)
{
  extern __shared__ double sm[];
  double *buf0 = sm;                 // [bs*bs] block i
  double *buf1 = buf0 + bs * bs;     // [bs*bs] block i+-1 (being fetched)
  double *v = buf1 + bs * bs;        // [bs] vector exchanged between the warps
  unsigned char *perm = (unsigned char *)(v + bs + (bs & 1)); // [nb][bs] gather permutations (bs <= 256)
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const size_t nb2 = (size_t)bs * bs;
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));
  auto fetch = [&](double *dst, const double *src) {
    for (int e = tid; e < bs * bs; e += nt)
      cp_async8(dst + e, src + e);
    cp_async_commit();
  };
  for (int sys = blockIdx.x; sys < nsys; sys += gridDim.x)
  {
    const double *D = d_factors + (size_t)sys * mat_stride;
    const double *Lv = l_values + (size_t)sys * nb * nb2;
    const int *piv = pivots + (size_t)sys * nb * bs;
    const double *sup = D + (size_t)nb * nb2 + (size_t)(nb - 1) * bs;
    const double *b = rhs + (size_t)sys * nb * bs;
    double *x = solution + (size_t)sys * nb * bs; // y is stored in x during the forward sweep
    __syncthreads();
    if (nb > 1)
      fetch(buf0, Lv + nb2); // L_1
    // gather permutation of every block: v_new[k] = v_old[perm[k]] reproduces the sequential interchanges of dgetrs
    for (int i = tid; i < nb; i += nt)
    {
      unsigned char *pm = perm + (size_t)i * bs;
      for (int k = 0; k < bs; ++k)
        pm[k] = (unsigned char)k;
      for (int k = 0; k < bs; ++k)
      {
        const int p = piv[(size_t)i * bs + k] - 1;
        const unsigned char t = pm[k];
        pm[k] = pm[p];
        pm[p] = t;
      }
    }
    for (int j = tid; j < bs; j += nt)
    {
      v[j] = b[j];
      x[j] = b[j];
    }
    // ---- forward sweep -----------------------------------------------------------------------------------------------------
    for (int i = 1; i < nb; ++i)
    {
      double *cur = (i & 1) ? buf0 : buf1, *nxt = (i & 1) ? buf1 : buf0;
      cp_async_wait_all();
      __syncthreads(); // L_i has landed, v = y_{i-1} is complete
      if (i + 1 < nb)
        fetch(nxt, Lv + (size_t)(i + 1) * nb2);
      else
        fetch(nxt, D + (size_t)(nb - 1) * nb2); // first block of the back sweep
      double yj[2];
#pragma unroll
      for (int h = 0; h < 2; ++h)
      {
        const int j = tid + h * 64;
        if (j < bs)
        {
          double acc = b[(size_t)i * bs + j];
          for (int c = 0; c < bs; ++c)
            acc = acc + cur[(size_t)c * bs + j] * (-1. * v[c]); // matrix_vector_multiply order, blas_lapack_kernels.h:60-77
          yj[h] = acc;
        }
      }
      __syncthreads(); // everyone is done reading v
#pragma unroll
      for (int h = 0; h < 2; ++h)
      {
        const int j = tid + h * 64;
        if (j < bs)
        {
          v[j] = yj[h];
          x[(size_t)i * bs + j] = yj[h];
        }
      }
    }
    if (nb == 1)
      fetch(buf0, D);
    // ---- back sweep ----------------------------------------------------------------------------------------------------------
    // after the forward sweep the LU factors of block nb-1 are in buffer (nb & 1 ? buf1 : buf0) for nb > 1, buf0 for nb == 1
    for (int i = nb - 1, step = 0; i >= 0; --i, ++step)
    {
      double *cur, *nxt;
      if (nb == 1)
        cur = buf0, nxt = buf1;
      else
      {
        const bool first_in_buf1 = ((nb - 1) & 1) != 0; // block nb-1 was fetched as `nxt` of forward step nb-1
        const bool in1 = first_in_buf1 ? ((step & 1) == 0) : ((step & 1) != 0);
        cur = in1 ? buf1 : buf0;
        nxt = in1 ? buf0 : buf1;
      }
      cp_async_wait_all();
      __syncthreads(); // factors of block i have landed; v = x_{i+1} (or y_{nb-1}) is complete
      if (i > 0)
        fetch(nxt, D + (size_t)(i - 1) * nb2);
      if (warp == 0)
      {
        const unsigned char *pm = perm + (size_t)i * bs;
        double w[RPW];
#pragma unroll
        for (int h = 0; h < RPW; ++h)
        {
          const int r = lane + 32 * h;
          w[h] = 0.;
          if (r < bs)
          { // right-hand side y_i - sup_i o x_{i+1}, row interchanges applied as a gather
            const int s = pm[r];
            const double ys = x[(size_t)i * bs + s];
            w[h] = (i == nb - 1) ? ys : ys - sup[(size_t)i * bs + s] * v[s];
          }
        }
        double rdiag[RPW]; // reciprocals of this lane's diagonal entries of U
#pragma unroll
        for (int h = 0; h < RPW; ++h)
        {
          const int r = lane + 32 * h;
          rdiag[h] = r < bs ? 1. / cur[r + (size_t)r * bs] : 0.;
        }
        // unit-lower forward substitution
#pragma unroll
        for (int h = 0; h < RPW; ++h)
        {
#pragma unroll 4
          for (int l = 0; l < 32; ++l)
          {
            const int k = 32 * h + l;
            if (k < bs - 1)
            {
              const double wk = __shfl_sync(0xffffffffu, w[h], l);
#pragma unroll
              for (int h2 = h; h2 < RPW; ++h2)
              {
                const int r = lane + 32 * h2;
                if (r > k && r < bs)
                  w[h2] -= cur[r + (size_t)k * bs] * wk;
              }
            }
          }
        }
        // upper backward substitution
#pragma unroll
        for (int h = RPW - 1; h >= 0; --h)
        {
#pragma unroll 4
          for (int l = 31; l >= 0; --l)
          {
            const int k = 32 * h + l;
            if (k < bs)
            {
              if (lane == l)
                w[h] = w[h] * rdiag[h];
              const double wk = __shfl_sync(0xffffffffu, w[h], l);
#pragma unroll
              for (int h2 = 0; h2 <= h; ++h2)
              {
                const int r = lane + 32 * h2;
                if (r < k)
                  w[h2] -= cur[r + (size_t)k * bs] * wk;
              }
            }
          }
        }
        __syncwarp(); // every lane has gathered its right-hand side from v (the shuffles above order this already)
#pragma unroll
        for (int h = 0; h < RPW; ++h)
        {
          const int r = lane + 32 * h;
          if (r < bs)
          {
            v[r] = w[h];
            x[(size_t)i * bs + r] = w[h];
          }
        }
      }
    }
    cp_async_wait_all();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_btddod_solve_inv (extension, not in the reference API): the same block-Thomas solve with the explicit inverses
// D_i^{-1} that the factorisation forms anyway for L_{i+1}: the back sweep x_i = D_i^{-1} (y_i - sup_i o x_{i+1})
// becomes a matrix-vector product, fully parallel over the rows, instead of two dependent triangular solves. Used
// inside the Newton loops of the steady solvers and of the batched ESDIRK integrator, where the linear solve only
// proposes the update of an iteration that converges on the true residual.
//
// One CTA of eight warps per system. The 2*nb-1 blocks (L_1..L_{nb-1}, then Dinv_{nb-1}..Dinv_0) stream through a
// four-deep ring of shared-memory buffers filled with 16-byte cp.async (the blocks are only 8-byte aligned when the
// block size is odd: the first / last element then travel separately). Every row's dot product is split over four
// lanes (fma, partial sums combined by shuffle), so the dependent chain per block step is bs/4 long; the vector
// ping-pongs between two shared arrays, which leaves one barrier per step. The right-hand side is staged in shared
// memory once and overwritten with y_i, so the back sweep never waits for a global read of its own earlier writes.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SI_STAGES = 4;
constexpr int SI_THREADS = 256;

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gmem_src)
{
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "MBAR_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra MBAR_DONE;\n"
               "bra MBAR_WAIT;\n"
               "MBAR_DONE:\n"
               "}\n" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned addresses and size)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared-memory load that keeps its place in the instruction stream (ptxas otherwise sinks each pair of loads to just
// before the fma that consumes it, which serialises the shared-memory latency of the whole dot product)
__device__ __forceinline__ double lds_f64(const double *p)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ double lds32_f64(unsigned int a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds32_v2f64(unsigned int a)
{
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(const double *p)
{
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)));
  return v;
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group()
{
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(SI_THREADS) k_btddod_solve_inv(int nsys, const double *d_factors, const double *l_values,
                                                                 const double *dinv, const double *rhs, int nb, int bs,
                                                                 double *solution, int staged, const int *rows, int pairs)
{
  // pairs != 0: launched as clusters of two CTAs per system. Factors from the TWISTED elimination (gb_btinv.cu: the
  // tag in block 0 of l_values) are then applied from both ends at once -- CTA 0 sweeps the blocks 0..m downwards and
  // back, CTA 1 the blocks nb-1..m+1 upwards and back, the two meeting at block m through two cluster barriers -- which
  // halves the dependent chain of a solve. Plain factors: CTA 0 alone runs the one-sided sweep (m = nb-1).
  const unsigned int crank = pairs ? cluster_ctarank() : 0u;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb2 = bs * bs;
  const int bufsz = (nb2 + 3) & ~1; // room for the one-element shift of an odd-aligned block, multiple of 16 bytes
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm); // [SI_STAGES] "block landed" barriers
  double *ring = sm + 8;
  double *v0 = ring + (size_t)SI_STAGES * bufsz; // [2][bsp] ping-pong vector
  const int bsp = (bs + 63) & ~63; // padded with zeros to whole groups of 64 columns (the dot product does not mask)
  double *yv = v0 + 2 * bsp;                     // [nb*bs] right-hand side, overwritten with y_i (if `staged`: it fits)
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));
  const int sub = lane & 7, part = lane >> 3;
  if (tid == 0)
  {
    for (int k = 0; k < SI_STAGES; ++k)
      mbar_init(bars + k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  int gs0 = 0; // blocks streamed by this CTA before the current system (ring slot and barrier phase follow from it)
  int pair_iter = 0; // systems this pair of CTAs has looked at (selects one of two decision words)
  // byte offsets, inside a block, of the sixteen matrix entries this thread multiplies in the fast path (bs <= 64):
  // the same for every step, kept in registers so that a step's loads are plain 32-bit shared addresses
  unsigned int mofs[16];
  {
    const int cp0 = 8 * (part & 1) + 4 * (part >> 1);
    const int r0 = warp * 8 + sub;
#pragma unroll
    for (int u = 0; u < 16; ++u)
    {
      const int c = cp0 + 16 * (u >> 2) + (u & 3);
      mofs[u] = (unsigned int)(((c < bs ? c : 0) * bs + (r0 < bs ? r0 : 0)) * 8);
    }
  }

  for (int sys = pairs ? (int)clusterid_x() : blockIdx.x; sys < nsys; sys += pairs ? (int)nclusterid_x() : gridDim.x)
  {
    const int fsys = rows ? rows[sys] : sys; // where this system's factors live (right-hand sides are compact)
    const double *D = d_factors + (size_t)fsys * mat_stride;
    const double *Lv = l_values + (size_t)fsys * nb * nb2;
    const double *Di = dinv + (size_t)fsys * nb * nb2;
    const double *b = rhs + (size_t)sys * nb * bs;
    double *x = solution + (size_t)sys * nb * bs;
    // meeting block m of the sweep(s): nb-1 for plain factors. CTA 0 reads the tag and the pair agrees on ITS reading
    // through distributed shared memory (a caller that refreshes factors on another stream while a solve of the whole
    // batch is in flight -- the asynchronous integrator does, for members that are not inside a step -- must not be able
    // to split the pair over a half-written tag: one CTA would wait for the other at a cluster barrier for ever)
    int mdec = -1;
    if (pairs)
    {
      int *sdec = reinterpret_cast<int *>(sm + 4) + (pair_iter & 1);
      if (crank == 0 && tid == 0)
      {
        const int mt = (nb >= 4 && bs >= 2 && __ldg(Lv + 1) == BT_TWIST_MAGIC) ? (int)__ldg(Lv) : -1;
        *sdec = (mt >= 1 && mt <= nb - 2) ? mt : -1;
      }
      cluster_sync_all();
      unsigned int remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(sdec)), "r"(0));
      asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(mdec) : "r"(remote) : "memory");
      ++pair_iter;
    }
    const bool twisted = mdec >= 0;
    const int m = twisted ? mdec : nb - 1;
    if (crank != 0 && !twisted)
      continue;
    const bool top = crank == 0;
    // this CTA: nf forward steps (block i, multiplier slot i for the top sweep / i+1 for the bottom one), then nbk
    // back steps
    const int nf = top ? m : nb - 1 - m, nbk = top ? m + 1 : nb - 1 - m;
    // off-diagonal that couples a solved block to the next one of the back sweep: sup_{i-1} going up, sub_i going down
    const double *offd = D + (size_t)nb * nb2 + (top ? (size_t)(nb - 1) * bs : (size_t)0);
    const int nstream = nf + nbk;
    auto blk_of = [&](int s) { return s < nf ? (top ? s + 1 : nb - 2 - s) : (top ? m - (s - nf) : m + 1 + (s - nf)); };
    auto src_of = [&](int s) {
      const int i = blk_of(s);
      return s < nf ? Lv + (size_t)(top ? i : i + 1) * nb2 : Di + (size_t)i * nb2;
    };
    // Block s lands at ring[(gs0+s) % STAGES] + (1 if its source address is an odd multiple of 8 bytes): one TMA bulk
    // copy of the 16-byte aligned span that lies inside the block's neighbourhood (a single SM cannot keep enough
    // LDGSTS requests in flight to stream 22 KB per step), plus the odd last element, if any, by an 8-byte cp.async.
    // parity (in units of 8 bytes) of the address of block s: decides the one-element shift of its ring slot
    const unsigned int parL = (unsigned int)(reinterpret_cast<size_t>(Lv) >> 3) & 1u;
    const unsigned int parD = (unsigned int)(reinterpret_cast<size_t>(Di) >> 3) & 1u;
    auto par_of = [&](int s) {
      const int i = blk_of(s);
      return s < nf ? (parL + (unsigned int)(top ? i : i + 1) * (unsigned int)nb2) & 1u
                    : (parD + (unsigned int)i * (unsigned int)nb2) & 1u;
    };
    auto fetch = [&](int s) {
      if (s < nstream && warp == SI_THREADS / 32 - 1)
      {
        const double *src = src_of(s);
        const int g = gs0 + s;
        double *dst = ring + (size_t)(g % SI_STAGES) * bufsz;
        const int off = (int)((reinterpret_cast<size_t>(src) >> 3) & 1);
        const int span = nb2 + off, n16 = span >> 1; // elements from src - off
        // (issued from the last warp: it owns rows 56..63, so for block sizes up to 56 it has no dot product to start
        // and the ~350 cycles of address arithmetic and issue stay off the critical path of the sweep)
        if (tid == SI_THREADS - 32)
        {
          mbar_expect_tx(bars + g % SI_STAGES, (unsigned int)n16 * 16u);
          bulk_g2s(dst, src - off, (unsigned int)n16 * 16u, bars + g % SI_STAGES);
        }
        if ((span & 1) && tid == SI_THREADS - 31)
          cp_async8(dst + 2 * n16, src - off + 2 * n16);
      }
      cp_async_commit(); // (an empty group keeps the group count in step)
    };
    auto buf_of = [&](int s) { return ring + (unsigned int)((gs0 + s) % SI_STAGES) * (unsigned int)bufsz + par_of(s); };
    __syncthreads(); // the previous system is done with the shared arrays
    if (staged)
    { // (the blocks this CTA's sweep touches)
      const int e0 = top ? 0 : (m + 1) * bs, e1 = top ? (m + 1) * bs : nb * bs;
      for (int e = e0 + tid; e < e1; e += SI_THREADS)
        yv[e] = b[e];
    }
    for (int s = 0; s < SI_STAGES - 1; ++s)
      fetch(s);
    for (int j = tid; j < 2 * bsp; j += SI_THREADS)
      if (j % bsp >= bs)
        v0[j] = 0.; // the padding of both vectors stays zero for the whole sweep
    {
      const size_t first = top ? (size_t)0 : (size_t)(nb - 1) * bs;
      for (int j = tid; j < bs; j += SI_THREADS)
      {
        v0[j] = b[first + j];
        x[first + j] = b[first + j];
      }
    }
    for (int s = 0; s < nstream; ++s)
    {
      const bool fwd = s < nf;
      const int i = blk_of(s);
      // back steps: is there a next block, which one, and the off-diagonal entry that couples it
      const bool has_next = top ? i > 0 : i < nb - 1;
      const int inext = top ? i - 1 : i + 1, ioff = top ? i - 1 : i;
      // bottom sweep, last forward step: only t = U_m z_{m+1} is wanted (handed to the top sweep through x_m's slot)
      const bool hand_t = twisted && !top && fwd && s == nf - 1;
      if (twisted && !top && s == nf)
      { // the top sweep has published x_m: right-hand side of the first downward step
        __syncthreads();
        cluster_sync_all();
        double *vfirst = v0 + (s & 1) * bsp;
        for (int j = tid; j < bs; j += SI_THREADS)
          vfirst[j] = (staged ? yv[(size_t)(m + 1) * bs + j] : b[(size_t)(m + 1) * bs + j]) -
                      __ldg(offd + (size_t)m * bs + j) * __ldcg(x + (size_t)m * bs + j);
      }
      // operand of the epilogue that does not depend on this sweep: issued before the wait (first 64 rows)
      double sup0 = 0.;
      if (!fwd && has_next && part == 0 && warp * 8 + sub < bs)
        sup0 = __ldg(offd + (size_t)ioff * bs + warp * 8 + sub);
#ifdef GB_JAC_TIMELINE
      long long c0 = clock64();
#endif
      cp_async_wait_group<SI_STAGES - 2>();
      mbar_wait(bars + (gs0 + s) % SI_STAGES, (unsigned int)(((gs0 + s) / SI_STAGES) & 1));
#ifdef GB_JAC_TIMELINE
      long long c1 = clock64();
#endif
      __syncthreads(); // block s has landed for everyone; the vector of this step is complete; ring slot s-1 is free
#ifdef GB_JAC_TIMELINE
      if (*((volatile double *)v0) == 1.2345e300) // (a read that needs the barrier's release: the clock below is taken after it)
        printf("x");
      long long c2 = clock64();
#endif
      fetch(s + SI_STAGES - 1);
#ifdef GB_JAC_TIMELINE
      long long c3 = clock64();
#endif
      const double *cur = buf_of(s);
      const double *vin = v0 + (s & 1) * bsp;
      double *vout = v0 + ((s + 1) & 1) * bsp;
      for (int row0 = 0; row0 < bs; row0 += 64)
      {
        const int row = row0 + warp * 8 + sub;
        const bool live = row < bs;
        const int rr = live ? row : 0;
        // operands of the epilogue, fetched ahead of the dot product
        double e0 = 0., e1 = sup0;
        if (part == 0 && live)
        {
          if (fwd)
            e0 = hand_t ? 0. : (staged ? yv[(size_t)i * bs + row] : b[(size_t)i * bs + row]);
          else if (has_next)
          {
            e0 = staged ? yv[(size_t)inext * bs + row] : x[(size_t)inext * bs + row];
            if (row0 > 0)
              e1 = __ldg(offd + (size_t)ioff * bs + row);
          }
        }
        // part p takes the columns 16 j + 8 (p & 1) + 4 (p >> 1) + t, t < 4: the two parts of a half-warp are eight
        // columns apart, which for an odd block size puts their eight rows in disjoint banks. All loads of a pair of
        // j are issued before the first fma (two accumulators).
        const int cpart = 8 * (part & 1) + 4 * (part >> 1);
        const double *cb = cur + rr;
        double acc;
        if (bs <= 64)
        {
          // All sixteen matrix entries and the sixteen vector entries (four 16-byte loads per pair of groups) are
          // requested before the first fma; four accumulators of depth four. Columns past the end re-read column 0
          // against the zero padding of the vector.
          double mm[16];
          double2 wv[8];
          const unsigned int cur32 = smem_u32(cur), vin32 = smem_u32(vin) + (unsigned int)cpart * 8u;
#pragma unroll
          for (int u = 0; u < 16; ++u)
            mm[u] = lds32_f64(cur32 + mofs[u]);
#pragma unroll
          for (int q = 0; q < 4; ++q)
          {
            wv[2 * q] = lds32_v2f64(vin32 + 128u * q);
            wv[2 * q + 1] = lds32_v2f64(vin32 + 128u * q + 16u);
          }
          // (the accumulators depend on the LAST loads, so that the in-order issue cannot stall on the first fma
          // before every load is in flight)
          double a0 = fma(mm[15], 0., 0.), a1 = fma(wv[7].y, 0., 0.), a2 = 0., a3 = 0.;
#pragma unroll
          for (int q = 0; q < 4; ++q)
          {
            a0 = fma(mm[4 * q], wv[2 * q].x, a0);
            a1 = fma(mm[4 * q + 1], wv[2 * q].y, a1);
            a2 = fma(mm[4 * q + 2], wv[2 * q + 1].x, a2);
            a3 = fma(mm[4 * q + 3], wv[2 * q + 1].y, a3);
          }
          acc = (a0 + a1) + (a2 + a3);
        }
        else
        {
          double a0 = 0., a1 = 0.;
          for (int j = 0; j < bs; j += 32)
          {
            double mm[8], ww[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
              const int c = cpart + j + 16 * (u >> 2) + (u & 3);
              const int cc = c < bs ? c : 0; // (a column past the end re-reads column 0 with weight zero)
              mm[u] = lds_f64(cb + (size_t)cc * bs);
              ww[u] = lds_f64(vin + cc);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (j + 16 * (u >> 2) + (u & 3) + cpart >= bs)
                ww[u] = 0.;
            a0 = fma(mm[7], 0., a0);
            a1 = fma(ww[7], 0., a1);
#pragma unroll
            for (int u = 0; u < 8; u += 2)
            {
              a0 = fma(mm[u], ww[u], a0);
              a1 = fma(mm[u + 1], ww[u + 1], a1);
            }
          }
          acc = a0 + a1;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
#ifdef GB_JAC_TIMELINE
        if (blockIdx.x == 0 && tid == 0 && acc != 1.2345e300)
          g_bt_timeline2[3] += clock64() - c3; // (overwrites the LU 'barrier wait' slot while the solve runs)
#endif
        if (part == 0 && live)
        {
          if (fwd)
          { // y_i = b_i - L_i y_{i-1}
            const double y = e0 - acc;
            if (staged && !hand_t)
              yv[(size_t)i * bs + row] = y;
            else
              x[(size_t)i * bs + row] = y;
            vout[row] = y; // feeds the next forward step, or is the right-hand side of the first back step
          }
          else
          { // x_i = Dinv_i v; the next right-hand side is y_{i-1} - sup_{i-1} o x_i
            x[(size_t)i * bs + row] = acc;
            if (has_next)
              vout[row] = e0 - e1 * acc;
          }
        }
      }
      if (twisted)
      {
        if (top && s == nf - 1)
        { // y_m is complete; the bottom sweep hands over -U_m z_{m+1}: the first back step solves for x_m
          cluster_sync_all();
          for (int j = tid; j < bs; j += SI_THREADS)
            vout[j] += __ldcg(x + (size_t)m * bs + j);
        }
        else if (top ? s == nf : s == nf - 1)
        { // publish x_m (top) / -t (bottom) to the other CTA
          __threadfence();
          cluster_sync_all();
        }
      }
#ifdef GB_JAC_TIMELINE
      if (blockIdx.x == 0 && tid == 0)
      {
        g_bt_timeline2[4] += c1 - c0;
        g_bt_timeline2[5] += c2 - c1;
        g_bt_timeline2[6] += c3 - c2;
        g_bt_timeline2[7] += clock64() - c3;
      }
#endif
    }
    cp_async_wait_group<0>();
    gs0 += nstream;
  }
  // (a CTA's shared memory must outlive the other CTA's reads of it: leave together)
  if (pairs)
    cluster_sync_all();
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
  const size_t need = sizeof(double) * (2 * (size_t)LD(bs) * bs + bs + 2) + sizeof(int) * bs + 64;
  if (need > 227 * 1024 || bs > 120)
  {
    set_error("block size too large for the in-shared-memory block LU");
    return GB_ERR_UNSUPPORTED;
  }
  return GB_OK;
}
} // namespace
bool gb::bt_twist_ok(int nb, int bs)
{
  static int enabled = -1;
  if (enabled < 0)
  {
    const char *e = getenv("GB_BT_TWIST");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || nb < 4 || bs < 2)
    return false;
  // (k_btddod_solve_inv with the right-hand side staged in shared memory, see gb_btddod_full_solve_inv_batch)
  const size_t base = sizeof(double) * ((size_t)SI_STAGES * (((size_t)bs * bs + 3) & ~(size_t)1) + 2 * (((size_t)bs + 63) & ~(size_t)63)) + 64 + 16;
  return base + sizeof(double) * (size_t)nb * bs <= (size_t)227 * 1024;
}
namespace
{
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
  static int factorize_impl(int n, double *d_factors, int nb, int bs, double *l_values, int *pivots, double *dinv,
                            void *stream);

  int gb_btddod_full_factorize_batch(int n, double *d_factors, int nb, int bs, double *l_values, int *pivots,
                                     void *stream)
  {
    return factorize_impl(n, d_factors, nb, bs, l_values, pivots, nullptr, stream);
  }

  int gb_btddod_full_factorize_inv_batch(int n, double *d_factors, int nb, int bs, double *l_values, int *pivots,
                                         double *out_dinv, void *stream)
  {
    if (out_dinv == nullptr)
    {
      set_error("gb_btddod_full_factorize_inv_batch needs the out_dinv array");
      return GB_ERR_ARG;
    }
    return factorize_impl(n, d_factors, nb, bs, l_values, pivots, out_dinv, stream);
  }

  int gb_btddod_full_solve_inv_batch(int n, const double *d_factors, const double *l_values, const double *dinv,
                                     const double *rhs, int nb, int bs, double *solution, const int *system_rows,
                                     void *stream)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t base = sizeof(double) * ((size_t)SI_STAGES * (((size_t)bs * bs + 3) & ~(size_t)1) + 2 * (((size_t)bs + 63) & ~(size_t)63)) + 64 + 16;
    const size_t with_rhs = base + sizeof(double) * (size_t)nb * bs;
    if (base > (size_t)227 * 1024)
    {
      set_error("block size too large for k_btddod_solve_inv (a ring of four blocks lives in shared memory)");
      return GB_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<size_t>(l_values) | reinterpret_cast<size_t>(dinv)) & 15)
    { // the bulk copies start at the 16-byte boundary at or just below a block
      set_error("gb_btddod_full_solve_inv_batch: l_values and dinv must be 16-byte aligned");
      return GB_ERR_ARG;
    }
    const int staged = with_rhs <= (size_t)227 * 1024 ? 1 : 0; // the right-hand side too, when it fits
    const size_t smem = staged ? with_rhs : base;
    BCK(cudaFuncSetAttribute(k_btddod_solve_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (bt_twist_ok(nb, bs))
    { // clusters of two CTAs per system: factors of the twisted elimination are applied from both ends at once
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (unsigned int)n), cfg.blockDim = dim3(SI_THREADS);
      cfg.dynamicSmemBytes = smem, cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
      cfg.attrs = at, cfg.numAttrs = 1;
      BCK(cudaLaunchKernelEx(&cfg, k_btddod_solve_inv, n, d_factors, l_values, dinv, rhs, nb, bs, solution, staged,
                             system_rows, 1));
    }
    else
      k_btddod_solve_inv<<<n, SI_THREADS, smem, (cudaStream_t)stream>>>(n, d_factors, l_values, dinv, rhs, nb, bs, solution,
                                                                        staged, system_rows, 0);
    ++g_btddod_launches;
    BCK(cudaGetLastError());
    return GB_OK;
  }

  static int factorize_impl(int n, double *d_factors, int nb, int bs, double *l_values, int *pivots, double *dinv,
                            void *stream)
  {
    int rc = bt_check(n, nb, bs);
    if (rc != GB_OK || n == 0)
      return rc;
    const size_t smem = sizeof(double) * (2 * (size_t)LD(bs) * bs + bs + 2) + sizeof(int) * bs + 64;
#define GB_BT_LAUNCH(Q, RPL)                                                                                          \
  do                                                                                                                  \
  {                                                                                                                   \
    BCK(cudaFuncSetAttribute(k_btddod_factorize<Q, RPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    k_btddod_factorize<Q, RPL><<<n, ((Q * bs + 31) / 32) * 32, smem, (cudaStream_t)stream>>>(n, d_factors, nb, bs,    \
                                                                                             l_values, pivots, dinv); \
  } while (0)
    if (bs <= 16)
      GB_BT_LAUNCH(16, 1);
    else if (bs <= 32)
      GB_BT_LAUNCH(16, 2);
    else if (bs <= 64)
      GB_BT_LAUNCH(16, 4);
    else
      GB_BT_LAUNCH(8, 15);
#undef GB_BT_LAUNCH
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
    const size_t smem = sizeof(double) * (2 * (size_t)bs * bs + bs + 2) + (size_t)nb * bs + 64;
    if (smem > 227 * 1024)
    {
      set_error("system too large for the shared-memory resident block-Thomas solve");
      return GB_ERR_UNSUPPORTED;
    }
    if (bs <= 64)
    {
      BCK(cudaFuncSetAttribute(k_btddod_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_btddod_solve<2><<<n, 64, smem, (cudaStream_t)stream>>>(n, d_factors, l_values, pivots, rhs, nb, bs, solution);
    }
    else
    {
      BCK(cudaFuncSetAttribute(k_btddod_solve<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_btddod_solve<4><<<n, 64, smem, (cudaStream_t)stream>>>(n, d_factors, l_values, pivots, rhs, nb, bs, solution);
    }
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
