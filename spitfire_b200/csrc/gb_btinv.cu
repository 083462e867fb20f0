// gb_btinv.cu -- k_btddod_invert: block-Thomas elimination of BTDDOD systems by explicit inverses, FP64, sm_100a.
//
// Extension of the block-Thomas path (btddod_matrix_kernels.cpp:19-80) for the Newton / PsiTC / ESDIRK loops that run
// on the device: those loops only ever apply the factorisation through gb_btddod_full_solve_inv_batch, whose forward
// sweep needs L_i = diag(sub_{i-1}) D'_{i-1}^{-1} and whose back sweep needs D'_i^{-1}, with
//     D'_0 = D_0,   D'_i = D_i - L_i diag(sup_{i-1})                 (btddod_matrix_kernels.cpp:48-75)
// so the LU factors and pivot vectors of the reference's layout are never read. This kernel therefore skips dgetrf +
// dgetrs-on-the-identity (two dependent sweeps of ~3 bs steps per block) and inverts every D'_i in place with
// Gauss-Jordan elimination and partial pivoting: bs steps per block, ONE block barrier per step.
//
// The block recurrence is sequential, so a system is a latency chain of nb * bs elimination steps; the design goal is
// the shortest possible step:
//   * one CTA of 8 warps per system; the block lives in REGISTERS, thread (warp w, lane l) owns rows l + 32a and
//     columns w + 8b -- 14 doubles for a 53 x 53 block;
//   * implicit pivoting: rows are never exchanged. The pivot of step k is the largest entry of column k among the rows
//     not used yet; the warp that owns column k finds it with three redux.sync on the bit pattern of |a| and
//     publishes the pivot row index, the reciprocal pivot and the column (one 8-byte store per lane) -- the only
//     shared-memory traffic of a step; the other warps pick the pivot row out of their own registers by shuffle;
//   * the row / column permutation is undone once per block when the inverse goes to shared memory in its true
//     layout, from where it is stored (coalesced) and folded into the next block's D'.
// The original matrix is NOT overwritten (the reference's factorisation is in place; here the Jacobian survives).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "../../include/griffon_b200.h"
#include "gb_kernels.cuh"
#include "gb_mech.h"

namespace gb
{
extern std::atomic<long> g_btddod_launches;

namespace
{

#ifdef GB_JAC_TIMELINE
__device__ long long g_inv_timeline[8];
#endif

template <int N, int I = 0, class F>
__device__ __forceinline__ void static_for(F &&f)
{
  if constexpr (I < N)
  {
    f(std::integral_constant<int, I>{});
    static_for<N, I + 1>(f);
  }
}

// reciprocal by MUFU.RCP64H and one cubic correction: within an ulp or two of 1/x for normal x,
// without the range checks and the slow path of the IEEE division (the pivot's reciprocal sits on the critical path of
// every elimination step). 0 -> inf and non-finite x behave as in a plain division for the purposes of this kernel
// (the block is singular: the result is non-finite either way).
__device__ __forceinline__ double fast_rcp(double x)
{
  // r0 = 1/x (1 + O(2^-23)); with e = 1 - x r0: 1/x = r0 (1 + e + e^2 + ...) -- the cubic step r0 (1 + e + e^2) leaves
  // 2^-69 in three dependent FP64 operations (the FP64 pipe's dependent latency is what a pivot step pays for), and
  // one more residual correction brings the result within an ulp
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.);
  const double p = fma(e, e, e);
  return fma(r, p, r);
}

// IW warps per CTA (a power of two), R rows per lane (bs <= 32 R), C column slots per warp (bs <= IW C)
template <int R, int C, int IW>
__global__ void __launch_bounds__(IW * 32) k_btddod_invert(int nsys, const double *__restrict__ mats, int nb, int bs,
                                                        double *__restrict__ l_values, double *__restrict__ dinv, int twist)
{
  // twist != 0 (launched as clusters of two CTAs per system): TWISTED elimination. CTA 0 eliminates downwards from
  // block 0, CTA 1 upwards from block nb-1 with the roles of the two off-diagonals exchanged,
  //     D''_{nb-1} = D_{nb-1},   U_i = diag(sup_i) D''_{i+1}^{-1},   D''_i = D_i - U_i diag(sub_i),
  // and the two meet at block m = (nb-1)/2:   D*_m = D_m - L_m diag(sup_{m-1}) - U_m diag(sub_m).
  // Slot i of l_values holds L_i for i <= m and U_{i-1} for i > m; dinv holds the inverses of D'_i, D*_m, D''_i; block
  // 0 of l_values carries the tag {m, BT_TWIST_MAGIC} by which k_btddod_solve_inv recognises the format. The dependent
  // chain of a system is halved.
  const unsigned int crank = twist ? cluster_ctarank() : 0u;
  const bool top = crank == 0;
  const int m = twist ? (nb - 1) / 2 : nb - 1;
  const int nblk = top ? m + 1 : nb - 1 - m;
  constexpr int INT_ = IW * 32;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb2 = bs * bs;
  double *sInv = sm;                        // [bs*bs] inverse of the previous block, true column-major layout
  double *sf = sInv + nb2 + (nb2 & 1);      // [2][64 R... ] published pivot column, ping-pong
  const int fstride = 32 * R;
  double *srinv = sf + 2 * fstride;         // [2] reciprocal pivot
  int *sp = reinterpret_cast<int *>(srinv + 2);  // [2] pivot row of the step
  int *sperm = sp + 2;                      // [bs] pivot row of step k
  int *sinvp = sperm + 32 * R;              // [bs] step at which row r was the pivot
  const size_t mat_stride = (size_t)bs * ((size_t)nb * bs + 2 * (nb - 1));

  for (int sys = twist ? (int)clusterid_x() : blockIdx.x; sys < nsys; sys += twist ? (int)nclusterid_x() : gridDim.x)
  {
    const double *M = mats + (size_t)sys * mat_stride;
    const double *subd = M + (size_t)nb * nb2, *supd = subd + (size_t)(nb - 1) * bs;
    // row / column scaling of the previous inverse for the j-th block of this CTA's sweep (block i = blk_of(j)):
    // downwards sub_{i-1} by row and sup_{i-1} by column, upwards sup_i by row and sub_i by column
    const double *rowd = top ? subd : supd, *cold = top ? supd : subd;
    auto blk_of = [&](int j) { return top ? j : nb - 1 - j; };
    double *Lv = l_values + (size_t)sys * nb * nb2;
    double *Di = dinv + (size_t)sys * nb * nb2;
    double S[R][C], Dn[R][C], subn[R], supn[C];
    // block i into registers, with the off-diagonal entries that turn it into D'_i (sub_{i-1} by row, sup_{i-1} by column)
    auto load_block = [&](int j) {
      const int i = blk_of(j), io = top ? i - 1 : i;
      const double *D = M + (size_t)i * nb2;
#pragma unroll
      for (int b = 0; b < C; ++b)
      {
        const int col = warp + IW * b;
#pragma unroll
        for (int a = 0; a < R; ++a)
        {
          const int row = lane + 32 * a;
          Dn[a][b] = (row < bs && col < bs) ? __ldg(D + (size_t)col * bs + row) : 0.;
        }
        supn[b] = (j > 0 && col < bs) ? __ldg(cold + (size_t)io * bs + col) : 0.;
      }
#pragma unroll
      for (int a = 0; a < R; ++a)
        subn[a] = (j > 0 && lane + 32 * a < bs) ? __ldg(rowd + (size_t)io * bs + lane + 32 * a) : 0.;
    };
    load_block(0);
    __syncthreads(); // the previous system is done with the shared arrays
    for (int j = 0; j < nblk; ++j)
    {
      const int i = blk_of(j);
      const size_t slot = top ? (size_t)i : (size_t)i + 1; // where this block's multiplier goes
      const bool meet = twist && top && j == m;           // the block where the two sweeps meet
      if (meet)
        cluster_sync_all(); // U_m (slot m+1) has been written by the other CTA
      // ---- D'_i: the first block as it is; then D_i - (sub o Dinv_{i-1}) o sup, L_i on the way (:48-75) --------------
#pragma unroll
      for (int b = 0; b < C; ++b)
      {
        const int col = warp + IW * b;
#pragma unroll
        for (int a = 0; a < R; ++a)
        {
          const int row = lane + 32 * a;
          double v = Dn[a][b];
          if (j > 0 && row < bs && col < bs)
          {
            const double l = subn[a] * sInv[(size_t)col * bs + row];
            Lv[slot * nb2 + (size_t)col * bs + row] = l;
            v = v - l * supn[b];
          }
          if (top && j == 0 && row < bs && col < bs) // block 0 of l_values is never read as a multiplier: zero, or the tag
            Lv[(size_t)col * bs + row] = (twist && col == 0 && row < 2) ? (row == 0 ? (double)m : BT_TWIST_MAGIC) : 0.;
          if (meet && row < bs && col < bs)
            v = v - __ldcg(Lv + (size_t)(m + 1) * nb2 + (size_t)col * bs + row) * __ldg(subd + (size_t)m * bs + col);
          S[a][b] = v;
        }
      }
      if (j + 1 < nblk)
        load_block(j + 1); // lands during the elimination
      // ---- Gauss-Jordan with implicit partial pivoting ---------------------------------------------------------------
      // Step k: the warp that owns column k ("owner") finds the pivot and publishes {pivot row, 1/pivot, column k};
      // every warp then updates its columns. The owner of step k+1 runs ahead: as soon as step k's data is there it
      // updates column k+1 alone, searches, publishes and ARRIVES at step k+1's named barrier without waiting, and
      // only then finishes its other columns -- so the pivot search of the next step hides under the updates of this
      // one. The other warps SYNC on that barrier when they get there. Two barrier ids and two publish buffers
      // alternate; an owner keeps what it published in registers and never re-reads it.
      unsigned int used = 0; // bit a: my row lane + 32a has been a pivot
      int p_cur = 0;
      double r_cur = 0., c_cur[R], m_cur[R], k_cur[R], rs[R];
#pragma unroll
      for (int a = 0; a < R; ++a)
        c_cur[a] = 0., rs[a] = 1.;
      // pivot search on my column slot B (compile-time) for step k1, published into buffer k1 & 1
      auto search_publish = [&](auto B, int k1, int &p_o, double &r_o, double (&c_o)[R]) {
        constexpr int b1 = decltype(B)::value;
        // Row of maximum modulus among the unused rows, compared on the upper 32 bits of |a| (sign-free exponent and
        // 20 mantissa bits: entries within 1e-6 of each other may tie, the lowest lane wins -- harmless for the
        // stability of the elimination, and one redux.sync instead of three): key = bits + 1, 0 for rows out of play.
        unsigned int key = 0u;
        int krow = 0;
        double kval = 0.;
#pragma unroll
        for (int a = 0; a < R; ++a)
        {
          const int row = lane + 32 * a;
          if (row < bs && !((used >> a) & 1u))
          {
            const unsigned int kk = (unsigned int)__double2hiint(fabs(S[a][b1])) + 1u;
            if (kk > key)
              key = kk, krow = row, kval = S[a][b1];
          }
        }
        const unsigned int m1 = __reduce_max_sync(0xffffffffu, key);
        const int src = __ffs(__ballot_sync(0xffffffffu, key == m1)) - 1;
        p_o = __shfl_sync(0xffffffffu, krow, src);
        const double pv = __shfl_sync(0xffffffffu, kval, src);
        r_o = fast_rcp(pv);
        const int buf1 = k1 & 1;
#pragma unroll
        for (int a = 0; a < R; ++a)
        {
          c_o[a] = S[a][b1];
          sf[buf1 * fstride + lane + 32 * a] = c_o[a];
        }
        if (lane == 0)
        {
          srinv[buf1] = r_o;
          sp[buf1] = p_o;
          sperm[k1] = p_o;
          sinvp[p_o] = k1;
        }
        // (bar.arrive orders the stores above before the barrier completes for the threads that bar.sync on it: the
        // producer / consumer pattern of the PTX manual; an explicit fence here costs ~150 cycles on the critical path)
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + buf1), "r"(INT_) : "memory");
      };
      // One column slot of the step's update; kcol: this is column k itself (owner only). Row scaling is DEFERRED: a
      // pivot row keeps its values (and a 1 in slot k) and remembers the reciprocal of its pivot in rs; Gauss-Jordan
      // is invariant under row scaling as long as every row is updated with ITS OWN entry of column k, which is what
      // m_cur holds (0 for the pivot row itself), so the update is one fma per element for every row, and the rows are
      // scaled once when the block is finished.
      auto update_col = [&](auto B, bool kcol, int pa, int pl) {
        constexpr int b = decltype(B)::value;
        double sel = S[0][b];
#pragma unroll
        for (int a = 1; a < R; ++a)
          if (pa == a)
            sel = S[a][b];
        const double pr = __shfl_sync(0xffffffffu, sel, pl) * r_cur; // scaled pivot-row entry of my column
#pragma unroll
        for (int a = 0; a < R; ++a)
          S[a][b] = kcol ? k_cur[a] : fma(m_cur[a], pr, S[a][b]);
      };
      if (warp == 0)
        search_publish(std::integral_constant<int, 0>{}, 0, p_cur, r_cur, c_cur);
      static_for<C>([&](auto BK) {
        constexpr int bk = decltype(BK)::value;
        constexpr int bk1 = bk + 1 < C ? bk + 1 : bk;
        for (int wk = 0; wk < IW; ++wk)
        {
          const int k = wk + IW * bk;
          if (k >= bs)
            break;
          const int buf = k & 1;
          const bool own = warp == wk;
          if (!own)
          {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + buf), "r"(INT_) : "memory");
            p_cur = sp[buf];
            r_cur = srinv[buf];
#pragma unroll
            for (int a = 0; a < R; ++a)
              c_cur[a] = sf[buf * fstride + lane + 32 * a];
          }
          const int pa = p_cur >> 5, pl = p_cur & 31;
          const bool mine_p = lane == pl;
          if (mine_p)
            used |= 1u << pa;
#pragma unroll
          for (int a = 0; a < R; ++a)
          {
            const bool prow = mine_p && pa == a;
            m_cur[a] = prow ? 0. : -c_cur[a];
            k_cur[a] = prow ? 1. : -(c_cur[a] * r_cur); // what slot k of my rows becomes
            if (prow)
              rs[a] = r_cur;
          }
          // look-ahead: the owner of step k+1 brings its column k+1 up to date first and publishes
          const bool own1 = (k + 1 < bs) && warp == ((wk + 1) & (IW - 1));
          const bool ahead_same = own1 && wk < IW - 1;               // column k+1 sits in my slot bk
          const bool ahead_next = own1 && wk == IW - 1 && bk + 1 < C; // ... in my slot bk + 1
          int p_n = 0;
          double r_n = 0., c_n[R];
#pragma unroll
          for (int a = 0; a < R; ++a)
            c_n[a] = 0.;
          if (ahead_same)
          {
            update_col(std::integral_constant<int, bk>{}, false, pa, pl);
            search_publish(std::integral_constant<int, bk>{}, k + 1, p_n, r_n, c_n);
          }
          if (ahead_next)
          {
            update_col(std::integral_constant<int, bk1>{}, false, pa, pl);
            search_publish(std::integral_constant<int, bk1>{}, k + 1, p_n, r_n, c_n);
          }
          static_for<C>([&](auto B) {
            constexpr int b = decltype(B)::value;
            if (!((ahead_same && b == bk) || (ahead_next && b == bk + 1)))
              update_col(B, own && b == bk, pa, pl);
          });
          if (own1)
          {
            p_cur = p_n;
            r_cur = r_n;
#pragma unroll
            for (int a = 0; a < R; ++a)
              c_cur[a] = c_n[a];
          }
        }
      });
      __syncthreads(); // sperm / sinvp complete; everybody is done reading sInv of the previous block
      // ---- undo the permutation: S[r][m] is entry (step of r, pivot row of step m) of the inverse -----------------------
#pragma unroll
      for (int b = 0; b < C; ++b)
      {
        const int m = warp + IW * b;
#pragma unroll
        for (int a = 0; a < R; ++a)
        {
          const int r = lane + 32 * a;
          if (r < bs && m < bs)
            sInv[(size_t)sperm[m] * bs + sinvp[r]] = S[a][b] * rs[a]; // (the deferred row scaling)
        }
      }
      __syncthreads();
      for (int e = tid; e < nb2; e += INT_)
        Di[(size_t)i * nb2 + e] = sInv[e];
    }
    if (twist && !top)
    { // hand U_m = diag(sup_m) D''_{m+1}^{-1} to the CTA that eliminates the meeting block
      for (int e = tid; e < nb2; e += INT_)
        Lv[(size_t)(m + 1) * nb2 + e] = __ldg(supd + (size_t)m * bs + e % bs) * sInv[e];
      __threadfence();
      cluster_sync_all();
    }
  }
}

int inv_fail(cudaError_t e, const char *what)
{
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  cudaGetLastError();
  return GB_ERR_CUDA;
}

template <int R, int C, int IW>
int launch_invert(int n, const double *mats, int nb, int bs, double *l_values, double *dinv, cudaStream_t st, bool twist)
{
  const size_t nb2 = (size_t)bs * bs;
  const size_t smem = sizeof(double) * (nb2 + (nb2 & 1) + 2 * 32 * R + 2) + sizeof(int) * (2 + 2 * 32 * R) + 16;
  cudaError_t e = cudaFuncSetAttribute(k_btddod_invert<R, C, IW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return inv_fail(e, "k_btddod_invert attribute");
  int dev = 0, sms = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / smem));
  if (twist)
  { // one cluster of two CTAs per system
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (unsigned int)std::min(n, std::max(1, sms * per_sm / 2))), cfg.blockDim = dim3(IW * 32);
    cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, k_btddod_invert<R, C, IW>, n, mats, nb, bs, l_values, dinv, 1);
    ++g_btddod_launches;
    return e == cudaSuccess ? GB_OK : inv_fail(e, "k_btddod_invert (clusters)");
  }
  const int grid = std::min(n, sms * per_sm);
  k_btddod_invert<R, C, IW><<<grid, IW * 32, smem, st>>>(n, mats, nb, bs, l_values, dinv, 0);
  ++g_btddod_launches;
  e = cudaGetLastError();
  return e == cudaSuccess ? GB_OK : inv_fail(e, "k_btddod_invert");
}
} // namespace
} // namespace gb

using namespace gb;

static int invert_impl(int n, const double *matrix, int nb, int bs, double *out_l_values, double *out_dinv, void *stream,
                       bool tw)
{
  if (n < 0 || nb < 1 || bs < 1 || (n > 0 && (!matrix || !out_l_values || !out_dinv)))
  {
    set_error("gb_btddod_full_invert_batch: bad dimensions or null array");
    return GB_ERR_ARG;
  }
  if (n == 0)
    return GB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // warps per CTA: 8; GB_INVERT_WARPS=16 selects the variants with half the columns per thread (measured 3.24 ms
  // against 3.01 ms per GRI-128 system: the step is bound by the owner-to-owner dependent chain, not by the width of
  // a warp's update)
  static int w16 = -1;
  if (w16 < 0)
  {
    const char *e = getenv("GB_INVERT_WARPS");
    w16 = (e && atoi(e) == 16) ? 1 : 0;
  }
  if (bs <= 16)
    return launch_invert<1, 2, 8>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw);
  if (bs <= 32)
    return w16 ? launch_invert<1, 2, 16>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw)
               : launch_invert<1, 4, 8>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw);
  if (bs <= 56)
    return w16 ? launch_invert<2, 4, 16>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw)
               : launch_invert<2, 7, 8>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw);
  if (bs <= 64)
    return w16 ? launch_invert<2, 4, 16>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw)
               : launch_invert<2, 8, 8>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw);
  if (bs <= 96)
    return launch_invert<3, 6, 16>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw);
  if (bs <= 128)
    return launch_invert<4, 8, 16>(n, matrix, nb, bs, out_l_values, out_dinv, st, tw);
  set_error("gb_btddod_full_invert_batch: block size above 128 is not supported");
  return GB_ERR_UNSUPPORTED;
}

extern "C" int gb_btddod_full_invert_batch(int n, const double *matrix, int nb, int bs, double *out_l_values,
                                           double *out_dinv, void *stream)
{
  return invert_impl(n, matrix, nb, bs, out_l_values, out_dinv, stream, false);
}

extern "C" int gb_btddod_full_invert_twisted_batch(int n, const double *matrix, int nb, int bs, double *out_l_values,
                                                   double *out_dinv, void *stream)
{
  return invert_impl(n, matrix, nb, bs, out_l_values, out_dinv, stream, bt_twist_ok(nb, bs));
}
