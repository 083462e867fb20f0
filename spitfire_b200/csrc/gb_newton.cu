// gb_newton.cu -- vector kernels of the batched implicit time integrator (host solver loops).
//
// The reference advances one flamelet at a time with numpy expressions: the ESDIRK stage loop of
// KennedyCarpenterS6P4Q3.single_step (time/methods.py:502-612), SimpleNewtonSolver (time/nonlinear.py:185-268: residual
// dt*(gamma*f(x) + explicit) - (x - q_n), update x <- x - solve(residual), weighted infinity norm against the
// tolerance) and the embedded error estimate handed to the PI controller (time/stepcontrol.py:84-101). For a batch of
// F independent members on the device those expressions are a dozen eager tensor operations and one host
// synchronisation per Newton iteration; here each of them is ONE kernel over all members, the convergence flags stay
// on the device and the host only reads one integer (the number of members still iterating).
//
// Members are the rows of [n][ndof] arrays (row stride ndof). One CTA per member: a GRI-3.0 flamelet row is 6678
// doubles, a batch has at most a few hundred members, so every kernel is a single wave of CTAs that each stream
// ~0.4 MB through one SM -- launch-latency sized, which is the point. The arithmetic follows the reference's
// expressions operation by operation (the library is compiled without FMA contraction), so a member of the batch
// sees exactly the numbers the serial code would produce for it.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <string>

#include "../../include/griffon_b200.h"
#include "gb_kernels.cuh"
#include "gb_mech.h"

namespace gb
{
extern std::atomic<long> g_btddod_launches; // (shared launch counter of the solver-side kernels, gb_btddod.cu)

namespace
{
constexpr int NT = 512;
constexpr int MAXK = 6;

struct KPtrs
{
  const double *k[MAXK];
  double c[MAXK];
  double c2[MAXK];
};

// explicit = c[nk-1]*k[nk-1]; then for j = nk-2 .. 0: explicit = explicit + c[j]*k[j]   (methods.py:560-575)
// res = dt*(gamma*f + explicit) - (x - q)                                               (nonlinear.py:204)
__global__ void __launch_bounds__(NT) k_stage_begin(int ndof, int nk, KPtrs kp, double gamma, const double *__restrict__ dt,
                                                    const double *__restrict__ x, const double *__restrict__ q,
                                                    const double *__restrict__ f, double *__restrict__ expl,
                                                    double *__restrict__ res, int *__restrict__ conv)
{
  const int m = blockIdx.x;
  const size_t base = (size_t)m * ndof;
  const double h = dt[m];
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    double e = kp.c[nk - 1] * kp.k[nk - 1][base + i];
    for (int j = nk - 2; j >= 0; --j)
      e = e + kp.c[j] * kp.k[j][base + i];
    expl[base + i] = e;
    res[base + i] = h * (gamma * f[base + i] + e) - (x[base + i] - q[base + i]);
  }
  if (threadIdx.x == 0)
    conv[m] = 0;
}

// xn = x - dx for the members still iterating, xn = x for the converged ones; resets the unconverged counter that the
// tail kernel of the same iteration accumulates
__global__ void __launch_bounds__(NT) k_newton_update(int ndof, const double *__restrict__ x, const double *__restrict__ dx,
                                                      const int *__restrict__ conv, double *__restrict__ xn,
                                                      int *__restrict__ n_unconverged)
{
  const int m = blockIdx.x;
  const size_t base = (size_t)m * ndof;
  const bool done = conv[m] != 0;
  for (int i = threadIdx.x; i < ndof; i += NT)
    xn[base + i] = done ? x[base + i] : x[base + i] - dx[base + i];
  if (m == 0 && threadIdx.x == 0)
    *n_unconverged = 0;
}

__device__ __forceinline__ void block_max_nan(double &v, int &bad)
{
  __shared__ double sv[NT / 32];
  __shared__ int sb[NT / 32];
  for (int o = 16; o > 0; o >>= 1)
  {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  __syncthreads(); // (the arrays may still be read from a previous call)
  if ((threadIdx.x & 31) == 0)
    sv[threadIdx.x >> 5] = v, sb[threadIdx.x >> 5] = bad;
  __syncthreads();
  v = sv[0], bad = sb[0];
  for (int w = 1; w < NT / 32; ++w)
    v = fmax(v, sv[w]), bad |= sb[w];
}

// rn = dt*(gamma*fn + explicit) - (xn - q); members still iterating take (xn, fn, rn) as their new (x, f, res);
// conv |= max|rn*w| < tol (a NaN anywhere in rn leaves the member unconverged, as the comparison does in the reference)
__global__ void __launch_bounds__(NT) k_newton_tail(int ndof, const double *__restrict__ fn, const double *__restrict__ xn,
                                                    const double *__restrict__ expl, const double *__restrict__ q,
                                                    const double *__restrict__ dt, double gamma,
                                                    const double *__restrict__ w, double tol, double *__restrict__ x,
                                                    double *__restrict__ f, double *__restrict__ res, int *__restrict__ conv,
                                                    int *__restrict__ n_unconverged)
{
  const int m = blockIdx.x;
  const size_t base = (size_t)m * ndof;
  const bool done = conv[m] != 0;
  if (done)
    return; // (uniform over the CTA) a converged member keeps its values
  const double h = dt[m];
  double nrm = 0.;
  int bad = 0;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    const double fi = fn[base + i], xi = xn[base + i];
    const double r = h * (gamma * fi + expl[base + i]) - (xi - q[base + i]);
    x[base + i] = xi;
    f[base + i] = fi;
    res[base + i] = r;
    const double a = fabs(r * w[base + i]);
    bad |= (a != a);
    nrm = fmax(nrm, a);
  }
  block_max_nan(nrm, bad);
  if (threadIdx.x == 0)
  {
    const bool ok = !bad && nrm < tol;
    if (ok)
      conv[m] = 1;
    else
      atomicAdd(n_unconverged, 1);
  }
}

// The Newton tail with a per-member STAGE MACHINE: a member that converges (or runs out of iterations) at stage s stores
// its stage derivative K[s] = f, and, if stages remain, forms the explicit part and the first residual of stage s+1
// right here -- so the members of a batch walk through the implicit stages of a step independently of each other and
// a round of kernels (solve, update, rhs, this one) serves every member at whatever stage it is in. The number of
// rounds of a step is then the largest per-member SUM of Newton iterations, not the sum over the stages of the largest
// per-member count (measured on config 5: 6.4 k against 15.8 k iterations for the slowest member and the lock-step
// batch). Each member's arithmetic is unchanged: same expressions, same order (methods.py:560-575, nonlinear.py:204-257).
struct StageCoef
{
  double a[MAXK][MAXK]; // a[s][j], the Butcher tableau rows of the implicit stages
};
__global__ void __launch_bounds__(NT) k_newton_tail_staged(int n, int ndof, int nstages, StageCoef sc, int max_iter,
                                                           const double *__restrict__ fn, const double *__restrict__ xn,
                                                           const double *__restrict__ q, const double *__restrict__ dt,
                                                           double gamma, const double *__restrict__ w, double tol,
                                                           double *__restrict__ x, double *__restrict__ f,
                                                           double *__restrict__ res, double *__restrict__ expl,
                                                           double *__restrict__ K, int *__restrict__ stage,
                                                           int *__restrict__ iters, int *__restrict__ nlfail,
                                                           int *__restrict__ done, int *__restrict__ n_left)
{
  const int m = blockIdx.x;
  if (done[m])
    return;
  const size_t base = (size_t)m * ndof, kstride = (size_t)n * ndof;
  const double h = dt[m];
  double nrm = 0.;
  int bad = 0;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    const double fi = fn[base + i], xi = xn[base + i];
    const double r = h * (gamma * fi + expl[base + i]) - (xi - q[base + i]);
    x[base + i] = xi;
    f[base + i] = fi;
    res[base + i] = r;
    const double a = fabs(r * w[base + i]);
    bad |= (a != a);
    nrm = fmax(nrm, a);
  }
  block_max_nan(nrm, bad);
  const bool ok = !bad && nrm < tol;
  const int it = iters[m] + 1, s = stage[m];
  const bool stage_end = ok || it >= max_iter;
  __syncthreads(); // (everybody has read iters / stage before thread 0 updates them)
  if (!stage_end)
  {
    if (threadIdx.x == 0)
    {
      iters[m] = it;
      atomicAdd(n_left, 1);
    }
    return;
  }
  const int s1 = s + 1;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    const double fi = f[base + i];
    K[(size_t)s * kstride + base + i] = fi;
    if (s1 < nstages)
    {
      double e = sc.a[s1][s1 - 1] * fi; // (K[s] = f, just stored)
      for (int j = s1 - 2; j >= 0; --j)
        e = e + sc.a[s1][j] * K[(size_t)j * kstride + base + i];
      expl[base + i] = e;
      res[base + i] = h * (gamma * fi + e) - (x[base + i] - q[base + i]);
    }
  }
  if (threadIdx.x == 0)
  {
    if (!ok)
      nlfail[m] = 1;
    iters[m] = 0;
    stage[m] = s1;
    if (s1 < nstages)
      atomicAdd(n_left, 1);
    else
      done[m] = 1;
  }
}

// ---- members that advance independently of each other ACROSS steps (time/batched.py: integrate_batch_async) ----------
// state[m]: 0 = not taking part in this round (waiting for host control / factors, or finished), 1 = BEGIN (the round's
// right-hand side is f(q): it becomes K[0] and the member enters its first implicit stage), 2 = inside the stages.
__global__ void __launch_bounds__(NT) k_async_update(int ndof, const double *__restrict__ x, const double *__restrict__ dx,
                                                     const double *__restrict__ q, const int *__restrict__ state,
                                                     double *__restrict__ xn)
{
  const int m = blockIdx.x, st = state[m];
  if (st == 0)
    return; // (xn keeps a finite, stale value: the round's right-hand side of this member is not used)
  const size_t base = (size_t)m * ndof;
  for (int i = threadIdx.x; i < ndof; i += NT)
    xn[base + i] = st == 1 ? q[base + i] : x[base + i] - dx[base + i];
}

__global__ void __launch_bounds__(NT) k_async_tail(int n, int ndof, int nstages, StageCoef sc, int max_iter,
                                                   const double *__restrict__ fn, const double *__restrict__ xn,
                                                   const double *__restrict__ q, const double *__restrict__ dt,
                                                   double gamma, const double *__restrict__ w, double tol,
                                                   double *__restrict__ x, double *__restrict__ f, double *__restrict__ res,
                                                   double *__restrict__ expl, double *__restrict__ K, int *__restrict__ state,
                                                   int *__restrict__ stage, int *__restrict__ iters,
                                                   int *__restrict__ nlfail, int *__restrict__ newton_its)
{
  const int m = blockIdx.x, st = state[m];
  if (st == 0)
    return;
  const size_t base = (size_t)m * ndof, kstride = (size_t)n * ndof;
  const double h = dt[m];
  if (st == 1)
  { // first stage of a new step: K[0] = f(q), (x, f) = (q, K[0]), explicit part and residual of stage 1
    for (int i = threadIdx.x; i < ndof; i += NT)
    {
      const double fi = fn[base + i], qi = q[base + i];
      K[base + i] = fi;
      x[base + i] = qi;
      f[base + i] = fi;
      const double e = sc.a[1][0] * fi;
      expl[base + i] = e;
      res[base + i] = h * (gamma * fi + e) - (qi - qi);
    }
    if (threadIdx.x == 0)
    {
      state[m] = 2;
      stage[m] = 1;
      iters[m] = 0;
      nlfail[m] = 0;
    }
    return;
  }
  double nrm = 0.;
  int bad = 0;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    const double fi = fn[base + i], xi = xn[base + i];
    const double r = h * (gamma * fi + expl[base + i]) - (xi - q[base + i]);
    x[base + i] = xi;
    f[base + i] = fi;
    res[base + i] = r;
    const double a = fabs(r * w[base + i]);
    bad |= (a != a);
    nrm = fmax(nrm, a);
  }
  block_max_nan(nrm, bad);
  const bool ok = !bad && nrm < tol;
  const int it = iters[m] + 1, s = stage[m];
  const bool stage_end = ok || it >= max_iter;
  __syncthreads();
  if (threadIdx.x == 0)
    newton_its[m] += 1;
  if (!stage_end)
  {
    if (threadIdx.x == 0)
      iters[m] = it;
    return;
  }
  const int s1 = s + 1;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    const double fi = f[base + i];
    K[(size_t)s * kstride + base + i] = fi;
    if (s1 < nstages)
    {
      double e = sc.a[s1][s1 - 1] * fi;
      for (int j = s1 - 2; j >= 0; --j)
        e = e + sc.a[s1][j] * K[(size_t)j * kstride + base + i];
      expl[base + i] = e;
      res[base + i] = h * (gamma * fi + e) - (x[base + i] - q[base + i]);
    }
  }
  if (threadIdx.x == 0)
  {
    if (!ok)
      nlfail[m] = 1;
    iters[m] = 0;
    stage[m] = s1;
    if (s1 >= nstages)
      state[m] = 0; // the step's stages are complete: the host takes over (stage[m] == nstages tells it so)
  }
}

// members the host sends into a new step: state 1 (BEGIN), stage 0, their step size
__global__ void k_async_start(int n, const int *__restrict__ start, const double *__restrict__ dt_in, int *__restrict__ state,
                              int *__restrict__ stage, double *__restrict__ dt)
{
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n && start[m])
  {
    state[m] = 1;
    stage[m] = 0;
    dt[m] = dt_in[m];
  }
}

// members that have completed the stages of a step (state 0, stage == nstages) and whose update is finite (stats[2]):
// q <- q + dq, clipped at zero if asked (the accepted step of integrate_batch); stage is reset so that the completion
// is reported once
__global__ void __launch_bounds__(NT) k_async_accept(int n, int ndof, int nstages, const double *__restrict__ dq,
                                                     const double *__restrict__ stats, int clip, const int *__restrict__ state,
                                                     int *__restrict__ stage, double *__restrict__ q)
{
  const int m = blockIdx.x;
  if (state[m] != 0 || stage[m] != nstages)
    return;
  const bool ok = stats[2 * n + m] > 0.5;
  const size_t base = (size_t)m * ndof;
  if (ok)
    for (int i = threadIdx.x; i < ndof; i += NT)
    {
      double v = q[base + i] + dq[base + i];
      if (clip && v < 0.)
        v = 0.;
      q[base + i] = v;
    }
  __syncthreads();
  if (threadIdx.x == 0)
    stage[m] = 0;
}

// k_async_accept that also PUBLISHES the outcome of the round to the host: every member's state and stage (as they
// stood before the reset) and, for the members that completed their step, the statistics, the Newton-failure flag and
// the new state row -- written straight into mapped pinned host memory, so that a round of the asynchronous integrator
// ends with one synchronisation and no copy operations (gb_flamelet_async_tick_batch).
__global__ void __launch_bounds__(NT) k_async_accept_publish(int n, int ndof, int nstages, const double *__restrict__ dq,
                                                             const double *__restrict__ stats, int clip,
                                                             const int *__restrict__ state, int *__restrict__ stage,
                                                             double *__restrict__ q, const int *__restrict__ nlfail,
                                                             int *__restrict__ h_state, int *__restrict__ h_stage,
                                                             double *__restrict__ h_stats, int *__restrict__ h_nlfail,
                                                             double *__restrict__ h_q)
{
  const int m = blockIdx.x;
  const int st = state[m], sg = stage[m];
  const bool done = st == 0 && sg == nstages;
  if (done)
  {
    const bool ok = stats[2 * n + m] > 0.5;
    const size_t base = (size_t)m * ndof;
    for (int i = threadIdx.x; i < ndof; i += NT)
    {
      double v = q[base + i];
      if (ok)
      {
        v = v + dq[base + i];
        if (clip && v < 0.)
          v = 0.;
        q[base + i] = v;
      }
      h_q[base + i] = v;
    }
    if (threadIdx.x < 3)
      h_stats[threadIdx.x * n + m] = stats[threadIdx.x * n + m];
    if (threadIdx.x == 3)
      h_nlfail[m] = nlfail[m];
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    h_state[m] = st;
    h_stage[m] = sg;
    if (done)
      stage[m] = 0;
  }
}

// dq = dt*(b0*k0 + b1*k1 + ... ), dqh likewise with bh (left to right, methods.py:598-610); stats[0][m] = max|(dq-dqh)*w|
// (the error estimate of the PI controller), stats[1][m] = max|dq*w|, stats[2][m] = 1 if every dq is finite else 0
__global__ void __launch_bounds__(NT) k_esdirk_finish(int ndof, int n, int nk, KPtrs kp, const double *__restrict__ dt,
                                                      const double *__restrict__ w, double *__restrict__ dq,
                                                      double *__restrict__ stats)
{
  const int m = blockIdx.x;
  const size_t base = (size_t)m * ndof;
  const double h = dt[m];
  double e_err = 0., e_dq = 0.;
  int nan_err = 0, nan_dq = 0, nonfinite = 0;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    double s = kp.c[0] * kp.k[0][base + i], sh = kp.c2[0] * kp.k[0][base + i];
    for (int j = 1; j < nk; ++j)
    {
      s = s + kp.c[j] * kp.k[j][base + i];
      sh = sh + kp.c2[j] * kp.k[j][base + i];
    }
    const double d = h * s, dh = h * sh;
    dq[base + i] = d;
    const double wi = w[base + i];
    const double a = fabs((d - dh) * wi), b = fabs(d * wi);
    nan_err |= (a != a);
    nan_dq |= (b != b);
    nonfinite |= !isfinite(d);
    e_err = fmax(e_err, a);
    e_dq = fmax(e_dq, b);
  }
  block_max_nan(e_err, nan_err);
  block_max_nan(e_dq, nan_dq);
  double dummy = 0.;
  block_max_nan(dummy, nonfinite);
  if (threadIdx.x == 0)
  {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    stats[m] = nan_err ? qnan : e_err;
    stats[n + m] = nan_dq ? qnan : e_dq;
    stats[2 * n + m] = nonfinite ? 0. : 1.;
  }
}

// q <- max(q + dq, 0) (or q + dq) for the accepted members, in place (integrator.py:598-612 with the flamelet's
// clip-negative post-step); members with accept[m] == 0 keep their state
__global__ void __launch_bounds__(NT) k_accept_step(int ndof, const double *__restrict__ dq, const int *__restrict__ accept,
                                                    int clip, double *__restrict__ q)
{
  const int m = blockIdx.x;
  if (!accept[m])
    return;
  const size_t base = (size_t)m * ndof;
  for (int i = threadIdx.x; i < ndof; i += NT)
  {
    double v = q[base + i] + dq[base + i];
    if (clip && v < 0.)
      v = 0.;
    q[base + i] = v;
  }
}

// flags[m] = member m of a (and b) holds a non-finite value; *count += 1 per such member. One warp per member.
__global__ void __launch_bounds__(256) k_nonfinite_members(int n, long la, const double *__restrict__ a, long lb,
                                                           const double *__restrict__ b, int *__restrict__ flags,
                                                           int *__restrict__ count)
{
  const int lane = threadIdx.x & 31;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long m = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < n; m += nw)
  {
    int bad = 0;
    const unsigned long long *pa = reinterpret_cast<const unsigned long long *>(a) + m * la;
    for (long i = lane; i < la; i += 32)
      bad |= ((pa[i] >> 52) & 0x7ffull) == 0x7ffull;
    if (b)
    {
      const unsigned long long *pb = reinterpret_cast<const unsigned long long *>(b) + m * lb;
      for (long i = lane; i < lb; i += 32)
        bad |= ((pb[i] >> 52) & 0x7ffull) == 0x7ffull;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0)
    {
      if (flags)
        flags[m] = bad;
      if (bad)
        atomicAdd(count, 1);
    }
  }
}

int *device_counter()
{
  static int *ctr[16] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16)
    return nullptr;
  if (!ctr[dev] && cudaMalloc(&ctr[dev], sizeof(int)) != cudaSuccess)
    ctr[dev] = nullptr;
  return ctr[dev];
}

int check_n(int n, int ndof)
{
  if (n < 0 || ndof <= 0)
  {
    set_error("batched integrator kernels: n >= 0 and ndof > 0 required");
    return GB_ERR_ARG;
  }
  return GB_OK;
}
int cuda_rc(const char *what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return GB_ERR_CUDA;
  }
  return GB_OK;
}
} // namespace

// see k_async_accept_publish; the h_* pointers are device addresses of mapped pinned host arrays
int async_accept_publish(int n, int ndof, int nstages, const double *dq, const double *stats, int clip, const int *state,
                         int *stage, double *q, const int *nlfail, int *h_state, int *h_stage, double *h_stats,
                         int *h_nlfail, double *h_q, cudaStream_t st)
{
  k_async_accept_publish<<<n, NT, 0, st>>>(n, ndof, nstages, dq, stats, clip, state, stage, q, nlfail, h_state, h_stage,
                                           h_stats, h_nlfail, h_q);
  ++g_btddod_launches;
  return cuda_rc("k_async_accept_publish");
}

// the non-finite member count in three steps, so that a pipelined host entry point can accumulate over its chunks
int nonfinite_reset(cudaStream_t st)
{
  int *ctr = device_counter();
  if (!ctr || cudaMemsetAsync(ctr, 0, sizeof(int), st) != cudaSuccess)
  {
    set_error("non-finite count: no device counter");
    return GB_ERR_CUDA;
  }
  return GB_OK;
}
int nonfinite_accumulate(int n, long la, const double *a, long lb, const double *b, int *flags, cudaStream_t st)
{
  int *ctr = device_counter();
  if (!ctr)
    return GB_ERR_CUDA;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<long>(((long)n + 7) / 8, (long)sms * 8);
  k_nonfinite_members<<<grid, 256, 0, st>>>(n, la, a, lb, b, flags, ctr);
  ++g_btddod_launches;
  return cuda_rc("k_nonfinite_members");
}
int nonfinite_read(cudaStream_t st)
{
  int host = 0;
  int *ctr = device_counter();
  if (!ctr || cudaMemcpyAsync(&host, ctr, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return cuda_rc("non-finite count read-back");
  return host;
}
} // namespace gb

using namespace gb;

extern "C"
{
  int gb_esdirk_stage_begin_batch(int n, int ndof, int nk, const double *const *k, const double *coef, double gamma,
                                  const double *dt, const double *x, const double *q, const double *f,
                                  double *explicit_out, double *res_out, int *conv, void *stream)
  {
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    if (nk < 1 || nk > MAXK || !k || !coef || !dt || !x || !q || !f || !explicit_out || !res_out || !conv)
    {
      set_error("gb_esdirk_stage_begin_batch: 1 <= nk <= 6 and non-null arrays required");
      return GB_ERR_ARG;
    }
    KPtrs kp{};
    for (int j = 0; j < nk; ++j)
      kp.k[j] = k[j], kp.c[j] = coef[j];
    k_stage_begin<<<n, NT, 0, (cudaStream_t)stream>>>(ndof, nk, kp, gamma, dt, x, q, f, explicit_out, res_out, conv);
    ++g_btddod_launches;
    return cuda_rc("k_stage_begin");
  }

  int gb_newton_update_batch(int n, int ndof, const double *x, const double *dx, const int *conv, double *xn,
                             int *n_unconverged, void *stream)
  {
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    if (!x || !dx || !conv || !xn || !n_unconverged)
    {
      set_error("gb_newton_update_batch: null array");
      return GB_ERR_ARG;
    }
    k_newton_update<<<n, NT, 0, (cudaStream_t)stream>>>(ndof, x, dx, conv, xn, n_unconverged);
    ++g_btddod_launches;
    return cuda_rc("k_newton_update");
  }

  int gb_newton_tail_batch(int n, int ndof, const double *fn, const double *xn, const double *explicit_, const double *q,
                           const double *dt, double gamma, const double *weights, double tolerance, double *x, double *f,
                           double *res, int *conv, int *n_unconverged, int *host_count, void *stream)
  {
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    if (!fn || !xn || !explicit_ || !q || !dt || !weights || !x || !f || !res || !conv || !n_unconverged)
    {
      set_error("gb_newton_tail_batch: null array");
      return GB_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_newton_tail<<<n, NT, 0, st>>>(ndof, fn, xn, explicit_, q, dt, gamma, weights, tolerance, x, f, res, conv,
                                    n_unconverged);
    ++g_btddod_launches;
    rc = cuda_rc("k_newton_tail");
    if (rc != GB_OK)
      return rc;
    if (host_count)
    { // synchronous read of the one integer the host loop branches on
      if (cudaMemcpyAsync(host_count, n_unconverged, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess)
        return cuda_rc("gb_newton_tail_batch: count read-back");
    }
    return GB_OK;
  }

  int gb_newton_tail_staged_batch(int n, int ndof, int nstages, const double *tableau, int max_iterations,
                                  const double *fn, const double *xn, const double *q, const double *dt, double gamma,
                                  const double *weights, double tolerance, double *x, double *f, double *res,
                                  double *explicit_, double *K, int *stage, int *iters, int *nlfail, int *done,
                                  int *n_left, int *host_count, void *stream)
  {
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    if (nstages < 2 || nstages > MAXK || !tableau || !fn || !xn || !q || !dt || !weights || !x || !f || !res || !explicit_ ||
        !K || !stage || !iters || !nlfail || !done || !n_left)
    {
      set_error("gb_newton_tail_staged_batch: 2 <= nstages <= 6 and non-null arrays required");
      return GB_ERR_ARG;
    }
    StageCoef sc{};
    for (int a = 0; a < nstages; ++a)
      for (int b = 0; b < nstages; ++b)
        sc.a[a][b] = tableau[a * nstages + b];
    cudaStream_t st = (cudaStream_t)stream;
    k_newton_tail_staged<<<n, NT, 0, st>>>(n, ndof, nstages, sc, max_iterations, fn, xn, q, dt, gamma, weights, tolerance,
                                           x, f, res, explicit_, K, stage, iters, nlfail, done, n_left);
    ++g_btddod_launches;
    rc = cuda_rc("k_newton_tail_staged");
    if (rc != GB_OK)
      return rc;
    if (host_count)
    {
      if (cudaMemcpyAsync(host_count, n_left, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess)
        return cuda_rc("gb_newton_tail_staged_batch: count read-back");
    }
    return GB_OK;
  }

  int gb_async_round_kernels(int n, int ndof, int nstages, const double *tableau, int max_iterations, int phase,
                             const double *fn, double *xn, const double *dx, const double *q, const double *dt,
                             double gamma, const double *weights, double tolerance, double *x, double *f, double *res,
                             double *explicit_, double *K, int *state, int *stage, int *iters, int *nlfail,
                             int *newton_its, void *stream)
  { // phase 0: the update kernel (before the right-hand side), phase 1: the tail kernel (after it)
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (phase == 0)
    {
      k_async_update<<<n, NT, 0, st>>>(ndof, x, dx, q, state, xn);
      ++g_btddod_launches;
      return cuda_rc("k_async_update");
    }
    if (phase == 2)
    { // start: fn = start flags (int, device), dx = step sizes (device); dt is written
      k_async_start<<<(n + 127) / 128, 128, 0, st>>>(n, reinterpret_cast<const int *>(fn), dx, state, stage,
                                                     const_cast<double *>(dt));
      ++g_btddod_launches;
      return cuda_rc("k_async_start");
    }
    if (phase == 3)
    { // accept: fn = dq, dx = stats [3][n], max_iterations = clip flag; q is written
      k_async_accept<<<n, NT, 0, st>>>(n, ndof, nstages, fn, dx, max_iterations, state, stage, const_cast<double *>(q));
      ++g_btddod_launches;
      return cuda_rc("k_async_accept");
    }
    if (nstages < 2 || nstages > MAXK || !tableau)
    {
      set_error("gb_async_round_kernels: 2 <= nstages <= 6 and a tableau required");
      return GB_ERR_ARG;
    }
    StageCoef sc{};
    for (int a = 0; a < nstages; ++a)
      for (int b = 0; b < nstages; ++b)
        sc.a[a][b] = tableau[a * nstages + b];
    k_async_tail<<<n, NT, 0, st>>>(n, ndof, nstages, sc, max_iterations, fn, xn, q, dt, gamma, weights, tolerance, x, f, res,
                                   explicit_, K, state, stage, iters, nlfail, newton_its);
    ++g_btddod_launches;
    return cuda_rc("k_async_tail");
  }

  int gb_esdirk_finish_batch(int n, int ndof, int nk, const double *const *k, const double *b, const double *bh,
                             const double *dt, const double *weights, double *dq, double *stats, void *stream)
  {
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    if (nk < 1 || nk > MAXK || !k || !b || !bh || !dt || !weights || !dq || !stats)
    {
      set_error("gb_esdirk_finish_batch: 1 <= nk <= 6 and non-null arrays required");
      return GB_ERR_ARG;
    }
    KPtrs kp{};
    for (int j = 0; j < nk; ++j)
      kp.k[j] = k[j], kp.c[j] = b[j], kp.c2[j] = bh[j];
    k_esdirk_finish<<<n, NT, 0, (cudaStream_t)stream>>>(ndof, n, nk, kp, dt, weights, dq, stats);
    ++g_btddod_launches;
    return cuda_rc("k_esdirk_finish");
  }

  int gb_count_nonfinite_members_batch(int n, long len_a, const double *a, long len_b, const double *b, int *flags_out,
                                       void *stream)
  {
    if (n < 0 || len_a <= 0 || (n > 0 && !a) || (b && len_b <= 0))
    {
      set_error("gb_count_nonfinite_members_batch: bad sizes or null array");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = nonfinite_reset(st);
    if (rc != GB_OK)
      return rc;
    rc = nonfinite_accumulate(n, len_a, a, len_b, b, flags_out, st);
    if (rc != GB_OK)
      return rc;
    return nonfinite_read(st);
  }

  int gb_accept_step_batch(int n, int ndof, const double *dq, const int *accept, int clip_negative, double *q,
                           void *stream)
  {
    int rc = check_n(n, ndof);
    if (rc != GB_OK || n == 0)
      return rc;
    if (!dq || !accept || !q)
    {
      set_error("gb_accept_step_batch: null array");
      return GB_ERR_ARG;
    }
    k_accept_step<<<n, NT, 0, (cudaStream_t)stream>>>(ndof, dq, accept, clip_negative, q);
    ++g_btddod_launches;
    return cuda_rc("k_accept_step");
  }
}
