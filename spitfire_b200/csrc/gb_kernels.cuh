// gb_kernels.cuh -- launch interface of the chemistry kernels (implemented in gb_kernels.cu)
#pragma once
#include <cuda_runtime.h>

#include "gb_mech.h"

namespace gb
{

enum ChemMode : int
{
  MODE_PRODRATES = 0,    // in: T[n], rho[n], y[n][ns]            out0: w[n][ns]
  MODE_REACTOR_RHS = 1,  // in: state[n][ns]                      out0: rhs[n][ns]
  MODE_FLAMELET_RHS = 2, // in: state[F][nzi][ns]                 out0: rhs[F][nzi][ns]
  MODE_SENS = 3,         // in: rho[n], T[n], y[n][ns]            out1: wsens[n][(ns+1)^2]
  MODE_REACTOR_JAC = 4,  // in: state[n][ns]                      out0: rhs[n][ns], out1: jac[n][ns*ns]
  MODE_FLAMELET_JAC = 5  // in: state[F][nzi][ns]                 out1: BTDDOD jac per flamelet
};

struct ReactorDev
{
  double p, T_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV;
  const double *y_in; // device [ns] or null
  int heat_option, open;
};

struct FlameletDev
{
  int nzi;
  const double *oxy, *fuel;                 // device [ns]
  int adiabatic;
  const double *T_conv, *h_conv, *T_rad, *h_rad; // device, per flamelet [nzi]
  const double *cmajor, *csub, *csup;       // device, per flamelet [nzi*ns]
  const double *mcoeff, *ncoeff;            // device, per flamelet [nzi]
  const double *chi;                        // device, per flamelet [nzi+2]
  long stride_heat, stride_coeff, stride_mn, stride_chi;
  int include_enthalpy_flux, include_variable_cp, use_scaled_heat_loss;
  // work arrays produced by the pre-pass (flamelet_prepass): cp at every interior point and at both streams
  const double *cp_grid;                    // [F][nzi]
  const double *maxT;                       // [F]
  double cp_oxy, cp_fuel;                   // filled on device: read through cp_bc
  const double *cp_bc;                      // [2] = {cp(oxy), cp(fuel)}
  // jacobian extras
  int scale_and_offset;
  double prefactor;
  int chem_only; // diagonal blocks as they stand at flamelet_kernels.cpp:1327 (before cmajor etc.): the eigenvalue pass
  // flamelet right-hand side of a SUBSET of the batch (asynchronous integrator: only the members inside a step): the
  // kernel's n = nmembers * nzi states are the points of the flamelets members[0..nmembers); 0 = all, in order
  int nmembers;
  unsigned char members[64];
};
// state index in the batch's arrays of the l-th state a flamelet right-hand side launch works on
__host__ __device__ inline int flamelet_state_index(const FlameletDev &fl, int l)
{
  if (fl.nmembers <= 0)
    return l;
  const int k = l / fl.nzi;
  return (int)fl.members[k] * fl.nzi + (l - k * fl.nzi);
}

struct ChemArgs
{
  DeviceMech dm;
  int mode;
  int n;           // number of thermochemical states (for flamelets F*nzi)
  const double *in_T, *in_rho, *in_y, *in_state;
  double p;        // pressure for state-based modes
  double *out0, *out1;
  ReactorDev rx;
  FlameletDev fl;
  int G;           // states per CTA tile
  int GS;          // smem stride per index (G padded to odd)
  int stagger;     // k_jac: start delay (cycles) per CTA phase slot, see launch_jac_g
};

// launches; return cudaError_t of the launch
cudaError_t launch_rates(const ChemArgs &a, cudaStream_t s);
cudaError_t launch_jac(const ChemArgs &a, cudaStream_t s);
// gb_jac4.cu: warp-specialised reactor-Jacobian kernel for large mechanisms (launch_jac dispatches to it)
bool jac4_applicable(const ChemArgs &a);
cudaError_t launch_jac4(const ChemArgs &a, cudaStream_t s);
cudaError_t launch_thermo(const DeviceMech &dm, int what, int n, const double *aux, const double *T, const double *y,
                          double *out, cudaStream_t s);
// flamelet pre-pass: cp_grid[F][nzi], maxT[F], cp_bc[2]
cudaError_t launch_flamelet_prepass(const DeviceMech &dm, int F, const double *state, const FlameletDev &fl,
                                    double *cp_grid, double *maxT, double *cp_bc, cudaStream_t s);
cudaError_t launch_flamelet_offdiag(const DeviceMech &dm, int F, const FlameletDev &fl, double *out_jac, cudaStream_t s);
// gb_eig.cu: out[b] = max Re(lambda) of the n x n matrix at base + (b / per_f) * stride_f + (b % per_f) * n * n
cudaError_t launch_block_max_real_eig(int nblocks, const double *base, long stride_f, int per_f, int n, double *out,
                                      cudaStream_t s);
// out[b * n + q] = max(maxre[b] - diffterm, 0)
cudaError_t launch_expand_expeig(int nblocks, int n, const double *maxre, double diffterm, double *out, cudaStream_t s);
// ---- twisted ("burn at both ends") block-Thomas elimination (gb_btinv.cu, k_btddod_solve_inv) ----------------------
// Factors made by the twisted elimination carry a tag in block 0 of l_values (never read otherwise): element 0 = the
// meeting block m, element 1 = BT_TWIST_MAGIC. bt_twist_ok: the one rule both kernels' launchers follow (nb >= 4,
// bs >= 2, right-hand side staged in shared memory, not disabled by GB_BT_TWIST=0).
constexpr double BT_TWIST_MAGIC = 2.718281828459045e-300;
bool bt_twist_ok(int nb, int bs);
#ifdef __CUDACC__
__device__ __forceinline__ unsigned int cluster_ctarank()
{
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned int clusterid_x()
{
  unsigned int r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned int nclusterid_x()
{
  unsigned int r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// all threads of both CTAs; orders global and shared-memory accesses across the barrier at cluster scope
__device__ __forceinline__ void cluster_sync_all()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::
                   : "memory");
}
#endif
// gb_newton.cu: acceptance kernel of the asynchronous integrator that writes the round's outcome into mapped host memory
int async_accept_publish(int n, int ndof, int nstages, const double *dq, const double *stats, int clip, const int *state,
                         int *stage, double *q, const int *nlfail, int *h_state, int *h_stage, double *h_stats,
                         int *h_nlfail, double *h_q, cudaStream_t st);
long kernel_launch_count();
void count_launch();
#ifdef GB_JAC_TIMELINE
int debug_jac_timeline(long long *out);
int debug_jac4_timeline(long long *out);
int debug_bt_timeline(long long *out);
#endif

} // namespace gb
