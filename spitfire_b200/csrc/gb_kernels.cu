// gb_kernels.cu -- FP64 sm_100a kernels for Griffon's batched chemistry:
//   k_rates : production rates / isobaric reactor RHS / flamelet RHS           (chemistry_kernels.cpp:35-463,
//             isobaric_reactor_kernels.cpp:170-219, flamelet_kernels.cpp:1039-1218)
//   (k_jac  : exact rate sensitivities / reactor Jacobian / flamelet Jacobian lives in gb_jac.cu)
//   k_thermo: batched thermodynamic helpers                                    (thermodynamics_kernels.cpp)
//
// Work decomposition: a CTA owns a tile of G thermochemical states whose working set lives
// in shared memory with the state index fastest ([quantity][g], stride GS odd => conflict-free both when lanes
// walk g and when they walk the quantity index). Three kinds of phases alternate, separated by __syncthreads():
//   species phase : one thread per (g, species)   -> NASA7 cp, h, dcp/dT, Gibbs, dB/dT
//   reaction phase: one thread per (g, reaction)  -> rate of progress q_r and its sparse sensitivity record
//                   {q, dq/drho, dq/dT, a, b, dq/dY_slot...}; the dense part of dq/dY of third-body / last-species
//                   reactions is carried as the two scalars a_r, b_r:  dq/dY_s = sparse_s + a_r*u_s + b_r with
//                   u_s = 1/M_s - 1/M_ns  (so no reaction ever touches all ns-1 columns)
//   row phase     : one thread per (g, species row), walks the reactions in which the species is a net species in
//                   ascending reaction order (the reference's accumulation order) and accumulates its own row of
//                   d w / d(rho,T,Y) in shared memory -- no atomics, bit-reproducible.
// The epilogue applies chem_jac_isobaric + transform_isobaric_primitive_jacobian and streams the ns x ns block to
// HBM fully coalesced. HBM traffic per state is the algorithmic minimum: ns doubles in, ns + ns^2 doubles out.
#include "gb_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "gb_device.cuh"

namespace gb
{

static std::atomic<long> g_launches{0};
void count_launch() { ++g_launches; } // kernels launched from other translation units (gb_eig.cu)
extern std::atomic<long> g_btddod_launches; // gb_btddod.cu
extern std::atomic<long> g_jac_launches;    // gb_jac.cu
long kernel_launch_count() { return g_launches.load() + g_btddod_launches.load() + g_jac_launches.load(); }

// ------------------------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------------------------
#define SM(arr, idx, g) (arr)[(idx)*GS + (g)]

// concentration of species s for state g: (y*rho)*invmw as in `C_R(i)` chemistry_kernels.cpp:370
#define CONC(s) (SM(sy, (s), g) * rho * dm.invmw[(s)])

// multiply v by the concentrations of list entries != skip. seq=true reproduces the reference's special-cased
// orders ((v*C)*C), seq=false its generic branch (v*(C*C), v*((C*C)*C)); use_pow handles |nu|>3 where the reference
// does (rates_sensitivities_exact.cpp:702-704).
__device__ __forceinline__ double mult_conc(double v, const DeviceMech &dm, const short *idx, const signed char *st,
                                            int n, int skip, bool seq, bool use_pow, const double *sy, int GS, int g,
                                            double rho)
{
  for (int i = 0; i < n; ++i)
  {
    if (i == skip)
      continue;
    const double c = CONC(idx[i]);
    const int nu = st[i];
    if (seq)
    {
      for (int k = 0; k < nu; ++k)
        v *= c;
    }
    else
    {
      if (nu == 1)
        v *= c;
      else if (nu == 2)
        v *= c * c;
      else if (nu == 3)
        v *= c * c * c;
      else if (use_pow)
        v *= pow(c, (double)nu);
    }
  }
  return v;
}

// d(prod C^nu)/dY at list position `which`, continuing the left-to-right product a * ... (a = k*rho/M_which)
__device__ __forceinline__ double dconc(double a, const DeviceMech &dm, const short *idx, const signed char *st, int n,
                                        int which, bool seq, bool use_pow, const double *sy, int GS, int g,
                                        double rho)
{
  const int nu = st[which];
  const double c = CONC(idx[which]);
  if (nu == 2)
    a = a * 2. * c;
  else if (nu == 3)
    a = a * 3. * c * c;
  else if (nu > 3)
    a = use_pow ? a * (double)nu * pow(c, (double)(nu - 1)) : 0.;
  return mult_conc(a, dm, idx, st, n, which, seq, use_pow, sy, GS, g, rho);
}

// ------------------------------------------------------------------------------------------------------------------
// tile prologue: load states, y_ns = 1 - sum, mixture molecular weight, density
// ------------------------------------------------------------------------------------------------------------------
template <bool JAC>
__device__ __forceinline__ void load_tile(const ChemArgs &a, int tile0, int gcount, double *sc, double *sy)
{
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, GS = a.GS;
  const bool state_mode = a.in_state != nullptr;
  if (state_mode)
  {
    // state = [T, Y_0..Y_{ns-2}]
    for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
    {
      const int g = item / ns, j = item - g * ns;
      const int sidx = a.mode == MODE_FLAMELET_RHS ? flamelet_state_index(a.fl, tile0 + g) : tile0 + g;
      const double v = a.in_state[(size_t)sidx * ns + j];
      if (j == 0)
        SM(sc, Tile::S_T, g) = v;
      else
        SM(sy, j - 1, g) = v;
    }
  }
  else
  {
    for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
    {
      const int g = item / ns, j = item - g * ns;
      SM(sy, j, g) = a.in_y[(size_t)(tile0 + g) * ns + j];
    }
    for (int g = threadIdx.x; g < gcount; g += blockDim.x)
    {
      SM(sc, Tile::S_T, g) = a.in_T[tile0 + g];
      SM(sc, Tile::S_RHO, g) = a.in_rho[tile0 + g];
    }
  }
}

// sequential per-state sums in the reference's order (extract_y combustion_kernels.h:505-515,
// mixture_molecular_weight :381-387, ideal_gas_density :526-530)
__device__ __forceinline__ void state_scalars(const ChemArgs &a, int g, double *sc, double *sy)
{
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, GS = a.GS;
  if (a.in_state != nullptr)
  {
    double yl = 1.;
    for (int j = 0; j < ns - 1; ++j)
      yl -= SM(sy, j, g);
    SM(sy, ns - 1, g) = yl;
  }
  double d = 0.;
  for (int i = 0; i < ns; ++i)
    d += dm.invmw[i] * SM(sy, i, g);
  const double mmw = 1. / d;
  const double T = SM(sc, Tile::S_T, g);
  SM(sc, Tile::S_MMW, g) = mmw;
  SM(sc, Tile::S_LOGT, g) = log(T);
  SM(sc, Tile::S_INVT, g) = 1. / T;
  if (a.in_state != nullptr)
    SM(sc, Tile::S_RHO, g) = a.p * mmw / (T * dm.Ru);
}

// ------------------------------------------------------------------------------------------------------------------
// k_rates
// ------------------------------------------------------------------------------------------------------------------
// rate of progress (k - kr) of reaction r for state g, with production_rates' own expressions
__device__ __forceinline__ double rate_of_progress(const DeviceMech &dm, int r, int g, int GS, const double *sc,
                                                   const double *sy, const double *sg)
{
  const int f = dm.flags[r];
  const double T = SM(sc, Tile::S_T, g), invT = SM(sc, Tile::S_INVT, g), logT = SM(sc, Tile::S_LOGT, g);
  const double rho = SM(sc, Tile::S_RHO, g);
  const double conc = rho / SM(sc, Tile::S_MMW, g);
  double k = rate_constant(f_kform(f), dm.kfA[r], dm.kfb[r], dm.kfE[r], T, invT, logT);
  const int type = f_type(f);
  if (type != RT_SIMPLE)
  {
    // third-body concentration with the association of the switch(n_tb) ladders, chemistry_kernels.cpp:164-199
    const int t0 = dm.tb_off[r], ntb = dm.tb_off[r + 1] - t0;
    const double base = dm.base_eff[r];
    double m;
    if (ntb == 0)
      m = base * conc;
    else
    {
      double s = dm.tb_eff[t0] * SM(sy, dm.tb_idx[t0], g);
      const int n8 = ntb < 8 ? ntb : 8;
      for (int i = 1; i < n8; ++i)
        s = s + dm.tb_eff[t0 + i] * SM(sy, dm.tb_idx[t0 + i], g);
      m = base * conc + rho * (s);
      for (int i = 8; i < ntb; ++i)
        m += rho * (dm.tb_eff[t0 + i] * SM(sy, dm.tb_idx[t0 + i], g));
    }
    if (type == RT_THIRD_BODY)
      k *= m;
    else
    {
      const double kp = dm.kpA[r] * exp(dm.kpb[r] * logT - dm.kpE[r] * invT);
      if (type == RT_LINDEMANN)
        k /= (1 + k / (kp * m));
      else
      { // TROE, chemistry_kernels.cpp:239-314
        const double pr = kp / k * m;
        const double *troe = dm.troe + 4 * (size_t)r;
        const int tb = f_troe(f);
        double fc = 0.0;
        fc = (tb & TROE_T3) ? (1 - troe[0]) * exp(-T / troe[1]) : 0.0;
        fc = fc + ((tb & TROE_T1) ? troe[0] * exp(-T / troe[2]) : 0.0);
        fc = fc + ((tb & TROE_T2) ? exp(-invT * troe[3]) : 0.0);
        const double logFCent = log10(fc);
        const double logPrC = log10(fmax(pr, 1.e-300)) + (-0.4 - 0.67 * logFCent);
        const double f1 = logPrC / ((0.75 - 1.27 * logFCent) - 0.14 * logPrC);
        k = k * pow(10., logFCent / (1 + f1 * f1)) * pr / (1 + pr);
      }
    }
  }
  double kr = 0.;
  if (f & F_HAS_ORDERS)
  { // chemistry_kernels.cpp:325-337
    const int n = dm.n_sp[r];
    for (int i = 0; i < n; ++i)
    {
      const double ord = dm.sp_order[NSR * (size_t)r + i];
      if (fabs(ord) > 1.e-12)
      {
        const int s = dm.sp_idx[NSR * (size_t)r + i];
        k *= pow(fmax(CONC(s), 0.), ord);
      }
    }
  }
  else
  {
    const short *rc = dm.rc_idx + NSR * (size_t)r, *pd = dm.pd_idx + NSR * (size_t)r;
    const signed char *rcs = dm.rc_st + NSR * (size_t)r, *pds = dm.pd_st + NSR * (size_t)r;
    const int nrc = dm.n_rc[r], npd = dm.n_pd[r];
    if (f & F_REVERSIBLE)
    { // chemistry_kernels.cpp:341-368
      const int nn = dm.n_net[r];
      const short *ni = dm.net_idx + NSR * (size_t)r;
      const signed char *nst = dm.net_st + NSR * (size_t)r;
      double gs = nst[0] * SM(sg, ni[0], g);
      for (int i = 1; i < nn; ++i)
        gs = gs + nst[i] * SM(sg, ni[i], g);
      const double invRu = 1. / dm.Ru;
      const double port = dm.p_ref * invT * invRu;
      kr = k * exp(dm.sum_stoich[r] * log(port) - invT * invRu * (gs));
    }
    // mass action, chemistry_kernels.cpp:373-452: special orders multiply k by the left-to-right product of
    // concentrations, the generic branch multiplies species by species
    if (f & F_FWD_SPECIAL)
    {
      double p = 1.;
      bool first = true;
      for (int i = 0; i < nrc; ++i)
        for (int q = 0; q < rcs[i]; ++q)
        {
          const double c = CONC(rc[i]);
          p = first ? c : p * c;
          first = false;
        }
      k *= p;
    }
    else
      k = mult_conc(k, dm, rc, rcs, nrc, -1, false, false, sy, GS, g, rho);
    if (f & F_REVERSIBLE)
    {
      if (f & F_REV_SPECIAL)
      {
        double p = 1.;
        bool first = true;
        for (int i = 0; i < npd; ++i)
          for (int q = 0; q < pds[i]; ++q)
          {
            const double c = CONC(pd[i]);
            p = first ? c : p * c;
            first = false;
          }
        kr *= p;
      }
      else
        kr = mult_conc(kr, dm, pd, pds, npd, -1, false, false, sy, GS, g, rho);
    }
  }
  return k - kr;
}

__global__ void __launch_bounds__(512, 1) k_rates(const ChemArgs a)
{
  extern __shared__ double smem[];
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, nr = dm.nr, G = a.G, GS = a.GS;
  double *sc = smem;                 // [NSC][GS]
  double *sy = sc + Tile::NSC * GS;  // [ns][GS]
  double *sg = sy + ns * GS;         // Gibbs
  double *sh = sg + ns * GS;         // enthalpies
  double *scp = sh + ns * GS;        // species cp
  double *sw = scp + ns * GS;        // production rates
  double *sq = sw + ns * GS;         // [nr][GS] rates of progress

  const int ntiles = (a.n + G - 1) / G;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const int tile0 = tile * G;
    const int gcount = min(G, a.n - tile0);
    __syncthreads();
    load_tile<false>(a, tile0, gcount, sc, sy);
    __syncthreads();
    if (threadIdx.x < gcount)
      state_scalars(a, threadIdx.x, sc, sy);
    __syncthreads();
    // species phase
    for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
    {
      const int i = item / gcount, g = item - i * gcount;
      const SpeciesThermo t =
          species_thermo<false>(dm, i, SM(sc, Tile::S_T, g), SM(sc, Tile::S_LOGT, g), SM(sc, Tile::S_INVT, g));
      SM(sg, i, g) = t.g;
      SM(sh, i, g) = t.h;
      SM(scp, i, g) = t.cp;
    }
    __syncthreads();
    // reaction phase (and, on the side, cp = sum y_i cp_i in species order, thermodynamics_kernels.cpp:45-131)
    if (a.mode != MODE_PRODRATES && threadIdx.x < gcount)
    {
      const int g = threadIdx.x;
      double cp = 0.;
      for (int i = 0; i < ns; ++i)
        cp += SM(sy, i, g) * SM(scp, i, g);
      SM(sc, Tile::S_CP, g) = cp;
    }
    for (int item = threadIdx.x; item < gcount * nr; item += blockDim.x)
    {
      const int r = item / gcount, g = item - r * gcount;
      SM(sq, r, g) = rate_of_progress(dm, r, g, GS, sc, sy, sg);
    }
    __syncthreads();
    // row phase: w_i -= nu*MW*(k-kr) over reactions in ascending order, chemistry_kernels.cpp:457-461
    for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
    {
      const int i = item / gcount, g = item - i * gcount;
      double w = 0.;
      for (int c = 0; c < dm.n_chunks; ++c)
      {
        const int p0 = dm.row_off[c * ns + i], p1 = dm.row_off[c * ns + i + 1];
        for (int p = p0; p < p1; ++p)
          w -= dm.row_stmw[p] * SM(sq, dm.row_rxn[p], g);
      }
      SM(sw, i, g) = w;
    }
    __syncthreads();
    // epilogue
    if (a.mode == MODE_PRODRATES)
    {
      for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
      {
        const int g = item / ns, i = item - g * ns;
        a.out0[(size_t)(tile0 + g) * ns + i] = SM(sw, i, g);
      }
    }
    else
    {
      // chem_rhs_isobaric, isobaric_reactor_kernels.cpp:19-29 (+ open / heat-transfer terms :39-56, 31-37, 194-218)
      if (threadIdx.x < gcount)
      {
        const int g = threadIdx.x;
        const double rho = SM(sc, Tile::S_RHO, g), cp = SM(sc, Tile::S_CP, g), T = SM(sc, Tile::S_T, g);
        double d = 0.;
        for (int i = 0; i < ns; ++i)
          d += SM(sh, i, g) * SM(sw, i, g);
        double rhs0 = -d / (rho * cp);
        if (a.mode == MODE_REACTOR_RHS)
        {
          if (a.rx.open)
          {
            // mass_rhs_isobaric: energy part; inflow enthalpies at T_in
            const double Tin = a.rx.T_in, logTin = log(Tin), invTin = 1. / Tin;
            double m0 = 0.;
            {
              const SpeciesThermo tl = species_thermo<false>(dm, ns - 1, Tin, logTin, invTin);
              m0 = (tl.h - SM(sh, ns - 1, g)) * a.rx.y_in[ns - 1];
            }
            for (int i = 0; i < ns - 1; ++i)
            {
              const SpeciesThermo ti = species_thermo<false>(dm, i, Tin, logTin, invTin);
              m0 += (ti.h - SM(sh, i, g)) * a.rx.y_in[i];
            }
            m0 /= cp;
            const double invTau = 1. / a.rx.tau;
            m0 *= invTau;
            rhs0 += m0;
          }
          if (a.rx.heat_option == 1)
            rhs0 = 0.;
          else if (a.rx.heat_option == 2)
          {
            const double Ts = a.rx.T_surf;
            rhs0 += a.rx.SoV / (rho * cp) *
                    (a.rx.h_conv * (a.rx.T_inf - T) + a.rx.eps_rad * 5.67e-8 * (Ts * Ts * Ts * Ts - T * T * T * T));
          }
        }
        SM(sc, Tile::S_AUX0, g) = rhs0;
      }
      __syncthreads();
      if (a.mode == MODE_REACTOR_RHS)
      {
        for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
        {
          const int g = item / ns, j = item - g * ns;
          double v;
          if (j == 0)
            v = SM(sc, Tile::S_AUX0, g);
          else
          {
            const double invRho = 1. / SM(sc, Tile::S_RHO, g);
            v = SM(sw, j - 1, g) * invRho;
            if (a.rx.open)
            {
              const double invTau = 1. / a.rx.tau;
              v += (a.rx.y_in[j - 1] - SM(sy, j - 1, g)) * invTau;
            }
          }
          a.out0[(size_t)(tile0 + g) * ns + j] = v;
        }
      }
      else
      { // MODE_FLAMELET_RHS, flamelet_kernels.cpp:1094-1217
        const FlameletDev &fl = a.fl;
        const int nzi = fl.nzi;
        if (threadIdx.x < gcount)
        {
          const int g = threadIdx.x;
          const int sidx = flamelet_state_index(fl, tile0 + g), F = sidx / nzi, i = sidx - F * nzi;
          const double rho = SM(sc, Tile::S_RHO, g), cp = SM(sc, Tile::S_CP, g), T = SM(sc, Tile::S_T, g);
          double rhs0 = SM(sc, Tile::S_AUX0, g);
          if (!fl.adiabatic)
          {
            const size_t ho = (size_t)F * fl.stride_heat + i;
            const double hc = fl.h_conv[ho], hr = fl.h_rad[ho], Tc = fl.T_conv[ho], Tr = fl.T_rad[ho];
            if (fl.use_scaled_heat_loss)
            {
              const double maxT = fl.maxT[F];
              const double maxT4 = maxT * maxT * maxT * maxT;
              const double Tr4 = Tr * Tr * Tr * Tr;
              const double q = hc * (Tc - T) / (maxT - Tc) + hr * 5.67e-8 * (Tr4 - T * T * T * T) / (maxT4 - Tr4);
              rhs0 += q / (rho * cp);
            }
            else
            {
              const double q = hc * (Tc - T) + hr * 5.67e-8 * (Tr * Tr * Tr * Tr - T * T * T * T);
              rhs0 += q / (rho * cp);
            }
          }
          if (fl.include_enthalpy_flux || fl.include_variable_cp)
          {
            const double *st = a.in_state + (size_t)F * nzi * ns;
            const double *nm1 = (i == 0) ? fl.oxy : st + (size_t)(i - 1) * ns;
            const double *np1 = (i == nzi - 1) ? fl.fuel : st + (size_t)(i + 1) * ns;
            const double mc = fl.mcoeff[(size_t)F * fl.stride_mn + i], nc = fl.ncoeff[(size_t)F * fl.stride_mn + i];
            const double chi = fl.chi[(size_t)F * fl.stride_chi + i]; // sic: chi[i], not chi[i+1]
            const double dTdZ = mc * nm1[0] + nc * np1[0];
            if (fl.include_enthalpy_flux)
            {
              const double cpn = SM(scp, ns - 1, g);
              double dYdZ_cpi = 0.;
              for (int j = 0; j < ns - 1; ++j)
                dYdZ_cpi += (SM(scp, j, g) - cpn) * (mc * nm1[1 + j] + nc * np1[1 + j]);
              rhs0 += 0.5 * chi / cp * dTdZ * dYdZ_cpi;
            }
            if (fl.include_variable_cp)
            {
              const double *cpg = fl.cp_grid + (size_t)F * nzi;
              const double cpm = (i == 0) ? fl.cp_bc[0] : cpg[i - 1];
              const double cpp = (i == nzi - 1) ? fl.cp_bc[1] : cpg[i + 1];
              const double cpz = mc * cpm + nc * cpp;
              rhs0 += 0.5 * chi * cpz / cp * dTdZ;
            }
          }
          SM(sc, Tile::S_AUX0, g) = rhs0;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
        {
          const int g = item / ns, j = item - g * ns;
          const int sidx = flamelet_state_index(fl, tile0 + g), F = sidx / nzi, i = sidx - F * nzi;
          double v;
          if (j == 0)
            v = SM(sc, Tile::S_AUX0, g);
          else
            v = SM(sw, j - 1, g) * (1. / SM(sc, Tile::S_RHO, g));
          // diffusion stencil, :1208-1217
          const double *st = a.in_state + (size_t)F * nzi * ns;
          const size_t co = (size_t)F * fl.stride_coeff + (size_t)i * ns + j;
          const double qm = (i == 0) ? fl.oxy[j] : st[(size_t)(i - 1) * ns + j];
          const double qp = (i == nzi - 1) ? fl.fuel[j] : st[(size_t)(i + 1) * ns + j];
          const double q0 = st[(size_t)i * ns + j];
          v += fl.cmajor[co] * q0 + fl.csub[co] * qm + fl.csup[co] * qp;
          a.out0[(size_t)sidx * ns + j] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_thermo: one thread per state (cold helper path; griffon.pyx:684-758)
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_thermo(const DeviceMech dm, int what, int n, const double *aux, const double *T, const double *y,
                         double *out)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  const int ns = dm.ns;
  const double *ys = y ? y + (size_t)s * ns : nullptr;
  const double t = T ? T[s] : 0.;
  const double logT = T ? log(t) : 0., invT = T ? 1. / t : 0.;
  double mmw = 0.;
  if (ys)
  {
    double d = 0.;
    for (int i = 0; i < ns; ++i)
      d += ys[i] * dm.invmw[i];
    mmw = 1. / d;
  }
  switch (what)
  {
  case 0: // mixture molecular weight
    out[s] = mmw;
    break;
  case 1: // density from p=aux
    out[s] = aux[s] * mmw / (t * dm.Ru);
    break;
  case 2: // pressure from rho=aux
    out[s] = aux[s] * t * dm.Ru / mmw;
    break;
  case 3:
  case 4:
  { // cp_mix / cv_mix
    double cp = 0.;
    for (int i = 0; i < ns; ++i)
      cp += ys[i] * species_thermo<false>(dm, i, t, logT, invT).cp;
    out[s] = (what == 3) ? cp : cp - dm.Ru / mmw;
    break;
  }
  case 5:
  case 6:
  { // enthalpy_mix / energy_mix : inner_product(h_i, y)
    double d = 0.;
    const double RT = dm.Ru * t;
    for (int i = 0; i < ns; ++i)
    {
      double h = species_thermo<false>(dm, i, t, logT, invT).h;
      if (what == 6)
        h -= RT * dm.invmw[i];
      d += ys[i] * h;
    }
    out[s] = d;
    break;
  }
  case 7:
  case 8:
    for (int i = 0; i < ns; ++i)
    {
      double cp = species_thermo<false>(dm, i, t, logT, invT).cp;
      if (what == 8)
        cp -= dm.Ru * dm.invmw[i];
      out[(size_t)s * ns + i] = cp;
    }
    break;
  case 9:
  case 10:
  {
    const double RT = dm.Ru * t;
    for (int i = 0; i < ns; ++i)
    {
      double h = species_thermo<false>(dm, i, t, logT, invT).h;
      if (what == 10)
        h -= RT * dm.invmw[i];
      out[(size_t)s * ns + i] = h;
    }
    break;
  }
  case 11:
    for (int i = 0; i < ns; ++i)
      out[(size_t)s * ns + i] = species_thermo<false>(dm, i, t, logT, invT).dcp;
    break;
  case 12:
    for (int i = 0; i < ns; ++i)
      out[(size_t)s * ns + i] = ys[i] * mmw * dm.invmw[i];
    break;
  }
}

// flamelet pre-pass: cp at every interior point (flamelet_kernels.cpp:1062-1070), max T per flamelet (:1049-1058),
// cp of the two boundary streams (:1076-1086). One WARP per job: jobs [0, F*nzi) are the grid points, the next two the
// streams, the last F the per-flamelet maxima. The species heat capacities of a point are evaluated by the lanes in
// parallel; the two sums that the reference forms in species order (Y_ns = 1 - sum Y_j, cp = sum Y_i cp_i) are then
// accumulated by lane 0 in that order from shared memory, so the value is bit-identical to the serial evaluation
// while the latency drops from ns dependent polynomial evaluations to one.
__global__ void __launch_bounds__(128) k_flamelet_prepass(const DeviceMech dm, int F, int nzi, const double *state,
                                                          const double *oxy, const double *fuel, double *cp_grid,
                                                          double *maxT, double *cp_bc)
{
  extern __shared__ double pp_smem[]; // [warps per block][ns] products Y_i cp_i
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int job = blockIdx.x * (blockDim.x >> 5) + wib;
  const int ns = dm.ns, npts = F * nzi;
  if (job < npts + 2)
  {
    const double *st = job < npts ? state + (size_t)job * ns : (job == npts ? oxy : fuel);
    double *prod = pp_smem + (size_t)wib * ns;
    const double t = st[0], logT = log(t), invT = 1. / t;
    for (int i = lane; i < ns; i += 32)
      prod[i] = species_thermo<false>(dm, i, t, logT, invT).cp; // cp_i for now
    __syncwarp();
    if (lane == 0)
    {
      double yl = 1.;
      for (int j = 0; j < ns - 1; ++j)
        yl -= st[1 + j];
      double cp = 0.;
      for (int i = 0; i < ns; ++i)
      {
        const double yi = (i < ns - 1) ? st[1 + i] : yl;
        cp += yi * prod[i];
      }
      if (job < npts)
        cp_grid[job] = cp;
      else
        cp_bc[job - npts] = cp;
    }
  }
  else if (job < npts + 2 + F)
  {
    const int f = job - npts - 2;
    double m = -1.;
    for (int i = lane; i < nzi; i += 32)
      m = fmax(m, state[((size_t)f * nzi + i) * ns]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0)
      maxT[f] = m;
  }
}

// sub/super-diagonal scalars of the BTDDOD flamelet Jacobian (flamelet_kernels.cpp:1385-1394, scaled :1395-1400)
__global__ void k_flamelet_offdiag(int ns, int F, FlameletDev fl, double *out_jac)
{
  const int nzi = fl.nzi;
  const size_t per = (size_t)(nzi - 1) * ns;
  const size_t jac_stride = (size_t)ns * ((size_t)nzi * ns + 2 * (nzi - 1));
  const size_t total = (size_t)F * per;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
  {
    const size_t f = t / per, e = t - f * per; // e = (iz-1)*ns + iq, iz = 1..nzi-1
    const double *csub = fl.csub + f * fl.stride_coeff, *csup = fl.csup + f * fl.stride_coeff;
    double *J = out_jac + f * jac_stride + (size_t)nzi * ns * ns;
    double a = csub[e + ns], b = csup[e];
    if (fl.scale_and_offset)
    {
      a *= fl.prefactor;
      b *= fl.prefactor;
    }
    J[e] = a;
    J[per + e] = b;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------------------------
static int sm_count()
{
  static int n = 0;
  if (!n)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static size_t rates_smem(const DeviceMech &dm, int GS)
{
  return sizeof(double) * (size_t)GS * (Tile::NSC + 6 * (size_t)dm.ns + dm.nr);
}
static int env_int(const char *name, int dflt)
{
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

cudaError_t launch_rates(const ChemArgs &a_in, cudaStream_t s)
{
  ChemArgs a = a_in;
  const int maxsm = 227 * 1024;
  // tile size: 31 states per CTA for large batches; small batches (a few flamelets) are spread over all SMs with
  // smaller tiles, whose latency is lower (GRI, 1008 states: 72 us with 7-state tiles against 115 us with 31)
  int G = env_int("GB_RATES_G", 0);
  bool dynamic_tiles = false;
  if (G <= 0)
  {
    // ... and batches of a few waves get equal tiles: 7168 states (56 flamelets of 128 points) are 2 x 148 tiles of 25
    // states rather than 148 + 84 tiles of 31 (the second wave costs as much as the first)
    const int sms = sm_count();
    const int waves = std::max(1, (a.n + 31 * sms - 1) / (31 * sms));
    G = std::min(31, std::max(7, (a.n + waves * sms - 1) / (waves * sms)));
    // GB_RATES_TPS = t > 1: batches of up to a few waves are cut into about t tiles per SM, one CTA per tile, so that
    // the hardware deals the tiles to whatever SMs are free (other streams' kernels may hold some: the Jacobian
    // refreshes of the asynchronous integrator do) instead of one tile per SM and a second wave for the unlucky ones
    static const int tps = std::max(1, env_int("GB_RATES_TPS", 1));
    if (tps > 1 && a.n <= 4 * 31 * sms)
    {
      G = std::min(31, std::max(7, (a.n + tps * sms - 1) / (tps * sms)));
      dynamic_tiles = true;
    }
  }
  while (G > 1 && rates_smem(a.dm, G | 1) > (size_t)maxsm)
    G -= 2;
  if (rates_smem(a.dm, G | 1) > (size_t)maxsm)
    return cudaErrorInvalidConfiguration;
  G = std::min(G, std::max(1, a.n));
  a.G = G;
  a.GS = G | 1;
  const size_t sm = rates_smem(a.dm, a.GS);
  static bool attr = false;
  if (!attr)
  {
    cudaFuncSetAttribute(k_rates, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
    attr = true;
  }
  const int ntiles = (a.n + G - 1) / G;
  const int grid = dynamic_tiles ? ntiles : std::min(ntiles, sm_count());
  const int threads = env_int("GB_RATES_THREADS", 512);
  k_rates<<<grid, threads, sm, s>>>(a);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t launch_thermo(const DeviceMech &dm, int what, int n, const double *aux, const double *T, const double *y,
                          double *out, cudaStream_t s)
{
  const int threads = 128;
  k_thermo<<<(n + threads - 1) / threads, threads, 0, s>>>(dm, what, n, aux, T, y, out);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t launch_flamelet_offdiag(const DeviceMech &dm, int F, const FlameletDev &fl, double *out_jac, cudaStream_t s)
{
  if (fl.nzi < 2)
    return cudaSuccess;
  const size_t total = (size_t)F * (fl.nzi - 1) * dm.ns;
  const int threads = 256;
  const int grid = (int)std::min<size_t>((total + threads - 1) / threads, (size_t)sm_count() * 8);
  k_flamelet_offdiag<<<grid, threads, 0, s>>>(dm.ns, F, fl, out_jac);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t launch_flamelet_prepass(const DeviceMech &dm, int F, const double *state, const FlameletDev &fl,
                                    double *cp_grid, double *maxT, double *cp_bc, cudaStream_t s)
{
  const int jobs = F * fl.nzi + 2 + F; // one warp each
  const int threads = 128, wpb = threads / 32;
  k_flamelet_prepass<<<(jobs + wpb - 1) / wpb, threads, sizeof(double) * wpb * dm.ns, s>>>(dm, F, fl.nzi, state, fl.oxy,
                                                                                          fl.fuel, cp_grid, maxT, cp_bc);
  ++g_launches;
  return cudaGetLastError();
}

} // namespace gb
