// gb_kernels.cu -- FP64 sm_100a kernels for Griffon's batched chemistry:
//   k_rates : production rates / isobaric reactor RHS / flamelet RHS           (chemistry_kernels.cpp:35-463,
//             isobaric_reactor_kernels.cpp:170-219, flamelet_kernels.cpp:1039-1218)
//   k_jac   : exact rate sensitivities / reactor Jacobian / flamelet Jacobian  (rates_sensitivities_exact.cpp,
//             isobaric_reactor_kernels.cpp:221-343, flamelet_kernels.cpp:1220-1409)
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
extern std::atomic<long> g_btddod_launches; // gb_btddod.cu
long kernel_launch_count() { return g_launches.load() + g_btddod_launches.load(); }

// ------------------------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------------------------
struct Tile
{
  // scalars per state, [NSC][GS]
  enum
  {
    S_T = 0,
    S_LOGT,
    S_INVT,
    S_RHO,
    S_MMW,
    S_CP,
    S_CPSENST,
    S_P,
    S_AUX0,
    S_AUX1,
    S_AUX2,
    S_AUX3,
    S_AUX4,
    S_AUX5,
    S_AUX6,
    S_AUX7,
    NSC
  };
};

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
      const double v = a.in_state[(size_t)(tile0 + g) * ns + j];
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
          const int sidx = tile0 + g, F = sidx / nzi, i = sidx - F * nzi;
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
          const int sidx = tile0 + g, F = sidx / nzi, i = sidx - F * nzi;
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
// k_jac
// ------------------------------------------------------------------------------------------------------------------
// Reaction phase of the Jacobian path: fills reaction r's record for state g.
//   rec[0]=q  rec[1]=dq/drho  rec[2]=dq/dT  rec[3]=a  rec[4]=b  rec[5+k]=sparse dq/dY_{slot k}
// rates_sensitivities_exact.cpp:128-1009 restated per reaction; the dense parts (`for s<ns-1` loops at :522-525,
// :807-810, :849-850, :859-862 ...) are folded into a, b.
__device__ __forceinline__ void jac_reaction(const DeviceMech &dm, int r, int g, int GS, const double *sc,
                                             const double *sy, const double *sg, const double *sdb, double *rec)
{
  const int f = dm.flags[r];
  const int ns = dm.ns;
  const double T = SM(sc, Tile::S_T, g), invT = SM(sc, Tile::S_INVT, g), logT = SM(sc, Tile::S_LOGT, g);
  const double rho = SM(sc, Tile::S_RHO, g);
  const double invM = 1. / SM(sc, Tile::S_MMW, g);
  const double ct = rho * invM;
  const double invRu = 1. / dm.Ru;
  const int nslots = dm.slot_off[r + 1] - dm.slot_off[r];
  for (int k = 0; k < nslots; ++k)
    rec[(5 + k) * GS] = 0.;

  const double kfA = dm.kfA[r], kfb = dm.kfb[r], kfE = dm.kfE[r];
  const double kf = rate_constant(f_kform(f), kfA, kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT); // ARRHENIUS_SENS_OVER_K, :25
  double Rnet, dRnetdrho, dRnetdT;
  double cR = 0.; // dense offset of dRnet/dY_s (all s): -(last species as reactant) +(last species as product)
  const int last = ns - 1;

  if (f & F_HAS_ORDERS)
  { // :198-281
    const int n = dm.n_sp[r];
    const short *sp = dm.sp_idx + NSR * (size_t)r;
    const double *ord = dm.sp_order + NSR * (size_t)r;
    const signed char *spslot = dm.sp_slot + NSR * (size_t)r;
    double sumOrders = 0.;
    Rnet = kf;
    for (int i = 0; i < n; ++i)
      if (fabs(ord[i]) > 1.e-12)
      {
        Rnet *= pow(fmax(CONC(sp[i]), 0.), ord[i]);
        sumOrders += ord[i];
      }
    dRnetdrho = Rnet / ct * invM * sumOrders;
    dRnetdT = Rnet * kf_sens;
    for (int j = 0; j < n; ++j)
    {
      if (!(fabs(ord[j]) > 1.e-12))
        continue;
      const bool is_last = sp[j] == last;
      double v = kf;
      for (int l = 0; l < n; ++l)
      {
        const double cl = CONC(sp[l]);
        if (l != j)
        {
          if (fabs(ord[l]) > 1.e-12)
            v *= is_last ? pow(cl, ord[l]) : pow(fmax(cl, 0.), ord[l]);
        }
        else
        {
          const double pre = ord[l] * rho * dm.invmw[sp[l]];
          if (ord[l] > 1 || is_last)
            v *= pre * pow(fmax(cl, 1.e-16), ord[l] - 1.);
          else
            v *= pre / pow(fmax(cl, 1.e-16), 1. - ord[l]);
        }
      }
      if (is_last)
        cR -= v;
      else
        rec[(5 + spslot[j]) * GS] = v;
    }
  }
  else
  {
    const short *rc = dm.rc_idx + NSR * (size_t)r, *pd = dm.pd_idx + NSR * (size_t)r;
    const signed char *rcs = dm.rc_st + NSR * (size_t)r, *pds = dm.pd_st + NSR * (size_t)r;
    const signed char *rcslot = dm.rc_slot + NSR * (size_t)r, *pdslot = dm.pd_slot + NSR * (size_t)r;
    const int nrc = dm.n_rc[r], npd = dm.n_pd[r];
    const bool fseq = (f & F_FWD_SPECIAL) != 0, rseq = (f & F_REV_SPECIAL) != 0;
    // forward, :287-526
    Rnet = mult_conc(kf, dm, rc, rcs, nrc, -1, fseq, false, sy, GS, g, rho);
    dRnetdrho = Rnet / ct * invM * dm.sum_rc[r];
    dRnetdT = Rnet * kf_sens;
    for (int i = 0; i < nrc; ++i)
    {
      const double d = dconc(kf * rho * dm.invmw[rc[i]], dm, rc, rcs, nrc, i, fseq, false, sy, GS, g, rho);
      if (rc[i] == last)
        cR -= d;
      else
        rec[(5 + rcslot[i]) * GS] = d;
    }
    if (f & F_REVERSIBLE)
    { // :528-812
      double Kc = 1., dKc = 0.;
      {
        const int nn = dm.n_net[r];
        const short *ni = dm.net_idx + NSR * (size_t)r;
        const signed char *nst = dm.net_st + NSR * (size_t)r;
        double gs = nst[0] * SM(sg, ni[0], g);
        double ds = nst[0] * SM(sdb, ni[0], g);
        for (int i = 1; i < nn; ++i)
        {
          gs = gs + nst[i] * SM(sg, ni[i], g);
          ds = ds + nst[i] * SM(sdb, ni[i], g);
        }
        Kc = exp(-(dm.sum_stoich[r] * log(dm.p_ref * invT * invRu) - invT * invRu * (gs)));
        dKc = -ds;
      }
      const double kr = kf / Kc;
      const double Rr = mult_conc(kr, dm, pd, pds, npd, -1, rseq, false, sy, GS, g, rho);
      Rnet -= Rr;
      dRnetdrho -= Rr / ct * invM * dm.sum_pd[r];
      dRnetdT -= Rr * (kf_sens - dKc);
      for (int i = 0; i < npd; ++i)
      {
        const double d = dconc(kr * rho * dm.invmw[pd[i]], dm, pd, pds, npd, i, rseq, true, sy, GS, g, rho);
        if (pd[i] == last)
          cR += d;
        else
          rec[(5 + pdslot[i]) * GS] -= d;
      }
    }
  }

  // third-body / falloff factor C_tbaf and its sensitivities, :826-1000
  const int type = f_type(f);
  double Ctbaf = 1., dCdrho = 0., dCdT = 0., coef = 0.; // dCtbaf/dY_s = coef*(base*u_s + eps_s - eps_last)
  const int t0 = dm.tb_off[r], ntb = dm.tb_off[r + 1] - t0;
  const double base = dm.base_eff[r];
  if (type != RT_SIMPLE)
  {
    double M = base * ct;
    double dMdrho = base * invM;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = dm.tb_eff[t0 + i] * SM(sy, dm.tb_idx[t0 + i], g);
      M = M + rho * e;
      dMdrho += e;
    }
    if (type == RT_THIRD_BODY)
    {
      Ctbaf = M;
      dCdrho = dMdrho;
      dCdT = 0.;
      coef = rho;
    }
    else
    {
      const double kpA = dm.kpA[r], kpb = dm.kpb[r], kpE = dm.kpE[r];
      const double kp_over_kf = kpA * exp(kpb * logT - kpE * invT) / kf;
      const double kp_sens = invT * (kpb + kpE * invT);
      const double pr = kp_over_kf * M;
      double nsTmp;
      if (type == RT_LINDEMANN)
      { // :867-903
        Ctbaf = pr / (1. + pr);
        dCdT = Ctbaf / (1. + pr) * (kp_sens - kf_sens);
        nsTmp = kp_over_kf / ((1. + pr) * (1. + pr));
      }
      else
      { // TROE, :905-995
        const double *troe = dm.troe + 4 * (size_t)r;
        const int tb = f_troe(f);
        double fCent = 0., dfCentdT = 0.;
        if (tb & TROE_T3)
        {
          const double t1exp = exp(-T / troe[1]);
          fCent = (1 - troe[0]) * t1exp;
          dfCentdT = (troe[0] - 1) / troe[1] * t1exp;
        }
        if (tb & TROE_T1)
        {
          const double t2exp = exp(-T / troe[2]);
          fCent = (tb & TROE_T3) ? fCent + troe[0] * t2exp : troe[0] * t2exp;
          dfCentdT = (tb & TROE_T3) ? dfCentdT - troe[0] / troe[2] * t2exp : -troe[0] / troe[2] * t2exp;
        }
        if (tb & TROE_T2)
        {
          const double t3exp = exp(-invT * troe[3]);
          const bool any = (tb & (TROE_T3 | TROE_T1)) != 0;
          fCent = any ? fCent + t3exp : t3exp;
          dfCentdT = any ? dfCentdT + t3exp * troe[3] * invT * invT : t3exp * troe[3] * invT * invT;
        }
        const double log10pr = log10(fmax(pr, 1.e-300));
        const double log10fcent = log10(fmax(fCent, 1.e-300));
        const double logfcent = log(fmax(fCent, 1.e-300));
        const double ln10 = log(10.);
        const double aTroe = log10pr - 0.67 * log10fcent - 0.4;
        const double bTroe = -0.14 * log10pr - 1.1762 * log10fcent + 0.806;
        const double gTroe = 1 / (1 + (aTroe / bTroe) * (aTroe / bTroe));
        const double fTroe = pow(fCent, gTroe);
        Ctbaf = fTroe * pr / (1 + pr);
        const double dfTroedT =
            fTroe * (gTroe / fCent * dfCentdT +
                     logfcent * (-2.0 * gTroe * gTroe / ln10 * aTroe / (bTroe * bTroe * bTroe) *
                                 ((bTroe + 0.14 * aTroe) * (kp_sens - kf_sens) -
                                  (0.67 * bTroe - 1.1762 * aTroe) * dfCentdT / fCent)));
        dCdT = 1. / (1. + 1. / pr) * dfTroedT + fTroe * pr / ((1. + pr) * (1. + pr)) * (kp_sens - kf_sens);
        nsTmp = kp_over_kf * (-2.0 / (1. + pr) * fTroe * logfcent * gTroe * gTroe / ln10 * aTroe /
                                  (bTroe * bTroe * bTroe) * (bTroe + 0.14 * aTroe) +
                              fTroe / ((1. + pr) * (1 + pr)));
      }
      dCdrho = nsTmp * dMdrho; // = nsTmp*base*invM + sum nsTmp*eps_i*y_i, distributed (:880-882)
      coef = nsTmp * rho;
    }
  }

  rec[0] = Rnet * Ctbaf;
  rec[GS] = dRnetdrho * Ctbaf + dCdrho * Rnet;
  rec[2 * GS] = dRnetdT * Ctbaf + dCdT * Rnet;
  double b = cR * Ctbaf;
  if (type != RT_SIMPLE)
  {
    for (int k = 0; k < nslots; ++k)
      rec[(5 + k) * GS] *= Ctbaf;
    const signed char *tbslot = dm.tb_slot + t0;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = coef * dm.tb_eff[t0 + i];
      if (tbslot[i] >= 0)
        rec[(5 + tbslot[i]) * GS] += e * Rnet;
      else
        b -= e * Rnet; // the last species is a third body: -eps_last on every column (:855-865)
    }
    rec[3 * GS] = coef * base * Rnet;
  }
  else
    rec[3 * GS] = 0.;
  rec[4 * GS] = b;
}

// ---- staged (shared-memory) parameter access ---------------------------------------------------------------------
__device__ __forceinline__ double as_double(unsigned long long u) { return __longlong_as_double((long long)u); }

// Reaction phase, fast path: identical arithmetic to jac_reaction() but every parameter comes from the chunk's
// parameter blob staged in shared memory (format: gb_mech.cu commit(), "per-chunk staged images").
__device__ __forceinline__ void jac_reaction_staged(const DeviceMech &dm, const unsigned long long *P, int g, int GS,
                                                    const double *sc, const double *sy, const double *sg,
                                                    const double *sdb, double *srec_g)
{
  const unsigned long long w0 = P[0], w1 = P[1];
  const int f = (int)(unsigned int)w0;
  double *rec = srec_g + (size_t)(unsigned int)(w0 >> 32) * GS;
  const int nrc = (int)(w1 & 255), npd = (int)((w1 >> 8) & 255), nn = (int)((w1 >> 16) & 255),
            ntb = (int)((w1 >> 24) & 255), nslots = (int)((w1 >> 32) & 255);
  const int sum_stoich = (int)(signed char)((w1 >> 40) & 255), sum_rc = (int)((w1 >> 48) & 255),
            sum_pd = (int)((w1 >> 56) & 255);
  const int type = f_type(f);
  const unsigned long long *Prc = P + (type == RT_SIMPLE ? 5 : 13);
  const unsigned long long *Ppd = Prc + 2 * nrc;
  const unsigned long long *Pnet = Ppd + 2 * npd;
  const unsigned long long *Ptb = Pnet + nn;

  const int ns = dm.ns, last = ns - 1;
  const double T = SM(sc, Tile::S_T, g), invT = SM(sc, Tile::S_INVT, g), logT = SM(sc, Tile::S_LOGT, g);
  const double rho = SM(sc, Tile::S_RHO, g);
  const double invM = 1. / SM(sc, Tile::S_MMW, g);
  const double ct = rho * invM;
  const double invRu = 1. / dm.Ru;
  for (int k = 0; k < nslots; ++k)
    rec[(5 + k) * GS] = 0.;

  const double kfb = as_double(P[3]), kfE = as_double(P[4]);
  const double kf = rate_constant(f_kform(f), as_double(P[2]), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT);
  double cR = 0.;

#define SP_IDX(Q, i) ((int)(Q[2 * (i)] & 0xffff))
#define SP_ST(Q, i) ((int)((Q[2 * (i)] >> 16) & 255))
#define SP_SLOT(Q, i) ((int)(signed char)((Q[2 * (i)] >> 24) & 255))
#define SP_INVMW(Q, i) as_double(Q[2 * (i) + 1])
#define SP_CONC(Q, i) (SM(sy, SP_IDX(Q, i), g) * rho * SP_INVMW(Q, i))

  // product of concentrations (except position skip) continuing v, and the derivative wrt position `which`
  auto mult = [&](double v, const unsigned long long *Q, int n, int skip, bool seq, bool use_pow) {
    for (int i = 0; i < n; ++i)
    {
      if (i == skip)
        continue;
      const double c = SP_CONC(Q, i);
      const int nu = SP_ST(Q, i);
      if (seq)
      {
        for (int k = 0; k < nu; ++k)
          v *= c;
      }
      else if (nu == 1)
        v *= c;
      else if (nu == 2)
        v *= c * c;
      else if (nu == 3)
        v *= c * c * c;
      else if (use_pow)
        v *= pow(c, (double)nu);
    }
    return v;
  };
  auto deriv = [&](double a, const unsigned long long *Q, int n, int which, bool seq, bool use_pow) {
    const int nu = SP_ST(Q, which);
    if (nu > 1)
    {
      const double c = SP_CONC(Q, which);
      if (nu == 2)
        a = a * 2. * c;
      else if (nu == 3)
        a = a * 3. * c * c;
      else
        a = use_pow ? a * (double)nu * pow(c, (double)(nu - 1)) : 0.;
    }
    return mult(a, Q, n, which, seq, use_pow);
  };

  const bool fseq = (f & F_FWD_SPECIAL) != 0, rseq = (f & F_REV_SPECIAL) != 0;
  double Rnet = mult(kf, Prc, nrc, -1, fseq, false);
  double dRnetdrho = Rnet / ct * invM * sum_rc;
  double dRnetdT = Rnet * kf_sens;
  for (int i = 0; i < nrc; ++i)
  {
    const double d = deriv(kf * rho * SP_INVMW(Prc, i), Prc, nrc, i, fseq, false);
    if (SP_IDX(Prc, i) == last)
      cR -= d;
    else
      rec[(5 + SP_SLOT(Prc, i)) * GS] = d;
  }
  if (f & F_REVERSIBLE)
  {
    double gs, ds;
    {
      const int i0 = (int)(Pnet[0] & 0xffff), s0 = (int)(signed char)((Pnet[0] >> 16) & 255);
      gs = s0 * SM(sg, i0, g);
      ds = s0 * SM(sdb, i0, g);
    }
    for (int i = 1; i < nn; ++i)
    {
      const int ii = (int)(Pnet[i] & 0xffff), si = (int)(signed char)((Pnet[i] >> 16) & 255);
      gs = gs + si * SM(sg, ii, g);
      ds = ds + si * SM(sdb, ii, g);
    }
    const double Kc = exp(-(sum_stoich * SM(sc, Tile::S_AUX7, g) - invT * invRu * (gs)));
    const double dKc = -ds;
    const double kr = kf / Kc;
    const double Rr = mult(kr, Ppd, npd, -1, rseq, false);
    Rnet -= Rr;
    dRnetdrho -= Rr / ct * invM * sum_pd;
    dRnetdT -= Rr * (kf_sens - dKc);
    for (int i = 0; i < npd; ++i)
    {
      const double d = deriv(kr * rho * SP_INVMW(Ppd, i), Ppd, npd, i, rseq, true);
      if (SP_IDX(Ppd, i) == last)
        cR += d;
      else
        rec[(5 + SP_SLOT(Ppd, i)) * GS] -= d;
    }
  }

  double Ctbaf = 1., dCdrho = 0., dCdT = 0., coef = 0.;
  const double base = (type != RT_SIMPLE) ? as_double(P[5]) : 0.;
  if (type != RT_SIMPLE)
  {
    double M = base * ct;
    double dMdrho = base * invM;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = as_double(Ptb[2 * i + 1]) * SM(sy, (int)(Ptb[2 * i] & 0xffff), g);
      M = M + rho * e;
      dMdrho += e;
    }
    if (type == RT_THIRD_BODY)
    {
      Ctbaf = M;
      dCdrho = dMdrho;
      coef = rho;
    }
    else
    {
      const double kpb = as_double(P[7]), kpE = as_double(P[8]);
      const double kp_over_kf = as_double(P[6]) * exp(kpb * logT - kpE * invT) / kf;
      const double kp_sens = invT * (kpb + kpE * invT);
      const double pr = kp_over_kf * M;
      double nsTmp;
      if (type == RT_LINDEMANN)
      {
        Ctbaf = pr / (1. + pr);
        dCdT = Ctbaf / (1. + pr) * (kp_sens - kf_sens);
        nsTmp = kp_over_kf / ((1. + pr) * (1. + pr));
      }
      else
      {
        const double tr0 = as_double(P[9]), tr1 = as_double(P[10]), tr2 = as_double(P[11]), tr3 = as_double(P[12]);
        const int tb = f_troe(f);
        double fCent = 0., dfCentdT = 0.;
        if (tb & TROE_T3)
        {
          const double t1exp = exp(-T / tr1);
          fCent = (1 - tr0) * t1exp;
          dfCentdT = (tr0 - 1) / tr1 * t1exp;
        }
        if (tb & TROE_T1)
        {
          const double t2exp = exp(-T / tr2);
          fCent = (tb & TROE_T3) ? fCent + tr0 * t2exp : tr0 * t2exp;
          dfCentdT = (tb & TROE_T3) ? dfCentdT - tr0 / tr2 * t2exp : -tr0 / tr2 * t2exp;
        }
        if (tb & TROE_T2)
        {
          const double t3exp = exp(-invT * tr3);
          const bool any = (tb & (TROE_T3 | TROE_T1)) != 0;
          fCent = any ? fCent + t3exp : t3exp;
          dfCentdT = any ? dfCentdT + t3exp * tr3 * invT * invT : t3exp * tr3 * invT * invT;
        }
        const double log10pr = log10(fmax(pr, 1.e-300));
        const double log10fcent = log10(fmax(fCent, 1.e-300));
        const double logfcent = log(fmax(fCent, 1.e-300));
        const double ln10 = log(10.);
        const double aTroe = log10pr - 0.67 * log10fcent - 0.4;
        const double bTroe = -0.14 * log10pr - 1.1762 * log10fcent + 0.806;
        const double gTroe = 1 / (1 + (aTroe / bTroe) * (aTroe / bTroe));
        const double fTroe = pow(fCent, gTroe);
        Ctbaf = fTroe * pr / (1 + pr);
        const double dfTroedT =
            fTroe * (gTroe / fCent * dfCentdT +
                     logfcent * (-2.0 * gTroe * gTroe / ln10 * aTroe / (bTroe * bTroe * bTroe) *
                                 ((bTroe + 0.14 * aTroe) * (kp_sens - kf_sens) -
                                  (0.67 * bTroe - 1.1762 * aTroe) * dfCentdT / fCent)));
        dCdT = 1. / (1. + 1. / pr) * dfTroedT + fTroe * pr / ((1. + pr) * (1. + pr)) * (kp_sens - kf_sens);
        nsTmp = kp_over_kf * (-2.0 / (1. + pr) * fTroe * logfcent * gTroe * gTroe / ln10 * aTroe /
                                  (bTroe * bTroe * bTroe) * (bTroe + 0.14 * aTroe) +
                              fTroe / ((1. + pr) * (1 + pr)));
      }
      dCdrho = nsTmp * dMdrho;
      coef = nsTmp * rho;
    }
  }
  rec[0] = Rnet * Ctbaf;
  rec[GS] = dRnetdrho * Ctbaf + dCdrho * Rnet;
  rec[2 * GS] = dRnetdT * Ctbaf + dCdT * Rnet;
  double b = cR * Ctbaf;
  if (type != RT_SIMPLE)
  {
    for (int k = 0; k < nslots; ++k)
      rec[(5 + k) * GS] *= Ctbaf;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = coef * as_double(Ptb[2 * i + 1]);
      const int slot = (int)(signed char)((Ptb[2 * i] >> 24) & 255);
      if (slot >= 0)
        rec[(5 + slot) * GS] += e * Rnet;
      else
        b -= e * Rnet;
    }
    rec[3 * GS] = coef * base * Rnet;
  }
  else
    rec[3 * GS] = 0.;
  rec[4 * GS] = b;
#undef SP_IDX
#undef SP_ST
#undef SP_SLOT
#undef SP_INVMW
#undef SP_CONC
}

// 16-byte asynchronous global->shared copies (LDGSTS) of a [words] x 8-byte image, spread over the CTA
__device__ __forceinline__ void stage_async(void *dst_smem, const void *src_gmem, int bytes)
{
  const unsigned base = (unsigned)__cvta_generic_to_shared(dst_smem);
  const char *src = (const char *)src_gmem;
  for (int o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(base + o), "l"(src + o) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void stage_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__global__ void __launch_bounds__(512, 1) k_jac(const ChemArgs a)
{
  extern __shared__ __align__(16) double smem[];
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, G = a.G, GS = a.GS;
  const int nsm1 = ns - 1;
  const int ncolx = nsm1 + 5; // extended row: Y_0..Y_{ns-2}, w, dw/drho, dw/dT, A, B
  // staged images first (16-byte aligned), then the per-state working set
  unsigned long long *sprm = (unsigned long long *)smem;                    // [max_prm_words]
  unsigned long long *ssegs = sprm + dm.max_prm_words;                      // [max_segs]
  unsigned int *sitems = (unsigned int *)(ssegs + dm.max_segs);             // [max_items]
  double *sc = (double *)(sitems + dm.max_items);                           // [NSC][GS]
  double *sy = sc + Tile::NSC * GS;    // [ns][GS]
  double *sg = sy + ns * GS;           // Gibbs
  double *sdb = sg + ns * GS;          // dB/dT
  double *sh = sdb + ns * GS;          // enthalpies
  double *scp = sh + ns * GS;          // species cp
  double *sdcp = scp + ns * GS;        // species dcp/dT
  double *sprho = sdcp + ns * GS;      // [ns][GS] primitive-Jacobian rho column P[:,rho]
  double *strow = sprho + ns * GS;     // [ns+1][GS] T-row of the primitive Jacobian (cols rho, T, Y_k)
  double *srec = strow + (ns + 1) * GS; // [rec_cap][GS]
  double *sJ = srec + dm.rec_cap * GS; // [ncolx][ns][GS] extended rows of dw_i/d(Y_k | w,rho,T,A,B)
  double *srow = sJ + (size_t)nsm1 * ns * GS; // alias: [5][ns][GS] = columns ns-1..ns+3 of sJ

  const int ntiles = (a.n + G - 1) / G;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const int tile0 = tile * G;
    const int gcount = min(G, a.n - tile0);
    __syncthreads();
    // stage chunk 0's parameters and gather schedule while the tile is loaded
    stage_async(sprm, dm.cprm + dm.cprm_off[0], 8 * (dm.cprm_off[1] - dm.cprm_off[0]));
    stage_async(ssegs, dm.csegs + dm.cseg_off[0], 8 * (dm.cseg_off[1] - dm.cseg_off[0]));
    stage_async(sitems, dm.citems + dm.citem_off[0], 4 * (dm.citem_off[1] - dm.citem_off[0]));
    stage_commit();
    load_tile<true>(a, tile0, gcount, sc, sy);
    for (int e = threadIdx.x; e < ncolx * ns * GS; e += blockDim.x)
      sJ[e] = 0.;
    __syncthreads();
    if (threadIdx.x < gcount)
    {
      state_scalars(a, threadIdx.x, sc, sy);
      const int g = threadIdx.x;
      SM(sc, Tile::S_AUX7, g) = log(dm.p_ref * SM(sc, Tile::S_INVT, g) * (1. / dm.Ru)); // log(p0/(R T)), :535
    }
    __syncthreads();
    for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
    {
      const int i = item / gcount, g = item - i * gcount;
      const SpeciesThermo t =
          species_thermo<true>(dm, i, SM(sc, Tile::S_T, g), SM(sc, Tile::S_LOGT, g), SM(sc, Tile::S_INVT, g));
      SM(sg, i, g) = t.g;
      SM(sdb, i, g) = t.dB;
      SM(sh, i, g) = t.h;
      SM(scp, i, g) = t.cp;
      SM(sdcp, i, g) = t.dcp;
    }
    stage_wait_all();
    __syncthreads();
    if (threadIdx.x >= blockDim.x - gcount)
    { // cp and dcp/dT of the mixture in species order (thermodynamics_kernels.cpp:45-131, 183-260)
      const int g = blockDim.x - 1 - threadIdx.x;
      double cp = 0., dcp = 0.;
      for (int i = 0; i < ns; ++i)
      {
        cp += SM(sy, i, g) * SM(scp, i, g);
        if (dm.cptype[i] == CP_CONST)
          dcp = 0.; // sic, thermodynamics_kernels.cpp:202
        else
          dcp += SM(sy, i, g) * SM(sdcp, i, g);
      }
      SM(sc, Tile::S_CP, g) = cp;
      SM(sc, Tile::S_CPSENST, g) = dcp;
    }
    for (int c = 0; c < dm.n_chunks; ++c)
    {
      const int r0 = dm.chunk_rxn[c], nrc = dm.chunk_rxn[c + 1] - r0;
      // reaction phase: parameters from the staged blob
      {
        const unsigned int *offs = (const unsigned int *)sprm;
        for (int item = threadIdx.x; item < gcount * nrc; item += blockDim.x)
        {
          const int rr = item / gcount, g = item - rr * gcount;
          const unsigned long long *P = sprm + offs[rr];
          if (((int)(unsigned int)P[0]) & F_HAS_ORDERS)
            jac_reaction(dm, r0 + rr, g, GS, sc, sy, sg, sdb, srec + (size_t)dm.rec_off[r0 + rr] * GS + g);
          else
            jac_reaction_staged(dm, P, g, GS, sc, sy, sg, sdb, srec + g);
        }
      }
      stage_wait_all(); // this chunk's gather schedule (issued one phase ago) has landed
      __syncthreads();
      // the parameter buffer is free: prefetch the next chunk's parameters under the gather phase
      if (c + 1 < dm.n_chunks)
      {
        stage_async(sprm, dm.cprm + dm.cprm_off[c + 1], 8 * (dm.cprm_off[c + 2] - dm.cprm_off[c + 1]));
        stage_commit();
      }
      // gather phase: balanced segments of (row, column)-sorted items; every entry is summed in reaction order in a
      // register and added to its shared-memory slot once per chunk (rates_sensitivities_exact.cpp:1014-1026)
      {
        const int nseg = dm.cseg_off[c + 1] - dm.cseg_off[c];
        for (int item = threadIdx.x; item < gcount * nseg; item += blockDim.x)
        {
          const int sidx = item / gcount, g = item - sidx * gcount;
          const unsigned long long sd = ssegs[sidx];
          const int row = (int)(sd & 0xffff), cnt = (int)((sd >> 16) & 0xffff);
          if (cnt == 0)
            continue;
          const unsigned int *it = sitems + (unsigned int)(sd >> 32);
          const double negmw = -dm.mw[row];
          const double *recg = srec + g;
          unsigned int u = it[0];
          int col = (int)((u >> 16) & 0xfff);
          double acc = 0.;
          for (int k = 0; k < cnt; ++k)
          {
            u = it[k];
            const int ck = (int)((u >> 16) & 0xfff);
            if (ck != col)
            {
              SM(sJ, col * ns + row, g) += acc;
              acc = 0.;
              col = ck;
            }
            const int nu = ((int)u) >> 28; // sign-extended 4-bit stoichiometric coefficient
            acc += ((double)nu * negmw) * recg[(size_t)(u & 0xffff) * GS];
          }
          SM(sJ, col * ns + row, g) += acc;
        }
      }
      stage_wait_all(); // next chunk's parameters have landed
      __syncthreads();
      if (c + 1 < dm.n_chunks)
      { // the schedule buffers are free: prefetch the next chunk's schedule under its reaction phase
        stage_async(ssegs, dm.csegs + dm.cseg_off[c + 1], 8 * (dm.cseg_off[c + 2] - dm.cseg_off[c + 1]));
        stage_async(sitems, dm.citems + dm.citem_off[c + 1], 4 * (dm.citem_off[c + 2] - dm.citem_off[c + 1]));
        stage_commit();
      }
    }

    if (a.mode == MODE_SENS)
    {
      // raw (ns+1)x(ns+1) col-major sensitivities, rates_sensitivities_exact.cpp:68,1011-1025
      const int nsp1 = ns + 1;
      for (int g = 0; g < gcount; ++g)
      {
        double *out = a.out1 + (size_t)(tile0 + g) * nsp1 * nsp1;
        for (int e = threadIdx.x; e < nsp1 * nsp1; e += blockDim.x)
        {
          const int i = e % nsp1, col = e / nsp1;
          double v = 0.;
          if (i < ns)
          {
            if (col == 0)
              v = SM(srow, ns + i, g);
            else if (col == 1)
              v = SM(srow, 2 * ns + i, g);
            else if (col - 2 < nsm1)
            {
              const int k = col - 2;
              const double u = dm.invmw[k] - dm.invmw[nsm1];
              v = SM(sJ, k * ns + i, g) + (SM(srow, 3 * ns + i, g) * u + SM(srow, 4 * ns + i, g));
            }
          }
          out[e] = v;
        }
      }
      continue;
    }

    // ---- epilogue: chem_jac_isobaric (isobaric_reactor_kernels.cpp:58-98) + transform (:319-343) --------------
    // E1: per-state inner products with the species enthalpies, in species order
    if (threadIdx.x < 6 * gcount)
    {
      const int which = threadIdx.x / gcount, g = threadIdx.x - which * gcount;
      double d = 0.;
      if (which == 3)
      {
        for (int i = 0; i < ns; ++i)
          d += SM(scp, i, g) * SM(srow, i, g); // inner_product(w, cpi)
      }
      else
      {
        const int q = which < 3 ? which : which - 1; // 0:w 1:dw/drho 2:dw/dT 4->3:A 5->4:B
        for (int i = 0; i < ns; ++i)
          d += SM(sh, i, g) * SM(srow, q * ns + i, g);
      }
      SM(sc, Tile::S_AUX0 + which, g) = d; // AUX0: w.h  AUX1: wrho.h  AUX2: wT.h  AUX3: w.cpi  AUX4: A.h  AUX5: B.h
    }
    __syncthreads();
    // E2: rho column of the primitive Jacobian and the T-row (cols rho, T, Y_k); one thread per (g, col)
    for (int item = threadIdx.x; item < gcount * (ns + 1); item += blockDim.x)
    {
      const int col = item / gcount, g = item - col * gcount;
      const double rho = SM(sc, Tile::S_RHO, g), cp = SM(sc, Tile::S_CP, g);
      const double invRhoCp = 1. / (rho * cp), invRho = 1. / rho, invCp = 1. / cp;
      const double rhs0 = -SM(sc, Tile::S_AUX0, g) / (rho * cp);
      double v;
      if (col == 0)
        v = -invRhoCp * SM(sc, Tile::S_AUX1, g) - invRho * rhs0;
      else if (col == 1)
        v = -invRhoCp * (SM(sc, Tile::S_AUX2, g) + SM(sc, Tile::S_AUX3, g)) - rhs0 * SM(sc, Tile::S_CPSENST, g) * invCp;
      else if (col - 2 < nsm1)
      {
        const int k = col - 2;
        double d = 0.;
        for (int i = 0; i < ns; ++i)
          d += SM(sh, i, g) * SM(sJ, k * ns + i, g);
        const double u = dm.invmw[k] - dm.invmw[nsm1];
        d += SM(sc, Tile::S_AUX4, g) * u + SM(sc, Tile::S_AUX5, g);
        v = -invRhoCp * d - rhs0 * (SM(scp, k, g) - SM(scp, nsm1, g)) * invCp;
      }
      else
        v = 0.;
      SM(strow, col, g) = v;
      // species rows of the rho column: P[1+i, rho] = (dw_i/drho - w_i/rho)/rho, :75-78
      if (col >= 1 && col - 1 < nsm1)
      {
        const int i = col - 1;
        SM(sprho, 1 + i, g) = invRho * (SM(srow, ns + i, g) - invRho * SM(srow, i, g));
      }
      if (col == 0)
        SM(sc, Tile::S_AUX6, g) = rhs0;
    }
    __syncthreads();
    // open-reactor and heat-transfer terms (mass_jac_isobaric :100-140, heat_jac_isobaric :142-168, :289-309)
    // and the flamelet heat-loss terms (flamelet_kernels.cpp:1290-1320) modify the T-row / rhs only
    if (threadIdx.x < gcount)
    {
      const int g = threadIdx.x;
      const double rho = SM(sc, Tile::S_RHO, g), cp = SM(sc, Tile::S_CP, g), T = SM(sc, Tile::S_T, g);
      const double cpsensT = SM(sc, Tile::S_CPSENST, g);
      const double invCp = 1. / cp, invRhoCp = 1. / (rho * cp);
      double rhs0 = SM(sc, Tile::S_AUX6, g);
      if (a.mode == MODE_REACTOR_JAC)
      {
        if (a.rx.open)
        {
          const double Tin = a.rx.T_in, logTin = log(Tin), invTin = 1. / Tin;
          const double invTau = 1. / a.rx.tau;
          double m0;
          {
            const SpeciesThermo tl = species_thermo<false>(dm, ns - 1, Tin, logTin, invTin);
            m0 = (tl.h - SM(sh, ns - 1, g)) * a.rx.y_in[ns - 1];
          }
          for (int i = 0; i < nsm1; ++i)
          {
            const SpeciesThermo ti = species_thermo<false>(dm, i, Tin, logTin, invTin);
            m0 += (ti.h - SM(sh, i, g)) * a.rx.y_in[i];
          }
          m0 /= cp;
          m0 *= invTau;
          double ycp = 0.;
          for (int i = 0; i < ns; ++i)
            ycp += SM(scp, i, g) * a.rx.y_in[i]; // inner_product(inflowY, cpi)
          SM(strow, 1, g) += -invCp * (cpsensT * m0 + invTau * ycp);
          for (int k = 0; k < nsm1; ++k)
            SM(strow, 2 + k, g) += -m0 * (SM(scp, k, g) - SM(scp, nsm1, g)) * invCp;
          rhs0 += m0;
        }
        if (a.rx.heat_option == 1)
        {
          for (int k = 0; k < ns + 1; ++k)
            SM(strow, k, g) = 0.;
          rhs0 = 0.;
        }
        else if (a.rx.heat_option == 2)
        {
          const double Ts = a.rx.T_surf;
          const double rate = a.rx.SoV / (rho * cp) *
                              (a.rx.h_conv * (a.rx.T_inf - T) + a.rx.eps_rad * 5.67e-8 * (Ts * Ts * Ts * Ts - T * T * T * T));
          SM(strow, 0, g) += -rate / rho;
          SM(strow, 1, g) +=
              -invCp * cpsensT * rate - a.rx.SoV * invRhoCp * (a.rx.h_conv + 4. * a.rx.eps_rad * 5.67e-8 * T * T * T);
          const double cpn = SM(scp, nsm1, g);
          for (int k = 0; k < nsm1; ++k)
            SM(strow, 2 + k, g) += invCp * rate * (cpn - SM(scp, k, g));
          rhs0 += rate;
        }
      }
      else if (!a.fl.adiabatic)
      { // MODE_FLAMELET_JAC heat loss
        const FlameletDev &fl = a.fl;
        const int sidx = tile0 + g, F = sidx / fl.nzi, iz = sidx - F * fl.nzi;
        const size_t ho = (size_t)F * fl.stride_heat + iz;
        const double Tc = fl.T_conv[ho], Tr = fl.T_rad[ho], hc = fl.h_conv[ho], hr = fl.h_rad[ho];
        double q;
        if (fl.use_scaled_heat_loss)
        {
          const double maxT = fl.maxT[F];
          const double maxT4 = maxT * maxT * maxT * maxT;
          const double Tr4 = Tr * Tr * Tr * Tr;
          q = (hc * (Tc - T) / (maxT - Tc) + hr * 5.67e-8 * (Tr4 - T * T * T * T) / (maxT4 - Tr4)) * invRhoCp;
          SM(strow, 1, g) -=
              invCp * cpsensT * q + invRhoCp * (hc / (maxT - Tc) + 4. * hr / (maxT4 - Tr4) * 5.67e-8 * T * T * T);
        }
        else
        {
          q = (hc * (Tc - T) + hr * 5.67e-8 * (Tr * Tr * Tr * Tr - T * T * T * T)) * invRhoCp;
          SM(strow, 1, g) -= invCp * cpsensT * q + invRhoCp * (hc + 4. * hr * 5.67e-8 * T * T * T);
        }
        SM(strow, 0, g) -= q / rho;
        const double cpn = SM(scp, nsm1, g);
        for (int k = 0; k < nsm1; ++k)
          SM(strow, 2 + k, g) += invCp * q * (cpn - SM(scp, k, g));
      }
      SM(sc, Tile::S_AUX6, g) = rhs0;
      SM(sprho, 0, g) = SM(strow, 0, g);
    }
    __syncthreads();

    // E3: transform to the (T, Y) Jacobian and stream out, coalesced over the column-major ns x ns block
    if (a.mode == MODE_REACTOR_JAC)
    {
      for (int g = 0; g < gcount; ++g)
      {
        const double rho = SM(sc, Tile::S_RHO, g), T = SM(sc, Tile::S_T, g), mmw = SM(sc, Tile::S_MMW, g);
        const double invRho = 1. / rho, roT = rho / T, negRhoMmw = -rho * mmw;
        const bool open = a.rx.open != 0;
        const double invTau = open ? 1. / a.rx.tau : 0.;
        double *out = a.out1 + (size_t)(tile0 + g) * ns * ns;
        for (int e = threadIdx.x; e < ns * ns; e += blockDim.x)
        {
          const int row = e % ns, col = e / ns;
          const double prho = SM(sprho, row, g);
          double v;
          if (col == 0)
          {
            const double pT = (row == 0) ? SM(strow, 1, g) : SM(srow, 2 * ns + row - 1, g) * invRho;
            v = pT - roT * prho;
          }
          else
          {
            const int k = col - 1;
            const double u = dm.invmw[k] - dm.invmw[nsm1];
            double pY;
            if (row == 0)
              pY = SM(strow, 2 + k, g);
            else
            {
              const int i = row - 1;
              pY = invRho * (SM(sJ, k * ns + i, g) + (SM(srow, 3 * ns + i, g) * u + SM(srow, 4 * ns + i, g)));
              if (open && i == k)
                pY += -invTau;
            }
            v = pY + negRhoMmw * u * prho;
          }
          out[e] = v;
        }
      }
      // rhs
      for (int item = threadIdx.x; item < gcount * ns; item += blockDim.x)
      {
        const int g = item / ns, j = item - g * ns;
        double v;
        if (j == 0)
          v = SM(sc, Tile::S_AUX6, g);
        else
        {
          v = SM(srow, j - 1, g) * (1. / SM(sc, Tile::S_RHO, g));
          if (a.rx.open)
            v += (a.rx.y_in[j - 1] - SM(sy, j - 1, g)) * (1. / a.rx.tau);
        }
        a.out0[(size_t)(tile0 + g) * ns + j] = v;
      }
    }
    else
    { // MODE_FLAMELET_JAC: diagonal block of the BTDDOD matrix + cmajor, (T,T) enthalpy-flux correction,
      // optional prefactor*J - I (flamelet_kernels.cpp:1322-1408); off-diagonals are written by the host wrapper
      const FlameletDev &fl = a.fl;
      const int nzi = fl.nzi;
      const size_t jac_stride = (size_t)ns * ((size_t)nzi * ns + 2 * (nzi - 1));
      for (int g = 0; g < gcount; ++g)
      {
        const int sidx = tile0 + g, F = sidx / nzi, iz = sidx - F * nzi;
        const double rho = SM(sc, Tile::S_RHO, g), T = SM(sc, Tile::S_T, g), mmw = SM(sc, Tile::S_MMW, g);
        const double cp = SM(sc, Tile::S_CP, g);
        const double invRho = 1. / rho, roT = rho / T, negRhoMmw = -rho * mmw;
        double tt_corr = 0.;
        if (fl.include_enthalpy_flux)
        { // :1350-1381
          const double *st = a.in_state + (size_t)F * nzi * ns;
          const double *cpg = fl.cp_grid + (size_t)F * nzi;
          const double mc = fl.mcoeff[(size_t)F * fl.stride_mn + iz], nc = fl.ncoeff[(size_t)F * fl.stride_mn + iz];
          const double Tm = (iz == 0) ? fl.oxy[0] : st[(size_t)(iz - 1) * ns];
          const double Tp = (iz == nzi - 1) ? fl.fuel[0] : st[(size_t)(iz + 1) * ns];
          const double cpm = (iz == 0) ? fl.cp_bc[0] : cpg[iz - 1];
          const double cpp = (iz == nzi - 1) ? fl.cp_bc[1] : cpg[iz + 1];
          const double dTdZ = mc * Tm + nc * Tp;
          const double dcpdZ = mc * cpm + nc * cpp;
          const double f1 = 0.5 * fl.chi[(size_t)F * fl.stride_chi + iz] / cp * dTdZ * dcpdZ;
          tt_corr = f1 / cp * SM(sc, Tile::S_CPSENST, g);
        }
        const double *cmaj = fl.cmajor + (size_t)F * fl.stride_coeff + (size_t)iz * ns;
        double *out = a.out1 + (size_t)F * jac_stride + (size_t)iz * ns * ns;
        for (int e = threadIdx.x; e < ns * ns; e += blockDim.x)
        {
          const int row = e % ns, col = e / ns;
          const double prho = SM(sprho, row, g);
          double v;
          if (col == 0)
          {
            const double pT = (row == 0) ? SM(strow, 1, g) : SM(srow, 2 * ns + row - 1, g) * invRho;
            v = pT - roT * prho;
          }
          else
          {
            const int k = col - 1;
            const double u = dm.invmw[k] - dm.invmw[nsm1];
            double pY;
            if (row == 0)
              pY = SM(strow, 2 + k, g);
            else
            {
              const int i = row - 1;
              pY = invRho * (SM(sJ, k * ns + i, g) + (SM(srow, 3 * ns + i, g) * u + SM(srow, 4 * ns + i, g)));
            }
            v = pY + negRhoMmw * u * prho;
          }
          if (row == col)
          {
            v += cmaj[row];
            if (row == 0)
              v -= tt_corr;
          }
          if (fl.scale_and_offset)
          {
            v *= fl.prefactor;
            if (row == col)
              v -= 1.;
          }
          out[e] = v;
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
// cp of the two boundary streams (:1076-1086). One thread per grid point.
__global__ void k_flamelet_prepass(const DeviceMech dm, int F, int nzi, const double *state, const double *oxy,
                                   const double *fuel, double *cp_grid, double *maxT, double *cp_bc)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = dm.ns;
  auto cp_of = [&](const double *st) {
    const double t = st[0], logT = log(t), invT = 1. / t;
    double yl = 1.;
    for (int j = 0; j < ns - 1; ++j)
      yl -= st[1 + j];
    double cp = 0.;
    for (int i = 0; i < ns; ++i)
    {
      const double yi = (i < ns - 1) ? st[1 + i] : yl;
      cp += yi * species_thermo<false>(dm, i, t, logT, invT).cp;
    }
    return cp;
  };
  if (s < F * nzi)
    cp_grid[s] = cp_of(state + (size_t)s * ns);
  if (s == 0)
  {
    cp_bc[0] = cp_of(oxy);
    cp_bc[1] = cp_of(fuel);
  }
  if (s < F)
  {
    double m = -1;
    for (int i = 0; i < nzi; ++i)
      m = fmax(m, state[((size_t)s * nzi + i) * ns]);
    maxT[s] = m;
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
static size_t jac_smem(const DeviceMech &dm, int GS)
{
  const size_t ns = dm.ns;
  return 8 * (size_t)dm.max_prm_words + 8 * (size_t)dm.max_segs + 4 * (size_t)dm.max_items +
         sizeof(double) * (size_t)GS * (Tile::NSC + 6 * ns + ns + (ns + 1) + dm.rec_cap + (ns - 1 + 5) * ns);
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
  int G = env_int("GB_RATES_G", 31);
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
  const int grid = std::min(ntiles, sm_count());
  const int threads = env_int("GB_RATES_THREADS", 512);
  k_rates<<<grid, threads, sm, s>>>(a);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t launch_jac(const ChemArgs &a_in, cudaStream_t s)
{
  ChemArgs a = a_in;
  const int maxsm = 227 * 1024;
  int G = env_int("GB_JAC_G", 15);
  while (G > 1 && jac_smem(a.dm, G | 1) > (size_t)maxsm)
    G -= 2;
  if (jac_smem(a.dm, G | 1) > (size_t)maxsm)
    return cudaErrorInvalidConfiguration;
  G = std::min(G, std::max(1, a.n));
  a.G = G;
  a.GS = G | 1;
  const size_t sm = jac_smem(a.dm, a.GS);
  static bool attr = false;
  if (!attr)
  {
    cudaFuncSetAttribute(k_jac, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
    attr = true;
  }
  const int ntiles = (a.n + G - 1) / G;
  const int grid = std::min(ntiles, sm_count());
  const int threads = env_int("GB_JAC_THREADS", 512);
  k_jac<<<grid, threads, sm, s>>>(a);
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
  const int n = std::max(F * fl.nzi, 1);
  const int threads = 128;
  k_flamelet_prepass<<<(n + threads - 1) / threads, threads, 0, s>>>(dm, F, fl.nzi, state, fl.oxy, fl.fuel, cp_grid,
                                                                     maxT, cp_bc);
  ++g_launches;
  return cudaGetLastError();
}

} // namespace gb
