// gb_jac.cu -- k_jac: exact rate sensitivities, isobaric-reactor Jacobian and flamelet Jacobian blocks, FP64, sm_100a.
//
// Replaces prod_rates_sens_exact (rates_sensitivities_exact.cpp:33-1028), chem_jac_isobaric + mass/heat_jac_isobaric
// + transform_isobaric_primitive_jacobian (isobaric_reactor_kernels.cpp:58-168, 221-343) and the per-point part of
// flamelet_jacobian (flamelet_kernels.cpp:1254-1408) for batches of states.
//
// A CTA owns a tile of G states ("state blocking": every thread does its unit of work for all G states, so mechanism
// data is decoded once per tile and each thread carries G independent dependency chains). Phases, separated by
// __syncthreads():
//   load    : coalesced read of the G state vectors
//   thermo  : thread per species -> cp_i, h_i, dcp_i/dT, Gibbs, dB_i/dT; meanwhile one warp does the order-sensitive
//             per-state sums (Y_ns = 1 - sum, mixture weight) sequentially as the reference does
//   react   : thread per reaction -> record {q, dq/drho, dq/dT, a, b, H, dq/dY_slot...} in shared memory
//   gather  : thread per balanced range of the static plan stream (gb_plan.cu): every destination is summed in a
//             register in ascending reaction order and stored once -- no atomics, no read-modify-write
//   fix/row : split destinations are recombined; per-row and per-state quantities of chem_jac_isobaric are formed
//   output  : the ns x ns block is transformed to (T, Y) variables and streamed to HBM, lanes along the contiguous
//             (column-major) dimension, G fully coalesced stores per thread
// HBM traffic is the algorithmic minimum: ns doubles in, ns + ns^2 doubles out per state.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "gb_device.cuh"
#include "gb_kernels.cuh"

namespace gb
{

extern std::atomic<long> g_jac_launches;
std::atomic<long> g_jac_launches{0};

#define SMG(arr, idx, g) (arr)[(idx)*G + (g)]

__device__ __forceinline__ double u2d(unsigned long long u) { return __longlong_as_double((long long)u); }

// ------------------------------------------------------------------------------------------------------------------
// reaction phase: record of reaction P (packed parameters in global memory, gb_plan.cu) for state g.
// rates_sensitivities_exact.cpp:128-1009 restated per reaction; the `for s < ns-1` dense loops (:522-525, :807-810,
// :849-850, :859-862, ...) are carried by the two scalars a, b.
// ------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void reaction_record(const DeviceMech &dm, const unsigned long long *__restrict__ P, int g,
                                                const double *sc, const double *sy, const double *sg,
                                                const double *sdb, const double *sh, double *srec)
{
  const unsigned long long w0 = P[0], w1 = P[1];
  const int f = (int)(unsigned int)w0;
  double *rec = srec + (size_t)(unsigned int)(w0 >> 32) * G + g;
  const int nrc = (int)(w1 & 255), npd = (int)((w1 >> 8) & 255), nn = (int)((w1 >> 16) & 255),
            ntb = (int)((w1 >> 24) & 255), nslots = (int)((w1 >> 32) & 255);
  const int sum_stoich = (int)(signed char)((w1 >> 40) & 255), sum_rc = (int)((w1 >> 48) & 255),
            sum_pd = (int)((w1 >> 56) & 255);
  const int type = f_type(f);
  const unsigned long long *Prc = P + (type == RT_SIMPLE ? 5 : 13);
  const unsigned long long *Ppd = Prc + 2 * nrc;
  const unsigned long long *Pnet = Ppd + 2 * npd;
  const unsigned long long *Ptb = Pnet + 2 * nn;

  const int last = dm.ns - 1;
  const double T = SMG(sc, Tile::S_T, g), invT = SMG(sc, Tile::S_INVT, g), logT = SMG(sc, Tile::S_LOGT, g);
  const double rho = SMG(sc, Tile::S_RHO, g);
  const double invM = 1. / SMG(sc, Tile::S_MMW, g);
  const double ct = rho * invM;
  const double invRu = 1. / dm.Ru;
  for (int k = 0; k < nslots; ++k)
    rec[(JP_REC_HDR + k) * G] = 0.;

  const double kfb = u2d(P[3]), kfE = u2d(P[4]);
  const double kf = rate_constant(f_kform(f), u2d(P[2]), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT); // ARRHENIUS_SENS_OVER_K, :25
  double cR = 0.; // dense offset of dRnet/dY_s: -(last species as reactant) + (last species as product)

#define SP_IDX(Q, i) ((int)(Q[2 * (i)] & 0xffff))
#define SP_ST(Q, i) ((int)((Q[2 * (i)] >> 16) & 255))
#define SP_SLOT(Q, i) ((int)(signed char)((Q[2 * (i)] >> 24) & 255))
#define SP_INVMW(Q, i) u2d(Q[2 * (i) + 1])
#define SP_CONC(Q, i) (SMG(sy, SP_IDX(Q, i), g) * rho * SP_INVMW(Q, i))

  // v * prod_{i != skip} C_i^nu_i; seq reproduces the reference's special-cased orders ((v*C)*C), otherwise its
  // generic branch (v*(C*C)); use_pow handles |nu| > 3 where the reference does (:702-704)
  auto mult = [&](double v, const unsigned long long *Q, int n, int skip, bool seq, bool use_pow) {
    for (int i = 0; i < n; ++i)
    {
      if (i == skip)
        continue;
      const double c = SP_CONC(Q, i);
      const int nu = SP_ST(Q, i);
      if (nu == 1)
        v *= c;
      else if (seq)
      {
        for (int k = 0; k < nu; ++k)
          v *= c;
      }
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
  double Rnet = mult(kf, Prc, nrc, -1, fseq, false); // :287-325
  double dRnetdrho = Rnet / ct * invM * sum_rc;
  double dRnetdT = Rnet * kf_sens;
  for (int i = 0; i < nrc; ++i)
  { // :332-526
    const double d = deriv(kf * rho * SP_INVMW(Prc, i), Prc, nrc, i, fseq, false);
    if (SP_IDX(Prc, i) == last)
      cR -= d;
    else
      rec[(JP_REC_HDR + SP_SLOT(Prc, i)) * G] = d;
  }
  if (f & F_REVERSIBLE)
  { // :528-812
    double gs, ds;
    {
      const int i0 = (int)(Pnet[0] & 0xffff), s0 = (int)(signed char)((Pnet[0] >> 16) & 255);
      gs = s0 * SMG(sg, i0, g);
      ds = s0 * SMG(sdb, i0, g);
    }
    for (int i = 1; i < nn; ++i)
    {
      const int ii = (int)(Pnet[2 * i] & 0xffff), si = (int)(signed char)((Pnet[2 * i] >> 16) & 255);
      gs = gs + si * SMG(sg, ii, g);
      ds = ds + si * SMG(sdb, ii, g);
    }
    const double Kc = exp(-(sum_stoich * SMG(sc, Tile::S_AUX7, g) - invT * invRu * (gs)));
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
        rec[(JP_REC_HDR + SP_SLOT(Ppd, i)) * G] -= d;
    }
  }

  // third-body / falloff factor C_tbaf and its sensitivities, :826-1000
  double Ctbaf = 1., dCdrho = 0., dCdT = 0., coef = 0.; // dCtbaf/dY_s = coef*(base*u_s + eps_s - eps_last)
  const double base = (type != RT_SIMPLE) ? u2d(P[5]) : 0.;
  if (type != RT_SIMPLE)
  {
    double M = base * ct;
    double dMdrho = base * invM;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = u2d(Ptb[2 * i + 1]) * SMG(sy, (int)(Ptb[2 * i] & 0xffff), g);
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
      const double kpb = u2d(P[7]), kpE = u2d(P[8]);
      const double kp_over_kf = u2d(P[6]) * exp(kpb * logT - kpE * invT) / kf;
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
        const double tr0 = u2d(P[9]), tr1 = u2d(P[10]), tr2 = u2d(P[11]), tr3 = u2d(P[12]);
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

  rec[0] = Rnet * Ctbaf;                              // q, :1002
  rec[G] = dRnetdrho * Ctbaf + dCdrho * Rnet;         // dq/drho
  rec[2 * G] = dRnetdT * Ctbaf + dCdT * Rnet;         // dq/dT
  double b = cR * Ctbaf;
  if (type != RT_SIMPLE)
  {
    for (int k = 0; k < nslots; ++k)
      rec[(JP_REC_HDR + k) * G] *= Ctbaf;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = coef * u2d(Ptb[2 * i + 1]);
      const int slot = (int)(signed char)((Ptb[2 * i] >> 24) & 255);
      if (slot >= 0)
        rec[(JP_REC_HDR + slot) * G] += e * Rnet;
      else
        b -= e * Rnet; // the last species is a third body: -eps_last on every column (:855-865)
    }
    rec[3 * G] = coef * base * Rnet;
  }
  else
    rec[3 * G] = 0.;
  rec[4 * G] = b;
  // reaction enthalpy H = sum_net h_i * (-nu_i M_i): sum_i h_i dw_i/dx = sum_r H_r dq_r/dx (isobaric_reactor_kernels.cpp:74-92)
  double H = 0.;
  for (int i = 0; i < nn; ++i)
    H += SMG(sh, (int)(Pnet[2 * i] & 0xffff), g) * u2d(Pnet[2 * i + 1]);
  rec[5 * G] = H;
#undef SP_IDX
#undef SP_ST
#undef SP_SLOT
#undef SP_INVMW
#undef SP_CONC
}

// reactions with non-elementary orders (rare): global-memory parameter path, rates_sensitivities_exact.cpp:198-281
template <int G>
__device__ void reaction_record_orders(const DeviceMech &dm, int r, const unsigned long long *__restrict__ P, int g,
                                       const double *sc, const double *sy, const double *sh, double *srec)
{
  const unsigned long long w0 = P[0], w1 = P[1];
  const int f = (int)(unsigned int)w0;
  double *rec = srec + (size_t)(unsigned int)(w0 >> 32) * G + g;
  const int nrc = (int)(w1 & 255), npd = (int)((w1 >> 8) & 255), nn = (int)((w1 >> 16) & 255),
            nslots = (int)((w1 >> 32) & 255);
  const int type = f_type(f);
  const unsigned long long *Pnet = P + (type == RT_SIMPLE ? 5 : 13) + 2 * nrc + 2 * npd;
  const int last = dm.ns - 1;
  const double T = SMG(sc, Tile::S_T, g), invT = SMG(sc, Tile::S_INVT, g), logT = SMG(sc, Tile::S_LOGT, g);
  const double rho = SMG(sc, Tile::S_RHO, g);
  const double invM = 1. / SMG(sc, Tile::S_MMW, g);
  const double ct = rho * invM;
  for (int k = 0; k < nslots; ++k)
    rec[(JP_REC_HDR + k) * G] = 0.;
  const double kfb = u2d(P[3]), kfE = u2d(P[4]);
  const double kf = rate_constant(f_kform(f), u2d(P[2]), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT);
  const int n = dm.n_sp[r];
  const short *sp = dm.sp_idx + NSR * (size_t)r;
  const double *ord = dm.sp_order + NSR * (size_t)r;
  const signed char *spslot = dm.sp_slot + NSR * (size_t)r;
#define CS(i) (SMG(sy, sp[i], g) * rho * dm.invmw[sp[i]])
  double sumOrders = 0., Rnet = kf, cR = 0.;
  for (int i = 0; i < n; ++i)
    if (fabs(ord[i]) > 1.e-12)
    {
      Rnet *= pow(fmax(CS(i), 0.), ord[i]);
      sumOrders += ord[i];
    }
  const double dRnetdrho = Rnet / ct * invM * sumOrders;
  const double dRnetdT = Rnet * kf_sens;
  for (int j = 0; j < n; ++j)
  {
    if (!(fabs(ord[j]) > 1.e-12))
      continue;
    const bool is_last = sp[j] == last;
    double v = kf;
    for (int l = 0; l < n; ++l)
    {
      const double cl = CS(l);
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
      rec[(JP_REC_HDR + spslot[j]) * G] = v;
  }
#undef CS
  // third-body factors of non-elementary reactions: only the plain third-body form is supported here
  double Ctbaf = 1., dCdrho = 0., coef = 0., b = 0.;
  const double base = (type != RT_SIMPLE) ? u2d(P[5]) : 0.;
  const int ntb = (int)((w1 >> 24) & 255);
  const unsigned long long *Ptb = Pnet + 2 * nn;
  if (type == RT_THIRD_BODY)
  {
    double M = base * ct, dMdrho = base * invM;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = u2d(Ptb[2 * i + 1]) * SMG(sy, (int)(Ptb[2 * i] & 0xffff), g);
      M = M + rho * e;
      dMdrho += e;
    }
    Ctbaf = M;
    dCdrho = dMdrho;
    coef = rho;
  }
  else if (type != RT_SIMPLE)
    Ctbaf = __longlong_as_double(0x7ff8000000000000LL); // falloff with non-elementary orders: not supported (NaN)
  rec[0] = Rnet * Ctbaf;
  rec[G] = dRnetdrho * Ctbaf + dCdrho * Rnet;
  rec[2 * G] = dRnetdT * Ctbaf;
  b = cR * Ctbaf;
  if (type == RT_THIRD_BODY)
  {
    for (int k = 0; k < nslots; ++k)
      rec[(JP_REC_HDR + k) * G] *= Ctbaf;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = coef * u2d(Ptb[2 * i + 1]);
      const int slot = (int)(signed char)((Ptb[2 * i] >> 24) & 255);
      if (slot >= 0)
        rec[(JP_REC_HDR + slot) * G] += e * Rnet;
      else
        b -= e * Rnet;
    }
    rec[3 * G] = coef * base * Rnet;
  }
  else
    rec[3 * G] = 0.;
  rec[4 * G] = b;
  double H = 0.;
  for (int i = 0; i < nn; ++i)
    H += SMG(sh, (int)(Pnet[2 * i] & 0xffff), g) * u2d(Pnet[2 * i + 1]);
  rec[5 * G] = H;
}

// ------------------------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(512, 1) k_jac(const ChemArgs a)
{
  extern __shared__ __align__(16) double smem[];
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, nr = dm.nr, nsm1 = ns - 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *sc = smem;                    // [NSC][G] per-state scalars
  double *sy = sc + Tile::NSC * G;      // [ns][G] mass fractions
  double *sg = sy + ns * G;             // Gibbs              -> later: row quantity P[1+i, rho]
  double *sdb = sg + ns * G;            // dB/dT              -> later: row quantity nm_i * RA_i
  double *sh = sdb + ns * G;            // enthalpies
  double *scp = sh + ns * G;            // species cp
  double *sdcp = scp + ns * G;          // species dcp/dT     -> later: row quantity nm_i * RB_i
  double *su = sdcp + ns * G;           // [ns] u_k = 1/M_k - 1/M_ns
  double *snm = su + ns;                // [ns] -M_i
  double *srec = snm + ns;              // [rec_total][G] reaction records
  double *sJ = srec + (size_t)dm.jp_rec_total * G; // [nslots][G] gathered sums
  unsigned short *semap = (unsigned short *)(sJ + (size_t)dm.jp_nslots * G); // [ns*(ns-1)]

  // per-CTA constants
  for (int i = tid; i < ns; i += nt)
  {
    su[i] = dm.invmw[i] - dm.invmw[nsm1];
    snm[i] = -dm.mw[i];
  }
  for (int e = tid; e < ns * nsm1; e += nt)
    semap[e] = dm.jp_emap[e];

  const int ntiles = (a.n + G - 1) / G;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const int tile0 = tile * G;
    const int gcount = min(G, a.n - tile0);
    __syncthreads();
    // ---- load (states past the end of the batch replicate the tile's first state; they are never written) ----------
    if (a.in_state != nullptr)
    {
      for (int item = tid; item < G * ns; item += nt)
      {
        const int g = item / ns, j = item - g * ns;
        const double v = a.in_state[(size_t)(tile0 + (g < gcount ? g : 0)) * ns + j];
        if (j == 0)
          SMG(sc, Tile::S_T, g) = v;
        else
          SMG(sy, j - 1, g) = v;
      }
    }
    else
    {
      for (int item = tid; item < G * ns; item += nt)
      {
        const int g = item / ns, j = item - g * ns;
        SMG(sy, j, g) = a.in_y[(size_t)(tile0 + (g < gcount ? g : 0)) * ns + j];
      }
      if (tid < G)
      {
        SMG(sc, Tile::S_T, tid) = a.in_T[tile0 + (tid < gcount ? tid : 0)];
        SMG(sc, Tile::S_RHO, tid) = a.in_rho[tile0 + (tid < gcount ? tid : 0)];
      }
    }
    __syncthreads();
    // ---- thermo (threads 32..) overlapped with the order-sensitive per-state sums (warp 0) ------------------------------
    if (tid < G)
    { // extract_y (combustion_kernels.h:505-515), mixture_molecular_weight (:381-387), ideal_gas_density (:526-530)
      const int g = tid;
      if (a.in_state != nullptr)
      {
        double yl = 1.;
        for (int j = 0; j < nsm1; ++j)
          yl -= SMG(sy, j, g);
        SMG(sy, nsm1, g) = yl;
      }
      double d = 0.;
      for (int i = 0; i < ns; ++i)
        d += dm.invmw[i] * SMG(sy, i, g);
      const double mmw = 1. / d, T = SMG(sc, Tile::S_T, g), invT = 1. / T;
      SMG(sc, Tile::S_MMW, g) = mmw;
      SMG(sc, Tile::S_LOGT, g) = log(T);
      SMG(sc, Tile::S_INVT, g) = invT;
      if (a.in_state != nullptr)
        SMG(sc, Tile::S_RHO, g) = a.p * mmw / (T * dm.Ru);
      SMG(sc, Tile::S_AUX7, g) = log(dm.p_ref * invT * (1. / dm.Ru)); // log(p0/(R T)), :535
    }
    else if (tid >= 32)
    {
      for (int i = tid - 32; i < ns; i += nt - 32)
      {
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
          const double T = SMG(sc, Tile::S_T, g);
          const SpeciesThermo t = species_thermo<true>(dm, i, T, log(T), 1. / T);
          SMG(sg, i, g) = t.g;
          SMG(sdb, i, g) = t.dB;
          SMG(sh, i, g) = t.h;
          SMG(scp, i, g) = t.cp;
          SMG(sdcp, i, g) = t.dcp;
        }
      }
    }
    __syncthreads();
    // ---- reaction phase (the last G threads first form cp and dcp/dT of the mixture, species order) ----------------------
    if (tid >= nt - G)
    { // thermodynamics_kernels.cpp:45-131, 183-260
      const int g = nt - 1 - tid;
      double cp = 0., dcp = 0.;
      for (int i = 0; i < ns; ++i)
      {
        cp += SMG(sy, i, g) * SMG(scp, i, g);
        if (dm.cptype[i] == CP_CONST)
          dcp = 0.; // sic, thermodynamics_kernels.cpp:202
        else
          dcp += SMG(sy, i, g) * SMG(sdcp, i, g);
      }
      SMG(sc, Tile::S_CP, g) = cp;
      SMG(sc, Tile::S_CPSENST, g) = dcp;
    }
    for (int r = tid; r < nr; r += nt)
    {
      const unsigned long long *P = dm.jp_prm + dm.jp_prm_off[r];
      if (((int)(unsigned int)P[0]) & F_HAS_ORDERS)
      {
        for (int g = 0; g < G; ++g)
          reaction_record_orders<G>(dm, r, P, g, sc, sy, sh, srec);
      }
      else
      {
#pragma unroll 1
        for (int g = 0; g < G; ++g)
          reaction_record<G>(dm, P, g, sc, sy, sg, sdb, sh, srec);
      }
    }
    __syncthreads();
    // ---- gather phase: this thread's range of the plan stream -----------------------------------------------------------------
    {
      const unsigned int *__restrict__ st = dm.jp_stream;
      int w = dm.jp_tstart[tid];
      const int wend = dm.jp_tstart[tid + 1];
      while (w < wend)
      {
        const unsigned int hd = st[w++];
        const int slot = (int)(hd & 0xfffff), cnt = (int)((hd >> 20) & 0x7ff);
        double acc[G];
#pragma unroll
        for (int g = 0; g < G; ++g)
          acc[g] = 0.;
        if (hd >> 31)
        { // product items: sum_r H_r * value_r
          for (int k = 0; k < cnt; ++k)
          {
            const unsigned int u = st[w + k];
            const double *pa = srec + (size_t)(u & 0xffff) * G, *pb = srec + (size_t)(u >> 16) * G;
#pragma unroll
            for (int g = 0; g < G; ++g)
              acc[g] += pa[g] * pb[g];
          }
        }
        else
        { // plain items: sum_r nu_r * value_r
          for (int k = 0; k < cnt; ++k)
          {
            const unsigned int u = st[w + k];
            const double nu = (double)(((int)u) >> 24);
            const double *pv = srec + (size_t)(u & 0xffff) * G;
#pragma unroll
            for (int g = 0; g < G; ++g)
              acc[g] += nu * pv[g];
          }
        }
        w += cnt;
#pragma unroll
        for (int g = 0; g < G; ++g)
          SMG(sJ, slot, g) = acc[g];
      }
    }
    __syncthreads();
    // ---- recombine split destinations in part order -------------------------------------------------------------------------------
    for (int item = tid; item < dm.jp_nfix * G; item += nt)
    {
      const int fi = item / G, g = item - fi * G;
      const int dst = dm.jp_fix[3 * fi], first = dm.jp_fix[3 * fi + 1], np = dm.jp_fix[3 * fi + 2];
      double v = SMG(sJ, dst, g);
      for (int p = 0; p < np; ++p)
        v += SMG(sJ, first + p, g);
      SMG(sJ, dst, g) = v;
    }
    __syncthreads();
    const double *R_w = sJ + (size_t)dm.jp_rbase * G; // [5][ns][G]: sums of nu*value for w, dw/drho, dw/dT, A, B
    const double *TH = sJ + (size_t)dm.jp_tbase * G;  // [ns+1][G]
    const double *SS = sJ + (size_t)dm.jp_sbase * G;  // [3][G]: w.h, A.h, B.h

    if (a.mode == MODE_SENS)
    { // raw (ns+1)x(ns+1) column-major sensitivities, rates_sensitivities_exact.cpp:68, 1011-1025
      const int nsp1 = ns + 1;
      int row = tid % nsp1, col = tid / nsp1;
      const int drow = nt % nsp1, dcol = nt / nsp1;
      for (int e = tid; e < nsp1 * nsp1; e += nt)
      {
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
          double v = 0.;
          if (row < ns)
          {
            const double nm = snm[row];
            if (col == 0)
              v = nm * SMG(R_w, ns + row, g);
            else if (col == 1)
              v = nm * SMG(R_w, 2 * ns + row, g);
            else if (col - 2 < nsm1)
            {
              const int k = col - 2;
              const unsigned short s = semap[k * ns + row];
              const double rv = (s == 0xffff) ? 0. : SMG(sJ, s, g);
              v = nm * (rv + (SMG(R_w, 3 * ns + row, g) * su[k] + SMG(R_w, 4 * ns + row, g)));
            }
          }
          if (g < gcount)
            a.out1[(size_t)(tile0 + g) * nsp1 * nsp1 + e] = v;
        }
        row += drow;
        col += dcol;
        if (row >= nsp1)
        {
          row -= nsp1;
          ++col;
        }
      }
      continue;
    }

    // ---- per-row quantities (threads 32..) and per-state quantities (threads < G) of chem_jac_isobaric (:58-98) -------
    if (tid < G)
    {
      const int g = tid;
      const double rho = SMG(sc, Tile::S_RHO, g), cp = SMG(sc, Tile::S_CP, g), T = SMG(sc, Tile::S_T, g);
      const double cpsensT = SMG(sc, Tile::S_CPSENST, g);
      const double invRhoCp = 1. / (rho * cp), invRho = 1. / rho, invCp = 1. / cp;
      double wcp = 0.; // inner_product(w, cpi)
      for (int i = 0; i < ns; ++i)
        wcp += SMG(scp, i, g) * (snm[i] * SMG(R_w, i, g));
      const double rhs0c = -SMG(SS, 0, g) / (rho * cp);
      double rhs0 = rhs0c;
      double P0rho = -invRhoCp * SMG(TH, 0, g) - invRho * rhs0c;
      double P0T = -invRhoCp * (SMG(TH, 1, g) + wcp) - rhs0c * cpsensT * invCp;
      double cextra = 0.; // extra coefficient of (cp_k - cp_ns) in the T-row of the Y_k columns
      if (a.mode == MODE_REACTOR_JAC)
      { // mass_jac_isobaric :100-140, heat_jac_isobaric :142-168, :289-309
        if (a.rx.open)
        {
          const double Tin = a.rx.T_in, logTin = log(Tin), invTin = 1. / Tin;
          const double invTau = 1. / a.rx.tau;
          double m0;
          {
            const SpeciesThermo tl = species_thermo<false>(dm, nsm1, Tin, logTin, invTin);
            m0 = (tl.h - SMG(sh, nsm1, g)) * a.rx.y_in[nsm1];
          }
          for (int i = 0; i < nsm1; ++i)
          {
            const SpeciesThermo ti = species_thermo<false>(dm, i, Tin, logTin, invTin);
            m0 += (ti.h - SMG(sh, i, g)) * a.rx.y_in[i];
          }
          m0 /= cp;
          m0 *= invTau;
          double ycp = 0.;
          for (int i = 0; i < ns; ++i)
            ycp += SMG(scp, i, g) * a.rx.y_in[i];
          P0T += -invCp * (cpsensT * m0 + invTau * ycp);
          cextra += -m0 * invCp;
          rhs0 += m0;
        }
        if (a.rx.heat_option == 2)
        {
          const double Ts = a.rx.T_surf;
          const double rate = a.rx.SoV / (rho * cp) *
                              (a.rx.h_conv * (a.rx.T_inf - T) + a.rx.eps_rad * 5.67e-8 * (Ts * Ts * Ts * Ts - T * T * T * T));
          P0rho += -rate / rho;
          P0T += -invCp * cpsensT * rate - a.rx.SoV * invRhoCp * (a.rx.h_conv + 4. * a.rx.eps_rad * 5.67e-8 * T * T * T);
          cextra += -invCp * rate;
          rhs0 += rate;
        }
      }
      else if (a.mode == MODE_FLAMELET_JAC && !a.fl.adiabatic)
      { // flamelet_kernels.cpp:1290-1320
        const FlameletDev &fl = a.fl;
        const int sidx = tile0 + (g < gcount ? g : 0), F = sidx / fl.nzi, iz = sidx - F * fl.nzi;
        const size_t ho = (size_t)F * fl.stride_heat + iz;
        const double Tc = fl.T_conv[ho], Tr = fl.T_rad[ho], hc = fl.h_conv[ho], hr = fl.h_rad[ho];
        double q;
        if (fl.use_scaled_heat_loss)
        {
          const double maxT = fl.maxT[F];
          const double maxT4 = maxT * maxT * maxT * maxT;
          const double Tr4 = Tr * Tr * Tr * Tr;
          q = (hc * (Tc - T) / (maxT - Tc) + hr * 5.67e-8 * (Tr4 - T * T * T * T) / (maxT4 - Tr4)) * invRhoCp;
          P0T -= invCp * cpsensT * q + invRhoCp * (hc / (maxT - Tc) + 4. * hr / (maxT4 - Tr4) * 5.67e-8 * T * T * T);
        }
        else
        {
          q = (hc * (Tc - T) + hr * 5.67e-8 * (Tr * Tr * Tr * Tr - T * T * T * T)) * invRhoCp;
          P0T -= invCp * cpsensT * q + invRhoCp * (hc + 4. * hr * 5.67e-8 * T * T * T);
        }
        P0rho -= q / rho;
        cextra += -invCp * q;
      }
      SMG(sc, Tile::S_AUX0, g) = rhs0;
      SMG(sc, Tile::S_AUX1, g) = P0rho;
      SMG(sc, Tile::S_AUX2, g) = P0T;
      SMG(sc, Tile::S_AUX3, g) = -rhs0c * invCp + cextra; // coefficient of (cp_k - cp_ns) in P[0, Y_k]
    }
    else if (tid >= 32)
    {
      for (int i = tid - 32; i < ns; i += nt - 32)
      {
        const double nm = snm[i];
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
          const double invRho = 1. / SMG(sc, Tile::S_RHO, g);
          const double w = nm * SMG(R_w, i, g), wr = nm * SMG(R_w, ns + i, g);
          SMG(sg, i, g) = invRho * (wr - invRho * w);        // P[1+i, rho], :75-78
          SMG(sdb, i, g) = nm * SMG(R_w, 3 * ns + i, g);     // -M_i * RA_i
          SMG(sdcp, i, g) = nm * SMG(R_w, 4 * ns + i, g);    // -M_i * RB_i
        }
      }
    }
    __syncthreads();

    // ---- output: transform (:319-343) and stream out ---------------------------------------------------------------------------------
    const bool reactor = a.mode == MODE_REACTOR_JAC;
    const bool isothermal = reactor && a.rx.heat_option == 1;
    const bool open = reactor && a.rx.open != 0;
    const double invTau = open ? 1. / a.rx.tau : 0.;
    double invRho_[G], roT_[G], nRM_[G];
#pragma unroll
    for (int g = 0; g < G; ++g)
    {
      const double rho = SMG(sc, Tile::S_RHO, g);
      invRho_[g] = 1. / rho;
      roT_[g] = rho / SMG(sc, Tile::S_T, g);
      nRM_[g] = -rho * SMG(sc, Tile::S_MMW, g);
    }
    // destination of state g's block, and the flamelet extras of its grid point
    const FlameletDev &fl = a.fl;
    size_t obase[G];
    double ttc[G];
    const double *cmaj[G];
#pragma unroll
    for (int g = 0; g < G; ++g)
    {
      const int sidx = tile0 + (g < gcount ? g : 0);
      ttc[g] = 0.;
      cmaj[g] = nullptr;
      if (reactor)
        obase[g] = (size_t)sidx * ns * ns;
      else
      { // block iz of flamelet F in BTDDOD storage
        const int nzi = fl.nzi, F = sidx / nzi, iz = sidx - F * nzi;
        obase[g] = (size_t)F * ((size_t)ns * ((size_t)nzi * ns + 2 * (nzi - 1))) + (size_t)iz * ns * ns;
        cmaj[g] = fl.cmajor + (size_t)F * fl.stride_coeff + (size_t)iz * ns;
        if (fl.include_enthalpy_flux)
        { // (T,T) correction, flamelet_kernels.cpp:1350-1381
          const double *stt = a.in_state + (size_t)F * nzi * ns;
          const double *cpg = fl.cp_grid + (size_t)F * nzi;
          const double mc = fl.mcoeff[(size_t)F * fl.stride_mn + iz], nc = fl.ncoeff[(size_t)F * fl.stride_mn + iz];
          const double Tm = (iz == 0) ? fl.oxy[0] : stt[(size_t)(iz - 1) * ns];
          const double Tp = (iz == nzi - 1) ? fl.fuel[0] : stt[(size_t)(iz + 1) * ns];
          const double cpm = (iz == 0) ? fl.cp_bc[0] : cpg[iz - 1];
          const double cpp = (iz == nzi - 1) ? fl.cp_bc[1] : cpg[iz + 1];
          const double cp = SMG(sc, Tile::S_CP, g);
          const double dTdZ = mc * Tm + nc * Tp, dcpdZ = mc * cpm + nc * cpp;
          const double f1 = 0.5 * fl.chi[(size_t)F * fl.stride_chi + iz] / cp * dTdZ * dcpdZ;
          ttc[g] = f1 / cp * SMG(sc, Tile::S_CPSENST, g);
        }
      }
    }
    {
      int row = tid % ns, col = tid / ns;
      const int drow = nt % ns, dcol = nt / ns;
      for (int e = tid; e < ns * ns; e += nt)
      {
        const double uk = col > 0 ? su[col - 1] : 0.;
        unsigned short s = 0xffff;
        if (row > 0 && col > 0)
          s = semap[(col - 1) * ns + (row - 1)];
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
          double v;
          if (row == 0)
          { // temperature row
            const double P0rho = SMG(sc, Tile::S_AUX1, g);
            if (isothermal)
              v = 0.;
            else if (col == 0)
              v = SMG(sc, Tile::S_AUX2, g) - roT_[g] * P0rho;
            else
            {
              const int k = col - 1;
              const double cp = SMG(sc, Tile::S_CP, g);
              const double sum = SMG(TH, 2 + k, g) + uk * SMG(SS, 1, g) + SMG(SS, 2, g);
              const double pY = -sum / (SMG(sc, Tile::S_RHO, g) * cp) +
                                SMG(sc, Tile::S_AUX3, g) * (SMG(scp, k, g) - SMG(scp, nsm1, g));
              v = pY + nRM_[g] * uk * P0rho;
            }
          }
          else
          {
            const int i = row - 1;
            const double prho = SMG(sg, i, g);
            if (col == 0)
              v = snm[i] * SMG(R_w, 2 * ns + i, g) * invRho_[g] - roT_[g] * prho;
            else
            {
              const double rv = (s == 0xffff) ? 0. : snm[i] * SMG(sJ, s, g);
              double pY = invRho_[g] * (rv + (SMG(sdb, i, g) * uk + SMG(sdcp, i, g)));
              if (open && row == col)
                pY += -invTau;
              v = pY + nRM_[g] * uk * prho;
            }
          }
          if (!reactor)
          {
            if (row == col)
            {
              v += cmaj[g][row];
              if (row == 0)
                v -= ttc[g];
            }
            if (fl.scale_and_offset)
            {
              v *= fl.prefactor;
              if (row == col)
                v -= 1.;
            }
          }
          if (g < gcount)
            a.out1[obase[g] + e] = v;
        }
        row += drow;
        col += dcol;
        if (row >= ns)
        {
          row -= ns;
          ++col;
        }
      }
    }
    if (reactor)
    { // right-hand side, chem_rhs_isobaric :19-29 (+ :194-218)
      for (int j = tid; j < ns; j += nt)
      {
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
          double v;
          if (j == 0)
            v = isothermal ? 0. : SMG(sc, Tile::S_AUX0, g);
          else
          {
            v = snm[j - 1] * SMG(R_w, j - 1, g) * invRho_[g];
            if (open)
              v += (a.rx.y_in[j - 1] - SMG(sy, j - 1, g)) * invTau;
          }
          if (g < gcount)
            a.out0[(size_t)(tile0 + g) * ns + j] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
static size_t jac_smem_bytes(const DeviceMech &dm, int G)
{
  const size_t ns = dm.ns;
  size_t doubles = (size_t)G * (Tile::NSC + 6 * ns + dm.jp_rec_total + dm.jp_nslots) + 2 * ns;
  return doubles * sizeof(double) + sizeof(unsigned short) * ns * (ns - 1) + 16;
}

static int jac_sm_count()
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

template <int G>
static cudaError_t launch_jac_g(const ChemArgs &a, size_t smem, cudaStream_t s)
{
  static bool attr = false;
  if (!attr)
  {
    cudaError_t e = cudaFuncSetAttribute(k_jac<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess)
      return e;
    attr = true;
  }
  const int threads = a.dm.jp_threads;
  const int ntiles = (a.n + G - 1) / G;
  // CTAs per SM: limited by shared memory and by 64K registers / (threads * ~128)
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>((size_t)(227 * 1024) / (smem + 1024), (size_t)(512 / threads)));
  per_sm = std::min(per_sm, 8);
  const int grid = std::min(ntiles, jac_sm_count() * per_sm);
  k_jac<G><<<grid, threads, smem, s>>>(a);
  ++g_jac_launches;
  return cudaGetLastError();
}

cudaError_t launch_jac(const ChemArgs &a_in, cudaStream_t s)
{
  ChemArgs a = a_in;
  const size_t maxsm = 227 * 1024;
  int G = 4;
  if (const char *e = getenv("GB_JAC_G"))
    G = std::max(1, std::min(4, atoi(e)));
  while (G > 1 && jac_smem_bytes(a.dm, G) > maxsm)
    --G;
  if (jac_smem_bytes(a.dm, G) > maxsm)
    return cudaErrorInvalidConfiguration; // mechanism too large for the shared-memory resident plan
  a.G = G;
  a.GS = G;
  const size_t smem = jac_smem_bytes(a.dm, G);
  switch (G)
  {
  case 4:
    return launch_jac_g<4>(a, smem, s);
  case 3:
    return launch_jac_g<3>(a, smem, s);
  case 2:
    return launch_jac_g<2>(a, smem, s);
  default:
    return launch_jac_g<1>(a, smem, s);
  }
}

} // namespace gb
