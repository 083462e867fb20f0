// gb_react.cuh -- per-reaction device functions of the Jacobian kernels (k_jac in gb_jac.cu, k_jac4 in gb_jac4.cu):
// a lane evaluates one reaction for one state of the tile and writes the reaction's record {q, dq/drho, dq/dT,
// [a, b,] dq/dY_slot...} into shared memory. rates_sensitivities_exact.cpp:33-1028 restated per reaction.
#pragma once
#include <cuda_runtime.h>

#include "gb_device.cuh"
#include "gb_kernels.cuh"

namespace gb
{

// per-state scalars of the tile, [JP_NSC][G]
enum JScalar : int
{
  J_T = 0,
  J_LOGT,
  J_INVT,
  J_RHO,
  J_MMW,
  J_INVM,  // 1/mmw
  J_CT,    // rho/mmw
  J_IRHO,  // 1/rho
  J_DRHOF, // (1/ct)*(1/mmw): d(prod C)/drho = R * sum_nu * J_DRHOF
  J_LPRT,  // log(p_ref/(Ru T))
  J_CP,
  J_DCP,
  J_DPART, // partial sum of Y_i/M_i over i < ns-1
  J_M0,    // open reactor: inflow enthalpy term
  J_YCP,   // open reactor: sum cp_i y_in,i
  J_OBASE, // (bits) offset of the state's output block
  J_CMOFF, // (bits) offset of the flamelet point's cmajor row
  J_TTC,   // flamelet (T,T) enthalpy-flux correction
  J_NLOGT, // log T and 1/T of the NEXT tile (formed while the last warp idles, see k_jac)
  J_NINVT
};
static_assert(J_NINVT < JP_NSC, "JP_NSC too small");

#define SMG(arr, idx, g) (arr)[(idx)*G + (g)]

__device__ __forceinline__ double u2d(unsigned long long u) { return __longlong_as_double((long long)u); }

struct JacSmem
{
  double *sc, *sy, *sC, *sg, *sdb, *sh, *scp, *sdcp, *su, *snm, *sim, *sR, *sTH;
  unsigned short *semap;
  double *sF = nullptr; // k_jac4: precomputed third-body / falloff factors {C_tbaf, dC/drho, dC/dT, coef}[G] per reaction
};

// ------------------------------------------------------------------------------------------------------------------
// fast path: simple reaction A + B (<)=> C + D with unit coefficients, none of them the last species.
// rates_sensitivities_exact.cpp:128-1009 specialised; record = {q, dq/drho, dq/dT, dq/dY_A, dq/dY_B, dq/dY_C, dq/dY_D}
// ------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void react_fast(const DeviceMech &dm, const unsigned long long *__restrict__ P, int g,
                                           const JacSmem &s)
{
  const ulonglong2 *P2 = reinterpret_cast<const ulonglong2 *>(P);
  const ulonglong2 q0 = __ldg(P2), q1 = __ldg(P2 + 1), q2 = __ldg(P2 + 2), q3 = __ldg(P2 + 3), q4 = __ldg(P2 + 4),
                   q5 = __ldg(P2 + 5);
  const int f = (int)(unsigned int)q0.x;
  double *rec = s.sR + (size_t)(unsigned int)(q0.x >> 32) * G + g;
  const int ia = (int)(q0.y & 0xffff), ib = (int)((q0.y >> 16) & 0xffff);
  const int ic = (int)((q0.y >> 32) & 0xffff), id = (int)((q0.y >> 48) & 0xffff);
  const bool rev = (f & F_REVERSIBLE) != 0;
  const double T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g), logT = SMG(s.sc, J_LOGT, g);
  const double rho = SMG(s.sc, J_RHO, g), drhof = SMG(s.sc, J_DRHOF, g);
  const double kfA = u2d(q2.x), kfb = u2d(q2.y), kfE = u2d(q3.x);
  const double cA = SMG(s.sC, ia, g), cB = SMG(s.sC, ib, g), cC = SMG(s.sC, ic, g), cD = SMG(s.sC, id, g);
  // sum_net nu_i g_i and sum_net nu_i dB_i/dT in ascending species order (:528-560)
  const unsigned long long ni = q1.x, nn = q1.y;
  double gs, ds;
  {
    const int i0 = (int)(ni & 0xffff);
    const double s0 = (double)(int)(signed char)(nn & 255);
    gs = s0 * SMG(s.sg, i0, g);
    ds = s0 * SMG(s.sdb, i0, g);
  }
#pragma unroll
  for (int i = 1; i < 4; ++i)
  {
    const int ii = (int)((ni >> (16 * i)) & 0xffff);
    const double si = (double)(int)(signed char)((nn >> (8 * i)) & 255);
    gs = fma(si, SMG(s.sg, ii, g), gs);
    ds = fma(si, SMG(s.sdb, ii, g), ds);
  }
  const double sum_stoich = (double)(int)(signed char)((nn >> 32) & 255);
  // the two exponentials are independent instruction streams: Arrhenius factor and 1/K_c (0 if irreversible)
  const int kform = f_kform(f);
  const double ninf = __longlong_as_double(0xfff0000000000000LL);
  const double arg_r = rev ? sum_stoich * SMG(s.sc, J_LPRT, g) - invT * dm.invRu * (gs) : ninf; // :535
  double ef = 1., invKc;
  if (kform == KF_ARRHENIUS)
  { // two independent exponentials back to back
    ef = exp(kfb * logT - kfE * invT);
    invKc = exp(arg_r);
  }
  else
    invKc = exp(arg_r);
  double kf; // chemistry_kernels.cpp:140-157
  switch (kform)
  {
  case KF_CONSTANT:
    kf = kfA;
    break;
  case KF_LINEAR:
    kf = kfA * T;
    break;
  case KF_QUADRATIC:
    kf = kfA * T * T;
    break;
  case KF_RECIPROCAL:
    kf = kfA * invT;
    break;
  default:
    kf = kfA * ef;
  }
  const double kf_sens = invT * (kfb + kfE * invT); // ARRHENIUS_SENS_OVER_K, :25
  const double kfr = kf * rho;
  const double Rf = kf * cA * cB;              // :287-325
  const double kr = kf * invKc;
  const double Rr = kr * cC * cD;
  const double krr = kr * rho;
  rec[0] = Rf - Rr;
  rec[G] = Rf * drhof * 2. - Rr * drhof * 2.; // sums of the coefficients = 2
  rec[2 * G] = Rf * kf_sens - Rr * (kf_sens + ds);
  rec[3 * G] = kfr * u2d(q3.y) * cB;           // :332-526
  rec[4 * G] = kfr * u2d(q4.x) * cA;
  if (rev)
  { // :528-812
    rec[5 * G] = -(krr * u2d(q4.y) * cD);
    rec[6 * G] = -(krr * u2d(q5.x) * cC);
  }
}

// third-body / falloff factor C_tbaf of a structured reaction and its sensitivities (rates_sensitivities_exact.cpp
// :826-1000): dC/drho, dC/dT and `coef` with dC/dY_s = coef*(base*u_s + eps_s - eps_last). kf, kf_sens: high-pressure
// rate constant and its ARRHENIUS_SENS_OVER_K.
template <int G>
__device__ __forceinline__ void falloff_eval(const DeviceMech &dm, const unsigned long long *__restrict__ P, int f, int ntb,
                                             double kf, double kf_sens, int g, const JacSmem &s, double &Ctbaf,
                                             double &dCdrho, double &dCdT, double &coef)
{
  const ulonglong2 *P2 = reinterpret_cast<const ulonglong2 *>(P);
  const int type = f_type(f);
  const double T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g), logT = SMG(s.sc, J_LOGT, g);
  const double rho = SMG(s.sc, J_RHO, g);
  const unsigned long long *Ptb = P + 20;
  const double base = u2d(P[12]);
  const double invM = SMG(s.sc, J_INVM, g), ct = SMG(s.sc, J_CT, g);
  dCdT = 0.;
  double M = base * ct;
  double dMdrho = base * invM;
  for (int i = 0; i < ntb; ++i)
  {
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(Ptb) + i);
    const double e = u2d(t.y) * SMG(s.sy, (int)(t.x & 0xffff), g);
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
    // Reciprocals are formed once and reused (1/(1+pr), 1/bTroe, 1/fCent): the reference divides each time, which
    // differs by an ulp or so per factor; the divisions sit on the critical path of the longest reaction groups
    const ulonglong2 k0 = __ldg(P2 + 6), k1 = __ldg(P2 + 7); // base, kpA | kpb, kpE
    const double kpb = u2d(k1.x), kpE = u2d(k1.y);
    const double kp_over_kf = u2d(k0.y) * exp(kpb * logT - kpE * invT) / kf;
    const double dsens = invT * (kpb + kpE * invT) - kf_sens; // kp_sens - kf_sens
    const double pr = kp_over_kf * M;
    const double inv1p = 1. / (1. + pr);
    double nsTmp;
    if (type == RT_LINDEMANN)
    { // :867-903
      Ctbaf = pr * inv1p;
      dCdT = Ctbaf * inv1p * dsens;
      nsTmp = kp_over_kf * (inv1p * inv1p);
    }
    else
    { // TROE, :905-995. Absent terms are evaluated as exp(-inf) = 0 so that the three exponentials and the
      // logarithms are independent instruction streams
      const ulonglong2 t0 = __ldg(P2 + 8), t1 = __ldg(P2 + 9);
      const double tr0 = u2d(t0.x), tr1 = u2d(t0.y), tr2 = u2d(t1.x), tr3 = u2d(t1.y);
      const int tb = f_troe(f);
      const double ninf = __longlong_as_double(0xfff0000000000000LL);
      const double a1 = (tb & TROE_T3) ? -T / tr1 : ninf, a2 = (tb & TROE_T1) ? -T / tr2 : ninf,
                   a3 = (tb & TROE_T2) ? -invT * tr3 : ninf;
      const double t1exp = exp(a1), t2exp = exp(a2), t3exp = exp(a3);
      const double log10pr = log10(fmax(pr, 1.e-300));
      double fCent = 0., dfCentdT = 0.;
      if (tb & TROE_T3)
      {
        fCent = (1 - tr0) * t1exp;
        dfCentdT = (tr0 - 1) / tr1 * t1exp;
      }
      if (tb & TROE_T1)
      {
        fCent = (tb & TROE_T3) ? fCent + tr0 * t2exp : tr0 * t2exp;
        dfCentdT = (tb & TROE_T3) ? dfCentdT - tr0 / tr2 * t2exp : -tr0 / tr2 * t2exp;
      }
      if (tb & TROE_T2)
      {
        const bool any = (tb & (TROE_T3 | TROE_T1)) != 0;
        fCent = any ? fCent + t3exp : t3exp;
        dfCentdT = any ? dfCentdT + t3exp * tr3 * invT * invT : t3exp * tr3 * invT * invT;
      }
      const double fc = fmax(fCent, 1.e-300);
      const double log10fcent = log10(fc);
      const double logfcent = log(fc);
      const double invfc = 1. / fCent;
      const double invln10 = 1. / 2.302585092994046; // 1/log(10.)
      const double aTroe = log10pr - 0.67 * log10fcent - 0.4;
      const double bTroe = -0.14 * log10pr - 1.1762 * log10fcent + 0.806;
      const double invb = 1. / bTroe;
      const double ab = aTroe * invb;
      const double gTroe = 1 / (1 + ab * ab);
      const double fTroe = pow(fCent, gTroe);
      const double prinv = pr * inv1p; // pr/(1+pr) = 1/(1+1/pr)
      Ctbaf = fTroe * prinv;
      const double common = -2.0 * gTroe * gTroe * invln10 * aTroe * (invb * invb * invb); // -2 g^2/ln10 a/b^3
      const double dfc = dfCentdT * invfc;
      const double dfTroedT =
          fTroe * (gTroe * dfc + logfcent * (common * ((bTroe + 0.14 * aTroe) * dsens - (0.67 * bTroe - 1.1762 * aTroe) * dfc)));
      dCdT = prinv * dfTroedT + fTroe * prinv * inv1p * dsens;
      nsTmp = kp_over_kf * (inv1p * fTroe * logfcent * common * (bTroe + 0.14 * aTroe) + fTroe * (inv1p * inv1p));
    }
    dCdrho = nsTmp * dMdrho;
    coef = nsTmp * rho;
  }
}

// k_jac4 producer task: the factor of one third-body / falloff reaction for one state -> s.sF (see react_struct<G, true>)
template <int G>
__device__ __forceinline__ void falloff_task(const DeviceMech &dm, const unsigned long long *__restrict__ P, int g,
                                             const JacSmem &s)
{
  const ulonglong2 *P2 = reinterpret_cast<const ulonglong2 *>(P);
  const ulonglong2 q0 = __ldg(P2), q1 = __ldg(P2 + 1), q2 = __ldg(P2 + 2);
  const int f = (int)(unsigned int)q0.x;
  const int ntb = (int)((q0.y >> 24) & 255);
  const double T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g), logT = SMG(s.sc, J_LOGT, g);
  const double kfb = u2d(q1.y), kfE = u2d(q2.x);
  const double kf = rate_constant(f_kform(f), u2d(q1.x), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT);
  double Ctbaf, dCdrho, dCdT, coef;
  falloff_eval<G>(dm, P, f, ntb, kf, kf_sens, g, s, Ctbaf, dCdrho, dCdT, coef);
  double *F = s.sF + (size_t)(unsigned int)P[11] * 4 * G + g;
  F[0] = Ctbaf, F[G] = dCdrho, F[2 * G] = dCdT, F[3 * G] = coef;
}

// ------------------------------------------------------------------------------------------------------------------
// structured path: up to three reactant and three product entries with coefficients 1..3, none of them the last
// species, optionally with a third-body / Lindemann / Troe factor (the last species may be a third body).
// rates_sensitivities_exact.cpp:128-1009 with every loop unrolled and the parameter record decoded once.
// Record = {q, dq/drho, dq/dT, dq/dY_slot...} (simple) or {q, dq/drho, dq/dT, a, b, dq/dY_slot...} (other types)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mul_pow(double v, double c, int nu, bool seq)
{ // v * c^nu with the association of the reference: special-cased orders (v*c)*c, generic branch v*(c*c)
  if (nu == 1)
    return v * c;
  if (seq)
  {
    v = v * c * c;
    return nu == 3 ? v * c : v;
  }
  return nu == 2 ? v * (c * c) : v * (c * c * c);
}

template <int G, bool PRE = false>
__device__ __forceinline__ void react_struct(const DeviceMech &dm, const unsigned long long *__restrict__ P, int g,
                                             const JacSmem &s)
{
  const ulonglong2 *P2 = reinterpret_cast<const ulonglong2 *>(P);
  const ulonglong2 q0 = __ldg(P2), q1 = __ldg(P2 + 1), q2 = __ldg(P2 + 2), q3 = __ldg(P2 + 3), q4 = __ldg(P2 + 4),
                   q5 = __ldg(P2 + 5);
  const int f = (int)(unsigned int)q0.x;
  const unsigned long long w1 = q0.y;
  const int nrc = (int)(w1 & 255), npd = (int)((w1 >> 8) & 255), ntb = (int)((w1 >> 24) & 255),
            nslots = (int)((w1 >> 32) & 255);
  const double sum_stoich = (double)(int)(signed char)((w1 >> 40) & 255);
  const double sum_rc = (double)(int)((w1 >> 48) & 255), sum_pd = (double)(int)((w1 >> 56) & 255);
  const int type = f_type(f);
  const int hdr = type == RT_SIMPLE ? JP_HDR_FAST : JP_HDR_GEN;
  double *rec = s.sR + (size_t)(unsigned int)(q0.x >> 32) * G + g;
  const unsigned int ent[6] = {(unsigned int)q2.y, (unsigned int)(q2.y >> 32), (unsigned int)q3.x,
                               (unsigned int)(q3.x >> 32), (unsigned int)q3.y, (unsigned int)(q3.y >> 32)};
  const unsigned int net[6] = {(unsigned int)q4.x, (unsigned int)(q4.x >> 32), (unsigned int)q4.y,
                               (unsigned int)(q4.y >> 32), (unsigned int)q5.x, (unsigned int)(q5.x >> 32)};

  const double T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g), logT = SMG(s.sc, J_LOGT, g);
  const double rho = SMG(s.sc, J_RHO, g), drhof = SMG(s.sc, J_DRHOF, g);
  const double kfb = u2d(q1.y), kfE = u2d(q2.x);
  const double kf = rate_constant(f_kform(f), u2d(q1.x), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT); // ARRHENIUS_SENS_OVER_K, :25
  const bool fseq = (f & F_FWD_SPECIAL) != 0, rseq = (f & F_REV_SPECIAL) != 0;

  for (int k = 0; k < nslots; ++k)
    rec[(hdr + k) * G] = 0.;

  // concentrations, coefficients, 1/M of the entries (unused entries: species 0, coefficient 0)
  double c[6], im[6];
  int nu[6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
  {
    const int idx = (int)(ent[i] & 0xffff);
    c[i] = SMG(s.sC, idx, g);
    im[i] = s.sim[idx];
    nu[i] = (int)((ent[i] >> 16) & 255);
  }
  // forward rate and its derivatives, :287-526
  double Rnet = kf;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < nrc)
      Rnet = mul_pow(Rnet, c[i], nu[i], fseq);
  double dRdrho = Rnet * drhof * sum_rc;
  double dRdT = Rnet * kf_sens;
  const double kfr = kf * rho;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < nrc)
    {
      double d = kfr * im[i];
      if (nu[i] == 2)
        d = d * 2. * c[i];
      else if (nu[i] == 3)
        d = d * 3. * c[i] * c[i];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j != i && j < nrc)
          d = mul_pow(d, c[j], nu[j], fseq);
      rec[(hdr + (int)(ent[i] >> 24)) * G] = d;
    }
  if (f & F_REVERSIBLE)
  { // :528-812
    double gs, ds;
    {
      const double s0 = (double)(int)(signed char)((net[0] >> 16) & 255);
      gs = s0 * SMG(s.sg, (int)(net[0] & 0xffff), g);
      ds = s0 * SMG(s.sdb, (int)(net[0] & 0xffff), g);
    }
#pragma unroll
    for (int i = 1; i < 6; ++i)
    {
      const double si = (double)(int)(signed char)((net[i] >> 16) & 255);
      gs = fma(si, SMG(s.sg, (int)(net[i] & 0xffff), g), gs);
      ds = fma(si, SMG(s.sdb, (int)(net[i] & 0xffff), g), ds);
    }
    const double invKc = exp(sum_stoich * SMG(s.sc, J_LPRT, g) - invT * dm.invRu * (gs)); // 1/K_c, :535
    const double kr = kf * invKc;
    double Rr = kr;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < npd)
        Rr = mul_pow(Rr, c[3 + i], nu[3 + i], rseq);
    Rnet -= Rr;
    dRdrho -= Rr * drhof * sum_pd;
    dRdT -= Rr * (kf_sens + ds);
    const double krr = kr * rho;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < npd)
      {
        double d = krr * im[3 + i];
        if (nu[3 + i] == 2)
          d = d * 2. * c[3 + i];
        else if (nu[3 + i] == 3)
          d = d * 3. * c[3 + i] * c[3 + i];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j != i && j < npd)
            d = mul_pow(d, c[3 + j], nu[3 + j], rseq);
        rec[(hdr + (int)(ent[3 + i] >> 24)) * G] -= d;
      }
  }
  if (type == RT_SIMPLE)
  {
    rec[0] = Rnet;
    rec[G] = dRdrho;
    rec[2 * G] = dRdT;
    return;
  }

  // third-body / falloff factor C_tbaf and its sensitivities, :826-1000
  const unsigned long long *Ptb = P + 20;
  const double base = u2d(P[12]);
  double Ctbaf, dCdrho, dCdT, coef; // dCtbaf/dY_s = coef*(base*u_s + eps_s - eps_last)
  if (PRE)
  { // k_jac4: evaluated ahead of the reaction phase by the producer warps (falloff_task)
    const double *F = s.sF + (size_t)(unsigned int)P[11] * 4 * G + g;
    Ctbaf = F[0], dCdrho = F[G], dCdT = F[2 * G], coef = F[3 * G];
  }
  else
    falloff_eval<G>(dm, P, f, ntb, kf, kf_sens, g, s, Ctbaf, dCdrho, dCdT, coef);
  rec[0] = Rnet * Ctbaf;                      // q, :1002
  rec[G] = dRdrho * Ctbaf + dCdrho * Rnet;    // dq/drho
  rec[2 * G] = dRdT * Ctbaf + dCdT * Rnet;    // dq/dT
  double b = 0.;
  for (int k = 0; k < nslots; ++k)
    rec[(JP_HDR_GEN + k) * G] *= Ctbaf;
  for (int i = 0; i < ntb; ++i)
  {
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(Ptb) + i);
    const double e = coef * u2d(t.y);
    const int slot = (int)(signed char)((t.x >> 24) & 255);
    if (slot >= 0)
      rec[(JP_HDR_GEN + slot) * G] += e * Rnet;
    else
      b -= e * Rnet; // the last species is a third body: -eps_last on every column (:855-865)
  }
  rec[3 * G] = coef * base * Rnet;
  rec[4 * G] = b;
}

// ------------------------------------------------------------------------------------------------------------------
// generic path: any reaction without non-elementary orders. rates_sensitivities_exact.cpp:128-1009 restated per
// reaction; the `for s < ns-1` dense loops (:522-525, :807-810, :849-850, :859-862, ...) are carried by the two
// scalars a, b. Record = {q, dq/drho, dq/dT, a, b, dq/dY_slot...}
// ------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void react_generic(const DeviceMech &dm, const unsigned long long *__restrict__ P, int g,
                                              const JacSmem &s)
{
  const unsigned long long w0 = P[0], w1 = P[1];
  const int f = (int)(unsigned int)w0;
  double *rec = s.sR + (size_t)(unsigned int)(w0 >> 32) * G + g;
  const int nrc = (int)(w1 & 255), npd = (int)((w1 >> 8) & 255), nn = (int)((w1 >> 16) & 255),
            ntb = (int)((w1 >> 24) & 255), nslots = (int)((w1 >> 32) & 255);
  const int sum_stoich = (int)(signed char)((w1 >> 40) & 255), sum_rc = (int)((w1 >> 48) & 255),
            sum_pd = (int)((w1 >> 56) & 255);
  const int type = f_type(f);
  const unsigned long long *Prc = P + (type == RT_SIMPLE ? 5 : 13);
  const unsigned long long *Ppd = Prc + 2 * nrc;
  const unsigned long long *Pnet = Ppd + 2 * npd;
  const unsigned long long *Ptb = Pnet + nn;

  const int last = dm.ns - 1;
  const double T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g), logT = SMG(s.sc, J_LOGT, g);
  const double rho = SMG(s.sc, J_RHO, g), drhof = SMG(s.sc, J_DRHOF, g);
  const double invM = SMG(s.sc, J_INVM, g), ct = SMG(s.sc, J_CT, g);
  const double invRu = dm.invRu;
  for (int k = 0; k < nslots; ++k)
    rec[(JP_HDR_GEN + k) * G] = 0.;

  const double kfb = u2d(P[3]), kfE = u2d(P[4]);
  const double kf = rate_constant(f_kform(f), u2d(P[2]), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT); // ARRHENIUS_SENS_OVER_K, :25
  double cR = 0.; // dense offset of dRnet/dY_s: -(last species as reactant) + (last species as product)

#define SP_IDX(Q, i) ((int)(Q[2 * (i)] & 0xffff))
#define SP_ST(Q, i) ((int)((Q[2 * (i)] >> 16) & 255))
#define SP_SLOT(Q, i) ((int)(signed char)((Q[2 * (i)] >> 24) & 255))
#define SP_INVMW(Q, i) u2d(Q[2 * (i) + 1])
#define SP_CONC(Q, i) SMG(s.sC, SP_IDX(Q, i), g)

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
  double dRnetdrho = Rnet * drhof * sum_rc;
  double dRnetdT = Rnet * kf_sens;
  for (int i = 0; i < nrc; ++i)
  { // :332-526
    const double d = deriv(kf * rho * SP_INVMW(Prc, i), Prc, nrc, i, fseq, false);
    if (SP_IDX(Prc, i) == last)
      cR -= d;
    else
      rec[(JP_HDR_GEN + SP_SLOT(Prc, i)) * G] = d;
  }
  if (f & F_REVERSIBLE)
  { // :528-812
    double gs, ds;
    {
      const int i0 = (int)(Pnet[0] & 0xffff), s0 = (int)(signed char)((Pnet[0] >> 16) & 255);
      gs = s0 * SMG(s.sg, i0, g);
      ds = s0 * SMG(s.sdb, i0, g);
    }
    for (int i = 1; i < nn; ++i)
    {
      const int ii = (int)(Pnet[i] & 0xffff), si = (int)(signed char)((Pnet[i] >> 16) & 255);
      gs = gs + si * SMG(s.sg, ii, g);
      ds = ds + si * SMG(s.sdb, ii, g);
    }
    const double invKc = exp(sum_stoich * SMG(s.sc, J_LPRT, g) - invT * invRu * (gs));
    const double dKc = -ds;
    const double kr = kf * invKc;
    const double Rr = mult(kr, Ppd, npd, -1, rseq, false);
    Rnet -= Rr;
    dRnetdrho -= Rr * drhof * sum_pd;
    dRnetdT -= Rr * (kf_sens - dKc);
    for (int i = 0; i < npd; ++i)
    {
      const double d = deriv(kr * rho * SP_INVMW(Ppd, i), Ppd, npd, i, rseq, true);
      if (SP_IDX(Ppd, i) == last)
        cR += d;
      else
        rec[(JP_HDR_GEN + SP_SLOT(Ppd, i)) * G] -= d;
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
      const double e = u2d(Ptb[2 * i + 1]) * SMG(s.sy, (int)(Ptb[2 * i] & 0xffff), g);
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
      rec[(JP_HDR_GEN + k) * G] *= Ctbaf;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = coef * u2d(Ptb[2 * i + 1]);
      const int slot = (int)(signed char)((Ptb[2 * i] >> 24) & 255);
      if (slot >= 0)
        rec[(JP_HDR_GEN + slot) * G] += e * Rnet;
      else
        b -= e * Rnet; // the last species is a third body: -eps_last on every column (:855-865)
    }
    rec[3 * G] = coef * base * Rnet;
  }
  else
    rec[3 * G] = 0.;
  rec[4 * G] = b;
#undef SP_IDX
#undef SP_ST
#undef SP_SLOT
#undef SP_INVMW
#undef SP_CONC
}

// reactions with non-elementary orders (rare), rates_sensitivities_exact.cpp:198-281; parameters from the SoA tables
template <int G>
__device__ __noinline__ void react_orders(int ns, const int *__restrict__ dm_n_sp, const short *__restrict__ dm_sp_idx,
                                          const double *__restrict__ dm_sp_order,
                                          const signed char *__restrict__ dm_sp_slot,
                                          const double *__restrict__ dm_invmw, int r,
                                          const unsigned long long *__restrict__ P, int g, double *sc, double *sC,
                                          double *sy, double *sR)
{
  JacSmem s;
  s.sc = sc, s.sC = sC, s.sy = sy, s.sR = sR;
  const unsigned long long w0 = P[0], w1 = P[1];
  const int f = (int)(unsigned int)w0;
  double *rec = s.sR + (size_t)(unsigned int)(w0 >> 32) * G + g;
  const int nrc = (int)(w1 & 255), npd = (int)((w1 >> 8) & 255), nn = (int)((w1 >> 16) & 255),
            nslots = (int)((w1 >> 32) & 255);
  const int type = f_type(f);
  const unsigned long long *Pnet = P + (type == RT_SIMPLE ? 5 : 13) + 2 * nrc + 2 * npd;
  const int last = ns - 1;
  const double T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g), logT = SMG(s.sc, J_LOGT, g);
  const double rho = SMG(s.sc, J_RHO, g);
  const double invM = SMG(s.sc, J_INVM, g), ct = SMG(s.sc, J_CT, g);
  for (int k = 0; k < nslots; ++k)
    rec[(JP_HDR_GEN + k) * G] = 0.;
  const double kfb = u2d(P[3]), kfE = u2d(P[4]);
  const double kf = rate_constant(f_kform(f), u2d(P[2]), kfb, kfE, T, invT, logT);
  const double kf_sens = invT * (kfb + kfE * invT);
  const int n = dm_n_sp[r];
  const short *sp = dm_sp_idx + NSR * (size_t)r;
  const double *ord = dm_sp_order + NSR * (size_t)r;
  const signed char *spslot = dm_sp_slot + NSR * (size_t)r;
#define CS(i) SMG(s.sC, sp[i], g)
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
        const double pre = ord[l] * rho * dm_invmw[sp[l]];
        if (ord[l] > 1 || is_last)
          v *= pre * pow(fmax(cl, 1.e-16), ord[l] - 1.);
        else
          v *= pre / pow(fmax(cl, 1.e-16), 1. - ord[l]);
      }
    }
    if (is_last)
      cR -= v;
    else
      rec[(JP_HDR_GEN + spslot[j]) * G] = v;
  }
#undef CS
  // third-body factors of non-elementary reactions: only the plain third-body form is supported here
  double Ctbaf = 1., dCdrho = 0., coef = 0., b = 0.;
  const double base = (type != RT_SIMPLE) ? u2d(P[5]) : 0.;
  const int ntb = (int)((w1 >> 24) & 255);
  const unsigned long long *Ptb = Pnet + nn;
  if (type == RT_THIRD_BODY)
  {
    double M = base * ct, dMdrho = base * invM;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = u2d(Ptb[2 * i + 1]) * SMG(s.sy, (int)(Ptb[2 * i] & 0xffff), g);
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
      rec[(JP_HDR_GEN + k) * G] *= Ctbaf;
    for (int i = 0; i < ntb; ++i)
    {
      const double e = coef * u2d(Ptb[2 * i + 1]);
      const int slot = (int)(signed char)((Ptb[2 * i] >> 24) & 255);
      if (slot >= 0)
        rec[(JP_HDR_GEN + slot) * G] += e * Rnet;
      else
        b -= e * Rnet;
    }
    rec[3 * G] = coef * base * Rnet;
  }
  else
    rec[3 * G] = 0.;
  rec[4 * G] = b;
}


} // namespace gb
