// gb_device.cuh -- device functions shared by the chemistry kernels (thermo polynomials, rate constants,
// third-body concentrations). FP64 throughout. Formulas and, where it is free, the order of floating-point
// operations follow the reference so results agree to the last few ulp; citations are to
// /root/reference/src/spitfire/griffon/src/*.cpp.
#pragma once
#include "gb_mech.h"

namespace gb
{

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

__device__ __forceinline__ int f_type(int f) { return f & F_TYPE_MASK; }
__device__ __forceinline__ int f_kform(int f) { return (f >> F_KFORM_SHIFT) & 7; }
__device__ __forceinline__ int f_troe(int f) { return (f >> F_TROE_SHIFT) & 7; }

// NASA7 / constant-cp species properties at temperature T.
//  cp_i, dcp_i/dT : thermodynamics_kernels.cpp:45-131, 183-260 (frozen outside [Tmin,Tmax])
//  h_i            : thermodynamics_kernels.cpp:262-351 (linear extension outside [Tmin,Tmax])
//  g_i (molar Gibbs), dB_i/dT : chemistry_kernels.cpp:56-97, rates_sensitivities_exact.cpp:82-126 (only T<=Tmid test)
struct SpeciesThermo
{
  double cp, h, dcp, g, dB;
};

template <bool JAC>
__device__ __forceinline__ SpeciesThermo species_thermo(const DeviceMech &dm, int i, double T, double logT,
                                                        double invT)
{
  SpeciesThermo o;
  const double *c = dm.cpc + (size_t)i * NCP;
  const double iw = dm.invmw[i];
  const int type = dm.cptype[i];
  if (type == CP_NASA7)
  {
    const double Tmid = c[0], minT = dm.tmin[i], maxT = dm.tmax[i];
    const bool low = T <= Tmid;
    const double b0 = low ? c[8] : c[1], b1 = low ? c[9] : c[2], b2 = low ? c[10] : c[3], b3 = low ? c[11] : c[4],
                 b4 = low ? c[12] : c[5], b5 = low ? c[13] : c[6], b6 = low ? c[14] : c[7];
    // Gibbs and dB/dT are not clipped
    o.g = b5 + T * (b0 - b6 - b0 * logT - T * (b1 + T * (b2 + T * (b3 + T * b4))));
    if (JAC)
    {
      const double invRu = dm.invRu, Ru = dm.RuR;
      o.dB = invRu * ((b0 - Ru) * invT + b1 + T * (2 * b2 + T * (3 * b3 + T * 4 * b4)) + b5 * invT * invT);
    }
    else
      o.dB = 0.;
    if ((low && T >= minT) || (!low && T <= maxT))
    {
      o.cp = iw * (b0 + T * (2. * b1 + T * (6. * b2 + T * (12. * b3 + 20. * T * b4))));
      o.h = iw * (b5 + T * (b0 + T * (b1 + T * (2. * b2 + T * (3. * b3 + T * 4. * b4)))));
      o.dcp = iw * ((2. * b1 + T * (12. * b2 + T * (36. * b3 + 80. * T * b4))));
    }
    else
    {
      // below Tmin the low branch is active (T < Tmin <= Tmid), above Tmax the high branch
      const double tb = low ? minT : maxT;
      o.cp = iw * (b0 + tb * (2. * b1 + tb * (6. * b2 + tb * (12. * b3 + 20. * tb * b4))));
      o.h = iw * (b5 + b0 * T +
                  tb * (2. * b1 * T +
                        tb * (3. * 2. * b2 * T - b1 +
                              tb * (4. * 3. * b3 * T - 2. * 2. * b2 +
                                    tb * (5. * 4. * b4 * T - 3. * 3. * b3 + tb * -4. * 4. * b4)))));
      o.dcp = 0.;
    }
  }
  else if (type == CP_NASA9)
  {
    // NASA9, several temperature regions. cp, dcp/dT: thermodynamics_kernels.cpp:91-124, 229-253 (frozen outside
    // [Tmin, Tmax]); h: :305-347 (linear extension); Gibbs and dB/dT: rates_sensitivities_exact.cpp:101-116 (the first
    // region with T < Thi, else the last one -- chemistry_kernels.cpp:74-88 picks the same region whenever T lies
    // strictly inside one and leaves the value unset otherwise).
    const double *c9 = dm.n9 + dm.n9_off[i];
    const int nreg = (int)c9[0];
    const double minT = dm.tmin[i], maxT = dm.tmax[i];
    int kg = nreg - 1;
    for (int k = 0; k < nreg - 1; ++k)
      if (T < c9[1 + k * 11 + 1])
      {
        kg = k;
        break;
      }
    {
      const double *a = c9 + 1 + kg * 11 + 2;
      o.g = a[7] - 0.5 * a[0] * invT + a[1] * (logT + 1.0) -
            T * (a[2] * (logT - 1.0) + a[8] +
                 T * (0.5 * a[3] + T * (0.1666666666666666 * a[4] + T * (0.0833333333333333 * a[5] + T * 0.05 * a[6]))));
      if (JAC)
      {
        const double invRu = dm.invRu, Ru = dm.RuR;
        o.dB = invRu * (invT * (a[2] - Ru + invT * (a[7] + a[1] * logT - invT * a[0])) + 0.5 * a[3] +
                        T * (a[4] * 0.3333333333333333 + T * (0.25 * a[5] + T * 0.2 * a[6])));
      }
      else
        o.dB = 0.;
    }
    auto cp_of = [](const double *a, double t, double it) {
      return it * (a[1] + it * a[0]) + a[2] + t * (a[3] + t * (a[4] + t * (a[5] + t * a[6])));
    };
    auto h_over_t = [](const double *a, double t, double it, double lt) {
      return it * (a[7] + lt * a[1] - a[0] * it) + a[2] +
             t * (0.5 * a[3] + t * (0.3333333333333333 * a[4] + t * (0.25 * a[5] + t * 0.2 * a[6])));
    };
    if (T < minT || T > maxT)
    {
      const double tb = T < minT ? minT : maxT;
      const double *a = c9 + 1 + (T < minT ? 0 : nreg - 1) * 11 + 2;
      const double itb = 1. / tb, ltb = log(tb);
      const double hb = iw * tb * h_over_t(a, tb, itb, ltb);
      o.cp = iw * cp_of(a, tb, itb);
      o.h = hb + o.cp * (T - tb);
      o.dcp = 0.;
    }
    else
    {
      int kr = nreg - 1; // (T == Tmax matches no region in the reference; the last one is used here)
      for (int k = 0; k < nreg; ++k)
        if (T >= c9[1 + k * 11] && T < c9[1 + k * 11 + 1])
        {
          kr = k;
          break;
        }
      const double *a = c9 + 1 + kr * 11 + 2;
      o.cp = iw * cp_of(a, T, invT);
      o.h = iw * T * h_over_t(a, T, invT, logT);
      o.dcp = iw * (-invT * invT * (a[1] + invT * 2.0 * a[0]) + a[3] + T * (2.0 * a[4] + T * (3.0 * a[5] + T * 4.0 * a[6])));
    }
  }
  else
  { // CP_CONST: c = {T0, h0, s0, cp}
    o.cp = iw * c[3];
    o.h = iw * (c[1] + c[3] * (T - c[0]));
    o.dcp = 0.;
    o.g = c[1] + c[3] * (T - c[0]) - T * (c[2] + c[3] * (logT - log(c[0])));
    if (JAC)
    {
      const double invRu = dm.invRu;
      o.dB = invT * (dm.mw[i] * invRu * (c[3] - invT * (c[3] * c[0] - c[1])) - 1);
    }
    else
      o.dB = 0.;
  }
  return o;
}

// forward rate constant by temperature form, chemistry_kernels.cpp:140-157
__device__ __forceinline__ double rate_constant(int kform, double A, double b, double E, double T, double invT,
                                                double logT)
{
  switch (kform)
  {
  case KF_CONSTANT:
    return A;
  case KF_LINEAR:
    return A * T;
  case KF_QUADRATIC:
    return A * T * T;
  case KF_RECIPROCAL:
    return A * invT;
  default:
    return A * exp(b * logT - E * invT);
  }
}

} // namespace gb
