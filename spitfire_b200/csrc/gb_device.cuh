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
