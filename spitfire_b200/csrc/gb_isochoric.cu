// gb_isochoric.cu -- isochoric (constant-volume) reactor right-hand side and Jacobian for batches of states, FP64.
//
// Replaces reactor_rhs_isochoric / reactor_jac_isochoric (isochoric_reactor_kernels.cpp:192-335; griffon.pyx:831-866).
// State [rho, T, Y_0..Y_{ns-2}]; the Jacobian is (ns+1) x (ns+1) column-major in the primitive variables themselves,
// so it is a direct function of the production rates and their exact sensitivities: the heavy kernels are the ones
// of the isobaric path (k_rates MODE_PRODRATES, k_jac MODE_SENS), bracketed by two light kernels of this file:
//   k_iso_split    : state -> rho, T, y (Y_ns = 1 - sum, extract_y combustion_kernels.h:505-515)
//   k_iso_assemble : cv, cv_i, e_i, dcv/dT (thermodynamics_kernels.cpp:169-181, 353-364) and chem/mass/heat_*_isochoric
//                    (:19-190); one warp per state, lanes over species / Jacobian columns, the inner products of the
//                    reference summed in species order.
#include <cuda_runtime.h>

#include "gb_device.cuh"
#include "gb_kernels.cuh"

namespace gb
{

__global__ void k_iso_split(int ns, int n, const double *__restrict__ state, double *__restrict__ rho,
                            double *__restrict__ T, double *__restrict__ y)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  const double *q = state + (size_t)s * (ns + 1);
  rho[s] = q[0];
  T[s] = q[1];
  double *ys = y + (size_t)s * ns;
  double last = 1.;
  for (int i = 0; i < ns - 1; ++i)
  {
    const double v = q[2 + i];
    ys[i] = v;
    last -= v;
  }
  ys[ns - 1] = last;
}

// dynamic shared memory: per warp 4*ns doubles (cv_i, e_i, inflow e_i, w)
__global__ void __launch_bounds__(128) k_iso_assemble(const DeviceMech dm, int n, const double *__restrict__ state,
                                                      const double *__restrict__ y, const double *__restrict__ w_all,
                                                      const double *__restrict__ wsens_all, const ReactorDev rx,
                                                      double rho_in, double *__restrict__ out_rhs,
                                                      double *__restrict__ out_jac)
{
  extern __shared__ double sm[];
  const int ns = dm.ns, n1 = ns + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double *cvi = sm + (size_t)warp * 4 * ns, *e = cvi + ns, *ein = e + ns, *sw = ein + ns;
  const bool open = rx.open != 0;
  for (int s = blockIdx.x * wpb + warp; s < n; s += gridDim.x * wpb)
  {
    const double rho = state[(size_t)s * n1], T = state[(size_t)s * n1 + 1];
    const double *ys = y + (size_t)s * ns, *w = w_all + (size_t)s * ns;
    const double logT = log(T), invT = 1. / T, RT = dm.Ru * T;
    __syncwarp();
    for (int i = lane; i < ns; i += 32)
    {
      const SpeciesThermo t = species_thermo<false>(dm, i, T, logT, invT);
      cvi[i] = t.cp - dm.Ru * dm.invmw[i]; // cv_mix_and_species, thermodynamics_kernels.cpp:169-181
      e[i] = t.h - RT * dm.invmw[i];      // species_energies, :353-364
      sw[i] = w[i];
      if (open)
      {
        const double Ti = rx.T_in;
        ein[i] = species_thermo<false>(dm, i, Ti, log(Ti), 1. / Ti).h - dm.Ru * Ti * dm.invmw[i];
      }
    }
    __syncwarp();
    // mixture quantities in species order (every lane redundantly: no divergence, no broadcast)
    double cv = 0., cvsensT = 0., d = 0., we = 0., wcv = 0.;
    for (int i = 0; i < ns; ++i)
    {
      const SpeciesThermo t = species_thermo<false>(dm, i, T, logT, invT);
      cv += ys[i] * t.cp; // cp_mix_and_species :45-131
      if (dm.cptype[i] == CP_CONST)
        cvsensT = 0.; // sic, cp_sens_T thermodynamics_kernels.cpp:202
      else
        cvsensT += ys[i] * t.dcp;
      d += ys[i] * dm.invmw[i];
      we += e[i] * sw[i];   // inner_product(nSpec, w, e)
      wcv += cvi[i] * sw[i]; // inner_product(nSpec, w, cvi)
    }
    const double mmw = 1. / d;
    cv -= dm.Ru / mmw;
    const double invRho = 1. / rho, invCv = 1. / cv, invRhoCv = 1. / (rho * cv);
    const double rhs1c = -we / (rho * cv); // chem_rhs_isochoric :19-30
    // open-reactor and heat-transfer terms of the temperature equation (mass_rhs_isochoric :40-62, heat_rhs :32-38)
    double m1 = 0., ycv = 0., rate = 0.;
    const double invTau = open ? 1. / rx.tau : 0.;
    const double mfac = open ? invTau * rho_in * invRho : 0.;
    if (open)
    {
      m1 = (ein[ns - 1] - e[ns - 1]) * rx.y_in[ns - 1];
      for (int i = 0; i < ns - 1; ++i)
        m1 += (ein[i] - e[i]) * rx.y_in[i];
      m1 /= cv;
      m1 *= mfac;
      for (int i = 0; i < ns; ++i)
        ycv += cvi[i] * rx.y_in[i]; // inner_product(nSpec, inflowY, cvi)
    }
    if (rx.heat_option == 2)
      rate = rx.SoV / (rho * cv) *
             (rx.h_conv * (rx.T_inf - T) + rx.eps_rad * 5.67e-8 * (rx.T_surf * rx.T_surf * rx.T_surf * rx.T_surf - T * T * T * T));
    double *rhs = out_rhs + (size_t)s * n1;
    // ---- right-hand side ------------------------------------------------------------------------------------------------------
    for (int i = lane; i < n1; i += 32)
    {
      double v;
      if (i == 0)
        v = open ? 0. + (rho_in - rho) * invTau : 0.;
      else if (i == 1)
      {
        v = rhs1c;
        if (open)
          v += m1;
        if (rx.heat_option == 1)
          v = 0.;
        else if (rx.heat_option == 2)
          v += rate;
      }
      else
      {
        v = sw[i - 2] * invRho;
        if (open)
          v += (rx.y_in[i - 2] - ys[i - 2]) * mfac;
      }
      rhs[i] = v;
    }
    if (!out_jac)
      continue;
    // ---- Jacobian: lane per column (chem_jac_isochoric :64-110, mass_jac :112-160, heat_jac :162-190) ----------------
    const double *ws = wsens_all + (size_t)s * n1 * n1;
    double *J = out_jac + (size_t)s * n1 * n1;
    const double cvn = cvi[ns - 1];
    for (int c = lane; c < n1; c += 32)
    {
      const double *wc = ws + (size_t)c * n1;
      double *Jc = J + (size_t)c * n1;
      double dot = 0.;
      for (int i = 0; i < ns; ++i)
        dot += e[i] * wc[i]; // inner_product(nSpec, &wsens[c*(ns+1)], e)
      double jrho = 0., jT;
      if (c == 0)
      {
        jT = -invRhoCv * dot - invRho * rhs1c;
        if (open)
        {
          jrho = -invTau;
          jT += -invRho * m1;
        }
        if (rx.heat_option == 2)
          jT += -rate / rho;
        for (int i = 0; i < ns - 1; ++i)
        {
          double v = invRho * (wc[i] - invRho * sw[i]);
          if (open)
            v += -invRho * ((rx.y_in[i] - ys[i]) * mfac);
          Jc[2 + i] = v;
        }
      }
      else if (c == 1)
      {
        jT = -invCv * (invRho * (dot + wcv) + rhs1c * cvsensT);
        if (open)
          jT += -invCv * (invTau * invRho * (rho_in * ycv) + cvsensT * m1);
        if (rx.heat_option == 2)
          jT += -invCv * cvsensT * rate - rx.SoV * invRhoCv * (rx.h_conv + 4. * rx.eps_rad * 5.67e-8 * T * T * T);
        for (int i = 0; i < ns - 1; ++i)
          Jc[2 + i] = invRho * wc[i] + (open ? 0. : 0.);
      }
      else
      {
        const int k = c - 2;
        jT = -invRhoCv * dot - rhs1c * (cvi[k] - cvn) * invCv;
        if (open)
          jT += -invCv * m1 * (cvi[k] - cvn);
        if (rx.heat_option == 2)
          jT += invCv * rate * (cvn - cvi[k]);
        for (int i = 0; i < ns - 1; ++i)
        {
          double v = invRho * wc[i];
          if (open)
            v += (i == k) ? -invTau * rho_in / rho : 0.;
          Jc[2 + i] = v;
        }
      }
      if (rx.heat_option == 1)
        jT = 0.;
      Jc[0] = jrho;
      Jc[1] = jT;
    }
  }
}

cudaError_t launch_iso_split(int ns, int n, const double *state, double *rho, double *T, double *y, cudaStream_t s)
{
  k_iso_split<<<(n + 127) / 128, 128, 0, s>>>(ns, n, state, rho, T, y);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_iso_assemble(const DeviceMech &dm, int n, const double *state, const double *y, const double *w,
                                const double *wsens, const ReactorDev &rx, double rho_in, double *out_rhs,
                                double *out_jac, cudaStream_t s)
{
  const int threads = 128, wpb = threads / 32;
  const size_t smem = sizeof(double) * 4 * (size_t)dm.ns * wpb;
  if (smem > 48 * 1024)
  {
    static bool attr = false;
    if (!attr)
    {
      cudaError_t e = cudaFuncSetAttribute(k_iso_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess)
        return e;
      attr = true;
    }
  }
  const int blocks = std::max(1, std::min((n + wpb - 1) / wpb, 148 * 8));
  k_iso_assemble<<<blocks, threads, smem, s>>>(dm, n, state, y, w, wsens, rx, rho_in, out_rhs, out_jac);
  count_launch();
  return cudaGetLastError();
}

} // namespace gb
