/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE. extern "C" shim around the UNMODIFIED reference Griffon C++.
 *
 * Compiled by oracle/build_oracle.py together with the reference's own sources where they lie
 * (/root/reference/src/spitfire/griffon/src/*.cpp, include/*.h) into oracle/_ref/libref_griffon.so.
 * Nothing from the reference is copied into this repository; this file only forwards calls to
 * griffon::CombustionKernels (combustion_kernels.h:33) and griffon::btddod::* (btddod_matrix_kernels.h).
 * It plays the role of the reference's Cython layer (griffon.pyx) without needing Cython-generated code.
 */
#include "griffon_oracle.h"

#include <map>
#include <string>
#include <vector>

#include "btddod_matrix_kernels.h"
#include "combustion_kernels.h"

struct go_mech
{
  griffon::CombustionKernels ck;
  std::map<std::string, double> element_mw;
  std::vector<double> mw; // recomputed the way chemistry_setup.cpp:52-56 does, for go_mech_molecular_weights
  int n_reactions = 0;
};

namespace
{
template <class T>
std::map<std::string, T> to_map(int n, const char *const *names, const T *vals)
{
  std::map<std::string, T> out;
  for (int i = 0; i < n; ++i)
    out[names[i]] = vals[i];
  return out;
}
} // namespace

extern "C"
{

  const char *go_kind(void) { return "reference"; }

  go_mech *go_mech_create(void) { return new go_mech(); }
  void go_mech_destroy(go_mech *m) { delete m; }

  int go_mech_set_ref_pressure(go_mech *m, double p)
  {
    m->ck.mechanism_set_ref_pressure(p);
    return 0;
  }
  int go_mech_set_ref_temperature(go_mech *m, double T)
  {
    m->ck.mechanism_set_ref_temperature(T);
    return 0;
  }
  int go_mech_set_gas_constant(go_mech *m, double Ru)
  {
    m->ck.mechanism_set_gas_constant(Ru);
    return 0;
  }
  int go_mech_set_element_mw(go_mech *m, const char *element, double mw)
  {
    m->element_mw[element] = mw;
    m->ck.mechanism_set_element_mw_map(m->element_mw);
    return 0;
  }
  int go_mech_add_element(go_mech *m, const char *element)
  {
    m->ck.mechanism_add_element(element);
    return 0;
  }
  int go_mech_add_species(go_mech *m, const char *name, int n_atoms, const char *const *atom_names,
                          const double *atom_counts)
  {
    try
    {
      const auto am = to_map<double>(n_atoms, atom_names, atom_counts);
      m->ck.mechanism_add_species(name, am);
      double mw = 0.;
      for (const auto &a : am)
        mw += m->element_mw.at(a.first) * a.second;
      m->mw.push_back(mw);
    }
    catch (...)
    {
      return -1;
    }
    return 0;
  }
  int go_mech_resize_heat_capacity_data(go_mech *m)
  {
    m->ck.mechanism_resize_heat_capacity_data();
    return 0;
  }
  int go_mech_add_const_cp(go_mech *m, const char *s, double Tmin, double Tmax, double T0, double h0, double s0,
                           double cp)
  {
    try
    {
      m->ck.mechanism_add_const_cp(s, Tmin, Tmax, T0, h0, s0, cp);
    }
    catch (...)
    {
      return -1;
    }
    return 0;
  }
  int go_mech_add_nasa7_cp(go_mech *m, const char *s, double Tmin, double Tmid, double Tmax, const double *low7,
                           const double *high7)
  {
    try
    {
      m->ck.mechanism_add_nasa7_cp(s, Tmin, Tmid, Tmax, std::vector<double>(low7, low7 + 7),
                                   std::vector<double>(high7, high7 + 7));
    }
    catch (...)
    {
      return -1;
    }
    return 0;
  }
  int go_mech_add_nasa9_cp(go_mech *m, const char *s, double Tmin, double Tmax, int n, const double *c)
  {
    try
    {
      m->ck.mechanism_add_nasa9_cp(s, Tmin, Tmax, std::vector<double>(c, c + n));
    }
    catch (...)
    {
      return -1;
    }
    return 0;
  }

  int go_mech_add_reaction(go_mech *m, int type, int reversible, int n_reactants, const char *const *reactant_names,
                           const int *reactant_stoich, int n_products, const char *const *product_names,
                           const int *product_stoich, double fwd_A, double fwd_b, double fwd_Ea, int n_eff,
                           const char *const *eff_names, const double *eff_values, double default_eff, double flf_A,
                           double flf_b, double flf_Ea, const double *troe4, int n_orders,
                           const char *const *order_names, const double *order_values)
  {
    try
    {
      const auto rs = to_map<int>(n_reactants, reactant_names, reactant_stoich);
      const auto ps = to_map<int>(n_products, product_names, product_stoich);
      const auto eff = to_map<double>(n_eff, eff_names, eff_values);
      const auto ord = to_map<double>(n_orders, order_names, order_values);
      std::vector<double> troe(4, 0.);
      if (troe4)
        troe.assign(troe4, troe4 + 4);
      const bool rev = reversible != 0;
      auto &ck = m->ck;
      if (n_orders == 0)
      {
        switch (type)
        {
        case 1:
          ck.mechanism_add_reaction_simple(rs, ps, rev, fwd_A, fwd_b, fwd_Ea);
          break;
        case 2:
          ck.mechanism_add_reaction_three_body(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, eff, default_eff);
          break;
        case 3:
          ck.mechanism_add_reaction_Lindemann(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, eff, default_eff, flf_A, flf_b,
                                              flf_Ea);
          break;
        case 4:
          ck.mechanism_add_reaction_Troe(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, eff, default_eff, flf_A, flf_b, flf_Ea,
                                         troe);
          break;
        default:
          return -1;
        }
      }
      else
      {
        switch (type)
        {
        case 1:
          ck.mechanism_add_reaction_simple_with_special_orders(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, ord);
          break;
        case 2:
          ck.mechanism_add_reaction_three_body_with_special_orders(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, eff,
                                                                   default_eff, ord);
          break;
        case 3:
          ck.mechanism_add_reaction_Lindemann_with_special_orders(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, eff, default_eff,
                                                                  flf_A, flf_b, flf_Ea, ord);
          break;
        case 4:
          ck.mechanism_add_reaction_Troe_with_special_orders(rs, ps, rev, fwd_A, fwd_b, fwd_Ea, eff, default_eff,
                                                             flf_A, flf_b, flf_Ea, troe, ord);
          break;
        default:
          return -1;
        }
      }
      ++m->n_reactions;
    }
    catch (...)
    {
      return -2;
    }
    return 0;
  }

  int go_mech_n_species(const go_mech *m) { return (int)m->mw.size(); }
  int go_mech_n_reactions(const go_mech *m) { return m->n_reactions; }
  int go_mech_molecular_weights(const go_mech *m, double *out)
  {
    for (size_t i = 0; i < m->mw.size(); ++i)
      out[i] = m->mw[i];
    return 0;
  }

  double go_mixture_molecular_weight(const go_mech *m, const double *y) { return m->ck.mixture_molecular_weight(y); }
  void go_mole_fractions(const go_mech *m, const double *y, double *x) { m->ck.mole_fractions(y, x); }
  double go_ideal_gas_density(const go_mech *m, double p, double T, const double *y)
  {
    return m->ck.ideal_gas_density(p, T, y);
  }
  double go_ideal_gas_pressure(const go_mech *m, double rho, double T, const double *y)
  {
    return m->ck.ideal_gas_pressure(rho, T, y);
  }
  double go_cp_mix(const go_mech *m, double T, const double *y) { return m->ck.cp_mix(T, y); }
  double go_cv_mix(const go_mech *m, double T, const double *y) { return m->ck.cv_mix(T, y); }
  double go_enthalpy_mix(const go_mech *m, double T, const double *y) { return m->ck.enthalpy_mix(T, y); }
  double go_energy_mix(const go_mech *m, double T, const double *y) { return m->ck.energy_mix(T, y); }
  void go_species_cp(const go_mech *m, double T, double *out) { m->ck.species_cp(T, out); }
  void go_species_cv(const go_mech *m, double T, double *out) { m->ck.species_cv(T, out); }
  void go_species_enthalpies(const go_mech *m, double T, double *out) { m->ck.species_enthalpies(T, out); }
  void go_species_energies(const go_mech *m, double T, double *out) { m->ck.species_energies(T, out); }
  void go_cp_sens_T(const go_mech *m, double T, const double *y, double *a, double *b) { m->ck.cp_sens_T(T, y, a, b); }

  void go_production_rates(const go_mech *m, double T, double rho, const double *y, double *out_w)
  {
    m->ck.production_rates(T, rho, y, out_w);
  }
  void go_prod_rates_primitive_sensitivities(const go_mech *m, double rho, double T, const double *y, int option,
                                             double *out)
  {
    m->ck.prod_rates_primitive_sensitivities(rho, T, y, option, out);
  }

  void go_reactor_rhs_isobaric(const go_mech *m, const double *state, double p, double T_in, const double *y_in,
                               double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                               int heat_option, int open, double *out_rhs)
  {
    m->ck.reactor_rhs_isobaric(state, p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open != 0,
                               out_rhs);
  }
  void go_reactor_jac_isobaric(const go_mech *m, const double *state, double p, double T_in, const double *y_in,
                               double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                               int heat_option, int open, int rso, int sto, double *out_rhs, double *out_jac)
  {
    m->ck.reactor_jac_isobaric(state, p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open != 0,
                               rso, sto, out_rhs, out_jac);
  }

  void go_reactor_rhs_isochoric(const go_mech *m, const double *state, double rho_in, double T_in, const double *y_in,
                                double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                                int heat_option, int open, double *out_rhs)
  {
    m->ck.reactor_rhs_isochoric(state, rho_in, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option,
                                open != 0, out_rhs);
  }
  void go_reactor_jac_isochoric(const go_mech *m, const double *state, double rho_in, double T_in, const double *y_in,
                                double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                                int heat_option, int open, int rso, double *out_rhs, double *out_jac)
  {
    m->ck.reactor_jac_isochoric(state, rho_in, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option,
                                open != 0, rso, out_rhs, out_jac);
  }

  void go_reactor_jac_isobaric_many(const go_mech *m, int n, const double *state, double p, int rso, double *out_rhs,
                                    double *out_jac)
  {
    const int ns = (int)m->mw.size();
    const double dummy = 0.;
    for (int i = 0; i < n; ++i)
      m->ck.reactor_jac_isobaric(state + (size_t)i * ns, p, 0., &dummy, 0., 0., 0., 0., 0., 0., 0, false, rso, 0,
                                 out_rhs + (size_t)i * ns, out_jac + (size_t)i * ns * ns);
  }
  void go_reactor_rhs_isobaric_many(const go_mech *m, int n, const double *state, double p, double *out_rhs)
  {
    const int ns = (int)m->mw.size();
    const double dummy = 0.;
    for (int i = 0; i < n; ++i)
      m->ck.reactor_rhs_isobaric(state + (size_t)i * ns, p, 0., &dummy, 0., 0., 0., 0., 0., 0., 0, false,
                                 out_rhs + (size_t)i * ns);
  }

  void go_flamelet_stencils(const go_mech *m, const double *dz, int nzi, const double *chi, const double *inv_lewis,
                            double *cmajor, double *csub, double *csup, double *mcoeff, double *ncoeff)
  {
    m->ck.flamelet_stencils(dz, nzi, chi, inv_lewis, cmajor, csub, csup, mcoeff, ncoeff);
  }
  void go_flamelet_jac_indices(const go_mech *m, int nzi, int *rows, int *cols)
  {
    m->ck.flamelet_jac_indices(nzi, rows, cols);
  }
  void go_flamelet_rhs(const go_mech *m, const double *state, double p, const double *oxy, const double *fuel,
                       int adiabatic, const double *T_conv, const double *h_conv, const double *T_rad,
                       const double *h_rad, int nzi, const double *cmajor, const double *csub, const double *csup,
                       const double *mcoeff, const double *ncoeff, const double *chi, int ief, int ivc, int ushl,
                       double *out_rhs)
  {
    m->ck.flamelet_rhs(state, p, oxy, fuel, adiabatic != 0, T_conv, h_conv, T_rad, h_rad, nzi, cmajor, csub, csup,
                       mcoeff, ncoeff, chi, ief != 0, ivc != 0, ushl != 0, out_rhs);
  }
  void go_flamelet_jacobian(const go_mech *m, const double *state, double p, const double *oxy, const double *fuel,
                            int adiabatic, const double *T_conv, const double *h_conv, const double *T_rad,
                            const double *h_rad, int nzi, const double *cmajor, const double *csub,
                            const double *csup, const double *mcoeff, const double *ncoeff, const double *chi,
                            int compute_eigenvalues, double diffterm, int scale_and_offset, double prefactor, int rso,
                            int sto, int ief, int ivc, int ushl, double *out_expeig, double *out_jac)
  {
    m->ck.flamelet_jacobian(state, p, oxy, fuel, adiabatic != 0, T_conv, h_conv, T_rad, h_rad, nzi, cmajor, csub, csup,
                            mcoeff, ncoeff, chi, compute_eigenvalues != 0, diffterm, scale_and_offset != 0, prefactor,
                            rso, sto, ief != 0, ivc != 0, ushl != 0, out_expeig, out_jac);
  }

  void go_btddod_full_factorize(double *d, int nb, int bs, double *l, int *piv)
  {
    griffon::btddod::btddod_full_factorize(d, nb, bs, l, piv);
  }
  void go_btddod_full_solve(const double *d, const double *l, const int *piv, const double *rhs, int nb, int bs,
                            double *x)
  {
    griffon::btddod::btddod_full_solve(d, l, piv, rhs, nb, bs, x);
  }
  void go_btddod_full_matvec(const double *a, const double *v, int nb, int bs, double *out)
  {
    griffon::btddod::btddod_full_matvec(a, v, nb, bs, out);
  }
  void go_btddod_scale_and_add_diagonal(double *a, double ms, const double *d, double ds, int nb, int bs)
  {
    griffon::btddod::btddod_scale_and_add_diagonal(a, ms, d, ds, nb, bs);
  }
}
