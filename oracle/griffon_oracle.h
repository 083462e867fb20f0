/*
 * griffon_oracle.h -- C interface shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in spitfire_b200/ (the product) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Two shared objects export exactly these symbols:
 *   oracle/liboracle_port.so   <- oracle/griffon_oracle.c : plain-C restatement of the reference algorithm
 *   oracle/_ref/libref_griffon.so <- oracle/ref_shim.cpp  : thin extern "C" shim around the UNMODIFIED reference
 *                                    C++ (compiled from /root/reference/src/spitfire/griffon where it lies)
 * The functions are single-state and mirror the reference C++ methods one to one
 * (reference: src/spitfire/griffon/include/combustion_kernels.h:292-503, btddod_matrix_kernels.h).
 */
#ifndef GRIFFON_ORACLE_H
#define GRIFFON_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct go_mech go_mech;

const char *go_kind(void); /* "port" or "reference" */

go_mech *go_mech_create(void);
void go_mech_destroy(go_mech *m);
int go_mech_set_ref_pressure(go_mech *m, double p_ref);
int go_mech_set_ref_temperature(go_mech *m, double T_ref);
int go_mech_set_gas_constant(go_mech *m, double Ru);
int go_mech_set_element_mw(go_mech *m, const char *element, double mw);
int go_mech_add_element(go_mech *m, const char *element);
int go_mech_add_species(go_mech *m, const char *name, int n_atoms, const char *const *atom_names,
                        const double *atom_counts);
int go_mech_resize_heat_capacity_data(go_mech *m);
int go_mech_add_const_cp(go_mech *m, const char *species, double Tmin, double Tmax, double T0, double h0, double s0,
                         double cp);
int go_mech_add_nasa7_cp(go_mech *m, const char *species, double Tmin, double Tmid, double Tmax, const double *low7,
                         const double *high7);
int go_mech_add_nasa9_cp(go_mech *m, const char *species, double Tmin, double Tmax, int n_coeffs,
                         const double *coeffs);
int go_mech_add_reaction(go_mech *m, int type, int reversible, int n_reactants, const char *const *reactant_names,
                         const int *reactant_stoich, int n_products, const char *const *product_names,
                         const int *product_stoich, double fwd_A, double fwd_b, double fwd_Ea_over_R, int n_eff,
                         const char *const *eff_names, const double *eff_values, double default_eff, double flf_A,
                         double flf_b, double flf_Ea_over_R, const double *troe4, int n_orders,
                         const char *const *order_names, const double *order_values);
int go_mech_n_species(const go_mech *m);
int go_mech_n_reactions(const go_mech *m);
int go_mech_molecular_weights(const go_mech *m, double *out_mw);

/* thermodynamics */
double go_mixture_molecular_weight(const go_mech *m, const double *y);
void go_mole_fractions(const go_mech *m, const double *y, double *x);
double go_ideal_gas_density(const go_mech *m, double p, double T, const double *y);
double go_ideal_gas_pressure(const go_mech *m, double rho, double T, const double *y);
double go_cp_mix(const go_mech *m, double T, const double *y);
double go_cv_mix(const go_mech *m, double T, const double *y);
double go_enthalpy_mix(const go_mech *m, double T, const double *y);
double go_energy_mix(const go_mech *m, double T, const double *y);
void go_species_cp(const go_mech *m, double T, double *out);
void go_species_cv(const go_mech *m, double T, double *out);
void go_species_enthalpies(const go_mech *m, double T, double *out);
void go_species_energies(const go_mech *m, double T, double *out);
void go_cp_sens_T(const go_mech *m, double T, const double *y, double *out_cpmixsens, double *out_cpspeciessens);

/* kinetics */
void go_production_rates(const go_mech *m, double T, double rho, const double *y, double *out_w);
void go_prod_rates_primitive_sensitivities(const go_mech *m, double rho, double T, const double *y, int option,
                                           double *out_sens);

/* isobaric reactor */
void go_reactor_rhs_isobaric(const go_mech *m, const double *state, double p, double T_in, const double *y_in,
                             double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                             int heat_option, int open, double *out_rhs);
void go_reactor_jac_isobaric(const go_mech *m, const double *state, double p, double T_in, const double *y_in,
                             double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                             int heat_option, int open, int rates_sens_option, int sens_transform_option,
                             double *out_rhs, double *out_jac);

/* convenience for timing the CPU path without Python overhead: a plain loop of the single-state call over n states
 * (state [n][ns], out_rhs [n][ns], out_jac [n][ns*ns]); closed adiabatic reactor */
/* isochoric reactor (isochoric_reactor_kernels.cpp:192-335): state [rho, T, Y_0..Y_{ns-2}], rhs [ns+1], jac [(ns+1)^2] */
void go_reactor_rhs_isochoric(const go_mech *m, const double *state, double rho_in, double T_in, const double *y_in,
                              double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                              int heat_option, int open, double *out_rhs);
void go_reactor_jac_isochoric(const go_mech *m, const double *state, double rho_in, double T_in, const double *y_in,
                              double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                              int heat_option, int open, int rates_sens_option, double *out_rhs, double *out_jac);
void go_reactor_jac_isobaric_many(const go_mech *m, int n, const double *state, double p, int rates_sens_option,
                                  double *out_rhs, double *out_jac);
void go_reactor_rhs_isobaric_many(const go_mech *m, int n, const double *state, double p, double *out_rhs);

/* flamelet (argument order of the C++ methods: T_conv, h_conv, T_rad, h_rad) */
void go_flamelet_stencils(const go_mech *m, const double *dz, int nzi, const double *chi, const double *inv_lewis,
                          double *out_cmajor, double *out_csub, double *out_csup, double *out_mcoeff,
                          double *out_ncoeff);
void go_flamelet_jac_indices(const go_mech *m, int nzi, int *out_rows, int *out_cols);
void go_flamelet_rhs(const go_mech *m, const double *state, double p, const double *oxy, const double *fuel,
                     int adiabatic, const double *T_conv, const double *h_conv, const double *T_rad,
                     const double *h_rad, int nzi, const double *cmajor, const double *csub, const double *csup,
                     const double *mcoeff, const double *ncoeff, const double *chi, int include_enthalpy_flux,
                     int include_variable_cp, int use_scaled_heat_loss, double *out_rhs);
void go_flamelet_jacobian(const go_mech *m, const double *state, double p, const double *oxy, const double *fuel,
                          int adiabatic, const double *T_conv, const double *h_conv, const double *T_rad,
                          const double *h_rad, int nzi, const double *cmajor, const double *csub, const double *csup,
                          const double *mcoeff, const double *ncoeff, const double *chi, int compute_eigenvalues,
                          double diffterm, int scale_and_offset, double prefactor, int rates_sens_option,
                          int sens_transform_option, int include_enthalpy_flux, int include_variable_cp,
                          int use_scaled_heat_loss, double *out_expeig, double *out_jac);

/* BTDDOD block Thomas */
void go_btddod_full_factorize(double *d_factors, int num_blocks, int block_size, double *out_l_values,
                              int *out_d_pivots);
void go_btddod_full_solve(const double *d_factors, const double *l_values, const int *d_pivots, const double *rhs,
                          int num_blocks, int block_size, double *out_solution);
void go_btddod_full_matvec(const double *matrix, const double *vec, int num_blocks, int block_size,
                           double *out_matvec);
void go_btddod_scale_and_add_diagonal(double *matrix, double matrix_scale, const double *diagonal, double diag_scale,
                                      int num_blocks, int block_size);

#ifdef __cplusplus
}
#endif
#endif
