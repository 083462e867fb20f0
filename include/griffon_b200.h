/*
 * griffon_b200.h -- C-ABI of the B200-native Griffon hot path.
 *
 * This is the drop-in boundary: every entry point replaces one method of the
 * reference's Cython class `PyCombustionKernels` / module functions `py_btddod_*`
 * (reference: src/spitfire/griffon/griffon.pyx) or the C++ method behind it
 * (src/spitfire/griffon/include/combustion_kernels.h, btddod_matrix_kernels.h).
 * The citation next to each declaration names the interface it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary.
 *  - all functions return int: 0 = ok, <0 = argument / state / CUDA error
 *    (text via gb_last_error()). The synchronous `*_host` entry points of the
 *    reactor / flamelet right-hand sides and Jacobians return > 0 = the number
 *    of members (states, flamelets) whose output holds an Inf or NaN; for the
 *    asynchronous `*_batch` ones ask gb_count_nonfinite_members_batch. Nothing
 *    throws.
 *  - a gb_mech handle serves ONE stream and ONE host thread at a time: the
 *    flamelet and isochoric entry points and every `*_host` one keep work
 *    arrays in the handle (grown on demand with cudaMalloc, which synchronises
 *    the device the first time a larger batch arrives). Use one handle per
 *    stream / thread; handles are cheap (the tables are ~150 KB).
 *  - `*_batch` entry points take DEVICE pointers and a cudaStream_t passed as
 *    void* (NULL = default stream); they are asynchronous w.r.t. the host.
 *  - `*_host` entry points take HOST pointers, stage through device buffers
 *    owned by the handle, run the same kernels and copy back; synchronous.
 *  - single-state methods of the reference are `*_host` calls with n = 1.
 *  - layouts are the reference's: reactor state [T, Y_0..Y_{ns-2}] per state
 *    (AoS, stride ns), Jacobian ns x ns column-major per state, flamelet state
 *    [nzi][ns], flamelet Jacobian in BTDDOD storage (nzi col-major ns x ns
 *    blocks, then (nzi-1)*ns sub-diagonal, then (nzi-1)*ns super-diagonal).
 *  - there is NO CPU fallback: if no CUDA device is usable every compute entry
 *    point returns GB_ERR_CUDA.
 */
#ifndef GRIFFON_B200_H
#define GRIFFON_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GB_OK 0
#define GB_ERR_ARG (-1)
#define GB_ERR_STATE (-2)
#define GB_ERR_CUDA (-3)
#define GB_ERR_UNSUPPORTED (-4)

/* reaction types, as RateType in combustion_kernels.h:60-67 */
#define GB_RXN_SIMPLE 1
#define GB_RXN_THIRD_BODY 2
#define GB_RXN_LINDEMANN 3
#define GB_RXN_TROE 4

typedef struct gb_mech gb_mech; /* opaque; owns host tables + device copies (CombustionKernels*, griffon.pyx:220-229) */

const char *gb_last_error(void);
/* number of usable CUDA devices (0 if none), never fails */
int gb_cuda_device_count(void);

/* ---- mechanism construction: replaces the 18 `mechanism_*` setters ------------------------------------- */
gb_mech *gb_mech_create(void);                                  /* griffon.pyx:225 __cinit__            */
void gb_mech_destroy(gb_mech *m);                               /* griffon.pyx:228 __dealloc__          */
int gb_mech_set_ref_pressure(gb_mech *m, double p_ref);         /* chemistry_setup.cpp:64               */
int gb_mech_set_ref_temperature(gb_mech *m, double T_ref);      /* chemistry_setup.cpp:74               */
int gb_mech_set_gas_constant(gb_mech *m, double Ru);            /* chemistry_setup.cpp:69               */
int gb_mech_set_element_mw(gb_mech *m, const char *element, double mw); /* one entry of mechanism_set_element_mw_map, :21 */
int gb_mech_add_element(gb_mech *m, const char *element);       /* chemistry_setup.cpp:26               */
int gb_mech_add_species(gb_mech *m, const char *name, int n_atoms, const char *const *atom_names,
                        const double *atom_counts);             /* chemistry_setup.cpp:36               */
int gb_mech_resize_heat_capacity_data(gb_mech *m);              /* chemistry_setup.cpp:79               */
int gb_mech_add_const_cp(gb_mech *m, const char *species, double Tmin, double Tmax, double T0, double h0,
                         double s0, double cp);                 /* chemistry_setup.cpp:89               */
int gb_mech_add_nasa7_cp(gb_mech *m, const char *species, double Tmin, double Tmid, double Tmax,
                         const double *low7, const double *high7); /* chemistry_setup.cpp:102           */
int gb_mech_add_nasa9_cp(gb_mech *m, const char *species, double Tmin, double Tmax, int n_coeffs,
                         const double *coeffs);                 /* chemistry_setup.cpp:132; coeffs = {nregions, (Tlo, Thi, a0..a8) * nregions} */
/* One generic adder covers mechanism_add_reaction_{simple,three_body,Lindemann,Troe}[_with_special_orders]
 * (chemistry_setup.cpp:156-345). Unused groups are passed with n = 0 / NULL. Ea is Ea/Ru (K), as the
 * reference's Python passes it (mechanism.py:185). troe4 = [A, T3, T1, T2] zero padded (griffon.pyx:410-416). */
int gb_mech_add_reaction(gb_mech *m, int type, int reversible,
                         int n_reactants, const char *const *reactant_names, const int *reactant_stoich,
                         int n_products, const char *const *product_names, const int *product_stoich,
                         double fwd_A, double fwd_b, double fwd_Ea_over_R,
                         int n_eff, const char *const *eff_names, const double *eff_values, double default_eff,
                         double flf_A, double flf_b, double flf_Ea_over_R, const double *troe4,
                         int n_orders, const char *const *order_names, const double *order_values);
int gb_mech_n_species(const gb_mech *m);
int gb_mech_n_reactions(const gb_mech *m);
int gb_mech_molecular_weights(const gb_mech *m, double *out_mw /* [ns] host */);
/* packs the tables and uploads them to the current CUDA device (done lazily by every compute call) */
int gb_mech_commit(gb_mech *m);

/* ---- thermodynamics (griffon.pyx:684-758; thermodynamics_kernels.cpp) ----------------------------------- */
/* what: selects the quantity; T[n], y[n*ns] (full mass-fraction vectors), out sized per `what`.             */
#define GB_THERMO_MMW 0          /* mixture_molecular_weight(y)          out[n]    combustion_kernels.h:381 */
#define GB_THERMO_DENSITY 1      /* ideal_gas_density(p=aux,T,y)         out[n]    thermodynamics_kernels.cpp:29 */
#define GB_THERMO_PRESSURE 2     /* ideal_gas_pressure(rho=aux,T,y)      out[n]    thermodynamics_kernels.cpp:37 */
#define GB_THERMO_CP_MIX 3       /* cp_mix(T,y)                          out[n]    :133 */
#define GB_THERMO_CV_MIX 4       /* cv_mix(T,y)                          out[n]    :149 */
#define GB_THERMO_H_MIX 5        /* enthalpy_mix(T,y)                    out[n]    :366 */
#define GB_THERMO_E_MIX 6        /* energy_mix(T,y)                      out[n]    :375 */
#define GB_THERMO_CP_SPECIES 7   /* species_cp(T)                        out[n*ns] :143 */
#define GB_THERMO_CV_SPECIES 8   /* species_cv(T)                        out[n*ns] :156 */
#define GB_THERMO_H_SPECIES 9    /* species_enthalpies(T)                out[n*ns] :262 */
#define GB_THERMO_E_SPECIES 10   /* species_energies(T)                  out[n*ns] :353 */
#define GB_THERMO_DCPDT_SPECIES 11 /* dcpdT_species(T,y)                 out[n*ns] :183 */
#define GB_THERMO_MOLE_FRACTIONS 12 /* mole_fractions(y)                 out[n*ns] :17  */
int gb_thermo_batch(gb_mech *m, int what, int n, const double *aux /* [n] or NULL */, const double *T,
                    const double *y, double *out, void *stream);
int gb_thermo_host(gb_mech *m, int what, int n, const double *aux, const double *T, const double *y, double *out);

/* ---- kinetics (griffon.pyx:763-783; chemistry_kernels.cpp:35-483) ---------------------------------------- */
/* production_rates(T, rho, y, out_w): y full [n*ns], out_w [n*ns] */
int gb_production_rates_batch(gb_mech *m, int n, const double *T, const double *rho, const double *y,
                              double *out_w, void *stream);
int gb_production_rates_host(gb_mech *m, int n, const double *T, const double *rho, const double *y, double *out_w);
/* prod_rates_primitive_sensitivities(rho, T, y, option, out[(ns+1)^2]) col-major, ld ns+1 per state */
int gb_prod_rates_sens_batch(gb_mech *m, int n, const double *rho, const double *T, const double *y,
                             int rates_sensitivity_option, double *out_sens, void *stream);
int gb_prod_rates_sens_host(gb_mech *m, int n, const double *rho, const double *T, const double *y,
                            int rates_sensitivity_option, double *out_sens);

/* ---- isobaric reactor (griffon.pyx:788-824; isobaric_reactor_kernels.cpp:170-343) ------------------------ */
typedef struct gb_reactor_params {
  double pressure;            /* Pa, shared by the batch                                         */
  double inflow_temperature;  /* T_in                                                            */
  const double *inflow_y;     /* [ns] full inflow mass fractions; only read when open != 0       */
  double tau;                 /* mixing time                                                     */
  double fluid_temperature;   /* T_inf  (convection)                                             */
  double surf_temperature;    /* T_surf (radiation)                                              */
  double h_conv;
  double eps_rad;
  double surface_area_over_volume;
  int heat_transfer_option;   /* 0 adiabatic, 1 isothermal, 2 diathermal (:206-218)              */
  int open;                   /* bool                                                            */
} gb_reactor_params;
/* state [n*ns] = [T, Y_0..Y_{ns-2}] per state; out_rhs [n*ns]. inflow_y is a DEVICE pointer for _batch. */
int gb_reactor_rhs_isobaric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                  double *out_rhs, void *stream);
int gb_reactor_rhs_isobaric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                 double *out_rhs);
/* out_jac [n*ns*ns], per state column-major (row i, col j at i + j*ns). rates_sensitivity_option 0 (dense),
 * 2 (sparse) give identical results; 1 (no-TBAF, inexact) is mapped to exact. sensitivity_transform_option 0. */
int gb_reactor_jac_isobaric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                  int rates_sensitivity_option, int sensitivity_transform_option,
                                  double *out_rhs, double *out_jac, void *stream);
int gb_reactor_jac_isobaric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                 int rates_sensitivity_option, int sensitivity_transform_option, double *out_rhs,
                                 double *out_jac);

/* ---- isochoric reactor (griffon.pyx:831-866; isochoric_reactor_kernels.cpp:192-335) ------------------------- */
/* state [n*(ns+1)] = [rho, T, Y_0..Y_{ns-2}] per state; out_rhs [n*(ns+1)]; out_jac [n*(ns+1)^2], per state
 * column-major in the primitive variables (row i, col j at i + j*(ns+1)). `prm->pressure` is not read; the inflow
 * density replaces it (reactor_rhs_isochoric's inflowDensity). rates_sensitivity_option as for the isobaric call.
 * Work arrays live in the handle (scratch slots): one stream at a time per handle. */
int gb_reactor_rhs_isochoric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                   double inflow_density, double *out_rhs, void *stream);
int gb_reactor_rhs_isochoric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                  double inflow_density, double *out_rhs);
int gb_reactor_jac_isochoric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                   double inflow_density, int rates_sensitivity_option, double *out_rhs,
                                   double *out_jac, void *stream);
int gb_reactor_jac_isochoric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                  double inflow_density, int rates_sensitivity_option, double *out_rhs,
                                  double *out_jac);

/* ---- flamelet (griffon.pyx:556-679; flamelet_kernels.cpp:31-90, 1039-1409) -------------------------------- */
/* Host-side setup helpers (cold; run once per Flamelet) */
int gb_flamelet_stencils(const gb_mech *m, const double *dz, int nzi, const double *dissipation_rate,
                         const double *inv_lewis, double *out_cmajor, double *out_csub, double *out_csup,
                         double *out_mcoeff, double *out_ncoeff);
int gb_flamelet_jac_indices(const gb_mech *m, int nzi, int *out_rows, int *out_cols);

typedef struct gb_flamelet_params {
  int nzi;                    /* interior grid points                                                        */
  double pressure;
  const double *oxy_state;    /* [ns]  = [T, Y_0..Y_{ns-2}] of the oxidizer stream (Z=0 boundary)            */
  const double *fuel_state;   /* [ns]  fuel stream (Z=1 boundary)                                            */
  int adiabatic;              /* bool                                                                        */
  const double *T_convection; /* per flamelet [nzi]; the four heat-loss arrays are read only if !adiabatic   */
  const double *h_convection;
  const double *T_radiation;
  const double *h_radiation;
  const double *cmajor;       /* per flamelet [nzi*ns]                                                       */
  const double *csub;
  const double *csup;
  const double *mcoeff;       /* per flamelet [nzi]                                                          */
  const double *ncoeff;
  const double *chi;          /* per flamelet [nzi+2] full-grid dissipation rate, indexed chi[i] as the
                                 reference does (flamelet_kernels.cpp:1175)               */
  int include_enthalpy_flux;
  int include_variable_cp;
  int use_scaled_heat_loss;
  /* strides (in doubles) between consecutive flamelets of a batch for each per-flamelet array; 0 = shared   */
  long stride_heat;           /* T_convection,h_convection,T_radiation,h_radiation                           */
  long stride_coeff;          /* cmajor,csub,csup                                                            */
  long stride_mn;             /* mcoeff,ncoeff                                                               */
  long stride_chi;            /* chi                                                                         */
} gb_flamelet_params;
/* state, out_rhs: [F][nzi*ns] */
int gb_flamelet_rhs_batch(gb_mech *m, int n_flamelets, const double *state, const gb_flamelet_params *prm,
                          double *out_rhs, void *stream);
int gb_flamelet_rhs_host(gb_mech *m, int n_flamelets, const double *state, const gb_flamelet_params *prm,
                         double *out_rhs);
/* out_jac: [F][ns*(nzi*ns + 2*(nzi-1))] BTDDOD; written completely (no need to pre-zero, unlike the reference
 * which accumulates into the off-diagonals, flamelet_kernels.cpp:1386-1394). out_expeig [F][nzi*ns] is written
 * only if compute_eigenvalues: max(max_i Re(lambda_i) - diffterm, 0) of the point's transformed chemical block,
 * repeated for the point's ns unknowns (flamelet_kernels.cpp:1329-1341, where LAPACK dgeev does it on the host;
 * here balancing + Householder-Hessenberg + double-shift QR run on the device, one warp per block). */
int gb_flamelet_jacobian_batch(gb_mech *m, int n_flamelets, const double *state, const gb_flamelet_params *prm,
                               int compute_eigenvalues, double diffterm, int scale_and_offset, double prefactor,
                               int rates_sensitivity_option, int sensitivity_transform_option, double *out_expeig,
                               double *out_jac, void *stream);
int gb_flamelet_jacobian_host(gb_mech *m, int n_flamelets, const double *state, const gb_flamelet_params *prm,
                              int compute_eigenvalues, double diffterm, int scale_and_offset, double prefactor,
                              int rates_sensitivity_option, int sensitivity_transform_option, double *out_expeig,
                              double *out_jac);

/* Extension (the reference keeps this inside flamelet_jacobian): largest real part of the eigenvalues of nblocks
 * dense n x n matrices stored back to back (either major order: the spectrum of the transpose is the same);
 * replaces griffon::lapack::eigenvalues (blas_lapack_kernels.h:157-180) + the max over Re (flamelet_kernels.cpp:1332-
 * 1336). blocks, out: device. n <= 169 (the matrix lives in shared memory). */
int gb_max_real_eigenvalue_batch(int nblocks, int n, const double *blocks, double *out, void *stream);

/* ---- BTDDOD block-Thomas (griffon.pyx:1006-1113; btddod_matrix_kernels.cpp:19-165, 429-465) --------------- */
/* Batched over n_systems independent matrices stored back to back (stride = block_size*(num_blocks*block_size
 * + 2*(num_blocks-1)) doubles for matrices, num_blocks*block_size^2 for l_values, num_blocks*block_size for
 * pivots / vectors). Pivots are LAPACK-style 1-based row indices (dgetrf semantics). */
int gb_btddod_full_factorize_batch(int n_systems, double *d_factors, int num_blocks, int block_size,
                                   double *out_l_values, int *out_d_pivots, void *stream);
int gb_btddod_full_solve_batch(int n_systems, const double *d_factors, const double *l_values, const int *d_pivots,
                               const double *rhs, int num_blocks, int block_size, double *out_solution,
                               void *stream);
int gb_btddod_full_matvec_batch(int n_systems, const double *matrix, const double *vec, int num_blocks,
                                int block_size, double *out_matvec, void *stream);
/* A <- matrix_scale*A + diag_scale*diag(diagonal) */
int gb_btddod_scale_and_add_diagonal_batch(int n_systems, double *matrix, double matrix_scale,
                                           const double *diagonal, double diag_scale, int num_blocks,
                                           int block_size, void *stream);
/* Extension (no counterpart in the reference API): the factorisation also returns the explicit inverses of the
 * factorised diagonal blocks, out_dinv [n_systems][num_blocks*block_size^2] (it forms them anyway for L_{i+1},
 * btddod_matrix_kernels.cpp:48-63), and gb_btddod_full_solve_inv_batch runs the back sweep as matrix-vector
 * products with them (no pivots, no triangular solves). Same result up to rounding; used inside Newton loops.
 * system_rows (device, may be NULL): the factors of the k-th right-hand side are those of system system_rows[k] of the
 * factor arrays, so a solver can address a subset of a batch without gathering 8 MB of factors per member.
 * The solve is launched as thread-block clusters of two CTAs per system when the sizes allow it (num_blocks >= 4, the
 * right-hand side fits in shared memory): factors tagged by gb_btddod_full_invert_twisted_batch (below) are then
 * applied from both ends at once; factors of the two one-sided eliminations are swept by the first CTA alone.
 * block 0 of l_values is never a multiplier: it is zero (one-sided eliminations) or holds the tag. */
int gb_btddod_full_factorize_inv_batch(int n_systems, double *d_factors, int num_blocks, int block_size,
                                       double *out_l_values, int *out_d_pivots, double *out_dinv, void *stream);
int gb_btddod_full_solve_inv_batch(int n_systems, const double *d_factors, const double *l_values, const double *dinv,
                                   const double *rhs, int num_blocks, int block_size, double *out_solution,
                                   const int *system_rows, void *stream);
/* Extension: the same elimination by explicit inverses only -- Gauss-Jordan with partial pivoting on every
 * D'_i = D_i - L_i diag(sup_{i-1}), L_i = diag(sub_{i-1}) D'_{i-1}^{-1} (btddod_matrix_kernels.cpp:48-75) -- for solvers
 * that apply the result through gb_btddod_full_solve_inv_batch alone (pass `matrix` as its d_factors: only the
 * super-diagonal is read from it). `matrix` is NOT overwritten; no LU factors or pivots are produced. About 5x
 * shorter latency chain per system than factorize_inv (one barrier and bs, not ~3 bs, dependent steps per block). */
int gb_btddod_full_invert_batch(int n_systems, const double *matrix, int num_blocks, int block_size,
                                double *out_l_values, double *out_dinv, void *stream);
/* Extension: the TWISTED ("burn at both ends") form of the elimination above, for the same consumers. Two CTAs of a
 * thread-block cluster share a system: one eliminates downwards from block 0 (L_i, D'_i as above), the other upwards
 * from block nb-1 (U_i = diag(sup_i) D''_{i+1}^{-1}, D''_i = D_i - U_i diag(sub_i)); they meet at block m = (nb-1)/2,
 * D*_m = D_m - L_m diag(sup_{m-1}) - U_m diag(sub_m). out_l_values holds L_i in slot i <= m and U_{i-1} in slot i > m,
 * out_dinv the inverses of D'_i (i < m), D*_m, D''_i (i > m); block 0 of out_l_values (never read as a multiplier)
 * carries the tag {m, magic} by which gb_btddod_full_solve_inv_batch recognises the format and sweeps from both ends
 * at once (two CTAs, two cluster barriers): the dependent chain of the elimination AND of every solve is halved.
 * Same solution as the one-sided elimination up to rounding (a different, equally stable elimination order). Falls
 * back to the one-sided form for num_blocks < 4 or when GB_BT_TWIST=0 is set in the environment. */
int gb_btddod_full_invert_twisted_batch(int n_systems, const double *matrix, int num_blocks, int block_size,
                                        double *out_l_values, double *out_dinv, void *stream);
int gb_btddod_full_factorize_host(int n_systems, double *d_factors, int num_blocks, int block_size,
                                  double *out_l_values, int *out_d_pivots);
int gb_btddod_full_solve_host(int n_systems, const double *d_factors, const double *l_values, const int *d_pivots,
                              const double *rhs, int num_blocks, int block_size, double *out_solution);
int gb_btddod_full_matvec_host(int n_systems, const double *matrix, const double *vec, int num_blocks,
                               int block_size, double *out_matvec);
int gb_btddod_scale_and_add_diagonal_host(int n_systems, double *matrix, double matrix_scale,
                                          const double *diagonal, double diag_scale, int num_blocks,
                                          int block_size);

/* ---- vector kernels of the batched implicit integrator (host solver loops) ---------------------
 * Replace, for a batch of independent members on the device, the numpy expressions of the reference's ESDIRK stage loop
 * (time/methods.py:502-612), SimpleNewtonSolver (time/nonlinear.py:185-268) and the embedded error estimate for the
 * PI controller (time/stepcontrol.py:84-101). Members are the rows of [n][ndof] device arrays; the arithmetic follows
 * the reference's expressions operation by operation. No reference C++ counterpart (the reference does this in numpy).
 *
 * stage begin: explicit = coef[nk-1]*k[nk-1] + ... + coef[0]*k[0] (accumulated from j = nk-1 down, methods.py:560-575),
 *              res = dt*(gamma*f + explicit) - (x - q) (nonlinear.py:204), conv[m] = 0.
 *              k: HOST array of nk (<= 6) device pointers, coef: HOST array. */
int gb_esdirk_stage_begin_batch(int n, int ndof, int nk, const double *const *k, const double *coef, double gamma,
                                const double *dt, const double *x, const double *q, const double *f,
                                double *explicit_out, double *res_out, int *conv, void *stream);
/* xn = conv ? x : x - dx; also resets *n_unconverged (device int) for the tail kernel of the same iteration */
int gb_newton_update_batch(int n, int ndof, const double *x, const double *dx, const int *conv, double *xn,
                           int *n_unconverged, void *stream);
/* rn = dt*(gamma*fn + explicit) - (xn - q); members with conv == 0 take (xn, fn, rn) as their new (x, f, res) and set
 * conv when max|rn*weights| < tolerance (nonlinear.py:236-257); *n_unconverged counts the members still iterating.
 * host_count (may be NULL): if given, the count is copied there and the stream is synchronised -- the one integer
 * the host loop branches on. */
int gb_newton_tail_batch(int n, int ndof, const double *fn, const double *xn, const double *explicit_, const double *q,
                         const double *dt, double gamma, const double *weights, double tolerance, double *x, double *f,
                         double *res, int *conv, int *n_unconverged, int *host_count, void *stream);
/* dq = dt*(b[0]*k[0] + ... + b[nk-1]*k[nk-1]), dqh likewise with bh (methods.py:598-610); stats [3][n]:
 * max|(dq - dqh)*weights| (error estimate), max|dq*weights|, 1/0 = every dq finite / not. k, b, bh: HOST arrays. */
int gb_esdirk_finish_batch(int n, int ndof, int nk, const double *const *k, const double *b, const double *bh,
                           const double *dt, const double *weights, double *dq, double *stats, void *stream);
/* The whole Newton loop of one implicit stage for F flamelets (time/nonlinear.py:185-268 for a batch): per iteration
 * dx = solve_inv(factors, res); xn = x - dx; fn = flamelet_rhs(xn); gb_newton_tail_batch(...), until every member has
 * converged or max_iterations is reached. Arguments as in gb_btddod_full_solve_inv_batch (num_blocks = prm->nzi,
 * block_size = n_species), gb_flamelet_rhs_batch and gb_newton_tail_batch; work: 3*F*nzi*ns doubles. Synchronous.
 * Returns the number of members that did not converge (>= 0) or a negative error code; *out_iterations (may be NULL)
 * receives the number of iterations taken. The host cost of an iteration is four launches and one synchronisation. */
int gb_flamelet_newton_stage_batch(gb_mech *m, int F, const gb_flamelet_params *prm, const double *d_factors,
                                   const double *l_values, const double *dinv, const int *system_rows,
                                   const double *explicit_, const double *q, const double *dt, double gamma,
                                   const double *weights, double tolerance, int max_iterations, double *x, double *f,
                                   double *res, int *conv, double *work, int *n_unconverged, int *out_iterations,
                                   void *stream);
/* The Newton tail with a per-member stage machine: like gb_newton_tail_batch, but a member that converges (or has taken
 * max_iterations) at its stage s stores K[s] = f and prepares stage s+1 itself (explicit part from tableau row s+1 and
 * K[0..s], first residual), or sets done[m] after the last stage; nlfail[m] records a stage that ran out of
 * iterations (nonlinear.py:259-268). K: [nstages][n][ndof]; tableau: HOST, nstages x nstages row-major (a[s][j]);
 * stage / iters / nlfail / done: device int [n]; *n_left counts the members that are not done. The members of a batch
 * then take the stages of a step independently of each other. */
int gb_newton_tail_staged_batch(int n, int ndof, int nstages, const double *tableau, int max_iterations, const double *fn,
                                const double *xn, const double *q, const double *dt, double gamma,
                                const double *weights, double tolerance, double *x, double *f, double *res,
                                double *explicit_, double *K, int *stage, int *iters, int *nlfail, int *done,
                                int *n_left, int *host_count, void *stream);
/* All implicit stages of one ESDIRK step for F flamelets (methods.py:502-612 for a batch): rounds of {solve_inv, update,
 * flamelet rhs, staged tail} until every member is done. On entry: K[0] = f(q), (x, f) = (q, K[0]), explicit_ / res
 * prepared for stage 1 (gb_esdirk_stage_begin_batch), stage[m] = 1, iters = nlfail = done = 0; work: 3*F*nzi*ns
 * doubles. Synchronous; returns the number of members not done (0) or a negative error code; *out_rounds = rounds of
 * kernels taken (the largest per-member sum of Newton iterations over the stages). */
int gb_flamelet_esdirk_stages_batch(gb_mech *m, int F, const gb_flamelet_params *prm, const double *d_factors,
                                    const double *l_values, const double *dinv, const int *system_rows, int nstages,
                                    const double *tableau, const double *q, const double *dt, double gamma,
                                    const double *weights, double tolerance, int max_iterations, double *x, double *f,
                                    double *res, double *explicit_, double *K, int *stage, int *iters, int *nlfail,
                                    int *done, double *work, int *n_left, int *out_rounds, void *stream);
/* The ASYNCHRONOUS batch integrator (spitfire_b200/time/batched.py: integrate_batch_async): the members of a batch
 * advance independently of each other across steps. state[m]: 0 = not taking part (waiting for host control or new
 * factors, or finished), 1 = BEGIN (this round's right-hand side is f(q): it becomes K[0] and the member enters stage 1),
 * 2 = inside the implicit stages; a member that completes its last stage gets state 0 and stage == nstages.
 * gb_async_round_kernels launches one kernel of a round: phase 0 update, 1 tail, 2 start (fn = int start flags, dx =
 * step sizes, both device), 3 accept (fn = dq, dx = stats, max_iterations = clip flag).
 * gb_flamelet_async_tick_batch does one "tick" on all F members: start the members flagged in host_start with the
 * step sizes host_dt; rounds {solve_inv, update, flamelet rhs, tail} until a member has completed its stages, no member
 * is active or max_rounds is reached (state / stage copied to host_state / host_stage after each round); then, if
 * members completed, the embedded error estimate (b, bh: HOST arrays; dq, stats: device), q <- q + dq (clipped at zero
 * if clip_negative) for the completed members whose update is finite, and host_stats [3][F], host_nlfail [F] and the
 * rows of the completed members in host_q [F][ndof] are filled. Returns the rounds taken or a negative error code.
 * newton_its[m] accumulates the member's Newton iterations; start_d / dtin_d: device work arrays of F ints / doubles.
 * host_members (may be NULL) / n_members: a superset of the members that are inside a step during this tick; the
 * flamelet right-hand side of a round is evaluated for those only (at most 64; otherwise, or if NULL, for all F). */
int gb_async_round_kernels(int n, int ndof, int nstages, const double *tableau, int max_iterations, int phase,
                           const double *fn, double *xn, const double *dx, const double *q, const double *dt,
                           double gamma, const double *weights, double tolerance, double *x, double *f, double *res,
                           double *explicit_, double *K, int *state, int *stage, int *iters, int *nlfail,
                           int *newton_its, void *stream);
int gb_flamelet_async_tick_batch(gb_mech *m, int F, const gb_flamelet_params *prm, const double *d_factors,
                                 const double *l_values, const double *dinv, int nstages, const double *tableau,
                                 const double *b, const double *bh, double *q, double *dt, double gamma,
                                 const double *weights, double tolerance, int max_iterations, int clip_negative, double *x,
                                 double *f, double *res, double *explicit_, double *K, int *state, int *stage, int *iters,
                                 int *nlfail, int *newton_its, double *work, double *dq, double *stats, int *start_d,
                                 double *dtin_d, int max_rounds, const int *host_start, const double *host_dt,
                                 int *host_state, int *host_stage, double *host_stats, int *host_nlfail, double *host_q,
                                 const int *host_members, int n_members, void *stream);
/* Status "> 0 = number of members with non-finite output" (SURVEY 8(b)): flags_out[m] (device, may be NULL) = 1 if row m of
 * a [n][len_a] -- or of b [n][len_b], if given -- holds an Inf or NaN. Synchronises the stream and returns the number of
 * such members (>= 0) or a negative error code. The asynchronous *_batch entry points cannot report it themselves; the
 * synchronous *_host entry points of the reactor and flamelet right-hand sides / Jacobians return it as their status.
 * Replaces the NaN / Inf scans the reference does in Python (flamelet.py:1435, 1443). */
int gb_count_nonfinite_members_batch(int n, long len_a, const double *a, long len_b, const double *b, int *flags_out,
                                     void *stream);
/* q <- q + dq (clipped at zero if clip_negative) for the members with accept[m] != 0, in place */
int gb_accept_step_batch(int n, int ndof, const double *dq, const int *accept, int clip_negative, double *q,
                         void *stream);

/* ---- instrumentation ------------------------------------------------------------------------------------ */
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long gb_kernel_launch_count(void);
/* host-clock accounting of gb_flamelet_async_tick_batch since the last reset: out[0] ticks, [1] rounds, [2] seconds
 * inside the calls, [3] seconds inside the round loops, [4] / [5] seconds until the solve / the right-hand side of a
 * round had finished (only with GB_TICK_PROFILE set in the environment, which adds two synchronisations per round);
 * out: 8 doubles. tools/dev/dev_tick_stats.py */
void gb_debug_tick_stats(double *out, int reset);
/* name of the build (arch, flags), for logs */
const char *gb_build_info(void);
/* FP64 vector peak of the current device, measured with a dependent-free DFMA (kind 0) or DMUL+DADD (kind 1)
 * micro-benchmark: Tflop/s (2 flops per multiply-add) and thread-level FP64 instructions per clock and SM.
 * bench.py's second roofline bound. No reference counterpart. */
int gb_measure_fp64_peak(int kind, double *out_tflops, double *out_inst_per_clk_sm);
/* dependent-issue latency of the FP64 pipe: cycles per instruction of one warp running one dependent chain of DFMA
 * (kind 2), DADD (3) or DMUL (4) -- the number every latency-bound kernel of this path is designed around */
int gb_measure_fp64_latency(int kind, double *out_cycles);

#ifdef __cplusplus
}
#endif
#endif /* GRIFFON_B200_H */
