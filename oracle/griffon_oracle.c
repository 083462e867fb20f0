/*
 * griffon_oracle.c -- TEST INFRASTRUCTURE: plain-C restatement of the reference's Griffon hot path.
 *
 * This is the CPU oracle ("port") that the CUDA product is checked against. It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning: every function below is checked (tests/test_oracle_vs_reference.py) to be BIT-IDENTICAL to the
 * unmodified reference C++ (oracle/_ref/libref_griffon.so, compiled from /root/reference) on the reference's own
 * fixture mechanisms, and through it to the reference's gold files (tests/golden/). Expression order and
 * association follow the reference line by line for that reason; compile with -ffp-contract=off.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/src/spitfire/griffon/). Third-party arithmetic: LAPACK dgetrf/dgetrs/dgeev
 * (blas_lapack_kernels.h:84-180), taken -- like the reference build here -- from SciPy's bundled OpenBLAS.
 *
 * Not restated (out of the hot path): the inexact no-TBAF sensitivities
 * (option 1 is mapped to the exact option 0), isochoric reactors, 2-D flamelets, block-Jacobi/GS helpers.
 */
#include "griffon_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NSR 8 /* combustion_kernels.h:518 MechanismData<8,15> */
#define NCP 15

enum { CP_UNKNOWN = 0, CP_CONST = 1, CP_NASA7 = 2, CP_NASA9 = 3 };                 /* combustion_kernels.h:43 */
enum { RT_SIMPLE = 1, RT_THIRD_BODY = 2, RT_LINDEMANN = 3, RT_TROE = 4 };          /* :60 */
enum { KF_CONSTANT, KF_LINEAR, KF_QUADRATIC, KF_RECIPROCAL, KF_ARRHENIUS };        /* :75 */
enum { TR_NONE, TR_T1, TR_T2, TR_T12, TR_T3, TR_T13, TR_T23, TR_T123 };            /* :88 */
enum { RO_ONE, RO_TWO, RO_ONE_ONE, RO_TWO_ONE, RO_ONE_TWO, RO_ONE_ONE_ONE, RO_OTHER }; /* :105 */

typedef struct
{
  int rc_idx[NSR], rc_st[NSR];
  double rc_invmw[NSR];
  int n_rc;
  int pd_idx[NSR], pd_st[NSR]; /* product stoich stored NEGATIVE, chemistry_setup.cpp:391 */
  double pd_invmw[NSR];
  int n_pd;
  int net_idx[NSR], net_st[NSR];
  double net_mw[NSR];
  int n_net;
  int *tb_idx;
  double *tb_eff; /* invMW_i * (eff_i - default), chemistry_setup.cpp:416 */
  int n_tb;
  int has_orders;
  int sp_idx[NSR];
  double sp_order[NSR], sp_invmw[NSR];
  int sp_nonzero[NSR];
  int n_sp;
  double base_eff, kf[3], kp[3], troe[4];
  int reversible, type, kform, troeform, fwd_order, rev_order;
  int sum_stoich, sum_rc_stoich, sum_pd_stoich;
} rxn_t;

struct go_mech
{
  int n_elem_mw;
  char **elem_mw_names;
  double *elem_mw;
  int n_elements;
  char **element_names;
  int ns;
  char **species_names;
  double *mw, *invmw;
  double (*cpc)[NCP];
  double *tmin, *tmax;
  int *cptype;
  double p_ref, T_ref, Ru;
  int nr, cap_r;
  rxn_t *rx;
  int unsupported;
  double **c9; /* NASA9: per species {nregions, (Tlo, Thi, a0..a8 [a2.. times R]) * nregions}, chemistry_setup.cpp:132-152 */
};

const char *go_kind(void) { return "port"; }

/* ----------------------------------------------------------------------------------------------------------------
 * mechanism construction -- chemistry_setup.cpp
 * -------------------------------------------------------------------------------------------------------------- */
static char *dupstr(const char *s)
{
  char *d = (char *)malloc(strlen(s) + 1);
  strcpy(d, s);
  return d;
}

go_mech *go_mech_create(void) { return (go_mech *)calloc(1, sizeof(go_mech)); }

void go_mech_destroy(go_mech *m)
{
  if (!m)
    return;
  for (int i = 0; i < m->n_elem_mw; ++i)
    free(m->elem_mw_names[i]);
  for (int i = 0; i < m->n_elements; ++i)
    free(m->element_names[i]);
  for (int i = 0; i < m->ns; ++i)
    free(m->species_names[i]);
  for (int r = 0; r < m->nr; ++r)
  {
    free(m->rx[r].tb_idx);
    free(m->rx[r].tb_eff);
  }
  free(m->elem_mw_names);
  free(m->elem_mw);
  free(m->element_names);
  free(m->species_names);
  free(m->mw);
  free(m->invmw);
  free(m->cpc);
  free(m->tmin);
  free(m->tmax);
  free(m->cptype);
  if (m->c9)
    for (int i = 0; i < m->ns; ++i)
      free(m->c9[i]);
  free(m->c9);
  free(m->rx);
  free(m);
}

int go_mech_set_ref_pressure(go_mech *m, double p) /* chemistry_setup.cpp:64 */
{
  m->p_ref = p;
  return 0;
}
int go_mech_set_ref_temperature(go_mech *m, double T) /* :74 */
{
  m->T_ref = T;
  return 0;
}
int go_mech_set_gas_constant(go_mech *m, double Ru) /* :69 */
{
  m->Ru = Ru;
  return 0;
}
int go_mech_set_element_mw(go_mech *m, const char *e, double mw) /* :21 */
{
  for (int i = 0; i < m->n_elem_mw; ++i)
    if (!strcmp(m->elem_mw_names[i], e))
    {
      m->elem_mw[i] = mw;
      return 0;
    }
  m->elem_mw_names = (char **)realloc(m->elem_mw_names, sizeof(char *) * (m->n_elem_mw + 1));
  m->elem_mw = (double *)realloc(m->elem_mw, sizeof(double) * (m->n_elem_mw + 1));
  m->elem_mw_names[m->n_elem_mw] = dupstr(e);
  m->elem_mw[m->n_elem_mw++] = mw;
  return 0;
}
int go_mech_add_element(go_mech *m, const char *e) /* :26 */
{
  for (int i = 0; i < m->n_elements; ++i)
    if (!strcmp(m->element_names[i], e))
      return 0;
  m->element_names = (char **)realloc(m->element_names, sizeof(char *) * (m->n_elements + 1));
  m->element_names[m->n_elements++] = dupstr(e);
  return 0;
}

static int species_index(const go_mech *m, const char *name)
{
  for (int i = 0; i < m->ns; ++i)
    if (!strcmp(m->species_names[i], name))
      return i;
  return -1;
}

/* indices that visit `names` in std::map<std::string,...> order (byte-wise lexicographic) */
static void sorted_order(int n, const char *const *names, int *order)
{
  for (int i = 0; i < n; ++i)
    order[i] = i;
  for (int i = 1; i < n; ++i)
  {
    int k = order[i], j = i - 1;
    while (j >= 0 && strcmp(names[order[j]], names[k]) > 0)
    {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = k;
  }
}

int go_mech_add_species(go_mech *m, const char *name, int n_atoms, const char *const *atom_names,
                        const double *atom_counts) /* :36-62 */
{
  if (species_index(m, name) >= 0)
    return -1;
  int order[64];
  if (n_atoms > 64)
    return -1;
  sorted_order(n_atoms, atom_names, order);
  double mw = 0.;
  for (int a = 0; a < n_atoms; ++a)
  {
    const char *an = atom_names[order[a]];
    int found = 0;
    for (int i = 0; i < m->n_elements; ++i)
      if (!strcmp(m->element_names[i], an))
        found = 1;
    if (!found)
      return -1;
    int e = -1;
    for (int i = 0; i < m->n_elem_mw; ++i)
      if (!strcmp(m->elem_mw_names[i], an))
        e = i;
    if (e < 0)
      return -1;
    mw += m->elem_mw[e] * atom_counts[order[a]];
  }
  m->species_names = (char **)realloc(m->species_names, sizeof(char *) * (m->ns + 1));
  m->mw = (double *)realloc(m->mw, sizeof(double) * (m->ns + 1));
  m->invmw = (double *)realloc(m->invmw, sizeof(double) * (m->ns + 1));
  m->species_names[m->ns] = dupstr(name);
  m->mw[m->ns] = mw;
  m->invmw[m->ns] = 1. / mw;
  ++m->ns;
  return 0;
}

int go_mech_resize_heat_capacity_data(go_mech *m) /* :79 */
{
  free(m->cpc);
  free(m->tmin);
  free(m->tmax);
  free(m->cptype);
  m->cpc = (double(*)[NCP])calloc(m->ns > 0 ? m->ns : 1, sizeof(double[NCP]));
  m->tmin = (double *)calloc(m->ns > 0 ? m->ns : 1, sizeof(double));
  m->tmax = (double *)calloc(m->ns > 0 ? m->ns : 1, sizeof(double));
  m->cptype = (int *)calloc(m->ns > 0 ? m->ns : 1, sizeof(int));
  if (m->c9)
    for (int i = 0; i < m->ns; ++i)
      free(m->c9[i]);
  free(m->c9);
  m->c9 = (double **)calloc(m->ns > 0 ? m->ns : 1, sizeof(double *));
  return 0;
}

int go_mech_add_const_cp(go_mech *m, const char *s, double Tmin, double Tmax, double T0, double h0, double s0,
                         double cp) /* :89-100 */
{
  const int i = species_index(m, s);
  if (i < 0 || !m->cpc)
    return -1;
  m->cptype[i] = CP_CONST;
  m->tmin[i] = Tmin;
  m->tmax[i] = Tmax;
  m->cpc[i][0] = T0;
  m->cpc[i][1] = h0;
  m->cpc[i][2] = s0;
  m->cpc[i][3] = cp;
  return 0;
}

int go_mech_add_nasa7_cp(go_mech *m, const char *s, double Tmin, double Tmid, double Tmax, const double *low7,
                         const double *high7) /* :102-130 */
{
  const int i = species_index(m, s);
  if (i < 0 || !m->cpc)
    return -1;
  m->cptype[i] = CP_NASA7;
  m->tmin[i] = Tmin;
  m->tmax[i] = Tmax;
  double *c = m->cpc[i];
  c[0] = Tmid;
  const double R = m->Ru;
  for (int k = 0; k < 7; ++k)
    c[1 + k] = high7[k] * R;
  for (int k = 0; k < 7; ++k)
    c[8 + k] = low7[k] * R;
  c[2] /= 2.;
  c[3] /= 6.;
  c[4] /= 12.;
  c[5] /= 20.;
  c[9] /= 2.;
  c[10] /= 6.;
  c[11] /= 12.;
  c[12] /= 20.;
  return 0;
}

int go_mech_add_nasa9_cp(go_mech *m, const char *s, double Tmin, double Tmax, int n, const double *c) /* :132-152 */
{
  const int i = species_index(m, s);
  if (i < 0 || !m->cpc || n < 1)
    return -1;
  const int nregions = (int)c[0];
  if (n < 1 + 11 * nregions)
    return -1;
  m->cptype[i] = CP_NASA9;
  m->tmin[i] = Tmin;
  m->tmax[i] = Tmax;
  free(m->c9[i]);
  m->c9[i] = (double *)calloc((size_t)n, sizeof(double));
  const double R = m->Ru;
  m->c9[i][0] = c[0];
  for (int k = 0; k < nregions; ++k)
    for (int j = 0; j < 11; ++j)
      m->c9[i][1 + k * 11 + j] = (j < 2 ? 1.0 : R) * c[1 + k * 11 + j];
  return 0;
}

/* NASA9 polynomial pieces as written at the cited lines; a = region coefficients a0..a8 */
static double n9_cp(const double *a, double t, double invT)
{ /* thermodynamics_kernels.cpp:99, 106, 118 */
  return invT * (a[1] + invT * a[0]) + a[2] + t * (a[3] + t * (a[4] + t * (a[5] + t * a[6])));
}
static double n9_h_over_t(const double *a, double t, double invT, double logT)
{ /* :314, 324, 337 (the factor invMW * T is applied by the caller) */
  return invT * (a[7] + logT * a[1] - a[0] * invT) + a[2] +
         t * (0.5 * a[3] + t * (0.3333333333333333 * a[4] + t * (0.25 * a[5] + t * 0.2 * a[6])));
}
static double n9_gibbs(const double *a, double T, double invT, double logT)
{ /* chemistry_kernels.cpp:83, rates_sensitivities_exact.cpp:110 */
  return a[7] - 0.5 * a[0] * invT + a[1] * (logT + 1.0) -
         T * (a[2] * (logT - 1.0) + a[8] +
              T * (0.5 * a[3] + T * (0.1666666666666666 * a[4] + T * (0.0833333333333333 * a[5] + T * 0.05 * a[6]))));
}

/* ReactionRateData::finalize, chemistry_setup.cpp:460-732 */
static int finalize_reaction(const go_mech *m, rxn_t *x)
{
  /* net species: std::map<int, (stoich, invmw)> keyed by species index, :463-541 */
  int idx[2 * NSR], st[2 * NSR], n = 0;
  double inv[2 * NSR];
  x->sum_stoich = 0;
  for (int i = 0; i < x->n_rc; ++i)
  {
    x->sum_stoich += x->rc_st[i];
    int f = -1;
    for (int k = 0; k < n; ++k)
      if (idx[k] == x->rc_idx[i])
        f = k;
    if (f < 0) /* std::map::insert does nothing on an existing key */
    {
      idx[n] = x->rc_idx[i];
      st[n] = x->rc_st[i];
      inv[n] = x->rc_invmw[i];
      ++n;
    }
  }
  for (int i = 0; i < x->n_pd; ++i)
  {
    x->sum_stoich += x->pd_st[i];
    int f = -1;
    for (int k = 0; k < n; ++k)
      if (idx[k] == x->pd_idx[i])
        f = k;
    if (f < 0)
    {
      idx[n] = x->pd_idx[i];
      st[n] = x->pd_st[i];
      inv[n] = x->pd_invmw[i];
      ++n;
    }
    else
      st[f] += x->pd_st[i];
  }
  /* ascending species index */
  for (int i = 1; i < n; ++i)
  {
    int ki = idx[i], ks = st[i], j = i - 1;
    double kv = inv[i];
    while (j >= 0 && idx[j] > ki)
    {
      idx[j + 1] = idx[j];
      st[j + 1] = st[j];
      inv[j + 1] = inv[j];
      --j;
    }
    idx[j + 1] = ki;
    st[j + 1] = ks;
    inv[j + 1] = kv;
  }
  x->n_net = 0;
  for (int k = 0; k < n; ++k)
    if (abs(st[k]) > 0) /* |stoich| > 1e-14 on integers */
      ++x->n_net;
  if (x->n_net < 2 || x->n_net > 8)
    return -2; /* :499-527 throws */
  int q = 0;
  for (int k = 0; k < n; ++k)
    if (abs(st[k]) > 0)
    {
      x->net_idx[q] = idx[k];
      x->net_st[q] = st[k];
      x->net_mw[q] = 1. / inv[k];
      ++q;
    }
  (void)m;

  /* forward / reverse order classification, :604-672 */
  switch (x->n_rc)
  {
  case 1:
    x->fwd_order = x->rc_st[0] == 1 ? RO_ONE : (x->rc_st[0] == 2 ? RO_TWO : RO_OTHER);
    break;
  case 2:
    if (x->rc_st[0] == 1 && x->rc_st[1] == 1)
      x->fwd_order = RO_ONE_ONE;
    else if (x->rc_st[0] == 1 && x->rc_st[1] == 2)
      x->fwd_order = RO_ONE_TWO;
    else if (x->rc_st[0] == 2 && x->rc_st[1] == 1)
      x->fwd_order = RO_TWO_ONE;
    else
      x->fwd_order = RO_OTHER;
    break;
  case 3:
    x->fwd_order = (x->rc_st[0] == 1 && x->rc_st[1] == 1 && x->rc_st[2] == 1) ? RO_ONE_ONE_ONE : RO_OTHER;
    break;
  default:
    x->fwd_order = RO_OTHER;
  }
  x->sum_rc_stoich = 0;
  for (int s = 0; s < x->n_rc; ++s)
    x->sum_rc_stoich += abs(x->rc_st[s]);
  switch (x->n_pd)
  {
  case 1:
    x->rev_order = x->pd_st[0] == -1 ? RO_ONE : (x->pd_st[0] == -2 ? RO_TWO : RO_OTHER);
    break;
  case 2:
    if (x->pd_st[0] == -1 && x->pd_st[1] == -1)
      x->rev_order = RO_ONE_ONE;
    else if (x->pd_st[0] == -1 && x->pd_st[1] == -2)
      x->rev_order = RO_ONE_TWO;
    else if (x->pd_st[0] == -2 && x->pd_st[1] == -1)
      x->rev_order = RO_TWO_ONE;
    else
      x->rev_order = RO_OTHER;
    break;
  case 3:
    x->rev_order = (x->pd_st[0] == -1 && x->pd_st[1] == -1 && x->pd_st[2] == -1) ? RO_ONE_ONE_ONE : RO_OTHER;
    break;
  default:
    x->rev_order = RO_OTHER;
  }
  x->sum_pd_stoich = 0;
  for (int s = 0; s < x->n_pd; ++s)
    x->sum_pd_stoich += abs(x->pd_st[s]);

  /* rate-constant temperature form, :674-688 */
  if (fabs(x->kf[2]) < 1.e-6)
  {
    if (fabs(x->kf[1]) < 1.e-6)
      x->kform = KF_CONSTANT;
    else if (fabs(x->kf[1] - 1) < 1.e-6)
      x->kform = KF_LINEAR;
    else if (fabs(x->kf[1] - 2) < 1.e-6)
      x->kform = KF_QUADRATIC;
    else if (fabs(x->kf[1] + 1) < 1.e-6)
      x->kform = KF_RECIPROCAL;
    else
      x->kform = KF_ARRHENIUS;
  }
  else
    x->kform = KF_ARRHENIUS;

  /* Troe terms present, :690-730 (names refer to troeParams indices 1,2,3 = T3,T1,T2 of the Troe form) */
  x->troeform = TR_NONE;
  if (x->type == RT_TROE)
  {
    const int a = fabs(x->troe[1]) > 1.e-8, b = fabs(x->troe[2]) > 1.e-8, c = fabs(x->troe[3]) > 1.e-8;
    if (a)
      x->troeform = b ? (c ? TR_T123 : TR_T12) : (c ? TR_T13 : TR_T1);
    else
      x->troeform = b ? (c ? TR_T23 : TR_T2) : (c ? TR_T3 : TR_NONE);
  }
  return 0;
}

int go_mech_add_reaction(go_mech *m, int type, int reversible, int n_reactants, const char *const *reactant_names,
                         const int *reactant_stoich, int n_products, const char *const *product_names,
                         const int *product_stoich, double fwd_A, double fwd_b, double fwd_Ea, int n_eff,
                         const char *const *eff_names, const double *eff_values, double default_eff, double flf_A,
                         double flf_b, double flf_Ea, const double *troe4, int n_orders,
                         const char *const *order_names, const double *order_values) /* :156-345 */
{
  if (type < RT_SIMPLE || type > RT_TROE || n_reactants > NSR || n_products > NSR || n_orders > NSR)
    return -1;
  rxn_t x;
  memset(&x, 0, sizeof(x));
  x.type = type;
  x.reversible = reversible != 0;
  x.has_orders = n_orders > 0;
  x.kf[0] = fwd_A;
  x.kf[1] = fwd_b;
  x.kf[2] = fwd_Ea;
  int order[64];
  /* set_reactants_or_products, :374-404: iterate the std::map in name order */
  sorted_order(n_reactants, reactant_names, order);
  for (int i = 0; i < n_reactants; ++i)
  {
    const int s = species_index(m, reactant_names[order[i]]);
    if (s < 0)
      return -1;
    x.rc_idx[i] = s;
    x.rc_st[i] = reactant_stoich[order[i]];
    x.rc_invmw[i] = m->invmw[s];
  }
  x.n_rc = n_reactants;
  sorted_order(n_products, product_names, order);
  for (int i = 0; i < n_products; ++i)
  {
    const int s = species_index(m, product_names[order[i]]);
    if (s < 0)
      return -1;
    x.pd_idx[i] = s;
    x.pd_st[i] = -product_stoich[order[i]];
    x.pd_invmw[i] = m->invmw[s];
  }
  x.n_pd = n_products;
  if (n_orders > 0)
  { /* set_special_orders, :420-435 */
    sorted_order(n_orders, order_names, order);
    for (int i = 0; i < n_orders; ++i)
    {
      const int s = species_index(m, order_names[order[i]]);
      if (s < 0)
        return -1;
      x.sp_idx[i] = s;
      x.sp_invmw[i] = m->invmw[s];
      x.sp_order[i] = order_values[order[i]];
      x.sp_nonzero[i] = fabs(x.sp_order[i]) > 1.e-12;
    }
    x.n_sp = n_orders;
  }
  if (type != RT_SIMPLE)
  { /* set_three_body_efficiencies, :405-419 */
    if (n_eff > 64)
      return -1;
    x.base_eff = default_eff;
    x.tb_idx = (int *)malloc(sizeof(int) * (n_eff > 0 ? n_eff : 1));
    x.tb_eff = (double *)malloc(sizeof(double) * (n_eff > 0 ? n_eff : 1));
    sorted_order(n_eff, eff_names, order);
    for (int i = 0; i < n_eff; ++i)
    {
      const int s = species_index(m, eff_names[order[i]]);
      if (s < 0)
      {
        free(x.tb_idx);
        free(x.tb_eff);
        return -1;
      }
      x.tb_idx[i] = s;
      x.tb_eff[i] = m->invmw[s] * (eff_values[order[i]] - x.base_eff);
    }
    x.n_tb = n_eff;
  }
  if (type == RT_LINDEMANN || type == RT_TROE)
  {
    x.kp[0] = flf_A;
    x.kp[1] = flf_b;
    x.kp[2] = flf_Ea;
  }
  if (type == RT_TROE)
    for (int k = 0; k < 4; ++k)
      x.troe[k] = troe4 ? troe4[k] : 0.;
  const int rc = finalize_reaction(m, &x);
  if (rc)
  {
    free(x.tb_idx);
    free(x.tb_eff);
    return rc;
  }
  if (m->nr == m->cap_r)
  {
    m->cap_r = m->cap_r ? 2 * m->cap_r : 64;
    m->rx = (rxn_t *)realloc(m->rx, sizeof(rxn_t) * m->cap_r);
  }
  m->rx[m->nr++] = x;
  return 0;
}

int go_mech_n_species(const go_mech *m) { return m->ns; }
int go_mech_n_reactions(const go_mech *m) { return m->nr; }
int go_mech_molecular_weights(const go_mech *m, double *out)
{
  for (int i = 0; i < m->ns; ++i)
    out[i] = m->mw[i];
  return 0;
}

/* ----------------------------------------------------------------------------------------------------------------
 * thermodynamics -- thermodynamics_kernels.cpp, combustion_kernels.h:381-387, 505-535
 * -------------------------------------------------------------------------------------------------------------- */
static double inner_product(int n, const double *x, const double *y) /* blas_lapack_kernels.h:41-48 */
{
  double d = 0.;
  for (int i = 0; i < n; ++i)
    d += y[i] * x[i];
  return d;
}

static void extract_y(const go_mech *m, const double *ynm1, double *y) /* combustion_kernels.h:505-515 */
{
  const int ns = m->ns;
  y[ns - 1] = 1.;
  for (int j = 0; j < ns - 1; ++j)
  {
    y[j] = ynm1[j];
    y[ns - 1] -= y[j];
  }
}

double go_mixture_molecular_weight(const go_mech *m, const double *y) /* combustion_kernels.h:381-387 */
{
  return 1. / inner_product(m->ns, y, m->invmw);
}

void go_mole_fractions(const go_mech *m, const double *y, double *x) /* thermodynamics_kernels.cpp:17-27 */
{
  const double mmw = go_mixture_molecular_weight(m, y);
  for (int i = 0; i < m->ns; ++i)
    x[i] = y[i] * mmw * m->invmw[i];
}

static double density_from(const go_mech *m, double p, double T, double mmw) /* combustion_kernels.h:526-530 */
{
  return p * mmw / (T * m->Ru);
}

double go_ideal_gas_density(const go_mech *m, double p, double T, const double *y)
{
  return density_from(m, p, T, go_mixture_molecular_weight(m, y));
}

double go_ideal_gas_pressure(const go_mech *m, double rho, double T, const double *y) /* :531-535 */
{
  return rho * T * m->Ru / go_mixture_molecular_weight(m, y);
}

/* thermodynamics_kernels.cpp:45-131 (CONST and NASA7 branches). NOTE out_cpspecies may alias y (species_cp). */
static void cp_mix_and_species(const go_mech *m, double t, const double *y, double *out_cpmix, double *out_cpi)
{
  *out_cpmix = 0.;
  for (int i = 0; i < m->ns; ++i)
  {
    const double *c = m->cpc[i];
    const double maxT = m->tmax[i], minT = m->tmin[i], iw = m->invmw[i];
    const double yi = y[i];
    if (m->cptype[i] == CP_CONST)
      out_cpi[i] = iw * c[3];
    else if (m->cptype[i] == CP_NASA7)
    {
      if (t <= c[0] && t >= minT)
        out_cpi[i] = iw * (c[8] + t * (2. * c[9] + t * (6. * c[10] + t * (12. * c[11] + 20. * t * c[12]))));
      else if (t > c[0] && t <= maxT)
        out_cpi[i] = iw * (c[1] + t * (2. * c[2] + t * (6. * c[3] + t * (12. * c[4] + 20. * t * c[5]))));
      else if (t < minT)
        out_cpi[i] =
            iw * (c[8] + minT * (2. * c[9] + minT * (6. * c[10] + minT * (12. * c[11] + 20. * minT * c[12]))));
      else
        out_cpi[i] = iw * (c[1] + maxT * (2. * c[2] + maxT * (6. * c[3] + maxT * (12. * c[4] + 20. * maxT * c[5]))));
    }
    else if (m->cptype[i] == CP_NASA9)
    { /* :91-124: frozen outside [Tmin, Tmax]; inside, the region with Tlo <= t < Thi (t == Tmax matches none) */
      const double *c9 = m->c9[i];
      const int nregions = (int)c9[0];
      if (t < minT)
        out_cpi[i] = iw * n9_cp(c9 + 3, minT, 1. / minT);
      else if (t > maxT)
        out_cpi[i] = iw * n9_cp(c9 + 1 + (nregions - 1) * 11 + 2, maxT, 1. / maxT);
      else
      {
        int found = 0;
        for (int k = 0; k < nregions && !found; ++k)
          if (t >= c9[1 + k * 11] && t < c9[1 + k * 11 + 1])
          {
            out_cpi[i] = iw * n9_cp(c9 + 1 + k * 11 + 2, t, 1. / t);
            found = 1;
          }
        if (!found)
          continue;
      }
    }
    else
      continue;
    *out_cpmix += yi * out_cpi[i];
  }
}

double go_cp_mix(const go_mech *m, double T, const double *y) /* :133-141 */
{
  double cpmix = 0.;
  double cpi[m->ns];
  cp_mix_and_species(m, T, y, &cpmix, cpi);
  return cpmix;
}

void go_species_cp(const go_mech *m, double T, double *out) /* :143-147; y aliases out as in the reference */
{
  double garbage;
  double tmp[m->ns];
  memcpy(tmp, out, sizeof(double) * m->ns);
  cp_mix_and_species(m, T, tmp, &garbage, out);
}

double go_cv_mix(const go_mech *m, double T, const double *y) /* :149-154 */
{
  return go_cp_mix(m, T, y) - m->Ru / go_mixture_molecular_weight(m, y);
}

void go_species_cv(const go_mech *m, double T, double *out) /* :156-167 */
{
  go_species_cp(m, T, out);
  for (int i = 0; i < m->ns; ++i)
    out[i] -= m->Ru * m->invmw[i];
}

void go_cp_sens_T(const go_mech *m, double t, const double *y, double *out_mix, double *out_spec) /* :183-260 */
{
  *out_mix = 0.;
  for (int i = 0; i < m->ns; ++i)
  {
    const double *c = m->cpc[i];
    const double minT = m->tmin[i], maxT = m->tmax[i], iw = m->invmw[i];
    if (m->cptype[i] == CP_CONST)
    {
      out_spec[i] = 0.;
      *out_mix = 0.; /* sic, :202 resets the running mixture sum */
    }
    else if (m->cptype[i] == CP_NASA7)
    {
      if (t <= c[0] && t >= minT)
      {
        out_spec[i] = iw * ((2. * c[9] + t * (12. * c[10] + t * (36. * c[11] + 80. * t * c[12]))));
        *out_mix += y[i] * out_spec[i];
      }
      else if (t > c[0] && t <= maxT)
      {
        out_spec[i] = iw * ((2. * c[2] + t * (12. * c[3] + t * (36. * c[4] + 80. * t * c[5]))));
        *out_mix += y[i] * out_spec[i];
      }
      else
      {
        out_spec[i] = 0.;
        *out_mix += 0.;
      }
    }
    else if (m->cptype[i] == CP_NASA9)
    { /* :229-253 */
      const double *c9 = m->c9[i];
      const int nregions = (int)c9[0];
      if (t < minT || t > maxT)
      {
        out_spec[i] = 0.;
        *out_mix += 0.;
      }
      else
      {
        const double invT = 1. / t;
        for (int k = 0; k < nregions; ++k)
          if (t >= c9[1 + k * 11] && t < c9[1 + k * 11 + 1])
          {
            const double *a = c9 + 1 + k * 11 + 2;
            out_spec[i] = iw * (-invT * invT * (a[1] + invT * 2.0 * a[0]) + a[3] +
                                t * (2.0 * a[4] + t * (3.0 * a[5] + t * 4.0 * a[6])));
            *out_mix += y[i] * out_spec[i];
            break;
          }
      }
    }
  }
}

void go_species_enthalpies(const go_mech *m, double temp, double *out) /* :262-351 */
{
  for (int i = 0; i < m->ns; ++i)
  {
    const double *c = m->cpc[i];
    const double minT = m->tmin[i], maxT = m->tmax[i], iw = m->invmw[i];
    if (m->cptype[i] == CP_CONST)
      out[i] = iw * (c[1] + c[3] * (temp - c[0]));
    else if (m->cptype[i] == CP_NASA7)
    {
      if (temp <= c[0] && temp >= minT)
        out[i] = iw * (c[13] +
                       temp * (c[8] + temp * (c[9] + temp * (2. * c[10] + temp * (3. * c[11] + temp * 4. * c[12])))));
      else if (temp > c[0] && temp <= maxT)
        out[i] =
            iw * (c[6] + temp * (c[1] + temp * (c[2] + temp * (2. * c[3] + temp * (3. * c[4] + temp * 4. * c[5])))));
      else if (temp < minT)
        out[i] = iw * (c[13] + c[8] * temp +
                       minT * (2. * c[9] * temp +
                               minT * (3. * 2. * c[10] * temp - c[9] +
                                       minT * (4. * 3. * c[11] * temp - 2. * 2. * c[10] +
                                               minT * (5. * 4. * c[12] * temp - 3. * 3. * c[11] +
                                                       minT * -4. * 4. * c[12])))));
      else
        out[i] = iw * (c[6] + c[1] * temp +
                       maxT * (2. * c[2] * temp +
                               maxT * (3. * 2. * c[3] * temp - c[2] +
                                       maxT * (4. * 3. * c[4] * temp - 2. * 2. * c[3] +
                                               maxT * (5. * 4. * c[5] * temp - 3. * 3. * c[4] +
                                                       maxT * -4. * 4. * c[5])))));
    }
    else if (m->cptype[i] == CP_NASA9)
    { /* :305-347: linear extension with the frozen cp outside [Tmin, Tmax] */
      const double *c9 = m->c9[i];
      const int nregions = (int)c9[0];
      if (temp < minT)
      {
        const double invT = 1. / minT, logT = log(minT);
        const double *a = c9 + 3;
        const double hmin = iw * minT * n9_h_over_t(a, minT, invT, logT);
        const double cpmin = iw * n9_cp(a, minT, invT);
        out[i] = hmin + cpmin * (temp - minT);
      }
      else if (temp > maxT)
      {
        const double invT = 1. / maxT, logT = log(maxT);
        const double *a = c9 + 1 + (nregions - 1) * 11 + 2;
        const double hmax = iw * maxT * n9_h_over_t(a, maxT, invT, logT);
        const double cpmax = iw * n9_cp(a, maxT, invT);
        out[i] = hmax + cpmax * (temp - maxT);
      }
      else
      {
        const double invT = 1. / temp, logT = log(temp);
        for (int k = 0; k < nregions; ++k)
          if (temp >= c9[1 + k * 11] && temp < c9[1 + k * 11 + 1])
          {
            out[i] = iw * temp * n9_h_over_t(c9 + 1 + k * 11 + 2, temp, invT, logT);
            break;
          }
      }
    }
  }
}

void go_species_energies(const go_mech *m, double T, double *out) /* :353-364 */
{
  go_species_enthalpies(m, T, out);
  const double RT = m->Ru * T;
  for (int i = 0; i < m->ns; ++i)
    out[i] -= RT * m->invmw[i];
}

double go_enthalpy_mix(const go_mech *m, double T, const double *y) /* :366-373 */
{
  double hi[m->ns];
  go_species_enthalpies(m, T, hi);
  return inner_product(m->ns, hi, y);
}

double go_energy_mix(const go_mech *m, double T, const double *y) /* :375-382 */
{
  double ei[m->ns];
  go_species_energies(m, T, ei);
  return inner_product(m->ns, ei, y);
}

/* ----------------------------------------------------------------------------------------------------------------
 * production rates -- chemistry_kernels.cpp:35-463
 * -------------------------------------------------------------------------------------------------------------- */
#define ARRHENIUS(coef) ((coef)[0] * exp((coef)[1] * logT - (coef)[2] * invT)) /* chemistry_kernels.cpp:22 */

/* Gibbs function per species, chemistry_kernels.cpp:56-97 (no Tmin/Tmax clipping, only T <= Tmid) */
static void species_gibbs(const go_mech *m, double T, double logT, double *g)
{
  for (int n = 0; n < m->ns; ++n)
  {
    const double *c = m->cpc[n];
    if (m->cptype[n] == CP_NASA7)
    {
      if (T <= c[0])
        g[n] = c[13] + T * (c[8] - c[14] - c[8] * logT - T * (c[9] + T * (c[10] + T * (c[11] + T * c[12]))));
      else
        g[n] = c[6] + T * (c[1] - c[7] - c[1] * logT - T * (c[2] + T * (c[3] + T * (c[4] + T * c[5]))));
    }
    else if (m->cptype[n] == CP_CONST)
      g[n] = c[1] + c[3] * (T - c[0]) - T * (c[2] + c[3] * (logT - log(c[0])));
    else if (m->cptype[n] == CP_NASA9)
    { /* :74-88: strictly inside a region, Tlo < T < Thi (otherwise the reference leaves the value unset; 0 here) */
      const double *c9 = m->c9[n];
      const int nregions = (int)c9[0];
      g[n] = 0.;
      for (int k = 0; k < nregions; ++k)
        if (T < c9[1 + k * 11 + 1] && T > c9[1 + k * 11])
        {
          g[n] = n9_gibbs(c9 + 1 + k * 11 + 2, T, 1. / T, logT);
          break;
        }
    }
    else
      g[n] = 0.;
  }
}

/* third-body concentration sum as written in the switch(n_tb) ladders, chemistry_kernels.cpp:164-199:
 * returns baseEff*conc + rho*(((t0+t1)+t2)+...) for n_tb <= 8 and the `default:` association beyond. */
static double third_body_conc(const rxn_t *x, double conc, double rho, const double *y)
{
  const int n = x->n_tb;
  if (n == 0)
    return x->base_eff * conc;
  double s = x->tb_eff[0] * y[x->tb_idx[0]];
  const int n8 = n < 8 ? n : 8;
  for (int i = 1; i < n8; ++i)
    s = s + x->tb_eff[i] * y[x->tb_idx[i]];
  double mm = x->base_eff * conc + rho * (s);
  for (int i = 8; i < n; ++i)
    mm += rho * (x->tb_eff[i] * y[x->tb_idx[i]]);
  return mm;
}

static void production_rates_mmw(const go_mech *m, double T, double rho, double mmw, const double *y, double *w)
{
  const double invT = 1 / T;
  const double logT = log(T);
  const double conc = rho / mmw;
  const int ns = m->ns;
  const double invGasConstant = 1. / m->Ru;
  const double port = m->p_ref * invT * invGasConstant;

  for (int i = 0; i < ns; ++i)
    w[i] = 0.;
  double specG[ns];
  species_gibbs(m, T, logT, specG);

  double k = 0, kr, pr, logPrC, logFCent;
  for (int r = 0; r < m->nr; ++r)
  {
    const rxn_t *x = &m->rx[r];
    switch (x->kform)
    { /* :140-157 */
    case KF_CONSTANT:
      k = x->kf[0];
      break;
    case KF_LINEAR:
      k = x->kf[0] * T;
      break;
    case KF_QUADRATIC:
      k = x->kf[0] * T * T;
      break;
    case KF_RECIPROCAL:
      k = x->kf[0] * invT;
      break;
    default:
      k = ARRHENIUS(x->kf);
      break;
    }
    switch (x->type)
    {
    case RT_SIMPLE:
      break;
    case RT_THIRD_BODY: /* :163-200 */
      if (x->n_tb == 0)
        k *= x->base_eff * conc; /* note: k * baseEff first, then * conc (left-assoc of `k *= a * b` is k*(a*b)) */
      else
        k *= third_body_conc(x, conc, rho, y);
      break;
    case RT_LINDEMANN: /* :201-238 */
      if (x->n_tb == 0)
        k /= (1 + k / (ARRHENIUS(x->kp) * (x->base_eff * conc)));
      else
        k /= (1 + k / (ARRHENIUS(x->kp) * third_body_conc(x, conc, rho, y)));
      break;
    case RT_TROE: /* :239-314 */
    {
      if (x->n_tb == 0)
        pr = ARRHENIUS(x->kp) / k * (x->base_eff * conc);
      else
        pr = ARRHENIUS(x->kp) / k * third_body_conc(x, conc, rho, y);
      const double *troe = x->troe;
      switch (x->troeform)
      {
      case TR_T123:
        logFCent = log10((1 - troe[0]) * exp(-T / troe[1]) + troe[0] * exp(-T / troe[2]) + exp(-invT * troe[3]));
        break;
      case TR_T12:
        logFCent = log10((1 - troe[0]) * exp(-T / troe[1]) + troe[0] * exp(-T / troe[2]) + 0.0);
        break;
      case TR_T1:
        logFCent = log10((1 - troe[0]) * exp(-T / troe[1]) + 0.0 + 0.0);
        break;
      case TR_T23:
        logFCent = log10(0.0 + troe[0] * exp(-T / troe[2]) + exp(-invT * troe[3]));
        break;
      case TR_T2:
        logFCent = log10(0.0 + troe[0] * exp(-T / troe[2]) + 0.0);
        break;
      case TR_T13:
        logFCent = log10((1 - troe[0]) * exp(-T / troe[1]) + 0.0 + exp(-invT * troe[3]));
        break;
      case TR_T3:
        logFCent = log10(0.0 + 0.0 + exp(-invT * troe[3]));
        break;
      default:
        logFCent = NAN; /* the reference throws */
      }
#define CTROE (-0.4 - 0.67 * logFCent)
#define NTROE (0.75 - 1.27 * logFCent)
#define F1 (logPrC / (NTROE - 0.14 * logPrC))
      logPrC = log10(fmax(pr, 1.e-300)) + CTROE;
      k = k * pow(10, logFCent / (1 + F1 * F1)) * pr / (1 + pr);
#undef CTROE
#undef NTROE
#undef F1
      break;
    }
    }

    if (x->has_orders)
    { /* :325-337 */
      for (int i = 0; i < x->n_sp; ++i)
        if (x->sp_nonzero[i])
          k *= pow(fmax(y[x->sp_idx[i]] * rho * x->sp_invmw[i], 0.), x->sp_order[i]);
      kr = 0.;
    }
    else
    {
      kr = 0.;
      if (x->reversible && x->n_net >= 2 && x->n_net <= 8)
      { /* :341-368 */
        double gs = x->net_st[0] * specG[x->net_idx[0]];
        for (int i = 1; i < x->n_net; ++i)
          gs = gs + x->net_st[i] * specG[x->net_idx[i]];
        kr = k * exp(x->sum_stoich * log(port) - invT * invGasConstant * (gs));
      }
#define C_R(i) (y[x->rc_idx[i]] * rho * x->rc_invmw[i])
#define C_P(i) (y[x->pd_idx[i]] * rho * x->pd_invmw[i])
      switch (x->fwd_order)
      { /* :373-410 */
      case RO_ONE:
        k *= C_R(0);
        break;
      case RO_TWO:
        k *= C_R(0) * C_R(0);
        break;
      case RO_ONE_ONE:
        k *= C_R(0) * C_R(1);
        break;
      case RO_ONE_ONE_ONE:
        k *= C_R(0) * C_R(1) * C_R(2);
        break;
      case RO_TWO_ONE:
        k *= C_R(0) * C_R(0) * C_R(1);
        break;
      case RO_ONE_TWO:
        k *= C_R(0) * C_R(1) * C_R(1);
        break;
      default:
        for (int i = 0; i < x->n_rc; ++i)
          switch (x->rc_st[i])
          {
          case 1:
            k *= C_R(i);
            break;
          case 2:
            k *= C_R(i) * C_R(i);
            break;
          case 3:
            k *= C_R(i) * C_R(i) * C_R(i);
            break;
          }
      }
      if (x->reversible)
      { /* :412-452 */
        switch (x->rev_order)
        {
        case RO_ONE:
          kr *= C_P(0);
          break;
        case RO_TWO:
          kr *= C_P(0) * C_P(0);
          break;
        case RO_ONE_ONE:
          kr *= C_P(0) * C_P(1);
          break;
        case RO_ONE_ONE_ONE:
          kr *= C_P(0) * C_P(1) * C_P(2);
          break;
        case RO_TWO_ONE:
          kr *= C_P(0) * C_P(0) * C_P(1);
          break;
        case RO_ONE_TWO:
          kr *= C_P(0) * C_P(1) * C_P(1);
          break;
        default:
          for (int i = 0; i < x->n_pd; ++i)
            switch (x->pd_st[i])
            {
            case -1:
              kr *= C_P(i);
              break;
            case -2:
              kr *= C_P(i) * C_P(i);
              break;
            case -3:
              kr *= C_P(i) * C_P(i) * C_P(i);
              break;
            }
        }
      }
#undef C_R
#undef C_P
    }
    for (int i = 0; i < x->n_net; ++i) /* :457-461 */
      w[x->net_idx[i]] -= x->net_st[i] * x->net_mw[i] * (k - kr);
  }
}

void go_production_rates(const go_mech *m, double T, double rho, const double *y, double *out_w) /* :29-33 */
{
  production_rates_mmw(m, T, rho, go_mixture_molecular_weight(m, y), y, out_w);
}

/* ----------------------------------------------------------------------------------------------------------------
 * exact rate sensitivities -- rates_sensitivities_exact.cpp:33-1028 (== rates_sensitivities_sparse.cpp: the sparse
 * variant only skips adding exact zeros, rates_sensitivities_sparse.cpp:1026-1040)
 * -------------------------------------------------------------------------------------------------------------- */
#define ARR_SENS_OVER_K(coef) (invT * ((coef)[1] + (coef)[2] * invT)) /* rates_sensitivities_exact.cpp:25 */

/* product of concentrations of all reactants (or products) except `skip`, multiplied into *v in index order,
 * as the `default:` branches do (:406-423, 503-520, 687-707, 788-805) */
static void mult_other_conc(double *v, int n, const int *idx, const int *st, const double *invmw, int skip,
                            const double *y, double rho, int allow_pow)
{
  for (int i = 0; i < n; ++i)
  {
    if (i == skip)
      continue;
    const double c = y[idx[i]] * rho * invmw[i];
    switch (abs(st[i]))
    {
    case 1:
      *v *= c;
      break;
    case 2:
      *v *= c * c;
      break;
    case 3:
      *v *= c * c * c;
      break;
    default:
      if (allow_pow)
        *v *= pow(c, abs(st[i]));
      break;
    }
  }
}

/* d(prod C^nu)/dY_s for the special-cased orders: continues the left-to-right product `a * ...` exactly as the
 * expressions at :341-391 / :438-489 / :617-667 / :723-773 are written. `which` = position of the differentiated
 * species, a = k * rho * invmw already multiplied left to right. */
static double order_tail(double a, int order_kind, int which, const int *idx, const double *invmw, const double *y,
                         double rho)
{
#define CC(i) (y[idx[i]] * rho * invmw[i])
  switch (order_kind)
  {
  case RO_ONE:
    return a;
  case RO_TWO:
    return a * 2. * CC(which);
  case RO_ONE_ONE:
    return which == 0 ? a * CC(1) : a * CC(0);
  case RO_ONE_ONE_ONE:
    return which == 0 ? a * CC(1) * CC(2) : (which == 1 ? a * CC(0) * CC(2) : a * CC(0) * CC(1));
  case RO_TWO_ONE:
    return which == 0 ? a * 2. * CC(0) * CC(1) : a * CC(0) * CC(0);
  case RO_ONE_TWO:
    return which == 0 ? a * CC(1) * CC(1) : a * 2. * CC(1) * CC(0);
  }
#undef CC
  return 0.;
}

static void prod_rates_sens_exact(const go_mech *m, double T, double rho, double Mmix, const double *y, double *w,
                                  double *wsens)
{
  const int ns = m->ns, nr = m->nr;
  const double invRu = 1. / m->Ru;
  const double *Msp = m->mw, *invMsp = m->invmw;

  double kf = 0, Kc = 0, kr = 0, Rr = 0, Rnet = 0, q = 0, Ctbaf = 0, pr = 0, fCent = 0, flfConc = 0, fTroe = 0,
         gTroe = 0;
  double dRnetdrho = 0, dRnetdT = 0, dKcdToverKc = 0, dCtbafdrho = 0, dCtbafdT = 0;
  double dqdrho = 0, dqdT = 0, dfTroedT = 0, dfCentdT = 0, aTroe = 0, bTroe = 0;
  double nsTmp = 0;

  double specG[ns], dBdTSpec[ns], dCtbafdY[ns], dRnetdY[ns], dqdY[ns];

  for (int i = 0; i < ns; ++i)
  {
    w[i] = 0.;
    specG[i] = 0.;
  }
  for (int i = 0; i < (ns + 1) * (ns + 1); ++i)
    wsens[i] = 0.;

  const double invT = 1. / T;
  const double logT = log(T);
  const double invM = 1. / Mmix;
  const double ct = rho * invM;
  const double Ru = 1. / invRu;

  for (int n = 0; n < ns; ++n)
  { /* :82-126 */
    const double *c = m->cpc[n];
    if (m->cptype[n] == CP_NASA7)
    {
      if (T <= c[0])
      {
        specG[n] = c[13] + T * (c[8] - c[14] - c[8] * logT - T * (c[9] + T * (c[10] + T * (c[11] + T * c[12]))));
        dBdTSpec[n] =
            invRu * ((c[8] - Ru) * invT + c[9] + T * (2 * c[10] + T * (3 * c[11] + T * 4 * c[12])) + c[13] * invT * invT);
      }
      else
      {
        specG[n] = c[6] + T * (c[1] - c[7] - c[1] * logT - T * (c[2] + T * (c[3] + T * (c[4] + T * c[5]))));
        dBdTSpec[n] =
            invRu * ((c[1] - Ru) * invT + c[2] + T * (2 * c[3] + T * (3 * c[4] + T * 4 * c[5])) + c[6] * invT * invT);
      }
    }
    else if (m->cptype[n] == CP_CONST)
    {
      specG[n] = c[1] + c[3] * (T - c[0]) - T * (c[2] + c[3] * (logT - log(c[0])));
      dBdTSpec[n] = invT * (Msp[n] * invRu * (c[3] - invT * (c[3] * c[0] - c[1])) - 1);
    }
    else if (m->cptype[n] == CP_NASA9)
    { /* :101-116: the first region with T < Thi, else the last one */
      const double *c9 = m->c9[n];
      const int nregions = (int)c9[0];
      for (int k = 0; k < nregions; ++k)
        if (T < c9[1 + k * 11 + 1] || k == nregions - 1)
        {
          const double *a = c9 + 1 + k * 11 + 2;
          specG[n] = n9_gibbs(a, T, invT, logT);
          dBdTSpec[n] = invRu * (invT * (a[2] - Ru + invT * (a[7] + a[1] * logT - invT * a[0])) + 0.5 * a[3] +
                                 T * (a[4] * 0.3333333333333333 + T * (0.25 * a[5] + T * 0.2 * a[6])));
          break;
        }
    }
    else
      dBdTSpec[n] = 0.;
  }

  for (int r = 0; r < nr; ++r)
  {
    const rxn_t *x = &m->rx[r];
    Rnet = 0.;
    dRnetdrho = 0.;
    dRnetdT = 0.;
    dqdrho = 0.;
    dqdT = 0.;
    dCtbafdrho = 0.0;
    dCtbafdT = 0.0;
    for (int i = 0; i < ns - 1; ++i)
    {
      dqdY[i] = 0.0;
      dRnetdY[i] = 0.0;
      dCtbafdY[i] = 0.0;
    }
    switch (x->kform)
    {
    case KF_CONSTANT:
      kf = x->kf[0];
      break;
    case KF_LINEAR:
      kf = x->kf[0] * T;
      break;
    case KF_QUADRATIC:
      kf = x->kf[0] * T * T;
      break;
    case KF_RECIPROCAL:
      kf = x->kf[0] * invT;
      break;
    default:
      kf = ARRHENIUS(x->kf);
      break;
    }

    if (x->has_orders)
    { /* :198-281 */
#define C_S(i) (y[x->sp_idx[i]] * rho * x->sp_invmw[i])
      double sumOrders = 0.;
      Rnet = kf;
      for (int i = 0; i < x->n_sp; ++i)
        if (x->sp_nonzero[i])
        {
          Rnet *= pow(fmax(C_S(i), 0.), x->sp_order[i]);
          sumOrders += x->sp_order[i];
        }
      dRnetdrho = Rnet / ct * invM * sumOrders;
      dRnetdT = Rnet * ARR_SENS_OVER_K(x->kf);
      int nsIsReactant = 0, nsReactantIdx = -1;
      for (int j = 0; j < x->n_sp; ++j)
      {
        double mod_rate = kf;
        const int s = x->sp_idx[j];
        if (s == ns - 1)
        {
          nsIsReactant = 1;
          nsReactantIdx = j;
        }
        else if (x->sp_nonzero[j])
        {
          for (int l = 0; l < x->n_sp; ++l)
          {
            if (l != j)
            {
              if (x->sp_nonzero[l])
                mod_rate *= pow(fmax(C_S(l), 0.), x->sp_order[l]);
            }
            else
            {
              if (x->sp_order[l] > 1)
                mod_rate *= x->sp_order[l] * rho * x->sp_invmw[l] * pow(fmax(C_S(l), 1.e-16), x->sp_order[l] - 1.);
              else
                mod_rate *= x->sp_order[l] * rho * x->sp_invmw[l] / pow(fmax(C_S(l), 1.e-16), 1. - x->sp_order[l]);
            }
          }
          dRnetdY[s] = mod_rate;
        }
      }
      if (nsIsReactant && x->sp_nonzero[nsReactantIdx])
      {
        nsTmp = kf;
        for (int i = 0; i < x->n_sp; ++i)
        {
          if (i != nsReactantIdx)
          {
            if (x->sp_nonzero[i])
              nsTmp *= pow(C_S(i), x->sp_order[i]);
          }
          else
            nsTmp *= x->sp_order[i] * rho * x->sp_invmw[i] * pow(fmax(C_S(i), 1.e-16), x->sp_order[i] - 1.);
        }
        for (int s = 0; s < ns - 1; ++s)
          dRnetdY[s] -= nsTmp;
      }
#undef C_S
    }
    else
    {
#define C_R(i) (y[x->rc_idx[i]] * rho * x->rc_invmw[i])
#define C_P(i) (y[x->pd_idx[i]] * rho * x->pd_invmw[i])
      switch (x->fwd_order)
      { /* :287-325 */
      case RO_ONE:
        Rnet = kf * C_R(0);
        break;
      case RO_TWO:
        Rnet = kf * C_R(0) * C_R(0);
        break;
      case RO_ONE_ONE:
        Rnet = kf * C_R(0) * C_R(1);
        break;
      case RO_ONE_ONE_ONE:
        Rnet = kf * C_R(0) * C_R(1) * C_R(2);
        break;
      case RO_TWO_ONE:
        Rnet = kf * C_R(0) * C_R(0) * C_R(1);
        break;
      case RO_ONE_TWO:
        Rnet = kf * C_R(0) * C_R(1) * C_R(1);
        break;
      default:
        Rnet = kf;
        mult_other_conc(&Rnet, x->n_rc, x->rc_idx, x->rc_st, x->rc_invmw, -1, y, rho, 0);
      }
      dRnetdrho = Rnet / ct * invM * x->sum_rc_stoich;
      dRnetdT = Rnet * ARR_SENS_OVER_K(x->kf);

      int nsIsReactant = 0, nsReactantIdx = -1;
      for (int sridx = 0; sridx < x->n_rc; ++sridx)
      { /* :332-431 */
        const int s = x->rc_idx[sridx];
        if (s != ns - 1)
        {
          if (x->fwd_order != RO_OTHER)
          {
            dRnetdY[s] = order_tail(kf * rho * x->rc_invmw[sridx], x->fwd_order, sridx, x->rc_idx, x->rc_invmw, y, rho);
          }
          else
          {
            switch (x->rc_st[sridx])
            {
            case 1:
              dRnetdY[s] = kf * rho * x->rc_invmw[sridx];
              break;
            case 2:
              dRnetdY[s] = kf * rho * x->rc_invmw[sridx] * 2. * C_R(sridx);
              break;
            case 3:
              dRnetdY[s] = kf * rho * x->rc_invmw[sridx] * 3. * C_R(sridx) * C_R(sridx);
              break;
            }
            mult_other_conc(&dRnetdY[s], x->n_rc, x->rc_idx, x->rc_st, x->rc_invmw, sridx, y, rho, 0);
          }
        }
        else
        {
          nsIsReactant = 1;
          nsReactantIdx = sridx;
        }
      }
      if (nsIsReactant)
      { /* :433-526 */
        if (x->fwd_order != RO_OTHER)
        {
          nsTmp = order_tail(kf * rho * invMsp[ns - 1], x->fwd_order, nsReactantIdx, x->rc_idx, x->rc_invmw, y, rho);
        }
        else
        {
          nsTmp = kf * rho * invMsp[ns - 1];
          switch (x->rc_st[nsReactantIdx])
          {
          case 2:
            nsTmp *= 2. * C_R(nsReactantIdx);
            break;
          case 3:
            nsTmp *= 3. * C_R(nsReactantIdx) * C_R(nsReactantIdx);
            break;
          }
          mult_other_conc(&nsTmp, x->n_rc, x->rc_idx, x->rc_st, x->rc_invmw, nsReactantIdx, y, rho, 0);
        }
        for (int s = 0; s < ns - 1; ++s)
          dRnetdY[s] -= nsTmp;
      }

      if (x->reversible)
      { /* :528-812 */
        if (x->n_net >= 2 && x->n_net <= 6)
        { /* K_c only for 2..6 net species, stale otherwise */
          double gs = x->net_st[0] * specG[x->net_idx[0]];
          double ds = x->net_st[0] * dBdTSpec[x->net_idx[0]];
          for (int i = 1; i < x->n_net; ++i)
          {
            gs = gs + x->net_st[i] * specG[x->net_idx[i]];
            ds = ds + x->net_st[i] * dBdTSpec[x->net_idx[i]];
          }
          Kc = exp(-(x->sum_stoich * log(m->p_ref * invT * invRu) - invT * invRu * (gs)));
          dKcdToverKc = -ds;
        }
        kr = kf / Kc;
        switch (x->rev_order)
        {
        case RO_ONE:
          Rr = kr * C_P(0);
          break;
        case RO_TWO:
          Rr = kr * C_P(0) * C_P(0);
          break;
        case RO_ONE_ONE:
          Rr = kr * C_P(0) * C_P(1);
          break;
        case RO_ONE_ONE_ONE:
          Rr = kr * C_P(0) * C_P(1) * C_P(2);
          break;
        case RO_TWO_ONE:
          Rr = kr * C_P(0) * C_P(0) * C_P(1);
          break;
        case RO_ONE_TWO:
          Rr = kr * C_P(0) * C_P(1) * C_P(1);
          break;
        default:
          Rr = kr;
          mult_other_conc(&Rr, x->n_pd, x->pd_idx, x->pd_st, x->pd_invmw, -1, y, rho, 0);
        }
        Rnet -= Rr;
        dRnetdrho -= Rr / ct * invM * x->sum_pd_stoich;
        dRnetdT -= Rr * (ARR_SENS_OVER_K(x->kf) - dKcdToverKc);

        int nsIsProduct = 0, nsProductIdx = -1;
        for (int sridx = 0; sridx < x->n_pd; ++sridx)
        {
          const int s = x->pd_idx[sridx];
          if (s != ns - 1)
          {
            if (x->rev_order != RO_OTHER)
            {
              dRnetdY[s] -= order_tail(kr * rho * x->pd_invmw[sridx], x->rev_order, sridx, x->pd_idx, x->pd_invmw, y, rho);
            }
            else
            { /* :669-709 */
              const int as = abs(x->pd_st[sridx]);
              switch (as)
              {
              case 1:
                nsTmp = kr * rho * x->pd_invmw[sridx];
                break;
              case 2:
                nsTmp = kr * rho * x->pd_invmw[sridx] * 2. * C_P(sridx);
                break;
              case 3:
                nsTmp = kr * rho * x->pd_invmw[sridx] * 3. * C_P(sridx) * C_P(sridx);
                break;
              default:
                nsTmp = kr * rho * x->pd_invmw[sridx] * as * pow(C_P(sridx), as - 1);
                break;
              }
              mult_other_conc(&nsTmp, x->n_pd, x->pd_idx, x->pd_st, x->pd_invmw, sridx, y, rho, 1);
              dRnetdY[s] -= nsTmp;
            }
          }
          else
          {
            nsIsProduct = 1;
            nsProductIdx = sridx;
          }
        }
        if (nsIsProduct)
        { /* :718-811 */
          if (x->rev_order != RO_OTHER)
          {
            nsTmp = order_tail(kr * rho * invMsp[ns - 1], x->rev_order, nsProductIdx, x->pd_idx, x->pd_invmw, y, rho);
          }
          else
          {
            switch (abs(x->pd_st[nsProductIdx]))
            {
            case 1:
              nsTmp = kr * rho * invMsp[ns - 1];
              break;
            case 2:
              nsTmp = kr * rho * invMsp[ns - 1] * 2. * C_P(nsProductIdx);
              break;
            case 3: /* sic: `*=` in the reference, :785 */
              nsTmp *= kr * rho * invMsp[ns - 1] * 3. * C_P(nsProductIdx) * C_P(nsProductIdx);
              break;
            }
            mult_other_conc(&nsTmp, x->n_pd, x->pd_idx, x->pd_st, x->pd_invmw, nsProductIdx, y, rho, 0);
          }
          for (int s = 0; s < ns - 1; ++s)
            dRnetdY[s] += nsTmp;
        }
      }
#undef C_R
#undef C_P
    }

    /* third-body / falloff factor and its sensitivities, :817-1000 */
    switch (x->type)
    {
    case RT_SIMPLE:
      Ctbaf = 1.0;
      dCtbafdrho = 0.0;
      dCtbafdT = 0.0;
      for (int s = 0; s < ns - 1; ++s)
        dCtbafdY[s] = 0.0;
      break;
    case RT_THIRD_BODY:
    {
      Ctbaf = x->base_eff * ct;
      for (int i = 0; i < x->n_tb; ++i)
        Ctbaf += rho * (x->tb_eff[i] * y[x->tb_idx[i]]);
      dCtbafdrho = x->base_eff * invM;
      for (int i = 0; i < x->n_tb; ++i)
        dCtbafdrho += (x->tb_eff[i] * y[x->tb_idx[i]]);
      dCtbafdT = 0.0;
      const double rho_baseeff = rho * x->base_eff;
      for (int s = 0; s < ns - 1; ++s)
        dCtbafdY[s] = rho_baseeff * (invMsp[s] - invMsp[ns - 1]);
      for (int i = 0; i < x->n_tb; ++i)
        dCtbafdY[x->tb_idx[i]] += rho * (x->tb_eff[i]);
      for (int s = 0; s < x->n_tb; ++s)
        if (x->tb_idx[s] == ns - 1)
        {
          for (int ss = 0; ss < ns - 1; ++ss)
            dCtbafdY[ss] -= rho * (x->tb_eff[s]);
          break;
        }
      break;
    }
    case RT_LINDEMANN:
    {
      flfConc = x->base_eff * ct;
      for (int i = 0; i < x->n_tb; ++i)
        flfConc = flfConc + rho * (x->tb_eff[i] * y[x->tb_idx[i]]);
      const double kp_over_kf = ARRHENIUS(x->kp) / kf;
      pr = kp_over_kf * flfConc;
      Ctbaf = pr / (1. + pr);
      dCtbafdT = Ctbaf / (1. + pr) * (ARR_SENS_OVER_K(x->kp) - ARR_SENS_OVER_K(x->kf));
      nsTmp = kp_over_kf / ((1. + pr) * (1. + pr));
      dCtbafdrho = nsTmp * x->base_eff * invM;
      for (int i = 0; i < x->n_tb; ++i)
        dCtbafdrho += nsTmp * (x->tb_eff[i] * y[x->tb_idx[i]]);
      nsTmp *= rho;
      for (int s = 0; s < ns - 1; ++s)
        dCtbafdY[s] = nsTmp * x->base_eff * (invMsp[s] - invMsp[ns - 1]);
      for (int i = 0; i < x->n_tb; ++i)
        dCtbafdY[x->tb_idx[i]] += nsTmp * (x->tb_eff[i]);
      for (int s = 0; s < x->n_tb; ++s)
        if (x->tb_idx[s] == ns - 1)
        {
          for (int ss = 0; ss < ns - 1; ++ss)
            dCtbafdY[ss] -= nsTmp * (x->tb_eff[s]);
          break;
        }
      break;
    }
    case RT_TROE:
    {
      const double *troe = x->troe;
      const double t1exp = exp(-T / troe[1]);
      const double t2exp = exp(-T / troe[2]);
      const double t3exp = exp(-invT * troe[3]);
      switch (x->troeform)
      {
      case TR_T123:
        fCent = (1 - troe[0]) * t1exp + troe[0] * t2exp + t3exp;
        dfCentdT = (troe[0] - 1) / troe[1] * t1exp - troe[0] / troe[2] * t2exp + t3exp * troe[3] * invT * invT;
        break;
      case TR_T12:
        fCent = (1 - troe[0]) * t1exp + troe[0] * t2exp;
        dfCentdT = (troe[0] - 1) / troe[1] * t1exp - troe[0] / troe[2] * t2exp;
        break;
      case TR_T1:
        fCent = (1 - troe[0]) * t1exp;
        dfCentdT = (troe[0] - 1) / troe[1] * t1exp;
        break;
      case TR_T23:
        fCent = troe[0] * t2exp + t3exp;
        dfCentdT = -troe[0] / troe[2] * t2exp + t3exp * troe[3] * invT * invT;
        break;
      case TR_T2:
        fCent = troe[0] * t2exp;
        dfCentdT = -troe[0] / troe[2] * t2exp;
        break;
      case TR_T13:
        fCent = (1 - troe[0]) * t1exp + t3exp;
        dfCentdT = (troe[0] - 1) / troe[1] * t1exp + t3exp * troe[3] * invT * invT;
        break;
      case TR_T3:
        fCent = t3exp;
        dfCentdT = t3exp * troe[3] * invT * invT;
        break;
      default:
        fCent = NAN;
        dfCentdT = NAN;
      }
      flfConc = x->base_eff * ct;
      for (int i = 0; i < x->n_tb; ++i)
        flfConc = flfConc + rho * (x->tb_eff[i] * y[x->tb_idx[i]]);
      const double kp_over_kf = ARRHENIUS(x->kp) / kf;
      pr = kp_over_kf * flfConc;
      const double log10pr = log10(fmax(pr, 1.e-300));
      const double log10fcent = log10(fmax(fCent, 1.e-300));
      const double logfcent = log(fmax(fCent, 1.e-300));
      const double ln10 = log(10.);
      aTroe = log10pr - 0.67 * log10fcent - 0.4;
      bTroe = -0.14 * log10pr - 1.1762 * log10fcent + 0.806;
      gTroe = 1 / (1 + (aTroe / bTroe) * (aTroe / bTroe));
      fTroe = pow(fCent, gTroe);
      Ctbaf = fTroe * pr / (1 + pr);
      dfTroedT = fTroe * (gTroe / fCent * dfCentdT +
                          logfcent * (-2.0 * gTroe * gTroe / ln10 * aTroe / (bTroe * bTroe * bTroe) *
                                      ((bTroe + 0.14 * aTroe) * (ARR_SENS_OVER_K(x->kp) - ARR_SENS_OVER_K(x->kf)) -
                                       (0.67 * bTroe - 1.1762 * aTroe) * dfCentdT / fCent)));
      dCtbafdT = 1. / (1. + 1. / pr) * dfTroedT +
                 fTroe * pr / ((1. + pr) * (1. + pr)) * (ARR_SENS_OVER_K(x->kp) - ARR_SENS_OVER_K(x->kf));
      nsTmp = kp_over_kf * (-2.0 / (1. + pr) * fTroe * logfcent * gTroe * gTroe / ln10 * aTroe /
                                (bTroe * bTroe * bTroe) * (bTroe + 0.14 * aTroe) +
                            fTroe / ((1. + pr) * (1 + pr)));
      dCtbafdrho = nsTmp * x->base_eff * invM;
      for (int i = 0; i < x->n_tb; ++i)
        dCtbafdrho += nsTmp * (x->tb_eff[i] * y[x->tb_idx[i]]);
      nsTmp *= rho;
      for (int s = 0; s < ns - 1; ++s)
        dCtbafdY[s] = nsTmp * x->base_eff * (invMsp[s] - invMsp[ns - 1]);
      for (int i = 0; i < x->n_tb; ++i)
        dCtbafdY[x->tb_idx[i]] += nsTmp * (x->tb_eff[i]);
      for (int s = 0; s < x->n_tb; ++s)
        if (x->tb_idx[s] == ns - 1)
        {
          for (int ss = 0; ss < ns - 1; ++ss)
            dCtbafdY[ss] -= nsTmp * (x->tb_eff[s]);
          break;
        }
      break;
    }
    }

    q = Rnet * Ctbaf; /* :1002-1026 */
    dqdrho = dRnetdrho * Ctbaf + dCtbafdrho * Rnet;
    dqdT = dRnetdT * Ctbaf + dCtbafdT * Rnet;
    for (int s = 0; s < ns - 1; ++s)
      dqdY[s] = dRnetdY[s] * Ctbaf + dCtbafdY[s] * Rnet;
    const int nsp1 = ns + 1, nsm1 = ns - 1;
    for (int i = 0; i < x->n_net; ++i)
    {
      const int index = x->net_idx[i];
      const int offset = index + 2 * nsp1;
      const double factor = -x->net_st[i] * x->net_mw[i];
      w[index] += factor * q;
      wsens[index] += factor * dqdrho;
      wsens[index + nsp1] += factor * dqdT;
      for (int s = 0; s < nsm1; ++s)
        wsens[offset + nsp1 * s] += factor * dqdY[s];
    }
  }
}

void go_prod_rates_primitive_sensitivities(const go_mech *m, double rho, double T, const double *y, int option,
                                           double *out_sens) /* chemistry_kernels.cpp:464-483 */
{
  (void)option;
  double w[m->ns];
  prod_rates_sens_exact(m, T, rho, go_mixture_molecular_weight(m, y), y, w, out_sens);
}

/* ----------------------------------------------------------------------------------------------------------------
 * isobaric reactor -- isobaric_reactor_kernels.cpp
 * -------------------------------------------------------------------------------------------------------------- */
static void chem_rhs_isobaric(const go_mech *m, double rho, double cp, const double *h, const double *w,
                              double *out_rhs) /* :19-29 */
{
  const int ns = m->ns;
  out_rhs[0] = -inner_product(ns, w, h) / (rho * cp);
  const double invRho = 1. / rho;
  for (int i = 0; i < ns - 1; ++i)
    out_rhs[1 + i] = w[i] * invRho;
}

static double heat_rhs_isobaric(double T, double rho, double cp, double Tf, double Ts, double hConv, double epsRad,
                                double SoV) /* :31-37 */
{
  return SoV / (rho * cp) * (hConv * (Tf - T) + epsRad * 5.67e-8 * (Ts * Ts * Ts * Ts - T * T * T * T));
}

static void mass_rhs_isobaric(const go_mech *m, const double *y, const double *h, const double *hin, double rho,
                              double cp, const double *yin, double tau, double *out_rhs) /* :39-56 */
{
  const int ns = m->ns;
  (void)rho;
  out_rhs[0] = (hin[ns - 1] - h[ns - 1]) * yin[ns - 1];
  for (int i = 0; i < ns - 1; ++i)
  {
    out_rhs[0] += (hin[i] - h[i]) * yin[i];
    out_rhs[1 + i] = yin[i] - y[i];
  }
  out_rhs[0] /= cp;
  const double invTau = 1. / tau;
  for (int i = 0; i < ns; ++i)
    out_rhs[i] *= invTau;
}

/* :58-98. primJac is ns x (ns+1) column-major; the reference's Y_k loop writes one element past each column
 * (i < nSpec) which is overwritten by the next column / falls off the end; the in-bounds result is restated. */
static void chem_jac_isobaric(const go_mech *m, double rho, double cp, const double *cpi, double cpsensT,
                              const double *h, const double *w, const double *wsens, double *out_rhs, double *P)
{
  const int ns = m->ns;
  chem_rhs_isobaric(m, rho, cp, h, w, out_rhs);
  const double invRhoCp = 1. / (rho * cp);
  const double invRho = 1. / rho;
  const double invCp = 1. / cp;
  P[0] = -invRhoCp * inner_product(ns, wsens, h) - invRho * out_rhs[0];
  for (int i = 0; i < ns - 1; ++i)
    P[1 + i] = invRho * (wsens[i] - invRho * w[i]);
  P[ns] = -invRhoCp * (inner_product(ns, &wsens[ns + 1], h) + inner_product(ns, w, cpi)) - out_rhs[0] * cpsensT * invCp;
  for (int i = 0; i < ns - 1; ++i)
    P[ns + 1 + i] = wsens[ns + 1 + i] * invRho;
  const double cpn = cpi[ns - 1];
  for (int k = 0; k < ns - 1; ++k)
  {
    const int firstRow = (2 + k) * ns;
    P[firstRow] = -invRhoCp * inner_product(ns, &wsens[(2 + k) * (ns + 1)], h) - out_rhs[0] * (cpi[k] - cpn) * invCp;
    for (int i = 0; i < ns - 1; ++i)
      P[firstRow + 1 + i] = invRho * wsens[(2 + k) * (ns + 1) + i];
  }
}

static void mass_jac_isobaric(const go_mech *m, const double *y, double rho, double cp, double cpsensT,
                              const double *cpi, const double *h, const double *hin, const double *yin, double tau,
                              double *out_rhs, double *P) /* :100-140 */
{
  const int ns = m->ns;
  mass_rhs_isobaric(m, y, h, hin, rho, cp, yin, tau, out_rhs);
  const double invCp = 1. / cp;
  const double invTau = 1. / tau;
  for (int i = 0; i < ns; ++i)
    P[i] = 0.;
  P[ns] = -invCp * (cpsensT * out_rhs[0] + invTau * inner_product(ns, yin, cpi));
  for (int i = 0; i < ns - 1; ++i)
    P[ns + 1 + i] = 0.;
  for (int k = 0; k < ns - 1; ++k)
  {
    P[(2 + k) * ns] = -out_rhs[0] * (cpi[k] - cpi[ns - 1]) * invCp;
    for (int i = 0; i < ns - 1; ++i)
      P[(2 + k) * ns + 1 + i] = 0.;
    P[(2 + k) * ns + 1 + k] = -invTau;
  }
}

static void heat_jac_isobaric(const go_mech *m, double T, double rho, double cp, double cpsensT, const double *cpi,
                              double Tc, double Tr, double hc, double eps, double SoV, double *rate,
                              double *PJ) /* :142-168 */
{
  const int ns = m->ns;
  *rate = heat_rhs_isobaric(T, rho, cp, Tc, Tr, hc, eps, SoV);
  const double invRhoCp = 1. / (rho * cp);
  const double invCp = 1. / cp;
  PJ[0] = -*rate / rho;
  PJ[1] = -invCp * cpsensT * *rate - SoV * invRhoCp * (hc + 4. * eps * 5.67e-8 * T * T * T);
  const double cpn = cpi[ns - 1];
  for (int k = 0; k < ns - 1; ++k)
    PJ[2 + k] = invCp * *rate * (cpn - cpi[k]);
}

static void transform_isobaric_primitive_jacobian(const go_mech *m, double rho, double T, double mmw,
                                                  const double *P, double *J) /* :319-343 */
{
  const int ns = m->ns;
  for (int i = 0; i < ns * ns; ++i)
    J[i] = 0.;
  const double roT = rho / T;
  for (int i = 0; i < ns; ++i)
    J[i] = P[ns + i] - roT * P[i];
  const double negRhoMmw = -rho * mmw;
  for (int k = 0; k < ns - 1; ++k)
    for (int i = 0; i < ns; ++i)
      J[(1 + k) * ns + i] = P[(2 + k) * ns + i] + negRhoMmw * (m->invmw[k] - m->invmw[ns - 1]) * P[i];
}

void go_reactor_rhs_isobaric(const go_mech *m, const double *state, double p, double T_in, const double *y_in,
                             double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                             int heat_option, int open, double *out_rhs) /* :170-219 */
{
  const int ns = m->ns;
  double h[ns], w[ns], y[ns];
  const double T = state[0];
  extract_y(m, &state[1], y);
  const double mmw = go_mixture_molecular_weight(m, y);
  const double rho = density_from(m, p, T, mmw);
  const double cp = go_cp_mix(m, T, y);
  go_species_enthalpies(m, T, h);
  production_rates_mmw(m, T, rho, mmw, y, w);
  chem_rhs_isobaric(m, rho, cp, h, w, out_rhs);
  if (open)
  {
    double massRhs[ns], hin[ns];
    go_species_enthalpies(m, T_in, hin);
    mass_rhs_isobaric(m, y, h, hin, rho, cp, y_in, tau, massRhs);
    for (int i = 0; i < ns; ++i)
      out_rhs[i] += massRhs[i];
  }
  switch (heat_option)
  {
  case 1:
    out_rhs[0] = 0.;
    break;
  case 2:
    out_rhs[0] += heat_rhs_isobaric(T, rho, cp, T_inf, T_surf, h_conv, eps_rad, SoV);
    break;
  }
}

void go_reactor_jac_isobaric(const go_mech *m, const double *state, double p, double T_in, const double *y_in,
                             double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                             int heat_option, int open, int rates_sens_option, int sens_transform_option,
                             double *out_rhs, double *out_jac) /* :221-317 */
{
  (void)rates_sens_option;
  const int ns = m->ns;
  double cp, cpsensT, heatRate;
  double cpi[ns], cpisensT[ns], h[ns], w[ns], y[ns], heatPJ[ns + 1];
  double *wsens = (double *)malloc(sizeof(double) * (ns + 1) * (ns + 1));
  double *P = (double *)malloc(sizeof(double) * ns * (ns + 1));
  const double T = state[0];
  extract_y(m, &state[1], y);
  const double mmw = go_mixture_molecular_weight(m, y);
  const double rho = density_from(m, p, T, mmw);
  cp_mix_and_species(m, T, y, &cp, cpi);
  go_species_enthalpies(m, T, h);
  go_cp_sens_T(m, T, y, &cpsensT, cpisensT);
  prod_rates_sens_exact(m, T, rho, mmw, y, w, wsens);
  chem_jac_isobaric(m, rho, cp, cpi, cpsensT, h, w, wsens, out_rhs, P);
  if (open)
  {
    double massRhs[ns], hin[ns];
    double *mP = (double *)malloc(sizeof(double) * ns * (ns + 1));
    go_species_enthalpies(m, T_in, hin);
    mass_jac_isobaric(m, y, rho, cp, cpsensT, cpi, h, hin, y_in, tau, massRhs, mP);
    for (int i = 0; i < ns * (ns + 1); ++i)
      P[i] += mP[i];
    for (int i = 0; i < ns; ++i)
      out_rhs[i] += massRhs[i];
    free(mP);
  }
  switch (heat_option)
  {
  case 1:
    for (int k = 0; k < ns + 1; ++k)
      P[k * ns] = 0.;
    out_rhs[0] = 0.;
    break;
  case 2:
    heat_jac_isobaric(m, T, rho, cp, cpsensT, cpi, T_inf, T_surf, h_conv, eps_rad, SoV, &heatRate, heatPJ);
    for (int k = 0; k < ns + 1; ++k)
      P[k * ns] += heatPJ[k];
    out_rhs[0] += heatRate;
    break;
  }
  if (sens_transform_option == 0)
    transform_isobaric_primitive_jacobian(m, rho, T, mmw, P, out_jac);
  free(wsens);
  free(P);
}

/* ----------------------------------------------------------------------------------------------------------------
 * isochoric reactor -- isochoric_reactor_kernels.cpp. State [rho, T, Y_0..Y_{ns-2}], Jacobian (ns+1) x (ns+1)
 * column-major in the primitive variables themselves (no transform).
 * -------------------------------------------------------------------------------------------------------------- */
static void cv_mix_and_species(const go_mech *m, double T, const double *y, double mmw, double *cv,
                               double *cvi) /* thermodynamics_kernels.cpp:169-181 */
{
  cp_mix_and_species(m, T, y, cv, cvi);
  *cv -= m->Ru / mmw;
  for (int i = 0; i < m->ns; ++i)
    cvi[i] -= m->Ru * m->invmw[i];
}

static void species_energies(const go_mech *m, double T, double *e) /* thermodynamics_kernels.cpp:353-364 */
{
  go_species_enthalpies(m, T, e);
  const double RT = m->Ru * T;
  for (int i = 0; i < m->ns; ++i)
    e[i] -= RT * m->invmw[i];
}

static void chem_rhs_isochoric(const go_mech *m, double rho, double cv, const double *e, const double *w,
                               double *out_rhs) /* :19-30 */
{
  const int ns = m->ns;
  out_rhs[0] = 0.;
  out_rhs[1] = -inner_product(ns, w, e) / (rho * cv);
  const double invRho = 1. / rho;
  for (int i = 0; i < ns - 1; ++i)
    out_rhs[2 + i] = w[i] * invRho;
}

static double heat_rhs_isochoric(double T, double rho, double cv, double Tf, double Ts, double hConv, double epsRad,
                                 double SoV) /* :32-38 */
{
  return SoV / (rho * cv) * (hConv * (Tf - T) + epsRad * 5.67e-8 * (Ts * Ts * Ts * Ts - T * T * T * T));
}

static void mass_rhs_isochoric(const go_mech *m, const double *y, const double *e, const double *ein, double rho,
                               double rhoin, double cv, const double *yin, double tau, double *out_rhs) /* :40-62 */
{
  const int ns = m->ns;
  out_rhs[1] = (ein[ns - 1] - e[ns - 1]) * yin[ns - 1];
  for (int i = 0; i < ns - 1; ++i)
  {
    out_rhs[1] += (ein[i] - e[i]) * yin[i];
    out_rhs[2 + i] = yin[i] - y[i];
  }
  const double invRho = 1. / rho;
  const double invTau = 1. / tau;
  out_rhs[0] = (rhoin - rho) * invTau;
  out_rhs[1] /= cv;
  for (int i = 1; i < ns + 1; ++i)
    out_rhs[i] *= invTau * rhoin * invRho;
}

static void chem_jac_isochoric(const go_mech *m, double rho, double cv, const double *cvi, double cvsensT,
                               const double *e, const double *w, const double *wsens, double *out_rhs,
                               double *J) /* :64-110 */
{
  const int ns = m->ns, n1 = ns + 1;
  chem_rhs_isochoric(m, rho, cv, e, w, out_rhs);
  const double invRho = 1. / rho;
  const double invCv = 1. / cv;
  const double invRhoCv = 1. / (rho * cv);
  J[0] = 0.;
  J[1] = -invRhoCv * inner_product(ns, wsens, e) - invRho * out_rhs[1];
  for (int i = 0; i < ns - 1; ++i)
    J[2 + i] = invRho * (wsens[i] - invRho * w[i]);
  J[n1] = 0.;
  J[n1 + 1] = -invCv * (invRho * (inner_product(ns, &wsens[n1], e) + inner_product(ns, w, cvi)) + out_rhs[1] * cvsensT);
  for (int i = 0; i < ns - 1; ++i)
    J[n1 + 2 + i] = invRho * wsens[n1 + i];
  const double cvn = cvi[ns - 1];
  for (int k = 0; k < ns - 1; ++k)
  {
    const int fr = (2 + k) * n1;
    J[fr] = 0.;
    J[fr + 1] = -invRhoCv * inner_product(ns, &wsens[fr], e) - out_rhs[1] * (cvi[k] - cvn) * invCv;
    for (int i = 0; i < ns - 1; ++i)
      J[fr + 2 + i] = invRho * wsens[fr + i];
  }
}

static void mass_jac_isochoric(const go_mech *m, const double *y, double rho, double rhoin, double cv, double cvsensT,
                               const double *cvi, const double *e, const double *ein, const double *yin, double tau,
                               double *out_rhs, double *J) /* :112-160 */
{
  const int ns = m->ns, n1 = ns + 1;
  mass_rhs_isochoric(m, y, e, ein, rho, rhoin, cv, yin, tau, out_rhs);
  const double invRho = 1. / rho;
  const double invCv = 1. / cv;
  const double invTau = 1. / tau;
  J[0] = -invTau;
  for (int i = 0; i < ns; ++i)
    J[1 + i] = -invRho * out_rhs[1 + i];
  J[n1] = 0.;
  J[n1 + 1] = -invCv * (invTau * invRho * (rhoin * inner_product(ns, yin, cvi)) + cvsensT * out_rhs[1]);
  for (int i = 0; i < ns - 1; ++i)
    J[n1 + 2 + i] = 0.;
  const double cvn = cvi[ns - 1];
  for (int k = 0; k < ns - 1; ++k)
  {
    const int fr = (2 + k) * n1;
    J[fr] = 0.;
    J[fr + 1] = -invCv * out_rhs[1] * (cvi[k] - cvn);
    for (int i = 0; i < ns - 1; ++i)
      J[fr + 2 + i] = 0.;
    J[fr + 2 + k] = -invTau * rhoin / rho;
  }
}

static void heat_jac_isochoric(const go_mech *m, double T, double rho, double cv, double cvsensT, const double *cvi,
                               double Tc, double Tr, double hc, double eps, double SoV, double *rate,
                               double *PJ) /* :162-190 */
{
  const int ns = m->ns;
  *rate = heat_rhs_isochoric(T, rho, cv, Tc, Tr, hc, eps, SoV);
  const double invRhoCv = 1. / (rho * cv);
  const double invCv = 1. / cv;
  PJ[0] = -*rate / rho;
  PJ[1] = -invCv * cvsensT * *rate - SoV * invRhoCv * (hc + 4. * eps * 5.67e-8 * T * T * T);
  const double cvn = cvi[ns - 1];
  for (int k = 0; k < ns - 1; ++k)
    PJ[2 + k] = invCv * *rate * (cvn - cvi[k]);
}

void go_reactor_rhs_isochoric(const go_mech *m, const double *state, double rho_in, double T_in, const double *y_in,
                              double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                              int heat_option, int open, double *out_rhs) /* :192-246 */
{
  const int ns = m->ns;
  double cv, cvi[ns], e[ns], w[ns], y[ns];
  const double rho = state[0], T = state[1];
  extract_y(m, &state[2], y);
  const double mmw = go_mixture_molecular_weight(m, y);
  cv_mix_and_species(m, T, y, mmw, &cv, cvi);
  species_energies(m, T, e);
  production_rates_mmw(m, T, rho, mmw, y, w);
  chem_rhs_isochoric(m, rho, cv, e, w, out_rhs);
  if (open)
  {
    double massRhs[ns + 1], ein[ns];
    species_energies(m, T_in, ein);
    mass_rhs_isochoric(m, y, e, ein, rho, rho_in, cv, y_in, tau, massRhs);
    for (int i = 0; i < ns + 1; ++i)
      out_rhs[i] += massRhs[i];
  }
  switch (heat_option)
  {
  case 1:
    out_rhs[1] = 0.;
    break;
  case 2:
    out_rhs[1] += heat_rhs_isochoric(T, rho, cv, T_inf, T_surf, h_conv, eps_rad, SoV);
    break;
  }
}

void go_reactor_jac_isochoric(const go_mech *m, const double *state, double rho_in, double T_in, const double *y_in,
                              double tau, double T_inf, double T_surf, double h_conv, double eps_rad, double SoV,
                              int heat_option, int open, int rates_sens_option, double *out_rhs,
                              double *out_jac) /* :248-335 */
{
  (void)rates_sens_option;
  const int ns = m->ns, n1 = ns + 1;
  double cv, cvsensT, heatRate;
  double cvi[ns], cvisensT[ns], e[ns], w[ns], y[ns], heatPJ[ns + 1];
  double *wsens = (double *)malloc(sizeof(double) * n1 * n1);
  const double rho = state[0], T = state[1];
  extract_y(m, &state[2], y);
  const double mmw = go_mixture_molecular_weight(m, y);
  cv_mix_and_species(m, T, y, mmw, &cv, cvi);
  go_cp_sens_T(m, T, y, &cvsensT, cvisensT); /* cv_sens_T = cp_sens_T, combustion_kernels.h:542-545 */
  species_energies(m, T, e);
  prod_rates_sens_exact(m, T, rho, mmw, y, w, wsens);
  chem_jac_isochoric(m, rho, cv, cvi, cvsensT, e, w, wsens, out_rhs, out_jac);
  if (open)
  {
    double massRhs[ns + 1], ein[ns];
    double *mJ = (double *)malloc(sizeof(double) * n1 * n1);
    species_energies(m, T_in, ein);
    mass_jac_isochoric(m, y, rho, rho_in, cv, cvsensT, cvi, e, ein, y_in, tau, massRhs, mJ);
    for (int i = 0; i < ns + 1; ++i)
      out_rhs[i] += massRhs[i];
    for (int i = 0; i < n1 * n1; ++i)
      out_jac[i] += mJ[i];
    free(mJ);
  }
  switch (heat_option)
  {
  case 1:
    for (int k = 0; k < n1; ++k)
      out_jac[k * n1 + 1] = 0.;
    out_rhs[1] = 0.;
    break;
  case 2:
    heat_jac_isochoric(m, T, rho, cv, cvsensT, cvi, T_inf, T_surf, h_conv, eps_rad, SoV, &heatRate, heatPJ);
    out_rhs[1] += heatRate;
    for (int k = 0; k < n1; ++k)
      out_jac[k * n1 + 1] += heatPJ[k];
    break;
  }
  free(wsens);
}

void go_reactor_jac_isobaric_many(const go_mech *m, int n, const double *state, double p, int rates_sens_option,
                                  double *out_rhs, double *out_jac)
{
  const int ns = m->ns;
  const double dummy = 0.;
  for (int i = 0; i < n; ++i)
    go_reactor_jac_isobaric(m, state + (size_t)i * ns, p, 0., &dummy, 0., 0., 0., 0., 0., 0., 0, 0, rates_sens_option,
                            0, out_rhs + (size_t)i * ns, out_jac + (size_t)i * ns * ns);
}

void go_reactor_rhs_isobaric_many(const go_mech *m, int n, const double *state, double p, double *out_rhs)
{
  const int ns = m->ns;
  const double dummy = 0.;
  for (int i = 0; i < n; ++i)
    go_reactor_rhs_isobaric(m, state + (size_t)i * ns, p, 0., &dummy, 0., 0., 0., 0., 0., 0., 0, 0,
                            out_rhs + (size_t)i * ns);
}

/* ----------------------------------------------------------------------------------------------------------------
 * flamelet -- flamelet_kernels.cpp
 * -------------------------------------------------------------------------------------------------------------- */
void go_flamelet_stencils(const go_mech *m, const double *dz, int nzi, const double *chi, const double *invLe,
                          double *cmajor, double *csub, double *csup, double *mcoeff, double *ncoeff) /* :31-48 */
{
  const int ns = m->ns;
  for (int i = 0; i < nzi; ++i)
  {
    const double dzt = dz[i] + dz[i + 1];
    for (int l = 0; l < ns; ++l)
    {
      cmajor[i * ns + l] = -chi[1 + i] / (dz[i] * dz[i + 1]) * invLe[l];
      csub[i * ns + l] = chi[1 + i] / (dzt * dz[i]) * invLe[l];
      csup[i * ns + l] = chi[1 + i] / (dzt * dz[i + 1]) * invLe[l];
    }
    ncoeff[i] = 1 / (dz[i] + dz[i + 1]);
    mcoeff[i] = -ncoeff[i];
  }
}

void go_flamelet_jac_indices(const go_mech *m, int nzi, int *rows, int *cols) /* :50-90 */
{
  const int ns = m->ns;
  int idx = 0;
  for (int iz = 0; iz < nzi; ++iz)
    for (int iq = 0; iq < ns; ++iq)
      for (int jq = 0; jq < ns; ++jq)
      {
        rows[idx] = iz * ns + jq;
        cols[idx] = iz * ns + iq;
        ++idx;
      }
  for (int iz = 1; iz < nzi; ++iz)
    for (int iq = 0; iq < ns; ++iq)
    {
      rows[idx] = iz * ns + iq;
      cols[idx] = iz * ns + iq - ns;
      ++idx;
    }
  for (int iz = 0; iz < nzi - 1; ++iz)
    for (int iq = 0; iq < ns; ++iq)
    {
      rows[idx] = iz * ns + iq;
      cols[idx] = iz * ns + iq + ns;
      ++idx;
    }
}

static double cp_mix_of_state(const go_mech *m, const double *st)
{
  double y[m->ns];
  extract_y(m, &st[1], y);
  return go_cp_mix(m, st[0], y);
}

void go_flamelet_rhs(const go_mech *m, const double *state, double p, const double *oxy, const double *fuel,
                     int adiabatic, const double *T_conv, const double *h_conv, const double *T_rad,
                     const double *h_rad, int nzi, const double *cmajor, const double *csub, const double *csup,
                     const double *mcoeff, const double *ncoeff, const double *chi, int include_enthalpy_flux,
                     int include_variable_cp, int use_scaled_heat_loss, double *out_rhs) /* :1039-1218 */
{
  const int ns = m->ns;
  double maxT = -1, maxT4 = -1;
  if (use_scaled_heat_loss)
  {
    for (int i = 0; i < nzi; ++i)
      maxT = fmax(maxT, state[i * ns]);
    maxT4 = maxT * maxT * maxT * maxT;
  }
  double *cpz_grid = (double *)malloc(sizeof(double) * nzi);
  if (include_variable_cp)
  { /* :1062-1092 */
    double *cp_grid = (double *)malloc(sizeof(double) * nzi);
    for (int i = 0; i < nzi; ++i)
      cp_grid[i] = cp_mix_of_state(m, &state[i * ns]);
    for (int i = 0; i < nzi; ++i)
    {
      if (i == 0)
        cpz_grid[i] = mcoeff[i] * cp_mix_of_state(m, oxy) + ncoeff[i] * cp_grid[1];
      else if (i == nzi - 1)
        cpz_grid[i] = mcoeff[i] * cp_grid[nzi - 2] + ncoeff[i] * cp_mix_of_state(m, fuel);
      else
        cpz_grid[i] = mcoeff[i] * cp_grid[i - 1] + ncoeff[i] * cp_grid[i + 1];
    }
    free(cp_grid);
  }
  for (int i = 0; i < nzi; ++i)
  { /* :1094-1206 */
    double h[ns], w[ns], cpi[ns], y[ns];
    double cp;
    const double T = state[i * ns];
    extract_y(m, &state[i * ns + 1], y);
    const double mmw = go_mixture_molecular_weight(m, y);
    const double rho = density_from(m, p, T, mmw);
    if (include_enthalpy_flux)
      cp_mix_and_species(m, T, y, &cp, cpi);
    else
      cp = go_cp_mix(m, T, y);
    go_species_enthalpies(m, T, h);
    production_rates_mmw(m, T, rho, mmw, y, w);
    chem_rhs_isobaric(m, rho, cp, h, w, &out_rhs[i * ns]);
    if (!adiabatic)
    {
      const double hc = h_conv[i], hr = h_rad[i], Tc = T_conv[i], Tr = T_rad[i];
      if (use_scaled_heat_loss)
      {
        const double Tr4 = Tr * Tr * Tr * Tr;
        const double q = hc * (Tc - T) / (maxT - Tc) + hr * 5.67e-8 * (Tr4 - T * T * T * T) / (maxT4 - Tr4);
        out_rhs[i * ns] += q / (rho * cp);
      }
      else
      {
        const double q = hc * (Tc - T) + hr * 5.67e-8 * (Tr * Tr * Tr * Tr - T * T * T * T);
        out_rhs[i * ns] += q / (rho * cp);
      }
    }
    const double *state_nm1 = (i == 0) ? oxy : &state[(i - 1) * ns];
    const double *state_np1 = (i == nzi - 1) ? fuel : &state[(i + 1) * ns];
    if (include_enthalpy_flux)
    {
      const double cpn = cpi[ns - 1];
      double dYdZ_cpi = 0.;
      const double dTdZ = mcoeff[i] * state_nm1[0] + ncoeff[i] * state_np1[0];
      for (int j = 0; j < ns - 1; ++j)
        dYdZ_cpi += (cpi[j] - cpn) * (mcoeff[i] * state_nm1[1 + j] + ncoeff[i] * state_np1[1 + j]);
      out_rhs[i * ns] += 0.5 * chi[i] / cp * dTdZ * dYdZ_cpi;
      if (include_variable_cp)
        out_rhs[i * ns] += 0.5 * chi[i] * cpz_grid[i] / cp * dTdZ;
    }
    if (include_variable_cp && !include_enthalpy_flux)
    {
      const double dTdZ = mcoeff[i] * state_nm1[0] + ncoeff[i] * state_np1[0];
      out_rhs[i * ns] += 0.5 * chi[i] * cpz_grid[i] / cp * dTdZ;
    }
  }
  const int endIdx = (nzi - 1) * ns; /* :1208-1217 */
  for (int i = ns; i < endIdx; ++i)
    out_rhs[i] += cmajor[i] * state[i] + csub[i] * state[i - ns] + csup[i] * state[i + ns];
  for (int j = 0; j < ns; ++j)
  {
    out_rhs[j] += cmajor[j] * state[j] + csub[j] * oxy[j] + csup[j] * state[j + ns];
    out_rhs[endIdx + j] += cmajor[endIdx + j] * state[endIdx + j] + csub[endIdx + j] * state[endIdx + j - ns] +
                           csup[endIdx + j] * fuel[j];
  }
  free(cpz_grid);
}

extern void dgeev_(const char *, const char *, const int *, double *, const int *, double *, double *, double *,
                   const int *, double *, const int *, double *, int *, int *);
extern void dgetrf_(const int *, const int *, double *, const int *, int *, int *);
extern void dgetrs_(const char *, const int *, const int *, const double *, const int *, const int *, double *,
                    const int *, int *);

static void eigenvalues(int n, const double *matrix, double *re, double *im) /* blas_lapack_kernels.h:157-180 */
{
  const char N = 'N';
  double null[1], wkopt;
  int lwork = -1, info;
  double *copy = (double *)malloc(sizeof(double) * n * n);
  memcpy(copy, matrix, sizeof(double) * n * n);
  dgeev_(&N, &N, &n, copy, &n, re, im, null, &n, null, &n, &wkopt, &lwork, &info);
  lwork = (int)wkopt;
  double *work = (double *)malloc(sizeof(double) * lwork);
  dgeev_(&N, &N, &n, copy, &n, re, im, null, &n, null, &n, work, &lwork, &info);
  free(work);
  free(copy);
}

void go_flamelet_jacobian(const go_mech *m, const double *state, double p, const double *oxy, const double *fuel,
                          int adiabatic, const double *T_conv, const double *h_conv, const double *T_rad,
                          const double *h_rad, int nzi, const double *cmajor, const double *csub, const double *csup,
                          const double *mcoeff, const double *ncoeff, const double *chi, int compute_eigenvalues,
                          double diffterm, int scale_and_offset, double prefactor, int rates_sens_option,
                          int sens_transform_option, int include_enthalpy_flux, int include_variable_cp,
                          int use_scaled_heat_loss, double *out_expeig, double *out_jac) /* :1220-1409 */
{
  (void)rates_sens_option, (void)include_variable_cp;
  const int ns = m->ns;
  const int nelements = ns * (nzi * ns + 2 * (nzi - 1));
  const int blocksize = ns * ns;
  double rhsTemp[ns], re[ns], im[ns];
  double *cp = (double *)malloc(sizeof(double) * nzi);
  double *cpsensT = (double *)malloc(sizeof(double) * nzi);
  double *wsens = (double *)malloc(sizeof(double) * (ns + 1) * (ns + 1));
  double *P = (double *)malloc(sizeof(double) * ns * (ns + 1));
  double maxT = -1, maxT4 = -1;
  if (use_scaled_heat_loss)
  {
    for (int i = 0; i < nzi; ++i)
      maxT = fmax(maxT, state[i * ns]);
    maxT4 = maxT * maxT * maxT * maxT;
  }
  int idx = 0;
  for (int iz = 0; iz < nzi; ++iz)
  {
    double cpi[ns], cpisensT[ns], h[ns], w[ns], y[ns];
    const double T = state[iz * ns];
    extract_y(m, &state[iz * ns + 1], y);
    const double mmw = go_mixture_molecular_weight(m, y);
    const double rho = density_from(m, p, T, mmw);
    cp_mix_and_species(m, T, y, &cp[iz], cpi);
    go_species_enthalpies(m, T, h);
    go_cp_sens_T(m, T, y, &cpsensT[iz], cpisensT);
    prod_rates_sens_exact(m, T, rho, mmw, y, w, wsens);
    chem_jac_isobaric(m, rho, cp[iz], cpi, cpsensT[iz], h, w, wsens, rhsTemp, P);
    if (!adiabatic)
    { /* :1290-1320; the reference's k-loop runs two columns out of bounds, the in-bounds part is k < ns-1 */
      const double Tc = T_conv[iz], Tr = T_rad[iz], hc = h_conv[iz], hr = h_rad[iz];
      const double invRhoCp = 1. / (rho * cp[iz]);
      const double invCp = 1. / cp[iz];
      double q;
      if (use_scaled_heat_loss)
      {
        const double Tr4 = Tr * Tr * Tr * Tr;
        q = (hc * (Tc - T) / (maxT - Tc) + hr * 5.67e-8 * (Tr4 - T * T * T * T) / (maxT4 - Tr4)) * invRhoCp;
        P[ns] -= invCp * cpsensT[iz] * q + invRhoCp * (hc / (maxT - Tc) + 4. * hr / (maxT4 - Tr4) * 5.67e-8 * T * T * T);
      }
      else
      {
        q = (hc * (Tc - T) + hr * 5.67e-8 * (Tr * Tr * Tr * Tr - T * T * T * T)) * invRhoCp;
        P[ns] -= invCp * cpsensT[iz] * q + invRhoCp * (hc + 4. * hr * 5.67e-8 * T * T * T);
      }
      P[0] -= q / rho;
      const double cpn = cpi[ns - 1];
      for (int k = 0; k < ns - 1; ++k)
        P[(2 + k) * ns] += invCp * q * (cpn - cpi[k]);
    }
    if (sens_transform_option == 0)
      transform_isobaric_primitive_jacobian(m, rho, T, mmw, P, &out_jac[idx]);
    if (compute_eigenvalues)
    { /* :1329-1341 */
      eigenvalues(ns, &out_jac[idx], re, im);
      double exp_eig = 0.;
      for (int iq = 0; iq < ns; ++iq)
        exp_eig = fmax(exp_eig, fmax(re[iq] - diffterm, 0.));
      for (int iq = 0; iq < ns; ++iq)
        out_expeig[iz * ns + iq] = exp_eig;
    }
    for (int iq = 0; iq < ns; ++iq)
      out_jac[idx + iq * (ns + 1)] += cmajor[iz * ns + iq];
    idx += blocksize;
  }
  if (include_enthalpy_flux)
  { /* :1350-1381 */
    for (int i = 1; i < nzi - 1; ++i)
    {
      const double dTdZ = mcoeff[i] * state[(i - 1) * ns] + ncoeff[i] * state[(i + 1) * ns];
      const double dcpdZ = mcoeff[i] * cp[i - 1] + ncoeff[i] * cp[i + 1];
      const double f1 = 0.5 * chi[i] / cp[i] * dTdZ * dcpdZ;
      out_jac[i * blocksize] -= f1 / cp[i] * cpsensT[i];
    }
    {
      const double cp_oxy = cp_mix_of_state(m, oxy);
      const int i = 0;
      const double dTdZ = mcoeff[i] * oxy[0] + ncoeff[i] * state[(i + 1) * ns];
      const double dcpdZ = mcoeff[i] * cp_oxy + ncoeff[i] * cp[i + 1];
      const double f1 = 0.5 * chi[i] / cp[i] * dTdZ * dcpdZ;
      out_jac[i * blocksize] -= f1 / cp[i] * cpsensT[i];
    }
    {
      const double cp_fuel = cp_mix_of_state(m, fuel);
      const int i = nzi - 1;
      const double dTdZ = mcoeff[i] * state[(i - 1) * ns] + ncoeff[i] * fuel[0];
      const double dcpdZ = mcoeff[i] * cp[i - 1] + ncoeff[i] * cp_fuel;
      const double f1 = 0.5 * chi[i] / cp[i] * dTdZ * dcpdZ;
      out_jac[i * blocksize] -= f1 / cp[i] * cpsensT[i];
    }
  }
  const int off_diag_offset = (nzi - 1) * ns; /* :1385-1394: ACCUMULATES, the caller zeroes the buffer */
  for (int iz = 1; iz < nzi; ++iz)
    for (int iq = 0; iq < ns; ++iq)
    {
      out_jac[idx] += csub[iz * ns + iq];
      out_jac[off_diag_offset + idx] += csup[(iz - 1) * ns + iq];
      ++idx;
    }
  if (scale_and_offset)
  { /* :1395-1408 */
    for (int i = 0; i < nelements; ++i)
      out_jac[i] *= prefactor;
    for (int iz = 0; iz < nzi; ++iz)
      for (int iq = 0; iq < ns; ++iq)
        out_jac[iz * blocksize + iq * (ns + 1)] -= 1.;
  }
  free(cp);
  free(cpsensT);
  free(wsens);
  free(P);
}

/* ----------------------------------------------------------------------------------------------------------------
 * BTDDOD block Thomas -- btddod_matrix_kernels.cpp
 * -------------------------------------------------------------------------------------------------------------- */
void go_btddod_full_factorize(double *d, int num_blocks, int bs, double *l_values, int *pivots) /* :19-80 */
{
  const int nb2 = bs * bs;
  const int nelem_offdiagonal = (num_blocks - 1) * bs;
  const int nelem_blockdiagonals = num_blocks * nb2;
  const double *sub = &d[nelem_blockdiagonals];
  const double *sup = &d[nelem_blockdiagonals + nelem_offdiagonal];
  int info;
  const char trans = 'N';
  dgetrf_(&bs, &bs, d, &bs, pivots, &info);
  for (int i = 1; i < num_blocks; ++i)
  {
    const int o1 = i * nb2;
    const int om = o1 - nb2;
    for (int l = 0; l < nb2; ++l)
      l_values[o1 + l] = 0.;
    for (int l = 0; l < bs; ++l)
      l_values[o1 + l * (bs + 1)] = 1.;
    dgetrs_(&trans, &bs, &bs, &d[om], &bs, &pivots[(i - 1) * bs], &l_values[o1], &bs, &info);
    const int im1_base = (i - 1) * bs;
    for (int j = 0; j < bs; ++j)
      for (int k = 0; k < bs; ++k)
        l_values[o1 + j * bs + k] *= sub[im1_base + k];
    for (int j = 0; j < bs; ++j)
    {
      const int o3 = o1 + j * bs;
      const double fac = -sup[im1_base + j];
      for (int k = 0; k < bs; ++k)
        d[o3 + k] += fac * l_values[o3 + k];
    }
    dgetrf_(&bs, &bs, &d[o1], &bs, &pivots[i * bs], &info);
  }
}

static void lu_solve_with_copy(int n, const double *factor, const int *ipiv, const double *rhs, double *solution)
{ /* blas_lapack_kernels.h:113-123 */
  const int one = 1;
  const char trans = 'N';
  int info;
  for (int i = 0; i < n; ++i)
    solution[i] = rhs[i];
  dgetrs_(&trans, &n, &one, factor, &n, ipiv, solution, &n, &info);
}

void go_btddod_full_solve(const double *d, const double *l_values, const int *pivots, const double *rhs,
                          int num_blocks, int bs, double *x) /* :82-119 */
{
  const int nb2 = bs * bs;
  const double *sup = &d[num_blocks * nb2 + (num_blocks - 1) * bs];
  double *y = (double *)malloc(sizeof(double) * num_blocks * bs);
  double *tmp = (double *)malloc(sizeof(double) * bs);
  for (int i = 0; i < bs; ++i)
    y[i] = rhs[i];
  for (int i = 1; i < num_blocks; ++i)
  {
    double *yi = &y[i * bs];
    const double *ym = &y[(i - 1) * bs];
    const double *L = &l_values[i * nb2];
    for (int l = 0; l < bs; ++l)
      yi[l] = rhs[i * bs + l];
    /* matrix_vector_multiply(n, y, -1, L, x, 1), blas_lapack_kernels.h:60-77 */
    for (int j = 0; j < bs; ++j)
      yi[j] *= 1.;
    for (int c = 0; c < bs; ++c)
    {
      const double axi = -1. * ym[c];
      for (int j = 0; j < bs; ++j)
        yi[j] = yi[j] + L[c * bs + j] * axi;
    }
  }
  int i = num_blocks - 1;
  lu_solve_with_copy(bs, &d[i * nb2], &pivots[i * bs], &y[i * bs], &x[i * bs]);
  for (i = num_blocks - 2; i >= 0; --i)
  {
    for (int j = 0; j < bs; ++j)
      tmp[j] = y[i * bs + j] - sup[i * bs + j] * x[(i + 1) * bs + j];
    lu_solve_with_copy(bs, &d[i * nb2], &pivots[i * bs], tmp, &x[i * bs]);
  }
  free(y);
  free(tmp);
}

void go_btddod_full_matvec(const double *a, const double *vec, int num_blocks, int bs, double *out) /* :121-165 */
{
  const int nb2 = bs * bs;
  for (int i = 0; i < num_blocks; ++i)
  {
    double *yi = &out[i * bs];
    for (int j = 0; j < bs; ++j)
      yi[j] *= 0.;
    for (int c = 0; c < bs; ++c)
    {
      const double axi = 1. * vec[i * bs + c];
      for (int j = 0; j < bs; ++j)
        yi[j] = yi[j] + a[i * nb2 + c * bs + j] * axi;
    }
  }
  const double *sub = &a[num_blocks * nb2];
  const double *sup = &a[num_blocks * nb2 + (num_blocks - 1) * bs];
  for (int iz = 1; iz < num_blocks - 1; ++iz)
    for (int iq = 0; iq < bs; ++iq)
      out[iz * bs + iq] +=
          sup[iz * bs + iq] * vec[(iz + 1) * bs + iq] + sub[(iz - 1) * bs + iq] * vec[(iz - 1) * bs + iq];
  const int iz1 = num_blocks - 1;
  for (int iq = 0; iq < bs; ++iq)
    out[iz1 * bs + iq] += sub[(iz1 - 1) * bs + iq] * vec[(iz1 - 1) * bs + iq];
  for (int iq = 0; iq < bs; ++iq)
    out[iq] += sup[iq] * vec[bs + iq];
}

void go_btddod_scale_and_add_diagonal(double *a, double matrix_scale, const double *diagonal, double diag_scale,
                                      int num_blocks, int bs) /* :447-465 */
{
  const int nb2 = bs * bs;
  const int nelem_matrix = bs * (num_blocks * bs + 2 * (num_blocks - 1));
  for (int i = 0; i < nelem_matrix; ++i)
    a[i] *= matrix_scale;
  for (int iz = 0; iz < num_blocks; ++iz)
    for (int iq = 0; iq < bs; ++iq)
      a[iz * nb2 + iq * (bs + 1)] += diag_scale * diagonal[iz * bs + iq];
}
