// gb_mech.h -- host-side mechanism tables of the B200 Griffon path and their packed device image.
//
// Replaces the reference's MechanismData<8,15> (combustion_kernels.h:127-276) and chemistry_setup.cpp.
// The host keeps an array-of-reactions description (built by the gb_mech_* setters, finalized the way
// ReactionRateData::finalize does, chemistry_setup.cpp:460-732); gb_mech_commit() flattens it into SoA device
// tables laid out for the kernels in gb_kernels.cu: per-reaction parameter columns, CSR third-body lists, the
// per-reaction "slot" lists that index the sparse d q_r / d Y_s records, and the per-species row schedules that
// drive the deterministic (reaction-ordered) row gather of the Jacobian.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace gb
{

constexpr int NSR = 8;   // max reactants / products / net species per reaction (combustion_kernels.h:518)
constexpr int NCP = 16;  // 15 coefficients as in the reference (+1 pad so a species' row is 128 B)

enum CpType : int { CP_UNKNOWN = 0, CP_CONST = 1, CP_NASA7 = 2, CP_NASA9 = 3 };
enum RateType : int { RT_SIMPLE = 1, RT_THIRD_BODY = 2, RT_LINDEMANN = 3, RT_TROE = 4 };
enum KForm : int { KF_CONSTANT = 0, KF_LINEAR, KF_QUADRATIC, KF_RECIPROCAL, KF_ARRHENIUS };
// bit0: troeParams[1] (T3) term present, bit1: [2] (T1), bit2: [3] (T2); 0 = none (chemistry_setup.cpp:697-725)
enum TroeBits : int { TROE_T3 = 1, TROE_T1 = 2, TROE_T2 = 4 };

struct HostReaction
{
  int type = 0;
  bool reversible = false;
  bool has_orders = false;
  double kf[3] = {0, 0, 0};
  double kp[3] = {0, 0, 0};
  double troe[4] = {0, 0, 0, 0};
  double base_eff = 0.;
  int n_rc = 0, rc_idx[NSR], rc_st[NSR];
  int n_pd = 0, pd_idx[NSR], pd_st[NSR]; // product stoichiometry stored positive here
  int n_net = 0, net_idx[NSR], net_st[NSR];
  std::vector<int> tb_idx;
  std::vector<double> tb_eff; // invMW*(eff - default)
  int n_sp = 0, sp_idx[NSR];
  double sp_order[NSR];
  int kform = KF_ARRHENIUS, troebits = 0;
  bool fwd_special = false, rev_special = false; // forwardOrder/reverseOrder != OTHER
  int sum_stoich = 0, sum_rc = 0, sum_pd = 0;
};

// flags word per reaction (device)
constexpr int F_TYPE_MASK = 0x7;        // RateType
constexpr int F_KFORM_SHIFT = 3;        // 3 bits
constexpr int F_TROE_SHIFT = 6;         // 3 bits
constexpr int F_REVERSIBLE = 1 << 9;
constexpr int F_HAS_ORDERS = 1 << 10;
constexpr int F_FWD_SPECIAL = 1 << 11;  // sequential multiply of repeated concentrations (special-cased orders)
constexpr int F_REV_SPECIAL = 1 << 12;
constexpr int F_KC_VALID = 1 << 13;     // n_net in 2..6: the sensitivity code evaluates K_c (else stale, App. A.5)

// Device image. All pointers are device pointers into one allocation; see gb_mech.cu for the packing.
struct DeviceMech
{
  int ns, nr;
  double Ru, p_ref;
  double invRu, RuR; // 1/Ru and 1/(1/Ru), rounded as the reference's expressions round them
  // species
  const double *mw, *invmw, *tmin, *tmax; // [ns]
  const double *netmw;                    // [ns] 1/invmw: the molecular weight the net factors -nu*M use (chemistry_setup.cpp:538)
  const double *cpc;                      // [ns][NCP]
  const int *cptype;                      // [ns]
  const int *n9_off;                      // [ns] NASA9: offset of the species' block in n9, -1 otherwise
  const double *n9;                       // {nregions, (Tlo, Thi, a0..a8 [times R]) * nregions} per NASA9 species
  // reactions, SoA [nr]
  const int *flags;
  const double *kfA, *kfb, *kfE, *kpA, *kpb, *kpE, *troe /*[nr][4]*/, *base_eff;
  const int *sum_stoich, *sum_rc, *sum_pd;
  const int *n_rc, *n_pd, *n_net, *n_sp;       // [nr]
  const short *rc_idx, *pd_idx, *net_idx, *sp_idx; // [nr][NSR]
  const signed char *rc_st, *pd_st;                // [nr][NSR]  (positive)
  const double *net_fac;                           // [nr][NSR]  -nu_net*MW  (rates_sensitivities_exact.cpp:1018)
  const double *net_stmw;                          // [nr][NSR]  nu_net*MW   (chemistry_kernels.cpp:460)
  const signed char *net_st;                       // [nr][NSR]
  const double *sp_order;                          // [nr][NSR]
  const int *tb_off;                               // [nr+1]
  const short *tb_idx;                             // [tb_off[nr]]
  const double *tb_eff;
  // sparse-record slots: for reaction r the species (never the last one) whose dq/dY is non-zero beyond the
  // collapsed dense part; rc_slot/pd_slot/tb_slot map list positions to slot numbers (-1: last species)
  const int *slot_off;                             // [nr+1]
  const short *slot_species;                       // [slot_off[nr]]
  const signed char *rc_slot, *pd_slot;            // [nr][NSR]
  const signed char *tb_slot;                      // [tb_off[nr]]
  const signed char *sp_slot;                      // [nr][NSR]
  // records: chunk c covers reactions [chunk_rxn[c], chunk_rxn[c+1]); rec_off[r] = offset (in doubles) of reaction
  // r's record inside its chunk; record = {q, dq/drho, dq/dT, a, b, dqdY[slots]}
  int n_chunks, rec_cap;                           // rec_cap = max doubles per chunk per state
  const int *chunk_rxn;                            // [n_chunks+1]
  const int *rec_off;                              // [nr]
  // row schedules: species i is a net species of reactions row_rxn[row_off[c*ns+i] .. row_off[c*ns+i+1]) of chunk c
  // (ascending reaction order), with factor row_fac = -nu*MW
  const int *row_off;                              // [n_chunks*ns + 1]
  const int *row_rxn;
  const double *row_fac;
  const double *row_stmw;                          // nu*MW, for production_rates' `w -= nu*MW*(k-kr)`
  // row processing order (heaviest first) for load balance
  const short *row_order;                          // [ns]
  // ---- Jacobian plan (gb_plan.cu), consumed by k_jac (gb_jac.cu); see JacPlanHost ----
  const unsigned long long *jp_prm; // packed reaction parameters (fast records: 12 words, generic: variable)
  const unsigned int *jp_items;     // [round][step][lane]: record row (16) | nu (int8) << 16
  const unsigned short *jp_emap;    // [(ns+1)*(ns-1)] entry (row r: 0 = T, 1+i = species i <= ns-1; column c >= 1) at
                                    // r + (ns+1)*(c-1) -> row of the gathered-sum array
  // small tables, copied to shared memory by every CTA (offsets in ints into jp_tab):
  //  t_wg [nwarps+1] reaction groups of every warp; t_groups: per group kind (0 fast, 1 structured, 2 generic) and 32/G parameter
  //  offsets (-1: idle); t_wr [nwarps+1] gather rounds of every warp; t_rounds: per round first item and number of
  //  steps; t_rdest (u16) [round][lane] destination row, t_rspec (u16) its species; t_fix (dst row, first extra part row, extra parts);
  //  t_rowsrc (u16) [5][ns] rows of sum_r nu*{q, dq/drho, dq/dT, a, b}; t_csparts (dest, begin, end);
  //  t_cspfirst [ncs+1] first part of every column destination; t_csitems: row (16) | species (16) << 16
  const int *jp_tab;
  int jp_tab_words, jp_t_wg, jp_t_groups, jp_t_wr, jp_t_rounds, jp_t_rdest, jp_t_fix, jp_t_rowsrc, jp_t_csparts,
      jp_t_cspfirst, jp_t_csitems, jp_t_rspec;
  int jp_G, jp_threads, jp_rec_rows, jp_rows, jp_nfix, jp_ncs, jp_ncsp, jp_t0base, jp_c0base, jp_zrow, jp_smem;
  // ---- schedule of k_jac4 (gb_plan4.cu, gb_jac4.cu); j4_threads == 0: not available for this mechanism ----
  const unsigned int *j4_items, *j4_rdest;
  const int *j4_tab;
  int j4_tab_words, j4_t_wg, j4_t_groups, j4_t_fgroups, j4_t_wr, j4_t_rounds, j4_t_wfix, j4_t_cfxoff, j4_t_cfx, j4_nzero;
  const unsigned short *j4_zlist;
  int j4_threads, j4_ncons, j4_rec_rows, j4_nF, j4_nfg, j4_bufsz, j4_nwx, j4_smem;
};

constexpr int JP_FAST_WORDS = 12; // fast-path parameter record, 8-byte words
constexpr int JP_HDR_FAST = 3;    // fast record = {q, dq/drho, dq/dT, dq/dY_slot...}
constexpr int JP_HDR_GEN = 5;     // generic record = {q, dq/drho, dq/dT, a, b, dq/dY_slot...}
constexpr int JP_NSC = 20;        // per-state scalars of a k_jac tile
constexpr int JP_BLK = 2;         // gather steps per prefetch block (round lengths are multiples of it)

struct JacPlanHost
{
  int G = 8, threads = 512;
  std::vector<unsigned long long> prm;
  std::vector<int> wg_off, groups, wr_off, rounds, fix, cs_off;
  std::vector<unsigned int> items, cs_items;
  std::vector<unsigned short> rdest, rspec, rowsrc, emap;
  std::vector<int> tab;
  int t_wg = 0, t_groups = 0, t_wr = 0, t_rounds = 0, t_rdest = 0, t_fix = 0, t_rowsrc = 0, t_csparts = 0, t_cspfirst = 0,
      t_csitems = 0, t_rspec = 0;
  int rec_rows = 0, rows = 0, ncs = 0, ncsp = 0, t0base = 0, c0base = 0, zrow = 0;
  // statistics (printed with GB_PLAN_VERBOSE=1)
  int n_fast = 0, n_struct = 0, n_generic = 0, n_dest = 0, n_parts = 0, n_items = 0, n_steps = 0, max_rounds = 0;
};

// what both Jacobian plans share: classification of the reactions by code path, packed parameter records, record rows
// and the logical destinations (gb_plan.cu)
struct PlanCommon
{
  std::vector<char> fast, last_involved, kind; // kind: 0 fast, 1 structured, 2 generic
  std::vector<int> hdr_of, rec_off, prm_off, fidx;
  int rec_rows = 0, n_falloff = 0;
  std::vector<unsigned long long> prm;
  // logical destinations: [0, ns*(ns-1)) R[i][k] at k*ns + i; then 5*ns row scalars q*ns + i (sums of nu * {q, dq/drho,
  // dq/dT, a, b}); items = record row (16) | nu (int8) << 16, ascending reaction order
  std::vector<std::vector<unsigned int>> dest;
};

// schedule of k_jac4 (gb_plan4.cu)
struct JacPlan4Host
{
  int ncons = 0, nprod = 0, threads = 0;
  int rec_rows = 0, nF = 0, nfg = 0, bufsz = 0, nwx = 0;
  std::vector<unsigned int> items; // [round][step pair][lane][2]: record row | sign << 31
  std::vector<unsigned int> rdest; // [round][lane]: destination code | species << 16
  std::vector<int> tab;            // small tables, copied to shared memory
  int t_wg = 0, t_groups = 0, t_fgroups = 0, t_wr = 0, t_rounds = 0, t_wfix = 0, t_cfxoff = 0, t_cfx = 0, nzero = 0;
  std::vector<unsigned short> zlist; // entries (c*ns + r) without a destination
  int n_fast = 0, n_struct = 0, n_generic = 0, n_parts = 0, n_items = 0, n_steps = 0;
};

struct HostMech
{
  std::map<std::string, double> element_mw;
  std::vector<std::string> elements;
  std::vector<std::string> species;
  std::map<std::string, int> species_index;
  std::vector<double> mw, invmw;
  std::vector<int> cptype;
  std::vector<double> tmin, tmax;
  std::vector<double> cpc; // [ns][NCP]
  bool heat_capacity_sized = false;
  bool has_nasa9 = false;
  std::vector<int> n9_off;   // per species offset into n9 (-1: not NASA9)
  std::vector<double> n9;
  double p_ref = 101325., T_ref = 298.15, Ru = 8314.46261815324;
  std::vector<HostReaction> reactions;

  // device side
  bool committed = false;
  int device = -1;
  void *d_blob = nullptr;
  size_t blob_bytes = 0;
  DeviceMech dm{};
  int max_slots = 0;
  // scratch device/pinned buffers for the *_host entry points (grown on demand)
  void *d_scratch[8] = {nullptr};
  size_t d_scratch_bytes[8] = {0};
};

// builds the plan for tiles of G states and CTAs of `threads` threads; GB_ERR_UNSUPPORTED if it does not fit
int build_jac_plan(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                   const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                   const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                   const std::vector<int> &tb_off, int G, int threads, JacPlanHost &out);
size_t jac_smem_bytes(int ns, const JacPlanHost &p);
int build_jac4_plan(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                    const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                    const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                    const std::vector<int> &tb_off, int ncons, int nprod, JacPlan4Host &out);
size_t jac4_smem_bytes(int ns, const JacPlan4Host &p);
struct HostMech;
int build_plan_common(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                      const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                      const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                      const std::vector<int> &tb_off, PlanCommon &out);

// returns 0 or a negative GB_ERR_* code; message in gb::last_error
int finalize_reaction(const HostMech &m, HostReaction &x);
int commit(HostMech &m);
void release_device(HostMech &m);
void set_error(const std::string &msg);
const char *get_error();

} // namespace gb

struct gb_mech
{
  gb::HostMech h;
};
