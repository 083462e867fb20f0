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
  // species
  const double *mw, *invmw, *tmin, *tmax; // [ns]
  const double *cpc;                      // [ns][NCP]
  const int *cptype;                      // [ns]
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
  // ---- per-chunk images staged into shared memory by k_jac (cp.async), see gb_mech.cu pack_chunks() ----
  // parameter blob of chunk c: 8-byte words cprm[cprm_off[c] .. cprm_off[c+1]); it starts with one 32-bit word offset
  // per reaction of the chunk (padded to a whole number of 8-byte words), followed by the reactions' packed records
  const unsigned long long *cprm;
  const int *cprm_off;                             // [n_chunks+1], even (16-byte aligned chunks)
  // gather items of chunk c, sorted by (row, column, reaction): rec(16) | col(12) << 16 | (nu & 15) << 28, where rec
  // is the record slot (in doubles, relative to the chunk's record base), col the column of the extended row
  // (0..ns-2: Y_k, ns-1..ns+3: w, dw/drho, dw/dT, A, B) and nu the net stoichiometric coefficient (factor -nu*MW_row)
  const unsigned int *citems;
  const int *citem_off;                            // [n_chunks+1], multiples of 4
  // balanced segments of chunk c: row(16) | count(16) << 16 | begin(32) << 32 (begin relative to the chunk's items);
  // a segment never splits a (row, column) entry
  const unsigned long long *csegs;
  const int *cseg_off;                             // [n_chunks+1], even
  int max_prm_words, max_items, max_segs;
  // ---- Jacobian plan (gb_plan.cu), consumed by k_jac (gb_jac.cu) ----
  const unsigned long long *jp_prm; // packed parameters of all reactions
  const int *jp_prm_off;            // [nr] word offset of reaction r in jp_prm
  const unsigned int *jp_stream;    // gather stream: header {slot:20, count:11, product:1} followed by `count` items
  const int *jp_tstart;             // [jp_threads+1] stream range of every CTA thread
  const int *jp_fix;                // [3*jp_nfix] (dest slot, first extra slot, number of extra parts)
  const unsigned short *jp_emap;    // [ns*(ns-1)] logical R entry k*ns+i -> compact slot, 0xffff = structurally zero
  int jp_threads, jp_rec_total, jp_nslots, jp_rbase, jp_tbase, jp_sbase, jp_nfix;
};

constexpr int JP_REC_HDR = 6; // record = {q, dq/drho, dq/dT, a, b, H, dq/dY_slot...}

struct JacPlanHost
{
  std::vector<unsigned long long> prm;
  std::vector<int> prm_off, tstart, fix;
  std::vector<unsigned int> stream;
  std::vector<unsigned short> emap;
  int threads = 512, rec_total = 0, nslots = 0, rbase = 0, tbase = 0, sbase = 0;
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

int build_jac_plan(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                   const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                   const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                   const std::vector<int> &tb_off, JacPlanHost &out);

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
