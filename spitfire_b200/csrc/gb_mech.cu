// gb_mech.cu -- mechanism setters, reaction finalization and device packing (see gb_mech.h).
#include "gb_mech.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "../../include/griffon_b200.h"

namespace gb
{

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
const char *get_error() { return g_error.c_str(); }

// ReactionRateData::finalize (chemistry_setup.cpp:460-732): net stoichiometry in ascending species index, order
// classification, rate-constant form, Troe terms present.
int finalize_reaction(const HostMech &m, HostReaction &x)
{
  std::map<int, int> net; // species index -> nu_reactant - nu_product
  x.sum_stoich = 0;
  x.sum_rc = 0;
  x.sum_pd = 0;
  for (int i = 0; i < x.n_rc; ++i)
  {
    x.sum_stoich += x.rc_st[i];
    x.sum_rc += std::abs(x.rc_st[i]);
    net[x.rc_idx[i]] += x.rc_st[i];
  }
  for (int i = 0; i < x.n_pd; ++i)
  {
    x.sum_stoich -= x.pd_st[i];
    x.sum_pd += std::abs(x.pd_st[i]);
    net[x.pd_idx[i]] -= x.pd_st[i];
  }
  x.n_net = 0;
  for (const auto &kv : net)
    if (kv.second != 0)
    {
      if (x.n_net == NSR)
      {
        x.n_net = NSR + 1;
        break;
      }
      x.net_idx[x.n_net] = kv.first;
      x.net_st[x.n_net] = kv.second;
      ++x.n_net;
    }
  if (x.n_net < 2 || x.n_net > 8)
  {
    set_error("bad number of net reacting species (n < 2 or n > 8): " + std::to_string(x.n_net));
    return GB_ERR_ARG;
  }
  auto special = [](int n, const int *st) {
    if (n == 1)
      return st[0] == 1 || st[0] == 2;
    if (n == 2)
      return (st[0] == 1 && st[1] == 1) || (st[0] == 1 && st[1] == 2) || (st[0] == 2 && st[1] == 1);
    if (n == 3)
      return st[0] == 1 && st[1] == 1 && st[2] == 1;
    return false;
  };
  x.fwd_special = special(x.n_rc, x.rc_st);
  x.rev_special = special(x.n_pd, x.pd_st);

  if (std::abs(x.kf[2]) < 1.e-6)
  {
    if (std::abs(x.kf[1]) < 1.e-6)
      x.kform = KF_CONSTANT;
    else if (std::abs(x.kf[1] - 1) < 1.e-6)
      x.kform = KF_LINEAR;
    else if (std::abs(x.kf[1] - 2) < 1.e-6)
      x.kform = KF_QUADRATIC;
    else if (std::abs(x.kf[1] + 1) < 1.e-6)
      x.kform = KF_RECIPROCAL;
    else
      x.kform = KF_ARRHENIUS;
  }
  else
    x.kform = KF_ARRHENIUS;

  x.troebits = 0;
  if (x.type == RT_TROE)
  {
    if (std::abs(x.troe[1]) > 1.e-8)
      x.troebits |= TROE_T3;
    if (std::abs(x.troe[2]) > 1.e-8)
      x.troebits |= TROE_T1;
    if (std::abs(x.troe[3]) > 1.e-8)
      x.troebits |= TROE_T2;
    if (x.troebits == 0)
    {
      set_error("Troe reaction without any Troe term (the reference throws at evaluation)");
      return GB_ERR_ARG;
    }
  }
  (void)m;
  return GB_OK;
}

void release_device(HostMech &m)
{
  if (m.d_blob)
    cudaFree(m.d_blob);
  m.d_blob = nullptr;
  for (int i = 0; i < 8; ++i)
  {
    if (m.d_scratch[i])
      cudaFree(m.d_scratch[i]);
    m.d_scratch[i] = nullptr;
    m.d_scratch_bytes[i] = 0;
  }
  m.committed = false;
}

namespace
{
// Appends arrays to a host byte blob, 256-byte aligned; pointers are patched after the single device allocation.
struct Blob
{
  std::vector<unsigned char> bytes;
  template <class T>
  size_t add(const std::vector<T> &v)
  {
    size_t off = (bytes.size() + 255) & ~size_t(255);
    bytes.resize(off + std::max<size_t>(v.size() * sizeof(T), 8));
    if (!v.empty())
      std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
};
template <class T>
const T *at(void *base, size_t off)
{
  return reinterpret_cast<const T *>(static_cast<unsigned char *>(base) + off);
}
} // namespace

int commit(HostMech &m)
{
  if (m.committed)
    return GB_OK;
  const int ns = (int)m.species.size();
  const int nr = (int)m.reactions.size();
  if (ns < 2)
  {
    set_error("mechanism needs at least two species");
    return GB_ERR_STATE;
  }
  if (!m.heat_capacity_sized)
  {
    set_error("heat capacity data was never set (mechanism_resize_heat_capacity_data)");
    return GB_ERR_STATE;
  }
  for (int i = 0; i < ns; ++i)
    if (m.cptype[i] != CP_CONST && m.cptype[i] != CP_NASA7 && m.cptype[i] != CP_NASA9)
    {
      set_error("species " + m.species[i] + " has no heat capacity model");
      return GB_ERR_STATE;
    }
  if (ns > 32767)
  {
    set_error("too many species");
    return GB_ERR_UNSUPPORTED;
  }

  const int last = ns - 1;
  std::vector<int> flags(nr), sum_stoich(nr), sum_rc(nr), sum_pd(nr), n_rc(nr), n_pd(nr), n_net(nr), n_sp(nr);
  std::vector<double> kfA(nr), kfb(nr), kfE(nr), kpA(nr), kpb(nr), kpE(nr), troe(4 * (size_t)nr), base_eff(nr);
  std::vector<short> rc_idx(NSR * (size_t)nr, 0), pd_idx(NSR * (size_t)nr, 0), net_idx(NSR * (size_t)nr, 0),
      sp_idx(NSR * (size_t)nr, 0);
  std::vector<signed char> rc_st(NSR * (size_t)nr, 0), pd_st(NSR * (size_t)nr, 0), net_st(NSR * (size_t)nr, 0);
  std::vector<signed char> rc_slot(NSR * (size_t)nr, -1), pd_slot(NSR * (size_t)nr, -1), sp_slot(NSR * (size_t)nr, -1);
  std::vector<double> net_fac(NSR * (size_t)nr, 0.), net_stmw(NSR * (size_t)nr, 0.), sp_order(NSR * (size_t)nr, 0.);
  std::vector<int> tb_off(nr + 1, 0), slot_off(nr + 1, 0);
  std::vector<short> tb_idx, slot_species;
  std::vector<double> tb_eff;
  std::vector<signed char> tb_slot;

  m.max_slots = 0;
  for (int r = 0; r < nr; ++r)
  {
    const HostReaction &x = m.reactions[r];
    int f = x.type & F_TYPE_MASK;
    f |= x.kform << F_KFORM_SHIFT;
    f |= x.troebits << F_TROE_SHIFT;
    if (x.reversible)
      f |= F_REVERSIBLE;
    if (x.has_orders)
      f |= F_HAS_ORDERS;
    if (x.fwd_special)
      f |= F_FWD_SPECIAL;
    if (x.rev_special)
      f |= F_REV_SPECIAL;
    if (x.n_net >= 2 && x.n_net <= 6)
      f |= F_KC_VALID;
    flags[r] = f;
    kfA[r] = x.kf[0], kfb[r] = x.kf[1], kfE[r] = x.kf[2];
    kpA[r] = x.kp[0], kpb[r] = x.kp[1], kpE[r] = x.kp[2];
    for (int k = 0; k < 4; ++k)
      troe[4 * (size_t)r + k] = x.troe[k];
    base_eff[r] = x.base_eff;
    sum_stoich[r] = x.sum_stoich, sum_rc[r] = x.sum_rc, sum_pd[r] = x.sum_pd;
    n_rc[r] = x.n_rc, n_pd[r] = x.n_pd, n_net[r] = x.n_net, n_sp[r] = x.n_sp;

    // slots: unique species in order of first appearance among reactants (or special-order species), products,
    // third bodies; the last species has no column
    std::vector<int> slots;
    auto slot_of = [&](int s) -> int {
      if (s == last)
        return -1;
      for (size_t k = 0; k < slots.size(); ++k)
        if (slots[k] == s)
          return (int)k;
      slots.push_back(s);
      return (int)slots.size() - 1;
    };
    for (int i = 0; i < x.n_rc; ++i)
    {
      rc_idx[NSR * (size_t)r + i] = (short)x.rc_idx[i];
      rc_st[NSR * (size_t)r + i] = (signed char)x.rc_st[i];
      if (!x.has_orders)
        rc_slot[NSR * (size_t)r + i] = (signed char)slot_of(x.rc_idx[i]);
    }
    for (int i = 0; i < x.n_sp; ++i)
    {
      sp_idx[NSR * (size_t)r + i] = (short)x.sp_idx[i];
      sp_order[NSR * (size_t)r + i] = x.sp_order[i];
      sp_slot[NSR * (size_t)r + i] = (signed char)slot_of(x.sp_idx[i]);
    }
    for (int i = 0; i < x.n_pd; ++i)
    {
      pd_idx[NSR * (size_t)r + i] = (short)x.pd_idx[i];
      pd_st[NSR * (size_t)r + i] = (signed char)x.pd_st[i];
      if (!x.has_orders && x.reversible)
        pd_slot[NSR * (size_t)r + i] = (signed char)slot_of(x.pd_idx[i]);
    }
    for (int i = 0; i < x.n_net; ++i)
    {
      net_idx[NSR * (size_t)r + i] = (short)x.net_idx[i];
      net_st[NSR * (size_t)r + i] = (signed char)x.net_st[i];
      const double mw = 1. / m.invmw[x.net_idx[i]]; // net_mw = 1/invmw, chemistry_setup.cpp:538
      net_fac[NSR * (size_t)r + i] = -x.net_st[i] * mw;
      net_stmw[NSR * (size_t)r + i] = x.net_st[i] * mw;
    }
    tb_off[r + 1] = tb_off[r] + (int)x.tb_idx.size();
    for (size_t j = 0; j < x.tb_idx.size(); ++j)
    {
      tb_idx.push_back((short)x.tb_idx[j]);
      tb_eff.push_back(x.tb_eff[j]);
      tb_slot.push_back((signed char)slot_of(x.tb_idx[j]));
    }
    if (slots.size() > 120)
    {
      set_error("reaction with more than 120 distinct species");
      return GB_ERR_UNSUPPORTED;
    }
    slot_off[r + 1] = slot_off[r] + (int)slots.size();
    for (int s : slots)
      slot_species.push_back((short)s);
    m.max_slots = std::max(m.max_slots, (int)slots.size());
  }

  // record chunks: at most `chunk_rxn_max` reactions per chunk
  int chunk_rxn_max = 64;
  if (const char *e = std::getenv("GB_CHUNK_REACTIONS"))
    chunk_rxn_max = std::max(1, std::atoi(e));
  std::vector<int> chunk_rxn(1, 0), rec_off(nr, 0);
  int rec_cap = 8;
  {
    int cur = 0, count = 0;
    for (int r = 0; r < nr; ++r)
    {
      const int len = 5 + (slot_off[r + 1] - slot_off[r]);
      if (count == chunk_rxn_max)
      {
        chunk_rxn.push_back(r);
        rec_cap = std::max(rec_cap, cur);
        cur = 0;
        count = 0;
      }
      rec_off[r] = cur;
      cur += len;
      ++count;
    }
    rec_cap = std::max(rec_cap, cur);
    chunk_rxn.push_back(nr);
  }
  const int n_chunks = (int)chunk_rxn.size() - 1;

  // row schedules per chunk
  std::vector<int> row_off((size_t)n_chunks * ns + 1, 0), row_rxn;
  std::vector<double> row_fac, row_stmw;
  std::vector<int> row_total(ns, 0);
  for (int c = 0; c < n_chunks; ++c)
    for (int i = 0; i < ns; ++i)
    {
      for (int r = chunk_rxn[c]; r < chunk_rxn[c + 1]; ++r)
      {
        const HostReaction &x = m.reactions[r];
        for (int k = 0; k < x.n_net; ++k)
          if (x.net_idx[k] == i)
          {
            row_rxn.push_back(r);
            row_fac.push_back(net_fac[NSR * (size_t)r + k]);
            row_stmw.push_back(net_stmw[NSR * (size_t)r + k]);
            row_total[i] += 1 + (slot_off[r + 1] - slot_off[r]);
          }
      }
      row_off[(size_t)c * ns + i + 1] = (int)row_rxn.size();
    }
  std::vector<short> row_order(ns);
  for (int i = 0; i < ns; ++i)
    row_order[i] = (short)i;
  std::stable_sort(row_order.begin(), row_order.end(),
                   [&](short a, short b) { return row_total[a] > row_total[b]; });

  // Jacobian plan: the largest tile (G states) whose working set fits in shared memory and whose gather rounds fit in
  // the register-resident accumulators; CTA size by the amount of work per tile
  JacPlanHost jp;
  {
    int rc = GB_ERR_UNSUPPORTED;
    const char *eg = std::getenv("GB_JAC_G"), *et = std::getenv("GB_JAC_THREADS");
    for (int G = eg ? std::max(1, std::min(8, std::atoi(eg))) : 8; G >= 1; G /= 2)
    {
      if (G & (G - 1))
        continue;
      // first pass at full width to learn the amount of work, then pick the CTA size
      rc = build_jac_plan(m, flags, slot_off, slot_species, rc_slot, pd_slot, tb_slot, tb_off, G, 512, jp);
      if (rc != GB_OK)
        continue;
      int threads = 512;
      if (et)
        threads = std::max(64, std::min(512, (std::atoi(et) / 32) * 32));
      else
      {
        const int LPR = 32 / G;
        const int ngroups = (int)jp.groups.size() / (1 + LPR), nrounds = (int)jp.rounds.size() / 2;
        const int warps = std::max((ngroups + 3) / 4, (nrounds + 1) / 2);
        threads = std::max(128, std::min(512, 32 * warps));
      }
      if (threads != 512)
        rc = build_jac_plan(m, flags, slot_off, slot_species, rc_slot, pd_slot, tb_slot, tb_off, G, threads, jp);
      if (rc == GB_OK && jac_smem_bytes(ns, jp) <= (size_t)227 * 1024 - 64)
        break;
      rc = GB_ERR_UNSUPPORTED;
      set_error("mechanism too large for the shared-memory resident Jacobian plan");
    }
    if (rc != GB_OK)
      return rc;
  }

  // schedule of the warp-specialised reactor-Jacobian kernel (k_jac4): used when the mechanism is large enough to fill
  // its CTA and the working set of a four-state tile fits in shared memory
  JacPlan4Host j4;
  bool have_j4 = false;
  {
    int threads = 768, nprod = 4;
    if (const char *e = std::getenv("GB_JAC4_THREADS"))
      threads = std::max(128, std::min(1024, (std::atoi(e) / 32) * 32));
    if (const char *e = std::getenv("GB_JAC4_PRODUCERS"))
      nprod = std::max(2, std::min(threads / 32 - 1, std::atoi(e)));
    int min_nr = 100;
    if (const char *e = std::getenv("GB_JAC4_MIN_REACTIONS"))
      min_nr = std::atoi(e);
    // opt-in while it is slower than k_jac (GB_JAC4=1): see profiles/r02_kjac4_*.txt
    const char *on = std::getenv("GB_JAC4");
    if (on && std::atoi(on) != 0 && nr >= min_nr && ns <= 250 &&
        build_jac4_plan(m, flags, slot_off, slot_species, rc_slot, pd_slot, tb_slot, tb_off, threads / 32 - nprod, nprod,
                        j4) == GB_OK &&
        jac4_smem_bytes(ns, j4) <= (size_t)227 * 1024)
      have_j4 = true;
    if (std::getenv("GB_PLAN_VERBOSE"))
      fprintf(stderr, "[gb plan4] %s, shared memory %zu bytes\n", have_j4 ? "enabled" : "not used",
              j4.threads ? jac4_smem_bytes(ns, j4) : (size_t)0);
  }

  // (everything above is pure host work, so GB_PLAN_VERBOSE=1 shows the plan without a device)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    cudaGetLastError();
    set_error("no CUDA device is usable; the B200 Griffon path has no CPU fallback");
    return GB_ERR_CUDA;
  }
  std::vector<double> cpc = m.cpc;
  Blob b;
  std::vector<double> netmw(ns);
  for (int i = 0; i < ns; ++i)
    netmw[i] = 1. / m.invmw[i];
  const size_t o_netmw = b.add(netmw);
  const size_t o_mw = b.add(m.mw), o_invmw = b.add(m.invmw), o_tmin = b.add(m.tmin), o_tmax = b.add(m.tmax);
  const size_t o_cpc = b.add(cpc), o_cptype = b.add(m.cptype);
  std::vector<int> n9_off = m.n9_off;
  n9_off.resize(ns, -1);
  std::vector<double> n9 = m.n9;
  if (n9.empty())
    n9.push_back(0.);
  const size_t o_n9off = b.add(n9_off), o_n9 = b.add(n9);
  const size_t o_flags = b.add(flags), o_kfA = b.add(kfA), o_kfb = b.add(kfb), o_kfE = b.add(kfE), o_kpA = b.add(kpA),
               o_kpb = b.add(kpb), o_kpE = b.add(kpE), o_troe = b.add(troe), o_base = b.add(base_eff);
  const size_t o_ss = b.add(sum_stoich), o_src = b.add(sum_rc), o_spd = b.add(sum_pd);
  const size_t o_nrc = b.add(n_rc), o_npd = b.add(n_pd), o_nnet = b.add(n_net), o_nsp = b.add(n_sp);
  const size_t o_rcidx = b.add(rc_idx), o_pdidx = b.add(pd_idx), o_netidx = b.add(net_idx), o_spidx = b.add(sp_idx);
  const size_t o_rcst = b.add(rc_st), o_pdst = b.add(pd_st), o_netst = b.add(net_st);
  const size_t o_netfac = b.add(net_fac), o_netstmw = b.add(net_stmw), o_sporder = b.add(sp_order);
  const size_t o_tboff = b.add(tb_off), o_tbidx = b.add(tb_idx), o_tbeff = b.add(tb_eff), o_tbslot = b.add(tb_slot);
  const size_t o_slotoff = b.add(slot_off), o_slotsp = b.add(slot_species);
  const size_t o_rcslot = b.add(rc_slot), o_pdslot = b.add(pd_slot), o_spslot = b.add(sp_slot);
  const size_t o_chunk = b.add(chunk_rxn), o_recoff = b.add(rec_off);
  const size_t o_rowoff = b.add(row_off), o_rowrxn = b.add(row_rxn), o_rowfac = b.add(row_fac),
               o_rowstmw = b.add(row_stmw), o_roworder = b.add(row_order);
  const size_t o_jpprm = b.add(jp.prm), o_jpitems = b.add(jp.items), o_jpemap = b.add(jp.emap),
               o_jptab = b.add(jp.tab);
  const size_t o_j4items = b.add(j4.items), o_j4rdest = b.add(j4.rdest), o_j4tab = b.add(j4.tab), o_j4z = b.add(j4.zlist);

  release_device(m);
  if (cudaMalloc(&m.d_blob, b.bytes.size()) != cudaSuccess ||
      cudaMemcpy(m.d_blob, b.bytes.data(), b.bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess)
  {
    set_error(std::string("device upload of the mechanism failed: ") + cudaGetErrorString(cudaGetLastError()));
    return GB_ERR_CUDA;
  }
  m.blob_bytes = b.bytes.size();
  cudaGetDevice(&m.device);
  void *base = m.d_blob;
  DeviceMech &d = m.dm;
  d.ns = ns, d.nr = nr, d.Ru = m.Ru, d.p_ref = m.p_ref;
  d.invRu = 1. / m.Ru, d.RuR = 1. / d.invRu;
  d.mw = at<double>(base, o_mw), d.invmw = at<double>(base, o_invmw), d.netmw = at<double>(base, o_netmw);
  d.tmin = at<double>(base, o_tmin), d.tmax = at<double>(base, o_tmax);
  d.cpc = at<double>(base, o_cpc), d.cptype = at<int>(base, o_cptype);
  d.n9_off = at<int>(base, o_n9off), d.n9 = at<double>(base, o_n9);
  d.flags = at<int>(base, o_flags);
  d.kfA = at<double>(base, o_kfA), d.kfb = at<double>(base, o_kfb), d.kfE = at<double>(base, o_kfE);
  d.kpA = at<double>(base, o_kpA), d.kpb = at<double>(base, o_kpb), d.kpE = at<double>(base, o_kpE);
  d.troe = at<double>(base, o_troe), d.base_eff = at<double>(base, o_base);
  d.sum_stoich = at<int>(base, o_ss), d.sum_rc = at<int>(base, o_src), d.sum_pd = at<int>(base, o_spd);
  d.n_rc = at<int>(base, o_nrc), d.n_pd = at<int>(base, o_npd), d.n_net = at<int>(base, o_nnet);
  d.n_sp = at<int>(base, o_nsp);
  d.rc_idx = at<short>(base, o_rcidx), d.pd_idx = at<short>(base, o_pdidx), d.net_idx = at<short>(base, o_netidx);
  d.sp_idx = at<short>(base, o_spidx);
  d.rc_st = at<signed char>(base, o_rcst), d.pd_st = at<signed char>(base, o_pdst);
  d.net_st = at<signed char>(base, o_netst);
  d.net_fac = at<double>(base, o_netfac), d.net_stmw = at<double>(base, o_netstmw);
  d.sp_order = at<double>(base, o_sporder);
  d.tb_off = at<int>(base, o_tboff), d.tb_idx = at<short>(base, o_tbidx), d.tb_eff = at<double>(base, o_tbeff);
  d.tb_slot = at<signed char>(base, o_tbslot);
  d.slot_off = at<int>(base, o_slotoff), d.slot_species = at<short>(base, o_slotsp);
  d.rc_slot = at<signed char>(base, o_rcslot), d.pd_slot = at<signed char>(base, o_pdslot);
  d.sp_slot = at<signed char>(base, o_spslot);
  d.n_chunks = n_chunks, d.rec_cap = rec_cap;
  d.chunk_rxn = at<int>(base, o_chunk), d.rec_off = at<int>(base, o_recoff);
  d.row_off = at<int>(base, o_rowoff), d.row_rxn = at<int>(base, o_rowrxn);
  d.row_fac = at<double>(base, o_rowfac), d.row_stmw = at<double>(base, o_rowstmw);
  d.row_order = at<short>(base, o_roworder);
  d.jp_prm = at<unsigned long long>(base, o_jpprm);
  d.jp_items = at<unsigned int>(base, o_jpitems);
  d.jp_emap = at<unsigned short>(base, o_jpemap), d.jp_tab = at<int>(base, o_jptab);
  d.jp_tab_words = (int)jp.tab.size();
  d.jp_t_wg = jp.t_wg, d.jp_t_groups = jp.t_groups, d.jp_t_wr = jp.t_wr, d.jp_t_rounds = jp.t_rounds;
  d.jp_t_rdest = jp.t_rdest, d.jp_t_fix = jp.t_fix, d.jp_t_rowsrc = jp.t_rowsrc, d.jp_t_csparts = jp.t_csparts;
  d.jp_t_cspfirst = jp.t_cspfirst, d.jp_t_csitems = jp.t_csitems, d.jp_t_rspec = jp.t_rspec;
  d.jp_G = jp.G, d.jp_threads = jp.threads, d.jp_rec_rows = jp.rec_rows, d.jp_rows = jp.rows;
  d.jp_nfix = (int)jp.fix.size() / 3, d.jp_ncs = jp.ncs, d.jp_ncsp = jp.ncsp, d.jp_t0base = jp.t0base;
  d.jp_c0base = jp.c0base, d.jp_zrow = jp.zrow, d.jp_smem = (int)jac_smem_bytes(ns, jp);
  d.j4_items = at<unsigned int>(base, o_j4items), d.j4_rdest = at<unsigned int>(base, o_j4rdest);
  d.j4_tab = at<int>(base, o_j4tab);
  d.j4_tab_words = (int)j4.tab.size();
  d.j4_t_wg = j4.t_wg, d.j4_t_groups = j4.t_groups, d.j4_t_fgroups = j4.t_fgroups, d.j4_t_wr = j4.t_wr;
  d.j4_t_rounds = j4.t_rounds, d.j4_t_wfix = j4.t_wfix, d.j4_t_cfxoff = j4.t_cfxoff, d.j4_t_cfx = j4.t_cfx, d.j4_nzero = j4.nzero;
  d.j4_zlist = at<unsigned short>(base, o_j4z);
  d.j4_threads = have_j4 ? j4.threads : 0, d.j4_ncons = j4.ncons, d.j4_rec_rows = j4.rec_rows, d.j4_nF = j4.nF;
  d.j4_nfg = j4.nfg, d.j4_bufsz = j4.bufsz, d.j4_nwx = j4.nwx, d.j4_smem = have_j4 ? (int)jac4_smem_bytes(ns, j4) : 0;
  m.committed = true;
  return GB_OK;
}

} // namespace gb

// ------------------------------------------------------------------------------------------------------------------
// C-ABI: mechanism construction (include/griffon_b200.h)
// ------------------------------------------------------------------------------------------------------------------
using gb::HostMech;
using gb::HostReaction;

extern "C"
{

  const char *gb_last_error(void) { return gb::get_error(); }

  int gb_cuda_device_count(void)
  {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
    return n;
  }

  gb_mech *gb_mech_create(void) { return new gb_mech(); }

  void gb_mech_destroy(gb_mech *m)
  {
    if (!m)
      return;
    gb::release_device(m->h);
    delete m;
  }

#define GB_TOUCH(m)                 \
  if (!(m))                         \
  {                                 \
    gb::set_error("null handle");   \
    return GB_ERR_ARG;              \
  }                                 \
  (m)->h.committed = false;

  int gb_mech_set_ref_pressure(gb_mech *m, double p)
  {
    GB_TOUCH(m);
    m->h.p_ref = p;
    return GB_OK;
  }
  int gb_mech_set_ref_temperature(gb_mech *m, double T)
  {
    GB_TOUCH(m);
    m->h.T_ref = T;
    return GB_OK;
  }
  int gb_mech_set_gas_constant(gb_mech *m, double Ru)
  {
    GB_TOUCH(m);
    m->h.Ru = Ru;
    return GB_OK;
  }
  int gb_mech_set_element_mw(gb_mech *m, const char *e, double mw)
  {
    GB_TOUCH(m);
    m->h.element_mw[e] = mw;
    return GB_OK;
  }
  int gb_mech_add_element(gb_mech *m, const char *e)
  {
    GB_TOUCH(m);
    auto &v = m->h.elements;
    if (std::find(v.begin(), v.end(), std::string(e)) == v.end())
      v.push_back(e);
    return GB_OK;
  }
  int gb_mech_add_species(gb_mech *m, const char *name, int n_atoms, const char *const *atom_names,
                          const double *atom_counts)
  {
    GB_TOUCH(m);
    HostMech &h = m->h;
    if (h.species_index.count(name))
    {
      gb::set_error(std::string("species ") + name + " cannot be added twice");
      return GB_ERR_ARG;
    }
    std::map<std::string, double> atoms; // name-ordered summation, chemistry_setup.cpp:52-56
    for (int i = 0; i < n_atoms; ++i)
      atoms[atom_names[i]] = atom_counts[i];
    double mw = 0.;
    for (const auto &a : atoms)
    {
      if (std::find(h.elements.begin(), h.elements.end(), a.first) == h.elements.end() ||
          !h.element_mw.count(a.first))
      {
        gb::set_error("cannot find atom " + a.first);
        return GB_ERR_ARG;
      }
      mw += h.element_mw.at(a.first) * a.second;
    }
    h.species_index[name] = (int)h.species.size();
    h.species.push_back(name);
    h.mw.push_back(mw);
    h.invmw.push_back(1. / mw);
    return GB_OK;
  }
  int gb_mech_resize_heat_capacity_data(gb_mech *m)
  {
    GB_TOUCH(m);
    HostMech &h = m->h;
    const size_t ns = h.species.size();
    h.cpc.assign(ns * gb::NCP, 0.);
    h.tmin.assign(ns, 0.);
    h.tmax.assign(ns, 0.);
    h.cptype.assign(ns, gb::CP_UNKNOWN);
    h.heat_capacity_sized = true;
    return GB_OK;
  }
  static int species_or_error(gb_mech *m, const char *s)
  {
    auto it = m->h.species_index.find(s);
    if (it == m->h.species_index.end() || !m->h.heat_capacity_sized)
    {
      gb::set_error(std::string("unknown species (or heat capacity data not sized): ") + s);
      return -1;
    }
    return it->second;
  }
  int gb_mech_add_const_cp(gb_mech *m, const char *s, double Tmin, double Tmax, double T0, double h0, double s0,
                           double cp)
  {
    GB_TOUCH(m);
    const int i = species_or_error(m, s);
    if (i < 0)
      return GB_ERR_ARG;
    HostMech &h = m->h;
    h.cptype[i] = gb::CP_CONST;
    h.tmin[i] = Tmin, h.tmax[i] = Tmax;
    double *c = &h.cpc[(size_t)i * gb::NCP];
    c[0] = T0, c[1] = h0, c[2] = s0, c[3] = cp;
    return GB_OK;
  }
  int gb_mech_add_nasa7_cp(gb_mech *m, const char *s, double Tmin, double Tmid, double Tmax, const double *low7,
                           const double *high7)
  {
    GB_TOUCH(m);
    const int i = species_or_error(m, s);
    if (i < 0)
      return GB_ERR_ARG;
    HostMech &h = m->h;
    h.cptype[i] = gb::CP_NASA7;
    h.tmin[i] = Tmin, h.tmax[i] = Tmax;
    double *c = &h.cpc[(size_t)i * gb::NCP];
    // pre-scaled storage, chemistry_setup.cpp:111-129: c[0]=Tmid, c[1..7]=R*high (a1/2,a2/6,a3/12,a4/20), c[8..14]=low
    c[0] = Tmid;
    for (int k = 0; k < 7; ++k)
      c[1 + k] = high7[k] * h.Ru;
    for (int k = 0; k < 7; ++k)
      c[8 + k] = low7[k] * h.Ru;
    c[2] /= 2., c[3] /= 6., c[4] /= 12., c[5] /= 20.;
    c[9] /= 2., c[10] /= 6., c[11] /= 12., c[12] /= 20.;
    return GB_OK;
  }
  int gb_mech_add_nasa9_cp(gb_mech *m, const char *s, double Tmin, double Tmax, int n, const double *c)
  {
    GB_TOUCH(m);
    const int i = species_or_error(m, s);
    if (i < 0)
      return GB_ERR_ARG;
    const int nregions = (n >= 1 && c) ? (int)c[0] : 0;
    if (nregions < 1 || n < 1 + 11 * nregions)
    {
      gb::set_error("NASA9 coefficient list must be {nregions, (Tlo, Thi, a0..a8) * nregions}");
      return GB_ERR_ARG;
    }
    HostMech &h = m->h;
    h.cptype[i] = gb::CP_NASA9;
    h.tmin[i] = Tmin, h.tmax[i] = Tmax;
    h.has_nasa9 = true;
    // chemistry_setup.cpp:132-152: region bounds as given, a0..a8 times R
    if ((int)h.n9_off.size() != (int)h.species.size())
      h.n9_off.assign(h.species.size(), -1);
    h.n9_off[i] = (int)h.n9.size();
    h.n9.push_back(c[0]);
    for (int k = 0; k < nregions; ++k)
      for (int j = 0; j < 11; ++j)
        h.n9.push_back((j < 2 ? 1.0 : h.Ru) * c[1 + k * 11 + j]);
    return GB_OK;
  }

  int gb_mech_add_reaction(gb_mech *m, int type, int reversible, int n_reactants, const char *const *reactant_names,
                           const int *reactant_stoich, int n_products, const char *const *product_names,
                           const int *product_stoich, double fwd_A, double fwd_b, double fwd_Ea, int n_eff,
                           const char *const *eff_names, const double *eff_values, double default_eff, double flf_A,
                           double flf_b, double flf_Ea, const double *troe4, int n_orders,
                           const char *const *order_names, const double *order_values)
  {
    GB_TOUCH(m);
    HostMech &h = m->h;
    if (type < gb::RT_SIMPLE || type > gb::RT_TROE)
    {
      gb::set_error("unknown reaction type");
      return GB_ERR_ARG;
    }
    if (n_reactants > gb::NSR || n_products > gb::NSR || n_orders > gb::NSR || n_reactants < 0 || n_products < 0)
    {
      gb::set_error("more than 8 reactants/products/orders in one reaction");
      return GB_ERR_ARG;
    }
    HostReaction x;
    x.type = type;
    x.reversible = reversible != 0;
    x.has_orders = n_orders > 0;
    x.kf[0] = fwd_A, x.kf[1] = fwd_b, x.kf[2] = fwd_Ea;
    auto lookup = [&](const char *name, int &out) {
      auto it = h.species_index.find(name);
      if (it == h.species_index.end())
      {
        gb::set_error(std::string("reaction refers to unknown species ") + name);
        return false;
      }
      out = it->second;
      return true;
    };
    // the reference iterates std::map<std::string,...>: species appear in byte-wise name order
    std::map<std::string, int> rs, ps;
    std::map<std::string, double> es, os;
    for (int i = 0; i < n_reactants; ++i)
      rs[reactant_names[i]] = reactant_stoich[i];
    for (int i = 0; i < n_products; ++i)
      ps[product_names[i]] = product_stoich[i];
    for (int i = 0; i < n_eff; ++i)
      es[eff_names[i]] = eff_values[i];
    for (int i = 0; i < n_orders; ++i)
      os[order_names[i]] = order_values[i];
    for (const auto &kv : rs)
    {
      if (!lookup(kv.first.c_str(), x.rc_idx[x.n_rc]))
        return GB_ERR_ARG;
      x.rc_st[x.n_rc++] = kv.second;
    }
    for (const auto &kv : ps)
    {
      if (!lookup(kv.first.c_str(), x.pd_idx[x.n_pd]))
        return GB_ERR_ARG;
      x.pd_st[x.n_pd++] = kv.second;
    }
    for (const auto &kv : os)
    {
      if (!lookup(kv.first.c_str(), x.sp_idx[x.n_sp]))
        return GB_ERR_ARG;
      x.sp_order[x.n_sp++] = kv.second;
    }
    if (type != gb::RT_SIMPLE)
    {
      x.base_eff = default_eff;
      for (const auto &kv : es)
      {
        int s;
        if (!lookup(kv.first.c_str(), s))
          return GB_ERR_ARG;
        x.tb_idx.push_back(s);
        x.tb_eff.push_back(h.invmw[s] * (kv.second - default_eff)); // chemistry_setup.cpp:416
      }
    }
    if (type == gb::RT_LINDEMANN || type == gb::RT_TROE)
      x.kp[0] = flf_A, x.kp[1] = flf_b, x.kp[2] = flf_Ea;
    if (type == gb::RT_TROE && troe4)
      for (int k = 0; k < 4; ++k)
        x.troe[k] = troe4[k];
    const int rc = gb::finalize_reaction(h, x);
    if (rc != GB_OK)
      return rc;
    h.reactions.push_back(x);
    return GB_OK;
  }

  int gb_mech_n_species(const gb_mech *m) { return m ? (int)m->h.species.size() : 0; }
  int gb_mech_n_reactions(const gb_mech *m) { return m ? (int)m->h.reactions.size() : 0; }
  int gb_mech_molecular_weights(const gb_mech *m, double *out)
  {
    if (!m || !out)
      return GB_ERR_ARG;
    for (size_t i = 0; i < m->h.mw.size(); ++i)
      out[i] = m->h.mw[i];
    return GB_OK;
  }
  int gb_mech_commit(gb_mech *m)
  {
    if (!m)
      return GB_ERR_ARG;
    return gb::commit(m->h);
  }
}
