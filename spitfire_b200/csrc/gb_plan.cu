// gb_plan.cu -- host-side construction of the static "Jacobian plan" consumed by k_jac (gb_jac.cu).
//
// The plan turns the accumulation loop of the reference (rates_sensitivities_exact.cpp:1014-1026: every reaction
// scatters factor*dq into the rows of its net species) into a gather that needs no atomics and no read-modify-write:
//   * every reaction owns a record {q, dq/drho, dq/dT, a, b, H, dq/dY_slot...} in shared memory (per state), where
//     a, b carry the dense part of dq/dY (dq/dY_s = sparse_s + a*u_s + b, u_s = 1/M_s - 1/M_ns) and
//     H = sum_i h_i * (-nu_i M_i) is the reaction enthalpy used for the temperature row;
//   * every destination -- a structurally non-zero entry of R[i][k] = sum_r nu_ri dq_r/dY_k, the five row scalars
//     (w, dw/drho, dw/dT, A, B) of every species, the enthalpy-weighted temperature-row sums and three per-state
//     scalars -- is a run of 32-bit items in one static stream; an item names a record value and the integer net
//     stoichiometric coefficient multiplying it (plain: rec | nu << 16) or two record values to be multiplied
//     (product: rec_a | rec_b << 16); a run starts with a header word slot:20 | count:11 << 20 | product << 31;
//   * the stream is cut into one contiguous range per CTA thread with balanced cost; destinations with many items are
//     split into parts that are added in a fixed order afterwards, so results are bit-reproducible run to run.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/griffon_b200.h"
#include "gb_mech.h"

namespace gb
{

int build_jac_plan(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                   const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                   const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                   const std::vector<int> &tb_off, JacPlanHost &out)
{
  const int ns = (int)m.species.size(), nr = (int)m.reactions.size(), last = ns - 1;
  out = JacPlanHost();
  auto dbits = [](double v) {
    unsigned long long u;
    std::memcpy(&u, &v, 8);
    return u;
  };

  // ---- records and parameter blob -----------------------------------------------------------------------------
  std::vector<int> rec_off(nr, 0);
  int rec_total = 0;
  for (int r = 0; r < nr; ++r)
  {
    rec_off[r] = rec_total;
    rec_total += JP_REC_HDR + (slot_off[r + 1] - slot_off[r]);
  }
  if (rec_total > 32767)
  {
    set_error("mechanism too large for the packed Jacobian plan (more than 32767 record values per state)");
    return GB_ERR_UNSUPPORTED;
  }
  out.rec_total = rec_total;
  out.prm_off.assign(nr, 0);
  for (int r = 0; r < nr; ++r)
  {
    const HostReaction &x = m.reactions[r];
    out.prm_off[r] = (int)out.prm.size();
    const int ntb = (int)x.tb_idx.size(), nsl = slot_off[r + 1] - slot_off[r];
    if (x.n_rc > 255 || ntb > 255 || nsl > 120 || std::abs(x.sum_stoich) > 127)
    {
      set_error("reaction too large for the packed parameter format");
      return GB_ERR_UNSUPPORTED;
    }
    out.prm.push_back((unsigned long long)(unsigned int)flags[r] | ((unsigned long long)(unsigned int)rec_off[r] << 32));
    unsigned long long w1 = 0;
    w1 |= (unsigned long long)(x.n_rc & 255);
    w1 |= (unsigned long long)(x.n_pd & 255) << 8;
    w1 |= (unsigned long long)(x.n_net & 255) << 16;
    w1 |= (unsigned long long)(ntb & 255) << 24;
    w1 |= (unsigned long long)(nsl & 255) << 32;
    w1 |= (unsigned long long)((unsigned char)(signed char)x.sum_stoich) << 40;
    w1 |= (unsigned long long)(x.sum_rc & 255) << 48;
    w1 |= (unsigned long long)(x.sum_pd & 255) << 56;
    out.prm.push_back(w1);
    for (int k = 0; k < 3; ++k)
      out.prm.push_back(dbits(x.kf[k]));
    if (x.type != RT_SIMPLE)
    {
      out.prm.push_back(dbits(x.base_eff));
      for (int k = 0; k < 3; ++k)
        out.prm.push_back(dbits(x.kp[k]));
      for (int k = 0; k < 4; ++k)
        out.prm.push_back(dbits(x.troe[k]));
    }
    for (int i = 0; i < x.n_rc; ++i)
    {
      out.prm.push_back((unsigned long long)(unsigned short)x.rc_idx[i] |
                        ((unsigned long long)(unsigned char)x.rc_st[i] << 16) |
                        ((unsigned long long)(unsigned char)rc_slot[NSR * (size_t)r + i] << 24));
      out.prm.push_back(dbits(m.invmw[x.rc_idx[i]]));
    }
    for (int i = 0; i < x.n_pd; ++i)
    {
      out.prm.push_back((unsigned long long)(unsigned short)x.pd_idx[i] |
                        ((unsigned long long)(unsigned char)x.pd_st[i] << 16) |
                        ((unsigned long long)(unsigned char)pd_slot[NSR * (size_t)r + i] << 24));
      out.prm.push_back(dbits(m.invmw[x.pd_idx[i]]));
    }
    for (int i = 0; i < x.n_net; ++i)
    { // net species: index | nu, then the factor -nu*MW (rates_sensitivities_exact.cpp:1018)
      out.prm.push_back((unsigned long long)(unsigned short)x.net_idx[i] |
                        ((unsigned long long)(unsigned char)(signed char)x.net_st[i] << 16));
      out.prm.push_back(dbits(-x.net_st[i] * (1. / m.invmw[x.net_idx[i]])));
    }
    for (int j = 0; j < ntb; ++j)
    {
      out.prm.push_back((unsigned long long)(unsigned short)x.tb_idx[j] |
                        ((unsigned long long)(unsigned char)tb_slot[tb_off[r] + j] << 24));
      out.prm.push_back(dbits(x.tb_eff[j]));
    }
  }

  // ---- logical destinations -------------------------------------------------------------------------------------------
  // [0, ns*(ns-1))            R[i][k], logical id k*ns + i (column-major like the output block)
  // [yend, yend + 5*ns)       row scalars q*ns + i, q = 0..4: w, dw/drho, dw/dT, A, B (as sums of nu * value)
  // [tbase, tbase + ns + 1)   temperature-row sums: sum_r H_r dq_r/d(rho | T | Y_k)
  // [sbase, sbase + 3)        per-state scalars: w.h, A.h, B.h
  const int yend = ns * (ns - 1), rbase = yend, tbase = rbase + 5 * ns, sbase = tbase + ns + 1, nlogical = sbase + 3;
  std::vector<std::vector<unsigned int>> dest(nlogical);
  for (int r = 0; r < nr; ++r)
  {
    const HostReaction &x = m.reactions[r];
    const int base = rec_off[r], nsl = slot_off[r + 1] - slot_off[r];
    bool last_involved = false;
    for (int i = 0; i < x.n_rc; ++i)
      last_involved |= (x.rc_idx[i] == last && !x.has_orders);
    for (int i = 0; i < x.n_pd; ++i)
      last_involved |= (x.pd_idx[i] == last && x.reversible && !x.has_orders);
    for (int i = 0; i < x.n_sp; ++i)
      last_involved |= (x.sp_idx[i] == last);
    for (size_t j = 0; j < x.tb_idx.size(); ++j)
      last_involved |= (x.tb_idx[j] == last);
    const bool tbtype = x.type != RT_SIMPLE;
    auto plain = [&](int rec, int nu) { return (unsigned int)rec | ((unsigned int)(nu & 255) << 24); };
    auto prod = [&](int a, int b) { return (unsigned int)a | ((unsigned int)b << 16); };
    for (int k = 0; k < x.n_net; ++k)
    {
      const int row = x.net_idx[k], nu = x.net_st[k];
      if (nu < -128 || nu > 127)
      {
        set_error("net stoichiometric coefficient outside [-128, 127]");
        return GB_ERR_UNSUPPORTED;
      }
      for (int q = 0; q < 3; ++q)
        dest[rbase + q * ns + row].push_back(plain(base + q, nu));
      if (tbtype)
        dest[rbase + 3 * ns + row].push_back(plain(base + 3, nu));
      if (last_involved)
        dest[rbase + 4 * ns + row].push_back(plain(base + 4, nu));
      for (int q = 0; q < nsl; ++q)
        dest[(int)slot_species[slot_off[r] + q] * ns + row].push_back(plain(base + JP_REC_HDR + q, nu));
    }
    dest[tbase + 0].push_back(prod(base + 5, base + 1));
    dest[tbase + 1].push_back(prod(base + 5, base + 2));
    for (int q = 0; q < nsl; ++q)
      dest[tbase + 2 + (int)slot_species[slot_off[r] + q]].push_back(prod(base + 5, base + JP_REC_HDR + q));
    dest[sbase + 0].push_back(prod(base + 5, base + 0));
    if (tbtype)
      dest[sbase + 1].push_back(prod(base + 5, base + 3));
    if (last_involved)
      dest[sbase + 2].push_back(prod(base + 5, base + 4));
  }

  // ---- compact slots: structurally zero R entries get no storage ---------------------------------------------------------
  out.emap.assign(yend, (unsigned short)0xffff);
  int nslots = 0;
  std::vector<int> slot_of(nlogical, -1);
  for (int s = 0; s < yend; ++s)
    if (!dest[s].empty())
    {
      out.emap[s] = (unsigned short)nslots;
      slot_of[s] = nslots++;
    }
  if (nslots >= 0xffff)
  {
    set_error("too many non-zero Jacobian entries for the compact slot map");
    return GB_ERR_UNSUPPORTED;
  }
  out.rbase = nslots;
  for (int s = rbase; s < tbase; ++s)
    slot_of[s] = nslots++;
  out.tbase = nslots;
  for (int s = tbase; s < sbase; ++s)
    slot_of[s] = nslots++;
  out.sbase = nslots;
  for (int s = sbase; s < nlogical; ++s)
    slot_of[s] = nslots++;

  // ---- parts, stream, balanced thread partition ------------------------------------------------------------------------------
  int split = 24;
  if (const char *e = std::getenv("GB_JAC_SPLIT"))
    split = std::max(4, std::atoi(e));
  struct Part
  {
    int logical, slot, begin, count, prod;
  };
  std::vector<Part> parts;
  size_t nitems = 0;
  for (int s = 0; s < nlogical; ++s)
  {
    if (slot_of[s] < 0)
      continue;
    const int n = (int)dest[s].size(), pr = s >= tbase ? 1 : 0;
    nitems += n;
    if (n <= split + split / 2)
      parts.push_back({s, slot_of[s], 0, n, pr});
    else
    {
      const int np = (n + split - 1) / split;
      out.fix.push_back(slot_of[s]);
      out.fix.push_back(nslots);
      out.fix.push_back(np - 1);
      for (int p = 0; p < np; ++p)
      {
        const int b = (int)((long long)n * p / np), e = (int)((long long)n * (p + 1) / np);
        parts.push_back({s, p == 0 ? slot_of[s] : nslots + p - 1, b, e - b, pr});
      }
      nslots += np - 1;
    }
  }
  if (nslots >= (1 << 20))
  {
    set_error("too many Jacobian plan destinations");
    return GB_ERR_UNSUPPORTED;
  }
  out.nslots = nslots;
  // CTA size: enough threads for one reaction each and ~16+ items each, a multiple of 32 in [64, 512]
  int threads = (int)((nitems + parts.size()) / 48);
  threads = std::min(256, std::max(64, ((threads + 31) / 32) * 32));
  if (const char *e = std::getenv("GB_JAC_THREADS"))
    threads = std::max(32, std::min(1024, (std::atoi(e) / 32) * 32));
  out.threads = threads;

  std::vector<double> cost(parts.size());
  double total = 0.;
  for (size_t p = 0; p < parts.size(); ++p)
  {
    cost[p] = 2.0 + parts[p].count * (parts[p].prod ? 1.5 : 1.0);
    total += cost[p];
  }
  out.tstart.assign(threads + 1, 0);
  size_t p = 0;
  double done = 0.;
  for (int t = 0; t < threads; ++t)
  {
    out.tstart[t] = (int)out.stream.size();
    const double target = total * (t + 1) / threads;
    while (p < parts.size() && (done + 0.5 * cost[p] <= target || t == threads - 1))
    {
      const Part &pt = parts[p];
      if (pt.count > 2047)
      {
        set_error("Jacobian plan part too long");
        return GB_ERR_UNSUPPORTED;
      }
      out.stream.push_back((unsigned int)pt.slot | ((unsigned int)pt.count << 20) | ((unsigned int)pt.prod << 31));
      for (int k = 0; k < pt.count; ++k)
        out.stream.push_back(dest[pt.logical][pt.begin + k]);
      done += cost[p];
      ++p;
    }
  }
  out.tstart[threads] = (int)out.stream.size();
  return GB_OK;
}

} // namespace gb
