// gb_plan.cu -- host-side construction of the static "Jacobian plan" consumed by k_jac (gb_jac.cu).
//
// The plan turns the accumulation loop of the reference (rates_sensitivities_exact.cpp:1014-1026: every reaction
// scatters factor*dq into the rows of its net species) into a schedule without atomics or read-modify-write:
//   * reaction phase: the reactions are sorted by code path and packed into groups of 32/G reactions; a warp evaluates a
//     group for the G states of the tile at once (lane = state + G * position in the group). The dominant shape
//     "simple, A + B (<)=> C + D, unit coefficients" has a straight-line fast path with a fixed 96-byte parameter
//     record; everything else goes through the generic path (variable-length parameter record). Groups are assigned to
//     warps by estimated cost (longest processing time first).
//     Every reaction owns a record in shared memory (per state): fast {q, dq/drho, dq/dT, dq/dY_slot...}, generic
//     {q, dq/drho, dq/dT, a, b, dq/dY_slot...}; a, b carry the dense part of dq/dY (dq/dY_s = sparse_s + a*u_s + b,
//     u_s = 1/M_s - 1/M_ns) of third-body reactions and of reactions involving the last species.
//   * gather phase: every destination -- a structurally non-zero entry of R[i][k] = sum_r nu_ri dq_r/dY_k or one of the
//     five row scalars (sums of nu * {q, dq/drho, dq/dT, a, b}) of a species -- is a list of items (record row, nu) in
//     ascending reaction order, the reference's accumulation order. Long lists are cut into parts that are added
//     afterwards in part order. Parts are sorted by length and dealt out 32 at a time ("rounds": the 32 lanes of a
//     warp gather 32 parts of equal padded length in lock step); rounds are assigned to warps by length (LPT).
//   * column sums: the temperature row needs sum_i h_i dw_i/dx (isobaric_reactor_kernels.cpp:74-92), formed from the
//     gathered rows in species order like the reference's inner products.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/griffon_b200.h"
#include "gb_mech.h"

namespace gb
{

size_t jac_smem_bytes(int ns, const JacPlanHost &p)
{
  const size_t region = (size_t)std::max(p.rec_rows, p.rows) + 2;
  const size_t doubles = (size_t)p.G * (JP_NSC + 7 * (size_t)ns + region + p.ncsp) + 3 * (size_t)ns + (ns & 1);
  return doubles * sizeof(double) + sizeof(int) * p.tab.size() +
         sizeof(unsigned short) * (size_t)(ns + 1) * (ns - 1) + 16;
}

int build_plan_common(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                      const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                      const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                      const std::vector<int> &tb_off, PlanCommon &out)
{
  const int ns = (int)m.species.size(), nr = (int)m.reactions.size(), last = ns - 1;
  out = PlanCommon();
  auto dbits = [](double v) {
    unsigned long long u;
    std::memcpy(&u, &v, 8);
    return u;
  };
  // ---- classification, records, parameter blob -----------------------------------------------------------------------
  std::vector<char> &fast = out.fast, &last_involved = out.last_involved, &kind = out.kind;
  std::vector<int> &hdr_of = out.hdr_of, &rec_off = out.rec_off, &prm_off = out.prm_off;
  fast.assign(nr, 0), last_involved.assign(nr, 0), kind.assign(nr, 2);
  hdr_of.assign(nr, JP_HDR_GEN), rec_off.assign(nr, 0), prm_off.assign(nr, 0);
  out.fidx.assign(nr, -1);
  int rec_rows = 0;
  for (int r = 0; r < nr; ++r)
  {
    const HostReaction &x = m.reactions[r];
    bool li = false;
    for (int i = 0; i < x.n_rc; ++i)
      li |= (x.rc_idx[i] == last && !x.has_orders);
    for (int i = 0; i < x.n_pd; ++i)
      li |= (x.pd_idx[i] == last && x.reversible && !x.has_orders);
    for (int i = 0; i < x.n_sp; ++i)
      li |= (x.sp_idx[i] == last);
    for (size_t j = 0; j < x.tb_idx.size(); ++j)
      li |= (x.tb_idx[j] == last);
    last_involved[r] = li;
    bool f = x.type == RT_SIMPLE && !x.has_orders && !li && x.n_rc == 2 && x.rc_st[0] == 1 && x.rc_st[1] == 1 &&
             x.rc_idx[0] != x.rc_idx[1] && x.n_net <= 4 && (flags[r] & F_KC_VALID);
    if (f && x.reversible)
    {
      f = x.n_pd == 2 && x.pd_st[0] == 1 && x.pd_st[1] == 1 && x.pd_idx[0] != x.pd_idx[1];
      for (int i = 0; f && i < 2; ++i)
        for (int j = 0; j < 2; ++j)
          f = f && x.rc_idx[i] != x.pd_idx[j];
    }
    const int nsl = slot_off[r + 1] - slot_off[r];
    if (f && nsl != (x.reversible ? 4 : 2))
      f = false;
    if (getenv("GB_JAC_NOFAST"))
      f = false;
    fast[r] = f;
    // structured path: straight-line code for up to three reactant and three product entries with coefficients
    // <= 3, optionally with a third-body / falloff factor; the last species may only appear as a third body
    bool st = !f && !x.has_orders && x.n_rc <= 3 && x.n_net <= 6 && nsl <= 100 && (flags[r] & F_KC_VALID);
    for (int i = 0; st && i < x.n_rc; ++i)
      st = x.rc_st[i] >= 1 && x.rc_st[i] <= 3 && x.rc_idx[i] != last;
    if (st && x.reversible)
    {
      st = x.n_pd <= 3;
      for (int i = 0; st && i < x.n_pd; ++i)
        st = x.pd_st[i] >= 1 && x.pd_st[i] <= 3 && x.pd_idx[i] != last;
    }
    if (st && x.type == RT_SIMPLE && li)
      st = false;
    if (getenv("GB_JAC_NOSTRUCT"))
      st = false;
    kind[r] = f ? 0 : (st ? 1 : 2);
    hdr_of[r] = (f || (st && x.type == RT_SIMPLE)) ? JP_HDR_FAST : JP_HDR_GEN;
    rec_off[r] = rec_rows;
    rec_rows += hdr_of[r] + nsl;
  }
  if (nr >= (1 << 18))
  {
    set_error("too many reactions for the packed parameter format");
    return GB_ERR_UNSUPPORTED;
  }
  if (rec_rows > 65000)
  {
    set_error("mechanism too large for the packed Jacobian plan (more than 65000 record values per state)");
    return GB_ERR_UNSUPPORTED;
  }
  out.rec_rows = rec_rows;
  for (int r = 0; r < nr; ++r)
  {
    const HostReaction &x = m.reactions[r];
    if (out.prm.size() & 1)
      out.prm.push_back(0ull); // 16-byte alignment of every record
    prm_off[r] = (int)out.prm.size();
    const int ntb = (int)x.tb_idx.size(), nsl = slot_off[r + 1] - slot_off[r];
    if (x.n_rc > 255 || ntb > 255 || nsl > 120 || std::abs(x.sum_stoich) > 127)
    {
      set_error("reaction too large for the packed parameter format");
      return GB_ERR_UNSUPPORTED;
    }
    if (fast[r])
    {
      // w0: flags | record row << 32; w1: species a, b, c, d; w2: net species; w3: net nu (int8 x4), sum_stoich, n_net
      // w4..6: A, b, Ea/R; w7..10: 1/M of a, b, c, d; w11: pad
      const int c = x.reversible ? x.pd_idx[0] : 0, d = x.reversible ? x.pd_idx[1] : 0;
      out.prm.push_back((unsigned long long)((unsigned int)flags[r] | ((unsigned int)r << 14)) |
                        ((unsigned long long)(unsigned int)rec_off[r] << 32));
      out.prm.push_back((unsigned long long)(unsigned short)x.rc_idx[0] |
                        ((unsigned long long)(unsigned short)x.rc_idx[1] << 16) |
                        ((unsigned long long)(unsigned short)c << 32) | ((unsigned long long)(unsigned short)d << 48));
      unsigned long long w2 = 0, w3 = 0;
      for (int i = 0; i < 4; ++i)
      {
        const int idx = i < x.n_net ? x.net_idx[i] : x.net_idx[0];
        const int nu = i < x.n_net ? x.net_st[i] : 0;
        w2 |= (unsigned long long)(unsigned short)idx << (16 * i);
        w3 |= (unsigned long long)(unsigned char)(signed char)nu << (8 * i);
      }
      w3 |= (unsigned long long)(unsigned char)(signed char)x.sum_stoich << 32;
      w3 |= (unsigned long long)(x.n_net & 255) << 40;
      out.prm.push_back(w2);
      out.prm.push_back(w3);
      for (int k = 0; k < 3; ++k)
        out.prm.push_back(dbits(x.kf[k]));
      out.prm.push_back(dbits(m.invmw[x.rc_idx[0]]));
      out.prm.push_back(dbits(m.invmw[x.rc_idx[1]]));
      out.prm.push_back(dbits(m.invmw[c]));
      out.prm.push_back(dbits(m.invmw[d]));
      out.prm.push_back(0ull);
      continue;
    }
    if (kind[r] == 1)
    {
      // w0: flags | reaction << 14 | record row << 32; w1: counts; w2..4: A, b, Ea/R; w5..7: reactant and product
      // entries (idx 16 | coefficient 8 | slot 8), 32 bits each; w8..10: net entries (idx 16 | nu 8); w11: pad;
      // third-body / falloff types continue with w12: default efficiency, w13..15: low-pressure A, b, Ea/R,
      // w16..19: Troe parameters, then (idx 16 | slot 8 << 24, efficiency) per third body
      out.prm.push_back((unsigned long long)((unsigned int)flags[r] | ((unsigned int)r << 14)) |
                        ((unsigned long long)(unsigned int)rec_off[r] << 32));
      unsigned long long w1 = 0;
      w1 |= (unsigned long long)(x.n_rc & 255);
      w1 |= (unsigned long long)((x.reversible ? x.n_pd : 0) & 255) << 8;
      w1 |= (unsigned long long)(x.n_net & 255) << 16;
      w1 |= (unsigned long long)(ntb & 255) << 24;
      w1 |= (unsigned long long)(nsl & 255) << 32;
      w1 |= (unsigned long long)((unsigned char)(signed char)x.sum_stoich) << 40;
      w1 |= (unsigned long long)(x.sum_rc & 255) << 48;
      w1 |= (unsigned long long)(x.sum_pd & 255) << 56;
      out.prm.push_back(w1);
      for (int k = 0; k < 3; ++k)
        out.prm.push_back(dbits(x.kf[k]));
      unsigned int e[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 0; i < x.n_rc; ++i)
        e[i] = (unsigned int)(unsigned short)x.rc_idx[i] | ((unsigned int)(unsigned char)x.rc_st[i] << 16) |
               ((unsigned int)(unsigned char)rc_slot[NSR * (size_t)r + i] << 24);
      if (x.reversible)
        for (int i = 0; i < x.n_pd; ++i)
          e[3 + i] = (unsigned int)(unsigned short)x.pd_idx[i] | ((unsigned int)(unsigned char)x.pd_st[i] << 16) |
                     ((unsigned int)(unsigned char)pd_slot[NSR * (size_t)r + i] << 24);
      for (int i = 0; i < 6; i += 2)
        out.prm.push_back((unsigned long long)e[i] | ((unsigned long long)e[i + 1] << 32));
      unsigned int ne[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 6; ++i)
      { // unused entries repeat species 0 with nu = 0 (adds an exact zero)
        const int idx = i < x.n_net ? x.net_idx[i] : x.net_idx[0];
        const int nu = i < x.n_net ? x.net_st[i] : 0;
        ne[i] = (unsigned int)(unsigned short)idx | ((unsigned int)(unsigned char)(signed char)nu << 16);
      }
      for (int i = 0; i < 6; i += 2)
        out.prm.push_back((unsigned long long)ne[i] | ((unsigned long long)ne[i + 1] << 32));
      // w11: index of the reaction among the structured third-body / falloff reactions (k_jac4 evaluates their
      // factors ahead of the reaction phase, falloff_task in gb_react.cuh)
      out.fidx[r] = x.type != RT_SIMPLE ? out.n_falloff++ : -1;
      out.prm.push_back((unsigned long long)(unsigned int)std::max(0, out.fidx[r]));
      if (x.type != RT_SIMPLE)
      {
        out.prm.push_back(dbits(x.base_eff));
        for (int k = 0; k < 3; ++k)
          out.prm.push_back(dbits(x.kp[k]));
        for (int k = 0; k < 4; ++k)
          out.prm.push_back(dbits(x.troe[k]));
        for (int j = 0; j < ntb; ++j)
        {
          out.prm.push_back((unsigned long long)(unsigned short)x.tb_idx[j] |
                            ((unsigned long long)(unsigned char)tb_slot[tb_off[r] + j] << 24));
          out.prm.push_back(dbits(x.tb_eff[j]));
        }
      }
      continue;
    }
    out.prm.push_back((unsigned long long)((unsigned int)flags[r] | ((unsigned int)r << 14)) |
                        ((unsigned long long)(unsigned int)rec_off[r] << 32));
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
    for (int i = 0; i < x.n_net; ++i) // net species: index | nu
      out.prm.push_back((unsigned long long)(unsigned short)x.net_idx[i] |
                        ((unsigned long long)(unsigned char)(signed char)x.net_st[i] << 16));
    for (int j = 0; j < ntb; ++j)
    {
      out.prm.push_back((unsigned long long)(unsigned short)x.tb_idx[j] |
                        ((unsigned long long)(unsigned char)tb_slot[tb_off[r] + j] << 24));
      out.prm.push_back(dbits(x.tb_eff[j]));
    }
  }
  for (int k = 0; k < 66; ++k)
    out.prm.push_back(0ull); // the group prefetch touches up to 512 bytes past the start of the last record

  // ---- logical destinations and their items ----------------------------------------------------------------------------------
  // [0, ns*(ns-1))       R[i][k], logical id k*ns + i
  // [rbase, rbase+5*ns)  row scalars q*ns + i, q = 0..4: sums of nu * {q, dq/drho, dq/dT, a, b}
  const int yend = ns * (ns - 1), rbase = yend, nlogical = rbase + 5 * ns;
  std::vector<std::vector<unsigned int>> &dest = out.dest;
  dest.assign(nlogical, std::vector<unsigned int>());
  for (int r = 0; r < nr; ++r)
  {
    const HostReaction &x = m.reactions[r];
    const int base = rec_off[r], nsl = slot_off[r + 1] - slot_off[r], hdr = hdr_of[r];
    const bool tbtype = x.type != RT_SIMPLE;
    for (int k = 0; k < x.n_net; ++k)
    {
      const int row = x.net_idx[k], nu = x.net_st[k];
      if (nu < -128 || nu > 127)
      {
        set_error("net stoichiometric coefficient outside [-128, 127]");
        return GB_ERR_UNSUPPORTED;
      }
      auto item = [&](int rec) { return (unsigned int)rec | ((unsigned int)(nu & 255) << 16); };
      for (int q = 0; q < 3; ++q)
        dest[rbase + q * ns + row].push_back(item(base + q));
      if (tbtype && hdr == JP_HDR_GEN)
        dest[rbase + 3 * ns + row].push_back(item(base + 3));
      if (last_involved[r] && hdr == JP_HDR_GEN)
        dest[rbase + 4 * ns + row].push_back(item(base + 4));
      for (int q = 0; q < nsl; ++q)
        dest[(int)slot_species[slot_off[r] + q] * ns + row].push_back(item(base + hdr + q));
    }
  }

  return GB_OK;
}

int build_jac_plan(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                   const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                   const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                   const std::vector<int> &tb_off, int G, int threads, JacPlanHost &out)
{
  const int ns = (int)m.species.size(), nr = (int)m.reactions.size(), last = ns - 1;
  out = JacPlanHost();
  out.G = G;
  out.threads = threads;
  const int nwarps = threads / 32, LPR = 32 / G, RMAX = 32 / G;

  PlanCommon pc;
  {
    const int rc = build_plan_common(m, flags, slot_off, slot_species, rc_slot, pd_slot, tb_slot, tb_off, pc);
    if (rc != GB_OK)
      return rc;
  }
  const std::vector<char> &fast = pc.fast, &last_involved = pc.last_involved, &kind = pc.kind;
  const std::vector<int> &hdr_of = pc.hdr_of, &rec_off = pc.rec_off, &prm_off = pc.prm_off;
  const int rec_rows = pc.rec_rows;
  out.rec_rows = rec_rows;
  out.prm = pc.prm;
  (void)last_involved;
  (void)hdr_of;
  (void)rec_off;
  (void)last;

  // ---- reaction groups --------------------------------------------------------------------------------------------------
  {
    auto key = [&](int r) {
      const HostReaction &x = m.reactions[r];
      long k = kind[r];
      k = k * 8 + x.type;
      k = k * 8 + (x.kform == KF_ARRHENIUS ? 0 : 1 + x.kform);
      k = k * 2 + (x.reversible ? 0 : 1);
      k = k * 2 + (x.has_orders ? 1 : 0);
      if (!fast[r])
      {
        k = k * 16 + x.n_rc;
        k = k * 16 + x.n_pd;
        k = k * 8 + x.troebits;
        k = k * 64 + std::min<int>(63, (int)x.tb_idx.size());
      }
      return k;
    };
    auto cost = [&](int r) { // measured group times on B200 (tools/timeline.py, GRI-3.0), cycles
      const HostReaction &x = m.reactions[r];
      double c;
      if (kind[r] == 0)
        c = 600. + (x.kform == KF_ARRHENIUS ? 300. : 0.) + (x.reversible ? 460. : 0.);
      else
      {
        c = x.type == RT_SIMPLE ? 4200. : (x.type == RT_THIRD_BODY ? 7900. : (x.type == RT_LINDEMANN ? 10000. : 13600.));
        if (kind[r] == 2)
          c = 1.5 * c + (x.has_orders ? 6000. : 0.);
      }
      return c;
    };
    std::vector<int> order(nr);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
    struct Group
    {
      int kind;
      double cost;
      std::vector<int> rx;
    };
    std::vector<Group> groups;
    for (int p = 0; p < nr;)
    {
      Group g;
      g.kind = kind[order[p]];
      g.cost = 0.;
      while (p < nr && (int)g.rx.size() < LPR && kind[order[p]] == g.kind)
      {
        g.cost = std::max(g.cost, cost(order[p]));
        g.rx.push_back(order[p]);
        ++p;
      }
      // lanes of a generic group diverge: the warp pays for (part of) the union of the code paths
      if (g.kind != 0)
      {
        double extra = 0.;
        for (size_t i = 1; i < g.rx.size(); ++i)
          if (key(g.rx[i]) / 64 != key(g.rx[i - 1]) / 64) // same code path up to the number of third bodies
            extra += 0.25 * cost(g.rx[i]);
        g.cost += extra;
      }
      groups.push_back(g);
    }
    std::vector<int> gorder(groups.size());
    std::iota(gorder.begin(), gorder.end(), 0);
    std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) { return groups[a].cost > groups[b].cost; });
    std::vector<double> load(nwarps, 0.);
    // (Round 2, measured and NOT adopted: the timeline shows ~450 cycles of loop overhead per group and ~3,800 cycles
    // for the last warp's chain, which this model leaves out -- warp arrivals at the end of the phase spread from 11.5 k
    // to 23.5 k cycles. Adding both makes the arrivals even and the kernel 0.9 % SLOWER (6.76 against 6.70 ms, paired
    // runs): the phase is bound by the issue slots of the schedulers, not by its longest warp.)
    load[nwarps - 1] = 1500.; // the last warp starts with the mixture cp chain (k_jac)
    std::vector<std::vector<int>> per_warp(nwarps);
    for (int gi : gorder)
    {
      const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      per_warp[w].push_back(gi);
      load[w] += groups[gi].cost;
    }
    out.wg_off.assign(nwarps + 1, 0);
    for (int w = 0; w < nwarps; ++w)
    {
      out.wg_off[w] = (int)out.groups.size() / (1 + LPR);
      for (int gi : per_warp[w])
      {
        out.groups.push_back(groups[gi].kind);
        for (int i = 0; i < LPR; ++i)
          out.groups.push_back(i < (int)groups[gi].rx.size() ? prm_off[groups[gi].rx[i]] : -1);
      }
    }
    out.wg_off[nwarps] = (int)out.groups.size() / (1 + LPR);
    for (int r = 0; r < nr; ++r)
      (kind[r] == 0 ? out.n_fast : (kind[r] == 1 ? out.n_struct : out.n_generic))++;
    if (getenv("GB_PLAN_VERBOSE"))
    {
      fprintf(stderr, "[gb plan] G=%d threads=%d reactions: %d fast, %d structured, %d generic, %d groups; warp loads:", G,
              threads, out.n_fast, out.n_struct, out.n_generic, (int)groups.size());
      for (int w = 0; w < nwarps; ++w)
        fprintf(stderr, " %.0f", load[w]);
      fprintf(stderr, "\n");
    }
  }

  // ---- logical destinations and their items (build_plan_common) --------------------------------------------------------
  const int yend = ns * (ns - 1), rbase = yend, nlogical = rbase + 5 * ns;
  const std::vector<std::vector<unsigned int>> &dest = pc.dest;

  // ---- rows of the gathered-sum array -----------------------------------------------------------------------------------------
  std::vector<int> row_of(nlogical, -1);
  int rows = 0;
  for (int s = 0; s < nlogical; ++s)
    if (!dest[s].empty())
      row_of[s] = rows++;
  out.n_dest = rows;

  int split = 12;
  if (const char *e = std::getenv("GB_JAC_SPLIT"))
    split = std::max(2, std::atoi(e));
  struct Part
  {
    int logical, row, begin, count;
  };
  std::vector<Part> parts;
  for (int s = 0; s < nlogical; ++s)
  {
    if (row_of[s] < 0)
      continue;
    const int n = (int)dest[s].size();
    out.n_items += n;
    if (n <= split + split / 2)
      parts.push_back({s, row_of[s], 0, n});
    else
    {
      const int np = (n + split - 1) / split;
      out.fix.push_back(row_of[s]);
      out.fix.push_back(rows);
      out.fix.push_back(np - 1);
      for (int p = 0; p < np; ++p)
      {
        const int b = (int)((long long)n * p / np), e = (int)((long long)n * (p + 1) / np);
        parts.push_back({s, p == 0 ? row_of[s] : rows + p - 1, b, e - b});
      }
      rows += np - 1;
    }
  }
  out.n_parts = (int)parts.size();
  out.t0base = rows;
  rows += ns; // final temperature-row values, column c at t0base + c
  out.c0base = rows;
  rows += ns; // final column-0 values of the species rows, row 1+i at c0base + i
  out.rows = rows;
  const int region = std::max(rows, rec_rows);
  out.zrow = region; // always zero; region + 1 is the write-only dummy row of idle lanes
  if (region + 2 > 65535)
  {
    set_error("too many Jacobian plan rows");
    return GB_ERR_UNSUPPORTED;
  }

  // ---- rounds: 32 parts of similar length per round, rounds dealt to warps by LPT ------------------------------------------
  {
    const int BLK = JP_BLK; // steps are issued in blocks (item prefetch distance), so lengths are padded to BLK
    std::vector<int> porder(parts.size());
    std::iota(porder.begin(), porder.end(), 0);
    std::stable_sort(porder.begin(), porder.end(), [&](int a, int b) { return parts[a].count > parts[b].count; });
    struct Round
    {
      int len;
      std::vector<int> p;
    };
    std::vector<Round> rds;
    for (size_t i = 0; i < porder.size(); i += 32)
    {
      Round rd;
      rd.len = ((parts[porder[i]].count + BLK - 1) / BLK) * BLK;
      for (size_t j = i; j < std::min(porder.size(), i + 32); ++j)
        rd.p.push_back(porder[j]);
      rds.push_back(rd);
    }
    if ((int)rds.size() > nwarps * RMAX)
    {
      set_error("Jacobian plan needs more gather rounds than the register-resident accumulators allow for this tile size");
      return GB_ERR_UNSUPPORTED;
    }
    std::vector<double> load(nwarps, 0.);
    std::vector<std::vector<int>> per_warp(nwarps);
    for (size_t i = 0; i < rds.size(); ++i) // already in descending length
    {
      int w = -1;
      for (int c = 0; c < nwarps; ++c)
        if ((int)per_warp[c].size() < RMAX && (w < 0 || load[c] < load[w]))
          w = c;
      per_warp[w].push_back((int)i);
      load[w] += rds[i].len + 3.;
    }
    out.wr_off.assign(nwarps + 1, 0);
    for (int w = 0; w < nwarps; ++w)
    {
      out.wr_off[w] = (int)out.rounds.size() / 2;
      out.max_rounds = std::max(out.max_rounds, (int)per_warp[w].size());
      for (int ri : per_warp[w])
      {
        const Round &rd = rds[ri];
        out.rounds.push_back((int)out.items.size());
        out.rounds.push_back(rd.len);
        out.n_steps += rd.len;
        for (int k = 0; k < rd.len; ++k)
          for (int l = 0; l < 32; ++l)
          {
            unsigned int it = (unsigned int)out.zrow; // nu = 0 on the zero row: a no-op
            if (l < (int)rd.p.size() && k < parts[rd.p[l]].count)
              it = dest[parts[rd.p[l]].logical][parts[rd.p[l]].begin + k];
            out.items.push_back(it);
          }
        for (int l = 0; l < 32; ++l)
        {
          out.rdest.push_back((unsigned short)(l < (int)rd.p.size() ? parts[rd.p[l]].row : out.zrow + 1));
          // species whose row the destination belongs to (its factor -nu*M_i multiplies every item)
          const int lg = l < (int)rd.p.size() ? parts[rd.p[l]].logical : 0;
          out.rspec.push_back((unsigned short)(lg < yend ? lg % ns : (lg - rbase) % ns));
        }
      }
    }
    out.wr_off[nwarps] = (int)out.rounds.size() / 2;
    for (int k = 0; k < 32 * (3 * BLK + 2); ++k)
      out.items.push_back((unsigned int)out.zrow); // the prefetch runs up to three blocks past the end
    if (getenv("GB_PLAN_VERBOSE"))
    {
      fprintf(stderr, "[gb plan] rec_rows=%d rows=%d dests=%d parts=%d items=%d steps*32=%d rounds=%d fix=%d; warp steps:",
              rec_rows, rows, out.n_dest, out.n_parts, out.n_items, out.n_steps * 32, (int)rds.size(),
              (int)out.fix.size() / 3);
      for (int w = 0; w < nwarps; ++w)
        fprintf(stderr, " %.0f", load[w]);
      fprintf(stderr, "\n");
    }
  }

  // ---- row-scalar sources, column sums, output map ---------------------------------------------------------------------------------
  out.rowsrc.assign(5 * (size_t)ns, (unsigned short)out.zrow);
  for (int q = 0; q < 5; ++q)
    for (int i = 0; i < ns; ++i)
      if (row_of[rbase + q * ns + i] >= 0)
        out.rowsrc[(size_t)q * ns + i] = (unsigned short)row_of[rbase + q * ns + i];
  // column destinations: 0..ns-2: sum_i h_i R[i][k]; then sums over species of h_i * {W, Wrho, WT, A, B}_i and of
  // cp_i * W_i (the gathered rows already carry the factor -nu*M_i); items in species order
  out.ncs = ns - 1 + 6;
  out.cs_off.assign(1, 0);
  for (int k = 0; k < ns - 1; ++k)
  {
    for (int i = 0; i < ns; ++i)
      if (row_of[k * ns + i] >= 0)
        out.cs_items.push_back((unsigned int)row_of[k * ns + i] | ((unsigned int)i << 16));
    out.cs_off.push_back((int)out.cs_items.size());
  }
  for (int q = 0; q < 6; ++q)
  {
    const int src = q < 5 ? q : 0;
    for (int i = 0; i < ns; ++i)
      if (row_of[rbase + src * ns + i] >= 0)
        out.cs_items.push_back((unsigned int)row_of[rbase + src * ns + i] | ((unsigned int)i << 16));
    out.cs_off.push_back((int)out.cs_items.size());
  }
  out.cs_items.push_back(0u);
  // column sums are formed by one lane per part (G accumulators in registers); parts of a destination are added
  // in part order afterwards
  std::vector<int> csparts, cspfirst(1, 0);
  {
    int csplit = 8;
    if (const char *e = std::getenv("GB_JAC_CSPLIT"))
      csplit = std::max(2, std::atoi(e));
    for (int d = 0; d < out.ncs; ++d)
    {
      const int b0 = out.cs_off[d], n = out.cs_off[d + 1] - b0;
      const int np = std::max(1, (n + csplit - 1) / csplit);
      for (int q = 0; q < np; ++q)
      {
        csparts.push_back(d);
        csparts.push_back(b0 + (int)((long long)n * q / np));
        csparts.push_back(b0 + (int)((long long)n * (q + 1) / np));
      }
      cspfirst.push_back((int)csparts.size() / 3);
    }
    out.ncsp = (int)csparts.size() / 3;
  }
  // ---- small tables copied to shared memory by every CTA -----------------------------------------------------------------
  {
    auto add = [&](const std::vector<int> &v) {
      const int off = (int)out.tab.size();
      out.tab.insert(out.tab.end(), v.begin(), v.end());
      return off;
    };
    auto add16 = [&](const std::vector<unsigned short> &v) {
      const int off = (int)out.tab.size();
      for (size_t i = 0; i < v.size(); i += 2)
        out.tab.push_back((int)((unsigned int)v[i] | ((unsigned int)(i + 1 < v.size() ? v[i + 1] : 0) << 16)));
      return off;
    };
    out.t_wg = add(out.wg_off);
    out.t_groups = add(out.groups);
    out.t_wr = add(out.wr_off);
    out.t_rounds = add(out.rounds);
    out.t_rdest = add16(out.rdest);
    out.t_rspec = add16(out.rspec);
    out.t_fix = add(out.fix);
    out.t_rowsrc = add16(out.rowsrc);
    out.t_csparts = add(csparts);
    out.t_cspfirst = add(cspfirst);
    out.t_csitems = add(std::vector<int>(out.cs_items.begin(), out.cs_items.end()));
    if (out.tab.size() & 1)
      out.tab.push_back(0);
  }
  out.emap.assign((size_t)(ns + 1) * (ns - 1), (unsigned short)out.zrow);
  for (int c = 1; c < ns; ++c)
  {
    out.emap[(size_t)(ns + 1) * (c - 1)] = (unsigned short)(out.t0base + c);
    for (int r = 1; r <= ns; ++r)
      if (row_of[(c - 1) * ns + (r - 1)] >= 0)
        out.emap[(size_t)(ns + 1) * (c - 1) + r] = (unsigned short)row_of[(c - 1) * ns + (r - 1)];
  }
  return GB_OK;
}

} // namespace gb
