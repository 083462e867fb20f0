// gb_plan4.cu -- host-side schedule of k_jac4 (gb_jac4.cu), the warp-specialised reactor-Jacobian kernel.
//
// Same decomposition as the plan of k_jac (gb_plan.cu: every reaction writes a compact record, every structurally
// non-zero entry of R[i][k] = sum_r nu_ri dq_r/dY_k and every row scalar is a list of (record row, nu) items in ascending
// reaction order), scheduled for tiles of FOUR states whose Jacobian blocks are assembled in shared memory in their
// final column-major layout and leave the SM by one bulk copy:
//   * consumer warps: reaction groups of 8 reactions x 4 states (third-body / falloff factors arrive precomputed),
//     gather rounds of 32 destination parts that write straight into the Jacobian tile / the row-scalar array,
//     column jobs of the output transform;
//   * producer warps (one tile ahead): thermodynamics, concentrations and the third-body / falloff factor tasks.
// Destination codes (u16): [0, ns*ns) entry c*ns + r of the state's block (r = 0 temporarily holds the row of the last
// species, which only enters the temperature row); [ns*ns, ns*ns + 5*ns) row scalars q*ns + i; above: extra parts of
// split destinations (added in part order by their consumers).
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

size_t jac4_smem_bytes(int ns, const JacPlan4Host &p)
{
  const size_t G = 4;
  size_t doubles = 2 * (size_t)p.bufsz;                      // double-buffered producer output
  doubles += (size_t)ns * G;                                 // dcp_i/dT (producer scratch)
  doubles += 3 * (size_t)ns + ((4 - (3 * ns) % 4) & 3);       // u, -M, 1/M (padded: 32-byte aligned rows follow)
  doubles += (size_t)p.nwx * G;                              // row scalars + extra parts
  doubles += 2 * (size_t)ns * G + 16 * G;                     // c2, c3, per-state scalars of the transform
  doubles += (size_t)p.rec_rows * G;                         // reaction records
  doubles += G * (size_t)ns * ns;                            // Jacobian tile
  return doubles * sizeof(double) + sizeof(int) * p.tab.size() + 16;
}

int build_jac4_plan(const HostMech &m, const std::vector<int> &flags, const std::vector<int> &slot_off,
                    const std::vector<short> &slot_species, const std::vector<signed char> &rc_slot,
                    const std::vector<signed char> &pd_slot, const std::vector<signed char> &tb_slot,
                    const std::vector<int> &tb_off, int ncons, int nprod, JacPlan4Host &out)
{
  const int ns = (int)m.species.size(), nr = (int)m.reactions.size(), last = ns - 1, nsns = ns * ns;
  const int G = 4, LPR = 8;
  out = JacPlan4Host();
  out.ncons = ncons, out.nprod = nprod, out.threads = 32 * (ncons + nprod);
  PlanCommon pc;
  {
    const int rc = build_plan_common(m, flags, slot_off, slot_species, rc_slot, pd_slot, tb_slot, tb_off, pc);
    if (rc != GB_OK)
      return rc;
  }
  out.rec_rows = pc.rec_rows + 1; // + one always-zero row for padding items
  const int zrow = pc.rec_rows;
  out.nF = pc.n_falloff;
  out.bufsz = G * (JP_NSC + 6 * ns + 4 * std::max(1, pc.n_falloff));

  // ---- consumer reaction groups (8 reactions x 4 states), dealt to the consumer warps by estimated cost -------------
  std::vector<int> wg_off, groups;
  {
    auto key = [&](int r) {
      const HostReaction &x = m.reactions[r];
      long k = pc.kind[r];
      k = k * 8 + x.type;
      k = k * 8 + (x.kform == KF_ARRHENIUS ? 0 : 1 + x.kform);
      k = k * 2 + (x.reversible ? 0 : 1);
      k = k * 2 + (x.has_orders ? 1 : 0);
      if (!pc.fast[r])
      {
        k = k * 16 + x.n_rc;
        k = k * 16 + x.n_pd;
      }
      return k;
    };
    auto cost = [&](int r) { // cycles per group, measured with tiles of four states (tools/timeline.py)
      const HostReaction &x = m.reactions[r];
      if (pc.kind[r] == 0)
        return 700. + (x.kform == KF_ARRHENIUS ? 300. : 0.) + (x.reversible ? 460. : 0.);
      if (pc.kind[r] == 1)
        return x.type == RT_SIMPLE ? 6000. : 7000.;
      return (x.type == RT_SIMPLE ? 9000. : 16000.) + (x.has_orders ? 6000. : 0.);
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
    std::vector<Group> gs;
    for (int p = 0; p < nr;)
    {
      Group g;
      g.kind = pc.kind[order[p]];
      g.cost = 0.;
      while (p < nr && (int)g.rx.size() < LPR && pc.kind[order[p]] == g.kind)
      {
        g.cost = std::max(g.cost, cost(order[p]));
        g.rx.push_back(order[p]);
        ++p;
      }
      gs.push_back(g);
    }
    std::vector<int> gorder(gs.size());
    std::iota(gorder.begin(), gorder.end(), 0);
    std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) { return gs[a].cost > gs[b].cost; });
    std::vector<double> load(ncons, 0.);
    std::vector<std::vector<int>> per_warp(ncons);
    for (int gi : gorder)
    {
      const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      per_warp[w].push_back(gi);
      load[w] += gs[gi].cost;
    }
    wg_off.assign(ncons + 1, 0);
    for (int w = 0; w < ncons; ++w)
    {
      wg_off[w] = (int)groups.size() / (1 + LPR);
      for (int gi : per_warp[w])
      {
        groups.push_back(gs[gi].kind);
        for (int i = 0; i < LPR; ++i)
          groups.push_back(i < (int)gs[gi].rx.size() ? pc.prm_off[gs[gi].rx[i]] : -1);
      }
    }
    wg_off[ncons] = (int)groups.size() / (1 + LPR);
    for (int r = 0; r < nr; ++r)
      (pc.kind[r] == 0 ? out.n_fast : (pc.kind[r] == 1 ? out.n_struct : out.n_generic))++;
    if (getenv("GB_PLAN_VERBOSE"))
    {
      fprintf(stderr, "[gb plan4] consumers=%d producers=%d reactions: %d fast, %d structured (%d with factor), %d generic, "
                      "%d groups; warp loads:", ncons, nprod, out.n_fast, out.n_struct, out.nF, out.n_generic, (int)gs.size());
      for (int w = 0; w < ncons; ++w)
        fprintf(stderr, " %.0f", load[w]);
      fprintf(stderr, "\n");
    }
  }

  // ---- producer factor tasks: 8 reactions x 4 states, Troe first ------------------------------------------------------------
  std::vector<int> fgroups;
  {
    std::vector<int> fr;
    for (int r = 0; r < nr; ++r)
      if (pc.fidx[r] >= 0)
        fr.push_back(r);
    std::stable_sort(fr.begin(), fr.end(), [&](int a, int b) {
      const HostReaction &x = m.reactions[a], &y = m.reactions[b];
      if (x.type != y.type)
        return x.type > y.type;
      return x.troebits > y.troebits;
    });
    for (size_t p = 0; p < fr.size(); p += LPR)
      for (int i = 0; i < LPR; ++i)
        fgroups.push_back(p + i < fr.size() ? pc.prm_off[fr[p + i]] : -1);
    out.nfg = (int)fgroups.size() / LPR;
  }

  // ---- destinations: codes, parts, fix tables --------------------------------------------------------------------------------------
  const int yend = ns * (ns - 1), rbase = yend, nlogical = rbase + 5 * ns;
  auto code_of = [&](int lg) {
    if (lg < yend)
    {
      const int k = lg / ns, i = lg - k * ns;
      return (1 + k) * ns + (i == last ? 0 : 1 + i);
    }
    return nsns + (lg - rbase);
  };
  int split = 12;
  if (const char *e = std::getenv("GB_JAC_SPLIT"))
    split = std::max(2, std::atoi(e));
  struct Part
  {
    int code;
    std::vector<unsigned int> items; // record row | sign << 31, |nu| > 1 repeated
  };
  std::vector<Part> parts;
  int nx = 0; // extra part rows
  std::vector<int> wfix(5 * (size_t)ns, 0);                 // first extra row << 8 | extra parts
  std::vector<std::vector<int>> cfx(ns);                    // per column: (row, first extra row, extra parts)
  for (int lg = 0; lg < nlogical; ++lg)
  {
    const std::vector<unsigned int> &d = pc.dest[lg];
    if (d.empty())
      continue;
    std::vector<unsigned int> ex;
    for (unsigned int it : d)
    {
      const int nu = (int)(signed char)((it >> 16) & 255);
      const unsigned int w = ((it & 0xffffu) * 32u) | (nu < 0 ? 0x80000000u : 0u); // byte offset of the row | sign
      for (int k = 0; k < std::abs(nu); ++k)
        ex.push_back(w);
    }
    out.n_items += (int)ex.size();
    const int n = (int)ex.size(), code = code_of(lg);
    if (n <= split + split / 2)
    {
      parts.push_back({code, ex});
      continue;
    }
    const int np = (n + split - 1) / split;
    if (np - 1 > 255 || nx + np - 1 > 0xffff)
    {
      set_error("Jacobian plan: destination with too many parts");
      return GB_ERR_UNSUPPORTED;
    }
    if (lg < yend)
    {
      const int c = code / ns, r = code - c * ns;
      cfx[c].push_back(r);
      cfx[c].push_back(nx);
      cfx[c].push_back(np - 1);
    }
    else
      wfix[lg - rbase] = (nx << 8) | (np - 1);
    for (int p = 0; p < np; ++p)
    {
      const int b = (int)((long long)n * p / np), e = (int)((long long)n * (p + 1) / np);
      parts.push_back({p == 0 ? code : nsns + 5 * ns + nx + p - 1, std::vector<unsigned int>(ex.begin() + b, ex.begin() + e)});
    }
    nx += np - 1;
  }
  out.nwx = 5 * ns + nx;
  out.n_parts = (int)parts.size();
  if (nsns + out.nwx >= 0xffff || out.rec_rows > 0xffff)
  {
    set_error("mechanism too large for the packed Jacobian plan");
    return GB_ERR_UNSUPPORTED;
  }

  // ---- gather rounds: 32 parts of similar length, dealt to the consumer warps by length ----------------------------------
  std::vector<int> wr_off, rounds;
  {
    const int BLK = 2;
    std::vector<int> porder(parts.size());
    std::iota(porder.begin(), porder.end(), 0);
    std::stable_sort(porder.begin(), porder.end(),
                     [&](int a, int b) { return parts[a].items.size() > parts[b].items.size(); });
    struct Round
    {
      int len;
      std::vector<int> p;
    };
    std::vector<Round> rds;
    for (size_t i = 0; i < porder.size(); i += 32)
    {
      Round rd;
      rd.len = (((int)parts[porder[i]].items.size() + BLK - 1) / BLK) * BLK;
      for (size_t j = i; j < std::min(porder.size(), i + 32); ++j)
        rd.p.push_back(porder[j]);
      rds.push_back(rd);
    }
    std::vector<double> load(ncons, 0.);
    std::vector<std::vector<int>> per_warp(ncons);
    for (size_t i = 0; i < rds.size(); ++i)
    {
      const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      per_warp[w].push_back((int)i);
      load[w] += rds[i].len + 3.;
    }
    wr_off.assign(ncons + 1, 0);
    for (int w = 0; w < ncons; ++w)
    {
      wr_off[w] = (int)rounds.size() / 2;
      for (int ri : per_warp[w])
      {
        const Round &rd = rds[ri];
        rounds.push_back((int)out.items.size());
        rounds.push_back(rd.len);
        out.n_steps += rd.len;
        // items of a round: [step pair][lane][2]
        for (int k = 0; k < rd.len; k += 2)
          for (int l = 0; l < 32; ++l)
            for (int q = 0; q < 2; ++q)
            {
              unsigned int it = (unsigned int)zrow * 32u;
              if (l < (int)rd.p.size() && k + q < (int)parts[rd.p[l]].items.size())
                it = parts[rd.p[l]].items[k + q];
              out.items.push_back(it);
            }
        for (int l = 0; l < 32; ++l)
        {
          unsigned int code = 0xffffu, spec = 0;
          if (l < (int)rd.p.size())
          {
            code = (unsigned int)parts[rd.p[l]].code;
            // species whose row the destination belongs to (its factor -M_i multiplies every item)
            spec = 0;
          }
          out.rdest.push_back(code | (spec << 16));
        }
      }
    }
    wr_off[ncons] = (int)rounds.size() / 2;
    for (int k = 0; k < 64 * 6; ++k)
      out.items.push_back((unsigned int)zrow * 32u); // the item prefetch runs up to three blocks past the end
    if (getenv("GB_PLAN_VERBOSE"))
    {
      fprintf(stderr, "[gb plan4] rec_rows=%d parts=%d items=%d steps*32=%d rounds=%d extra parts=%d; warp steps:", out.rec_rows,
              out.n_parts, out.n_items, out.n_steps * 32, (int)rds.size(), nx);
      for (int w = 0; w < ncons; ++w)
        fprintf(stderr, " %.0f", load[w]);
      fprintf(stderr, "\n");
    }
  }
  // species of every part's destination: recover from the code
  {
    // code -> species row: entries c*ns + r: r == 0 -> last species, else r - 1; row scalars q*ns + i -> i; extra parts
    // inherit the species of their destination
    std::vector<int> xspec(std::max(1, nx), 0);
    for (int q = 0; q < 5 * ns; ++q)
      if (wfix[q])
        for (int p = 0; p < (wfix[q] & 255); ++p)
          xspec[(wfix[q] >> 8) + p] = q % ns;
    for (int c = 0; c < ns; ++c)
      for (size_t e = 0; e < cfx[c].size(); e += 3)
        for (int p = 0; p < cfx[c][e + 2]; ++p)
          xspec[cfx[c][e + 1] + p] = cfx[c][e] == 0 ? last : cfx[c][e] - 1;
    for (unsigned int &w : out.rdest)
    {
      const int code = (int)(w & 0xffffu);
      int spec = 0;
      if (code == 0xffff)
        spec = 0;
      else if (code < nsns)
      {
        const int r = code % ns;
        spec = r == 0 ? last : r - 1;
      }
      else if (code < nsns + 5 * ns)
        spec = (code - nsns) % ns;
      else
        spec = xspec[code - nsns - 5 * ns];
      w = (unsigned int)code | ((unsigned int)spec << 16);
    }
  }

  // ---- small tables (shared memory) ------------------------------------------------------------------------------------------------------
  {
    auto add = [&](const std::vector<int> &v) {
      const int off = (int)out.tab.size();
      out.tab.insert(out.tab.end(), v.begin(), v.end());
      return off;
    };
    out.t_wg = add(wg_off);
    out.t_groups = add(groups);
    out.t_fgroups = add(fgroups);
    out.t_wr = add(wr_off);
    out.t_rounds = add(rounds);
    out.t_wfix = add(wfix);
    std::vector<int> cfx_off(1, 0), cfx_flat;
    for (int c = 0; c < ns; ++c)
    {
      cfx_flat.insert(cfx_flat.end(), cfx[c].begin(), cfx[c].end());
      cfx_off.push_back((int)cfx_flat.size() / 3);
    }
    cfx_flat.push_back(0);
    out.t_cfxoff = add(cfx_off);
    out.t_cfx = add(cfx_flat);
    // entries of the columns 1..ns-1 without a destination: structural zeros, written by the gather phase
    {
      std::vector<char> has((size_t)nsns, 0);
      for (int lg = 0; lg < yend; ++lg)
        if (!pc.dest[lg].empty())
          has[code_of(lg)] = 1;
      for (int e = ns; e < nsns; ++e)
        if (!has[e])
          out.zlist.push_back((unsigned short)e);
      out.nzero = (int)out.zlist.size();
      out.zlist.push_back(0);
    }
    if (out.tab.size() & 1)
      out.tab.push_back(0);
  }
  return GB_OK;
}

} // namespace gb
