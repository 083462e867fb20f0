// gb_jac4.cu -- k_jac4: isobaric-reactor RHS + analytical Jacobian for large mechanisms, FP64, sm_100a.
//
// Same mathematics and the same per-reaction device functions as k_jac (gb_jac.cu / gb_react.cuh; replaces
// prod_rates_sens_exact, rates_sensitivities_exact.cpp:33-1028, and chem/mass/heat_jac_isobaric +
// transform_isobaric_primitive_jacobian, isobaric_reactor_kernels.cpp:58-343), organised differently:
//
//   * a CTA owns a tile of FOUR states whose ns x ns Jacobian blocks are assembled in shared memory in their final
//     column-major layout (4 * ns^2 doubles, contiguous like the four blocks in HBM) and leave the SM by ONE bulk
//     asynchronous copy (cp.async.bulk shared -> global) that drains while the next tile is being computed;
//   * warp specialisation: PRODUCER warps run one tile ahead and prepare everything that only depends on the state --
//     thermodynamic polynomials, the order-sensitive mixture sums, concentrations and the third-body / Lindemann / Troe
//     factors with their long dependent chains of transcendentals -- into a double-buffered set of arrays; CONSUMER
//     warps evaluate the mass-action part of the reactions into records, gather the records straight into the
//     Jacobian tile (static plan, gb_plan4.cu: no atomics, ascending reaction order), and apply the output transform
//     J = c1_row * R + u_col * c2_row + c3_row in place, forming the temperature row's inner products on the way.
//     Producers and consumers meet at named barriers (bar.sync / bar.arrive), never at a CTA-wide barrier.
//
// HBM traffic is the algorithmic minimum: ns doubles in, ns + ns^2 doubles out per state.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "gb_react.cuh"

namespace gb
{

extern std::atomic<long> g_jac_launches;

namespace
{
constexpr int G4 = 4;
enum NamedBarrier : int
{
  BAR_CONS = 1,  // consumer warps only
  BAR_PROD = 2,  // producer warps only
  BAR_FULL0 = 3, // + buffer: producer output of a tile is complete (producers arrive, consumers wait)
  BAR_EMPTY0 = 5 // + buffer: consumers are done with a buffer (consumers arrive, producers wait)
};
// per-state scalars of the transform phase, [8][4]
enum TScalar : int
{
  TS_INVRHO = 0,
  TS_INVRHOCP,
  TS_KY,    // -rhs0c/cp + cextra: coefficient of (cp_k - cp_ns) in the temperature row
  TS_SA,
  TS_SB,
  TS_NRM,   // -rho * M_mix
  TS_P0RHO,
  TS_SPARE
};

__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

#ifdef GB_JAC_TIMELINE
__device__ long long g_jac4_timeline[16 * 32];
#define TL4(k)                                                                     \
  if (blockIdx.x == 0 && kt == 2 && lane == 0)                                     \
    g_jac4_timeline[(k)*32 + warp] = clock64();
#else
#define TL4(k)
#endif

struct Bufs
{
  double *sc, *sy, *sC, *sg, *sdb, *sh, *scp, *sF;
};
__device__ __forceinline__ Bufs buf_ptrs(double *B, int ns)
{
  Bufs b;
  b.sc = B;
  b.sy = b.sc + JP_NSC * G4;
  b.sC = b.sy + ns * G4;
  b.sg = b.sC + ns * G4;
  b.sdb = b.sg + ns * G4;
  b.sh = b.sdb + ns * G4;
  b.scp = b.sh + ns * G4;
  b.sF = b.scp + ns * G4;
  return b;
}
} // namespace

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_jac4(const ChemArgs a)
{
  constexpr int G = G4;
  extern __shared__ __align__(16) double smem[];
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, nsm1 = ns - 1, nsns = ns * ns;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncons = dm.j4_ncons, nprod = NT / 32 - ncons;
  const int nct = ncons * 32, npt = nprod * 32;

  double *const bufs = smem;
  const int bufsz = dm.j4_bufsz;
  double *const sdcp = bufs + 2 * bufsz;
  double *const su = sdcp + ns * G, *const snm = su + ns, *const sim = snm + ns;
  double *const sWX = sim + ns + (ns & 1);
  double *const sc2 = sWX + (size_t)dm.j4_nwx * G, *const sc3 = sc2 + ns * G, *const sS = sc3 + ns * G;
  double *const sR = sS + 8 * G;
  double *const sJ = sR + (size_t)dm.j4_rec_rows * G;
  int *const stab = (int *)(sJ + (size_t)G * nsns);

  for (int i = tid; i < ns; i += NT)
  {
    su[i] = dm.invmw[i] - dm.invmw[nsm1];
    snm[i] = -dm.netmw[i];
    sim[i] = dm.invmw[i];
  }
  for (int e = tid; e < dm.j4_tab_words; e += NT)
    stab[e] = dm.j4_tab[e];
  if (tid < G)
    sR[(size_t)(dm.j4_rec_rows - 1) * G + tid] = 0.; // the zero row of padding items
  for (int e = tid; e < dm.j4_nwx * G; e += NT)
    sWX[e] = 0.; // row scalars without items (inert species) are never written by the gather: they stay zero
  __syncthreads();

  const int ntiles = (a.n + G - 1) / G;
  const bool open = a.rx.open != 0, isothermal = a.rx.heat_option == 1;
  const double invTau = open ? 1. / a.rx.tau : 0.;

  if (warp >= ncons)
  {
    // ================================================ producers ====================================================
    const int pw = warp - ncons, ptid = tid - nct;
    const int *t_fgroups = stab + dm.j4_t_fgroups;
    int kt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kt)
    {
      const int b = kt & 1;
      const Bufs B = buf_ptrs(bufs + (size_t)b * bufsz, ns);
      JacSmem s;
      s.sc = B.sc, s.sy = B.sy, s.sC = B.sC, s.sg = B.sg, s.sdb = B.sdb, s.sh = B.sh, s.scp = B.scp, s.sdcp = sdcp;
      s.su = su, s.snm = snm, s.sim = sim, s.sR = sR, s.sTH = nullptr, s.semap = nullptr, s.sF = B.sF;
      const int tile0 = tile * G, gcount = min(G, a.n - tile0);
      if (kt >= 2)
        nbar_sync(BAR_EMPTY0 + b, NT);
      // ---- load (states past the end of the batch replicate the tile's first state; they are never written) -------
      for (int item = ptid; item < G * ns; item += npt)
      {
        const int g = item / ns, j = item - g * ns;
        const double v = a.in_state[(size_t)(tile0 + (g < gcount ? g : 0)) * ns + j];
        if (j != 0)
          SMG(s.sy, j - 1, g) = v;
        else
        {
          SMG(s.sc, J_T, g) = v;
          SMG(s.sc, J_LOGT, g) = log(v);
          SMG(s.sc, J_INVT, g) = 1. / v;
        }
      }
      nbar_sync(BAR_PROD, npt);
      // ---- order-sensitive chains (first producer warp) next to the thermodynamic polynomials (the others) ----------
      if (pw == 0)
      {
        // lanes [0, G): Y_ns = 1 - sum_j Y_j (extract_y, combustion_kernels.h:505-515); lanes [G, 2G): sum_i Y_i/M_i over
        // all but the last species (mixture_molecular_weight, :381-387)
        double d = 0.;
        if (lane < 2 * G)
        {
          const bool first = lane < G;
          const int g = first ? lane : lane - G;
          d = first ? 1. : 0.;
#pragma unroll 4
          for (int j = 0; j < nsm1; ++j)
            d = d + (first ? -1. : sim[j]) * SMG(s.sy, j, g);
          if (first)
            SMG(s.sy, nsm1, g) = d;
        }
        const double dpart = __shfl_down_sync(0xffffffffu, d, G);
        if (lane < G)
        {
          const int g = lane;
          const double mmw = 1. / (dpart + sim[nsm1] * d), T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g);
          const double rho = a.p * mmw / (T * dm.Ru); // ideal_gas_density (:526-530)
          const double invM = 1. / mmw, ct = rho * invM;
          SMG(s.sc, J_MMW, g) = mmw;
          SMG(s.sc, J_INVM, g) = invM;
          SMG(s.sc, J_CT, g) = ct;
          SMG(s.sc, J_RHO, g) = rho;
          SMG(s.sc, J_IRHO, g) = 1. / rho;
          SMG(s.sc, J_DRHOF, g) = 1. / ct * invM;
          SMG(s.sc, J_LPRT, g) = log(dm.p_ref * invT * dm.invRu); // log(p0/(R T)), :535
        }
      }
      else
      {
        for (int item = ptid - 32; item < ns * G; item += npt - 32)
        {
          const int i = item / G, g = item - i * G;
          const SpeciesThermo t = species_thermo<true>(dm, i, SMG(s.sc, J_T, g), SMG(s.sc, J_LOGT, g), SMG(s.sc, J_INVT, g));
          SMG(s.sg, i, g) = t.g;
          SMG(s.sdb, i, g) = t.dB;
          SMG(s.sh, i, g) = t.h;
          SMG(s.scp, i, g) = t.cp;
          SMG(sdcp, i, g) = t.dcp;
        }
      }
      nbar_sync(BAR_PROD, npt);
      // ---- concentrations; mixture cp, dcp/dT and the open-reactor terms in species order (last producer warp) ------
      for (int item = ptid; item < ns * G; item += npt)
      {
        const int i = item / G, g = item - i * G;
        SMG(s.sC, i, g) = SMG(s.sy, i, g) * SMG(s.sc, J_RHO, g) * sim[i];
      }
      if (pw == nprod - 1 && lane < G)
      { // thermodynamics_kernels.cpp:45-131, 183-260
        const int g = lane;
        double cp = 0., dcp = 0.;
        for (int i = 0; i < ns; ++i)
        {
          cp += SMG(s.sy, i, g) * SMG(s.scp, i, g);
          if (dm.cptype[i] == CP_CONST)
            dcp = 0.; // sic, thermodynamics_kernels.cpp:202
          else
            dcp += SMG(s.sy, i, g) * SMG(sdcp, i, g);
        }
        SMG(s.sc, J_CP, g) = cp;
        SMG(s.sc, J_DCP, g) = dcp;
        if (open)
        { // mass_jac_isobaric :100-140: inflow enthalpy term and sum cp_i y_in,i
          const double Tin = a.rx.T_in, logTin = log(Tin), invTin = 1. / Tin;
          double m0;
          {
            const SpeciesThermo tl = species_thermo<false>(dm, nsm1, Tin, logTin, invTin);
            m0 = (tl.h - SMG(s.sh, nsm1, g)) * a.rx.y_in[nsm1];
          }
          for (int i = 0; i < nsm1; ++i)
          {
            const SpeciesThermo ti = species_thermo<false>(dm, i, Tin, logTin, invTin);
            m0 += (ti.h - SMG(s.sh, i, g)) * a.rx.y_in[i];
          }
          m0 /= cp;
          m0 *= invTau;
          double ycp = 0.;
          for (int i = 0; i < ns; ++i)
            ycp += SMG(s.scp, i, g) * a.rx.y_in[i];
          SMG(s.sc, J_M0, g) = m0;
          SMG(s.sc, J_YCP, g) = ycp;
        }
      }
      // ---- third-body / falloff factors: 8 reactions x 4 states per task -----------------------------------------------
      for (int fg = pw; fg < dm.j4_nfg; fg += nprod)
      {
        const int off = t_fgroups[fg * 8 + (lane >> 2)];
        if (off >= 0)
          falloff_task<G>(dm, dm.jp_prm + off, lane & 3, s);
      }
      nbar_arrive(BAR_FULL0 + b, NT);
    }
    return;
  }

  // ================================================== consumers ====================================================
  const int *t_wg = stab + dm.j4_t_wg, *t_groups = stab + dm.j4_t_groups, *t_wr = stab + dm.j4_t_wr;
  const int *t_rounds = stab + dm.j4_t_rounds, *t_wfix = stab + dm.j4_t_wfix, *t_cfxoff = stab + dm.j4_t_cfxoff;
  const int *t_cfx = stab + dm.j4_t_cfx;
  const int nwrow = 5 * ns; // row-scalar rows of sWX; extra parts follow
  int kt = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kt)
  {
    const int b = kt & 1;
    const Bufs B = buf_ptrs(bufs + (size_t)b * bufsz, ns);
    JacSmem s;
    s.sc = B.sc, s.sy = B.sy, s.sC = B.sC, s.sg = B.sg, s.sdb = B.sdb, s.sh = B.sh, s.scp = B.scp, s.sdcp = sdcp;
    s.su = su, s.snm = snm, s.sim = sim, s.sR = sR, s.sTH = nullptr, s.semap = nullptr, s.sF = B.sF;
    const int tile0 = tile * G, gcount = min(G, a.n - tile0);
    TL4(0)
    // ---- the previous tile's bulk copy must have finished reading the Jacobian tile ---------------------------------------
    if (tid == 0)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    nbar_sync(BAR_CONS, nct);
    {
      double2 *z = reinterpret_cast<double2 *>(sJ);
      for (int i = tid; i < G * nsns / 2; i += nct)
        z[i] = make_double2(0., 0.);
    }
    TL4(1)
    nbar_sync(BAR_FULL0 + b, NT);
    TL4(2)
    // ---- reaction phase: records --------------------------------------------------------------------------------------------------
    {
      const int g = lane & 3, sub = lane >> 2;
      const int g0 = t_wg[warp], g1 = t_wg[warp + 1];
      auto prefetch_group = [&](int gi) {
        if (gi < g1)
        {
          const int off = t_groups[gi * 9 + 1 + sub];
          if (off >= 0)
          {
            const char *p = reinterpret_cast<const char *>(dm.jp_prm + off) + g * 128;
#pragma unroll
            for (int k = 0; k < 128; k += 32)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p + k));
          }
        }
      };
      prefetch_group(g0);
      for (int gi = g0; gi < g1; ++gi)
      {
        prefetch_group(gi + 1);
        const int *grp = t_groups + gi * 9;
        const int kind = grp[0], off = grp[1 + sub];
        if (off < 0)
          continue;
        const unsigned long long *P = dm.jp_prm + off;
        if (kind == 0)
          react_fast<G>(dm, P, g, s);
        else if (kind == 1)
          react_struct<G, true>(dm, P, g, s);
        else if (((int)(unsigned int)P[0]) & F_HAS_ORDERS)
          react_orders<G>(ns, dm.n_sp, dm.sp_idx, dm.sp_order, dm.sp_slot, dm.invmw, (int)(((unsigned int)P[0]) >> 14), P, g,
                          s.sc, s.sC, s.sy, s.sR);
        else
          react_generic<G>(dm, P, g, s);
      }
    }
    TL4(3)
    nbar_sync(BAR_CONS, nct);
    TL4(4)
    // ---- gather: lane per destination part, four accumulators, results straight into the Jacobian tile ---------
    {
      const int r0 = t_wr[warp], r1 = t_wr[warp + 1];
      const int rot = lane & 1;
      const uint2 *__restrict__ it =
          reinterpret_cast<const uint2 *>(dm.j4_items) + (r1 > r0 ? t_rounds[2 * r0] / 2 : 0) + lane;
      uint2 cur = __ldg(it), nxt = __ldg(it + 32);
      it += 64;
      for (int r = r0; r < r1; ++r)
      {
        const int L = t_rounds[2 * r + 1];
        const unsigned int rd = __ldg(dm.j4_rdest + (size_t)r * 32 + lane);
        const int code = (int)(rd & 0xffffu);
        const double nm = snm[rd >> 16];
        const int nm_hi = __double2hiint(nm), nm_lo = __double2loint(nm);
        double acc0 = 0., acc1 = 0., acc2 = 0., acc3 = 0.; // slots: states 2*rot, 2*rot+1, 2*(rot^1), 2*(rot^1)+1
        for (int k = 0; k < L; k += 2)
        {
          const uint2 nn = __ldg(it);
          it += 32;
#pragma unroll
          for (int q = 0; q < 2; ++q)
          {
            const unsigned int u = q == 0 ? cur.x : cur.y;
            // factor -nu*M_i of the reference's `wsens += factor * dq` (rates_sensitivities_exact.cpp:1014-1026);
            // |nu| > 1 arrives as repeated items
            const double c = __hiloint2double(nm_hi ^ (int)(u & 0x80000000u), nm_lo);
            const double2 *p = reinterpret_cast<const double2 *>(sR + (size_t)(u & 0xffffu) * G);
            const double2 x = p[rot], y = p[rot ^ 1];
            acc0 = fma(c, x.x, acc0);
            acc1 = fma(c, x.y, acc1);
            acc2 = fma(c, y.x, acc2);
            acc3 = fma(c, y.y, acc3);
          }
          cur = nxt;
          nxt = nn;
        }
        if (code < nsns)
        {
          double *p = sJ + code;
          p[(size_t)(2 * rot) * nsns] = acc0;
          p[(size_t)(2 * rot + 1) * nsns] = acc1;
          p[(size_t)(2 * (rot ^ 1)) * nsns] = acc2;
          p[(size_t)(2 * (rot ^ 1) + 1) * nsns] = acc3;
        }
        else if (code != 0xffff)
        {
          double2 *p = reinterpret_cast<double2 *>(sWX + (size_t)(code - nsns) * G);
          p[rot] = make_double2(acc0, acc1);
          p[rot ^ 1] = make_double2(acc2, acc3);
        }
      }
    }
    TL4(5)
    nbar_sync(BAR_CONS, nct);
    TL4(6)
    // ---- row constants of the transform, column 0, right-hand side; per-state sums for the temperature row -------
    {
      // row scalar q of species i: the gathered value plus its extra parts in part order
      auto wval = [&](int q, int i, int g) {
        double v = sWX[(size_t)(q * ns + i) * G + g];
        const int fx = t_wfix[q * ns + i];
        for (int p = 0; p < (fx & 255); ++p)
          v += sWX[(size_t)(nwrow + (fx >> 8) + p) * G + g];
        return v;
      };
      if (warp == ncons - 1)
      {
        // lanes (q, g): sum_i h_i * {W, Wrho, WT, A, B}_i and sum_i cp_i * W_i in species order, the inner products of
        // isobaric_reactor_kernels.cpp:74-92
        const int q = lane >> 2, g = lane & 3;
        double acc = 0.;
        if (q < 6)
        {
          const double *wsrc = q == 5 ? s.scp : s.sh;
          const int qq = q == 5 ? 0 : q;
          for (int i = 0; i < ns; ++i)
            acc = fma(SMG(wsrc, i, g), wval(qq, i, g), acc);
        }
        const double SW = __shfl_sync(0xffffffffu, acc, g), SWr = __shfl_sync(0xffffffffu, acc, 4 + g);
        const double SWT = __shfl_sync(0xffffffffu, acc, 8 + g), SA = __shfl_sync(0xffffffffu, acc, 12 + g);
        const double SB = __shfl_sync(0xffffffffu, acc, 16 + g), wcp = __shfl_sync(0xffffffffu, acc, 20 + g);
        if (lane < G)
        { // chem_jac_isobaric :58-98, mass_jac_isobaric :100-140, heat_jac_isobaric :142-168, transform :319-343
          const double rho = SMG(s.sc, J_RHO, g), cp = SMG(s.sc, J_CP, g), T = SMG(s.sc, J_T, g);
          const double cpsensT = SMG(s.sc, J_DCP, g);
          const double invRhoCp = 1. / (rho * cp), invRho = SMG(s.sc, J_IRHO, g), invCp = 1. / cp;
          const double rhs0c = -SW * invRhoCp;
          double rhs0 = rhs0c;
          double P0rho = -invRhoCp * SWr - invRho * rhs0c;
          double P0T = -invRhoCp * (SWT + wcp) - rhs0c * cpsensT * invCp;
          double cextra = 0.;
          if (open)
          {
            const double m0 = SMG(s.sc, J_M0, g);
            P0T += -invCp * (cpsensT * m0 + invTau * SMG(s.sc, J_YCP, g));
            cextra += -m0 * invCp;
            rhs0 += m0;
          }
          if (a.rx.heat_option == 2)
          {
            const double Ts = a.rx.T_surf;
            const double rate = a.rx.SoV / (rho * cp) *
                                (a.rx.h_conv * (a.rx.T_inf - T) + a.rx.eps_rad * 5.67e-8 * (Ts * Ts * Ts * Ts - T * T * T * T));
            P0rho += -rate / rho;
            P0T += -invCp * cpsensT * rate - a.rx.SoV * invRhoCp * (a.rx.h_conv + 4. * a.rx.eps_rad * 5.67e-8 * T * T * T);
            cextra += -invCp * rate;
            rhs0 += rate;
          }
          const double roT = rho * SMG(s.sc, J_INVT, g), nRM = -rho * SMG(s.sc, J_MMW, g);
          sJ[(size_t)g * nsns] = isothermal ? 0. : P0T - roT * P0rho;
          if (g < gcount)
            a.out0[(size_t)(tile0 + g) * ns] = isothermal ? 0. : rhs0;
          SMG(sS, TS_INVRHO, g) = invRho;
          SMG(sS, TS_INVRHOCP, g) = invRhoCp;
          SMG(sS, TS_KY, g) = -rhs0c * invCp + cextra;
          SMG(sS, TS_SA, g) = SA;
          SMG(sS, TS_SB, g) = SB;
          SMG(sS, TS_NRM, g) = nRM;
          SMG(sS, TS_P0RHO, g) = P0rho;
        }
      }
      else
      {
        for (int item = tid; item < ns * G; item += nct - 32)
        { // chem_jac_isobaric rows (:75-98) folded with transform_isobaric_primitive_jacobian (:319-343)
          const int i = item / G, g = item - i * G;
          const double invRho = SMG(s.sc, J_IRHO, g), rho = SMG(s.sc, J_RHO, g);
          const double w = wval(0, i, g), wr = wval(1, i, g), wT = wval(2, i, g);
          const double nmA = wval(3, i, g), nmB = wval(4, i, g);
          const double prho = invRho * (wr - invRho * w); // P[1+i, rho]
          const double nRM = -rho * SMG(s.sc, J_MMW, g), roT = rho / SMG(s.sc, J_T, g);
          SMG(sc2, i, g) = invRho * nmA + nRM * prho;
          SMG(sc3, i, g) = invRho * nmB;
          if (i < nsm1)
          {
            sJ[(size_t)g * nsns + 1 + i] = wT * invRho - roT * prho; // J[1+i, 0]
            if (g < gcount)
            { // right-hand side, chem_rhs_isobaric :19-29 (+ :194-218)
              double v = w * invRho;
              if (open)
                v += (a.rx.y_in[i] - SMG(s.sy, i, g)) * invTau;
              a.out0[(size_t)(tile0 + g) * ns + 1 + i] = v;
            }
          }
        }
      }
    }
    TL4(7)
    nbar_sync(BAR_CONS, nct);
    TL4(8)
    // ---- transform in place, one column per warp pass: lanes (j, g) own the rows j, j+8, ... of state g -----------------
    {
      const int j = lane >> 2, g = lane & 3;
      const double c1 = SMG(sS, TS_INVRHO, g);
      double *const Jg = sJ + (size_t)g * nsns;
      for (int c = 1 + warp; c < ns; c += ncons)
      {
        const int k = c - 1;
        const double uk = su[k];
        // split destinations of this column: add the extra parts in part order
        {
          const int f0 = t_cfxoff[c], f1 = t_cfxoff[c + 1];
          if (f1 > f0)
          {
            for (int f = f0 + j; f < f1; f += 8)
            {
              double *p = Jg + c * ns + t_cfx[3 * f];
              double v = *p;
              const int first = t_cfx[3 * f + 1], np = t_cfx[3 * f + 2];
              for (int q = 0; q < np; ++q)
                v += sWX[(size_t)(nwrow + first + q) * G + g];
              *p = v;
            }
            __syncwarp();
          }
        }
        double acc = 0.;
        for (int r = j; r < ns; r += 8)
        {
          double *p = Jg + c * ns + r;
          const double v = *p;
          const int sp = r ? r - 1 : nsm1; // row 0 temporarily holds the row of the last species
          acc = fma(SMG(s.sh, sp, g), v, acc);
          if (r)
          {
            double o = fma(c1, v, fma(uk, SMG(sc2, sp, g), SMG(sc3, sp, g)));
            if (r == c && open)
              o += -invTau;
            *p = o;
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (j == 0)
        { // temperature row, isobaric_reactor_kernels.cpp:74-98 with the transform :319-343
          double v = 0.;
          if (!isothermal)
          {
            const double sum = acc + uk * SMG(sS, TS_SA, g) + SMG(sS, TS_SB, g);
            const double pY = -sum * SMG(sS, TS_INVRHOCP, g) + SMG(sS, TS_KY, g) * (SMG(s.scp, k, g) - SMG(s.scp, nsm1, g));
            v = pY + SMG(sS, TS_NRM, g) * uk * SMG(sS, TS_P0RHO, g);
          }
          Jg[c * ns] = v;
        }
      }
    }
    TL4(9)
    if (tile + 2 * (int)gridDim.x < ntiles)
      nbar_arrive(BAR_EMPTY0 + b, NT);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    nbar_sync(BAR_CONS, nct);
    TL4(10)
    // ---- output: one bulk copy of the four blocks (contiguous in HBM), plain stores for a ragged last tile --------------
    {
      double *dst = a.out1 + (size_t)tile0 * nsns;
      if (gcount == G && (reinterpret_cast<size_t>(dst) & 15) == 0)
      {
        if (tid == 0)
        {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(sJ)),
                       "r"((unsigned int)(G * nsns * sizeof(double)))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      else
      {
        for (int e = tid; e < gcount * nsns; e += nct)
          dst[e] = sJ[e];
      }
    }
  }
  if (tid == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------------
static int jac4_sm_count()
{
  static int n = 0;
  if (!n)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int NT>
static cudaError_t launch_jac4_nt(const ChemArgs &a, cudaStream_t s)
{
  static bool attr = false;
  if (!attr)
  {
    cudaError_t e = cudaFuncSetAttribute(k_jac4<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess)
      return e;
    attr = true;
  }
  const int ntiles = (a.n + G4 - 1) / G4;
  const int grid = std::max(1, std::min(ntiles, jac4_sm_count()));
  k_jac4<NT><<<grid, NT, (size_t)a.dm.j4_smem, s>>>(a);
  ++g_jac_launches;
  return cudaGetLastError();
}

bool jac4_applicable(const ChemArgs &a)
{
  return a.mode == MODE_REACTOR_JAC && a.dm.j4_threads > 0 && a.in_state != nullptr;
}

cudaError_t launch_jac4(const ChemArgs &a, cudaStream_t s)
{
  switch (a.dm.j4_threads)
  {
  case 512:
    return launch_jac4_nt<512>(a, s);
  case 768:
    return launch_jac4_nt<768>(a, s);
  case 1024:
    return launch_jac4_nt<1024>(a, s);
  default:
    return cudaErrorInvalidConfiguration;
  }
}

#ifdef GB_JAC_TIMELINE
int debug_jac4_timeline(long long *out)
{
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_jac4_timeline, sizeof(long long) * 16 * 32) == cudaSuccess ? 0 : -3;
}
#endif

} // namespace gb
