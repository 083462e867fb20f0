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
__device__ __forceinline__ double lds_f64(unsigned int addr)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds_v2f64(unsigned int addr, double &x, double &y)
{
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts_f64(unsigned int addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(unsigned int addr, double x, double y)
{
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
constexpr int JT_MAXM = 8; // rows per lane of the transform phase held in registers: mechanisms up to 64 species

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
  extern __shared__ __align__(128) double smem[];
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, nsm1 = ns - 1, nsns = ns * ns;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncons = dm.j4_ncons, nprod = NT / 32 - ncons;
  const int nct = ncons * 32, npt = nprod * 32;

  double *const bufs = smem;
  const int bufsz = dm.j4_bufsz;
  double *const sdcp = bufs + 2 * bufsz;
  double *const su = sdcp + ns * G, *const snm = su + ns, *const sim = snm + ns;
  double *const sWX = sim + ns + ((4 - (3 * ns) % 4) & 3); // 32-byte aligned rows from here on
  double *const sc2 = sWX + (size_t)dm.j4_nwx * G, *const sc3 = sc2 + ns * G, *const sS = sc3 + ns * G;
  double *const sR = sS + 16 * G;
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
      TL4(11)
      if (kt >= 2)
        nbar_sync(BAR_EMPTY0 + b, NT);
      TL4(12)
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
      TL4(13)
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
      TL4(14)
      // ---- third-body / falloff factors: 8 reactions x 4 states per task -----------------------------------------------
      for (int fg = pw; fg < dm.j4_nfg; fg += nprod)
      {
        const int off = t_fgroups[fg * 8 + (lane >> 2)];
        if (off >= 0)
          falloff_task<G>(dm, dm.jp_prm + off, lane & 3, s);
      }
      TL4(15)
      nbar_arrive(BAR_FULL0 + b, NT);
    }
    return;
  }

  // ================================================== consumers ====================================================
  const int *t_wg = stab + dm.j4_t_wg, *t_groups = stab + dm.j4_t_groups, *t_wr = stab + dm.j4_t_wr;
  const int *t_rounds = stab + dm.j4_t_rounds, *t_wfix = stab + dm.j4_t_wfix, *t_cfxoff = stab + dm.j4_t_cfxoff;
  const int *t_cfx = stab + dm.j4_t_cfx;
  const int nwrow = 5 * ns; // row-scalar rows of sWX; extra parts follow
  const unsigned int sJ_u32 = smem_u32(sJ), sR_u32 = smem_u32(sR), sWX_u32 = smem_u32(sWX);
  const unsigned int jstride = 8u * (unsigned int)nsns; // bytes between the blocks of two states
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
    // the previous tile's bulk copy must have finished reading the Jacobian tile before the gather writes into it
    if (tid == 0)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    nbar_sync(BAR_CONS, nct);
    TL4(4)
    // ---- gather: lane per destination part, four accumulators, results straight into the Jacobian tile ---------
    {
      // structural zeros of R (entries without a destination) are written here: nobody else touches them
      for (int z = tid; z < dm.j4_nzero; z += nct)
      {
        const unsigned int addr = sJ_u32 + 8u * (unsigned int)__ldg(dm.j4_zlist + z);
        sts_f64(addr, 0.);
        sts_f64(addr + jstride, 0.);
        sts_f64(addr + 2 * jstride, 0.);
        sts_f64(addr + 3 * jstride, 0.);
      }
      const int r0 = t_wr[warp], r1 = t_wr[warp + 1];
      const unsigned int rot16 = (lane & 1) * 16u;
      // the rounds of a warp are contiguous in the item stream: items are fetched three step pairs ahead (they come
      // from L2: what is left of L1 next to 225 KB of shared memory does not hold the plan)
      const uint2 *__restrict__ it =
          reinterpret_cast<const uint2 *>(dm.j4_items) + (r1 > r0 ? t_rounds[2 * r0] / 2 : 0) + lane;
      uint2 cur = __ldg(it), n1 = __ldg(it + 32), n2 = __ldg(it + 64);
      it += 96;
      unsigned int rd = r1 > r0 ? __ldg(dm.j4_rdest + (size_t)r0 * 32 + lane) : 0xffffu;
      for (int r = r0; r < r1; ++r)
      {
        const int L = t_rounds[2 * r + 1];
        const unsigned int code = rd & 0xffffu;
        const double nm = snm[rd >> 16];
        if (r + 1 < r1)
          rd = __ldg(dm.j4_rdest + (size_t)(r + 1) * 32 + lane);
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
            // |nu| > 1 arrives as repeated items. Item = byte offset of the record row | sign << 31
            const double c = __hiloint2double(nm_hi ^ (int)(u & 0x80000000u), nm_lo);
            const unsigned int addr = sR_u32 + (u & 0x7fffffffu) + rot16;
            double x0, x1, y0, y1;
            lds_v2f64(addr, x0, x1);
            lds_v2f64(addr ^ 16u, y0, y1);
            acc0 = fma(c, x0, acc0);
            acc1 = fma(c, x1, acc1);
            acc2 = fma(c, y0, acc2);
            acc3 = fma(c, y1, acc3);
          }
          cur = n1;
          n1 = n2;
          n2 = nn;
        }
        if (code < (unsigned int)nsns)
        {
          const unsigned int addr = sJ_u32 + 8u * code + (rot16 >> 3) * jstride; // state 2*rot
          sts_f64(addr, acc0);
          sts_f64(addr + jstride, acc1);
          const unsigned int addr2 = sJ_u32 + 8u * code + (2u - (rot16 >> 3)) * jstride; // state 2*(rot^1)
          sts_f64(addr2, acc2);
          sts_f64(addr2 + jstride, acc3);
        }
        else if (code != 0xffffu)
        {
          const unsigned int addr = sWX_u32 + 32u * (code - (unsigned int)nsns) + rot16;
          sts_v2f64(addr, acc0, acc1);
          sts_v2f64(addr ^ 16u, acc2, acc3);
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
        double v = sWX[(q * ns + i) * G + g];
        const int fx = t_wfix[q * ns + i];
        if (fx)
          for (int p = 0; p < (fx & 255); ++p)
            v += sWX[(nwrow + (fx >> 8) + p) * G + g];
        return v;
      };
      // (q, g) pairs dealt to the warps, lanes along the species: sum_i h_i * {W, Wrho, WT, A, B}_i and sum_i cp_i * W_i,
      // the inner products of isobaric_reactor_kernels.cpp:74-92 (pairwise instead of sequential summation)
      for (int pr = warp; pr < 6 * G; pr += ncons)
      {
        const int q = pr >> 2, g = pr & 3;
        const double *wsrc = q == 5 ? s.scp : s.sh;
        const int qq = q == 5 ? 0 : q;
        double acc = 0.;
        for (int i = lane; i < ns; i += 32)
          acc = fma(wsrc[i * G + g], wval(qq, i, g), acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (lane == 0)
          sS[(8 + q) * G + g] = acc;
      }
      for (int item = tid; item < ns * G; item += nct)
      { // chem_jac_isobaric rows (:75-98) folded with transform_isobaric_primitive_jacobian (:319-343)
        const int i = item >> 2, g = item & 3;
        const double invRho = s.sc[J_IRHO * G + g], rho = s.sc[J_RHO * G + g];
        const double w = wval(0, i, g), wr = wval(1, i, g), wT = wval(2, i, g);
        const double nmA = wval(3, i, g), nmB = wval(4, i, g);
        const double prho = invRho * (wr - invRho * w); // P[1+i, rho]
        const double nRM = -rho * s.sc[J_MMW * G + g], roT = rho / s.sc[J_T * G + g];
        sc2[item] = invRho * nmA + nRM * prho;
        sc3[item] = invRho * nmB;
        if (i < nsm1)
        {
          sJ[g * nsns + 1 + i] = wT * invRho - roT * prho; // J[1+i, 0]
          if (g < gcount)
          { // right-hand side, chem_rhs_isobaric :19-29 (+ :194-218)
            double v = w * invRho;
            if (open)
              v += (a.rx.y_in[i] - s.sy[item]) * invTau;
            a.out0[(size_t)(tile0 + g) * ns + 1 + i] = v;
          }
        }
      }
    }
    TL4(7)
    nbar_sync(BAR_CONS, nct);
    TL4(8)
    // ---- transform in place, one column per warp pass: lanes (j, g) own the rows j, j+8, ... of state g. Row 0 of a
    // column temporarily holds the row of the last species (it only enters the temperature row) and receives the
    // column's inner product sum_i h_i R[i][k]; the temperature row is finished after the next barrier. -----------
    {
      if (warp == ncons - 1 && lane < G)
      { // chem_jac_isobaric :58-98, mass_jac_isobaric :100-140, heat_jac_isobaric :142-168, transform :319-343
        const int g = lane;
        const double SW = sS[8 * G + g], SWr = sS[9 * G + g], SWT = sS[10 * G + g], SA = sS[11 * G + g];
        const double SB = sS[12 * G + g], wcp = sS[13 * G + g];
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
        sJ[g * nsns] = isothermal ? 0. : P0T - roT * P0rho;
        if (g < gcount)
          a.out0[(size_t)(tile0 + g) * ns] = isothermal ? 0. : rhs0;
        SMG(sS, TS_INVRHOCP, g) = invRhoCp;
        SMG(sS, TS_KY, g) = -rhs0c * invCp + cextra;
        SMG(sS, TS_SA, g) = SA;
        SMG(sS, TS_SB, g) = SB;
        SMG(sS, TS_NRM, g) = nRM;
        SMG(sS, TS_P0RHO, g) = P0rho;
      }
      const int j = lane >> 2, g = lane & 3;
      const double c1 = s.sc[J_IRHO * G + g];
      // this lane's rows: r = j + 8 m; their h, c2, c3 stay in registers for all columns of the tile
      double hr[JT_MAXM], c2r[JT_MAXM], c3r[JT_MAXM];
      const int mall = ns >= 8 ? (ns - 8) / 8 + 1 : 0; // iterations in which every lane has a row
      const bool tail = j + 8 * mall < ns;
#pragma unroll
      for (int m = 0; m < JT_MAXM; ++m)
      {
        const int r = j + 8 * m;
        hr[m] = c2r[m] = c3r[m] = 0.;
        if (r < ns)
        {
          const int sp = r ? r - 1 : nsm1;
          hr[m] = s.sh[sp * G + g], c2r[m] = sc2[sp * G + g], c3r[m] = sc3[sp * G + g];
        }
      }
      const unsigned int lane_u32 = sJ_u32 + 8u * (unsigned int)(g * nsns + j);
      for (int c = 1 + warp; c < ns; c += ncons)
      {
        const double uk = su[c - 1];
        // split destinations of this column: add the extra parts in part order
        {
          const int f0 = t_cfxoff[c], f1 = t_cfxoff[c + 1];
          if (f1 > f0)
          {
            for (int f = f0 + j; f < f1; f += 8)
            {
              double *p = sJ + g * nsns + c * ns + t_cfx[3 * f];
              double v = *p;
              const int first = t_cfx[3 * f + 1], np = t_cfx[3 * f + 2];
              for (int q = 0; q < np; ++q)
                v += sWX[(nwrow + first + q) * G + g];
              *p = v;
            }
            __syncwarp();
          }
        }
        const unsigned int col = lane_u32 + 8u * (unsigned int)(c * ns);
        double acc = 0.;
#pragma unroll
        for (int m = 0; m < JT_MAXM; ++m)
        {
          if (m < mall || (m == mall && tail))
          {
            const double v = lds_f64(col + 64u * m);
            acc = fma(hr[m], v, acc);
            if (m > 0 || j > 0)
            {
              double o = fma(c1, v, fma(uk, c2r[m], c3r[m]));
              if (open && j + 8 * m == c)
                o += -invTau;
              sts_f64(col + 64u * m, o);
            }
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (j == 0)
          sts_f64(col, acc);
      }
    }
    TL4(9)
    nbar_sync(BAR_CONS, nct);
    // ---- temperature row, isobaric_reactor_kernels.cpp:74-98 with the transform :319-343 ------------------------------------
    for (int item = tid; item < (ns - 1) * G; item += nct)
    {
      const int k = item >> 2, g = item & 3;
      double *p = sJ + g * nsns + (k + 1) * ns;
      double v = 0.;
      if (!isothermal)
      {
        const double uk = su[k];
        const double sum = *p + uk * SMG(sS, TS_SA, g) + SMG(sS, TS_SB, g);
        const double pY = -sum * SMG(sS, TS_INVRHOCP, g) + SMG(sS, TS_KY, g) * (s.scp[k * G + g] - s.scp[nsm1 * G + g]);
        v = pY + SMG(sS, TS_NRM, g) * uk * SMG(sS, TS_P0RHO, g);
      }
      *p = v;
    }
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
  case 640:
    return launch_jac4_nt<640>(a, s);
  case 768:
    return launch_jac4_nt<768>(a, s);
  case 896:
    return launch_jac4_nt<896>(a, s);
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
