// gb_jac.cu -- k_jac: exact rate sensitivities, isobaric-reactor Jacobian and flamelet Jacobian blocks, FP64, sm_100a.
//
// Replaces prod_rates_sens_exact (rates_sensitivities_exact.cpp:33-1028), chem_jac_isobaric + mass/heat_jac_isobaric
// + transform_isobaric_primitive_jacobian (isobaric_reactor_kernels.cpp:58-168, 221-343) and the per-point part of
// flamelet_jacobian (flamelet_kernels.cpp:1254-1408) for batches of states.
//
// A CTA owns a tile of G states whose working set lives in shared memory with the state index fastest ([row][g]).
// Phases, separated by __syncthreads() (the static schedule is built on the host, gb_plan.cu):
//   load     : coalesced read of the G state vectors
//   thermo   : thread per (species, state) -> cp_i, h_i, dcp_i/dT, Gibbs, dB_i/dT; two groups of G threads do the
//              order-sensitive sums (Y_ns = 1 - sum, mixture weight) sequentially as the reference does
//   conc     : thread per (species, state) -> concentrations; per-state scalars
//   react    : warp per reaction group, lane = state + G * reaction-in-group -> reaction records in shared memory
//   gather   : lane per destination part, G accumulators in registers, items in ascending reaction order; the sums
//              stay in registers until every warp is done reading records, then overwrite the record region
//   fix      : split destinations are recombined in part order
//   rows/cols: per-row constants of the output transform, column sums for the temperature row (species order)
//   T-row    : final temperature-row values
//   output   : J = c1_row * R + u_col * c2_row + c3_row streamed to HBM, lanes along the contiguous (column-major)
//              dimension, G coalesced stores per thread and entry
// HBM traffic is the algorithmic minimum: ns doubles in, ns + ns^2 doubles out per state.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "gb_react.cuh"

namespace gb
{

extern std::atomic<long> g_jac_launches;
std::atomic<long> g_jac_launches{0};


// Debug build only (-DGB_JAC_TIMELINE): every warp of CTA 0 records clock64() when it ARRIVES at each barrier of its
// second tile; read back through gb_debug_jac_timeline (gb_api.cu), tools/timeline.py prints the table.
#ifdef GB_JAC_TIMELINE
__device__ long long g_jac_timeline[20 * 32];
#define TL_MARK(k)                                                                  \
  if (blockIdx.x == 0 && tile == blockIdx.x + gridDim.x && lane == 0)               \
    g_jac_timeline[(k)*32 + warp] = clock64();
#else
#define TL_MARK(k)
#endif

// ------------------------------------------------------------------------------------------------------------------
// Rows of G doubles are read and written 16 bytes at a time. Lanes of a warp usually address different rows at the
// same moment, and a row is G*8 bytes, so without care all lanes hit the same few banks. Every lane therefore visits
// the G/2 chunks of a row in an order rotated by its lane number; its registers hold the states in that rotated
// order ("slots"): slot s <-> state 2*(((s>>1) + rot) & (G/2-1)) + (s&1).
// ------------------------------------------------------------------------------------------------------------------
template <int G>
struct Rows
{
  static constexpr int NCH = G >= 2 ? G / 2 : 1;
  __device__ __forceinline__ static int chunk(int j, int rot) { return (j + rot) & (NCH - 1); }
  __device__ __forceinline__ static int state_of(int slot, int rot) { return G == 1 ? 0 : 2 * chunk(slot >> 1, rot) + (slot & 1); }
  __device__ __forceinline__ static void load(const double *p, int rot, double (&v)[G])
  {
    if (G == 1)
      v[0] = p[0];
    else
    {
#pragma unroll
      for (int j = 0; j < NCH; ++j)
      {
        const double2 x = *reinterpret_cast<const double2 *>(p + 2 * chunk(j, rot));
        v[2 * j] = x.x;
        v[2 * j + (G >= 2 ? 1 : 0)] = x.y;
      }
    }
  }
  // The gathered sums are stored with the chunks of row R rotated by R ("swizzled"): chunk c sits at position
  // (c + R) & (NCH-1). Lanes that read different rows in canonical register order (slot = state, needed for
  // coalesced global stores) then spread over all banks as well.
  __device__ __forceinline__ static int swz(int row, int g)
  {
    return G == 1 ? row : row * G + 2 * chunk(g >> 1, row) + (g & 1);
  }
  // slot j <-> state chunk (j + rot): rot = 0 gives canonical order
  __device__ __forceinline__ static void load_swz(const double *base, int row, int rot, double (&v)[G])
  {
    if (G == 1)
      v[0] = base[row];
    else
    {
#pragma unroll
      for (int j = 0; j < NCH; ++j)
      {
        const double2 x = *reinterpret_cast<const double2 *>(base + (size_t)row * G + 2 * chunk(j, rot + row));
        v[2 * j] = x.x;
        v[2 * j + (G >= 2 ? 1 : 0)] = x.y;
      }
    }
  }
  __device__ __forceinline__ static void store_swz(double *base, int row, int rot, const double (&v)[G])
  {
    if (G == 1)
      base[row] = v[0];
    else
    {
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        *reinterpret_cast<double2 *>(base + (size_t)row * G + 2 * chunk(j, rot + row)) =
            make_double2(v[2 * j], v[2 * j + (G >= 2 ? 1 : 0)]);
    }
  }
  __device__ __forceinline__ static void store(double *p, int rot, const double (&v)[G])
  {
    if (G == 1)
      p[0] = v[0];
    else
    {
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        *reinterpret_cast<double2 *>(p + 2 * chunk(j, rot)) = make_double2(v[2 * j], v[2 * j + (G >= 2 ? 1 : 0)]);
    }
  }
};

// one gather step: acc[slot] += (nu * -M_row) * record[row][state(slot)], the factor and the two roundings of the
// reference's `wsens += factor * dq` (rates_sensitivities_exact.cpp:1014-1026)
template <int G>
__device__ __forceinline__ void gather_step(unsigned int u, const double *sR, int rot, double nm, double (&acc)[G])
{
  const double coef = (double)(((int)(u << 8)) >> 24) * nm;
  double v[G];
  Rows<G>::load(sR + (size_t)(u & 0xffff) * G, rot, v);
#pragma unroll
  for (int g = 0; g < G; ++g)
    acc[g] = acc[g] + coef * v[g];
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int JAC_SLOTS = 8; // start-time slots of the persistent CTAs (see `stagger`)

template <int G>
__global__ void __launch_bounds__(512, 1) k_jac(const ChemArgs a)
{
  extern __shared__ __align__(16) double smem[];
  constexpr int LPR = 32 / G, RMAX = 32 / G;
  const DeviceMech &dm = a.dm;
  const int ns = dm.ns, nsm1 = ns - 1;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int rot = lane & (Rows<G>::NCH - 1);
  const int region = max(dm.jp_rec_rows, dm.jp_rows) + 2;
  JacSmem s;
  s.sc = smem;                         // [JP_NSC][G] per-state scalars
  s.sy = s.sc + JP_NSC * G;            // [ns][G] mass fractions
  s.sC = s.sy + ns * G;                // concentrations
  s.sg = s.sC + ns * G;                // Gibbs              -> later: c1 of the output transform
  s.sdb = s.sg + ns * G;               // dB/dT              -> later: c2
  s.sh = s.sdb + ns * G;               // enthalpies
  s.scp = s.sh + ns * G;               // species cp
  s.sdcp = s.scp + ns * G;             // species dcp/dT     -> later: c3
  s.su = s.sdcp + ns * G;              // [ns] u_k = 1/M_k - 1/M_ns
  s.snm = s.su + ns;                   // [ns] -M_i
  s.sim = s.snm + ns;                  // [ns] 1/M_i
  s.sR = s.sim + ns + (ns & 1);        // [region][G] reaction records, then gathered sums (16-byte aligned)
  s.sTH = s.sR + (size_t)region * G;   // [ncsp][G] column-sum parts
  int *stab = (int *)(s.sTH + (size_t)dm.jp_ncsp * G); // plan tables (gb_mech.h)
  s.semap = (unsigned short *)(stab + dm.jp_tab_words); // [(ns+1)*(ns-1)]
  const int zrow = dm.jp_zrow;
  const int *t_groups = stab + dm.jp_t_groups, *t_rounds = stab + dm.jp_t_rounds, *t_fix = stab + dm.jp_t_fix;
  const int *t_csparts = stab + dm.jp_t_csparts, *t_cspfirst = stab + dm.jp_t_cspfirst;
  const unsigned int *t_csitems = (const unsigned int *)(stab + dm.jp_t_csitems);
  const unsigned short *t_rdest = (const unsigned short *)(stab + dm.jp_t_rdest);
  const unsigned short *t_rspec = (const unsigned short *)(stab + dm.jp_t_rspec);
  const unsigned short *rowsrc = (const unsigned short *)(stab + dm.jp_t_rowsrc);

  // per-CTA constants
  for (int i = tid; i < ns; i += nt)
  {
    s.su[i] = dm.invmw[i] - dm.invmw[nsm1];
    s.snm[i] = -dm.netmw[i];
    s.sim[i] = dm.invmw[i];
  }
  for (int e = tid; e < dm.jp_tab_words; e += nt)
    stab[e] = dm.jp_tab[e];
  for (int e = tid; e < (ns + 1) * nsm1; e += nt)
    s.semap[e] = dm.jp_emap[e];
  if (tid < 2 * G)
    s.sR[(size_t)zrow * G + tid] = 0.;
#define SJ(row, g) s.sR[Rows<G>::swz((row), (g))] /* gathered sums: swizzled rows */

  const bool state_mode = a.in_state != nullptr;
  const bool reactor = a.mode == MODE_REACTOR_JAC;
  const bool flamelet = a.mode == MODE_FLAMELET_JAC;
  const bool isothermal = reactor && a.rx.heat_option == 1;
  const bool open = reactor && a.rx.open != 0;
  const double invTau = open ? 1. / a.rx.tau : 0.;
  const FlameletDev &fl = a.fl;

  const int ntiles = (a.n + G - 1) / G;
  // The parameter records of a reaction group are pulled into L1 one group ahead: the G lanes that share a reaction
  // touch 512 bytes of its record between them, so the (dependent) parameter reads of the generic path hit L1.
  auto prefetch_group = [&](int gi, int gend) {
    if (gi < gend)
    {
      const int off = t_groups[gi * (1 + LPR) + 1 + lane / G];
      if (off >= 0)
      {
        const char *p = reinterpret_cast<const char *>(dm.jp_prm + off) + (lane % G) * (512 / G);
#pragma unroll
        for (int k = 0; k < 512 / G; k += 32)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(p + k));
      }
    }
  };
  auto fetch_state = [&](int tile) { // this thread's first element of a tile's state block
    double v = 0.;
    if (state_mode && tile < ntiles && tid < G * ns)
    {
      const int g = tid / ns, j = tid - g * ns, gc = min(G, a.n - tile * G);
      v = a.in_state[(size_t)(tile * G + (g < gc ? g : 0)) * ns + j];
    }
    return v;
  };
  double pre = fetch_state(blockIdx.x);
  // The last warp is idle during the temperature-row phase: its first G lanes use that time to fetch the NEXT tile's
  // temperatures and to form log T and 1/T, which otherwise sit (a ~1.5 k cycle dependent chain on G threads) between
  // two block barriers of the load phase.
  const bool tkeeper = state_mode && warp == (nt >> 5) - 1 && lane < G;
  double nxtT = 0.;
  auto fetch_T = [&](int tile) {
    if (tkeeper && tile < ntiles)
    {
      const int gc = min(G, a.n - tile * G);
      nxtT = a.in_state[(size_t)(tile * G + (lane < gc ? lane : 0)) * ns];
    }
  };
  auto finish_T = [&]() { // (parked in two spare scalar rows so that nothing but T itself stays in registers)
    if (tkeeper)
    {
      SMG(s.sc, J_NLOGT, lane) = log(nxtT);
      SMG(s.sc, J_NINVT, lane) = 1. / nxtT;
    }
  };
  fetch_T(blockIdx.x);
  finish_T();
  // De-phasing: all CTAs of the persistent grid start together and take the same time per tile, so without a start
  // offset every SM reaches its store-bound output phase at the same moment and the chip-wide write path, not the
  // SM's own, sets the length of that phase. CTA b waits (b mod JAC_SLOTS) * stagger cycles once.
  if (a.stagger > 0)
  {
    if (tid == 0)
    {
      const long long t0 = clock64(), d = (long long)(blockIdx.x % JAC_SLOTS) * a.stagger;
      while (clock64() - t0 < d)
        __nanosleep(256);
    }
    __syncthreads();
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const int tile0 = tile * G;
    const int gcount = min(G, a.n - tile0);
    TL_MARK(0)
    __syncthreads();
    // ---- load (states past the end of the batch replicate the tile's first state; they are never written). The
    // first G*ns values of the tile were fetched into registers during the previous tile's output phase ------------
    if (state_mode)
    {
      for (int item = tid, k = 0; item < G * ns; item += nt, ++k)
      {
        const int g = item / ns, j = item - g * ns;
        const double v = (k == 0) ? pre : a.in_state[(size_t)(tile0 + (g < gcount ? g : 0)) * ns + j];
        if (j != 0)
          SMG(s.sy, j - 1, g) = v;
      }
      if (tkeeper)
      {
        SMG(s.sc, J_T, lane) = nxtT;
        SMG(s.sc, J_LOGT, lane) = SMG(s.sc, J_NLOGT, lane);
        SMG(s.sc, J_INVT, lane) = SMG(s.sc, J_NINVT, lane);
      }
    }
    else
    {
      for (int item = tid; item < G * ns; item += nt)
      {
        const int g = item / ns, j = item - g * ns;
        SMG(s.sy, j, g) = a.in_y[(size_t)(tile0 + (g < gcount ? g : 0)) * ns + j];
      }
      if (tid < G)
      {
        const double T = a.in_T[tile0 + (tid < gcount ? tid : 0)];
        SMG(s.sc, J_T, tid) = T;
        SMG(s.sc, J_LOGT, tid) = log(T);
        SMG(s.sc, J_INVT, tid) = 1. / T;
        SMG(s.sc, J_RHO, tid) = a.in_rho[tile0 + (tid < gcount ? tid : 0)];
      }
    }
    TL_MARK(1)
    __syncthreads();
    // ---- thermo, overlapped with the two order-sensitive chains and the per-state scalars (warp 0) ---------------------
    prefetch_group(stab[dm.jp_t_wg + warp], stab[dm.jp_t_wg + warp + 1]);
    if (warp == 0)
    {
      // lanes [0, G): Y_ns = 1 - sum_j Y_j (extract_y, combustion_kernels.h:505-515); lanes [G, 2G): sum_i Y_i/M_i over
      // all but the last species (mixture_molecular_weight, :381-387). One uniform loop d <- d + c_j * Y_j serves
      // both chains (c_j = -1 makes the update the exact subtraction of the reference).
      double d = 0.;
      if (lane < 2 * G)
      {
        const bool first = lane < G;
        const int g = first ? lane : lane - G;
        if (first && !state_mode)
          d = SMG(s.sy, nsm1, g);
        else
        {
          d = first ? 1. : 0.;
#pragma unroll 4
          for (int j = 0; j < nsm1; ++j)
            d = d + (first ? -1. : s.sim[j]) * SMG(s.sy, j, g);
          if (first)
            SMG(s.sy, nsm1, g) = d;
        }
      }
      const double dpart = __shfl_down_sync(0xffffffffu, d, G);
      if (lane < G)
      {
        const int g = lane;
        const double mmw = 1. / (dpart + s.sim[nsm1] * d), T = SMG(s.sc, J_T, g), invT = SMG(s.sc, J_INVT, g);
        // ideal_gas_density (:526-530)
        const double rho = state_mode ? a.p * mmw / (T * dm.Ru) : SMG(s.sc, J_RHO, g);
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
      for (int item = tid - 32; item < ns * G; item += nt - 32)
      {
        const int i = item / G, g = item - i * G;
        const SpeciesThermo t = species_thermo<true>(dm, i, SMG(s.sc, J_T, g), SMG(s.sc, J_LOGT, g), SMG(s.sc, J_INVT, g));
        SMG(s.sg, i, g) = t.g;
        SMG(s.sdb, i, g) = t.dB;
        SMG(s.sh, i, g) = t.h;
        SMG(s.scp, i, g) = t.cp;
        SMG(s.sdcp, i, g) = t.dcp;
      }
    }
    TL_MARK(2)
    __syncthreads();
    // ---- concentrations ---------------------------------------------------------------------------------------------------------------
    for (int item = tid; item < ns * G; item += nt)
    {
      const int i = item / G, g = item - i * G;
      SMG(s.sC, i, g) = SMG(s.sy, i, g) * SMG(s.sc, J_RHO, g) * s.sim[i];
    }
    TL_MARK(3)
    __syncthreads();
    // ---- reaction phase (the last warp first forms cp, dcp/dT and the open-reactor terms in species order) -------------
    if (tid >= nt - 32 && lane < G)
    { // thermodynamics_kernels.cpp:45-131, 183-260
      const int g = lane;
      double cp = 0., dcp = 0.;
      for (int i = 0; i < ns; ++i)
      {
        cp += SMG(s.sy, i, g) * SMG(s.scp, i, g);
        if (dm.cptype[i] == CP_CONST)
          dcp = 0.; // sic, thermodynamics_kernels.cpp:202
        else
          dcp += SMG(s.sy, i, g) * SMG(s.sdcp, i, g);
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
    {
      const int g = lane % G, sub = lane / G;
      const int g0 = stab[dm.jp_t_wg + warp], g1 = stab[dm.jp_t_wg + warp + 1];
      for (int gi = g0; gi < g1; ++gi)
      {
        prefetch_group(gi + 1, g1);
        const int *grp = t_groups + gi * (1 + LPR);
        const int kind = grp[0], off = grp[1 + sub];
        if (off < 0)
          continue;
        const unsigned long long *P = dm.jp_prm + off;
#ifdef GB_JAC_TIMELINE
        const long long tl0 = clock64();
#endif
        if (kind == 0)
          react_fast<G>(dm, P, g, s);
        else if (kind == 1)
          react_struct<G>(dm, P, g, s);
        else if (((int)(unsigned int)P[0]) & F_HAS_ORDERS)
          react_orders<G>(ns, dm.n_sp, dm.sp_idx, dm.sp_order, dm.sp_slot, dm.invmw, (int)(((unsigned int)P[0]) >> 14), P, g,
                          s.sc, s.sC, s.sy, s.sR);
        else
          react_generic<G>(dm, P, g, s);
#ifdef GB_JAC_TIMELINE
        if (blockIdx.x == 0 && tile == blockIdx.x + gridDim.x && lane == 0)
        {
          // bucket: 0 fast, 1 structured simple, 2 third body, 3 Lindemann, 4 Troe, 5 generic (rows 11.., counts 17..)
          const int bucket = kind == 0 ? 0 : (kind == 2 ? 5 : f_type((int)(unsigned int)P[0]));
          atomicAdd((unsigned long long *)&g_jac_timeline[(11 + bucket) * 32], (unsigned long long)(clock64() - tl0));
          atomicAdd((unsigned long long *)&g_jac_timeline[(11 + bucket) * 32 + 1], 1ull);
          atomicAdd((unsigned long long *)&g_jac_timeline[17 * 32 + warp], (unsigned long long)(clock64() - tl0));
          atomicAdd((unsigned long long *)&g_jac_timeline[18 * 32 + warp], 1ull);
        }
#endif
      }
    }
    TL_MARK(4)
    __syncthreads();
    // ---- gather: every lane sums its parts in registers --------------------------------------------------------------------------
    double hold[RMAX][G];
    const int r0 = stab[dm.jp_t_wr + warp], nround = stab[dm.jp_t_wr + warp + 1] - r0;
    {
      // The warp's rounds are contiguous in the item stream, so the item words are prefetched two blocks (four steps)
      // ahead across round boundaries: they come from L2 (the plan is larger than what is left of L1).
      const unsigned int *__restrict__ it = dm.jp_items + (nround > 0 ? t_rounds[2 * r0] : 0) + lane;
      unsigned int u0 = __ldg(it), u1 = __ldg(it + 32), n0 = __ldg(it + 64), n1 = __ldg(it + 96);
      it += 128;
#pragma unroll
      for (int j = 0; j < RMAX; ++j)
      {
#pragma unroll
        for (int g = 0; g < G; ++g)
          hold[j][g] = 0.;
        if (j < nround)
        {
          const int L = t_rounds[2 * (r0 + j) + 1];
          const double nm = s.snm[t_rspec[(r0 + j) * 32 + lane]];
          for (int k = 0; k < L; k += JP_BLK)
          {
            const unsigned int m0 = __ldg(it), m1 = __ldg(it + 32);
            it += 32 * JP_BLK;
            gather_step<G>(u0, s.sR, rot, nm, hold[j]);
            gather_step<G>(u1, s.sR, rot, nm, hold[j]);
            u0 = n0, u1 = n1;
            n0 = m0, n1 = m1;
          }
        }
      }
    }
    TL_MARK(5)
    __syncthreads();
    // ---- the sums overwrite the record region -----------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (j < nround)
        Rows<G>::store_swz(s.sR, t_rdest[(r0 + j) * 32 + lane], rot, hold[j]);
    TL_MARK(6)
    __syncthreads();
    // ---- recombine split destinations in part order -------------------------------------------------------------------------------
    if (dm.jp_nfix > 0)
    {
#ifdef GB_JAC_FIX_VEC2
      // (measured, not adopted: 6.74-6.77 against 6.61 ms per 262,144 states, paired runs)
      if (G >= 2)
      { // two states (one 16-byte chunk of the swizzled rows) per thread: one pass over the block for GRI-3.0
        constexpr int H = G >= 2 ? G / 2 : 1;
        for (int item = tid; item < dm.jp_nfix * H; item += nt)
        {
          const int fi = item / H, j = item - fi * H;
          const int dst = t_fix[3 * fi], first = t_fix[3 * fi + 1], np = t_fix[3 * fi + 2];
          double2 *pd = reinterpret_cast<double2 *>(s.sR + (size_t)dst * G + 2 * Rows<G>::chunk(j, dst));
          double2 v = *pd;
          for (int p = 0; p < np; ++p)
          {
            const int row = first + p;
            const double2 x = *reinterpret_cast<const double2 *>(s.sR + (size_t)row * G + 2 * Rows<G>::chunk(j, row));
            v.x += x.x;
            v.y += x.y;
          }
          *pd = v;
        }
      }
      else
#endif
      for (int item = tid; item < dm.jp_nfix * G; item += nt)
      {
        const int fi = item / G, g = item - fi * G;
        const int dst = t_fix[3 * fi], first = t_fix[3 * fi + 1], np = t_fix[3 * fi + 2];
        double v = SJ(dst, g);
        for (int p = 0; p < np; ++p)
          v += SJ(first + p, g);
        SJ(dst, g) = v;
      }
      TL_MARK(7)
      __syncthreads();
    }

    if (a.mode == MODE_SENS)
    { // raw (ns+1)x(ns+1) column-major sensitivities, rates_sensitivities_exact.cpp:68, 1011-1025
      const int nsp1 = ns + 1;
      for (int e = tid; e < nsp1 * nsp1; e += nt)
      {
        const int col = e / nsp1, row = e - col * nsp1;
        for (int g = 0; g < gcount; ++g)
        {
          double v = 0.;
          if (row < ns)
          {
            if (col == 0)
              v = SJ(rowsrc[ns + row], g);
            else if (col == 1)
              v = SJ(rowsrc[2 * ns + row], g);
            else if (col - 2 < nsm1)
            {
              const int k = col - 2;
              const double rv = SJ(s.semap[k * (ns + 1) + row + 1], g);
              v = rv + (SJ(rowsrc[3 * ns + row], g) * s.su[k] + SJ(rowsrc[4 * ns + row], g));
            }
          }
          a.out1[(size_t)(tile0 + g) * nsp1 * nsp1 + e] = v;
        }
      }
      continue;
    }

    // ---- column-sum parts (lane per part, G accumulators) and row constants (thread per (species, state)) --------------
    fetch_T(tile + gridDim.x);
#ifdef GB_JAC_EARLY_PREFETCH
    // (measured, not adopted: requesting the next tile's state here, two phases before the store-bound output phase,
    // instead of at its start: 6.80 against 6.73 ms per 262,144 states, paired runs)
    pre = fetch_state(tile + gridDim.x);
#endif
    {
      const int ncsp = dm.jp_ncsp, ncs = dm.jp_ncs;
      [[maybe_unused]] const int nw = nt >> 5;
      // parts first (lane per part), then the row constants on all threads. Measured alternatives (262,144 GRI-3.0
      // states, whole kernel): parts dealt round-robin to ALL warps (-DGB_JAC_CS_SPREAD: a part is a dependent chain of
      // shared-memory reads, so sixteen warps with fourteen parts each should beat seven full warps) 6.82 ms against
      // 6.78 ms; the loop over a part's items software-pipelined by hand (-DGB_JAC_CS_PIPE) spills (128 registers)
      // and costs 18 %.
#ifdef GB_JAC_CS_SPREAD
      for (int job = lane * nw + warp; job < ncsp; job += 32 * nw)
#else
      for (int job = tid; job < ncsp; job += nt)
#endif
      {
        const int d = t_csparts[3 * job], p0 = t_csparts[3 * job + 1], p1 = t_csparts[3 * job + 2];
        double acc[G];
#pragma unroll
        for (int g = 0; g < G; ++g)
          acc[g] = 0.;
        // sum over species (ascending) of h_i (cp_i for the last destination) times the row value: the inner
        // products of isobaric_reactor_kernels.cpp:74-92
        const double *wsrc = (d == ncs - 1) ? s.scp : s.sh;
        auto fetch = [&](int p, double (&w)[G], double (&v)[G]) {
          const unsigned int u = t_csitems[p];
          Rows<G>::load(wsrc + (size_t)(u >> 16) * G, rot, w);
          Rows<G>::load_swz(s.sR, (int)(u & 0xffff), rot, v);
        };
        auto mac = [&](const double (&w)[G], const double (&v)[G]) {
#pragma unroll
          for (int g = 0; g < G; ++g)
            acc[g] += w[g] * v[g];
        };
#ifndef GB_JAC_CS_PIPE
        for (int p = p0; p < p1; ++p)
        {
          double w[G], v[G];
          fetch(p, w, v);
          mac(w, v);
        }
#else
        if (p0 < p1)
        {
          double wa[G], va[G], wb[G], vb[G];
          fetch(p0, wa, va);
          for (int p = p0;; p += 2)
          {
            const bool hb = p + 1 < p1;
            if (hb)
              fetch(p + 1, wb, vb);
            mac(wa, va);
            if (!hb)
              break;
            const bool ha = p + 2 < p1;
            if (ha)
              fetch(p + 2, wa, va);
            mac(wb, vb);
            if (!ha)
              break;
          }
        }
#endif
        Rows<G>::store(s.sTH + (size_t)job * G, rot, acc);
      }
      // row constants: chem_jac_isobaric rows (:75-98) folded with transform_isobaric_primitive_jacobian (:319-343).
      // (Dealing them first to the threads of the warps that carry no parts, -DGB_JAC_CS_OFFSET, is 1.4 % slower on the
      // whole kernel: 6.85 against 6.75 ms, paired runs.)
#ifdef GB_JAC_CS_OFFSET
      for (int item = (tid + nt - (((ncsp + 31) & ~31) % nt)) % nt; item < ns * G; item += nt)
#else
      for (int item = tid; item < ns * G; item += nt)
#endif
      {
        const int i = item / G, g = item - i * G;
        const double invRho = SMG(s.sc, J_IRHO, g), rho = SMG(s.sc, J_RHO, g);
        const double w = SJ(rowsrc[i], g), wr = SJ(rowsrc[ns + i], g);
        const double wT = SJ(rowsrc[2 * ns + i], g);
        const double nmA = SJ(rowsrc[3 * ns + i], g), nmB = SJ(rowsrc[4 * ns + i], g);
        const double prho = invRho * (wr - invRho * w); // P[1+i, rho]
        const double nRM = -rho * SMG(s.sc, J_MMW, g), roT = rho / SMG(s.sc, J_T, g);
        SMG(s.sg, i, g) = invRho;
        SMG(s.sdb, i, g) = invRho * nmA + nRM * prho;
        SMG(s.sdcp, i, g) = invRho * nmB;
        if (i < nsm1)
        {
          SJ(dm.jp_c0base + i, g) = wT * invRho - roT * prho; // J[1+i, 0]
          if (reactor && g < gcount)
          { // right-hand side, chem_rhs_isobaric :19-29 (+ :194-218)
            double v = w * invRho;
            if (open)
              v += (a.rx.y_in[i] - SMG(s.sy, i, g)) * invTau;
            a.out0[(size_t)(tile0 + g) * ns + 1 + i] = v;
          }
        }
      }
    }
    TL_MARK(8)
    __syncthreads();
    // ---- column sums: the parts of a destination added up in part order, once per (destination, state), into the slot
    // of its first part (every temperature-row entry needs seven of these sums: formed inside that phase they were
    // seven dependent loops on its critical path, and the six per-state ones were formed by all ns threads of a state)
    for (int item = tid; item < dm.jp_ncs * G; item += nt)
    {
      const int d = item / G, g = item - d * G;
      const int p0 = t_cspfirst[d], p1 = t_cspfirst[d + 1];
      double v = SMG(s.sTH, p0, g);
      for (int p = p0 + 1; p < p1; ++p)
        v += SMG(s.sTH, p, g);
      SMG(s.sTH, p0, g) = v;
    }
    __syncthreads();
    // ---- temperature row (:58-98, 142-168; flamelet_kernels.cpp:1290-1320) ----------------------------------------------------
    finish_T();
    for (int item = tid; item < ns * G; item += nt)
    {
      const int c = item / G, g = item - c * G;
      const int ncs = dm.jp_ncs;
      auto colsum = [&](int d) { return SMG(s.sTH, t_cspfirst[d], g); };
      const double rho = SMG(s.sc, J_RHO, g), cp = SMG(s.sc, J_CP, g), T = SMG(s.sc, J_T, g);
      const double cpsensT = SMG(s.sc, J_DCP, g);
      const double invRhoCp = 1. / (rho * cp), invRho = SMG(s.sc, J_IRHO, g), invCp = 1. / cp;
      const double SW = colsum(nsm1 + 0), SWr = colsum(nsm1 + 1), SWT = colsum(nsm1 + 2);
      const double SA = colsum(nsm1 + 3), SB = colsum(nsm1 + 4), wcp = colsum(ncs - 1);
      const double rhs0c = -SW * invRhoCp;
      double rhs0 = rhs0c;
      double P0rho = -invRhoCp * SWr - invRho * rhs0c;
      double P0T = -invRhoCp * (SWT + wcp) - rhs0c * cpsensT * invCp;
      double cextra = 0.; // extra coefficient of (cp_k - cp_ns) in the T-row of the Y_k columns
      const int sidx = tile0 + (g < gcount ? g : 0);
      if (reactor)
      { // mass_jac_isobaric :100-140, heat_jac_isobaric :142-168, :289-309
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
      }
      else if (flamelet && !fl.adiabatic)
      { // flamelet_kernels.cpp:1290-1320
        const int F = sidx / fl.nzi, iz = sidx - F * fl.nzi;
        const size_t ho = (size_t)F * fl.stride_heat + iz;
        const double Tc = fl.T_conv[ho], Tr = fl.T_rad[ho], hc = fl.h_conv[ho], hr = fl.h_rad[ho];
        double q;
        if (fl.use_scaled_heat_loss)
        {
          const double maxT = fl.maxT[F];
          const double maxT4 = maxT * maxT * maxT * maxT;
          const double Tr4 = Tr * Tr * Tr * Tr;
          q = (hc * (Tc - T) / (maxT - Tc) + hr * 5.67e-8 * (Tr4 - T * T * T * T) / (maxT4 - Tr4)) * invRhoCp;
          P0T -= invCp * cpsensT * q + invRhoCp * (hc / (maxT - Tc) + 4. * hr / (maxT4 - Tr4) * 5.67e-8 * T * T * T);
        }
        else
        {
          q = (hc * (Tc - T) + hr * 5.67e-8 * (Tr * Tr * Tr * Tr - T * T * T * T)) * invRhoCp;
          P0T -= invCp * cpsensT * q + invRhoCp * (hc + 4. * hr * 5.67e-8 * T * T * T);
        }
        P0rho -= q / rho;
        cextra += -invCp * q;
      }
      const double roT = rho * SMG(s.sc, J_INVT, g), nRM = -rho * SMG(s.sc, J_MMW, g);
      double v;
      if (isothermal)
        v = 0.;
      else if (c == 0)
        v = P0T - roT * P0rho;
      else
      {
        const int k = c - 1;
        const double uk = s.su[k];
        const double sum = colsum(k) + uk * SA + SB;
        const double pY = -sum * invRhoCp + (-rhs0c * invCp + cextra) * (SMG(s.scp, k, g) - SMG(s.scp, nsm1, g));
        v = pY + nRM * uk * P0rho;
      }
      SJ(dm.jp_t0base + c, g) = v;
      if (c == 0)
      {
        if (reactor && g < gcount)
          a.out0[(size_t)(tile0 + g) * ns] = isothermal ? 0. : rhs0;
        // destination of the state's block, and the flamelet extras of its grid point
        size_t obase;
        double ttc = 0.;
        if (!flamelet || fl.chem_only)
          obase = (size_t)sidx * ns * ns; // dense blocks (the eigenvalue pass reads them in place)
        else
        { // block iz of flamelet F in BTDDOD storage
          const int nzi = fl.nzi, F = sidx / nzi, iz = sidx - F * nzi;
          obase = (size_t)F * ((size_t)ns * ((size_t)nzi * ns + 2 * (nzi - 1))) + (size_t)iz * ns * ns;
          SMG(s.sc, J_CMOFF, g) = __longlong_as_double((long long)((size_t)F * fl.stride_coeff + (size_t)iz * ns));
          if (fl.include_enthalpy_flux)
          { // (T,T) correction, flamelet_kernels.cpp:1350-1381
            const double *stt = a.in_state + (size_t)F * nzi * ns;
            const double *cpg = fl.cp_grid + (size_t)F * nzi;
            const double mc = fl.mcoeff[(size_t)F * fl.stride_mn + iz], nc = fl.ncoeff[(size_t)F * fl.stride_mn + iz];
            const double Tm = (iz == 0) ? fl.oxy[0] : stt[(size_t)(iz - 1) * ns];
            const double Tp = (iz == nzi - 1) ? fl.fuel[0] : stt[(size_t)(iz + 1) * ns];
            const double cpm = (iz == 0) ? fl.cp_bc[0] : cpg[iz - 1];
            const double cpp = (iz == nzi - 1) ? fl.cp_bc[1] : cpg[iz + 1];
            const double dTdZ = mc * Tm + nc * Tp, dcpdZ = mc * cpm + nc * cpp;
            const double f1 = 0.5 * fl.chi[(size_t)F * fl.stride_chi + iz] / cp * dTdZ * dcpdZ;
            ttc = f1 / cp * cpsensT;
          }
        }
        SMG(s.sc, J_OBASE, g) = __longlong_as_double((long long)obase);
        SMG(s.sc, J_TTC, g) = ttc;
      }
    }
    TL_MARK(9)
    __syncthreads();

    // ---- output: column 0, then the columns 1..ns-1 with one row per thread ----------------------------------------------------
#ifndef GB_JAC_EARLY_PREFETCH
    pre = fetch_state(tile + gridDim.x);
#endif
    for (int item = tid; item < ns * G; item += nt)
    {
      const int g = item / ns, r = item - g * ns;
      if (g >= gcount)
        continue;
      double v = (r == 0) ? SJ(dm.jp_t0base, g) : SJ(dm.jp_c0base + r - 1, g);
      const size_t ob = (size_t)__double_as_longlong(SMG(s.sc, J_OBASE, g));
      if (flamelet && !fl.chem_only)
      {
        if (r == 0)
        {
          v += fl.cmajor[(size_t)__double_as_longlong(SMG(s.sc, J_CMOFF, g))];
          v -= SMG(s.sc, J_TTC, g);
        }
        if (fl.scale_and_offset)
        {
          v *= fl.prefactor;
          if (r == 0)
            v -= 1.;
        }
      }
      a.out1[ob + r] = v;
    }
    {
      const int cpi = nt / ns; // columns per iteration
      if (tid < cpi * ns)
      {
        const int r = tid % ns;
        int c = 1 + tid / ns;
        // registers hold the states in canonical order (slot = state): every store instruction writes one state
        double c1[G], c2[G], c3[G];
        double *ob[G];
        if (r == 0)
        {
#pragma unroll
          for (int g = 0; g < G; ++g)
            c1[g] = 1., c2[g] = 0., c3[g] = 0.;
        }
        else
        {
          Rows<G>::load(s.sg + (size_t)(r - 1) * G, 0, c1);
          Rows<G>::load(s.sdb + (size_t)(r - 1) * G, 0, c2);
          Rows<G>::load(s.sdcp + (size_t)(r - 1) * G, 0, c3);
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
          ob[g] = a.out1 + (size_t)__double_as_longlong(SMG(s.sc, J_OBASE, g)) + r;
        for (; c < ns; c += cpi)
        {
          const int e = r + (ns + 1) * (c - 1);
          const double uk = s.su[c - 1];
          double v[G];
          Rows<G>::load_swz(s.sR, (int)s.semap[e], 0, v);
#pragma unroll
          for (int sl = 0; sl < G; ++sl)
            v[sl] = fma(c1[sl], v[sl], fma(uk, c2[sl], c3[sl]));
          if (r == c)
          {
            if (open)
            {
#pragma unroll
              for (int sl = 0; sl < G; ++sl)
                v[sl] += -invTau;
            }
            if (flamelet && !fl.chem_only)
            {
#pragma unroll
              for (int sl = 0; sl < G; ++sl)
                v[sl] += fl.cmajor[(size_t)__double_as_longlong(SMG(s.sc, J_CMOFF, sl)) + r];
            }
          }
          if (flamelet && fl.scale_and_offset)
          {
#pragma unroll
            for (int sl = 0; sl < G; ++sl)
            {
              v[sl] *= fl.prefactor;
              if (r == c)
                v[sl] -= 1.;
            }
          }
          const int off = ns * c;
#pragma unroll
          for (int sl = 0; sl < G; ++sl)
            if (sl < gcount)
            {
              // (streaming store: the Jacobian is written once and never read back by this kernel; measured 1 % on the
              // whole kernel, 6.88 against 6.95 ms per 262,144 states)
#ifdef GB_JAC_NO_STCS
              ob[sl][off] = v[sl];
#else
              __stcs(ob[sl] + off, v[sl]);
#endif
            }
        }
      }
    }
    TL_MARK(10)
  }
}

// ------------------------------------------------------------------------------------------------------------------
static int jac_sm_count()
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

template <int G>
static cudaError_t launch_jac_g(const ChemArgs &a, size_t smem, cudaStream_t s)
{
  static bool attr = false;
  if (!attr)
  {
    cudaError_t e = cudaFuncSetAttribute(k_jac<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
    if (e != cudaSuccess)
      return e;
    attr = true;
  }
  const int threads = a.dm.jp_threads;
  const int ntiles = (a.n + G - 1) / G;
  // CTAs per SM: limited by shared memory and by 64K registers / (threads * 128)
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>((size_t)(227 * 1024) / (smem + 1024), (size_t)(512 / threads)));
  per_sm = std::min(per_sm, 8);
  if (const char *e = getenv("GB_JAC_CTAS"))
    per_sm = std::max(1, atoi(e));
  int grid = std::max(1, std::min(ntiles, jac_sm_count() * per_sm));
  if (const char *e = getenv("GB_JAC_GRID"))
    grid = std::max(1, std::min(grid, atoi(e)));
  ChemArgs b = a;
  b.stagger = 0;
  if (const char *e = getenv("GB_JAC_STAGGER"))
    b.stagger = std::max(0, atoi(e));
  k_jac<G><<<grid, threads, smem, s>>>(b);
  ++g_jac_launches;
  return cudaGetLastError();
}

#ifdef GB_JAC_TIMELINE
int debug_jac_timeline(long long *out)
{
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_jac_timeline, sizeof(long long) * 20 * 32) == cudaSuccess ? 0 : -3;
}
#endif

cudaError_t launch_jac(const ChemArgs &a_in, cudaStream_t s)
{
  if (jac4_applicable(a_in))
    return launch_jac4(a_in, s);
  ChemArgs a = a_in;
  const size_t smem = (size_t)a.dm.jp_smem;
  a.G = a.dm.jp_G;
  a.GS = a.G;
  switch (a.G)
  {
  case 8:
    return launch_jac_g<8>(a, smem, s);
  case 4:
    return launch_jac_g<4>(a, smem, s);
  case 2:
    return launch_jac_g<2>(a, smem, s);
  case 1:
    return launch_jac_g<1>(a, smem, s);
  default:
    return cudaErrorInvalidConfiguration;
  }
}

} // namespace gb
