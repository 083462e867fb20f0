// gb_api.cu -- C-ABI compute entry points (include/griffon_b200.h): argument checks, lazy mechanism commit,
// kernel launches, and the *_host variants that stage host buffers through device scratch owned by the handle.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/griffon_b200.h"
#include "gb_kernels.cuh"
#include "gb_mech.h"

namespace gb
{
int nonfinite_reset(cudaStream_t st);
int nonfinite_accumulate(int n, long la, const double *a, long lb, const double *b, int *flags, cudaStream_t st);
int nonfinite_read(cudaStream_t st);
} // namespace gb

using namespace gb;

#define GB_STR2(x) #x
#define GB_STR(x) GB_STR2(x)
#ifndef GB_FLAGS
#define GB_FLAGS ""
#endif

namespace
{

int cuda_fail(cudaError_t e, const char *what)
{
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  cudaGetLastError();
  return GB_ERR_CUDA;
}

int ready(gb_mech *m)
{
  if (!m)
  {
    set_error("null mechanism handle");
    return GB_ERR_ARG;
  }
  return commit(m->h);
}

// device scratch slot k of at least `bytes` bytes
int scratch(gb_mech *m, int k, size_t bytes, void **out)
{
  HostMech &h = m->h;
  if (h.d_scratch_bytes[k] < bytes)
  {
    if (h.d_scratch[k])
      cudaFree(h.d_scratch[k]);
    h.d_scratch[k] = nullptr;
    h.d_scratch_bytes[k] = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&h.d_scratch[k], want);
    if (e != cudaSuccess)
      return cuda_fail(e, "scratch allocation");
    h.d_scratch_bytes[k] = want;
  }
  *out = h.d_scratch[k];
  return GB_OK;
}

#define CK(call)                         \
  do                                     \
  {                                      \
    cudaError_t e__ = (call);            \
    if (e__ != cudaSuccess)              \
      return cuda_fail(e__, #call);      \
  } while (0)
#define RC(call)            \
  do                        \
  {                         \
    int rc__ = (call);      \
    if (rc__ != GB_OK)      \
      return rc__;          \
  } while (0)

ReactorDev reactor_dev(const gb_reactor_params *p, const double *d_yin)
{
  ReactorDev r;
  r.p = p->pressure;
  r.T_in = p->inflow_temperature;
  r.tau = p->tau;
  r.T_inf = p->fluid_temperature;
  r.T_surf = p->surf_temperature;
  r.h_conv = p->h_conv;
  r.eps_rad = p->eps_rad;
  r.SoV = p->surface_area_over_volume;
  r.y_in = d_yin;
  r.heat_option = p->heat_transfer_option;
  r.open = p->open ? 1 : 0;
  return r;
}

int check_reactor(const gb_mech *m, int n, const void *state, const gb_reactor_params *p, const void *out)
{
  (void)m;
  if (n < 0 || (n > 0 && (!state || !out)) || !p)
  {
    set_error("bad reactor arguments");
    return GB_ERR_ARG;
  }
  if (p->heat_transfer_option < 0 || p->heat_transfer_option > 2)
  {
    set_error("heat_transfer_option must be 0, 1 or 2");
    return GB_ERR_ARG;
  }
  if (p->open && !p->inflow_y)
  {
    set_error("open reactor needs inflow_y");
    return GB_ERR_ARG;
  }
  return GB_OK;
}

} // namespace

extern "C"
{

  long gb_kernel_launch_count(void) { return kernel_launch_count(); }
  // host-clock accounting of gb_flamelet_async_tick_batch (one integrator at a time): {ticks, rounds, seconds inside
  // the calls, seconds inside the round loops, seconds from launch to synchronised end of the solves (GB_TICK_PROFILE=1
  // only: adds a synchronisation per round)}
  static double g_tick_stats[8] = {0};
  void gb_debug_tick_stats(double *out, int reset)
  {
    for (int k = 0; k < 8; ++k)
    {
      out[k] = g_tick_stats[k];
      if (reset)
        g_tick_stats[k] = 0.;
    }
  }
#ifdef GB_JAC_TIMELINE
  int gb_debug_jac_timeline(long long *out /* [16*32] */) { return gb::debug_jac_timeline(out); }
  int gb_debug_bt_timeline(long long *out /* [8] */) { return gb::debug_bt_timeline(out); }
  int gb_debug_jac4_timeline(long long *out /* [16*32] */) { return gb::debug_jac4_timeline(out); }
#endif
  const char *gb_build_info(void) { return "griffon_b200 sm_100a fp64, nvcc " GB_STR(__CUDACC_VER_MAJOR__) "." GB_STR(__CUDACC_VER_MINOR__) GB_FLAGS; }

  // ---- thermo ----------------------------------------------------------------------------------------------------
  int gb_thermo_batch(gb_mech *m, int what, int n, const double *aux, const double *T, const double *y, double *out,
                      void *stream)
  {
    RC(ready(m));
    if (what < 0 || what > 12 || n < 0 || !out)
    {
      set_error("bad thermo arguments");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return GB_OK;
    CK(launch_thermo(m->h.dm, what, n, aux, T, y, out, (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_thermo_host(gb_mech *m, int what, int n, const double *aux, const double *T, const double *y, double *out)
  {
    RC(ready(m));
    if (what < 0 || what > 12 || n < 0 || !out)
    {
      set_error("bad thermo arguments");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    const bool per_species = what >= 7;
    void *d_aux = nullptr, *d_T = nullptr, *d_y = nullptr, *d_out = nullptr;
    RC(scratch(m, 0, sizeof(double) * n, &d_aux));
    RC(scratch(m, 1, sizeof(double) * n, &d_T));
    RC(scratch(m, 2, sizeof(double) * n * ns, &d_y));
    RC(scratch(m, 3, sizeof(double) * n * (per_species ? ns : 1), &d_out));
    if (aux)
      CK(cudaMemcpy(d_aux, aux, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (T)
      CK(cudaMemcpy(d_T, T, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (y)
      CK(cudaMemcpy(d_y, y, sizeof(double) * n * ns, cudaMemcpyHostToDevice));
    CK(launch_thermo(m->h.dm, what, n, aux ? (double *)d_aux : nullptr, T ? (double *)d_T : nullptr,
                     y ? (double *)d_y : nullptr, (double *)d_out, 0));
    CK(cudaMemcpy(out, d_out, sizeof(double) * n * (per_species ? ns : 1), cudaMemcpyDeviceToHost));
    return GB_OK;
  }

  // ---- kinetics --------------------------------------------------------------------------------------------------
  int gb_production_rates_batch(gb_mech *m, int n, const double *T, const double *rho, const double *y, double *out_w,
                                void *stream)
  {
    RC(ready(m));
    if (n < 0 || (n > 0 && (!T || !rho || !y || !out_w)))
    {
      set_error("bad production_rates arguments");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return GB_OK;
    ChemArgs a{};
    a.dm = m->h.dm;
    a.mode = MODE_PRODRATES;
    a.n = n;
    a.in_T = T, a.in_rho = rho, a.in_y = y;
    a.out0 = out_w;
    CK(launch_rates(a, (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_production_rates_host(gb_mech *m, int n, const double *T, const double *rho, const double *y, double *out_w)
  {
    RC(ready(m));
    if (n <= 0)
      return n == 0 ? GB_OK : GB_ERR_ARG;
    const int ns = m->h.dm.ns;
    void *d_T, *d_rho, *d_y, *d_w;
    RC(scratch(m, 0, sizeof(double) * n, &d_T));
    RC(scratch(m, 1, sizeof(double) * n, &d_rho));
    RC(scratch(m, 2, sizeof(double) * n * ns, &d_y));
    RC(scratch(m, 3, sizeof(double) * n * ns, &d_w));
    CK(cudaMemcpy(d_T, T, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_rho, rho, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_y, y, sizeof(double) * n * ns, cudaMemcpyHostToDevice));
    RC(gb_production_rates_batch(m, n, (double *)d_T, (double *)d_rho, (double *)d_y, (double *)d_w, nullptr));
    CK(cudaMemcpy(out_w, d_w, sizeof(double) * n * ns, cudaMemcpyDeviceToHost));
    return GB_OK;
  }

  int gb_prod_rates_sens_batch(gb_mech *m, int n, const double *rho, const double *T, const double *y,
                               int rates_sensitivity_option, double *out_sens, void *stream)
  {
    RC(ready(m));
    if (n < 0 || (n > 0 && (!T || !rho || !y || !out_sens)) || rates_sensitivity_option < 0 ||
        rates_sensitivity_option > 2)
    {
      set_error("bad prod_rates_sens arguments");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return GB_OK;
    ChemArgs a{};
    a.dm = m->h.dm;
    a.mode = MODE_SENS;
    a.n = n;
    a.in_T = T, a.in_rho = rho, a.in_y = y;
    a.out1 = out_sens;
    CK(launch_jac(a, (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_prod_rates_sens_host(gb_mech *m, int n, const double *rho, const double *T, const double *y,
                              int rates_sensitivity_option, double *out_sens)
  {
    RC(ready(m));
    if (n <= 0)
      return n == 0 ? GB_OK : GB_ERR_ARG;
    const int ns = m->h.dm.ns;
    const size_t so = (size_t)(ns + 1) * (ns + 1);
    void *d_T, *d_rho, *d_y, *d_s;
    RC(scratch(m, 0, sizeof(double) * n, &d_T));
    RC(scratch(m, 1, sizeof(double) * n, &d_rho));
    RC(scratch(m, 2, sizeof(double) * n * ns, &d_y));
    RC(scratch(m, 3, sizeof(double) * n * so, &d_s));
    CK(cudaMemcpy(d_T, T, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_rho, rho, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_y, y, sizeof(double) * n * ns, cudaMemcpyHostToDevice));
    RC(gb_prod_rates_sens_batch(m, n, (double *)d_rho, (double *)d_T, (double *)d_y, rates_sensitivity_option,
                                (double *)d_s, nullptr));
    CK(cudaMemcpy(out_sens, d_s, sizeof(double) * n * so, cudaMemcpyDeviceToHost));
    return GB_OK;
  }

  // ---- isobaric reactor ----------------------------------------------------------------------------------------
  int gb_reactor_rhs_isobaric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                    double *out_rhs, void *stream)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n == 0)
      return GB_OK;
    ChemArgs a{};
    a.dm = m->h.dm;
    a.mode = MODE_REACTOR_RHS;
    a.n = n;
    a.in_state = state;
    a.p = prm->pressure;
    a.out0 = out_rhs;
    a.rx = reactor_dev(prm, prm->inflow_y);
    CK(launch_rates(a, (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_reactor_jac_isobaric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                    int rates_sensitivity_option, int sensitivity_transform_option, double *out_rhs,
                                    double *out_jac, void *stream)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n > 0 && !out_jac)
    {
      set_error("out_jac is null");
      return GB_ERR_ARG;
    }
    if (rates_sensitivity_option < 0 || rates_sensitivity_option > 2 || sensitivity_transform_option != 0)
    {
      set_error("rates_sensitivity_option must be 0..2 and sensitivity_transform_option 0");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return GB_OK;
    ChemArgs a{};
    a.dm = m->h.dm;
    a.mode = MODE_REACTOR_JAC;
    a.n = n;
    a.in_state = state;
    a.p = prm->pressure;
    a.out0 = out_rhs;
    a.out1 = out_jac;
    a.rx = reactor_dev(prm, prm->inflow_y);
    CK(launch_jac(a, (cudaStream_t)stream));
    return GB_OK;
  }

  static int stage_reactor(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                           gb_reactor_params *dprm, double **d_state)
  {
    const int ns = m->h.dm.ns;
    void *p_state, *p_yin;
    RC(scratch(m, 0, sizeof(double) * n * ns, &p_state));
    RC(scratch(m, 1, sizeof(double) * ns, &p_yin));
    CK(cudaMemcpy(p_state, state, sizeof(double) * n * ns, cudaMemcpyHostToDevice));
    *dprm = *prm;
    if (prm->open)
    {
      CK(cudaMemcpy(p_yin, prm->inflow_y, sizeof(double) * ns, cudaMemcpyHostToDevice));
      dprm->inflow_y = (const double *)p_yin;
    }
    else
      dprm->inflow_y = nullptr;
    *d_state = (double *)p_state;
    return GB_OK;
  }

  int gb_reactor_rhs_isobaric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                   double *out_rhs)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    gb_reactor_params dprm;
    double *d_state;
    RC(stage_reactor(m, n, state, prm, &dprm, &d_state));
    void *d_rhs;
    RC(scratch(m, 2, sizeof(double) * n * ns, &d_rhs));
    RC(gb_reactor_rhs_isobaric_batch(m, n, d_state, &dprm, (double *)d_rhs, nullptr));
    const int bad = gb_count_nonfinite_members_batch(n, ns, (const double *)d_rhs, 0, nullptr, nullptr, nullptr);
    CK(cudaMemcpy(out_rhs, d_rhs, sizeof(double) * n * ns, cudaMemcpyDeviceToHost));
    return bad; // > 0: states whose right-hand side holds an Inf or NaN
  }

  // Pipeline resources of the host entry points: three streams (host->device, kernels, device->host) and the events
  // that order the two halves of the double-buffered device staging area. Process-wide, created on first use; the
  // *_host entry points are synchronous, so one set serves every handle of the calling thread's device.
  struct HostPipe
  {
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int device = -1;
  };
  static int host_pipe(HostPipe **out)
  {
    static HostPipe pipes[16];
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16)
    {
      set_error("device ordinal out of range for the host pipeline");
      return GB_ERR_CUDA;
    }
    HostPipe &p = pipes[dev];
    if (p.device != dev)
    {
      CK(cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b)
      {
        CK(cudaEventCreateWithFlags(&p.ev_in[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p.ev_k[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p.ev_out[b], cudaEventDisableTiming));
      }
      p.device = dev;
    }
    *out = &p;
    return GB_OK;
  }

  // states per chunk of the pipelined host path: about 256 MB of Jacobian per chunk, a whole number of waves of
  // 8-state tiles over 148 SMs; batches of up to two chunks take the serial path (nothing to overlap)
  static int host_chunk_states(int ns)
  {
    const size_t per = sizeof(double) * (size_t)ns * ns;
    size_t c = ((size_t)256 << 20) / per;
    const size_t wave = 8 * 148;
    c = std::max<size_t>(wave, c / wave * wave);
    if (const char *e = getenv("GB_HOST_CHUNK"))
      c = std::max(1, atoi(e));
    return (int)std::min<size_t>(c, (size_t)1 << 30);
  }

  int gb_reactor_jac_isobaric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                   int rates_sensitivity_option, int sensitivity_transform_option, double *out_rhs,
                                   double *out_jac)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    const int chunk = host_chunk_states(ns);
    if (n <= 2 * chunk)
    { // small batch: copy in, one launch, copy out
      gb_reactor_params dprm;
      double *d_state;
      RC(stage_reactor(m, n, state, prm, &dprm, &d_state));
      void *d_rhs, *d_jac;
      RC(scratch(m, 2, sizeof(double) * n * ns, &d_rhs));
      RC(scratch(m, 3, sizeof(double) * (size_t)n * ns * ns, &d_jac));
      RC(gb_reactor_jac_isobaric_batch(m, n, d_state, &dprm, rates_sensitivity_option, sensitivity_transform_option,
                                       (double *)d_rhs, (double *)d_jac, nullptr));
      const int bad = gb_count_nonfinite_members_batch(n, ns, (const double *)d_rhs, (long)ns * ns, (const double *)d_jac,
                                                       nullptr, nullptr);
      CK(cudaMemcpy(out_rhs, d_rhs, sizeof(double) * n * ns, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out_jac, d_jac, sizeof(double) * (size_t)n * ns * ns, cudaMemcpyDeviceToHost));
      return bad; // > 0: states whose right-hand side or Jacobian holds an Inf or NaN
    }
    // Large batch: chunks flow through a double-buffered staging area on three streams, so the device->host copy of
    // chunk c (the 8*ns^2 bytes per state that bound this entry point) overlaps the kernel of chunk c+1 and the
    // host->device copy of chunk c+2. Pinned host buffers make the copies asynchronous; pageable ones still work
    // (the runtime stages them), only without overlap.
    HostPipe *pp;
    RC(host_pipe(&pp));
    HostPipe &P = *pp;
    void *p_state, *p_yin, *p_rhs, *p_jac;
    const size_t sb = sizeof(double) * (size_t)chunk * ns, jb = sizeof(double) * (size_t)chunk * ns * ns;
    RC(scratch(m, 0, 2 * sb, &p_state));
    RC(scratch(m, 1, sizeof(double) * ns, &p_yin));
    RC(scratch(m, 2, 2 * sb, &p_rhs));
    RC(scratch(m, 3, 2 * jb, &p_jac));
    gb_reactor_params dprm = *prm;
    if (prm->open)
    {
      CK(cudaMemcpy(p_yin, prm->inflow_y, sizeof(double) * ns, cudaMemcpyHostToDevice));
      dprm.inflow_y = (const double *)p_yin;
    }
    else
      dprm.inflow_y = nullptr;
    CK(cudaDeviceSynchronize()); // earlier work on the handle's buffers (legacy stream) is done
    RC(gb::nonfinite_reset(P.s_k));
    int rc = GB_OK;
    const int nchunks = (n + chunk - 1) / chunk;
    for (int c = 0; c < nchunks && rc == GB_OK; ++c)
    {
      const int b = c & 1, lo = c * chunk, cnt = std::min(chunk, n - lo);
      double *d_state = (double *)p_state + (size_t)b * chunk * ns, *d_rhs = (double *)p_rhs + (size_t)b * chunk * ns;
      double *d_jac = (double *)p_jac + (size_t)b * chunk * ns * ns;
      if (c >= 2)
      { // buffer b is free once chunk c-2 has left the device (its kernel, which read d_state, finished before that)
        CK(cudaStreamWaitEvent(P.s_in, P.ev_out[b], 0));
        CK(cudaStreamWaitEvent(P.s_k, P.ev_out[b], 0));
      }
      CK(cudaMemcpyAsync(d_state, state + (size_t)lo * ns, sizeof(double) * (size_t)cnt * ns, cudaMemcpyHostToDevice,
                         P.s_in));
      CK(cudaEventRecord(P.ev_in[b], P.s_in));
      CK(cudaStreamWaitEvent(P.s_k, P.ev_in[b], 0));
      rc = gb_reactor_jac_isobaric_batch(m, cnt, d_state, &dprm, rates_sensitivity_option,
                                         sensitivity_transform_option, d_rhs, d_jac, P.s_k);
      if (rc != GB_OK)
        break;
      rc = gb::nonfinite_accumulate(cnt, ns, d_rhs, (long)ns * ns, d_jac, nullptr, P.s_k);
      if (rc != GB_OK)
        break;
      CK(cudaEventRecord(P.ev_k[b], P.s_k));
      CK(cudaStreamWaitEvent(P.s_out, P.ev_k[b], 0));
      CK(cudaMemcpyAsync(out_rhs + (size_t)lo * ns, d_rhs, sizeof(double) * (size_t)cnt * ns, cudaMemcpyDeviceToHost,
                         P.s_out));
      CK(cudaMemcpyAsync(out_jac + (size_t)lo * ns * ns, d_jac, sizeof(double) * (size_t)cnt * ns * ns,
                         cudaMemcpyDeviceToHost, P.s_out));
      CK(cudaEventRecord(P.ev_out[b], P.s_out));
    }
    CK(cudaStreamSynchronize(P.s_in));
    const int bad = rc == GB_OK ? gb::nonfinite_read(P.s_k) : rc;
    CK(cudaStreamSynchronize(P.s_k));
    CK(cudaStreamSynchronize(P.s_out));
    return bad; // > 0: states whose right-hand side or Jacobian holds an Inf or NaN
  }
}

// ---- isochoric reactor ---------------------------------------------------------------------------------------------
namespace gb
{
cudaError_t launch_iso_split(int ns, int n, const double *state, double *rho, double *T, double *y, cudaStream_t s);
cudaError_t launch_iso_assemble(const DeviceMech &dm, int n, const double *state, const double *y, const double *w,
                                const double *wsens, const ReactorDev &rx, double rho_in, double *out_rhs,
                                double *out_jac, cudaStream_t s);
} // namespace gb

namespace
{
// rates (and, with out_jac, exact sensitivities) of the batch through the kernels of the isobaric path, then the
// isochoric assembly. Work arrays live in the handle's scratch slots 4 and 5 (a handle serves one stream at a time).
int isochoric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm, double rho_in, double *out_rhs,
                    double *out_jac, void *stream)
{
  const int ns = m->h.dm.ns;
  cudaStream_t st = (cudaStream_t)stream;
  void *p_work, *p_sens = nullptr;
  RC(scratch(m, 4, sizeof(double) * (size_t)n * (2 + 2 * (size_t)ns), &p_work));
  if (out_jac)
    RC(scratch(m, 5, sizeof(double) * (size_t)n * (ns + 1) * (ns + 1), &p_sens));
  double *rho = (double *)p_work, *T = rho + n, *y = T + n, *w = y + (size_t)n * ns;
  CK(gb::launch_iso_split(ns, n, state, rho, T, y, st));
  ChemArgs a{};
  a.dm = m->h.dm;
  a.mode = MODE_PRODRATES;
  a.n = n;
  a.in_T = T, a.in_rho = rho, a.in_y = y;
  a.out0 = w;
  CK(launch_rates(a, st));
  if (out_jac)
  {
    a.mode = MODE_SENS;
    a.out0 = nullptr;
    a.out1 = (double *)p_sens;
    CK(launch_jac(a, st));
  }
  const ReactorDev rx = reactor_dev(prm, prm->inflow_y);
  CK(gb::launch_iso_assemble(m->h.dm, n, state, y, w, (const double *)p_sens, rx, rho_in, out_rhs, out_jac, st));
  return GB_OK;
}
} // namespace

extern "C"
{
  int gb_reactor_rhs_isochoric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                     double inflow_density, double *out_rhs, void *stream)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n == 0)
      return GB_OK;
    return isochoric_batch(m, n, state, prm, inflow_density, out_rhs, nullptr, stream);
  }

  int gb_reactor_jac_isochoric_batch(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                     double inflow_density, int rates_sensitivity_option, double *out_rhs,
                                     double *out_jac, void *stream)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if ((n > 0 && !out_jac) || rates_sensitivity_option < 0 || rates_sensitivity_option > 2)
    {
      set_error("bad isochoric Jacobian arguments");
      return GB_ERR_ARG;
    }
    if (n == 0)
      return GB_OK;
    return isochoric_batch(m, n, state, prm, inflow_density, out_rhs, out_jac, stream);
  }

  static int stage_isochoric(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                             gb_reactor_params *dprm, double **d_state)
  {
    const int ns = m->h.dm.ns;
    void *p_state, *p_yin;
    RC(scratch(m, 0, sizeof(double) * (size_t)n * (ns + 1), &p_state));
    RC(scratch(m, 1, sizeof(double) * ns, &p_yin));
    CK(cudaMemcpy(p_state, state, sizeof(double) * (size_t)n * (ns + 1), cudaMemcpyHostToDevice));
    *dprm = *prm;
    if (prm->open)
    {
      CK(cudaMemcpy(p_yin, prm->inflow_y, sizeof(double) * ns, cudaMemcpyHostToDevice));
      dprm->inflow_y = (const double *)p_yin;
    }
    else
      dprm->inflow_y = nullptr;
    *d_state = (double *)p_state;
    return GB_OK;
  }

  int gb_reactor_rhs_isochoric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                    double inflow_density, double *out_rhs)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    gb_reactor_params dprm;
    double *d_state;
    RC(stage_isochoric(m, n, state, prm, &dprm, &d_state));
    void *d_rhs;
    RC(scratch(m, 2, sizeof(double) * (size_t)n * (ns + 1), &d_rhs));
    RC(gb_reactor_rhs_isochoric_batch(m, n, d_state, &dprm, inflow_density, (double *)d_rhs, nullptr));
    const int bad = gb_count_nonfinite_members_batch(n, ns + 1, (const double *)d_rhs, 0, nullptr, nullptr, nullptr);
    CK(cudaMemcpy(out_rhs, d_rhs, sizeof(double) * (size_t)n * (ns + 1), cudaMemcpyDeviceToHost));
    return bad;
  }

  int gb_reactor_jac_isochoric_host(gb_mech *m, int n, const double *state, const gb_reactor_params *prm,
                                    double inflow_density, int rates_sensitivity_option, double *out_rhs,
                                    double *out_jac)
  {
    RC(ready(m));
    RC(check_reactor(m, n, state, prm, out_rhs));
    if (n == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    gb_reactor_params dprm;
    double *d_state;
    RC(stage_isochoric(m, n, state, prm, &dprm, &d_state));
    void *d_rhs, *d_jac;
    RC(scratch(m, 2, sizeof(double) * (size_t)n * (ns + 1), &d_rhs));
    RC(scratch(m, 3, sizeof(double) * (size_t)n * (ns + 1) * (ns + 1), &d_jac));
    RC(gb_reactor_jac_isochoric_batch(m, n, d_state, &dprm, inflow_density, rates_sensitivity_option, (double *)d_rhs,
                                      (double *)d_jac, nullptr));
    const int bad = gb_count_nonfinite_members_batch(n, ns + 1, (const double *)d_rhs, (long)(ns + 1) * (ns + 1),
                                                     (const double *)d_jac, nullptr, nullptr);
    CK(cudaMemcpy(out_rhs, d_rhs, sizeof(double) * (size_t)n * (ns + 1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_jac, d_jac, sizeof(double) * (size_t)n * (ns + 1) * (ns + 1), cudaMemcpyDeviceToHost));
    return bad;
  }
}

// ---- flamelet ------------------------------------------------------------------------------------------------------
namespace
{

int check_flamelet(int F, const void *state, const gb_flamelet_params *p, const void *out)
{
  if (F < 0 || !p || (F > 0 && (!state || !out)))
  {
    set_error("bad flamelet arguments");
    return GB_ERR_ARG;
  }
  if (p->nzi < 2)
  {
    set_error("a flamelet needs at least two interior grid points");
    return GB_ERR_ARG;
  }
  if (!p->oxy_state || !p->fuel_state || !p->cmajor || !p->csub || !p->csup || !p->mcoeff || !p->ncoeff || !p->chi)
  {
    set_error("flamelet parameter arrays must not be null");
    return GB_ERR_ARG;
  }
  if (!p->adiabatic && (!p->T_convection || !p->h_convection || !p->T_radiation || !p->h_radiation))
  {
    set_error("non-adiabatic flamelets need the four heat-loss arrays");
    return GB_ERR_ARG;
  }
  return GB_OK;
}

// fills the device-side flamelet descriptor and runs the pre-pass (cp grid, max T, boundary cp)
int flamelet_dev(gb_mech *m, int F, const double *d_state, const gb_flamelet_params *p, FlameletDev *fl,
                 cudaStream_t s)
{
  fl->nzi = p->nzi;
  fl->oxy = p->oxy_state, fl->fuel = p->fuel_state;
  fl->adiabatic = p->adiabatic ? 1 : 0;
  fl->T_conv = p->T_convection, fl->h_conv = p->h_convection, fl->T_rad = p->T_radiation, fl->h_rad = p->h_radiation;
  fl->cmajor = p->cmajor, fl->csub = p->csub, fl->csup = p->csup;
  fl->mcoeff = p->mcoeff, fl->ncoeff = p->ncoeff, fl->chi = p->chi;
  fl->stride_heat = p->stride_heat, fl->stride_coeff = p->stride_coeff, fl->stride_mn = p->stride_mn;
  fl->stride_chi = p->stride_chi;
  fl->include_enthalpy_flux = p->include_enthalpy_flux ? 1 : 0;
  fl->include_variable_cp = p->include_variable_cp ? 1 : 0;
  fl->use_scaled_heat_loss = p->use_scaled_heat_loss ? 1 : 0;
  fl->scale_and_offset = 0;
  fl->chem_only = 0;
  fl->prefactor = 1.;
  void *cpg, *mt, *bc;
  RC(scratch(m, 4, sizeof(double) * (size_t)F * p->nzi, &cpg));
  RC(scratch(m, 5, sizeof(double) * (size_t)F, &mt));
  RC(scratch(m, 6, sizeof(double) * 2, &bc));
  fl->cp_grid = (const double *)cpg, fl->maxT = (const double *)mt, fl->cp_bc = (const double *)bc;
  CK(launch_flamelet_prepass(m->h.dm, F, d_state, *fl, (double *)cpg, (double *)mt, (double *)bc, s));
  return GB_OK;
}

// temporary device copies of the host flamelet arrays for the *_host entry points
struct FlameletStage
{
  std::vector<void *> bufs;
  gb_flamelet_params dp;
  ~FlameletStage()
  {
    for (void *b : bufs)
      cudaFree(b);
  }
  int put(const double *h, size_t n, const double **out)
  {
    *out = nullptr;
    if (!h)
      return GB_OK;
    void *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double) * (n ? n : 1)));
    bufs.push_back(d);
    CK(cudaMemcpy(d, h, sizeof(double) * n, cudaMemcpyHostToDevice));
    *out = (const double *)d;
    return GB_OK;
  }
  int stage(int ns, int F, const gb_flamelet_params *p)
  {
    dp = *p;
    const size_t nzi = p->nzi;
    const size_t nh = p->stride_heat ? (size_t)(F - 1) * p->stride_heat + nzi : nzi;
    const size_t nc = p->stride_coeff ? (size_t)(F - 1) * p->stride_coeff + nzi * ns : nzi * ns;
    const size_t nm = p->stride_mn ? (size_t)(F - 1) * p->stride_mn + nzi : nzi;
    const size_t nx = p->stride_chi ? (size_t)(F - 1) * p->stride_chi + nzi + 2 : nzi + 2;
    RC(put(p->oxy_state, ns, &dp.oxy_state));
    RC(put(p->fuel_state, ns, &dp.fuel_state));
    if (!p->adiabatic)
    {
      RC(put(p->T_convection, nh, &dp.T_convection));
      RC(put(p->h_convection, nh, &dp.h_convection));
      RC(put(p->T_radiation, nh, &dp.T_radiation));
      RC(put(p->h_radiation, nh, &dp.h_radiation));
    }
    RC(put(p->cmajor, nc, &dp.cmajor));
    RC(put(p->csub, nc, &dp.csub));
    RC(put(p->csup, nc, &dp.csup));
    RC(put(p->mcoeff, nm, &dp.mcoeff));
    RC(put(p->ncoeff, nm, &dp.ncoeff));
    RC(put(p->chi, nx, &dp.chi));
    return GB_OK;
  }
};

} // namespace

extern "C"
{

  // flamelet_stencils, flamelet_kernels.cpp:31-48 (host, once per Flamelet)
  int gb_flamelet_stencils(const gb_mech *m, const double *dz, int nzi, const double *chi, const double *inv_lewis,
                           double *cmajor, double *csub, double *csup, double *mcoeff, double *ncoeff)
  {
    if (!m || !dz || !chi || !inv_lewis || !cmajor || !csub || !csup || !mcoeff || !ncoeff || nzi < 1)
    {
      set_error("bad flamelet_stencils arguments");
      return GB_ERR_ARG;
    }
    const int ns = (int)m->h.species.size();
    for (int i = 0; i < nzi; ++i)
    {
      const double dzt = dz[i] + dz[i + 1];
      for (int l = 0; l < ns; ++l)
      {
        cmajor[i * ns + l] = -chi[1 + i] / (dz[i] * dz[i + 1]) * inv_lewis[l];
        csub[i * ns + l] = chi[1 + i] / (dzt * dz[i]) * inv_lewis[l];
        csup[i * ns + l] = chi[1 + i] / (dzt * dz[i + 1]) * inv_lewis[l];
      }
      ncoeff[i] = 1 / (dz[i] + dz[i + 1]);
      mcoeff[i] = -ncoeff[i];
    }
    return GB_OK;
  }

  // flamelet_jac_indices, flamelet_kernels.cpp:50-90: COO indices of the BTDDOD storage
  int gb_flamelet_jac_indices(const gb_mech *m, int nzi, int *rows, int *cols)
  {
    if (!m || !rows || !cols || nzi < 1)
    {
      set_error("bad flamelet_jac_indices arguments");
      return GB_ERR_ARG;
    }
    const int ns = (int)m->h.species.size();
    size_t idx = 0;
    for (int iz = 0; iz < nzi; ++iz)
      for (int iq = 0; iq < ns; ++iq)
        for (int jq = 0; jq < ns; ++jq, ++idx)
        {
          rows[idx] = iz * ns + jq;
          cols[idx] = iz * ns + iq;
        }
    for (int iz = 1; iz < nzi; ++iz)
      for (int iq = 0; iq < ns; ++iq, ++idx)
      {
        rows[idx] = iz * ns + iq;
        cols[idx] = iz * ns + iq - ns;
      }
    for (int iz = 0; iz < nzi - 1; ++iz)
      for (int iq = 0; iq < ns; ++iq, ++idx)
      {
        rows[idx] = iz * ns + iq;
        cols[idx] = iz * ns + iq + ns;
      }
    return GB_OK;
  }

  // members (host, may be NULL): evaluate only the flamelets members[0..n_members) of the batch (rows of the others
  // are left untouched); at most 64 of them, else everything is evaluated
  static int flamelet_rhs_members(gb_mech *m, int F, const double *state, const gb_flamelet_params *prm, double *out_rhs,
                                  const int *members, int n_members, void *stream)
  {
    RC(ready(m));
    RC(check_flamelet(F, state, prm, out_rhs));
    if (F == 0)
      return GB_OK;
    ChemArgs a{};
    a.dm = m->h.dm;
    a.mode = MODE_FLAMELET_RHS;
    a.n = F * prm->nzi;
    a.in_state = state;
    a.p = prm->pressure;
    a.out0 = out_rhs;
    RC(flamelet_dev(m, F, state, prm, &a.fl, (cudaStream_t)stream));
    if (members && n_members > 0 && n_members < F && n_members <= 64 && F <= 256)
    {
      a.fl.nmembers = n_members;
      for (int k = 0; k < n_members; ++k)
      {
        if (members[k] < 0 || members[k] >= F)
        {
          set_error("flamelet right-hand side: member index out of range");
          return GB_ERR_ARG;
        }
        a.fl.members[k] = (unsigned char)members[k];
      }
      a.n = n_members * prm->nzi;
    }
    CK(launch_rates(a, (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_flamelet_rhs_batch(gb_mech *m, int F, const double *state, const gb_flamelet_params *prm, double *out_rhs,
                            void *stream)
  {
    return flamelet_rhs_members(m, F, state, prm, out_rhs, nullptr, 0, stream);
  }

  // The Newton loop of one implicit stage for F flamelets, entirely behind the C-ABI: per iteration the inverse-based
  // block-Thomas solve, the update, the flamelet right-hand side and the fused residual / norm / convergence kernel
  // are launched back to back and the host reads one integer. (time/nonlinear.py:185-268 restated for a batch; the
  // host cost of an iteration is four launches and one synchronisation instead of a Python loop body.)
  int gb_flamelet_newton_stage_batch(gb_mech *m, int F, const gb_flamelet_params *prm, const double *d_factors,
                                     const double *l_values, const double *dinv, const int *system_rows,
                                     const double *explicit_, const double *q, const double *dt, double gamma,
                                     const double *weights, double tolerance, int max_iterations, double *x, double *f,
                                     double *res, int *conv, double *work, int *n_unconverged, int *out_iterations,
                                     void *stream)
  {
    RC(ready(m));
    RC(check_flamelet(F, x, prm, f));
    if (!d_factors || !l_values || !dinv || !explicit_ || !q || !dt || !weights || !res || !conv || !work || !n_unconverged)
    {
      set_error("gb_flamelet_newton_stage_batch: null array");
      return GB_ERR_ARG;
    }
    if (out_iterations)
      *out_iterations = 0;
    if (F == 0)
      return 0;
    const int ns = m->h.dm.ns, nzi = prm->nzi, ndof = ns * nzi;
    double *dx = work, *xn = dx + (size_t)F * ndof, *fn = xn + (size_t)F * ndof;
    int left = F;
    for (int it = 0; it < max_iterations && left > 0; ++it)
    {
      RC(gb_btddod_full_solve_inv_batch(F, d_factors, l_values, dinv, res, nzi, ns, dx, system_rows, stream));
      RC(gb_newton_update_batch(F, ndof, x, dx, conv, xn, n_unconverged, stream));
      RC(gb_flamelet_rhs_batch(m, F, xn, prm, fn, stream));
      RC(gb_newton_tail_batch(F, ndof, fn, xn, explicit_, q, dt, gamma, weights, tolerance, x, f, res, conv,
                              n_unconverged, &left, stream));
      if (out_iterations)
        *out_iterations = it + 1;
    }
    return left; // > 0: members that did not converge within max_iterations
  }

  // All implicit stages of one ESDIRK step for F flamelets, the members walking through the stages independently (the
  // stage machine of gb_newton_tail_staged_batch): rounds of {solve_inv, update, flamelet rhs, staged tail} until every
  // member has finished its last stage. On entry stage[m] = first implicit stage (1), iters = nlfail = done = 0, K[0] =
  // f(q), (x, f) = (q, K[0]), explicit / res prepared for stage 1 (gb_esdirk_stage_begin_batch).
  int gb_flamelet_esdirk_stages_batch(gb_mech *m, int F, const gb_flamelet_params *prm, const double *d_factors,
                                      const double *l_values, const double *dinv, const int *system_rows, int nstages,
                                      const double *tableau, const double *q, const double *dt, double gamma,
                                      const double *weights, double tolerance, int max_iterations, double *x, double *f,
                                      double *res, double *explicit_, double *K, int *stage, int *iters, int *nlfail,
                                      int *done, double *work, int *n_left, int *out_rounds, void *stream)
  {
    RC(ready(m));
    RC(check_flamelet(F, x, prm, f));
    if (!d_factors || !l_values || !dinv || !tableau || !q || !dt || !weights || !res || !explicit_ || !K || !stage ||
        !iters || !nlfail || !done || !work || !n_left)
    {
      set_error("gb_flamelet_esdirk_stages_batch: null array");
      return GB_ERR_ARG;
    }
    if (out_rounds)
      *out_rounds = 0;
    if (F == 0)
      return 0;
    const int ns = m->h.dm.ns, nzi = prm->nzi, ndof = ns * nzi;
    double *dx = work, *xn = dx + (size_t)F * ndof, *fn = xn + (size_t)F * ndof;
    int left = F;
    const int max_rounds = (nstages - 1) * max_iterations;
    for (int round = 0; round < max_rounds && left > 0; ++round)
    {
      RC(gb_btddod_full_solve_inv_batch(F, d_factors, l_values, dinv, res, nzi, ns, dx, system_rows, stream));
      RC(gb_newton_update_batch(F, ndof, x, dx, done, xn, n_left, stream));
      RC(gb_flamelet_rhs_batch(m, F, xn, prm, fn, stream));
      RC(gb_newton_tail_staged_batch(F, ndof, nstages, tableau, max_iterations, fn, xn, q, dt, gamma, weights, tolerance,
                                     x, f, res, explicit_, K, stage, iters, nlfail, done, n_left, &left, stream));
      if (out_rounds)
        *out_rounds = round + 1;
    }
    return left; // 0 once every member has finished its last stage
  }

  // One "tick" of the asynchronous batch integrator for F flamelets, everything device-side in one call:
  //   1. members flagged in host_start get state 1 (BEGIN) and their step size host_dt[m];
  //   2. rounds {solve_inv, update, flamelet rhs, tail} for the members in state 1 or 2; after each round the state /
  //      stage arrays are copied to the host; stop when a member has completed its stages (state 0, stage == nstages),
  //      no member is active, or max_rounds rounds were taken;
  //   3. if members completed: embedded error estimate (gb_esdirk_finish_batch), q <- q + dq for those whose update is
  //      finite, and their statistics, Newton-failure flags and new states are copied to the host.
  // Returns the number of rounds taken (>= 0) or a negative error code.
  int gb_flamelet_async_tick_batch(gb_mech *m, int F, const gb_flamelet_params *prm, const double *d_factors,
                                   const double *l_values, const double *dinv, int nstages, const double *tableau,
                                   const double *b, const double *bh, double *q, double *dt, double gamma,
                                   const double *weights, double tolerance, int max_iterations, int clip_negative,
                                   double *x, double *f, double *res, double *explicit_, double *K, int *state, int *stage,
                                   int *iters, int *nlfail, int *newton_its, double *work, double *dq, double *stats,
                                   int *start_d, double *dtin_d, int max_rounds, const int *host_start,
                                   const double *host_dt, int *host_state, int *host_stage, double *host_stats,
                                   int *host_nlfail, double *host_q, const int *host_members, int n_members,
                                   void *stream)
  {
    RC(ready(m));
    RC(check_flamelet(F, x, prm, f));
    if (!d_factors || !l_values || !dinv || !tableau || !b || !bh || !q || !dt || !weights || !res || !explicit_ || !K ||
        !state || !stage || !iters || !nlfail || !newton_its || !work || !dq || !stats || !start_d || !dtin_d ||
        !host_state || !host_stage || !host_stats || !host_nlfail || !host_q)
    {
      set_error("gb_flamelet_async_tick_batch: null array");
      return GB_ERR_ARG;
    }
    if (F == 0)
      return 0;
    const int ns = m->h.dm.ns, nzi = prm->nzi, ndof = ns * nzi;
    double *dx = work, *xn = dx + (size_t)F * ndof, *fn = xn + (size_t)F * ndof;
    cudaStream_t st = (cudaStream_t)stream;
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    static const bool tick_profile = getenv("GB_TICK_PROFILE") != nullptr;
    const clk::time_point tick_t0 = clk::now();
    if (host_start)
    {
      bool any = false;
      for (int k = 0; k < F; ++k)
        any = any || host_start[k] != 0;
      if (any)
      {
        if (!host_dt)
        {
          set_error("gb_flamelet_async_tick_batch: host_start needs host_dt");
          return GB_ERR_ARG;
        }
        CK(cudaMemcpyAsync(start_d, host_start, sizeof(int) * F, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dtin_d, host_dt, sizeof(double) * F, cudaMemcpyHostToDevice, st));
        RC(gb_async_round_kernels(F, ndof, nstages, tableau, 0, 2, reinterpret_cast<const double *>(start_d), nullptr,
                                  dtin_d, q, dt, gamma, weights, tolerance, x, f, res, explicit_, K, state, stage, iters,
                                  nlfail, newton_its, stream));
      }
    }
    int rounds = 0;
    bool any_complete = false;
    // Opt-in (GB_TICK_PUBLISH=1), measured not to pay: larger batches complete a member's step in nearly every round, so
    // the error estimate and the acceptance kernel can run at the end of EVERY round, that kernel writing the round's
    // outcome (states, stages, and the statistics and new state rows of the completed members) straight into the mapped
    // pinned host arrays -- one synchronisation per round and no copy operations instead of two synchronisations and
    // 4 + k copies. 56 GRI-128 trajectories: 0.417 ms per tick against 0.422 ms (the 53 KB rows written over PCIe by
    // the kernel cost what the second synchronisation did); bit-identical results.
    static const bool publish_enabled = getenv("GB_TICK_PUBLISH") && atoi(getenv("GB_TICK_PUBLISH")) != 0;
    int *hd_state = nullptr, *hd_stage = nullptr, *hd_nlfail = nullptr;
    double *hd_stats = nullptr, *hd_q = nullptr;
    bool publish = publish_enabled && !tick_profile && F >= 16;
    if (publish)
    {
      publish = cudaHostGetDevicePointer((void **)&hd_state, host_state, 0) == cudaSuccess &&
                cudaHostGetDevicePointer((void **)&hd_stage, host_stage, 0) == cudaSuccess &&
                cudaHostGetDevicePointer((void **)&hd_stats, host_stats, 0) == cudaSuccess &&
                cudaHostGetDevicePointer((void **)&hd_nlfail, host_nlfail, 0) == cudaSuccess &&
                cudaHostGetDevicePointer((void **)&hd_q, host_q, 0) == cudaSuccess;
      if (!publish)
        cudaGetLastError(); // (host arrays that are not mapped pinned memory: the two-step form)
    }
    const double *kp_all[6];
    for (int j = 0; j < nstages && j < 6; ++j)
      kp_all[j] = K + (size_t)j * F * ndof;
    const clk::time_point loop_t0 = clk::now();
    while (rounds < max_rounds)
    {
      const clk::time_point r0 = clk::now();
      RC(gb_btddod_full_solve_inv_batch(F, d_factors, l_values, dinv, res, nzi, ns, dx, nullptr, stream));
      if (tick_profile)
      {
        CK(cudaStreamSynchronize(st));
        g_tick_stats[4] += secs(r0, clk::now());
      }
      RC(gb_async_round_kernels(F, ndof, nstages, tableau, max_iterations, 0, fn, xn, dx, q, dt, gamma, weights, tolerance, x,
                                f, res, explicit_, K, state, stage, iters, nlfail, newton_its, stream));
      const clk::time_point r1 = clk::now();
      RC(flamelet_rhs_members(m, F, xn, prm, fn, host_members, n_members, stream));
      if (tick_profile)
      {
        CK(cudaStreamSynchronize(st));
        g_tick_stats[5] += secs(r1, clk::now());
      }
      RC(gb_async_round_kernels(F, ndof, nstages, tableau, max_iterations, 1, fn, xn, dx, q, dt, gamma, weights, tolerance, x,
                                f, res, explicit_, K, state, stage, iters, nlfail, newton_its, stream));
      if (publish)
      {
        RC(gb_esdirk_finish_batch(F, ndof, nstages, kp_all, b, bh, dt, weights, dq, stats, stream));
        RC(async_accept_publish(F, ndof, nstages, dq, stats, clip_negative, state, stage, q, nlfail, hd_state, hd_stage,
                                hd_stats, hd_nlfail, hd_q, st));
      }
      else
      {
        CK(cudaMemcpyAsync(host_state, state, sizeof(int) * F, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(host_stage, stage, sizeof(int) * F, cudaMemcpyDeviceToHost, st));
      }
      CK(cudaStreamSynchronize(st));
      ++rounds;
      bool any_active = false;
      for (int k = 0; k < F; ++k)
      {
        any_active = any_active || host_state[k] != 0;
        any_complete = any_complete || (host_state[k] == 0 && host_stage[k] == nstages);
      }
      if (any_complete || !any_active)
        break;
    }
    g_tick_stats[3] += secs(loop_t0, clk::now());
    g_tick_stats[1] += rounds;
    struct TickEnd
    {
      clk::time_point t0;
      ~TickEnd()
      {
        g_tick_stats[0] += 1.;
        g_tick_stats[2] += std::chrono::duration<double>(clk::now() - t0).count();
      }
    } tick_end{tick_t0};
    if (any_complete && !publish)
    {
      const double *kp[6];
      for (int j = 0; j < nstages; ++j)
        kp[j] = K + (size_t)j * F * ndof;
      RC(gb_esdirk_finish_batch(F, ndof, nstages, kp, b, bh, dt, weights, dq, stats, stream));
      RC(gb_async_round_kernels(F, ndof, nstages, tableau, clip_negative, 3, dq, nullptr, stats, q, dt, gamma, weights,
                                tolerance, x, f, res, explicit_, K, state, stage, iters, nlfail, newton_its, stream));
      CK(cudaMemcpyAsync(host_stats, stats, sizeof(double) * 3 * F, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(host_nlfail, nlfail, sizeof(int) * F, cudaMemcpyDeviceToHost, st));
      for (int k = 0; k < F; ++k)
        if (host_state[k] == 0 && host_stage[k] == nstages)
          CK(cudaMemcpyAsync(host_q + (size_t)k * ndof, q + (size_t)k * ndof, sizeof(double) * ndof,
                             cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    return rounds;
  }

  int gb_flamelet_jacobian_batch(gb_mech *m, int F, const double *state, const gb_flamelet_params *prm,
                                 int compute_eigenvalues, double diffterm, int scale_and_offset, double prefactor,
                                 int rates_sensitivity_option, int sensitivity_transform_option, double *out_expeig,
                                 double *out_jac, void *stream)
  {
    RC(ready(m));
    RC(check_flamelet(F, state, prm, out_jac));
    if (compute_eigenvalues && !out_expeig)
    {
      set_error("compute_eigenvalues needs out_expeig");
      return GB_ERR_ARG;
    }
    if (rates_sensitivity_option < 0 || rates_sensitivity_option > 2 || sensitivity_transform_option != 0)
    {
      set_error("rates_sensitivity_option must be 0..2 and sensitivity_transform_option 0");
      return GB_ERR_ARG;
    }
    if (F == 0)
      return GB_OK;
    ChemArgs a{};
    a.dm = m->h.dm;
    a.mode = MODE_FLAMELET_JAC;
    a.n = F * prm->nzi;
    a.in_state = state;
    a.p = prm->pressure;
    a.out1 = out_jac;
    RC(flamelet_dev(m, F, state, prm, &a.fl, (cudaStream_t)stream));
    a.fl.scale_and_offset = scale_and_offset ? 1 : 0;
    a.fl.prefactor = prefactor;
    if (compute_eigenvalues)
    { // flamelet_kernels.cpp:1329-1341: the eigenvalues are those of the transformed chemical block as it stands before
      // the diffusion diagonal, the enthalpy-flux correction and the scaling are applied; a first pass writes exactly
      // those blocks (densely) into scratch memory, one warp per block reduces each to its largest real part
      const int ns = m->h.dm.ns, nblocks = F * prm->nzi;
      void *d_blocks, *d_maxre;
      RC(scratch(m, 7, sizeof(double) * ((size_t)nblocks * ns * ns + nblocks), &d_blocks));
      d_maxre = (double *)d_blocks + (size_t)nblocks * ns * ns;
      ChemArgs e = a;
      e.out1 = (double *)d_blocks;
      e.fl.chem_only = 1;
      e.fl.scale_and_offset = 0;
      CK(launch_jac(e, (cudaStream_t)stream));
      CK(launch_block_max_real_eig(nblocks, (const double *)d_blocks, (long)prm->nzi * ns * ns, prm->nzi, ns,
                                   (double *)d_maxre, (cudaStream_t)stream));
      CK(launch_expand_expeig(nblocks, ns, (const double *)d_maxre, diffterm, out_expeig, (cudaStream_t)stream));
    }
    CK(launch_jac(a, (cudaStream_t)stream));
    CK(launch_flamelet_offdiag(m->h.dm, F, a.fl, out_jac, (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_max_real_eigenvalue_batch(int nblocks, int n, const double *blocks, double *out, void *stream)
  {
    if (nblocks < 0 || n < 1 || (nblocks > 0 && (!blocks || !out)))
    {
      set_error("gb_max_real_eigenvalue_batch: bad arguments");
      return GB_ERR_ARG;
    }
    if (sizeof(double) * ((size_t)n * (n | 1) + n + 2) > (size_t)227 * 1024)
    {
      set_error("gb_max_real_eigenvalue_batch: matrix does not fit in shared memory (n <= 169)");
      return GB_ERR_UNSUPPORTED;
    }
    CK(launch_block_max_real_eig(nblocks, blocks, (long)nblocks * n * n, nblocks > 0 ? nblocks : 1, n, out,
                                 (cudaStream_t)stream));
    return GB_OK;
  }

  int gb_flamelet_rhs_host(gb_mech *m, int F, const double *state, const gb_flamelet_params *prm, double *out_rhs)
  {
    RC(ready(m));
    RC(check_flamelet(F, state, prm, out_rhs));
    if (F == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    const size_t nv = (size_t)F * prm->nzi * ns;
    FlameletStage st;
    RC(st.stage(ns, F, prm));
    void *d_state, *d_rhs;
    RC(scratch(m, 0, sizeof(double) * nv, &d_state));
    RC(scratch(m, 2, sizeof(double) * nv, &d_rhs));
    CK(cudaMemcpy(d_state, state, sizeof(double) * nv, cudaMemcpyHostToDevice));
    RC(gb_flamelet_rhs_batch(m, F, (double *)d_state, &st.dp, (double *)d_rhs, nullptr));
    const int bad = gb_count_nonfinite_members_batch(F, (long)prm->nzi * ns, (const double *)d_rhs, 0, nullptr, nullptr,
                                                     nullptr);
    CK(cudaMemcpy(out_rhs, d_rhs, sizeof(double) * nv, cudaMemcpyDeviceToHost));
    return bad; // > 0: flamelets whose right-hand side holds an Inf or NaN
  }

  int gb_flamelet_jacobian_host(gb_mech *m, int F, const double *state, const gb_flamelet_params *prm,
                                int compute_eigenvalues, double diffterm, int scale_and_offset, double prefactor,
                                int rates_sensitivity_option, int sensitivity_transform_option, double *out_expeig,
                                double *out_jac)
  {
    RC(ready(m));
    RC(check_flamelet(F, state, prm, out_jac));
    if (F == 0)
      return GB_OK;
    const int ns = m->h.dm.ns;
    const size_t nv = (size_t)F * prm->nzi * ns;
    const size_t nj = (size_t)F * ns * ((size_t)prm->nzi * ns + 2 * (prm->nzi - 1));
    FlameletStage st;
    RC(st.stage(ns, F, prm));
    void *d_state, *d_jac, *d_eig;
    RC(scratch(m, 0, sizeof(double) * nv, &d_state));
    RC(scratch(m, 3, sizeof(double) * nj, &d_jac));
    RC(scratch(m, 2, sizeof(double) * nv, &d_eig));
    CK(cudaMemcpy(d_state, state, sizeof(double) * nv, cudaMemcpyHostToDevice));
    RC(gb_flamelet_jacobian_batch(m, F, (double *)d_state, &st.dp, compute_eigenvalues, diffterm, scale_and_offset,
                                  prefactor, rates_sensitivity_option, sensitivity_transform_option, (double *)d_eig,
                                  (double *)d_jac, nullptr));
    const int bad = gb_count_nonfinite_members_batch(F, (long)(nj / F), (const double *)d_jac, 0, nullptr, nullptr, nullptr);
    CK(cudaMemcpy(out_jac, d_jac, sizeof(double) * nj, cudaMemcpyDeviceToHost));
    if (compute_eigenvalues && out_expeig)
      CK(cudaMemcpy(out_expeig, d_eig, sizeof(double) * nv, cudaMemcpyDeviceToHost));
    return bad; // > 0: flamelets whose Jacobian holds an Inf or NaN
  }
}
