// C ABI of libeftb200: plan management, stage entry points, the fused per-batch pipeline.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

static thread_local char g_error[512] = "";

void eftb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

namespace {

inline bool c_has_ap(const eftb_plan* p) { return p->cfg.has_ap != 0; }

template <typename T>
int upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0 || !src) return EFTB_OK;
  EFTB_CUDA_CHECK(cudaMalloc((void**)dst, n * sizeof(T)));
  EFTB_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return EFTB_OK;
}

struct Sizes {
  size_t u, F, D, P22, Dg, T, Cr, scal, out, ap, rs, Dcf;
};

Sizes sizes(const eftb_plan* p, int Bp) {
  const eftb_config& c = p->cfg;
  Sizes z;
  z.u = (size_t)p->K * Bp;
  z.F = (size_t)c.front_rows * Bp;
  z.D = (size_t)EFTB_NCH * (c.Nmax + 1) * 2 * Bp;
  z.P22 = (size_t)EFTB_N22 * c.Nk * Bp;
  z.Dg = (size_t)c.Nl * 12 * (c.Nmax + 1) * 2 * Bp;
  z.T = (size_t)c.Nl * c.Nk * c.nterm * Bp;
  z.Cr = (size_t)c.Nl * (14 + (c.with_nnlo ? 1 : 0)) * c.Ns * Bp;
  z.scal = (size_t)3 * Bp;
  z.out = c.has_project ? (size_t)c.nout * c.nterm * Bp : 0;
  z.ap = c.has_ap ? ap_scratch_doubles(p, Bp) : 0;
  z.rs = c.has_resum ? resum_scratch_doubles(p, Bp) : 0;
  z.Dcf = c.row_cre_cf >= 0 ? z.D : 0;
  return z;
}

__global__ void probe_fp64_kernel(double* sink, int iters) {
  // 16 independent DFMA chains per thread
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double m = 1.0 + 1e-12, c = 1e-13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678) sink[0] = s;
}

}  // namespace

unsigned long long g_eftb_launches = 0;

int eftb_current_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { eftb_set_error("cudaGetDevice -> %s", cudaGetErrorString(e)); return -1; }
  return dev;
}

int eftb_sm_count() {
  static int sms[EFTB_MAX_DEVICES] = {};
  const int dev = eftb_current_device();
  if (dev < 0) return 0;
  if (dev < EFTB_MAX_DEVICES && __atomic_load_n(&sms[dev], __ATOMIC_ACQUIRE)) return sms[dev];
  int n = 0;
  cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { eftb_set_error("cudaDeviceGetAttribute -> %s", cudaGetErrorString(e)); return 0; }
  if (dev < EFTB_MAX_DEVICES) __atomic_store_n(&sms[dev], n, __ATOMIC_RELEASE);
  return n;
}

extern "C" {

int eftb_abi_version(void) { return EFTB_ABI_VERSION; }
unsigned long long eftb_launch_count(void) { return __atomic_load_n(&g_eftb_launches, __ATOMIC_RELAXED); }
const char* eftb_last_error(void) { return g_error; }
int eftb_padded_batch(int B) { return B < 1 ? 0 : eftb_round_up(B, 32); }

int eftb_probe_fp64(int iters, double* tflops, void* stream) {
  if (!tflops || iters < 1) { eftb_set_error("eftb_probe_fp64: bad argument"); return EFTB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const int sms = eftb_sm_count();
  if (!sms) return EFTB_ERR_CUDA;
  double* sink = nullptr;
  EFTB_CUDA_CHECK(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  EFTB_CUDA_CHECK(cudaEventCreate(&e0));
  EFTB_CUDA_CHECK(cudaEventCreate(&e1));
  const int blocks = sms * 4, threads = 256;
  probe_fp64_kernel<<<blocks, threads, 0, s>>>(sink, iters / 8 + 1);  // warm-up
  EFTB_CUDA_CHECK(cudaEventRecord(e0, s));
  probe_fp64_kernel<<<blocks, threads, 0, s>>>(sink, iters);
  EFTB_CUDA_CHECK(cudaEventRecord(e1, s));
  EFTB_CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  EFTB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = 2.0 * 16.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return EFTB_OK;
}

int eftb_plan_create(const eftb_config* cfg, const eftb_constants* h, eftb_plan** out) {
  if (!cfg || !h || !out) { eftb_set_error("eftb_plan_create: NULL argument"); return EFTB_ERR_ARG; }
  const eftb_config& c = *cfg;
  if (c.Nl < 2 || c.Nl > 3 || c.Nk < 8 || c.Ns < 4 || c.Nmax < 8 || (c.Nmax & 1) || c.nterm < 24 || c.nin < 4) {
    eftb_set_error("eftb_plan_create: unsupported sizes Nl=%d Nk=%d Ns=%d Nmax=%d nterm=%d", c.Nl, c.Nk, c.Ns, c.Nmax, c.nterm);
    return EFTB_ERR_ARG;
  }
  if (c.nterm != 24 + (c.with_nnlo ? 3 : 0)) { eftb_set_error("eftb_plan_create: nterm inconsistent with with_nnlo"); return EFTB_ERR_ARG; }
  if ((c.row_cre_cf >= 0) != (c.row_cim_cf >= 0) || c.row_cre_cf + c.Nmax / 2 + 1 > c.front_rows || c.row_cim_cf + c.Nmax / 2 + 1 > c.front_rows) {
    eftb_set_error("eftb_plan_create: inconsistent rows of the second coefficient set");
    return EFTB_ERR_ARG;
  }
  eftb_plan* p = new eftb_plan();
  p->cfg = c;
  p->K = c.nin + c.ntail + c.ntailx;
  int rc = 0;
  rc |= upload(&p->k, h->k, c.Nk);
  rc |= upload(&p->l11, h->l11, (size_t)c.Nl * 3);
  rc |= upload(&p->lct, h->lct, (size_t)c.Nl * 6);
  rc |= upload(&p->lctnnlo, h->lctnnlo, (size_t)c.Nl * 3);
  rc |= upload(&p->l22, h->l22, (size_t)c.Nl * EFTB_N22);
  rc |= upload(&p->l13, h->l13, (size_t)c.Nl * EFTB_N13);
  rc |= upload(&p->lr, h->lr, c.ntail);
  rc |= upload(&p->lrx, h->lrx, c.ntailx);
  rc |= gemm_upload(h->Wf, 1, c.front_rows, p->K, &p->Wf);
  if (!h->pair_table || !h->pair_offsets) { eftb_set_error("eftb_plan_create: pair table missing"); rc = EFTB_ERR_ARG; }
  else rc |= antidiag_pack(p, h->pair_table, h->pair_offsets);
  rc |= gemm_upload(h->Ak, 1, c.Nk, 2 * (c.Nmax + 1), &p->Ak);
  rc |= gemm_upload(h->As, c.Nl, c.Ns, 2 * (c.Nmax + 1), &p->As);
  if (c.has_resum) {
    if (!h->R || !h->q) { eftb_set_error("eftb_plan_create: resummation constants missing"); rc = EFTB_ERR_ARG; }
    else rc |= resum_pack(p, h->R, h->q);
    rc |= upload(&p->kr2, h->kr2, c.Nkr);
  }
  if (c.has_ap) {
    rc |= gemm_upload(h->Cinv, 1, c.Nk, c.Nk, &p->Cinv);
    rc |= upload(&p->knot_lo, h->knot_lo, c.nint);
    rc |= upload(&p->basis, h->basis, (size_t)c.nint * 16);
    rc |= upload(&p->mu, h->mu, c.nmu);
    rc |= upload(&p->wl, h->wl, (size_t)c.Nl * c.nmu);
  }
  int nl_out = c.Nl, nk_out = c.Nk;
  if (c.has_project) {
    rc |= gemm_upload(h->project, 1, c.nout, c.Nl * c.Nk, &p->project);
    if (h->project_st) rc |= gemm_upload(h->project_st, 1, c.nout, c.Nl * c.Nk, &p->project_st);
    nl_out = c.nl_out;
    nk_out = c.nl_out > 0 ? c.nout / c.nl_out : 0;
    if (nl_out < 1 || nl_out * nk_out != c.nout) { eftb_set_error("eftb_plan_create: nout != nl_out*nk_out"); rc = EFTB_ERR_ARG; }
  }
  if (!rc) {
    // point-major export order [l][term][k]  <-  batch-minor row (l*nk + k)*nterm + term
    std::vector<int32_t> perm((size_t)nl_out * c.nterm * nk_out);
    for (int l = 0; l < nl_out; ++l)
      for (int i = 0; i < c.nterm; ++i)
        for (int k = 0; k < nk_out; ++k) perm[((size_t)l * c.nterm + i) * nk_out + k] = (l * nk_out + k) * c.nterm + i;
    p->perm_rows = (int)perm.size();
    rc |= upload(&p->perm_out, perm.data(), perm.size());
  }
  if (!rc) {
    cudaError_t e = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_loops, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) { eftb_set_error("eftb_plan_create: %s", cudaGetErrorString(e)); rc = EFTB_ERR_CUDA; }
  }
  if (rc) { eftb_plan_destroy(p); return rc < 0 ? rc : EFTB_ERR_CUDA; }
  *out = p;
  return EFTB_OK;
}

void eftb_plan_destroy(eftb_plan* p) {
  if (!p) return;
  void* ptrs[] = {p->k, p->l11, p->lct, p->lctnnlo, p->l22, p->l13, p->lr, p->lrx,
                  p->rs.Rt, p->rs.Rk, p->rs.qpack, p->kr2, p->knot_lo, p->basis, p->mu, p->wl, p->perm_out};
  for (void* q : ptrs) if (q) cudaFree(q);
  antidiag_free(p);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_loops) cudaEventDestroy(p->ev_loops);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  if (p->side) cudaStreamDestroy(p->side);
  gemm_free(&p->Wf); gemm_free(&p->Ak); gemm_free(&p->As); gemm_free(&p->Cinv); gemm_free(&p->project); gemm_free(&p->project_st);
  delete p;
}

size_t eftb_workspace_bytes(const eftb_plan* p, int B) {
  if (!p || B < 1) return 0;
  Sizes z = sizes(p, eftb_padded_batch(B));
  // u | F | D (later reused for the spline coefficients and the AP output) | P22 | Dg | T | Cr | f,DA,H | out | AP operator
  const size_t apreg = c_has_ap(p) ? ap_coef_doubles(p, eftb_padded_batch(B)) + z.T : 0;  // coefficients | AP output
  size_t dreg = z.D > apreg ? z.D : apreg;
  return (z.u + z.F + dreg + z.P22 + z.Dg + z.T + z.Cr + z.scal + z.out + z.ap + z.rs + z.Dcf) * sizeof(double);
}

int eftb_to_batch_minor(const double* in, int B, int R, double* out, void* stream) {
  if (!in || !out || B < 1 || R < 1) { eftb_set_error("eftb_to_batch_minor: bad argument"); return EFTB_ERR_ARG; }
  return launch_to_batch_minor(in, B, eftb_padded_batch(B), R, out, (cudaStream_t)stream);
}

int eftb_to_point_major(const double* in, int B, int R, const int32_t* perm, double* out, void* stream) {
  if (!in || !out || B < 1 || R < 1) { eftb_set_error("eftb_to_point_major: bad argument"); return EFTB_ERR_ARG; }
  return launch_to_point_major(in, B, eftb_padded_batch(B), R, perm, out, (cudaStream_t)stream);
}

#define EFTB_NEED(cond, what)                                   \
  if (!(cond)) {                                                \
    eftb_set_error("%s: %s", __func__, what);                   \
    return EFTB_ERR_ARG;                                        \
  }

int eftb_front(const eftb_plan* p, int B, const double* plin, double* u, double* F, void* stream) {
  EFTB_NEED(p && plin && u && F && B >= 1, "NULL/invalid argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int Bp = eftb_padded_batch(B);
  int rc = launch_front_prepare(p, B, Bp, plin, u, s);
  if (rc) return rc;
  return gemm_run(p->Wf, u, F, Bp, 1, 1, 0, 0, 0, 0, s);
}

int eftb_antidiag(const eftb_plan* p, int B, const double* F, double* D, void* stream) {
  EFTB_NEED(p && F && D && B >= 1, "NULL/invalid argument");
  return launch_antidiag(p, eftb_padded_batch(B), F, D, (cudaStream_t)stream);
}

int eftb_antidiag_cf(const eftb_plan* p, int B, const double* F, double* D, void* stream) {
  EFTB_NEED(p && F && D && B >= 1, "NULL/invalid argument");
  return launch_antidiag(p, eftb_padded_batch(B), F, D, (cudaStream_t)stream, true);
}

int eftb_has_cf_set(const eftb_plan* p) { return p && p->cfg.row_cre_cf >= 0; }

int eftb_spectral(const eftb_plan* p, int B, const double* D, const double* Dcf, double* P22, double* Cs, void* stream) {
  EFTB_NEED(p && D && P22 && Cs && B >= 1, "NULL/invalid argument");
  if (!Dcf) Dcf = D;
  cudaStream_t s = (cudaStream_t)stream;
  const eftb_config& c = p->cfg;
  const int Bp = eftb_padded_batch(B);
  const size_t dch = (size_t)(c.Nmax + 1) * 2 * Bp;
  int rc = gemm_run(p->Ak, D, P22, Bp, EFTB_N22, EFTB_N22, dch, 0, (size_t)c.Nk * Bp, 0, s);
  if (rc) return rc;
  return gemm_run(p->As, Dcf, Cs, Bp, c.Nl * EFTB_NCH, EFTB_NCH, dch, 0, (size_t)c.Ns * Bp, (size_t)EFTB_NCH * c.Ns * Bp, s);
}

// configuration-space half of eftb_spectral_grouped: regroup the channels, then Cloopl[l][r] = As[l] @ Dg[l][r], written
// straight into rows 2..13 of the point-major Cr[b][l][ncr][Ns]
static int spectral_cf(const eftb_plan* p, int Bp, const double* Dcf, const double* f, double* Dg, double* Cr, cudaStream_t s) {
  const eftb_config& c = p->cfg;
  const size_t dch = (size_t)(c.Nmax + 1) * 2 * Bp;
  const int ncr = 14 + (c.with_nnlo ? 1 : 0);
  int rc = launch_regroup(p, Bp, Dcf, f, Dg, s);
  if (rc) return rc;
  GemmPointMajor pm{Bp, (size_t)c.Nl * ncr * c.Ns, 0};
  return gemm_run(p->As, Dg, Cr + 2 * (size_t)c.Ns, Bp, c.Nl * 12, 12, dch, 12 * dch, (size_t)c.Ns, (size_t)ncr * c.Ns, s, &pm);
}

int eftb_spectral_grouped(const eftb_plan* p, int B, const double* D, const double* Dcf, const double* f, double* Dg, double* P22,
                          double* Cr, void* stream) {
  EFTB_NEED(p && D && f && Dg && P22 && Cr && B >= 1, "NULL/invalid argument");
  if (!Dcf) Dcf = D;
  cudaStream_t s = (cudaStream_t)stream;
  const eftb_config& c = p->cfg;
  const int Bp = eftb_padded_batch(B);
  const size_t dch = (size_t)(c.Nmax + 1) * 2 * Bp;
  int rc = gemm_run(p->Ak, D, P22, Bp, EFTB_N22, EFTB_N22, dch, 0, (size_t)c.Nk * Bp, 0, s);
  if (rc) return rc;
  return spectral_cf(p, Bp, Dcf, f, Dg, Cr, s);
}

int eftb_group(const eftb_plan* p, int B, const double* F, const double* P22, const double* Cs, const double* f, double* T,
               double* Cr, void* stream) {
  EFTB_NEED(p && F && P22 && f && T && Cr && B >= 1, "NULL/invalid argument");
  return launch_group(p, eftb_padded_batch(B), F, P22, Cs, f, T, Cr, (cudaStream_t)stream);
}

size_t eftb_resum_scratch_bytes(const eftb_plan* p, int B) {
  if (!p || B < 1 || !p->cfg.has_resum) return 0;
  return resum_scratch_doubles(p, eftb_padded_batch(B)) * sizeof(double);
}

int eftb_resum(const eftb_plan* p, int B, const double* F, const double* Cr, const double* f, double* T, double* scratch,
               void* stream) {
  EFTB_NEED(p && F && Cr && f && T && scratch && B >= 1, "NULL/invalid argument");
  if (!p->cfg.has_resum) { eftb_set_error("eftb_resum: plan built without IR resummation"); return EFTB_ERR_NOT_BUILT; }
  return launch_resum(p, B, eftb_padded_batch(B), F, Cr, f, T, scratch, (cudaStream_t)stream);
}

size_t eftb_ap_scratch_bytes(const eftb_plan* p, int B) {
  if (!p || B < 1 || !p->cfg.has_ap) return 0;
  const int Bp = eftb_padded_batch(B);
  return (ap_coef_doubles(p, Bp) + ap_scratch_doubles(p, Bp)) * sizeof(double);
}

static int ap_stage(const eftb_plan* p, int B, const double* Tin, const double* DA, const double* H, double* coef,
                    double* gscratch, double* Tout, cudaStream_t s, int phase = EFTB_PHASE_ALL) {
  const eftb_config& c = p->cfg;
  const int Bp = eftb_padded_batch(B);
  const size_t per_l = (size_t)c.Nk * c.nterm * Bp;
  // B-spline coefficients of every term row: coef[l] = Cinv @ T[l]  (N = nterm*Bp columns), stored point-major
  // [b][l][term][j] for the one-CTA-per-cosmology apply kernel
  GemmPointMajor pm{Bp, ap_coef_doubles(p, 1), (size_t)c.Nk};
  int rc = gemm_run(p->Cinv, Tin, coef, c.nterm * Bp, c.Nl, c.Nl, per_l, 0, (size_t)c.nterm * c.Nk, 0, s, &pm);
  if (rc) return rc;
  if (Bp > B) {  // pad lanes are not processed by the per-cosmology kernels: keep them finite
    EFTB_CUDA_CHECK(cudaMemcpyAsync(Tout, Tin, (size_t)c.Nl * per_l * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  return launch_ap(p, B, Bp, coef, Tin, DA, H, gscratch, Tout, s, phase);
}

int eftb_ap(const eftb_plan* p, int B, const double* Tin, const double* DA, const double* H, double* scratch, double* Tout,
            void* stream) {
  EFTB_NEED(p && Tin && DA && H && scratch && Tout && B >= 1 && Tin != Tout, "NULL/invalid argument");
  if (!p->cfg.has_ap) { eftb_set_error("eftb_ap: plan built without AP"); return EFTB_ERR_NOT_BUILT; }
  const eftb_config& c = p->cfg;
  return ap_stage(p, B, Tin, DA, H, scratch, scratch + ap_coef_doubles(p, eftb_padded_batch(B)), Tout, (cudaStream_t)stream);
}

int eftb_project(const eftb_plan* p, int B, const double* T, double* out, void* stream) {
  EFTB_NEED(p && T && out && B >= 1, "NULL/invalid argument");
  if (!p->cfg.has_project) { eftb_set_error("eftb_project: plan built without projection"); return EFTB_ERR_NOT_BUILT; }
  const int Bp = eftb_padded_batch(B);
  const size_t ld = (size_t)p->cfg.nterm * Bp;
  int rc = gemm_run(p->project, T, out, p->cfg.nterm * Bp, 1, 1, 0, 0, 0, 0, (cudaStream_t)stream);
  if (rc || !p->project_st.d) return rc;
  // the stochastic rows (terms 21..23) see their own operator (window_st = False, fiberst = False): a column window
  return gemm_run(p->project_st, T + (size_t)21 * Bp, out + (size_t)21 * Bp, 3 * Bp, 1, 1, 0, 0, 0, 0, (cudaStream_t)stream, nullptr,
                  ld, ld);
}

int eftb_eval_terms(const eftb_plan* p, int B, const double* plin, const double* f, const double* DA, const double* H,
                    double* terms_bm, double* terms_pm, void* workspace, size_t workspace_bytes, void* stream) {
  EFTB_NEED(p && plin && f && workspace && B >= 1, "NULL/invalid argument");
  const eftb_config& c = p->cfg;
  EFTB_NEED(!c.has_ap || (DA && H), "plan has AP but DA/H missing");
  if (workspace_bytes < eftb_workspace_bytes(p, B)) { eftb_set_error("eftb_eval_terms: workspace too small"); return EFTB_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int Bp = eftb_padded_batch(B);
  Sizes z = sizes(p, Bp);
  double* w = (double*)workspace;
  double* u = w;            w += z.u;
  double* F = w;            w += z.F;
  const size_t ncoef = c.has_ap ? ap_coef_doubles(p, Bp) : 0;
  double* D = w;            w += (z.D > ncoef + z.T ? z.D : ncoef + z.T);
  double* P22 = w;          w += z.P22;
  double* Dg = w;           w += z.Dg;
  double* T = w;            w += z.T;
  double* Cr = w;           w += z.Cr;
  double* scal = w;         w += z.scal;
  double* out = w;
  double* Dcf = z.Dcf ? out + z.out + z.ap + z.rs : nullptr;
  int rc;
  // Side stream (fork/join by events; capturable into a CUDA graph together with `stream`): everything that depends on
  // the cosmology scalars only - the Q(f) expansion of the resummation and the AP resampling geometry - runs beside the
  // front end and the loop kernels, and the D -> P22(k) GEMM runs beside the regrouping of the configuration-space
  // channels.  All of them write buffers nothing else touches until the join.
  static const int side_mask = getenv("EFTB_SIDE") ? atoi(getenv("EFTB_SIDE")) : 3;  // tuning knob: 1 = Q(f) + AP geometry, 2 = P22 GEMM
  const bool split_ap = c.has_ap && ap_chunk_count(p, B) == 1;
  cudaStream_t s2 = p->side;
  cudaStream_t sq = (side_mask & 1) ? s2 : s, sg = (side_mask & 2) ? s2 : s;
  double* qf = out + z.out + z.ap;
  if ((rc = launch_scalars_to_batch_minor(f, c.has_ap ? DA : nullptr, c.has_ap ? H : nullptr, B, Bp, scal, s))) return rc;
  EFTB_CUDA_CHECK(cudaEventRecord(p->ev_fork, s));
  EFTB_CUDA_CHECK(cudaStreamWaitEvent(s2, p->ev_fork, 0));
  if (c.has_resum && (rc = launch_resum(p, B, Bp, F, Cr, scal, T, qf, sq, EFTB_PHASE_FIRST))) return rc;
  if (split_ap && (rc = launch_ap(p, B, Bp, nullptr, nullptr, scal + Bp, scal + 2 * (size_t)Bp, out + z.out, nullptr, sq,
                                  EFTB_PHASE_FIRST))) return rc;
  if ((rc = eftb_front(p, B, plin, u, F, stream))) return rc;
  if ((rc = launch_antidiag(p, Bp, F, D, s))) return rc;
  EFTB_CUDA_CHECK(cudaEventRecord(p->ev_loops, s));
  EFTB_CUDA_CHECK(cudaStreamWaitEvent(s2, p->ev_loops, 0));
  {  // D -> P22(k) on the side stream
    const size_t dch = (size_t)(c.Nmax + 1) * 2 * Bp;
    if ((rc = gemm_run(p->Ak, D, P22, Bp, EFTB_N22, EFTB_N22, dch, 0, (size_t)c.Nk * Bp, 0, sg))) return rc;
  }
  EFTB_CUDA_CHECK(cudaEventRecord(p->ev_join, s2));
  if (Dcf && (rc = launch_antidiag(p, Bp, F, Dcf, s, true))) return rc;
  if ((rc = spectral_cf(p, Bp, Dcf ? Dcf : D, scal, Dg, Cr, s))) return rc;
  EFTB_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_join, 0));
  if ((rc = launch_group(p, Bp, F, P22, nullptr, scal, T, Cr, s))) return rc;
  if (c.has_resum && (rc = launch_resum(p, B, Bp, F, Cr, scal, T, qf, s, EFTB_PHASE_SECOND))) return rc;
  double* cur = T;
  if (c.has_ap) {
    double* coef = D;
    double* T2 = D + ncoef;
    if ((rc = ap_stage(p, B, T, scal + Bp, scal + 2 * (size_t)Bp, coef, out + z.out, T2, s,
                       split_ap ? EFTB_PHASE_SECOND : EFTB_PHASE_ALL))) return rc;
    cur = T2;
  }
  if (c.has_project) {
    if ((rc = eftb_project(p, B, cur, out, stream))) return rc;
    cur = out;
  }
  const size_t nfinal = c.has_project ? z.out : z.T;
  if (terms_bm) EFTB_CUDA_CHECK(cudaMemcpyAsync(terms_bm, cur, nfinal * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (terms_pm && (rc = launch_to_point_major(cur, B, Bp, p->perm_rows, p->perm_out, terms_pm, s))) return rc;
  return EFTB_OK;
}

int eftb_workspace_terms(const eftb_plan* p, int B, const void* workspace, size_t workspace_bytes, int stage, double* terms_bm,
                         void* stream) {
  EFTB_NEED(p && workspace && terms_bm && B >= 1 && (stage == 0 || stage == 1), "NULL/invalid argument");
  const eftb_config& c = p->cfg;
  if (stage == 1 && !c.has_ap) { eftb_set_error("eftb_workspace_terms: plan built without AP"); return EFTB_ERR_NOT_BUILT; }
  if (workspace_bytes < eftb_workspace_bytes(p, B)) { eftb_set_error("eftb_workspace_terms: workspace too small"); return EFTB_ERR_WORKSPACE; }
  // the layout of eftb_eval_terms: u | F | D (coef | T2 after the AP stage) | P22 | Dg | T | ...
  const int Bp = eftb_padded_batch(B);
  Sizes z = sizes(p, Bp);
  const double* w = (const double*)workspace;
  const size_t ncoef = c.has_ap ? ap_coef_doubles(p, Bp) : 0;
  const double* D = w + z.u + z.F;
  const double* T = D + (z.D > ncoef + z.T ? z.D : ncoef + z.T) + z.P22 + z.Dg;
  const double* src = stage == 0 ? T : D + ncoef;
  EFTB_CUDA_CHECK(cudaMemcpyAsync(terms_bm, src, z.T * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return EFTB_OK;
}

struct eftb_operator {
  GemmMatrix A;
};

int eftb_operator_create(int M, int K, const double* host, eftb_operator** out) {
  if (M < 1 || K < 1 || !host || !out) { eftb_set_error("eftb_operator_create: bad argument"); return EFTB_ERR_ARG; }
  eftb_operator* op = new eftb_operator();
  int rc = gemm_upload(host, 1, M, K, &op->A);
  if (rc) { delete op; return rc; }
  *out = op;
  return EFTB_OK;
}

void eftb_operator_destroy(eftb_operator* op) {
  if (!op) return;
  gemm_free(&op->A);
  delete op;
}

int eftb_operator_apply(const eftb_operator* op, const double* X, double* C, int N, void* stream) {
  if (!op || !X || !C || N < 32 || N % 32) { eftb_set_error("eftb_operator_apply: bad argument (N must be a multiple of 32)"); return EFTB_ERR_ARG; }
  return gemm_run(op->A, X, C, N, 1, 1, 0, 0, 0, 0, (cudaStream_t)stream);
}

}  // extern "C"
