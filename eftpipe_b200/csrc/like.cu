// Likelihood stage: bias reduction to the non-Gaussian model vector PNG and the Gaussian-parameter
// derivative rows PG (parambasis.py:42-136, :249-316; likelihood.py:483-549), C^-1 product on the DMMA GEMM,
// and the analytic marginalisation  -2 ln P = -F1^T F2^-1 F1 + F0 + ln det(F2 / 2 pi)  (marginal.py:79-196)
// by a per-point Cholesky factorisation in shared memory.
#include <math.h>
#include <vector>
#include "common.cuh"

#define EFTB_MAX_TRACERS 8

struct eftb_like {
  eftb_like_config cfg;
  int32_t *nout = nullptr, *nterm = nullptr, *par_index = nullptr, *eastcoast = nullptr;
  int32_t h_nout[EFTB_MAX_TRACERS], h_nterm[EFTB_MAX_TRACERS];
  double* scales = nullptr;
  int32_t *d_tracer = nullptr, *d_row = nullptr, *d_row_g = nullptr;
  int32_t* res_perm = nullptr;  // rows d * (ngauss + 1) of the vector block: the residual PNG - data of data point d
  double *data = nullptr, *picc = nullptr;
  GemmMatrix invcov;
  int32_t *g_count = nullptr, *g_tracer = nullptr, *g_term = nullptr, *g_var = nullptr;
  double *g_coef = nullptr, *sigma_inv = nullptr, *sigma_inv_mu = nullptr;
  double mu_sigma_mu = 0.0;
};

namespace {

struct TracerPtrs {
  const double* terms[EFTB_MAX_TRACERS];
  const double* fg[EFTB_MAX_TRACERS];
};

struct VecArgs {
  TracerPtrs tp;
  const int32_t *nterm, *par_index, *eastcoast, *d_tracer, *d_row, *d_row_g, *g_count, *g_tracer, *g_term, *g_var;
  const double *scales, *data, *picc, *g_coef, *nuis;
  double* V;  // [ndata][ngauss+1][Bp]
  int Bp, ndata, ngauss;
};

// V[d][0] = PNG[d] - data[d],  V[d][1+g] = PG[g][d]
__global__ void __launch_bounds__(128) like_vectors_kernel(VecArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
  if (b >= a.Bp) return;
  const size_t Bp = a.Bp;
  const int tr = a.d_tracer[d], nt = a.nterm[tr];
  const double* term = a.tp.terms[tr] + ((size_t)a.d_row[d] * nt) * Bp + b;
  const double* termg = a.tp.terms[tr] + ((size_t)a.d_row_g[d] * nt) * Bp + b;  // rows of the marginalised derivatives
  const double f = a.tp.fg[tr][b];
  double par[EFTB_NPAR];
#pragma unroll
  for (int i = 0; i < EFTB_NPAR; ++i) {
    const int ix = a.par_index[tr * EFTB_NPAR + i];
    par[i] = ix < 0 ? 0.0 : a.nuis[(size_t)ix * Bp + b];
  }
  const double b1A = par[0], b2A = par[1], b3A = par[2], b4A = par[3], cctA = par[4], cr1A = par[5], cr2A = par[6];
  const double b1B = par[7], b2B = par[8], b3B = par[9], b4B = par[10], cctB = par[11], cr1B = par[12], cr2B = par[13];
  const double ce0 = par[14], cemono = par[15], cequad = par[16];
  const double cn0 = par[17], cn1 = par[18];  // west: cr4, cr6; east: ctilde, - (parambasis.py:96-107)
  const double* sc = a.scales + tr * 6;
  const double kmA = sc[0], krA = sc[1], ndA = sc[2], kmB = sc[3], krB = sc[4], ndB = sc[5];
  double bias[24];
  // parambasis.py:84-126
  bias[0] = b1A * b1B;
  bias[1] = (b1A + b1B) * f;
  bias[2] = f * f;
  if (!a.eastcoast[tr]) {
    bias[3] = b1A * cctB / (kmB * kmB) + b1B * cctA / (kmA * kmA);
    bias[4] = b1B * cr1A / (krA * krA) + b1A * cr1B / (krB * krB);
    bias[5] = b1B * cr2A / (krA * krA) + b1A * cr2B / (krB * krB);
    bias[6] = (cctA / (kmA * kmA) + cctB / (kmB * kmB)) * f;
    bias[7] = (cr1A / (krA * krA) + cr1B / (krB * krB)) * f;
    bias[8] = (cr2A / (krA * krA) + cr2B / (krB * krB)) * f;
  } else {
    bias[3] = -cctA - cctB;
    bias[4] = -(cr1A + cr1B) * f;
    bias[5] = -(cr2A + cr2B) * f * f;
    bias[6] = bias[7] = bias[8] = 0.0;
  }
  bias[9] = 1.0;
  bias[10] = 0.5 * (b1A + b1B);
  bias[11] = 0.5 * (b2A + b2B);
  bias[12] = 0.5 * (b3A + b3B);
  bias[13] = 0.5 * (b4A + b4B);
  bias[14] = b1A * b1B;
  bias[15] = 0.5 * (b1A * b2B + b1B * b2A);
  bias[16] = 0.5 * (b1A * b3B + b1B * b3A);
  bias[17] = 0.5 * (b1A * b4B + b1B * b4A);
  bias[18] = b2A * b2B;
  bias[19] = 0.5 * (b2A * b4B + b2B * b4A);
  bias[20] = b4A * b4B;
  const double x1 = 0.5 * (1.0 / ndA + 1.0 / ndB);
  const double x2 = 0.5 * (1.0 / ndA / (kmA * kmA) + 1.0 / ndB / (kmB * kmB));
  bias[21] = ce0 * x1;
  bias[22] = cemono * x2;
  bias[23] = cequad * x2;
  double tv[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) tv[i] = term[(size_t)i * Bp];
  const double f2 = f * f, f4 = f2 * f2;
  // same summation grouping as the reference: Plin + Ploop + Pct + Pst + Picc (parambasis.py:38-39)
  double plin = 0.0, ploop = 0.0, pct = 0.0, pst = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) plin += bias[i] * tv[i];
#pragma unroll
  for (int i = 3; i < 9; ++i) pct += bias[i] * tv[i];
#pragma unroll
  for (int i = 9; i < 21; ++i) ploop += bias[i] * tv[i];
#pragma unroll
  for (int i = 21; i < 24; ++i) pst += bias[i] * tv[i];
  if (nt > 24) {  // with_NNLO: Pct += bctNNLOAB . PctNNLOl (parambasis.py:96-107, :132-134)
    double bn0, bn1, bn2;
    if (!a.eastcoast[tr]) {
      const double kr4 = (krA * krA) * (krA * krA);
      bn0 = 0.25 * (b1A * b1A) / kr4 * cn0;
      bn1 = 0.25 * b1A / kr4 * cn1;
      bn2 = 0.0;
    } else {
      bn0 = cn0 * (-(b1A * b1A) * f4);
      bn1 = cn0 * (-2.0 * b1A * (f4 * f));
      bn2 = cn0 * (-(f4 * f2));
    }
    pct += (bn0 * term[(size_t)24 * Bp] + bn1 * term[(size_t)25 * Bp]) + bn2 * term[(size_t)26 * Bp];
  }
  const int nc = a.ngauss + 1;
  double* out = a.V + ((size_t)d * nc) * Bp + b;
  out[0] = (plin + ploop + pct + pst + a.picc[d]) - a.data[d];
  // dP/dg factors b1A^pa b1B^pb f^pf, code = pa | pb << 2 | pf << 4 (eftb200.h g_var)
  const double pw_a[4] = {1.0, b1A, b1A * b1A, b1A * b1A * b1A};
  const double pw_b[4] = {1.0, b1B, b1B * b1B, b1B * b1B * b1B};
  const double pw_f[8] = {1.0, f, f2, f2 * f, f4, f4 * f, f4 * f2, f4 * f2 * f};
  for (int g = 0; g < a.ngauss; ++g) {
    double v = 0.0;
    for (int e = 0; e < a.g_count[g]; ++e) {
      if (a.g_tracer[g * 2 + e] != tr) continue;
      const int base = (g * 2 + e) * 3;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double c = a.g_coef[base + q];
        if (c != 0.0) {
          const int code = a.g_var[base + q];
          v += c * (pw_a[code & 3] * pw_b[(code >> 2) & 3] * pw_f[(code >> 4) & 7]) * termg[(size_t)a.g_term[base + q] * Bp];
        }
      }
    }
    out[(size_t)(1 + g) * Bp] = v;
  }
}

struct FinArgs {
  const double *V, *Y, *sigma_inv, *sigma_inv_mu;
  const double *pp_loc, *pp_sinv;  // per-point Gaussian prior (callable loc / scale, marginal.py:13-20, :60-77): [B][nG] or NULL
  double mu_sigma_mu;
  double *logp, *bestfit, *fullchi2;
  int32_t* status;
  int B, Bp, ndata, ngauss, jeffreys;
};

// block (LF_PX points, ny entry-rows): every entry of the packed lower triangle of F2, of F1 and F0 is one dot product
// over the data index, sum_d V[d][ra] * Y[d][rb]; entry-rows are spread over threadIdx.y, loads are coalesced
// over the LF_PX points and unrolled 4x (independent partial sums) to keep several loads in flight
// LF_PX points per CTA (Bp is a multiple of 32, hence of LF_PX): 8 rather than a full warp of points, so that a batch of
// 1024 spreads over 128 CTAs instead of 32 - the kernel is latency bound and there are 148 SMs to hide it on
constexpr int LF_PX = 8;

__global__ void like_finish_kernel(FinArgs a) {
  extern __shared__ double sm[];
  const int nG = a.ngauss, nc = nG + 1;
  double* F2 = sm;                    // [nG][nG][LF_PX]
  double* F1 = F2 + (size_t)nG * nG * LF_PX;  // [nG][LF_PX]
  double* F0 = F1 + (size_t)nG * LF_PX;       // [LF_PX]
  double* F1o = F0 + LF_PX;                    // [nG][LF_PX]  F1 and diag(F2) before the factorisation (fullchi2 only)
  double* F2d = F1o + (size_t)nG * LF_PX;      // [nG][LF_PX]
  const int lx = threadIdx.x, g = threadIdx.y;
  const int b = blockIdx.x * LF_PX + lx;
  const size_t Bp = a.Bp, stride = (size_t)nc * Bp;
  const int ntri = nG * (nG + 1) / 2, nent = ntri + nG + 1;
  // prior terms: plan constants, or this point's own location / inverse variance (diagonal Sigma^-1; padding lanes: none)
  const bool pp = a.pp_sinv != nullptr;
  const double* ps = pp ? a.pp_sinv + (size_t)(b < a.B ? b : 0) * nG : nullptr;
  const double* pl = pp ? a.pp_loc + (size_t)(b < a.B ? b : 0) * nG : nullptr;
  auto sinv = [&](int i, int j) { return pp ? (i == j ? ps[i] : 0.0) : a.sigma_inv[i * nG + j]; };
  auto sinv_mu = [&](int i) { return pp ? ps[i] * pl[i] : a.sigma_inv_mu[i]; };
  auto mu_s_mu = [&]() {
    if (!pp) return a.mu_sigma_mu;
    double s = 0.0;
    for (int i = 0; i < nG; ++i) s = fma(ps[i] * pl[i], pl[i], s);
    return s;
  };
  for (int e = threadIdx.y; e < nent; e += blockDim.y) {
    int eg = 0, ej = 0, ra = 0, rb = 0;
    if (e < ntri) {
      while ((eg + 1) * (eg + 2) / 2 <= e) ++eg;
      ej = e - eg * (eg + 1) / 2;
      ra = 1 + eg; rb = 1 + ej;
    } else if (e < ntri + nG) {
      eg = e - ntri;
      ra = 1 + eg;
    }
    const double* pv = a.V + (size_t)ra * Bp + b;
    const double* py = a.Y + (size_t)rb * Bp + b;
    // 8 independent partial sums: 16 loads in flight per thread (the loop is a chain of L2 round trips otherwise)
    double sp[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int d = 0;
    for (; d + 8 <= a.ndata; d += 8) {
      double v[8], y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { v[u] = pv[(size_t)(d + u) * stride]; y[u] = py[(size_t)(d + u) * stride]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) sp[u] = fma(v[u], y[u], sp[u]);
    }
    double st = 0.0;
    for (; d < a.ndata; ++d) st = fma(pv[(size_t)d * stride], py[(size_t)d * stride], st);
    const double sum = (((sp[0] + sp[1]) + (sp[2] + sp[3])) + ((sp[4] + sp[5]) + (sp[6] + sp[7]))) + st;
    if (e < ntri) {
      const double v = sum + sinv(eg, ej);  // marginal.py:167-175
      F2[((size_t)eg * nG + ej) * LF_PX + lx] = v;
      F2[((size_t)ej * nG + eg) * LF_PX + lx] = v;
    } else if (e < ntri + nG) {
      F1[(size_t)eg * LF_PX + lx] = -sum + sinv_mu(eg);  // marginal.py:177-185
    } else {
      F0[lx] = sum + mu_s_mu();  // marginal.py:187-196
    }
  }
  __syncthreads();
  if (g != 0 || b >= a.B) return;
  if (a.fullchi2)
    for (int i = 0; i < nG; ++i) {
      F1o[(size_t)i * LF_PX + lx] = F1[(size_t)i * LF_PX + lx];
      F2d[(size_t)i * LF_PX + lx] = F2[((size_t)i * nG + i) * LF_PX + lx];
    }
  // in-place Cholesky F2 = L L^T on this point's column of shared memory
  bool ok = true;
  double logdet = 0.0;
  for (int j = 0; j < nG && ok; ++j) {
    double dj = F2[((size_t)j * nG + j) * LF_PX + lx];
    for (int k = 0; k < j; ++k) { const double l = F2[((size_t)j * nG + k) * LF_PX + lx]; dj -= l * l; }
    if (!(dj > 0.0)) { ok = false; break; }
    const double ljj = sqrt(dj), rjj = 1.0 / ljj;
    F2[((size_t)j * nG + j) * LF_PX + lx] = rjj;  // the diagonal holds 1 / L_jj: the solves below multiply instead of dividing
    if (!a.jeffreys) logdet += log(dj);           // 2 ln L_jj; not needed without the ln det term (marginal.py:119-120)
    for (int i = j + 1; i < nG; ++i) {
      double s = F2[((size_t)i * nG + j) * LF_PX + lx];
      for (int k = 0; k < j; ++k) s -= F2[((size_t)i * nG + k) * LF_PX + lx] * F2[((size_t)j * nG + k) * LF_PX + lx];
      F2[((size_t)i * nG + j) * LF_PX + lx] = s * rjj;
    }
  }
  if (!ok) {  // reference raises RuntimeError("det of F2ij <= 0") (marginal.py:113-116); here: flag the point
    a.logp[b] = -INFINITY;
    a.status[b] = 1;
    if (a.bestfit) for (int i = 0; i < nG; ++i) a.bestfit[(size_t)b * nG + i] = NAN;
    if (a.fullchi2) a.fullchi2[b] = NAN;
    return;
  }
  // y = L^-1 F1 ;  F1^T F2^-1 F1 = |y|^2
  double quad = 0.0;
  for (int i = 0; i < nG; ++i) {
    double s = F1[(size_t)i * LF_PX + lx];
    for (int k = 0; k < i; ++k) s -= F2[((size_t)i * nG + k) * LF_PX + lx] * F1[(size_t)k * LF_PX + lx];
    s *= F2[((size_t)i * nG + i) * LF_PX + lx];
    F1[(size_t)i * LF_PX + lx] = s;
    quad += s * s;
  }
  logdet -= nG * log(2.0 * M_PI);  // ln det(F2 / 2 pi)
  const double chi2 = -quad + F0[lx] + (a.jeffreys ? 0.0 : logdet);  // marginal.py:118-122
  a.logp[b] = -0.5 * chi2;
  a.status[b] = 0;
  if (a.bestfit || a.fullchi2) {  // bG = L^-T y (marginal.py:117)
    for (int i = nG - 1; i >= 0; --i) {
      double s = F1[(size_t)i * LF_PX + lx];
      for (int k = i + 1; k < nG; ++k) s -= F2[((size_t)k * nG + i) * LF_PX + lx] * F1[(size_t)k * LF_PX + lx];
      s *= F2[((size_t)i * nG + i) * LF_PX + lx];
      F1[(size_t)i * LF_PX + lx] = s;
      if (a.bestfit) a.bestfit[(size_t)b * nG + i] = s;
    }
  }
  if (a.fullchi2) {
    // marginal.py:129-131: chi^2 of the data at the best-fit bG, r = PNG + bG.PG - d, without the prior terms:
    //   r^T C^-1 r = F0' + 2 bG.g + bG^T F2' bG,  F2' = F2 - Sigma^-1,  g = PG C^-1 (PNG - d) = -(F1 - Sigma^-1 mu),
    //   F0' = F0 - mu^T Sigma^-1 mu.  The strict upper triangle of F2 still holds the values from before the factorisation.
    double full = F0[lx] - mu_s_mu();
    for (int i = 0; i < nG; ++i) {
      const double bi = F1[(size_t)i * LF_PX + lx];
      full -= 2.0 * bi * (F1o[(size_t)i * LF_PX + lx] - sinv_mu(i));
      full += bi * bi * (F2d[(size_t)i * LF_PX + lx] - sinv(i, i));
      for (int j = i + 1; j < nG; ++j)
        full += 2.0 * bi * F1[(size_t)j * LF_PX + lx] * (F2[((size_t)i * nG + j) * LF_PX + lx] - sinv(i, j));
    }
    a.fullchi2[b] = full;
  }
}

template <typename T>
int upload(T** dst, const T* src, size_t n) {
  if (n == 0) { *dst = nullptr; return EFTB_OK; }
  EFTB_CUDA_CHECK(cudaMalloc((void**)dst, n * sizeof(T)));
  EFTB_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return EFTB_OK;
}

int fill_vectors(const eftb_like* L, int Bp, const double* const* terms, const double* const* fg, const double* nuis,
                 double* V, cudaStream_t s) {
  VecArgs a;
  for (int t = 0; t < EFTB_MAX_TRACERS; ++t) {
    a.tp.terms[t] = t < L->cfg.ntracer ? terms[t] : nullptr;
    a.tp.fg[t] = t < L->cfg.ntracer ? fg[t] : nullptr;
  }
  a.nterm = L->nterm; a.par_index = L->par_index; a.eastcoast = L->eastcoast; a.d_tracer = L->d_tracer; a.d_row = L->d_row; a.d_row_g = L->d_row_g;
  a.g_count = L->g_count; a.g_tracer = L->g_tracer; a.g_term = L->g_term; a.g_var = L->g_var; a.scales = L->scales;
  a.data = L->data; a.picc = L->picc; a.g_coef = L->g_coef; a.nuis = nuis; a.V = V; a.Bp = Bp; a.ndata = L->cfg.ndata;
  a.ngauss = L->cfg.ngauss;
  dim3 grid((Bp + 127) / 128, L->cfg.ndata);
  like_vectors_kernel<<<grid, 128, 0, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

}  // namespace

extern "C" {

int eftb_like_create(const eftb_like_config* cfg, const eftb_like_constants* h, eftb_like** out) {
  if (!cfg || !h || !out) { eftb_set_error("eftb_like_create: NULL argument"); return EFTB_ERR_ARG; }
  if (cfg->ntracer < 1 || cfg->ntracer > EFTB_MAX_TRACERS || cfg->ndata < 1 || cfg->ngauss < 0 || cfg->ngauss > 31) {
    eftb_set_error("eftb_like_create: unsupported sizes (ntracer=%d ndata=%d ngauss=%d)", cfg->ntracer, cfg->ndata, cfg->ngauss);
    return EFTB_ERR_ARG;
  }
  eftb_like* L = new eftb_like();
  L->cfg = *cfg;
  const int nt = cfg->ntracer, nd = cfg->ndata, ng = cfg->ngauss;
  for (int t = 0; t < nt; ++t) { L->h_nout[t] = h->nout[t]; L->h_nterm[t] = h->nterm[t]; }
  int rc = 0;
  rc |= upload(&L->nout, h->nout, nt);
  rc |= upload(&L->nterm, h->nterm, nt);
  rc |= upload(&L->scales, h->scales, (size_t)nt * 6);
  rc |= upload(&L->par_index, h->par_index, (size_t)nt * EFTB_NPAR);
  rc |= upload(&L->eastcoast, h->eastcoast, nt);
  rc |= upload(&L->d_tracer, h->d_tracer, nd);
  rc |= upload(&L->d_row, h->d_row, nd);
  rc |= upload(&L->d_row_g, h->d_row_g ? h->d_row_g : h->d_row, nd);
  {
    std::vector<int32_t> perm(nd);
    for (int d = 0; d < nd; ++d) perm[d] = d * (ng + 1);
    rc |= upload(&L->res_perm, perm.data(), nd);
  }
  rc |= upload(&L->data, h->data, nd);
  rc |= upload(&L->picc, h->picc, nd);
  rc |= gemm_upload(h->invcov, 1, nd, nd, &L->invcov);
  rc |= upload(&L->g_count, h->g_count, ng);
  rc |= upload(&L->g_tracer, h->g_tracer, (size_t)ng * 2);
  rc |= upload(&L->g_term, h->g_term, (size_t)ng * 6);
  rc |= upload(&L->g_var, h->g_var, (size_t)ng * 6);
  rc |= upload(&L->g_coef, h->g_coef, (size_t)ng * 6);
  rc |= upload(&L->sigma_inv, h->sigma_inv, (size_t)ng * ng);
  rc |= upload(&L->sigma_inv_mu, h->sigma_inv_mu, ng);
  L->mu_sigma_mu = h->mu_sigma_mu;
  if (rc) { eftb_like_destroy(L); return EFTB_ERR_CUDA; }
  *out = L;
  return EFTB_OK;
}

void eftb_like_destroy(eftb_like* L) {
  if (!L) return;
  void* ptrs[] = {L->nout, L->nterm, L->scales, L->par_index, L->eastcoast, L->d_tracer, L->d_row, L->d_row_g, L->res_perm, L->data, L->picc,
                  L->g_count, L->g_tracer, L->g_term, L->g_var, L->g_coef, L->sigma_inv, L->sigma_inv_mu};
  for (void* p : ptrs) if (p) cudaFree(p);
  gemm_free(&L->invcov);
  delete L;
}

size_t eftb_like_workspace_bytes(const eftb_like* L, int B) {
  if (!L || B < 1) return 0;
  size_t Bp = eftb_padded_batch(B);
  return 2 * (size_t)L->cfg.ndata * (L->cfg.ngauss + 1) * Bp * sizeof(double);
}

int eftb_like_eval(const eftb_like* L, int B, const double* const* terms, const double* const* fgrowth, const double* nuis,
                   double* logp, double* bestfit, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  return eftb_like_eval_full(L, B, terms, fgrowth, nuis, logp, bestfit, nullptr, status, workspace, workspace_bytes, stream);
}

int eftb_like_eval_full(const eftb_like* L, int B, const double* const* terms, const double* const* fgrowth, const double* nuis,
                        double* logp, double* bestfit, double* fullchi2, int32_t* status, void* workspace, size_t workspace_bytes,
                        void* stream) {
  return eftb_like_eval_priors(L, B, terms, fgrowth, nuis, nullptr, nullptr, logp, bestfit, fullchi2, status, workspace,
                               workspace_bytes, stream);
}

int eftb_like_eval_priors(const eftb_like* L, int B, const double* const* terms, const double* const* fgrowth, const double* nuis,
                          const double* prior_loc, const double* prior_sigma_inv, double* logp, double* bestfit, double* fullchi2,
                          int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  if ((prior_loc == nullptr) != (prior_sigma_inv == nullptr)) {
    eftb_set_error("eftb_like_eval_priors: prior_loc and prior_sigma_inv go together");
    return EFTB_ERR_ARG;
  }
  if (!L || !terms || !fgrowth || !nuis || !logp || !status || !workspace || B < 1) {
    eftb_set_error("eftb_like_eval: NULL/invalid argument");
    return EFTB_ERR_ARG;
  }
  if (workspace_bytes < eftb_like_workspace_bytes(L, B)) { eftb_set_error("eftb_like_eval: workspace too small"); return EFTB_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int Bp = eftb_padded_batch(B), nd = L->cfg.ndata, nc = L->cfg.ngauss + 1;
  double* V = (double*)workspace;
  double* Y = V + (size_t)nd * nc * Bp;
  int rc = fill_vectors(L, Bp, terms, fgrowth, nuis, V, s);
  if (rc) return rc;
  rc = gemm_run(L->invcov, V, Y, nc * Bp, 1, 1, 0, 0, 0, 0, s);
  if (rc) return rc;
  FinArgs a{V, Y, L->sigma_inv, L->sigma_inv_mu, prior_loc, prior_sigma_inv, L->mu_sigma_mu, logp, bestfit, fullchi2, status, B, Bp, nd,
            L->cfg.ngauss, L->cfg.jeffreys};
  const int nG = L->cfg.ngauss;
  size_t smem = sizeof(double) * ((size_t)nG * nG * LF_PX + (size_t)3 * nG * LF_PX + LF_PX);
  static DeviceSmem configured;
  EFTB_SET_SMEM(configured, like_finish_kernel, smem);
  const int nent = nG * (nG + 1) / 2 + nG + 1;
  dim3 block(LF_PX, nent < 64 ? nent : 64), grid(Bp / LF_PX);
  like_finish_kernel<<<grid, block, smem, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int eftb_like_vectors(const eftb_like* L, int B, const double* const* terms, const double* const* fgrowth, const double* nuis,
                      double* vec, void* workspace, size_t workspace_bytes, void* stream) {
  if (!L || !terms || !fgrowth || !nuis || !vec || !workspace || B < 1) { eftb_set_error("eftb_like_vectors: NULL/invalid argument"); return EFTB_ERR_ARG; }
  if (workspace_bytes < eftb_like_workspace_bytes(L, B)) { eftb_set_error("eftb_like_vectors: workspace too small"); return EFTB_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int Bp = eftb_padded_batch(B), nd = L->cfg.ndata, nc = L->cfg.ngauss + 1;
  double* V = (double*)workspace;
  int rc = fill_vectors(L, Bp, terms, fgrowth, nuis, V, s);
  if (rc) return rc;
  return launch_to_point_major(V, B, Bp, nd * nc, nullptr, vec, s);
}

int eftb_like_residuals(const eftb_like* L, int B, const void* workspace, double* out, void* stream) {
  if (!L || !workspace || !out || B < 1) { eftb_set_error("eftb_like_residuals: NULL/invalid argument"); return EFTB_ERR_ARG; }
  return launch_to_point_major((const double*)workspace, B, eftb_padded_batch(B), L->cfg.ndata, L->res_perm, out, (cudaStream_t)stream);
}

}  // extern "C"
