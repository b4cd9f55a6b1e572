// Likelihood stage: bias reduction to the non-Gaussian model vector PNG and the Gaussian-parameter
// derivative rows PG (parambasis.py:42-136, :249-316; likelihood.py:483-549), C^-1 product on the DMMA GEMM,
// and the analytic marginalisation  -2 ln P = -F1^T F2^-1 F1 + F0 + ln det(F2 / 2 pi)  (marginal.py:79-196)
// by a per-point Cholesky factorisation in shared memory.
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"

#define EFTB_MAX_TRACERS 8

struct eftb_like {
  eftb_like_config cfg;
  int32_t *nout = nullptr, *nterm = nullptr, *par_index = nullptr, *eastcoast = nullptr;
  int32_t h_nout[EFTB_MAX_TRACERS], h_nterm[EFTB_MAX_TRACERS];
  double* scales = nullptr;
  int32_t *d_tracer = nullptr, *d_row = nullptr, *d_row_g = nullptr;
  int32_t *mode = nullptr, *xb_off = nullptr, *xg_off = nullptr;  // explicit-bias tracers (custom EFT bases)
  int32_t* res_perm = nullptr;  // rows d * (ngauss + 1) of the vector block: the residual PNG - data of data point d
  double *data = nullptr, *picc = nullptr;
  GemmMatrix invcov;   // C^-1 (two-operand form: A = V, Bm = C^-1 V), used when C^-1 has no Cholesky factor
  GemmMatrix factor;   // L^T with C^-1 = L L^T: W = L^T V and the Gram matrix is W^T W (one operand, half the traffic)
  bool has_factor = false;
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
  const int32_t *mode, *xb_off, *xg_off;
  int ntracer;
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
  if (a.mode && a.mode[tr]) {
    // a custom basis: the coefficient of every term row comes with the point (probed on the host from the basis' own
    // reduce_Plk / reduce_Plk_gaussian_table, parambasis.py:139-162)
    const double* xb = a.nuis + (size_t)a.xb_off[tr] * Bp + b;
    double png = 0.0;
    for (int i = 0; i < nt; ++i) png = fma(xb[(size_t)i * Bp], term[(size_t)i * Bp], png);
    double* out = a.V + ((size_t)d * (a.ngauss + 1)) * Bp + b;
    out[0] = (png + a.picc[d]) - a.data[d];
    for (int g = 0; g < a.ngauss; ++g) {
      const int off = a.xg_off[g * a.ntracer + tr];
      double v = 0.0;
      if (off >= 0) {
        const double* xg = a.nuis + (size_t)off * Bp + b;
        for (int i = 0; i < nt; ++i) v = fma(xg[(size_t)i * Bp], termg[(size_t)i * Bp], v);
      }
      out[(size_t)(1 + g) * Bp] = v;
    }
    return;
  }
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
  const double *A, *Bm;  // [ndata][ngauss+1][Bp]: Gram[a][b] = sum_d A[d][a] Bm[d][b]  (A == Bm = L^T V, or A = V, Bm = C^-1 V)
  const double *sigma_inv, *sigma_inv_mu;
  const double *pp_loc, *pp_sinv;  // per-point Gaussian prior (callable loc / scale, marginal.py:13-20, :60-77): [B][nG] or NULL
  double mu_sigma_mu;
  double *logp, *bestfit, *fullchi2;
  int32_t* status;
  int B, Bp, ndata, ngauss, jeffreys;
};

__device__ __forceinline__ void cp_async16z(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = pred ? 16 : 0;  // 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

// Marginalisation of one tile of PX points (marginal.py:79-196).
//
// Phase 1 - Gram matrix.  Every entry of F2 (packed upper triangle), of F1 and F0 is a dot product over the data index of
// two columns of the vector block, i.e. the (nc x nc) Gram matrix G = A^T Bm per point, nc = ngauss + 1 (column 0 = the
// residual PNG - d).  The tile's rows stream ONCE from HBM through a double-buffered cp.async ring of LG_DCH data rows
// (the first version re-read every row from L2 for each of the ~130 entries: 2.5 GB of L2 traffic per 8192 points, the
// whole cost of the stage); thread = (point x, 4 x 4 block of G's upper triangle): 8 conflict-free shared loads per 16 FMA.
// Phase 2 - one warp per point: F2 = G + Sigma^-1, left-looking Cholesky with lane = row (the lanes' partial dots run in
// parallel, the pivot of column j is broadcast with __shfl_sync), forward / backward substitution and the best-fit
// chi^2 the same way.
constexpr int LG_DCH = 8;  // data rows per stage

template <int NC, int PX>  // NC: padded nc (16 or 32); PX: points per CTA
__global__ void __launch_bounds__(PX * (NC / 4) * (NC / 4 + 1) / 2) like_gram_kernel(FinArgs a) {
  constexpr int NBR = NC / 4, NBLK = NBR * (NBR + 1) / 2, NTHR = PX * NBLK, SP = NC + 1;
  extern __shared__ __align__(16) double sm[];
  const int nG = a.ngauss, nc = nG + 1;
  const bool two = a.A != a.Bm;
  double* tileA = sm;                                          // [2][LG_DCH][NC][PX]
  double* tileB = tileA + (two ? 2 * LG_DCH * NC * PX : 0);    // second operand (or the same tile)
  double* S = tileB + 2 * LG_DCH * NC * PX;                    // [PX][NC][SP]: upper triangle = G, strict lower = L
  double* dg = S + (size_t)PX * NC * SP;                       // [PX][NC]: diag(F2), later 1 / L_jj
  const int tid = threadIdx.x, x = tid % PX, blk = tid / PX;
  const int b0 = blockIdx.x * PX;
  const size_t Bp = a.Bp;
  // block (bi <= bj) of the upper triangle
  int bi = 0, rem = blk;
  while (rem >= NBR - bi) { rem -= NBR - bi; ++bi; }
  const int bj = bi + rem;

  auto load_stage = [&](int d0, int buf) {
    // rows d0 .. d0 + LG_DCH: for every (d, c < NC) the PX consecutive points, 16 bytes at a time; c >= nc and d >= ndata are zero-filled
    constexpr int CH = LG_DCH * NC * (PX / 2);
    for (int i = tid; i < CH; i += NTHR) {
      const int q = i % (PX / 2), c = (i / (PX / 2)) % NC, d = i / ((PX / 2) * NC);
      const bool ok = c < nc && d0 + d < a.ndata;
      const size_t off = ok ? ((size_t)(d0 + d) * nc + c) * Bp + b0 + 2 * q : 0;
      const int so = buf * LG_DCH * NC * PX + (d * NC + c) * PX + 2 * q;
      cp_async16z(tileA + so, a.A + off, ok);
      if (two) cp_async16z(tileB + so, a.Bm + off, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const int nst = (a.ndata + LG_DCH - 1) / LG_DCH;
  load_stage(0, 0);
  for (int st = 0; st < nst; ++st) {
    const int buf = st & 1;
    if (st + 1 < nst) {
      load_stage((st + 1) * LG_DCH, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    const double* ta = tileA + buf * LG_DCH * NC * PX + x;
    const double* tb = tileB + buf * LG_DCH * NC * PX + x;
#pragma unroll
    for (int d = 0; d < LG_DCH; ++d) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = ta[(d * NC + 4 * bi + i) * PX];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = tb[(d * NC + 4 * bj + j) * PX];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  {  // G into shared memory (entries below the diagonal of a diagonal block are simply not used)
    double* Sx = S + (size_t)x * NC * SP;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * bi + i <= 4 * bj + j) Sx[(4 * bi + i) * SP + 4 * bj + j] = acc[i][j];
  }
  __syncthreads();

  // ---- one warp per point ----
  const int warp = tid >> 5, lane = tid & 31, nwarp = NTHR / 32;
  for (int px = warp; px < PX; px += nwarp) {
    const int b = b0 + px;
    if (b >= a.B) continue;  // padding lanes of the batch
    double* Sx = S + (size_t)px * NC * SP;
    double* dx = dg + (size_t)px * NC;
    // rows / columns of Sx: 0 = residual, 1 + i = Gaussian parameter i; lane = Gaussian row i
    const int i = lane;
    const bool row = i < nG;
    const bool pp = a.pp_sinv != nullptr;
    const double* ps = pp ? a.pp_sinv + (size_t)b * nG : nullptr;
    const double* pl = pp ? a.pp_loc + (size_t)b * nG : nullptr;
    auto sinv = [&](int r, int c) { return pp ? (r == c ? ps[r] : 0.0) : a.sigma_inv[r * nG + c]; };
    // F2 = G + Sigma^-1 (marginal.py:167-175): strict lower triangle <- upper + prior, diagonal to dx
    double f1 = 0.0, musmu_part = 0.0;
    if (row) {
      for (int j = 0; j < i; ++j) Sx[(1 + i) * SP + 1 + j] = Sx[(1 + j) * SP + 1 + i] + sinv(i, j);
      dx[i] = Sx[(1 + i) * SP + 1 + i] + sinv(i, i);
      const double smu = pp ? ps[i] * pl[i] : a.sigma_inv_mu[i];
      f1 = -Sx[1 + i] + smu;  // F1 = -PG C^-1 (PNG - d) + Sigma^-1 mu (marginal.py:177-185); G[0][1+i]
      if (pp) musmu_part = smu * pl[i];
    }
    double musmu = a.mu_sigma_mu;
    if (pp) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) musmu_part += __shfl_xor_sync(0xffffffffu, musmu_part, o);
      musmu = musmu_part;
    }
    const double F0 = Sx[0] + musmu;  // marginal.py:187-196
    __syncwarp();
    // left-looking Cholesky F2 = L L^T: column j of L from the rows' partial dot products, pivot broadcast by shuffle
    bool ok = true;
    double logdet = 0.0, rdiag = 0.0;
    for (int j = 0; j < nG; ++j) {
      double sacc = 0.0;
      if (row && i >= j) {
        sacc = i == j ? dx[j] : Sx[(1 + i) * SP + 1 + j];
        const double* li = Sx + (1 + i) * SP + 1;
        const double* lj = Sx + (1 + j) * SP + 1;
        for (int k = 0; k < j; ++k) sacc = fma(-li[k], lj[k], sacc);
      }
      const double dj = __shfl_sync(0xffffffffu, sacc, j);
      if (!(dj > 0.0)) { ok = false; break; }
      const double rj = 1.0 / sqrt(dj);
      if (!a.jeffreys) logdet += log(dj);  // 2 ln L_jj; not needed without the ln det term (marginal.py:119-120)
      if (row && i > j) Sx[(1 + i) * SP + 1 + j] = sacc * rj;
      if (i == j) rdiag = rj;
      __syncwarp();
    }
    if (!ok) {  // reference raises RuntimeError("det of F2ij <= 0") (marginal.py:113-116); here: flag the point
      if (lane == 0) {
        a.logp[b] = -INFINITY;
        a.status[b] = 1;
        if (a.fullchi2) a.fullchi2[b] = NAN;
      }
      if (a.bestfit && row) a.bestfit[(size_t)b * nG + i] = NAN;
      continue;
    }
    // y = L^-1 F1 ;  F1^T F2^-1 F1 = |y|^2
    double quad = 0.0, y = 0.0, bcur = f1;
    for (int j = 0; j < nG; ++j) {
      const double yj = __shfl_sync(0xffffffffu, bcur, j) * __shfl_sync(0xffffffffu, rdiag, j);
      quad = fma(yj, yj, quad);
      if (row && i > j) bcur = fma(-Sx[(1 + i) * SP + 1 + j], yj, bcur);
      if (i == j) y = yj;
    }
    if (lane == 0) {
      logdet -= nG * log(2.0 * M_PI);  // ln det(F2 / 2 pi)
      const double chi2 = -quad + F0 + (a.jeffreys ? 0.0 : logdet);  // marginal.py:118-122
      a.logp[b] = -0.5 * chi2;
      a.status[b] = 0;
    }
    if (a.bestfit || a.fullchi2) {  // bG = L^-T y (marginal.py:117)
      double xb = 0.0, ycur = y;
      for (int j = nG - 1; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, ycur, j) * __shfl_sync(0xffffffffu, rdiag, j);
        if (row && i < j) ycur = fma(-Sx[(1 + j) * SP + 1 + i], xj, ycur);
        if (i == j) xb = xj;
      }
      if (a.bestfit && row) a.bestfit[(size_t)b * nG + i] = xb;
      if (a.fullchi2) {
        // marginal.py:129-131: chi^2 of the data at the best-fit bG, r = PNG + bG.PG - d, without the prior terms:
        //   r^T C^-1 r = G00 + 2 sum_i bG_i G[0][i] + sum_ij bG_i bG_j G[i][j]   (G = the upper triangle, untouched)
        double t = 0.0;
        if (row) dx[i] = xb;  // diag(F2) is no longer needed: the best fit, for the other lanes
        __syncwarp();
        if (row) {
          double off = 0.0;
          for (int j = i + 1; j < nG; ++j) off = fma(dx[j], Sx[(1 + i) * SP + 1 + j], off);
          t = xb * (2.0 * Sx[1 + i] + xb * Sx[(1 + i) * SP + 1 + i] + 2.0 * off);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) a.fullchi2[b] = Sx[0] + t;
      }
    }
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
  a.mode = L->mode; a.xb_off = L->xb_off; a.xg_off = L->xg_off; a.ntracer = L->cfg.ntracer;
  a.data = L->data; a.picc = L->picc; a.g_coef = L->g_coef; a.nuis = nuis; a.V = V; a.Bp = Bp; a.ndata = L->cfg.ndata;
  a.ngauss = L->cfg.ngauss;
  dim3 grid((Bp + 127) / 128, L->cfg.ndata);
  like_vectors_kernel<<<grid, 128, 0, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

template <int NC, int PX>
int launch_gram_t(const FinArgs& a, cudaStream_t s) {
  constexpr int NBR = NC / 4, NTHR = PX * NBR * (NBR + 1) / 2;
  const bool two = a.A != a.Bm;
  const size_t smem = sizeof(double) * ((size_t)(two ? 2 : 1) * 2 * LG_DCH * NC * PX + (size_t)PX * NC * (NC + 1) + (size_t)PX * NC);
  static DeviceSmem configured;
  EFTB_SET_SMEM(configured, (like_gram_kernel<NC, PX>), smem);
  like_gram_kernel<NC, PX><<<a.Bp / PX, NTHR, smem, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int launch_gram(const FinArgs& a, cudaStream_t s) {
  return a.ngauss + 1 <= 16 ? launch_gram_t<16, 16>(a, s) : launch_gram_t<32, 8>(a, s);
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
  if (h->mode) {
    if (!h->xb_off || (ng && !h->xg_off)) { eftb_set_error("eftb_like_create: mode given without xb_off / xg_off"); eftb_like_destroy(L); return EFTB_ERR_ARG; }
    rc |= upload(&L->mode, h->mode, nt);
    rc |= upload(&L->xb_off, h->xb_off, nt);
    rc |= upload(&L->xg_off, h->xg_off, (size_t)ng * nt);
  }
  {
    std::vector<int32_t> perm(nd);
    for (int d = 0; d < nd; ++d) perm[d] = d * (ng + 1);
    rc |= upload(&L->res_perm, perm.data(), nd);
  }
  rc |= upload(&L->data, h->data, nd);
  rc |= upload(&L->picc, h->picc, nd);
  rc |= gemm_upload(h->invcov, 1, nd, nd, &L->invcov);
  {  // C^-1 = L L^T on the host (nd x nd, once); a matrix that is not positive definite keeps the two-operand form
    std::vector<double> lo((size_t)nd * nd, 0.0);
    bool pd = true;
    for (int j = 0; j < nd && pd; ++j) {
      double dj = h->invcov[(size_t)j * nd + j];
      for (int k = 0; k < j; ++k) dj -= lo[(size_t)j * nd + k] * lo[(size_t)j * nd + k];
      if (!(dj > 0.0)) { pd = false; break; }
      const double ljj = sqrt(dj);
      lo[(size_t)j * nd + j] = ljj;
      for (int i = j + 1; i < nd; ++i) {
        double v = 0.5 * (h->invcov[(size_t)i * nd + j] + h->invcov[(size_t)j * nd + i]);
        for (int k = 0; k < j; ++k) v -= lo[(size_t)i * nd + k] * lo[(size_t)j * nd + k];
        lo[(size_t)i * nd + j] = v / ljj;
      }
    }
    if (pd) {
      std::vector<double> lt((size_t)nd * nd, 0.0);
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j <= i; ++j) lt[(size_t)j * nd + i] = lo[(size_t)i * nd + j];
      rc |= gemm_upload(lt.data(), 1, nd, nd, &L->factor);
      L->has_factor = true;
    }
  }
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
  void* ptrs[] = {L->mode, L->xb_off, L->xg_off,
                  L->nout, L->nterm, L->scales, L->par_index, L->eastcoast, L->d_tracer, L->d_row, L->d_row_g, L->res_perm, L->data, L->picc,
                  L->g_count, L->g_tracer, L->g_term, L->g_var, L->g_coef, L->sigma_inv, L->sigma_inv_mu};
  for (void* p : ptrs) if (p) cudaFree(p);
  gemm_free(&L->invcov);
  gemm_free(&L->factor);
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
  static const bool force_two = getenv("EFTB_LIKE_TWO_OPERAND") != nullptr;  // tuning / test knob: the C^-1 V form
  const bool one = L->has_factor && !force_two;
  rc = gemm_run(one ? L->factor : L->invcov, V, Y, nc * Bp, 1, 1, 0, 0, 0, 0, s);
  if (rc) return rc;
  FinArgs a{one ? Y : V, Y, L->sigma_inv, L->sigma_inv_mu, prior_loc, prior_sigma_inv, L->mu_sigma_mu, logp, bestfit, fullchi2, status, B, Bp,
            nd, L->cfg.ngauss, L->cfg.jeffreys};
  return launch_gram(a, s);
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
