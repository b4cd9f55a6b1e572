// IR resummation (Resum.Ps, pybird.py:1413-1464), restructured.
//
// Reference: for each of the 3*(1+1+12) correlation-function rows C_r(s) and each of the 2*NIR filter powers
// XpYp_j(s) it runs a 192-point FFTLog and a Bessel back-transform (1344 python iterations), then contracts
// with the bulk coefficients Q^{ll'}_u(f).  For fixed grids that FFTLog chain is a fixed real operator
// R[v,k,s] (plan.py resum_operator), and k2p[j,k] * XpYp[j,s] = z^{p+1} or Y k^2 z^p with z = k^2 X(s), so
//
//   out[l,i,k] += sum_{l',s} T_a[l,l',k,s] C[l',i,s],
//   T_a[l,l',k,s] = sum_v R[v,k,s] * ( z * A(z) + Y k^2 * B(z) ),   A, B: degree NIR-1 polynomials in z whose
//                   coefficients are Q_a[l,l',p*Na+v](f) and Q_a[l,l',(NIR+p)*Na+v](f).
//
// One CTA per cosmology: Q(f) is expanded once into shared memory, then each thread owns one (a, l, k)
// output column and sweeps s in chunks of 4 (Horner in z with broadcast coefficient loads, 8 DFMA per
// 16-byte shared load).  FP64-FMA bound: ~7e6 DFMA per cosmology at Nl=3.
#include "common.cuh"

namespace {

struct ResumArgs {
  const double *F, *Cr, *f, *R, *q, *kr2, *l11, *lct, *lctnnlo;
  double* T;
  int B, Bp, Nk, Ns, nterm, ncr, with_nnlo, Nkr, Nklow, qdeg, row_x, row_y;
};

template <int NL, int NIR, int NA>
__global__ void __launch_bounds__(288) resum_kernel(ResumArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NN = 2 * NIR * NA;
  double2* Qf = reinterpret_cast<double2*>(sm);       // [2][NL][NL][NA][NIR] (x: X^{p+1} coefficient, y: Y X^p)
  double* Xs = sm + 2 * (2 * NL * NL * NA * NIR);     // [Ns]
  double* Ys = Xs + a.Ns;                             // [Ns]
  double* Cs = Ys + a.Ns;                             // [NL][ncr][Ns]
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t Bp = a.Bp;
  const double f1 = a.f[b];

  for (int s = tid; s < a.Ns; s += blockDim.x) {
    Xs[s] = a.F[(size_t)(a.row_x + s) * Bp + b];
    Ys[s] = a.F[(size_t)(a.row_y + s) * Bp + b];
  }
  for (int i = tid; i < NL * a.ncr * a.Ns; i += blockDim.x) Cs[i] = a.Cr[(size_t)i * Bp + b];
  // Q^{ll'}_u(f): polynomial in f (pybird.py:1367-1380 evaluates the reference's lambdas)
  for (int i = tid; i < 2 * NL * NL * NN; i += blockDim.x) {
    const double* qc = a.q + (size_t)i * a.qdeg;
    double v = 0.0;
    for (int d = a.qdeg - 1; d >= 0; --d) v = fma(v, f1, qc[d]);
    const int u = i % NN, all = i / NN;  // all = (a*NL + l)*NL + lp
    const int j = u / NA, vv = u % NA;
    double* dst = reinterpret_cast<double*>(Qf + ((size_t)all * NA + vv) * NIR + (j < NIR ? j : j - NIR));
    dst[j < NIR ? 0 : 1] = v;
  }
  __syncthreads();

  const int ntask = 2 * NL * a.Nkr;
  for (int task = tid; task < ntask; task += blockDim.x) {
    const int ia = task / (NL * a.Nkr), l = (task / a.Nkr) % NL, ik = task % a.Nkr;
    const double k2 = a.kr2[ik];
    double acc[12], lin[NL], nnlo[NL];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) lin[i] = nnlo[i] = 0.0;
    for (int s0 = 0; s0 < a.Ns; s0 += 4) {
      double z[4], yk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool ok = s0 + c < a.Ns;
        z[c] = ok ? k2 * Xs[s0 + c] : 0.0;
        yk[c] = ok ? k2 * Ys[s0 + c] : 0.0;
      }
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) {
        double Tl[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int v = 0; v < NA; ++v) {
          const double2* qv = Qf + ((size_t)((ia * NL + l) * NL + lp) * NA + v) * NIR;
          double A[4] = {0.0, 0.0, 0.0, 0.0}, Bq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int p = NIR - 1; p >= 0; --p) {
            const double2 qq = qv[p];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              A[c] = fma(A[c], z[c], qq.x);
              Bq[c] = fma(Bq[c], z[c], qq.y);
            }
          }
          const double* rr = a.R + ((size_t)v * a.Nkr + ik) * a.Ns + s0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double r = (s0 + c < a.Ns) ? __ldg(rr + c) : 0.0;
            Tl[c] = fma(r, fma(z[c], A[c], yk[c] * Bq[c]), Tl[c]);
          }
        }
        const double* crow = Cs + (size_t)lp * a.ncr * a.Ns + s0;
        const int ns = min(4, a.Ns - s0);
        if (ia == 0) {
          for (int c = 0; c < ns; ++c) lin[lp] = fma(Tl[c], crow[c], lin[lp]);
        } else {
          for (int c = 0; c < ns; ++c) lin[lp] = fma(Tl[c], crow[a.Ns + c], lin[lp]);
#pragma unroll
          for (int i = 0; i < 12; ++i)
            for (int c = 0; c < ns; ++c) acc[i] = fma(Tl[c], crow[(size_t)(2 + i) * a.Ns + c], acc[i]);
          if (a.with_nnlo)
            for (int c = 0; c < ns; ++c) nnlo[lp] = fma(Tl[c], crow[(size_t)14 * a.Ns + c], nnlo[lp]);
        }
      }
    }
    double* out = a.T + ((size_t)(l * a.Nk + a.Nklow + ik) * a.nterm) * Bp + b;
    if (ia == 0) {
      for (int i = 0; i < 3; ++i) {
        double v = 0.0;
#pragma unroll
        for (int lp = 0; lp < NL; ++lp) v = fma(a.l11[lp * 3 + i], lin[lp], v);
        out[(size_t)i * Bp] += v;  // pybird.py:1442, :1445
      }
    } else {
      for (int i = 0; i < 6; ++i) {
        double v = 0.0;
#pragma unroll
        for (int lp = 0; lp < NL; ++lp) v = fma(a.lct[lp * 6 + i], lin[lp], v);
        out[(size_t)(3 + i) * Bp] += v;  // pybird.py:1443, :1446
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) out[(size_t)(9 + i) * Bp] += acc[i];  // pybird.py:1444, :1462
      if (a.with_nnlo)
        for (int i = 0; i < 3; ++i) {
          double v = 0.0;
#pragma unroll
          for (int lp = 0; lp < NL; ++lp) v = fma(a.lctnnlo[lp * 3 + i], nnlo[lp], v);
          out[(size_t)(24 + i) * Bp] += v;  // pybird.py:1455-1458
        }
    }
  }
}

template <int NL, int NIR, int NA>
int run(const ResumArgs& a, cudaStream_t s) {
  size_t smem = sizeof(double) * (2 * (2 * NL * NL * NA * NIR) + 2 * a.Ns + (size_t)NL * a.ncr * a.Ns);
  static bool configured = false;
  if (!configured) {
    EFTB_CUDA_CHECK(cudaFuncSetAttribute(resum_kernel<NL, NIR, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  resum_kernel<NL, NIR, NA><<<a.B, 288, smem, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

}  // namespace

int launch_resum(const eftb_plan* p, int B, int Bp, const double* F, const double* Cr, const double* f, double* T,
                 cudaStream_t s) {
  const eftb_config& c = p->cfg;
  ResumArgs a;
  a.F = F; a.Cr = Cr; a.f = f; a.R = p->R; a.q = p->q; a.kr2 = p->kr2; a.l11 = p->l11; a.lct = p->lct;
  a.lctnnlo = p->lctnnlo; a.T = T; a.B = B; a.Bp = Bp; a.Nk = c.Nk; a.Ns = c.Ns; a.nterm = c.nterm;
  a.ncr = 14 + (c.with_nnlo ? 1 : 0); a.with_nnlo = c.with_nnlo; a.Nkr = c.Nkr; a.Nklow = c.Nklow; a.qdeg = c.qdeg;
  a.row_x = c.row_x; a.row_y = c.row_y;
  if (c.Nl == 3 && c.NIR == 16 && c.Na == 3) return run<3, 16, 3>(a, s);
  if (c.Nl == 2 && c.NIR == 8 && c.Na == 2) return run<2, 8, 2>(a, s);
  eftb_set_error("resum: unsupported (Nl, NIR, Na) = (%d, %d, %d)", c.Nl, c.NIR, c.Na);
  return EFTB_ERR_ARG;
}
