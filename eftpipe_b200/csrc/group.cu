// Term assembly: Legendre weighting of every loop term, regrouping by powers of the growth rate f into
// the 12 bias-independent loop rows, stochastic basis, shot-noise subtraction
// (Bird.setPsCfl / reducePsCfl / setPstl / subtractShotNoise, pybird.py:737-866).
// Element-wise over the batch (one lane per cosmology), one CTA row per k node or per s node.
#include "common.cuh"

namespace {

// (row, f-power, term) triples of reducePsCfl (pybird.py:762-846), as X-macros so that every index is a
// literal after preprocessing (keeps the accumulators in registers)
#define EFTB_G22(X)                                                                                               \
  X(0, 2, 20) X(0, 3, 23) X(0, 3, 24) X(0, 4, 25) X(0, 4, 26) X(0, 4, 27) X(1, 1, 9) X(1, 2, 14) X(1, 2, 15)      \
  X(1, 3, 21) X(1, 3, 22) X(2, 1, 10) X(2, 2, 16) X(2, 2, 17) X(4, 1, 11) X(4, 2, 18) X(4, 2, 19) X(5, 0, 0)      \
  X(5, 1, 6) X(5, 2, 12) X(5, 2, 13) X(6, 0, 1) X(6, 1, 7) X(8, 0, 2) X(8, 1, 8) X(9, 0, 3) X(10, 0, 4) X(11, 0, 5)
#define EFTB_G13(X) \
  X(0, 2, 7) X(0, 3, 8) X(0, 3, 9) X(1, 1, 3) X(1, 2, 5) X(1, 2, 6) X(3, 1, 4) X(5, 0, 0) X(5, 1, 2) X(7, 0, 1)

struct GroupArgs {
  const double *F, *P22, *Cs, *f, *k, *l11, *lct, *lctnnlo, *l22, *l13;
  double *T, *Cr;
  int Bp, Nl, Nk, Ns, nterm, with_nnlo, ncr;
  int row_p11, row_p13, row_c11, row_cct, row_cctnnlo;
};

__global__ void __launch_bounds__(128) group_kernel(GroupArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bp) return;
  const size_t Bp = a.Bp;
  const double f1 = a.f[b];
  double fp[5];
  fp[0] = 1.0;
#pragma unroll
  for (int i = 1; i < 5; ++i) fp[i] = fp[i - 1] * f1;

  if ((int)blockIdx.y < a.Nk) {
    const int k = blockIdx.y;
    const double kv = a.k[k], k0 = a.k[0];
    const double P11k = a.F[(size_t)(a.row_p11 + k) * Bp + b], P110 = a.F[(size_t)a.row_p11 * Bp + b];
    double v22[EFTB_N22], z22[EFTB_N22], v13[EFTB_N13], z13[EFTB_N13];
#pragma unroll
    for (int i = 0; i < EFTB_N22; ++i) {
      v22[i] = a.P22[((size_t)i * a.Nk + k) * Bp + b];
      z22[i] = a.P22[((size_t)i * a.Nk) * Bp + b];
    }
    // P13[b,k] = k^3 P11(k) Re sum_n c_n k^{eta_n} M13[b,n]   (pybird.py:1080-1086)
    const double s13 = kv * kv * kv * P11k, s130 = k0 * k0 * k0 * P110;
#pragma unroll
    for (int i = 0; i < EFTB_N13; ++i) {
      v13[i] = s13 * a.F[(size_t)(a.row_p13 + i * a.Nk + k) * Bp + b];
      z13[i] = s130 * a.F[(size_t)(a.row_p13 + i * a.Nk) * Bp + b];
    }
    for (int l = 0; l < a.Nl; ++l) {
      double* out = a.T + ((size_t)(l * a.Nk + k) * a.nterm) * Bp + b;
      for (int i = 0; i < 3; ++i) out[(size_t)i * Bp] = a.l11[l * 3 + i] * P11k;
      for (int i = 0; i < 6; ++i) out[(size_t)(3 + i) * Bp] = a.lct[l * 6 + i] * (kv * kv) * P11k;
      double row[12], row0[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) row[i] = row0[i] = 0.0;
#define ACC22(r, p, t) { const double w = fp[p] * a.l22[l * EFTB_N22 + t]; row[r] += w * v22[t]; row0[r] += w * z22[t]; }
#define ACC13(r, p, t) { const double w = fp[p] * a.l13[l * EFTB_N13 + t]; row[r] += w * v13[t]; row0[r] += w * z13[t]; }
      EFTB_G22(ACC22)
      EFTB_G13(ACC13)
#undef ACC22
#undef ACC13
#pragma unroll
      for (int i = 0; i < 12; ++i) out[(size_t)(9 + i) * Bp] = row[i] - row0[i];  // pybird.py:861-866
      // stochastic basis {1, k^2} on l=0 and k^2 on l=2 (pybird.py:850-859)
      out[(size_t)21 * Bp] = (l == 0) ? 1.0 : 0.0;
      out[(size_t)22 * Bp] = (l == 0) ? kv * kv : 0.0;
      out[(size_t)23 * Bp] = (l == 1) ? kv * kv : 0.0;
      if (a.with_nnlo)
        for (int i = 0; i < 3; ++i) out[(size_t)(24 + i) * Bp] = a.lctnnlo[l * 3 + i] * (kv * kv * kv * kv) * P11k;
    }
  } else {
    const int s = blockIdx.y - a.Nk;
    // Cr is point-major [b][l][ncr][Ns]: the one-CTA-per-cosmology resummation kernel reads it contiguously
    const size_t ld = (size_t)a.Nl * a.ncr * a.Ns;
    for (int l = 0; l < a.Nl; ++l) {
      double* out = a.Cr + (size_t)b * ld + (size_t)(l * a.ncr) * a.Ns + s;
      const size_t rs = (size_t)a.Ns;
      out[0] = a.F[(size_t)(a.row_c11 + l * a.Ns + s) * Bp + b];
      out[rs] = a.F[(size_t)(a.row_cct + l * a.Ns + s) * Bp + b];
      if (a.with_nnlo) out[14 * rs] = a.F[(size_t)(a.row_cctnnlo + l * a.Ns + s) * Bp + b];
      if (!a.Cs) continue;  // Cloopl rows were produced by the grouped spectral GEMM
      const double* cs = a.Cs + ((size_t)(l * EFTB_NCH) * a.Ns + s) * Bp + b;
      const size_t crs = (size_t)a.Ns * Bp;
      double row[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) row[i] = 0.0;
#define ACC22(r, p, t) row[r] += fp[p] * a.l22[l * EFTB_N22 + t] * cs[(size_t)(t) * crs];
#define ACC13(r, p, t) row[r] += fp[p] * a.l13[l * EFTB_N13 + t] * cs[(size_t)(EFTB_N22 + t) * crs];
      EFTB_G22(ACC22)
      EFTB_G13(ACC13)
#undef ACC22
#undef ACC13
#pragma unroll
      for (int i = 0; i < 12; ++i) out[(size_t)(2 + i) * rs] = row[i];
    }
  }
}

// Dg[l][r][t2][b] = sum over the members (r, p, ch) of reducePsCfl's row r of  f^p * l22|l13[l][ch] * D[ch][t2][b]:
// the Legendre weighting and f-power grouping of the configuration-space loop terms (pybird.py:805-846) applied
// BEFORE the linear D -> C(s) transform, so that the transform runs on Nl*12 instead of Nl*38 channels.
__global__ void __launch_bounds__(128) regroup_kernel(const double* __restrict__ D, const double* __restrict__ f,
                                                      const double* __restrict__ l22, const double* __restrict__ l13, int Nl,
                                                      int nt2, int Bp, double* __restrict__ Dg) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, t2 = blockIdx.y;
  if (b >= Bp) return;
  const size_t chs = (size_t)nt2 * Bp;
  const double f1 = f[b];
  double fp[5];
  fp[0] = 1.0;
#pragma unroll
  for (int i = 1; i < 5; ++i) fp[i] = fp[i - 1] * f1;
  double v[EFTB_NCH];
  const double* src = D + (size_t)t2 * Bp + b;
#pragma unroll
  for (int c = 0; c < EFTB_NCH; ++c) v[c] = src[(size_t)c * chs];
  for (int l = 0; l < Nl; ++l) {
    double row[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) row[i] = 0.0;
#define ACC22(r, p, t) row[r] = fma(fp[p] * l22[l * EFTB_N22 + t], v[t], row[r]);
#define ACC13(r, p, t) row[r] = fma(fp[p] * l13[l * EFTB_N13 + t], v[EFTB_N22 + t], row[r]);
    EFTB_G22(ACC22)
    EFTB_G13(ACC13)
#undef ACC22
#undef ACC13
    double* dst = Dg + ((size_t)(l * 12) * nt2 + t2) * Bp + b;
#pragma unroll
    for (int i = 0; i < 12; ++i) dst[(size_t)i * chs] = row[i];
  }
}

}  // namespace

int launch_regroup(const eftb_plan* p, int Bp, const double* D, const double* f, double* Dg, cudaStream_t s) {
  const eftb_config& c = p->cfg;
  const int nt2 = 2 * (c.Nmax + 1);
  dim3 grid((Bp + 127) / 128, nt2);
  regroup_kernel<<<grid, 128, 0, s>>>(D, f, p->l22, p->l13, c.Nl, nt2, Bp, Dg);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int launch_group(const eftb_plan* p, int Bp, const double* F, const double* P22, const double* Cs, const double* f,
                 double* T, double* Cr, cudaStream_t s) {
  const eftb_config& c = p->cfg;
  GroupArgs a;
  a.F = F; a.P22 = P22; a.Cs = Cs; a.f = f; a.k = p->k; a.l11 = p->l11; a.lct = p->lct; a.lctnnlo = p->lctnnlo;
  a.l22 = p->l22; a.l13 = p->l13; a.T = T; a.Cr = Cr; a.Bp = Bp; a.Nl = c.Nl; a.Nk = c.Nk; a.Ns = c.Ns;
  a.nterm = c.nterm; a.with_nnlo = c.with_nnlo; a.ncr = 14 + (c.with_nnlo ? 1 : 0);
  a.row_p11 = c.row_p11; a.row_p13 = c.row_p13; a.row_c11 = c.row_c11; a.row_cct = c.row_cct;
  a.row_cctnnlo = c.row_cctnnlo;
  dim3 grid((Bp + 127) / 128, c.Nk + c.Ns);
  group_kernel<<<grid, 128, 0, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}
