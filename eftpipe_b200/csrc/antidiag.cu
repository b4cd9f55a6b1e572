// Anti-diagonal reduction of the one-loop kernels.
//
// The reference contracts  sum_{n,m} c_n c_m M_b[n,m] x^{eta_n + eta_m}  separately for every k and s node
// (pybird.py:1074-1078, :1103-1125): 28 (+30) complex 257x257 quadratic forms per node.  Because
// eta_n + eta_m depends on n+m only, and so does the Bessel factor Ml[l,n,m] (pybird.py:1035-1038), all
// of them follow from
//        D_ch[t] = sum_{n+m=t} c_n c_m M_ch[n,m],      t = 0 .. 2 Nmax,
// which is what this kernel computes (Hermitian half t <= Nmax, symmetric pairs n <= m folded into the
// table on the host).  One lane = one cosmology, one warp = 32 cosmologies, 76 FP64 accumulators (38
// complex channels: 28 22-type, 10 13-type) per lane; the pair table is read with warp-uniform 16-byte
// loads (one L1 transaction per warp) and the coefficients with coalesced loads over the batch.
// Work is FP64-FMA bound: 4 + 38*4 DFMA per (pair, cosmology), 16641 pairs at NFFT=256.
#include "common.cuh"

namespace {

constexpr int AD_WARPS = 4;

__global__ void __launch_bounds__(AD_WARPS * 32) antidiag_kernel(const double* __restrict__ cre, const double* __restrict__ cim,
                                                                const double2* __restrict__ table,
                                                                const int32_t* __restrict__ offsets, int Nmax, int Bp,
                                                                int npair, double* __restrict__ D) {
  const int lane_b = (blockIdx.x * AD_WARPS + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31);
  const bool live = lane_b < Bp;
  const int b = live ? lane_b : Bp - 1;
  const int Nh = Nmax >> 1;
  // this block's range of anti-diagonals, balanced by pair count
  const long lo = (long)npair * blockIdx.y / gridDim.y, hi = (long)npair * (blockIdx.y + 1) / gridDim.y;
  int t0 = 0, t1 = 0;
  {
    // first t with offsets[t] >= lo  (offsets is increasing, offsets[0] = 0, offsets[Nmax+1] = npair)
    int a = 0, z = Nmax + 1;
    while (a < z) { int m = (a + z) >> 1; if (offsets[m] >= lo) z = m; else a = m + 1; }
    t0 = a;
    a = 0; z = Nmax + 1;
    while (a < z) { int m = (a + z) >> 1; if (offsets[m] >= hi) z = m; else a = m + 1; }
    t1 = a;
  }
  const size_t tstride = (size_t)2 * Bp, chstride = (size_t)(Nmax + 1) * 2 * Bp;
  for (int t = t0; t < t1; ++t) {
    double ar[EFTB_NCH], ai[EFTB_NCH];
#pragma unroll
    for (int c = 0; c < EFTB_NCH; ++c) ar[c] = ai[c] = 0.0;
    const double2* row = table + (size_t)offsets[t] * EFTB_NCH;
    const int np = (t >> 1) + 1;
    for (int p = 0; p < np; ++p) {
      const int m = t - p;
      const int mi = m <= Nh ? m : Nmax - m;  // c_m = conj(c_{Nmax-m}) for m > Nmax/2
      const double xr = cre[(size_t)p * Bp + b], xi = cim[(size_t)p * Bp + b];
      const double yr = cre[(size_t)mi * Bp + b];
      double yi = cim[(size_t)mi * Bp + b];
      yi = m <= Nh ? yi : -yi;
      const double pr = xr * yr - xi * yi, pi = xr * yi + xi * yr;
      const double2* mrow = row + (size_t)p * EFTB_NCH;
#pragma unroll
      for (int c = 0; c < EFTB_NCH; ++c) {
        const double2 mv = __ldg(mrow + c);
        ar[c] = fma(pr, mv.x, ar[c]);
        ar[c] = fma(-pi, mv.y, ar[c]);
        ai[c] = fma(pr, mv.y, ai[c]);
        ai[c] = fma(pi, mv.x, ai[c]);
      }
    }
    if (live) {
      double* out = D + (size_t)t * tstride + b;
#pragma unroll
      for (int c = 0; c < EFTB_NCH; ++c) {
        out[(size_t)c * chstride] = ar[c];
        out[(size_t)c * chstride + Bp] = ai[c];
      }
    }
  }
}

}  // namespace

int launch_antidiag(const eftb_plan* p, int Bp, const double* F, double* D, cudaStream_t s) {
  const eftb_config& c = p->cfg;
  const double* cre = F + (size_t)c.row_cre * Bp;
  const double* cim = F + (size_t)c.row_cim * Bp;
  int nbx = (Bp / 32 + AD_WARPS - 1) / AD_WARPS;
  // enough t-chunks to put >= ~3 CTAs on each of the 148 SMs, at most one chunk per anti-diagonal pair of rows
  int want = (3 * 148 + nbx - 1) / nbx;
  int nchunk = want < 1 ? 1 : (want > (c.Nmax + 1) / 2 ? (c.Nmax + 1) / 2 : want);
  dim3 grid(nbx, nchunk);
  antidiag_kernel<<<grid, AD_WARPS * 32, 0, s>>>(cre, cim, p->pair_table, p->pair_offsets, c.Nmax, Bp, c.npair, D);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}
