// Alcock-Paczynski distortion (APeffect.AP / integrAP, pybird.py:1581-1621).
//
//   P'_l(k) = (q_perp^2 q_par)^-1 * 2 * trapz_mu[ (2l+1)/2 L_l(mu) * sum_l' spline_l'(k'(k,mu)) L_l'(mu'(k,mu)) ]
//
// The reference's `interp1d(kind="cubic")` is the not-a-knot B-spline interpolant on the fixed nodes co.k:
// its coefficient vector is a fixed matrix times the values (done by the DMMA GEMM before this kernel), and
// on knot interval j the value is sum_r coef[j+r] * (cubic basis polynomial r of interval j)(k' - knot_j).
// One CTA per cosmology; per k node: phase 1 evaluates, once per mu node, the interval, the four basis values
// and L_l'(mu'), phase 2 contracts them with the coefficients of all term rows (thread = term row x mu slice;
// coefficients stay in registers while the interval index does not change, which it rarely does because k'
// sweeps a few per cent around k), phase 3 reduces the mu slices.  This is the only stage whose resampling
// abscissae depend on the cosmology.
#include "common.cuh"

namespace {

struct ApArgs {
  const double *coef, *Tin, *DA, *H, *k, *knot_lo, *basis, *mu, *wl;
  double* Tout;
  int B, Bp, Nk, nterm, nmu, nint, ap_st;
  double da_fid, h_fid;
};

template <int NL>
__global__ void __launch_bounds__(256) ap_kernel(ApArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int nsl = 256 / a.nterm, mps = (a.nmu + nsl - 1) / nsl;
  double* coefs = sm;                                   // [NL][Nk][nterm]
  double* pt = coefs + (size_t)NL * a.Nk * a.nterm;     // [nmu][4*NL]
  double* red = pt + (size_t)a.nmu * 4 * NL;            // [nsl][NL][nterm]
  double* knots = red + (size_t)nsl * NL * a.nterm;     // [nint]
  double* bas = knots + a.nint;                         // [nint][4][4]
  double* wls = bas + (size_t)a.nint * 16;              // [NL][nmu]
  double* mus = wls + (size_t)NL * a.nmu;               // [nmu]
  int* pj = reinterpret_cast<int*>(mus + a.nmu);        // [nmu]
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t Bp = a.Bp;

  for (int i = tid; i < NL * a.Nk * a.nterm; i += 256) coefs[i] = a.coef[(size_t)i * Bp + b];
  for (int i = tid; i < a.nint; i += 256) knots[i] = a.knot_lo[i];
  for (int i = tid; i < a.nint * 16; i += 256) bas[i] = a.basis[i];
  for (int i = tid; i < NL * a.nmu; i += 256) wls[i] = a.wl[i];
  for (int i = tid; i < a.nmu; i += 256) mus[i] = a.mu[i];
  const double qperp = a.DA[b] / a.da_fid, qpar = a.h_fid / a.H[b];  // pybird.py:1560-1561
  const double Fap = qpar / qperp;
  const double iF2m1 = 1.0 / (Fap * Fap) - 1.0;
  const double norm = 1.0 / (qperp * qperp * qpar);
  __syncthreads();

  const int ti = tid % a.nterm, sl = tid / a.nterm;
  for (int ik = 0; ik < a.Nk; ++ik) {
    // ---- phase 1: geometry of every mu node ------------------------------------------------------
    if (tid < a.nmu) {
      const double m = mus[tid];
      const double root = 1.0 + m * m * iF2m1;
      const double sq = sqrt(root);
      const double kp = a.k[ik] / qperp * sq;   // pybird.py:1608
      const double mup = m / Fap / sq;          // pybird.py:1609
      int lo = 0, hi = a.nint - 1;              // largest j with knots[j] <= kp, clamped (extrapolation)
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (knots[mid] <= kp) lo = mid; else hi = mid - 1;
      }
      const double x = kp - knots[lo];
      const double* bj = bas + lo * 16;
      double bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) bv[r] = fma(fma(fma(bj[r * 4 + 3], x, bj[r * 4 + 2]), x, bj[r * 4 + 1]), x, bj[r * 4]);
      const double m2 = mup * mup;
      double L[3];
      L[0] = 1.0;
      L[1] = 0.5 * (3.0 * m2 - 1.0);
      L[2] = (35.0 * m2 * m2 - 30.0 * m2 + 3.0) * 0.125;
#pragma unroll
      for (int lp = 0; lp < NL; ++lp)
#pragma unroll
        for (int r = 0; r < 4; ++r) pt[(size_t)tid * 4 * NL + lp * 4 + r] = L[lp] * bv[r];
      pj[tid] = lo;
    }
    __syncthreads();
    // ---- phase 2: contract with the spline coefficients of term row ti over this thread's mu slice ---
    if (sl < nsl) {
      double acc[NL], cf[4 * NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) acc[l] = 0.0;
      int jc = -1;
      const int t1 = min(a.nmu, (sl + 1) * mps);
      for (int t = sl * mps; t < t1; ++t) {
        const int j = pj[t];
        if (j != jc) {
#pragma unroll
          for (int lp = 0; lp < NL; ++lp)
#pragma unroll
            for (int r = 0; r < 4; ++r) cf[lp * 4 + r] = coefs[((size_t)lp * a.Nk + j + r) * a.nterm + ti];
          jc = j;
        }
        const double* w = pt + (size_t)t * 4 * NL;
        double val = 0.0;
#pragma unroll
        for (int q = 0; q < 4 * NL; ++q) val = fma(w[q], cf[q], val);
#pragma unroll
        for (int l = 0; l < NL; ++l) acc[l] = fma(wls[l * a.nmu + t], val, acc[l]);
      }
#pragma unroll
      for (int l = 0; l < NL; ++l) red[((size_t)sl * NL + l) * a.nterm + ti] = acc[l];
    }
    __syncthreads();
    // ---- phase 3: reduce the slices, normalise, store --------------------------------------------
    if (tid < NL * a.nterm) {
      const int l = tid / a.nterm, i = tid % a.nterm;
      const size_t o = ((size_t)(l * a.Nk + ik) * a.nterm + i) * Bp + b;
      const bool apply = a.ap_st || i < 21 || i >= 24;  // Pstl only with APst (pybird.py:1618-1619)
      if (apply) {
        double v = 0.0;
        for (int s = 0; s < nsl; ++s) v += red[((size_t)s * NL + l) * a.nterm + i];
        a.Tout[o] = norm * v;
      } else {
        a.Tout[o] = a.Tin[o];
      }
    }
    // no barrier needed here: phase 1 of the next node writes pt/pj only, which phase 3 does not read, and
    // the barrier after that phase 1 orders this phase 3 before the next phase 2's writes to red
  }
}

template <int NL>
int run(const ApArgs& a, cudaStream_t s) {
  const int nsl = 256 / a.nterm;
  size_t n = (size_t)NL * a.Nk * a.nterm + (size_t)a.nmu * 4 * NL + (size_t)nsl * NL * a.nterm + a.nint + (size_t)a.nint * 16 +
             (size_t)NL * a.nmu + a.nmu;
  size_t smem = n * sizeof(double) + sizeof(int) * a.nmu + 16;
  if (a.nmu > 256 || NL * a.nterm > 256 || smem > 200 * 1024) {
    eftb_set_error("ap: unsupported sizes nmu=%d nterm=%d smem=%zu", a.nmu, a.nterm, smem);
    return EFTB_ERR_ARG;
  }
  static size_t configured = 0;
  if (smem > configured) {
    EFTB_CUDA_CHECK(cudaFuncSetAttribute(ap_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  ap_kernel<NL><<<a.B, 256, smem, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

}  // namespace

int launch_ap(const eftb_plan* p, int B, int Bp, const double* coef, const double* Tin, const double* DA, const double* H,
              double* Tout, cudaStream_t s) {
  const eftb_config& c = p->cfg;
  ApArgs a;
  a.coef = coef; a.Tin = Tin; a.DA = DA; a.H = H; a.k = p->k; a.knot_lo = p->knot_lo; a.basis = p->basis; a.mu = p->mu;
  a.wl = p->wl; a.Tout = Tout; a.B = B; a.Bp = Bp; a.Nk = c.Nk; a.nterm = c.nterm; a.nmu = c.nmu; a.nint = c.nint;
  a.ap_st = c.ap_st; a.da_fid = c.da_fid; a.h_fid = c.h_fid;
  if (c.Nl == 3) return run<3>(a, s);
  if (c.Nl == 2) return run<2>(a, s);
  eftb_set_error("ap: unsupported Nl=%d", c.Nl);
  return EFTB_ERR_ARG;
}
