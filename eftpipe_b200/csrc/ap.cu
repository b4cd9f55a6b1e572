// Alcock-Paczynski distortion (APeffect.AP / integrAP, pybird.py:1581-1621).
//
//   P'_l(k) = (q_perp^2 q_par)^-1 * 2 * trapz_mu[ (2l+1)/2 L_l(mu) * sum_l' spline_l'(k'(k,mu)) L_l'(mu'(k,mu)) ]
//
// The reference's `interp1d(kind="cubic")` is the not-a-knot B-spline interpolant on the fixed nodes co.k:
// spline_l'(x) = sum_j coef_l'[j] B_j(x), coef = Cinv @ values (a fixed matrix, done by the DMMA GEMM before
// these kernels).  The resampling geometry (k', mu') depends on the cosmology but NOT on which of the 24-27
// term rows is being resampled, so instead of re-evaluating the spline for every row (the reference's order,
// 15 FMA per (row, k, mu)) the mu-quadrature is folded into a per-cosmology banded operator first:
//
//   G[k][l][l'][j] = sum_mu w_l(mu) L_l'(mu'(k,mu)) B_j(k'(k,mu))        (ap_geom_kernel, one thread per (b, k))
//   P'_l(k)[row]   = norm * sum_l' sum_j G[k][l][l'][j] coef_l'[j][row]  (ap_apply_kernel, one CTA per cosmology)
//
// k'(mu) is monotone in mu, so for one k node the B-splines that are touched form a window of
// W = |j(mu=1) - j(mu=0)| + 4 consecutive j; the geometry thread keeps the 4 live columns of that window in
// registers (NL*NL x 4 accumulators), retires one column to G each time k' crosses a knot and rotates.
// Exact re-association of the reference sum: ~7x fewer FP64 operations than the row-by-row order.
#include "common.cuh"

namespace {

struct ApArgs {
  const double *coef, *Tin, *DA, *H, *k, *knot_lo, *basis, *mu, *wl;
  double *Tout, *G;
  int2* meta;       // per (b, k): first B-spline index of the window, window length
  int b0, nb;       // this launch handles cosmologies [b0, b0 + nb)
  int Bp, Nk, nterm, nmu, nint, ap_st, wcap;
  double da_fid, h_fid;
};

// Branch-free 1/sqrt(x): single-precision seed (one MUFU) and ONE Newton step with an FMA residual -> relative
// error ~1e-14.  That is six orders below the 1e-8 parity bar for everything it feeds (k', mu'^2: smooth functions),
// and it halves the dependent FP64 chain of a mu node, which is what bounds this kernel (a DFMA has ~40 cycles of
// latency on this part and there are only ~3 warps per scheduler to hide it).  CUDA's rsqrt(double) also carries a
// special-case branch + call that splits the loop body into basic blocks.
__device__ __forceinline__ double rsqrt_newton(double x) {
  float yf;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x));
  const double y = (double)yf;
  const double e = fma(-(x * y), y, 1.0);
  return fma(0.5 * y, e, y);
}

constexpr int GEOM_THREADS = 128;
constexpr int APPLY_THREADS = 256;

// Everything of the resampling geometry that depends on (cosmology, mu) only - NOT on the k node:
//   k'(k, mu) = (k / q_perp) * sqrt(root),  root = 1 + mu^2 (F^-2 - 1)                 (pybird.py:1608)
//   mu'^2 = mu^2 F^-2 / root  ->  even Legendre L_l'(mu'), times the quadrature weight w_l(mu)   (pybird.py:1609, :1595)
// AP_TAB doubles per node: sqrt(root), then w_l L_l' for (l, l').  ap_geom_kernel builds these rows in shared memory for
// the (at most GEOM_COS) cosmologies its threads belong to, one tile of mu nodes at a time.
constexpr int AP_TAB = 10;
constexpr int GEOM_MU_TILE = 200;

template <int NL>
__device__ __forceinline__ void mu_row(const ApArgs& a, double invF2, int t, double* out) {
  const double m = a.mu[t], m2 = m * m;
  const double root = fma(m2, invF2 - 1.0, 1.0);
  const double rs = rsqrt_newton(root);
  const double mp2 = m2 * invF2 * (rs * rs);
  double L[3];
  L[0] = 1.0;
  L[1] = 0.5 * (3.0 * mp2 - 1.0);
  L[2] = (35.0 * mp2 * mp2 - 30.0 * mp2 + 3.0) * 0.125;
  out[0] = root * rs;
#pragma unroll
  for (int l = 0; l < NL; ++l)
#pragma unroll
    for (int lp = 0; lp < NL; ++lp) out[1 + l * NL + lp] = a.wl[l * a.nmu + t] * L[lp];
}

__device__ __forceinline__ double ap_invF2(const ApArgs& a, int b) {
  const double qperp = a.DA[b] / a.da_fid, qpar = a.h_fid / a.H[b];  // pybird.py:1560-1561
  const double Fap = qpar / qperp;
  return 1.0 / (Fap * Fap);
}

template <int NL>
__global__ void __launch_bounds__(GEOM_THREADS, 3) ap_geom_kernel(ApArgs a) {
  constexpr int NQ = NL * NL;
  static_assert(1 + NL * NL <= AP_TAB, "mu table too narrow");
  extern __shared__ __align__(16) double sm[];
  double* knots = sm;                      // [nint]
  double* bas = knots + a.nint;            // [nint][4][4]
  double* tabs = bas + a.nint * 16;        // [ncos][tile][AP_TAB] (16-byte aligned: nint * 17 is even or padded below)
  tabs += (a.nint * 17) & 1;
  const int tid = threadIdx.x;
  for (int i = tid; i < a.nint; i += GEOM_THREADS) knots[i] = a.knot_lo[i];
  for (int i = tid; i < a.nint * 16; i += GEOM_THREADS) bas[i] = a.basis[i];
  const int ntot = a.nb * a.Nk;
  const int gid0 = blockIdx.x * GEOM_THREADS, gid = gid0 + tid;
  const bool active = gid < ntot;
  const int blA = gid0 / a.Nk, blB = (min(gid0 + GEOM_THREADS, ntot) - 1) / a.Nk, ncos = blB - blA + 1;
  const int tile = min(a.nmu, GEOM_MU_TILE);
  // fill the rows of mu nodes [t0, t0 + nt) of this CTA's cosmologies (every thread helps, also inactive ones)
  auto fill_tile = [&](int t0, int nt) {
    for (int i = tid; i < ncos * nt; i += GEOM_THREADS) {
      const int c = i / nt, t = i - c * nt;
      mu_row<NL>(a, ap_invF2(a, a.b0 + blA + c), t0 + t, tabs + ((size_t)c * tile + t) * AP_TAB);
    }
  };
  fill_tile(0, tile);
  __syncthreads();
  const int gidc = active ? gid : ntot - 1;  // inactive threads shadow the last node (no stores) so that they reach the barriers
  const int bl = gidc / a.Nk, ik = gidc - bl * a.Nk, b = a.b0 + bl;

  const double qperp = a.DA[b] / a.da_fid;  // pybird.py:1560
  const double kq = a.k[ik] / qperp;
  // this cosmology's rows; the lanes of a warp are consecutive k of (mostly) one cosmology: broadcast loads
  const double* tab = tabs + (size_t)(bl - blA) * tile * AP_TAB;

  auto locate = [&](double x) {  // largest j with knots[j] <= x, clamped (end polynomials extrapolate)
    int lo = 0, hi = a.nint - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (knots[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
  };
  double endrow[AP_TAB];
  mu_row<NL>(a, ap_invF2(a, b), a.nmu - 1, endrow);
  const int jfirst = locate(kq * tab[0]);
  const int jlast = locate(kq * endrow[0]);
  const int jlo = min(jfirst, jlast), jhi = max(jfirst, jlast);
  const int wn = jhi - jlo + 4;
  double* Grow = a.G + ((size_t)bl * a.Nk + ik) * NQ * a.wcap;
  if (active) a.meta[(size_t)bl * a.Nk + ik] = make_int2(jlo, wn);
  // k'(mu) is monotone, so the live window is only ever moved in the direction jfirst -> jlast (a k' that dips back
  // across a knot by rounding keeps its current interval: the spline is C2, the value agrees to ~1e-14).  Every
  // column of the window is therefore retired exactly once, with a plain store: no zero-fill, no atomics.
  const bool up = jlast > jfirst, down = jlast < jfirst;

  double acc[NQ][4];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[q][r] = 0.0;
  int j = jfirst;
  double bc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) bc[i] = bas[j * 16 + i];
  double knot = knots[j];

  for (int t0 = 0; t0 < a.nmu; t0 += tile) {
   const int nt = min(tile, a.nmu - t0);
   if (t0 > 0) {  // next tile of mu nodes (nmu > GEOM_MU_TILE only)
     __syncthreads();
     fill_tile(t0, nt);
     __syncthreads();
   }
   for (int t = 0; t < nt; ++t) {
    // 80 bytes per node, 16-byte aligned: sqrt(root) | w_l L_l'
    const double2* row = reinterpret_cast<const double2*>(tab + (size_t)t * AP_TAB);
    double wL[AP_TAB];
#pragma unroll
    for (int i = 0; i < AP_TAB / 2; ++i) {
      const double2 v = row[i];
      wL[2 * i] = v.x;
      wL[2 * i + 1] = v.y;
    }
    const double kp = kq * wL[0];
    while (up && j < jhi && kp >= knots[j + 1]) {
      double* g = Grow + (j - jlo);          // B-spline j has no support beyond this knot: retire its column
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (active) g[q * a.wcap] = acc[q][0];
        acc[q][0] = acc[q][1]; acc[q][1] = acc[q][2]; acc[q][2] = acc[q][3]; acc[q][3] = 0.0;
      }
      ++j;
#pragma unroll
      for (int i = 0; i < 16; ++i) bc[i] = bas[j * 16 + i];
      knot = knots[j];
    }
    while (down && j > jlo && kp < knot) {
      double* g = Grow + (j + 3 - jlo);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (active) g[q * a.wcap] = acc[q][3];
        acc[q][3] = acc[q][2]; acc[q][2] = acc[q][1]; acc[q][1] = acc[q][0]; acc[q][0] = 0.0;
      }
      --j;
#pragma unroll
      for (int i = 0; i < 16; ++i) bc[i] = bas[j * 16 + i];
      knot = knots[j];
    }
    const double x = kp - knot;
    double bv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) bv[r] = fma(fma(fma(bc[r * 4 + 3], x, bc[r * 4 + 2]), x, bc[r * 4 + 1]), x, bc[r * 4]);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[q][r] = fma(wL[1 + q], bv[r], acc[q][r]);
   }
  }
  if (!active) return;
  double* g = Grow + (j - jlo);
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) g[q * a.wcap + r] = acc[q][r];
}

constexpr int APPLY_WS = 16;  // window columns of a node staged in shared memory (wider windows: rest from global)

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

// One CTA per cosmology.  Thread = (k group, l, term row): the B-spline coefficients of the cosmology live in shared
// memory ([l'][term][j], point-major from the Cinv GEMM), and the banded-operator rows of the NEXT k node of each
// group are copied global -> shared with cp.async while the current node is contracted (double buffer), so every
// operator word is fetched once per CTA and its DRAM/L2 latency is hidden behind the previous node.
template <int NL>
__global__ void __launch_bounds__(APPLY_THREADS) ap_apply_kernel(ApArgs a) {
  constexpr int NQ = NL * NL;
  extern __shared__ __align__(16) double sm[];
  const int nslot = NL * a.nterm, ngrp = APPLY_THREADS / nslot;
  double* coefs = sm;                                                   // [NL][nterm][Nk]
  double* Gs = coefs + (size_t)NL * a.Nk * a.nterm;                     // [2][ngrp][NQ][APPLY_WS]
  int2* metas = reinterpret_cast<int2*>(Gs + (size_t)2 * ngrp * NQ * APPLY_WS);  // [Nk]
  const int bl = blockIdx.x, b = a.b0 + bl, tid = threadIdx.x;
  const size_t Bp = a.Bp;
  for (int i = tid; i < a.Nk; i += APPLY_THREADS) metas[i] = a.meta[(size_t)bl * a.Nk + i];
  {  // B-spline coefficients, point-major [b][l][term][j]: contiguous, coalesced
    const double* cb = a.coef + (size_t)b * NL * a.Nk * a.nterm;
#pragma unroll 8
    for (int i = tid; i < NL * a.Nk * a.nterm; i += APPLY_THREADS) coefs[i] = cb[i];
  }
  const double qperp = a.DA[b] / a.da_fid, qpar = a.h_fid / a.H[b];
  const double norm = 1.0 / (qperp * qperp * qpar);  // pybird.py:1611
  __syncthreads();

  const int grp = tid / nslot, slot = tid - grp * nslot;
  const bool active = grp < ngrp;
  const int l = active ? slot / a.nterm : 0, i = active ? slot - l * a.nterm : 0;
  const bool apply = a.ap_st || i < 21 || i >= 24;  // Pstl only with APst (pybird.py:1618-1619)
  const double* Gb = a.G + (size_t)bl * a.Nk * NQ * a.wcap;
  auto prefetch = [&](int ik, int buf) {
    if (active && ik < a.Nk) {
      const int wn = min(metas[ik].y, APPLY_WS);
      const double* Gk = Gb + (size_t)ik * NQ * a.wcap;
      double* dst = Gs + (size_t)((buf * ngrp + grp) * NQ) * APPLY_WS;
      for (int e = slot; e < NQ * APPLY_WS; e += nslot) {
        const int q = e / APPLY_WS, c = e - q * APPLY_WS;
        if (c < wn) cp_async8(dst + e, Gk + (size_t)q * a.wcap + c);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  prefetch(grp, 0);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  const int niter = (a.Nk + ngrp - 1) / ngrp;
  for (int it = 0; it < niter; ++it) {
    const int buf = it & 1, ik = grp + it * ngrp;
    prefetch(ik + ngrp, buf ^ 1);
    if (active && ik < a.Nk) {
      const size_t o = ((size_t)(l * a.Nk + ik) * a.nterm + i) * Bp + b;
      if (!apply) {
        a.Tout[o] = a.Tin[o];
      } else {
        const int2 mw = metas[ik];
        const double* gq = Gs + (size_t)((buf * ngrp + grp) * NQ + l * NL) * APPLY_WS;
        const double* cf = coefs + (size_t)i * a.Nk + mw.x;
        const int wn = min(mw.y, APPLY_WS);
        double acc = 0.0;
#pragma unroll
        for (int lp = 0; lp < NL; ++lp)
          for (int c = 0; c < wn; ++c) acc = fma(gq[lp * APPLY_WS + c], cf[(size_t)lp * a.nterm * a.Nk + c], acc);
        if (mw.y > APPLY_WS) {  // strong AP distortion: the columns beyond the staged window straight from global
          const double* Gk = Gb + ((size_t)ik * NQ + l * NL) * a.wcap;
          for (int c = APPLY_WS; c < mw.y; ++c)
#pragma unroll
            for (int lp = 0; lp < NL; ++lp) acc = fma(__ldg(Gk + lp * a.wcap + c), cf[(size_t)lp * a.nterm * a.Nk + c], acc);
        }
        a.Tout[o] = norm * acc;
      }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
  }
}

// cosmologies per launch: the banded operator G is stored dense in j (any window fits), bounded to ~256 MB
int ap_chunk(const eftb_config& c, int B) {
  const size_t per = (size_t)c.Nk * c.Nl * c.Nl * c.Nk;
  size_t n = ((size_t)32 << 20) / per;  // doubles
  if (n < 1) n = 1;
  return (int)(n < (size_t)B ? n : (size_t)B);
}

template <int NL>
int run(ApArgs a, int B, cudaStream_t s, int phase) {
  const int geom_cos = (GEOM_THREADS - 2 + a.Nk) / a.Nk + 1;  // cosmologies a CTA's GEOM_THREADS consecutive (b, k) nodes can touch
  const int mu_tile = a.nmu < GEOM_MU_TILE ? a.nmu : GEOM_MU_TILE;
  const size_t smem_g = sizeof(double) * (a.nint + (size_t)a.nint * 16 + 1 + (size_t)geom_cos * mu_tile * AP_TAB);
  const size_t smem_a = sizeof(double) * ((size_t)NL * a.Nk * a.nterm + (size_t)2 * (APPLY_THREADS / (NL * a.nterm)) * NL * NL * APPLY_WS) +
                        sizeof(int2) * a.Nk;
  if (NL * a.nterm > APPLY_THREADS || smem_g > 200 * 1024 || smem_a > 200 * 1024) {
    eftb_set_error("ap: unsupported sizes nmu=%d nterm=%d Nk=%d", a.nmu, a.nterm, a.Nk);
    return EFTB_ERR_ARG;
  }
  static size_t conf_g = 0, conf_a = 0;
  if (smem_g > conf_g) {
    EFTB_CUDA_CHECK(cudaFuncSetAttribute(ap_geom_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    conf_g = smem_g;
  }
  if (smem_a > conf_a) {
    EFTB_CUDA_CHECK(cudaFuncSetAttribute(ap_apply_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    conf_a = smem_a;
  }
  const int chunk = a.nb;  // capacity of the scratch, set by the caller
  for (int b0 = 0; b0 < B; b0 += chunk) {
    a.b0 = b0;
    a.nb = B - b0 < chunk ? B - b0 : chunk;
    const int nthreads = a.nb * a.Nk;
    if (phase & EFTB_PHASE_FIRST) {
      ap_geom_kernel<NL><<<(nthreads + GEOM_THREADS - 1) / GEOM_THREADS, GEOM_THREADS, smem_g, s>>>(a);
      EFTB_LAUNCH_CHECK();
    }
    if (phase & EFTB_PHASE_SECOND) {
      ap_apply_kernel<NL><<<a.nb, APPLY_THREADS, smem_a, s>>>(a);
      EFTB_LAUNCH_CHECK();
    }
  }
  return EFTB_OK;
}

}  // namespace

size_t ap_scratch_doubles(const eftb_plan* p, int B) {
  const eftb_config& c = p->cfg;
  const size_t chunk = ap_chunk(c, B);
  return chunk * c.Nk * c.Nl * c.Nl * c.Nk + chunk * c.Nk;  // G | meta (int2 = 8 bytes each)
}

int ap_chunk_count(const eftb_plan* p, int B) {
  const int chunk = ap_chunk(p->cfg, B);
  return (B + chunk - 1) / chunk;
}

int launch_ap(const eftb_plan* p, int B, int Bp, const double* coef, const double* Tin, const double* DA, const double* H,
              double* scratch, double* Tout, cudaStream_t s, int phase) {
  const eftb_config& c = p->cfg;
  if (phase != EFTB_PHASE_ALL && ap_chunk_count(p, B) != 1) {
    eftb_set_error("ap: split phases need the whole batch in one chunk");
    return EFTB_ERR_ARG;
  }
  ApArgs a;
  a.coef = coef; a.Tin = Tin; a.DA = DA; a.H = H; a.k = p->k; a.knot_lo = p->knot_lo; a.basis = p->basis; a.mu = p->mu;
  a.wl = p->wl; a.Tout = Tout; a.Bp = Bp; a.Nk = c.Nk; a.nterm = c.nterm; a.nmu = c.nmu; a.nint = c.nint;
  a.ap_st = c.ap_st; a.da_fid = c.da_fid; a.h_fid = c.h_fid; a.wcap = c.Nk;
  a.nb = ap_chunk(c, B);
  a.b0 = 0;
  a.G = scratch;
  a.meta = reinterpret_cast<int2*>(scratch + (size_t)a.nb * c.Nk * c.Nl * c.Nl * c.Nk);
  if (c.Nl == 3) return run<3>(a, B, s, phase);
  if (c.Nl == 2) return run<2>(a, B, s, phase);
  eftb_set_error("ap: unsupported Nl=%d", c.Nl);
  return EFTB_ERR_ARG;
}
