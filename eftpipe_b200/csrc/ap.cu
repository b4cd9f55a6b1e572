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
  double *Tout, *G;  // G: dense overflow operator [nb][Nk][NQ][wcap] (columns >= APPLY_WS of wide windows only)
  double* Gc;       // compact operator [nb][Nk][NL][KP]: row (k, l) = [l'][c < APPLY_WS] zero padded - what ap_apply streams
  double* Tpm;      // point-major results of ap_apply [nb][NL * Nk * nterm]: contiguous per cosmology, transposed afterwards
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
constexpr int BAS_P = 17;     // shared-memory pitch of one interval's basis polynomials
constexpr int APPLY_WS = 10;  // window columns of a node held in the compact operator (wider windows: rest in the dense overflow)
__host__ __device__ constexpr int apply_kp(int NL) { return (NL * APPLY_WS + 3) / 4 * 4; }  // K of the per-node product, padded to k4 steps
// doubles between the B-spline coefficient blocks [l'][term][j] of consecutive cosmologies: whole 16-byte units (TMA source)
__host__ __device__ inline size_t coef_stride(int NL, int nterm, int Nk) { return ((size_t)NL * nterm * Nk + 1) & ~(size_t)1; }
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

// Two threads per (cosmology, k) node: the live window has 4 columns (B-splines j .. j+3 of the current interval); lane
// half h of a pair keeps columns 2h, 2h+1 - NL*NL x 2 accumulators and 2 basis polynomials instead of x 4 and 4 (the
// one-thread form sat at 168 registers, 3 CTAs per SM, FP64 pipe 40 % busy: the serial chain load -> k' -> compare ->
// Horner of a node had only 12 warps per SM to hide behind).  When k' crosses a knot the window shifts by one column:
// the column that leaves is stored by the half that owns it, the one that changes sides moves with a shuffle.
template <int NL>
__global__ void __launch_bounds__(GEOM_THREADS, 4) ap_geom_kernel(ApArgs a) {
  constexpr int NQ = NL * NL;
  static_assert(1 + NL * NL <= AP_TAB, "mu table too narrow");
  extern __shared__ __align__(16) double sm[];
  double* knots = sm;                      // [nint]
  double* bas = knots + a.nint;            // [nint][BAS_P]: 4 x 4 coefficients per interval, pitch 17 (lanes of a warp sit in
                                           // different intervals: a pitch of 16 doubles puts all of them on the same banks)
  double* tabs = bas + a.nint * BAS_P;     // [ncos][tile][AP_TAB] (16-byte aligned: nint * 18 is even)
  const int tid = threadIdx.x;
  for (int i = tid; i < a.nint; i += GEOM_THREADS) knots[i] = a.knot_lo[i];
  for (int i = tid; i < a.nint * 16; i += GEOM_THREADS) bas[(i >> 4) * BAS_P + (i & 15)] = a.basis[i];
  constexpr int NODES = GEOM_THREADS / 2;
  const int ntot = a.nb * a.Nk;
  const int node0 = blockIdx.x * NODES, node = node0 + (tid >> 1), half = tid & 1;
  const bool active = node < ntot;
  const int blA = node0 / a.Nk, blB = (min(node0 + NODES, ntot) - 1) / a.Nk, ncos = blB - blA + 1;
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
  const int nodec = active ? node : ntot - 1;  // inactive threads shadow the last node (no stores) so that they reach the barriers
  const int bl = nodec / a.Nk, ik = nodec - bl * a.Nk, b = a.b0 + bl;

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
  double* Gcrow = a.Gc + ((size_t)bl * a.Nk + ik) * NL * apply_kp(NL);
  // column `col` of the (l, l') = q window: compact row [l][l' * APPLY_WS + col] or, beyond APPLY_WS, the dense overflow
  auto put = [&](int q, int col, double v) {
    if (col < APPLY_WS) Gcrow[(q / NL) * apply_kp(NL) + (q % NL) * APPLY_WS + col] = v;
    else Grow[q * a.wcap + col] = v;
  };
  if (active && half == 0) a.meta[(size_t)bl * a.Nk + ik] = make_int2(jlo, wn);
  // k'(mu) is monotone, so the live window is only ever moved in the direction jfirst -> jlast (a k' that dips back
  // across a knot by rounding keeps its current interval: the spline is C2, the value agrees to ~1e-14).  Every
  // column of the window is therefore retired exactly once, with a plain store: no zero-fill, no atomics.
  const bool up = jlast > jfirst, down = jlast < jfirst;

  double acc[NQ][2];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q][0] = acc[q][1] = 0.0;
  int j = jfirst;
  double bc[8];
  auto load_basis = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) bc[i] = bas[j * BAS_P + half * 8 + i];
  };
  load_basis();
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
    // both lanes of a pair take the same branches (same node): the shuffles below see their partner
    while (up && j < jhi && kp >= knots[j + 1]) {
      const unsigned m = __activemask();
      const int col = j - jlo;               // B-spline j has no support beyond this knot: column 0 leaves the window
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const double other = __shfl_xor_sync(m, acc[q][0], 1);  // half 0 receives column 2 (half 1's first)
        if (half == 0) {
          if (active) put(q, col, acc[q][0]);
          acc[q][0] = acc[q][1];
          acc[q][1] = other;
        } else {
          acc[q][0] = acc[q][1];
          acc[q][1] = 0.0;
        }
      }
      ++j;
      load_basis();
      knot = knots[j];
    }
    while (down && j > jlo && kp < knot) {
      const unsigned m = __activemask();
      const int col = j + 3 - jlo;           // column 3 leaves the window
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const double other = __shfl_xor_sync(m, acc[q][1], 1);  // half 1 receives column 1 (half 0's second)
        if (half == 1) {
          if (active) put(q, col, acc[q][1]);
          acc[q][1] = acc[q][0];
          acc[q][0] = other;
        } else {
          acc[q][1] = acc[q][0];
          acc[q][0] = 0.0;
        }
      }
      --j;
      load_basis();
      knot = knots[j];
    }
    const double x = kp - knot;
    double bv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) bv[r] = fma(fma(fma(bc[r * 4 + 3], x, bc[r * 4 + 2]), x, bc[r * 4 + 1]), x, bc[r * 4]);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int r = 0; r < 2; ++r) acc[q][r] = fma(wL[1 + q], bv[r], acc[q][r]);
   }
  }
  if (!active) return;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int r = 0; r < 2; ++r) put(q, j - jlo + 2 * half + r, acc[q][r]);
  if (half) return;
  // zero padding of the compact rows: columns [wn, APPLY_WS) of every (l, l') and the K padding of every l
  for (int c = wn; c < APPLY_WS; ++c)
#pragma unroll
    for (int q = 0; q < NQ; ++q) put(q, c, 0.0);
#pragma unroll
  for (int l = 0; l < NL; ++l)
    for (int c = NL * APPLY_WS; c < apply_kp(NL); ++c) Gcrow[l * apply_kp(NL) + c] = 0.0;
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int APPLY_SLACK = 16;  // zeroed doubles after the coefficient block: windows may run past the last row's end

// One CTA per cosmology.  Its compact banded operator (Nk x NL rows of K = NL * APPLY_WS columns, written by ap_geom) and
// its B-spline coefficients ([l'][term][j], point-major from the Cinv GEMM) are both contiguous in global memory: two TMA
// bulk copies (cp.async.bulk + mbarrier complete_tx) bring them to shared memory - no per-element load instructions
// (an earlier cp.async version was bound by L1TEX request traffic).  Then one warp per k node evaluates
//   out[l][term] = sum_{(l', c)} G[k][l][(l', c)] * coef[l'][term][jlo(k) + c]
// as an (8 x K)(K x 8 NT) DMMA product: rows l (NL of 8 used; the FP64 pipe is not the bound here, shared-memory and
// issue traffic are, and the fragment form needs 3x fewer of both than per-output dot products).  Zero operator columns
// beyond a window's length make out-of-window coefficient reads harmless (they stay inside the block + slack).
template <int NL, int NT>  // NT: 8-term tiles, nterm <= 8 NT
__global__ void __launch_bounds__(APPLY_THREADS, 3) ap_apply_kernel(ApArgs a) {
  constexpr int NQ = NL * NL, KK = NL * APPLY_WS, KP = apply_kp(NL), NKS = KP / 4;
  constexpr int NT_MAX = NT;
  extern __shared__ __align__(128) double sm[];
  const int nslot = NL * a.nterm;
  const uint32_t gbytes = (uint32_t)((size_t)a.Nk * NL * KP * sizeof(double)), cbytes = (uint32_t)(coef_stride(NL, a.nterm, a.Nk) * sizeof(double));
  double* Gs = sm;                                                      // [Nk][NL][KP]
  double* coefs = Gs + (size_t)a.Nk * NL * KP;                          // [NL][nterm][Nk] + slack
  int2* metas = reinterpret_cast<int2*>(coefs + coef_stride(NL, a.nterm, a.Nk) + APPLY_SLACK);  // [Nk]
  uint64_t* bar = reinterpret_cast<uint64_t*>(metas + a.Nk);
  const int bl = blockIdx.x, b = a.b0 + bl, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Gb = a.G + (size_t)bl * a.Nk * NQ * a.wcap;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(gbytes + cbytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(Gs)),
                 "l"(a.Gc + (size_t)bl * a.Nk * NL * KP), "r"(gbytes), "r"(smem_u32(bar))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(coefs)),
                 "l"(a.coef + (size_t)b * coef_stride(NL, a.nterm, a.Nk)), "r"(cbytes), "r"(smem_u32(bar))
                 : "memory");
  }
  for (int i = tid; i < a.Nk; i += APPLY_THREADS) metas[i] = a.meta[(size_t)bl * a.Nk + i];
  // the slack after the block; with an odd block the copy also brings one pad word, which is never written by the GEMM: zero it
  if (tid < APPLY_SLACK) coefs[coef_stride(NL, a.nterm, a.Nk) + tid] = 0.0;
  const double qperp = a.DA[b] / a.da_fid, qpar = a.h_fid / a.H[b];
  const double norm = 1.0 / (qperp * qperp * qpar);  // pybird.py:1611
  __syncthreads();  // metas, slack and the barrier initialisation are visible to every thread
  {
    uint32_t ok;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    } while (!ok);
  }
  if ((nslot * a.Nk) & 1) {  // odd block: the copy brought one pad word the GEMM never writes; out-of-window reads may touch it
    if (tid == 0) coefs[(size_t)nslot * a.Nk] = 0.0;
    __syncthreads();
  }

  // fragment roles: A[row = r][col = c4] = G row of multipole l = r; B[row = c4][col = r] = coefficient of term 8 t + r.
  // Everything that does not depend on the node is hoisted: K-index offsets into the coefficient block, term offsets,
  // output offsets and flags.  Padded K columns and padded term columns read valid (finite) coefficients: they meet zero
  // operator columns or land in accumulator columns that are never stored.
  const int r = lane >> 2, c4 = lane & 3;
  const int lstride = a.nterm * a.Nk;
  int boff[NKS];  // K index kappa = 4 s + c4  ->  (l', c)  ->  l' * lstride + c
#pragma unroll
  for (int s = 0; s < NKS; ++s) {
    const int kap = 4 * s + c4, lp = kap / APPLY_WS;
    boff[s] = kap < KK ? lp * lstride + (kap - lp * APPLY_WS) : 0;
  }
  int ioff[NT_MAX];          // B operand: term 8 t + r (clamped)
  unsigned store = 0;  // per (t, h): result stored; the Pstl rows pass through unchanged without APst (pybird.py:1618-1619): ap_transpose_kernel
#pragma unroll
  for (int t = 0; t < NT_MAX; ++t) {
    ioff[t] = min(8 * t + r, a.nterm - 1) * a.Nk;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 8 * t + 2 * c4 + h;
      const bool valid = i < a.nterm && r < NL;
      const bool st = !a.ap_st && i >= 21 && i < 24;
      if (valid && !st) store |= 1u << (2 * t + h);
    }
  }
  const double* grow0 = Gs + (size_t)(r < NL ? r : 0) * KP + c4;
  // results go to a point-major block [NL * Nk][nterm] of this cosmology: the terms 2 c4, 2 c4 + 1 of a tile are adjacent, so
  // a lane stores 16 bytes and the four lanes of a row fill whole sectors (the batch-minor array took 3600 scattered 8-byte
  // stores per cosmology: L1/TEX 91 % busy); ap_transpose_kernel brings them to [row][Bp] and passes the un-resampled rows through
  double* tpm = a.Tpm + (size_t)bl * NL * a.Nk * a.nterm;
  const bool vec2 = (a.nterm & 1) == 0;
  auto put_terms = [&](double* node, const double (&acc)[NT_MAX][2], const unsigned flags) {
#pragma unroll
    for (int t = 0; t < NT_MAX; ++t) {
      const unsigned f2 = (flags >> (2 * t)) & 3u;
      double* o = node + 8 * t + 2 * c4;
      if (f2 == 3u && vec2) *reinterpret_cast<double2*>(o) = make_double2(norm * acc[t][0], norm * acc[t][1]);
      else {
        if (f2 & 1u) o[0] = norm * acc[t][0];
        if (f2 & 2u) o[1] = norm * acc[t][1];
      }
    }
  };
  // one k node alone: rows l of the m8 tile (NL of 8 used) - the fallback for nodes that cannot be paired
  auto single = [&](const int ik) {
    const int2 mw = metas[ik];
    const double* grow = grow0 + (size_t)ik * NL * KP;
    const double* cbase = coefs + mw.x;
    double acc[NT_MAX][2];
#pragma unroll
    for (int t = 0; t < NT_MAX; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
      const double af = r < NL ? grow[4 * s] : 0.0;
#pragma unroll
      for (int t = 0; t < NT_MAX; ++t) dmma884(acc[t][0], acc[t][1], af, cbase[ioff[t] + boff[s]]);
    }
    // C[row = r (multipole)][col = 2 c4 + h (term in tile)]
    double* node = tpm + (size_t)(r * a.Nk + ik) * a.nterm;
    if (mw.y > APPLY_WS && r < NL) {  // strong AP distortion: the columns beyond the compact window from the dense overflow
      const double* Gk = Gb + ((size_t)ik * NQ + r * NL) * a.wcap;
#pragma unroll
      for (int t = 0; t < NT_MAX; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!((store >> (2 * t + h)) & 1u)) continue;
          const double* cg = a.coef + (size_t)b * coef_stride(NL, a.nterm, a.Nk) + (size_t)(8 * t + 2 * c4 + h) * a.Nk + mw.x;
          for (int c = APPLY_WS; c < mw.y; ++c)
#pragma unroll
            for (int lp = 0; lp < NL; ++lp)
              acc[t][h] = fma(__ldg(Gk + lp * a.wcap + c), __ldg(cg + (size_t)lp * lstride + c), acc[t][h]);
        }
    }
    put_terms(node, acc, store);
  };
  // Two neighbouring k nodes share one m8 tile (rows 0..NL-1: node A, rows 4..4+NL-1: node B) whenever both windows fit
  // the compact width from the smaller of the two first B-spline indices: the coefficient (B) fragments are then common
  // and node X reads its operator row shifted by jlo(X) - jmin columns.  Half the DMMAs and half the B-fragment loads.
  const int pl = r & 3, pn = r >> 2;             // multipole and node-in-pair of this lane's A / C row
  unsigned pstore = 0;
#pragma unroll
  for (int t = 0; t < NT_MAX; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 8 * t + 2 * c4 + h;
      const bool valid = i < a.nterm && pl < NL;
      const bool st = !a.ap_st && i >= 21 && i < 24;
      if (valid && !st) pstore |= 1u << (2 * t + h);
    }
  const int npair = (a.Nk + 1) / 2;
  for (int ip = warp; ip < npair; ip += APPLY_THREADS / 32) {
    const int ikA = 2 * ip, ikB = ikA + 1;
    if (ikB >= a.Nk) { single(ikA); continue; }
    const int2 mA = metas[ikA], mB = metas[ikB];
    const int jmin = min(mA.x, mB.x), shA = mA.x - jmin, shB = mB.x - jmin;
    if (mA.y + shA > APPLY_WS || mB.y + shB > APPLY_WS) { single(ikA); single(ikB); continue; }
    const int ikX = pn ? ikB : ikA, sh = pn ? shB : shA;
    const double* grow = Gs + ((size_t)ikX * NL + (pl < NL ? pl : 0)) * KP;
    const double* cbase = coefs + jmin;
    double acc[NT_MAX][2];
#pragma unroll
    for (int t = 0; t < NT_MAX; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
      const int kap = 4 * s + c4, lp = kap / APPLY_WS, c = kap - lp * APPLY_WS - sh;  // this node's own window column
      const double af = (pl < NL && kap < KK && c >= 0) ? grow[lp * APPLY_WS + c] : 0.0;
#pragma unroll
      for (int t = 0; t < NT_MAX; ++t) dmma884(acc[t][0], acc[t][1], af, cbase[ioff[t] + boff[s]]);
    }
    put_terms(tpm + (size_t)(pl * a.Nk + ikX) * a.nterm, acc, pstore);
  }
}

// Tpm[nb][R] (row = (l, k, term)) -> Tout[R][Bp], columns b0 .. b0 + nb; rows of terms that are not resampled (Pstl without
// APst, pybird.py:1618-1619) are copied from Tin instead; in the last chunk the padded lanes b >= B replicate point B - 1
__global__ void __launch_bounds__(256) ap_transpose_kernel(ApArgs a, int R, int B) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;  // column block relative to b0, row block
  const bool last = a.b0 + a.nb >= B;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = min(c0 + j, a.nb - 1), r = r0 + threadIdx.x;
    tile[j][threadIdx.x] = r < R ? a.Tpm[(size_t)c * R + r] : 0.0;
  }
  __syncthreads();
  const int ncol = last ? a.Bp - a.b0 : a.nb;  // the last chunk also fills the padding
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r >= R || c >= ncol) continue;
    const int term = r % a.nterm;
    const size_t o = (size_t)r * a.Bp + a.b0 + c;
    const bool pass = !a.ap_st && term >= 21 && term < 24;
    a.Tout[o] = pass ? a.Tin[(size_t)r * a.Bp + min(a.b0 + c, B - 1)] : tile[threadIdx.x][j];
  }
}

// cosmologies per launch: the overflow operator G is stored dense in j (any window fits) and only ever touched for windows
// wider than APPLY_WS, i.e. it is address space, not traffic: bounded to 2 GB so that a config-4 shard (8192 points) is
// ONE launch of each kernel (6 chunks of 1491 points ran 1.3 waves each: a third of every geometry launch was tail)
int ap_chunk(const eftb_config& c, int B) {
  const size_t per = (size_t)c.Nk * c.Nl * c.Nl * c.Nk;
  size_t n = ((size_t)256 << 20) / per;  // doubles
  if (n < 1) n = 1;
  return (int)(n < (size_t)B ? n : (size_t)B);
}

template <int NL>
int run(ApArgs a, int B, cudaStream_t s, int phase) {
  const int geom_cos = (GEOM_THREADS / 2 - 2 + a.Nk) / a.Nk + 1;  // cosmologies a CTA's GEOM_THREADS / 2 consecutive (b, k) nodes can touch
  const int mu_tile = a.nmu < GEOM_MU_TILE ? a.nmu : GEOM_MU_TILE;
  const size_t smem_g = sizeof(double) * (a.nint + (size_t)a.nint * BAS_P + (size_t)geom_cos * mu_tile * AP_TAB);
  const size_t smem_a = sizeof(double) * (coef_stride(NL, a.nterm, a.Nk) + APPLY_SLACK + (size_t)a.Nk * NL * apply_kp(NL)) + sizeof(int2) * a.Nk + 16;
  if (a.nterm > 32 || smem_g > 200 * 1024 || smem_a > 200 * 1024) {
    eftb_set_error("ap: unsupported sizes nmu=%d nterm=%d Nk=%d", a.nmu, a.nterm, a.Nk);
    return EFTB_ERR_ARG;
  }
  static DeviceSmem conf_g, conf_a3, conf_a4;
  EFTB_SET_SMEM(conf_g, ap_geom_kernel<NL>, smem_g);
  EFTB_SET_SMEM(conf_a3, (ap_apply_kernel<NL, 3>), smem_a);
  EFTB_SET_SMEM(conf_a4, (ap_apply_kernel<NL, 4>), smem_a);
  const int chunk = a.nb;  // capacity of the scratch, set by the caller
  for (int b0 = 0; b0 < B; b0 += chunk) {
    a.b0 = b0;
    a.nb = B - b0 < chunk ? B - b0 : chunk;
    const int nthreads = a.nb * a.Nk * 2;  // two lanes per node
    if (phase & EFTB_PHASE_FIRST) {
      ap_geom_kernel<NL><<<(nthreads + GEOM_THREADS - 1) / GEOM_THREADS, GEOM_THREADS, smem_g, s>>>(a);
      EFTB_LAUNCH_CHECK();
    }
    if (phase & EFTB_PHASE_SECOND) {
      if (a.nterm <= 24) ap_apply_kernel<NL, 3><<<a.nb, APPLY_THREADS, smem_a, s>>>(a);
      else ap_apply_kernel<NL, 4><<<a.nb, APPLY_THREADS, smem_a, s>>>(a);
      EFTB_LAUNCH_CHECK();
      const int R = NL * a.Nk * a.nterm;
      const int ncol = b0 + a.nb >= B ? a.Bp - b0 : a.nb;
      ap_transpose_kernel<<<dim3((ncol + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(a, R, B);
      EFTB_LAUNCH_CHECK();
    }
  }
  return EFTB_OK;
}

}  // namespace

size_t ap_scratch_doubles(const eftb_plan* p, int B) {
  const eftb_config& c = p->cfg;
  const size_t chunk = ap_chunk(c, B);
  // dense overflow G | meta (int2 = 8 bytes each) | pad to 16 bytes | compact operator Gc
  // dense overflow G | meta (int2 = 8 bytes each) | pad | compact operator Gc | pad | point-major results Tpm
  const size_t n = chunk * c.Nk * c.Nl * c.Nl * c.Nk + chunk * c.Nk + 1 + chunk * c.Nk * c.Nl * apply_kp(c.Nl) + 1 +
                   chunk * c.Nl * c.Nk * c.nterm;
  return (n + 1) & ~(size_t)1;  // whole 16-byte units: the buffers that follow in the workspace are TMA sources too
}

size_t ap_coef_doubles(const eftb_plan* p, int Bp) { return coef_stride(p->cfg.Nl, p->cfg.nterm, p->cfg.Nk) * (size_t)Bp; }

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
  {
    const size_t off = (size_t)a.nb * c.Nk * c.Nl * c.Nl * c.Nk + (size_t)a.nb * c.Nk;
    a.Gc = scratch + off + (off & 1);  // 16-byte aligned (the scratch base is): TMA bulk copies read it
    const size_t off2 = (size_t)(a.Gc - scratch) + (size_t)a.nb * c.Nk * c.Nl * apply_kp(c.Nl);
    a.Tpm = scratch + off2 + (off2 & 1);  // 16-byte aligned: double2 stores
  }
  if ((phase & EFTB_PHASE_SECOND) && ((reinterpret_cast<uintptr_t>(coef) | reinterpret_cast<uintptr_t>(a.Gc)) & 15)) {
    eftb_set_error("ap: the scratch / coefficient buffers must be 16-byte aligned (TMA sources)");
    return EFTB_ERR_ARG;
  }
  if (c.Nl == 3) return run<3>(a, B, s, phase);
  if (c.Nl == 2) return run<2>(a, B, s, phase);
  eftb_set_error("ap: unsupported Nl=%d", c.Nl);
  return EFTB_ERR_ARG;
}
