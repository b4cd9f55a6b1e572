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
// Most of those polynomials vanish identically (the X^{p+1} terms of Q^{ll'} couple to a single Bessel order
// v, the Y X^p terms to 2-3): plan creation keeps, per l', the <= 4 non-zero (kind, v) "slots" (11 of 18
// polynomials at Nl=3).  One CTA (128 threads) per cosmology and a: Q(f) is expanded once (resum_q_kernel) and staged
// in shared memory with the rows C[l',i,s] by TMA bulk copies.  a = 1 (counterterm + loop rows), default form
// (resum_body_mma): a warp owns 8 output columns (l,k) and sweeps s in passes of 16; every lane runs 3-4 slots x
// 4 points of Horner chains fed by broadcast 16-byte shared loads, and its 4 results are the A fragment of
// mma.m8n8k4.f64, so the contraction with the 13 (14) rows runs on DMMA.  The scalar form (resum_body: thread = one
// (l,k) column, s in chunks of 4, 13 + 3 accumulators per thread) remains as the fallback for unaligned / odd row
// sizes and sweeps the columns left over by the groups of 8 (3 x 43 = 129 = 16 x 8 + 1), one warp per column with the
// s-chunks spread over its lanes and a shuffle reduction.  FP64 bound: ~2.3e6 DFMA per cosmology at Nl=3 for a = 1.
//
// The a = 0 half (linear terms) contracts with ONE row per l' (C11), so there the separable form of the reference
// itself (pybird.py:1409-1441) is 3x cheaper than the sweep:  z^p = k^{2p} X(s)^p, hence
//   IR0[l,l',k] = sum_{slot,p} Q_0[l,l',p,slot] k^{2(p+1)} G_slot[k,p],   G_slot[k,p] = sum_s R_v[k,s] W_slot[p,s],
//   W_(l',X)[p,s] = X^{p+1} C11_l',  W_(l',Y,v)[p,s] = Y X^p C11_l'
// - per slot a (Nkr x Ns)(Ns x NIR) product with a FIXED left operand: DMMA m8n8k4 with the B fragments built on the
// fly as base_slot[s] * X(s)^p from two small shared-memory tables (resum_linear_body); the fold over p with
// Q_0 k^{2(p+1)} is a second, small DMMA product fed straight from the C fragments of the first.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace {

#ifndef EFTB_RS_THREADS
#define EFTB_RS_THREADS 128
#endif
constexpr int RS_THREADS = EFTB_RS_THREADS;
constexpr int RS_C = 4;      // s points per chunk
constexpr int RS_SLOTS = 4;  // polynomial slots per l'

struct ResumArgs {
  const double *F, *Cr, *f, *Rt, *Rk, *qpack, *kr2, *l11, *lct, *lctnnlo;
  double *T, *Qf;   // Qf: [B][NQ] expanded Q(f) (scratch)
  int B, Bp, Nk, Ns, NsP, nterm, ncr, Nkr, Nklow, qdeg, row_x, row_y, NQ, KPAD;
  int nslot[3];     // canonical slots of l': 0 = (X, v = l'), 1 + v = (Y, v); nslot = 1 + number of Y orders used
  int nslots;       // every non-zero polynomial as (l', canonical slot)
  int mma;          // 1: row contraction of the a = 1 half on DMMA (resum_body_mma), 0: scalar sweep (resum_body)
  int mix;          // CTA order: 0 = all a = 1 halves, then all a = 0 halves; k > 0 = one a = 0 CTA after every k a = 1 CTAs
  signed char slot_lp[12], slot_s[12];
};

constexpr int RL_MCH = 3;  // m8 tiles (k rows) per warp task of the linear-term GEMM
#ifndef EFTB_RS_MC
#define EFTB_RS_MC 4
#endif
#ifndef EFTB_RS_MINB
#define EFTB_RS_MINB 4
#endif
constexpr int RS_MINB = EFTB_RS_MINB;  // resident CTAs per SM the kernel is compiled for
constexpr int RS_MC = EFTB_RS_MC;      // quads (of 4 s points, the m8n8k4 K dimension) per warp pass of the DMMA form
constexpr int RS_PASS = 4 * RS_MC;     // s points per warp pass; NsP is a multiple of 16

__host__ __device__ inline int rl_pitch(int NsP) { return NsP + ((4 - NsP % 16) + 16) % 16; }  // = 4 mod 16: conflict-free B fragments

// accumulators of one (l, k) output column: IA = 0 (linear terms, Q_0, contracted with C11) keeps one sum per l';
// IA = 1 (Q_1) keeps Cct per l', the 12 loop rows and, with NNLO, CctNNLO per l'
// which cosmology / half a CTA works on.  mix = 0: blocks [0, B) are the a = 1 halves, [B, 2B) the a = 0 halves (the block
// scheduler hands out the heavier halves first and the light ones fill the tail).  mix = k: the halves alternate in groups -
// k a = 1 CTAs, then the a = 0 CTAs of the same cosmologies - so that every SM holds both kinds at once: the a = 0 half is
// latency bound (barriers, short DMMA chains) and leaves FP64-pipe time that the Horner sweeps of the a = 1 half can use.
__device__ __forceinline__ void rs_block(const int B, const int mix, int& b, int& half) {
  const int i = blockIdx.x;
  if (mix <= 0) { half = i >= B; b = half ? i - B : i; return; }
  const int per = 2 * mix, g = i / per, r = i - g * per;
  const int base = g * mix, n = min(mix, B - base);  // the last group may be short
  half = r >= n;
  b = base + (half ? r - n : r);
  if (r >= 2 * n) { b = -1; }  // padding blocks of a short last group
}

template <int NL, bool NNLO, int IA>
struct Accum {
  double lin[NL], loop[IA ? 12 : 1], nnlo[(IA && NNLO) ? NL : 1];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NL; ++i) lin[i] = 0.0;
#pragma unroll
    for (int i = 0; i < (IA ? 12 : 1); ++i) loop[i] = 0.0;
#pragma unroll
    for (int i = 0; i < ((IA && NNLO) ? NL : 1); ++i) nnlo[i] = 0.0;
  }
};

// acc = acc * z + q as a volatile instruction: volatile asm statements keep their program order, which pins the
// breadth-first schedule of the Horner sweep (NS * RS_C independent chains between two steps of one chain)
__device__ __forceinline__ void horner_step(double& acc, double z, double q) {
  asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc) : "d"(z), "d"(q));
}

// Horner sweep of the NS polynomial slots of one (a, l, l') over the RS_C points of a chunk, then the weighted slot
// sum  T[c] = R[l'][c] z[c] P[0][c] + Y k^2 [c] sum_v R[v][c] P[1+v][c]
template <int NIR, int NS, int NL, int RS_C>
__device__ __forceinline__ void horner(const double* __restrict__ q, const double (&z)[RS_C], const double (&yk)[RS_C],
                                       const double (&Rv)[NL][RS_C], int lp, double (&T)[RS_C]) {
  double P[NS][RS_C];
  {  // the two leading coefficients in one step: P = q[NIR-1] z + q[NIR-2] (no multiply-add onto a zero accumulator)
    const double2 t01 = *reinterpret_cast<const double2*>(q + (NIR - 1) * RS_SLOTS);
    const double2 t23 = *reinterpret_cast<const double2*>(q + (NIR - 1) * RS_SLOTS + 2);
    const double2 a01 = *reinterpret_cast<const double2*>(q + (NIR - 2) * RS_SLOTS);
    const double2 a23 = *reinterpret_cast<const double2*>(q + (NIR - 2) * RS_SLOTS + 2);
    const double qt[4] = {t01.x, t01.y, t23.x, t23.y}, qa[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int c = 0; c < RS_C; ++c) P[s][c] = fma(qt[s], z[c], qa[s]);
  }
#pragma unroll
  for (int p = NIR - 3; p >= 0; --p) {
    const double2 a01 = *reinterpret_cast<const double2*>(q + p * RS_SLOTS);
    const double2 a23 = *reinterpret_cast<const double2*>(q + p * RS_SLOTS + 2);
    const double qa[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int c = 0; c < RS_C; ++c) horner_step(P[s][c], z[c], qa[s]);
  }
#pragma unroll
  for (int c = 0; c < RS_C; ++c) {
    double y = 0.0;
#pragma unroll
    for (int v = 0; v < NS - 1; ++v) y = fma(Rv[v][c], P[1 + v][c], y);
    T[c] = fma(Rv[lp][c] * z[c], P[0][c], yk[c] * y);
  }
}

__device__ __forceinline__ double dot4(const double (&T)[RS_C], const double* row, double acc) {
  const double2 c01 = *reinterpret_cast<const double2*>(row), c23 = *reinterpret_cast<const double2*>(row + 2);
  return fma(T[3], c23.y, fma(T[2], c23.x, fma(T[1], c01.y, fma(T[0], c01.x, acc))));
}

// Cs holds the rows this IA contracts with: IA = 0: [NL][1][NsP] (C11); IA = 1: [NL][ncr-1][NsP] (Cct, Cloopl x12[, CctNNLO])
template <int NL, int NIR, bool NNLO, int IA>
__device__ __forceinline__ void sweep_chunk(const ResumArgs& a, const double* Ql, const double* Xs, const double* Ys,
                                            const double* Cs, int cp, int ik, double k2, int s0, Accum<NL, NNLO, IA>& A) {
  // R[v,k,s] of this chunk, shared by every l' (Rt is [v][s][k]: lanes = k read contiguously).  Issued first and
  // consumed only after the first Horner sweep, which hides the L1/L2 latency.
  double Rv[NL][RS_C];
  const double* rbase = a.Rt + (size_t)s0 * a.Nkr + ik;
#pragma unroll
  for (int v = 0; v < NL; ++v)
#pragma unroll
    for (int c = 0; c < RS_C; ++c) Rv[v][c] = __ldg(rbase + ((size_t)v * a.NsP + c) * a.Nkr);
  double z[RS_C], yk[RS_C];
  {
    const double2 x01 = *reinterpret_cast<const double2*>(Xs + s0), x23 = *reinterpret_cast<const double2*>(Xs + s0 + 2);
    const double2 y01 = *reinterpret_cast<const double2*>(Ys + s0), y23 = *reinterpret_cast<const double2*>(Ys + s0 + 2);
    z[0] = k2 * x01.x; z[1] = k2 * x01.y; z[2] = k2 * x23.x; z[3] = k2 * x23.y;
    yk[0] = k2 * y01.x; yk[1] = k2 * y01.y; yk[2] = k2 * y23.x; yk[3] = k2 * y23.y;
  }
  constexpr int NROW = IA ? 0 : 1;  // (unused marker to keep the two layouts visible)
  (void)NROW;
  const int nrow = IA ? a.ncr - 1 : 1;
#pragma unroll
  for (int lp = 0; lp < NL; ++lp) {
    const double* q = Ql + (size_t)(lp * NIR) * RS_SLOTS;
    double T[RS_C];
    if (a.nslot[lp] > 3) horner<NIR, 4, NL>(q, z, yk, Rv, lp, T);
    else horner<NIR, 3, NL>(q, z, yk, Rv, lp, T);
    const double* crow = Cs + (size_t)lp * nrow * cp + s0;  // cp: row pitch of Cs
    A.lin[lp] = dot4(T, crow, A.lin[lp]);  // IA = 0: C11, IA = 1: Cct
    if (IA) {
#pragma unroll
      for (int i = 0; i < 12; ++i) A.loop[i] = dot4(T, crow + (size_t)(1 + i) * cp, A.loop[i]);
      if (NNLO) A.nnlo[lp] = dot4(T, crow + (size_t)13 * cp, A.nnlo[lp]);
    }
  }
}

template <int NL, bool NNLO, int IA>
__device__ __forceinline__ void write_out(const ResumArgs& a, int b, int l, int ik, const Accum<NL, NNLO, IA>& A) {
  double* out = a.T + ((size_t)(l * a.Nk + a.Nklow + ik) * a.nterm) * a.Bp + b;
  const size_t Bp = a.Bp;
  // fire-and-forget reductions (RED): no other thread touches these elements
  if (!IA) {
    for (int i = 0; i < 3; ++i) {
      double v = 0.0;
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) v = fma(a.l11[lp * 3 + i], A.lin[lp], v);
      atomicAdd(out + (size_t)i * Bp, v);  // pybird.py:1442, :1445
    }
  } else {
    for (int i = 0; i < 6; ++i) {
      double v = 0.0;
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) v = fma(a.lct[lp * 6 + i], A.lin[lp], v);
      atomicAdd(out + (size_t)(3 + i) * Bp, v);  // pybird.py:1443, :1446
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) atomicAdd(out + (size_t)(9 + i) * Bp, A.loop[i]);  // pybird.py:1444, :1462
    if (NNLO)
      for (int i = 0; i < 3; ++i) {
        double v = 0.0;
#pragma unroll
        for (int lp = 0; lp < NL; ++lp) v = fma(a.lctnnlo[lp * 3 + i], A.nnlo[lp], v);
        atomicAdd(out + (size_t)(24 + i) * Bp, v);  // pybird.py:1455-1458
      }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// one CTA = one cosmology and one a in {0, 1}; 3 CTAs per SM (a single warp cannot issue DFMAs at the full pipe
// rate, it takes 3-4 warps per scheduler)
template <int NL, int NIR, bool NNLO, int IA>
__device__ __forceinline__ void resum_body(const ResumArgs& a, const int b) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NQH = NL * NL * NIR * RS_SLOTS;  // this a's half of the expanded Q table
  constexpr int qls = NL * NIR * RS_SLOTS;  // Q table of one l
  double* Qs = sm;                       // [NL][NL][NIR][4]
  double* Xs = Qs + NQH;                 // [NsP]
  double* Ys = Xs + a.NsP;               // [NsP]
  double* Cs = Ys + a.NsP;               // [NL][nrow][NsP]
  const int tid = threadIdx.x;
  const size_t Bp = a.Bp;
  const int nrow = IA ? a.ncr - 1 : 1, row0 = IA ? 1 : 0;

  for (int s = tid; s < a.NsP; s += RS_THREADS) {
    const bool ok = s < a.Ns;
    Xs[s] = ok ? a.F[(size_t)(a.row_x + s) * Bp + b] : 0.0;
    Ys[s] = ok ? a.F[(size_t)(a.row_y + s) * Bp + b] : 0.0;
  }
  const double* crb = a.Cr + (size_t)b * NL * a.ncr * a.Ns;               // point-major [b][l][ncr][Ns]
  const double* qf = a.Qf + (size_t)b * (2 * NQH) + (size_t)IA * NQH;     // Q^{ll'}(f) of this cosmology (resum_q_kernel)
  // The rows this half contracts with are contiguous per multipole and, like the Q table, whole 16-byte units: TMA bulk
  // copies (one thread issues NL + 1 of them, one mbarrier wait) instead of ~30 dependent load/store pairs per thread.
  const bool bulk = a.NsP == a.Ns && ((size_t)nrow * a.Ns) % 2 == 0 && ((size_t)a.ncr * a.Ns) % 2 == 0 && (row0 * a.Ns) % 2 == 0 &&
                    ((reinterpret_cast<uintptr_t>(crb) | reinterpret_cast<uintptr_t>(qf)) & 15) == 0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Cs + (size_t)NL * nrow * a.NsP);
  if (bulk) {
    if (tid == 0) {
      const uint32_t cbytes = (uint32_t)((size_t)nrow * a.Ns * sizeof(double)), qbytes = (uint32_t)(NQH * sizeof(double));
      const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar32));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar32), "r"(NL * cbytes + qbytes) : "memory");
#pragma unroll
      for (int l = 0; l < NL; ++l)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(Cs + (size_t)l * nrow * a.NsP)),
                     "l"(crb + ((size_t)l * a.ncr + row0) * a.Ns), "r"(cbytes), "r"(bar32)
                     : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(Qs)),
                   "l"(qf), "r"(qbytes), "r"(bar32)
                   : "memory");
    }
  } else {
#pragma unroll 8
    for (int i = tid; i < NL * nrow * a.NsP; i += RS_THREADS) {
      const int s = i % a.NsP, r = (i / a.NsP) % nrow, l = i / (a.NsP * nrow);
      Cs[i] = s < a.Ns ? crb[((size_t)l * a.ncr + row0 + r) * a.Ns + s] : 0.0;
    }
#pragma unroll
    for (int i = tid; i < NQH; i += RS_THREADS) Qs[i] = qf[i];
  }
  __syncthreads();  // X, Y (and the fallback copies) are in place; the barrier initialisation is visible
  if (bulk) {
    uint32_t ok;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
    } while (!ok);
  }

  const int ntask = NL * a.Nkr, nchunk = a.NsP / RS_C;
  const int rem = ntask % RS_THREADS;
  const bool coop = rem > 0 && rem <= RS_THREADS / 32;
  const int nmain = coop ? ntask - rem : ntask;
  Accum<NL, NNLO, IA> A;
  for (int task = tid; task < nmain; task += RS_THREADS) {
    const int l = task / a.Nkr, ik = task - l * a.Nkr;
    const double k2 = a.kr2[ik];
    A.zero();
    for (int ch = 0; ch < nchunk; ++ch) sweep_chunk<NL, NIR, NNLO, IA>(a, Qs + (size_t)l * qls, Xs, Ys, Cs, a.NsP, ik, k2, ch * RS_C, A);
    write_out<NL, NNLO, IA>(a, b, l, ik, A);
  }
  const int warp = tid >> 5, lane = tid & 31;
  if (coop && warp < rem) {
    const int task = nmain + warp;
    const int l = task / a.Nkr, ik = task - l * a.Nkr;
    const double k2 = a.kr2[ik];
    A.zero();
    for (int ch = lane; ch < nchunk; ch += 32) sweep_chunk<NL, NIR, NNLO, IA>(a, Qs + (size_t)l * qls, Xs, Ys, Cs, a.NsP, ik, k2, ch * RS_C, A);
#pragma unroll
    for (int i = 0; i < NL; ++i) A.lin[i] = warp_sum(A.lin[i]);
    if (IA) {
#pragma unroll
      for (int i = 0; i < 12; ++i) A.loop[i] = warp_sum(A.loop[i]);
      if (NNLO) {
#pragma unroll
        for (int i = 0; i < NL; ++i) A.nnlo[i] = warp_sum(A.nnlo[i]);
      }
    }
    if (lane == 0) write_out<NL, NNLO, IA>(a, b, l, ik, A);
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


// a = 1 half with the row contraction on the tensor pipe.  A warp owns 8 output columns (l,k) at a time
// and sweeps s in passes of 16: lane = (column r = lane>>2, s offset c4 = lane&3) runs its Horner chains at the 4 points
// s0 + 4c + c4 (c = 0..3) - which is the A fragment (8 columns x 4 s) of mma.m8n8k4.f64 for "quad" c - so the contraction
// over s with the 13 (14) rows C[l',i,s] is, per l' and quad, one DMMA per n8 tile of rows; the B fragment is one 8-byte
// shared load (row pitch = 4 mod 16 doubles: conflict-free) instead of 26 broadcast 16-byte loads per thread and l'.
// n columns: 0..11 the loop rows, 12 + l' the Cct row of l' (zero for the other l', so the Cct sums stay separate per l'),
// 16 + l' CctNNLO.  The columns left over by the groups of 8 (129 = 16 x 8 + 1) take the scalar sweep, one warp each.
template <int NL, int NIR, bool NNLO>
__device__ __forceinline__ void resum_body_mma(const ResumArgs& a, const int b) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NQL = NL * NIR * RS_SLOTS, qls = NQL;  // Q table of one l
  constexpr int NT = NNLO ? 3 : 2;
  const int NsP = a.NsP, CP = NsP + 4, nrow = a.ncr - 1;
  double* Qs = sm;                 // [NL][NL][NIR][4]
  double* Xs = Qs + NL * NQL;      // [NsP]
  double* Ys = Xs + NsP;           // [NsP]
  double* Cs = Ys + NsP;           // [NL][nrow][CP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(Cs + (size_t)NL * nrow * CP);
  const int tid = threadIdx.x;
  const size_t Bp = a.Bp;
  for (int s = tid; s < NsP; s += RS_THREADS) {
    const bool ok = s < a.Ns;
    Xs[s] = ok ? a.F[(size_t)(a.row_x + s) * Bp + b] : 0.0;
    Ys[s] = ok ? a.F[(size_t)(a.row_y + s) * Bp + b] : 0.0;
  }
  const int npad = NsP - a.Ns;  // the bulk copies bring Ns doubles per row: zero the rest (disjoint from what they write)
  for (int i = tid; i < NL * nrow * npad; i += RS_THREADS) Cs[(size_t)(i / npad) * CP + a.Ns + i % npad] = 0.0;
  const double* crb = a.Cr + (size_t)b * NL * a.ncr * a.Ns;             // point-major [b][l][ncr][Ns]
  const double* qf = a.Qf + (size_t)b * (2 * NL * NQL) + (size_t)NL * NQL;  // a = 1 half of Q^{ll'}(f)
  if (tid == 0) {
    const uint32_t rbytes = (uint32_t)(a.Ns * sizeof(double)), qbytes = (uint32_t)(NL * NQL * sizeof(double));
    const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar32));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar32), "r"(NL * nrow * rbytes + qbytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(Qs)),
                 "l"(qf), "r"(qbytes), "r"(bar32)
                 : "memory");
    for (int l = 0; l < NL; ++l) {
      for (int r = 0; r < nrow; ++r)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(Cs + ((size_t)l * nrow + r) * CP)),
                     "l"(crb + ((size_t)l * a.ncr + 1 + r) * a.Ns), "r"(rbytes), "r"(bar32)
                     : "memory");
    }
  }
  __syncthreads();  // X, Y are in place; the barrier initialisation is visible
  {
    uint32_t ok;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
    } while (!ok);
  }

  const int warp = tid >> 5, lane = tid & 31, r = lane >> 2, c4 = lane & 3;
  const int ntask = NL * a.Nkr, nrg = ntask / 8, rem = ntask - 8 * nrg;
  // B-fragment rows of this lane (n = 8 t + r): tile 0 -> loop row r; tile 1 -> loop row 8 + r (r < 4) or Cct of l' = r - 4;
  // tile 2 -> CctNNLO of l' = r.  Offsets are relative to the block of l' and include the lane's s offset.
  const int off0 = (1 + r) * CP + c4, off1 = (r < 4 ? 9 + r : 0) * CP + c4, off2 = 13 * CP + c4;
  for (int g = warp; g < nrg; g += RS_THREADS / 32) {
    const int task = 8 * g + r;
    const int l = task / a.Nkr, ik = task - l * a.Nkr;
    const double k2 = a.kr2[ik];
    const double* Ql = Qs + (size_t)l * qls;
    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
    for (int s0 = 0; s0 < NsP; s0 += RS_PASS) {
      double Rv[NL][RS_MC], z[RS_MC], yk[RS_MC];
      const double* rbase = a.Rt + (size_t)(s0 + c4) * a.Nkr + ik;
#pragma unroll
      for (int v = 0; v < NL; ++v)
#pragma unroll
        for (int c = 0; c < RS_MC; ++c) Rv[v][c] = __ldg(rbase + ((size_t)v * NsP + 4 * c) * a.Nkr);
#pragma unroll
      for (int c = 0; c < RS_MC; ++c) {
        z[c] = k2 * Xs[s0 + 4 * c + c4];
        yk[c] = k2 * Ys[s0 + 4 * c + c4];
      }
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) {
        const double* q = Ql + (size_t)(lp * NIR) * RS_SLOTS;
        double T[RS_MC];
        if (a.nslot[lp] > 3) horner<NIR, 4, NL>(q, z, yk, Rv, lp, T);
        else horner<NIR, 3, NL>(q, z, yk, Rv, lp, T);
        const double* cb = Cs + (size_t)lp * nrow * CP + s0;
        const bool v1 = r < 4 || r - 4 == lp, v2 = r == lp;
#pragma unroll
        for (int c = 0; c < RS_MC; ++c) {
          const double b0 = cb[off0 + 4 * c];
          const double b1 = v1 ? cb[off1 + 4 * c] : 0.0;
          dmma884(acc[0][0], acc[0][1], T[c], b0);
          dmma884(acc[1][0], acc[1][1], T[c], b1);
          if (NNLO) {
            const double b2 = v2 ? cb[off2 + 4 * c] : 0.0;
            dmma884(acc[NT - 1][0], acc[NT - 1][1], T[c], b2);
          }
        }
      }
    }
    // C fragment: row r (this lane's column (l,k)), n = 8 t + 2 c4 + {0, 1}
    double* out = a.T + ((size_t)(l * a.Nk + a.Nklow + ik) * a.nterm) * Bp + b;
    atomicAdd(out + (size_t)(9 + 2 * c4) * Bp, acc[0][0]);      // pybird.py:1444, :1462
    atomicAdd(out + (size_t)(10 + 2 * c4) * Bp, acc[0][1]);
    const double lin2 = __shfl_sync(0xffffffffu, acc[1][0], (lane & ~3) | 3);  // n = 14: Cct sum of l' = 2
    if (c4 < 2) {
      atomicAdd(out + (size_t)(17 + 2 * c4) * Bp, acc[1][0]);
      atomicAdd(out + (size_t)(18 + 2 * c4) * Bp, acc[1][1]);
    } else if (c4 == 2) {
      const double lin[3] = {acc[1][0], acc[1][1], lin2};
      for (int i = 0; i < 6; ++i) {
        double v = 0.0;
#pragma unroll
        for (int lp = 0; lp < NL; ++lp) v = fma(a.lct[lp * 6 + i], lin[lp], v);
        atomicAdd(out + (size_t)(3 + i) * Bp, v);  // pybird.py:1443, :1446
      }
    }
    if (NNLO) {
      const double nn2 = __shfl_sync(0xffffffffu, acc[NT - 1][0], (lane & ~3) | 1);  // n = 18
      if (c4 == 0) {
        const double nn[3] = {acc[NT - 1][0], acc[NT - 1][1], nn2};
        for (int i = 0; i < 3; ++i) {
          double v = 0.0;
#pragma unroll
          for (int lp = 0; lp < NL; ++lp) v = fma(a.lctnnlo[lp * 3 + i], nn[lp], v);
          atomicAdd(out + (size_t)(24 + i) * Bp, v);  // pybird.py:1455-1458
        }
      }
    }
  }
  // left-over columns: one warp each, lane = one s point at a time (a light sweep: one Horner chain per slot, the 13 (14)
  // rows read at consecutive s - conflict-free), then a shuffle reduction
  for (int t = warp; t < rem; t += RS_THREADS / 32) {
    const int task = 8 * nrg + t;
    const int l = task / a.Nkr, ik = task - l * a.Nkr;
    const double k2 = a.kr2[ik];
    const double* Ql = Qs + (size_t)l * qls;
    Accum<NL, NNLO, 1> A;
    A.zero();
    for (int s = lane; s < NsP; s += 32) {
      double Rv[NL][1], z[1], yk[1];
#pragma unroll
      for (int v = 0; v < NL; ++v) Rv[v][0] = __ldg(a.Rt + ((size_t)v * NsP + s) * a.Nkr + ik);
      z[0] = k2 * Xs[s];
      yk[0] = k2 * Ys[s];
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) {
        double T[1];
        if (a.nslot[lp] > 3) horner<NIR, 4, NL>(Ql + (size_t)(lp * NIR) * RS_SLOTS, z, yk, Rv, lp, T);
        else horner<NIR, 3, NL>(Ql + (size_t)(lp * NIR) * RS_SLOTS, z, yk, Rv, lp, T);
        const double* crow = Cs + (size_t)lp * nrow * CP + s;
        A.lin[lp] = fma(T[0], crow[0], A.lin[lp]);
#pragma unroll
        for (int i = 0; i < 12; ++i) A.loop[i] = fma(T[0], crow[(size_t)(1 + i) * CP], A.loop[i]);
        if (NNLO) A.nnlo[lp] = fma(T[0], crow[(size_t)13 * CP], A.nnlo[lp]);
      }
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) A.lin[i] = warp_sum(A.lin[i]);
#pragma unroll
    for (int i = 0; i < 12; ++i) A.loop[i] = warp_sum(A.loop[i]);
    if (NNLO) {
#pragma unroll
      for (int i = 0; i < NL; ++i) A.nnlo[i] = warp_sum(A.nnlo[i]);
    }
    if (lane == 0) write_out<NL, NNLO, 1>(a, b, l, ik, A);
  }
}

// a = 0 half on the tensor pipe: one CTA = one cosmology; a warp task = (slot, chunk of RL_MCH m8-tiles of k) with all
// NIR/8 n8-tiles of p, K = s in steps of 4.  Fragment layout of mma.m8n8k4.f64: a = A[lane>>2][lane&3],
// b = B[lane&3][lane>>2], c = C[lane>>2][2(lane&3) + {0,1}].  The A fragments (the fixed operator, read through L1) are
// double-buffered in registers two K-steps ahead of the DMMAs; the tasks left over by whole rounds of the warps are
// split along K (partial sums in extra planes of `part`, added in a fixed order) so that every warp gets the same work.
__host__ __device__ inline int rl_kpitch(int KP) { return KP + 2; }  // = 2 mod 8: conflict-free k2p columns
__host__ __device__ inline int rl_planes(int nslots) { return nslots + RS_THREADS / 32 - 1; }

template <int NL, int NIR>
__device__ __forceinline__ void resum_linear_body(const ResumArgs& a, const int b) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NT = NIR / 8, NW = RS_THREADS / 32;
  const int NsP = a.NsP, pitch = rl_pitch(NsP), KP = a.KPAD, KPP = rl_kpitch(KP);
  double* Xpow = sm;                           // [NIR][pitch]   X(s)^p
  double* base = Xpow + NIR * pitch;           // [2][NL][NsP]   X C11_l' | Y C11_l'
  double* k2p = base + 2 * NL * NsP;           // [NIR][KPP]     k^{2(p+1)}
  double* part = k2p + NIR * KPP;              // [planes][NL][KP] partial sums per slot (+ K-split pieces), fixed summation order
  double* Qs = part + rl_planes(a.nslots) * NL * KP;  // [NL][NL][NIR][4]  Q_0(f)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t Bp = a.Bp;

  {
    const double* crb = a.Cr + (size_t)b * NL * a.ncr * a.Ns;  // point-major Cr[b][l][ncr][Ns], row 0 = C11
    for (int s = tid; s < NsP; s += RS_THREADS) {
      const bool ok = s < a.Ns;
      const double x = ok ? a.F[(size_t)(a.row_x + s) * Bp + b] : 0.0;
      const double y = ok ? a.F[(size_t)(a.row_y + s) * Bp + b] : 0.0;
      double xp = 1.0;
#pragma unroll
      for (int p = 0; p < NIR; ++p) { Xpow[p * pitch + s] = xp; xp *= x; }
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) {
        const double c = ok ? crb[(size_t)lp * a.ncr * a.Ns + s] : 0.0;
        base[lp * NsP + s] = x * c;
        base[(NL + lp) * NsP + s] = y * c;
      }
    }
    for (int k = tid; k < KP; k += RS_THREADS) {
      const double k2 = k < a.Nkr ? a.kr2[k] : 0.0;
      double kp = k2;
#pragma unroll
      for (int p = 0; p < NIR; ++p) { k2p[p * KPP + k] = kp; kp *= k2; }
    }
    constexpr int NQH = NL * NL * NIR * RS_SLOTS;
    const double* qf = a.Qf + (size_t)b * (2 * NQH);  // a = 0 half of the expanded table
    for (int i = tid; i < NQH; i += RS_THREADS) Qs[i] = qf[i];
  }
  __syncthreads();

  const int mchunks = KP / (8 * RL_MCH), ntask = a.nslots * mchunks, nkb = NsP / 8;  // K blocks of 2 steps (8 points of s)
  const int nfull = (ntask / NW) * NW, rem = ntask - nfull;
  const int nsplit = (rem > 0 && NW % rem == 0) ? NW / rem : 1;  // pieces of a left-over task
  const int r = lane >> 2, c4 = lane & 3;
  const int nunit = nsplit > 1 ? nfull + NW : ntask;
  for (int unit = warp; unit < nunit; unit += NW) {
    int task = unit, kb0 = 0, kb1 = nkb, plane = -1;
    if (unit >= nfull && nsplit > 1) {
      const int u = unit - nfull, h = u % nsplit;
      task = nfull + u / nsplit;
      kb0 = h * nkb / nsplit; kb1 = (h + 1) * nkb / nsplit;
      if (h > 0) plane = a.nslots + (task - nfull) * (nsplit - 1) + (h - 1);
    }
    const int slot = task / mchunks, mh = task - slot * mchunks;
    if (plane < 0) plane = slot;
    const int lp = a.slot_lp[slot], sc = a.slot_s[slot];
    const int v = sc == 0 ? lp : sc - 1;                       // Bessel order of this slot
    const double* bb = base + ((sc == 0 ? 0 : NL) + lp) * NsP + c4;
    const double* xp = Xpow + r * pitch + c4;
    const double* ap = a.Rk + ((size_t)v * KP + mh * (8 * RL_MCH) + r) * NsP + c4;
    double acc[RL_MCH][NT][2];
#pragma unroll
    for (int i = 0; i < RL_MCH; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    auto load_a = [&](double (&af)[2][RL_MCH], const int kb) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int i = 0; i < RL_MCH; ++i) af[u][i] = __ldg(ap + (size_t)i * 8 * NsP + 8 * kb + 4 * u);
    };
    auto block = [&](const double (&af)[2][RL_MCH], const int kb) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        double bf[NT];
        const double bs = bb[8 * kb + 4 * u];
#pragma unroll
        for (int j = 0; j < NT; ++j) bf[j] = bs * xp[j * 8 * pitch + 8 * kb + 4 * u];
#pragma unroll
        for (int i = 0; i < RL_MCH; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[u][i], bf[j]);
      }
    };
    double a0[2][RL_MCH], a1[2][RL_MCH];
    int kb = kb0;
    load_a(a0, kb);
    for (; kb + 1 < kb1; kb += 2) {
      load_a(a1, kb + 1);
      block(a0, kb);
      if (kb + 2 < kb1) load_a(a0, kb + 2);
      block(a1, kb + 1);
    }
    if (kb < kb1) block(a0, kb);  // odd number of blocks: a0 holds the last one
    // fold  sum_p Q_0[l,l',p,slot] k^{2(p+1)} G[k,p]  as a second DMMA product: the C fragment scaled by k^{2(p+1)} is the A
    // fragment of K-step (j, e) for the column order p = 8j + 2(lane&3) + e (a contraction does not care about the order),
    // B[p][n] = Q_0[l = n] for n < NL, zero beyond; C2[k][l] lands in lane (k, l/2)
    double qb[NT][2];
    {
      const int n = lane >> 2;
      const double* q = Qs + (size_t)(((n < NL ? n : 0) * NL + lp) * NIR) * RS_SLOTS + sc;
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) qb[j][e] = n < NL ? q[(j * 8 + 2 * c4 + e) * RS_SLOTS] : 0.0;
    }
    double acc2[RL_MCH][2];
#pragma unroll
    for (int i = 0; i < RL_MCH; ++i) acc2[i][0] = acc2[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int i = 0; i < RL_MCH; ++i) {
          const int k = (mh * RL_MCH + i) * 8 + r;
          const double w = k2p[(j * 8 + 2 * c4 + e) * KPP + k] * acc[i][j][e];
          dmma884(acc2[i][0], acc2[i][1], w, qb[j][e]);
        }
#pragma unroll
    for (int i = 0; i < RL_MCH; ++i) {
      const int k = (mh * RL_MCH + i) * 8 + r;
      if (2 * c4 < NL) part[(size_t)(plane * NL + 2 * c4) * KP + k] = acc2[i][0];
      if (2 * c4 + 1 < NL) part[(size_t)(plane * NL + 2 * c4 + 1) * KP + k] = acc2[i][1];
    }
  }
  __syncthreads();
  for (int task = tid; task < NL * a.Nkr; task += RS_THREADS) {
    const int l = task / a.Nkr, ik = task - l * a.Nkr;
    double lin[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) lin[i] = 0.0;
    for (int slot = 0; slot < a.nslots; ++slot) {
      const double val = part[(size_t)(slot * NL + l) * KP + ik];
#pragma unroll
      for (int i = 0; i < NL; ++i) lin[i] += a.slot_lp[slot] == i ? val : 0.0;
    }
    if (nsplit > 1) {  // the later K pieces of the split tasks: each covers one chunk of k of one slot
      for (int e = 0; e < rem * (nsplit - 1); ++e) {
        const int t = nfull + e / (nsplit - 1), slot = t / mchunks, mh = t - slot * mchunks;
        if (ik / (8 * RL_MCH) != mh) continue;
        const double val = part[(size_t)((a.nslots + e) * NL + l) * KP + ik];
#pragma unroll
        for (int i = 0; i < NL; ++i) lin[i] += a.slot_lp[slot] == i ? val : 0.0;
      }
    }
    double* out = a.T + ((size_t)(l * a.Nk + a.Nklow + ik) * a.nterm) * Bp + b;
    for (int i = 0; i < 3; ++i) {
      double val = 0.0;
#pragma unroll
      for (int lp = 0; lp < NL; ++lp) val = fma(a.l11[lp * 3 + i], lin[lp], val);
      atomicAdd(out + (size_t)i * Bp, val);  // pybird.py:1442, :1445
    }
  }
}

// one CTA = one cosmology and one half (a = 1: counterterm + loop rows, a = 0: linear terms); order of the CTAs: rs_block
template <int NL, int NIR, bool NNLO, int MINB, bool MMA>
__global__ void __launch_bounds__(RS_THREADS, MINB * 128 / RS_THREADS) resum_kernel(ResumArgs a) {
  int b, half;
  rs_block(a.B, a.mix, b, half);
  if (b < 0) return;
#ifdef EFTB_TIMING_ONLY_HALF  // timing builds only (-DEFTB_TIMING_ONLY_HALF=0|1, wrong results): one half alone, same launch
  if (half != EFTB_TIMING_ONLY_HALF) return;
#endif
  if (half == 0) {
    if (MMA) resum_body_mma<NL, NIR, NNLO>(a, b);
    else resum_body<NL, NIR, NNLO, 1>(a, b);
  } else {
    resum_linear_body<NL, NIR>(a, b);
  }
}

// Q^{ll'}_u(f) = sum_d q[u][d] f^d for every cosmology (pybird.py:1367-1380 evaluates the reference's lambdas);
// qpack is [qdeg][NQ] so that consecutive threads read consecutive entries; Qf is [B][NQ]
__global__ void __launch_bounds__(256) resum_q_kernel(const double* __restrict__ qpack, const double* __restrict__ f, int NQ,
                                                      int qdeg, int B, double* __restrict__ Qf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= NQ || b >= B) return;
  const double f1 = f[b];
  double c[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) c[d] = d < qdeg ? __ldg(qpack + (size_t)d * NQ + i) : 0.0;
  double v = 0.0;
#pragma unroll
  for (int d = 15; d >= 0; --d) v = fma(v, f1, c[d]);
  Qf[(size_t)b * NQ + i] = v;
}

template <int NL, int NIR, bool NNLO>
int run(ResumArgs a, cudaStream_t s, int phase) {
  if (phase & EFTB_PHASE_FIRST) {
    dim3 qgrid((a.NQ + 255) / 256, a.B);
    resum_q_kernel<<<qgrid, 256, 0, s>>>(a.qpack, a.f, a.NQ, a.qdeg, a.B, a.Qf);
    EFTB_LAUNCH_CHECK();
  }
  if (!(phase & EFTB_PHASE_SECOND)) return EFTB_OK;
  // tuning knob (A/B runs): EFTB_RESUM_DOTS=scalar keeps the row contraction of the a = 1 half on the DFMA pipe
  static const bool scalar_dots = getenv("EFTB_RESUM_DOTS") && !strcmp(getenv("EFTB_RESUM_DOTS"), "scalar");
  // the DMMA form sweeps s in passes of RS_PASS and stages its rows by TMA bulk copies (16-byte units)
  a.mma = !scalar_dots && a.NsP % RS_PASS == 0 && a.Ns % 2 == 0 &&
          ((reinterpret_cast<uintptr_t>(a.Cr) | reinterpret_cast<uintptr_t>(a.Qf)) & 15) == 0;
  size_t smem = sizeof(double) * ((size_t)NL * NL * NIR * RS_SLOTS + 2 * a.NsP + (size_t)NL * (a.ncr - 1) * (a.NsP + 4) + 2);  // + mbarrier
  const size_t smem_lin = sizeof(double) * ((size_t)NIR * rl_pitch(a.NsP) + 2 * NL * a.NsP + (size_t)rl_kpitch(a.KPAD) * NIR +
                                            (size_t)rl_planes(a.nslots) * NL * a.KPAD + (size_t)NL * NL * NIR * RS_SLOTS);
  if (smem_lin > smem) smem = smem_lin;
  if (smem > 200 * 1024) { eftb_set_error("resum: %zu bytes of shared memory needed", smem); return EFTB_ERR_ARG; }
  static DeviceSmem conf_mma, conf_scalar;
  EFTB_SET_SMEM(conf_mma, (resum_kernel<NL, NIR, NNLO, RS_MINB, true>), smem);
  EFTB_SET_SMEM(conf_scalar, (resum_kernel<NL, NIR, NNLO, RS_MINB, false>), smem);
  // CTA order (rs_block); measured at 8192 points x 3 tracers: 0 -> 6.41 ms, 1 -> 6.29, 2 -> 6.28, 4 -> 6.27, 16 -> 6.27; at
  // 1024 / 2048 points the plain order is better (0.281 / 0.556 ms against 0.315 / 0.567: the a = 0 halves fill the tail)
  static const int mix_env = getenv("EFTB_RESUM_MIX") ? atoi(getenv("EFTB_RESUM_MIX")) : -1;
  const int mix = mix_env >= 0 ? mix_env : (a.B >= 4096 ? 4 : 0);
  a.mix = mix;
  const int nblk = mix <= 0 ? 2 * a.B : ((a.B + mix - 1) / mix) * 2 * mix;
  if (a.mma) resum_kernel<NL, NIR, NNLO, RS_MINB, true><<<nblk, RS_THREADS, smem, s>>>(a);
  else resum_kernel<NL, NIR, NNLO, RS_MINB, false><<<nblk, RS_THREADS, smem, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

}  // namespace

// Host-side packing at plan creation: transposed, s-padded operator Rt[v][NsP][Nkr]; the canonical slot
// structure of Q^{ll'} (slot 0 = (X, v = l'), slot 1 + v = (Y, v)) is verified against the table; coefficient
// table qpack[d][a][l][l'][p][slot].
int resum_pack(eftb_plan* p, const double* R, const double* q) {
  const eftb_config& c = p->cfg;
  const int Nl = c.Nl, NIR = c.NIR, Na = c.Na, Nn = 2 * NIR * Na, NsP = eftb_round_up(c.Ns, 16);
  if (Nl > 3 || Na != Nl || Na + 1 > RS_SLOTS || c.qdeg > 16) {
    eftb_set_error("resum: unsupported sizes Nl=%d Na=%d qdeg=%d", Nl, Na, c.qdeg);
    return EFTB_ERR_ARG;
  }
  auto qat = [&](int a, int l, int lp, int u, int d) { return q[((((size_t)a * Nl + l) * Nl + lp) * Nn + u) * c.qdeg + d]; };
  ResumPack& P = p->rs;
  for (int lp = 0; lp < Nl; ++lp) {
    int ymax = -1;
    for (int kind = 0; kind < 2; ++kind)
      for (int v = 0; v < Na; ++v) {
        bool any = false;
        for (int a = 0; a < 2 && !any; ++a)
          for (int l = 0; l < Nl && !any; ++l)
            for (int pp = 0; pp < NIR && !any; ++pp)
              for (int d = 0; d < c.qdeg && !any; ++d) any = qat(a, l, lp, (kind * NIR + pp) * Na + v, d) != 0.0;
        if (!any) continue;
        if (kind == 0 && v != lp) {
          eftb_set_error("resum: Q table couples X^p of l'=%d to Bessel order %d (expected %d only)", lp, v, lp);
          return EFTB_ERR_ARG;
        }
        if (kind == 1) ymax = v;
      }
    P.nslot[lp] = 2 + ymax < 3 ? 3 : 2 + ymax;  // kernels are instantiated for 3 and 4 slots
    P.slot_lp[P.nslots] = lp; P.slot_s[P.nslots++] = 0;        // (X, v = l')
    for (int v = 0; v <= ymax; ++v) { P.slot_lp[P.nslots] = lp; P.slot_s[P.nslots++] = 1 + v; }  // (Y, v)
  }
  const size_t NQ = (size_t)2 * Nl * Nl * NIR * RS_SLOTS;
  std::vector<double> qp(NQ * c.qdeg, 0.0), rt((size_t)Na * NsP * c.Nkr, 0.0);
  for (int a = 0; a < 2; ++a)
    for (int l = 0; l < Nl; ++l)
      for (int lp = 0; lp < Nl; ++lp)
        for (int pp = 0; pp < NIR; ++pp)
          for (int s = 0; s < 1 + Na; ++s) {
            const size_t e = ((((size_t)a * Nl + l) * Nl + lp) * NIR + pp) * RS_SLOTS + s;
            const int u = s == 0 ? pp * Na + lp : (NIR + pp) * Na + (s - 1);
            for (int d = 0; d < c.qdeg; ++d) qp[(size_t)d * NQ + e] = qat(a, l, lp, u, d);
          }
  for (int v = 0; v < Na; ++v)
    for (int k = 0; k < c.Nkr; ++k)
      for (int s = 0; s < c.Ns; ++s) rt[((size_t)v * NsP + s) * c.Nkr + k] = R[((size_t)v * c.Nkr + k) * c.Ns + s];
  // k-major copy for the linear-term GEMM, rows padded with zeros to whole warp tasks
  const int KPAD = eftb_round_up(c.Nkr, 8 * RL_MCH);
  std::vector<double> rk((size_t)Na * KPAD * NsP, 0.0);
  for (int v = 0; v < Na; ++v)
    for (int k = 0; k < c.Nkr; ++k)
      for (int s = 0; s < c.Ns; ++s) rk[((size_t)v * KPAD + k) * NsP + s] = R[((size_t)v * c.Nkr + k) * c.Ns + s];
  P.KPAD = KPAD;
  EFTB_CUDA_CHECK(cudaMalloc((void**)&P.Rk, rk.size() * sizeof(double)));
  EFTB_CUDA_CHECK(cudaMemcpy(P.Rk, rk.data(), rk.size() * sizeof(double), cudaMemcpyHostToDevice));
  P.NsP = NsP;
  P.NQ = (int)NQ;
  EFTB_CUDA_CHECK(cudaMalloc((void**)&P.qpack, qp.size() * sizeof(double)));
  EFTB_CUDA_CHECK(cudaMemcpy(P.qpack, qp.data(), qp.size() * sizeof(double), cudaMemcpyHostToDevice));
  EFTB_CUDA_CHECK(cudaMalloc((void**)&P.Rt, rt.size() * sizeof(double)));
  EFTB_CUDA_CHECK(cudaMemcpy(P.Rt, rt.data(), rt.size() * sizeof(double), cudaMemcpyHostToDevice));
  return EFTB_OK;
}

size_t resum_scratch_doubles(const eftb_plan* p, int B) { return (size_t)p->rs.NQ * B; }

int launch_resum(const eftb_plan* p, int B, int Bp, const double* F, const double* Cr, const double* f, double* T,
                 double* scratch, cudaStream_t s, int phase) {
  const eftb_config& c = p->cfg;
  ResumArgs a;
  a.F = F; a.Cr = Cr; a.f = f; a.Rt = p->rs.Rt; a.qpack = p->rs.qpack; a.kr2 = p->kr2; a.l11 = p->l11; a.lct = p->lct;
  a.lctnnlo = p->lctnnlo; a.T = T; a.Qf = scratch; a.B = B; a.Bp = Bp; a.Nk = c.Nk; a.Ns = c.Ns; a.NsP = p->rs.NsP;
  a.nterm = c.nterm; a.ncr = 14 + (c.with_nnlo ? 1 : 0); a.Nkr = c.Nkr; a.Nklow = c.Nklow; a.qdeg = c.qdeg;
  a.row_x = c.row_x; a.row_y = c.row_y; a.NQ = p->rs.NQ; a.Rk = p->rs.Rk; a.KPAD = p->rs.KPAD; a.nslots = p->rs.nslots;
  for (int i = 0; i < 12; ++i) { a.slot_lp[i] = (signed char)p->rs.slot_lp[i]; a.slot_s[i] = (signed char)p->rs.slot_s[i]; }
  for (int lp = 0; lp < 3; ++lp) a.nslot[lp] = lp < c.Nl ? p->rs.nslot[lp] : 0;
  const bool nnlo = c.with_nnlo != 0;
  if (c.Nl == 3 && c.NIR == 16 && c.Na == 3) return nnlo ? run<3, 16, true>(a, s, phase) : run<3, 16, false>(a, s, phase);
  if (c.Nl == 2 && c.NIR == 8 && c.Na == 2) return nnlo ? run<2, 8, true>(a, s, phase) : run<2, 8, false>(a, s, phase);
  eftb_set_error("resum: unsupported (Nl, NIR, Na) = (%d, %d, %d)", c.Nl, c.NIR, c.Na);
  return EFTB_ERR_ARG;
}
