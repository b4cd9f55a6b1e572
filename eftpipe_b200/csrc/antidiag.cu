// Anti-diagonal reduction of the one-loop kernels on the FP64 tensor-core (DMMA) path.
//
// The reference contracts  sum_{n,m} c_n c_m M_b[n,m] x^{eta_n + eta_m}  separately for every k and s node
// (pybird.py:1074-1078, :1103-1125): 28 (+30) complex 257x257 quadratic forms per node.  Because
// eta_n + eta_m depends on n+m only, and so does the Bessel factor Ml[l,n,m] (pybird.py:1035-1038), all
// of them follow from
//        D_ch[t] = sum_{n+m=t} c_n c_m M_ch[n,m],      t = 0 .. 2 Nmax,
// (Hermitian half t <= Nmax, symmetric pairs n <= m folded into the table on the host).  For one
// anti-diagonal t this is a real GEMM
//        [Re D; Im D] (76 x B) = [[Mr, -Mi], [Mi, Mr]] (76 x 2 np(t)) * [Re(c_p c_{t-p}); Im(c_p c_{t-p})] (2 np(t) x B)
// whose right-hand operand (the coefficient products) is formed in registers directly in the DMMA B-fragment
// layout, so only the kernel table streams from memory.
//
// Work decomposition: one warp = 32 cosmologies (4 n-tiles) x all 38 complex channels (10 m-tiles of 8 rows =
// 4 channels x {re, im}) = 80 FP64 accumulators per lane; a k-step of `mma.sync.m8n8k4` consumes 2 pairs.
// The 4 warps of a CTA share the FFTLog coefficients of their 32 cosmologies in shared memory and each
// works through its own list of anti-diagonals (host-balanced by pair count, longest-first).  The table is
// stored in fragment order (1280 B per k-step, compact complex) and streamed through a per-warp ring of
// shared-memory stages with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), one stage (4 k-steps, 5 KB) ahead.
//
// Scheduling: a task = (group of 32 cosmologies, one of `nb` bins of 4 warp lists).  The grid is persistent - at most
// SMs x 2 CTAs, each taking a contiguous range of the group-major task list, so that it re-stages the coefficients only
// when its range crosses into the next group - and nb is chosen so that the tasks fill whole rounds of the resident CTAs:
// at a config-4 shard (256 groups) one CTA per group left 40 of the 148 SMs with a single CTA; 256 x 15 tasks over 296
// CTAs are 12 or 13 tasks each.
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace {

constexpr int AD_WARPS = 4;                  // warps per CTA (8 when the coefficient table leaves room for one CTA per SM only)
constexpr int AD_MT = 10;                    // m-tiles: 80 rows >= 2 * 38
constexpr int AD_NT = 4;                     // n-tiles per warp: 32 cosmologies
constexpr int AD_KSTEP_BYTES = AD_MT * 8 * 16;  // 10 m-tiles x (2 pairs x 4 channels) complex = 1280 B
constexpr int AD_S = 4;                      // k-steps per stage
constexpr int AD_NST = 2;                    // stages in the ring

struct AdArgs {
  const double* cre;      // F rows: Re c_n, n = 0..Nmax/2, batch-minor [.][Bp]
  const double* cim;
  const double2* tab;     // fragment-ordered table
  const int4* descs;      // stage descriptors {first k-step, count | last<<8, t, first pair}
  const int32_t* bin_off; // [nbins + 1]
  double* D;
  int Nmax, Bp;
  int nb, ntasks;         // bins of AD_WARPS warp lists per cosmology group; ngroups * nb
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double v, uint32_t mask) {  // mask = 0 or 0x80000000
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}

// One task: this warp's list `bin` of anti-diagonals for the 32 cosmologies staged in `cs`, written to D at column b0.
// Kept out of line on purpose: inlined into the persistent task loop the 80 accumulators plus the loop's own state hit the
// 255-register ceiling and ptxas serialises the ten A-fragment loads of a k-step on one register pair (measured: +25 %).
// iter0 = ring stages this warp has consumed so far (all tasks): ring slot and mbarrier parity follow from it.
__device__ __noinline__ uint32_t antidiag_task(const double* __restrict__ cs, unsigned char* ring, uint64_t* bars, const double2* tab,
                                               const int4* descs, int d0, int d1, double* D, int Nmax, int Bp, int b0, uint32_t iter0) {
  const int lane = threadIdx.x & 31;
  const int Nh = Nmax >> 1;
  auto issue = [&](int d) {  // lane 0 only
    const int4 ds = __ldg(descs + d);
    const int slot = (int)((iter0 + (uint32_t)(d - d0)) % AD_NST);
    const uint32_t bytes = (uint32_t)(ds.y & 0xff) * AD_KSTEP_BYTES;
    mbar_expect_tx(bars + slot, bytes);
    tma_bulk_load(ring + (size_t)slot * AD_S * AD_KSTEP_BYTES, reinterpret_cast<const unsigned char*>(tab) + (size_t)ds.x * AD_KSTEP_BYTES,
                  bytes, bars + slot);
  };
  if (lane == 0)
    for (int i = 0; i < AD_NST && d0 + i < d1; ++i) issue(d0 + i);

  // lane roles inside the fragments
  const int grp = lane >> 2, col = lane & 3;
  const int a_cc = grp >> 1, a_part = grp & 1;      // A row: channel-in-tile, {re, im}
  const int q = col >> 1, comp = col & 1;           // A col / B row: pair-in-step, {re, im} of the product
  const bool a_same = a_part == comp;
  const uint32_t a_neg = a_part ? 0u : 0x80000000u; // row re, col im -> -Mi ; row im, col re -> +Mi
  const int a_word = q * 4 + a_cc;                  // double2 index inside an m-tile block

  double acc[AD_MT][AD_NT][2];
#pragma unroll
  for (int i = 0; i < AD_MT; ++i)
#pragma unroll
    for (int j = 0; j < AD_NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int d = d0; d < d1; ++d) {
    const uint32_t it = iter0 + (uint32_t)(d - d0);
    const int slot = (int)(it % AD_NST);
    const int4 ds = __ldg(descs + d);
    const int cnt = ds.y & 0xff, last = ds.y >> 8, t = ds.z;
    mbar_wait(bars + slot, (it / AD_NST) & 1u);
    const double2* stage = reinterpret_cast<const double2*>(ring + (size_t)slot * AD_S * AD_KSTEP_BYTES);
    const int half = t >> 1;
    for (int ks = 0; ks < cnt; ++ks) {
      // B fragments: element (row = col, column = grp) of [Re(c_p c_m); Im(c_p c_m)] for pairs p0+0, p0+1
      const int p = min(ds.w + 2 * ks + q, half);   // a padded second pair multiplies zero table entries
      const int m = t - p;
      const bool folded = m > Nh;                    // c_m = conj(c_{Nmax-m})
      const int mi = folded ? Nmax - m : m;
      // value = xr * (comp ? yi : yr) + (comp ? xi : -xi) * (comp ? yr : yi),   yi carries the fold sign
      const double* xrow = cs + (size_t)(2 * p) * 32 + grp;
      const double* yrow = cs + (size_t)(2 * mi) * 32 + grp;
      const uint32_t s1 = (comp && folded) ? 0x80000000u : 0u;
      const uint32_t s2 = (comp ? 0u : 0x80000000u) ^ ((!comp && folded) ? 0x80000000u : 0u);
      double bf[AD_NT];
#pragma unroll
      for (int j = 0; j < AD_NT; ++j) {
        const double xr = xrow[j * 8], xi = xrow[32 + j * 8];
        const double y1 = yrow[(comp ? 32 : 0) + j * 8], y2 = yrow[(comp ? 0 : 32) + j * 8];
        bf[j] = fma(xr, flip_sign(y1, s1), flip_sign(xi, s2) * y2);
      }
      const double2* blk = stage + (size_t)ks * (AD_KSTEP_BYTES / 16) + a_word;
#pragma unroll
      for (int i = 0; i < AD_MT; ++i) {
        const double2 mv = blk[i * 8];
        const double af = a_same ? mv.x : flip_sign(mv.y, a_neg);
#pragma unroll
        for (int j = 0; j < AD_NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af, bf[j]);
      }
    }
    __syncwarp();
    if (lane == 0 && d + AD_NST < d1) issue(d + AD_NST);
    if (last) {
      // D[ch][t][part][b]: row R = 8 i + grp -> channel R >> 1, part R & 1; columns 8 j + 2 col + {0, 1}
      const size_t tstride = (size_t)2 * Bp, chstride = (size_t)(Nmax + 1) * tstride;
#pragma unroll
      for (int i = 0; i < AD_MT; ++i) {
        const int ch = 4 * i + a_cc;
        if (ch < EFTB_NCH) {
          double* out = D + (size_t)ch * chstride + (size_t)t * tstride + (size_t)a_part * Bp + b0 + 2 * col;
#pragma unroll
          for (int j = 0; j < AD_NT; ++j) *reinterpret_cast<double2*>(out + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
#pragma unroll
        for (int j = 0; j < AD_NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }
  }
  return iter0 + (uint32_t)(d1 - d0);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) antidiag_kernel(AdArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int Nh = a.Nmax >> 1;
  const int nrow = 2 * (Nh + 1);
  double* cs = reinterpret_cast<double*>(smraw);                                   // [Nh+1][2][32]
  unsigned char* ring0 = smraw + (size_t)nrow * 32 * sizeof(double);               // [warps][NST][S * 1280]
  uint64_t* bars0 = reinterpret_cast<uint64_t*>(ring0 + (size_t)WARPS * AD_NST * AD_S * AD_KSTEP_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* ring = ring0 + (size_t)warp * AD_NST * AD_S * AD_KSTEP_BYTES;
  uint64_t* bars = bars0 + warp * AD_NST;

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < AD_NST; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  uint32_t iter0 = 0;
  const int task0 = (int)((long long)blockIdx.x * a.ntasks / gridDim.x), task1 = (int)((long long)(blockIdx.x + 1) * a.ntasks / gridDim.x);
  int group = -1;
  for (int task = task0; task < task1; ++task) {
    const int b0 = (task / a.nb) * 32;
    if (task / a.nb != group) {
      group = task / a.nb;
      if (task != task0) __syncthreads();  // every warp is done with the previous group's coefficients
      // FFTLog coefficients of this group's 32 cosmologies: rows (n, re), (n, im)
      for (int i = tid; i < nrow * 32; i += WARPS * 32) {
        const int row = i >> 5, n = i & 31, idx = row >> 1;
        const double* src = (row & 1) ? a.cim : a.cre;
        cs[i] = src[(size_t)idx * a.Bp + b0 + n];
      }
      __syncthreads();
    }
    const int bin = (task % a.nb) * WARPS + warp;
    const int d0 = a.bin_off[bin], d1 = a.bin_off[bin + 1];
    if (d0 < d1) iter0 = antidiag_task(cs, ring, bars, a.tab, a.descs, d0, d1, a.D, a.Nmax, a.Bp, b0, iter0);
  }
}

}  // namespace

// schedule of anti-diagonals over `nbins` warps (longest-first greedy), as TMA stage descriptors
struct AdSchedule {
  int4* descs = nullptr;
  int32_t* bin_off = nullptr;
};

struct AntidiagPack {
  double2* tab = nullptr;
  std::vector<int> kstart;  // first k-step of anti-diagonal t, [Nmax + 2]
  std::map<int, AdSchedule> sched;
  std::mutex mu;
};

int antidiag_pack(eftb_plan* p, const double* pair_table, const int32_t* offsets) {
  const int Nmax = p->cfg.Nmax;
  AntidiagPack* P = new AntidiagPack();
  p->ad = P;
  P->kstart.assign(Nmax + 2, 0);
  for (int t = 0; t <= Nmax; ++t) P->kstart[t + 1] = P->kstart[t] + ((t / 2 + 1) + 1) / 2;
  const size_t nks = P->kstart[Nmax + 1];
  std::vector<double> tab(nks * (AD_KSTEP_BYTES / 8), 0.0);
  for (int t = 0; t <= Nmax; ++t) {
    const int np = t / 2 + 1;
    for (int pi = 0; pi < np; ++pi) {
      const size_t ks = P->kstart[t] + pi / 2;
      const int j = pi & 1;
      const double* src = pair_table + ((size_t)offsets[t] + pi) * EFTB_NCH * 2;
      for (int ch = 0; ch < EFTB_NCH; ++ch) {
        const size_t w = (ks * AD_MT + ch / 4) * 8 + j * 4 + (ch & 3);
        tab[2 * w] = src[2 * ch];
        tab[2 * w + 1] = src[2 * ch + 1];
      }
    }
  }
  EFTB_CUDA_CHECK(cudaMalloc((void**)&P->tab, tab.size() * sizeof(double)));
  EFTB_CUDA_CHECK(cudaMemcpy(P->tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  return EFTB_OK;
}

void antidiag_free(eftb_plan* p) {
  AntidiagPack* P = p->ad;
  if (!P) return;
  if (P->tab) cudaFree(P->tab);
  for (auto& kv : P->sched) {
    if (kv.second.descs) cudaFree(kv.second.descs);
    if (kv.second.bin_off) cudaFree(kv.second.bin_off);
  }
  delete P;
  p->ad = nullptr;
}

static int get_schedule(AntidiagPack* P, int Nmax, int nbins, AdSchedule* out) {
  std::lock_guard<std::mutex> lock(P->mu);
  auto it = P->sched.find(nbins);
  if (it != P->sched.end()) { *out = it->second; return EFTB_OK; }
  std::vector<int> order(Nmax + 1);
  for (int t = 0; t <= Nmax; ++t) order[t] = t;
  auto nk = [&](int t) { return P->kstart[t + 1] - P->kstart[t]; };
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return nk(x) > nk(y); });
  std::vector<std::vector<int>> bins(nbins);
  std::vector<long> load(nbins, 0);
  for (int t : order) {
    int best = 0;
    for (int b = 1; b < nbins; ++b) if (load[b] < load[best]) best = b;
    bins[best].push_back(t);
    load[best] += nk(t) + 2;  // + epilogue cost of one anti-diagonal, in k-step units
  }
  std::vector<int4> descs;
  std::vector<int32_t> off(nbins + 1, 0);
  for (int b = 0; b < nbins; ++b) {
    for (int t : bins[b]) {
      const int n = nk(t);
      for (int k0 = 0; k0 < n; k0 += AD_S) {
        const int cnt = std::min(AD_S, n - k0);
        const int last = k0 + cnt >= n;
        descs.push_back(make_int4(P->kstart[t] + k0, cnt | (last << 8), t, 2 * k0));
      }
    }
    off[b + 1] = (int32_t)descs.size();
  }
  AdSchedule s;
  EFTB_CUDA_CHECK(cudaMalloc((void**)&s.descs, descs.size() * sizeof(int4)));
  EFTB_CUDA_CHECK(cudaMemcpy(s.descs, descs.data(), descs.size() * sizeof(int4), cudaMemcpyHostToDevice));
  EFTB_CUDA_CHECK(cudaMalloc((void**)&s.bin_off, off.size() * sizeof(int32_t)));
  EFTB_CUDA_CHECK(cudaMemcpy(s.bin_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  P->sched[nbins] = s;
  *out = s;
  return EFTB_OK;
}

int launch_antidiag(const eftb_plan* p, int Bp, const double* F, double* D, cudaStream_t s, bool cf_set) {
  const eftb_config& c = p->cfg;
  AntidiagPack* P = p->ad;
  if (!P) { eftb_set_error("antidiag: plan has no pair table"); return EFTB_ERR_ARG; }
  const int ngroups = Bp / 32;
  const size_t cs_bytes = (size_t)2 * (c.Nmax / 2 + 1) * 32 * sizeof(double);
  auto smem_for = [&](int warps) { return cs_bytes + (size_t)warps * AD_NST * AD_S * AD_KSTEP_BYTES + warps * AD_NST * sizeof(uint64_t); };
  // two CTAs of 4 warps per SM when they fit (NFFT = 256); a longer coefficient table (NFFT = 512: 132 kB) leaves room for
  // one CTA only, which then runs 8 warps on the same table - the kernel needs two warps per scheduler to keep the DMMA
  // pipe fed (measured at NFFT = 512: 4 warps 16.1 ms, 0.58 of peak)
  static const int force_w = getenv("EFTB_AD_WARPS") ? atoi(getenv("EFTB_AD_WARPS")) : 0;  // tuning knob (A/B runs)
  int W = AD_WARPS;
  if (2 * (smem_for(4) + 1024) > 227 * 1024 && smem_for(8) + 1024 <= 227 * 1024) W = 8;
  if (force_w == 4 || (force_w == 8 && smem_for(8) + 1024 <= 227 * 1024)) W = force_w;
  const size_t smem = smem_for(W);
  if (smem > 227 * 1024) { eftb_set_error("antidiag: Nmax=%d needs %zu bytes of shared memory", c.Nmax, smem); return EFTB_ERR_ARG; }
  static DeviceSmem configured4, configured8;
  EFTB_SET_SMEM(configured4, antidiag_kernel<4>, smem_for(4));
  if (W == 8) EFTB_SET_SMEM(configured8, antidiag_kernel<8>, smem_for(8));
  const int sms = eftb_sm_count();
  if (!sms) return EFTB_ERR_CUDA;
  // resident CTAs: limited by shared memory (and to 2 x 4 warps per SM by registers)
  const int per_sm = W == 8 ? 1 : std::max(1, std::min(2, (int)((227 * 1024) / (smem + 1024))));
  const int slots = sms * per_sm;
  // bins per group: estimated makespan = rounds of the resident CTAs x time of one task, where a task lasts as long as its
  // longest warp list - the average list, but never less than the longest single anti-diagonal (lists are whole
  // anti-diagonals) - plus ~2 k-steps of task overhead; ties go to fewer, larger tasks.  One group (the
  // B = 1 latency case) spreads over up to 64 CTAs: 0.088 -> 0.047 ms.
  const int maxnb = std::min(64, (c.Nmax + 1 + W - 1) / W);
  const double total_steps = (double)P->kstart[c.Nmax + 1], longest = (double)(P->kstart[c.Nmax + 1] - P->kstart[c.Nmax]);
  int nb = 1;
  double best = 1e300;
  for (int cand = 1; cand <= maxnb; ++cand) {
    const long nt = (long)ngroups * cand;
    const long rounds = (nt + slots - 1) / slots;
    const double cost = (double)rounds * (std::max(total_steps / (cand * W), longest) + 2.0);
    if (cost < best * (1.0 - 1e-3)) { best = cost; nb = cand; }
  }
  static const int force_nb = getenv("EFTB_AD_NB") ? atoi(getenv("EFTB_AD_NB")) : 0;  // tuning knob (A/B runs)
  if (force_nb > 0) nb = std::min(force_nb, maxnb);
  const int ntasks = ngroups * nb;
  AdSchedule sc;
  int rc = get_schedule(P, c.Nmax, nb * W, &sc);
  if (rc) return rc;
  AdArgs a;
  const bool second = cf_set && c.row_cre_cf >= 0;  // pybird.py:1151-1160: coef_cf differs from coef_pk
  a.cre = F + (size_t)(second ? c.row_cre_cf : c.row_cre) * Bp;
  a.cim = F + (size_t)(second ? c.row_cim_cf : c.row_cim) * Bp;
  a.tab = P->tab; a.descs = sc.descs; a.bin_off = sc.bin_off; a.D = D; a.Nmax = c.Nmax; a.Bp = Bp;
  a.nb = nb; a.ntasks = ntasks;
  if (W == 8) antidiag_kernel<8><<<std::min(ntasks, slots), 8 * 32, smem, s>>>(a);
  else antidiag_kernel<4><<<std::min(ntasks, slots), 4 * 32, smem, s>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}
