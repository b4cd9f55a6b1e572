// Fixed-operator FP64 GEMM on the DMMA tensor-core path (mma.sync m8n8k4 f64; tcgen05 has no f64 kind).
//
//   C[M][N] = A[M][K] * X[K][N]      A: plan constant (row-major, zero-padded to [Mp][Kp] at upload)
//                                    X, C: batch-minor activations, N = (rows-per-point) * Bp, N % 32 == 0
//
// One CTA = 4 warps side by side along N (4 x 32 columns); every warp owns all MT m8-tiles of the CTA's
// M-slab, i.e. a (8*MT) x 32 accumulator block = MT*4 DMMA tiles, so per k4-step a warp issues MT + 4
// shared-memory fragment loads for 4*MT DMMAs.  A and X tiles are staged with 16-byte cp.async into
// double-buffered shared memory (row pitches 20 and 132 doubles = 4 mod 16 -> conflict-free fragment
// reads).  Used for: front operator, D -> P22/C22/C13 spectral transforms, B-spline collocation, the
// window/binning/chained projection and the inverse-covariance product.
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"

namespace {

constexpr int BK = 16;
constexpr int BN = 128;  // columns per CTA of the default 4-warp tile (the 2-warp variant covers 64)
constexpr int APITCH = BK + 4;

struct GemmArgs {
  const double* A;
  const double* X;
  double* C;
  int M, K, Kp, Mp, N;
  int a_batched, zdiv;
  size_t xs, xs2, cs, cs2;
  int pm_bp;         // > 0: point-major epilogue, column n = i * pm_bp + b -> C[b * pm_ld + i * pm_is + row]
  size_t pm_ld, pm_is;
  size_t ldx, ldc;   // row pitch of X and of the batch-minor C (>= N: a column window of a wider array)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MT, int NW>
__global__ void __launch_bounds__(32 * NW) gemm_f64_kernel(GemmArgs g) {
  constexpr int BM = 8 * MT, BN = 32 * NW, XPITCH = BN + 4, NTHR = 32 * NW;  // XPITCH = 4 mod 16 for NW = 2, 4
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                        // [2][BM][APITCH]
  double* Xs = smem + 2 * BM * APITCH;      // [2][BK][XPITCH]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, zb = blockIdx.z;
  const int zq = zb / g.zdiv, zr = zb % g.zdiv;
  const double* A = g.A + (g.a_batched ? (size_t)zq * g.Mp * g.Kp : 0) + (size_t)m0 * g.Kp;
  const double* X = g.X + (size_t)zr * g.xs + (size_t)zq * g.xs2;
  double* C = g.C + (size_t)zr * g.cs + (size_t)zq * g.cs2;

  double acc[MT][4][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nkt = g.Kp / BK;
  auto load_stage = [&](int kt, int buf) {
    // A tile: BM rows x 16 doubles = BM*8 chunks of 16 B
    double* as = As + buf * BM * APITCH;
    for (int c = tid; c < BM * 8; c += NTHR) {
      int r = c >> 3, q = c & 7;
      const bool ok = m0 + r < g.Mp;  // A is padded to a multiple of 8 rows only: the last slab may be partial
      cp_async16(as + r * APITCH + q * 2, ok ? A + (size_t)r * g.Kp + kt * BK + q * 2 : g.A, ok);
    }
    double* xs = Xs + buf * BK * XPITCH;
    for (int c = tid; c < BK * (BN / 2); c += NTHR) {
      int r = c / (BN / 2), q = c % (BN / 2);
      int kk = kt * BK + r, col = n0 + q * 2;
      bool ok = (kk < g.K) && (col < g.N);
      const double* src = ok ? X + (size_t)kk * g.ldx + col : X;
      cp_async16(xs + r * XPITCH + q * 2, src, ok);
    }
    cp_async_commit();
  };

  load_stage(0, 0);
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) {
      load_stage(kt + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* as = As + buf * BM * APITCH;
    const double* xs = Xs + buf * BK * XPITCH + warp * 32;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double a[MT], b[4];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = as[(i * 8 + (lane >> 2)) * APITCH + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = xs[(k4 * 4 + (lane & 3)) * XPITCH + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  if (g.pm_bp > 0) {
    // point-major epilogue: the 8 lanes that share a column pair hold 8 consecutive rows -> 64-byte runs per column
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + warp * 32 + j * 8 + (lane & 3) * 2;
      if (col >= g.N) continue;
      const int ii = col / g.pm_bp, b = col - ii * g.pm_bp;
      double* c0 = C + (size_t)b * g.pm_ld + (size_t)ii * g.pm_is;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int row = m0 + i * 8 + (lane >> 2);
        if (row < g.M) {
          c0[row] = acc[i][j][0];
          c0[g.pm_ld + row] = acc[i][j][1];
        }
      }
    }
    return;
  }
  // epilogue: each lane owns 2 adjacent columns of every tile row -> 16-byte stores
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    int row = m0 + i * 8 + (lane >> 2);
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + warp * 32 + j * 8 + (lane & 3) * 2;
      if (col < g.N) *reinterpret_cast<double2*>(C + (size_t)row * g.ldc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

template <int MT, int NW>
int launch(const GemmArgs& g, int nz, cudaStream_t s) {
  constexpr int BM = 8 * MT, BNW = 32 * NW;
  size_t smem = sizeof(double) * (2 * BM * APITCH + 2 * BK * (BNW + 4));
  static DeviceSmem configured;
  EFTB_SET_SMEM(configured, (gemm_f64_kernel<MT, NW>), smem);
  dim3 grid((g.N + BNW - 1) / BNW, (g.Mp + BM - 1) / BM, nz);
  gemm_f64_kernel<MT, NW><<<grid, 32 * NW, smem, s>>>(g);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int g_sms = 0;  // SMs of the device the current gemm_run call targets (refreshed per call from the per-device cache)

// slab height (in m8 tiles) for one launch: CTAs are dealt round-robin to the SMs, so the makespan is about
// ceil(nCTA / SMs) slabs of (8 MT + fixed per-slab overhead) rows; pick the MT that minimises it
int pick_mt(int M, int ncta_per_slab) {
  int best = 10;
  double best_cost = 1e30;
  for (int mt = 4; mt <= 11; ++mt) {
    const int bm = 8 * mt, slabs = (M + bm - 1) / bm;
    const long ncta = (long)slabs * ncta_per_slab;
    const double cost = (double)((ncta + g_sms - 1) / g_sms) * (bm + 16.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = mt; }
  }
  return best;
}

template <int NW>
int launch_mt_nw(int mt, const GemmArgs& g, int nz, cudaStream_t stream) {
  switch (mt) {
    case 11: return launch<11, NW>(g, nz, stream);
    case 10: return launch<10, NW>(g, nz, stream);
    case 9: return launch<9, NW>(g, nz, stream);
    case 8: return launch<8, NW>(g, nz, stream);
    case 7: return launch<7, NW>(g, nz, stream);
    case 6: return launch<6, NW>(g, nz, stream);
    case 5: return launch<5, NW>(g, nz, stream);
    default: return launch<4, NW>(g, nz, stream);
  }
}

// tile = (slab height in m8 tiles) * 16 + warps per CTA (4: 128 columns, 2: 64 columns)
int launch_mt(int tile, const GemmArgs& g, int nz, cudaStream_t stream) {
  return (tile & 15) == 2 ? launch_mt_nw<2>(tile >> 4, g, nz, stream) : launch_mt_nw<4>(tile >> 4, g, nz, stream);
}

// Slab height per GEMM shape: the analytic pick above ignores that several CTAs share an SM, so the first call of every
// shape (outside stream capture: the engine warms its pipelines up before capturing them) times all slab heights once
// and keeps the fastest.  The slab height only re-tiles the rows; the K summation order, hence every output bit, is the
// same for all of them.  EFTB_GEMM_AUTOTUNE=0 keeps the analytic pick.
int choose_mt(const GemmArgs& g, int nz, cudaStream_t stream) {
  static std::map<std::tuple<int, int, int, int, int, int>, int> tuned;
  static std::mutex mu;
  static const bool enabled = !(getenv("EFTB_GEMM_AUTOTUNE") && atoi(getenv("EFTB_GEMM_AUTOTUNE")) == 0);
  const int model = pick_mt(g.M, ((g.N + BN - 1) / BN) * nz) * 16 + 4;
  if (!enabled) return model;
  const auto key = std::make_tuple(g.Mp, g.Kp, g.N, nz, g.pm_bp > 0 ? 1 : 0, g.zdiv);
  std::lock_guard<std::mutex> lock(mu);
  auto it = tuned.find(key);
  if (it != tuned.end()) return it->second;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return model;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return model;
  int best = model;
  float best_ms = 1e30f;
  for (int nw = 4; nw >= 2; nw -= 2)
    for (int mt = 4; mt <= 11; ++mt) {
      const int tile = mt * 16 + nw;
      if (launch_mt(tile, g, nz, stream) != EFTB_OK) continue;  // warm-up (function attributes, caches)
      float ms = 1e30f;
      for (int round = 0; round < 2; ++round) {  // best of two rounds of 3 launches: one-off hiccups do not decide
        cudaEventRecord(e0, stream);
        for (int r = 0; r < 3; ++r) launch_mt(tile, g, nz, stream);
        cudaEventRecord(e1, stream);
        if (cudaEventSynchronize(e1) != cudaSuccess) { ms = -1.f; break; }
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        ms = t < ms ? t : ms;
      }
      if (ms < 0.f) { best = model; nw = 0; break; }
      if (ms < best_ms * 0.98f) { best_ms = ms; best = tile; }  // ties go to the variant tried first
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  tuned[key] = best;
  if (getenv("EFTB_GEMM_VERBOSE"))
    fprintf(stderr, "eftb gemm autotune: M=%d K=%d N=%d nz=%d -> MT=%d x %d warps (model MT=%d), %.1f us\n", g.M, g.K, g.N, nz, best >> 4,
            best & 15, model >> 4, best_ms / 3 * 1e3);
  return best;
}

}  // namespace

int gemm_upload(const double* host, int nbatch, int M, int K, GemmMatrix* out) {
  int Mp = eftb_round_up(M, 8), Kp = eftb_round_up(K, BK);
  size_t n = (size_t)nbatch * Mp * Kp;
  double* tmp = (double*)calloc(n, sizeof(double));
  if (!tmp) return EFTB_ERR_ARG;
  for (int b = 0; b < nbatch; ++b)
    for (int r = 0; r < M; ++r)
      for (int c = 0; c < K; ++c) tmp[((size_t)b * Mp + r) * Kp + c] = host[((size_t)b * M + r) * K + c];
  double* d = nullptr;
  cudaError_t e = cudaMalloc(&d, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(d, tmp, n * sizeof(double), cudaMemcpyHostToDevice);
  free(tmp);
  if (e != cudaSuccess) {
    eftb_set_error("gemm_upload: %s", cudaGetErrorString(e));
    return EFTB_ERR_CUDA;
  }
  out->d = d; out->M = M; out->K = K; out->Mp = Mp; out->Kp = Kp; out->MT = 0; out->nbatch = nbatch;
  return EFTB_OK;
}

void gemm_free(GemmMatrix* m) {
  if (m->d) cudaFree(m->d);
  m->d = nullptr;
}

int gemm_run(const GemmMatrix& A, const double* X, double* C, int N, int nz, int zdiv, size_t xs, size_t xs2,
             size_t cs, size_t cs2, cudaStream_t stream, const GemmPointMajor* pm, size_t ldx, size_t ldc) {
  if (!A.d || !X || !C || N % 2 || ldx % 2 || ldc % 2 || (pm && (pm->bp < 2 || pm->bp % 2))) { eftb_set_error("gemm_run: bad arguments"); return EFTB_ERR_ARG; }
  g_sms = eftb_sm_count();
  if (!g_sms) return EFTB_ERR_CUDA;
  GemmArgs g{A.d, X, C, A.M, A.K, A.Kp, A.Mp, N, A.nbatch > 1 ? 1 : 0, zdiv, xs, xs2, cs, cs2,
             pm ? pm->bp : 0, pm ? pm->ld : 0, pm ? pm->is : 0, ldx ? ldx : (size_t)N, ldc ? ldc : (size_t)N};
  return launch_mt(choose_mt(g, nz, stream), g, nz, stream);
}
