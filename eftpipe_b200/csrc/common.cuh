// Shared declarations of libeftb200 (sm_100a).  Internal header - the public ABI is include/eftb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/eftb200.h"

#define EFTB_NCH 38
#define EFTB_N22 28
#define EFTB_N13 10

void eftb_set_error(const char* fmt, ...);

#define EFTB_CUDA_CHECK(call)                                                              \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      eftb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return EFTB_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

extern unsigned long long g_eftb_launches;  // kernels launched (or recorded into a capturing stream) by this library

#define EFTB_LAUNCH_CHECK()                                                                \
  do {                                                                                     \
    __atomic_fetch_add(&g_eftb_launches, 1ull, __ATOMIC_RELAXED);                          \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      eftb_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));    \
      return EFTB_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

static inline int eftb_round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- per-device launch configuration -------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to a device, not to the process: a process
// that drives several GPUs must configure every kernel on each of them.  One DeviceSmem per kernel (family) remembers
// the largest size configured per device; `needs(dev, bytes)` says whether the attribute has to be set (again) on that
// device.
#define EFTB_MAX_DEVICES 64
int eftb_current_device();  // cudaGetDevice, -1 on error (error text set)
int eftb_sm_count();        // SMs of the current device (cached per device), 0 on error
struct DeviceSmem {
  size_t configured[EFTB_MAX_DEVICES] = {};
  // true: `bytes` exceeds what this device was configured for - the caller sets the attribute and then calls done()
  bool needs(int dev, size_t bytes) const { return dev < 0 || dev >= EFTB_MAX_DEVICES || bytes > __atomic_load_n(&configured[dev], __ATOMIC_ACQUIRE); }
  void done(int dev, size_t bytes) { if (dev >= 0 && dev < EFTB_MAX_DEVICES) __atomic_store_n(&configured[dev], bytes, __ATOMIC_RELEASE); }
};
#define EFTB_SET_SMEM(state, kernel, bytes)                                                                         \
  do {                                                                                                              \
    const int _dev = eftb_current_device();                                                                         \
    if (_dev < 0) return EFTB_ERR_CUDA;                                                                             \
    if ((state).needs(_dev, (bytes))) {                                                                             \
      EFTB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      (state).done(_dev, (bytes));                                                                                  \
    }                                                                                                               \
  } while (0)

// ---- fixed-operator GEMM:  C[M][N] = A[M][K] X[K][N], A zero-padded to [Mp][Kp] on upload -------
struct GemmMatrix {       // device copy of a fixed operator, padded for the kernel's tiles
  double* d = nullptr;    // [nbatch][Mp][Kp]
  int M = 0, K = 0, Mp = 0, Kp = 0, MT = 0, nbatch = 0;
};
int gemm_upload(const double* host, int nbatch, int M, int K, GemmMatrix* out);
void gemm_free(GemmMatrix* m);
// z-batch: zb in [0, nz), zq = zb / zdiv, zr = zb % zdiv: A matrix index = zq (if A has >1 batch),
// X offset = zr*xs + zq*xs2, C offset = zr*cs + zq*cs2
// pm != NULL: point-major output.  Column n = i * bp + b (b = cosmology) goes to C[b * ld + i * is + row] (+ the
// z offsets zr*cs + zq*cs2): per-cosmology contiguous rows for the one-CTA-per-cosmology consumers.
struct GemmPointMajor { int bp; size_t ld, is; };
int gemm_run(const GemmMatrix& A, const double* X, double* C, int N, int nz, int zdiv, size_t xs, size_t xs2,
             size_t cs, size_t cs2, cudaStream_t stream, const GemmPointMajor* pm = nullptr, size_t ldx = 0,
             size_t ldc = 0);  // ldx / ldc: row pitch of X / C when the N columns are a window of a wider array (0 = N)

struct AntidiagPack;  // fragment-ordered pair table + warp schedules (antidiag.cu)

// IR-resummation constants in kernel layout (resum.cu resum_pack)
struct ResumPack {
  double *qpack = nullptr, *Rt = nullptr;  // [qdeg][2][Nl][Nl][NIR][4],  [Na][NsP][Nkr]
  double* Rk = nullptr;                    // [Na][KPAD][NsP]: k-major copy, rows zero-padded to KPAD (linear-term GEMM)
  int NsP = 0, NQ = 0, KPAD = 0;
  int nslot[3] = {};  // canonical slots per l': (X, v = l'), (Y, 0), (Y, 1)[, (Y, 2)]
  int nslots = 0;     // all non-zero (l', kind, v) polynomials, listed as (l', canonical slot) pairs
  int slot_lp[12] = {}, slot_s[12] = {};
};

// ---- plan ---------------------------------------------------------------------------------------
struct eftb_plan {
  eftb_config cfg;
  int K;  // nin + ntail + ntailx
  double *k = nullptr, *l11 = nullptr, *lct = nullptr, *lctnnlo = nullptr, *l22 = nullptr, *l13 = nullptr;
  double *lr = nullptr, *lrx = nullptr;
  GemmMatrix Wf, Ak, As, Cinv, project, project_st;  // project_st: operator of the stochastic rows when it differs
  AntidiagPack* ad = nullptr;
  double* kr2 = nullptr;
  ResumPack rs;
  double *knot_lo = nullptr, *basis = nullptr, *mu = nullptr, *wl = nullptr;
  int32_t* perm_out = nullptr;  // point-major export permutation
  int perm_rows = 0;
  // fused pipeline only: side stream + fork/join events (one eftb_eval_terms call at a time per plan)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_loops = nullptr, ev_join = nullptr;
};

// kernels' host launchers (defined in the respective .cu files)
int resum_pack(eftb_plan* p, const double* R, const double* q);
int antidiag_pack(eftb_plan* p, const double* pair_table, const int32_t* offsets);
void antidiag_free(eftb_plan* p);
int launch_front_prepare(const eftb_plan* p, int B, int Bp, const double* plin, double* u, cudaStream_t s);
int launch_antidiag(const eftb_plan* p, int Bp, const double* F, double* D, cudaStream_t s, bool cf_set = false);
int launch_group(const eftb_plan* p, int Bp, const double* F, const double* P22, const double* Cs,
                 const double* f, double* T, double* Cr, cudaStream_t s);  // Cs == NULL: Cloopl rows of Cr already filled
int launch_regroup(const eftb_plan* p, int Bp, const double* D, const double* f, double* Dg, cudaStream_t s);
// phases of the two-kernel stages, so that the fused pipeline can run the cosmology-only halves (the Q(f) expansion,
// the AP resampling geometry) on a side stream while the loop kernels run
enum { EFTB_PHASE_ALL = 3, EFTB_PHASE_FIRST = 1, EFTB_PHASE_SECOND = 2 };
int launch_resum(const eftb_plan* p, int B, int Bp, const double* F, const double* Cr, const double* f,
                 double* T, double* scratch, cudaStream_t s, int phase = EFTB_PHASE_ALL);  // FIRST: Q(f); SECOND: the sweep
size_t resum_scratch_doubles(const eftb_plan* p, int B);  // expanded Q(f): [B][2 Nl Nl NIR 4]
int launch_ap(const eftb_plan* p, int B, int Bp, const double* coef, const double* Tin, const double* DA,
              const double* H, double* scratch, double* Tout, cudaStream_t s, int phase = EFTB_PHASE_ALL);  // FIRST: geometry
int ap_chunk_count(const eftb_plan* p, int B);
size_t ap_coef_doubles(const eftb_plan* p, int Bp);  // size of the point-major B-spline coefficient buffer (even per-point stride)  // launches the AP stage splits the batch into (phases need 1)
size_t ap_scratch_doubles(const eftb_plan* p, int B);  // banded AP operator G + window metadata
int launch_to_batch_minor(const double* in, int B, int Bp, int R, double* out, cudaStream_t s);
int launch_scalars_to_batch_minor(const double* s0, const double* s1, const double* s2, int B, int Bp, double* out, cudaStream_t s);
int launch_to_point_major(const double* in, int B, int Bp, int R, const int32_t* perm, double* out,
                          cudaStream_t s);
