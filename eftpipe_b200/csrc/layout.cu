// Layout kernels: point-major <-> batch-minor transposes and the front-end input vector.
#include "common.cuh"

namespace {

// in[B][R] -> out[R][Bp]; lanes beyond B replicate point B-1 so that padded lanes stay finite
__global__ void to_batch_minor_kernel(const double* __restrict__ in, int B, int Bp, int R, double* __restrict__ out) {
  __shared__ double tile[32][33];
  int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int b = min(b0 + j, B - 1), r = r0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R) ? in[(size_t)b * R + r] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, b = b0 + threadIdx.x;
    if (r < R && b < Bp) out[(size_t)r * Bp + b] = tile[threadIdx.x][j];
  }
}

// in[rows][Bp] -> out[B][R] with out[b][r] = in[perm ? perm[r] : r][b]
__global__ void to_point_major_kernel(const double* __restrict__ in, int B, int Bp, int R,
                                      const int32_t* __restrict__ perm, double* __restrict__ out) {
  __shared__ double tile[32][33];
  int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, b = b0 + threadIdx.x;
    if (r < R && b < Bp) {
      int src = perm ? perm[r] : r;
      tile[j][threadIdx.x] = in[(size_t)src * Bp + b];
    }
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int b = b0 + j, r = r0 + threadIdx.x;
    if (b < B && r < R) out[(size_t)b * R + r] = tile[threadIdx.x][j];
  }
}

// power-law tails beyond the last input sample (fftlog.py:146-151) for the loop FFTLog and for the
// 32-point IR-filter FFTLog of P exp(-k^2/Lambda^2)/k^2 (pybird.py:1321-1325).  One thread per (tail node, point).
__global__ void front_tails_kernel(const double* __restrict__ plin, int B, int Bp, int nin, int ntail, int ntailx,
                                   const double* __restrict__ lr, const double* __restrict__ lrx, double inv_dlog,
                                   double wx_last, double wx_prev, double* __restrict__ u) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (b >= Bp) return;
  const int src = min(b, B - 1);
  double last = plin[(size_t)src * nin + nin - 1], prev = plin[(size_t)src * nin + nin - 2];
  double x;
  if (i < ntail) {
    x = lr[i];
  } else {
    last *= wx_last;
    prev *= wx_prev;
    x = lrx[i - ntail];
  }
  const double slope = (log(last) - log(prev)) * inv_dlog;
  u[(size_t)(nin + i) * Bp + b] = last * exp(slope * x);
}

// up to three per-point scalars (f, D_A, H) to their batch-minor rows in one launch: out[r][Bp], pad lanes replicate B-1
__global__ void scalars_to_batch_minor_kernel(const double* __restrict__ s0, const double* __restrict__ s1,
                                              const double* __restrict__ s2, int B, int Bp, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bp) return;
  const int src = min(b, B - 1);
  out[b] = s0[src];
  if (s1) out[(size_t)Bp + b] = s1[src];
  if (s2) out[2 * (size_t)Bp + b] = s2[src];
}

}  // namespace

int launch_scalars_to_batch_minor(const double* s0, const double* s1, const double* s2, int B, int Bp, double* out, cudaStream_t s) {
  scalars_to_batch_minor_kernel<<<(Bp + 255) / 256, 256, 0, s>>>(s0, s1, s2, B, Bp, out);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int launch_to_batch_minor(const double* in, int B, int Bp, int R, double* out, cudaStream_t s) {
  dim3 grid(Bp / 32, (R + 31) / 32), block(32, 8);
  to_batch_minor_kernel<<<grid, block, 0, s>>>(in, B, Bp, R, out);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int launch_to_point_major(const double* in, int B, int Bp, int R, const int32_t* perm, double* out, cudaStream_t s) {
  dim3 grid(Bp / 32, (R + 31) / 32), block(32, 8);
  to_point_major_kernel<<<grid, block, 0, s>>>(in, B, Bp, R, perm, out);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}

int launch_front_prepare(const eftb_plan* p, int B, int Bp, const double* plin, double* u, cudaStream_t s) {
  const eftb_config& c = p->cfg;
  int rc = launch_to_batch_minor(plin, B, Bp, c.nin, u, s);
  if (rc) return rc;
  front_tails_kernel<<<dim3((Bp + 127) / 128, c.ntail + c.ntailx), 128, 0, s>>>(plin, B, Bp, c.nin, c.ntail, c.ntailx, p->lr, p->lrx, c.inv_dlog,
                                                     c.wx_last, c.wx_prev, u);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}
