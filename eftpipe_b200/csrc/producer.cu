// Input side on the device: a batched linear-power producer standing where CLASS / CAMB stand in the
// reference (boltzmann.py:22-101 `BoltzmannExtractor`: Pkh, f, DA, H per evaluation).  No Boltzmann code is available here,
// so the producer is the Eisenstein & Hu (1998) with-wiggles fitting formula in flat LCDM - the same model, formula for
// formula, as eftpipe_b200/synthetic.py (the generator of every synthetic input of the tests and the bench).  What it
// buys: the per-point input of the pipeline shrinks from 1.65 kB (P_lin on 200 nodes + scalars) to the three sampled
// cosmological parameters, and P_lin never crosses PCIe.
//
// One CTA per cosmology: the k nodes and the sigma8 quadrature nodes are spread over the threads, the growth and distance
// integrals are Gauss-Legendre sums over the first warps, block reductions in shared memory.
#include <math.h>
#include "common.cuh"

namespace {

struct EhConst {  // per cosmology, everything of eh98_transfer that does not depend on k
  double h, fb, fc, keq, s, ksilk, alpha_c, beta_c, alpha_b, beta_b, beta_node;
};

__device__ EhConst eh_setup(double Om, double Ob, double h, double Tcmb) {
  EhConst c;
  const double om = Om * h * h, ob = Ob * h * h;
  c.h = h;
  c.fb = Ob / Om;
  c.fc = 1.0 - c.fb;
  const double th = Tcmb / 2.7, th2 = th * th, th4 = th2 * th2;
  const double zeq = 2.50e4 * om / th4;
  c.keq = 7.46e-2 * om / th2;
  const double b1 = 0.313 * pow(om, -0.419) * (1.0 + 0.607 * pow(om, 0.674));
  const double b2 = 0.238 * pow(om, 0.223);
  const double zd = 1291.0 * pow(om, 0.251) / (1.0 + 0.659 * pow(om, 0.828)) * (1.0 + b1 * pow(ob, b2));
  const double Req = 31.5 * ob / th4 * (1e3 / zeq), Rd = 31.5 * ob / th4 * (1e3 / zd);
  c.s = 2.0 / (3.0 * c.keq) * sqrt(6.0 / Req) * log((sqrt(1.0 + Rd) + sqrt(Rd + Req)) / (1.0 + sqrt(Req)));
  c.ksilk = 1.6 * pow(ob, 0.52) * pow(om, 0.73) * (1.0 + pow(10.4 * om, -0.95));
  const double a1 = pow(46.9 * om, 0.670) * (1.0 + pow(32.1 * om, -0.532));
  const double a2 = pow(12.0 * om, 0.424) * (1.0 + pow(45.0 * om, -0.582));
  c.alpha_c = pow(a1, -c.fb) * pow(a2, -(c.fb * c.fb * c.fb));
  const double bb1 = 0.944 / (1.0 + pow(458.0 * om, -0.708));
  const double bb2 = pow(0.395 * om, -0.0266);
  c.beta_c = 1.0 / (1.0 + bb1 * (pow(c.fc, bb2) - 1.0));
  const double y = (1.0 + zeq) / (1.0 + zd), sy = sqrt(1.0 + y);
  const double G = y * (-6.0 * sy + (2.0 + 3.0 * y) * log((sy + 1.0) / (sy - 1.0)));
  c.alpha_b = 2.07 * c.keq * c.s * pow(1.0 + Rd, -0.75) * G;
  c.beta_node = 8.41 * pow(om, 0.435);
  c.beta_b = 0.5 + c.fb + (3.0 - 2.0 * c.fb) * sqrt((17.2 * om) * (17.2 * om) + 1.0);
  return c;
}

__device__ __forceinline__ double eh_T0(double q, double ac, double bc) {
  const double C = 14.2 / ac + 386.0 / (1.0 + 69.9 * pow(q, 1.08));
  const double L = log(M_E + 1.8 * bc * q);
  return L / (L + C * q * q);
}

__device__ double eh_transfer(const EhConst& c, double k_hmpc) {
  const double k = k_hmpc * c.h;  // 1 / Mpc
  const double q = k / (13.41 * c.keq), ks = k * c.s;
  const double x4 = (ks / 5.4) * (ks / 5.4), fint = 1.0 / (1.0 + x4 * x4);
  const double Tc = fint * eh_T0(q, 1.0, c.beta_c) + (1.0 - fint) * eh_T0(q, c.alpha_c, c.beta_c);
  const double bn = c.beta_node / ks;
  const double st = c.s / cbrt(1.0 + bn * bn * bn);
  const double bb = c.beta_b / ks, arg = k * st;
  const double Tb = (eh_T0(q, 1.0, 1.0) / (1.0 + (ks / 5.2) * (ks / 5.2)) + c.alpha_b / (1.0 + bb * bb * bb) * exp(-pow(k / c.ksilk, 1.4))) *
                    (arg == 0.0 ? 1.0 : sin(arg) / arg);
  return c.fb * Tb + c.fc * Tc;
}

__device__ __forceinline__ double Efun(double Om, double a) { return sqrt(Om / a + a * a * (1.0 - Om)); }

struct EhArgs {
  const double *theta, *kh, *gl_u, *gl_w;  // theta [B][3]: Om, h, sigma8; GL nodes / weights on [0, 1]
  double *pkh, *f, *DA, *H;
  double* sig2;   // [B] variance of the un-normalised spectrum in 8 Mpc/h spheres (redshift independent)
  int have_sig2;  // 1: read sig2 (another tracer of the same cosmologies computed it), 0: compute and store it
  int B, nk, ngl, nsig;
  double z, omega_b, ns, Tcmb;
};

constexpr int EH_THREADS = 128;

__device__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < EH_THREADS / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(EH_THREADS) eh_power_kernel(EhArgs a) {
  __shared__ EhConst cs;
  __shared__ double red[EH_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double Om = a.theta[3 * b], h = a.theta[3 * b + 1], s8 = a.theta[3 * b + 2];
  if (tid == 0) cs = eh_setup(Om, a.omega_b / (h * h), h, a.Tcmb);
  __syncthreads();
  const EhConst c = cs;
  // sigma8 of the un-normalised spectrum: trapezoid over ln k on logspace(-4, 2, nsig) (synthetic._sigma8_unnorm).  It does
  // not depend on the redshift: the tracers of one evaluation share it (ten times the work of the 200 output nodes)
  double sig2;
  if (a.have_sig2) {
    sig2 = a.sig2[b];
  } else {
    const double dl = 6.0 * M_LN10 / (a.nsig - 1);
    double part = 0.0;
    for (int i = tid; i < a.nsig; i += EH_THREADS) {
      const double k = exp(-4.0 * M_LN10 + dl * i), T = eh_transfer(c, k), x = 8.0 * k;
      const double W = 3.0 * (sin(x) - x * cos(x)) / (x * x * x);
      const double y = k * k * k * pow(k, a.ns) * T * T * W * W / (2.0 * M_PI * M_PI);
      part += (i == 0 || i == a.nsig - 1) ? 0.5 * y : y;
    }
    sig2 = block_sum(part, red) * dl;
    if (tid == 0 && a.sig2) a.sig2[b] = sig2;
  }
  // growth D(a) = 2.5 Om E(a) / a * int_0^a E(x)^-3 dx with x = a t^2 (synthetic.make_batch_fast), at a = 1 / (1 + z) and a = 1
  const double az = 1.0 / (1.0 + a.z);
  double g1 = 0.0, g0 = 0.0, da = 0.0;
  for (int i = tid; i < a.ngl; i += EH_THREADS) {
    const double u = a.gl_u[i], w = a.gl_w[i];
    const double e1 = Efun(Om, az * u * u), e0 = Efun(Om, u * u);
    g1 += w * 2.0 * az * u / (e1 * e1 * e1);
    g0 += w * 2.0 * u / (e0 * e0 * e0);
    const double zz = a.z * u;
    da += w * a.z / sqrt(Om * (1.0 + zz) * (1.0 + zz) * (1.0 + zz) + (1.0 - Om));
  }
  g1 = block_sum(g1, red);
  g0 = block_sum(g0, red);
  da = block_sum(da, red);
  const double D = 2.5 * Om * Efun(Om, az) / az * g1, D0 = 2.5 * Om * Efun(Om, 1.0) * g0;
  const double norm = s8 * s8 / sig2 * (D / D0) * (D / D0);
  for (int i = tid; i < a.nk; i += EH_THREADS) {
    const double k = a.kh[i], T = eh_transfer(c, k);
    a.pkh[(size_t)b * a.nk + i] = norm * pow(k, a.ns) * T * T;
  }
  if (tid == 0) {
    a.f[b] = (Om * (5.0 * az - 3.0 * D)) / (2.0 * (az * az * az * (1.0 - Om) + Om) * D);  // synthetic.growth_rate
    a.DA[b] = da / (1.0 + a.z);
    a.H[b] = sqrt(Om * (1.0 + a.z) * (1.0 + a.z) * (1.0 + a.z) + (1.0 - Om));
  }
}

}  // namespace

extern "C" int eftb_eh_power(int B, const double* theta, double z, double omega_b, double ns, double Tcmb, const double* kh, int nk,
                             const double* gl_u, const double* gl_w, int ngl, int nsig, double* pkh, double* f, double* DA, double* H,
                             double* sigma2, int have_sigma2, void* stream) {
  if (B < 1 || !theta || !kh || !gl_u || !gl_w || !pkh || !f || !DA || !H || nk < 1 || ngl < 1 || nsig < 2) {
    eftb_set_error("eftb_eh_power: NULL/invalid argument");
    return EFTB_ERR_ARG;
  }
  if (have_sigma2 && !sigma2) { eftb_set_error("eftb_eh_power: have_sigma2 without sigma2"); return EFTB_ERR_ARG; }
  EhArgs a{theta, kh, gl_u, gl_w, pkh, f, DA, H, sigma2, have_sigma2, B, nk, ngl, nsig, z, omega_b, ns, Tcmb};
  eh_power_kernel<<<B, EH_THREADS, 0, (cudaStream_t)stream>>>(a);
  EFTB_LAUNCH_CHECK();
  return EFTB_OK;
}
