/*
 * eftb200.h - C ABI of libeftb200.so, the sm_100a implementation of eftpipe's PyBird hot path.
 *
 * The reference (zhaoruiyang98/eftpipe) is pure Python and has NO FFI boundary for this path;
 * its seams are Python protocols.  This header is therefore the boundary a
 * maintainer would bind with ctypes; every entry point names the reference
 * function(s) whose per-evaluation arithmetic it replaces (paths relative to eftpipe/).
 *
 * Conventions
 *  - plain C, no torch types; all data pointers are DEVICE pointers owned by the caller unless a
 *    parameter is documented as host memory; the library owns only the immutable plan constants.
 *  - every call enqueues work on `stream` (a cudaStream_t passed as void*) and does not synchronise.
 *    eftb_eval_terms additionally forks part of its work onto a plan-owned side stream and joins it back into
 *    `stream` with events before it returns (capturable into a CUDA graph): one eftb_eval_terms call at a time per
 *    plan; the stage entry points may run concurrently on one plan with distinct workspaces.
 *  - return value: 0 on success, negative eftb_status on error (never throws, never aborts);
 *    per-point numerical failures (non positive-definite F2) are reported in the `status` array.
 *  - "batch-minor" arrays: logical shape [rows][Bp] with the cosmology index fastest and
 *    Bp = eftb_padded_batch(B) (B rounded up to a multiple of 32; pad lanes replicate point B-1).
 *  - workspaces and scratch buffers must be 16-byte aligned (any cudaMalloc / torch allocation is): parts of them are
 *    sources of TMA bulk copies.
 *  - all arithmetic is IEEE binary64.
 */
#ifndef EFTB200_H
#define EFTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFTB_ABI_VERSION 5
#define EFTB_NPAR 19 /* nuisance columns per tracer read by the bias reduction (eftb_like_constants.par_index) */

typedef enum {
  EFTB_OK = 0,
  EFTB_ERR_ARG = -1,      /* NULL pointer / inconsistent sizes (reference: ValueError) */
  EFTB_ERR_CUDA = -2,     /* CUDA runtime error, see eftb_last_error() */
  EFTB_ERR_NOT_BUILT = -3,/* stage not present in this plan (e.g. AP requested, plan has none) */
  EFTB_ERR_WORKSPACE = -4 /* workspace too small */
} eftb_status;

typedef struct eftb_plan eftb_plan; /* one tracer pipeline (pybird Common+NonLinear+Resum+APeffect+Window+Binning) */
typedef struct eftb_like eftb_like; /* one likelihood (likelihood.py EFTLike + marginal.py) */

/* ---- plan description (host memory) ------------------------------------------------------------ */
typedef struct {
  /* grids: pybird.Common (pybird/pybird.py:486-582) */
  int32_t Nl, Nk, Ns, Nmax, nterm, with_nnlo;
  /* front operator layout (eftpipe_b200/plan.py front_operator; fftlog.py:84-166, pybird.py:1316-1353) */
  int32_t nin, ntail, ntailx, front_rows;
  int32_t row_cre, row_cim, row_p11, row_p13, row_c11, row_cct, row_cctnnlo, row_x, row_y;
  double inv_dlog, wx_last, wx_prev;
  /* anti-diagonal pair table */
  int32_t npair;
  /* IR resummation (pybird.py:1174-1464) */
  int32_t has_resum, NIR, Na, Nkr, Nklow, qdeg;
  /* Alcock-Paczynski (pybird.py:1467-1628) */
  int32_t has_ap, nmu, nint, ap_st;
  double da_fid, h_fid;
  /* window (+ICC) / binning / chained projection (window.py:371-415, icc.py:471-484,
     binning.py:131-162, chained.py:32-68) */
  int32_t has_project, nout, nl_out; /* nout = nl_out * nk_out */
  /* IRcutoff "loop" / "resum" (pybird.py:1151-1160): the configuration-space terms use a second set of FFTLog
     coefficients, emitted by the front operator at these rows; row_cre_cf < 0: one set serves both */
  int32_t row_cre_cf, row_cim_cf;
} eftb_config;

typedef struct {
  const double* k;          /* [Nk]                      co.k */
  const double* l11;        /* [Nl][3]   pybird.py:570   */
  const double* lct;        /* [Nl][6]   pybird.py:571   */
  const double* lctnnlo;    /* [Nl][3]   pybird.py:572   */
  const double* l22;        /* [Nl][28]  pybird.py:573-581 */
  const double* l13;        /* [Nl][10]  pybird.py:582   */
  const double* Wf;         /* [front_rows][nin+ntail+ntailx] */
  const double* lr;         /* [ntail]  log(x_i/x_last) of the FFTLog-NFFT tail nodes */
  const double* lrx;        /* [ntailx] same for the 32-point IR-filter FFTLog */
  const double* pair_table; /* [npair][38][2] (re,im) */
  const int32_t* pair_offsets; /* [Nmax+2] */
  const double* Ak;         /* [Nk][2(Nmax+1)] */
  const double* As;         /* [Nl][Ns][2(Nmax+1)] */
  const double* R;          /* [Na][Nkr][Ns]              resum operator */
  const double* q;          /* [2][Nl][Nl][2*NIR*Na][qdeg] Q^{ll'}(f) polynomial coefficients */
  const double* kr2;        /* [Nkr] */
  const double* Cinv;       /* [Nk][Nk]   B-spline collocation inverse */
  const double* knot_lo;    /* [nint] */
  const double* basis;      /* [nint][4][4] */
  const double* mu;         /* [nmu] */
  const double* wl;         /* [Nl][nmu]  2*trapz weight*(2l+1)/2*L_l(mu) */
  const double* project;    /* [nout][Nl*Nk] */
  const double* project_st; /* [nout][Nl*Nk] or NULL: operator of the stochastic terms (21..23) when it differs from
                               `project` (window_st=False window.py:401-403, fiberst=False pybird.py:1798-1806) */
} eftb_constants;

/* ---- library ----------------------------------------------------------------------------------- */
int eftb_abi_version(void);
const char* eftb_last_error(void);
int eftb_padded_batch(int B);
/* number of kernels this library has launched - or recorded into a capturing stream - since it was loaded (bench.py's
   `gpu_launches`: the difference across the capture of one step is the kernel count of the replayed graph) */
unsigned long long eftb_launch_count(void);
/* FP64 FMA-pipe peak probe: runs `iters` dependent-chain DFMA bundles on every SM and returns the
   measured TFLOP/s through *tflops (synchronises; used by bench.py for the roofline denominator) */
int eftb_probe_fp64(int iters, double* tflops, void* stream);

/* ---- plan -------------------------------------------------------------------------------------- */
int eftb_plan_create(const eftb_config* cfg, const eftb_constants* host_constants, eftb_plan** out);
void eftb_plan_destroy(eftb_plan* plan);
/* bytes of scratch `eftb_eval_terms` needs for a batch of B points */
size_t eftb_workspace_bytes(const eftb_plan* plan, int B);

/* ---- layout helpers ---------------------------------------------------------------------------- */
/* out[r][Bp] (batch-minor) <- in[B][R] (point-major); pad lanes replicate point B-1 */
int eftb_to_batch_minor(const double* in, int B, int R, double* out, void* stream);
/* out[B][R] (point-major) <- in[perm ? perm[r] : r][Bp]; perm is a DEVICE int32 array or NULL */
int eftb_to_point_major(const double* in, int B, int R, const int32_t* perm, double* out, void* stream);

/* ---- stage entry points (batch-minor device arrays) ---------------------------------------------
 * F    [front_rows][Bp]           front-end products: c_n (Hermitian half), P11, 13-loop, C11, Cct, X, Y
 * D    [38][Nmax+1][2][Bp]        anti-diagonal sums of c_n c_m M_ch[n,m]
 * P22  [28][Nk][Bp]               bird.P22          (pybird.py:1074-1078)
 * Cs   [Nl][38][Ns][Bp]           bird.C22 (ch<28), bird.C13 (ch>=28), before Legendre weights.  Ns counts the points
 *                                 of the resummation grid: with optiresum every configuration-space array is the
 *                                 extracted BAO peak (Resum.extractBAO, pybird.py:1382-1400) of the reference's
 * T    [Nl][Nk][nterm][Bp]        term index: 0-2 P11l, 3-8 Pctl, 9-20 Ploopl, 21-23 Pstl, 24-26 PctNNLOl
 * Cr   [Bp][Nl][ncr][Ns]          POINT-major; rows: C11, Cct, Cloopl x12 [, CctNNLO]; ncr = 14 + with_nnlo
 */
/* Bird.__init__ interpolation + FFTLog.Coef + IRFilters + makeP13/C11/Cct (pybird.py:694-695,
   :1127-1141, :1080-1101, :1316-1353; fftlog.py:84-166).  plin: point-major [B][nin]. */
int eftb_front(const eftb_plan*, int B, const double* plin, double* u_scratch, double* F, void* stream);
/* the quadratic part of makeP22 / makeC22 / makeC13 (pybird.py:1074-1078, :1103-1125).  eftb_antidiag uses the
   k-space coefficient set (coef_pk, pybird.py:1162), eftb_antidiag_cf the configuration-space one (coef_cf, :1163);
   they coincide unless the plan was built with IRcutoff "loop" or "resum" (eftb_has_cf_set) */
int eftb_antidiag(const eftb_plan*, int B, const double* F, double* D, void* stream);
int eftb_antidiag_cf(const eftb_plan*, int B, const double* F, double* D, void* stream);
int eftb_has_cf_set(const eftb_plan*);
/* D -> P22(k), Dcf -> C22/C13(s); Dcf == NULL: D serves both */
int eftb_spectral(const eftb_plan*, int B, const double* D, const double* Dcf, double* P22, double* Cs, void* stream);
/* fused-path variant: the Legendre weighting + f-power grouping of the configuration-space loop terms
   (reducePsCfl, pybird.py:805-846) is applied to D first (Dg_scratch: [Nl][12][Nmax+1][2][Bp]), so the
   D -> C(s) transform runs on Nl*12 channels and writes Cloopl straight into rows 2..13 of Cr; f: [Bp] */
int eftb_spectral_grouped(const eftb_plan*, int B, const double* D, const double* Dcf, const double* f,
                          double* Dg_scratch, double* P22, double* Cr, void* stream);
/* Bird.setPsCfl / reducePsCfl / setPstl / subtractShotNoise (pybird.py:737-866); f: [Bp].
   Cs == NULL: the Cloopl rows of Cr were already produced by eftb_spectral_grouped */
int eftb_group(const eftb_plan*, int B, const double* F, const double* P22, const double* Cs,
               const double* f, double* T, double* Cr, void* stream);
/* Resum.Ps (pybird.py:1413-1464), in place on T; scratch: eftb_resum_scratch_bytes(plan, B) bytes
   (the bulk coefficients Q^{ll'}(f) of every point, pybird.py:1367-1380) */
size_t eftb_resum_scratch_bytes(const eftb_plan*, int B);
int eftb_resum(const eftb_plan*, int B, const double* F, const double* Cr, const double* f, double* T,
               double* scratch, void* stream);
/* APeffect.AP (pybird.py:1598-1621); DA,H: [Bp]; scratch: eftb_ap_scratch_bytes(plan, B) bytes (B-spline
   coefficients [Nl][Nk][nterm][Bp] + the per-cosmology banded resampling operator); Tout may not alias Tin */
size_t eftb_ap_scratch_bytes(const eftb_plan*, int B);
int eftb_ap(const eftb_plan*, int B, const double* Tin, const double* DA, const double* H,
            double* scratch, double* Tout, void* stream);
/* Window.Window (+ICC) -> Binning.transform -> Chained.transform as one operator; out: [nout][nterm][Bp] */
int eftb_project(const eftb_plan*, int B, const double* T, double* out, void* stream);

/* ---- fused pipeline ------------------------------------------------------------------------------
 * theory.py:557-609 `calculate_power_spectrum` for a batch: plin [B][nin], f/DA/H [B] (point-major,
 * device).  Results (either may be NULL):
 *   terms_bm : batch-minor, [nout][nterm][Bp] if the plan has a projection else [Nl][Nk][nterm][Bp]
 *   terms_pm : point-major [B][Nl_out][nterm][nk_out] (the reference's Bird/PlainBird array order,
 *              concatenated P11l|Pctl|Ploopl|Pstl[|PctNNLOl] along the term axis)
 */
int eftb_eval_terms(const eftb_plan*, int B, const double* plin, const double* f, const double* DA,
                    const double* H, double* terms_bm, double* terms_pm, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Intermediate term arrays of the LAST eftb_eval_terms call that used `workspace` (the reference's Bird snapshots,
 * pybird.py:726-735, taken by Resum / APeffect with snapshot=True, :1463-1464, :1620-1621): stage 0 = after the IR
 * resummation (after setPsCfl when the plan has none), stage 1 = after the AP stage.  terms_bm: [Nl][Nk][nterm][Bp].
 * Nothing is recomputed: the fused pipeline leaves both in its workspace. */
int eftb_workspace_terms(const eftb_plan*, int B, const void* workspace, size_t workspace_bytes, int stage,
                         double* terms_bm, void* stream);

/* ---- standalone fixed operators -------------------------------------------------------------------
 * A single stage of the projection chain applied on its own (Window.integrWindow window.py:371-387,
 * IntegralConstraint.integrWindow icc.py:471-484, Binning.integrBinning binning.py:131-144,
 * Chained.transform chained.py:56-68): C[M][N] = A[M][K] X[K][N] with X, C batch-minor, N % 32 == 0. */
typedef struct eftb_operator eftb_operator;
int eftb_operator_create(int M, int K, const double* host_matrix, eftb_operator** out);
void eftb_operator_destroy(eftb_operator*);
int eftb_operator_apply(const eftb_operator*, const double* X, double* C, int N, void* stream);

/* ---- input side (boltzmann.py:22-101 BoltzmannExtractor: Pkh, f, DA, H per evaluation) ------------------------------
 * A batched producer of the pipeline's inputs on the device, standing where CLASS / CAMB stand in the reference: the
 * Eisenstein & Hu (1998) with-wiggles linear power spectrum of flat LCDM, sigma8-normalised, at redshift z, and the
 * matching growth rate f, angular distance DA (in c / H0) and H / H0 - the same model as eftpipe_b200/synthetic.py.
 * theta: [B][3] = (Omega_m, h, sigma8) point-major; kh: [nk] wavenumbers in h / Mpc; gl_u / gl_w: Gauss-Legendre nodes and
 * weights on [0, 1] (ngl of them) for the growth and distance integrals; nsig: nodes of the sigma8 quadrature on
 * logspace(-4, 2).  Outputs: pkh [B][nk] (point-major, what eftb_eval_terms takes), f, DA, H [B].
 * sigma2 [B] (optional): the variance of the un-normalised spectrum - redshift independent, ten times the work of the nk
 * output nodes.  have_sigma2 = 0: computed and stored there; 1: read from there (the tracers of one evaluation share it). */
int eftb_eh_power(int B, const double* theta, double z, double omega_b, double ns, double Tcmb, const double* kh, int nk,
                  const double* gl_u, const double* gl_w, int ngl, int nsig, double* pkh, double* f, double* DA, double* H,
                  double* sigma2, int have_sigma2, void* stream);

/* ---- likelihood (parambasis.py:42-136,:249-316; likelihood.py:483-594; marginal.py:79-196) ------- */
typedef struct {
  int32_t ntracer, ndata, ngauss, npar, jeffreys;
} eftb_like_config;

typedef struct {
  /* per tracer */
  const int32_t* nout;        /* [ntracer] rows of that tracer's projected terms */
  const int32_t* nterm;       /* [ntracer] */
  const double* scales;       /* [ntracer][6] kmA, krA, ndA, kmB, krB, ndB (parambasis.py:68-75) */
  const int32_t* par_index;   /* [ntracer][EFTB_NPAR] columns of the nuisance array holding
                                 b1A,b2A,b3A,b4A,cctA,cr1A,cr2A, b1B..cr2B, ce0,cemono,cequad, and the NNLO
                                 counterterm coefficients cr4,cr6 (west) / ctilde,- (east; parambasis.py:96-107),
                                 read only when the tracer has 27 term rows (with_NNLO); -1 = 0.0 */
  const int32_t* eastcoast;   /* [ntracer] counterform flag (parambasis.py:102) */
  /* per data point */
  const int32_t* d_tracer;    /* [ndata] */
  const int32_t* d_row;       /* [ndata] row of the tracer's projected terms */
  const double* data;         /* [ndata] */
  const double* picc;         /* [ndata] constant integral-constraint contribution */
  const double* invcov;       /* [ndata][ndata] */
  /* gaussian (marginalised) parameters: dP/dg = sum_{q<3} c_q * v_q * term[i_q] on tracer g_tracer,
     v = b1A^pa * b1B^pb * f^pf coded as g_var = pa | pb << 2 | pf << 4 (pa, pb <= 3, pf <= 7)
     (parambasis.py:249-316, :403-454; the NNLO rows cr4, cr6, ctilde need b1^2 and f^4..f^6) */
  const int32_t* g_count;     /* [ngauss] number of (tracer) entries, <= 2 */
  const int32_t* g_tracer;    /* [ngauss][2] */
  const int32_t* g_term;      /* [ngauss][2][3] */
  const int32_t* g_var;       /* [ngauss][2][3] */
  const double* g_coef;       /* [ngauss][2][3] (0 = unused slot) */
  const double* sigma_inv;    /* [ngauss][ngauss]  (marginal.py:69-77) */
  const double* sigma_inv_mu; /* [ngauss] */
  double mu_sigma_mu;
  const int32_t* d_row_g;     /* [ndata] or NULL (= d_row): row of the projected terms the marginalised-parameter
                                 derivatives PG read.  Differs from d_row for the un-binned interpolated products,
                                 where PNG goes through PlkInterpolator (theory.py:75-106, origin inserted) and PG
                                 through a plain cubic interpolation (likelihood.py:510-513) */
  /* custom EFT bases (parambasis.py:139-162 `EFTBasis` by dotted path, :457-465): a basis the kernel has no formulas for
     hands its bias vector over explicitly.  mode[t] = 1: PNG of tracer t = sum_i nuis[xb_off[t] + i] * term[i] (nterm[t]
     columns of the nuisance array, the coefficient of every term row for every point), and the derivative row of
     Gaussian parameter g on tracer t = sum_i nuis[xg_off[g * ntracer + t] + i] * term[i] (xg_off < 0: none).  All three
     NULL: every tracer uses the built-in West / East-coast formulas (mode 0). */
  const int32_t* mode;        /* [ntracer] or NULL */
  const int32_t* xb_off;      /* [ntracer] or NULL */
  const int32_t* xg_off;      /* [ngauss][ntracer] or NULL */
} eftb_like_constants;

int eftb_like_create(const eftb_like_config*, const eftb_like_constants* host, eftb_like** out);
void eftb_like_destroy(eftb_like*);
size_t eftb_like_workspace_bytes(const eftb_like*, int B);
/* terms[t]: batch-minor projected terms of tracer t ([nout_t][nterm_t][Bp]); fgrowth[t]: [Bp];
 * nuis: batch-minor [npar][Bp].  Outputs (device, [B]): logp, chi2-like pieces and status
 * (0 ok, 1 = F2 not positive definite -> logp = -inf; marginal.py:113-116 raises there).
 * bestfit (optional, point-major [B][ngauss]) = F2^-1 F1 (marginal.py:117). */
int eftb_like_eval(const eftb_like*, int B, const double* const* terms, const double* const* fgrowth,
                   const double* nuis, double* logp, double* bestfit, int32_t* status, void* workspace,
                   size_t workspace_bytes, void* stream);
/* the same plus fullchi2 [B] (optional): chi^2 of the data at the best-fit marginalised parameters, without the prior
 * terms (marginal.py:129-131; EFTLike's derived `{prefix}fullchi2`, likelihood.py:583-590) */
int eftb_like_eval_full(const eftb_like*, int B, const double* const* terms, const double* const* fgrowth,
                        const double* nuis, double* logp, double* bestfit, double* fullchi2, int32_t* status,
                        void* workspace, size_t workspace_bytes, void* stream);
/* the same with a per-point Gaussian prior: prior_loc [B][ngauss] and prior_sigma_inv [B][ngauss] (the diagonal of
 * Sigma^-1 = 1 / scale^2, all zeros for a point whose scales are infinite) replace the plan constants.  This is the
 * reference's callable `loc` / `scale` (strings eval'ed against the sampled EFT parameters at every evaluation,
 * marginal.py:13-20, :60-77, likelihood.py:560-564); both NULL = eftb_like_eval_full */
int eftb_like_eval_priors(const eftb_like*, int B, const double* const* terms, const double* const* fgrowth,
                          const double* nuis, const double* prior_loc, const double* prior_sigma_inv, double* logp,
                          double* bestfit, double* fullchi2, int32_t* status, void* workspace, size_t workspace_bytes,
                          void* stream);
/* out [B][ndata] (point-major) = PNG - data of the LAST eftb_like_eval / eval_full call that used `workspace`
 * (likelihood.py:528-549): read back from the vectors that call left there, nothing is recomputed */
int eftb_like_residuals(const eftb_like*, int B, const void* workspace, double* out, void* stream);
/* un-marginalised pieces for parity tests: vec [B][ndata][ngauss+1] (point-major) with
 * vec[.,d,0] = PNG[d] - data[d] (likelihood.py:528-549) and vec[.,d,1+g] = PG[g][d] (likelihood.py:483-525) */
int eftb_like_vectors(const eftb_like*, int B, const double* const* terms, const double* const* fgrowth,
                      const double* nuis, double* vec, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EFTB200_H */
