"""Synthetic inputs for tests and benchmarks.

No Boltzmann code exists in this image, so the linear power spectrum handed to the hot path
is the Eisenstein & Hu (1998) with-wiggles fitting formula, sigma8-normalised and scaled by
the flat-LCDM growth factor; `f`, `DA`, `H` follow flat LCDM in the same (dimensionless)
conventions as the reference's own helpers (pybird.py:18-42).  The product kernels, the
oracle and the reference all consume the *same arrays* produced here, so the physical
accuracy of the fitting formula is irrelevant to parity.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.integrate import quad

KIN = np.logspace(-5, 0, 200)  # theory.py:562
FIDUCIAL = dict(Om=0.307115, h=0.6777, omega_b=0.02214, ns=0.9611, sigma8=0.8)
RDRAG = 147.66
# tests/compare/default_params.yaml centre, (b1, c2, b3, c4, cct, cr1, cr2, ce0, cemono, cequad)
NUISANCE_CENTRE = dict(b1=2.14, c2=0.776, b3=0.770, c4=0.0, cct=-1.84, cr1=-1.89, cr2=-1.49,
                       ce0=0.260, cemono=0.0, cequad=-0.929)


def eh98_transfer(k_hmpc, Om, Ob, h, Tcmb=2.7255):
    """Eisenstein & Hu 1998 (ApJ 496, 605) eqs. 2-24, baryon wiggles included."""
    k = np.asarray(k_hmpc, dtype=float) * h  # 1/Mpc
    om, ob = Om * h * h, Ob * h * h
    fb = Ob / Om
    fc = 1.0 - fb
    th = Tcmb / 2.7
    zeq = 2.50e4 * om * th**-4
    keq = 7.46e-2 * om * th**-2
    b1 = 0.313 * om**-0.419 * (1 + 0.607 * om**0.674)
    b2 = 0.238 * om**0.223
    zd = 1291 * om**0.251 / (1 + 0.659 * om**0.828) * (1 + b1 * ob**b2)
    R = lambda z: 31.5 * ob * th**-4 * (1e3 / z)
    Req, Rd = R(zeq), R(zd)
    s = 2.0 / (3.0 * keq) * np.sqrt(6.0 / Req) * np.log((np.sqrt(1 + Rd) + np.sqrt(Rd + Req)) / (1 + np.sqrt(Req)))
    ksilk = 1.6 * ob**0.52 * om**0.73 * (1 + (10.4 * om) ** -0.95)
    q = k / (13.41 * keq)
    a1 = (46.9 * om) ** 0.670 * (1 + (32.1 * om) ** -0.532)
    a2 = (12.0 * om) ** 0.424 * (1 + (45.0 * om) ** -0.582)
    alpha_c = a1**-fb * a2 ** (-(fb**3))
    bb1 = 0.944 / (1 + (458 * om) ** -0.708)
    bb2 = (0.395 * om) ** -0.0266
    beta_c = 1.0 / (1 + bb1 * (fc**bb2 - 1))

    def T0(ac, bc):
        C = 14.2 / ac + 386.0 / (1 + 69.9 * q**1.08)
        L = np.log(np.e + 1.8 * bc * q)
        return L / (L + C * q * q)

    fint = 1.0 / (1 + (k * s / 5.4) ** 4)
    Tc = fint * T0(1.0, beta_c) + (1 - fint) * T0(alpha_c, beta_c)
    y = (1 + zeq) / (1 + zd)
    G = y * (-6 * np.sqrt(1 + y) + (2 + 3 * y) * np.log((np.sqrt(1 + y) + 1) / (np.sqrt(1 + y) - 1)))
    alpha_b = 2.07 * keq * s * (1 + Rd) ** -0.75 * G
    beta_node = 8.41 * om**0.435
    beta_b = 0.5 + fb + (3 - 2 * fb) * np.sqrt((17.2 * om) ** 2 + 1)
    st = s / (1 + (beta_node / (k * s)) ** 3) ** (1.0 / 3.0)
    Tb = (T0(1.0, 1.0) / (1 + (k * s / 5.2) ** 2)
          + alpha_b / (1 + (beta_b / (k * s)) ** 3) * np.exp(-((k / ksilk) ** 1.4))) * np.sinc(k * st / np.pi)
    return fb * Tb + fc * Tc


def _E(Om, a):
    return np.sqrt(Om / a + a**2 * (1 - Om))


def growth_factor(Om, a):
    return 2.5 * Om * _E(Om, a) / a * quad(lambda x: _E(Om, x) ** -3, 0, a)[0]


def growth_rate(Om, z):
    a = 1.0 / (1.0 + z)
    D = growth_factor(Om, a)
    return (Om * (5 * a - 3 * D)) / (2.0 * (a**3 * (1 - Om) + Om) * D)


def hubble(Om, z):
    return (Om * (1 + z) ** 3 + (1 - Om)) ** 0.5


def angular_distance(Om, z):
    return quad(lambda x: 1.0 / hubble(Om, x), 0, z)[0] / (1 + z)


def _sigma8_unnorm(Om, Ob, h, ns):
    k = np.logspace(-4, 2, 2000)
    T = eh98_transfer(k, Om, Ob, h)
    x = 8.0 * k
    W = 3 * (np.sin(x) - x * np.cos(x)) / x**3
    integrand = k**3 * k**ns * T**2 * W**2 / (2 * np.pi**2)
    return np.sqrt(np.trapezoid(integrand, np.log(k)))


def linear_power(kh, Om, h, sigma8, z, omega_b=FIDUCIAL["omega_b"], ns=FIDUCIAL["ns"]):
    Ob = omega_b / h**2
    norm = (sigma8 / _sigma8_unnorm(Om, Ob, h, ns)) ** 2
    D = growth_factor(Om, 1.0 / (1 + z)) / growth_factor(Om, 1.0)
    return norm * D**2 * kh**ns * eh98_transfer(kh, Om, Ob, h) ** 2


@dataclass
class SyntheticBatch:
    kin: np.ndarray  # (200,)
    plin: np.ndarray  # (B, 200)
    f: np.ndarray  # (B,)
    DA: np.ndarray  # (B,)
    H: np.ndarray  # (B,)
    h: np.ndarray  # (B,)
    rdrag: np.ndarray  # (B,)
    theta: np.ndarray  # (B, 3) Om, h, sigma8

    def __len__(self):
        return self.plin.shape[0]


def draw_cosmologies(B, seed, fiducial_first=True):
    rng = np.random.default_rng(seed)
    out = []
    if fiducial_first:
        out.append((FIDUCIAL["Om"], FIDUCIAL["h"], FIDUCIAL["sigma8"]))
    while len(out) < B:
        Om = rng.normal(FIDUCIAL["Om"], 0.02)
        h = rng.normal(FIDUCIAL["h"], 0.02)
        s8 = rng.normal(FIDUCIAL["sigma8"], 0.05)
        if 0.2 <= Om <= 0.45:
            out.append((Om, h, s8))
    return np.array(out[:B])


def make_batch(B, z, seed=20261018, unique=None):
    """B synthetic cosmologies at redshift z.  `unique` (< B) computes only that many distinct
    spectra and tiles them with a smooth per-point rescaling (keeps large benchmark batches
    cheap to generate while every point still carries different numbers)."""
    nu = B if unique is None else min(unique, B)
    theta = draw_cosmologies(nu, seed)
    plin = np.empty((nu, KIN.size))
    f = np.empty(nu)
    DA = np.empty(nu)
    H = np.empty(nu)
    for i, (Om, h, s8) in enumerate(theta):
        plin[i] = linear_power(KIN, Om, h, s8, z)
        f[i] = growth_rate(Om, z)
        DA[i] = angular_distance(Om, z)
        H[i] = hubble(Om, z)
    if nu < B:
        reps = -(-B // nu)
        idx = np.tile(np.arange(nu), reps)[:B]
        scale = 1.0 + 0.02 * np.sin(0.37 * np.arange(B))
        plin = plin[idx] * scale[:, None]
        f, DA, H, theta = f[idx], DA[idx], H[idx], theta[idx]
    return SyntheticBatch(kin=KIN.copy(), plin=plin, f=f, DA=DA, H=H, h=theta[:, 1].copy(),
                          rdrag=np.full(B, RDRAG), theta=theta)


def make_batch_fast(B, z, seed=20261018):
    """B DISTINCT synthetic cosmologies, vectorised (benchmark-size batches: `make_batch` runs scipy.quad per point).
    Same formulas as `make_batch`; the growth and distance integrals use Gauss-Legendre nodes instead of adaptive
    quadrature (agreement ~1e-12, tests/test_host_mirror.py), the sigma8 normalisation the same 2000-node trapezoid."""
    theta = draw_cosmologies(B, seed)
    Om, h, s8 = theta[:, 0:1], theta[:, 1:2], theta[:, 2:3]
    Ob = FIDUCIAL["omega_b"] / h**2
    ns = FIDUCIAL["ns"]
    u, w = np.polynomial.legendre.leggauss(96)
    u, w = 0.5 * (u + 1.0), 0.5 * w  # nodes on [0, 1]

    def growth(a):  # x = a t^2 removes the x^(3/2) end-point behaviour of E(x)^-3
        x = a * u[None, :] ** 2
        integral = np.sum(w[None, :] * 2.0 * a * u[None, :] * _E(Om, x) ** -3, axis=1, keepdims=True)
        return 2.5 * Om * _E(Om, a) / a * integral

    a = 1.0 / (1.0 + z)
    D, D0 = growth(a), growth(1.0)
    f = (Om * (5 * a - 3 * D)) / (2.0 * (a**3 * (1 - Om) + Om) * D)
    zz = z * u[None, :]
    DA = np.sum(w[None, :] * z / hubble(Om, zz), axis=1, keepdims=True) / (1 + z)
    k = np.logspace(-4, 2, 2000)[None, :]
    T = eh98_transfer(k, Om, Ob, h)
    x = 8.0 * k
    W = 3 * (np.sin(x) - x * np.cos(x)) / x**3
    sig = np.sqrt(np.trapezoid(k**3 * k**ns * T**2 * W**2 / (2 * np.pi**2), np.log(k[0]), axis=1))[:, None]
    plin = (s8 / sig) ** 2 * (D / D0) ** 2 * KIN[None, :] ** ns * eh98_transfer(KIN[None, :], Om, Ob, h) ** 2
    return SyntheticBatch(kin=KIN.copy(), plin=np.ascontiguousarray(plin), f=f[:, 0].copy(), DA=DA[:, 0].copy(),
                          H=hubble(Om, z)[:, 0].copy(), h=theta[:, 1].copy(), rdrag=np.full(B, RDRAG), theta=theta)


def draw_nuisance(B, seed=20261018, scatter=0.05):
    """(B, 10) nuisance vectors (b1, c2, b3, c4, cct, cr1, cr2, ce0, cemono, cequad) around the
    reference's default centre."""
    rng = np.random.default_rng(seed + 7)
    centre = np.array(list(NUISANCE_CENTRE.values()))
    width = scatter * np.maximum(np.abs(centre), 0.5)
    out = centre[None, :] + width[None, :] * rng.standard_normal((B, centre.size))
    out[0] = centre
    return out


def c2c4_to_b2b4(c2, c4):
    return (c2 + c4) / np.sqrt(2.0), (c2 - c4) / np.sqrt(2.0)
