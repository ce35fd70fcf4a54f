"""TEST INFRASTRUCTURE ONLY (tests/ and development scripts; the product never imports it).

numpy restatement of the rho-series behind gwsem_b200/csrc/marg_series.cu: by Mehler's expansion the rectangle
probability of an ordinal item given the standardised genotype u is pi(rho, u) = pi(0) sum_k rho^k He_k(u) C_k / k!, so
the polyserial score psi = d log pi / d rho and its products with d log pi / du, d log pi / dz1, d log pi / dz0 are
power series in rho whose coefficients are polynomials in u with per-person coefficients.  `direct` holds the closed
forms (the per-person terms of lavaan's muthen1984 polyserial score, which oracle/marginals_oracle.py restates for the
reference's WLS marginals statistics, R/model.R:413-436 through OpenMx); tests/test_series_math.py checks one against
the other, which is what the truncation orders and the acceptance bound of the CUDA path rest on.
Run as a script it prints the truncation error by order, |rho| and u."""
import math
import numpy as np
from scipy.stats import norm

MMAX = 6
L = MMAX + 3


def hermite(k):
    """coefficients (ascending) of the probabilists' Hermite polynomial He_k"""
    a = [np.array([1.0]), np.array([0.0, 1.0])]
    for n in range(1, k + 1):
        nxt = np.zeros(n + 2)
        nxt[1:] += a[n]
        nxt[:n] -= n * a[n - 1]
        a.append(nxt)
    return a[k]


def he_val(k, x):
    c = hermite(k)
    return sum(c[i] * x ** i for i in range(len(c)))


def smul(A, B, M):
    """product of two series [N, M+1, L] truncated at rho^M"""
    out = np.zeros_like(A)
    for m1 in range(M + 1):
        for m2 in range(M + 1 - m1):
            for l1 in range(L):
                if not np.any(A[:, m1, l1]):
                    continue
                for l2 in range(L - l1):
                    out[:, m1 + m2, l1 + l2] += A[:, m1, l1] * B[:, m2, l2]
    return out


def sdiv(A, R, M):
    """A / R for a series R with constant term exactly 1"""
    Q = np.zeros_like(A)
    for m in range(M + 1):
        acc = A[:, m, :].copy()
        for k in range(1, m + 1):
            for l1 in range(L):
                if not np.any(R[:, k, l1]):
                    continue
                for l2 in range(L - l1):
                    acc[:, l1 + l2] -= R[:, k, l1] * Q[:, m - k, l2]
        Q[:, m, :] = acc
    return Q


def person_series(z1, z0, M):
    N = len(z1)
    f1 = np.where(z1 >= 12, 0.0, norm.pdf(z1))
    f0 = np.where(z0 <= -12, 0.0, norm.pdf(z0))
    pi0 = np.where(z0 > 0, norm.sf(z0) - norm.sf(z1), norm.cdf(z1) - norm.cdf(z0))
    z1c = np.where(z1 >= 12, 0.0, z1)
    z0c = np.where(z0 <= -12, 0.0, z0)
    C = [np.ones(N)] + [(he_val(k - 1, z0c) * f0 - he_val(k - 1, z1c) * f1) / pi0 for k in range(1, M + 3)]
    R = np.zeros((N, M + 1, L)); Rr = np.zeros_like(R); Rrr = np.zeros_like(R); U = np.zeros_like(R)
    Z1 = np.zeros_like(R); Z0 = np.zeros_like(R)
    for k in range(M + 3):
        hk = hermite(k)
        for l, c in enumerate(hk):
            if c == 0:
                continue
            if k <= M:
                R[:, k, l] += c * C[k] / math.factorial(k)
                Z1[:, k, l] += c * f1 / pi0 * he_val(k, z1c) / math.factorial(k)
                Z0[:, k, l] -= c * f0 / pi0 * he_val(k, z0c) / math.factorial(k)
            if 1 <= k <= M + 1:
                Rr[:, k - 1, l] += c * C[k] / math.factorial(k - 1)
            if 2 <= k <= M + 2:
                Rrr[:, k - 2, l] += c * C[k] / math.factorial(k - 2)
        if 1 <= k <= M:
            hk1 = hermite(k - 1)
            for l, c in enumerate(hk1):
                U[:, k, l] += c * C[k] / math.factorial(k - 1)
    psi = sdiv(Rr, R, M)
    dpsi = sdiv(Rrr, R, M) - smul(psi, psi, M)
    F3 = smul(psi, sdiv(U, R, M), M)
    F4 = smul(psi, sdiv(Z1, R, M), M)
    F5 = smul(psi, sdiv(Z0, R, M), M)
    return dict(psi=psi, dpsi=dpsi, F3=F3, F4=F4, F5=F5)


def evaluate(S, rho, u, M):
    out = np.zeros(S.shape[0])
    for m in range(M + 1):
        for l in range(L):
            out += S[:, m, l] * rho ** m * u ** l
    return out


def direct(z1, z0, rho, u):
    s = math.sqrt(1 - rho * rho)
    a1, a0 = (z1 - rho * u) / s, (z0 - rho * u) / s
    f1 = np.where(z1 >= 12, 0.0, norm.pdf(a1)); f0 = np.where(z0 <= -12, 0.0, norm.pdf(a0))
    pi = np.where(a0 > 0, norm.sf(a0) - norm.sf(a1), norm.cdf(a1) - norm.cdf(a0))
    s3 = 1 / s ** 3; r3 = 3 * rho / s ** 2
    g1, g0 = (rho * z1 - u) * s3, (rho * z0 - u) * s3
    h1, h0 = r3 * g1 + z1 * s3, r3 * g0 + z0 * s3
    psi = (f1 * g1 - f0 * g0) / pi
    dpsi = (f1 * (h1 - a1 * g1 * g1) - f0 * (h0 - a0 * g0 * g0)) / pi - psi * psi
    dl_a = f1 / s / pi; dl_b = -f0 / s / pi
    dl_du = -u - rho / s * (f1 - f0) / pi
    return dict(psi=psi, dpsi=dpsi, F3=psi * (dl_du + u), F4=psi * dl_a, F5=psi * dl_b)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    N = 20000
    eta = rng.normal(size=N) * 0.35
    ystar = eta + rng.normal(size=N)
    cuts = np.array([-0.84, -0.25, 0.25, 0.84])
    cat = np.searchsorted(cuts, ystar)
    ext = np.concatenate([[-12.0 + 0], cuts, [12.0]])
    z1 = np.where(cat == len(cuts), 12.0, ext[cat + 1] - eta)
    z0 = np.where(cat == 0, -12.0, ext[cat] - eta)
    S = person_series(z1, z0, MMAX)
    for rho in (0.004, 0.012, 0.03):
        for u in (-0.7, 1.5, 3.0, 6.0):
            d = direct(z1, z0, rho, u)
            row = []
            for M in (2, 3, 4, 5):
                errs = []
                for key in ("psi", "dpsi", "F3", "F4", "F5"):
                    e = evaluate(S[key], rho, u, M) - d[key]
                    errs.append(f"{key} {np.abs(e).mean():.1e}/{np.abs(e).max():.1e}")
                row.append(f"M={M}: " + " ".join(errs))
            print(f"rho={rho} u={u}  |psi| mean {np.abs(d['psi']).mean():.2f}")
            for r in row:
                print("   ", r)
