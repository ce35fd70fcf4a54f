"""The Mehler series behind the class-sum statistics path (gwsem_b200/csrc/marg_series.cu) against the closed forms of
the polyserial score terms, on CPU: the truncation orders the kernel uses (weighted sums at rho^3, the score root at
rho^5) meet the accuracy its acceptance bound assumes."""
import numpy as np
import series_oracle as so


def people(n, seed):
    rng = np.random.default_rng(seed)
    eta = rng.normal(size=n) * 0.35
    cuts = np.array([-0.84, -0.25, 0.25, 0.84])
    cat = np.searchsorted(cuts, eta + rng.normal(size=n))
    ext = np.concatenate([[-12.0], cuts, [12.0]])
    z1 = np.where(cat == len(cuts), 12.0, ext[cat + 1] - eta)
    z0 = np.where(cat == 0, -12.0, ext[cat] - eta)
    return z1, z0


def test_series_reproduces_closed_forms():
    z1, z0 = people(4000, 1)
    S = so.person_series(z1, z0, so.MMAX)
    for rho, u in [(0.004, -0.7), (0.012, 1.5), (0.012, 6.0), (-0.02, 3.0)]:
        d = so.direct(z1, z0, rho, u)
        for key in ("psi", "dpsi", "F3", "F4", "F5"):
            full = so.evaluate(S[key], rho, u, so.MMAX)
            assert np.max(np.abs(full - d[key])) < 1e-9 * max(1.0, abs(u)) ** 2, (rho, u, key)


def test_truncation_orders_of_the_kernel():
    z1, z0 = people(4000, 2)
    S = so.person_series(z1, z0, so.MMAX)
    for rho, u, tol3, tol5 in [(0.005, 1.5, 2e-8, 1e-12), (0.012, 1.5, 1e-6, 3e-10), (0.012, 6.0, 3e-5, 1e-8)]:
        d = so.direct(z1, z0, rho, u)
        scale = np.abs(d["psi"]).mean()
        for key in ("psi", "F3", "F4", "F5"):      # order 3: the SNP-dependent rows of Muthen's A and B
            assert np.abs(so.evaluate(S[key], rho, u, 3) - d[key]).mean() < tol3 * scale * 30, (rho, u, key)
        assert np.abs(so.evaluate(S["psi"], rho, u, 5) - d["psi"]).mean() < tol5 * scale * 30, (rho, u)   # order 5: the score root


def test_first_neglected_term_is_the_error_estimate():
    """The acceptance bound of marg_series_kernel takes rho^4 * rms(coefficient of rho^4) for the error of the order-3
    series; check that it is within a factor of two of the real thing."""
    z1, z0 = people(4000, 3)
    S = so.person_series(z1, z0, so.MMAX)
    for rho, u in [(0.01, 1.2), (0.02, -0.8), (0.015, 4.0)]:
        d = so.direct(z1, z0, rho, u)
        err = np.sqrt(np.mean((so.evaluate(S["psi"], rho, u, 3) - d["psi"]) ** 2))
        coef = sum(S["psi"][:, 4, l] * u ** l for l in range(so.L))
        est = rho ** 4 * np.sqrt(np.mean(coef ** 2))
        assert 0.5 < err / est < 2.0, (rho, u, err, est)
