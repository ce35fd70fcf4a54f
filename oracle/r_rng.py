"""TEST INFRASTRUCTURE ONLY -- a bit-exact restatement of the part of R's random number stack
that the reference's tests use to simulate their phenotypes, so that the numeric pins of
tests/testthat/test-model.R, test-formats.R, test-gxe.R and test-covariate.R (all written as
`set.seed(1)` followed by rbeta / MASS::mvrnorm / rnorm / runif / quantile / cut) become
known-answer tests for the per-SNP fit.

Restated from R's published sources (R >= 3.6 with RNGversion("3.5"), which only changes
sample()):
  * set.seed(seed): RNG.c RNG_Init -- 50 rounds of the LCG 69069*s+1 "initial scrambling",
    then 625 more for i_seed[]; FixupSeeds sets dummy[0] = mti = 624
  * unif_rand(): RNG.c MT_genrand (Matsumoto & Nishimura MT19937) * 2.3283064365386963e-10,
    fixup() into (0, 1)
  * norm_rand(), INVERSION: snorm.c -- u = unif_rand(); u = (int)(2^27 u) + unif_rand();
    qnorm5(u / 2^27)
  * qnorm5: qnorm.c -- Wichura (1988) algorithm AS 241, PPND16
  * rbeta(aa, bb): rbeta.c -- Cheng (1978) algorithms BB (min(a,b) > 1) and BC
  * MASS::mvrnorm(n, mu, Sigma): eigen(Sigma, symmetric=TRUE) (values descending = LAPACK's
    ascending order reversed), X <- matrix(rnorm(p*n), n);  mu + V diag(sqrt(ev)) t(X)
  * quantile(x, probs) type 7, cut(x, breaks, right=TRUE)

Checked in tests/test_r_rng.py against values of R that are public knowledge
(set.seed(1): runif -> 0.2655087 0.3721239 0.5728534; rnorm -> -0.6264538 0.1836433 -0.8356286),
against scipy's ndtri for AS 241, and -- end to end -- by the reference's own pins.

Only tests/ may import this.
"""
import math

import numpy as np

_N, _M = 624, 397
_MATRIX_A, _UPPER, _LOWER = 0x9908B0DF, 0x80000000, 0x7FFFFFFF
_I2_32M1 = 2.328306437080797e-10  # = 1/(2^32 - 1)


class RRng:
    """R's default generators: Mersenne-Twister + Inversion."""

    def __init__(self, seed):
        self.set_seed(seed)

    # RNG.c: RNG_Init + FixupSeeds
    def set_seed(self, seed):
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        buf = []
        for _ in range(_N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            buf.append(s)
        self.mt = buf[1:]
        self.mti = _N            # dummy[0] = 624: the whole state is regenerated at the first draw
        # rbeta.c keeps its set-up constants in statics keyed on (aa, bb)
        self._beta_key = None

    def _genrand(self):
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N - _M):
                y = (mt[kk] & _UPPER) | (mt[kk + 1] & _LOWER)
                mt[kk] = mt[kk + _M] ^ (y >> 1) ^ (_MATRIX_A if y & 1 else 0)
            for kk in range(_N - _M, _N - 1):
                y = (mt[kk] & _UPPER) | (mt[kk + 1] & _LOWER)
                mt[kk] = mt[kk + (_M - _N)] ^ (y >> 1) ^ (_MATRIX_A if y & 1 else 0)
            y = (mt[_N - 1] & _UPPER) | (mt[0] & _LOWER)
            mt[_N - 1] = mt[_M - 1] ^ (y >> 1) ^ (_MATRIX_A if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y * 2.3283064365386963e-10

    def unif_rand(self):
        v = self._genrand()
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def norm_rand(self):
        big = 134217728.0
        u = self.unif_rand()
        u = int(big * u) + self.unif_rand()
        return qnorm(u / big)

    def runif(self, n, lo=0.0, hi=1.0):
        return np.array([lo + (hi - lo) * self.unif_rand() for _ in range(n)])

    def rnorm(self, n, mean=0.0, sd=1.0):
        return np.array([mean + sd * self.norm_rand() for _ in range(n)])

    # rbeta.c
    def rbeta1(self, aa, bb):
        expmax = 1024 * math.log(2.0)
        dblmax = 1.7976931348623157e308
        a, b = min(aa, bb), max(aa, bb)
        alpha = a + b

        def v_w(u1, beta, AA):
            v = beta * math.log(u1 / (1.0 - u1))
            if v <= expmax:
                w = AA * math.exp(v)
                if not math.isfinite(w):
                    w = dblmax
            else:
                w = dblmax
            return v, w

        if a <= 1.0:  # algorithm BC
            if self._beta_key != (aa, bb):
                self._beta_key = (aa, bb)
                beta = 1.0 / a
                delta = 1.0 + b - a
                k1 = delta * (0.0138889 + 0.0416667 * a) / (b * beta - 0.777778)
                k2 = 0.25 + (0.5 + 0.25 / delta) * a
                self._bc = (beta, delta, k1, k2)
            beta, delta, k1, k2 = self._bc
            while True:
                u1 = self.unif_rand()
                u2 = self.unif_rand()
                if u1 < 0.5:
                    y = u1 * u2
                    z = u1 * y
                    if 0.25 * u2 + z - y >= k1:
                        continue
                else:
                    z = u1 * u1 * u2
                    if z <= 0.25:
                        v, w = v_w(u1, beta, b)
                        break
                    if z >= k2:
                        continue
                v, w = v_w(u1, beta, b)
                if alpha * (math.log(alpha / (a + w)) + v) - 1.3862944 >= math.log(z):
                    break
            return a / (a + w) if aa == a else w / (a + w)
        # algorithm BB
        if self._beta_key != (aa, bb):
            self._beta_key = (aa, bb)
            beta = math.sqrt((alpha - 2.0) / (2.0 * a * b - alpha))
            gamma = a + 1.0 / beta
            self._bb = (beta, gamma)
        beta, gamma = self._bb
        while True:
            u1 = self.unif_rand()
            u2 = self.unif_rand()
            v, w = v_w(u1, beta, a)
            z = u1 * u1 * u2
            r = gamma * v - 1.3862944
            s = a + r - w
            if s + 2.609438 >= 5.0 * z:
                break
            t = math.log(z)
            if s > t:
                break
            if not (r + alpha * math.log(alpha / (b + w)) < t):
                break
        return b / (b + w) if aa != a else w / (b + w)

    def rbeta(self, n, a, b):
        return np.array([self.rbeta1(a, b) for _ in range(n)])

    def mvrnorm_identity(self, n, p):
        """MASS::mvrnorm(n, mu = rep(0, p), Sigma = diag(p)).

        eigen(diag(p), symmetric=TRUE): LAPACK dsyevr on the identity returns the unit vectors
        (every Householder reflector of the tridiagonalisation has tau = 0, the tridiagonal matrix
        splits into 1x1 blocks) with the eigenvalues in ascending order; R reverses that order
        (`ord <- rev(seq_along(z$values))`), so `vectors` is the exchange matrix and
        mu + vectors %*% diag(sqrt(ev)) %*% t(X) reverses the columns of X = matrix(rnorm(p*n), n)."""
        X = self.rnorm(p * n).reshape(p, n).T        # matrix(., n) fills column-major
        return X[:, ::-1].copy()


def qnorm(p):
    """qnorm5(p, 0, 1, lower=TRUE, log=FALSE): AS 241 PPND16."""
    if p <= 0.0:
        return -math.inf
    if p >= 1.0:
        return math.inf
    q = p - 0.5
    if abs(q) <= 0.425:
        r = 0.180625 - q * q
        return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                        + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                     + 133.14166789178437745) * r + 3.387132872796366608) / \
            (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r
                 + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
              + 42.313330701600911252) * r + 1.0)
    r = p if q < 0 else 1.0 - p
    r = math.sqrt(-math.log(r))
    if r <= 5.0:
        r -= 1.6
        val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r
                   + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
                + 4.6303378461565452959) * r + 1.42343711074968357734) / \
            (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r
                 + 0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r
              + 2.05319162663775882187) * r + 1.0)
    else:
        r -= 5.0
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r
                   + 0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r
                + 5.4637849111641143699) * r + 6.6579046435011037772) / \
            (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                 + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r
              + 0.59983220655588793769) * r + 1.0)
    return -val if q < 0 else val


def quantile7(x, probs):
    """quantile(x, probs, type = 7) (stats::quantile.default)."""
    x = np.sort(np.asarray(x, float))
    n = len(x)
    out = []
    for pr in np.atleast_1d(probs):
        index = 1 + max(n - 1, 0) * pr
        lo, hi = int(math.floor(index)), int(math.ceil(index))
        q = x[lo - 1]
        if index > lo and x[hi - 1] != q:
            h = index - lo
            q = (1 - h) * q + h * x[hi - 1]
        out.append(q)
    return np.array(out)


def cut_codes(x, breaks):
    """cut(x, breaks, right = TRUE): 0-based level codes, intervals (b_k, b_{k+1}]."""
    b = np.asarray(breaks, float)
    return np.searchsorted(b[1:-1], np.asarray(x, float), side="left")
