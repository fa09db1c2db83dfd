"""Integer tables that turn a raw Philox uint32 into the variates the step path needs.

The reference draws from `random` / `np.random` (SURVEY.md section 8c lists every call site).
The kernels consume counter-based draws x = Philox4x32-10(seed; env, epoch, site, k) instead and
map x to the SAME distributions with integer comparisons only, so that a float64 host and an
integer GPU kernel cannot disagree on a branch:

  random() < p            <=>  x < ceil(p * 2**32)                       (CyberDefenseEnv.py:679, :690)
  np.random.poisson(lam)   =   #{j : x >= floor(CDF(j) * 2**32)}          (CyberDefenseEnv.py:668)
  ceil(triangular(0,m,h))  =   min(h, 1 + #{v : x >= floor(F(v) * 2**32)}) (CDSimulator.py:308)
  randint / choice / sample / shuffle: floor(x * n / 2**32) picks, in the reference's order
"""
import math

M32 = 0xFFFFFFFF


def bernoulli_threshold(p):
    p = float(p)
    if p <= 0.0:
        return 0
    if p >= 1.0:
        return 1 << 32
    return min(1 << 32, int(math.ceil(p * 4294967296.0)))


def poisson_table(lam, n=16):
    lam = float(lam)
    if lam <= 0.0:
        return [M32] * n
    out, term, cdf = [], math.exp(-lam), 0.0
    for j in range(n):
        cdf += term
        out.append(min(M32, int(math.floor(min(cdf, 1.0) * 4294967296.0))))
        term = term * lam / (j + 1)
    return out


def triangular_ceil_table(mode, high, n=8):
    mode, high = float(mode), float(high)
    out = []
    for v in range(1, n + 1):
        if v >= high:
            out.append(M32)
            continue
        if v <= mode:
            f = (v * v) / (high * mode)
        else:
            f = 1.0 - ((high - v) ** 2) / (high * (high - mode))
        out.append(min(M32, int(math.floor(f * 4294967296.0))))
    return out
