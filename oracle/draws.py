"""Counter-based draw contract, Python side (TEST INFRASTRUCTURE ONLY).

Every random decision on the step path is a pure function of
    x = Philox4x32-10(key=(seed_lo, seed_hi), ctr=(env, epoch, site, k >> 2))[k & 3]
with `epoch` a per-env counter bumped once per step()/step_grouped()/
randomize_compromise_and_ownership()/sample_action() call, `site` one of the
SITE_* ids below (each names one RNG call site of the reference, file:line in
the table) and `k` the running count of draws that site has made inside the
epoch.  All transforms are defined on the raw uint32 so that a float64 host and
an integer-only GPU kernel cannot disagree on a branch.

The distributions are the reference's (uniform ints, Bernoulli, triangular,
Poisson, uniform sampling without replacement); only the bit-level mapping from
the generator to the variate is ours, because Mersenne-Twister bit tricks cannot
be mirrored by a parallel kernel.
"""
import math

M32 = 0xFFFFFFFF

# site ids -> reference RNG call site
SITE_STALL = 1        # volt_typhoon_env.py:138   random.randint in _stall
SITE_BLOCK = 2        # volt_typhoon_env.py:505   random.choice(pool) (block edge)
SITE_UNBLOCK = 3      # volt_typhoon_env.py:511   random.choice(pool) (unblock edge)
SITE_ZDAY = 4         # volt_typhoon_env.py:1136  random.choice(owned_indices)
SITE_PROBE = 5        # volt_typhoon_env.py:1189  random.choice(compromised_devices)
SITE_WL_SAMPLE = 6    # CDSimulator.py:298        random.sample(free_candidates, k)
SITE_WL_TRI = 7       # CDSimulator.py:308        np.random.triangular(0, mode, high, 1)
SITE_WL_LAZY = 8      # CDSimulator.py:328        random.random() (M > 500, placement-neutral)
SITE_EV_POISSON = 9   # CyberDefenseEnv.py:668    np.random.poisson(lam)
SITE_EV_ADD = 10      # CyberDefenseEnv.py:679    random.random() < p_add
SITE_EV_PICK = 11     # CyberDefenseEnv.py:675    random.choice(tuple(set)) (sorted ascending first)
SITE_EV_ATT = 12      # CyberDefenseEnv.py:690    random.random() < p_attacker
SITE_SHUFFLE = 13     # volt_typhoon_env.py:359   random.shuffle(non_dcs)
SITE_DETECT = 14      # CDSimulator.py:699,716    random.choice(["A","D"])
SITE_SA_TYPE = 15     # CyberDefenseEnv.py:558/560 action_space.sample()
SITE_SA_NDEV = 16     # CyberDefenseEnv.py:567    random.randint(1, numOfDevice)
SITE_SA_DEVS = 17     # CyberDefenseEnv.py:565    random.sample(keys, k)
SITE_SA_EXP = 18      # CyberDefenseEnv.py:571    random.randrange(MaxExploits)
SITE_SA_APP = 19      # CyberDefenseEnv.py:576    random.randint(0, num_apps - 1)
SITE_WL_ASSIGN = 20   # CDSimulator.py:207        random.random() (unreachable: origin==target is free)

_PM0 = 0xD2511F53
_PM1 = 0xCD9E8D57
_PW0 = 0x9E3779B9
_PW1 = 0xBB67AE85


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11).  ctr: 4 uint32, key: 2 uint32."""
    c0, c1, c2, c3 = [int(c) & M32 for c in ctr]
    k0, k1 = [int(k) & M32 for k in key]
    for r in range(10):
        p0 = _PM0 * c0
        p1 = _PM1 * c2
        hi0, lo0 = p0 >> 32, p0 & M32
        hi1, lo1 = p1 >> 32, p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + _PW0) & M32
        k1 = (k1 + _PW1) & M32
    return (c0, c1, c2, c3)


def draw_u32(seed, env, epoch, site, k):
    key = (seed & M32, (seed >> 32) & M32)
    out = philox4x32_10((env, epoch, site, k >> 2), key)
    return out[k & 3]


def below(x, n):
    """floor(x * n / 2**32): uniform integer in [0, n)."""
    return (int(x) * int(n)) >> 32


def bernoulli_threshold(p):
    """`random.random() < p`  <=>  x < threshold (threshold in [0, 2**32])."""
    p = float(p)
    if p <= 0.0:
        return 0
    if p >= 1.0:
        return 1 << 32
    return min(1 << 32, int(math.ceil(p * 4294967296.0)))


def poisson_table(lam, n=16):
    """Inverse-CDF thresholds: variate = #{j : x >= T[j]}, T[j] = floor(CDF(j) * 2**32)."""
    lam = float(lam)
    out = []
    if lam <= 0.0:
        return [M32] * n  # always 0 events (x >= 2**32-1 has probability 2**-32)
    term = math.exp(-lam)
    cdf = 0.0
    for j in range(n):
        cdf += term
        out.append(min(M32, int(math.floor(min(cdf, 1.0) * 4294967296.0))))
        term = term * lam / (j + 1)
    return out


def poisson_from(x, table):
    return sum(1 for t in table if x >= t)


def triangular_ceil_table(mode, high, n=7):
    """ceil(triangular(0, mode, high)) in 1..high: value = 1 + #{v : x >= T[v-1]}.

    T[v-1] = floor(F(v) * 2**32) for v = 1..high-1 with F the triangular CDF
    (left=0); unused entries are 2**32-1 padded so they never count.
    """
    mode = float(mode)
    high = float(high)
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


def triangular_ceil_from(x, table, high):
    return min(int(high), 1 + sum(1 for t in table if x >= t))
