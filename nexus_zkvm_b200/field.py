"""Tiny pure-Python M31 / CM31 / QM31 helpers for host-side parameter bookkeeping (lookup-element powers,
cumsum shifts).  Field definitions: reference spec zkvm-spec-3.0.pdf §3.1.  Not on any hot path."""
P = (1 << 31) - 1


def m31_inv(a):
    return pow(a % P, P - 2, P)


def cm31_mul(x, y):
    return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def cm31_add(x, y):
    return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)


def cm31_sub(x, y):
    return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)


def cm31_inv(x):
    n = m31_inv((x[0] * x[0] + x[1] * x[1]) % P)
    return ((x[0] * n) % P, (-x[1] * n) % P)


R = (2, 1)


def qm31_mul(x, y):
    a, b, c, d = (x[0], x[1]), (x[2], x[3]), (y[0], y[1]), (y[2], y[3])
    lo = cm31_add(cm31_mul(a, c), cm31_mul(R, cm31_mul(b, d)))
    hi = cm31_add(cm31_mul(a, d), cm31_mul(b, c))
    return (lo[0], lo[1], hi[0], hi[1])


def qm31_add(x, y):
    return tuple((a + b) % P for a, b in zip(x, y))


def qm31_sub(x, y):
    return tuple((a - b) % P for a, b in zip(x, y))


def qm31_neg(x):
    return tuple((-a) % P for a in x)


def qm31_inv(x):
    a, b = (x[0], x[1]), (x[2], x[3])
    denom = cm31_sub(cm31_mul(a, a), cm31_mul(R, cm31_mul(b, b)))
    di = cm31_inv(denom)
    lo = cm31_mul(a, di)
    hi = cm31_mul(b, di)
    return (lo[0], lo[1], (-hi[0]) % P, (-hi[1]) % P)


def qm31_mul_m31(x, s):
    return tuple((a * s) % P for a in x)
