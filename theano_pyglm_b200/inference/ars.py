"""Derivative-free adaptive rejection sampling for a 1-D log-concave density (Gilks 1992).

Stands in for `hips.inference.ars.adaptive_rejection_sample`, which the reference imports
(inference/gibbs.py:14) but does not vendor; same call shape:
    adaptive_rejection_sample(func, xs, v_xs, domain, stepsz=..., debug=False)
`func` may be expensive (each call is a pass over the data on the GPU), so evaluations are few:
the hull starts from the points the caller already has (the Gauss-Hermite abscissae).
"""
import numpy as np


def _hull(xs, hs, lb, ub):
    """Piecewise-linear upper hull from chords of a concave function.

    On [x_i, x_{i+1}] concavity bounds h by the extensions of both neighbouring chords; their pointwise
    minimum is linear up to their intersection z_i.  Returns the segments as (left, right, x0, h0, slope)
    where the bound on [left, right] is h0 + slope*(x - x0)."""
    n = len(xs)
    slopes = np.diff(hs) / np.diff(xs)               # chord i joins points i, i+1
    segs = []
    if np.isinf(lb):                                  # left tail: chord 0 extended (needs slope > 0)
        segs.append((-np.inf, xs[0], xs[0], hs[0], slopes[0]))
    elif lb < xs[0]:
        segs.append((lb, xs[0], xs[0], hs[0], slopes[0]))
    for i in range(n - 1):
        left_ok, right_ok = i - 1 >= 0, i + 1 <= n - 2
        if left_ok and right_ok:
            sl, sr = slopes[i - 1], slopes[i + 1]
            # lines: hs[i] + sl*(x - xs[i])  and  hs[i+1] + sr*(x - xs[i+1])
            if sl - sr > 1e-300:
                z = (hs[i + 1] - hs[i] + sl * xs[i] - sr * xs[i + 1]) / (sl - sr)
                z = min(max(z, xs[i]), xs[i + 1])
            else:
                z = 0.5 * (xs[i] + xs[i + 1])
            segs.append((xs[i], z, xs[i], hs[i], sl))
            segs.append((z, xs[i + 1], xs[i + 1], hs[i + 1], sr))
        elif left_ok:
            segs.append((xs[i], xs[i + 1], xs[i], hs[i], slopes[i - 1]))
        elif right_ok:
            segs.append((xs[i], xs[i + 1], xs[i + 1], hs[i + 1], slopes[i + 1]))
        else:                                        # only two points: bound by the larger value
            segs.append((xs[i], xs[i + 1], xs[i], max(hs[i], hs[i + 1]), 0.0))
    if np.isinf(ub):
        segs.append((xs[-1], np.inf, xs[-1], hs[-1], slopes[-1]))
    elif ub > xs[-1]:
        segs.append((xs[-1], ub, xs[-1], hs[-1], slopes[-1]))
    return [s for s in segs if s[1] > s[0]]


def _seg_logmass(seg):
    l, r, x0, h0, m = seg
    if abs(m) < 1e-12:
        return h0 + np.log(r - l)
    a = h0 + m * (l - x0) if np.isfinite(l) else -np.inf
    b = h0 + m * (r - x0) if np.isfinite(r) else -np.inf
    hi, lo = max(a, b), min(a, b)
    return hi + np.log1p(-np.exp(lo - hi)) - np.log(abs(m))


def _sample_seg(seg, u):
    l, r, x0, h0, m = seg
    if abs(m) < 1e-12:
        return l + u * (r - l)
    if m > 0:                                         # density grows to the right end (finite)
        span = -np.inf if not np.isfinite(l) else m * (l - r)
        return r + np.log(u + (1 - u) * np.exp(span)) / m
    span = -np.inf if not np.isfinite(r) else m * (r - l)
    return l + np.log((1 - u) + u * np.exp(span)) / m


def adaptive_rejection_sample(func, xs, v_xs, domain, stepsz=1.0, debug=False, rng=None, max_evals=200):
    """One draw from exp(func)."""
    rng = np.random if rng is None else rng
    lb, ub = domain
    order = np.argsort(xs)
    xs = list(np.asarray(xs, dtype=np.float64)[order])
    hs = list(np.asarray(v_xs, dtype=np.float64)[order])
    keep = [i for i in range(len(xs)) if np.isfinite(hs[i]) and (i == 0 or xs[i] > xs[i - 1])]
    xs, hs = [xs[i] for i in keep], [hs[i] for i in keep]
    evals = 0
    if len(xs) == 0:
        x0 = 0.0 if np.isinf(lb) or np.isinf(ub) else 0.5 * (lb + ub)
        xs, hs = [x0], [float(func(x0))]
        evals += 1
    while len(xs) < 3:
        xn = xs[-1] + stepsz
        xs.append(xn); hs.append(float(func(xn))); evals += 1
    # the unbounded tails need an increasing first chord and a decreasing last chord
    while np.isinf(lb) and hs[1] - hs[0] <= 0 and evals < max_evals:
        xn = xs[0] - stepsz * (1 + evals)
        xs.insert(0, xn); hs.insert(0, float(func(xn))); evals += 1
    while np.isinf(ub) and hs[-1] - hs[-2] >= 0 and evals < max_evals:
        xn = xs[-1] + stepsz * (1 + evals)
        xs.append(xn); hs.append(float(func(xn))); evals += 1
    while True:
        xa, ha = np.array(xs), np.array(hs)
        segs = _hull(xa, ha - ha.max(), lb, ub)
        lm = np.array([_seg_logmass(s) for s in segs])
        p = np.exp(lm - lm.max())
        k = int(np.searchsorted(np.cumsum(p / p.sum()), rng.rand()))
        k = min(k, len(segs) - 1)
        x = float(_sample_seg(segs[k], rng.rand()))
        l, r, x0, h0, m = segs[k]
        upper = h0 + m * (x - x0) + ha.max()
        i = int(np.searchsorted(xa, x))
        lw = -np.inf                                   # squeeze: the chord under the curve
        if 0 < i < len(xa):
            lw = ha[i - 1] + (ha[i] - ha[i - 1]) * (x - xa[i - 1]) / (xa[i] - xa[i - 1])
        logu = np.log(rng.rand())
        if logu <= lw - upper:
            return x
        hx = float(func(x)); evals += 1
        if logu <= hx - upper or evals >= max_evals:
            return x
        if np.isfinite(hx) and x not in xs:
            xs.insert(i, x); hs.insert(i, hx)
