"""ORACLE — test infrastructure only (never imported by pi3_slam_b200/).

Restatement of what the reference's focal / shift solve actually executes (utils/geometry_numpy.py:79-96:
`least_squares(fn, x0=0, ftol=1e-3, method='lm')`), for ONE parameter: scipy (>= 1.16, here 1.18.1) forms the Jacobian by
its own 2-point finite differences (step sqrt(eps) * sign(x) * max(1, |x|), scipy/optimize/_numdiff.py) and drives MINPACK
`lmder` with internal variable scaling (x_scale='jac' -> mode 1), factor 100, xtol = gtol = 1e-8, max_nfev = 100.
`lmder` / `lmpar` / `qrsolv` are restated below for n = 1 (MINPACK, Argonne 1980: Moré, Garbow, Hillstrom), where every
matrix is a scalar: R = -||J||, Q^T f = -(J.f)/||J|| (Householder signs cancel everywhere they are used).

The device solver (pi3b_focal_shift_lm, csrc/post.cu) is a line-by-line port of `lm_focal_shift`; tests pin THIS file
against scipy itself (tests/test_oracle.py) and the kernel against this file and against scipy (tests/test_creator_gpu.py).
"""
from __future__ import annotations

import numpy as np

EPSMCH = np.finfo(np.float64).eps
DWARF = np.finfo(np.float64).tiny
FACTOR, XTOL, GTOL, MAXFEV = 100.0, 1e-8, 1e-8, 100


def residual(uv, xy, z, shift):
    """fn of geometry_numpy.py:84-88 (uv, xy (S,2) and z (S,) float32; shift float64 -> float64 arithmetic)."""
    xy_proj = xy / (z + shift)[:, None]
    f = (xy_proj * uv).sum() / np.square(xy_proj).sum()
    return (f * xy_proj - uv).ravel()


def _qrsolv1(r, diag_in, qtb):
    sdiag = diag_in
    wa = qtb
    rn = r
    if sdiag != 0.0:
        if abs(r) < abs(sdiag):
            cotan = r / sdiag
            sin = 0.5 / np.sqrt(0.25 + 0.25 * cotan * cotan)
            cos = sin * cotan
        else:
            tan = sdiag / r
            cos = 0.5 / np.sqrt(0.25 + 0.25 * tan * tan)
            sin = cos * tan
        rn = cos * r + sin * sdiag
        wa = cos * qtb
    x = wa / rn if rn != 0.0 else 0.0
    return x, rn


def _lmpar1(r, d, qtb, delta, par):
    x = qtb / r if r != 0.0 else 0.0
    dxnorm = abs(d * x)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:
        return 0.0, x
    parl = 0.0
    if r != 0.0:
        wa1 = d * (d * x / dxnorm) / r
        parl = ((fp / delta) / abs(wa1)) / abs(wa1)
    gnorm = abs(r * qtb / d)
    paru = gnorm / delta
    if paru == 0.0:
        paru = DWARF / min(delta, 0.1)
    par = min(max(par, parl), paru)
    if par == 0.0:
        par = gnorm / dxnorm
    it = 0
    while True:
        it += 1
        if par == 0.0:
            par = max(DWARF, 0.001 * paru)
        x, sdiag = _qrsolv1(r, np.sqrt(par) * d, qtb)
        dxnorm = abs(d * x)
        temp = fp
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta or (parl == 0.0 and fp <= temp and temp < 0.0) or it == 10:
            break
        wa1 = d * (d * x / dxnorm) / sdiag
        parc = ((fp / delta) / abs(wa1)) / abs(wa1)
        if fp > 0.0:
            parl = max(parl, par)
        if fp < 0.0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return par, x


def lm_focal_shift(uv, xyz, ftol=1e-3):
    """(shift float32, focal, nfev) as utils/geometry_numpy.py:79-96 returns them (plus the evaluation count)."""
    uv = np.asarray(uv, np.float32).reshape(-1, 2)
    xy = np.asarray(xyz, np.float32)[..., :2].reshape(-1, 2)
    z = np.asarray(xyz, np.float32)[..., 2].reshape(-1)
    fun = lambda s: residual(uv, xy, z, np.float64(s))         # noqa: E731
    x = 0.0
    fvec = fun(x)
    nfev = 1
    fnorm = float(np.sqrt(fvec @ fvec))
    par = 0.0
    it = 1
    diag = delta = xnorm = 0.0
    info = 0
    while info == 0:
        # scipy's 2-point forward difference
        h = np.sqrt(EPSMCH) * (1.0 if x >= 0 else -1.0) * max(1.0, abs(x))
        dx = (x + h) - x
        J = (fun(x + h) - fvec) / dx
        jnorm = float(np.sqrt(J @ J))
        r = -jnorm                                   # qrfac of one column (sign convention irrelevant below)
        qtf = -float(J @ fvec) / jnorm if jnorm != 0.0 else 0.0
        if it == 1:
            diag = jnorm if jnorm != 0.0 else 1.0
            xnorm = abs(diag * x)
            delta = FACTOR * xnorm
            if delta == 0.0:
                delta = FACTOR
        gnorm = 0.0
        if fnorm != 0.0 and jnorm != 0.0:
            gnorm = abs(r * (qtf / fnorm) / jnorm)
        if gnorm <= GTOL:
            info = 4
            break
        diag = max(diag, jnorm)
        while True:
            par, p = _lmpar1(r, diag, qtf, delta, par)
            p = -p
            x_new = x + p
            pnorm = abs(diag * p)
            if it == 1:
                delta = min(delta, pnorm)
            f_new = fun(x_new)
            nfev += 1
            fnorm1 = float(np.sqrt(f_new @ f_new))
            actred = -1.0
            if 0.1 * fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            temp1 = abs(r * p) / fnorm
            temp2 = np.sqrt(par) * pnorm / fnorm
            prered = temp1 ** 2 + temp2 ** 2 / 0.5
            dirder = -(temp1 ** 2 + temp2 ** 2)
            ratio = actred / prered if prered != 0.0 else 0.0
            if ratio <= 0.25:
                temp = 0.5 if actred >= 0.0 else 0.5 * dirder / (dirder + 0.5 * actred)
                if 0.1 * fnorm1 >= fnorm or temp < 0.1:
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par /= temp
            elif par == 0.0 or ratio >= 0.75:
                delta = pnorm / 0.5
                par *= 0.5
            if ratio >= 1e-4:
                x, fvec, fnorm = x_new, f_new, fnorm1
                xnorm = abs(diag * x)
                it += 1
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0:
                info = 1
            if delta <= XTOL * xnorm:
                info = 2
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0 and info == 2:
                info = 3
            if info != 0:
                break
            if nfev >= MAXFEV:
                info = 5
            if abs(actred) <= EPSMCH and prered <= EPSMCH and 0.5 * ratio <= 1.0:
                info = 6
            if delta <= EPSMCH * xnorm:
                info = 7
            if gnorm <= EPSMCH:
                info = 8
            if info != 0 or ratio >= 1e-4:
                break
    shift = np.float32(x)
    xy_proj = xy / (z + shift)[:, None]                              # float32 arithmetic, as the reference
    focal = (xy_proj * uv).sum() / np.square(xy_proj).sum()
    return shift, focal, nfev
