"""TEST INFRASTRUCTURE ONLY (see oracle/README.md) -- parity unpinned.

Matrix kind `Sqrt` (reference: src/gui/matrix.rs:606-613 -> mat_sqrt, :909-985).

The reference does not compute a matrix square root in closed form: it MINIMISES
    cost(X) = sum over the 16 entries of (X*X - M)^2,     X affine (rows 0..2 free = 12 unknowns, row 3 = 0 0 0 1)
starting from X0 = M with a quasi-Newton BFGS iteration (inverse Hessian H0 = I, at most 60 iterations, argmin 0.8's
`BFGS` + `MoreThuenteLineSearch` with c1 = 1e-4, c2 = 0.9), and accepts the last iterate if its cost is below 1e-4,
otherwise the matrix is "missing" ("Can't calculate sqrt!", matrix.rs:610).  Which root is found (a matrix has several)
and whether one is found at all is a property of that minimisation, so this restatement keeps its structure:

  * the same cost, unknowns, start point, H0, iteration cap and acceptance threshold (matrix.rs:927-985);
  * BFGS in its published inverse-Hessian form, stopping on |grad| < sqrt(eps) or |delta cost| < eps (argmin's defaults);
  * the More-Thuente line search as published (MINPACK-2 dcsrch / dcstep, ftol 1e-4, gtol 0.9, first trial step 1);
  * the exact gradient 2 (R X^T + X^T R), R = X*X - M, where the reference differentiates the same cost numerically
    (finitediff forward differences, step sqrt(eps)) -- the reference's iterate carries that O(1e-8) noise, ours does not.

argmin and finitediff are crates.io dependencies absent from /root/reference (Cargo.lock pins argmin 0.8.1, finitediff
0.1.4); their sources are not available here, so this follows the published algorithms, not their code, and is NOT
bit-comparable with the reference: where the minimisation converges (gradient test) both land on the same root to about
1e-8; where it stops on the iteration cap the iterates differ.  One reference scene uses the kind
(portal_in_portal_plus_ultra.ron:961, behind `show_sqrt`).  Every operation below is a plain IEEE double operation in a
fixed order; portal_b200/csrc/host/ph_matsqrt.cpp performs the same operations in the same order and the matrix fuzz
(tests/test_host_frontend.py) requires bit-identical results from the two.
"""
import math

SQRT_EPS = 1.4901161193847656e-08   # f64::EPSILON.sqrt()
EPS = 2.220446049250313e-16         # f64::EPSILON
MAX_ITERS = 60                      # matrix.rs:972
ACCEPT_COST = 1e-4                  # matrix.rs:980
N = 12


def _div(a, b):
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _sqrt(x):
    return math.sqrt(x) if x >= 0.0 else math.nan


def _max(a, b):
    return b if b > a else a


def _min(a, b):
    return b if b < a else a


def _dot(a, b):
    s = 0.0
    for i in range(N):
        s = s + a[i] * b[i]
    return s


def _cost_grad(m, x):
    """m column-major [c][r]; x[4*r + c] = X[r][c] for r < 3 (matrix.rs:927-943: the vector is row-major)."""
    X = [[x[c], x[4 + c], x[8 + c], (1.0 if c == 3 else 0.0)] for c in range(4)]       # X[c][r]
    R = [[0.0] * 4 for _ in range(4)]
    for c in range(4):
        for r in range(4):
            acc = X[0][r] * X[c][0]
            acc = acc + X[1][r] * X[c][1]
            acc = acc + X[2][r] * X[c][2]
            acc = acc + X[3][r] * X[c][3]
            R[c][r] = acc - m[c][r]
    f = 0.0
    for c in range(4):                      # to_cols_array order (matrix.rs:947-951)
        for r in range(4):
            f = f + R[c][r] * R[c][r]
    g = [0.0] * N
    for r in range(3):
        for c in range(4):
            a = 0.0
            for j in range(4):              # (R X^T)[r][c]
                a = a + R[j][r] * X[j][c]
            for i in range(4):              # (X^T R)[r][c]
                a = a + X[r][i] * R[c][i]
            g[4 * r + c] = 2.0 * a
    return f, g


def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * _div(dx, abs(dx))
    if fp > fx:
        theta = _div(3.0 * (fx - fp), stp - stx) + dx + dp
        s = _max(_max(abs(theta), abs(dx)), abs(dp))
        ts = _div(theta, s)
        gamma = s * _sqrt(ts * ts - _div(dx, s) * _div(dp, s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = _div(p, q)
        stpc = stx + r * (stp - stx)
        stpq = stx + _div(_div(dx, _div(fx - fp, stp - stx) + dx), 2.0) * (stp - stx)
        if abs(stpc - stx) < abs(stpq - stx):
            stpf = stpc
        else:
            stpf = stpc + _div(stpq - stpc, 2.0)
        brackt = True
    elif sgnd < 0.0:
        theta = _div(3.0 * (fx - fp), stp - stx) + dx + dp
        s = _max(_max(abs(theta), abs(dx)), abs(dp))
        ts = _div(theta, s)
        gamma = s * _sqrt(ts * ts - _div(dx, s) * _div(dp, s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = _div(p, q)
        stpc = stp + r * (stx - stp)
        stpq = stp + _div(dp, dp - dx) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = _div(3.0 * (fx - fp), stp - stx) + dx + dp
        s = _max(_max(abs(theta), abs(dx)), abs(dp))
        ts = _div(theta, s)
        gamma = s * _sqrt(_max(0.0, ts * ts - _div(dx, s) * _div(dp, s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = _div(p, q)
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + _div(dp, dp - dx) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            if stp > stx:
                stpf = _min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = _max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = _min(stpmax, stpf)
            stpf = _max(stpmin, stpf)
    else:
        if brackt:
            theta = _div(3.0 * (fp - fy), sty - stp) + dy + dp
            s = _max(_max(abs(theta), abs(dy)), abs(dp))
            ts = _div(theta, s)
            gamma = s * _sqrt(ts * ts - _div(dy, s) * _div(dp, s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = _div(p, q)
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


def _line_search(m, x, p, f0, g0):
    """More-Thuente (dcsrch): returns (f, x_new, grad_new) or None when p is not a descent direction."""
    if not (g0 < 0.0):
        return None
    ftol, gtol, xtol = 1e-4, 0.9, 1e-10
    stpmin, stpmax = SQRT_EPS, math.inf
    stp = 1.0
    gtest = ftol * g0
    width = stpmax - stpmin
    width1 = 2.0 * width
    stx, fx, gx = 0.0, f0, g0
    sty, fy, gy = 0.0, f0, g0
    stmin, stmax = 0.0, stp + 4.0 * stp
    brackt, stage = False, 1
    for _ in range(100):
        xn = [x[i] + stp * p[i] for i in range(N)]
        f, gn = _cost_grad(m, xn)
        g = _dot(gn, p)
        ftest = f0 + stp * gtest
        if stage == 1 and f <= ftest and g >= 0.0:
            stage = 2
        if ((brackt and (stp <= stmin or stp >= stmax))
                or (brackt and stmax - stmin <= xtol * stmax)
                or (stp == stpmax and f <= ftest and g <= gtest)
                or (stp == stpmin and (f > ftest or g >= gtest))
                or (f <= ftest and abs(g) <= gtol * (-g0))):
            return f, xn, gn
        if stage == 1 and f <= fx and f > ftest:
            fm, fxm, fym = f - stp * gtest, fx - stx * gtest, fy - sty * gtest
            gm, gxm, gym = g - gtest, gx - gtest, gy - gtest
            stx, fxm, gxm, sty, fym, gym, stp, brackt = _dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax)
            fx, fy = fxm + stx * gtest, fym + sty * gtest
            gx, gy = gxm + gtest, gym + gtest
        else:
            stx, fx, gx, sty, fy, gy, stp, brackt = _dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax)
        if brackt:
            if abs(sty - stx) >= 0.66 * width1:
                stp = stx + 0.5 * (sty - stx)
            width1 = width
            width = abs(sty - stx)
            stmin = _min(stx, sty)
            stmax = _max(stx, sty)
        else:
            stmin = stp + 1.1 * (stp - stx)
            stmax = stp + 4.0 * (stp - stx)
        stp = _max(stp, stpmin)
        stp = _min(stp, stpmax)
        if (brackt and (stp <= stmin or stp >= stmax)) or (brackt and stmax - stmin <= xtol * stmax):
            stp = stx
    xn = [x[i] + stp * p[i] for i in range(N)]
    f, gn = _cost_grad(m, xn)
    return f, xn, gn


def mat_sqrt(m):
    """m: column-major 4x4 ([c][r]); returns X (same layout) with cost(X) < 1e-4, or None (matrix.rs:980-984)."""
    x = [m[c][r] for r in range(3) for c in range(4)]            # mat_to_vec2, matrix.rs:936-943
    H = [[1.0 if i == j else 0.0 for j in range(N)] for i in range(N)]
    f, g = _cost_grad(m, x)
    for _ in range(MAX_ITERS):
        if _sqrt(_dot(g, g)) < SQRT_EPS:
            break
        p = [-_dot(H[i], g) for i in range(N)]
        step = _line_search(m, x, p, f, _dot(g, p))
        if step is None:
            break
        fn, xn, gn = step
        s = [xn[i] - x[i] for i in range(N)]
        y = [gn[i] - g[i] for i in range(N)]
        rho = _div(1.0, _dot(y, s))
        Hy = [_dot(H[i], y) for i in range(N)]
        k = rho * rho * _dot(y, Hy) + rho
        # (I - rho s y^T) H (I - rho y s^T) + rho s s^T, expanded for a symmetric H
        H = [[H[i][j] - rho * (s[i] * Hy[j] + Hy[i] * s[j]) + k * s[i] * s[j] for j in range(N)] for i in range(N)]
        done = abs(f - fn) < EPS
        x, f, g = xn, fn, gn
        if done:
            break
    if not (f < ACCEPT_COST):
        return None
    return [[x[c], x[4 + c], x[8 + c], (1.0 if c == 3 else 0.0)] for c in range(4)]
