// Matrix kind `Sqrt` (reference: src/gui/matrix.rs:606-613 -> mat_sqrt, :909-985).
//
// The reference MINIMISES cost(X) = sum over the 16 entries of (X*X - M)^2 over affine X (rows 0..2 free, 12 unknowns)
// from X0 = M with argmin 0.8's BFGS (H0 = I, at most 60 iterations) and More-Thuente line search (c1 1e-4, c2 0.9),
// and accepts the last iterate when its cost is below 1e-4; otherwise the matrix is missing ("Can't calculate sqrt!").
// This keeps that structure -- same cost, unknowns, start, H0, cap and threshold; BFGS in its published inverse-Hessian
// form with argmin's default stops (|grad| < sqrt(eps), |delta cost| < eps); the More-Thuente search as published
// (MINPACK-2 dcsrch / dcstep) -- with the exact gradient 2 (R X^T + X^T R), R = X*X - M, where the reference takes
// forward differences of the same cost.  argmin / finitediff are crates.io dependencies whose sources are absent, so
// this is not bit-comparable with the reference: where the minimisation converges both reach the same root to ~1e-8.
// oracle/mat_sqrt.py performs the same IEEE double operations in the same order (tests require identical bits).
#include <cmath>
#include <limits>

#include "ph_scene.h"

namespace ph {
namespace {

constexpr int N = 12;
constexpr double kSqrtEps = 1.4901161193847656e-08;
constexpr double kEps = 2.220446049250313e-16;
constexpr double kInf = std::numeric_limits<double>::infinity();

inline double max2(double a, double b) { return b > a ? b : a; }
inline double min2(double a, double b) { return b < a ? b : a; }

double dot(const double* a, const double* b) {
    double s = 0.0;
    for (int i = 0; i < N; i++) s = s + a[i] * b[i];
    return s;
}

// x[4*r + c] = X[r][c], r < 3 (the reference's parameter vector is row-major, matrix.rs:927-943)
double cost_grad(const Mat4& m, const double* x, double* g) {
    double X[4][4], R[4][4];  // [c][r]
    for (int c = 0; c < 4; c++) {
        X[c][0] = x[c];
        X[c][1] = x[4 + c];
        X[c][2] = x[8 + c];
        X[c][3] = c == 3 ? 1.0 : 0.0;
    }
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            double acc = X[0][r] * X[c][0];
            acc = acc + X[1][r] * X[c][1];
            acc = acc + X[2][r] * X[c][2];
            acc = acc + X[3][r] * X[c][3];
            R[c][r] = acc - m[4 * c + r];
        }
    double f = 0.0;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) f = f + R[c][r] * R[c][r];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double a = 0.0;
            for (int j = 0; j < 4; j++) a = a + R[j][r] * X[j][c];
            for (int i = 0; i < 4; i++) a = a + X[r][i] * R[c][i];
            g[4 * r + c] = 2.0 * a;
        }
    return f;
}

struct Interval {
    double stx, fx, dx, sty, fy, dy, stp;
    bool brackt;
};

// MINPACK-2 dcstep: the safeguarded cubic / quadratic trial step and the interval update
void dcstep(Interval& v, double fp, double dp, double stpmin, double stpmax) {
    double &stx = v.stx, &fx = v.fx, &dx = v.dx, &sty = v.sty, &fy = v.fy, &dy = v.dy, &stp = v.stp;
    const double sgnd = dp * (dx / std::fabs(dx));
    double stpf;
    if (fp > fx) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = max2(max2(std::fabs(theta), std::fabs(dx)), std::fabs(dp));
        const double ts = theta / s;
        double gamma = s * std::sqrt(ts * ts - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        const double p = (gamma - dx) + theta;
        const double q = ((gamma - dx) + gamma) + dp;
        const double r = p / q;
        const double stpc = stx + r * (stp - stx);
        const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (std::fabs(stpc - stx) < std::fabs(stpq - stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        v.brackt = true;
    } else if (sgnd < 0.0) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = max2(max2(std::fabs(theta), std::fabs(dx)), std::fabs(dp));
        const double ts = theta / s;
        double gamma = s * std::sqrt(ts * ts - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta;
        const double q = ((gamma - dp) + gamma) + dx;
        const double r = p / q;
        const double stpc = stp + r * (stx - stp);
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = std::fabs(stpc - stp) > std::fabs(stpq - stp) ? stpc : stpq;
        v.brackt = true;
    } else if (std::fabs(dp) < std::fabs(dx)) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = max2(max2(std::fabs(theta), std::fabs(dx)), std::fabs(dp));
        const double ts = theta / s;
        double gamma = s * std::sqrt(max2(0.0, ts * ts - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta;
        const double q = (gamma + (dx - dp)) + gamma;
        const double r = p / q;
        double stpc;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (v.brackt) {
            stpf = std::fabs(stpc - stp) < std::fabs(stpq - stp) ? stpc : stpq;
            if (stp > stx) stpf = min2(stp + 0.66 * (sty - stp), stpf);
            else stpf = max2(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = std::fabs(stpc - stp) > std::fabs(stpq - stp) ? stpc : stpq;
            stpf = min2(stpmax, stpf);
            stpf = max2(stpmin, stpf);
        }
    } else {
        if (v.brackt) {
            const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            const double s = max2(max2(std::fabs(theta), std::fabs(dy)), std::fabs(dp));
            const double ts = theta / s;
            double gamma = s * std::sqrt(ts * ts - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            const double p = (gamma - dp) + theta;
            const double q = ((gamma - dp) + gamma) + dy;
            const double r = p / q;
            stpf = stp + r * (sty - stp);
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp, fy = fp, dy = dp;
    } else {
        if (sgnd < 0.0) sty = stx, fy = fx, dy = dx;
        stx = stp, fx = fp, dx = dp;
    }
    stp = stpf;
}

// More-Thuente (dcsrch) along p from x; false when p is not a descent direction
bool line_search(const Mat4& m, const double* x, const double* p, double f0, double g0, double& f_out, double* xn, double* gn) {
    if (!(g0 < 0.0)) return false;
    const double ftol = 1e-4, gtol = 0.9, xtol = 1e-10, stpmin = kSqrtEps, stpmax = kInf;
    const double gtest = ftol * g0;
    double width = stpmax - stpmin, width1 = 2.0 * width;
    Interval v{0.0, f0, g0, 0.0, f0, g0, 1.0, false};
    double stmin = 0.0, stmax = v.stp + 4.0 * v.stp;
    int stage = 1;
    for (int fev = 0; fev < 100; fev++) {
        const double stp = v.stp;
        for (int i = 0; i < N; i++) xn[i] = x[i] + stp * p[i];
        const double f = cost_grad(m, xn, gn);
        const double g = dot(gn, p);
        const double ftest = f0 + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        if ((v.brackt && (stp <= stmin || stp >= stmax)) || (v.brackt && stmax - stmin <= xtol * stmax) ||
            (stp == stpmax && f <= ftest && g <= gtest) || (stp == stpmin && (f > ftest || g >= gtest)) ||
            (f <= ftest && std::fabs(g) <= gtol * (-g0))) {
            f_out = f;
            return true;
        }
        if (stage == 1 && f <= v.fx && f > ftest) {
            const double fm = f - stp * gtest, gm = g - gtest;
            Interval w{v.stx, v.fx - v.stx * gtest, v.dx - gtest, v.sty, v.fy - v.sty * gtest, v.dy - gtest, stp, v.brackt};
            dcstep(w, fm, gm, stmin, stmax);
            v = Interval{w.stx, w.fx + w.stx * gtest, w.dx + gtest, w.sty, w.fy + w.sty * gtest, w.dy + gtest, w.stp, w.brackt};
        } else {
            dcstep(v, f, g, stmin, stmax);
        }
        if (v.brackt) {
            if (std::fabs(v.sty - v.stx) >= 0.66 * width1) v.stp = v.stx + 0.5 * (v.sty - v.stx);
            width1 = width;
            width = std::fabs(v.sty - v.stx);
            stmin = min2(v.stx, v.sty);
            stmax = max2(v.stx, v.sty);
        } else {
            stmin = v.stp + 1.1 * (v.stp - v.stx);
            stmax = v.stp + 4.0 * (v.stp - v.stx);
        }
        v.stp = max2(v.stp, stpmin);
        v.stp = min2(v.stp, stpmax);
        if ((v.brackt && (v.stp <= stmin || v.stp >= stmax)) || (v.brackt && stmax - stmin <= xtol * stmax)) v.stp = v.stx;
    }
    for (int i = 0; i < N; i++) xn[i] = x[i] + v.stp * p[i];
    f_out = cost_grad(m, xn, gn);
    return true;
}

}  // namespace

bool mat_sqrt(const Mat4& m, Mat4& out) {
    double x[N], g[N], p[N], xn[N], gn[N], s[N], y[N], Hy[N];
    static thread_local double H[N][N];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) x[4 * r + c] = m[4 * c + r];  // mat_to_vec2, matrix.rs:936-943
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) H[i][j] = i == j ? 1.0 : 0.0;
    double f = cost_grad(m, x, g);
    for (int it = 0; it < 60; it++) {  // max_iters(60), matrix.rs:972
        if (std::sqrt(dot(g, g)) < kSqrtEps) break;
        for (int i = 0; i < N; i++) p[i] = -dot(H[i], g);
        double fn;
        if (!line_search(m, x, p, f, dot(g, p), fn, xn, gn)) break;
        for (int i = 0; i < N; i++) s[i] = xn[i] - x[i], y[i] = gn[i] - g[i];
        const double rho = 1.0 / dot(y, s);
        for (int i = 0; i < N; i++) Hy[i] = dot(H[i], y);
        const double k = rho * rho * dot(y, Hy) + rho;
        // (I - rho s y^T) H (I - rho y s^T) + rho s s^T, expanded for a symmetric H
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) H[i][j] = H[i][j] - rho * (s[i] * Hy[j] + Hy[i] * s[j]) + k * s[i] * s[j];
        const bool done = std::fabs(f - fn) < kEps;
        f = fn;
        for (int i = 0; i < N; i++) x[i] = xn[i], g[i] = gn[i];
        if (done) break;
    }
    if (!(f < 1e-4)) return false;  // matrix.rs:980-984
    for (int c = 0; c < 4; c++) {
        out[4 * c + 0] = x[c];
        out[4 * c + 1] = x[4 + c];
        out[4 * c + 2] = x[8 + c];
        out[4 * c + 3] = c == 3 ? 1.0 : 0.0;
    }
    return true;
}

}  // namespace ph
