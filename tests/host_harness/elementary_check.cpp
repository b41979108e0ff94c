// The pinned elementary functions of the numeric profile: the DEVICE definitions
// (portal_b200/csrc/device/pe_glsl.cuh, namespace pe) against the ORACLE's (oracle/glsl_compat.h, namespace pe_oracle)
// bit for bit, and both against double-precision libm for accuracy.  Built and run by tests/test_elementary_host.py.
#include <cstdio>
#include <cstdlib>

#include "cuda_on_host.h"
#include "pe_glsl.cuh"
#include "glsl_compat.h"

static uint64_t st = 0x243F6A8885A308D3ull;
static inline uint64_t rnd() { st ^= st >> 12; st ^= st << 25; st ^= st >> 27; return st * 2685821657736338717ull; }
static inline float unit() { return float(rnd() >> 40) * (1.0f / 16777216.0f); }
static bool same(float a, float b) { return __float_as_int(a) == __float_as_int(b) || (a != a && b != b); }

struct Stat { double max_rel = 0.0; long long n = 0, mism = 0; };
static void acc(Stat& s, float got_dev, float got_orc, double want) {
    s.n++;
    if (!same(got_dev, got_orc)) s.mism++;
    if (std::isfinite(want) && want != 0.0 && std::isfinite(got_dev) && std::fabs(want) > 1e-37) {
        const double rel = std::fabs((double(got_dev) - want) / want);
        if (rel > s.max_rel) s.max_rel = rel;
    }
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? std::atoll(argv[1]) : 2000000;
    Stat e2, l2, ex, lg, pw, sn, cs, at;
    const float specials[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, NAN, 1e-45f, 1e-40f, 1.17549435e-38f, 3.4e38f, 127.999f, 128.0f,
                              -126.0f, -127.5f, -149.0f, -150.0f, -151.0f, 0.5f, 2.0f, 0.70710678f, 0.70710683f};
    for (long long i = 0; i < n; i++) {
        float x, y;
        if (i < (long long)(sizeof specials / sizeof specials[0])) { x = specials[i]; y = specials[(i * 7) % 22]; }
        else {
            const unsigned k = unsigned(rnd() % 8);
            x = k == 0 ? __int_as_float(int(rnd())) : (unit() * 2.0f - 1.0f) * (k < 4 ? 4.0f : (k < 7 ? 140.0f : 1e-3f));
            y = (unit() * 2.0f - 1.0f) * 8.0f;
        }
        acc(e2, pe::exp2(x), pe_oracle::exp2(x), std::exp2(double(x)));
        acc(ex, pe::exp(x * 0.5f), pe_oracle::exp(x * 0.5f), std::exp(double(x * 0.5f)));
        const float ax = std::fabs(x);
        acc(l2, pe::log2(ax), pe_oracle::log2(ax), std::log2(double(ax)));
        acc(lg, pe::log(ax), pe_oracle::log(ax), std::log(double(ax)));
        const float base = std::fabs(x) * 0.05f + 0.01f;
        acc(pw, pe::pow(base, y), pe_oracle::pow(base, y), std::pow(double(base), double(y)));
        const float iy = std::rint(y);                                           // negative base, integral exponent; pow(x, 0)
        acc(pw, pe::pow(-base, iy), pe_oracle::pow(-base, iy), std::pow(double(-base), double(iy)));
        acc(pw, pe::pow(x, 0.0f), pe_oracle::pow(x, 0.0f), 1.0);
        if (!same(pe::pow(-base, y), pe_oracle::pow(-base, y))) pw.mism++;       // non-integral exponent: NaN on both sides
        acc(sn, pe::sin(x), pe_oracle::sin(x), std::fabs(x) < 50.0f ? std::sin(double(x)) : NAN);
        acc(cs, pe::cos(x), pe_oracle::cos(x), std::fabs(x) < 50.0f ? std::cos(double(x)) : NAN);
        acc(at, pe::atan(x, y), pe_oracle::atan(x, y), NAN);
    }
    std::printf("cases %lld mismatches %lld exp2_rel %.3g log2_rel %.3g exp_rel %.3g log_rel %.3g pow_rel %.3g\n", e2.n,
                e2.mism + l2.mism + ex.mism + lg.mism + pw.mism + sn.mism + cs.mism + at.mism, e2.max_rel, l2.max_rel, ex.max_rel, lg.max_rel, pw.max_rel);
    return 0;
}
