// ptb_math.cuh -- device-side scalar arithmetic of the RoadRunner path (fp64).
//
// Each function names the reference arithmetic it implements (paths relative to the PyTransit
// v2.8.1 tree).  These are new implementations written for the GPU: reciprocals are hoisted into
// per-vector constants, branches are arranged for warp coherence, and table indices are clamped
// where the reference reads one element out of bounds.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ptb {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
constexpr double kHalfPi = 1.57079632679489661923;


// ---------------------------------------------------------------------------------------------
// Branch-free fp64 primitives for the per-sample hot loop.  CUDA's sqrt / division / atan2 carry a
// slow-path call that ends the basic block, so two independent samples in one lane cannot be interleaved
// by the scheduler; these straight-line versions can (the loop is bound by dependent-issue latency, not by
// the fp64 pipe).  Accuracy: <= 1 ulp (sqrt, div), <= 2 ulp (atan2) on the ranges the model produces.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_sqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RSQ64H, 2^-22 relative
    double g = x * r, h = 0.5 * r;
    const double e = fma(-h, g, 0.5);
    g = fma(g, e, g);          // sqrt(x)      to ~2^-43
    h = fma(h, e, h);          // 1/(2 sqrt x) to ~2^-43
    const double d = fma(-g, g, x);
    g = fma(d, h, g);          // residual step: full precision
    // 0 and subnormal operands (flushed by the approximation) give 0; inf and NaN pass through.
    // (fp64 min/max are multi-instruction on this pipe: compares and selects only.)
    g = (x >= 2.2250738585072014e-308) ? g : 0.0;
    return (x < __longlong_as_double(0x7ff0000000000000LL)) ? g : x;
}
__device__ __forceinline__ float fast_sqrt(float x) { return sqrtf(x); }

// sqrt for the separation: the operand is a sum of squares (+0, positive, or NaN).  MUFU.RSQ64H + one coupled Newton
// step (2^-43 relative); +0 and subnormal operands are lifted to the smallest normal number by an integer max on the
// high word (the separation becomes 1.5e-154 instead of 0), NaN passes through.  A sum of squares that overflows to +inf
// (sky-plane coordinates beyond 1e154 stellar radii) gives NaN where the reference's sqrt gives +inf: not a model any
// caller evaluates, and the one place the straight-line version departs from IEEE sqrt.
__device__ __forceinline__ double sqrt_sep(double x) {
    const unsigned hi = max((unsigned)__double2hiint(x), 0x00100000u);
    x = __hiloint2double((int)hi, __double2loint(x));
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double g = x * r, h = 0.5 * r;
    return fma(g, fma(-h, g, 0.5), g);
}
__device__ __forceinline__ float sqrt_sep(float x) { return sqrtf(x); }

// n / d for finite, normal d (not 0)
__device__ __forceinline__ double fast_div(double n, double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));  // MUFU.RCP64H, ~2^-20 relative
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    const double q = n * r;
    return fma(fma(-q, d, n), r, q);
}

// atan2(y, x) for y >= 0 (the lens-area kite: y = 2 * kite area).  One division after the argument
// reduction atan(q) = pi/4 + atan((q-1)/(q+1)) for q > tan(pi/8); odd polynomial of degree 23 in the
// reduced argument (near-minimax fit, profiles/tools/fit_atan.py: 1.5e-16 relative).
#define PTB_ATAN_COEFFS(X)                                                                                                        \
    X(3.79652574538659332e-02) X(-5.03510245660155203e-02) X(5.84687829733087222e-02) X(-6.66295181362919070e-02)               \
    X(7.69204533090222520e-02) X(-9.09089680906402658e-02) X(1.11111107449196583e-01) X(-1.42857142792502445e-01)               \
    X(1.99999999999408928e-01) X(-3.33333333333331205e-01)
__device__ __forceinline__ double atan2_pos(double y, double x) {
    const double ax = fabs(x);
    const bool ygt = y > ax;
    const double mx = ygt ? y : ax, mn = ygt ? ax : y;
    const bool big = mn > 0.41421356237309503 * mx;
    const double num = big ? mn - mx : mn;
    const double den = big ? mn + mx : mx;
    const double t = (den > 0.0) ? fast_div(num, den) : 0.0;
    const double u = t * t;
    double q = -1.78053972054194459e-02;
#define PTB_STEP(c) q = fma(q, u, c);
    PTB_ATAN_COEFFS(PTB_STEP)
#undef PTB_STEP
    double r = fma(t * u, q, t);                 // atan(t)
    r = big ? r + 0.78539816339744830962 : r;    // atan(mn / mx)
    r = ygt ? kHalfPi - r : r;
    return (x < 0.0) ? kPi - r : r;
}
__device__ __forceinline__ float atan2_pos(float y, float x) { return atan2f(y, x); }

// The lens area needs two of them with the same y: atan2(y, xa) and atan2(y, xb).  Evaluated side by side: the two
// dependency chains interleave, and every coefficient (an fp64 immediate costs two moves) is materialised once
// (constant-memory coefficients were measured: the LDC latency in front of the polynomial costs more than the moves).
__device__ __forceinline__ void atan2_pos2(double y, double xa, double xb, double &ra, double &rb) {
    const double axa = fabs(xa), axb = fabs(xb);
    const bool ygta = y > axa, ygtb = y > axb;
    const double mxa = ygta ? y : axa, mna = ygta ? axa : y, mxb = ygtb ? y : axb, mnb = ygtb ? axb : y;
    const double thr = 0.41421356237309503;
    const bool biga = mna > thr * mxa, bigb = mnb > thr * mxb;
    const double numa = biga ? mna - mxa : mna, dena = biga ? mna + mxa : mxa;
    const double numb = bigb ? mnb - mxb : mnb, denb = bigb ? mnb + mxb : mxb;
    const double ta = (dena > 0.0) ? fast_div(numa, dena) : 0.0, tb = (denb > 0.0) ? fast_div(numb, denb) : 0.0;
    const double ua = ta * ta, ub = tb * tb;
#ifndef PTB_ATAN_ESTRIN
#define PTB_ATAN_ESTRIN 1
#endif
#if PTB_ATAN_ESTRIN
    // the degree-10 polynomial in u as two interleaved Horner chains in u^2 (even / odd coefficients): half the depth
    const double c0 = -1.78053972054194459e-02, c1 = 3.79652574538659332e-02, c2 = -5.03510245660155203e-02,
                 c3 = 5.84687829733087222e-02, c4 = -6.66295181362919070e-02, c5 = 7.69204533090222520e-02,
                 c6 = -9.09089680906402658e-02, c7 = 1.11111107449196583e-01, c8 = -1.42857142792502445e-01,
                 c9 = 1.99999999999408928e-01, c10 = -3.33333333333331205e-01;
    const double va = ua * ua, vb = ub * ub;
    // q = c0 u^10 + c1 u^9 + ... + c10 = E(v) + u O(v),  E = c0 v^5 + c2 v^4 + ... + c10,  O = c1 v^4 + c3 v^3 + ... + c9
    double ea = fma(c0, va, c2), eb = fma(c0, vb, c2), oa = fma(c1, va, c3), ob = fma(c1, vb, c3);
    ea = fma(ea, va, c4); eb = fma(eb, vb, c4); oa = fma(oa, va, c5); ob = fma(ob, vb, c5);
    ea = fma(ea, va, c6); eb = fma(eb, vb, c6); oa = fma(oa, va, c7); ob = fma(ob, vb, c7);
    ea = fma(ea, va, c8); eb = fma(eb, vb, c8); oa = fma(oa, va, c9); ob = fma(ob, vb, c9);
    ea = fma(ea, va, c10); eb = fma(eb, vb, c10);
    const double qa = fma(oa, ua, ea), qb = fma(ob, ub, eb);
#else
    double qa = -1.78053972054194459e-02, qb = qa;
#define PTB_STEP(c) { const double cc = c; qa = fma(qa, ua, cc); qb = fma(qb, ub, cc); }
    PTB_ATAN_COEFFS(PTB_STEP)
#undef PTB_STEP
#endif
    ra = fma(ta * ua, qa, ta);
    rb = fma(tb * ub, qb, tb);
    const double q4 = 0.78539816339744830962, h = kHalfPi, pi = kPi;
    ra = biga ? ra + q4 : ra;
    rb = bigb ? rb + q4 : rb;
    ra = ygta ? h - ra : ra;
    rb = ygtb ? h - rb : rb;
    ra = (xa < 0.0) ? pi - ra : ra;
    rb = (xb < 0.0) ? pi - rb : rb;
}
__device__ __forceinline__ void atan2_pos2(float y, float xa, float xb, float &ra, float &rb) {
    ra = atan2f(y, xa);
    rb = atan2f(y, xb);
}

// ---------------------------------------------------------------------------------------------
// Geometry (models/roadrunner/common.py)
// ---------------------------------------------------------------------------------------------

// circle_circle_intersection_area, acos form (common.py:36-49).  Table construction only.
// The acos arguments approach +-1 at tangency, where a 1-ulp change of the argument is amplified
// ~1e4-fold, so the reference's operation order is kept rounding for rounding (explicit
// round-to-nearest intrinsics: no FMA contraction).
__device__ __forceinline__ double ccia_acos(double r1, double r2, double b) {
    if (r1 < b - r2) return 0.0;
    if (r1 >= b + r2) return kPi * (r2 * r2);
    if (b - r2 <= -r1) return kPi * (r1 * r1);
    const double b2 = __dmul_rn(b, b), r12 = __dmul_rn(r1, r1), r22 = __dmul_rn(r2, r2);
    const double x2 = __ddiv_rn(__dsub_rn(__dadd_rn(b2, r22), r12), __dmul_rn(__dmul_rn(2.0, b), r2));
    const double x1 = __ddiv_rn(__dsub_rn(__dadd_rn(b2, r12), r22), __dmul_rn(__dmul_rn(2.0, b), r1));
    const double q = __dmul_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dadd_rn(-b, r2), r1), __dsub_rn(__dadd_rn(b, r2), r1)),
                                         __dadd_rn(__dsub_rn(b, r2), r1)),
                               __dadd_rn(__dadd_rn(b, r2), r1));
    return __dsub_rn(__dadd_rn(__dmul_rn(r22, acos(x2)), __dmul_rn(r12, acos(x1))), __dmul_rn(0.5, sqrt(q)));
}

// circle_circle_intersection_area_kite(1, k, z) (common.py:52-73, tsort :5-33): lens area of the
// unit star and a planet of radius k at separation z, and kappa0.  k2 = k*k is passed in.
// The Kahan-ordered product keeps the reference's parenthesisation.
template <typename T>
__device__ __forceinline__ void kite_area(T k, T k2, T z, T &area, T &kappa0) {
    const T one = T(1), two = T(2), half = T(0.5), pi = T(kPi);
    if (one + k <= z) {
        area = T(0);
        kappa0 = T(0);
    } else if (fabs(one - k) < z) {
        // descending sort of (1, k, z), compares and selects only (tsort, common.py:5-33)
        const bool kg = k > one;
        const T hi = kg ? k : one, lo = kg ? one : k;
        const bool c1 = z > hi, c2 = z > lo;
        const T x = c1 ? z : hi;
        const T y = c1 ? hi : (c2 ? z : lo);
        const T zz = c2 ? lo : z;
        const T akite = half * fast_sqrt((x + (y + zz)) * (zz - (x - y)) * (zz + (x - y)) * (x + (y - zz)));
        const T z2 = z * z;
        T k0, k1;
        atan2_pos2(two * akite, (k - one) * (k + one) + z2, (one - k) * (one + k) + z2, k0, k1);
        area = k1 + k2 * k0 - akite;
        kappa0 = k0;
    } else if (z <= one - k) {
        area = pi * k2;
        kappa0 = pi;
    } else if (z <= k - one) {
        area = pi;
        kappa0 = T(0);
    } else {
        area = T(nan(""));
        kappa0 = T(nan(""));
    }
}

// The lens-area branch of kite_area alone, for a separation that is known to be on the limb (|1 - k| < z < 1 + k):
// straight-line code, so that the evaluations of two samples in one lane interleave.  Same arithmetic as above.
template <typename T>
__device__ __forceinline__ T kite_area_limb(T k, T k2, T z) {
    const T one = T(1), two = T(2), half = T(0.5);
    const bool kg = k > one;
    const T hi = kg ? k : one, lo = kg ? one : k;
    const bool c1 = z > hi, c2 = z > lo;
    const T x = c1 ? z : hi;
    const T y = c1 ? hi : (c2 ? z : lo);
    const T zz = c2 ? lo : z;
    // (2^-43 is ample for an area that enters the flux with a weight of ~0.1; a product that rounding made negative
    //  gives NaN, as the reference's sqrt does)
    const T akite = half * sqrt_sep((x + (y + zz)) * (zz - (x - y)) * (zz + (x - y)) * (x + (y - zz)));
    const T z2 = z * z;
    T k0, k1;
    atan2_pos2(two * akite, (k - one) * (k + one) + z2, (one - k) * (one + k) + z2, k0, k1);
    return k1 + k2 * k0 - akite;
}

// interpolate_mean_limb_darkening_s (common.py:225-233) with inv_dg = 1/dg hoisted and the upper
// node clamped to ng-1 (the reference reads lda[ng] for g in (1-1e-7, 1]; that term multiplies a
// lens area < 1e-10).  `row` may point to shared or global memory.
template <typename T>
__device__ __forceinline__ T ldm_lerp(T g, T dg, T inv_dg, const T *row, int ng) {
    if (g < T(0)) return T(nan(""));
    if (g > T(1)) return T(0);
    int i = (int)floor(g * inv_dg);
    const T a = (g - i * dg) * inv_dg;
    const int i0 = min(i, ng - 1);
    const int i1 = min(i + 1, ng - 1);
    return (T(1) - a) * row[i0] + a * row[i1];
}

// ---------------------------------------------------------------------------------------------
// Limb-darkening laws (models/numba/ldmodels.py:22-139) and analytic disk integrals (:27-123)
// ---------------------------------------------------------------------------------------------
enum : int {
    LD_UNIFORM = 0, LD_LINEAR = 1, LD_QUADRATIC = 2, LD_QUADRATIC_TRI = 3, LD_NONLINEAR = 4, LD_GENERAL = 5,
    LD_SQUARE_ROOT = 6, LD_LOGARITHMIC = 7, LD_EXPONENTIAL = 8, LD_POWER_2 = 9, LD_POWER_2_PM = 10,
    LD_PROFILES = 100
};

__device__ __forceinline__ double ld_intensity(int law, double mu, const double *pv, int nldc) {
    const double om = 1.0 - mu;
    switch (law) {
    case LD_UNIFORM: return 1.0;
    case LD_LINEAR: return 1.0 - pv[0] * om;
    case LD_QUADRATIC: return 1.0 - pv[0] * om - pv[1] * (om * om);
    case LD_QUADRATIC_TRI: {
        const double a = sqrt(pv[0]), b = 2.0 * pv[1];
        const double u = a * b, v = a * (1.0 - b);
        return 1.0 - u * om - v * (om * om);
    }
    case LD_NONLINEAR:
        return 1.0 - pv[0] * (1.0 - sqrt(mu)) - pv[1] * om - pv[2] * (1.0 - pow(mu, 1.5)) - pv[3] * (1.0 - mu * mu);
    case LD_GENERAL: {
        double s = 0.0;
        for (int i = 0; i < nldc; ++i) s += pv[i] * (1.0 - pow(mu, (double)(i + 1)));
        return s;
    }
    case LD_SQUARE_ROOT: return 1.0 - pv[0] * om - pv[1] * (1.0 - sqrt(mu));
    case LD_LOGARITHMIC: return 1.0 - pv[0] * om - pv[1] * mu * log(mu);
    case LD_EXPONENTIAL: return 1.0 - pv[0] * om - pv[1] / (1.0 - exp(mu));
    case LD_POWER_2: return 1.0 - pv[0] * (1.0 - pow(mu, pv[1]));
    case LD_POWER_2_PM: {
        const double c = 1.0 - pv[0] + pv[1];
        const double al = log2(c / pv[1]);
        return 1.0 - c * (1.0 - pow(mu, al));
    }
    default: return nan("");
    }
}

// true when the reference registry has an analytic integral for the law (rrmodel.py:48-58);
// 'linear' keeps the reference's 2 pi / 6 (3 - 2u) as coded (ldmodels.py:37-39).
__device__ __forceinline__ bool ld_integral(int law, const double *pv, double &istar) {
    switch (law) {
    case LD_UNIFORM: istar = kPi; return true;
    case LD_LINEAR: istar = 2.0 * kPi * 1.0 / 6.0 * (3.0 - 2.0 * pv[0]); return true;
    case LD_QUADRATIC: istar = 2.0 * kPi * 1.0 / 12.0 * (-2.0 * pv[0] - pv[1] + 6.0); return true;
    case LD_QUADRATIC_TRI: {
        const double a = sqrt(pv[0]), b = 2.0 * pv[1];
        const double u = a * b, v = a * (1.0 - b);
        istar = 2.0 * kPi * 1.0 / 12.0 * (-2.0 * u - v + 6.0);
        return true;
    }
    case LD_POWER_2: istar = 2.0 * kPi * (-pv[0] * pv[1] + pv[1] + 2.0) / (2.0 * pv[1] + 4.0); return true;
    default: return false;
    }
}

// ---------------------------------------------------------------------------------------------
// Orbit (meepmeep.backends.numba.point2d, restated from pytransit/orbits/taylor_z.py and
// orbits/orbits_py.py)
// ---------------------------------------------------------------------------------------------

// numpy.mod semantics (result has the sign of the divisor).
__device__ __forceinline__ double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

// mean_anomaly_offset (orbits_py.py:82-86)
__device__ __forceinline__ double mean_anomaly_offset(double e, double w) {
    double s, c;
    sincos(kHalfPi - w, &s, &c);
    double off = atan2(sqrt(1.0 - e * e) * s, e + c);
    off -= e * sin(off);
    return off;
}

// ta_newton_s (orbits_py.py:115-119,144-154,191-200) with t0 = 0: true anomaly at time t.
// `offset` = mean_anomaly_offset(e, w) is hoisted by the caller.
__device__ __forceinline__ double true_anomaly(double t, double p, double e, double offset) {
    const double Ma = pymod(kTwoPi * (t - (0.0 - offset * p / kTwoPi)) / p, kTwoPi);
    double Ea = Ma, err = 0.05, s, c;
    int it = 0;
    while (fabs(err) > 1e-8 && it < 1000) {
        sincos(Ea, &s, &c);
        err = Ea - e * s - Ma;
        Ea = Ea - err / (1.0 - e * c);
        ++it;
    }
    sincos(Ea, &s, &c);
    const double den = 1.0 - e * c;
    return atan2(sqrt(1.0 - e * e) * s / den, (c - e) / den);
}

// sky-plane position at time t (taylor_z.py:60-75)
__device__ __forceinline__ void sky_position(double t, double p, double ae, double ci, double e, double w,
                                             double offset, double &x, double &y) {
    const double f = true_anomaly(t, p, e, offset);
    const double r = ae / (1.0 + e * cos(f));
    double s, c;
    sincos(w + f, &s, &c);
    x = -r * c;
    y = -r * s * ci;
}

// 7-point central differences (taylor_z.py:77-100) -> monomial coefficients incl. 1/n!
// (layout evidenced by models/numba/gdmodel.py:441-442).
__device__ __forceinline__ void stencil_to_coeffs(const double *v, double *o) {
    const double dt = 2e-2;
    o[0] = v[3];
    o[1] = (1. / 60 * (v[6] - v[0]) + 9. / 60 * (v[1] - v[5]) + 45. / 60 * (v[4] - v[2])) / dt;
    o[2] = 0.5 * (1. / 90 * (v[0] + v[6]) - 3. / 20 * (v[1] + v[5]) + 3. / 2 * (v[2] + v[4]) - 49. / 18 * v[3]) /
           (dt * dt);
    o[3] = (1. / 8 * (v[0] - v[6]) + (v[5] - v[1]) + 13. / 8 * (v[2] - v[4])) / (dt * dt * dt) / 6.0;
    o[4] = (-1. / 6 * (v[0] + v[6]) + 2 * (v[1] + v[5]) - 13. / 2 * (v[2] + v[4]) + 28. / 3 * v[3]) /
           (dt * dt * dt * dt) / 24.0;
}

// sep_c (taylor_z.py:229-255): projected separation from the two quartics, Horner form.
template <typename T>
__device__ __forceinline__ T sep_poly(T t, const T *cx, const T *cy) {
    const T px = fma(t, fma(t, fma(t, fma(t, cx[4], cx[3]), cx[2]), cx[1]), cx[0]);
    const T py = fma(t, fma(t, fma(t, fma(t, cy[4], cy[3]), cy[2]), cy[1]), cy[0]);
    return fast_sqrt(fma(px, px, py * py));
}


__device__ __forceinline__ int floor_to_int(double x) { return __double2int_rd(x); }   // saturating, NaN -> 0
__device__ __forceinline__ int floor_to_int(float x) { return __float2int_rd(x); }

// LD-mean lerp (common.py:225-233) at grid position x = g/dg >= 0: node i = floor(x), weight x - i, upper node clamped to
// the last one (the reference reads one element past the row for g in (1-1e-7, 1]).  fp64: no conversion instructions
// (F2I / I2F run on the quarter-rate pipe) -- adding 1.5 * 2^52 rounds x - 0.5 to the nearest integer, which is
// floor(x) except when x is an integer (then it may be x - 1 with weight 1: the same value); the node index is the low
// word of the sum.  A position that is NaN or beyond the row gives a clamped node and a meaningless (NaN for NaN) weight:
// such samples are outside the stellar disk and their value is not used.
#ifndef SS_MAGIC_FLOOR
#define SS_MAGIC_FLOOR 1
#endif
__device__ __forceinline__ double ld_lerp(double x, const double *row, int ng) {
#if SS_MAGIC_FLOOR
    const double m = (x - 0.5) + 6755399441055744.0;            // 1.5 * 2^52: the sum stays in [2^52, 2^53), ulp 1
    const int i0 = (int)min((unsigned)__double2loint(m), (unsigned)(ng - 2));
    const double a = x - (m - 6755399441055744.0);
#else
    const int i0 = min(floor_to_int(x), ng - 2);
    const double a = x - (double)i0;
#endif
    const double r0 = row[i0], r1 = row[i0 + 1];
    return fma(a, r1 - r0, r0);
}
__device__ __forceinline__ float ld_lerp(float x, const float *row, int ng) {
    const int i0 = min(floor_to_int(x), ng - 2);     // >= 0; NaN -> 0 (the weight stays NaN)
    const float r0 = row[i0], r1 = row[i0 + 1];
    return fmaf(x - (float)i0, r1 - r0, r0);
}

// find_contact_point for points 1 (s=-1) and 4 (s=+1), target z = 1 + k (taylor_z.py:298-328).
__device__ __forceinline__ double contact_point(double k, double s, const double *cx, const double *cy) {
    const double zt = 1.0 + k;
    double t0 = 0.0;
    double t2 = s * 2.0 / cx[1];
    double t1 = 0.5 * t2;
    double z0 = sep_poly(t0, cx, cy) - zt;
    double z1 = sep_poly(t1, cx, cy) - zt;
    int i = 0;
    while (fabs(t2 - t0) > 1e-6 && i < 100) {
        if (z0 * z1 < 0.0) {
            t2 = t1;
            t1 = 0.5 * (t0 + t1);
        } else {
            t0 = t1;
            t1 = 0.5 * (t1 + t2);
            z0 = z1;
        }
        z1 = sep_poly(t1, cx, cy) - zt;
        ++i;
    }
    return t1;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (sm_90+/sm_100a PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy through the TMA engine; completion is signalled on `bar`.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ptb
