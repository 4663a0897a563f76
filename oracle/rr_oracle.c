/*
 * oracle/rr_oracle.c -- CPU restatement of PyTransit's RoadRunner / TSModel / white-noise lnL
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker* for the CUDA product path in pytransit_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.  The
 * product (pytransit_b200) never imports, links or calls anything in oracle/.
 *
 * Every function cites the reference file:line (relative to the PyTransit v2.8.1 tree) whose
 * arithmetic it restates, in the reference's evaluation order.  It is a restatement in C of the
 * algorithm, not a copy of the Python sources.
 *
 * Parity pinning: validated here against the reference's own Numba files executed side by side
 * (tests/golden/make_golden.py -> tests/golden/ fixtures) and against the reference's known-answer
 * tests (tests/test_uniform_model_nb.py:24-25, tests/test_limb_darkening.py).  The three orbit
 * functions solve2d / sep_c / bounding_box live in the third-party package meepmeep>=1.0.0, which
 * is NOT in the reference tree and not installable here: they are restated from the in-tree
 * ancestor pytransit/orbits/taylor_z.py + orbits/orbits_py.py with the (2,5) monomial layout
 * evidenced by pytransit/models/numba/gdmodel.py:441-442.  For those three: PARITY UNPINNED.
 * The eclipse sibling (model_eclipse.py) calls two more absent meepmeep functions: eclipse_time_offset
 * (restated from the in-tree eclipse_phase, orbits/orbits_py.py:544-555) and eclipse_light_travel_time
 * (no in-tree ancestor; restated from its physical definition): PARITY UNPINNED for both as well.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared; no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846
#define ORC_TWO_PI (2.0 * ORC_PI)
#define ORC_HALF_PI (0.5 * ORC_PI)

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* Threads                                                                                    */
/* ------------------------------------------------------------------------------------------ */
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Geometry: pytransit/models/roadrunner/common.py                                            */
/* ------------------------------------------------------------------------------------------ */

/* common.py:5-33 (tsort): the three numbers in descending order. */
static void orc_tsort(double r1, double r2, double b, double *x, double *y, double *z) {
    if (r1 > r2) {
        if (r1 > b) {
            *x = r1;
            if (r2 > b) { *y = r2; *z = b; } else { *y = b; *z = r2; }
        } else { *x = b; *y = r1; *z = r2; }
    } else {
        if (r2 > b) {
            *x = r2;
            if (r1 > b) { *y = r1; *z = b; } else { *y = b; *z = r1; }
        } else { *x = b; *y = r2; *z = r1; }
    }
}

/* common.py:36-49 (circle_circle_intersection_area): acos form, used for the weight tables. */
double orc_ccia(double r1, double r2, double b) {
    if (r1 < b - r2) return 0.0;
    else if (r1 >= b + r2) return ORC_PI * (r2 * r2);
    else if (b - r2 <= -r1) return ORC_PI * (r1 * r1);
    else
        return (r2 * r2) * acos((b * b + r2 * r2 - r1 * r1) / (2 * b * r2)) +
               (r1 * r1) * acos((b * b + r1 * r1 - r2 * r2) / (2 * b * r1)) -
               0.5 * sqrt((-b + r2 + r1) * (b + r2 - r1) * (b - r2 + r1) * (b + r2 + r1));
}

/* common.py:52-73 (circle_circle_intersection_area_kite): Agol+2020 kite form; returns the lens
 * area and kappa0. */
void orc_ccia_kite(double r1, double r2, double b, double *area, double *kappa) {
    if (r1 + r2 <= b) { *area = 0.0; *kappa = 0.0; }
    else if (fabs(r1 - r2) < b && b <= r1 + r2) {
        double x, y, z;
        orc_tsort(r1, r2, b, &x, &y, &z);
        double a_kite = 0.5 * sqrt((x + (y + z)) * (z - (x - y)) * (z + (x - y)) * (x + (y - z)));
        double k0 = atan2(2.0 * a_kite, (r2 - r1) * (r2 + r1) + b * b);
        double k1 = atan2(2.0 * a_kite, (r1 - r2) * (r1 + r2) + b * b);
        *area = r1 * r1 * k1 + r2 * r2 * k0 - a_kite;
        *kappa = k0;
    }
    else if (b <= r1 - r2) { *area = ORC_PI * (r2 * r2); *kappa = ORC_PI; }
    else if (b <= r2 - r1) { *area = ORC_PI * (r1 * r1); *kappa = 0.0; }
    else { *area = NAN; *kappa = NAN; }
}

/* common.py:131-149 (create_z_grid). ze, zm have nin + nedge entries. */
void orc_create_z_grid(double zcut, int nin, int nedge, double *ze, double *zm) {
    int n = nin + nedge;
    double mucut = sqrt(1.0 - zcut * zcut);
    double dz = zcut / nin;
    double dmu = mucut / nedge;
    for (int i = 0; i < n; ++i) { ze[i] = 0.0; zm[i] = 0.0; }
    for (int i = 0; i < nin - 1; ++i) ze[i] = (i + 1) * dz;
    for (int i = 0; i < nedge + 1; ++i) {
        double v = i * dmu;
        ze[n - 1 - i] = sqrt(1 - v * v);
    }
    for (int i = 0; i < n - 1; ++i) zm[i + 1] = 0.5 * (ze[i] + ze[i + 1]);
}

/* Numba's np.linspace (numba/np/arrayobj.py numpy_linspace): start + i*step, last = stop. */
static void orc_linspace(double start, double stop, int num, double *out) {
    int div = num - 1;
    if (div > 0) {
        double step = (stop - start) / div;
        for (int i = 0; i < num; ++i) out[i] = start + (i * step);
    } else if (num > 0) out[0] = start;
    if (num > 1) out[num - 1] = stop;
}

/* common.py:152-185 (calculate_weights_2d). weights[ng,nz]; returns dg = gs[1]-gs[0]. */
double orc_weights_2d(double k, const double *ze, int nz, int ng, double *weights) {
    double *gs = (double *)malloc(sizeof(double) * ng);
    orc_linspace(0.0, 1.0 - 1e-7, ng, gs);
    for (int ig = 0; ig < ng; ++ig) {
        double *w = weights + (size_t)ig * nz;
        double b = gs[ig] * (1.0 + k);
        double a0 = orc_ccia(ze[0], k, b);
        w[0] = a0;
        double s = w[0];
        for (int i = 1; i < nz; ++i) {
            double a1 = orc_ccia(ze[i], k, b);
            w[i] = a1 - a0;
            a0 = a1;
            s += w[i];
        }
        for (int i = 0; i < nz; ++i) w[i] /= s;
    }
    double dg = gs[1] - gs[0];
    free(gs);
    return dg;
}

/* common.py:188-223 (calculate_weights_3d). weights[nk,ng,nz]. Writes dk = (k1-k0)/nk (NOT the
 * linspace step: SURVEY Q1) and dg. */
void orc_weights_3d(int nk, double k0, double k1, const double *ze, int nz, int ng,
                    double *weights, double *dk, double *dg) {
    double *ks = (double *)malloc(sizeof(double) * nk);
    orc_linspace(k0, k1, nk, ks);
    double dgl = 0.0;
#pragma omp parallel for schedule(static)
    for (int ik = 0; ik < nk; ++ik) {
        double d = orc_weights_2d(ks[ik], ze, nz, ng, weights + (size_t)ik * ng * nz);
        if (ik == 0) dgl = d;
    }
    *dk = (k1 - k0) / nk;
    *dg = dgl;
    free(ks);
}

/* common.py:225-233 (interpolate_mean_limb_darkening_s). The reference reads lda[i+1] one past
 * the row when g is in (1-1e-7, 1]; here the upper index is clamped to ng-1 and the
 * product path does the same.  The affected term is multiplied by a lens area < 1e-10. */
static double orc_interp_ldm(double g, double dg, const double *lda, int ng) {
    if (g < 0.0) return NAN;
    if (g > 1.0) return 0.0;
    int i = (int)floor(g / dg);
    double a = (g - i * dg) / dg;
    int i1 = i + 1;
    if (i > ng - 1) i = ng - 1;
    if (i1 > ng - 1) i1 = ng - 1;
    return (1.0 - a) * lda[i] + a * lda[i1];
}

/* ------------------------------------------------------------------------------------------ */
/* Limb darkening laws: pytransit/models/numba/ldmodels.py                                    */
/* ------------------------------------------------------------------------------------------ */
enum {
    ORC_LD_UNIFORM = 0, ORC_LD_LINEAR = 1, ORC_LD_QUADRATIC = 2, ORC_LD_QUADRATIC_TRI = 3,
    ORC_LD_NONLINEAR = 4, ORC_LD_GENERAL = 5, ORC_LD_SQUARE_ROOT = 6, ORC_LD_LOGARITHMIC = 7,
    ORC_LD_EXPONENTIAL = 8, ORC_LD_POWER_2 = 9, ORC_LD_POWER_2_PM = 10
};

/* ldmodels.py:22-139 (ld_*): I(mu) for one coefficient vector pv[nldc]. */
static double orc_ld_eval(int law, double mu, const double *pv, int nldc) {
    switch (law) {
    case ORC_LD_UNIFORM: return 1.0;                                              /* :22-24 */
    case ORC_LD_LINEAR: return 1. - pv[0] * (1. - mu);                            /* :32-34 */
    case ORC_LD_QUADRATIC:                                                        /* :50-52 */
        return 1. - pv[0] * (1. - mu) - pv[1] * ((1. - mu) * (1. - mu));
    case ORC_LD_QUADRATIC_TRI: {                                                  /* :74-78 */
        double a = sqrt(pv[0]), b = 2 * pv[1];
        double u = a * b, v = a * (1. - b);
        return 1. - u * (1. - mu) - v * ((1. - mu) * (1. - mu));
    }
    case ORC_LD_NONLINEAR:                                                        /* :88-90 */
        return 1. - pv[0] * (1. - sqrt(mu)) - pv[1] * (1. - mu) - pv[2] * (1. - pow(mu, 1.5)) -
               pv[3] * (1. - mu * mu);
    case ORC_LD_GENERAL: {                                                        /* :93-98 */
        double s = 0.0;
        for (int i = 0; i < nldc; ++i) s += pv[i] * (1.0 - pow(mu, (double)(i + 1)));
        return s;
    }
    case ORC_LD_SQUARE_ROOT:                                                      /* :101-103 */
        return 1. - pv[0] * (1. - mu) - pv[1] * (1. - sqrt(mu));
    case ORC_LD_LOGARITHMIC:                                                      /* :106-108 */
        return 1. - pv[0] * (1. - mu) - pv[1] * mu * log(mu);
    case ORC_LD_EXPONENTIAL:                                                      /* :111-113 */
        return 1. - pv[0] * (1. - mu) - pv[1] / (1. - exp(mu));
    case ORC_LD_POWER_2:                                                          /* :116-118 */
        return 1. - pv[0] * (1. - pow(mu, pv[1]));
    case ORC_LD_POWER_2_PM: {                                                     /* :135-139 */
        double c = 1 - pv[0] + pv[1];
        double a = log2(c / pv[1]);
        return 1. - c * (1. - pow(mu, a));
    }
    default: return NAN;
    }
}

/* ldmodels.py:27-123 (ldi_*): analytic disk integral; returns 0 and sets *ok=0 when the law has
 * no analytic integral in the reference registry (rrmodel.py:48-58). 'linear' reproduces the
 * reference's 2*pi/6*(3-2u) as coded. */
static double orc_ldi_eval(int law, const double *pv, int *ok) {
    *ok = 1;
    switch (law) {
    case ORC_LD_UNIFORM: return ORC_PI;                                            /* :27-29 */
    case ORC_LD_LINEAR: return 2 * ORC_PI * 1 / 6 * (3 - 2 * pv[0]);               /* :37-39 */
    case ORC_LD_QUADRATIC: return 2 * ORC_PI * 1 / 12 * (-2 * pv[0] - pv[1] + 6);  /* :55-57 */
    case ORC_LD_QUADRATIC_TRI: {                                                   /* :81-85 */
        double a = sqrt(pv[0]), b = 2 * pv[1];
        double u = a * b, v = a * (1. - b);
        return 2 * ORC_PI * 1 / 12 * (-2 * u - v + 6);
    }
    case ORC_LD_POWER_2:                                                           /* :121-123 */
        return 2 * ORC_PI * (-pv[0] * pv[1] + pv[1] + 2) / (2 * pv[1] + 4);
    default: *ok = 0; return 0.0;
    }
}

/* ldmodels.py:142-175 (evaluate_ld / evaluate_ldi) + the numeric fallback of
 * rrmodel.py:151-152,220-227: istar = 2 pi trapezoid(z * I(mu(z)), z), mu = linspace(1,0,200).
 * ldc[npv,npb,nldc] -> ldp[npv,npb,nmu], istar[npv,npb]. */
void orc_evaluate_ld(int law, const double *mu, int nmu, const double *ldc, int64_t npv, int64_t npb,
                     int nldc, double *ldp, double *istar) {
    double ldmu[200], ldz[200];
    /* numpy.linspace(1, 0, 200): start + i*step with step = (0-1)/199, last element = stop */
    {
        double step = (0.0 - 1.0) / 199.0;
        for (int i = 0; i < 200; ++i) ldmu[i] = i * step + 1.0;
        ldmu[199] = 0.0;
        for (int i = 0; i < 200; ++i) ldz[i] = sqrt(1 - ldmu[i] * ldmu[i]);
    }
    for (int64_t ipv = 0; ipv < npv; ++ipv)
        for (int64_t ipb = 0; ipb < npb; ++ipb) {
            const double *pv = ldc + (ipv * npb + ipb) * nldc;
            double *out = ldp + (ipv * npb + ipb) * nmu;
            for (int i = 0; i < nmu; ++i) out[i] = orc_ld_eval(law, mu[i], pv, nldc);
            int ok;
            double is = orc_ldi_eval(law, pv, &ok);
            if (!ok) {
                /* scipy.integrate.trapezoid(y, x) = sum (x[i+1]-x[i]) * (y[i+1]+y[i]) / 2 */
                double s = 0.0;
                double yp = ldz[0] * orc_ld_eval(law, ldmu[0], pv, nldc);
                for (int i = 1; i < 200; ++i) {
                    double yc = ldz[i] * orc_ld_eval(law, ldmu[i], pv, nldc);
                    s += (ldz[i] - ldz[i - 1]) * (yc + yp) / 2.0;
                    yp = yc;
                }
                is = 2 * ORC_PI * s;
            }
            istar[ipv * npb + ipb] = is;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Tabulated (LDTk-style) profiles: pytransit/models/numba/ldtkldm.py                         */
/* ------------------------------------------------------------------------------------------ */

/* ldtkldm.py:22-60 (trilinear_interpolation_unit_cube/_single/_set) and :77-91
 * (integrate_profiles_set). profiles[nx,ny,nz,npb,nmu] (C order). nx,ny,nz here are the grid
 * *sizes*; the reference clamps to x0 + n*dx (one cell past the last node; kept as coded). */
void orc_ldtk_profiles(const double *profiles, int nx, int ny, int nz, int64_t npb, int nmu,
                       const double *xs, const double *ys, const double *zs, int64_t npv,
                       double x0, double dx, double y0, double dy, double z0, double dz,
                       const double *mu, double *ldp, double *istar) {
    size_t s_mu = 1, s_pb = (size_t)nmu, s_z = (size_t)npb * nmu, s_y = s_z * nz, s_x = s_y * ny;
    double *zz = (double *)malloc(sizeof(double) * nmu);
    for (int i = 0; i < nmu; ++i) zz[i] = sqrt(1.0 - mu[i] * mu[i]);
    (void)s_mu;
#pragma omp parallel for schedule(static)
    for (int64_t ipv = 0; ipv < npv; ++ipv) {
        double x = fmin(fmax(xs[ipv], x0), x0 + nx * dx);
        double y = fmin(fmax(ys[ipv], y0), y0 + ny * dy);
        double z = fmin(fmax(zs[ipv], z0), z0 + nz * dz);
        int ix = (int)floor((x - x0) / dx);
        int iy = (int)floor((y - y0) / dy);
        int iz = (int)floor((z - z0) / dz);
        double ax = (x - x0 - ix * dx) / dx;
        double ay = (y - y0 - iy * dy) / dy;
        double az = (z - z0 - iz * dz) / dz;
        double rx = 1.0 - ax, ry = 1.0 - ay, rz = 1.0 - az;
        double a1 = rx * ry * rz, a2 = ax * ry * rz, b1 = rx * ay * rz, b2 = rx * ry * az;
        double c1 = ax * ry * az, c2 = rx * ay * az, d1 = ax * ay * rz, d2 = ax * ay * az;
        const double *d000 = profiles + ix * s_x + iy * s_y + iz * s_z;
        for (int64_t ipb = 0; ipb < npb; ++ipb) {
            double *out = ldp + (ipv * npb + ipb) * nmu;
            for (int i = 0; i < nmu; ++i) {
                size_t o = ipb * s_pb + i;
                out[i] = (d000[o] * a1 + d000[o + s_x] * a2 + d000[o + s_y] * b1 + d000[o + s_z] * b2 +
                          d000[o + s_x + s_z] * c1 + d000[o + s_y + s_z] * c2 +
                          d000[o + s_x + s_y] * d1 + d000[o + s_x + s_y + s_z] * d2);
            }
            double s = 0.0;
            for (int i = 1; i < nmu; ++i)
                s += (zz[i] - zz[i - 1]) * 0.5 * (zz[i] * out[i] + zz[i - 1] * out[i - 1]);
            istar[ipv * npb + ipb] = 2.0 * ORC_PI * s;
        }
    }
    free(zz);
}

/* ------------------------------------------------------------------------------------------ */
/* Orbit: meepmeep.backends.numba.point2d (external; restated from the in-tree ancestor)       */
/* ------------------------------------------------------------------------------------------ */

/* Python-style float modulo (result takes the sign of the divisor), as numpy.mod. */
static double orc_pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

/* orbits/orbits_py.py:82-86 (mean_anomaly_offset). */
static double orc_mean_anomaly_offset(double e, double w) {
    double off = atan2(sqrt(1.0 - e * e) * sin(ORC_HALF_PI - w), e + cos(ORC_HALF_PI - w));
    off -= e * sin(off);
    return off;
}

/* orbits/orbits_py.py:115-119 (mean_anomaly), :144-154 (ea_newton_s), :191-200 (ta_from_ea_s,
 * ta_newton_s). */
static double orc_ta_newton(double t, double t0, double p, double e, double w) {
    double offset = orc_mean_anomaly_offset(e, w);
    double Ma = orc_pymod(ORC_TWO_PI * (t - (t0 - offset * p / ORC_TWO_PI)) / p, ORC_TWO_PI);
    double Ea = Ma;
    double err = 0.05;
    int k = 0;
    while (fabs(err) > 1e-8 && k < 1000) {
        err = Ea - e * sin(Ea) - Ma;
        Ea = Ea - err / (1.0 - e * cos(Ea));
        k += 1;
    }
    double sta = sqrt(1.0 - e * e) * sin(Ea) / (1.0 - e * cos(Ea));
    double cta = (cos(Ea) - e) / (1.0 - e * cos(Ea));
    return atan2(sta, cta);
}

/* solve2d(t, p, a, i, e, w) -> c[2][5]: 7-point central-difference stencil, dt = 0.02 d, of the
 * sky-plane position around t; orbits/taylor_z.py:23-102 (vajs_from_paiew) with the position
 * constant term kept and the 1/n! folded in (models/numba/gdmodel.py:441-442). */
void orc_solve2d(double t, double p, double a, double i, double e, double w, double *c) {
    const double dt = 2e-2;
    double ae = a * (1. - e * e);
    double ci = cos(i);
    double x[7], y[7];
    for (int j = 0; j < 7; ++j) {
        double f = orc_ta_newton(t + (j - 3) * dt, 0.0, p, e, w);
        double r = ae / (1. + e * cos(f));
        x[j] = -r * cos(w + f);
        y[j] = -r * sin(w + f) * ci;
    }
    for (int d = 0; d < 2; ++d) {
        const double *v = d == 0 ? x : y;
        double *o = c + d * 5;
        o[0] = v[3];
        o[1] = (1. / 60 * (v[6] - v[0]) + 9. / 60 * (v[1] - v[5]) + 45. / 60 * (v[4] - v[2])) / dt;
        o[2] = 0.5 * (1. / 90 * (v[0] + v[6]) - 3. / 20 * (v[1] + v[5]) + 3. / 2 * (v[2] + v[4]) -
                      49. / 18 * v[3]) / (dt * dt);
        o[3] = (1. / 8 * (v[0] - v[6]) + (v[5] - v[1]) + 13. / 8 * (v[2] - v[4])) / (dt * dt * dt) / 6.0;
        o[4] = (-1. / 6 * (v[0] + v[6]) + 2 * (v[1] + v[5]) - 13. / 2 * (v[2] + v[4]) + 28. / 3 * v[3]) /
               (dt * dt * dt * dt) / 24.0;
    }
}

/* sep_c(t, c): orbits/taylor_z.py:229-255 (z_taylor_st) in Horner form over c[2][5]. */
double orc_sep_c(double t, const double *c) {
    double px = c[0] + t * (c[1] + t * (c[2] + t * (c[3] + t * c[4])));
    double py = c[5] + t * (c[6] + t * (c[7] + t * (c[8] + t * c[9])));
    return sqrt(px * px + py * py);
}

/* find_contact_point: orbits/taylor_z.py:298-328 (array form models/numba/gdmodel.py:468-507). */
static double orc_find_contact_point(double k, int point, const double *c) {
    double s = (point == 1 || point == 2 || point == 12) ? -1.0 : 1.0;
    double zt = (point == 1 || point == 4) ? 1.0 + k : ((point == 2 || point == 3) ? 1.0 - k : 1.0);
    double t0 = 0.0;
    double t2 = s * 2.0 / c[1];
    double t1 = 0.5 * t2;
    double z0 = orc_sep_c(t0, c) - zt;
    double z1 = orc_sep_c(t1, c) - zt;
    int i = 0;
    while (fabs(t2 - t0) > 1e-6 && i < 100) {
        if (z0 * z1 < 0.0) {
            t2 = t1;
            t1 = 0.5 * (t0 + t1);
            z1 = orc_sep_c(t1, c) - zt;
        } else {
            t0 = t1;
            t1 = 0.5 * (t1 + t2);
            z0 = z1;
            z1 = orc_sep_c(t1, c) - zt;
        }
        i += 1;
    }
    return t1;
}

/* bounding_box(k, c): orbits/taylor_z.py:391-394.  The two contact times are returned in ascending
 * order: around a secondary eclipse the sky-plane x velocity c[0][1] is negative, which swaps the roles of
 * the "first" and "fourth" contact searches (for a transit the order is unchanged). */
void orc_bounding_box(double k, const double *c, double *t1, double *t4) {
    double a = orc_find_contact_point(k, 1, c);
    double b = orc_find_contact_point(k, 4, c);
    *t1 = a < b ? a : b;
    *t4 = a < b ? b : a;
}

/* eclipse_time_offset (meepmeep.backends.numba.utils; absent) restated from its in-tree ancestor
 * eclipse_phase, orbits/orbits_py.py:544-555: time from mid-transit to mid-eclipse. */
double orc_eclipse_time_offset(double p, double i, double e, double w) {
    (void)i;
    double etr = atan2(sqrt(1. - e * e) * sin(ORC_HALF_PI - w), e + cos(ORC_HALF_PI - w));
    double eec = atan2(sqrt(1. - e * e) * sin(ORC_HALF_PI + ORC_PI - w), e + cos(ORC_HALF_PI + ORC_PI - w));
    double mtr = etr - e * sin(etr);
    double mec = eec - e * sin(eec);
    double phase = (mec - mtr) * p / ORC_TWO_PI;
    return phase > 0. ? phase : p + phase;
}

/* eclipse_light_travel_time (meepmeep.backends.numba.newton; absent, no in-tree ancestor -- PARITY
 * UNPINNED): light travel time across the line-of-sight distance between the planet's positions at
 * mid-transit and mid-eclipse, (r_tr + r_ec) sin(i) stellar radii, in days.  R_sun as orbits_py.py:46. */
double orc_eclipse_light_travel_time(double p, double a, double i, double e, double w, double rstar) {
    (void)p;
    const double rsun = 0.5 * 1.392684e9, c_light = 299792458.0, d_s = 86400.0;
    double ae = a * (1. - e * e);
    double r_tr = ae / (1. + e * sin(w)), r_ec = ae / (1. - e * sin(w));
    return (r_tr + r_ec) * sin(i) * rstar * rsun / c_light / d_s;
}

/* eclipse_model: pytransit/models/roadrunner/model_eclipse.py:11-81.  k[npv]; t0[npv,nep]; flux[npv,npt] =
 * pi k^2 minus the occulted planet area, averaged over the exposure sub-samples. */
void orc_eclipse_model(const double *times, int64_t npt, const double *k, const double *t0, const double *p,
                       const double *a, const double *inc, const double *e, const double *w, double rstar,
                       int64_t npv, int64_t nlc, int64_t nep, const int64_t *lcids, const int64_t *epids,
                       const int64_t *nsamples, const double *exptimes, double *flux) {
    (void)nlc;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ipv = 0; ipv < npv; ++ipv) {
        double *f = flux + ipv * npt;
        if (isnan(a[ipv]) || a[ipv] <= 1.0 || e[ipv] < 0.0) {
            for (int64_t j = 0; j < npt; ++j) f[j] = NAN;
            continue;
        }
        double xyc[10], bt1, bt4;
        double shift = orc_eclipse_time_offset(p[ipv], inc[ipv], e[ipv], w[ipv]);
        orc_solve2d(shift, p[ipv], a[ipv], inc[ipv], e[ipv], w[ipv], xyc);
        double ltt = orc_eclipse_light_travel_time(p[ipv], a[ipv], inc[ipv], e[ipv], w[ipv], rstar);
        orc_bounding_box(k[ipv], xyc, &bt1, &bt4);
        const double pk2 = ORC_PI * k[ipv] * k[ipv];
        for (int64_t ipt = 0; ipt < npt; ++ipt) {
            int64_t ilc = lcids[ipt], iep = epids[ilc];
            double lo = bt1 - (0.003 + exptimes[ilc]), hi = bt4 + (0.003 + exptimes[ilc]);
            double te = t0[ipv * nep + iep] + shift + ltt;
            double epoch = floor((times[ipt] - te + 0.5 * p[ipv]) / p[ipv]);
            double tc = times[ipt] - (te + epoch * p[ipv]);
            if (!(lo <= tc && tc <= hi)) {
                f[ipt] = pk2;
            } else {
                double acc = 0.0;
                for (int64_t s = 1; s <= nsamples[ilc]; ++s) {
                    double off = exptimes[ilc] * ((s - 0.5) / nsamples[ilc] - 0.5);
                    double z = orc_sep_c(tc + off, xyc), area, kap;
                    orc_ccia_kite(1.0, k[ipv], z, &area, &kap);
                    acc += pk2 - area;
                }
                f[ipt] = acc / nsamples[ilc];
            }
        }
    }
}


/* esmodel: pytransit/models/roadrunner/model_ecspec.py:13-63 (eclipse spectroscopy).  k, t0, p, a, inc, e, w,
 * rstar[npv]; fratio[npv,npb]; flux[npv,npb,npt] = 1 - (f A / pi) / (1 + f k^2), averaged over the sub-samples. */
void orc_esmodel(const double *times, int64_t npt, const double *k, const double *t0, const double *p, const double *a,
                 const double *inc, const double *e, const double *w, const double *rstar, const double *fratio,
                 int64_t npv, int64_t npb, int64_t nsamples, double exptime, double *flux) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ipv = 0; ipv < npv; ++ipv) {
        double *f = flux + ipv * npb * npt;
        if (isnan(a[ipv]) || a[ipv] <= 1.0 || e[ipv] < 0.0) {
            for (int64_t j = 0; j < npb * npt; ++j) f[j] = NAN;
            continue;
        }
        double xyc[10], bt1, bt4;
        double shift = orc_eclipse_time_offset(p[ipv], inc[ipv], e[ipv], w[ipv]);
        orc_solve2d(shift, p[ipv], a[ipv], inc[ipv], e[ipv], w[ipv], xyc);
        double ltt = orc_eclipse_light_travel_time(p[ipv], a[ipv], inc[ipv], e[ipv], w[ipv], rstar[ipv]);
        double te = t0[ipv] + shift + ltt;
        orc_bounding_box(k[ipv], xyc, &bt1, &bt4);
        bt1 -= 0.0015 + exptime;
        bt4 += 0.0015 + exptime;
        for (int64_t ipt = 0; ipt < npt; ++ipt) {
            double epoch = floor((times[ipt] - te + 0.5 * p[ipv]) / p[ipv]);
            double tc = times[ipt] - (te + epoch * p[ipv]);
            if (!(bt1 <= tc && tc <= bt4)) {
                for (int64_t pb = 0; pb < npb; ++pb) f[pb * npt + ipt] = 1.0;
            } else {
                for (int64_t pb = 0; pb < npb; ++pb) f[pb * npt + ipt] = 0.0;
                for (int64_t s = 1; s <= nsamples; ++s) {
                    double off = exptime * ((s - 0.5) / nsamples - 0.5);
                    double z = orc_sep_c(tc + off, xyc), area, kap;
                    orc_ccia_kite(1.0, k[ipv], z, &area, &kap);
                    for (int64_t pb = 0; pb < npb; ++pb) {
                        double fr = fratio[ipv * npb + pb];
                        f[pb * npt + ipt] += 1.0 - (fr * area / ORC_PI) / (1.0 + fr * (k[ipv] * k[ipv]));
                    }
                }
                for (int64_t pb = 0; pb < npb; ++pb) f[pb * npt + ipt] /= nsamples;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* RoadRunner population model: pytransit/models/roadrunner/model_full.py:9-100 (rr_full)      */
/* ------------------------------------------------------------------------------------------ */

/* Stage outputs (any may be NULL): ldm_out[npv,npb,ng], xyc_out[npv,2,5], bbs_out[npv,nlc,2].
 * xyc_in (may be NULL): inject the Taylor coefficients instead of calling solve2d, so that the
 * downstream stages can be pinned independently of the unpinned orbit restatement.
 * k[npv,kcols] with kcols in {1,npb}; t0[npv,nep]; nsamples/exptimes have nlc entries
 * (the reference indexes the un-broadcast arrays, SURVEY Q16: callers pass full arrays).
 * Returns 0, or -1 for the reference's ValueError on the k shape (model_full.py:18-19). */
int orc_rr_full(const double *times, int64_t npt, const double *k, int64_t kcols, const double *t0,
                const double *p, const double *a, const double *inc, const double *e, const double *w,
                int64_t npv, int64_t nlc, int64_t npb, int64_t nep, const int64_t *lcids,
                const int64_t *pbids, const int64_t *epids, const int64_t *nsamples,
                const double *exptimes, const double *ldp, int nz, const double *istar,
                const double *weights, int nk, int ng, double dk, double kmin, double kmax, double dg,
                const double *ze, const double *xyc_in, double *flux, double *ldm_out, double *xyc_out,
                double *bbs_out) {
    if (kcols > 1 && kcols != npb) return -1;
    double *ks = (double *)malloc(sizeof(double) * npv * npb);
    for (int64_t ipv = 0; ipv < npv; ++ipv)
        for (int64_t ipb = 0; ipb < npb; ++ipb)
            ks[ipv * npb + ipb] = kcols == npb ? k[ipv * kcols + ipb] : k[ipv * kcols];

    char *good = (char *)malloc(npv);
    double *ldm = (double *)calloc((size_t)npv * npb * ng, sizeof(double));
    double *xyc = (double *)calloc((size_t)npv * 10, sizeof(double));
    double *bbs = (double *)calloc((size_t)npv * nlc * 2, sizeof(double));

    /* model_full.py:39-70 -- per parameter vector setup (serial in the reference). */
#pragma omp parallel
    {
        double *wg = (double *)malloc(sizeof(double) * ng * nz);
#pragma omp for schedule(dynamic, 16)
        for (int64_t ipv = 0; ipv < npv; ++ipv) {
            good[ipv] = 1;
            if (isnan(a[ipv]) || (a[ipv] <= 1.0) || (e[ipv] < 0.0) || isnan(ldp[ipv * npb * nz])) {
                good[ipv] = 0;
                continue;
            }
            double k0 = ks[ipv * npb];
            if (kmin <= k0 && k0 <= kmax) {                                   /* :47-51 */
                int ik = (int)floor((k0 - kmin) / dk);
                double ak = (k0 - kmin - ik * dk) / dk;
                int ik1 = ik + 1 < nk ? ik + 1 : nk - 1;  /* reference reads weights[nk] OOB (Q1) */
                const double *w0 = weights + (size_t)ik * ng * nz;
                const double *w1 = weights + (size_t)ik1 * ng * nz;
                for (int64_t ipb = 0; ipb < npb; ++ipb) {
                    const double *l = ldp + (ipv * npb + ipb) * nz;
                    double *o = ldm + (ipv * npb + ipb) * ng;
                    for (int ig = 0; ig < ng; ++ig) {
                        double d0 = 0.0, d1 = 0.0;
                        for (int iz = 0; iz < nz; ++iz) {
                            d0 += w0[ig * nz + iz] * l[iz];
                            d1 += w1[ig * nz + iz] * l[iz];
                        }
                        o[ig] = (1.0 - ak) * d0 + ak * d1;
                    }
                }
            } else {                                                          /* :52-55 */
                orc_weights_2d(k0, ze, nz, ng, wg);
                for (int64_t ipb = 0; ipb < npb; ++ipb) {
                    const double *l = ldp + (ipv * npb + ipb) * nz;
                    double *o = ldm + (ipv * npb + ipb) * ng;
                    for (int ig = 0; ig < ng; ++ig) {
                        double d0 = 0.0;
                        for (int iz = 0; iz < nz; ++iz) d0 += wg[ig * nz + iz] * l[iz];
                        o[ig] = d0;
                    }
                }
            }
            if (xyc_in) memcpy(xyc + ipv * 10, xyc_in + ipv * 10, sizeof(double) * 10);
            else orc_solve2d(0.0, p[ipv], a[ipv], inc[ipv], e[ipv], w[ipv], xyc + ipv * 10); /* :60 */
            double bt1, bt4;
            orc_bounding_box(k0, xyc + ipv * 10, &bt1, &bt4);                 /* :65-70 */
            for (int64_t ilc = 0; ilc < nlc; ++ilc) {
                bbs[(ipv * nlc + ilc) * 2 + 0] = bt1 - (0.003 + exptimes[ilc]);
                bbs[(ipv * nlc + ilc) * 2 + 1] = bt4 + (0.003 + exptimes[ilc]);
            }
        }
        free(wg);
    }

    /* model_full.py:75-99 -- the prange(npv*npt) loop. */
    int64_t ntot = npv * npt;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < ntot; ++j) {
        int64_t ipv = j / npt;
        int64_t ipt = j % npt;
        if (!good[ipv]) { flux[j] = NAN; continue; }
        int64_t ilc = lcids[ipt];
        int64_t ipb = pbids[ilc];
        int64_t iep = epids[ilc];
        double t0v = t0[ipv * nep + iep];
        double epoch = floor((times[ipt] - t0v + 0.5 * p[ipv]) / p[ipv]);
        double tc = times[ipt] - (t0v + epoch * p[ipv]);
        if (!(bbs[(ipv * nlc + ilc) * 2] <= tc && tc <= bbs[(ipv * nlc + ilc) * 2 + 1])) {
            flux[j] = 1.0;
        } else {
            double f = 0.0;
            double kk = ks[ipv * npb + ipb];
            double is = istar[ipv * npb + ipb];
            int64_t ns = nsamples[ilc];
            for (int64_t isample = 1; isample < ns + 1; ++isample) {
                double time_offset = exptimes[ilc] * ((isample - 0.5) / ns - 0.5);
                double z = orc_sep_c(tc + time_offset, xyc + ipv * 10);
                double iplanet = orc_interp_ldm(z / (1.0 + kk), dg, ldm + (ipv * npb + ipb) * ng, ng);
                double aplanet, kap;
                orc_ccia_kite(1.0, kk, z, &aplanet, &kap);
                f += (is - iplanet * aplanet) / is;
            }
            flux[j] = f / ns;
        }
    }

    if (ldm_out) memcpy(ldm_out, ldm, sizeof(double) * npv * npb * ng);
    if (xyc_out) memcpy(xyc_out, xyc, sizeof(double) * npv * 10);
    if (bbs_out) memcpy(bbs_out, bbs, sizeof(double) * npv * nlc * 2);
    free(ks); free(good); free(ldm); free(xyc); free(bbs);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Single-vector model: pytransit/models/roadrunner/model_simple.py:25-80 (rr_simple_serial)   */
/* ------------------------------------------------------------------------------------------ */
void orc_rr_simple(const double *times, int64_t npt, double k, double t0, double p, double a, double inc,
                   double e, double w, int64_t nsamples, double exptime, const double *ldp, int nz,
                   double istar, const double *weights, int nk, int ng, double dk, double kmin,
                   double kmax, double dg, const double *ze, double *flux) {
    if (isnan(a) || (a <= 1.0) || (e < 0.0) || isnan(ldp[0])) {                  /* :38-39 */
        for (int64_t i = 0; i < npt; ++i) flux[i] = NAN;
        return;
    }
    double *ldm = (double *)calloc(ng, sizeof(double));
    if (kmin <= k && k <= kmax) {                                               /* :44-47 */
        int ik = (int)floor((k - kmin) / dk);
        double ak = (k - kmin - ik * dk) / dk;
        int ik1 = ik + 1 < nk ? ik + 1 : nk - 1;
        for (int ig = 0; ig < ng; ++ig) {
            double d0 = 0.0, d1 = 0.0;
            for (int iz = 0; iz < nz; ++iz) {
                d0 += weights[((size_t)ik * ng + ig) * nz + iz] * ldp[iz];
                d1 += weights[((size_t)ik1 * ng + ig) * nz + iz] * ldp[iz];
            }
            ldm[ig] = (1.0 - ak) * d0 + ak * d1;
        }
    } else {                                                                    /* :48-50 */
        double *wg = (double *)malloc(sizeof(double) * ng * nz);
        orc_weights_2d(k, ze, nz, ng, wg);
        for (int ig = 0; ig < ng; ++ig) {
            double d0 = 0.0;
            for (int iz = 0; iz < nz; ++iz) d0 += wg[ig * nz + iz] * ldp[iz];
            ldm[ig] = d0;
        }
        free(wg);
    }
    double xyc[10];
    orc_solve2d(0.0, p, a, inc, e, w, xyc);                                     /* :54 */
    double bt1, bt4;
    orc_bounding_box(k, xyc, &bt1, &bt4);                                       /* :59-61 */
    bt1 -= 0.003 + exptime;
    bt4 += 0.003 + exptime;
    for (int64_t ipt = 0; ipt < npt; ++ipt) {                                   /* :66-79 */
        double epoch = floor((times[ipt] - t0 + 0.5 * p) / p);
        double tc = times[ipt] - (t0 + epoch * p);
        if (!(bt1 <= tc && tc <= bt4)) flux[ipt] = 1.0;
        else {
            double f = 0.0;
            for (int64_t isample = 1; isample < nsamples + 1; ++isample) {
                double time_offset = exptime * ((isample - 0.5) / nsamples - 0.5);
                double z = orc_sep_c(tc + time_offset, xyc);
                double iplanet = orc_interp_ldm(z / (1.0 + k), dg, ldm, ng);
                double aplanet, kap;
                orc_ccia_kite(1.0, k, z, &aplanet, &kap);
                f += (istar - iplanet * aplanet) / istar;
            }
            flux[ipt] = f / nsamples;
        }
    }
    free(ldm);
}

/* ------------------------------------------------------------------------------------------ */
/* Transmission spectroscopy: pytransit/models/roadrunner/model_trspec.py:11-93 (tsmodel_serial)*/
/* ------------------------------------------------------------------------------------------ */

/* weights == NULL means precompute_weights=False (tsmodel.py:117-120): direct 2-D weights at
 * kmean and *its* dg.  flux[npv,npb,npt].  The outer ipv loop is independent per vector and is
 * threaded here (the reference's serial kernel is single-threaded; its prange variant is broken,
 * SURVEY Q11). */
int orc_tsmodel(const double *times, int64_t npt, const double *k, const double *t0, const double *p,
                const double *a, const double *inc, const double *e, const double *w, int64_t npv,
                int64_t npb, int64_t nsamples, double exptime, const double *ldp, int nz,
                const double *istar, const double *weights, int nk, int ng, double dk, double kmin,
                double kmax_tab, double dg_tab, const double *ze, const double *xyc_in, double *flux) {
    (void)kmax_tab; /* shadowed by max(k[ipv]) in the reference (model_trspec.py:42) */
#pragma omp parallel
    {
        double *ldm = (double *)malloc(sizeof(double) * npb * ng);
        double *wg = (double *)malloc(sizeof(double) * ng * nz);
        double *afac = (double *)malloc(sizeof(double) * npb);
#pragma omp for schedule(dynamic, 1)
        for (int64_t ipv = 0; ipv < npv; ++ipv) {
            double *fl = flux + (size_t)ipv * npb * npt;
            if (isnan(a[ipv]) || (a[ipv] <= 1.0) || (e[ipv] < 0.0)) {            /* :37-39 */
                for (int64_t q = 0; q < npb * npt; ++q) fl[q] = NAN;
                continue;
            }
            const double *kv = k + ipv * npb;
            double ksum = 0.0, kmax = kv[0];
            for (int64_t ipb = 0; ipb < npb; ++ipb) { ksum += kv[ipb]; if (kv[ipb] > kmax) kmax = kv[ipb]; }
            double kmean = ksum / npb;                                          /* :41-43 */
            for (int64_t ipb = 0; ipb < npb; ++ipb) afac[ipb] = (kv[ipb] * kv[ipb]) / (kmean * kmean);

            double dg = dg_tab;
            if (weights != NULL && kmin <= kmean && kmean <= kmax) {             /* :48-52 */
                int ik = (int)floor((kmean - kmin) / dk);
                double ak = (kmean - kmin - ik * dk) / dk;
                int ik1 = ik + 1 < nk ? ik + 1 : nk - 1;
                if (ik > nk - 1) ik = nk - 1;
                const double *w0 = weights + (size_t)ik * ng * nz;
                const double *w1 = weights + (size_t)ik1 * ng * nz;
                for (int64_t ipb = 0; ipb < npb; ++ipb) {
                    const double *l = ldp + (ipv * npb + ipb) * nz;
                    for (int ig = 0; ig < ng; ++ig) {
                        double d0 = 0.0, d1 = 0.0;
                        for (int iz = 0; iz < nz; ++iz) {
                            d0 += w0[ig * nz + iz] * l[iz];
                            d1 += w1[ig * nz + iz] * l[iz];
                        }
                        ldm[ipb * ng + ig] = (1.0 - ak) * d0 + ak * d1;
                    }
                }
            } else {                                                            /* :53-56 */
                dg = orc_weights_2d(kmean, ze, nz, ng, wg);
                for (int64_t ipb = 0; ipb < npb; ++ipb) {
                    const double *l = ldp + (ipv * npb + ipb) * nz;
                    for (int ig = 0; ig < ng; ++ig) {
                        double d0 = 0.0;
                        for (int iz = 0; iz < nz; ++iz) d0 += wg[ig * nz + iz] * l[iz];
                        ldm[ipb * ng + ig] = d0;
                    }
                }
            }
            double xyc[10];
            if (xyc_in) memcpy(xyc, xyc_in + ipv * 10, sizeof(xyc));
            else orc_solve2d(0.0, p[ipv], a[ipv], inc[ipv], e[ipv], w[ipv], xyc); /* :61 */
            double bt1, bt4;
            orc_bounding_box(kmean, xyc, &bt1, &bt4);                           /* :66-68 */
            bt1 -= 0.0015 + exptime;
            bt4 += 0.0015 + exptime;

            for (int64_t ipt = 0; ipt < npt; ++ipt) {                           /* :73-92 */
                double epoch = floor((times[ipt] - t0[ipv] + 0.5 * p[ipv]) / p[ipv]);
                double tc = times[ipt] - (t0[ipv] + epoch * p[ipv]);
                if (!(bt1 <= tc && tc <= bt4)) {
                    for (int64_t ipb = 0; ipb < npb; ++ipb) fl[ipb * npt + ipt] = 1.0;
                } else {
                    for (int64_t ipb = 0; ipb < npb; ++ipb) fl[ipb * npt + ipt] = 0.0;
                    for (int64_t isample = 1; isample < nsamples + 1; ++isample) {
                        double time_offset = exptime * ((isample - 0.5) / nsamples - 0.5);
                        double z = orc_sep_c(tc + time_offset, xyc);
                        double ap0, kappa;
                        orc_ccia_kite(1.0, kmean, z, &ap0, &kappa);
                        double dadk = 2.0 * kmean * kappa;
                        double g = z / (1.0 + kmean);
                        if (z <= 1.0 - kmax) {
                            for (int64_t ipb = 0; ipb < npb; ++ipb) {
                                double ipl = orc_interp_ldm(g, dg, ldm + ipb * ng, ng);
                                double is = istar[ipv * npb + ipb];
                                fl[ipb * npt + ipt] += (is - ipl * ap0 * afac[ipb]) / is;
                            }
                        } else {
                            for (int64_t ipb = 0; ipb < npb; ++ipb) {
                                double ipl = orc_interp_ldm(g, dg, ldm + ipb * ng, ng);
                                double is = istar[ipv * npb + ipb];
                                fl[ipb * npt + ipt] += (is - ipl * (ap0 + (kv[ipb] - kmean) * dadk)) / is;
                            }
                        }
                    }
                    for (int64_t ipb = 0; ipb < npb; ++ipb) fl[ipb * npt + ipt] /= nsamples;
                }
            }
        }
        free(ldm); free(wg); free(afac);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* White-noise log likelihood: pytransit/lpf/loglikelihood/wnloglikelihood.py:22-35            */
/* ------------------------------------------------------------------------------------------ */

/* o[npt], m[npv,npt], e[npv,nerr], slices[nsl,2], nids[nsl] -> lnl[npv]. */
void orc_lnlike_normal(const double *o, const double *m, int64_t npv, int64_t npt, const double *e,
                       int64_t nerr, const int64_t *slices, const int64_t *nids, int64_t nsl,
                       double *lnl) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < npv; ++i) {
        double acc = 0.0;
        for (int64_t isl = 0; isl < nsl; ++isl) {
            double _e = e[i * nerr + nids[isl]];
            for (int64_t j = slices[isl * 2]; j < slices[isl * 2 + 1]; ++j) {
                double r = (o[j] - m[i * npt + j]) / _e;
                acc += -log(_e) - 0.5 * log(2 * ORC_PI) - 0.5 * (r * r);
            }
        }
        lnl[i] = acc;
    }
}

#ifdef __cplusplus
}
#endif
