// ptb_ts_kernels.cuh -- sm_100a kernels of the transmission-spectroscopy path
// (models/roadrunner/model_trspec.py:11-93, tsmodel.py:46-130).
//
//   k_ts_setup   per vector: k-mean / k-max, orbit, contact times, the (ng x nz) weight matrix at
//                k-mean (table blend or direct, model_trspec.py:48-56)
//   k_ts_ld      limb-darkening profile + I* per (vector, channel) for the named laws
//   k_ts_ldm     the dense contraction ldm[pb, ig] = sum_iz ldp[pb, iz] W[ig, iz] per vector on the
//                fp64 tensor-core path (mma.sync m8n8k4 DMMA; tcgen05 has no f64 kind), operands
//                staged in shared memory by TMA bulk copies
//   k_ts_flux    flux[npv, npb, npt]: one separation / lens area per (vector, time, sub-sample)
//                shared by all channels; per channel one lerp + first-order area correction
//
// HBM layout:
//   tsorb [npv][24]        cx[5] cy[5] p 1/p T1 T4 good - | kmean kmax 1/(1+kmean) kmean^2 ...
//   tsw   [npv][ng][nz]    weight matrix at kmean
//   tsldm [npv][npb][ldt]  limb-darkening means, ldt = ng rounded up to a multiple of 8
//   tsrec [npv][npb][4]    1/I*, k^2/kmean^2, k - kmean, -
//   flux  [npv][npb][npt]
#pragma once
#include "ptb_kernels.cuh"

namespace ptb {

constexpr int TSORB_STRIDE = 24;
constexpr int TSORB_KMEAN = 16, TSORB_KMAX = 17, TSORB_INV1K = 18, TSORB_K2 = 19;

struct TsSetupParams {
    const double *k;  // [npv][npb]
    const double *p, *a, *inc, *e, *w;
    const double *xyc_in;
    const double *W, *ze, *gs;
    double *tsorb, *tsw;
    int npv, npb, nk, ng, nz, use_table;
    double kmin, dk;
};

__global__ void __launch_bounds__(128) k_ts_setup(const __grid_constant__ TsSetupParams P) {
    __shared__ double s_sum[4], s_max[4];
    const int ipv = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double a = P.a[ipv], e = P.e[ipv];
    double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;
    const bool good = !(isnan(a) || (a <= 1.0) || (e < 0.0));  // model_trspec.py:37 (no ldp check)
    if (!good) {
        if (tid < TSORB_STRIDE) orb[tid] = (tid == ORB_GOOD) ? 0.0 : nan("");
        return;
    }
    // kmean, kmax over the channels (model_trspec.py:41-42)
    const double *kv = P.k + (size_t)ipv * P.npb;
    double s = 0.0, m = -INFINITY;
    bool anynan = false;
    for (int i = tid; i < P.npb; i += 128) {
        const double v = kv[i];
        s += v;
        m = fmax(m, v);
        anynan |= isnan(v);
    }
    if (anynan) m = nan("");
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, m, o);
        m = (isnan(m) || isnan(t)) ? nan("") : fmax(m, t);
    }
    if (lane == 0) { s_sum[warp] = s; s_max[warp] = m; }
    __syncthreads();
    const double kmean = (s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3]) / P.npb;
    double kmax = s_max[0];
    for (int i = 1; i < 4; ++i) kmax = (isnan(kmax) || isnan(s_max[i])) ? nan("") : fmax(kmax, s_max[i]);

    if (warp == 0) {
        solve_orbit_warp(lane, P.p[ipv], a, P.inc[ipv], e, P.w[ipv], kmean,
                         P.xyc_in ? P.xyc_in + (size_t)ipv * 10 : nullptr, orb);
        if (lane == 0) {
            orb[ORB_GOOD] = 1.0;
            orb[TSORB_KMEAN] = kmean;
            orb[TSORB_KMAX] = kmax;
            orb[TSORB_INV1K] = 1.0 / (1.0 + kmean);
            orb[TSORB_K2] = kmean * kmean;
            for (int j = 20; j < TSORB_STRIDE; ++j) orb[j] = 0.0;
        }
    }
    // weight matrix at kmean: table blend when precompute_weights and kmin <= kmean <= max(k)
    // (the reference's kmax argument is shadowed by max(k[ipv])), else direct.
    double *wout = P.tsw + (size_t)ipv * P.ng * P.nz;
    const int rowlen = P.ng * P.nz;
    if (P.use_table && P.kmin <= kmean && kmean <= kmax) {
        int ik = (int)floor((kmean - P.kmin) / P.dk);
        const double ak = (kmean - P.kmin - ik * P.dk) / P.dk;
        const int ik1 = min(ik + 1, P.nk - 1);
        ik = min(ik, P.nk - 1);
        const double *w0 = P.W + (size_t)ik * rowlen, *w1 = P.W + (size_t)ik1 * rowlen;
        for (int i = tid; i < rowlen; i += 128) wout[i] = (1.0 - ak) * w0[i] + ak * w1[i];
    } else {
        for (int ig = tid; ig < P.ng; ig += 128) weight_row(kmean, P.gs[ig], P.ze, P.nz, wout + (size_t)ig * P.nz, 1);
    }
}

// limb-darkening profile and I* for the named laws: one warp per (vector, channel)
struct TsLdParams {
    const double *ldc;  // [npv][npb][nld]
    const double *mu, *ldmu200, *ldz200;
    double *ldp, *istar;
    long long nrows;  // npv*npb
    int nld, law, nz;
};

__global__ void __launch_bounds__(128) k_ts_ld(const __grid_constant__ TsLdParams P) {
    __shared__ double scr[4][200];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * 4 + warp;
    if (row >= P.nrows) return;
    const double *pv = P.ldc + row * P.nld;
    for (int iz = lane; iz < P.nz; iz += 32) P.ldp[row * P.nz + iz] = ld_intensity(P.law, P.mu[iz], pv, P.nld);
    double is;
    if (!ld_integral(P.law, pv, is)) is = istar_numeric_warp(lane, P.law, pv, P.nld, P.ldmu200, P.ldz200, scr[warp]);
    if (lane == 0) P.istar[row] = is;
}

// ---------------------------------------------------------------------------------------------
// LD contraction on the fp64 tensor cores.
// CTA = (vector, tile of 64 channels), 8 warps; warp w owns channels [8w, 8w+8) x all ng.
// A = ldp tile [64][nz] (row-major, K = nz), B = W [ng][nz] ("col-major" K x N), both staged in
// shared memory with per-row TMA bulk copies into rows padded to nz+4 doubles (bank-conflict-free
// fragment loads).  D[pb][ig] accumulates over nz/4 DMMA k-steps.
// ---------------------------------------------------------------------------------------------
struct TsLdmParams {
    const double *tsw, *ldp, *istar, *k, *tsorb;
    double *tsldm, *tsrec;
    int npv, npb, ng, nz, ldt, rs;  // rs: padded shared-memory row stride (doubles), >= nz rounded up to 4
};

constexpr int TSL_PB = 64;

template <int NT>  // NT = number of 8-wide ig tiles (ldt/8), compile-time for register blocking
__global__ void __launch_bounds__(256) k_ts_ldm(const __grid_constant__ TsLdmParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TSL_RS = P.rs;
    double *sB = reinterpret_cast<double *>(smem_raw);   // [NT*8][rs]
    double *sA = sB + NT * 8 * TSL_RS;                    // [TSL_PB][rs]
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntile = (P.npb + TSL_PB - 1) / TSL_PB;
    const int ipv = blockIdx.x / ntile;
    const int pb0 = (blockIdx.x - ipv * ntile) * TSL_PB;
    const double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;
    if (orb[ORB_GOOD] == 0.0) return;
    const int nz = P.nz, ng = P.ng;
    const int nrowsA = min(TSL_PB, P.npb - pb0);

    // zero the padding (rows beyond ng / npb, columns beyond nz) so it contributes nothing
    for (int i = tid; i < (NT * 8 + TSL_PB) * TSL_RS; i += 256) {
        const int r = i / TSL_RS, c = i - r * TSL_RS;
        const bool isB = r < NT * 8;
        const bool live = isB ? (r < ng && c < nz) : ((r - NT * 8) < nrowsA && c < nz);
        if (!live) sB[i] = 0.0;
    }
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (warp == 0) {
        const uint32_t rb = (uint32_t)nz * 8u;
        if (lane == 0) mbar_expect_tx(&bar, rb * (uint32_t)(ng + nrowsA));
        __syncwarp();
        const double *wsrc = P.tsw + (size_t)ipv * ng * nz;
        const double *asrc = P.ldp + ((size_t)ipv * P.npb + pb0) * nz;
        for (int r = lane; r < ng; r += 32) tma_load_1d(sB + r * TSL_RS, wsrc + (size_t)r * nz, rb, &bar);
        for (int r = lane; r < nrowsA; r += 32) tma_load_1d(sA + r * TSL_RS, asrc + (size_t)r * nz, rb, &bar);
    }
    mbar_wait(&bar, 0);

    double acc[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = 0.0;
    const int fr = lane >> 2, fc = lane & 3;
    const double *arow = sA + (warp * 8 + fr) * TSL_RS + fc;
    const double *brow = sB + fr * TSL_RS + fc;
    const int ksteps = (nz + 3) / 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        const double av = arow[ks * 4];
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma_m8n8k4(acc[n][0], acc[n][1], av, brow[n * 8 * TSL_RS + ks * 4]);
    }
    const int pb = pb0 + warp * 8 + fr;
    if (pb < P.npb) {
        double *out = P.tsldm + ((size_t)ipv * P.npb + pb) * P.ldt + fc * 2;
#pragma unroll
        for (int n = 0; n < NT; ++n) *reinterpret_cast<double2 *>(out + n * 8) = make_double2(acc[n][0], acc[n][1]);
    }
    // per-channel record (model_trspec.py:43,87,91)
    if (tid < nrowsA) {
        const int q = pb0 + tid;
        const double kmean = orb[TSORB_KMEAN];
        const double kk = P.k[(size_t)ipv * P.npb + q];
        double *rec = P.tsrec + ((size_t)ipv * P.npb + q) * 4;
        rec[0] = 1.0 / P.istar[(size_t)ipv * P.npb + q];
        rec[1] = (kk * kk) / (kmean * kmean);
        rec[2] = kk - kmean;
        rec[3] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// flux[npv, npb, npt].  CTA = (vector, time tile, channel split); thread = VEC consecutive points.
// Geometry per (point, sub-sample) is computed once: registers when nsamples == 1, shared memory
// otherwise.  The channel loop then costs two gathered ldm loads, a lerp and ~6 flops per output,
// written with 16-byte streaming stores (npt is the contiguous axis).
// ---------------------------------------------------------------------------------------------
struct TsFluxParams {
    const double *time, *tsorb, *t0, *tsldm, *tsrec;
    double *flux;
    long long npt;
    int npv, npb, ng, ldt, ns, ntiles, pbsplit;
    double exptime, dg, inv_dg;
};

struct TsGeo {
    double alpha, ap0, dadk;
    int i0;    // lower ldm node; -1: planet off the disk (no contribution)
    int full;  // z <= 1 - kmax
};

__device__ __forceinline__ TsGeo ts_geometry(double t, const double *cx, const double *cy, double kmean, double k2,
                                             double inv1k, double kmax, double dg, double inv_dg, int ng) {
    TsGeo G;
    const double z = sep_poly(t, cx, cy);
    double kap;
    kite_area(kmean, k2, z, G.ap0, kap);
    G.dadk = 2.0 * kmean * kap;
    G.full = (z <= 1.0 - kmax) ? 1 : 0;
    const double g = z * inv1k;
    if (g > 1.0) {  // interpolate_mean_limb_darkening_s returns 0 (common.py:229-230)
        G.i0 = -1;
        G.alpha = 0.0;
    } else {
        const int i = (int)floor(g * inv_dg);
        G.alpha = (g - i * dg) * inv_dg;
        G.i0 = min(i, ng - 1);
    }
    return G;
}

template <int VEC, bool MULTI>
__global__ void __launch_bounds__(256) k_ts_flux(const __grid_constant__ TsFluxParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // MULTI: TsGeo[ns][256*VEC]
    const int tid = threadIdx.x;
    constexpr int TILE = 256 * VEC;
    int b = blockIdx.x;
    const int split = b % P.pbsplit;
    b /= P.pbsplit;
    const int tile = b % P.ntiles;
    const int ipv = b / P.ntiles;
    const long long npt = P.npt;
    const long long i0 = (long long)tile * TILE + (long long)tid * VEC;
    const bool inr = i0 < npt;
    const int pbper = (P.npb + P.pbsplit - 1) / P.pbsplit;
    const int pb_beg = split * pbper, pb_end = min(P.npb, pb_beg + pbper);
    double *fbase = P.flux + (size_t)ipv * P.npb * npt + i0;
    const double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;

    if (orb[ORB_GOOD] == 0.0) {  // flux[ipv, :, :] = nan (model_trspec.py:38)
        if (inr) {
            double v[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = nan("");
            for (int pb = pb_beg; pb < pb_end; ++pb) VecIO<VEC, double>::store(fbase + (size_t)pb * npt, v);
        }
        return;
    }
    const double p = orb[ORB_P], invp = orb[ORB_INVP];
    const double pad = 0.0015 + P.exptime;  // model_trspec.py:67-68
    const double lo = orb[ORB_T1] - pad, hi = orb[ORB_T4] + pad;
    const double t0 = P.t0[ipv];
    const double kmean = orb[TSORB_KMEAN], kmax = orb[TSORB_KMAX], inv1k = orb[TSORB_INV1K], k2 = orb[TSORB_K2];
    double cx[5], cy[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { cx[j] = orb[j]; cy[j] = orb[5 + j]; }

    bool inbox[VEC];
    TsGeo G[VEC];
    TsGeo *sG = reinterpret_cast<TsGeo *>(smem_raw);
    double tv[VEC];
    if (inr) VecIO<VEC, double>::load(P.time + i0, tv);
    bool any = false;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        inbox[j] = false;
        if (inr) {
            const double epoch = floor(fma(tv[j] - t0, invp, 0.5));
            const double tc = tv[j] - __dadd_rn(t0, __dmul_rn(epoch, p));
            inbox[j] = (lo <= tc) && (tc <= hi);
            if (inbox[j]) {
                if (!MULTI) {
                    G[j] = ts_geometry(tc, cx, cy, kmean, k2, inv1k, kmax, P.dg, P.inv_dg, P.ng);
                } else {
                    for (int s = 0; s < P.ns; ++s) {
                        const double off = P.exptime * (((s + 1) - 0.5) / P.ns - 0.5);
                        sG[(size_t)s * TILE + tid * VEC + j] =
                            ts_geometry(tc + off, cx, cy, kmean, k2, inv1k, kmax, P.dg, P.inv_dg, P.ng);
                    }
                }
            }
        }
        any |= inbox[j];
    }
    if (!inr) return;
    // (each thread reads back only what it wrote itself: no barrier needed)

    const double *ldm = P.tsldm + (size_t)ipv * P.npb * P.ldt;
    const double *rec = P.tsrec + (size_t)ipv * P.npb * 4;
    const int ngm1 = P.ng - 1;
    if (!any) {
        double v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = 1.0;
        for (int pb = pb_beg; pb < pb_end; ++pb) VecIO<VEC, double>::store(fbase + (size_t)pb * npt, v);
        return;
    }
    for (int pb = pb_beg; pb < pb_end; ++pb) {
        const double2 r01 = __ldg(reinterpret_cast<const double2 *>(rec + (size_t)pb * 4));
        const double dkk = __ldg(rec + (size_t)pb * 4 + 2);
        const double inv_istar = r01.x, afac = r01.y;
        const double *row = ldm + (size_t)pb * P.ldt;
        double v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            v[j] = 1.0;
            if (inbox[j]) {
                if (!MULTI) {
                    const TsGeo &g = G[j];
                    double ip = 0.0;
                    if (g.i0 >= 0) ip = (1.0 - g.alpha) * __ldg(row + g.i0) + g.alpha * __ldg(row + min(g.i0 + 1, ngm1));
                    const double x = g.full ? g.ap0 * afac : g.ap0 + dkk * g.dadk;
                    v[j] = 1.0 - ip * x * inv_istar;
                } else {
                    double acc = 0.0;
                    for (int s = 0; s < P.ns; ++s) {
                        const TsGeo g = sG[(size_t)s * TILE + tid * VEC + j];
                        double ip = 0.0;
                        if (g.i0 >= 0) ip = (1.0 - g.alpha) * __ldg(row + g.i0) + g.alpha * __ldg(row + min(g.i0 + 1, ngm1));
                        const double x = g.full ? g.ap0 * afac : g.ap0 + dkk * g.dadk;
                        acc += 1.0 - ip * x * inv_istar;
                    }
                    v[j] = acc / P.ns;
                }
            }
        }
        VecIO<VEC, double>::store(fbase + (size_t)pb * npt, v);
    }
}

// ---------------------------------------------------------------------------------------------
// Two-pass flux for one sample per point (the common spectroscopy case):
//   k_ts_geo    geometry per (vector, time): fold, box test, separation, lens area at k-mean, kappa0,
//               ld-mean node + weight -> geo arrays [npv][npt] (28 bytes per point, L2 resident)
//   k_ts_flux2  CTA = (vector, chunk of TS_CH channels): the chunk's ld-mean rows and per-channel
//               records arrive in shared memory by two TMA bulk copies; the CTA then walks the whole
//               time axis: per point one coalesced geometry load, per channel two shared-memory
//               gathers, a lerp, the first-order area correction and a 16-byte streaming store.
// The channel loop reads nothing from global memory but the geometry, so the kernel is bound by the
// flux write-out (8 B per point).
// ---------------------------------------------------------------------------------------------
struct TsGeoParams {
    const double *time, *tsorb, *t0;
    double *galpha, *gap0, *gdadk;  // [npv][npt]
    int *gi0;                       // [npv][npt]: ld-mean node | TS_FULL, or -1 (flux exactly 1)
    long long npt;
    int npv, ng;
    double exptime, dg, inv_dg;
};
constexpr int TS_FULL = 1 << 30;

template <int VEC>
__global__ void __launch_bounds__(256) k_ts_geo(const __grid_constant__ TsGeoParams P) {
    const int ntiles = (int)((P.npt + 256LL * VEC - 1) / (256LL * VEC));
    const int ipv = blockIdx.x / ntiles, tile = blockIdx.x - ipv * ntiles;
    const long long i0 = (long long)tile * 256 * VEC + (long long)threadIdx.x * VEC;
    if (i0 >= P.npt) return;
    const double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;
    const size_t o = (size_t)ipv * P.npt + i0;
    double al[VEC], ap[VEC], da[VEC];
    int id[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { al[j] = 0.0; ap[j] = 0.0; da[j] = 0.0; id[j] = -1; }
    if (orb[ORB_GOOD] != 0.0) {
        const double p = orb[ORB_P], invp = orb[ORB_INVP];
        const double pad = 0.0015 + P.exptime;  // model_trspec.py:67-68
        const double lo = orb[ORB_T1] - pad, hi = orb[ORB_T4] + pad, t0 = P.t0[ipv];
        const double kmean = orb[TSORB_KMEAN], kmax = orb[TSORB_KMAX], inv1k = orb[TSORB_INV1K], k2 = orb[TSORB_K2];
        double cx[5], cy[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) { cx[j] = orb[j]; cy[j] = orb[5 + j]; }
        double tv[VEC];
        VecIO<VEC, double>::load(P.time + i0, tv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const double epoch = floor(fma(tv[j] - t0, invp, 0.5));
            const double tc = tv[j] - __dadd_rn(t0, __dmul_rn(epoch, p));
            if ((lo <= tc) && (tc <= hi)) {
                const TsGeo g = ts_geometry(tc, cx, cy, kmean, k2, inv1k, kmax, P.dg, P.inv_dg, P.ng);
                if (g.i0 >= 0) {
                    al[j] = g.alpha; ap[j] = g.ap0; da[j] = g.dadk;
                    id[j] = g.i0 | (g.full ? TS_FULL : 0);
                }
            }
        }
    }
    VecIO<VEC, double>::store_keep(P.galpha + o, al);
    VecIO<VEC, double>::store_keep(P.gap0 + o, ap);
    VecIO<VEC, double>::store_keep(P.gdadk + o, da);
    if (VEC == 2) *reinterpret_cast<int2 *>(P.gi0 + o) = make_int2(id[0], id[VEC - 1]);
    else P.gi0[o] = id[0];
}

struct TsFlux2Params {
    const double *tsorb, *tsldm, *tsrec, *galpha, *gap0, *gdadk;
    const int *gi0;
    double *flux;
    long long npt;
    int npv, npb, ng, ldt, nchunks;
};
constexpr int TS_CH = 32;  // channels per CTA

template <int VEC>
__global__ void __launch_bounds__(256) k_ts_flux2(const __grid_constant__ TsFlux2Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // [TS_CH][ldt] ld means | [TS_CH][4] records
    __shared__ __align__(8) uint64_t bar;
    double *sLdm = reinterpret_cast<double *>(smem_raw);
    double *sRec = sLdm + (size_t)TS_CH * P.ldt;
    const int tid = threadIdx.x;
    const int ipv = blockIdx.x / P.nchunks, chunk = blockIdx.x - ipv * P.nchunks;
    const int pb0 = chunk * TS_CH, nch = min(TS_CH, P.npb - pb0);
    const long long npt = P.npt;
    double *fbase = P.flux + ((size_t)ipv * P.npb + pb0) * npt;
    const bool good = P.tsorb[(size_t)ipv * TSORB_STRIDE + ORB_GOOD] != 0.0;
    if (good) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            const uint32_t b1 = (uint32_t)nch * P.ldt * 8u, b2 = (uint32_t)nch * 32u;
            mbar_expect_tx(&bar, b1 + b2);
            tma_load_1d(sLdm, P.tsldm + ((size_t)ipv * P.npb + pb0) * P.ldt, b1, &bar);
            tma_load_1d(sRec, P.tsrec + ((size_t)ipv * P.npb + pb0) * 4, b2, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    }
    const int ngm1 = P.ng - 1;
    const size_t gbase = (size_t)ipv * npt;
    for (long long i0 = (long long)tid * VEC; i0 < npt; i0 += 256 * VEC) {
        double v[VEC];
        if (!good) {  // flux[ipv, :, :] = nan (model_trspec.py:38)
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = nan("");
            for (int c = 0; c < nch; ++c) VecIO<VEC, double>::store(fbase + (size_t)c * npt + i0, v);
            continue;
        }
        double al[VEC], ap[VEC], da[VEC];
        int id[VEC];
        if (VEC == 2) {
            const int2 t = __ldg(reinterpret_cast<const int2 *>(P.gi0 + gbase + i0));
            id[0] = t.x; id[VEC - 1] = t.y;
        } else {
            id[0] = __ldg(P.gi0 + gbase + i0);
        }
        bool any = false;
#pragma unroll
        for (int j = 0; j < VEC; ++j) any |= id[j] >= 0;
        if (!__any_sync(__activemask(), any)) {  // the whole warp is out of transit: ones for every channel
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = 1.0;
            for (int c = 0; c < nch; ++c) VecIO<VEC, double>::store(fbase + (size_t)c * npt + i0, v);
            continue;
        }
        VecIO<VEC, double>::load(P.galpha + gbase + i0, al);
        VecIO<VEC, double>::load(P.gap0 + gbase + i0, ap);
        VecIO<VEC, double>::load(P.gdadk + gbase + i0, da);
        int n0[VEC], n1[VEC];
        bool full[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            full[j] = (id[j] & TS_FULL) != 0 && id[j] >= 0;
            n0[j] = id[j] >= 0 ? (id[j] & ~TS_FULL) : 0;
            n1[j] = min(n0[j] + 1, ngm1);
        }
#pragma unroll 2
        for (int c = 0; c < nch; ++c) {
            const double *row = sLdm + (size_t)c * P.ldt;
            const double2 r01 = *reinterpret_cast<const double2 *>(sRec + c * 4);  // 1/I*, k^2/kmean^2
            const double dkk = sRec[c * 4 + 2];                                      // k - kmean
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                v[j] = 1.0;
                if (id[j] >= 0) {
                    const double ip = (1.0 - al[j]) * row[n0[j]] + al[j] * row[n1[j]];
                    const double x = full[j] ? ap[j] * r01.y : ap[j] + dkk * da[j];
                    v[j] = 1.0 - ip * x * r01.x;
                }
            }
            VecIO<VEC, double>::store(fbase + (size_t)c * npt + i0, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tabulated-profile limb darkening (models/numba/ldtkldm.py:22-60,77-91): trilinear blend of the
// 8 surrounding table nodes per vector, then I* = 2 pi trapezoid(z I, z).  One warp per
// (vector, channel) row; the nmu nodes are contiguous in the table, so loads coalesce.
// ---------------------------------------------------------------------------------------------
struct LdtkParams {
    const double *profiles, *xs, *ys, *zs, *mu;
    double *ldp, *istar;
    long long npv;
    int nx, ny, nz3, npb, nmu;
    double x0, dx, y0, dy, z0, dz;
};

__global__ void __launch_bounds__(128) k_ldtk_profiles(const __grid_constant__ LdtkParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * 4 + warp;
    if (row >= P.npv * P.npb) return;
    const long long ipv = row / P.npb;
    const int ipb = (int)(row - ipv * P.npb);
    // clamp as coded: upper limit x0 + n*dx, one cell past the last node (ldtkldm.py:42-44)
    const double x = fmin(fmax(P.xs[ipv], P.x0), P.x0 + P.nx * P.dx);
    const double y = fmin(fmax(P.ys[ipv], P.y0), P.y0 + P.ny * P.dy);
    const double z = fmin(fmax(P.zs[ipv], P.z0), P.z0 + P.nz3 * P.dz);
    int ix = (int)floor((x - P.x0) / P.dx), iy = (int)floor((y - P.y0) / P.dy), iz = (int)floor((z - P.z0) / P.dz);
    const double ax = (x - P.x0 - ix * P.dx) / P.dx, ay = (y - P.y0 - iy * P.dy) / P.dy, az = (z - P.z0 - iz * P.dz) / P.dz;
    const double rx = 1.0 - ax, ry = 1.0 - ay, rz = 1.0 - az;
    const double a1 = rx * ry * rz, a2 = ax * ry * rz, b1 = rx * ay * rz, b2 = rx * ry * az;
    const double c1 = ax * ry * az, c2 = rx * ay * az, d1 = ax * ay * rz, d2 = ax * ay * az;
    // keep the 2x2x2 cell inside the table (the reference slices past the end for values on the
    // upper boundary; weights of the missing nodes are then zero)
    const int ix1 = min(ix + 1, P.nx - 1), iy1 = min(iy + 1, P.ny - 1), iz1 = min(iz + 1, P.nz3 - 1);
    ix = min(ix, P.nx - 1); iy = min(iy, P.ny - 1); iz = min(iz, P.nz3 - 1);
    const size_t s_pb = (size_t)P.nmu, s_z = (size_t)P.npb * P.nmu, s_y = s_z * P.nz3, s_x = s_y * P.ny;
    const double *base = P.profiles + (size_t)ipb * s_pb;
    auto at = [&](int jx, int jy, int jz, int i) { return __ldg(base + jx * s_x + jy * s_y + jz * s_z + i); };
    double *out = P.ldp + row * P.nmu;
    double s = 0.0;
    for (int i0 = 0; i0 < P.nmu; i0 += 32) {
        const int i = i0 + lane;
        double v = 0.0, zi = 0.0;
        if (i < P.nmu) {
            v = at(ix, iy, iz, i) * a1 + at(ix1, iy, iz, i) * a2 + at(ix, iy1, iz, i) * b1 + at(ix, iy, iz1, i) * b2 +
                at(ix1, iy, iz1, i) * c1 + at(ix, iy1, iz1, i) * c2 + at(ix1, iy1, iz, i) * d1 + at(ix1, iy1, iz1, i) * d2;
            out[i] = v;
            zi = sqrt(1.0 - P.mu[i] * P.mu[i]);
        }
        // trapezoid term between node i-1 and i (integrate_profiles_set, ldtkldm.py:86-89)
        double vp = __shfl_up_sync(0xffffffffu, v, 1), zp = __shfl_up_sync(0xffffffffu, zi, 1);
        if (lane == 0 && i0 > 0) {
            const double mp = P.mu[i - 1];
            zp = sqrt(1.0 - mp * mp);
            vp = out[i - 1];  // written by lane 31 in the previous pass (made visible by __syncwarp)
        }
        if (i < P.nmu && i > 0) s += (zi - zp) * 0.5 * (zi * v + zp * vp);
        __syncwarp();
    }
    s = warp_sum(s);
    if (lane == 0) P.istar[row] = 2.0 * kPi * s;
}

}  // namespace ptb
