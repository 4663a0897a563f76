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
//   tsw   [npv][ng][rs]    weight matrix at kmean, rows padded to rs = nz rounded up to 4, + 4 (pad = 0)
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
    double *tsorb, *tsw;   // tsw[npv][ng][rs]: rows padded to rs doubles (pad columns zero) so that k_ts_ldm can
    int npv, npb, nk, ng, nz, use_table, rs;  // stage the whole matrix with ONE bulk copy, bank-conflict free
    double kmin, dk;
};

constexpr int TSS_THREADS = 256;

__global__ void __launch_bounds__(TSS_THREADS) k_ts_setup(const __grid_constant__ TsSetupParams P) {
    __shared__ double s_sum[TSS_THREADS / 32], s_max[TSS_THREADS / 32];
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
    for (int i = tid; i < P.npb; i += TSS_THREADS) {
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
    double ksum = s_sum[0];
    double kmax = s_max[0];
    for (int i = 1; i < TSS_THREADS / 32; ++i) {
        ksum += s_sum[i];
        kmax = (isnan(kmax) || isnan(s_max[i])) ? nan("") : fmax(kmax, s_max[i]);
    }
    const double kmean = ksum / P.npb;

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
    const int rs = P.rs, nz = P.nz;
    double *wout = P.tsw + (size_t)ipv * P.ng * rs;
    const int rowlen = P.ng * nz;
    if (P.use_table && P.kmin <= kmean && kmean <= kmax) {
        int ik = (int)floor((kmean - P.kmin) / P.dk);
        const double ak = (kmean - P.kmin - ik * P.dk) / P.dk;
        const int ik1 = min(ik + 1, P.nk - 1);
        ik = min(ik, P.nk - 1);
        const double *w0 = P.W + (size_t)ik * rowlen, *w1 = P.W + (size_t)ik1 * rowlen;
        for (int i = tid; i < P.ng * rs; i += TSS_THREADS) {
            const int g = i / rs, c = i - g * rs;
            wout[i] = c < nz ? (1.0 - ak) * w0[g * nz + c] + ak * w1[g * nz + c] : 0.0;
        }
    } else {
        // calculate_weights_2d (common.py:152-185) in shared memory: the ng x nz annulus areas in parallel, the
        // running difference and running-sum normalisation of each row in index order (as weight_row does), then
        // one coalesced write of the padded matrix
        extern __shared__ __align__(16) double s_w[];   // [ng][nz + 1]  (odd stride: conflict-free row walks)
        const int ws = nz + 1;
        for (int i = tid; i < P.ng * nz; i += TSS_THREADS) {
            const int g = i / nz, c = i - g * nz;
            s_w[g * ws + c] = ccia_acos(P.ze[c], kmean, P.gs[g] * (1.0 + kmean));
        }
        __syncthreads();
        for (int g = tid; g < P.ng; g += TSS_THREADS) {
            double *w = s_w + (size_t)g * ws;
            double a0 = w[0], sum = a0;
            for (int i = 1; i < nz; ++i) {
                const double a1 = w[i];
                const double d = a1 - a0;
                w[i] = d;
                a0 = a1;
                sum += d;
            }
            for (int i = 0; i < nz; ++i) w[i] /= sum;
        }
        __syncthreads();
        for (int i = tid; i < P.ng * rs; i += TSS_THREADS) {
            const int g = i / rs, c = i - g * rs;
            wout[i] = c < nz ? s_w[g * ws + c] : 0.0;
        }
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
// CTA = (vector, run of `tpc` tiles of 64 channels), 8 warps; warp w owns channels [8w, 8w+8) x all ng of a tile.
// The vector's weight matrix is staged once per CTA; profile tiles are double buffered (tile t+1 is in flight
// while tile t is contracted).
// A = ldp tile [64][nz] (row-major, K = nz), B = W [ng][nz] ("col-major" K x N), both staged in
// shared memory with per-row TMA bulk copies into rows padded to nz+4 doubles (bank-conflict-free
// fragment loads).  D[pb][ig] accumulates over nz/4 DMMA k-steps.
// ---------------------------------------------------------------------------------------------
struct TsLdmParams {
    const double *tsw, *ldp, *istar, *k, *tsorb;
    double *tsldm, *tsrec;
    int npv, npb, ng, nz, ldt, rs;  // rs: padded shared-memory row stride (doubles), >= nz rounded up to 4
    int tpc;                        // channel tiles per CTA (the weight matrix is staged once per CTA)
};

constexpr int TSL_PB = 64;

template <int NT>  // NT = number of 8-wide ig tiles (ldt/8), compile-time for register blocking
__global__ void __launch_bounds__(256) k_ts_ldm(const __grid_constant__ TsLdmParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TSL_RS = P.rs;
    double *sB = reinterpret_cast<double *>(smem_raw);   // [NT*8][rs]  weight matrix, rows padded in global memory too
    double *sA0 = sB + NT * 8 * TSL_RS;                   // 2 x [TSL_PB][ars] profile tiles; ars = nz (bulk) or rs (per-row)
    __shared__ __align__(8) uint64_t bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntile = (P.npb + TSL_PB - 1) / TSL_PB;
    const int ngrp = (ntile + P.tpc - 1) / P.tpc;         // CTAs per vector
    const int ipv = blockIdx.x / ngrp;
    const int tile0 = (blockIdx.x - ipv * ngrp) * P.tpc, tile1 = min(ntile, tile0 + P.tpc);
    const double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;
    if (orb[ORB_GOOD] == 0.0) return;
    const int nz = P.nz, ng = P.ng;
    // nz a multiple of 4: the k-steps never run past a row, so a profile tile needs no padding and arrives
    // as ONE bulk copy (its rows are contiguous in ldp); otherwise rows are staged one by one into padded rows
    const bool bulkA = (nz & 3) == 0;
    const int ars = bulkA ? nz : TSL_RS;
    const size_t atile = (size_t)TSL_PB * ars;

    // zero what the copies never write: B rows beyond ng; A pad columns (per-row staging only)
    for (int i = tid; i < (NT * 8 - ng) * TSL_RS; i += 256) sB[ng * TSL_RS + i] = 0.0;
    if (!bulkA)
        for (int i = tid; i < 2 * TSL_PB * ars; i += 256)
            if (i % ars >= nz) sA0[i] = 0.0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncthreads();

    // stage tile `t` into buffer `b` (warp 0); the weight matrix rides on the first tile's barrier
    auto stage = [&](int t, int b, bool with_w) {
        const int pb0 = t * TSL_PB, nrows = min(TSL_PB, P.npb - pb0);
        const uint32_t rb = (uint32_t)nz * 8u, wb = (uint32_t)ng * TSL_RS * 8u;
        const double *asrc = P.ldp + ((size_t)ipv * P.npb + pb0) * nz;
        double *dst = sA0 + (size_t)b * atile;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar[b], (with_w ? wb : 0u) + rb * (uint32_t)nrows);
            if (with_w) tma_load_1d(sB, P.tsw + (size_t)ipv * ng * TSL_RS, wb, &bar[b]);
            if (bulkA) tma_load_1d(dst, asrc, rb * (uint32_t)nrows, &bar[b]);
        }
        __syncwarp();
        if (!bulkA)
            for (int r = lane; r < nrows; r += 32) tma_load_1d(dst + r * ars, asrc + (size_t)r * nz, rb, &bar[b]);
    };
    if (warp == 0) stage(tile0, 0, true);

    const int fr = lane >> 2, fc = lane & 3;
    const int ksteps = (nz + 3) / 4;
    const double kmean = orb[TSORB_KMEAN];
    for (int t = tile0; t < tile1; ++t) {
        const int it = t - tile0, b = it & 1;
        const int pb0 = t * TSL_PB, nrowsA = min(TSL_PB, P.npb - pb0);
        // every warp has finished reading buffer b^1 (tile t-1): refill it with tile t+1
        __syncthreads();
        if (warp == 0 && t + 1 < tile1) stage(t + 1, b ^ 1, false);
        mbar_wait(&bar[b], (it >> 1) & 1);
        double *sA = sA0 + (size_t)b * atile;
        // rows beyond the last tile's channels hold stale data: finite garbage times computed-never-stored is
        // fine, NaN is not a problem either (those accumulators are discarded)
        double acc[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = 0.0;
        const double *arow = sA + (warp * 8 + fr) * ars + fc;
        const double *brow = sB + fr * TSL_RS + fc;
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
            const double kk = P.k[(size_t)ipv * P.npb + q];
            double *rec = P.tsrec + ((size_t)ipv * P.npb + q) * 4;
            // a disk integral that is not finite makes every in-box point NaN in the reference: (I* - x) / I* (model_trspec.py:91)
            const double is = P.istar[(size_t)ipv * P.npb + q];
            rec[0] = isfinite(is) ? 1.0 / is : nan("");
            rec[1] = (kk * kk) / (kmean * kmean);
            rec[2] = kk - kmean;
            rec[3] = 0.0;
        }
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
    void *flux;   // [npv][npb][npt] fp64, or fp32 in the opt-in fp32 output mode
    long long npt;
    int npv, npb, ng, ldt, ns, ntiles, pbsplit;
    double exptime, dg, inv_dg;
};

// flux values are computed in fp64; TO is the stored type (double, or float in the opt-in fp32 output mode)
template <int VEC, typename TO>
__device__ __forceinline__ void ts_store(TO *p, const double *v) {
    TO o[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[j] = (TO)v[j];
    VecIO<VEC, TO>::store(p, o);
}

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

template <int VEC, bool MULTI, typename TO>
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
    TO *fbase = reinterpret_cast<TO *>(P.flux) + (size_t)ipv * P.npb * npt + i0;
    const double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;

    if (orb[ORB_GOOD] == 0.0) {  // flux[ipv, :, :] = nan (model_trspec.py:38)
        if (inr) {
            double v[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = nan("");
            for (int pb = pb_beg; pb < pb_end; ++pb) ts_store<VEC, TO>(fbase + (size_t)pb * npt, v);
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
        for (int pb = pb_beg; pb < pb_end; ++pb) ts_store<VEC, TO>(fbase + (size_t)pb * npt, v);
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
        ts_store<VEC, TO>(fbase + (size_t)pb * npt, v);
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
    int *gi0;                       // [npv][npt]: ld-mean node | TS_FULL; -1: outside the box (flux exactly 1); -2: in the box, off the disk
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
                } else {
                    id[j] = -2;   // inside the box, planet off the disk: (I* - 0) / I* = 1, or NaN when I* is not finite
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
    void *flux;
    long long npt;
    int npv, npb, ng, ldt, nchunks;
    // fused contraction (FUSED): the CTA computes its chunk's LD means itself from the vector's weight matrix and the
    // chunk's limb-darkening profiles, and the per-channel records from k and I*
    const double *tsw, *ldp, *istar, *k;
    int nz, rs;
};
constexpr int TS_CH = 32;  // channels per CTA

template <int VEC, typename TO, bool FUSED, int NT>
__global__ void __launch_bounds__(256) k_ts_flux2(const __grid_constant__ TsFlux2Params P) {
    // shared memory: [TS_CH][ldt] ld means | [TS_CH][4] records.  FUSED: the prologue's operands -- the vector's weight
    // matrix [ng][rs] and the chunk's profiles [TS_CH][nz] -- use the same bytes first (they are dead once the
    // accumulators sit in registers), so the kernel keeps its ~45 KB footprint and five resident CTAs per SM.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    double *sLdm = reinterpret_cast<double *>(smem_raw);
    double *sRec = sLdm + (size_t)TS_CH * P.ldt;
    const int tid = threadIdx.x;
    const int ipv = blockIdx.x / P.nchunks, chunk = blockIdx.x - ipv * P.nchunks;
    const int pb0 = chunk * TS_CH, nch = min(TS_CH, P.npb - pb0);
    const long long npt = P.npt;
    TO *fbase = reinterpret_cast<TO *>(P.flux) + ((size_t)ipv * P.npb + pb0) * npt;
    const double *orb = P.tsorb + (size_t)ipv * TSORB_STRIDE;
    const bool good = orb[ORB_GOOD] != 0.0;
    if (good && !FUSED) {
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
    if (good && FUSED) {
        // ---- prologue: ldm[c][ig] = sum_iz ldp[c][iz] W[ig][iz] for the chunk's channels on the fp64 tensor-core path ----
        // (model_trspec.py:58-59; mma.sync.m8n8k4: a warp owns 8 channels x half of the g-node tiles; nz % 4 == 0)
        const int nz = P.nz, rs = P.rs, lane = tid & 31, warp = tid >> 5;
        double *sW = reinterpret_cast<double *>(smem_raw);       // [ng][rs]   (rows padded with zeros in global memory)
        double *sA = sW + (size_t)P.ng * rs;                      // [TS_CH][nz]
        if (tid == 0) {
            mbar_init(&bar, 1);
            const uint32_t wb = (uint32_t)P.ng * rs * 8u, ab = (uint32_t)nch * nz * 8u;
            mbar_expect_tx(&bar, wb + ab);
            tma_load_1d(sW, P.tsw + (size_t)ipv * P.ng * rs, wb, &bar);
            tma_load_1d(sA, P.ldp + ((size_t)ipv * P.npb + pb0) * nz, ab, &bar);
        }
        // per-channel records (model_trspec.py:43,87,91) while the copies are in flight
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        if (tid < nch) {
            const double kmean = orb[TSORB_KMEAN];
            const double kk = P.k[(size_t)ipv * P.npb + pb0 + tid];
            const double is = P.istar[(size_t)ipv * P.npb + pb0 + tid];
            r0 = isfinite(is) ? 1.0 / is : nan("");   // a non-finite disk integral: NaN for every in-box point, (I* - x) / I*
            r1 = (kk * kk) / (kmean * kmean);
            r2 = kk - kmean;
        }
        __syncthreads();
        mbar_wait(&bar, 0);
        const int fr = lane >> 2, fc = lane & 3;
        const int cg = warp & 3, half = warp >> 2;                 // channel group (8 channels), first / second half of the tiles
        constexpr int NH = (NT + 1) / 2;
        const int n0 = half * NH, n1 = min(NT, n0 + NH);
        double acc[NH][2];
#pragma unroll
        for (int n = 0; n < NH; ++n) acc[n][0] = acc[n][1] = 0.0;
        const double *arow = sA + (size_t)(cg * 8 + fr) * nz + fc;   // rows beyond nch: stale bytes, results never read
        for (int ks = 0; ks < nz / 4; ++ks) {
            const double av = arow[ks * 4];
#pragma unroll
            for (int n = 0; n < NH; ++n) {
                const int ig = min((n0 + n) * 8 + fr, P.ng - 1);     // g nodes beyond ng alias the last row: never read
                dmma_m8n8k4(acc[n][0], acc[n][1], av, sW[(size_t)ig * rs + ks * 4 + fc]);
            }
        }
        __syncthreads();   // every warp is done with the operands: their bytes become the ld means and records
#pragma unroll
        for (int n = 0; n < NH; ++n)
            if (n0 + n < n1)
                *reinterpret_cast<double2 *>(sLdm + (size_t)(cg * 8 + fr) * P.ldt + (n0 + n) * 8 + fc * 2) = make_double2(acc[n][0], acc[n][1]);
        if (tid < nch) {
            sRec[tid * 4 + 0] = r0;
            sRec[tid * 4 + 1] = r1;
            sRec[tid * 4 + 2] = r2;
            sRec[tid * 4 + 3] = 0.0;
        }
        __syncthreads();
    }
    const int ngm1 = P.ng - 1;
    const size_t gbase = (size_t)ipv * npt;
    for (long long i0 = (long long)tid * VEC; i0 < npt; i0 += 256 * VEC) {
        double v[VEC];
        if (!good) {  // flux[ipv, :, :] = nan (model_trspec.py:38)
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = nan("");
            for (int c = 0; c < nch; ++c) ts_store<VEC, TO>(fbase + (size_t)c * npt + i0, v);
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
        for (int j = 0; j < VEC; ++j) any |= id[j] != -1;
        if (!__any_sync(__activemask(), any)) {  // the whole warp is out of transit: ones for every channel
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = 1.0;
            for (int c = 0; c < nch; ++c) ts_store<VEC, TO>(fbase + (size_t)c * npt + i0, v);
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
                } else if (id[j] == -2) {
                    v[j] = fma(0.0, r01.x, 1.0);
                }
            }
            ts_store<VEC, TO>(fbase + (size_t)c * npt + i0, v);
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

// ---------------------------------------------------------------------------------------------
// The same interpolation with the table slab of a few channels resident in shared memory.
// k_ldtk_profiles reads 8 table rows from L2 for every (vector, channel) row it writes (C4: 2.6 GB of L2
// reads for 0.33 GB of output).  Here a CTA owns `chunk` channels: the slab profiles[:, :, :, pb0:pb0+chunk, :]
// (nodes x chunk x nmu doubles) is staged ONCE with one TMA bulk copy per table node, and the CTA then
// walks the vectors, one warp per vector: L2 traffic drops to the table size and the kernel is bound by
// its coalesced output stream.  k_ldtk_cells precomputes each vector's cell (8 node ids + 8 weights,
// ldtkldm.py:22-50) so the per-element work is 8 shared-memory loads and 8 FMAs.
// ---------------------------------------------------------------------------------------------
struct LdtkCell {
    double w[8];
    int node[8];
};

__global__ void k_ldtk_cells(const __grid_constant__ LdtkParams P, LdtkCell *__restrict__ cells) {
    const long long ipv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ipv >= P.npv) return;
    const double x = fmin(fmax(P.xs[ipv], P.x0), P.x0 + P.nx * P.dx);
    const double y = fmin(fmax(P.ys[ipv], P.y0), P.y0 + P.ny * P.dy);
    const double z = fmin(fmax(P.zs[ipv], P.z0), P.z0 + P.nz3 * P.dz);
    int ix = (int)floor((x - P.x0) / P.dx), iy = (int)floor((y - P.y0) / P.dy), iz = (int)floor((z - P.z0) / P.dz);
    const double ax = (x - P.x0 - ix * P.dx) / P.dx, ay = (y - P.y0 - iy * P.dy) / P.dy, az = (z - P.z0 - iz * P.dz) / P.dz;
    const double rx = 1.0 - ax, ry = 1.0 - ay, rz = 1.0 - az;
    const int ix1 = min(ix + 1, P.nx - 1), iy1 = min(iy + 1, P.ny - 1), iz1 = min(iz + 1, P.nz3 - 1);
    ix = min(ix, P.nx - 1); iy = min(iy, P.ny - 1); iz = min(iz, P.nz3 - 1);
    auto nd = [&](int jx, int jy, int jz) { return (jx * P.ny + jy) * P.nz3 + jz; };
    LdtkCell c;
    // the term order of trilinear_interpolation (ldtkldm.py:45-50)
    c.w[0] = rx * ry * rz; c.node[0] = nd(ix, iy, iz);
    c.w[1] = ax * ry * rz; c.node[1] = nd(ix1, iy, iz);
    c.w[2] = rx * ay * rz; c.node[2] = nd(ix, iy1, iz);
    c.w[3] = rx * ry * az; c.node[3] = nd(ix, iy, iz1);
    c.w[4] = ax * ry * az; c.node[4] = nd(ix1, iy, iz1);
    c.w[5] = rx * ay * az; c.node[5] = nd(ix, iy1, iz1);
    c.w[6] = ax * ay * rz; c.node[6] = nd(ix1, iy1, iz);
    c.w[7] = ax * ay * az; c.node[7] = nd(ix1, iy1, iz1);
    cells[ipv] = c;
}

constexpr int LDS_THREADS = 512, LDS_WARPS = LDS_THREADS / 32;

__global__ void __launch_bounds__(LDS_THREADS, 2) k_ldtk_profiles_slab(const __grid_constant__ LdtkParams P, const LdtkCell *__restrict__ cells,
                                                           int chunk, int vsplit) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nmu = P.nmu, nodes = P.nx * P.ny * P.nz3;
    const int nchunks = (P.npb + chunk - 1) / chunk;
    const int ichunk = blockIdx.x % nchunks, isplit = blockIdx.x / nchunks;
    const int pb0 = ichunk * chunk, nch = min(chunk, P.npb - pb0);
    const int slab = chunk * nmu;        // doubles per node in shared memory
    const int ne = nch * nmu;            // live elements per node / per vector
    double *sT = reinterpret_cast<double *>(smem_raw);   // [nodes][chunk][nmu]
    double *sZ = sT + (size_t)nodes * slab;               // [nmu]  z = sqrt(1 - mu^2)
    double *sI = sZ + ((nmu + 1) & ~1);                   // [nodes][chunk] disk integral of every node profile

    const bool tma_ok = ((nmu & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.profiles) & 15) == 0);
    if (tma_ok) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, (uint32_t)nodes * (uint32_t)ne * 8u);
        }
        __syncthreads();
        for (int n = tid; n < nodes; n += LDS_THREADS)
            tma_load_1d(sT + (size_t)n * slab, P.profiles + ((size_t)n * P.npb + pb0) * nmu, (uint32_t)ne * 8u, &bar);
    } else {
        for (int idx = tid; idx < nodes * ne; idx += LDS_THREADS) {
            const int n = idx / ne, r = idx - n * ne;
            sT[(size_t)n * slab + r] = __ldg(P.profiles + ((size_t)n * P.npb + pb0) * nmu + r);
        }
    }
    for (int i = tid; i < nmu; i += LDS_THREADS) sZ[i] = sqrt(1.0 - P.mu[i] * P.mu[i]);
    __syncthreads();
    if (tma_ok) mbar_wait(&bar, 0);

    // The trapezoid rule (integrate_profiles_set, ldtkldm.py:86-89) is linear in the profile and the interpolated
    // profile is a weighted sum of eight node profiles: the disk integral of every (node, channel) of the slab is
    // taken ONCE per CTA, and a vector's integral is the same weighted sum of eight of them.
    for (int idx = tid; idx < nodes * nch; idx += LDS_THREADS) {
        const int n = idx / nch, ch = idx - n * nch;
        const double *pr = sT + (size_t)n * slab + ch * nmu;
        double sum = 0.0;
        for (int i = 1; i < nmu; ++i) sum += (sZ[i] - sZ[i - 1]) * 0.5 * (sZ[i] * pr[i] + sZ[i - 1] * pr[i - 1]);
        sI[n * chunk + ch] = 2.0 * kPi * sum;
    }
    __syncthreads();

    for (long long ipv = (long long)isplit * LDS_WARPS + warp; ipv < P.npv; ipv += (long long)LDS_WARPS * vsplit) {
        const LdtkCell c = cells[ipv];   // uniform across the warp
        const double *t0 = sT + (size_t)c.node[0] * slab, *t1 = sT + (size_t)c.node[1] * slab, *t2 = sT + (size_t)c.node[2] * slab,
                     *t3 = sT + (size_t)c.node[3] * slab, *t4 = sT + (size_t)c.node[4] * slab, *t5 = sT + (size_t)c.node[5] * slab,
                     *t6 = sT + (size_t)c.node[6] * slab, *t7 = sT + (size_t)c.node[7] * slab;
        double *out = P.ldp + ((size_t)ipv * P.npb + pb0) * nmu;
        for (int e = lane; e < ne; e += 32)   // the term order of trilinear_interpolation (ldtkldm.py:45-50)
            out[e] = t0[e] * c.w[0] + t1[e] * c.w[1] + t2[e] * c.w[2] + t3[e] * c.w[3] + t4[e] * c.w[4] + t5[e] * c.w[5] +
                     t6[e] * c.w[6] + t7[e] * c.w[7];
        for (int ch = lane; ch < nch; ch += 32)
            P.istar[(size_t)ipv * P.npb + pb0 + ch] =
                sI[c.node[0] * chunk + ch] * c.w[0] + sI[c.node[1] * chunk + ch] * c.w[1] + sI[c.node[2] * chunk + ch] * c.w[2] +
                sI[c.node[3] * chunk + ch] * c.w[3] + sI[c.node[4] * chunk + ch] * c.w[4] + sI[c.node[5] * chunk + ch] * c.w[5] +
                sI[c.node[6] * chunk + ch] * c.w[6] + sI[c.node[7] * chunk + ch] * c.w[7];
    }
}

}  // namespace ptb
