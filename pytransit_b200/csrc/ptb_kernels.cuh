// ptb_kernels.cuh -- sm_100a kernels of the RoadRunner population path.
//
//   k_weight_table   W[nk,ng,nz]  (common.py:188-223)                         once per model
//   k_rr_orbit       per-vector Taylor orbit coefficients + contact times       \
//   k_bin_scan/scatter  counting sort of the vectors by weight-table row          > once per evaluate
//   k_rr_ldm         LD profile, I*, LD means (TMA-staged table rows shared      /  (model_full.py:39-70)
//                    by the vectors of a group)
//   k_rr_points      the npv x npt pass: phase fold, box test, supersampled flux, optional fused
//                    chi^2 reduction (model_full.py:76-99, wnloglikelihood.py:22-35)
//   k_lnl_finish     chi^2 partials -> lnL[npv]
//   k_lnl_model      lnlike_normal on a materialised model flux
//
// HBM layout (all fp64 unless noted):
//   orb  [npv][16]            cx[5] cy[5] p 1/p T1 T4 good -
//   ldrec[npv][npb][lds]      ldm[ng] | k 1/(1+k) 1/I* k^2 | pad      (lds = ng+4 rounded up to even)
//   flux [npv][npt]           row-major, written with 16-byte stores
#pragma once
#include "ptb_math.cuh"

namespace ptb {

constexpr int ORB_STRIDE = 16;
constexpr int ORB_P = 10, ORB_INVP = 11, ORB_T1 = 12, ORB_T4 = 13, ORB_GOOD = 14;

// ---------------------------------------------------------------------------------------------
// Weight table: one thread per (ik, ig) row; annuli in index order with the reference's running
// difference and running-sum normalisation (common.py:152-185).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void weight_row(double k, double g, const double *__restrict__ ze, int nz, double *w,
                                           int stride) {
    const double b = g * (1.0 + k);
    double a0 = ccia_acos(ze[0], k, b);
    w[0] = a0;
    double s = a0;
    for (int i = 1; i < nz; ++i) {
        const double a1 = ccia_acos(ze[i], k, b);
        const double d = a1 - a0;
        w[i * stride] = d;
        a0 = a1;
        s += d;
    }
    for (int i = 0; i < nz; ++i) w[i * stride] /= s;
}

__global__ void k_weight_table(const double *__restrict__ ks, const double *__restrict__ gs,
                               const double *__restrict__ ze, int nk, int ng, int nz, double *__restrict__ W) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nk * ng) return;
    const int ik = idx / ng, ig = idx % ng;
    weight_row(ks[ik], gs[ig], ze, nz, W + (size_t)idx * nz, 1);
}

// ---------------------------------------------------------------------------------------------
// Per-vector setup (model_full.py:39-70), four launches:
//
//   k_rr_orbit    8 lanes per vector: validity, the 7 Kepler solves of the Taylor stencil (one per lane),
//                 coefficients from sub-warp shuffles, T1/T4 bisection on two lanes; also the weight-table
//                 row `ik` of the vector and a histogram of the rows in use.
//   k_bin_scan    one CTA: prefix sums of the histogram -> where each table row's vectors and its
//                 groups of up to RR_GROUP vectors start.
//   k_bin_scatter counting-sort scatter: vectors ordered by table row.
//   k_rr_ldm      one CTA per group of vectors that share a table row pair: the two rows W[ik], W[ik+1]
//                 (ng*nz*8 bytes each) arrive in shared memory through ONE pair of TMA bulk copies per
//                 group instead of one per vector, overlapped with the limb-darkening profile evaluation;
//                 then the (ng x nz).(nz) contractions, register-blocked over the group's vectors.
// Sorting by table row cuts the L2 -> SM traffic of the contraction by the group size and removes the
// serial orbit solve from the critical path of the table staging.
// ---------------------------------------------------------------------------------------------
constexpr int RR_GROUP = 8;  // vectors per k_rr_ldm CTA (register blocking factor)

template <int WIDTH>
__device__ __forceinline__ void solve_orbit_lanes(int sl, bool valid, double p, double a, double inc, double e, double w,
                                                  double kbox, const double *xyc_in, double *orb_out) {
    double cx[5], cy[5];
    if (xyc_in != nullptr) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { cx[j] = xyc_in[j]; cy[j] = xyc_in[5 + j]; }
    } else {
        double x = 0.0, y = 0.0;
        if (valid && sl < 7) {
            const double offset = mean_anomaly_offset(e, w);
            sky_position((sl - 3) * 2e-2, p, a * (1.0 - e * e), cos(inc), e, w, offset, x, y);
        }
        double vx[7], vy[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            vx[j] = __shfl_sync(0xffffffffu, x, j, WIDTH);
            vy[j] = __shfl_sync(0xffffffffu, y, j, WIDTH);
        }
        stencil_to_coeffs(vx, cx);
        stencil_to_coeffs(vy, cy);
    }
    double tcon = 0.0;
    if (valid && sl < 2) tcon = contact_point(kbox, sl == 0 ? -1.0 : 1.0, cx, cy);
    const double t1 = __shfl_sync(0xffffffffu, tcon, 0, WIDTH);
    const double t4 = __shfl_sync(0xffffffffu, tcon, 1, WIDTH);
    if (valid && sl == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { orb_out[j] = cx[j]; orb_out[5 + j] = cy[j]; }
        orb_out[ORB_P] = p;
        orb_out[ORB_INVP] = 1.0 / p;
        orb_out[ORB_T1] = t1;
        orb_out[ORB_T4] = t4;
        orb_out[ORB_GOOD] = 1.0;
        orb_out[15] = 0.0;
    }
}

// full-warp form used by the TSModel setup
__device__ __forceinline__ void solve_orbit_warp(int lane, double p, double a, double inc, double e, double w,
                                                 double kbox, const double *xyc_in, double *orb_out) {
    solve_orbit_lanes<32>(lane, true, p, a, inc, e, w, kbox, xyc_in, orb_out);
}

struct OrbitParams {
    const double *k;  // [npv][kcols]
    const double *p, *a, *inc, *e, *w;
    const double *xyc_in;  // optional injected coefficients [npv][10]
    double *orb;
    int *bin;   // [npv]  table row ik, nk = direct weights, nk+1 = invalid vector
    int *hist;  // [nk+2]
    int npv, kcols, nk;
    double kmin, kmax, dk;
};

__global__ void __launch_bounds__(256) k_rr_orbit(const __grid_constant__ OrbitParams P) {
    const int lane = threadIdx.x & 31, sl = lane & 7;
    const int ipv = (blockIdx.x * 256 + threadIdx.x) >> 3;
    const bool inr = ipv < P.npv;
    double a = 0, e = 0, k0 = 0, p = 1, inc = 0, w = 0;
    if (inr) {
        a = P.a[ipv];
        e = P.e[ipv];
        k0 = P.k[(size_t)ipv * P.kcols];
        p = P.p[ipv];
        inc = P.inc[ipv];
        w = P.w[ipv];
    }
    const bool good0 = inr && !(isnan(a) || (a <= 1.0) || (e < 0.0));  // model_full.py:40 (ldp checked later)
    double *orb = P.orb + (size_t)(inr ? ipv : 0) * ORB_STRIDE;
    solve_orbit_lanes<8>(sl, good0, p, a, inc, e, w, k0, (P.xyc_in && inr) ? P.xyc_in + (size_t)ipv * 10 : nullptr, orb);
    if (inr && sl == 0) {
        int bin = P.nk + 1;
        if (good0) {
            if ((P.kmin <= k0) && (k0 <= P.kmax)) bin = min((int)floor((k0 - P.kmin) / P.dk), P.nk - 1);
            else bin = P.nk;
        } else {
            for (int j = 0; j < ORB_STRIDE; ++j) orb[j] = (j == ORB_GOOD) ? 0.0 : nan("");
        }
        P.bin[ipv] = bin;
        atomicAdd(&P.hist[bin], 1);
    }
}

// offsets[b] = first slot of bin b in `perm`; gstart[b] = first group of bin b; gstart[nk+1] = group count.
__global__ void __launch_bounds__(256) k_bin_scan(int *hist, int *offsets, int *gstart, int *cursor, int nbins_valid, int grp) {
    __shared__ int s_cnt[1024 + 8];
    const int n = nbins_valid;  // nk + 1 bins carry work (table rows + direct); the invalid bin is last
    for (int i = threadIdx.x; i < n + 1; i += 256) s_cnt[i] = hist[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0, g = 0;
        for (int i = 0; i < n; ++i) {
            offsets[i] = off;
            gstart[i] = g;
            off += s_cnt[i];
            g += (s_cnt[i] + grp - 1) / grp;
        }
        offsets[n] = off;  // the invalid bin: listed after all work, never grouped
        gstart[n] = g;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n + 1; i += 256) {
        cursor[i] = 0;
        hist[i] = 0;  // ready for the next call; counts live on in offsets[]
    }
}

__global__ void k_bin_scatter(const int *__restrict__ bin, const int *__restrict__ offsets, int *cursor, int *perm, int npv) {
    const int ipv = blockIdx.x * blockDim.x + threadIdx.x;
    if (ipv >= npv) return;
    const int b = bin[ipv];
    perm[offsets[b] + atomicAdd(&cursor[b], 1)] = ipv;
}

// I* by the reference's numeric fallback, 2 pi trapezoid(z I(mu(z)), z) on 200 nodes
// (rrmodel.py:151-152,223-227); one warp per (pv, pb), scratch[200] in shared memory.
__device__ __forceinline__ double istar_numeric_warp(int lane, int law, const double *pv, int nld,
                                                     const double *__restrict__ ldmu, const double *__restrict__ ldz,
                                                     double *scratch) {
    for (int i = lane; i < 200; i += 32) scratch[i] = ldz[i] * ld_intensity(law, ldmu[i], pv, nld);
    __syncwarp();
    double s = 0.0;
    for (int i = 1 + lane; i < 200; i += 32) s += (ldz[i] - ldz[i - 1]) * (scratch[i] + scratch[i - 1]) * 0.5;
    __syncwarp();
    return 2.0 * kPi * warp_sum(s);
}

struct LdmParams {
    const double *k;      // [npv][kcols]
    const double *ld;     // ldc[npv][npb][nld] or ldp[npv][npb][nz]
    const double *istar;  // [npv][npb] (profiles only)
    const double *W, *ze, *mu, *gs, *ldmu200, *ldz200;
    const int *offsets, *gstart, *perm;
    double *orb, *ldrec, *ldp_out, *istar_out;
    int npv, kcols, npb, nld, law, nk, ng, nz, lds, grp;  // grp <= RR_GROUP vectors per CTA
    double kmin, dk;
};

__global__ void __launch_bounds__(256) k_rr_ldm(const __grid_constant__ LdmParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_bin, s_first, s_cnt;
    __shared__ int s_pv[RR_GROUP];
    __shared__ double s_ak[RR_GROUP];
    const int ng = P.ng, nz = P.nz, npb = P.npb;
    const int rowlen = ng * nz;
    double *sW = reinterpret_cast<double *>(smem_raw);       // [2][ng][nz]
    double *sLdp = sW + 2 * rowlen;                           // [grp][npb][nz]
    double *sIstar = sLdp + (size_t)P.grp * npb * nz;         // [grp][npb]
    double *sOut = sIstar + P.grp * npb;                      // [2][grp][ng] partial contractions
    double *sScr = sOut + 2 * P.grp * ng;                     // [8][200] trapezoid scratch
    uint64_t *bar = reinterpret_cast<uint64_t *>(sScr + 8 * 200);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbins = P.nk + 1;
    if (tid == 0) {
        // which bin does this group belong to?  (binary search over the group starts)
        const int g = blockIdx.x;
        int lo = 0, hi = nbins;  // gstart[nbins] = total number of groups
        if (g >= P.gstart[nbins]) {
            s_bin = -1;
        } else {
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (P.gstart[mid] <= g) lo = mid; else hi = mid;
            }
            // skip empty bins that share the same start
            while (lo + 1 < nbins && P.gstart[lo + 1] <= g) ++lo;
            s_bin = lo;
            const int j = g - P.gstart[lo];
            const int cnt = P.offsets[lo + 1] - P.offsets[lo];
            s_first = P.offsets[lo] + j * P.grp;
            s_cnt = min(P.grp, cnt - j * P.grp);
        }
        mbar_init(bar, 1);
    }
    __syncthreads();
    const int bin = s_bin;
    if (bin < 0) return;
    const int cnt = s_cnt;
    const bool in_table = bin < P.nk;
    if (in_table && tid == 0) {
        const int ik1 = min(bin + 1, P.nk - 1);  // the reference reads weights[nk] here
        const uint32_t bytes = (uint32_t)rowlen * 8u;
        mbar_expect_tx(bar, 2u * bytes);
        tma_load_1d(sW, P.W + (size_t)bin * rowlen, bytes, bar);
        tma_load_1d(sW + rowlen, P.W + (size_t)ik1 * rowlen, bytes, bar);
    }
    if (tid < cnt) {
        const int ipv = P.perm[s_first + tid];
        s_pv[tid] = ipv;
        const double k0 = P.k[(size_t)ipv * P.kcols];
        const int ikraw = (int)floor((k0 - P.kmin) / P.dk);
        s_ak[tid] = (k0 - P.kmin - ikraw * P.dk) / P.dk;  // model_full.py:49
    }
    __syncthreads();

    // limb-darkening profile at the mu nodes (evaluate_ld, ldmodels.py:142-157) while the TMA is in flight
    const int nprof = cnt * npb * nz;
    for (int idx = tid; idx < nprof; idx += 256) {
        const int q = idx / (npb * nz), r = idx - q * npb * nz;
        const int pb = r / nz, iz = r - pb * nz;
        const int ipv = s_pv[q];
        double v;
        if (P.law == LD_PROFILES) v = P.ld[((size_t)ipv * npb + pb) * nz + iz];
        else v = ld_intensity(P.law, P.mu[iz], P.ld + ((size_t)ipv * npb + pb) * P.nld, P.nld);
        sLdp[idx] = v;
        if (P.ldp_out) P.ldp_out[((size_t)ipv * npb + pb) * nz + iz] = v;
    }
    // disk-integrated intensity (evaluate_ldi, ldmodels.py:160-175; numeric fallback): warp per (vector, pb)
    for (int r = warp; r < cnt * npb; r += 8) {
        const int q = r / npb, pb = r - q * npb;
        const int ipv = s_pv[q];
        double is;
        if (P.law == LD_PROFILES) {
            is = P.istar[(size_t)ipv * npb + pb];
        } else {
            const double *pv = P.ld + ((size_t)ipv * npb + pb) * P.nld;
            if (!ld_integral(P.law, pv, is)) is = istar_numeric_warp(lane, P.law, pv, P.nld, P.ldmu200, P.ldz200, sScr + warp * 200);
        }
        if (lane == 0) {
            sIstar[r] = is;
            if (P.istar_out) P.istar_out[(size_t)ipv * npb + pb] = is;
        }
    }
    __syncthreads();
    // isnan(ldp[ipv,0,0]) invalidates the vector (model_full.py:40)
    if (tid < cnt && isnan(sLdp[(size_t)tid * npb * nz])) P.orb[(size_t)s_pv[tid] * ORB_STRIDE + ORB_GOOD] = 0.0;

    if (in_table) mbar_wait(bar, 0);  // every thread observes the TMA completion

    const int t = tid >> 7, ig0 = tid & 127;  // thread = (table row pair member, g index)
    for (int q0 = 0; q0 < cnt; q0 += (in_table ? cnt : 1)) {
        const int nq = in_table ? cnt : 1;
        if (!in_table) {
            // direct weights for this radius ratio (calculate_weights_2d, common.py:152-185), one vector at a time
            __syncthreads();
            const double k0 = P.k[(size_t)s_pv[q0] * P.kcols];
            for (int ig = tid; ig < ng; ig += 256) weight_row(k0, P.gs[ig], P.ze, nz, sW + (size_t)ig * nz, 1);
            __syncthreads();
        }
        for (int pb = 0; pb < npb; ++pb) {
            for (int ig = ig0; ig < ng; ig += 128) {
                if (t == 0 || in_table) {
                    const double *wr = sW + (size_t)t * rowlen + (size_t)ig * nz;
                    double acc[RR_GROUP];
#pragma unroll
                    for (int q = 0; q < RR_GROUP; ++q) acc[q] = 0.0;
                    int iz = ig % nz;  // rotated start: the lanes of a warp hit distinct banks
                    for (int j = 0; j < nz; ++j) {
                        const double wv = wr[iz];
#pragma unroll
                        for (int q = 0; q < RR_GROUP; ++q)
                            if (q < nq) acc[q] = fma(wv, sLdp[((size_t)(q0 + q) * npb + pb) * nz + iz], acc[q]);
                        iz = (iz + 1 == nz) ? 0 : iz + 1;
                    }
#pragma unroll
                    for (int q = 0; q < RR_GROUP; ++q)
                        if (q < nq) sOut[((size_t)t * P.grp + q) * ng + ig] = acc[q];
                }
            }
            __syncthreads();
            for (int idx = tid; idx < nq * ng; idx += 256) {
                const int q = idx / ng, ig = idx - q * ng;
                const int ipv = s_pv[q0 + q];
                double v = sOut[(size_t)q * ng + ig];
                if (in_table) {
                    const double ak = s_ak[q0 + q];
                    v = (1.0 - ak) * v + ak * sOut[((size_t)P.grp + q) * ng + ig];
                }
                P.ldrec[((size_t)ipv * npb + pb) * P.lds + ig] = v;
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < cnt * npb; r += 256) {
        const int q = r / npb, pb = r - q * npb;
        const int ipv = s_pv[q];
        const double kk = P.k[(size_t)ipv * P.kcols + (P.kcols == npb ? pb : 0)];
        double *tail = P.ldrec + ((size_t)ipv * npb + pb) * P.lds + ng;
        tail[0] = kk;
        tail[1] = 1.0 / (1.0 + kk);
        tail[2] = 1.0 / sIstar[r];
        tail[3] = kk * kk;
        for (int j = ng + 4; j < P.lds; ++j) tail[j - ng] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// The npv x npt pass.
//
// One CTA = one parameter vector x one chunk of the time axis, 8 warps; the time axis is cut into
// blocks of 64 consecutive points and each warp walks its own blocks.
//
//  * Block classification (warp-uniform, ~10 fp64 ops per 64 points): set_data stores the smallest and
//    largest time stamp of every block.  A block whose [tmin, tmax] misses every transit window
//    [t0 + n p + lo, t0 + n p + hi] holds no in-box point, so its 64 fluxes are 1.0: the warp issues
//    one 16-byte streaming store per lane and moves on (likelihood mode: adds the block's
//    pre-summed (obs-1)^2).  ~93 % of a TESS sector takes this path at HBM-store speed.
//  * Other blocks: each point is folded and box-tested in fp64 in the reference's operation order
//    (model_full.py:88-91); in-box points go to a warp-private queue in shared memory.
//  * Batched drain: when >= 128 exposure sub-samples are queued the warp evaluates them together --
//    separation + limb-darkening lerp for every sample, then the samples on the stellar limb (the
//    ones that need the sqrt + 2 atan2 lens area, ~1/4 of them) are compacted a second time so the
//    expensive code runs with full warps; per-point sums over sub-samples are taken in exposure order.
//
// The per-vector record (ldm rows + k, 1/(1+k), 1/I*, k^2 per passband) is staged into shared memory
// with one TMA bulk copy per CTA.
// ---------------------------------------------------------------------------------------------
struct PointsParams {
    const double *time;
    const int32_t *lcids, *pbids, *epids, *nsamples;
    const double *exptimes;
    const double *orb, *t0, *ldrec;
    double *flux;
    const double *obs;
    const int32_t *blk;
    const double *isig2;  // [npv][nblocks]
    double *partial;      // [npv][nchunks]
    const double *bmin, *bmax;  // per 64-point block: smallest / largest time stamp
    const int32_t *blc;         // light curve of the block, -1 when it straddles light curves
    const double *bchi;         // likelihood: sum of (obs-1)^2 over the block's points
    const int32_t *bnoise;      // likelihood: noise id of the block, -1 none, -2 mixed (slow path)
    long long npt;
    int npv, nlc, npb, nep, ng, lds, nblocks, ns_max, nchunks, blocks_per_chunk, nblk64, stage_ld;
    double dg, inv_dg;
};

#ifndef PT_MINB
#define PT_MINB 3
#endif
constexpr int PT_THREADS = 256;
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_BLOCK = 64;           // points per classification block
constexpr int PT_BATCH = 128;          // exposure sub-samples evaluated per drain
constexpr int PT_QCAP = PT_BATCH + PT_BLOCK;
constexpr int PT_MAXBLK = 8192;        // blocks per CTA chunk (hit bitmap in shared memory)
constexpr double PT_EPS = 1e-9;        // classification margin, in periods (>> rounding, << the 0.003 d pad)

struct alignas(16) WarpScratch {
    double q_tc[PT_QCAP];
    double contrib[PT_BATCH];
    double l_z[PT_BATCH], l_ip[PT_BATCH];
    int q_ipt[PT_QCAP];
    int l_it[PT_BATCH];
};

template <int VEC>
struct VecIO;
template <>
struct VecIO<1> {
    __device__ static __forceinline__ void load(const double *p, double *v) { v[0] = __ldg(p); }
    __device__ static __forceinline__ void store(double *p, const double *v) { __stcs(p, v[0]); }
};
template <>
struct VecIO<2> {
    __device__ static __forceinline__ void load(const double *p, double *v) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        v[0] = t.x;
        v[1] = t.y;
    }
    __device__ static __forceinline__ void store(double *p, const double *v) {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
    }
};

// Batched evaluation of `np` queued in-box points starting at queue slot `base` (warp-cooperative;
// model_full.py:93-99).  Inlined at its single call site (an out-of-line call was measured 15 % slower:
// the shared-memory address space of the staged record is lost across the call).  Returns the warp
// lane's chi^2 increment (likelihood mode).
struct DrainCtx {
    const PointsParams *P;
    WarpScratch *ws;
    const double *orb;
    const double *ld;  // per-vector record in shared memory: [npb][lds]
    double *frow;
    const double *isig2;
    int lane, S;
};

template <bool SINGLE_LC, bool LNL>
__device__ __forceinline__ double drain_batch(const DrainCtx &c, int base, int np) {
    const PointsParams &P = *c.P;
    WarpScratch &ws = *c.ws;
    const int lane = c.lane, S = c.S, ng = P.ng, lds = P.lds;
    const double dg = P.dg, inv_dg = P.inv_dg;
    const unsigned lt_mask = (1u << lane) - 1u;
    double chi = 0.0;
    double cx[5], cy[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { cx[j] = c.orb[j]; cy[j] = c.orb[5 + j]; }
    // single light curve: the per-light-curve metadata are block constants
    const int ns1 = SINGLE_LC ? P.nsamples[0] : 1;
    const double et1 = SINGLE_LC ? P.exptimes[0] : 0.0;
    const double *row1 = c.ld + (SINGLE_LC ? (size_t)P.pbids[0] * lds : 0);
    double bigsum = 0.0;  // S > PT_BATCH only: running sum over sample chunks (np == 1)
    for (int s0 = 0; s0 < S; s0 += PT_BATCH) {
        const int SS = min(S - s0, PT_BATCH);  // sample slots in this pass
        const int nitems = np * SS;
        // stage A: separation, limb-darkening lerp, cheap area cases; limb samples -> second queue
        int nl = 0;
        for (int it0 = 0; it0 < nitems; it0 += 32) {
            const int it = it0 + lane;
            bool limb = false;
            double z = 0.0, ip = 0.0;
            if (it < nitems) {
                const int pt = (SS == 1) ? it : it / SS;
                const int s = s0 + (it - pt * SS);
                int ns = ns1;
                double et = et1;
                const double *row = row1;
                if (!SINGLE_LC) {
                    const int lc = P.lcids[ws.q_ipt[base + pt]];
                    ns = P.nsamples[lc];
                    et = P.exptimes[lc];
                    row = c.ld + (size_t)P.pbids[lc] * lds;
                }
                double cc = 0.0;
                if (s < ns) {
                    // exposure offset exptime*((s+1-0.5)/ns - 0.5) (model_full.py:94); exactly 0 for ns == 1
                    const double off = (ns == 1) ? 0.0 : et * (((s + 1) - 0.5) / ns - 0.5);
                    z = sep_poly(ws.q_tc[base + pt] + off, cx, cy);
                    const double k = row[ng];
                    ip = ldm_lerp(z * row[ng + 1], dg, inv_dg, row, ng);
                    if (1.0 + k <= z) cc = 1.0;                                   // no overlap: area 0
                    else if (fabs(1.0 - k) < z) limb = true;                      // lens: kite formula
                    else if (z <= 1.0 - k) cc = 1.0 - ip * (kPi * row[ng + 3]) * row[ng + 2];
                    else if (z <= k - 1.0) cc = 1.0 - ip * kPi * row[ng + 2];    // planet covers the star
                    else cc = nan("");
                }
                ws.contrib[it] = cc;
            }
            const unsigned m = __ballot_sync(0xffffffffu, limb);
            if (limb) {
                const int pos = nl + __popc(m & lt_mask);
                ws.l_it[pos] = it;
                ws.l_z[pos] = z;
                ws.l_ip[pos] = ip;
            }
            nl += __popc(m);
        }
        __syncwarp();
        // stage B: lens area on the limb (sqrt + 2 atan2), full warps
        for (int j = lane; j < nl; j += 32) {
            const int it = ws.l_it[j];
            const double *row = row1;
            if (!SINGLE_LC) {
                const int pt = (SS == 1) ? it : it / SS;
                row = c.ld + (size_t)P.pbids[P.lcids[ws.q_ipt[base + pt]]] * lds;
            }
            double area, kap;
            kite_area(row[ng], row[ng + 3], ws.l_z[j], area, kap);
            ws.contrib[it] = 1.0 - ws.l_ip[j] * area * row[ng + 2];
        }
        __syncwarp();
        // stage C: per-point sum over sub-samples in exposure order (model_full.py:93-99)
        for (int pt = lane; pt < np; pt += 32) {
            const int ipt = ws.q_ipt[base + pt];
            const int ns = SINGLE_LC ? ns1 : P.nsamples[P.lcids[ipt]];
            const int m = min(SS, ns - s0);
            double sum = bigsum;
            for (int j = 0; j < m; ++j) sum += ws.contrib[pt * SS + j];
            if (s0 + SS >= S) {
                const double f = sum / ns;
                if (LNL) {
                    const int b = P.blk ? P.blk[ipt] : 0;
                    if (b >= 0) {
                        const double d = P.obs[ipt] - f;
                        chi = fma(d * d, c.isig2[b], chi);
                    }
                } else {
                    c.frow[ipt] = f;
                }
            } else {
                bigsum = sum;  // only reached with np == 1 (lane 0)
            }
        }
        __syncwarp();
    }
    return chi;
}

template <int VEC, bool SINGLE_LC, bool LNL>
__global__ void __launch_bounds__(PT_THREADS, PT_MINB) k_rr_points(const __grid_constant__ PointsParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_red[PT_WARPS];
    __shared__ unsigned s_hit[PT_MAXBLK / 32];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ipv = blockIdx.x / P.nchunks;
    const int chunk = blockIdx.x - ipv * P.nchunks;
    const long long npt = P.npt;

    // dynamic smem: [WarpScratch x 8] [ldrec copy npb*lds] [lc_lo nlc] [lc_hi nlc] [lc_t0 nlc]
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(smem_raw)[warp];
    double *sLd = reinterpret_cast<double *>(smem_raw + sizeof(WarpScratch) * PT_WARPS);
    double *sLo = sLd + (size_t)P.npb * P.lds;
    double *sHi = sLo + (SINGLE_LC ? 0 : P.nlc);
    double *sT0 = sHi + (SINGLE_LC ? 0 : P.nlc);

    const double *orb = P.orb + (size_t)ipv * ORB_STRIDE;
    const bool good = orb[ORB_GOOD] != 0.0;
    const double *ldg = P.ldrec + (size_t)ipv * P.npb * P.lds;

    const int bbeg = chunk * P.blocks_per_chunk;
    const int bend = min(P.nblk64, bbeg + P.blocks_per_chunk);

    if (!good) {  // invalid parameter vector: NaN row (model_full.py:80-82)
        if (LNL) {
            if (tid == 0) P.partial[(size_t)ipv * P.nchunks + chunk] = nan("");
        } else {
            double *frow = P.flux + (size_t)ipv * npt;
            double v[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = nan("");
            const long long cend = min(npt, (long long)bend * PT_BLOCK);
            for (long long i = (long long)bbeg * PT_BLOCK + (long long)tid * VEC; i < cend; i += PT_THREADS * VEC)
                VecIO<VEC>::store(frow + i, v);
        }
        return;
    }

    if (tid == 0) {
        mbar_init(&bar, 1);
        const uint32_t bytes = (uint32_t)(P.npb * P.lds * 8);
        mbar_expect_tx(&bar, bytes);
        tma_load_1d(sLd, ldg, bytes, &bar);
    }
    const double p = orb[ORB_P], invp = orb[ORB_INVP], T1 = orb[ORB_T1], T4 = orb[ORB_T4];
    double lo1 = 0, hi1 = 0, t01 = 0;
    if (SINGLE_LC) {
        const double pad = 0.003 + P.exptimes[0];
        lo1 = T1 - pad;
        hi1 = T4 + pad;
        t01 = P.t0[(size_t)ipv * P.nep + P.epids[0]];
    } else {
        for (int lc = tid; lc < P.nlc; lc += PT_THREADS) {
            const double pad = 0.003 + P.exptimes[lc];
            sLo[lc] = T1 - pad;
            sHi[lc] = T4 + pad;
            sT0[lc] = P.t0[(size_t)ipv * P.nep + P.epids[lc]];
        }
    }
    __syncthreads();  // mbarrier init + per-light-curve tables visible

    bool ld_ready = false;
    DrainCtx dctx;

    const int S = P.ns_max;                        // sub-sample slots per queued point
    const int PB = S >= PT_BATCH ? 1 : PT_BATCH / S;  // points per drain batch
    int qn = 0;
    double chi = 0.0;
    const double *isig2 = LNL ? P.isig2 + (size_t)ipv * P.nblocks : nullptr;
    double *frow = LNL ? nullptr : P.flux + (size_t)ipv * npt;
    const unsigned lt_mask = (1u << lane) - 1u;
    dctx.P = &P; dctx.ws = &ws; dctx.orb = orb; dctx.ld = sLd; dctx.frow = frow; dctx.isig2 = isig2;
    dctx.lane = lane; dctx.S = S;

    // ---- classification of every block of this chunk, in parallel over the CTA -----------------------
    // Can any transit window [t0 + n p + lo, t0 + n p + hi] touch [tmin, tmax] of the block?  If not, the
    // block holds no in-box point.  One bit per block goes to shared memory; in likelihood mode the
    // pre-summed (obs-1)^2 of the untouched blocks is added right here.
    const int nbc = bend - bbeg;
    for (int bb0 = 0; bb0 < nbc; bb0 += PT_THREADS) {
        const int bb = bb0 + tid, b = bbeg + bb;
        bool hit = false;
        if (bb < nbc) {
            hit = true;
            const int lcb = SINGLE_LC ? 0 : P.blc[b];
            int nz_id = 0;
            if (LNL) nz_id = P.bnoise ? P.bnoise[b] : 0;
            const bool partial = (b == P.nblk64 - 1) && (npt % PT_BLOCK != 0);
            if (lcb >= 0 && nz_id != -2 && !partial) {
                const double lo = SINGLE_LC ? lo1 : sLo[lcb], hi = SINGLE_LC ? hi1 : sHi[lcb];
                const double t0 = SINGLE_LC ? t01 : sT0[lcb];
                const double n1 = ceil(fma(P.bmin[b] - t0 - hi, invp, -PT_EPS));
                const double n2 = floor(fma(P.bmax[b] - t0 - lo, invp, PT_EPS));
                hit = !(n1 > n2) || !(p > 0.0);  // NaNs and p <= 0 fall through to the exact per-point path
            }
            if (LNL && !hit && nz_id >= 0) chi = fma(P.bchi[b], isig2[nz_id], chi);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_hit[bb >> 5] = m;
    }
    __syncthreads();

    // ---- main loop: a warp takes groups of 8 blocks (512 points) ------------------------------------
    // One extra pass after the last group flushes the queue, so that drain() has exactly ONE call site
    // and batches are always cut from the top of the queue: a point's arithmetic does not depend on the
    // chunking and results are bit-reproducible across population splits.
    const int ngroups = (nbc + 7) >> 3;
    for (int g = warp;; g += PT_WARPS) {
        const bool live = g < ngroups;
        unsigned bits = live ? (s_hit[g >> 2] >> ((g & 3) * 8)) & 0xffu : 0u;
        const int b0 = bbeg + g * 8;
        if (!LNL && live) {
            // untouched blocks: 64 fluxes of exactly 1.0, one 16-byte streaming store per lane
            double one[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) one[j] = 1.0;
            double *fb = frow + (long long)b0 * PT_BLOCK + lane * VEC;
            if (bits == 0u && b0 + 8 <= bend) {  // the common case: eight untouched blocks, no predicates
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int h = 0; h < 2 / VEC; ++h) VecIO<VEC>::store(fb + j * PT_BLOCK + h * 32 * VEC, one);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!((bits >> j) & 1u) && b0 + j < bend) {
#pragma unroll
                        for (int h = 0; h < 2 / VEC; ++h) VecIO<VEC>::store(fb + j * PT_BLOCK + h * 32 * VEC, one);
                    }
                }
            }
        }
        // touched blocks: exact per-point path; the body runs once more on the flush pass (no block)
        do {
            if (bits) {
                const int jb = __ffs(bits) - 1;
                bits &= bits - 1;
                const long long base = (long long)(b0 + jb) * PT_BLOCK;
#pragma unroll
                for (int h = 0; h < 2 / VEC; ++h) {
                    const long long i0 = base + (long long)(h * 32 + lane) * VEC;
                    double tv[VEC], fv[VEC];
                    const bool inr = i0 < npt;  // VEC == 2 requires an even npt: vectors are all-in or all-out
                    if (inr) VecIO<VEC>::load(P.time + i0, tv);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        bool inbox = false;
                        double tc = 0.0;
                        if (inr) {
                            double lo = lo1, hi = hi1, t0 = t01;
                            if (!SINGLE_LC) {
                                const int lc = P.lcids[i0 + j];
                                lo = sLo[lc];
                                hi = sHi[lc];
                                t0 = sT0[lc];
                            }
                            // epoch = floor((t - t0 + p/2)/p); tc = t - (t0 + epoch p)  (model_full.py:88-89)
                            // The division is a multiplication by 1/p: the two can only disagree half a
                            // period away from the transit, where the point is outside the box either way.
                            const double epoch = floor(fma(tv[j] - t0, invp, 0.5));
                            tc = tv[j] - __dadd_rn(t0, __dmul_rn(epoch, p));
                            inbox = (lo <= tc) && (tc <= hi);
                            fv[j] = 1.0;
                            if (LNL && !inbox) {
                                const int nb = P.blk ? P.blk[i0 + j] : 0;
                                if (nb >= 0) {
                                    const double d = P.obs[i0 + j] - 1.0;
                                    chi = fma(d * d, isig2[nb], chi);
                                }
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, inbox);
                        if (inbox) {
                            const int pos = qn + __popc(m & lt_mask);
                            ws.q_ipt[pos] = (int)(i0 + j);
                            ws.q_tc[pos] = tc;
                        }
                        qn += __popc(m);
                    }
                    if (!LNL && inr) VecIO<VEC>::store(frow + i0, fv);
                }
                __syncwarp();
            }
            // drain full batches; the remainder (< PB points) waits for more, except on the flush pass
            while (qn >= PB || (!live && qn > 0)) {
                const int n = min(qn, PB);
                qn -= n;
                if (!ld_ready) {
                    mbar_wait(&bar, 0);
                    ld_ready = true;
                }
                chi += drain_batch<SINGLE_LC, LNL>(dctx, qn, n);
            }
        } while (bits);
        if (!live) break;
    }
    if (!ld_ready) mbar_wait(&bar, 0);  // never exit with the bulk copy in flight

    if (LNL) {
        chi = warp_sum(chi);
        if (lane == 0) s_red[warp] = chi;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int wq = 0; wq < PT_WARPS; ++wq) s += s_red[wq];
            P.partial[(size_t)ipv * P.nchunks + chunk] = s;
        }
    }
}

// chi^2 partials -> lnL (wnloglikelihood.py:34): lnL = sum_b n_b (-log s_b - log(2 pi)/2) - chi^2/2
__global__ void k_lnl_finish(const double *__restrict__ partial, int nchunks, const double *__restrict__ sigma,
                             const double *__restrict__ nblk, int nblocks, int npv, double *__restrict__ lnl) {
    const int ipv = blockIdx.x * blockDim.x + threadIdx.x;
    if (ipv >= npv) return;
    double chi = 0.0;
    for (int c = 0; c < nchunks; ++c) chi += partial[(size_t)ipv * nchunks + c];
    double cst = 0.0;
    for (int b = 0; b < nblocks; ++b) cst += nblk[b] * (-log(sigma[(size_t)ipv * nblocks + b]) - 0.5 * log(kTwoPi));
    lnl[ipv] = cst - 0.5 * chi;
}

__global__ void k_inv_sigma2(const double *__restrict__ sigma, long long n, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double s = sigma[i];
        out[i] = 1.0 / (s * s);
    }
}

// lnlike_normal on a materialised model (wnloglikelihood.py:22-35): one CTA per vector.
__global__ void __launch_bounds__(256) k_lnl_model(const double *__restrict__ model, const double *__restrict__ obs,
                                                   const int32_t *__restrict__ blk, const double *__restrict__ isig2,
                                                   long long npt, int nblocks, double *__restrict__ partial) {
    __shared__ double s_w[8];
    const int ipv = blockIdx.x;
    const double *m = model + (size_t)ipv * npt;
    const double *w = isig2 + (size_t)ipv * nblocks;
    double chi = 0.0;
    for (long long j = threadIdx.x; j < npt; j += 256) {
        const int b = blk ? blk[j] : 0;
        if (b >= 0) {
            const double d = obs[j] - m[j];
            chi = fma(d * d, w[b], chi);
        }
    }
    chi = warp_sum(chi);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = chi;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += s_w[i];
        partial[ipv] = s;
    }
}

}  // namespace ptb
