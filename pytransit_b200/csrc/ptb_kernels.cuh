// ptb_kernels.cuh -- sm_100a kernels of the RoadRunner population path.
//
//   k_weight_table   W[nk,ng,nz]  (common.py:188-223)                         once per model
//   k_rr_orbit       per-vector Taylor orbit coefficients + contact times       \
//   k_bin_sort       counting sort of the vectors by weight-table row            > once per evaluate
//   k_rr_ldm         LD profile, I*, LD means (TMA-staged table rows shared      /  (model_full.py:39-70)
//                    by the vectors of a group)
//   k_rr_points      the npv x npt pass: phase fold, box test, supersampled flux, optional fused
//                    chi^2 reduction (model_full.py:76-99, wnloglikelihood.py:22-35)
//   k_lnl_finish     chi^2 partials -> lnL[npv]
//   k_lnl_model      lnlike_normal on a materialised model flux
//
// HBM layout (all fp64 unless noted):
//   rec  [npv][recstride]     one record per parameter vector, fetched by ONE TMA bulk copy:
//                               orb[16]         cx[5] cy[5] p 1/p T1 T4 good -
//                               t0[nep]         transit centres (padded to an even count)
//                               ld[npb][lds]    ldm[ng] | k 1/(1+k) 1/I* k^2 | pad   (lds = ng+4, even)
//   flux [npv][npt]           row-major, written with 16-byte stores
#pragma once
#include <cooperative_groups.h>

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
// Per-vector setup (model_full.py:39-70), three launches on two streams:
//
//   k_rr_orbit    (side stream) 8 lanes per vector: validity, the 7 Kepler solves of the Taylor stencil
//                 (one per lane), coefficients from sub-warp shuffles, T1/T4 bisection on two lanes;
//                 copies the transit centres into the record.
//   k_bin_sort    one CTA: weight-table row `ik` of every vector, counting sort of the vectors by row
//                 (shared-memory histogram, scan, scatter) and one descriptor per group of up to
//                 RR_GROUP vectors that share a row pair.
//   k_rr_ldm      one CTA per group: the two rows W[ik], W[ik+1] (ng*nz*8 bytes each) arrive in shared
//                 memory through ONE pair of TMA bulk copies per group instead of one per vector,
//                 overlapped with the limb-darkening profile evaluation; then the (ng x nz).(nz)
//                 contractions, register-blocked over the group's vectors and both table rows.
// Sorting by table row cuts the L2 -> SM traffic of the contraction by the group size; the orbit solve
// (latency-bound fp64 chains) runs concurrently with the table staging and the contraction.
// ---------------------------------------------------------------------------------------------
constexpr int RR_GROUP = 8;  // vectors per k_rr_ldm CTA (register blocking factor)
constexpr int ORB_LDNAN = 15;  // record slot: 1.0 when the limb-darkening profile is NaN (model_full.py:40)

template <int WIDTH>
__device__ __forceinline__ void solve_orbit_lanes(int sl, bool valid, double p, double a, double inc, double e, double w,
                                                  double kbox, const double *xyc_in, double *orb_out, double t_centre = 0.0) {
    double cx[5], cy[5];
    if (xyc_in != nullptr) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { cx[j] = xyc_in[j]; cy[j] = xyc_in[5 + j]; }
    } else {
        double x = 0.0, y = 0.0;
        if (valid && sl < 7) {
            const double offset = mean_anomaly_offset(e, w);
            sky_position(t_centre + (sl - 3) * 2e-2, p, a * (1.0 - e * e), cos(inc), e, w, offset, x, y);
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
    // ascending order: around a secondary eclipse the x velocity is negative and the two searches swap roles
    const double ta = __shfl_sync(0xffffffffu, tcon, 0, WIDTH);
    const double tb = __shfl_sync(0xffffffffu, tcon, 1, WIDTH);
    const double t1 = (tb < ta) ? tb : ta, t4 = (tb < ta) ? ta : tb;
    if (valid && sl == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { orb_out[j] = cx[j]; orb_out[5 + j] = cy[j]; }
        orb_out[ORB_P] = p;
        orb_out[ORB_INVP] = 1.0 / p;
        orb_out[ORB_T1] = t1;
        orb_out[ORB_T4] = t4;
        orb_out[ORB_GOOD] = 1.0;
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
    const double *t0;      // [npv][nep]
    double *rec;           // per-vector records (orb at offset 0, t0 at ORB_STRIDE)
    int npv, kcols, nep, recstride;
    int eclipse;           // secondary-eclipse geometry (model_eclipse.py:38-44): expansion about mid-eclipse
    double rstar;          // stellar radius [R_sun] for the light-travel-time shift
    const double *rstar_v; // per-vector stellar radii (eclipse spectroscopy, model_ecspec.py:41); overrides rstar
};

// eclipse_time_offset: time from mid-transit to mid-eclipse (the reference's eclipse_phase, orbits_py.py:544-555)
__device__ __forceinline__ double eclipse_time_offset(double p, double e, double w) {
    double s, c;
    sincos(kHalfPi - w, &s, &c);
    const double q = sqrt(1.0 - e * e);
    const double etr = atan2(q * s, e + c);
    sincos(kHalfPi + kPi - w, &s, &c);
    const double eec = atan2(q * s, e + c);
    const double mtr = etr - e * sin(etr), mec = eec - e * sin(eec);
    const double phase = (mec - mtr) * p / kTwoPi;
    return phase > 0.0 ? phase : p + phase;
}

// eclipse_light_travel_time [d]: light crossing the line-of-sight distance between the mid-transit and
// mid-eclipse positions, (r_tr + r_ec) sin i stellar radii (meepmeep function, restated; R_sun as orbits_py.py:46)
__device__ __forceinline__ double eclipse_light_travel_time(double a, double inc, double e, double w, double rstar) {
    const double ae = a * (1.0 - e * e), sw = sin(w);
    return (ae / (1.0 + e * sw) + ae / (1.0 - e * sw)) * sin(inc) * rstar * (0.5 * 1.392684e9) / 299792458.0 / 86400.0;
}

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
    const bool good0 = inr && !(isnan(a) || (a <= 1.0) || (e < 0.0));  // model_full.py:40 (ldp checked in k_rr_ldm)
    double *orb = P.rec + (size_t)(inr ? ipv : 0) * P.recstride;
    double shift = 0.0, tadd = 0.0;
    if (P.eclipse && good0) {
        shift = eclipse_time_offset(p, e, w);
        tadd = shift + eclipse_light_travel_time(a, inc, e, w, P.rstar_v ? P.rstar_v[ipv] : P.rstar);  // te = t0 + shift + ltt (model_eclipse.py:71)
    }
    solve_orbit_lanes<8>(sl, good0, p, a, inc, e, w, k0, (P.xyc_in && inr) ? P.xyc_in + (size_t)ipv * 10 : nullptr, orb, shift);
    if (inr) {  // transit centres travel with the record (one TMA bulk copy per vector in k_rr_points)
        for (int j = sl; j < P.nep; j += 8) orb[ORB_STRIDE + j] = P.eclipse ? P.t0[(size_t)ipv * P.nep + j] + tadd : P.t0[(size_t)ipv * P.nep + j];
        if (sl == 0 && (P.nep & 1)) orb[ORB_STRIDE + P.nep] = 0.0;
        if (sl == 0 && !good0) {
            for (int j = 0; j < ORB_LDNAN; ++j) orb[j] = (j == ORB_GOOD) ? 0.0 : nan("");
        }
    }
}

// Counting sort of the vectors by weight-table row + group descriptors: ONE thread-block cluster of
// SORT_CTAS CTAs.  Every CTA histograms its share of the vectors in its own shared memory; after a
// cluster barrier each CTA reads its peers' histograms through distributed shared memory to get the
// bin totals and its own start slot inside every bin, scans the totals, and scatters its vectors.
//   bin = table row ik, nk = direct weights (k outside the table), nk+1 = invalid vector (never grouped)
//   perm[npv]   vectors ordered by bin
//   gdesc[g]    (bin, first slot in perm, count, -) for every group; *ngroups = number of groups
struct SortParams {
    const double *k, *a, *e;
    int *perm;
    int4 *gdesc;
    int *ngroups;
    int npv, kcols, nk, grp;
    double kmin, kmax, dk;
};

constexpr int SORT_THREADS = 1024;
constexpr int SORT_CTAS = 8;        // cluster size (portable maximum)
constexpr int SORT_CACHE = 4;       // bins kept in registers per thread (covers npv <= 32768 without recomputation)
constexpr int SORT_MAXBINS = 1024;  // nk + 2 <= SORT_MAXBINS

__device__ __forceinline__ int table_bin(const SortParams &P, int ipv) {
    const double a = P.a[ipv], e = P.e[ipv], k0 = P.k[(size_t)ipv * P.kcols];
    if (isnan(a) || (a <= 1.0) || (e < 0.0)) return P.nk + 1;
    if ((P.kmin <= k0) && (k0 <= P.kmax)) return min((int)floor((k0 - P.kmin) / P.dk), P.nk - 1);
    return P.nk;
}

__global__ void __cluster_dims__(SORT_CTAS, 1, 1) __launch_bounds__(SORT_THREADS) k_bin_sort(const __grid_constant__ SortParams P) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int s_cnt[SORT_MAXBINS];      // this CTA's histogram (read by the peers)
    __shared__ int s_tot[SORT_MAXBINS];      // bin totals over the cluster
    __shared__ int s_base[SORT_MAXBINS];     // first slot of this CTA's vectors inside each bin
    __shared__ int s_off[SORT_MAXBINS + 1], s_gst[SORT_MAXBINS + 1], s_cur[SORT_MAXBINS];
    __shared__ int s_wsum[32], s_wsum_g[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int gtid = rank * SORT_THREADS + tid, gthreads = SORT_CTAS * SORT_THREADS;
    const int nb = P.nk + 1;  // bins that carry work; the invalid bin nk+1 is listed last and never grouped
    for (int i = tid; i < nb + 1; i += SORT_THREADS) { s_cnt[i] = 0; s_cur[i] = 0; }
    __syncthreads();
    int cache[SORT_CACHE];
#pragma unroll
    for (int u = 0; u < SORT_CACHE; ++u) {
        const int i = gtid + u * gthreads;
        cache[u] = (i < P.npv) ? table_bin(P, i) : -1;
    }
#pragma unroll
    for (int u = 0; u < SORT_CACHE; ++u)
        if (cache[u] >= 0) atomicAdd(&s_cnt[cache[u]], 1);
    for (int i = gtid + SORT_CACHE * gthreads; i < P.npv; i += gthreads) atomicAdd(&s_cnt[table_bin(P, i)], 1);
    cluster.sync();

    // bin totals and this CTA's start inside each bin (peers' histograms through DSMEM)
    int c = 0, mine = 0;
    if (tid < nb + 1) {
#pragma unroll
        for (int r = 0; r < SORT_CTAS; ++r) {
            const int v = cluster.map_shared_rank(s_cnt, r)[tid];
            if (r < rank) mine += v;
            c += v;
        }
        s_tot[tid] = c;
    }
    // exclusive scans of the totals (slots) and of the group counts; nb + 1 <= 1024 entries, one per thread
    {
        const int g = (tid < nb) ? (c + P.grp - 1) / P.grp : 0;
        int ic = c, ig = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, ic, o), u = __shfl_up_sync(0xffffffffu, ig, o);
            if (lane >= o) { ic += t; ig += u; }
        }
        if (lane == 31) { s_wsum[warp] = ic; s_wsum_g[warp] = ig; }
        __syncthreads();
        if (warp == 0) {
            int wc = s_wsum[lane], wg = s_wsum_g[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wc, o), u = __shfl_up_sync(0xffffffffu, wg, o);
                if (lane >= o) { wc += t; wg += u; }
            }
            s_wsum[lane] = wc;
            s_wsum_g[lane] = wg;
        }
        __syncthreads();
        const int basec = warp ? s_wsum[warp - 1] : 0, baseg = warp ? s_wsum_g[warp - 1] : 0;
        if (tid < nb + 1) {
            s_off[tid] = basec + ic - c;
            s_base[tid] = basec + ic - c + mine;
            s_gst[tid] = baseg + ig - g;
        }
        if (tid == nb && rank == 0) *P.ngroups = baseg + ig - g;  // groups of the bins 0..nb-1
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SORT_CACHE; ++u)
        if (cache[u] >= 0) P.perm[s_base[cache[u]] + atomicAdd(&s_cur[cache[u]], 1)] = gtid + u * gthreads;
    for (int i = gtid + SORT_CACHE * gthreads; i < P.npv; i += gthreads) {
        const int b = table_bin(P, i);
        P.perm[s_base[b] + atomicAdd(&s_cur[b], 1)] = i;
    }
    const int ngroups = s_gst[nb];
    for (int g = gtid; g < ngroups; g += gthreads) {
        int lo = 0, hi = nb;  // last bin with s_gst[bin] <= g (an empty bin shares its successor's start, so
        while (hi - lo > 1) {  //  the search lands on the non-empty one that owns the group)
            const int mid = (lo + hi) >> 1;
            if (s_gst[mid] <= g) lo = mid; else hi = mid;
        }
        const int j = g - s_gst[lo];
        P.gdesc[g] = make_int4(lo, s_off[lo] + j * P.grp, min(P.grp, s_tot[lo] - j * P.grp), 0);
    }
    cluster.sync();  // no CTA leaves while a peer may still read its histogram
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

// fp64 tensor-core MMA (DMMA): D[8x8] += A[8x4] . B[4x8]; lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)+{0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct LdmParams {
    const double *k;      // [npv][kcols]
    const double *ld;     // ldc[npv][npb][nld] or ldp[npv][npb][nz]
    const double *istar;  // [npv][npb] (profiles only)
    const double *W, *ze, *mu, *gs, *ldmu200, *ldz200;
    const int *perm, *ngroups;
    const int4 *gdesc;
    double *rec, *ldp_out, *istar_out;
    int npv, kcols, npb, nld, law, nk, ng, nz, lds, grp;  // grp <= RR_GROUP vectors per CTA
    int recstride, rec_ld;                                // record stride / offset of the ld rows (doubles)
    int numeric_istar;                                    // the law has no analytic disk integral
    double kmin, dk;
};

__global__ void __launch_bounds__(256) k_rr_ldm(const __grid_constant__ LdmParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_pv[RR_GROUP];
    __shared__ double s_ak[RR_GROUP];
    __shared__ __align__(8) uint64_t bar;
    const int ng = P.ng, nz = P.nz, npb = P.npb;
    const int rowlen = ng * nz;
    double *sW = reinterpret_cast<double *>(smem_raw);       // [2][ng][nz]
    double *sLdp = sW + 2 * rowlen;                           // [grp][npb][nz]
    double *sIstar = sLdp + (size_t)P.grp * npb * nz;         // [grp][npb]
    double *sScr = sIstar + P.grp * npb;                      // [8][200] trapezoid scratch (numeric I* only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int)blockIdx.x >= *P.ngroups) return;
    const int4 gd = P.gdesc[blockIdx.x];
    const int bin = gd.x, first = gd.y, cnt = gd.z;
    const bool in_table = bin < P.nk;
    if (tid == 0) {
        mbar_init(&bar, 1);
        if (in_table) {
            const int ik1 = min(bin + 1, P.nk - 1);  // the reference reads weights[nk] here
            const uint32_t bytes = (uint32_t)rowlen * 8u;
            mbar_expect_tx(&bar, 2u * bytes);
            tma_load_1d(sW, P.W + (size_t)bin * rowlen, bytes, &bar);
            tma_load_1d(sW + rowlen, P.W + (size_t)ik1 * rowlen, bytes, &bar);
        }
    }
    if (tid < cnt) {
        const int ipv = P.perm[first + tid];
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
    // isnan(ldp[ipv,0,0]) invalidates the vector (model_full.py:40); k_rr_orbit owns ORB_GOOD, this flag is ours
    if (tid < cnt) P.rec[(size_t)s_pv[tid] * P.recstride + ORB_LDNAN] = isnan(sLdp[(size_t)tid * npb * nz]) ? 1.0 : 0.0;

    // Contraction ldm[q, ig] = sum_iz ldp[q, iz] W[ig, iz] on the fp64 tensor-core path: the group's (up to)
    // 8 vectors are the M = 8 rows of mma.sync.m8n8k4, a tile of 8 g-nodes the N = 8 columns, the mu nodes K.
    // A warp takes (passband, g-tile) jobs and accumulates both table rows of the pair.
    const int fr = lane >> 2, fc = lane & 3;  // fragment row (vector / g-node) and k index of this lane
    if (in_table) {
        mbar_wait(&bar, 0);  // every thread observes the TMA completion
        const int ksteps = (nz + 3) >> 2, ntiles = (ng + 7) >> 3;
        const int qa = min(fr, P.grp - 1);  // rows beyond the group alias the last one: computed, never stored
        for (int job = warp; job < npb * ntiles; job += 8) {
            const int pb = job / ntiles, ig0 = (job - pb * ntiles) * 8;
            const double *ap = sLdp + ((size_t)qa * npb + pb) * nz + fc;
            const double *bp = sW + (size_t)min(ig0 + fr, ng - 1) * nz + fc;
            double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 5
            for (int ks = 0; ks < ksteps; ++ks) {
                const bool inz = ks * 4 + fc < nz;
                const double av = inz ? ap[ks * 4] : 0.0;
                const double b0 = inz ? bp[ks * 4] : 0.0, b1 = inz ? bp[rowlen + ks * 4] : 0.0;
                dmma_m8n8k4(c00, c01, av, b0);
                dmma_m8n8k4(c10, c11, av, b1);
            }
            const int ig = ig0 + fc * 2;
            if (fr < cnt && ig < ng) {
                const double ak = s_ak[fr];
                double *out = P.rec + (size_t)s_pv[fr] * P.recstride + P.rec_ld + (size_t)pb * P.lds + ig;
                const double v0 = (1.0 - ak) * c00 + ak * c10;
                if (ig + 1 < ng) *reinterpret_cast<double2 *>(out) = make_double2(v0, (1.0 - ak) * c01 + ak * c11);
                else out[0] = v0;
            }
        }
    } else {
        // direct weights for this radius ratio (calculate_weights_2d, common.py:152-185), one vector at a time
        for (int q = 0; q < cnt; ++q) {
            __syncthreads();
            const double k0 = P.k[(size_t)s_pv[q] * P.kcols];
            for (int g = tid; g < ng; g += 256) weight_row(k0, P.gs[g], P.ze, nz, sW + (size_t)g * nz, 1);
            __syncthreads();
            for (int idx = tid; idx < npb * ng; idx += 256) {
                const int pb = idx / ng, g = idx - pb * ng;
                const double *wr = sW + (size_t)g * nz, *lp = sLdp + ((size_t)q * npb + pb) * nz;
                double acc = 0.0;
                int iz = g % nz;
                for (int j = 0; j < nz; ++j) {
                    acc = fma(wr[iz], lp[iz], acc);
                    iz = (iz + 1 == nz) ? 0 : iz + 1;
                }
                P.rec[(size_t)s_pv[q] * P.recstride + P.rec_ld + (size_t)pb * P.lds + g] = acc;
            }
        }
    }
    __syncthreads();  // the contraction's stores precede the (rare) NaN overwrite below
    for (int r = tid; r < cnt * npb; r += 256) {
        const int q = r / npb, pb = r - q * npb;
        const int ipv = s_pv[q];
        const double kk = P.k[(size_t)ipv * P.kcols + (P.kcols == npb ? pb : 0)];
        double *rowp = P.rec + (size_t)ipv * P.recstride + P.rec_ld + (size_t)pb * P.lds;
        double *tail = rowp + ng;
        tail[0] = kk;
        tail[1] = 1.0 / (1.0 + kk);
        // a disk integral that is not finite (the reference's numeric fallback gives NaN for 'logarithmic', -inf for
        // 'exponential': mu = 0 is a node) turns every in-box point into NaN there: (I* - x) / I*  (model_full.py:97)
        tail[2] = isfinite(sIstar[r]) ? 1.0 / sIstar[r] : nan("");
        tail[3] = kk * kk;
        for (int j = ng + 4; j < P.lds; ++j) tail[j - ng] = 0.0;
        // k < -1 makes g = z / (1 + k) negative, for which interpolate_mean_limb_darkening_s returns NaN
        // (common.py:227): the row itself carries the NaN, so the per-sample code needs no sign test
        if (1.0 + kk < 0.0)
            for (int g = 0; g < ng; ++g) rowp[g] = nan("");
    }
}

// ---------------------------------------------------------------------------------------------
// The npv x npt pass.
//
// Persistent kernel, one wave of CTAs (SM count x resident CTAs per SM).  Every WARP is an independent
// worker: it pulls work items (parameter vector x chunk of the time axis) from a global counter and
// never synchronises with the other warps of its CTA, so a vector with more transits than its
// neighbours delays nobody.  The item's per-vector record (orbit coefficients, transit centres, ld
// rows) is fetched by ONE TMA bulk copy into the warp's double-buffered shared-memory slot while the
// warp is still working on the previous item: no item starts with a chain of dependent global loads.
//
// Per item (the time axis is cut into blocks of 64 consecutive points):
//  1. Block classification: set_data stores the smallest and largest time stamp of every block.  A
//     block whose [tmin, tmax] misses every transit window [t0 + n p + lo, t0 + n p + hi] holds no
//     in-box point.  One ballot per 32 blocks -> a bitmap in shared memory (likelihood mode: the
//     pre-summed (obs-1)^2 of the untouched blocks is added right here).
//  2. Untouched blocks (~93 % of a TESS sector): 64 fluxes of exactly 1.0 -- predicate-free 16-byte
//     streaming stores, eight blocks (4 KB) per step.
//  3. Touched blocks: each point is folded and box-tested in fp64 in the reference's operation order
//     (model_full.py:88-91); in-box points go to the warp's queue; the next touched block's time
//     stamps are already in flight while the current one is folded.
//  4. Point-major drain: whenever 32 points are queued every lane takes ONE point and walks its
//     exposure sub-samples in order -- separation (two Horner quartics + sqrt) and ld-mean lerp per
//     sample, no integer division or per-sample metadata reload.  Samples on the stellar limb (the
//     ones that need the sqrt + 2 atan2 lens area, ~1/4 of them) are compacted into a second queue
//     and evaluated 32 at a time with full warps.  With one sample per point (S1) the limb queue is
//     carried across drains and flushed once per item; with supersampling the per-point sum over
//     sub-samples is taken in exposure order (model_full.py:93-99) from a per-pass buffer.
// The drain has a single call site and a point's arithmetic does not depend on batch composition or
// on which warp runs the item: results are bit-reproducible across launches and population splits.
// ---------------------------------------------------------------------------------------------
struct PointsParams {
    const double *time;
    const int32_t *lcids, *pbids, *epids, *nsamples;
    const double *exptimes;
    const double *rec;    // [npv][recstride] per-vector records
    void *flux;           // [npv][npt] fp64, or fp32 in the opt-in fp32 mode
    const double *obs;
    const int32_t *blk;
    const double *isig2;  // [npv][nblocks]
    double *partial;      // [npv][nchunks]
    const double *bmin, *bmax;  // per 64-point block: smallest / largest time stamp
    const int32_t *blc;         // light curve of the block, -1 when it straddles light curves
    const double *bchi;         // likelihood: sum of (obs-1)^2 over the block's points
    const int32_t *bnoise;      // likelihood: noise id of the block, -1 none, -2 mixed (slow path)
    const double *cmin, *cmax;  // the same per 16-point cell (supersampled kernel, ptb_ss_kernels.cuh)
    const int32_t *clc;
    const double *cchi;
    const int32_t *cnoise;
    int *work;                  // [0] next item, [1] CTAs finished (the last one re-arms both)
    long long npt;
    int npv, nlc, npb, nep, ng, lds, nblocks, ns_max, nchunks, blocks_per_chunk, nblk64;
    int recstride, rec_ld, ssc, frac_tab, ncell;
    double dg, inv_dg;
};

// resident CTAs per SM of the one-sample kernel (<= 85 registers); the supersampled kernel has its own bound
// (ptb_ss_kernels.cuh)
#ifndef PT_MINB_S1
#define PT_MINB_S1 3
#endif
constexpr int PT_THREADS = 256;
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_BLOCK = 64;           // points per classification block
constexpr int SS_CELL = 16;            // ... of the supersampled kernel
constexpr int PT_MAXBLK = 2048;        // blocks per item (hit bitmap in shared memory)
constexpr int PT_QCAP = 96;            // in-box point queue: < 32 carried + 64 from one block
constexpr int PT_LCAP = 64;            // limb sample queue: < 32 carried + 32 from one sub-sample step
constexpr int PT_COLS = 33;            // stride of the supersampled kernel's per-lane limb columns (odd: a lane's samples and
                                       // consecutive lanes' samples fall into different banks)
#ifndef PT_SSC_MAX_
#define PT_SSC_MAX_ 12
#endif
constexpr int PT_SSC_MAX = PT_SSC_MAX_;  // exposure sub-samples buffered per pass
constexpr int PT_FRAC_MAX = 1024;      // entries of the tabulated sub-sample offsets
constexpr double PT_EPS = 1e-9;        // classification margin, in periods (>> rounding, << the 0.003 d pad)

// Warp-private shared memory of k_rr_points: queues, two record slots (double buffered), the hit bitmap, two
// mbarriers and, in fp32 mode, the record converted to float.  T is the sample arithmetic / output type (double, or
// float in the opt-in fp32 mode).
__host__ __device__ inline size_t pt_warp_bytes(int recstride, int tsize) {
    const size_t rec_t = (tsize == 4) ? (((size_t)recstride * 4 + 15) & ~size_t(15)) : 0;
    const size_t n = (size_t)(PT_QCAP + 2 * PT_LCAP) * tsize + (size_t)(2 * PT_QCAP + 2 * PT_LCAP) * 4 + PT_MAXBLK / 8 + 16 +
                     (size_t)2 * recstride * 8 + rec_t;
    return (n + 15) & ~size_t(15);   // the next warp's mbarriers and TMA record slots stay 16-byte aligned
}
// CTA-wide part: the per-light-curve tables
__host__ __device__ inline size_t pt_shared_bytes(int nlc, int tsize) {
    (void)tsize;
    return (((size_t)nlc) * 8 + 2 * (size_t)nlc * 4 + 127) & ~size_t(127);   // box pad, ld-row offset, epoch id
}

template <typename T>
struct WarpScratch {
    unsigned char *base;
    int nrec;  // record slots (2: double buffered)
    int lcap;  // entries of the limb queue
    __device__ __forceinline__ WarpScratch(unsigned char *b, int nrec_, int lcap_) : base(b), nrec(nrec_), lcap(lcap_) {}
    __device__ __forceinline__ T *q_tc() const { return reinterpret_cast<T *>(base); }
    __device__ __forceinline__ T *l_z() const { return q_tc() + PT_QCAP; }
    __device__ __forceinline__ T *l_ip() const { return l_z() + lcap; }
    __device__ __forceinline__ int *q_ipt() const { return reinterpret_cast<int *>(l_ip() + lcap); }
    __device__ __forceinline__ int *q_lc() const { return q_ipt() + PT_QCAP; }
    __device__ __forceinline__ int *l_slot() const { return q_lc() + PT_QCAP; }   // the point index
    __device__ __forceinline__ int *l_row() const { return l_slot() + lcap; }
    __device__ __forceinline__ unsigned *hit() const { return reinterpret_cast<unsigned *>(l_row() + lcap); }
    __device__ __forceinline__ uint64_t *bar() const { return reinterpret_cast<uint64_t *>(hit() + PT_MAXBLK / 32); }
    __device__ __forceinline__ double *rec(int slot, int recstride) const {
        return reinterpret_cast<double *>(bar() + 2) + (size_t)slot * recstride;
    }
    // fp32 mode: the current record converted to float (16-byte aligned); fp64: unused
    __device__ __forceinline__ T *rec_t(int recstride) const { return reinterpret_cast<T *>(rec(nrec, recstride)); }
};

// VEC consecutive points per lane: time stamps are always fp64, fluxes are T.
template <int VEC, typename T>
struct VecIO;
template <typename T>
struct VecIO<1, T> {
    __device__ static __forceinline__ void load(const double *p, double *v) { v[0] = __ldg(p); }
    __device__ static __forceinline__ void store(T *p, const T *v) { __stcs(p, v[0]); }
    __device__ static __forceinline__ void store_keep(T *p, const T *v) { *p = v[0]; }
};
template <>
struct VecIO<2, double> {
    __device__ static __forceinline__ void load(const double *p, double *v) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        v[0] = t.x;
        v[1] = t.y;
    }
    __device__ static __forceinline__ void store(double *p, const double *v) {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
    }
    __device__ static __forceinline__ void store_keep(double *p, const double *v) {
        *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    }
};
template <>
struct VecIO<2, float> {
    __device__ static __forceinline__ void load(const double *p, double *v) { VecIO<2, double>::load(p, v); }
    __device__ static __forceinline__ void store(float *p, const float *v) {
        __stcs(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
    }
    __device__ static __forceinline__ void store_keep(float *p, const float *v) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};

// Per-CTA constants of the drain (shared-memory tables are item independent).
template <typename T>
struct DrainCtx {
    const PointsParams *P;
    const int *sRow;               // per light curve: offset of the passband row in the ld block
    int row1;                      // single light curve: the same as a scalar
    int lane;
};

// One sample at time `t` from mid-transit, straight-line code (selects only): separation (taylor_z.py:229-255; rsqrt +
// one coupled Newton step, 2^-43), LD-mean lerp (common.py:225-233) at grid position z * xs (xs = 1/((1+k) dg)) and the
// area cases that need no lens formula (common.py:52-73).  `limb` marks a sample on the limb: its contribution comes
// from limb_pass.
template <typename T>
__device__ __forceinline__ void sample_eval(T t, const T *cx, const T *cy, const T *row, int ng, T k, T xs, T inv_istar,
                                            T k2, T &z, T &ip, T &cc, bool &limb) {
    const T one = T(1), pi = T(kPi), qnan = T(nan(""));
    const T px = fma(t, fma(t, fma(t, fma(t, cx[4], cx[3]), cx[2]), cx[1]), cx[0]);
    const T py = fma(t, fma(t, fma(t, fma(t, cy[4], cy[3]), cy[2]), cy[1]), cy[0]);
    z = sqrt_sep(fma(px, px, py * py));
    ip = ld_lerp(z * xs, row, ng);                            // used only where the planet overlaps the disk
    const bool out = (one + k <= z);
    limb = !out && (fabs(one - k) < z);
    const bool covers = (z <= k - one);                       // planet covers the star: area pi; else pi k^2
    const bool inside = covers || (z <= one - k);
    const T c = one - ip * (covers ? pi : pi * k2) * inv_istar;
    // no overlap: area 0 -> (I* - 0) / I* = 1, or NaN when 1/I* is NaN (non-finite disk integral, see k_rr_ldm)
    cc = out ? fma(T(0), inv_istar, one) : (inside ? c : qnan);
}

// Lens-area pass over `take` limb samples at the top of the limb queue (warp-cooperative, one-sample kernels):
// writes the flux / accumulates chi^2 directly.
template <bool SINGLE_LC, bool LNL, typename T>
__device__ __forceinline__ double limb_pass(const PointsParams &P, const WarpScratch<T> &ws, const T *ld, const T *row1,
                                            T *frow, const double *isig2, double w_one, int lane, int first, int take) {
    double chi = 0.0;
    if (lane < take) {
        const int q = first + lane, ng = P.ng;
        const T *r2 = SINGLE_LC ? row1 : ld + ws.l_row()[q];
        const T v = T(1) - ws.l_ip()[q] * kite_area_limb<T>(r2[ng], r2[ng + 3], ws.l_z()[q]) * r2[ng + 2];
        const int ipt = ws.l_slot()[q];
        if (LNL) {
            const int b = P.blk ? P.blk[ipt] : 0;
            if (b >= 0) {   // the point's (obs - 1)^2 is already in the block baseline: swap it for (obs - model)^2
                const double o = P.obs[ipt], d1 = o - (double)v, d0 = o - 1.0;
                chi = fma(d1, d1, -d0 * d0) * (P.blk ? isig2[b] : w_one);
            }
        } else {
            frow[ipt] = v;
        }
    }
    return chi;
}

// One sample per point (nsamples == 1): point-major evaluation of `n` (<= 32) queued in-box points starting at queue
// slot `base` (model_full.py:93-99).  `rt` is the record in the arithmetic type.  `nl` is the fill of the limb queue,
// carried across calls; `flush` empties it.  Inlined at its single call site.  Returns the lane's chi^2 increment.
template <bool SINGLE_LC, bool LNL, typename T>
__device__ __forceinline__ double drain_points(const DrainCtx<T> &c, const WarpScratch<T> &ws, const T *rt, T *frow,
                                               const double *isig2, double w_one, int base, int n, int &nl, bool flush) {
    const PointsParams &P = *c.P;
    const int lane = c.lane, ng = P.ng;
    const T inv_dg = (T)P.inv_dg;
    const unsigned lt_mask = (1u << lane) - 1u;
    const T *ld = rt + P.rec_ld;
    T cx[5], cy[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { cx[j] = rt[j]; cy[j] = rt[5 + j]; }

    const bool valid = lane < n;
    int ipt = 0, rowoff = c.row1;
    T tc = T(0);
    if (valid) {
        ipt = ws.q_ipt()[base + lane];
        tc = ws.q_tc()[base + lane];
        if (!SINGLE_LC) rowoff = c.sRow[ws.q_lc()[base + lane]];
    }
    // likelihood: the observed flux (and noise id) of the lane's point, requested now so that the load is in flight
    // under the sample evaluation instead of in front of the chi^2 update
    double obs_v = 1.0;
    int nz_b = 0;
    if (LNL && valid) {
        obs_v = __ldg(P.obs + ipt);
        if (P.blk) nz_b = __ldg(P.blk + ipt);
    }
    __syncwarp();   // the queue slots are read: the next fold step may reuse them
    const T *row1 = ld + c.row1;
    const T *row = ld + rowoff;
    const T k = row[ng], inv1k = row[ng + 1], inv_istar = row[ng + 2], k2 = row[ng + 3];

    double chi = 0.0;
    T z, ip, cc;
    bool limb;
    // the exposure offset exptime*((1-0.5)/1 - 0.5) is exactly 0 (model_full.py:94), as the reference's own product is
    sample_eval<T>(tc, cx, cy, row, ng, k, inv1k * inv_dg, inv_istar, k2, z, ip, cc, limb);
    limb = limb && valid;
    const unsigned m = __ballot_sync(0xffffffffu, limb);
    if (m) {
        if (limb) {
            const int pos = nl + __popc(m & lt_mask);
            ws.l_slot()[pos] = ipt;
            ws.l_z()[pos] = z;
            ws.l_ip()[pos] = ip;
            if (!SINGLE_LC) ws.l_row()[pos] = rowoff;
        }
        nl += __popc(m);
        __syncwarp();
    }
    if (valid && !limb) {   // everything that is not on the limb is final
        if (LNL) {
            if (nz_b >= 0) {   // swap the baseline's (obs - 1)^2 for (obs - model)^2
                const double d1 = obs_v - (double)cc, d0 = obs_v - 1.0;
                chi += fma(d1, d1, -d0 * d0) * (P.blk ? isig2[nz_b] : w_one);   // one noise block: its weight is an item constant
            }
        } else {
            frow[ipt] = cc;
        }
    }
    // lens area on the limb (sqrt + 2 atan2): a full warp at a time from the top of the limb queue (single code
    // copy: one rounding behaviour); leftovers when the item ends
    while (nl >= 32 || (flush && nl > 0)) {
        const int take = min(nl, 32);
        nl -= take;
        chi += limb_pass<SINGLE_LC, LNL, T>(P, ws, ld, row1, frow, isig2, w_one, lane, nl, take);
        __syncwarp();
    }
    return chi;
}

template <int VEC, bool SINGLE_LC, bool LNL, typename T>
__global__ void __launch_bounds__(PT_THREADS, PT_MINB_S1) k_rr_points(const __grid_constant__ PointsParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool F32 = sizeof(T) == 4;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long npt = P.npt;
    const int nlc = P.nlc;
    const long long nitems = (long long)P.npv * P.nchunks;
    const unsigned lt_mask = (1u << lane) - 1u;

    // dynamic smem: [per-light-curve tables] [warp-private area x 8]
    double *sPad = reinterpret_cast<double *>(smem_raw);
    int *sRow = reinterpret_cast<int *>(sPad + nlc);
    int *sEp = sRow + nlc;
    constexpr int NREC = 2;
    const WarpScratch<T> ws(smem_raw + pt_shared_bytes(nlc, (int)sizeof(T)) + (size_t)warp * pt_warp_bytes(P.recstride, (int)sizeof(T)), NREC, PT_LCAP);
    unsigned *s_hit = ws.hit();
    uint64_t *bar = ws.bar();

    const uint32_t rec_bytes = (uint32_t)P.recstride * 8u;
    long long item = 0;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        item = atomicAdd(&P.work[0], 1);
        if (NREC == 2 && item < nitems) {
            mbar_expect_tx(&bar[0], rec_bytes);
            tma_load_1d(ws.rec(0, P.recstride), P.rec + (size_t)(item / P.nchunks) * P.recstride, rec_bytes, &bar[0]);
        }
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    // item-independent per-light-curve tables
    for (int lc = tid; lc < nlc; lc += PT_THREADS) {
        sPad[lc] = 0.003 + P.exptimes[lc];  // model_full.py:69-70
        sRow[lc] = P.pbids[lc] * P.lds;
        sEp[lc] = P.epids[lc];
    }
    __syncthreads();  // the only CTA-wide barrier: from here on the warps are independent workers

    DrainCtx<T> dctx;
    dctx.P = &P; dctx.sRow = sRow; dctx.row1 = sRow[0]; dctx.lane = lane;
    const double pad1 = sPad[0];
    const int ep1 = sEp[0];
    T *flux = reinterpret_cast<T *>(P.flux);

    for (int iter = 0; item < nitems; ++iter) {
        const int buf = (NREC == 2) ? (iter & 1) : 0;
        long long next = 0;
        if (lane == 0) {  // fetch the next item and start its record copy into the other slot
            if (NREC == 1) {  // single slot: this item's record (every lane left the slot at the end of the last item)
                mbar_expect_tx(&bar[0], rec_bytes);
                tma_load_1d(ws.rec(0, P.recstride), P.rec + (size_t)(item / P.nchunks) * P.recstride, rec_bytes, &bar[0]);
            }
            next = atomicAdd(&P.work[0], 1);
            if (NREC == 2 && next < nitems) {
                mbar_expect_tx(&bar[buf ^ 1], rec_bytes);
                tma_load_1d(ws.rec(buf ^ 1, P.recstride), P.rec + (size_t)(next / P.nchunks) * P.recstride, rec_bytes,
                            &bar[buf ^ 1]);
            }
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        const int ipv = (int)(item / P.nchunks);
        const int chunk = (int)(item - (long long)ipv * P.nchunks);
        item = next;
        const int bbeg = chunk * P.blocks_per_chunk;
        const int bend = min(P.nblk64, bbeg + P.blocks_per_chunk);
        const int nbc = bend - bbeg;
        const double *rec = ws.rec(buf, P.recstride);
        T *frow = LNL ? nullptr : flux + (size_t)ipv * npt;
        const double *isig2 = LNL ? P.isig2 + (size_t)ipv * P.nblocks : nullptr;
        double chi = 0.0;

        mbar_wait(&bar[buf], (NREC == 2) ? ((iter >> 1) & 1) : (iter & 1));
        if (rec[ORB_GOOD] == 0.0 || rec[ORB_LDNAN] != 0.0) {  // invalid parameter vector: NaN row (model_full.py:40,80-82)
            if (LNL) {
                if (lane == 0) P.partial[(size_t)ipv * P.nchunks + chunk] = nan("");
            } else {
                T v[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) v[j] = T(nan(""));
                const long long cend = min(npt, (long long)bend * PT_BLOCK);
                for (long long i = (long long)bbeg * PT_BLOCK + (long long)lane * VEC; i < cend; i += 32 * VEC)
                    VecIO<VEC, T>::store(frow + i, v);
            }
            __syncwarp();
            continue;
        }
        // the record in the arithmetic type: fp64 -> the TMA slot itself; fp32 -> converted once per item
        const T *rt;
        if (F32) {
            T *dst = ws.rec_t(P.recstride);
            for (int i = lane; i < P.recstride; i += 32) dst[i] = (T)rec[i];
            __syncwarp();
            rt = dst;
        } else {
            rt = reinterpret_cast<const T *>(rec);
        }

        const double p = rec[ORB_P], invp = rec[ORB_INVP], T1 = rec[ORB_T1], T4 = rec[ORB_T4];
        const double *t0v = rec + ORB_STRIDE;
        const double lo1 = T1 - pad1, hi1 = T4 + pad1, t01 = t0v[ep1];

        const double w_one = (LNL && !P.blk) ? isig2[0] : 0.0;  // single noise block: its weight is an item constant

        // ---- 1. classification of every block of this item --------------------------------------------
        for (int bb0 = 0; bb0 < nbc; bb0 += 32) {
            const int bb = bb0 + lane, b = bbeg + bb;
            bool hit = false;
            if (bb < nbc) {
                hit = true;
                const int lcb = SINGLE_LC ? 0 : P.blc[b];
                int nz_id = 0;
                if (LNL && P.blk) nz_id = P.bnoise ? P.bnoise[b] : 0;  // one block over all points: id 0 everywhere
                const bool partial = (b == P.nblk64 - 1) && (npt % PT_BLOCK != 0);
                if (lcb >= 0 && nz_id != -2 && !partial) {
                    const double lo = SINGLE_LC ? lo1 : T1 - sPad[lcb], hi = SINGLE_LC ? hi1 : T4 + sPad[lcb];
                    const double t0 = SINGLE_LC ? t01 : t0v[sEp[lcb]];
                    const double n1 = ceil(fma(P.bmin[b] - t0 - hi, invp, -PT_EPS));
                    const double n2 = floor(fma(P.bmax[b] - t0 - lo, invp, PT_EPS));
                    hit = !(n1 > n2) || !(p > 0.0);  // NaNs and p <= 0 fall through to the exact per-point path
                }
                // likelihood baseline: sum of (obs - 1)^2 of EVERY block that has one noise id, touched or not (the
                // drain swaps the in-box points' terms for (obs - model)^2); blocks with several ids: per point, below
                if (LNL && nz_id >= 0) chi = fma(P.bchi[b], P.blk ? isig2[nz_id] : w_one, chi);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_hit[bb0 >> 5] = m;
        }
        __syncwarp();

        // ---- 2. untouched blocks: 64 fluxes of exactly 1.0, vectorised streaming stores -------------------
        // (TMA bulk stores out of a constant shared-memory buffer were measured: same kernel time, and the
        //  16 KB buffer costs the supersampled variants a resident CTA -- DESIGN.md section 3.1)
        if (!LNL) {
            T one[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) one[j] = T(1);
            const int ngroups = (nbc + 7) >> 3;
            for (int g = 0; g < ngroups; ++g) {
                const unsigned bits = (s_hit[g >> 2] >> ((g & 3) * 8)) & 0xffu;
                const int b0 = bbeg + g * 8;
                T *fb = frow + (long long)b0 * PT_BLOCK + lane * VEC;
                if (bits == 0u && b0 + 8 <= bend) {  // the common case: eight untouched blocks, no predicates
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
#pragma unroll
                        for (int h = 0; h < 2 / VEC; ++h) VecIO<VEC, T>::store(fb + j * PT_BLOCK + h * 32 * VEC, one);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (!((bits >> j) & 1u) && b0 + j < bend) {
#pragma unroll
                            for (int h = 0; h < 2 / VEC; ++h) VecIO<VEC, T>::store(fb + j * PT_BLOCK + h * 32 * VEC, one);
                        }
                    }
                }
            }
        }

        // ---- 3./4. touched blocks: exact per-point path + drain -----------------------------------------
        // One extra pass after the last block flushes the queues, so that the drain has exactly ONE call
        // site and batches are always cut from the top of the queue.
        int qn = 0, nl = 0;
        const int nwords = (nbc + 31) >> 5;
        int wi = 0;
        unsigned wbits = s_hit[0];
        // find the first touched block and start loading its time stamps
        auto next_block = [&]() -> int {
            while (wbits == 0u) {
                if (++wi >= nwords) return -1;
                wbits = s_hit[wi];
            }
            const int j = __ffs(wbits) - 1;
            wbits &= wbits - 1;
            return wi * 32 + j;
        };
        int cur = next_block();
        // A lane folds two points of every 64-point block: A and B are neighbours (VEC == 2, one 16-byte load) or 32
        // points apart (VEC == 1, odd npt or unaligned arrays).  Scalars and 32-bit offsets from the lane's corner of
        // the item (64-bit addresses once per item).
        constexpr int DB = (VEC == 2) ? 1 : 32;
        const long long lbase = (long long)bbeg * PT_BLOCK + lane * VEC;
        const double *tl = P.time + lbase;
        const double *ol = LNL ? P.obs + lbase : nullptr;
        T *fl = LNL ? nullptr : frow + lbase;
        const int32_t *bl = SINGLE_LC ? nullptr : P.blc + bbeg;
        const long long left = npt - lbase;                                   // offsets < rem are inside the time axis
        const int rem = (int)(left > 0x7fffffffll ? 0x7fffffffll : left);   // (VEC == 2 requires an even npt: vectors are all-in or all-out)
        const int ipt0 = (int)lbase;
        double tnA = 0.0, tnB = 0.0;   // the prefetched block: time stamps,
        int lcbn = 0, nzn = 0;         // its light curve (-1: mixed, per-point lookup) and (likelihood) noise id (-2: mixed)
        const bool nz_lookup = LNL && P.blk && P.bnoise;
        auto load_block = [&](int bb) {
            const int off = bb * PT_BLOCK;
            if (!SINGLE_LC) lcbn = __ldg(bl + bb);
            if (nz_lookup) nzn = __ldg(P.bnoise + bbeg + bb);
            tnA = 0.0; tnB = 0.0;
            if (VEC == 2) {
                if (off < rem) {
                    const double2 t = __ldg(reinterpret_cast<const double2 *>(tl + off));
                    tnA = t.x; tnB = t.y;
                }
            } else {
                if (off < rem) tnA = __ldg(tl + off);
                if (off + DB < rem) tnB = __ldg(tl + off + DB);
            }
        };
        if (cur >= 0) load_block(cur);
        int lc_last = SINGLE_LC ? 0 : -2;           // the block constants below belong to this light curve
        double lob = lo1, hib = hi1, t0b = t01;
        T ones[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) ones[j] = T(1);
        for (;;) {
            const bool live = cur >= 0;
            if (live) {
                const int offA = cur * PT_BLOCK, offB = offA + DB;
                const double tA = tnA, tB = tnB;
                const int lcb = SINGLE_LC ? 0 : lcbn, nzb = nzn;
                if (!SINGLE_LC && lcb >= 0 && lcb != lc_last) {   // window and transit centre of the block's light curve
                    const double pd = sPad[lcb];
                    lob = T1 - pd;
                    hib = T4 + pd;
                    t0b = t0v[sEp[lcb]];
                    lc_last = lcb;
                }
                cur = next_block();
                if (cur >= 0) load_block(cur);  // in flight while this block is folded
                const bool inrA = offA < rem, inrB = offB < rem;
                bool inA, inB;
                int lcA = lcb, lcB = lcb;
                double tcA, tcB;
                // epoch = floor((t - t0 + p/2)/p); tc = t - (t0 + epoch p)  (model_full.py:88-89).  The division is a
                // multiplication by 1/p: the two can only disagree half a period away from the transit, where the
                // point is outside the box either way.  Always fp64: time stamps need all their digits; tc is small.
                if (SINGLE_LC || lcb >= 0) {   // the rule: window and transit centre are block constants
                    const double eA = floor(fma(tA - t0b, invp, 0.5)), eB = floor(fma(tB - t0b, invp, 0.5));
                    tcA = tA - __dadd_rn(t0b, __dmul_rn(eA, p));
                    tcB = tB - __dadd_rn(t0b, __dmul_rn(eB, p));
                    inA = inrA && (lob <= tcA) && (tcA <= hib);
                    inB = inrB && (lob <= tcB) && (tcB <= hib);
                } else {                       // a block that straddles light curves: per-point lookup
                    lcA = inrA ? P.lcids[(long long)ipt0 + offA] : 0;
                    lcB = inrB ? P.lcids[(long long)ipt0 + offB] : 0;
                    const double pdA = sPad[lcA], t0A = t0v[sEp[lcA]], pdB = sPad[lcB], t0B = t0v[sEp[lcB]];
                    const double eA = floor(fma(tA - t0A, invp, 0.5)), eB = floor(fma(tB - t0B, invp, 0.5));
                    tcA = tA - __dadd_rn(t0A, __dmul_rn(eA, p));
                    tcB = tB - __dadd_rn(t0B, __dmul_rn(eB, p));
                    inA = inrA && (T1 - pdA <= tcA) && (tcA <= T4 + pdA);
                    inB = inrB && (T1 - pdB <= tcB) && (tcB <= T4 + pdB);
                }
                if (LNL && nzb == -2) {   // several noise ids in this block: its baseline, point by point
                    if (inrA) {
                        const int nb = P.blk[(long long)ipt0 + offA];
                        const double d = ol[offA] - 1.0;
                        if (nb >= 0) chi = fma(d * d, isig2[nb], chi);
                    }
                    if (inrB) {
                        const int nb = P.blk[(long long)ipt0 + offB];
                        const double d = ol[offB] - 1.0;
                        if (nb >= 0) chi = fma(d * d, isig2[nb], chi);
                    }
                }
                // one compaction for the lane's two points
                const unsigned mA = __ballot_sync(0xffffffffu, inA), mB = __ballot_sync(0xffffffffu, inB);
                int pos = qn + __popc(mA & lt_mask) + __popc(mB & lt_mask);
                qn += __popc(mA) + __popc(mB);
                if (inA) {
                    ws.q_ipt()[pos] = ipt0 + offA;
                    ws.q_tc()[pos] = (T)tcA;
                    if (!SINGLE_LC) ws.q_lc()[pos] = lcA;
                    ++pos;
                }
                if (inB) {
                    ws.q_ipt()[pos] = ipt0 + offB;
                    ws.q_tc()[pos] = (T)tcB;
                    if (!SINGLE_LC) ws.q_lc()[pos] = lcB;
                }
                // 1.0 for the block's points, default cache policy (not evict-first): the line is still in L2 when the
                // drain updates its in-box points, so it reaches DRAM once
                if (!LNL) {
                    if (VEC == 2) {
                        if (inrA) VecIO<VEC, T>::store_keep(fl + offA, ones);
                    } else {
                        if (inrA) fl[offA] = T(1);
                        if (inrB) fl[offB] = T(1);
                    }
                }
                __syncwarp();
            }
            // drain full warps of points; the remainder (< 32) waits for more, except on the flush pass
            while (qn >= 32 || (!live && (qn > 0 || nl > 0))) {
                const int n = min(qn, 32);
                qn -= n;
                chi += drain_points<SINGLE_LC, LNL, T>(dctx, ws, rt, frow, isig2, w_one, qn, n, nl, !live && qn == 0);
            }
            if (!live) break;
        }

        if (LNL) {
            chi = warp_sum(chi);
            if (lane == 0) P.partial[(size_t)ipv * P.nchunks + chunk] = chi;
        }
        __syncwarp();  // every lane is done with this record slot before its next refill is issued
    }

    // the last CTA to leave re-arms the work counters for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(&P.work[1], 1);
        if (done == (int)gridDim.x - 1) {
            P.work[0] = 0;
            P.work[1] = 0;
            __threadfence();
        }
    }
}

// chi^2 partials -> lnL (wnloglikelihood.py:34): lnL = sum_b n_b (-log s_b - log(2 pi)/2) - chi^2/2.
// The result goes to `nout` destinations: one (the caller's buffer), or -- the fused all-gather of a sharded
// population -- slot `rank` of the gathered array on EVERY GPU of the box, written through NVLink peer
// mappings (each rank's buffer is symmetric memory mapped into this process).
constexpr int LNL_MAXPEERS = 16;
struct LnlOut {
    double *ptr[LNL_MAXPEERS];
    // Device-side ordering of the fused all-gather (no host-issued barrier): flag[r] is destination r's arrival
    // array [world] of 64-bit step numbers.  The last CTA of k_lnl_finish publishes `seq` into slot `rank` of every
    // destination after all the shard's stores (release at system scope over NVLink); the last CTA of the destination's own k_lnl_finish
    // acquires them.  null = no signalling (single destination, or the caller orders the ranks itself).
    unsigned long long *flag[LNL_MAXPEERS];
    unsigned long long seq;
    unsigned long long timeout_ns;  // a peer that never arrives must not hang the GPU: give up and raise *err
    int *done;  // CTAs of this launch that have stored their part (re-armed by the last one)
    int *err;
    int nout, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void k_lnl_finish(const double *__restrict__ partial, int nchunks, const double *__restrict__ sigma,
                             const double *__restrict__ nblk, int nblocks, int npv, const __grid_constant__ LnlOut out) {
    __shared__ int s_last;
    const int ipv = blockIdx.x * blockDim.x + threadIdx.x;
    if (ipv < npv) {
        double chi = 0.0;
        for (int c = 0; c < nchunks; ++c) chi += partial[(size_t)ipv * nchunks + c];
        double cst = 0.0;
        for (int b = 0; b < nblocks; ++b) cst += nblk[b] * (-log(sigma[(size_t)ipv * nblocks + b]) - 0.5 * log(kTwoPi));
        const double v = cst - 0.5 * chi;
        for (int r = 0; r < out.nout; ++r) out.ptr[r][ipv] = v;
    }
    if (out.flag[0] == nullptr) return;
    // ---- fused all-gather, ordered on the device -------------------------------------------------------------
    __threadfence_system();  // this thread's peer stores are visible system-wide before the CTA reports
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(out.done, 1);
        s_last = (done == (int)gridDim.x - 1);
        if (s_last) {        // every CTA's stores have been fenced: publish the step number everywhere
            *out.done = 0;
            __threadfence_system();
            for (int r = 0; r < out.nout; ++r) st_release_sys(out.flag[r] + out.rank, out.seq);
        }
    }
    __syncthreads();
    // The last CTA stays until every rank's shard of this step has landed in THIS GPU's gathered array (lane r polls
    // source r): when the kernel retires, work queued behind it on the stream sees the complete array.  No separate
    // wait kernel, nothing for the host to do.
    if (s_last && (int)threadIdx.x < out.nout) {
        const unsigned long long *mine = out.flag[out.rank] + threadIdx.x;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int spins = 0;
        while (ld_acquire_sys(mine) < out.seq) {
            if ((++spins & 1023) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > out.timeout_ns) {
                    atomicExch(out.err, 1 + (int)threadIdx.x);
                    break;
                }
            }
        }
    }
}

__global__ void k_inv_sigma2(const double *__restrict__ sigma, long long n, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double s = sigma[i];
        out[i] = 1.0 / (s * s);
    }
}

// Multiplicative baseline of the LPF layer (lpf/lpf.py:418-428) as a linear model in per-point basis functions:
//   bl[ipv, j] = sum_c pvp[ipv, i_bl + cstart[lc(j)] + c] * basis[c][j],  c < ncoef[lc(j)]   (1 where ncoef = 0)
// which covers both baselines of the reference: LegendreBaseline (lpf/baselines/legendrebaseline.py:23-40, basis =
// Legendre polynomials of the normalised time, summed in the same order) and LinearModelBaseline
// (lpf/baselines/linearbaseline.py:22-36, basis = 1 and the covariates).
struct BaselineParams {
    const double *pvp;      // [npv][npar]
    const double *basis;    // [nbasis][npt]; null: no baseline
    const int32_t *lcids;   // [npt], or null with a single light curve
    const int32_t *cstart, *ncoef;  // [nlc]
    long long npt;
    int npar, i_bl;
};

__device__ __forceinline__ double baseline_at(const BaselineParams &B, const double *pv, long long ipt) {
    const int lc = B.lcids ? B.lcids[ipt] : 0;
    const int nc = B.ncoef[lc];
    if (nc == 0) return 1.0;
    const double *c = pv + B.i_bl + B.cstart[lc];
    double bl = 0.0;
    for (int j = 0; j < nc; ++j) bl += c[j] * B.basis[(size_t)j * B.npt + ipt];
    return bl;
}

// flux[ipv, :] *= baseline (BaseLPF.flux_model, lpf.py:445-449, trends = 0), or flux = baseline (BaseLPF.baseline)
__global__ void __launch_bounds__(256) k_lpf_baseline(const __grid_constant__ BaselineParams B, double *__restrict__ flux, int only_baseline) {
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= B.npt) return;
    const int ipv = blockIdx.y;
    const double bl = baseline_at(B, B.pvp + (size_t)ipv * B.npar, j);
    double *f = flux + (size_t)ipv * B.npt + j;
    *f = only_baseline ? bl : bl * *f;
}

// lnlike_normal on a materialised model (wnloglikelihood.py:22-35): one CTA per vector.  With a baseline the model
// value is baseline * transit flux (lpf.py:445-449), evaluated on the fly -- the baseline itself is never stored.
__global__ void __launch_bounds__(256) k_lnl_model(const double *__restrict__ model, const double *__restrict__ obs,
                                                   const int32_t *__restrict__ blk, const double *__restrict__ isig2,
                                                   long long npt, int nblocks, double *__restrict__ partial,
                                                   const __grid_constant__ BaselineParams B) {
    __shared__ double s_w[8];
    const int ipv = blockIdx.x;
    const double *m = model + (size_t)ipv * npt;
    const double *w = isig2 + (size_t)ipv * nblocks;
    const double *pv = B.basis ? B.pvp + (size_t)ipv * B.npar : nullptr;
    double chi = 0.0;
    for (long long j = threadIdx.x; j < npt; j += 256) {
        const int b = blk ? blk[j] : 0;
        if (b >= 0) {
            const double mv = pv ? baseline_at(B, pv, j) * m[j] : m[j];
            const double d = obs[j] - mv;
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

// ---------------------------------------------------------------------------------------------
// k_host_delta -- keeps a page-locked HOST copy of a device result array up to date by writing only what
// changed (ptb_bind_host_result).  A transit model flux is exactly 1.0 outside the transit windows
// (model_full.py:91): the array is cut into blocks of HD_BLOCK elements; a block is "lit" when any element
// differs from 1.0 (NaN rows included).  Lit blocks are written straight into the host array through
// its device mapping (coalesced 16-byte stores, HD_BLOCK * 8 B per block over PCIe); blocks that were lit after
// the previous call and are not any more are reset to 1.0; everything else already holds 1.0 on the
// host.  `lit` keeps one bit per block between calls.  mode 0: delta; mode 1: only rebuild `lit`
// (the caller ships the whole array with the copy engine).
// One warp per word of 32 blocks, eight 16-byte loads in flight per lane; a warp-wide load covers 64 elements,
// i.e. HD_BPI blocks, told apart by sub-masks of one ballot.
// ---------------------------------------------------------------------------------------------
constexpr int HD_BLOCK = 16;                 // elements per block: one 128-byte line (fp64) per PCIe write burst (measured: 64 -> 7.8e10, 32 -> 9.2e10, 16 -> 9.9e10, 8 -> 9.9e10 pts/s e2e on C2)
constexpr int HD_LPB = HD_BLOCK / 2;         // lanes per block (two elements per lane)
constexpr int HD_BPI = 32 / HD_LPB;          // blocks covered by one warp-wide load
constexpr int HD_UNR = (32 / HD_BPI < 8) ? 32 / HD_BPI : 8;   // loads in flight per lane

template <typename T>
__global__ void __launch_bounds__(256) k_host_delta(const T *__restrict__ src, T *__restrict__ host, unsigned *__restrict__ lit,
                                                    unsigned long long *__restrict__ nwritten, long long count,
                                                    long long word0, long long nwords, int mode) {
    const int lane = threadIdx.x & 31;
    const int sub = lane / HD_LPB;                                   // which of the HD_BPI blocks of a load is mine
    const unsigned submask = (HD_LPB == 32 ? 0xffffffffu : ((1u << HD_LPB) - 1u)) << (sub * HD_LPB);
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = word0 + warp0; w < nwords; w += nwarps) {   // words [word0, nwords)
        const unsigned old = mode == 0 ? lit[w] : 0u;
        unsigned cur = 0u;
        const long long e0 = w * 32 * HD_BLOCK + lane * 2;
#pragma unroll
        for (int j0 = 0; j0 < 32 / HD_BPI; j0 += HD_UNR) {   // 32 / HD_BPI warp-wide loads cover the word's 32 blocks
            T v[HD_UNR][2];
#pragma unroll
            for (int j = 0; j < HD_UNR; ++j) {
                const long long e = e0 + (long long)(j0 + j) * 64;
                v[j][0] = T(1);
                v[j][1] = T(1);
                if (e + 1 < count) {
                    if (sizeof(T) == 8) {
                        const double2 t = __ldcs(reinterpret_cast<const double2 *>(src + e));
                        v[j][0] = (T)t.x;
                        v[j][1] = (T)t.y;
                    } else {
                        const float2 t = __ldcs(reinterpret_cast<const float2 *>(src + e));
                        v[j][0] = (T)t.x;
                        v[j][1] = (T)t.y;
                    }
                } else if (e < count) {
                    v[j][0] = __ldcs(src + e);
                }
            }
#pragma unroll
            for (int j = 0; j < HD_UNR; ++j) {
                const long long e = e0 + (long long)(j0 + j) * 64;
                const bool ne = (v[j][0] != T(1)) || (v[j][1] != T(1));
                const unsigned m = __ballot_sync(0xffffffffu, ne);
                const int b0 = (j0 + j) * HD_BPI;                 // first block of this load inside the word
#pragma unroll
                for (int q = 0; q < HD_BPI; ++q) {
                    const unsigned qm = (HD_LPB == 32 ? 0xffffffffu : ((1u << HD_LPB) - 1u)) << (q * HD_LPB);
                    if (m & qm) cur |= 1u << (b0 + q);
                }
                const bool is_lit = (m & submask) != 0u;
                const bool was_lit = (old >> (b0 + sub)) & 1u;
                if (mode == 0 && (is_lit || was_lit)) {  // a block that went dark holds exactly 1.0 in v
                    if (e + 1 < count) {
                        if (sizeof(T) == 8) *reinterpret_cast<double2 *>(host + e) = make_double2((double)v[j][0], (double)v[j][1]);
                        else *reinterpret_cast<float2 *>(host + e) = make_float2((float)v[j][0], (float)v[j][1]);
                    } else if (e < count) {
                        host[e] = v[j][0];
                    }
                }
            }
        }
        if (lane == 0) {
            lit[w] = cur;
            if (mode == 0 && (cur | old)) atomicAdd(nwritten, (unsigned long long)__popc(cur | old));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_lpf_map -- BaseLPF.transit_model's parameter mapping (lpf/lpf.py:435-443) on the device: one thread
// per parameter vector turns a row of pvp[npv, npar] into the arguments of the RoadRunner evaluation.
// as_from_rhop: orbits_py.py:604-618 with D_S = 86400 and G = scipy.constants.G; map_ldc: lpf.py:84-91;
// sigma = 10**pv (wnloglikelihood.py:80).
// ---------------------------------------------------------------------------------------------
struct LpfLayout {  // mirror of ptb_lpf_layout (include/ptb200.h)
    int32_t npar, i_tc, i_p, i_rho, i_b, i_k2, nk2, i_ld, nldc, ld_map, i_secw, i_sesw, inc_mode, i_loge, nloge, ntc, i_bl, pad_;
    double tref;
};

struct LpfMapParams {
    const double *pvp;
    double *k, *ldc, *t0, *p, *a, *inc, *e, *w, *sigma;
    int npv, npb;
    LpfLayout L;
};

__global__ void k_lpf_map(const __grid_constant__ LpfMapParams P) {
    const int ipv = blockIdx.x * blockDim.x + threadIdx.x;
    if (ipv >= P.npv) return;
    const LpfLayout &L = P.L;
    const double *pv = P.pvp + (size_t)ipv * L.npar;
    const double per = pv[L.i_p], rho = pv[L.i_rho], b = pv[L.i_b];
    const double ps = per * 86400.0;
    const double a = pow(6.67430e-11 / (3.0 * kPi), 1.0 / 3.0) * pow(ps * ps * 1e3 * rho, 1.0 / 3.0);
    double e = 0.0, w = 0.0;
    if (L.i_secw >= 0) {
        const double c = pv[L.i_secw], s = pv[L.i_sesw];
        e = c * c + s * s;
        w = atan2(s, c);
    }
    const double inc = (L.inc_mode == 1) ? acos(b / (a * ((1.0 - e * e) / (1.0 + e * sin(w))))) : acos(b / a);
    for (int j = 0; j < L.ntc; ++j) P.t0[(size_t)ipv * L.ntc + j] = pv[L.i_tc + j] - L.tref;   // one per epoch (ttvlpf.py:83)
    P.p[ipv] = per;
    P.a[ipv] = a;
    P.inc[ipv] = inc;
    P.e[ipv] = e;
    P.w[ipv] = w;
    for (int j = 0; j < L.nk2; ++j) P.k[(size_t)ipv * L.nk2 + j] = sqrt(pv[L.i_k2 + j]);
    double *ld = P.ldc + (size_t)ipv * P.npb * L.nldc;
    const double *q = pv + L.i_ld;
    if (L.ld_map) {
        for (int pb = 0; pb < P.npb; ++pb) {
            const double sa = sqrt(q[2 * pb]), tb = 2.0 * q[2 * pb + 1];
            ld[2 * pb] = sa * tb;
            ld[2 * pb + 1] = sa * (1.0 - tb);
        }
    } else {
        for (int j = 0; j < P.npb * L.nldc; ++j) ld[j] = q[j];
    }
    if (P.sigma)
        for (int j = 0; j < L.nloge; ++j) P.sigma[(size_t)ipv * L.nloge + j] = pow(10.0, pv[L.i_loge + j]);
}

// ---------------------------------------------------------------------------------------------
// k_dfma_peak -- measured fp64 FMA throughput of this GPU (the roofline denominator of the fused-likelihood kernel,
// whose bound is instruction issue on the fp64 pipe, not memory): 8 independent DFMA chains per thread, no memory
// traffic.  2 flop per DFMA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}

// ---------------------------------------------------------------------------------------------
// Model derivatives dfdk / dfdb (common.py:104-128), as coded: the partial derivatives of the single-sample flux
// F = (I* - l(g) A(k, b)) / I*, g = b / (1 + k), with respect to the radius ratio (at fixed l: dA/dk = 2 k kappa0)
// and to the separation b (dA/db = -2 A_kite / b, dl/db from the two neighbouring LD-mean nodes).  One thread per
// (vector, separation); the vector's LD-mean row, k and I* come from the records of the last evaluation.  The
// interpolation weight is `g - ig*dg` as in the reference (not divided by dg, unlike common.py:231).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rr_derivs(const double *__restrict__ rec, int recstride, int rec_ld, int lds, int ng, int pb,
                                                   double dg, const double *__restrict__ b, int nb, int npv,
                                                   double *__restrict__ dfdk, double *__restrict__ dfdb) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)npv * nb) return;
    const int ipv = (int)(idx / nb);
    const double *row = rec + (size_t)ipv * recstride + rec_ld + (size_t)pb * lds;
    const double k = row[ng], ist = 1.0 / row[ng + 2], z = b[idx];
    double dk = 0.0, db = 0.0;
    if (z < 1.0 + k - 1e-5) {
        const double g = z / (1.0 + k);
        const int ig = min(max((int)floor(g / dg), 0), ng - 1), ig1 = min(ig + 1, ng - 1);
        const double ag = g - ig * dg;
        const double l1 = row[ig], l2 = row[ig1];
        const double l = (1.0 - ag) * l1 + ag * l2;
        // lens area, kappa0 and the kite area (circle_circle_intersection_area_kite, common.py:52-73)
        double area, kap, akite = 0.0;
        kite_area<double>(k, k * k, z, area, kap);
        if (fabs(1.0 - k) < z) {
            const bool kg = k > 1.0;
            const double hi = kg ? k : 1.0, lo = kg ? 1.0 : k;
            const bool c1 = z > hi, c2 = z > lo;
            const double x = c1 ? z : hi, y = c1 ? hi : (c2 ? z : lo), zz = c2 ? lo : z;
            akite = 0.5 * sqrt((x + (y + zz)) * (zz - (x - y)) * (zz + (x - y)) * (x + (y - zz)));
        }
        dk = -2.0 * k * kap * l / ist;                                   // dfdk, common.py:105-113
        if (z >= 0.005) {                                                // dfdb, common.py:117-128
            const double dldb = -(l2 - l1) / (dg * (1.0 + k));
            db = 2.0 * akite * l / (z * ist) + dldb * area / ist;
        }
    }
    if (dfdk) dfdk[idx] = dk;
    if (dfdb) dfdb[idx] = db;
}

// ---------------------------------------------------------------------------------------------
// Secondary eclipse (model_eclipse.py:72-80): flux = pi k^2 - A(1, k, z), averaged over the sub-samples.
// The points kernel has just produced F = 1 - A / pi for a uniform stellar disk at the eclipse geometry
// (ldm = 1, I* = pi), so E = pi (k^2 - 1 + F): one in-place pass.  Out of eclipse F = 1 exactly -> pi k^2.
// ---------------------------------------------------------------------------------------------
__global__ void k_ecl_finish(double *__restrict__ flux, const double *__restrict__ k, long long npt, long long total) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= total) return;
    const double k0 = k[i / npt];
    if (i + 1 < total && (i + 1) / npt == i / npt && ((reinterpret_cast<uintptr_t>(flux + i) & 15) == 0)) {
        double2 f = *reinterpret_cast<double2 *>(flux + i);
        f.x = kPi * ((k0 * k0 - 1.0) + f.x);
        f.y = kPi * ((k0 * k0 - 1.0) + f.y);
        *reinterpret_cast<double2 *>(flux + i) = f;
    } else {
        flux[i] = kPi * ((k0 * k0 - 1.0) + flux[i]);
        if (i + 1 < total) {
            const double k1 = k[(i + 1) / npt];
            flux[i + 1] = kPi * ((k1 * k1 - 1.0) + flux[i + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Eclipse spectroscopy (model_ecspec.py:55-62): flux[ipv, pb, t] = mean_s [1 - (f_pb A_s / pi) / (1 + f_pb k^2)]
// = 1 - c_pb (1 - F[ipv, t]) with c = f / (1 + f k^2) and F the uniform-disk shape 1 - mean_s(A_s) / pi that the
// points kernel has just written.  A streaming expansion: F (one row per vector, L2 resident) is read once per
// channel, the output is written with 16-byte stores.  One CTA per (vector, channel chunk).
// ---------------------------------------------------------------------------------------------
constexpr int ES_CH = 16;
__global__ void __launch_bounds__(256) k_es_expand(const double *__restrict__ shape, const double *__restrict__ fratio,
                                                   const double *__restrict__ k, double *__restrict__ flux, long long npt,
                                                   int npb, int nchunks) {
    const int ipv = blockIdx.x / nchunks, chunk = blockIdx.x - ipv * nchunks;
    const int pb0 = chunk * ES_CH, nch = min(ES_CH, npb - pb0);
    __shared__ double s_c[ES_CH];
    const double k0 = k[ipv];
    if (threadIdx.x < nch) {
        const double f = fratio[(size_t)ipv * npb + pb0 + threadIdx.x];
        s_c[threadIdx.x] = f / (1.0 + f * (k0 * k0));
    }
    __syncthreads();
    const double *F = shape + (size_t)ipv * npt;
    double *out = flux + ((size_t)ipv * npb + pb0) * npt;
    const bool vec = ((npt & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(F) & 15) == 0);
    if (vec) {
        for (long long i = (long long)threadIdx.x * 2; i < npt; i += 512) {
            const double2 f = *reinterpret_cast<const double2 *>(F + i);
            const double d0 = 1.0 - f.x, d1 = 1.0 - f.y;
            for (int c = 0; c < nch; ++c)
                __stcs(reinterpret_cast<double2 *>(out + (size_t)c * npt + i), make_double2(1.0 - s_c[c] * d0, 1.0 - s_c[c] * d1));
        }
    } else {
        for (long long i = threadIdx.x; i < npt; i += 256) {
            const double d0 = 1.0 - F[i];
            for (int c = 0; c < nch; ++c) out[(size_t)c * npt + i] = 1.0 - s_c[c] * d0;
        }
    }
}

}  // namespace ptb
