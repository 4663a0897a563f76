// ptb_ss_kernels.cuh -- the npv x npt pass for supersampled data sets (some light curve has nsamples > 1):
// phase fold, box test, the exposure-averaged flux (model_full.py:76-99) and the optional fused chi^2.
//
// Same decomposition as k_rr_points (ptb_kernels.cuh): a persistent grid whose warps pull items (parameter vector x
// chunk of the time axis) from a global counter, classify the item's time axis against the transit windows (here in 16-point
// cells), fill the untouched cells with 1.0 and fold the touched ones point by point.  What differs is where the time goes:
// with ten sub-samples per point the kernel is bound by instruction issue in the per-sample arithmetic, not by the
// flux store, so this kernel is organised around the sample loop's registers and occupancy:
//
//   * fold and sample evaluation alternate in long PHASES instead of block by block: the fold phase walks touched
//     cells until the warp's queue holds ~SS_QCAP in-box points (4 bytes each: only the point index is queued, the
//     folded time and the light curve are recomputed / looked up by the drain), the drain phase then evaluates them 32 at a time.  Each phase is a real
//     function (`ss_fold`, `ss_drain_phase`: __noinline__) with its own register allocation; what survives a phase
//     boundary is warp-uniform and lives in a small shared-memory frame, so the driver loop keeps nothing alive
//     across the calls: 80 registers, three CTAs of 256 threads per SM (the block-by-block version needed 128, with
//     the fold state spilled around the sample loop);
//   * everything the phases need that does not depend on the item lives in shared memory (a copy of the kernel
//     parameters, the per-light-curve tables, the warp's record slot);
//   * classification works on CELLS of 16 points, not on 64-point blocks (a Kepler long-cadence block spans 1.3 d:
//     42 % of the blocks of a 3.5 d orbit are touched, against 14 % of the cells); touched cells are compacted into a
//     small ring buffer and a fold step takes four of them (eight lanes each, two points per lane, ONE warp-wide
//     compaction for all 64 points); cells that straddle light curves take a separate general path;
//   * the sample step is straight fp64 arithmetic: separation (two Horner quartics, rsqrt + one coupled Newton
//     step: 2^-43), LD-mean node and weight by adding 1.5 * 2^52 (no conversion instructions), lerp as one fused multiply-add on the node
//     difference; a sample on the stellar limb only drops its separation into the lane's shared-memory column;
//   * limb samples (sqrt + 2 atan2 lens area, common.py:52-73) are numbered by a warp scan and evaluated 32 at a
//     time by full warps, each value written back into its owner's column; every lane then adds its column in
//     exposure order.  A point's arithmetic depends on nothing but the point.
#pragma once
#include "ptb_kernels.cuh"

namespace ptb {

// CTA shape.  Measured on C3 (profiles/tools/ss_variants.sh): 256 threads x 3 CTAs (80 registers, 256-entry queue)
// 3.78 ms; 128 x 6 (80 registers) 3.78 ms with the 256-entry queue, 3.91 ms with 128 entries; 128 x 7 (72 registers,
// 128 entries: 28 warps instead of 24) 4.06 ms -- the extra warps do not pay for the spills and the shorter phases.
#ifndef SS_THREADS_
#define SS_THREADS_ 256
#endif
#ifndef PT_MINB_SS2
#define PT_MINB_SS2 3
#endif
constexpr int SS_THREADS = SS_THREADS_;
constexpr int SS_WARPS = SS_THREADS / 32;
#ifndef SS_BRANCHY_TAIL
#define SS_BRANCHY_TAIL 0
#endif
#ifndef SS_QCAP_
#define SS_QCAP_ 256
#endif
constexpr int SS_QCAP = SS_QCAP_;   // in-box points queued per fold phase (a fold step adds up to 64)
constexpr int SS_FRAME = 16;        // ints of the warp's phase frame
constexpr int SS_LIST = 128;        // ring buffer of touched cells (16-bit indices relative to the item's first cell)
constexpr int SS_CPB = PT_BLOCK / SS_CELL;   // cells per 64-point block

// Warp-private shared memory of the supersampled kernel: limb columns, point queue (point indices), hit bitmap
// (one bit per cell), phase frame, ring of touched cells, mbarrier, record slot (+ its float copy in fp32 mode).
__host__ __device__ inline size_t ss_tpart(int ssc, int tsize) { return ((size_t)(ssc * PT_COLS) * tsize + 15) & ~size_t(15); }
__host__ __device__ inline size_t ss_warp_bytes(int ssc, int recstride, int tsize) {
    const size_t rec_t = (tsize == 4) ? (((size_t)recstride * 4 + 15) & ~size_t(15)) : 0;
    return ss_tpart(ssc, tsize) + (size_t)SS_QCAP * 4 + PT_MAXBLK / 8 + SS_FRAME * 4 + SS_LIST * 2 + 16 + (size_t)recstride * 8 + rec_t;
}
// CTA-wide part: a copy of the kernel parameters, then the per-light-curve tables
constexpr size_t SS_PARAM_BYTES = (sizeof(PointsParams) + 127) & ~size_t(127);
__host__ __device__ inline size_t ss_shared_bytes(int nlc, int nfrac, int tsize) {
    return SS_PARAM_BYTES + ((((size_t)nlc) * 8 + ((size_t)nlc + nfrac) * tsize + 3 * (size_t)nlc * 4 + 127) & ~size_t(127));
}

template <typename T>
struct SsWarp {
    unsigned char *base;
    int ssc, recstride;
    __device__ __forceinline__ SsWarp(unsigned char *b, int ssc_, int recstride_) : base(b), ssc(ssc_), recstride(recstride_) {}
    __device__ __forceinline__ T *colz() const { return reinterpret_cast<T *>(base); }   // per-lane columns [ssc][PT_COLS] of limb samples
    __device__ __forceinline__ int *q_ipt() const { return reinterpret_cast<int *>(base + ss_tpart(ssc, (int)sizeof(T))); }
    __device__ __forceinline__ unsigned *hit() const { return reinterpret_cast<unsigned *>(q_ipt() + SS_QCAP); }
    __device__ __forceinline__ volatile int *frame() const { return reinterpret_cast<volatile int *>(hit() + PT_MAXBLK / 32); }
    __device__ __forceinline__ unsigned short *clist() const { return reinterpret_cast<unsigned short *>(hit() + PT_MAXBLK / 32 + SS_FRAME); }
    __device__ __forceinline__ uint64_t *bar() const { return reinterpret_cast<uint64_t *>(clist() + SS_LIST); }
    __device__ __forceinline__ double *rec() const { return reinterpret_cast<double *>(bar() + 2); }
    __device__ __forceinline__ T *rec_t() const { return reinterpret_cast<T *>(rec() + recstride); }   // fp32 mode only
};
// slots of the phase frame (warp-uniform values)
enum : int { FR_WI = 0, FR_LHEAD, FR_LTAIL, FR_QN, FR_DONE, FR_IPV, FR_CHUNK, FR_BBEG, FR_NCL, FR_NEXT_LO, FR_NEXT_HI, FR_ITER };

// the per-light-curve tables behind the parameter copy
template <typename T>
struct SsTables {
    double *sPad;
    T *sEt, *sFrac;
    int *sNs, *sRow, *sEp;
    int nfrac;
    __device__ __forceinline__ SsTables(unsigned char *smem, int nlc, int S, int frac_tab) {
        sPad = reinterpret_cast<double *>(smem + SS_PARAM_BYTES);
        sEt = reinterpret_cast<T *>(sPad + nlc);
        sFrac = sEt + nlc;
        nfrac = frac_tab ? nlc * S : 0;
        sNs = reinterpret_cast<int *>(sFrac + nfrac);
        sRow = sNs + nlc;
        sEp = sRow + nlc;
    }
};

// what every phase function starts with: the parameter copy, the tables and the warp's private area
#define SS_PHASE_PROLOGUE(T)                                                                                                   \
    extern __shared__ __align__(128) unsigned char smem_raw[];                                                                 \
    const PointsParams &P = *reinterpret_cast<const PointsParams *>(smem_raw);                                                 \
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;                                                                \
    const SsTables<T> tb(smem_raw, P.nlc, P.ns_max, P.frac_tab);                                                               \
    const SsWarp<T> ws(smem_raw + ss_shared_bytes(P.nlc, tb.nfrac, (int)sizeof(T)) +                                           \
                           (size_t)warp * ss_warp_bytes(P.ssc, P.recstride, (int)sizeof(T)),                                   \
                       P.ssc, P.recstride);                                                                                    \
    volatile int *fr = ws.frame()

// sum / ns by the straight-line division (<= 1 ulp from the IEEE quotient; CUDA's '/' carries a slow-path call)
__device__ __forceinline__ double ss_mean(double sum, int ns) { return fast_div(sum, (double)ns); }
__device__ __forceinline__ float ss_mean(float sum, int ns) { return sum / (float)ns; }

// ---- fold phase: exact per-point fold + box test over touched cells (model_full.py:88-91) until the queue holds
// more than SS_QCAP - 64 in-box points or the item's cells are exhausted.  The fluxes of the cells' points are set to
// 1.0 here (the drain overwrites the in-box ones); in likelihood mode the out-of-box points add their chi^2 here.
// Returns the lane's chi^2 increment.
template <int VEC, bool SINGLE_LC, bool LNL, typename T>
__device__ __noinline__ double ss_fold() {
    SS_PHASE_PROLOGUE(T);
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned *s_hit = ws.hit();
    unsigned short *clist = ws.clist();
    int *q_ipt = ws.q_ipt();
    int qn = fr[FR_QN];
    const int ipv = fr[FR_IPV], bbeg = fr[FR_BBEG], nwords = (fr[FR_NCL] + 31) >> 5;
    int wi = fr[FR_WI], head = fr[FR_LHEAD], tail = fr[FR_LTAIL];   // next bitmap word; ring positions (monotonic)
    const double *rec = ws.rec();
    const double p = rec[ORB_P], invp = rec[ORB_INVP], T1 = rec[ORB_T1], T4 = rec[ORB_T4];
    const double *t0v = rec + ORB_STRIDE;
    // the item's corner of the arrays: 64-bit addresses once per phase, 32-bit offsets per cell
    const long long pbase = (long long)bbeg * PT_BLOCK;
    const double *tl = P.time + pbase;
    const double *ol = LNL ? P.obs + pbase : nullptr;
    T *fl = LNL ? nullptr : reinterpret_cast<T *>(P.flux) + (size_t)ipv * P.npt + pbase;
    const int32_t *cl = P.clc + bbeg * SS_CPB;
    const long long left = P.npt - pbase;                                   // offsets < rem are inside the time axis
    const int rem = (int)(left > 0x7fffffffll ? 0x7fffffffll : left);     // (VEC == 2 requires an even npt: vectors are all-in or all-out)
    const int ipt0 = (int)pbase;
    const double *isig2 = LNL ? P.isig2 + (size_t)ipv * P.nblocks : nullptr;
    double chi = 0.0;
    // A fold step takes four cells, eight lanes each; a lane folds two points of its cell: A and B are neighbours
    // (VEC == 2, one 16-byte load) or 8 points apart (VEC == 1, odd npt or unaligned arrays).
    constexpr int DB = (VEC == 2) ? 1 : 8;
    const int g = lane >> 3, loff = (lane & 7) * VEC;
    T ones[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) ones[j] = T(1);

    // the set bits of the next bitmap words go to the ring while it has room for a whole word
    auto refill = [&]() {
        if (tail - head <= SS_LIST - 32 && wi < nwords) {
            __syncwarp();   // the slots about to be reused have been read by every lane
            do {
                const unsigned w = s_hit[wi];
                if ((w >> lane) & 1u) clist[(tail + __popc(w & lt_mask)) & (SS_LIST - 1)] = (unsigned short)(wi * 32 + lane);
                tail += __popc(w);
                ++wi;
            } while (tail - head <= SS_LIST - 32 && wi < nwords);
            __syncwarp();
        }
    };
    // the next group of up to four cells with its time stamps (and observed fluxes) in flight
    int celln = -1, cntn = 0, lcn = 0, nzn = 0;   // cells, count, light curve (-1 mixed), (likelihood) noise id (-2 mixed)
    double tnA = 0.0, tnB = 0.0;
    const bool nz_lookup = LNL && P.blk && P.cnoise;
    const int32_t *cnz = nz_lookup ? P.cnoise + bbeg * SS_CPB : nullptr;
    auto fetch = [&]() {
        cntn = min(4, tail - head);
        celln = (g < cntn) ? (int)clist[(head + g) & (SS_LIST - 1)] : -1;
        head += cntn;
        tnA = 0.0; tnB = 0.0; lcn = 0; nzn = 0;
        if (celln >= 0) {
            const int off = celln * SS_CELL + loff;
            if (!SINGLE_LC) lcn = __ldg(cl + celln);
            if (nz_lookup) nzn = __ldg(cnz + celln);
            if (VEC == 2) {
                if (off < rem) {
                    const double2 t = __ldg(reinterpret_cast<const double2 *>(tl + off));
                    tnA = t.x; tnB = t.y;
                }
            } else {
                if (off < rem) tnA = __ldg(tl + off);
                if (off + DB < rem) tnB = __ldg(tl + off + DB);
            }
        }
    };

    double lob = 0.0, hib = 0.0, t0b = 0.0;   // window and transit centre of the lane's cell
    if (SINGLE_LC) {
        const double pd = tb.sPad[0];
        lob = T1 - pd;
        hib = T4 + pd;
        t0b = t0v[tb.sEp[0]];
    }
    refill();
    int head0 = head;   // ring position of the first cell that has not been folded
    fetch();
    while (cntn > 0 && qn <= SS_QCAP - PT_BLOCK) {
        const int cell = celln, lcb = lcn, nzb = nzn;
        const double tA = tnA, tB = tnB;
        head0 = head;
        refill();
        fetch();  // in flight while this group is folded
        const int offA = (cell >= 0 ? cell : 0) * SS_CELL + loff, offB = offA + DB;
        const bool inrA = cell >= 0 && offA < rem, inrB = cell >= 0 && offB < rem;
        bool inA, inB;
        // epoch = floor((t - t0 + p/2)/p); tc = t - (t0 + epoch p)  (model_full.py:88-89).  The division is a
        // multiplication by 1/p: the two can only disagree half a period away from the transit, where the
        // point is outside the box either way.  Always fp64: time stamps need all their digits.
        if (SINGLE_LC || !__any_sync(0xffffffffu, lcb < 0)) {   // the rule: window and transit centre are cell constants
            if (!SINGLE_LC) {
                const double pd = tb.sPad[lcb];
                lob = T1 - pd;
                hib = T4 + pd;
                t0b = t0v[tb.sEp[lcb]];
            }
            const double eA = floor(fma(tA - t0b, invp, 0.5)), eB = floor(fma(tB - t0b, invp, 0.5));
            const double tcA = tA - __dadd_rn(t0b, __dmul_rn(eA, p)), tcB = tB - __dadd_rn(t0b, __dmul_rn(eB, p));
            inA = inrA && (lob <= tcA) && (tcA <= hib);
            inB = inrB && (lob <= tcB) && (tcB <= hib);
        } else {                       // some cell of the group straddles light curves: per-point lookup
            const int lcA = lcb >= 0 ? lcb : (inrA ? P.lcids[(long long)ipt0 + offA] : 0);
            const int lcB = lcb >= 0 ? lcb : (inrB ? P.lcids[(long long)ipt0 + offB] : 0);
            const double pdA = tb.sPad[lcA], t0A = t0v[tb.sEp[lcA]], pdB = tb.sPad[lcB], t0B = t0v[tb.sEp[lcB]];
            const double eA = floor(fma(tA - t0A, invp, 0.5)), eB = floor(fma(tB - t0B, invp, 0.5));
            const double tcA = tA - __dadd_rn(t0A, __dmul_rn(eA, p)), tcB = tB - __dadd_rn(t0B, __dmul_rn(eB, p));
            inA = inrA && (T1 - pdA <= tcA) && (tcA <= T4 + pdA);
            inB = inrB && (T1 - pdB <= tcB) && (tcB <= T4 + pdB);
        }
        if (LNL && nzb == -2) {   // several noise ids in this cell: its baseline, point by point
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
        // one compaction for the warp's 64 points
        const unsigned mA = __ballot_sync(0xffffffffu, inA), mB = __ballot_sync(0xffffffffu, inB);
        int pos = qn + __popc(mA & lt_mask) + __popc(mB & lt_mask);
        qn += __popc(mA) + __popc(mB);
        if (inA) q_ipt[pos++] = ipt0 + offA;   // only the index is queued: the drain looks the light curve up again
        if (inB) q_ipt[pos] = ipt0 + offB;
        // 1.0 for the cells' points, default cache policy (not evict-first): the line is still in L2 when the drain
        // updates its in-box points, so it reaches DRAM once
        if (!LNL) {
            if (VEC == 2) {
                if (inrA) VecIO<VEC, T>::store_keep(fl + offA, ones);
            } else {
                if (inrA) fl[offA] = T(1);
                if (inrB) fl[offB] = T(1);
            }
        }
    }
    __syncwarp();   // every lane has read the frame (and left the ring) before lane 0 rewrites it
    if (lane == 0) {
        fr[FR_QN] = qn;
        fr[FR_DONE] = cntn == 0;                  // nothing fetched: ring and bitmap are exhausted
        fr[FR_WI] = wi;
        fr[FR_LHEAD] = cntn > 0 ? head0 : head;   // a group fetched but not folded (queue full) is fetched again
        fr[FR_LTAIL] = tail;
    }
    return chi;
}

// The drain of `n` (<= 32) queued in-box points starting at queue slot `base`, one point per lane; the lane folds its
// point again (the same fp64 operations as the fold phase) and walks its exposure sub-samples in order
// (model_full.py:88-99).  Returns the lane's chi^2 increment (likelihood); the flux goes to the row of vector `ipv`.
template <bool SINGLE_LC, bool LNL, typename T>
__device__ __forceinline__ double ss_drain(const PointsParams &P, const SsTables<T> &tb, const SsWarp<T> &ws, int ipv, int base,
                                           int n, int lane) {
    const int ng = P.ng, S = P.ns_max, SSC = P.ssc;
    const double *rec = ws.rec();
    const T *rt = (sizeof(T) == 4) ? ws.rec_t() : reinterpret_cast<const T *>(rec);
    const T *ld = rt + P.rec_ld;
    T *colz = ws.colz();
    const T inv_dg = (T)P.inv_dg, one = T(1), pi = T(kPi);

    const bool valid = lane < n;
    int ipt = 0, lc = 0, ns = 0, rowoff = tb.sRow[0];
    T tc = T(0), et = T(0);
    if (valid) {
        ipt = ws.q_ipt()[base + lane];
        if (!SINGLE_LC) {   // the point's light curve: its cell's, or (cell straddling light curves) its own
            lc = __ldg(P.clc + (ipt >> 4));
            if (lc < 0) lc = __ldg(P.lcids + ipt);
            rowoff = tb.sRow[lc];
        }
        ns = tb.sNs[lc];
        et = tb.sEt[lc];
        // epoch = floor((t - t0 + p/2)/p); tc = t - (t0 + epoch p)  (model_full.py:88-89), as in the fold phase
        const double t = __ldg(P.time + ipt), t0 = rec[ORB_STRIDE + tb.sEp[lc]];
        const double epoch = floor(fma(t - t0, rec[ORB_INVP], 0.5));
        tc = (T)(t - __dadd_rn(t0, __dmul_rn(epoch, rec[ORB_P])));
    }
    __syncwarp();   // queue and frame are read
    const T *row = ld + rowoff;
    const T k = row[ng], inv1k = row[ng + 1], inv_istar = row[ng + 2], k2 = row[ng + 3];
    // per-point constants of the area cases that need no lens formula (common.py:52-73)
    const T zout = one + k, zin = fabs(one - k);
    const T qfull = ((k > one) ? pi : pi * k2) * inv_istar;   // planet covers the star: area pi; else pi k^2
    const T c_out = fma(T(0), inv_istar, one);                 // no overlap: (I* - 0) / I* = 1, NaN when 1/I* is NaN
    const T xs = inv1k * inv_dg;                               // separation -> position on the LD-mean grid
    const T *off = tb.nfrac ? tb.sFrac + lc * S : nullptr;
    T cx[5], cy[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { cx[j] = rt[j]; cy[j] = rt[5 + j]; }

    T sum = T(0);
    for (int s0 = 0; s0 < S; s0 += SSC) {
        const int SS = min(min(SSC, S - s0), ns - s0);   // ns = 0 for lanes without a point
        T *mz = colz + lane;
        // one exposure sub-sample at offset `o` from the point's folded time
        auto step = [&](T o) {
            const T t = tc + o;
            const T px = fma(t, fma(t, fma(t, fma(t, cx[4], cx[3]), cx[2]), cx[1]), cx[0]);
            const T py = fma(t, fma(t, fma(t, fma(t, cy[4], cy[3]), cy[2]), cy[1]), cy[0]);
            const T z = sqrt_sep(fma(px, px, py * py));      // sep_c (taylor_z.py:229-255)
            const T ip = ld_lerp(z * xs, row, ng);
#if SS_BRANCHY_TAIL
            if (zout <= z) {
                sum += c_out;
            } else if (zin < z) {
                *mz = z;
                mz += PT_COLS;
            } else {
                sum += fma(-ip, qfull, one);
            }
#else
            // straight-line tail (selects and predicated stores: consecutive steps interleave): no overlap, limb
            // (deferred to the lens-area passes), or full overlap -- a NaN separation lands there with a NaN weight
            const bool out = zout <= z, limb = !out && (zin < z);
            const T v = out ? c_out : fma(-ip, qfull, one);
            if (!limb) sum += v;
            if (limb) {
                *mz = z;
                mz += PT_COLS;
            }
#endif
        };
        // exposure offset exptime*((s+1-0.5)/ns - 0.5) (model_full.py:94): tabulated per light curve, or on the fly
        if (off) {
            const T *o = off + s0;
            for (int j = 0; j < SS; ++j) step(o[j]);
        } else {
            for (int j = 0; j < SS; ++j) step((T)__dmul_rn((double)et, ((s0 + j + 1) - 0.5) / ns - 0.5));
        }
        const int cnt = (int)((smem_u32(mz) - smem_u32(colz + lane)) / (unsigned)(PT_COLS * sizeof(T)));   // this lane's limb samples of the pass
        // number the warp's limb samples: inclusive scan of the per-lane counts
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        // owner (lane) and row (position in the owner's column) of limb sample q: the owner is the number of lanes whose
        // inclusive count is <= q
        auto locate = [&](int q, int &o, int &r, int &ro) {
            o = 0;
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, o + st - 1);
                if (v <= q) o += st;
            }
            o = min(o, 31);
            r = q - (__shfl_sync(0xffffffffu, incl, o) - __shfl_sync(0xffffffffu, cnt, o));
            ro = SINGLE_LC ? 0 : __shfl_sync(0xffffffffu, rowoff, o);
        };
        for (int q0 = 0; q0 < total; q0 += 32) {
            const int q = q0 + lane;
            int o, r, ro;
            locate(q, o, r, ro);
            if (q < total) {
                const T *r2 = SINGLE_LC ? row : ld + ro;
                T *slot = colz + r * PT_COLS + o;
                const T z = *slot;
                const T ip = ld_lerp(z * (r2[ng + 1] * inv_dg), r2, ng);
                *slot = one - ip * kite_area_limb<T>(r2[ng], r2[ng + 3], z) * r2[ng + 2];
            }
        }
        __syncwarp();
        // this lane's limb values, in exposure order (fixed trip count, predicated: the per-lane counts differ)
#pragma unroll
        for (int r = 0; r < PT_SSC_MAX; ++r)
            if (r < cnt) sum += colz[r * PT_COLS + lane];
        __syncwarp();
    }
    double chi = 0.0;
    if (valid) {
        const T f = ss_mean(sum, ns);   // flux / nsamples (model_full.py:99)
        if (LNL) {
            const int b = P.blk ? P.blk[ipt] : 0;
            if (b >= 0) {   // the point's (obs - 1)^2 is already in the cell baseline: swap it for (obs - model)^2
                const double o = P.obs[ipt], d1 = o - (double)f, d0 = o - 1.0;
                chi = fma(d1, d1, -d0 * d0) * (P.isig2 + (size_t)ipv * P.nblocks)[P.blk ? b : 0];
            }
        } else {
            (reinterpret_cast<T *>(P.flux) + (size_t)ipv * P.npt)[ipt] = f;
        }
    }
    return chi;
}

// ---- drain phase: full warps of points from the top of the queue; a remainder (< 32) waits for the next fold phase
// unless that was the item's last one.  Returns the lane's chi^2 increment.
template <bool SINGLE_LC, bool LNL, typename T>
__device__ __noinline__ double ss_drain_phase() {
    SS_PHASE_PROLOGUE(T);
    int qn = fr[FR_QN];
    const int done = fr[FR_DONE], ipv = fr[FR_IPV];
    double chi = 0.0;
    while (qn >= 32 || (done && qn > 0)) {
        const int n = min(qn, 32);
        qn -= n;
        chi += ss_drain<SINGLE_LC, LNL, T>(P, tb, ws, ipv, qn, n, lane);
    }
    __syncwarp();   // every lane has read the frame before lane 0 rewrites it
    if (lane == 0) fr[FR_QN] = qn;
    return chi;
}

template <int VEC, bool SINGLE_LC, bool LNL, typename T>
__global__ void __launch_bounds__(SS_THREADS, PT_MINB_SS2) k_rr_points_ss(const __grid_constant__ PointsParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int NH = 2 / VEC;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long npt = P.npt;
    const int S = P.ns_max, nlc = P.nlc;
    const long long nitems = (long long)P.npv * P.nchunks;

    // dynamic smem: [parameter copy] [per-light-curve tables] [sub-sample offsets] [warp-private area x 8]
    {
        const int *src = reinterpret_cast<const int *>(&P);
        int *dst = reinterpret_cast<int *>(smem_raw);
        for (int i = tid; i < (int)(sizeof(PointsParams) / 4); i += SS_THREADS) dst[i] = src[i];
    }
    const SsTables<T> tb(smem_raw, nlc, S, P.frac_tab);
    double *sPad = tb.sPad;
    int *sEp = tb.sEp;
    const SsWarp<T> ws(smem_raw + ss_shared_bytes(nlc, tb.nfrac, (int)sizeof(T)) + (size_t)warp * ss_warp_bytes(P.ssc, P.recstride, (int)sizeof(T)),
                       P.ssc, P.recstride);
    unsigned *s_hit = ws.hit();
    uint64_t *bar = ws.bar();
    volatile int *fr = ws.frame();

    const uint32_t rec_bytes = (uint32_t)P.recstride * 8u;
    long long item = 0;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        item = atomicAdd(&P.work[0], 1);
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    // item-independent per-light-curve tables
    for (int lc = tid; lc < nlc; lc += SS_THREADS) {
        sPad[lc] = 0.003 + P.exptimes[lc];  // model_full.py:69-70
        tb.sEt[lc] = (T)P.exptimes[lc];
        tb.sNs[lc] = P.nsamples[lc];
        tb.sRow[lc] = P.pbids[lc] * P.lds;
        sEp[lc] = P.epids[lc];
    }
    for (int i = tid; i < tb.nfrac; i += SS_THREADS) {   // exptimes[ilc]*((isample-0.5)/nsamples[ilc] - 0.5), model_full.py:94
        const int lc = i / S, s = i - lc * S;
        tb.sFrac[i] = (T)__dmul_rn(P.exptimes[lc], ((s + 1) - 0.5) / P.nsamples[lc] - 0.5);
    }
    __syncthreads();  // the only CTA-wide barrier: from here on the warps are independent workers

    int iter = 0;
    while (item < nitems) {
        const double *rec = ws.rec();
        double chi = 0.0;
        {   // ---- item header; what the phases and the item's end need is parked in the frame ------------------
            // this item's record: one TMA bulk copy into the warp's slot (every lane left the slot at the end of the
            // last item; the items are tens of thousands of sub-samples long, the fetch is noise)
            long long next = 0;
            if (lane == 0) {
                mbar_expect_tx(&bar[0], rec_bytes);
                tma_load_1d(ws.rec(), P.rec + (size_t)(item / P.nchunks) * P.recstride, rec_bytes, &bar[0]);
                next = atomicAdd(&P.work[0], 1);
            }
            next = __shfl_sync(0xffffffffu, next, 0);
            const int ipv = (int)(item / P.nchunks);
            const int chunk = (int)(item - (long long)ipv * P.nchunks);
            const int bbeg = chunk * P.blocks_per_chunk;
            const int bend = min(P.nblk64, bbeg + P.blocks_per_chunk);
            const int cbeg = bbeg * SS_CPB;                      // the item's cells
            const int ncl = min(P.ncell, bend * SS_CPB) - cbeg;
            if (lane == 0) {
                fr[FR_IPV] = ipv; fr[FR_CHUNK] = chunk; fr[FR_BBEG] = bbeg; fr[FR_NCL] = ncl;
                fr[FR_NEXT_LO] = (int)(unsigned)(next & 0xffffffffll); fr[FR_NEXT_HI] = (int)(next >> 32);
                fr[FR_ITER] = iter + 1;
                fr[FR_WI] = 0; fr[FR_LHEAD] = 0; fr[FR_LTAIL] = 0; fr[FR_QN] = 0; fr[FR_DONE] = 0;
            }
            T *frow = LNL ? nullptr : reinterpret_cast<T *>(P.flux) + (size_t)ipv * npt;
            const double *isig2 = LNL ? P.isig2 + (size_t)ipv * P.nblocks : nullptr;

            mbar_wait(&bar[0], iter & 1);
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
                item = next;
                ++iter;
                continue;
            }
            if (F32) {   // the record in the arithmetic type, converted once per item
                T *dst = ws.rec_t();
                for (int i = lane; i < P.recstride; i += 32) dst[i] = (T)rec[i];
            }

            const double p = rec[ORB_P], invp = rec[ORB_INVP], T1 = rec[ORB_T1], T4 = rec[ORB_T4];
            const double *t0v = rec + ORB_STRIDE;
            const double w_one = (LNL && !P.blk) ? isig2[0] : 0.0;  // single noise block: its weight is an item constant

            // ---- 1. classification of every cell of this item (see k_rr_points) -------------------------------
            for (int cc0 = 0; cc0 < ncl; cc0 += 32) {
                const int cc = cc0 + lane, c = cbeg + cc;
                bool hit = false;
                if (cc < ncl) {
                    hit = true;
                    const int lcb = SINGLE_LC ? 0 : P.clc[c];
                    int nz_id = 0;
                    if (LNL && P.blk) nz_id = P.cnoise ? P.cnoise[c] : 0;
                    const bool partial = (c == P.ncell - 1) && (npt % SS_CELL != 0);
                    if (lcb >= 0 && nz_id != -2 && !partial) {
                        const double pd = sPad[lcb], t0 = t0v[sEp[lcb]];
                        const double n1 = ceil(fma(P.cmin[c] - t0 - (T4 + pd), invp, -PT_EPS));
                        const double n2 = floor(fma(P.cmax[c] - t0 - (T1 - pd), invp, PT_EPS));
                        hit = !(n1 > n2) || !(p > 0.0);  // NaNs and p <= 0 fall through to the exact per-point path
                    }
                    // likelihood baseline: sum of (obs - 1)^2 of EVERY cell that has one noise id, touched or not (the
                    // drain swaps the in-box points' terms for (obs - model)^2); cells with several ids: per point, in ss_fold
                    if (LNL && nz_id >= 0) chi = fma(P.cchi[c], P.blk ? isig2[nz_id] : w_one, chi);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_hit[cc0 >> 5] = m;
            }
            __syncwarp();

            // ---- 2. untouched cells: 16 fluxes of exactly 1.0 each, vectorised streaming stores; a bitmap word
            //         covers 32 cells = 512 points = eight 64-point rows -------------------------------------------
            if (!LNL) {
                T one[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) one[j] = T(1);
                const int nw = (ncl + 31) >> 5;
                T *f0 = frow + (long long)cbeg * SS_CELL + lane * VEC;
                for (int w = 0; w < nw; ++w) {
                    const unsigned bits = s_hit[w];
                    T *fw = f0 + w * (32 * SS_CELL);
                    if (bits == 0u && (w + 1) * 32 <= ncl) {  // the common case: 32 untouched cells, no predicates
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
#pragma unroll
                            for (int h = 0; h < NH; ++h) VecIO<VEC, T>::store(fw + j * PT_BLOCK + h * 32 * VEC, one);
                        }
                    } else if (bits != 0xffffffffu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
#pragma unroll
                            for (int h = 0; h < NH; ++h) {
                                const int ci = (j * PT_BLOCK + h * 32 * VEC + lane * VEC) / SS_CELL;   // cell within the word
                                if (!((bits >> ci) & 1u) && w * 32 + ci < ncl) VecIO<VEC, T>::store(fw + j * PT_BLOCK + h * 32 * VEC, one);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- 3. touched blocks: fold and drain phases alternate until the item's blocks are exhausted ------------
        do {
            chi += ss_fold<VEC, SINGLE_LC, LNL, T>();
            __syncwarp();
            chi += ss_drain_phase<SINGLE_LC, LNL, T>();
            __syncwarp();
        } while (!fr[FR_DONE]);

        if (LNL) {
            chi = warp_sum(chi);
            if (lane == 0) P.partial[(size_t)fr[FR_IPV] * P.nchunks + fr[FR_CHUNK]] = chi;
        }
        item = ((long long)fr[FR_NEXT_HI] << 32) | (long long)(unsigned)fr[FR_NEXT_LO];
        iter = fr[FR_ITER];
        __syncwarp();  // every lane is done with this record slot and frame before the next item refills them
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

}  // namespace ptb
