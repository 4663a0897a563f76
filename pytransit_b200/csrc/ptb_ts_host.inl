// ptb_ts_host.inl -- host side of the TSModel and tabulated-profile entry points (included by
// ptb200.cu).

namespace {

template <int NT>
int launch_ts_ldm_t(ptb_model *h, const TsLdmParams &P, size_t smem, unsigned grid, cudaStream_t st) {
    auto kern = k_ts_ldm<NT>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(P);
    h->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

template <int VEC, typename TO, bool FUSED, int NT>
int launch_ts_flux2_t(ptb_model *h, const TsFlux2Params &P, size_t smem, unsigned grid, cudaStream_t st) {
    auto kern = k_ts_flux2<VEC, TO, FUSED, NT>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(P);
    h->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

template <int VEC, bool MULTI, typename TO>
int launch_ts_flux_t(ptb_model *h, const TsFluxParams &P, size_t smem, unsigned grid, cudaStream_t st) {
    auto kern = k_ts_flux<VEC, MULTI, TO>;
    if (smem > 32 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(P);
    h->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

}  // namespace

extern "C" {

int ptb_ts_evaluate(ptb_model *h, int64_t npv, int64_t npb, const double *k, const double *ld, int64_t nld,
                    const double *istar, const double *t0, const double *p, const double *a, const double *inc,
                    const double *e, const double *w, void *flux, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (npb < 1) return fail(h, PTB_ESHAPE, "ts_evaluate: npb must be >= 1");
    ModelArgs A{npv, npb, nld, k, ld, istar, t0, p, a, inc, e, w};
    if (int rc = check_model_args(h, "ts_evaluate", A, npb)) return rc;
    const int ng = h->cfg.ng, nz = h->nz;
    if (nz % 2) return fail(h, PTB_EINVAL, "ts_evaluate: nz = nzin + nzlimb must be even (TMA rows are 16-byte multiples)");
    if (ng > 128) return fail(h, PTB_EINVAL, "ts_evaluate: ng > 128 is not supported by the contraction kernel");
    const int ns = (int)h->h_nsamples[0];
    const double exptime = h->h_exptimes[0];
    if (ns > 24) return fail(h, PTB_EINVAL, "ts_evaluate: nsamples=%d > 24 is not supported", ns);
    if ((double)npv * (double)npb * (double)h->npt >= 9.0e18) return fail(h, PTB_EINVAL, "ts_evaluate: output too large");

    Staged D{};
    if (int rc = stage_model_args(h, A, npb, 1, nullptr, 0, st, D)) return rc;
    if (h->xyc_injected && h->xyc_npv != npv)
        return fail(h, PTB_ESHAPE, "injected xyc has npv=%lld but evaluate was called with npv=%lld", (long long)h->xyc_npv, (long long)npv);

    const int NTs[] = {4, 8, 13, 16};
    int nt = 16;
    for (int c : NTs) if (c * 8 >= ng) { nt = c; break; }
    const int ldt = nt * 8;
    const int rs = ((nz + 3) / 4) * 4 + 4;

    CU(h->d_orb.reserve((size_t)npv * TSORB_STRIDE * 8));
    CU(h->d_tsw.reserve((size_t)npv * ng * rs * 8));
    CU(h->d_ldrec.reserve((size_t)npv * npb * ldt * 8));
    CU(h->d_tsrec.reserve((size_t)npv * npb * 4 * 8));

    // 1. per-vector setup
    TsSetupParams SP{};
    SP.k = D.k; SP.p = D.p; SP.a = D.a; SP.inc = D.inc; SP.e = D.e; SP.w = D.w;
    SP.xyc_in = h->xyc_injected ? h->d_xyc.as<double>() : nullptr;
    SP.W = h->d_W.as<double>(); SP.ze = h->d_ze; SP.gs = h->d_gs;
    SP.tsorb = h->d_orb.as<double>(); SP.tsw = h->d_tsw.as<double>();
    SP.npv = (int)npv; SP.npb = (int)npb; SP.nk = h->cfg.nk; SP.ng = ng; SP.nz = nz;
    SP.use_table = h->cfg.precompute_weights ? 1 : 0;
    SP.rs = rs;
    SP.kmin = h->cfg.kmin; SP.dk = h->dk;
    mark(h, 0, st);
    const size_t smem_setup = SP.use_table ? 0 : (size_t)ng * (nz + 1) * 8;   // direct weights are built in shared memory
    if (smem_setup > 48 * 1024) CU(cudaFuncSetAttribute(k_ts_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_setup));
    k_ts_setup<<<(unsigned)npv, TSS_THREADS, smem_setup, st>>>(SP);
    h->launches++;
    CU(cudaGetLastError());

    // 2. limb-darkening profiles (named laws) or the caller's tabulated profiles
    const double *ldp = D.ld, *ist = D.istar;
    if (h->cfg.ldlaw != PTB_LD_PROFILES) {
        CU(h->d_ldp.reserve((size_t)npv * npb * nz * 8));
        CU(h->d_istar.reserve((size_t)npv * npb * 8));
        TsLdParams LP{};
        LP.ldc = D.ld; LP.mu = h->d_mu; LP.ldmu200 = h->d_ldmu; LP.ldz200 = h->d_ldz;
        LP.ldp = h->d_ldp.as<double>(); LP.istar = h->d_istar.as<double>();
        LP.nrows = (long long)npv * npb; LP.nld = (int)nld; LP.law = h->cfg.ldlaw; LP.nz = nz;
        k_ts_ld<<<(unsigned)((LP.nrows + 3) / 4), 128, 0, st>>>(LP);
        h->launches++;
        CU(cudaGetLastError());
        ldp = LP.ldp;
        ist = LP.istar;
    } else if (reinterpret_cast<uintptr_t>(ldp) & 15) {
        // TMA sources must be 16-byte aligned: realign a misaligned caller tensor once
        CU(h->d_ldp.reserve((size_t)npv * npb * nz * 8));
        CU(cudaMemcpyAsync(h->d_ldp.ptr, ldp, (size_t)npv * npb * nz * 8, cudaMemcpyDeviceToDevice, st));
        ldp = h->d_ldp.as<double>();
    }

    // 3. dense LD contraction (DMMA): fused into the flux kernel's prologue in the common case (one sample per point,
    //    default grid) -- the ld means then never touch HBM (0.83 GB written + 0.91 GB read back at C4) and a launch goes
    const bool multi = ns > 1;
    const size_t geo_bytes = (size_t)npv * h->npt * 28;
    const bool two_pass = !multi && geo_bytes <= (size_t)2 << 30;
    static const bool fuse_env = [] {
        const char *e = getenv("PTB_TS_FUSE");
        return !(e && atoi(e) == 0);
    }();
    const size_t smem_fused = std::max((size_t)ng * rs + (size_t)TS_CH * nz, (size_t)TS_CH * (ldt + 4)) * 8;
    const bool fused = fuse_env && two_pass && nt == 13 && (nz & 3) == 0 && smem_fused <= 200 * 1024;
    int rc = PTB_OK;
    if (!fused) {
        TsLdmParams MP{};
        MP.tsw = h->d_tsw.as<double>(); MP.ldp = ldp; MP.istar = ist; MP.k = D.k; MP.tsorb = h->d_orb.as<double>();
        MP.tsldm = h->d_ldrec.as<double>(); MP.tsrec = h->d_tsrec.as<double>();
        MP.npv = (int)npv; MP.npb = (int)npb; MP.ng = ng; MP.nz = nz; MP.ldt = ldt; MP.rs = rs;
        const size_t smem_ldm = (size_t)(nt * 8 + 2 * TSL_PB) * rs * 8;
        const long long ntile_ldm = (npb + TSL_PB - 1) / TSL_PB;
        MP.tpc = (int)std::min<long long>(4, ntile_ldm);
        const unsigned grid_ldm = (unsigned)(npv * ((ntile_ldm + MP.tpc - 1) / MP.tpc));
        switch (nt) {
        case 4: rc = launch_ts_ldm_t<4>(h, MP, smem_ldm, grid_ldm, st); break;
        case 8: rc = launch_ts_ldm_t<8>(h, MP, smem_ldm, grid_ldm, st); break;
        case 13: rc = launch_ts_ldm_t<13>(h, MP, smem_ldm, grid_ldm, st); break;
        default: rc = launch_ts_ldm_t<16>(h, MP, smem_ldm, grid_ldm, st); break;
        }
        if (rc) return rc;
    }
    mark(h, 1, st);

    // 4. flux
    const size_t count = (size_t)npv * npb * h->npt;
    // opt-in fp32 output mode (ptb_config.precision = 1): geometry and per-channel arithmetic stay fp64, the flux is
    // stored as float -- half the HBM and PCIe bytes of the 8 B/point write-out that bounds this model
    const bool f32 = h->cfg.precision == 1;
    const size_t esize = f32 ? 4 : 8;
    void *dflux = flux;
    const bool direct = flux && is_device_ptr(flux);
    if (!direct) {
        CU(h->d_flux.reserve(count * esize));
        dflux = h->d_flux.ptr;
    }
    const bool aligned = !multi && (h->npt % 2 == 0) && ((reinterpret_cast<uintptr_t>(h->d_time) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(dflux) & (f32 ? 7 : 15)) == 0);
    const int vec = aligned ? 2 : 1;
    if (two_pass) {
        // ---- one sample per point: geometry pass + channel-chunk flux pass ------------------------------
        const size_t npts = (size_t)npv * h->npt;
        CU(h->d_tsgeo.reserve(geo_bytes + 64));
        double *ga = h->d_tsgeo.as<double>();
        TsGeoParams GP{};
        GP.time = h->d_time; GP.tsorb = h->d_orb.as<double>(); GP.t0 = D.t0;
        GP.galpha = ga; GP.gap0 = ga + npts; GP.gdadk = ga + 2 * npts; GP.gi0 = reinterpret_cast<int *>(ga + 3 * npts);
        GP.npt = h->npt; GP.npv = (int)npv; GP.ng = ng; GP.exptime = exptime; GP.dg = h->dg; GP.inv_dg = 1.0 / h->dg;
        const long long gtiles = (h->npt + 256LL * vec - 1) / (256LL * vec);
        if (npv * gtiles > 0x7fffffffLL) return fail(h, PTB_EINVAL, "ts_evaluate: grid too large");
        if (vec == 2) k_ts_geo<2><<<(unsigned)(npv * gtiles), 256, 0, st>>>(GP);
        else k_ts_geo<1><<<(unsigned)(npv * gtiles), 256, 0, st>>>(GP);
        h->launches++;
        CU(cudaGetLastError());
        TsFlux2Params FP{};
        FP.tsorb = h->d_orb.as<double>(); FP.tsldm = h->d_ldrec.as<double>(); FP.tsrec = h->d_tsrec.as<double>();
        FP.galpha = GP.galpha; FP.gap0 = GP.gap0; FP.gdadk = GP.gdadk; FP.gi0 = GP.gi0; FP.flux = dflux;
        FP.npt = h->npt; FP.npv = (int)npv; FP.npb = (int)npb; FP.ng = ng; FP.ldt = ldt;
        FP.nchunks = (int)((npb + TS_CH - 1) / TS_CH);
        FP.tsw = h->d_tsw.as<double>(); FP.ldp = ldp; FP.istar = ist; FP.k = D.k; FP.nz = nz; FP.rs = rs;
        const long long grid = (long long)npv * FP.nchunks;
        if (grid > 0x7fffffffLL) return fail(h, PTB_EINVAL, "ts_evaluate: grid too large");
        const size_t smem_fl = fused ? smem_fused : (size_t)TS_CH * (ldt + 4) * 8;
        mark(h, 2, st);
#define PTB_TS_FLUX2(V, T)                                                                         \
    (fused ? launch_ts_flux2_t<V, T, true, 13>(h, FP, smem_fl, (unsigned)grid, st)                \
           : launch_ts_flux2_t<V, T, false, 16>(h, FP, smem_fl, (unsigned)grid, st))
        if (vec == 2) rc = f32 ? PTB_TS_FLUX2(2, float) : PTB_TS_FLUX2(2, double);
        else rc = f32 ? PTB_TS_FLUX2(1, float) : PTB_TS_FLUX2(1, double);
#undef PTB_TS_FLUX2
        if (rc) return rc;
        mark(h, 3, st);
    } else {
        // ---- supersampled (or very large) case: geometry in registers / shared memory per CTA ------------
        TsFluxParams FP{};
        FP.time = h->d_time; FP.tsorb = h->d_orb.as<double>(); FP.t0 = D.t0; FP.tsldm = h->d_ldrec.as<double>();
        FP.tsrec = h->d_tsrec.as<double>(); FP.flux = dflux; FP.npt = h->npt; FP.npv = (int)npv; FP.npb = (int)npb;
        FP.ng = ng; FP.ldt = ldt; FP.ns = ns; FP.exptime = exptime; FP.dg = h->dg; FP.inv_dg = 1.0 / h->dg;
        const long long tile = 256LL * vec;
        FP.ntiles = (int)((h->npt + tile - 1) / tile);
        const long long base_ctas = (long long)npv * FP.ntiles;
        const long long want = (long long)h->sm_count * 16;
        FP.pbsplit = (int)std::min<long long>(std::max<long long>(1, npb / 32), std::max<long long>(1, (want + base_ctas - 1) / base_ctas));
        const long long grid = base_ctas * FP.pbsplit;
        if (grid > 0x7fffffffLL) return fail(h, PTB_EINVAL, "ts_evaluate: grid too large");
        const size_t smem_fl = multi ? (size_t)ns * 256 * sizeof(TsGeo) : 0;
        mark(h, 2, st);
        if (f32) {
            if (vec == 2) rc = launch_ts_flux_t<2, false, float>(h, FP, smem_fl, (unsigned)grid, st);
            else if (!multi) rc = launch_ts_flux_t<1, false, float>(h, FP, smem_fl, (unsigned)grid, st);
            else rc = launch_ts_flux_t<1, true, float>(h, FP, smem_fl, (unsigned)grid, st);
        } else {
            if (vec == 2) rc = launch_ts_flux_t<2, false, double>(h, FP, smem_fl, (unsigned)grid, st);
            else if (!multi) rc = launch_ts_flux_t<1, false, double>(h, FP, smem_fl, (unsigned)grid, st);
            else rc = launch_ts_flux_t<1, true, double>(h, FP, smem_fl, (unsigned)grid, st);
        }
        if (rc) return rc;
        mark(h, 3, st);
    }
    h->last_npv = 0;  // RoadRunner stage taps do not describe a TS evaluation
    h->last_flux_count = direct ? 0 : (int64_t)count;
    if (flux && !direct) return deliver_host(h, flux, dflux, count, esize, st);
    return PTB_OK;
}

int ptb_ldtk_profiles(ptb_model *h, const double *profiles, int64_t nx, int64_t ny, int64_t nz3, int64_t npb,
                      int64_t nmu, const double *xs, const double *ys, const double *zs, int64_t npv, double x0,
                      double dx, double y0, double dy, double z0, double dz, const double *mu, double *ldp,
                      double *istar, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!profiles || !xs || !ys || !zs || !mu || !ldp || !istar) return fail(h, PTB_EINVAL, "ldtk_profiles: null argument");
    if (nx < 1 || ny < 1 || nz3 < 1 || npb < 1 || nmu < 2 || npv < 1) return fail(h, PTB_ESHAPE, "ldtk_profiles: bad shape");
    Stager S(h, st);
    auto rp = S.add(profiles, (size_t)nx * ny * nz3 * npb * nmu * 8);
    auto rx = S.add(xs, npv * 8), ry = S.add(ys, npv * 8), rz = S.add(zs, npv * 8), rm = S.add(mu, nmu * 8);
    if (int rc = S.commit()) return rc;
    const bool ldp_dev = is_device_ptr(ldp), is_dev = is_device_ptr(istar);
    double *dldp = ldp, *dis = istar;
    if (!ldp_dev) {
        CU(h->d_ldp.reserve((size_t)npv * npb * nmu * 8));
        dldp = h->d_ldp.as<double>();
    }
    if (!is_dev) {
        CU(h->d_istar.reserve((size_t)npv * npb * 8));
        dis = h->d_istar.as<double>();
    }
    LdtkParams P{};
    P.profiles = S.get<double>(rp); P.xs = S.get<double>(rx); P.ys = S.get<double>(ry); P.zs = S.get<double>(rz);
    P.mu = S.get<double>(rm); P.ldp = dldp; P.istar = dis; P.npv = npv; P.nx = (int)nx; P.ny = (int)ny; P.nz3 = (int)nz3;
    P.npb = (int)npb; P.nmu = (int)nmu; P.x0 = x0; P.dx = dx; P.y0 = y0; P.dy = dy; P.z0 = z0; P.dz = dz;
    // Slab kernel when a few channels of the whole table fit in shared memory (chunk <= 32 channels, <= 96 KB
    // so that two CTAs share an SM) and the population is large enough to amortise staging the table;
    // otherwise one warp per (vector, channel) row reading the 8 nodes from L2.
    const size_t node_bytes = (size_t)nmu * 8, nodes = (size_t)nx * ny * nz3;
    const int chunk = (int)std::min<size_t>({(size_t)32, (size_t)npb, (96 * 1024) / std::max<size_t>(1, nodes * node_bytes)});
    if (chunk >= 1 && npv >= 64 && nodes * nmu < ((size_t)1 << 30)) {
        CU(h->d_cells.reserve((size_t)npv * sizeof(LdtkCell)));
        LdtkCell *cells = h->d_cells.as<LdtkCell>();
        k_ldtk_cells<<<(unsigned)((npv + 127) / 128), 128, 0, st>>>(P, cells);
        const int nchunks = (int)((npb + chunk - 1) / chunk);
        const long long want = (long long)h->sm_count * 4;
        const int vsplit = (int)std::max<long long>(1, std::min<long long>((npv + 63) / 64, (want + nchunks - 1) / nchunks));
        const size_t smem = (nodes * chunk * nmu + ((nmu + 1) & ~(size_t)1) + nodes * chunk) * 8;
        CU(cudaFuncSetAttribute(k_ldtk_profiles_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ldtk_profiles_slab<<<(unsigned)(nchunks * vsplit), LDS_THREADS, smem, st>>>(P, cells, chunk, vsplit);
        h->launches += 2;
    } else {
        const long long rows = (long long)npv * npb;
        k_ldtk_profiles<<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(P);
        h->launches++;
    }
    CU(cudaGetLastError());
    if (!ldp_dev) CU(cudaMemcpyAsync(ldp, dldp, (size_t)npv * npb * nmu * 8, cudaMemcpyDeviceToHost, st));
    if (!is_dev) CU(cudaMemcpyAsync(istar, dis, (size_t)npv * npb * 8, cudaMemcpyDeviceToHost, st));
    if (!ldp_dev || !is_dev) CU(cudaStreamSynchronize(st));
    return PTB_OK;
}

}  // extern "C"
