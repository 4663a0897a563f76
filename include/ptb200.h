/*
 * ptb200.h -- C ABI of libptb200.so: the B200-native (sm_100a) batched RoadRunner / TSModel transit
 * model evaluation and fused white-noise population log-likelihood.
 *
 * This is the drop-in boundary for the hot path of hpparvi/PyTransit (v2.8.1).  Every entry point
 * cites the reference interface it replaces (paths relative to the PyTransit tree).  The reference
 * is pure Python + Numba, so "the FFI for this path" is what a ctypes/cffi binding inside
 * pytransit/models/roadrunner/ would call; INTEGRATION.md shows that binding.
 *
 * Conventions
 *  - Plain C types only: pointers and sizes.  No torch / numpy types cross this boundary.
 *  - Every data pointer may be a HOST pointer or a DEVICE pointer on the model's device; the
 *    library detects which (cudaPointerGetAttributes).  Device pointers are used in place
 *    (zero copy); host pointers are staged through buffers owned by the handle.
 *  - All floating-point arrays are C-contiguous float64, all index arrays int64 (what the
 *    reference's TransitModel.set_data produces, models/transitmodel.py:88-125).
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls whose
 *    outputs are device pointers are asynchronous on that stream; calls with host outputs return
 *    after the device-to-host copy has completed.
 *  - Return value: 0 on success, a negative ptb_status otherwise; ptb_last_error() gives the text.
 *    Numerical invalidity (a<=1, e<0, NaN) is DATA, not an error: the affected rows are NaN, as in
 *    the reference (models/roadrunner/model_full.py:40-42,80-82).
 *  - A handle is not thread safe (neither is the reference object); use one handle per host thread.
 *  - There is no CPU fallback: every entry point that computes fails with PTB_ECUDA when no
 *    CUDA device is usable.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_VERSION 100 /* 0.1.0 */

typedef struct ptb_model ptb_model; /* opaque */

enum ptb_status {
    PTB_OK = 0,
    PTB_EINVAL = -1, /* bad argument value        -> Python ValueError            */
    PTB_ESHAPE = -2, /* inconsistent array shapes -> Python ValueError            */
    PTB_ECUDA = -3,  /* CUDA runtime failure      -> Python RuntimeError          */
    PTB_ENOMEM = -4, /* allocation failure        -> Python MemoryError           */
    PTB_ESTATE = -5, /* call order (no set_data)  -> Python RuntimeError          */
    PTB_ENOTIMPL = -6 /* unknown limb darkening law -> Python NotImplementedError  */
};

/* Limb-darkening laws of RoadRunnerModel.ldmodels (models/roadrunner/rrmodel.py:48-58), evaluated
 * on the device from coefficient arrays (models/numba/ldmodels.py:22-175).  PTB_LD_PROFILES means
 * the caller supplies the tabulated profile ldp[npv,npb,nz] at ptb_get_tables().mu together with
 * istar[npv,npb] -- the LDModel protocol (models/ldmodel.py:21-39) and custom callables. */
enum ptb_ldlaw {
    PTB_LD_UNIFORM = 0, PTB_LD_LINEAR = 1, PTB_LD_QUADRATIC = 2, PTB_LD_QUADRATIC_TRI = 3,
    PTB_LD_NONLINEAR = 4, PTB_LD_GENERAL = 5, PTB_LD_SQUARE_ROOT = 6, PTB_LD_LOGARITHMIC = 7,
    PTB_LD_EXPONENTIAL = 8, PTB_LD_POWER_2 = 9, PTB_LD_POWER_2_PM = 10,
    PTB_LD_PROFILES = 100
};

/* Debug / parity taps for ptb_get_stage(): the per-vector intermediates of
 * models/roadrunner/model_full.py:34-70. */
enum ptb_stage {
    PTB_STAGE_LDP = 0,   /* [npv,npb,nz]   limb-darkening profile at the mu nodes              */
    PTB_STAGE_ISTAR = 1, /* [npv,npb]      disk-integrated intensity                           */
    PTB_STAGE_LDM = 2,   /* [npv,npb,ng]   limb-darkening means                                */
    PTB_STAGE_XYC = 3,   /* [npv,2,5]      Taylor-series orbit coefficients (solve2d)           */
    PTB_STAGE_BBOX = 4,  /* [npv,2]        un-padded contact times T1,T4 (bounding_box)         */
    PTB_STAGE_GOOD = 5   /* [npv]          1.0 valid / 0.0 invalid parameter vector            */
};

/* Constructor arguments: RoadRunnerModel.__init__ (models/roadrunner/rrmodel.py:60-63). */
typedef struct ptb_config {
    int32_t device;             /* CUDA device ordinal                                          */
    int32_t ldlaw;              /* enum ptb_ldlaw                                                */
    int32_t nk, nzin, nzlimb, ng; /* defaults 256, 20, 20, 100                                  */
    double kmin, kmax, zcut;    /* klims = (0.005, 0.5), zcut = 0.7                              */
    int32_t precompute_weights; /* TSModel only (models/roadrunner/tsmodel.py:117-120)           */
    int32_t precision;          /* 0 = fp64 (default); 1 = opt-in fp32 sample arithmetic         */
} ptb_config;

/* Fill *cfg with the reference defaults (rrmodel.py:60-63). */
void ptb_default_config(ptb_config *cfg);

int ptb_version(void);

/* Text of the last error on this handle (or of the last failed ptb_create when h is NULL). */
const char *ptb_last_error(const ptb_model *h);

/* RoadRunnerModel.__init__ + init_integration (rrmodel.py:60-173): builds the z grid
 * (common.py:131-149) on the host and the weight table W[nk,ng,nz] (common.py:188-223) on the
 * device. */
int ptb_create(const ptb_config *cfg, ptb_model **out);
void ptb_destroy(ptb_model *h);

/* The attributes RoadRunnerModel exposes after init_integration (rrmodel.py:165-173):
 * ze[nz], zm[nz], mu[nz], weights[nk,ng,nz], dk, dg.  Any output pointer may be NULL. */
int ptb_get_tables(ptb_model *h, double *ze, double *zm, double *mu, double *weights, double *dk,
                   double *dg);

/* TransitModel.set_data / RoadRunnerModel.set_data (models/transitmodel.py:56-125,
 * rrmodel.py:156-163).  lcids[npt] in [0,nlc); pbids[nlc] in [0,npb); epids[nlc] in [0,nep);
 * nsamples[nlc] >= 1; exptimes[nlc].  lcids may be NULL (all zeros, nlc must be 1). */
int ptb_set_data(ptb_model *h, const double *time, int64_t npt, const int64_t *lcids, int64_t nlc,
                 const int64_t *pbids, int64_t npb, const int64_t *epids, int64_t nep,
                 const int64_t *nsamples, const double *exptimes);

/* RoadRunnerModel.evaluate -> rr_full / rr_simple (rrmodel.py:175-238, model_full.py:9-100,
 * model_simple.py:11-80) on fully expanded arrays:
 *   k[npv,kcols] (kcols 1 or npb); ld = ldc[npv,npb,nld] for a named law, or ldp[npv,npb,nz] with
 *   istar[npv,npb] for PTB_LD_PROFILES (istar NULL otherwise); t0[npv,nep]; p,a,inc,e,w[npv].
 * flux[npv,npt] may be a device pointer (written in place), a host pointer, or NULL (the flux
 * stays in the handle's device buffer: ptb_flux_device_ptr).  It is a float64 array, or a float32
 * array when the handle was created with ptb_config.precision = 1 (opt-in fp32 mode: the phase
 * fold stays fp64, the per-sample geometry / limb-darkening arithmetic and the stored flux are
 * fp32; <= 1 ppm from the fp64 result). */
int ptb_rr_evaluate(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld,
                    int64_t nld, const double *istar, const double *t0, const double *p,
                    const double *a, const double *inc, const double *e, const double *w,
                    void *flux, void *stream);

/* EclipseModel.evaluate -> eclipse_model (models/new_eclipse_model.py:33-70, models/roadrunner/model_eclipse.py:11-81):
 * the secondary-eclipse sibling of the RoadRunner model.  k[npv]; t0[npv,nep]; p,a,inc,e,w[npv]; rstar [R_sun].
 * flux[npv,npt] = pi k^2 - (planet area occulted by the star), averaged over the exposure sub-samples: the
 * orbit is expanded about mid-eclipse (eclipse_time_offset = the reference's eclipse_phase,
 * orbits/orbits_py.py:544-555) and the eclipse centre is shifted by the light travel time.  The handle must
 * have been created with PTB_LD_UNIFORM; light-curve / epoch ids, nsamples and exptimes come from ptb_set_data. */
int ptb_eclipse_evaluate(ptb_model *h, int64_t npv, const double *k, const double *t0, const double *p,
                         const double *a, const double *inc, const double *e, const double *w, double rstar,
                         double *flux, void *stream);

/* EclipseSpectroscopyModel.evaluate -> esmodel (models/roadrunner/esmodel.py:46-89, model_ecspec.py:13-63): the
 * secondary eclipse in npb spectroscopic channels.  fratio[npv,npb] planet-to-star flux ratios; k,t0,p,a,inc,e,w,
 * rstar[npv].  flux[npv,npb,npt] = 1 - (f A / pi) / (1 + f k^2), averaged over the exposure sub-samples
 * (nsamples[0], exptimes[0] of ptb_set_data; one light curve, one epoch).  Handle created with PTB_LD_UNIFORM. */
int ptb_es_evaluate(ptb_model *h, int64_t npv, int64_t npb, const double *fratio, const double *k, const double *t0,
                    const double *p, const double *a, const double *inc, const double *e, const double *w,
                    const double *rstar, double *flux, void *stream);

/* Observed fluxes and noise blocks for the fused likelihood: the (o, slices, nids) arguments of
 * lnlike_normal (lpf/loglikelihood/wnloglikelihood.py:22-35,43-55).  obs[npt]; slices[nsl,2]
 * half-open point ranges; nids[nsl] in [0,nblocks).  Points outside every slice do not contribute.
 * slices NULL = one slice covering all points with noise id 0. */
int ptb_set_obs(ptb_model *h, const double *obs, const int64_t *slices, const int64_t *nids,
                int64_t nsl, int64_t nblocks);

/* BaseLPF.lnlikelihood with WNLogLikelihood (lpf/lpf.py:454-475, wnloglikelihood.py:79-81) fused
 * with the model: lnl[npv] = sum_j -log(s) - 0.5 log(2 pi) - 0.5 ((obs_j - model_ij)/s)^2 with
 * s = sigma[npv,nblocks] (already 10**pv).  No [npv,npt] flux is materialised.  Model arguments as
 * ptb_rr_evaluate. */
int ptb_rr_lnlike(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld,
                  int64_t nld, const double *istar, const double *t0, const double *p,
                  const double *a, const double *inc, const double *e, const double *w,
                  const double *sigma, double *lnl, void *stream);

/* The same, for a population sharded over the GPUs of one box: the likelihood and its
 * all-gather in one pass.  The reference has no multi-GPU path; the reduction it replaces is the prange over
 * the population of lnlike_normal (wnloglikelihood.py:30-34), whose rows are independent.
 * peer_bufs[r] (r < world <= 16) is rank r's gathered array lnl_all[world * npv] as a device pointer valid in
 * THIS process (symmetric / CUDA-IPC peer memory over NVLink); the finishing kernel stores this rank's lnl[npv]
 * into slot `rank` of every one of them.
 * Ordering.  peer_flags == NULL: the caller orders the ranks afterwards (e.g. one symmetric-memory barrier)
 * before anyone reads its gathered array.  peer_flags != NULL: peer_flags[r] is rank r's arrival array
 * uint64[world] (zero-initialised peer memory, same mapping rules); after all its stores the finishing kernel
 * publishes `seq` (>= 1, increasing by one per call on every rank) into slot `rank` of every arrival array with
 * a system-scope release, and the kernel's last thread block then waits until all `world` slots of THIS rank's
 * array have reached `seq` -- work queued on `stream` afterwards sees the complete gathered array,
 * with no host-side synchronisation.  Use at least two gathered arrays alternately (a peer overwrites the array
 * of step s at its step s+2, which it can only reach after this rank has published step s+1).  A peer that
 * never arrives is reported by ptb_gather_status after a 20 s device-side timeout.  Every rank must pass the
 * same npv. */
int ptb_rr_lnlike_allgather(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld,
                            int64_t nld, const double *istar, const double *t0, const double *p,
                            const double *a, const double *inc, const double *e, const double *w,
                            const double *sigma, double *const *peer_bufs, uint64_t *const *peer_flags,
                            uint64_t seq, int32_t world, int32_t rank, void *stream);

/* PTB_OK, or PTB_ESTATE with *timed_out_rank = the first rank whose shard a fused all-gather waited for in
 * vain (-1: none).  Synchronises the device. */
int ptb_gather_status(ptb_model *h, int32_t *timed_out_rank);

/* lnlike_normal (wnloglikelihood.py:22-35) on an existing model flux m[npv,npt] (device or host),
 * e.g. baseline-multiplied flux produced by the caller (lpf/lpf.py:445-449). */
int ptb_lnlike_normal(ptb_model *h, int64_t npv, const double *model, const double *sigma,
                      double *lnl, void *stream);

/* Parameter-vector layout of a log posterior function: where BaseLPF keeps the physical parameters in a
 * row of the population array pvp[npv, npar] (lpf/lpf.py:328-356: tc, p, rho, b, k2, (q1, q2) per
 * passband; WNLogLikelihood appends log10 sigma per noise block, wnloglikelihood.py:68-77).  Column
 * indices are explicit so that TransitAnalysis' per-planet blocks (lpf/transitanalysis.py:72-107) map too. */
typedef struct ptb_lpf_layout {
    int32_t npar;               /* columns of pvp                                                    */
    int32_t i_tc, i_p, i_rho, i_b; /* zero epoch, period [d], stellar density [g/cm^3], impact param. */
    int32_t i_k2, nk2;          /* area ratio column(s): 1, or npb consecutive (k = sqrt(k2))        */
    int32_t i_ld, nldc;         /* first limb-darkening column; npb x nldc consecutive                */
    int32_t ld_map;             /* 1: (q1,q2) -> (u,v) = (2 sqrt(q1) q2, sqrt(q1)(1 - 2 q2)), map_ldc
                                   (lpf/lpf.py:84-91); 0: coefficients passed through               */
    int32_t i_secw, i_sesw;     /* sqrt(e) cos w, sqrt(e) sin w columns; -1 = circular orbit         */
    int32_t inc_mode;           /* 0: i_from_ba (orbits_py.py:674-688); 1: i_from_baew (:654-670)   */
    int32_t i_loge, nloge;      /* log10 sigma columns, one per noise block (lnlike only)            */
    int32_t ntc;                /* consecutive transit-centre columns starting at i_tc: one per epoch of the
                                   dataset (TTVLPF, lpf/ttvlpf.py:70-83); 1 for BaseLPF                */
    int32_t i_bl;               /* first column of the baseline coefficients (ptb_set_baseline); -1 = none */
    int32_t reserved_;
    double tref;                /* reference time subtracted from tc (lpf/lpf.py:438)                */
} ptb_lpf_layout;

/* BaseLPF.transit_model (lpf/lpf.py:435-443): k = sqrt(k2), t0 = tc - tref, a = as_from_rhop(rho, p)
 * (orbits_py.py:604-618, G = scipy.constants.G = 6.67430e-11), i = arccos(b/a), ldc = map_ldc(...), all
 * computed on the device from pvp (host or device pointer), then the RoadRunner evaluation of
 * ptb_rr_evaluate.  lay->ntc must equal the number of epochs of the dataset (t0[npv, nep] = pvp[:, i_tc : i_tc
 * + ntc] - tref, TTVLPF.transit_model, lpf/ttvlpf.py:78-86). */
int ptb_lpf_transit_model(ptb_model *h, const double *pvp, int64_t npv, const ptb_lpf_layout *lay,
                          void *flux, void *stream);

/* Multiplicative baseline of the LPF layer (BaseLPF.baseline / flux_model, lpf/lpf.py:418-449) as a linear model in
 * per-point basis functions: bl[ipv, j] = sum_{c < ncoef[lc(j)]} pvp[ipv, i_bl + cstart[lc(j)] + c] * basis[c][j], and
 * 1 for light curves with ncoef = 0.  Covers LegendreBaseline (lpf/baselines/legendrebaseline.py:23-40: basis = Legendre
 * polynomials of the normalised time) and LinearModelBaseline (lpf/baselines/linearbaseline.py:22-36: basis = 1 and the
 * covariates).  basis[nbasis][npt] (host or device, copied), cstart / ncoef [nlc] with ncoef <= nbasis.  NULL basis
 * clears.  set_data clears it too (it is tied to the time axis).  fp64 handles only. */
int ptb_set_baseline(ptb_model *h, const double *basis, int64_t nbasis, const int64_t *cstart, const int64_t *ncoef);
/* BaseLPF.flux_model (lpf/lpf.py:445-449): baseline * transit_model (+ trends = 0); with only_baseline != 0 the
 * baseline alone (BaseLPF.baseline).  flux[npv, npt] host or device. */
int ptb_lpf_flux_model(ptb_model *h, const double *pvp, int64_t npv, const ptb_lpf_layout *lay, int32_t only_baseline,
                       double *flux, void *stream);
/* BaseLPF.lnlikelihood with one WNLogLikelihood (lpf/lpf.py:454-475, wnloglikelihood.py:79-81): the same
 * mapping, sigma = 10**pvp[:, i_loge : i_loge + nloge], and the fused model + likelihood of ptb_rr_lnlike.
 * The population never leaves the device when pvp and lnl are device pointers.  With a baseline registered and
 * lay->i_bl >= 0 the transit flux is materialised on the device and the likelihood kernel multiplies the baseline in on
 * the fly (the block-skipping fused kernel assumes an out-of-transit model value of exactly 1). */
int ptb_lpf_lnlike(ptb_model *h, const double *pvp, int64_t npv, const ptb_lpf_layout *lay, double *lnl,
                   void *stream);

/* TransmissionSpectroscopyModel.evaluate -> tsmodel_serial (models/roadrunner/tsmodel.py:46-130,
 * model_trspec.py:11-93): k[npv,npb]; ld as above with npb = the spectroscopic channel count;
 * t0,p,a,inc,e,w[npv]; flux[npv,npb,npt].  Uses nsamples[0], exptimes[0] of set_data.
 * `flux` is double[], or float[] when the handle was created with precision = 1 (opt-in fp32 output mode:
 * all arithmetic stays fp64, the stored value is rounded to float -- half the bytes of the write-out that
 * bounds this model; within 1 ppm of the fp64 result by construction). */
int ptb_ts_evaluate(ptb_model *h, int64_t npv, int64_t npb, const double *k, const double *ld,
                    int64_t nld, const double *istar, const double *t0, const double *p,
                    const double *a, const double *inc, const double *e, const double *w,
                    void *flux, void *stream);

/* LDTkLDModel.__call__ (models/ldtkldm.py:74-89) -> trilinear_interpolation_set +
 * integrate_profiles_set (models/numba/ldtkldm.py:53-60,77-91): profiles[nx,ny,nz3,npb,nmu],
 * stellar parameters xs,ys,zs[npv] -> ldp[npv,npb,nmu], istar[npv,npb] (device or host). */
int ptb_ldtk_profiles(ptb_model *h, const double *profiles, int64_t nx, int64_t ny, int64_t nz3,
                      int64_t npb, int64_t nmu, const double *xs, const double *ys,
                      const double *zs, int64_t npv, double x0, double dx, double y0, double dy,
                      double z0, double dz, const double *mu, double *ldp, double *istar,
                      void *stream);

/* Model derivatives dfdk(k, b, k0, lda, dg, ist) and dfdb(k, b, a, ak, lda, dg, ist) (models/roadrunner/common.py:104-128)
 * for the population of the LAST ptb_rr_evaluate / ptb_rr_lnlike: b[npv, nb] separations per vector (host or device);
 * k, the LD-mean row `lda` and I* of passband `pb` are the vector's own, kappa0 / lens area / kite area come from the
 * device's circle_circle_intersection_area_kite.  dfdk / dfdb [npv, nb], either may be NULL. */
int ptb_rr_derivatives(ptb_model *h, int64_t npv, int64_t nb, int64_t pb, const double *b, double *dfdk,
                       double *dfdb, void *stream);

/* Parity taps: copy a per-vector intermediate of the LAST evaluate/lnlike call to out (host or
 * device).  See enum ptb_stage for shapes. */
int ptb_get_stage(ptb_model *h, int32_t stage, double *out);

/* Inject Taylor coefficients xyc[npv,2,5] to be used instead of the device solve2d for the
 * following evaluations (NULL clears).  Lets everything downstream of the third-party
 * meepmeep.solve2d be pinned independently. */
int ptb_inject_xyc(ptb_model *h, const double *xyc, int64_t npv);

/* Device pointer / element count of the handle-owned result of the last evaluate call whose
 * output pointer was NULL (the RoadRunnerModelCL `_b_f` analogue, rrmodel_cl.py:365-369). */
int ptb_flux_device_ptr(ptb_model *h, void **ptr, int64_t *count);

/* Page-locked host memory for results (so device-to-host copies run at full PCIe rate). */
int ptb_host_alloc(void **ptr, size_t bytes);
int ptb_host_free(void *ptr);

/* Managed host result -- the analogue of RoadRunnerModelCL's persistent host array `f`
 * (models/roadrunner/rrmodel_cl.py:365-369: `enqueue_copy(queue, self.f, self._b_f); return self.f`).
 * `buf` is page-locked memory from ptb_host_alloc holding `count` result elements; the caller promises
 * that between calls nobody but this handle writes to it.  After binding, every ptb_rr_evaluate /
 * ptb_ts_evaluate whose `flux` argument is `buf` (and whose result has `count` elements) keeps `buf`
 * up to date by DELTA transfer: a transit model is exactly 1.0 outside the transit windows
 * (model_full.py:91), so after the first full copy only the 16-element blocks (128 B) that differ from 1.0
 * now -- or did after the previous call -- are written, by the GPU, straight into `buf` over PCIe.
 * The content of `buf` after each call is identical to a full copy.  Up to 8 buffers can be bound at the
 * same time, each with its own record of what it holds: a caller that still needs the previous result passes
 * another bound buffer to the next call (the Python layer rotates a small pool, so results never alias).
 * Binding a bound buffer again invalidates its record (next transfer is a full copy).  NULL unbinds all.  Any
 * other host pointer passed as `flux` takes the plain full-copy path.  A handle is used by one host thread at
 * a time and its calls are ordered on the stream they are given. */
int ptb_bind_host_result(ptb_model *h, void *buf, int64_t count);
/* Forget `buf` (NULL: every bound buffer).  Waits for the device first. */
int ptb_unbind_host_result(ptb_model *h, void *buf);
/* Bytes the last managed transfer moved to the host, and how many delta / full transfers ran. */
int ptb_host_result_stats(const ptb_model *h, int64_t *last_bytes, int64_t *delta_calls,
                          int64_t *full_calls);

/* Kernels launched by this handle since creation (bench.py's gpu_launches evidence; kernels inside a replayed
 * CUDA graph are counted). */
int64_t ptb_launch_count(const ptb_model *h);

/* Measured fp64 FMA throughput of the handle's GPU in TFLOP/s (8 independent DFMA chains per thread, no memory
 * traffic; best of 4 timed launches): the roofline denominator bench.py uses for the fused-likelihood kernel. */
int ptb_measure_fp64_peak(ptb_model *h, double *tflops);

/* CUDA-graph replay of launch-bound calls.  For populations of up to 2048 vectors the kernel sequence of
 * ptb_rr_evaluate / ptb_rr_lnlike / ptb_eclipse_evaluate (argument copy, orbit solve on its side stream,
 * sort, limb-darkening contraction, points kernel, likelihood finish) is captured the second time a call
 * signature (sizes, which arguments are host or device pointers, output pointer) is seen, and replayed with
 * one cudaGraphLaunch afterwards.  Results are identical to the eager path.  Enabled by default
 * (environment PTB_GRAPHS=0 disables it process-wide); never used while ptb_set_profiling is on. */
int ptb_set_graphs(ptb_model *h, int32_t enabled);
int ptb_graph_stats(const ptb_model *h, int64_t *replays, int64_t *captures);

/* Per-kernel device timing for bench.py's roofline: when enabled, CUDA events are recorded on the
 * launching stream around the per-vector setup kernel(s) and around the dominant npv x npt kernel of
 * every evaluate / lnlike call.  ptb_last_timing waits for the last call and returns both durations
 * in milliseconds. */
int ptb_set_profiling(ptb_model *h, int32_t enabled);
int ptb_last_timing(ptb_model *h, double *setup_ms, double *points_ms);
/* Totals over the calls since ptb_set_profiling(h, 1) (at most the last 256): no host
 * synchronisation happens between calls, only here. */
int ptb_timing_summary(ptb_model *h, int64_t *ncalls, double *setup_ms_total, double *points_ms_total);

/* Block until all work queued by this handle on `stream` has finished. */
int ptb_synchronize(ptb_model *h, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
