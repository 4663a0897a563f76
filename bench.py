#!/usr/bin/env python
"""bench.py -- throughput of the batched RoadRunner evaluation, BASELINE.json's metric:
model flux points/s (npv x npt) on configs[1]: RoadRunnerModel('power-2') population npv=8192 x 20 000
TESS 2-min cadence points, single passband, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]
                    [--precision fp64|fp32] [--host-result delta|copy] [--gather peer|peer-barrier|nccl]

One "step" = one pass of the hot path over the whole population: per-vector setup (limb-darkening
profile, LD-mean contraction, Kepler/Taylor orbit, contact times) + the npv x npt flux kernel.

Our arm prints ONE JSON line:
  value     whole-job flux points/s, parameters + time axis resident in HBM, flux written to HBM
            (device-resident output, `evaluate(copy=False)`), CUDA-event timed, max over ranks;
  e2e       same metric through the public API with HOST numpy inputs and a host result, H2D and D2H inside the
            timed region, in the opt-in delta mode (`host_result='delta'`); e2e_copy: the default mode (one full
            D2H copy into an array of the caller's own); both with a roofline against the device-to-host rate
            measured in the same run with every rank copying concurrently;
  roofline  the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic
            bytes = 8 B per flux point; `traffic` and the executed-instruction counters are measured
            in the run by an ncu child process (--no-counters skips it); C5 (nothing written): executed fp64 work
            against a DFMA peak measured on the same GPU;
  cpu_baseline        the CPU oracle port (oracle/rr_oracle.c, OpenMP, all host threads),
  cpu_baseline_numba  the reference's own Numba path (baseline/_ref), threads = cores -- both on bounded samples
            of the same workload (N=1, rank 0);
  collective  (default line) the C5 shard's fused lnL + all-gather measured in the same run: peer stores ordered
            by device-side flags, with the host-barrier and NCCL variants and a bit-identity check.
N > 1 (torchrun): the population is sharded, 8192 vectors per GPU (weak scaling), no data-path
collective for the flux; timing is barrier + synchronize bracketed, max over ranks.

`--impl reference` times the reference's CPU implementation of the path: the oracle port (kind "port") with all
host threads, and the reference's own Numba path beside it, on a bounded sample of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import workloads as wl  # noqa: E402

METRIC = 'model flux points/s (npv x npt)'
UNIT = 'points/s'


def workload(name: str, rank: int = 0):
    seed_shift = 1000 * rank
    if name == 'c2':
        c = wl.config2(seed=2 + seed_shift)
        desc = "C2: RoadRunnerModel('power-2') npv=8192 x npt=20000 (TESS 2-min), 1 passband, ns=1"
    elif name == 'c3':
        c = wl.config3(seed=3 + seed_shift)
        desc = "C3: RoadRunnerModel('quadratic') npv=16384 x npt=65536 (4 light curves = 4 passbands), ns=10"
    elif name == 'c5':
        c = wl.config5(npv=8192, seed=5 + seed_shift)
        desc = "C5 shard: eccentric 'power-2' npv=8192 x npt=100000 + fused Gaussian lnL"
    elif name == 'c4':
        c = wl.config4(seed=4 + seed_shift)
        desc = "C4: TSModel, tabulated (LDTk-style) profiles, npv=1024 x npb=1000 channels x npt=2000"
    elif name == 'c1':
        c = wl.config1()
        c.npv, c.npt = 1, c.time.size
        desc = "C1: RoadRunnerModel('quadratic') README example, npv=1 x npt=10000 (latency bound)"
    else:
        raise SystemExit(f'unknown workload {name}')
    return c, desc


def config_dict(args, desc: str, world: int, c=None) -> dict:
    """The `config` object -- identical for our arm and the reference arm of the same command line."""
    lnl = args.workload == 'c5'
    out_b = 4 if args.precision == 'fp32' else 8
    sizes = {'c1': 1e4, 'c2': 8192 * 2e4, 'c3': 16384 * 65536, 'c4': 1024 * 1000 * 2000, 'c5': 0}[args.workload]
    out_gb = out_b * 1e-9 * sizes
    gather = {'peer': 'lnL[npv] all-gathered by peer stores from the finishing kernel into symmetric memory over NVLink, ranks ordered '
                      'by device-side flags',
              'peer-barrier': 'lnL[npv] all-gathered by peer stores + one symmetric-memory barrier per step',
              'nccl': 'NCCL all-gather of lnL[npv] per step'}[args.gather]
    return {'workload': desc,
            'parallelism': f'population sharded over {world} GPU(s), ' +
                           (gather if (lnl and world > 1) else
                            'flux shards stay on their GPU: no data-path collective (the lnL all-gather of the path is measured in `collective`)'),
            'l2': ('time+obs (1.6 MB) are L2 resident by design; nothing is written' if lnl else
                   'each step streams %.2f GB of output through L2 (126 MB): inputs/outputs exceed L2, no explicit flush' % out_gb
                   if out_gb > 0.2 else
                   'latency-bound single vector: time axis and output (80 KB each) are L2 resident by design, no flush; '
                   'the timed loop replays the call as a CUDA graph, per-kernel durations are taken in a separate loop')}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake'}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ---------------------------------------------------------------------------------------------
# CPU side: oracle port
# ---------------------------------------------------------------------------------------------
def oracle_step(orc, tab, c, rows, lnl: bool):
    sl = slice(0, rows)
    law = c.ldmodel
    if c.name == 'C4':
        prof, (x0, dx), (y0, dy), (z0, dz) = c.table
        ldp, istar = orc.ldtk_profiles(prof, c.teff[sl], c.logg[sl], c.metal[sl], x0, dx, y0, dy, z0, dz, tab.mu)
        orc.tsmodel(tab, c.time, c.k[sl], c.t0[sl], c.p[sl], c.a[sl], c.i[sl], c.e[sl], c.w[sl], 1, 0.0, ldp, istar)
        return rows * c.npb * c.npt
    if c.name == 'C1':
        ldp, istar = orc.evaluate_ld(law, tab.mu, np.asarray(c.ldc).reshape(1, 1, -1))
        one = lambda v: np.full(1, float(v))
        orc.rr_full(tab, c.time, np.full((1, 1), c.k), np.full((1, 1), c.t0), one(c.p), one(c.a), one(c.i), one(c.e), one(c.w),
                    np.zeros(c.npt, np.int64), np.zeros(1, np.int64), np.zeros(1, np.int64), np.ones(1, np.int64), np.zeros(1),
                    ldp, istar)
        return c.npt
    ldp, istar = orc.evaluate_ld(law, tab.mu, c.ldc[sl])
    flux = orc.rr_full(tab, c.time, c.k[sl], c.t0[sl], c.p[sl], c.a[sl], c.i[sl], c.e[sl], c.w[sl], c.lcids, c.pbids,
                       c.epids, c.nsamples, c.exptimes, ldp, istar)
    if lnl:
        orc.lnlike_normal(c.obs, flux, c.sigma[sl], c.slices, c.nids)
    return rows * c.npt


# ---------------------------------------------------------------------------------------------
# CPU side: the reference's own Numba path, when its files travelled with the repo
# ---------------------------------------------------------------------------------------------
def numba_baseline(name: str, c, seconds: float, max_rows: int = 512):
    """Times the reference's RoadRunnerModel / TSModel / lnlike_normal -- its own files, loaded by
    baseline/refload.py from baseline/_ref (git-ignored pip install of the reference made by build()) -- on a
    bounded sample of the workload, all host cores.  -> dict, or {'unavailable': why}."""
    try:
        from baseline.refload import load_reference
        ref = load_reference()
        import numba
    except Exception as ex:   # no reference tree on this box, or no numba
        return {'unavailable': f'{type(ex).__name__}: {ex}'[:200]}
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    numba.set_num_threads(min(cores, numba.config.NUMBA_NUM_THREADS))
    nthreads = numba.get_num_threads()
    lnl = name == 'c5'
    if name == 'c4':
        # TSModel(nthreads > 1) is broken in the reference: serial, as BASELINE.md prescribes
        nthreads = 1
        m0 = ref.TSModel('power-2')
        prof, (x0, dx), (y0, dy), (z0, dz) = c.table
        from pytransit.models.ldmodel import LDModel

        class Tab(LDModel):      # stellar parameters -> profiles, as LDTkLDModel.__call__ does (models/ldtkldm.py:74-89)
            rows = 1

            def __call__(self, mu, x):
                sl = slice(0, self.rows)
                ldp = ref.ldtkldm.trilinear_interpolation_set(prof, c.teff[sl].copy(), c.logg[sl].copy(), c.metal[sl].copy(), x0, dx,
                                                              prof.shape[0], y0, dy, prof.shape[1], z0, dz, prof.shape[2])
                return ldp, ref.ldtkldm.integrate_profiles_set(mu, ldp)

            def _evaluate(self, mu, x):
                raise NotImplementedError

            def _integrate(self, x):
                raise NotImplementedError
        tab = Tab()
        m = ref.TSModel(tab, nthreads=1)
        m.set_data(c.time)
        del m0

        def run(rows):
            sl = slice(0, rows)
            tab.rows = rows
            # the reference insists on a 3-D coefficient array for a population; the tabulated model does not read it
            m.evaluate(c.k[sl], np.zeros((rows, c.npb, 3)), c.t0[sl], c.p[sl], c.a[sl], c.i[sl], c.e[sl], c.w[sl])
            return rows * c.npb * c.npt
        max_rows = min(max_rows, 16)
    elif name == 'c1':
        m = ref.RoadRunnerModel(c.ldmodel, nthreads=nthreads)
        m.set_data(c.time)

        def run(rows):
            m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w)
            return c.npt
        max_rows = 1
    else:
        m = ref.RoadRunnerModel(c.ldmodel, nthreads=nthreads)
        m.set_data(c.time, c.lcids, c.pbids, c.nsamples, c.exptimes, c.epids)

        def run(rows):
            sl = slice(0, rows)
            flux = m.evaluate(c.k[sl], c.ldc[sl], c.t0[sl], c.p[sl], c.a[sl], c.i[sl], c.e[sl], c.w[sl])
            if lnl:
                ref.lnlike_normal(c.obs, np.atleast_2d(flux), c.sigma[sl], c.slices, c.nids)
            return rows * c.npt
    rows = 1 if name == 'c1' else min(c.npv, 8)
    t = time.perf_counter()
    run(rows)                                   # JIT (or cache load) + first touch
    jit_s = time.perf_counter() - t
    t = time.perf_counter()
    run(rows)
    dt = max(time.perf_counter() - t, 1e-5)
    if name != 'c1':
        per = getattr(c, 'npb_out', 1) * c.npt
        rows = int(max(rows, min(c.npv, max_rows, rows * (seconds / 3.0) / dt, 2e9 // (8 * per))))
    best, total, passes, cpu_ratio = None, 0.0, 0, 0.0
    while passes < 3 or (total < seconds and passes < 50):
        t, tc = time.perf_counter(), time.process_time()
        n = run(rows)
        dt, dc = time.perf_counter() - t, time.process_time() - tc
        if best is None or dt < best:
            best, cpu_ratio = dt, dc / max(dt, 1e-9)
        total += dt
        passes += 1
    return {'value': n / best, 'unit': UNIT, 'cores': nthreads, 'kind': 'numba+standin',
            'sample': f'{rows} of {getattr(c, "npv", 1)} parameter vectors x {n // rows} points each, best of {passes} passes '
                      f'({best:.3f} s best, {total:.1f} s in all, first call incl. JIT {jit_s:.1f} s)',
            'threads_busy': round(cpu_ratio, 2), 'numba': numba.__version__, 'threading_layer': numba.threading_layer() if nthreads > 1 else 'serial',
            'note': "the reference's own files (rrmodel.py, model_full.py, model_trspec.py, common.py, ldmodels.py, wnloglikelihood.py:22-35) "
                    "run unmodified under Numba; the three absent third-party meepmeep functions come from baseline/_standin "
                    "(restated from the reference's orbits/taylor_z.py)"}


def host_threads(orc) -> int:
    """Use every host core this process may run on (torchrun exports OMP_NUM_THREADS=1; undo that here)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    orc.set_threads(n)
    return n


def cpu_sample_rows(orc, tab, c, lnl, target_s: float):
    """Pick a vector count so that one oracle pass takes about `target_s` seconds."""
    rows = min(64, c.npv)
    oracle_step(orc, tab, c, rows, lnl)          # warm-up (page faults, OpenMP pool)
    t = time.perf_counter()
    oracle_step(orc, tab, c, rows, lnl)
    dt = max(time.perf_counter() - t, 1e-4)
    rows = int(min(c.npv, max(rows, rows * target_s / dt)))
    # keep the sampled flux array below ~2 GB
    return max(1, min(rows, int(2e9 // (8 * c.npt * getattr(c, 'npb_out', 1)))))


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.lib()
    tab = orc.Tables()
    c, desc = workload(args.workload)
    if args.workload == 'c4':
        c.table, c.npb_out = wl.ldtk_style_table(c.npb, tab.mu), c.npb
    lnl = args.workload == 'c5'
    cores = host_threads(orc)
    budget = 120.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    rows = cpu_sample_rows(orc, tab, c, lnl, target_s=min(10.0, budget))
    for _ in range(args.warmup):
        oracle_step(orc, tab, c, rows, lnl)
    t = time.perf_counter()
    pts = 0
    for _ in range(args.steps):
        pts += oracle_step(orc, tab, c, rows, lnl)
    dt = time.perf_counter() - t
    value = pts / dt
    sample = f'{rows} of {c.npv} parameter vectors x {pts // (rows * args.steps)} points per step (rows 0..{rows - 1} of the same seeded workload)'
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(args, desc, world),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                             'note': 'oracle/rr_oracle.c (C restatement of the Numba path, OpenMP); the reference '
                                     'itself needs the absent meepmeep package'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    if not args.no_numba:
        line['cpu_baseline_numba'] = numba_baseline(args.workload, c, seconds=min(args.cpu_seconds, 15.0))
    emit(line)


# ---------------------------------------------------------------------------------------------
# hardware counters of the dominant kernel, measured on this box in this run (a child process under ncu)
# ---------------------------------------------------------------------------------------------
NCU_METRICS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum',
               'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
               'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__time_duration.sum']


def ncu_counters(args, kernel: str, timeout: float = 240.0):
    """One launch of `kernel` (regex) profiled with ncu in a child process that runs four steps of the same workload.
    Counters only -- no timing taken under the profiler is used as a bench value.  -> dict or {'unavailable': why}."""
    import csv
    import io
    import shutil
    import subprocess
    ncu = shutil.which('ncu') or '/usr/local/cuda/bin/ncu'
    if not Path(ncu).exists():
        return {'unavailable': 'ncu not found'}
    cmd = [ncu, '--metrics', ','.join(NCU_METRICS), '--clock-control', 'none', '-k', f'regex:{kernel}', '--launch-skip', '2', '-c', '1',
           '--csv', sys.executable, str(Path(__file__).resolve()), '--workload', args.workload, '--precision', args.precision,
           '--child-profile']
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, PTB_GRAPHS='0'))
    except Exception as ex:
        return {'unavailable': f'{type(ex).__name__}: {ex}'[:160]}
    rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 14]
    if not rows or 'Metric Name' not in rows[0]:
        return {'unavailable': ('ncu produced no counters: ' + (r.stderr or r.stdout)[-200:]).replace('\n', ' ')}
    h = rows[0]
    iname, iunit, ival, ikern = h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value'), h.index('Kernel Name')
    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12, 'msecond': 1.0, 'usecond': 1e-3, 'second': 1e3, 'nsecond': 1e-6}
    out = {'kernel': rows[1][ikern][:80]}
    for row in rows[1:]:
        try:
            out[row[iname]] = float(row[ival].replace(',', '')) * mult.get(row[iunit], 1.0)
        except ValueError:
            pass
    return out


def run_child_profile(args, local_rank: int = 0):
    """What ncu_counters profiles: four device-resident steps of the workload, nothing else."""
    import torch
    torch.cuda.set_device(local_rank)
    c, _ = workload(args.workload, 0)
    if args.workload == 'c4':
        import pytransit_b200 as pb
        c.table = wl.ldtk_style_table(c.npb, pb.TSModelCUDA('uniform', device=local_rank).mu)
        c.npb_out = c.npb
    m, step_device, *_ = make_steps(args, c, f'cuda:{local_rank}', local_rank, 1, None)
    for _ in range(4):
        out = step_device()
    torch.cuda.synchronize()
    del out


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------
def measured_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def make_steps(args, c, dev, local_rank, world, dist):
    """Build the model for the workload and return (model, step_device, step_host, pts_per_step, h2d_bytes, kernel name,
    algorithmic bytes of the dominant kernel per launch or None)."""
    import torch
    import pytransit_b200 as pb
    name = args.workload
    if name == 'c4':
        m = pb.TSModelCUDA(pb.TabulatedLDModel(*_table_args(c), device=local_rank), device=local_rank, host_result=args.host_result,
                           precision=args.precision)
        time_d = torch.as_tensor(c.time, device=dev)
        m.set_data(time_d)
        x = np.column_stack([c.teff, c.logg, c.metal])
        td = {k: torch.as_tensor(np.ascontiguousarray(getattr(c, k)), device=dev) for k in ('k', 't0', 'p', 'a', 'i', 'e', 'w')}
        pts = c.npv * c.npb * c.npt

        def step_device():
            return m.evaluate(td['k'], x, td['t0'], td['p'], td['a'], td['i'], td['e'], td['w'], copy=False)

        t0_alt = [c.t0, c.t0 + 0.02]
        nhost = [0]

        def step_host():      # alternates two populations (mid-transit times 29 min apart)
            nhost[0] += 1
            return m.evaluate(c.k, x, t0_alt[nhost[0] & 1], c.p, c.a, c.i, c.e, c.w, copy=True)

        h2d = sum(np.asarray(getattr(c, k)).nbytes for k in ('k', 't0', 'p', 'a', 'i', 'e', 'w')) + x.nbytes
        return m, step_device, step_host, pts, h2d, 'k_ts_flux', (4.0 if args.precision == 'fp32' else 8.0) * pts
    if name == 'c1':
        m = pb.RoadRunnerModelCUDA(c.ldmodel, device=local_rank)
        m.set_data(torch.as_tensor(c.time, device=dev))
        ldc_d = torch.as_tensor(c.ldc, device=dev)

        def step_device():
            return m.evaluate(c.k, ldc_d, c.t0, c.p, c.a, c.i, c.e, c.w, copy=False)

        def step_host():
            return m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w, copy=True)

        return m, step_device, step_host, c.npt, 8 * 8 + c.ldc.nbytes, 'k_rr_points', 8.0 * c.npt

    lnl = name == 'c5'
    m = pb.RoadRunnerModelCUDA(c.ldmodel, device=local_rank, precision=args.precision, host_result=args.host_result)
    esize = 4 if args.precision == 'fp32' else 8
    td = {k: torch.as_tensor(np.ascontiguousarray(getattr(c, k)), device=dev) for k in ('k', 'ldc', 't0', 'p', 'a', 'i', 'e', 'w')}
    time_d = torch.as_tensor(c.time, device=dev)
    if c.nlc > 1:
        m.set_data(time_d, c.lcids, c.pbids, c.nsamples, c.exptimes, c.epids)
    else:
        m.set_data(time_d, nsamples=c.nsamples, exptimes=c.exptimes)
    if lnl:
        m.set_obs(torch.as_tensor(c.obs, device=dev))
        sig_d = torch.as_tensor(c.sigma, device=dev)
    gathered = peer = None
    if lnl and world > 1:
        if args.gather != 'nccl':     # fused: the finishing kernel stores the shard into every rank's gathered array
            from pytransit_b200.distributed import PeerLnLGather
            peer = PeerLnLGather(m, c.npv, sync='flags' if args.gather == 'peer' else 'barrier')
        else:
            gathered = torch.empty((world * c.npv,), dtype=torch.float64, device=dev)

    def step_device():
        if lnl:
            if peer is not None:
                return peer.lnlikelihood(td['k'], td['ldc'], td['t0'], td['p'], td['a'], td['i'], td['e'], td['w'], sigma=sig_d)
            loc = m.lnlikelihood(td['k'], td['ldc'], td['t0'], td['p'], td['a'], td['i'], td['e'], td['w'], sigma=sig_d, copy=False)
            if gathered is not None:      # the one collective of the path: all-gather of lnL over NCCL/NVLink
                dist.all_gather_into_tensor(gathered, loc)
                return gathered
            return loc
        return m.evaluate(td['k'], td['ldc'], td['t0'], td['p'], td['a'], td['i'], td['e'], td['w'], copy=False)

    # the e2e loop alternates between two different populations (every vector gets its neighbour's
    # parameters), so that the delta host transfer re-sends AND resets every transit window each step
    names = ('k', 'ldc', 't0', 'p', 'a', 'i', 'e', 'w')
    pops = [tuple(np.ascontiguousarray(getattr(c, k)) for k in names),
            tuple(np.ascontiguousarray(np.roll(getattr(c, k), 1, axis=0)) for k in names)]
    nhost = [0]

    def step_host():
        nhost[0] += 1
        a = pops[nhost[0] & 1]
        if lnl:
            return m.lnlikelihood(*a, sigma=c.sigma, copy=True)
        return m.evaluate(*a, copy=True)

    h2d = sum(np.asarray(getattr(c, k)).nbytes for k in ('k', 'ldc', 't0', 'p', 'a', 'i', 'e', 'w')) + (c.sigma.nbytes if lnl else 0)
    pts = c.npv * c.npt
    kname = 'k_rr_points_ss' if int(np.max(c.nsamples)) > 1 else 'k_rr_points'     # supersampled data sets have their own kernel
    return m, step_device, step_host, pts, h2d, (kname + '<LNL>' if lnl else kname), (None if lnl else float(esize) * pts)


def _table_args(c):
    prof, (x0, dx), (y0, dy), (z0, dz) = c.table
    return prof, x0, dx, y0, dy, z0, dz



def run_collective(args, rank, world, local_rank, dev, dist):
    """The one exchange of the path, measured in the same run as the headline: the C5 shard (eccentric 'power-2',
    8192 vectors x 100 000 points per GPU) through the fused likelihood, lnL[npv] all-gathered over NVLink by peer
    stores from the finishing kernel, ranks ordered by device-side flags (no host-issued barrier).  Also timed:
    the same step without any exchange (what the gather costs), the round-1 symmetric-memory barrier variant and
    the NCCL all_gather_into_tensor variant.  Device-resident inputs, CUDA events, barrier + synchronize on both
    sides, max over ranks."""
    import torch
    import pytransit_b200 as pb
    from pytransit_b200.distributed import PeerLnLGather
    c, desc = workload('c5', rank)
    m = pb.RoadRunnerModelCUDA(c.ldmodel, device=local_rank)
    m.set_data(torch.as_tensor(c.time, device=dev))
    m.set_obs(torch.as_tensor(c.obs, device=dev))
    td = {k: torch.as_tensor(np.ascontiguousarray(getattr(c, k)), device=dev) for k in ('k', 'ldc', 't0', 'p', 'a', 'i', 'e', 'w', 'sigma')}
    a = (td['k'], td['ldc'], td['t0'], td['p'], td['a'], td['i'], td['e'], td['w'])
    steps = max(10, min(args.steps, 100))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            out = fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            out = fn()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, out

    local_ms, loc = timed(lambda: m.lnlikelihood(*a, sigma=td['sigma'], copy=False))
    m.set_profiling(True)
    for _ in range(5):
        m.lnlikelihood(*a, sigma=td['sigma'], copy=False)
    torch.cuda.synchronize()
    n, _, kp = m.timing_summary()
    m.set_profiling(False)
    res = {'workload': desc, 'n_gpus': world, 'steps': steps, 'unit': UNIT, 'dtype': 'f64',
           'per_gpu': f'npv={c.npv} x npt={c.npt}', 'local_only_ms_per_step': local_ms, 'kernel_ms': kp / n}
    pts = world * c.npv * c.npt
    if world == 1:
        res.update({'value': pts / (local_ms * 1e-3), 'ms_per_step': local_ms, 'gather': 'none (one GPU: the shard is the population)'})
        return res
    variants = {}
    gathered = torch.empty((world * c.npv,), dtype=torch.float64, device=dev)

    def step_nccl():
        dist.all_gather_into_tensor(gathered, m.lnlikelihood(*a, sigma=td['sigma'], copy=False))
        return gathered

    variants['nccl_all_gather_ms_per_step'], ref = timed(step_nccl)
    ref = ref.clone()
    pg_b = PeerLnLGather(m, c.npv, sync='barrier')
    variants['peer_stores_host_barrier_ms_per_step'], _ = timed(lambda: pg_b.lnlikelihood(*a, sigma=td['sigma']))
    pg = PeerLnLGather(m, c.npv, sync='flags')
    ms, got = timed(lambda: pg.lnlikelihood(*a, sigma=td['sigma']))
    same = bool(torch.equal(got, ref))
    m.gather_status()
    res.update({'value': pts / (ms * 1e-3), 'ms_per_step': ms,
                'gather': 'peer stores from k_lnl_finish into every rank\'s symmetric-memory array over NVLink, ranks ordered by '
                          'device-side release/acquire flags (the last block of k_lnl_finish waits for all peers); 8 B per vector',
                'gather_cost_ms_per_step': ms - local_ms, 'efficiency_vs_local_only': local_ms / ms,
                'bit_identical_to_nccl_all_gather': same, 'variants': variants})
    return res


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    from pytransit_b200.distributed import bind_to_gpu_numa
    cpus_all = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    numa_bound = bind_to_gpu_numa(local_rank)      # page-locked result arrays land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))

    c, desc = workload(args.workload, rank)
    lnl = args.workload == 'c5'
    if args.workload == 'c4':
        import pytransit_b200 as pb
        mu = pb.TSModelCUDA('uniform', device=local_rank).mu      # the model's own mu grid (device-built tables)
        c.table = wl.ldtk_style_table(c.npb, mu)
        c.npb_out = c.npb
    m, step_device, step_host, pts_per_step, h2d, kname, alg_bytes = make_steps(args, c, dev, local_rank, world, dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    del out
    barrier()
    # launch-bound single-vector workload: the timed loop runs without per-kernel events (they would also keep the
    # library from replaying the call as a CUDA graph); kernel durations come from a short separate loop afterwards
    separate_kernel_timing = args.workload == 'c1' and not args.no_kernel_timing
    m.set_profiling(not args.no_kernel_timing and not separate_kernel_timing)   # events only; nothing synchronises inside the timed loop
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = m.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_setup_ms, k_points_ms = [], []
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    launches = m.launch_count - launches0
    if separate_kernel_timing:
        m.set_profiling(True)
        for _ in range(min(args.steps, 50)):
            out = step_device()
        torch.cuda.synchronize()
    if not args.no_kernel_timing:
        n, a, b = m.timing_summary()
        k_setup_ms, k_points_ms = [a / n], [b / n]
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    del out
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * pts_per_step * args.steps / (ms_max * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ---------------------------------
    m.set_profiling(False)
    e2e_steps = max(2, min(args.steps, 20 if args.workload not in ('c3', 'c4') else 4))

    def measure_e2e(mode):
        """K calls of the public API with HOST numpy inputs and a host result; every result is read and dropped before the
        next call, as a consumer would (so one pooled page-locked buffer serves the loop and, in delta mode, has to
        follow the alternating populations)."""
        if hasattr(m, 'host_result'):
            m.host_result = mode
        try:
            chk = 0.0
            for _ in range(3):
                r = step_host()
                chk += float(r.reshape(-1)[-1])
                del r
            barrier()
            t0 = time.perf_counter()
            ev0.record()
            for _ in range(e2e_steps):
                r = step_host()
                nbytes = int(r.nbytes)
                chk += float(r.reshape(-1)[-1])      # the consumer reads the result ...
                del r                                # ... and lets go of it before the next call
            ev1.record()
            barrier()
            e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d2h = nbytes
            how = "host_result='copy' (the default): one full device-to-host copy of the result per call into a writable array of the caller's own"
            if lnl:
                how = 'lnL[npv] only: the flux is never materialised'
            elif mode == 'delta':
                last, ndelta, nfull = m.host_result_stats
                d2h = int(last)
                how = ("host_result='delta' (documented opt-in of the public API; results are read-only, never aliased): the full "
                       "result (%d bytes) is current in the caller's page-locked array after every step, but only the 16-point "
                       "blocks that differ from 1.0 now or did the last time this buffer was written cross PCIe (written by the "
                       "GPU); the timed steps alternate between two different populations; %d delta / %d full transfers so far"
                       % (nbytes, ndelta, nfull))
            return {'value': world * pts_per_step * e2e_steps / (float(t.item()) * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': d2h, 'result_bytes_per_step': nbytes,
                    'steps': e2e_steps, 'mode': mode if not lnl else 'lnl', 'd2h': how}
        except MemoryError as ex:   # page-locked result buffer too large for this host
            return {'value': None, 'unit': UNIT, 'error': str(ex)[:200]}

    # concurrent device-to-host copy rate of all ranks (copy engines, sequential 256 MB blocks): the PCIe ceiling the host
    # delivery is measured against
    pin = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pin.copy_(src, non_blocking=True)
    barrier()
    ev0.record()
    for _ in range(4):
        pin.copy_(src, non_blocking=True)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pcie_gbs = 4 * (256 << 20) / (float(t.item()) * 1e-3) / 1e9
    del pin, src

    def with_pcie_roofline(e):
        if e and e.get('value'):
            rate = e['d2h_bytes_per_step'] * e['value'] / (world * pts_per_step) / 1e9      # GB/s per GPU that crossed PCIe
            e['roofline'] = {'bound': 'pcie (device-to-host)', 'achieved': rate, 'peak': pcie_gbs, 'unit': 'GB/s per GPU',
                             'frac': rate / pcie_gbs,
                             'peak_source': f'measured in this run: {world} rank(s) copying 256 MB blocks device-to-host concurrently, max over ranks'}
        return e

    e2e = with_pcie_roofline(measure_e2e(args.host_result))
    e2e_copy = None
    if args.host_result == 'delta' and not lnl and args.workload != 'c4':
        # the same loop through the DEFAULT mode of the API (plain full copy, writable result): PCIe bound
        e2e_copy = with_pcie_roofline(measure_e2e('copy'))

    m_fp64_peak = None
    if lnl and rank == 0:
        try:
            m_fp64_peak = m.measure_fp64_peak()
        except Exception:
            m_fp64_peak = None

    # ---- the path's one collective, same run (default workload only) -----------------------------------
    collective = None
    if args.workload == 'c2' and not args.no_collective:
        del m, step_device, step_host
        torch.cuda.empty_cache()
        collective = run_collective(args, rank, world, local_rank, dev, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = measured_peaks()
    roof = None
    if k_points_ms:
        kp = float(np.mean(k_points_ms))
        ks = float(np.mean(k_setup_ms))
        counters = None
        if world == 1 and not args.no_counters:
            counters = ncu_counters(args, 'k_ts_flux' if args.workload == 'c4' else 'k_rr_points')
        have = counters is not None and 'unavailable' not in counters
        traffic, traffic_src = None, None
        if have and 'dram__bytes_write.sum' in counters:
            traffic = counters['dram__bytes_read.sum'] + counters['dram__bytes_write.sum']
            traffic_src = 'dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu child process of this run'
        else:
            tr = ROOT / 'profiles' / 'traffic.json'
            if tr.exists():
                try:
                    traffic = json.loads(tr.read_text()).get(args.workload)
                    traffic_src = 'profiles/traffic.json (earlier ncu --set full capture; no counters in this run: %s)' % \
                                  ((counters or {}).get('unavailable', 'N > 1 or --no-counters'))
                except Exception:
                    pass
        if alg_bytes is None:
            # Fused likelihood: nothing is written, time + obs (1.6 MB) stream from L2 -- the kernel is bound by instruction
            # issue on the fp64 pipe.  Roofline on EXECUTED work: fp64-pipe warp instructions of one launch (ncu counter)
            # x 32 lanes x 2 flop per FMA slot, over the kernel's event-timed duration, against the DFMA throughput
            # measured on this GPU (ptb_measure_fp64_peak).  The algorithmic gain of skipping untouched blocks (their
            # chi^2 is pre-summed) is reported separately as points per second; it is not part of the fraction.
            peak_tf = m_fp64_peak
            roof = {'bound': 'fp64 pipe (instruction issue); no HBM or tensor-core bound applies: nothing is written', 'kernel': kname,
                    'peak': peak_tf, 'unit': 'TFLOP/s', 'peak_source': 'measured on this GPU: DFMA microbenchmark k_dfma_peak (ptb_measure_fp64_peak)',
                    'traffic': traffic, 'traffic_source': traffic_src, 'kernel_ms': kp, 'setup_ms': ks,
                    'kernel_share_of_step': kp / (ms_max / args.steps), 'gpoints_per_s': pts_per_step / (kp * 1e-3) / 1e9}
            if have and 'sm__inst_executed_pipe_fp64.sum' in counters:
                f64 = counters['sm__inst_executed_pipe_fp64.sum']
                ach = f64 * 32 * 2 / (kp * 1e-3) / 1e12
                roof.update({'achieved': ach, 'frac': ach / peak_tf if peak_tf else None,
                             'executed_fp64_warp_inst_per_launch': f64, 'executed_warp_inst_per_launch': counters.get('smsp__inst_executed.sum'),
                             'fp64_pipe_active_pct': counters.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
                             'issue_active_pct': counters.get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                             'executed_fp64_lane_ops_per_point': f64 * 32 / pts_per_step,
                             'note': 'achieved = executed fp64-pipe warp instructions x 64 flop / kernel time; the kernel folds only the '
                                     '~7 % of points in blocks a transit window touches'})
            else:
                roof.update({'achieved': None, 'frac': None, 'note': 'no counters in this run: ' + str((counters or {}).get('unavailable', 'N > 1 or --no-counters'))})
        else:
            ach = alg_bytes / (kp * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'kernel': kname, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                    'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': kp, 'setup_ms': ks,
                    'kernel_share_of_step': kp / (ms_max / args.steps),
                    'step_frac_of_peak': alg_bytes / (ms_max / args.steps * 1e-3) / 1e9 / peak}
            if have:
                roof['issue_active_pct'] = counters.get('smsp__issue_active.avg.pct_of_peak_sustained_active')
                roof['fp64_pipe_active_pct'] = counters.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')

    # ---- CPU baseline on a bounded sample (N=1 only) -----------------------------------------------
    if cpus_all is not None:
        os.sched_setaffinity(0, cpus_all)          # the CPU arms use every host core again
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.lib()
        ncores = host_threads(orc)
        tab = orc.Tables()
        c0, _ = workload(args.workload, 0)
        if args.workload == 'c4':
            c0.table, c0.npb_out = c.table, c.npb
        rows = cpu_sample_rows(orc, tab, c0, lnl, target_s=args.cpu_seconds)
        best, total, passes = None, 0.0, 0
        while passes < 2 or total < args.cpu_seconds:      # about cpu_seconds of CPU work in all
            t1 = time.perf_counter()
            n = oracle_step(orc, tab, c0, rows, lnl)
            dt = time.perf_counter() - t1
            best = dt if best is None else min(best, dt)
            total += dt
            passes += 1
        cpu = {'value': n / best, 'unit': UNIT, 'cores': ncores, 'kind': 'port',
               'sample': f'{rows} of {c0.npv} parameter vectors x {n // rows} points each, best of {passes} passes '
                         f'({best:.3f} s best, {total:.1f} s of CPU work in all)'}

    cpu_numba = None
    if world == 1 and not args.no_cpu and not args.no_numba:
        c0, _ = workload(args.workload, 0)
        if args.workload == 'c4':
            c0.table, c0.npb_out = c.table, c.npb
        cpu_numba = numba_baseline(args.workload, c0, seconds=args.cpu_seconds)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64' if args.precision == 'fp64' else 'f32 (opt-in mode: fp64 phase fold, fp32 samples and output)', 'data': 'synthetic',
            'config': config_dict(args, desc, world), 'per_gpu': f'npv={c.npv} x npt={c.npt}' + (f' x npb={c.npb}' if args.workload == 'c4' else ''),
            'clocks': clocks, 'e2e': e2e, 'e2e_copy': e2e_copy, 'numa_bound': numa_bound,
            'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu, 'cpu_baseline_numba': cpu_numba,
            'collective': collective}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Everything except the result line goes to stderr: libraries (NCCL's version banner, torchrun notices) write to
    file descriptor 1 behind Python's back, and the contract is ONE JSON line on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-numba', action='store_true', help="skip timing the reference's own Numba path (baseline/_ref)")
    ap.add_argument('--no-kernel-timing', action='store_true')
    ap.add_argument('--no-counters', action='store_true', help='skip the ncu child process that measures DRAM traffic / executed instructions of the dominant kernel')
    ap.add_argument('--child-profile', action='store_true', help=argparse.SUPPRESS)
    ap.add_argument('--no-collective', action='store_true', help='skip the C5-shard fused lnL + all-gather block of the default line')
    ap.add_argument('--gather', default='peer', choices=['peer', 'peer-barrier', 'nccl'],
                    help='N>1 lnL workloads (--workload c5): peer stores ordered by device-side flags (default), peer stores + one '
                         'symmetric-memory barrier per step, or NCCL all_gather')
    ap.add_argument('--host-result', default='delta', choices=['delta', 'copy'], help='e2e: delta transfer (default) or plain full copy')
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'fp32'], help="'fp32' = the opt-in single-precision mode (c2/c3/c5 only)")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.child_profile:
        run_child_profile(args, local_rank)
    elif args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
