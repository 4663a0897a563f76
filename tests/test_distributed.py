"""CPU tests of the multi-GPU host logic under the gloo backend (world_size 2): population sharding and the
all-gather of the per-vector log likelihoods.  The per-rank compute is the CPU oracle here (tests may use
it as the checker); on the GPU box the same PopulationSharder drives RoadRunnerModelCUDA over NCCL
(bench.py --gpus N --workload c5)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pytransit_b200.distributed import shard_bounds, shard_population  # noqa: E402


def test_shard_bounds_cover_and_balance():
    for npv in (1, 7, 8, 13, 8192, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(npv, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == npv
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_shard_population_slices_only_population_arrays():
    npv = 10
    k = np.arange(npv * 2.0).reshape(npv, 2)
    ldc_shared = np.array([[0.2, 0.1], [0.3, 0.2]])
    s = shard_population(npv, 3, 1, k=k, ldc=ldc_shared, t0=np.arange(npv * 1.0), p=3.5, e=np.zeros(npv))
    lo, hi = shard_bounds(npv, 3, 1)
    assert np.array_equal(s['k'], k[lo:hi]) and np.array_equal(s['t0'], np.arange(lo, hi))
    assert s['ldc'] is ldc_shared and s['p'] == 3.5 and s['e'].shape == (hi - lo,)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, npv, npt, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import workloads as wl
        from oracle import oracle as orc
        from pytransit_b200.distributed import PopulationSharder
        orc.set_threads(1)
        tab = orc.Tables()
        c = wl.config5(npv=npv, npt=npt)

        def lnl_fn(k, ldc, t0, p, a, i, e, w, sigma):
            ldp, istar = orc.evaluate_ld('power-2', tab.mu, ldc)
            flux = orc.rr_full(tab, c.time, k, t0, p, a, i, e, w, c.lcids, c.pbids, c.epids, c.nsamples, c.exptimes,
                               ldp, istar)
            return orc.lnlike_normal(c.obs, flux, sigma, c.slices, c.nids)

        sh = PopulationSharder()
        assert (sh.world, sh.rank) == (world, rank)
        full = sh.lnlikelihood(lnl_fn, npv, k=c.k, ldc=c.ldc, t0=c.t0, p=c.p, a=c.a, i=c.i, e=c.e, w=c.w, sigma=c.sigma)
        ref = lnl_fn(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w, c.sigma)
        q.put((rank, bool(np.array_equal(full, ref)), full.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('npv', [8, 11, 1])      # even and ragged shards; one vector: rank 1 has an empty block
def test_allgather_lnl_world2_gloo(npv):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, npv, 600, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] == (npv,) for r in res)


def test_shard_population_goes_by_name_and_rank_not_by_size():
    """npb == npv or nblocks == npv must not get the shared arrays sliced."""
    from pytransit_b200.distributed import shard_population
    npv = npb = nblocks = 4
    rng = np.random.default_rng(0)
    k_shared, ldc_shared, sig_shared = rng.random(npb), rng.random((npb, 2)), rng.random(nblocks)
    k_pv, ldc_pv, sig_pv = rng.random((npv, npb)), rng.random((npv, npb, 2)), rng.random((npv, nblocks))
    p = rng.random(npv)
    s = shard_population(npv, 2, 1, k=k_shared, ldc=ldc_shared, sigma=sig_shared, p=p, t0=rng.random((npv, 3)), _sigma_blocks=nblocks)
    assert s['k'] is k_shared and s['ldc'] is ldc_shared and s['sigma'] is sig_shared
    assert s['p'].shape == (2,) and s['t0'].shape == (2, 3) and '_sigma_blocks' not in s
    s = shard_population(npv, 2, 1, k=k_pv, ldc=ldc_pv, sigma=sig_pv, p=p)
    assert np.array_equal(s['k'], k_pv[2:]) and np.array_equal(s['ldc'], ldc_pv[2:]) and np.array_equal(s['sigma'], sig_pv[2:])
    # one noise block: a 1-D sigma[npv] is per vector
    assert shard_population(npv, 2, 0, sigma=rng.random(npv))['sigma'].shape == (2,)
    # scalars and leading-1 arrays pass through
    s = shard_population(npv, 2, 0, e=0.0, k=np.full((1, 1), 0.1))
    assert s['e'] == 0.0 and s['k'].shape == (1, 1)
