"""Multi-GPU check of the fused likelihood + peer-memory all-gather (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_peer_gather_check.py

Every rank evaluates its shard of a seeded population with RoadRunnerModelCUDA; PeerLnLGather must leave the
same lnL[npv] on every rank as (1) the NCCL all-gather path (PopulationSharder), bit for bit, and (2) the
unsharded evaluation on one GPU to summation-order accuracy (the time axis is chunked by population size, so
the chi^2 partial sums group differently: ~1e-16 relative)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    import workloads as wl
    import pytransit_b200 as pb
    from pytransit_b200.distributed import PeerLnLGather, PopulationSharder, shard_population

    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    npv_local, npt = 96, 30_000
    npv = npv_local * world
    c = wl.config5(npv=npv, npt=npt)
    m = pb.RoadRunnerModelCUDA('power-2', device=local)
    m.set_data(c.time)
    m.set_obs(c.obs)
    args = dict(k=c.k, ldc=c.ldc, t0=c.t0, p=c.p, a=c.a, i=c.i, e=c.e, w=c.w, sigma=c.sigma)
    full = m.lnlikelihood(**args).copy()                                   # unsharded, this GPU
    shard = shard_population(npv, world, rank, **args)
    gathers = {mode: PeerLnLGather(m, npv_local, sync=mode) for mode in ('flags', 'barrier')}
    nccl = PopulationSharder().lnlikelihood(lambda **kw: m.lnlikelihood(copy=False, **kw), npv, **args).cpu().numpy()
    np.testing.assert_allclose(nccl, full, rtol=1e-12, atol=1e-9)   # lnL = cst - chi2/2 cancels: compare to the size of the terms
    for mode, pg in gathers.items():
        for it in range(5):                                                # repeated calls reuse the buffers
            sh = dict(shard, sigma=shard['sigma'] * (1.0 + it))            # new values every step: stale slots would show
            want = PopulationSharder().lnlikelihood(lambda **kw: m.lnlikelihood(copy=False, **kw), npv,
                                                    **dict(args, sigma=args['sigma'] * (1.0 + it))).cpu().numpy()
            got = pg.lnlikelihood(**sh).clone()                            # consumer on the same stream, no host sync
            got = got.cpu().numpy()
            assert np.array_equal(got, want), (mode, rank, it, np.abs(got - want).max())
        m.gather_status()
    dist.barrier()
    if rank == 0:
        print(f'peer gather ok ({", ".join(gathers)}): world={world} npv={npv} max|lnL|={np.abs(full).max():.3e}')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
