"""N>1 host path on CPU: two gloo ranks each run their shard of one ensemble (through the test double of libhsp2g) and
the gathered result equals the single-process run column for column -- segments are independent, so sharding must not
change a bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_shard(ens, cal, steps):
    from hspsquared_b200 import synth
    from hspsquared_b200.engine import LandBatch

    with LandBatch(ens.store(cal), cal, list(synth.PERLND_FORCINGS), ["PWATER_SURO", "PWATER_AGWO", "SNOW_PACKF"]) as b:
        return b.run(synth.forcing_chunk(ens, cal, 0, steps), chunk=100)


def _worker(rank, world, port, emu_path, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from hspsquared_b200 import _lib, synth
    from hspsquared_b200 import dist as D
    from hspsquared_b200.calendar import Calendar

    _lib.LIB_PATH = emu_path  # test double (tests/emu/README.md)
    _lib.ALLOW_BUILDS.add("host-emu")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cal = Calendar("1976-01-01", "1976-01-15", 60, "us")
    lo, hi = D.shard_bounds(n_total, rank, world)
    ens = synth.Ensemble("PERLND", n_total, seed=3, sections=("SNOW", "PWATER")).shard(lo, hi)
    out = _run_shard(ens, cal, cal.steps)
    full = D.gather_columns(torch.from_numpy(out), n_total)
    dist.barrier()
    if rank == 0:
        q.put(full.numpy())
    dist.destroy_process_group()


def test_shard_bounds():
    from hspsquared_b200.dist import shard_bounds

    for n, w in ((10, 3), (7, 8), (100000, 8), (1, 1)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_rank_shards_equal_single_process(emu_library):
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar

    n_total = 45
    cal = Calendar("1976-01-01", "1976-01-15", 60, "us")
    whole = _run_shard(synth.Ensemble("PERLND", n_total, seed=3, sections=("SNOW", "PWATER")), cal, cal.steps)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_library, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got, whole, equal_nan=True)
