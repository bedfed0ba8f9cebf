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


# ---- BASELINE.json configs[2] in the test suite: ONE PERLND + IMPLND ensemble split with dist.shard_bounds over two ranks
def _ensemble_results(rank, world, n_p, n_i, device, steps=48):
    """-> (PERLND columns, IMPLND columns): per-segment sums over the run of three saved series + the final state of this rank's
    shard of each operation's ensemble (exactly the numbers a calibration run keeps per ensemble member)"""
    import numpy as np

    from hspsquared_b200 import dist as D
    from hspsquared_b200 import synth
    from hspsquared_b200.calendar import Calendar

    cal = Calendar("1976-02-20", "1976-02-22", 60, "us")
    out = []
    for op, n_total, secs, save in (("PERLND", n_p, ("SNOW", "PWATER"), ["SNOW_PACKF", "PWATER_SURO", "PWATER_AGWO"]),
                                    ("IMPLND", n_i, ("SNOW", "IWATER"), ["SNOW_PACKF", "IWATER_SURO", "IWATER_RETS"])):
        ens = synth.Ensemble(op, n_total, seed=11, sections=secs)
        with D.ShardedBatch(lambda lo, hi: ens.shard(lo, hi).store(cal), n_total, cal, list(synth.PERLND_FORCINGS), save,
                            device=device) as sb:
            shard = ens.shard(sb.lo, sb.hi)
            o = sb.batch.run(synth.forcing_chunk(shard, cal, 0, steps), chunk=24)
            cols = np.concatenate([np.nansum(o, axis=0), sb.batch.get_state()], axis=0)  # [3 + states][n_local]
            out.append((sb, cols))
    return out


def _gpu_worker(rank, world, port, n_p, n_i, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    for sb, cols in _ensemble_results(rank, world, n_p, n_i, device=0):  # a one-GPU box: both ranks share cuda:0
        res.append(sb.gather(torch.from_numpy(cols)).numpy())
    dist.barrier()
    if rank == 0:
        q.put(res)
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_one_perlnd_implnd_ensemble_sharded_over_two_ranks_on_gpu():
    """configs[2] through the CUDA library: a 120,000 + 80,000-segment PERLND + IMPLND ensemble split with dist.shard_bounds over
    two ranks (processes), gathered, equals the one-process run column for column, bit for bit -- per-segment sums of the saved
    series and the complete final state."""
    from hspsquared_b200 import _lib

    assert torch.cuda.is_available() and _lib.build_kind() == "cuda/sm_100a"
    n_p, n_i = 120_000, 80_000
    whole = [cols for _sb, cols in _ensemble_results(0, 1, n_p, n_i, device=0)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, n_p, n_i, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for g, w in zip(got, whole):
        assert g.shape == w.shape and np.array_equal(g, w, equal_nan=True)


def _emu_worker(rank, world, port, emu_path, n_p, n_i, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from hspsquared_b200 import _lib

    _lib.LIB_PATH = emu_path
    _lib.ALLOW_BUILDS.add("host-emu")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = [sb.gather(torch.from_numpy(cols)).numpy() for sb, cols in _ensemble_results(rank, world, n_p, n_i, device=0)]
    dist.barrier()
    if rank == 0:
        q.put(res)
    dist.destroy_process_group()


def test_one_perlnd_implnd_ensemble_sharded_over_two_ranks(emu_library):
    """The same on the CPU test double with gloo (world_size 2): host-side sharding logic of configs[2]."""
    n_p, n_i = 75, 41
    whole = [cols for _sb, cols in _ensemble_results(0, 1, n_p, n_i, device=0)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_emu_worker, args=(r, 2, port, emu_library, n_p, n_i, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for g, w in zip(got, whole):
        assert g.shape == w.shape and np.array_equal(g, w, equal_nan=True)
