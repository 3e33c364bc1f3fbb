"""The N > 1 path on CPU: world_size-2 (and 3) gloo process groups exercise the host logic of
lens_flare_b200/sharding.py -- the per-rank job shards the C ABI hands out (lfb_list_jobs, host-only) and the
sum-reduce of the int64 sensor buffers.  The per-rank accumulators are produced here by the ORACLE (as the
stand-in for the GPU engine, which is what -m gpu tests cover): the reduced frame must equal the unsharded one
bit for bit, for any rank count."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, reduce_dst, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lens_flare_b200 import capi, sharding
    from oracle import bindings as ob
    port_oracle = ob.PortOracle()
    z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
    tex = z["pent_11"].astype(np.float32) * np.float32(1.0 / 255.0)
    lens = capi.builtin_lens(3, 550.0)
    lights = [capi.make_light(0.45, 0.55, theta=0.06), capi.make_light(0.7, 0.3, theta=0.1, radiance=(0.5, 1.0, 2.0))]
    params = capi.make_params(capi.MODE_EXACT_GRID, 96, 64, grid_n=16, pair_set=capi.PAIRS_ALL, include_direct=1, px_per_unit=0.1)
    mine = sharding.shard_params(params, rank, world)
    # the shards the C ABI deals out are a partition of the frame's jobs
    jobs = capi.list_jobs(lens, mine, len(lights))
    gathered = [None] * world
    dist.all_gather_object(gathered, [tuple(j) for j in jobs])
    if rank == 0:
        every = sorted(j for part in gathered for j in part)
        assert every == sorted(tuple(j) for j in capi.list_jobs(lens, params, len(lights)))
    _, acc = port_oracle.render(lens, tex, lights, mine, want_accum=True)
    t = torch.from_numpy(acc)
    sharding.reduce_accum(t, dst=reduce_dst)
    if reduce_dst is None or rank == reduce_dst:
        np.save(os.path.join(out_dir, f"reduced_{rank}.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,reduce_dst", [(2, None), (2, 0), (3, None)])
def test_sharded_frame_reduces_to_the_whole(port, apertures, tmp_path, world, reduce_dst):
    from lens_flare_b200 import capi
    mp.spawn(_worker, args=(world, _free_port(), reduce_dst, str(tmp_path)), nprocs=world, join=True)
    lens = capi.builtin_lens(3, 550.0)
    lights = [capi.make_light(0.45, 0.55, theta=0.06), capi.make_light(0.7, 0.3, theta=0.1, radiance=(0.5, 1.0, 2.0))]
    params = capi.make_params(capi.MODE_EXACT_GRID, 96, 64, grid_n=16, pair_set=capi.PAIRS_ALL, include_direct=1, px_per_unit=0.1)
    _, whole = port.render(lens, apertures["pent_11"], lights, params, want_accum=True)
    assert whole.any()
    ranks = range(world) if reduce_dst is None else [reduce_dst]
    for r in ranks:
        assert np.array_equal(np.load(tmp_path / f"reduced_{r}.npy"), whole)


def test_single_rank_is_a_no_op():
    from lens_flare_b200 import capi, sharding
    p = capi.make_params(capi.MODE_EXACT_GRID, 8, 8, grid_n=4)
    q = sharding.shard_params(p, 0, 1)
    assert (q.shard_index, q.shard_count) == (0, 0)
    t = torch.arange(6, dtype=torch.int64)
    assert sharding.reduce_accum(t) is t
