"""Multi-GPU parity check, launched with torchrun on a box with >= 2 GPUs (tests/test_multigpu.py does that when it can):
the sharded frame -- NCCL reduce path (ShardedFlare), fused peer-memory path (PeerFlare, with and without NVSwitch
multicast) and the tile-sparse peer path (PeerSparse) -- must equal the unsharded single-GPU frame bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_flare_b200 import capi, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
    tex = z["pentbig500_14"].astype(np.float32) * np.float32(1.0 / 255.0)
    lens = capi.builtin_lens(3, 550.0)
    eng = capi.Engine(local)
    eng.set_lens(lens)
    eng.set_aperture(tex)
    W, H = 960, 540
    params = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=128, pair_set=capi.PAIRS_ALL, include_direct=1)
    lights = [capi.make_light(0.45, 0.55, theta=capi.physical_theta(0.45, 0.55)), capi.make_light(0.7, 0.3, theta=0.1, radiance=(0.5, 1.0, 2.0))]
    whole = torch.from_numpy(eng.render_ghosts(lights, params, elem=capi.F32x3)).to(dev)  # unsharded, on every rank
    ok = True
    # --- NCCL path
    sh = sharding.ShardedFlare(eng, params, rank, world, dev, n_buffers=2)
    out = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    sh.begin()
    for _ in range(3):
        sh.frame(lights, out=out, elem=capi.F32x3, reduce_dst=0)
    sh.join()
    torch.cuda.synchronize()
    if rank == 0:
        same = torch.equal(out, whole)
        print("nccl reduce path == single GPU:", same)
        ok &= same
    # --- fused peer-memory path, plain peer loads then NVSwitch multicast
    fin = capi.Engine(local)
    for mc, two in ((False, False), (False, True), (True, True)):
        pf = sharding.PeerFlare(eng, params, rank, world, dev, dist.group.WORLD, n_buffers=3, use_multicast=mc,
                                finalize_engine=fin if two else None)
        if mc and not pf.mc:
            if rank == 0:
                print("multicast not supported on this box: skipped")
            continue
        pf.begin()
        for _ in range(7):
            b = pf.frame(lights, owner=0)
        pf.finish()
        torch.cuda.synchronize()
        if rank == 0:
            same = all(torch.equal(pf.result(q), whole) for q in range(3))
            print(f"peer path (multicast={mc}, two streams={two}) == single GPU:", same)
            ok &= same
    # --- tile-sparse peer path (the default of bench.py at N > 1): moving suns, so tiles of earlier frames must be re-zeroed
    suns = [[capi.make_light(0.3 + 0.05 * k, 0.4 + 0.03 * k, theta=0.05 + 0.01 * k, radiance=(1.0, 0.5 + 0.1 * k, 2.0 - 0.2 * k)),
             capi.make_light(0.7, 0.3, theta=0.1, radiance=(0.5, 1.0, 2.0))] for k in range(7)]
    wants = [torch.from_numpy(eng.render_ghosts(lt, params, elem=capi.F32x3)).to(dev) for lt in suns]
    for two in (False, True):
        ps = sharding.PeerSparse(eng, params, rank, world, dev, dist.group.WORLD, n_buffers=3, finalize_engine=fin if two else None)
        got = []
        ps.begin()
        for k, lt in enumerate(suns):
            b = ps.frame(lt, owner=0)
            if k >= 4:  # read every frame of the tail (a buffer's content is final only after a finish)
                ps.finish()
                torch.cuda.synchronize()
                got.append((k, ps.result(b).clone()))
        if rank == 0:
            same = all(torch.equal(g, wants[k]) for k, g in got)
            print(f"tile-sparse peer path (two streams={two}) == single GPU:", same)
            ok &= same
    # --- the same into shared page-locked HOST frames, frames in flight: direct stores by the reduce kernels, and the staged
    # reduce + paced drain (lfb_reduce_tiles_peers_staged / lfb_drain_tiles); a frame is read right after wait_frame()
    from multiprocessing import shared_memory
    cpu_group = dist.new_group(backend="gloo")
    frame_bytes, R = H * W * 24, 3
    name = [None]
    if rank == 0:
        shm = shared_memory.SharedMemory(create=True, size=R * frame_bytes)
        name[0] = shm.name
    dist.broadcast_object_list(name, src=0, group=cpu_group)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=name[0])
        try:
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    host = np.frombuffer(shm.buf, dtype=np.float64, count=R * H * W * 3).reshape(R, H, W, 3)
    L = capi.lib()
    assert L.lfb_host_register(host.ctypes.data, host.nbytes) == capi.OK
    base = L.lfb_host_device_pointer(host.ctypes.data)
    wants64 = [eng.render_ghosts(lt, params) for lt in suns] if rank == 0 else None
    drn = capi.Engine(local, stream_priority=1)
    for drain in (None, drn):
        if rank == 0:
            host[...] = 0.0
        dist.barrier(group=cpu_group)
        ps = sharding.PeerSparse(eng, params, rank, world, dev, dist.group.WORLD, n_buffers=R, finalize_engine=fin,
                                 host_out_ptrs=[base + b * frame_bytes for b in range(R)], drain_engine=drain, host_stride=24)
        same = True
        ps.begin()
        bufs = []
        for k, lt in enumerate(suns):
            if k >= 2:
                ps.wait_frame(k - 2)  # frame k - 2 is complete on every rank
                if rank == 0:
                    same &= bool(np.array_equal(host[bufs[k - 2]], wants64[k - 2]))
            bufs.append(ps.frame(lt, owner=0, elem=capi.F64x3, stride=24))
        ps.finish()
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            for k in (len(suns) - 2, len(suns) - 1):
                same &= bool(np.array_equal(host[bufs[k]], wants64[k]))
            print("tile-sparse peer path into host frames, two in flight (%s) == single GPU:" % ("staged + paced drain" if drain else "direct stores"), same)
            ok &= same
        del ps
        dist.barrier(group=cpu_group)
    drn.close()
    L.lfb_host_unregister(host.ctypes.data)
    del host
    shm.close()
    dist.barrier(group=cpu_group)
    if rank == 0:
        shm.unlink()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    fin.close()
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
