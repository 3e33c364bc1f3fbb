"""Multi-GPU sharding of a flare frame: one process per GPU, (light x ghost pair x wavelength)
jobs dealt to ranks by the engine (lfb_params.shard_index / shard_count, LPT order), each rank
splatting into its own full-frame fixed-point accumulators, then ONE sum-reduce of the int64
sensor buffers (NCCL over NVLink on the GPU box; gloo in the CPU tests).

Integer sums are order-independent, so the reduced frame is bit-identical for 1, 2, 4 or 8
ranks.  torch / torch.distributed are plumbing here (device memory, process group): every
pixel is produced by liblfb200.so.
"""
import torch
import torch.distributed as dist

from . import capi


def shard_params(params, rank, world_size):
    """This rank's view of the frame: same frame description, its share of the job list."""
    return capi.copy_params(params, shard=(rank, world_size) if world_size > 1 else (0, 0))


def accum_tensor(params, device):
    """The (H, W, 3) int64 fixed-point sensor accumulators the engine splats into."""
    return torch.zeros((params.height, params.width, 3), dtype=torch.int64, device=device)


def reduce_accum(accum, dst=None, group=None):
    """Sum the ranks' accumulators: all-reduce (dst=None) or reduce to `dst`.  No-op for 1 rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return accum
    if dst is None:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


class ShardedFlare:
    """Renders this rank's shard on its GPU and reduces the frame across ranks."""

    def __init__(self, engine, params, rank, world_size, device):
        self.engine, self.rank, self.world_size, self.device = engine, rank, world_size, device
        self.full_params = params
        self.params = shard_params(params, rank, world_size)
        self.accum = accum_tensor(params, device)
        # order the engine's stream after torch's current stream and vice versa with events
        self.ext_stream = torch.cuda.ExternalStream(engine.stream, device=device)

    def render(self, lights, reduce_dst=None):
        """Trace + splat this shard (engine stream), then the reduce (torch stream).  Asynchronous."""
        cur = torch.cuda.current_stream(self.device)
        self.ext_stream.wait_stream(cur)
        self.engine.render_ghosts_device(lights, self.params, self.accum.data_ptr(), clear_first=True)
        cur.wait_stream(self.ext_stream)
        reduce_accum(self.accum, dst=reduce_dst)
        return self.accum

    def finalize(self, out, elem=capi.F32x3):
        """accum -> (H, W, 3) float32/float64 pixels in `out` (a device tensor), on the engine stream."""
        cur = torch.cuda.current_stream(self.device)
        self.ext_stream.wait_stream(cur)
        self.engine.finalize_device(self.accum.data_ptr(), self.full_params, out.data_ptr(), out.stride(1) * out.element_size(), elem)
        cur.wait_stream(self.ext_stream)
        return out
