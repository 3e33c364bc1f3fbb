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
    """Renders this rank's shard on its GPU and reduces the frame across ranks, as a 3-stage pipeline over
    rotating accumulator buffers so that consecutive frames overlap:

        stream A (engine)          clear + trace + splat of frame k      -> accum[k % R]
        stream C (comm, N > 1)     NCCL sum-reduce of accum[k % R]
        stream B (finalize engine) fixed point -> pixels of frame k      -> the caller's device tensor

    Stage s of frame k only waits for stage s-1 of frame k and for the previous user of its buffer, so the
    steady-state frame time is max(trace, reduce, finalize) rather than their sum.  Everything is asynchronous;
    join() makes a torch stream wait for all three."""

    def __init__(self, engine, params, rank, world_size, device, n_buffers=3, finalize_engine=None):
        self.engine, self.rank, self.world_size, self.device = engine, rank, world_size, device
        self.fin_engine = finalize_engine if finalize_engine is not None else engine
        self.full_params = params
        self.params = shard_params(params, rank, world_size)
        self.accums = [accum_tensor(params, device) for _ in range(n_buffers)]
        self.accum = self.accums[0]
        self.A = torch.cuda.ExternalStream(engine.stream, device=device)
        self.B = torch.cuda.ExternalStream(self.fin_engine.stream, device=device)
        self.C = torch.cuda.Stream(device=device) if world_size > 1 else None
        self.traced = [torch.cuda.Event() for _ in range(n_buffers)]
        self.reduced = [torch.cuda.Event() for _ in range(n_buffers)]
        self.finalized = [None] * n_buffers
        self.k = 0

    def begin(self, stream=None):
        """Order the pipeline after `stream` (default: torch's current stream)."""
        cur = stream or torch.cuda.current_stream(self.device)
        self.A.wait_stream(cur)
        self.B.wait_stream(cur)
        if self.C is not None:
            self.C.wait_stream(cur)

    def join(self, stream=None):
        cur = stream or torch.cuda.current_stream(self.device)
        cur.wait_stream(self.A)
        cur.wait_stream(self.B)
        if self.C is not None:
            cur.wait_stream(self.C)

    def frame(self, lights, out=None, elem=capi.F32x3, reduce_dst=0):
        """Enqueue one frame: trace this shard, reduce (to `reduce_dst`, or all-reduce when None), and -- on the rank(s)
        holding the sum, when `out` (an (H, W, 3) device tensor) is given -- convert to pixels.  Returns the buffer index."""
        b = self.k % len(self.accums)
        self.k += 1
        acc = self.accums[b]
        self.accum = acc
        if self.finalized[b] is not None:
            self.A.wait_event(self.finalized[b])  # the buffer's previous frame has been read out
        self.engine.render_ghosts_device(lights, self.params, acc.data_ptr(), clear_first=True)
        self.traced[b].record(self.A)
        last = self.traced[b]
        if self.C is not None:
            self.C.wait_event(last)
            with torch.cuda.stream(self.C):
                reduce_accum(acc, dst=reduce_dst)
                self.reduced[b].record(self.C)
            last = self.reduced[b]
        self.B.wait_event(last)
        if out is not None and (reduce_dst is None or self.world_size == 1 or self.rank == reduce_dst):
            self.fin_engine.finalize_device(acc.data_ptr(), self.full_params, out.data_ptr(), out.stride(1) * out.element_size(), elem)
        ev = torch.cuda.Event()
        ev.record(self.B)
        self.finalized[b] = ev
        return b

    # --- the serial one-frame form (kept for callers that want a frame at a time) -----------------------
    def render(self, lights, reduce_dst=None):
        self.begin()
        b = self.frame(lights, out=None, reduce_dst=reduce_dst)
        self.join()
        return self.accums[b]

    def finalize(self, out, elem=capi.F32x3):
        cur = torch.cuda.current_stream(self.device)
        self.B.wait_stream(cur)
        self.fin_engine.finalize_device(self.accum.data_ptr(), self.full_params, out.data_ptr(), out.stride(1) * out.element_size(), elem)
        cur.wait_stream(self.B)
        return out
