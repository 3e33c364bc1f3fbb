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
    """A sensor accumulator buffer the engine splats into (lfb_accum_bytes: the H*W*3 int64 fixed-point sums followed by
    the dirty-tile map), as a flat int64 tensor; accum_pixels() views the sums."""
    words = (capi.lib().lfb_accum_bytes(params.width, params.height) + 7) // 8
    return torch.zeros((words,), dtype=torch.int64, device=device)


def accum_pixels(accum, params):
    """The (H, W, 3) int64 sums of an accumulator buffer."""
    n = params.height * params.width * 3
    return accum[:n].view(params.height, params.width, 3)


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

    def __init__(self, engine, params, rank, world_size, device, n_buffers=3, finalize_engine=None, fused_clear=False):
        self.engine, self.rank, self.world_size, self.device = engine, rank, world_size, device
        self.fin_engine = finalize_engine if finalize_engine is not None else engine
        # fused_clear: lfb_finalize_clear_device (one kernel converts and zeroes) instead of finalize + memset.  Measured on
        # cfg2 (B200): 19.0 us vs 12.9 + 8.7 us as kernels, but the pipelined step got SLOWER (0.1167 vs 0.1114 ms) -- the
        # read-modify-write kernel disturbs the co-running trace more than a pure streaming memset -- so it is off by default
        # and meant for serial hosts, where it saves a launch.
        self.fused_clear = fused_clear
        self.full_params = params
        self.params = shard_params(params, rank, world_size)
        self.accums = [accum_tensor(params, device) for _ in range(n_buffers)]
        self.accum = self.accums[0]
        self.A = torch.cuda.ExternalStream(engine.stream, device=device)
        self.B = torch.cuda.ExternalStream(self.fin_engine.stream, device=device)
        self.C = torch.cuda.Stream(device=device) if world_size > 1 else None
        self.traced = [torch.cuda.Event() for _ in range(n_buffers)]
        self.reduced = [torch.cuda.Event() for _ in range(n_buffers)]
        self.fin_events = [torch.cuda.Event() for _ in range(n_buffers)]  # reused: no per-frame allocations
        self.finalized = [None] * n_buffers
        self.dirty = [False] * n_buffers  # buffers whose sums were kept: cleared by the next trace into them
        self.k = 0

    def reset(self):
        """Forget the cross-frame dependencies (call with all streams idle, e.g. before capturing frames into a CUDA graph:
        a capture may not wait on events recorded outside it)."""
        self.finalized = [None] * len(self.accums)
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

    def frame(self, lights, out=None, elem=capi.F32x3, reduce_dst=0, keep=False):
        """Enqueue one frame: trace this shard, reduce (to `reduce_dst`, or all-reduce when None), and -- on the rank(s)
        holding the sum, when `out` (an (H, W, 3) device tensor) is given -- convert to pixels.  Returns the buffer index.
        keep=True leaves the summed accumulators in the buffer (it is then cleared at its next use instead of right away)."""
        b = self.k % len(self.accums)
        self.k += 1
        acc = self.accums[b]
        self.accum = acc
        if self.finalized[b] is not None:
            self.A.wait_event(self.finalized[b])  # the buffer's previous frame has been read out (and the buffer re-zeroed)
        # the accumulators are zero here: freshly allocated, or cleared on stream B right after they were read out, which
        # takes the 49.8 MB memset off the trace stream's critical path
        self.engine.render_ghosts_device(lights, self.params, acc.data_ptr(), clear_first=bool(self.dirty[b]))
        self.dirty[b] = keep
        self.traced[b].record(self.A)
        last = self.traced[b]
        if self.C is not None:
            self.C.wait_event(last)
            prev = torch.cuda.current_stream(self.device)
            torch.cuda.set_stream(self.C)  # cheaper than the context manager: this runs once per frame
            px = accum_pixels(acc, self.full_params)  # the sums only: the dirty-tile maps behind them are per rank
            if reduce_dst is None:
                dist.all_reduce(px, op=dist.ReduceOp.SUM)
            else:
                dist.reduce(px, dst=reduce_dst, op=dist.ReduceOp.SUM)
            self.reduced[b].record(self.C)
            torch.cuda.set_stream(prev)
            last = self.reduced[b]
        self.B.wait_event(last)
        cleared = False
        if out is not None and (reduce_dst is None or self.world_size == 1 or self.rank == reduce_dst):
            self.fin_engine.finalize_device(acc.data_ptr(), self.full_params, out.data_ptr(), out.stride(1) * out.element_size(), elem,
                                            clear=self.fused_clear and not keep)
            cleared = self.fused_clear and not keep
        if not keep and not cleared:
            prev = torch.cuda.current_stream(self.device)
            torch.cuda.set_stream(self.B)
            acc.zero_()
            torch.cuda.set_stream(prev)
        ev = self.fin_events[b]
        ev.record(self.B)
        self.finalized[b] = ev
        return b

    # --- the serial one-frame form (kept for callers that want a frame at a time) -----------------------
    def render(self, lights, reduce_dst=None):
        self.begin()
        b = self.frame(lights, out=None, reduce_dst=reduce_dst, keep=True)
        self.join()
        return accum_pixels(self.accums[b], self.full_params)

    def finalize(self, out, elem=capi.F32x3):
        cur = torch.cuda.current_stream(self.device)
        self.B.wait_stream(cur)
        self.fin_engine.finalize_device(self.accum.data_ptr(), self.full_params, out.data_ptr(), out.stride(1) * out.element_size(), elem)
        cur.wait_stream(self.B)
        return out


class PeerFlare:
    """The B200-native form of the multi-GPU frame: NO collective call.  Every rank's accumulators live in symmetric
    memory (mapped into all ranks over NVLink); after the trace, each rank runs ONE fused kernel
    (lfb_reduce_finalize_peers) that sums all ranks' accumulators for its 1/N slice of the pixels -- peer loads, or
    in-switch multimem.ld_reduce when the buffers are bound to an NVSwitch multicast object -- converts to pixels and
    stores them directly into the owner rank's output buffer.

    Two streams, R >= 3 rotating accumulator / output sets:
        stream A (engine)           trace frame k into accum[k % R]; device-side barrier (symmetric-memory signal pads)
        stream B (finalize engine)  fused reduce + finalize of frame k (reads every rank's accum[k % R])
    so the reduce of frame k overlaps the trace of frame k+1.  Buffer safety: trace(k) first waits for this rank's own
    reduce(k-R+1); a peer can only pass barrier(k+R-1) -- and then overwrite accum[k % R] in trace(k+R) -- after this
    rank arrived there, i.e. after this rank's reduce(k) finished reading it.  The owner's pixels of frame k are complete
    after finish() (or once R-1 further frames have passed their barrier).

    torch supplies the plumbing (symmetric allocation, rendezvous, barrier); the data path is liblfb200.so."""

    def __init__(self, engine, params, rank, world_size, device, group, n_buffers=3, out_dtype=torch.float32,
                 use_multicast=False, finalize_engine=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.engine, self.rank, self.world, self.device = engine, rank, world_size, device
        self.fin_engine = finalize_engine if finalize_engine is not None else engine
        self.full_params = params
        self.params = shard_params(params, rank, world_size)
        self.n_buffers = max(int(n_buffers), 3 if finalize_engine is not None else 2)
        H, W = params.height, params.width
        acc_words = (capi.lib().lfb_accum_bytes(W, H) + 7) // 8
        self.accum_all = symm_mem.empty((self.n_buffers, acc_words), dtype=torch.int64, device=device)
        self.out_all = symm_mem.empty((self.n_buffers, H, W, 3), dtype=out_dtype, device=device)
        self.h_acc = symm_mem.rendezvous(self.accum_all, group.group_name)
        self.h_out = symm_mem.rendezvous(self.out_all, group.group_name)
        self.flags = symm_mem.empty((16,), dtype=torch.int64, device=device)
        self.flags.zero_()
        self.h_flags = symm_mem.rendezvous(self.flags, group.group_name)
        self.epoch = 0
        self.accum_all.zero_()
        self.acc_bytes = acc_words * 8
        self.out_bytes = H * W * 3 * self.out_all.element_size()
        self.elem = capi.F32x3 if out_dtype == torch.float32 else capi.F64x3
        self.mc = int(self.h_acc.multicast_ptr) if (use_multicast and self.h_acc.has_multicast_support) else 0
        self.A = torch.cuda.ExternalStream(engine.stream, device=device)
        self.B = torch.cuda.ExternalStream(self.fin_engine.stream, device=device)
        self.two_streams = self.fin_engine is not engine
        self.reduce_done = {}
        self.k = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def begin(self, stream=None):
        cur = stream or torch.cuda.current_stream(self.device)
        self.A.wait_stream(cur)
        self.B.wait_stream(cur)

    def barrier(self):
        """Device-side barrier across the ranks, enqueued on the engine stream (lfb_peer_barrier: one warp, system-scope
        flags in symmetric memory; nothing blocks on the host)."""
        self.epoch += 1
        self.engine.peer_barrier([int(p) for p in self.h_flags.buffer_ptrs], self.rank, self.epoch)

    def finish(self, stream=None):
        """All frames enqueued so far are complete in the owner's buffers once `stream` passes this point."""
        self.A.wait_stream(self.B)
        self.barrier()
        (stream or torch.cuda.current_stream(self.device)).wait_stream(self.A)

    def frame(self, lights, owner=0):
        """Enqueue one frame; returns its buffer index."""
        k, R = self.k, self.n_buffers
        b = k % R
        self.k += 1
        if self.two_streams and (k - R + 1) in self.reduce_done:
            self.A.wait_event(self.reduce_done.pop(k - R + 1))
        my_acc = int(self.h_acc.buffer_ptrs[self.rank]) + b * self.acc_bytes
        self.engine.render_ghosts_device(lights, self.params, my_acc, clear_first=True)
        self.barrier()  # every rank has finished splatting into buffer b
        if self.two_streams:
            self.B.wait_stream(self.A)
        ptrs = [int(p) + b * self.acc_bytes for p in self.h_acc.buffer_ptrs]
        out_ptr = int(self.h_out.buffer_ptrs[owner]) + b * self.out_bytes
        self.fin_engine.reduce_finalize_peers(ptrs, self.rank, self.full_params, out_ptr, 3 * self.out_all.element_size(), self.elem,
                                              multicast_ptr=(self.mc + b * self.acc_bytes) if self.mc else None)
        if self.two_streams:
            ev = torch.cuda.Event()
            ev.record(self.B)
            self.reduce_done[k] = ev
        return b

    def result(self, b):
        """The owner's pixels of buffer b (valid after finish())."""
        return self.out_all[b]


class PeerSparse:
    """The multi-GPU frame, tile-sparse: NO collective call and NO full-frame traffic.  Every rank's accumulator buffers
    (sums + dirty-tile map) live in symmetric memory; after its trace each rank runs ONE kernel (lfb_reduce_tiles_peers) for
    its interleaved share of the dirty tiles: it ORs the ranks' tile maps, sums the tiles that are dirty anywhere -- reading
    only the ranks that have them, over NVLink peer memory -- zeroes them, and stores the pixels into the owner's output
    frame (a peer store) -- or, when `host_out` is given, straight into a page-locked HOST buffer every rank has mapped
    (a POSIX shared-memory segment each process registered): N GPUs then write the frame's dirty tiles over their own PCIe
    links at the same time.  ~1 % of the bytes an all-pixels reduce moves; integer sums, so the frame has the same bits for
    any rank count.

    Two streams, R >= 3 rotating buffer sets (accumulators, output frame, tile state):
        stream A (engine)           trace frame k into accum[k % R] -- nothing else: traces run back to back
        stream B (finalize engine)  device-side barrier(k) (symmetric-memory flags: every rank has traced frame k), then the tile
                                    reduce of frame k
    The barrier is on B, off the trace stream's critical path (on A it cost a cross-GPU round trip plus the slowest rank's
    jitter per frame: 0.123 ms per step at N = 8 against 0.090 on one GPU).  Buffer safety: every rank's reduce(k) reads -- and
    zeroes -- every rank's accum[k % R]; rank p may splat into its accum[k % R] again (trace(k+R)) once ALL ranks' reduce(k) are
    done, which rank p knows when it has passed barrier(k+1) on its own stream B (each rank arrives there after its reduce(k)).
    So trace(k+R) waits for the event recorded after barrier(k+1): two frames old for R = 3.  The owner's pixels of frame k are
    complete after finish()."""

    def __init__(self, engine, params, rank, world_size, device, group, n_buffers=3, out_dtype=torch.float32, finalize_engine=None,
                 host_out_ptrs=None, drain_engine=None, host_stride=24):
        """drain_engine (with host_out_ptrs; every rank alike): a third engine whose stream C carries the host writes.  The reduce
        then leaves this rank's tiles as pixels in a device staging buffer (lfb_reduce_tiles_peers_staged) and
        lfb_drain_tiles copies them into the host frame with a few clock-paced CTAs -- unpaced stores into host memory keep
        the GPU from fetching its next commands while they drain.  A frame is then complete on every rank once
        this rank has passed the barrier on C that follows the drain (wait_frame)."""
        import torch.distributed._symmetric_memory as symm_mem
        self.engine, self.rank, self.world, self.device = engine, rank, world_size, device
        self.fin_engine = finalize_engine if finalize_engine is not None else engine
        self.full_params = params
        self.params = shard_params(params, rank, world_size)
        self.n_buffers = max(int(n_buffers), 3 if finalize_engine is not None else 2)
        H, W = params.height, params.width
        acc_words = (capi.lib().lfb_accum_bytes(W, H) + 7) // 8
        self.accum_all = symm_mem.empty((self.n_buffers, acc_words), dtype=torch.int64, device=device)
        self.accum_all.zero_()
        self.out_all = symm_mem.empty((self.n_buffers, H, W, 3), dtype=out_dtype, device=device)
        self.out_all.zero_()
        self.h_acc = symm_mem.rendezvous(self.accum_all, group.group_name)
        self.h_out = symm_mem.rendezvous(self.out_all, group.group_name)
        self.flags = symm_mem.empty((16,), dtype=torch.int64, device=device)
        self.flags.zero_()
        self.h_flags = symm_mem.rendezvous(self.flags, group.group_name)
        st_words = (capi.lib().lfb_tile_state_bytes(W, H) + 3) // 4
        self.state = torch.zeros((self.n_buffers, st_words), dtype=torch.int32, device=device)  # this rank's own, per output frame
        self.epoch = 0
        self.acc_bytes = acc_words * 8
        self.out_bytes = H * W * 3 * self.out_all.element_size()
        self.elem = capi.F32x3 if out_dtype == torch.float32 else capi.F64x3
        self.host_out_ptrs = host_out_ptrs  # optional: per buffer, THIS rank's device mapping of a shared page-locked host frame
        self.A = torch.cuda.ExternalStream(engine.stream, device=device)
        self.B = torch.cuda.ExternalStream(self.fin_engine.stream, device=device)
        self.two_streams = self.fin_engine is not engine
        self.reduce_events = [torch.cuda.Event() for _ in range(self.n_buffers)]
        self.reduce_valid = [False] * self.n_buffers
        self.traced = [torch.cuda.Event() for _ in range(self.n_buffers)]
        self.passed = [torch.cuda.Event() for _ in range(self.n_buffers + 1)]  # [j % (R+1)]: this rank passed barrier(j) on B
        self.drain_engine = drain_engine if (drain_engine is not None and host_out_ptrs is not None and self.two_streams) else None
        if self.drain_engine is not None:
            self.host_stride = host_stride
            stage_bytes = capi.lib().lfb_tile_stage_bytes(W, H, host_stride)
            self.stage = torch.zeros((self.n_buffers, (stage_bytes + 15) // 16 * 16), dtype=torch.uint8, device=device)
            self.C = torch.cuda.ExternalStream(self.drain_engine.stream, device=device)
            self.flags2 = symm_mem.empty((16,), dtype=torch.int64, device=device)  # the barriers on C have their own flags and epochs
            self.flags2.zero_()
            self.h_flags2 = symm_mem.rendezvous(self.flags2, group.group_name)
            self.epoch2 = 0
            self.drained = [torch.cuda.Event() for _ in range(self.n_buffers)]
            self.drained_valid = [False] * self.n_buffers
            self.done = [torch.cuda.Event() for _ in range(self.n_buffers + 1)]  # [j % (R+1)]: frame j is in host memory on every rank
        self.k = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def begin(self, stream=None):
        cur = stream or torch.cuda.current_stream(self.device)
        self.A.wait_stream(cur)
        self.B.wait_stream(cur)
        if self.drain_engine is not None:
            self.C.wait_stream(cur)

    def barrier(self, engine=None):
        """Enqueue the device-side barrier on `engine`'s stream (default: the trace engine's)."""
        self.epoch += 1
        (engine or self.engine).peer_barrier([int(p) for p in self.h_flags.buffer_ptrs], self.rank, self.epoch)

    def finish(self, stream=None):
        """All frames enqueued so far are complete in the owner's buffers once `stream` passes this point."""
        self.A.wait_stream(self.B)
        if self.drain_engine is not None:
            self.A.wait_stream(self.C)
        self.barrier()  # on A, after everything on B (and C): every rank's reduces (and drains) are done
        self.B.wait_stream(self.A)  # later barriers on B come after this one
        if self.drain_engine is not None:
            self.C.wait_stream(self.A)
        (stream or torch.cuda.current_stream(self.device)).wait_stream(self.A)

    def frame(self, lights, owner=0, elem=None, stride=None):
        """Enqueue one frame; returns its buffer index."""
        k, R = self.k, self.n_buffers
        b = k % R
        self.k += 1
        my_acc = int(self.h_acc.buffer_ptrs[self.rank]) + b * self.acc_bytes
        if self.two_streams:
            if k - R >= 0:  # all ranks' reduce(k - R) are done once this rank has passed barrier(k - R + 1)
                self.A.wait_event(self.passed[(k - R + 1) % (R + 1)])
            self.engine.render_ghosts_device(lights, self.params, my_acc, clear_first=False)  # the reducers left it clear
            self.traced[b].record(self.A)
            self.B.wait_event(self.traced[b])
            self.barrier(self.fin_engine)  # on B: every rank has finished splatting into its buffer b
            self.passed[k % (R + 1)].record(self.B)
        else:
            self.engine.render_ghosts_device(lights, self.params, my_acc, clear_first=False)
            self.barrier()
        ptrs = [int(p) + b * self.acc_bytes for p in self.h_acc.buffer_ptrs]
        if self.drain_engine is not None:
            out_ptr = self.host_out_ptrs[b]
            el = elem if elem is not None else capi.F64x3
            if self.drained_valid[b]:
                self.B.wait_event(self.drained[b])  # the buffer's previous frame has left state[b] / stage[b]
            self.fin_engine.reduce_tiles_peers_staged(ptrs, self.rank, self.full_params, out_ptr, self.host_stride, el, self.state[b].data_ptr(),
                                                      self.stage[b].data_ptr())
            self.reduce_events[b].record(self.B)
            self.reduce_valid[b] = True
            self.C.wait_event(self.reduce_events[b])
            self.drain_engine.drain_tiles(self.full_params, out_ptr, self.host_stride, self.state[b].data_ptr(), self.stage[b].data_ptr())
            self.drained[b].record(self.C)
            self.drained_valid[b] = True
            self.epoch2 += 1  # every rank arrives here after its drain of frame k
            self.drain_engine.peer_barrier([int(p) for p in self.h_flags2.buffer_ptrs], self.rank, self.epoch2)
            self.done[k % (R + 1)].record(self.C)
            return b
        if self.host_out_ptrs is not None:
            out_ptr = self.host_out_ptrs[b]
            self.fin_engine.reduce_tiles_peers(ptrs, self.rank, self.full_params, out_ptr, stride or 24, elem if elem is not None else capi.F64x3,
                                               self.state[b].data_ptr())
        else:
            out_ptr = int(self.h_out.buffer_ptrs[owner]) + b * self.out_bytes
            self.fin_engine.reduce_tiles_peers(ptrs, self.rank, self.full_params, out_ptr, 3 * self.out_all.element_size(), self.elem,
                                               self.state[b].data_ptr())
        if self.two_streams:
            self.reduce_events[b].record(self.B)
            self.reduce_valid[b] = True
        return b

    def wait_frame(self, j):
        """Block the host until frame j (0-based, in enqueue order) is complete in its output buffer ON EVERY RANK: this rank has
        passed barrier(j + 1), where each rank arrives only after its reduce(j).  Needs frame j + 1 enqueued and at most R frames
        enqueued after j (two streams only)."""
        if self.drain_engine is not None:  # the barrier on C after the drain of frame j
            if not (j < self.k <= j + 1 + self.n_buffers):
                raise ValueError("wait_frame(%d): the frame must be enqueued and recent (k = %d)" % (j, self.k))
            self.done[j % (self.n_buffers + 1)].synchronize()
            return
        if not self.two_streams or not (j + 1 < self.k <= j + 1 + self.n_buffers):
            raise ValueError("wait_frame(%d): frame %d must be enqueued and recent (k = %d)" % (j, j + 1, self.k))
        self.passed[(j + 1) % (self.n_buffers + 1)].synchronize()

    def result(self, b):
        """The owner's pixels of buffer b (valid after finish())."""
        return self.out_all[b]
