"""Voice-sharded rendering across the GPUs of one box: one process per GPU (torchrun), each rank renders the
sub-graph of its own voices into a device-resident partial mix, then ONE reduce (NCCL over NVLink) sums the
mixed output blocks onto rank 0.  The feed-forward effect tree is linear in its top-level Sum2 mix, so the sum of
per-rank mixes equals the single-GPU render up to f32 summation order.

PyTorch is plumbing here: device tensors for the output block, torch.distributed for the reduce."""
import torch

from . import B200Renderer


def voices_of_rank(n_voices, rank, world_size):
    """voices v with v mod world_size == rank (round-robin shard)."""
    return list(range(rank, n_voices, world_size))


class ShardedRenderer:
    """`renderer` defaults to a B200Renderer on `device`.  Tests inject another object with the renderer interface
    (host path: fill_buffer) to exercise the shard/reduce logic over gloo without a GPU."""

    def __init__(self, rank=0, world_size=1, device=0, renderer=None, exchange="nccl", **kw):
        """exchange: "nccl" — torch.distributed reduce of the per-rank mix blocks (default);
        "p2p" — K5: every rank's stage kernel stores its mix block straight into rank 0's slab over NVLink
        (CUDA-IPC mapping), one barrier, rank 0 sums the rows in rank order (deterministic)."""
        self.rank, self.world_size, self.device = rank, world_size, device
        self.on_gpu = renderer is None
        self.r = B200Renderer(device=device, **kw) if renderer is None else renderer
        self.exchange = exchange if (self.on_gpu and world_size > 1) else "nccl"
        self._out = None
        self._host = None
        self._stream = None
        self._ev = None
        self._slabs = None          # p2p: (shape, [slab tensors on rank 0], [row pointers of this rank], step counter)

    def voices_of_rank(self, n_voices):
        return voices_of_rank(n_voices, self.rank, self.world_size)

    def _block(self, n_slots, n_times):
        dev = f"cuda:{self.device}" if self.on_gpu else "cpu"
        if self._out is None or tuple(self._out.shape) != (n_slots, n_times):
            self._out = torch.empty((n_slots, n_times), dtype=torch.float32, device=dev)
        return self._out

    def _reduce(self, out):
        if self.world_size > 1:
            import torch.distributed as dist
            if self.on_gpu:
                # Stream order on the device, no host round trip on either side (round 1 had two per step: r.sync() before
                # the collective, current_stream().synchronize() after): torch's current stream — the one the collective
                # is ordered on — waits for the render through an event, and the renderer's stream waits for the
                # collective the same way before the next render overwrites this block.
                ext, cur = self.cuda_stream(), torch.cuda.current_stream(self.device)
                if self._ev is None:
                    self._ev = (torch.cuda.Event(), torch.cuda.Event())
                self._ev[0].record(ext)
                cur.wait_event(self._ev[0])
                dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)   # the path's single exchange step
                self._ev[1].record(cur)
                ext.wait_event(self._ev[1])
            else:
                dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
        return out

    def cuda_stream(self):
        """The CUDA stream the renderer launches on, as a torch stream (CUDA-event timing; the collective is ordered on it)."""
        if self._stream is None:
            self._stream = torch.cuda.ExternalStream(self.r.stream(), device=f"cuda:{self.device}")
        return self._stream

    def _p2p_setup(self, n_slots, n_times):
        import torch.distributed as dist
        shape = (n_slots, n_times)
        if self._slabs is not None and self._slabs[0] == shape:
            return
        n = n_slots * n_times
        slabs, handles = [], [None, None]
        if self.rank == 0:
            for k in range(2):                      # double-buffered: one barrier per step is enough
                # its own cudaMalloc: an IPC handle names a whole allocation, not a sub-block of torch's pool
                slab = self.r.device_alloc(4 * n * self.world_size)
                slabs.append(slab)
                handles[k] = self.r.ipc_export(slab)
        dist.broadcast_object_list(handles, src=0)
        if self.rank == 0:
            rows = list(slabs)
        else:
            rows = [self.r.ipc_open(h) + 4 * n * self.rank for h in handles]
        self._slabs = [shape, slabs, rows, 0]

    def _fill_p2p(self, n_slots, n_times, idx):
        import torch.distributed as dist
        self._p2p_setup(n_slots, n_times)
        shape, slabs, rows, step = self._slabs
        k = step % 2
        self._slabs[3] = step + 1
        # the stage kernel's output stores ARE the transfer: row `rank` of rank 0's slab, over NVLink
        self.r.fill_buffer_device(rows[k], n_slots, n_times, idx)
        self.r.sync()
        dist.barrier()
        out = self._block(n_slots, n_times)
        if self.rank == 0:
            n = n_slots * n_times
            self.r.sum_rows(out.data_ptr(), slabs[k], self.world_size, n, n)
            self.r.sync()
        return out

    def fill_buffer_device(self, n_slots, n_times, idx, inputs=None):
        """Renders this rank's shard and reduces onto rank 0.  Returns the block tensor (valid on rank 0) with the work
        ENQUEUED on the renderer's stream (`cuda_stream()`): order later device work on that stream, or `r.sync()`."""
        if self.exchange == "p2p":
            assert not inputs
            return self._fill_p2p(n_slots, n_times, idx)
        out = self._block(n_slots, n_times)
        if self.on_gpu:
            assert not inputs, "device path: feed external inputs through B200Renderer.fill_buffer_device directly"
            self.r.fill_buffer_device(out.data_ptr(), n_slots, n_times, idx)      # enqueued on the renderer's stream
        else:
            out.copy_(torch.from_numpy(self.r.fill_buffer(n_slots, n_times, idx, inputs)))
        return self._reduce(out)

    def fill_buffer(self, n_slots, n_times, idx, inputs=None):
        """End-to-end: host ndarray on rank 0 (None elsewhere); includes the device->host copy.
        The returned array is a view of a pinned staging block that the next call overwrites."""
        out = self.fill_buffer_device(n_slots, n_times, idx, inputs)
        if not self.on_gpu:
            return out.numpy().copy() if self.rank == 0 else None
        if self.rank != 0:
            torch.cuda.synchronize(self.device)
            return None
        if self._host is None or tuple(self._host.shape) != (n_slots, n_times):
            self._host = torch.empty((n_slots, n_times), dtype=torch.float32, pin_memory=True)
        if self.world_size == 1:
            self.r.sync()                              # no collective ordered the copy's stream behind the render
        self._host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(self.device)
        return self._host.numpy()

    def render_stream(self, n_slots, idx, n_total, block, sink):
        """N4 at N GPUs: [idx, idx + n_total) in blocks of `block` samples; every block is rendered by all ranks, reduced
        onto rank 0 and handed to `sink(block_array, idx)` there (a view of pinned staging memory, valid during the
        call).  Two blocks in flight: while block k renders and reduces, block k-1 travels device->host on a copy
        stream and runs through the sink.  Same bits as the fill_buffer calls it stands for."""
        assert block > 0
        n_blocks = max(1, -(-n_total // block))
        if not self.on_gpu:
            t = idx
            for _ in range(n_blocks):
                n = min(block, idx + n_total - t)
                out = self.fill_buffer(n_slots, n, t)
                if self.rank == 0:
                    sink(out, t)
                t += n
            return
        dev = f"cuda:{self.device}"
        cap = n_slots * min(block, max(n_total, 1))
        copy_stream = torch.cuda.Stream(device=dev)
        dflat = [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(2)]
        hflat = [torch.empty(cap, dtype=torch.float32, pin_memory=True) for _ in range(2)] if self.rank == 0 else None
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        pending = [None, None]

        def deliver(b):                                  # wait for block b's copy, hand it to the sink
            if pending[b] is None:
                return
            t, n = pending[b]
            pending[b] = None
            copied[b].synchronize()
            if self.rank == 0:
                sink(hflat[b][:n_slots * n].view(n_slots, n).numpy(), t)

        t = idx
        for k in range(n_blocks):
            b = k & 1
            n = min(block, idx + n_total - t)
            # dflat[b] was last read by the copy of block k-2, delivered (hence complete) during iteration k-1
            self._out = dflat[b][:n_slots * n].view(n_slots, n)
            out = self.fill_buffer_device(n_slots, n, t)        # render + exchange, enqueued on the renderer's stream
            copy_stream.wait_stream(self.cuda_stream())
            copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            if self.rank == 0:
                with torch.cuda.stream(copy_stream):
                    hflat[b][:n_slots * n].view(n_slots, n).copy_(out, non_blocking=True)
            copied[b].record(copy_stream)
            pending[b] = (t, n)
            deliver(b ^ 1)                                       # block k-1 runs through the sink while k+1 renders
            t += n
        deliver((n_blocks - 1) & 1)
        self._out = None
