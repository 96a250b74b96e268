"""Voice-sharded rendering across the GPUs of one box: one process per GPU (torchrun), each rank renders the
sub-graph of its own voices into a device-resident partial mix, then ONE reduce (NCCL over NVLink) sums the
mixed output blocks onto rank 0.  The feed-forward effect tree is linear in its top-level Sum2 mix, so the sum of
per-rank mixes equals the single-GPU render up to f32 summation order.

PyTorch is plumbing here: device tensors for the output block, torch.distributed for the reduce."""
import numpy as np
import torch

from . import B200Renderer


class ShardedRenderer:
    def __init__(self, rank=0, world_size=1, device=0, **kw):
        self.rank, self.world_size, self.device = rank, world_size, device
        self.r = B200Renderer(device=device, **kw)
        self._out = None
        self._host = None

    def voices_of_rank(self, n_voices):
        """voices v with v mod world_size == rank (round-robin shard)."""
        return list(range(self.rank, n_voices, self.world_size))

    def _device_out(self, n_slots, n_times):
        if self._out is None or tuple(self._out.shape) != (n_slots, n_times):
            self._out = torch.empty((n_slots, n_times), dtype=torch.float32, device=f"cuda:{self.device}")
        return self._out

    def fill_buffer_device(self, n_slots, n_times, idx):
        """Renders this rank's shard and reduces onto rank 0.  Returns the device tensor (valid on rank 0)."""
        out = self._device_out(n_slots, n_times)
        self.r.fill_buffer_device(out.data_ptr(), n_slots, n_times, idx)
        self.r.sync()                                   # renderer stream -> visible to the collective's stream
        if self.world_size > 1:
            import torch.distributed as dist
            dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
        return out

    def fill_buffer(self, n_slots, n_times, idx):
        """End-to-end: host ndarray on rank 0 (None elsewhere); includes the device->host copy."""
        out = self.fill_buffer_device(n_slots, n_times, idx)
        if self.rank != 0:
            torch.cuda.synchronize(self.device)
            return None
        if self._host is None or tuple(self._host.shape) != (n_slots, n_times):
            self._host = torch.empty((n_slots, n_times), dtype=torch.float32, pin_memory=True)
        self._host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(self.device)
        return self._host.numpy()
