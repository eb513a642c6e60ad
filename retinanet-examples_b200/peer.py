"""Image-wise sharded inference without a collective launch: every rank's NMS kernel stores its detections straight into
every other rank's gather buffer over NVLink peer mappings (odtk_nms_gather / odtk_gather_wait, include/odtk_b200.h).
Replaces the all_gathers of the reference's odtk/infer.py:98-102 (five NCCL collectives after the loop) and round 1's
single host-launched ncclAllGather per step; the exchange is part of the step's kernels, so it sits inside the CUDA graph.

One process per GPU (torch.distributed initialised, any backend -- it only carries the 64-byte IPC handles once).
Buffers are ordinary torch CUDA tensors shared through CUDA IPC (torch's own cudaIpcGetMemHandle / OpenMemHandle path)."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def _export(t):
    """(ipc description of the tensor's storage, byte offset of the tensor inside it)."""
    st = t.untyped_storage()
    return st._share_cuda_(), t.storage_offset() * t.element_size()


def _import(desc, nbytes, device):
    handle, offset = desc
    st = torch.UntypedStorage._new_shared_cuda(*handle)
    return torch.empty(0, dtype=torch.uint8, device=device).set_(st, offset, (nbytes,))


class PeerGather:
    """Gather buffers of one rank + the peer mappings of all others.  `batch` = images per rank per step."""

    def __init__(self, batch, detections=100, nbox=4, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerGather needs torch.distributed (one process per GPU)")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("at most 8 peers (one NVSwitch domain)")
        self.batch, self.det, self.nbox = int(batch), int(detections), int(nbox)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.row = self.det * (2 + self.nbox)
        half = self.world * self.batch * self.row
        self.buf = torch.zeros(2 * half, dtype=torch.float32, device=self.device)          # two parity halves
        self.flags = torch.zeros(self.world, dtype=torch.int32, device=self.device)        # arrival counters, by source rank
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        mine = (_export(self.buf), _export(self.flags))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._peers = []
        self.desc = _lib.Gather()
        for r in range(self.world):
            if r == self.rank:
                pb, pf = self.buf, self.flags
            else:
                pb = _import(everyone[r][0], self.buf.numel() * 4, self.device)
                pf = _import(everyone[r][1], self.world * 4, self.device)
            self._peers.append((pb, pf))
            self.desc.packed[r], self.desc.flags[r] = pb.data_ptr(), pf.data_ptr()
        self.desc.epoch = self.epoch.data_ptr()
        self.desc.num_peers, self.desc.rank = self.world, self.rank
        self.steps = 0                                   # host mirror of the device step counter
        dist.barrier(group=group)                        # nobody stores into a buffer that is not mapped everywhere yet

    def gathered(self, step=None):
        """[world * batch, D, 2 + nbox] view of the rows that landed in step `step` (default: the last completed one).
        Valid until this rank completes two more steps."""
        k = (self.steps - 1) if step is None else step
        half = self.world * self.batch * self.row
        return self.buf[(k & 1) * half:(k & 1) * half + half].view(self.world * self.batch, self.det, 2 + self.nbox)
