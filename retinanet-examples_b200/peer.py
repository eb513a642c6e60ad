"""Image-wise sharded inference without a collective launch: every rank's NMS kernel stores its detections straight into
every other rank's gather buffer over NVLink peer mappings (odtk_nms_gather / odtk_gather_wait, include/odtk_b200.h).
Replaces the all_gathers of the reference's odtk/infer.py:98-102 (five NCCL collectives after the loop) and round 1's
single host-launched ncclAllGather per step; the exchange is part of the step's kernels, so it sits inside the CUDA graph.

One process per GPU (torch.distributed initialised, any backend -- it only carries the 64-byte IPC handles once).
The buffers are plain cudaMalloc memory shared through CUDA IPC by the library itself (odtk_peer_alloc / odtk_peer_open):
each rank opens its peers' handles with ITS OWN device current, which maps them into its address space with NVLink peer
access enabled."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class _DevArray:
    """Zero-copy view of raw device memory for torch (the CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


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
        L = _lib.lib()
        # one allocation per rank: [ two parity halves of the gather buffer | arrival counters, one per source rank ]
        self._buf_bytes = (2 * half * 4 + 255) // 256 * 256
        nbytes = self._buf_bytes + 256
        self._base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _lib.check(L.odtk_peer_alloc(nbytes, ctypes.byref(self._base), handle), "peer_alloc")
        self.buf = torch.as_tensor(_DevArray(self._base.value, 2 * half, "<f4"), device=self.device)
        self.flags = torch.as_tensor(_DevArray(self._base.value + self._buf_bytes, self.world, "<i4"), device=self.device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, bytes(handle), group=group)
        self._opened = []
        self.desc = _lib.Gather()
        for r in range(self.world):
            if r == self.rank:
                base = self._base.value
            else:
                p = ctypes.c_void_p()
                h = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[r])
                with torch.cuda.device(self.device):
                    _lib.check(L.odtk_peer_open(h, ctypes.byref(p)), "peer_open (rank %d)" % r)
                self._opened.append(p.value)
                base = p.value
            self.desc.packed[r], self.desc.flags[r] = base, base + self._buf_bytes
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

    def close(self):
        """Unmap the peers' buffers and free this rank's (after every rank is done with them: barrier first)."""
        L = _lib.lib()
        for p in self._opened:
            L.odtk_peer_close(ctypes.c_void_p(p))
        self._opened = []
        if self._base is not None and self._base.value:
            L.odtk_peer_free(self._base)
            self._base = None
