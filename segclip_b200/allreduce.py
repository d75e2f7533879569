"""Gradient mean over NVSwitch multicast (NVLS) -- host side of csrc/allreduce.cu.

The flat gradient buffer of the engine is allocated as SYMMETRIC memory (same size on every rank, peer-mapped and bound
to an NVSwitch multicast object).  The mapping itself is plumbing and is done by torch.distributed._symmetric_memory
(cuMemCreate / cuMulticastCreate / handle exchange over the process group); the arithmetic -- one two-shot
multimem.ld_reduce / multimem.st kernel per gradient bucket -- is the library's own (sc_nvls_allreduce).

`NvlsGradSync.create()` returns None when the GPUs have no multicast support (no NVSwitch) or the mapping fails; the
engine then keeps the NCCL bucket all-reduce (the comparator).  SEGCLIP_GRAD_SYNC=nccl forces the comparator."""
import os

import torch

from . import _lib as L


class NvlsGradSync:
    BLOCKS = int(os.environ.get("SEGCLIP_NVLS_BLOCKS", "48"))

    def __init__(self, group, device):
        import torch.distributed as dist
        self.group, self.dev = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.buf = self.handle = None
        self.mc = 0
        self.epoch = 1
        self.timeout_ms = int(float(os.environ.get("SEGCLIP_P2P_TIMEOUT_S", "600")) * 1000)
        self.stream = torch.cuda.Stream(device=device)
        self.err = torch.zeros(1, device=device, dtype=torch.int32)
        self.profile = None         # a list: all_reduce() appends (ready, start, end) CUDA events per bucket (tools/sync_timeline.py)

    @classmethod
    def create(cls, group, device):
        if os.environ.get("SEGCLIP_GRAD_SYNC", "").lower() == "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            self = cls(group, device)
            self._symm = symm
            # flag arrays of the per-CTA rank barriers
            nflags = L.SC_NVLS_MAX_BLOCKS * self.world
            self.flags = symm.empty(nflags, dtype=torch.int32, device=device)
            self.flags.zero_()
            fh = symm.rendezvous(self.flags, group)
            self._fh = fh
            if not getattr(fh, "has_multicast_support", lambda *a: True):
                return None
            self.peer_flags = torch.tensor([int(p) for p in fh.buffer_ptrs], dtype=torch.int64, device=device)
            torch.cuda.synchronize(device)
            fh.barrier()
            return self
        except Exception as e:      # no symmetric memory on this system: the caller keeps NCCL
            if os.environ.get("SEGCLIP_GRAD_SYNC", "").lower() == "nvls":
                raise
            import warnings
            warnings.warn("segclip_b200: NVLS gradient sync unavailable (%s: %s); using the NCCL bucket all-reduce" % (type(e).__name__, e))
            return None

    def alloc(self, numel):
        """Symmetric fp32 buffer of `numel` elements bound to a multicast object (collective over the group).
        Returns None if the multicast mapping is not available."""
        buf = self._symm.empty(numel, dtype=torch.float32, device=self.dev)
        buf.zero_()
        h = self._symm.rendezvous(buf, self.group)
        mc = int(getattr(h, "multicast_ptr", 0) or 0)
        if mc == 0:
            return None
        self.buf, self.handle, self.mc = buf, h, mc
        torch.cuda.synchronize(self.dev)
        h.barrier()
        return buf

    def all_reduce(self, start, end, after_stream):
        """Mean over ranks of gflat[start:end] (element offsets, multiples of 4), in place on every rank.  Runs on the sync
        stream after everything already queued on `after_stream`; join() makes a stream wait for all issued reductions."""
        prof = self.profile
        if prof is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record(after_stream)               # the bucket's last writer has been issued before this point
        self.stream.wait_stream(after_stream)
        if prof is not None:
            ev[1].record(self.stream)
        rc = L.lib().sc_nvls_allreduce(self.mc + 4 * start, end - start, 1.0 / self.world, self.rank, self.world,
                                       self.peer_flags.data_ptr(), self.epoch, self.BLOCKS, self.timeout_ms,
                                       self.err.data_ptr(), self.stream.cuda_stream)
        L.check(rc, "sc_nvls_allreduce")
        if prof is not None:
            ev[2].record(self.stream)
            prof.append(("bucket", 4 * (end - start), ev))
        self.epoch = (self.epoch + 2) & 0xFFFFFFFF

    def join(self, stream):
        stream.wait_stream(self.stream)

    def check(self):
        """Host-side check of the timeout flag (synchronises; tests / debugging only)."""
        if int(self.err.item()) != 0:
            raise L.SegclipB200Error("sc_nvls_allreduce: a rank did not arrive at the bucket barrier within the timeout")
