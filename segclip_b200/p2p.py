"""Multi-GPU exchange of the contrastive head: global-batch embeddings (forward) and per-row log-sum-exps (for
the reduction-free backward).

``P2PExchange``        one-shot NVLink peer-to-peer writes + epoch flags (libsegclip_b200 sc_p2p_*); the product path.
``CollectiveExchange`` torch.distributed.all_gather (NCCL on GPUs, gloo on CPU): the baseline comparator the
                       reference uses (diffdist all_gather, modules/util_module.py:180-190), kept for A/B runs and
                       for host-logic tests without GPUs.
Both expose: slot(B, E) -> per-batch-size buffers (t_all [N,E], v_all [N,E], lse_all [2,N]); gather_embeddings(slot);
gather_lse(slot); release(slot).  buffers(B, E) selects a slot for the slot-less forms of those calls.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L


class _DevMem:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class P2PChannel:
    """One symmetric buffer + signal pad per rank, peers mapped through CUDA IPC."""

    def __init__(self, group, device, nbytes):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = (nbytes + 255) // 256 * 256
        lib = L.lib()
        buf, pad = C.c_void_p(), C.c_void_p()
        hb, hp = C.create_string_buffer(64), C.create_string_buffer(64)
        with torch.cuda.device(device):
            L.check(lib.sc_p2p_alloc(self.nbytes, C.byref(buf), C.byref(pad), hb, hp), "sc_p2p_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, (hb.raw, hp.raw), group=group)
            self.bufs, self.pads = (C.c_void_p * self.world)(), (C.c_void_p * self.world)()
            self._opened = []
            for r, (b, p) in enumerate(handles):
                if r == self.rank:
                    self.bufs[r], self.pads[r] = buf.value, pad.value
                else:
                    pb, pp = C.c_void_p(), C.c_void_p()
                    L.check(lib.sc_p2p_open(b, C.byref(pb)), "sc_p2p_open")
                    L.check(lib.sc_p2p_open(p, C.byref(pp)), "sc_p2p_open")
                    self.bufs[r], self.pads[r] = pb.value, pp.value
                    self._opened += [pb.value, pp.value]
            self._own = (buf.value, pad.value)
            self.mem = torch.as_tensor(_DevMem(buf.value, self.nbytes), device=device)      # uint8 view of own buffer
            self.scratch = torch.zeros(2, dtype=torch.int32, device=device)
        self.epoch = 0
        self.released = 0
        dist.barrier(group=group)

    def view(self, byte_off, shape, dtype=torch.float32):
        n = int(torch.tensor(shape).prod()) * torch.empty((), dtype=dtype).element_size()
        return self.mem[byte_off:byte_off + n].view(dtype).view(*shape)

    def allgather(self, segments):
        """segments: [(src tensor, byte offset in the symmetric buffer)]"""
        if self.released < self.epoch:      # previous data never released (e.g. forward without backward)
            self.release()
        self.epoch += 1
        n = len(segments)
        srcs = (C.c_void_p * n)(*[t.data_ptr() for t, _ in segments])
        sizes = (C.c_int64 * n)(*[t.numel() * t.element_size() for t, _ in segments])
        offs = (C.c_int64 * n)(*[o for _, o in segments])
        L.check(L.lib().sc_p2p_allgather(n, srcs, sizes, offs, self.bufs, self.pads, self.rank, self.world, self.epoch,
                                         self.scratch.data_ptr(), L.stream()), "sc_p2p_allgather")

    def release(self):
        L.check(L.lib().sc_p2p_release(self.pads, self.rank, self.world, self.epoch, L.stream()), "sc_p2p_release")
        self.released = self.epoch

    def close(self):
        lib = L.lib()
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            lib.sc_p2p_close(p)
        lib.sc_p2p_free(*self._own)
        self._opened, self._own = [], (None, None)


class _Slot:
    """Buffers of one (per-rank batch, embedding width) pair.  An exchange keeps one slot per batch size it has seen, so a
    tail batch or an evaluation pass with another batch size never re-points the buffers a training plan was built on."""

    def __init__(self, B, E, world):
        self.B, self.E, self.N = B, E, B * world
        self.t_all = self.v_all = self.lse_all = None
        self.emb = self.lse = None          # P2PChannel pair (P2PExchange only)


class _ExchangeBase:
    def __init__(self, group, device):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slots = {}
        self.cur = None

    def slot(self, B, E):
        """Slot for per-rank batch B (created on first use; creation is COLLECTIVE for the P2P exchange: every rank must
        ask for a new batch size at the same point, which training loops do -- all ranks see the same batch sizes)."""
        key = (int(B), int(E))
        if key not in self.slots:
            self.slots[key] = self._make(_Slot(key[0], key[1], self.world))
        return self.slots[key]

    def buffers(self, B, E):
        """Selects (creating if needed) the slot for batch B and returns its (t_all [N,E], v_all [N,E], lse_all [2,N])."""
        self.cur = self.slot(B, E)
        return self.cur.t_all, self.cur.v_all, self.cur.lse_all

    # slot-less calls act on the slot selected by the last buffers() call
    def gather_embeddings(self, slot=None):
        self._gather_embeddings(slot or self.cur)

    def gather_lse(self, slot=None):
        self._gather_lse(slot or self.cur)

    def release(self, slot=None):
        self._release(slot or self.cur)


class P2PExchange(_ExchangeBase):
    """The product path: one NVLink P2P write kernel per exchange, no collective library call."""

    def _make(self, s):
        N, E = s.N, s.E
        s.emb = P2PChannel(self.group, self.device, 2 * N * E * 4)
        s.lse = P2PChannel(self.group, self.device, 2 * N * 4)
        s.t_all = s.emb.view(0, (N, E))
        s.v_all = s.emb.view(N * E * 4, (N, E))
        s.lse_all = s.lse.view(0, (2, N))
        return s

    def _gather_embeddings(self, s):
        B, E, N, r = s.B, s.E, s.N, self.rank
        lo = r * B
        s.emb.allgather([(s.t_all[lo:lo + B], lo * E * 4), (s.v_all[lo:lo + B], (N + lo) * E * 4)])

    def _gather_lse(self, s):
        B, N, r = s.B, s.N, self.rank
        lo = r * B
        s.lse.allgather([(s.lse_all[0, lo:lo + B], lo * 4), (s.lse_all[1, lo:lo + B], (N + lo) * 4)])

    def _release(self, s):
        s.emb.release()
        s.lse.release()


class CollectiveExchange(_ExchangeBase):
    """Baseline comparator: library all-gather (NCCL / gloo), as the reference does through diffdist."""

    def _make(self, s):
        dev = self.device
        s.t_all = torch.zeros(s.N, s.E, device=dev)
        s.v_all = torch.zeros(s.N, s.E, device=dev)
        s.lse_all = torch.zeros(2, s.N, device=dev)
        return s

    def _gather_rows(self, full, B):
        lo = self.rank * B
        mine = full[lo:lo + B].clone()
        dist.all_gather_into_tensor(full, mine, group=self.group) if full.is_cuda else \
            dist.all_gather(list(full.view(self.world, B, *full.shape[1:]).unbind(0)), mine, group=self.group)

    def _gather_embeddings(self, s):
        self._gather_rows(s.t_all, s.B)
        self._gather_rows(s.v_all, s.B)

    def _gather_lse(self, s):
        self._gather_rows(s.lse_all[0], s.B)
        self._gather_rows(s.lse_all[1], s.B)

    def _release(self, s):
        pass


def EmbeddingExchange(group, device):
    """Factory: SEGCLIP_EXCHANGE=nccl selects the library comparator, default is the P2P kernel."""
    if os.environ.get("SEGCLIP_EXCHANGE", "p2p") == "nccl" or torch.device(device).type != "cuda":
        return CollectiveExchange(group, device)
    return P2PExchange(group, device)
