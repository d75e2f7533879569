"""Multi-GPU exchange of the contrastive head: global-batch embeddings (forward) and per-row log-sum-exps (for
the reduction-free backward).

``P2PExchange``        one-shot NVLink peer-to-peer writes + epoch flags (libsegclip_b200 sc_p2p_*); the product path.
``CollectiveExchange`` torch.distributed.all_gather (NCCL on GPUs, gloo on CPU): the baseline comparator the
                       reference uses (diffdist all_gather, modules/util_module.py:180-190), kept for A/B runs and
                       for host-logic tests without GPUs.
Both expose: buffers(B, E) -> (t_all [N,E], v_all [N,E], lse_all [2,N]); gather_embeddings(); gather_lse(); release().
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


class _ExchangeBase:
    def __init__(self, group, device):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.B = self.E = None

    def _layout(self, B, E):
        self.B, self.E, self.N = B, E, B * self.world


class P2PExchange(_ExchangeBase):
    """The product path: one NVLink P2P write kernel per exchange, no collective library call."""

    def buffers(self, B, E):
        self._layout(B, E)
        N = self.N
        self.emb = P2PChannel(self.group, self.device, 2 * N * E * 4)
        self.lse = P2PChannel(self.group, self.device, 2 * N * 4)
        self.t_all = self.emb.view(0, (N, E))
        self.v_all = self.emb.view(N * E * 4, (N, E))
        self.lse_all = self.lse.view(0, (2, N))
        return self.t_all, self.v_all, self.lse_all

    def gather_embeddings(self):
        B, E, N, r = self.B, self.E, self.N, self.rank
        lo = r * B
        self.emb.allgather([(self.t_all[lo:lo + B], lo * E * 4), (self.v_all[lo:lo + B], (N + lo) * E * 4)])

    def gather_lse(self):
        B, N, r = self.B, self.N, self.rank
        lo = r * B
        self.lse.allgather([(self.lse_all[0, lo:lo + B], lo * 4), (self.lse_all[1, lo:lo + B], (N + lo) * 4)])

    def release(self):
        self.emb.release()
        self.lse.release()


class CollectiveExchange(_ExchangeBase):
    """Baseline comparator: library all-gather (NCCL / gloo), as the reference does through diffdist."""

    def buffers(self, B, E):
        self._layout(B, E)
        dev = self.device
        self.t_all = torch.zeros(self.N, E, device=dev)
        self.v_all = torch.zeros(self.N, E, device=dev)
        self.lse_all = torch.zeros(2, self.N, device=dev)
        return self.t_all, self.v_all, self.lse_all

    def _gather_rows(self, full):
        lo = self.rank * self.B
        mine = full[lo:lo + self.B].clone()
        dist.all_gather_into_tensor(full, mine, group=self.group) if full.is_cuda else \
            dist.all_gather(list(full.view(self.world, self.B, *full.shape[1:]).unbind(0)), mine, group=self.group)

    def gather_embeddings(self):
        self._gather_rows(self.t_all)
        self._gather_rows(self.v_all)

    def gather_lse(self):
        self._gather_rows(self.lse_all[0])
        self._gather_rows(self.lse_all[1])

    def release(self):
        pass


def EmbeddingExchange(group, device):
    """Factory: SEGCLIP_EXCHANGE=nccl selects the library comparator, default is the P2P kernel."""
    if os.environ.get("SEGCLIP_EXCHANGE", "p2p") == "nccl" or torch.device(device).type != "cuda":
        return CollectiveExchange(group, device)
    return P2PExchange(group, device)
