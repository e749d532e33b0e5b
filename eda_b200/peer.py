"""Host side of the NVLink peer-memory reduction (csrc/peer_reduce.cu): allocates one symmetric exchange buffer per
rank through torch's symmetric-memory rendezvous (each rank's buffer is mapped into every process of the node) and hands
the device array of peer pointers to the kernels.  PyTorch is plumbing here — allocation and the pointer exchange; the
reduction itself is this package's kernel."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class PeerReduce:
    def __init__(self, device, group=None, max_elems=4096):
        import torch.distributed._symmetric_memory as symm_mem

        lib = _lib.load()
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.max_elems = int(max_elems)
        nbytes = lib.eda_peer_buffer_bytes(self.world, self.max_elems)
        if nbytes == 0:
            raise RuntimeError("eda_b200.peer: unsupported world size")
        self.buf = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world:
            raise RuntimeError("eda_b200.peer: rendezvous returned an unexpected number of peers")
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64).to(self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # every rank's buffer is zeroed before anybody's first exchange

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def all_reduce_sum_(self, t):
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64)
        n = t.numel()
        if n > self.max_elems:
            raise RuntimeError(f"eda_b200.peer: vector of {n} elements exceeds the exchange buffer ({self.max_elems})")
        with torch.cuda.device(self.device):
            rc = _lib.load().eda_peer_allreduce(_p(self.ptrs), self.world, self.rank, self.max_elems, _p(t), n,
                                                1 if t.dtype == torch.float64 else 0, self._stream())
        _lib.check(rc, "peer_allreduce")
        return t

    def error_word(self):
        """Non-zero once a peer failed to arrive within the kernel's spin limit (debugging aid; synchronises)."""
        return int(self.buf[4:8].view(torch.int32).item())
