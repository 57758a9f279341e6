"""A raw NCCL communicator (ctypes on the libnccl.so.2 torch already ships) so the ONE collective of the retrieval path --
the all-gather of per-rank score blocks (reference semantics worker.py:359: the full candidate row is
needed before the softmax / sort) -- is enqueued on the engine's compute stream, right behind the kernels that produce the block:
no host synchronisation, no hop onto torch.distributed's internal communication stream.

torch.distributed (any backend) is only used once, to hand rank 0's ncclUniqueId to the other ranks.
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib.util
import os

import torch

_NCCL_FLOAT32 = 7


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _find_libnccl():
    cands = []
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec is not None and spec.submodule_search_locations:
        for p in spec.submodule_search_locations:
            cands += sorted(glob.glob(os.path.join(p, "lib", "libnccl.so*")))
    cands += sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so*")))
    cands += ["libnccl.so.2"]
    for c in cands:
        try:
            return C.CDLL(c)
        except OSError:
            continue
    raise RuntimeError("libnccl.so.2 not found (looked in the nvidia-nccl wheel, torch/lib and the loader path)")


class NcclComm(object):
    def __init__(self, group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NcclComm needs an initialised torch.distributed process group (to exchange the NCCL unique id)")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.lib = _find_libnccl()
        self.lib.ncclGetErrorString.restype = C.c_char_p
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self.lib.ncclAllGather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        uid = _UniqueId()
        if self.rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        box = [bytes(uid.internal) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        C.memmove(C.byref(uid), box[0], 128)
        self.comm = C.c_void_p()
        self.device = torch.cuda.current_device()
        self._check(self.lib.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL error %d: %s" % (rc, self.lib.ncclGetErrorString(rc).decode()))

    def all_gather_f32(self, send: torch.Tensor, recv: torch.Tensor, stream=None):
        """recv[r * n:(r + 1) * n] = rank r's send (n = send.numel()), enqueued on `stream` (default: the current stream)."""
        assert send.is_cuda and recv.is_cuda and send.dtype == torch.float32 and recv.dtype == torch.float32
        assert send.is_contiguous() and recv.is_contiguous() and recv.numel() == self.world * send.numel()
        st = stream if stream is not None else torch.cuda.current_stream(send.device).cuda_stream
        self._check(self.lib.ncclAllGather(send.data_ptr(), recv.data_ptr(), send.numel(), _NCCL_FLOAT32, self.comm, C.c_void_p(st)))

    def close(self):
        if getattr(self, "comm", None):
            torch.cuda.synchronize(self.device)
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None
