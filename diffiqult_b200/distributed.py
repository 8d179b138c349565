"""torch.distributed plumbing for the sharded hot path (one process per GPU, NCCL over NVLink).

The C library deals its ERI tiles round-robin to `world_size` ranks and asks the host for two kinds of exchange:
one sum-allreduce of the partial J|K buffer per Fock build and one of the gradient buffer.  The callback built
here wraps the device pointer it is handed as a torch tensor (no copy) and calls torch.distributed.all_reduce on
the stream the library hands it (the cudaStream_t of dq_options.stream), whatever torch's current stream is.  With the
gloo backend (CPU tests) the pointer is a host pointer.

Failure behaviour: a rank whose call fails before a collective leaves the other ranks blocked in it (NCCL semantics);
treat any exception / non-zero status as fatal for the job (torchrun tears the other ranks down).
"""
import ctypes as C

import numpy as np

from . import _capi


class _CudaBuffer(object):
    """Minimal __cuda_array_interface__ view of `count` doubles at a raw device pointer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def make_allreduce(group=None, on_device=True):
    """-> (ctypes callback, keep-alive): pass the callback as dq_options.allreduce."""
    import torch
    import torch.distributed as dist
    calls = {"n": 0, "doubles": 0}
    views = {}      # (pointer, count) -> tensor view: the library's pooled buffers keep their addresses from call to call

    def _cb(ptr, count, stream, user):
        if on_device:
            t = views.get((ptr, count))
            if t is None:
                t = torch.as_tensor(_CudaBuffer(ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
                if len(views) < 64:
                    views[(ptr, count)] = t
        else:
            buf = (C.c_double * count).from_address(ptr)
            t = torch.from_numpy(np.frombuffer(buf, dtype=np.float64))
        if on_device:
            # order the collective on the LIBRARY's stream: its kernels wrote the buffer there and will read it there
            cur = torch.cuda.current_stream()
            if int(stream or 0) == int(cur.cuda_stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                with torch.cuda.stream(torch.cuda.ExternalStream(int(stream or 0))):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        calls["n"] += 1
        calls["doubles"] += int(count)

    cb = _capi.ALLREDUCE_FN(_cb)
    cb.calls = calls
    return cb


def current_stream_handle():
    import torch
    return torch.cuda.current_stream().cuda_stream
