"""TEST INFRASTRUCTURE: run the package's host code (Compositor, stitcher, strips) on host
memory against the kernels compiled for the CPU (build_emul.py), so that the CPU test tier
exercises the real job tables, launch sequences and kernel source against the oracle.

Everything is done by monkeypatching from the test side; the package has no switch for it
and refuses to run without a CUDA device when used normally.
"""
import contextlib
import ctypes

import numpy as np
import torch

from . import build_emul


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def wait(self, stream=None):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return 0.0


class _FakeStream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass

    def synchronize(self):
        pass

    def record_event(self, event=None):
        return event or _FakeEvent()


def _load_library():
    from pano360_b200 import _lib
    lib = ctypes.CDLL(build_emul.build())
    for name, argtypes in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int64 if name == "p360_crop_scratch_bytes" else ctypes.c_int
    return lib


def install(monkeypatch):
    """Point the package at the host build of its kernels for the duration of a test."""
    from pano360_b200 import _lib, compositor, stitcher

    monkeypatch.setattr(_lib, "_lib", _load_library())
    stream = _FakeStream()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: stream)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda device: None)
    monkeypatch.setattr(compositor, "_require_cuda", lambda device: torch.device("cpu"))

    def to_host_tensor(self, array, pinned_key=None):
        return torch.from_numpy(np.array(array, copy=True, order="C"))

    monkeypatch.setattr(compositor.Compositor, "_to_device", to_host_tensor)

    # Fresh host allocations are zero pages, which would hide reads of memory no kernel wrote;
    # the GPU's caching allocator hands out stale data.  Poison every torch.empty instead.
    real_empty = torch.empty

    def poisoned_empty(*args, **kwargs):
        kwargs.pop("pin_memory", None)             # no page-locked memory without a driver
        t = real_empty(*args, **kwargs)
        if t.is_floating_point():
            t.fill_(float("nan"))
        elif t.dtype in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            t.fill_(0x5B)
        return t

    monkeypatch.setattr(torch, "empty", poisoned_empty)
    # the coarse pools are persistent and zeroed once in production; here every composite starts
    # from a large finite poison, so that a read of a cell nobody wrote shows wherever it matters
    monkeypatch.setattr(compositor.Compositor, "pool_fill", 1e30)
    monkeypatch.setattr(compositor.Compositor, "repoison", True)
    monkeypatch.setattr(stitcher, "_is_pinned_out", lambda out, shape: (
        out is not None and out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == tuple(shape) + (3,)))
    monkeypatch.setattr(stitcher, "_compositors", {})
    return compositor.Compositor()
