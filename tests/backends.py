"""Two ways to drive the C ABI of include/pdes_b200.h from the tests.

* EmuBackend  - the kernels compiled for the CPU emulator (tests/_emu/libpdes_emu.so, built by ./build.sh --emu);
                buffers are numpy arrays.  Test infrastructure for the CPU suite.
* CudaBackend - the real sm_100a library; buffers are torch CUDA tensors.  Used by the -m gpu parity tests.
Both expose the same small helper API so each parity test is written once.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from neural_pde_surrogates_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libpdes_emu.so")


class _Base:
    lib = None

    def check(self, code):
        _native.check(self.lib, code)

    def tables(self, H, W, m1, m2):
        n = self.lib.pdes_tables_floats(H, W, m1, m2)
        buf = np.zeros(n, dtype=np.float32)
        self.check(self.lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data))
        return self.upload(buf)


class EmuBackend(_Base):
    name = "emu"

    def __init__(self):
        srcs = [os.path.join(ROOT, "neural_pde_surrogates_b200", "csrc", f)
                for f in os.listdir(os.path.join(ROOT, "neural_pde_surrogates_b200", "csrc"))]
        stale = (not os.path.exists(EMU_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in srcs)
        if stale:
            subprocess.check_call([os.path.join(ROOT, "build.sh"), "--emu"])
        self.lib = _native.bind(ctypes.CDLL(EMU_LIB))
        assert self.lib.pdes_is_cuda_build() == 0
        self.stream = None

    def upload(self, a):
        a = np.asarray(a)
        if np.iscomplexobj(a):
            return np.ascontiguousarray(a.astype(np.complex64))
        return np.ascontiguousarray(a.astype(np.float32))

    def empty(self, shape, complex_=False):
        return np.full(shape, np.nan, dtype=np.complex64 if complex_ else np.float32)

    def ptr(self, a):
        return None if a is None else a.ctypes.data

    def download(self, a):
        return np.array(a)

    def sync(self):
        pass


class CudaBackend(_Base):
    name = "cuda"

    def __init__(self):
        import torch
        self.torch = torch
        self.lib = _native.library()
        self.dev = torch.device("cuda:0")

    @property
    def stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def upload(self, a):
        a = np.asarray(a)
        a = a.astype(np.complex64) if np.iscomplexobj(a) else a.astype(np.float32)
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def empty(self, shape, complex_=False):
        t = self.torch
        return t.full(tuple(shape), float("nan"), dtype=t.complex64 if complex_ else t.float32, device=self.dev)

    def ptr(self, a):
        return None if a is None else a.data_ptr()

    def download(self, a):
        self.torch.cuda.synchronize()
        return a.detach().cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()
