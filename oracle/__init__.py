"""oracle -- CPU checkers for the PLF newview path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and there only as the checker or the
reported CPU baseline.  The product (``amd-versal-phylogenetic-likelihood-function_b200``)
never imports it and has no CPU fallback.

Three independent statements of the reference algorithm (app/src/plf.cpp:8-68 in
/root/reference) live here:

* ``COracle``      -- ``liboracle.so`` built from ``plf_oracle.c`` (our C restatement),
* ``plf_numpy``    -- a vectorised numpy float32 restatement (``plf_numpy.py``),
* ``RefOracle``    -- ``_ref/libplf_ref.so``: the reference's own ``plf.cpp`` compiled in place
                      (built in the build container only; the prebuilt .so travels to the GPU box).

Parity status: PINNED -- see ``plf_oracle.h`` and ``tests/test_oracle.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libplf_ref.so")
LIB_REF_O0 = os.path.join(HERE, "_ref", "libplf_ref_O0.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_u8p = ctypes.POINTER(ctypes.c_ubyte)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and (when /root/reference is present) oracle/_ref/*.so."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


class COracle:
    """ctypes front end of liboracle.so (plf_oracle.h)."""

    def __init__(self, path: str = LIB_ORACLE):
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.plf_oracle_newview.restype = ctypes.c_int64
        L.plf_oracle_newview.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_size_t,
                                         _f32p, _f32p, _i32p, _u8p]
        L.plf_oracle_newview_ev4.restype = ctypes.c_int64
        L.plf_oracle_newview_ev4.argtypes = L.plf_oracle_newview.argtypes
        L.plf_oracle_newview_packed.restype = ctypes.c_int64
        L.plf_oracle_newview_packed.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_size_t,
                                                _f32p, _i32p, _u8p]
        L.plf_oracle_scaler_increment.restype = ctypes.c_int64
        L.plf_oracle_scaler_increment.argtypes = [_u8p, _i32p, ctypes.c_size_t]
        L.plf_oracle_transpose4.restype = None
        L.plf_oracle_transpose4.argtypes = [_f32p, _f32p]
        L.plf_oracle_newview_mt.restype = ctypes.c_int64
        L.plf_oracle_newview_mt.argtypes = L.plf_oracle_newview.argtypes + [ctypes.c_int]
        L.plf_oracle_newview_states.restype = ctypes.c_int64
        L.plf_oracle_newview_states.argtypes = [ctypes.c_int] + L.plf_oracle_newview.argtypes
        L.plf_oracle_newview_states_mt.restype = ctypes.c_int64
        L.plf_oracle_newview_states_mt.argtypes = L.plf_oracle_newview_states.argtypes + [ctypes.c_int]

    def newview(self, x1, x2, ev, left, right, wgt=None, nthreads: int = 1, ev4: bool = False):
        """Returns (x3[n,16] f32, scaler[n] u8, scaler_increment int)."""
        x1, x2, ev, left, right = map(_f32, (x1, x2, ev, left, right))
        n = x1.size // 16
        assert x1.size == n * 16 and x2.size == n * 16
        assert ev.size == (64 if ev4 else 16) and left.size == 64 and right.size == 64
        if wgt is not None:
            wgt = np.ascontiguousarray(wgt, dtype=np.int32)
            assert wgt.size == n
        x3 = np.empty((n, 16), dtype=np.float32)
        sc = np.empty(n, dtype=np.uint8)
        args = [_ptr(x1, _f32p), _ptr(x2, _f32p), _ptr(x3, _f32p), _ptr(ev, _f32p), n,
                _ptr(left, _f32p), _ptr(right, _f32p), _ptr(wgt, _i32p), _ptr(sc, _u8p)]
        if ev4:
            inc = self.lib.plf_oracle_newview_ev4(*args)
        elif nthreads > 1:
            inc = self.lib.plf_oracle_newview_mt(*args, nthreads)
        else:
            inc = self.lib.plf_oracle_newview(*args)
        return x3, sc, int(inc)

    def newview_states(self, states: int, x1, x2, ev, left, right, wgt=None, nthreads: int = 1):
        """General state count (4 = DNA, 20 = protein).  Returns (x3[n,4*S] f32, scaler[n] u8, increment)."""
        x1, x2, ev, left, right = map(_f32, (x1, x2, ev, left, right))
        S = int(states)
        n = x1.size // (4 * S)
        assert x1.size == n * 4 * S and x2.size == x1.size
        assert ev.size == S * S and left.size == 4 * S * S and right.size == 4 * S * S
        if wgt is not None:
            wgt = np.ascontiguousarray(wgt, dtype=np.int32)
            assert wgt.size == n
        x3 = np.empty((n, 4 * S), dtype=np.float32)
        sc = np.empty(n, dtype=np.uint8)
        args = [S, _ptr(x1, _f32p), _ptr(x2, _f32p), _ptr(x3, _f32p), _ptr(ev, _f32p), n,
                _ptr(left, _f32p), _ptr(right, _f32p), _ptr(wgt, _i32p), _ptr(sc, _u8p)]
        inc = (self.lib.plf_oracle_newview_states_mt(*args, nthreads) if nthreads > 1
               else self.lib.plf_oracle_newview_states(*args))
        if inc < 0:
            raise ValueError(f"unsupported state count {S}")
        return x3, sc, int(inc)

    def newview_packed(self, left_buf, right_buf, layout: int, n: int, wgt=None):
        left_buf, right_buf = _f32(left_buf), _f32(right_buf)
        assert left_buf.size >= 80 + 16 * n
        assert right_buf.size >= (80 if layout == 0 else 64) + 16 * n
        if wgt is not None:
            wgt = np.ascontiguousarray(wgt, dtype=np.int32)
        x3 = np.empty((n, 16), dtype=np.float32)
        sc = np.empty(n, dtype=np.uint8)
        inc = self.lib.plf_oracle_newview_packed(_ptr(left_buf, _f32p), _ptr(right_buf, _f32p),
                                                 layout, n, _ptr(x3, _f32p), _ptr(wgt, _i32p),
                                                 _ptr(sc, _u8p))
        return x3, sc, int(inc)

    def scaler_increment(self, scaler, wgt=None) -> int:
        scaler = np.ascontiguousarray(scaler, dtype=np.uint8)
        if wgt is not None:
            wgt = np.ascontiguousarray(wgt, dtype=np.int32)
        return int(self.lib.plf_oracle_scaler_increment(_ptr(scaler, _u8p), _ptr(wgt, _i32p),
                                                        scaler.size))

    def transpose4(self, m):
        m = _f32(m).reshape(16)
        out = np.empty(16, dtype=np.float32)
        self.lib.plf_oracle_transpose4(_ptr(m, _f32p), _ptr(out, _f32p))
        return out


class RefOracle:
    """ctypes front end of oracle/_ref/libplf_ref.so -- the reference's own plf()."""

    def __init__(self, path: str = LIB_REF):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not built (needs /root/reference; run `make -C oracle`)")
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.plf_ref_newview.restype = ctypes.c_int64
        L.plf_ref_newview.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_int, _f32p, _f32p, _i32p]
        L.plf_ref_newview_mt.restype = ctypes.c_int64
        L.plf_ref_newview_mt.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_size_t, _f32p, _f32p,
                                         _i32p, ctypes.c_int]

    @staticmethod
    def available(path: str = LIB_REF) -> bool:
        return os.path.exists(path)

    def newview(self, x1, x2, ev, left, right, wgt=None, nthreads: int = 1, out=None):
        """Returns (x3[n,16] f32, scaler_increment int).  The reference has no per-site bytes."""
        x1, x2, ev, left, right = map(_f32, (x1, x2, ev, left, right))
        n = x1.size // 16
        wgt = np.ones(n, dtype=np.int32) if wgt is None else np.ascontiguousarray(wgt, np.int32)
        x3 = np.empty((n, 16), dtype=np.float32) if out is None else out
        args = [_ptr(x1, _f32p), _ptr(x2, _f32p), _ptr(x3, _f32p), _ptr(ev, _f32p), n,
                _ptr(left, _f32p), _ptr(right, _f32p), _ptr(wgt, _i32p)]
        if nthreads > 1:
            inc = self.lib.plf_ref_newview_mt(*args, nthreads)
        else:
            inc = self.lib.plf_ref_newview(*args)
        return x3, int(inc)


# ---------------------------------------------------------------------------------------
# Synthetic stimulus: the recipe of app/src/host_mem.cpp:179-209, but seeded.
# ---------------------------------------------------------------------------------------

def host_mem_inputs(n: int, seed: int = 42):
    """EV[16], P_left[64], P_right[64], x1[n,16], x2[n,16], wgt[n] as host_mem.cpp:183-209.

    uniform(0,1) doubles cast to float; the left CLV of every 4th site (element index
    j % 64 < 16) is multiplied by float(1e-12) so that exactly those sites underflow.
    The random stream is numpy's RandomState (the reference uses an *unseeded* mt19937, so
    no particular stream is canonical)."""
    rng = np.random.RandomState(seed)
    ev = rng.random_sample(16).astype(np.float32)
    br = rng.random_sample(128)
    left = br[0::2].astype(np.float32)          # interleaved draws, host_mem.cpp:194-197
    right = br[1::2].astype(np.float32)
    d = rng.random_sample(2 * n * 16)
    j = np.arange(n * 16)
    scale = np.where(j % 64 < 16, np.float64(np.float32(1.0e-12)), 1.0)
    x1 = (d[0::2] * scale).astype(np.float32).reshape(n, 16)
    x2 = d[1::2].astype(np.float32).reshape(n, 16)
    wgt = np.ones(n, dtype=np.int32)
    return ev, left, right, x1, x2, wgt


def pack_buffers(ev, left, right, x1, x2, layout: int = 0):
    """[EV16|P64|CLV] / Sep right [P64|CLV] packing of host_mem.cpp:231-241."""
    x1 = _f32(x1).reshape(-1)
    x2 = _f32(x2).reshape(-1)
    lb = np.concatenate([_f32(ev).reshape(16), _f32(left).reshape(64), x1])
    if layout == 0:
        rb = np.concatenate([_f32(ev).reshape(16), _f32(right).reshape(64), x2])
    else:
        rb = np.concatenate([_f32(right).reshape(64), x2])
    return lb, rb
