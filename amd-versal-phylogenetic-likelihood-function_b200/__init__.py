"""B200-native Phylogenetic Likelihood Function (PLF) "newview" path.

Python face of ``libb200plf.so`` (C ABI: ``include/b200plf.h``).  It mirrors the operator
surface of the reference host program (GeertRoks/AMD-Versal-phylogenetic-likelihood-function,
``app/src/host_mem.cpp``): open device -> allocate per-instance buffers -> write packed
``[EV|P|CLV]`` inputs -> run -> read CLV + scaler bytes -> verify.  All arithmetic happens in
hand-written sm_100a CUDA kernels; this module only marshals pointers.  There is no CPU
fallback: if the shared library is missing or no GPU is present, calls raise ``PlfError``.

The directory name of this package is not a Python identifier; import it with
``importlib`` (see ``tests/conftest.py::load_pkg``) -- it registers itself as ``plf_b200``.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libb200plf.so")
HEADER_PATH = os.path.join(ROOT, "include", "b200plf.h")

SITE_FLOATS = 16
HEADER_COMB = 80
HEADER_SEP = 64
LAYOUT_COMB, LAYOUT_SEP = 0, 1
INPUT_MEM, INPUT_GEN = 0, 1
MATH_STRICT, MATH_FMA = 0, 1
GEN_WRITE, GEN_DISCARD = 0, 1
MARK_BEGIN, MARK_T1, MARK_T2, MARK_END = 0, 1, 2, 3
STATES_DNA, STATES_PROTEIN = 4, 20
LAUNCH_NO_PDL, LAUNCH_FENCED_RELEASE, LAUNCH_DEP_RELEASE, LAUNCH_SINGLE_CTA = 1, 2, 4, 8

# Algorithmic HBM bytes per site: read 64 (x1) + 64 (x2), write 64 (x3) + 1 scaler byte.
BYTES_PER_SITE = 193


class PlfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200plf error {code}: {msg}")
        self.code = code


class LaunchOpts(ctypes.Structure):
    _fields_ = [("math_mode", ctypes.c_int), ("variant", ctypes.c_int),
                ("threads_per_block", ctypes.c_int), ("blocks_per_sm", ctypes.c_int),
                ("ev_per_category", ctypes.c_int), ("flags", ctypes.c_int)]


_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_u = ctypes.c_uint
_i = ctypes.c_int

# name -> (restype, argtypes): every symbol include/b200plf.h declares.
PROTOTYPES = {
    "plf_device_count": (_i, [ctypes.POINTER(_i)]),
    "plf_device_info": (_i, [_i, ctypes.c_char_p, _sz, ctypes.c_char_p, _sz]),
    "plf_device_from_string": (_i, [ctypes.c_char_p, ctypes.POINTER(_i)]),
    "plf_ctx_create": (_i, [ctypes.POINTER(_vp), _i, _u, _i, _i]),
    "plf_ctx_create_states": (_i, [ctypes.POINTER(_vp), _i, _u, _i, _i, _i]),
    "plf_ctx_states": (_i, [_vp]),
    "plf_ctx_destroy": (_i, [_vp]),
    "plf_last_error": (ctypes.c_char_p, [_vp]),
    "plf_ctx_set_math": (_i, [_vp, _i]),
    "plf_ctx_set_gen_sink": (_i, [_vp, _i]),
    "plf_ctx_set_tuning": (_i, [_vp, _i, _i, _i]),
    "plf_ctx_instances": (_u, [_vp]),
    "plf_instance_alloc": (_i, [_vp, _u, _sz]),
    "plf_instance_free": (_i, [_vp, _u]),
    "plf_write_left": (_i, [_vp, _u, _vp, _sz, _sz]),
    "plf_write_right": (_i, [_vp, _u, _vp, _sz, _sz]),
    "plf_write_wgt": (_i, [_vp, _u, _vp, _sz]),
    "plf_run_async": (_i, [_vp, _u, _sz]),
    "plf_wait": (_i, [_vp, _u]),
    "plf_read_out": (_i, [_vp, _u, _vp, _sz, _sz]),
    "plf_read_scaler": (_i, [_vp, _u, _vp, _sz, _sz]),
    "plf_scaler_increment": (_i, [_vp, _u, ctypes.POINTER(ctypes.c_longlong)]),
    "plf_gen_checksum": (_i, [_vp, _u, ctypes.POINTER(ctypes.c_double)]),
    "plf_mark": (_i, [_vp, _u, _i]),
    "plf_elapsed_ms": (_i, [_vp, _u, _i, _i, ctypes.POINTER(ctypes.c_float)]),
    "plf_instance_device_ptrs": (_i, [_vp, _u] + [ctypes.POINTER(_vp)] * 4),
    "plf_instance_stream": (_i, [_vp, _u, ctypes.POINTER(_vp)]),
    "plf_newview_stream": (_i, [_vp] * 9 + [_sz, _sz, ctypes.POINTER(ctypes.c_longlong)]),
    "plf_host_alloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "plf_host_free": (_i, [_vp]),
    "plf_host_register": (_i, [_vp, _sz]),
    "plf_host_unregister": (_i, [_vp]),
    "plf_newview_device": (_i, [_vp] * 8 + [_sz, _vp, ctypes.POINTER(LaunchOpts), _vp]),
    "plf_newview_gen_device": (_i, [_vp, _vp, _sz, _vp, _vp, _i, ctypes.POINTER(LaunchOpts), _vp]),
    "plf_gen_pattern": (_i, [_vp] * 5),
    "plf_generate_device": (_i, [_vp, _vp, _sz, _sz, ctypes.c_uint64, _vp]),
    "plf_generate_host": (_i, [_vp, _vp, _sz, _sz, ctypes.c_uint64]),
    "plf_tree_create": (_i, [ctypes.POINTER(_vp), _i, _u, _vp, _vp, _sz]),
    "plf_tree_create_ex": (_i, [ctypes.POINTER(_vp), _i, _u, _vp, _vp, _sz, _i]),
    "plf_tree_create_states": (_i, [ctypes.POINTER(_vp), _i, _u, _vp, _vp, _sz, _i, _i]),
    "plf_tree_write_tip_codes": (_i, [_vp, _u, _vp, _sz, _sz]),
    "plf_tree_write_tip_vector": (_i, [_vp, _vp]),
    "plf_tree_destroy": (_i, [_vp]),
    "plf_tree_last_error": (ctypes.c_char_p, [_vp]),
    "plf_tree_set_math": (_i, [_vp, _i]),
    "plf_tree_set_tuning": (_i, [_vp, _i, _i]),
    "plf_tree_tip_ptr": (_i, [_vp, _u, ctypes.POINTER(_vp)]),
    "plf_tree_write_tip": (_i, [_vp, _u, _vp, _sz, _sz]),
    "plf_tree_write_matrices": (_i, [_vp, _vp, _vp, _vp]),
    "plf_tree_write_wgt": (_i, [_vp, _vp]),
    "plf_tree_run_async": (_i, [_vp]),
    "plf_tree_wait": (_i, [_vp]),
    "plf_tree_read_root": (_i, [_vp, _vp, _vp, _sz, _sz]),
    "plf_tree_total_scalings": (_i, [_vp, ctypes.POINTER(ctypes.c_longlong)]),
    "plf_tree_info": (_i, [_vp, ctypes.POINTER(_u), ctypes.POINTER(_u), ctypes.POINTER(_sz),
                           ctypes.POINTER(_sz)]),
    "plf_tree_last_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "plf_tree_evaluate_root": (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_double)]),
    "plf_evaluate_device": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "plf_evaluate_states_device": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "plf_newview_states_device": (_i, [_i] + [_vp] * 8 + [_sz, _vp, ctypes.POINTER(LaunchOpts), _vp]),
    "plf_generate_states_device": (_i, [_i, _vp, _vp, _sz, _sz, ctypes.c_uint64, _vp]),
    "plf_generate_states_host": (_i, [_i, _vp, _vp, _sz, _sz, ctypes.c_uint64]),
    "plf_states_kernel_info": (_i, [_i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_sz),
                                    ctypes.POINTER(_i)]),
    "plf_device_malloc": (_i, [_i, ctypes.POINTER(_vp), _sz]),
    "plf_device_free": (_i, [_vp]),
    "plf_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "plf_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "plf_memset_device": (_i, [_vp, _i, _sz, _vp]),
    "plf_stream_sync": (_i, [_vp]),
    "plf_kernel_info": (_i, [_i, _i] + [ctypes.POINTER(_i)] * 4),
    "plf_launch_count": (ctypes.c_ulonglong, []),
    "plf_set_release_mode": (_i, [_i]),
    "plf_range_push": (_i, [ctypes.c_char_p]),
    "plf_range_pop": (_i, []),
    "plf_multi_create": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_i), _i, _u, _i, _i]),
    "plf_multi_create_states": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_i), _i, _u, _i, _i, _i]),
    "plf_multi_destroy": (_i, [_vp]),
    "plf_multi_last_error": (ctypes.c_char_p, [_vp]),
    "plf_multi_size": (_i, [_vp]),
    "plf_multi_ctx": (_vp, [_vp, _i]),
    "plf_multi_partition": (_i, [_sz, _i, _i, ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "plf_multi_newview": (_i, [_vp] * 9 + [_sz, ctypes.POINTER(ctypes.c_longlong)]),
    "plf_multi_reduce": (_i, [_vp, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double),
                              ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double)]),
    "plf_multi_info": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_ulonglong)]),
    "plf_probe_host_link": (_i, [_i, _vp, _sz, _vp, _sz, _i, _i, ctypes.POINTER(ctypes.c_double)]),
}

_lib = None


def build(force: bool = False, quiet: bool = True) -> str:
    """Compile libb200plf.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["make", "-C", ROOT, "lib"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)
    return LIB_PATH


def load():
    """dlopen libb200plf.so and bind every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlfError(-2, f"{LIB_PATH} is not built (run `make lib`); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int, ctx=None):
    if rc != 0:
        msg = load().plf_last_error(ctx)
        raise PlfError(rc, msg.decode() if msg else "unknown error")


def device_count() -> int:
    n = _i(0)
    rc = load().plf_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


def device_info(device: int = 0):
    name = ctypes.create_string_buffer(256)
    bdf = ctypes.create_string_buffer(32)
    _check(load().plf_device_info(device, name, 256, bdf, 32))
    return name.value.decode(), bdf.value.decode()


def set_release_mode(mode: int) -> None:
    """Ring-slot release of the bulk-copy kernels: 1 fenced everywhere, 0 data dependency everywhere, -1 defaults."""
    _check(load().plf_set_release_mode(mode))


def probe_host_link(device: int, host_in=None, host_out=None, h2d_bytes: int = 0, d2h_bytes: int = 0, reps: int = 4,
                    pieces: int = 1) -> float:
    """Seconds for `reps` rounds of bare pinned H2D + D2H copies issued together (no kernels).  host_in / host_out:
    pinned numpy arrays (host_alloc) or None (allocated inside, h2d_bytes / d2h_bytes give the sizes)."""
    sec = ctypes.c_double(0.0)
    pi = _ptr(host_in) if host_in is not None else None
    po = _ptr(host_out) if host_out is not None else None
    nb_in = host_in.nbytes if host_in is not None else h2d_bytes
    nb_out = host_out.nbytes if host_out is not None else d2h_bytes
    _check(load().plf_probe_host_link(device, pi, nb_in, po, nb_out, reps, pieces, ctypes.byref(sec)))
    return sec.value


def launch_count() -> int:
    return int(load().plf_launch_count())


def _ptr(a):
    """numpy array / int / None -> void* value."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


# ------------------------------------------------------------------------------------------
# Host-side size / partition math of the reference test bench (app/src/include.h:150-268),
# with 64-bit sizes (the reference's `unsigned int` byte counts wrap at 4 GiB).
# ------------------------------------------------------------------------------------------
@dataclass
class TestbenchInfo:
    __test__ = False  # not a pytest class
    alignment_sites: int
    parallel_instances: int = 1
    plf_calls: int = 1
    window_size: int = 8192
    layout: int = LAYOUT_COMB
    elements_per_alignment: int = SITE_FLOATS
    word_size: int = 4

    def alignments_per_instance(self, instance: int | None = None) -> int:
        """ceil(n / instances); the last instance takes the remainder (include.h:181-192)."""
        per = math.ceil(self.alignment_sites / self.parallel_instances)
        if instance is None:
            return per
        if instance == self.parallel_instances - 1:
            return per - self.alignments_padding()
        return per

    def alignments_padding(self) -> int:
        return self.alignments_per_instance() * self.parallel_instances - self.alignment_sites

    def instance_offset(self, instance: int) -> int:
        """First site of an instance: k * alignments_per_instance(0) (host_mem.cpp:229,290)."""
        return instance * self.alignments_per_instance()

    @property
    def states(self) -> int:
        return self.elements_per_alignment // 4

    def header_left(self) -> int:
        """[EV S^2 | P 4S^2] in front of the CLV (80 floats for DNA, 2000 for protein)."""
        return 5 * self.states * self.states

    def header_right(self) -> int:
        return self.header_left() if self.layout == LAYOUT_COMB else 4 * self.states * self.states

    def instance_active_elements_left(self, instance: int) -> int:
        return self.alignments_per_instance(instance) * self.elements_per_alignment + self.header_left()

    def instance_active_elements_right(self, instance: int) -> int:
        return self.alignments_per_instance(instance) * self.elements_per_alignment + self.header_right()

    def elements_per_plf(self) -> int:
        return self.alignment_sites * self.elements_per_alignment

    def data_size(self) -> int:
        """Bytes of result CLV over all calls -- what timing.h:101-103 divides by time."""
        return self.elements_per_plf() * self.word_size * self.plf_calls

    def valid(self) -> bool:
        """Every instance must own at least one site (the reference silently underflows)."""
        return (self.alignment_sites > 0 and self.parallel_instances > 0 and
                self.alignments_per_instance(self.parallel_instances - 1) > 0)


def partition_sites(n: int, parts: int):
    """[(first_site, count)] per part with the reference rule (include.h:181-192).  Parts that
    would be empty or negative under that rule are returned with count 0."""
    per = math.ceil(n / parts) if parts > 0 else 0
    out = []
    for k in range(parts):
        lo = min(k * per, n)
        out.append((lo, max(0, min(per, n - lo))))
    return out


def pack_left(ev, p_left, x1):
    """[EV S^2 | P_left 4S^2 | CLV] (host_mem.cpp:231-233; S = 4: [EV16 | P64 | CLV]); S from the size of EV."""
    return np.concatenate([np.asarray(ev, np.float32).reshape(-1),
                           np.asarray(p_left, np.float32).reshape(-1),
                           np.asarray(x1, np.float32).reshape(-1)])


def pack_right(ev, p_right, x2, layout: int = LAYOUT_COMB):
    """Comb: [EV | P_right | CLV]; Sep: [P_right | CLV] (host_mem.cpp:234-241; 16 | 64 floats for DNA, 400 | 1600 for AA)."""
    parts = [np.asarray(p_right, np.float32).reshape(-1), np.asarray(x2, np.float32).reshape(-1)]
    if layout == LAYOUT_COMB:
        parts.insert(0, np.asarray(ev, np.float32).reshape(-1))
    return np.concatenate(parts)


class Context:
    """One GPU with NUM_ACCELERATORS independent PLF instances (CUDA streams).

    Mirrors the XRT objects the reference host holds: acap_info + per-instance kernels, buffer
    objects and run handles (host_mem.cpp:108-157)."""

    def __init__(self, device: int = 0, n_instances: int = 1, layout: int = LAYOUT_COMB,
                 input_src: int = INPUT_MEM, _borrowed: int | None = None, states: int = STATES_DNA):
        self.lib = load()
        self._owned = _borrowed is None
        self.states = states
        if _borrowed is None:
            self._ctx = _vp()
            _check(self.lib.plf_ctx_create_states(ctypes.byref(self._ctx), device, n_instances, layout, input_src, states))
        else:
            self._ctx = _vp(_borrowed)          # a context that belongs to a Multi
        self.device = device
        self.n_instances = n_instances
        self.layout = layout
        self.input_src = input_src

    def close(self):
        if self._ctx and self._owned:
            self.lib.plf_ctx_destroy(self._ctx)
        self._ctx = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _check(rc, self._ctx)

    def set_math(self, mode: int):
        self._ck(self.lib.plf_ctx_set_math(self._ctx, mode))

    def set_gen_sink(self, sink: int):
        self._ck(self.lib.plf_ctx_set_gen_sink(self._ctx, sink))

    def set_tuning(self, variant: int = 0, threads: int = 0, blocks_per_sm: int = 0):
        self._ck(self.lib.plf_ctx_set_tuning(self._ctx, variant, threads, blocks_per_sm))

    def instance_alloc(self, inst: int, max_sites: int):
        self._ck(self.lib.plf_instance_alloc(self._ctx, inst, max_sites))

    def instance_free(self, inst: int):
        self._ck(self.lib.plf_instance_free(self._ctx, inst))

    def write_left(self, inst: int, packed, nbytes: int | None = None, offset: int = 0):
        nbytes = packed.nbytes if nbytes is None else nbytes
        self._ck(self.lib.plf_write_left(self._ctx, inst, _ptr(packed), nbytes, offset))

    def write_right(self, inst: int, packed, nbytes: int | None = None, offset: int = 0):
        nbytes = packed.nbytes if nbytes is None else nbytes
        self._ck(self.lib.plf_write_right(self._ctx, inst, _ptr(packed), nbytes, offset))

    def write_wgt(self, inst: int, wgt):
        if wgt is None:
            self._ck(self.lib.plf_write_wgt(self._ctx, inst, None, 0))
        else:
            assert wgt.dtype == np.int32 and wgt.flags.c_contiguous
            self._ck(self.lib.plf_write_wgt(self._ctx, inst, _ptr(wgt), wgt.size))

    def run_async(self, inst: int, sites: int):
        self._ck(self.lib.plf_run_async(self._ctx, inst, sites))

    def wait(self, inst: int):
        self._ck(self.lib.plf_wait(self._ctx, inst))

    def read_out(self, inst: int, dst, nbytes: int | None = None, offset: int = 0):
        nbytes = dst.nbytes if nbytes is None else nbytes
        self._ck(self.lib.plf_read_out(self._ctx, inst, _ptr(dst), nbytes, offset))

    def read_scaler(self, inst: int, dst, nbytes: int | None = None, offset: int = 0):
        nbytes = dst.nbytes if nbytes is None else nbytes
        self._ck(self.lib.plf_read_scaler(self._ctx, inst, _ptr(dst), nbytes, offset))

    def scaler_increment(self, inst: int) -> int:
        v = ctypes.c_longlong(0)
        self._ck(self.lib.plf_scaler_increment(self._ctx, inst, ctypes.byref(v)))
        return v.value

    def gen_checksum(self, inst: int) -> float:
        v = ctypes.c_double(0)
        self._ck(self.lib.plf_gen_checksum(self._ctx, inst, ctypes.byref(v)))
        return v.value

    def mark(self, inst: int, mark_id: int):
        self._ck(self.lib.plf_mark(self._ctx, inst, mark_id))

    def elapsed_ms(self, inst: int, a: int, b: int) -> float:
        v = ctypes.c_float(0)
        self._ck(self.lib.plf_elapsed_ms(self._ctx, inst, a, b, ctypes.byref(v)))
        return v.value

    def device_ptrs(self, inst: int):
        p = [_vp() for _ in range(4)]
        self._ck(self.lib.plf_instance_device_ptrs(self._ctx, inst, *[ctypes.byref(x) for x in p]))
        return tuple(x.value for x in p)

    def stream(self, inst: int) -> int:
        s = _vp()
        self._ck(self.lib.plf_instance_stream(self._ctx, inst, ctypes.byref(s)))
        return s.value or 0

    def newview_stream(self, ev, p_left, p_right, x1, x2, x3, scaler=None, wgt=None, n_sites: int | None = None,
                       chunk_sites: int = 0) -> int:
        """Streamed round trip over host arrays (numpy arrays or raw pointers); returns the scaler increment."""
        ev = np.ascontiguousarray(ev, np.float32)
        pl = np.ascontiguousarray(p_left, np.float32)
        pr = np.ascontiguousarray(p_right, np.float32)
        n = n_sites if n_sites is not None else x1.size // (4 * self.states)
        inc = ctypes.c_longlong(0)
        self._ck(self.lib.plf_newview_stream(self._ctx, _ptr(ev), _ptr(pl), _ptr(pr), _ptr(x1), _ptr(x2), _ptr(x3),
                                             _ptr(scaler), _ptr(wgt), n, chunk_sites, ctypes.byref(inc)))
        return inc.value

    # -- the reference host's per-call sequence (host_mem.cpp:287-325) for numpy inputs --------
    def newview(self, ev, p_left, p_right, x1, x2, wgt=None, instances: int | None = None,
                timings: bool = False):
        """Packs [EV|P|CLV] per instance, writes, runs, reads back.  Returns
        (x3[n,16] f32, scaler[n] u8, scaler_increment)."""
        if self.input_src != INPUT_MEM:
            raise PlfError(-4, "Context.newview needs INPUT_SRC=mem")
        sf = 4 * self.states
        x1 = np.ascontiguousarray(x1, np.float32).reshape(-1, sf)
        x2 = np.ascontiguousarray(x2, np.float32).reshape(-1, sf)
        n = x1.shape[0]
        k_inst = self.n_instances if instances is None else instances
        tb = TestbenchInfo(n, k_inst, layout=self.layout, elements_per_alignment=sf)
        if n and not tb.valid():
            raise PlfError(-1, f"{n} sites cannot be split over {k_inst} instances "
                               "(last instance would be empty)")
        out = np.empty((n, sf), np.float32)
        sc = np.empty(n, np.uint8)
        if wgt is not None:
            wgt = np.ascontiguousarray(wgt, np.int32)
        keep = []
        total = 0
        if n == 0:
            return out, sc, 0
        for k in range(k_inst):
            lo, cnt = tb.instance_offset(k), tb.alignments_per_instance(k)
            lb = pack_left(ev, p_left, x1[lo:lo + cnt])
            rb = pack_right(ev, p_right, x2[lo:lo + cnt], self.layout)
            keep += [lb, rb]
            self.instance_alloc(k, cnt)
            self.mark(k, MARK_BEGIN)
            self.write_left(k, lb, tb.instance_active_elements_left(k) * 4)
            self.write_right(k, rb, tb.instance_active_elements_right(k) * 4)
            self.write_wgt(k, None if wgt is None else wgt[lo:lo + cnt])
            self.mark(k, MARK_T1)
            self.run_async(k, cnt)
            self.mark(k, MARK_T2)
            self.read_out(k, out[lo:lo + cnt])
            self.read_scaler(k, sc[lo:lo + cnt])
            self.mark(k, MARK_END)
        for k in range(k_inst):
            self.wait(k)
            total += self.scaler_increment(k)
        if timings:
            t = [(self.elapsed_ms(k, MARK_BEGIN, MARK_T1), self.elapsed_ms(k, MARK_T1, MARK_T2),
                  self.elapsed_ms(k, MARK_T2, MARK_END)) for k in range(k_inst)]
            return out, sc, total, t
        return out, sc, total


class Multi:
    """Several GPUs of one box from one process (plf_multi_*): the reference's ceil(n/parts) site split over the GPUs,
    one Context per GPU, and the final NCCL all-reduce of the scaler increments / log-likelihoods."""

    def __init__(self, devices, n_instances: int = 1, layout: int = LAYOUT_COMB, input_src: int = INPUT_MEM,
                 states: int = STATES_DNA):
        self.lib = load()
        self.devices = [int(d) for d in devices]
        self.states = states
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        self._m = _vp()
        rc = self.lib.plf_multi_create_states(ctypes.byref(self._m), arr, len(self.devices), n_instances, layout, input_src,
                                              states)
        if rc != 0:
            raise PlfError(rc, self.lib.plf_multi_last_error(None).decode())
        self.contexts = [Context(d, n_instances, layout, input_src, _borrowed=self.lib.plf_multi_ctx(self._m, r), states=states)
                         for r, d in enumerate(self.devices)]

    def _ck(self, rc):
        if rc != 0:
            raise PlfError(rc, self.lib.plf_multi_last_error(self._m).decode())

    def close(self):
        if self._m:
            for c in self.contexts:
                c.close()
            self.lib.plf_multi_destroy(self._m)
            self._m = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def partition(self, n_sites: int, rank: int):
        first, cnt = _sz(0), _sz(0)
        self._ck(self.lib.plf_multi_partition(n_sites, len(self.devices), rank, ctypes.byref(first), ctypes.byref(cnt)))
        return first.value, cnt.value

    def newview(self, ev, p_left, p_right, x1, x2, wgt=None):
        """Host arrays in, host arrays out: (x3[n,16], scaler[n], NCCL-reduced scaler increment)."""
        ev = np.ascontiguousarray(ev, np.float32)
        pl = np.ascontiguousarray(p_left, np.float32)
        pr = np.ascontiguousarray(p_right, np.float32)
        sf = 4 * self.states
        x1 = np.ascontiguousarray(x1, np.float32).reshape(-1, sf)
        x2 = np.ascontiguousarray(x2, np.float32).reshape(-1, sf)
        n = x1.shape[0]
        x3 = np.empty((n, sf), np.float32)
        sc = np.empty(n, np.uint8)
        if wgt is not None:
            wgt = np.ascontiguousarray(wgt, np.int32)
        inc = ctypes.c_longlong(0)
        self._ck(self.lib.plf_multi_newview(self._m, _ptr(ev), _ptr(pl), _ptr(pr), _ptr(x1), _ptr(x2), _ptr(x3), _ptr(sc),
                                            _ptr(wgt), n, ctypes.byref(inc)))
        return x3, sc, inc.value

    def reduce(self, increments=None, lnl=None):
        g = len(self.devices)
        ia = (ctypes.c_longlong * g)(*[int(v) for v in increments]) if increments is not None else None
        la = (ctypes.c_double * g)(*[float(v) for v in lnl]) if lnl is not None else None
        it, lt = ctypes.c_longlong(0), ctypes.c_double(0.0)
        self._ck(self.lib.plf_multi_reduce(self._m, ia, la, ctypes.byref(it), ctypes.byref(lt)))
        return it.value, lt.value

    def info(self):
        n, v, r = _i(0), _i(0), ctypes.c_ulonglong(0)
        self._ck(self.lib.plf_multi_info(self._m, ctypes.byref(n), ctypes.byref(v), ctypes.byref(r)))
        return {"n_devices": n.value, "nccl_version": v.value, "reductions": r.value}


def make_opts(math_mode: int = MATH_STRICT, variant: int = 0, threads: int = 0,
              blocks_per_sm: int = 0, ev_per_category: int = 0, flags: int = 0) -> LaunchOpts:
    return LaunchOpts(math_mode, variant, threads, blocks_per_sm, ev_per_category, flags)


def newview_device(x1, x2, x3, scaler, ev, p_left, p_right, wgt, n: int, scaler_sum,
                   opts: LaunchOpts | None = None, stream: int = 0):
    """Fused newview on caller-owned DEVICE memory; every pointer is an int (e.g. tensor.data_ptr())."""
    _check(load().plf_newview_device(x1, x2, x3, scaler, ev, p_left, p_right, wgt, n, scaler_sum,
                                     ctypes.byref(opts) if opts is not None else None, stream or None))


def newview_gen_device(x3, scaler, n: int, scaler_sum, checksum, sink: int = GEN_WRITE,
                       opts: LaunchOpts | None = None, stream: int = 0):
    _check(load().plf_newview_gen_device(x3, scaler, n, scaler_sum, checksum, sink,
                                         ctypes.byref(opts) if opts is not None else None,
                                         stream or None))


def gen_pattern():
    """(x1[16], x2[16], ev4[64], p_left[64], p_right[64]) of the INPUT_SRC=gen movers."""
    arrs = [np.empty(k, np.float32) for k in (16, 16, 64, 64, 64)]
    _check(load().plf_gen_pattern(*[_ptr(a) for a in arrs]))
    return tuple(arrs)


def generate_device(x1, x2, first_site: int, n: int, seed: int, stream: int = 0):
    _check(load().plf_generate_device(x1, x2, first_site, n, seed, stream or None))


def generate_host(first_site: int, n: int, seed: int):
    x1 = np.empty((n, 16), np.float32)
    x2 = np.empty((n, 16), np.float32)
    _check(load().plf_generate_host(_ptr(x1), _ptr(x2), first_site, n, seed))
    return x1, x2


def kernel_info(variant: int = 0, math_mode: int = MATH_STRICT, threads: int = 0):
    regs, thr, bps, sms = _i(0), _i(threads), _i(0), _i(0)
    _check(load().plf_kernel_info(variant, math_mode, ctypes.byref(regs), ctypes.byref(thr),
                                  ctypes.byref(bps), ctypes.byref(sms)))
    return {"regs": regs.value, "threads": thr.value, "blocks_per_sm": bps.value, "sms": sms.value}


# ---- the packed buffers as files (host/plfb_file.h) ------------------------------------------
PLFB_LEFT, PLFB_RIGHT, PLFB_OUT, PLFB_SCALER = 0, 1, 2, 3
_PLFB_HEADER = "<4sIIIIIQQ24s"          # magic, version, kind, layout, states, categories, sites, payload bytes, reserved


def plfb_payload_bytes(kind: int, layout: int, sites: int) -> int:
    if kind == PLFB_LEFT:
        return (80 + 16 * sites) * 4
    if kind == PLFB_RIGHT:
        return ((80 if layout == LAYOUT_COMB else 64) + 16 * sites) * 4
    if kind == PLFB_OUT:
        return 16 * sites * 4
    if kind == PLFB_SCALER:
        return sites
    raise ValueError(f"unknown PLFB buffer kind {kind}")


def save_plfb(path: str, kind: int, layout: int, sites: int, array) -> None:
    """Write one packed buffer ([EV|P|CLV] input, CLV output or scaler bytes) as a PLFB file."""
    import struct
    a = np.ascontiguousarray(array, dtype=np.uint8 if kind == PLFB_SCALER else np.float32)
    nbytes = plfb_payload_bytes(kind, layout, sites)
    if a.nbytes != nbytes:
        raise ValueError(f"buffer has {a.nbytes} bytes, kind {kind} / layout {layout} / {sites} sites needs {nbytes}")
    with open(path, "wb") as f:
        f.write(struct.pack(_PLFB_HEADER, b"PLFB", 1, kind, layout, 4, 4, sites, nbytes, b""))
        f.write(a.tobytes())


def load_plfb(path: str):
    """Read and validate a PLFB file -> {"kind", "layout", "sites", "data" (float32 or uint8 array)}."""
    import struct
    with open(path, "rb") as f:
        raw = f.read(64)
        if len(raw) != 64:
            raise ValueError(f"{path} is not a PLFB file")
        magic, version, kind, layout, states, cats, sites, nbytes, _ = struct.unpack(_PLFB_HEADER, raw)
        if magic != b"PLFB":
            raise ValueError(f"{path} is not a PLFB file")
        if version != 1 or states != 4 or cats != 4:
            raise ValueError(f"{path}: unsupported version/states/categories {version}/{states}/{cats}")
        if nbytes != plfb_payload_bytes(kind, layout, sites):
            raise ValueError(f"{path}: payload size does not match kind/layout/sites")
        payload = f.read(nbytes)
    if len(payload) != nbytes:
        raise ValueError(f"{path} is truncated")
    data = np.frombuffer(payload, dtype=np.uint8 if kind == PLFB_SCALER else np.float32).copy()
    return {"kind": kind, "layout": layout, "sites": sites, "data": data}


# ---- general state count: the reference's STATES knob (4 = DNA, 20 = protein; README.md:36,67,202) ----


def newview_states_device(states: int, x1, x2, x3, scaler, ev, p_left, p_right, wgt, n: int, scaler_sum,
                          opts: LaunchOpts | None = None, stream: int = 0):
    """plf() for `states` states.  CLVs / scaler / wgt / scaler_sum are DEVICE pointers (ints); ev [S*S],
    p_left / p_right [4*S*S] are HOST numpy arrays (they travel as kernel arguments).  For 20 states
    opts.variant = sites per lane (1, 2, 4; 0 = default for the math mode)."""
    S = int(states)
    mats = []
    for a, k in ((ev, S * S), (p_left, 4 * S * S), (p_right, 4 * S * S)):
        a = np.ascontiguousarray(a, np.float32).reshape(-1)
        if a.size != k:
            raise ValueError(f"matrix has {a.size} floats, expected {k} for {S} states")
        mats.append(a)
    _check(load().plf_newview_states_device(S, x1, x2, x3, scaler, _ptr(mats[0]), _ptr(mats[1]), _ptr(mats[2]), wgt, n,
                                            scaler_sum, ctypes.byref(opts) if opts is not None else None, stream or None))


def generate_states_device(states: int, x1, x2, first_site: int, n: int, seed: int, stream: int = 0):
    _check(load().plf_generate_states_device(states, x1, x2, first_site, n, seed, stream or None))


def generate_states_host(states: int, first_site: int, n: int, seed: int):
    x1 = np.empty((n, 4 * states), np.float32)
    x2 = np.empty((n, 4 * states), np.float32)
    _check(load().plf_generate_states_host(states, _ptr(x1), _ptr(x2), first_site, n, seed))
    return x1, x2


def states_kernel_info(states: int = STATES_PROTEIN, math_mode: int = MATH_STRICT, variant: int = 0, threads: int = 0):
    regs, thr, smem, tile = _i(0), _i(0), _sz(0), _i(0)
    _check(load().plf_states_kernel_info(states, math_mode, variant, threads, ctypes.byref(regs), ctypes.byref(thr),
                                         ctypes.byref(smem), ctypes.byref(tile)))
    return {"regs": regs.value, "threads": thr.value, "smem_bytes": smem.value, "tile_sites": tile.value}


def host_alloc(nbytes: int, dtype=np.uint8):
    """Pinned host memory as a numpy array (freed with host_free(arr))."""
    p = _vp()
    _check(load().plf_host_alloc(ctypes.byref(p), nbytes))
    buf = (ctypes.c_ubyte * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8).view(dtype)
    return arr, p.value


def host_free(ptr: int):
    _check(load().plf_host_free(ptr))


# ------------------------------------------------------------------------------------------
# Chained newview over a tree (plf_tree_* of include/b200plf.h)
# ------------------------------------------------------------------------------------------
def balanced_tree(n_tips: int):
    """(left, right) child arrays of a balanced binary tree over n_tips tips in post-order
    (level by level: pairs of the previous level's nodes)."""
    left, right, frontier, nxt = [], [], list(range(n_tips)), n_tips
    while len(frontier) > 1:
        new = []
        for a in range(0, len(frontier) - 1, 2):
            left.append(frontier[a])
            right.append(frontier[a + 1])
            new.append(nxt)
            nxt += 1
        if len(frontier) % 2:
            new.append(frontier[-1])
        frontier = new
    return np.asarray(left, np.int32), np.asarray(right, np.int32)


def random_tree(n_tips: int, seed: int = 0):
    """Random rooted binary tree by repeatedly joining two random live nodes (post-order)."""
    rng = np.random.RandomState(seed)
    live, left, right, nxt = list(range(n_tips)), [], [], n_tips
    while len(live) > 1:
        i, j = sorted(rng.choice(len(live), 2, replace=False))
        b = live.pop(j)
        a = live.pop(i)
        left.append(a)
        right.append(b)
        live.append(nxt)
        nxt += 1
    return np.asarray(left, np.int32), np.asarray(right, np.int32)


class Tree:
    """Post-order traversal of a rooted binary tree on one GPU: every inner node is one fused
    newview of its two children; all nodes of one level run in one launch."""

    def __init__(self, left, right, n_sites: int, device: int = 0, tip_codes: bool = False, states: int = STATES_DNA):
        self.lib = load()
        self.tip_codes = tip_codes
        self.states = states
        self.left = np.ascontiguousarray(left, np.int32)
        self.right = np.ascontiguousarray(right, np.int32)
        self.n_inner = self.left.size
        self.n_tips = self.n_inner + 1
        self.n_sites = n_sites
        self._t = _vp()
        rc = self.lib.plf_tree_create_states(ctypes.byref(self._t), device, self.n_tips, _ptr(self.left),
                                             _ptr(self.right), n_sites, 1 if tip_codes else 0, states)
        if rc != 0:
            raise PlfError(rc, self.lib.plf_tree_last_error(None).decode())

    def _ck(self, rc):
        if rc != 0:
            raise PlfError(rc, self.lib.plf_tree_last_error(self._t).decode())

    def close(self):
        if self._t:
            self.lib.plf_tree_destroy(self._t)
            self._t = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_math(self, mode: int):
        self._ck(self.lib.plf_tree_set_math(self._t, mode))

    def set_tuning(self, u: int = 0, chunk: int = 0):
        self._ck(self.lib.plf_tree_set_tuning(self._t, u, chunk))

    def tip_ptr(self, tip: int) -> int:
        p = _vp()
        self._ck(self.lib.plf_tree_tip_ptr(self._t, tip, ctypes.byref(p)))
        return p.value

    def write_tip(self, tip: int, clv, offset: int = 0):
        clv = np.ascontiguousarray(clv, np.float32)
        self._ck(self.lib.plf_tree_write_tip(self._t, tip, _ptr(clv), clv.nbytes, offset))
        self.wait()     # clv may be a temporary

    def write_tip_codes(self, tip: int, codes, first_site: int = 0):
        codes = np.ascontiguousarray(codes, np.uint8)
        self._ck(self.lib.plf_tree_write_tip_codes(self._t, tip, _ptr(codes), codes.size, first_site))
        self.wait()

    def write_tip_vector(self, tip_vector):
        tv = np.ascontiguousarray(tip_vector, np.float32).reshape(64)
        self._ck(self.lib.plf_tree_write_tip_vector(self._t, _ptr(tv)))

    def write_matrices(self, ev, p_left, p_right):
        S = self.states
        ev = np.ascontiguousarray(ev, np.float32).reshape(S * S)
        pl = np.ascontiguousarray(p_left, np.float32).reshape(self.n_inner, 4 * S * S)
        pr = np.ascontiguousarray(p_right, np.float32).reshape(self.n_inner, 4 * S * S)
        self._ck(self.lib.plf_tree_write_matrices(self._t, _ptr(ev), _ptr(pl), _ptr(pr)))

    def write_wgt(self, wgt):
        if wgt is None:
            self._ck(self.lib.plf_tree_write_wgt(self._t, None))
        else:
            wgt = np.ascontiguousarray(wgt, np.int32)
            assert wgt.size == self.n_sites
            self._ck(self.lib.plf_tree_write_wgt(self._t, _ptr(wgt)))

    def run_async(self):
        self._ck(self.lib.plf_tree_run_async(self._t))

    def wait(self):
        self._ck(self.lib.plf_tree_wait(self._t))

    def read_root(self, first_site: int = 0, n: int | None = None):
        n = self.n_sites - first_site if n is None else n
        clv = np.empty((n, 4 * self.states), np.float32)
        cnt = np.empty(n, np.int32)
        self._ck(self.lib.plf_tree_read_root(self._t, _ptr(clv), _ptr(cnt), first_site, n))
        return clv, cnt

    def total_scalings(self) -> int:
        v = ctypes.c_longlong(0)
        self._ck(self.lib.plf_tree_total_scalings(self._t, ctypes.byref(v)))
        return v.value

    def info(self):
        lv, sl, db, tb = _u(0), _u(0), _sz(0), _sz(0)
        self._ck(self.lib.plf_tree_info(self._t, ctypes.byref(lv), ctypes.byref(sl), ctypes.byref(db),
                                        ctypes.byref(tb)))
        return {"levels": lv.value, "clv_slots": sl.value, "device_bytes": db.value,
                "traversal_bytes": tb.value}

    def last_ms(self) -> float:
        v = ctypes.c_float(0)
        self._ck(self.lib.plf_tree_last_ms(self._t, ctypes.byref(v)))
        return v.value

    def evaluate_root(self, diag) -> float:
        """Log-likelihood of this rank's sites across the root branch (after run_async)."""
        diag = np.ascontiguousarray(diag, np.float32).reshape(4 * self.states)
        v = ctypes.c_double(0)
        self._ck(self.lib.plf_tree_evaluate_root(self._t, _ptr(diag), ctypes.byref(v)))
        return v.value


def evaluate_states_device(states: int, x1, x2, cnt1, cnt2, wgt, diag, n: int, lnl, stream: int = 0):
    """lnl (device double) += log-likelihood across a branch for S = 4 or 20 states; all pointers are device ints."""
    _check(load().plf_evaluate_states_device(states, _ptr(x1), _ptr(x2), _ptr(cnt1), _ptr(cnt2), _ptr(wgt), _ptr(diag), n,
                                             _ptr(lnl), _ptr(stream)))


def evaluate_device(x1, x2, cnt1, cnt2, wgt, diag, n: int, lnl, stream: int = 0):
    """Root log-likelihood on caller-owned DEVICE memory; pointers are ints; ADDS to the double *lnl."""
    _check(load().plf_evaluate_device(x1, x2, cnt1, cnt2, wgt, diag, n, lnl, stream or None))
