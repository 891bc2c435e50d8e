"""The C++ host drop-ins (host/host_mem.cpp, host/host_gen.cpp): argument handling and the golden
model on CPU; full runs (configure -> write -> run -> read -> verify) on the GPU box."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import PKG_DIR, ROOT, bits

HOST_MEM = os.path.join(PKG_DIR, "host_mem.exe")
HOST_GEN = os.path.join(PKG_DIR, "host_gen.exe")
CFG_COMB = "plf_128x9DNAwindow8192Comb_memDNAwindowComb"      # the reference's default artefact name
CFG_SEP = "plf_128x4DNAwindow8192Sep_memDNAwindowSep"
CFG_GEN = "plf_128x9DNAwindow8192Comb_genDNAwindowComb"


@pytest.fixture(scope="module")
def hosts():
    if not (os.path.exists(HOST_MEM) and os.path.exists(HOST_GEN)):
        subprocess.run(["make", "-C", ROOT, "lib", "host"], check=True, stdout=subprocess.DEVNULL)
    return HOST_MEM, HOST_GEN


def run(exe, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600, env=e)


def test_host_golden_model_equals_oracle(tmp_path, coracle):
    """host/golden_plf.cpp (verification phase only) is bit-identical to the pinned oracle."""
    shim = tmp_path / "shim.cpp"
    shim.write_text('#include "golden_plf.h"\nextern "C" long long g(const float*a,const float*b,float*c,const float*e,'
                    'unsigned long n,const float*l,const float*r,const int*w,unsigned char*s){long long i=0;'
                    'plfhost::golden_plf(a,b,c,e,n,l,r,w,i,s);return i;}\n')
    so = tmp_path / "libgolden.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                    f"-I{os.path.join(PKG_DIR, 'host')}", "-o", str(so), str(shim),
                    os.path.join(PKG_DIR, "host", "golden_plf.cpp")], check=True)
    lib = ctypes.CDLL(str(so))
    lib.g.restype = ctypes.c_longlong
    rng = np.random.RandomState(3)
    n = 20000
    x1 = (rng.standard_normal((n, 16)) * 10.0 ** rng.uniform(-12, 1, (n, 1))).astype(np.float32)
    x2 = (rng.standard_normal((n, 16)) * 10.0 ** rng.uniform(-12, 1, (n, 1))).astype(np.float32)
    x1[5] = -0.0
    x1[6, 3] = np.nan
    ev, left, right = (rng.standard_normal(k).astype(np.float32) for k in (16, 64, 64))
    wgt = rng.randint(0, 9, n).astype(np.int32)
    out = np.empty((n, 16), np.float32)
    sc = np.empty(n, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    inc = lib.g(p(x1), p(x2), p(out), p(ev), ctypes.c_ulong(n), p(left), p(right), p(wgt), p(sc))
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    assert inc == oinc and np.array_equal(sc, osc) and np.array_equal(bits(out), bits(o3))


def test_host_argument_errors_are_fatal(hosts):
    mem, gen = hosts
    assert run(mem).returncode == 2                                          # wrong argc
    r = run(mem, "plf_128x9DNAwindow8192Comb_fooDNAwindowComb", 0, 100, 1, 1)
    assert r.returncode == 2 and "neither mem nor gen" in r.stderr
    r = run(mem, CFG_GEN, 0, 100, 1, 1)
    assert r.returncode == 2 and "host_gen" in r.stderr
    r = run(gen, CFG_COMB, 0, 100, 1, 1)
    assert r.returncode == 2 and "host_mem" in r.stderr


def test_host_without_gpu_fails_loudly(hosts, pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    r = run(hosts[0], CFG_COMB, 0, 100, 1, 1)
    assert r.returncode == 2 and "cuda" in r.stderr.lower()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,sites,calls,inst", [(CFG_COMB, 100, 1, 1),        # BASELINE configs[0]
                                                   (CFG_COMB, 100000, 3, 9),
                                                   (CFG_SEP, 4099, 2, 4),
                                                   (CFG_COMB, 1000000, 1, 1)])  # BASELINE configs[1]
def test_host_mem_end_to_end(hosts, cfg, sites, calls, inst):
    r = run(hosts[0], cfg, 0, sites, calls, inst)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Test result: Passed" in r.stdout
    assert f"scalerIncrement (call 0): {(sites + 3) // 4}" in r.stdout        # every 4th site rescales
    assert "Speed up (excluding pcie transfer)" in r.stdout


@pytest.mark.gpu
def test_host_mem_rejects_bad_runs(hosts):
    mem = hosts[0]
    assert run(mem, CFG_COMB, 0, 10, 1, 8).returncode == 2       # last instance would be empty
    assert run(mem, CFG_COMB, 0, 100, 1, 10).returncode == 2     # more instances than NUM_ACCELERATORS
    assert run(mem, CFG_COMB, 0, "abc", 1, 1).returncode == 2
    assert run(mem, CFG_COMB, 99, 100, 1, 1).returncode == 2     # no such device


@pytest.mark.gpu
def test_host_mem_fma_mode_reports_mismatch_as_failure_or_pass(hosts):
    """FMA arithmetic is not bit-identical: the exact-compare verification of the host must notice."""
    r = run(hosts[0], CFG_COMB, 0, 50000, 1, 2, env={"PLF_MATH": "fma"})
    assert r.returncode in (0, 1)
    if r.returncode == 1:
        assert "ERROR: alignment data wrong" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("sink", ["write", "discard"])
def test_host_gen_runs(hosts, sink):
    r = run(hosts[1], CFG_GEN, 0, 1 << 20, 3, 4, env={"PLF_GEN_SINK": sink})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "G sites/s" in r.stdout and "scalerIncrement (instance 0, last call): 0" in r.stdout


HOST_STREAM = os.path.join(PKG_DIR, "host_stream.exe")


@pytest.mark.gpu
@pytest.mark.parametrize("sites,calls,chunk", [(100, 1, 0), (300001, 2, 65536), (1 << 20, 2, 0), (5000000, 1, 1 << 20)])
def test_host_stream_end_to_end(hosts, sites, calls, chunk):
    """The streamed round trip: chunked, triple-buffered H2D / kernel / D2H over unpacked host
    arrays, verified exactly against the host's CPU golden."""
    args = [CFG_COMB, 0, sites, calls] + ([chunk] if chunk else [])
    r = run(HOST_STREAM, *args)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Test result: Passed" in r.stdout
    assert f"scalerIncrement (last call): {(sites + 3) // 4}" in r.stdout


def test_host_stream_argument_errors(hosts):
    assert run(HOST_STREAM).returncode == 2
    r = run(HOST_STREAM, CFG_GEN, 0, 100, 1)
    assert r.returncode == 2 and "INPUT_SRC=mem" in r.stderr
