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


@pytest.mark.gpu
@pytest.mark.parametrize("sites,calls,chunk", [(100, 1, 0), (100003, 2, 16384)])
def test_host_stream_twenty_states(hosts, sites, calls, chunk):
    """The same streamed round trip with an AA configuration name (STATES knob): 80-float sites, exact against the
    host's golden loop nest."""
    cfg = "plf_128x9AAwindow8192Comb_memAAwindowComb"
    r = run(HOST_STREAM, *([cfg, 0, sites, calls] + ([chunk] if chunk else [])))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Test result: Passed" in r.stdout and "20 states x 4 rate categories, 80 floats per site" in r.stdout
    assert f"scalerIncrement (last call): {(sites + 3) // 4}" in r.stdout


def test_host_stream_argument_errors(hosts):
    assert run(HOST_STREAM).returncode == 2
    r = run(HOST_STREAM, CFG_GEN, 0, 100, 1)
    assert r.returncode == 2 and "INPUT_SRC=mem" in r.stderr


# ---- the packed buffers as files (host/plfb_file.h) -------------------------------------------
def test_plfb_python_round_trip_and_validation(pkg, tmp_path):
    rng = np.random.RandomState(1)
    n = 37
    left = rng.random_sample(80 + 16 * n).astype(np.float32)
    right_sep = rng.random_sample(64 + 16 * n).astype(np.float32)
    sc = rng.randint(0, 2, n).astype(np.uint8)
    pkg.save_plfb(str(tmp_path / "l.plfb"), pkg.PLFB_LEFT, pkg.LAYOUT_COMB, n, left)
    pkg.save_plfb(str(tmp_path / "r.plfb"), pkg.PLFB_RIGHT, pkg.LAYOUT_SEP, n, right_sep)
    pkg.save_plfb(str(tmp_path / "s.plfb"), pkg.PLFB_SCALER, pkg.LAYOUT_COMB, n, sc)
    got = pkg.load_plfb(str(tmp_path / "l.plfb"))
    assert (got["kind"], got["layout"], got["sites"]) == (pkg.PLFB_LEFT, pkg.LAYOUT_COMB, n) and np.array_equal(got["data"], left)
    got = pkg.load_plfb(str(tmp_path / "r.plfb"))
    assert got["layout"] == pkg.LAYOUT_SEP and np.array_equal(got["data"], right_sep)
    assert np.array_equal(pkg.load_plfb(str(tmp_path / "s.plfb"))["data"], sc)
    assert os.path.getsize(tmp_path / "l.plfb") == 64 + left.nbytes
    with pytest.raises(ValueError):                                   # Comb right buffer is 16 floats longer
        pkg.save_plfb(str(tmp_path / "x.plfb"), pkg.PLFB_RIGHT, pkg.LAYOUT_COMB, n, right_sep)
    raw = (tmp_path / "l.plfb").read_bytes()
    (tmp_path / "trunc.plfb").write_bytes(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        pkg.load_plfb(str(tmp_path / "trunc.plfb"))
    (tmp_path / "bad.plfb").write_bytes(b"NOPE" + raw[4:])
    with pytest.raises(ValueError, match="not a PLFB"):
        pkg.load_plfb(str(tmp_path / "bad.plfb"))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,layout", [(CFG_COMB, 0), (CFG_SEP, 1)])
def test_host_mem_dumps_buffers_the_oracle_reproduces(hosts, pkg, coracle, tmp_path, cfg, layout):
    """PLF_DUMP_DIR: the files hold exactly the packed inputs and the outputs; the oracle's packed front end maps
    one onto the other bit for bit."""
    n = 4099
    r = run(hosts[0], cfg, 0, n, 1, 3, env={"PLF_DUMP_DIR": str(tmp_path)})
    assert r.returncode == 0 and "Buffers written" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    f = {k: pkg.load_plfb(str(tmp_path / f"{k}.plfb")) for k in ("left", "right", "out", "scaler")}
    assert all(v["sites"] == n and v["layout"] == layout for v in f.values())
    o3, osc, oinc = coracle.newview_packed(f["left"]["data"], f["right"]["data"], layout, n)
    assert np.array_equal(bits(f["out"]["data"].reshape(n, 16)), bits(o3))
    assert np.array_equal(f["scaler"]["data"], osc) and oinc == (n + 3) // 4


@pytest.mark.gpu
def test_host_mem_loads_stimulus_from_files(hosts, pkg, coracle, tmp_path):
    """PLF_LOAD_DIR: a stimulus written by the Python package (signed values, no designed scaler pattern) runs
    through the host, passes its exact verification, and the dumped output equals the oracle's."""
    n = 3001
    rng = np.random.RandomState(8)
    mag = np.repeat(10.0 ** rng.uniform(-9, 0, n), 16)                 # per-site magnitudes: some sites rescale, some do not
    left = np.concatenate([rng.standard_normal(80), rng.standard_normal(16 * n) * mag]).astype(np.float32)
    right = np.concatenate([left[:16], rng.standard_normal(64), rng.standard_normal(16 * n) * mag]).astype(np.float32)
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    dst.mkdir()
    pkg.save_plfb(str(src / "left.plfb"), pkg.PLFB_LEFT, pkg.LAYOUT_COMB, n, left)
    pkg.save_plfb(str(src / "right.plfb"), pkg.PLFB_RIGHT, pkg.LAYOUT_COMB, n, right)
    r = run(hosts[0], CFG_COMB, 0, n, 1, 2, env={"PLF_LOAD_DIR": str(src), "PLF_DUMP_DIR": str(dst)})
    assert r.returncode == 0 and "Test result: Passed" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    o3, osc, oinc = coracle.newview_packed(left, right, 0, n)
    assert 0 < oinc < n
    assert np.array_equal(bits(pkg.load_plfb(str(dst / "out.plfb"))["data"].reshape(n, 16)), bits(o3))
    assert np.array_equal(pkg.load_plfb(str(dst / "scaler.plfb"))["data"], osc)
    assert f"scalerIncrement (call 0): {oinc}" in r.stdout
    # a file with the wrong site count is refused
    r = run(hosts[0], CFG_COMB, 0, n + 1, 1, 2, env={"PLF_LOAD_DIR": str(src)})
    assert r.returncode == 2 and "sites" in r.stderr


# ---- STATES knob at the host level (host/host_states.cpp) --------------------------------------
HOST_STATES = os.path.join(PKG_DIR, "host_states.exe")
CFG_AA = "plf_128x9AAwindow8192Comb_memAAwindowComb"


def test_host_golden_states_equals_oracle(tmp_path, coracle):
    """host/golden_plf.cpp with the state count as a parameter == the general-S oracle (S = 4 and 20), and at
    S = 4 == the DNA golden of the same file."""
    shim = tmp_path / "shim.cpp"
    shim.write_text('#include "golden_plf.h"\nextern "C" long long g(unsigned S,const float*a,const float*b,float*c,'
                    'const float*e,unsigned long n,const float*l,const float*r,const int*w,unsigned char*s){long long i=0;'
                    'plfhost::golden_plf_states(S,a,b,c,e,n,l,r,w,i,s);return i;}\n')
    so = tmp_path / "libgolden_states.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                    f"-I{os.path.join(PKG_DIR, 'host')}", "-o", str(so), str(shim),
                    os.path.join(PKG_DIR, "host", "golden_plf.cpp")], check=True)
    lib = ctypes.CDLL(str(so))
    lib.g.restype = ctypes.c_longlong
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for S, n in ((4, 5000), (20, 700)):
        rng = np.random.RandomState(S)
        x1 = (rng.standard_normal((n, 4 * S)) * 10.0 ** rng.uniform(-13, 1, (n, 1))).astype(np.float32)
        x2 = (rng.standard_normal((n, 4 * S)) * 10.0 ** rng.uniform(-3, 1, (n, 1))).astype(np.float32)
        ev, left, right = (rng.standard_normal(k).astype(np.float32) for k in (S * S, 4 * S * S, 4 * S * S))
        wgt = rng.randint(0, 9, n).astype(np.int32)
        out = np.empty((n, 4 * S), np.float32)
        sc = np.empty(n, np.uint8)
        inc = lib.g(S, p(x1), p(x2), p(out), p(ev), ctypes.c_ulong(n), p(left), p(right), p(wgt), p(sc))
        o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right, wgt)
        assert inc == oinc and 0 < oinc and np.array_equal(sc, osc) and np.array_equal(bits(out), bits(o3)), S
        if S == 4:
            d3, dsc, dinc = coracle.newview(x1, x2, ev, left, right, wgt)
            assert np.array_equal(bits(out), bits(d3)) and inc == dinc


def test_host_states_argument_errors(hosts):
    assert run(HOST_STATES).returncode == 2
    r = run(HOST_STATES, "plf_128x9RNAwindow8192Comb_memRNAwindowComb", 0, 100, 1)
    assert r.returncode == 2 and "DNA or AA" in r.stderr
    r = run(HOST_STATES, "plf_128x9AAwindow8192Comb_genAAwindowComb", 0, 100, 1)
    assert r.returncode == 2 and "INPUT_SRC=mem" in r.stderr
    r = run(hosts[1], CFG_AA, 0, 100, 1, 1)                         # the gen test bench is DNA only and says so
    assert r.returncode == 2 and ("host_states" in r.stderr or "Usage" in r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,sites,calls,instances", [
    (CFG_AA, 100, 1, 1), (CFG_AA, 200003, 2, 9), ("plf_128x4AAwindow8192Sep_memAAwindowSep", 50001, 1, 3)])
def test_host_mem_runs_the_AA_configuration(hosts, cfg, sites, calls, instances):
    """STATES=AA through the main drop-in host: the same five arguments, the instance API of a 20-state context
    (packed [EV400|P1600|CLV] buffers), exact verification against the host's S-state golden."""
    r = run(hosts[0], cfg, 0, sites, calls, instances)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2000:]
    assert "Test result: Passed" in r.stdout and f"scalerIncrement (call 0): {(sites + 3) // 4}" in r.stdout
    assert "AA (20 states x 4 rate categories, 80 floats per site)" in r.stdout and "961 B/site" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,sites,calls", [(CFG_AA, 100, 1), (CFG_AA, 200003, 2), (CFG_COMB, 100000, 2)])
def test_host_states_end_to_end(hosts, cfg, sites, calls):
    r = run(HOST_STATES, cfg, 0, sites, calls)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Test result: Passed (exact compare)" in r.stdout
    assert f"scalerIncrement (last call): {(sites + 3) // 4}" in r.stdout
    assert "G sites/s" in r.stdout


@pytest.mark.gpu
def test_host_states_fma_mode(hosts):
    r = run(HOST_STATES, CFG_AA, 0, 50000, 1, env={"PLF_MATH": "fma"})
    assert r.returncode == 0 and "Test result: Passed (1e-5 relative)" in r.stdout, r.stdout[-1500:]


@pytest.mark.gpu
def test_device_memory_helpers(pkg):
    lib = pkg.load()
    p = ctypes.c_void_p()
    assert lib.plf_device_malloc(0, ctypes.byref(p), 4096) == 0 and p.value
    src = np.arange(1024, dtype=np.float32)
    dst = np.zeros(1024, np.float32)
    assert lib.plf_memcpy_h2d(p, src.ctypes.data_as(ctypes.c_void_p), 4096, None) == 0
    assert lib.plf_memcpy_d2h(dst.ctypes.data_as(ctypes.c_void_p), p, 4096, None) == 0
    assert lib.plf_stream_sync(None) == 0 and np.array_equal(src, dst)
    assert lib.plf_memset_device(p, 0, 4096, None) == 0
    assert lib.plf_memcpy_d2h(dst.ctypes.data_as(ctypes.c_void_p), p, 4096, None) == 0
    assert lib.plf_stream_sync(None) == 0 and not dst.any()
    assert lib.plf_device_free(p) == 0
    assert lib.plf_memcpy_h2d(None, src.ctypes.data_as(ctypes.c_void_p), 16, None) != 0
    assert lib.plf_device_malloc(99, ctypes.byref(p), 16) != 0
