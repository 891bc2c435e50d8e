"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/b200plf.h declares, fails loudly without a GPU (no CPU fallback), and the host-side
partition math follows the reference test bench (app/src/include.h:150-268).  No compute calls."""
from __future__ import annotations

import math
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib(pkg):
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    return pkg.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "b200plf.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg, lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    assert set(syms) == set(pkg.PROTOTYPES), "header and Python prototypes disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True,
                        text=True, check=True).stdout
    exported = set(re.findall(r" T (plf_[a-z0-9_]+)", nm))
    assert set(syms) <= exported
    for s in syms:
        assert hasattr(lib, s)


def test_c_abi_has_no_cxx_or_torch_types(pkg):
    ldd = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "b200plf.h"\nint main(void){plf_launch_opts o; (void)o; return PLF_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only",
                    f"-I{os.path.join(ROOT, 'include')}", str(c)], check=True)


def test_no_cpu_fallback_without_gpu(pkg, lib):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.PlfError) as e:
        pkg.Context(0, 1)
    assert e.value.code in (-2, -1)
    with pytest.raises(pkg.PlfError):
        pkg.newview_device(16, 16, 16, None, 16, 16, 16, None, 8, None)


def test_error_reporting_for_bad_arguments(pkg, lib):
    with pytest.raises(pkg.PlfError):
        pkg.Context(0, 0)                      # n_instances == 0
    with pytest.raises(pkg.PlfError):
        pkg.Context(0, 1, layout=7)
    msg = lib.plf_last_error(None)
    assert msg and b"layout" in msg


def test_gen_pattern_is_the_movers_constant_site(pkg, lib):
    x1, x2, ev4, pl, pr = pkg.gen_pattern()
    # mm2sleft_genDNAwindowComb.cpp:44-49 / mm2sright_genDNAwindowComb.cpp:45-50
    assert np.allclose(x1[:4], [0.2135, 0.1427, 0.4139, 0.8301]) and np.isclose(x1[15], 0.2024)
    assert np.allclose(x2[:4], [0.123456, 0.234567, 0.345678, 0.789543]) and np.isclose(x2[15], 0.06789)
    ev4 = ev4.reshape(4, 4, 4)
    pl = pl.reshape(4, 4, 4)
    pr = pr.reshape(4, 4, 4)
    for j in range(4):
        # lane j sees its own slice of the site as both EV halves and as P^T rows
        assert np.array_equal(ev4[j, 0], x1[4 * j:4 * j + 4]) and np.array_equal(ev4[j, 1], x1[4 * j:4 * j + 4])
        assert np.array_equal(ev4[j, 2], x2[4 * j:4 * j + 4]) and np.array_equal(ev4[j, 3], x2[4 * j:4 * j + 4])
        for k in range(4):
            assert np.all(pl[j, k] == x1[4 * j + k]) and np.all(pr[j, k] == x2[4 * j + k])


def test_generate_host_distribution(pkg, lib):
    x1, x2 = pkg.generate_host(0, 4096, seed=42)
    assert x1.shape == (4096, 16)
    assert (x2 > 0).all() and (x2 < 1).all()
    big = x1[np.arange(4096) % 4 != 0]
    small = x1[np.arange(4096) % 4 == 0]
    assert (big > 0).all() and (big < 1).all() and abs(big.mean() - 0.5) < 0.01
    assert (small < 1.0001e-12).all() and (small > 0).all()      # host_mem.cpp:200-202
    # counter based: any sub-range reproduces the same values
    y1, y2 = pkg.generate_host(1000, 10, seed=42)
    assert np.array_equal(y1, x1[1000:1010]) and np.array_equal(y2, x2[1000:1010])
    z1, _ = pkg.generate_host(0, 16, seed=43)
    assert not np.array_equal(z1, x1[:16])


@pytest.mark.parametrize("n,inst", [(100, 1), (100, 9), (1000, 8), (1 << 20, 9), (7, 7), (64 << 20, 8)])
def test_testbench_partition_follows_reference_rule(pkg, n, inst):
    tb = pkg.TestbenchInfo(n, inst)
    per = math.ceil(n / inst)                       # include.h:184-186
    assert tb.alignments_per_instance() == per
    assert tb.alignments_padding() == per * inst - n
    sizes = [tb.alignments_per_instance(k) for k in range(inst)]
    assert sizes[:-1] == [per] * (inst - 1) and sum(sizes) == n
    assert [tb.instance_offset(k) for k in range(inst)] == [k * per for k in range(inst)]
    assert tb.instance_active_elements_left(0) == per * 16 + 80      # include.h:207-209
    assert pkg.TestbenchInfo(n, inst, layout=pkg.LAYOUT_SEP).instance_active_elements_right(0) == per * 16 + 64
    assert tb.data_size() == n * 64                 # 64-bit: 64 Mi sites is 4 GiB (no wrap)
    assert [c for _, c in pkg.partition_sites(n, inst)] == sizes


def test_partition_rejects_empty_last_instance(pkg):
    assert not pkg.TestbenchInfo(10, 8).valid()     # ceil(10/8)=2 -> last instance would get -4
    assert pkg.TestbenchInfo(16, 8).valid()
    assert pkg.partition_sites(10, 8) == [(0, 2), (2, 2), (4, 2), (6, 2), (8, 2), (10, 0), (10, 0), (10, 0)]


def test_pack_matches_oracle_packing(pkg):
    import oracle
    ev, left, right, x1, x2, _ = oracle.host_mem_inputs(33, seed=9)
    for layout in (0, 1):
        lb, rb = oracle.pack_buffers(ev, left, right, x1, x2, layout)
        assert np.array_equal(pkg.pack_left(ev, left, x1), lb)
        assert np.array_equal(pkg.pack_right(ev, right, x2, layout), rb)
