#!/usr/bin/env python
"""bench.py -- PLF newview throughput on B200 (sites/s, HBM GB/s), next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            # our arm  (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU plf()

Workload (config.workload): BASELINE.json configs[2] -- "DNA with 4 rate categories, 64M sites,
site-partitioned across 1/2/4/8 B200": 64 Mi sites TOTAL, split over the ranks with the reference's
instance rule (ceil(n/N), last takes the remainder), so scaling is "strong".  A step is one pass
of the fused newview kernel over the rank's whole site range (one launch).  Inputs (12 GiB/N per
rank) are far larger than L2, so no flush is needed between steps.

value      device-resident throughput: CUDA events on the launching stream, max over ranks.
e2e        the same sites through the reference-facing host API (plf_write_left/right -> plf_run_async
           -> plf_read_out/plf_read_scaler on NUM_ACCELERATORS=9 instances) from PINNED HOST buffers,
           H2D and D2H copies inside the timed region.
roofline   193 algorithmic bytes/site x sites per launch / mean launch time, against the measured
           HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline  oracle/_ref (the reference's plf.cpp compiled in place) on all host cores.

side       the other BASELINE.json configurations and the rows SURVEY.md section 8f adds, measured in the same
           run after the headline (about a minute in total; --no-side skips them), each with its own roofline:
           cfg2 (1 Mi sites, one instance: cold = rotating buffer sets, warm = one set, K launches between two
           events), cfg2_instances (the same 1 Mi sites split over NUM_ACCELERATORS=9 instance streams), cfg4a / cfg4b
           (INPUT_SRC=gen, sink write / discard), cfg5 (1024-taxon tree, 1 Mi sites split over the ranks, dense tips
           and state-code tips), protein (20 states, strict and FMA) and evaluate (root log-likelihood kernel).

Stand-alone side runs: --workload cfg2 (--buffer-sets 1 for the "same buffers" variant) and --workload protein.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG_DIR = os.path.join(ROOT, "amd-versal-phylogenetic-likelihood-function_b200")

TOTAL_SITES = {"cfg3": 64 << 20, "cfg2": 1 << 20}
WORKLOAD_NAME = {
    "cfg3": "BASELINE.json configs[2]: DNA, 4 rate categories, 64Mi sites, site-partitioned over N GPUs",
    "cfg2": "BASELINE.json configs[1]: single PLF instance, DNA, 1Mi sites, 1 GPU",
}
BYTES_PER_SITE = 193
NOMINAL_HBM_GBS = 8000.0
FALLBACK_HBM_GBS = 6650.0
SEED = 42


def buffer_sets_for(n_per_gpu: int, requested: int = 0) -> int:
    """cfg2's 192 MiB working set is only ~1.5x L2: rotate buffer sets so every step reads cold data."""
    return requested if requested > 0 else (1 if n_per_gpu * 128 >= (1 << 30) else 6)


def workload_config(workload: str, total_sites: int, world: int, math: str, buffer_sets: int = 0) -> dict:
    """`config` of the JSON line -- the SAME dict for our arm and for the reference arm (the driver compares them)."""
    n_max = -(-total_sites // world)
    sets = buffer_sets_for(n_max, buffer_sets)
    return {"workload": WORKLOAD_NAME[workload], "total_sites": total_sites, "sites_per_step": total_sites,
            "n_gpus": world, "partition": "contiguous site ranges, ceil(n/N) rule (reference include.h:181-192)",
            "math": math, "bytes_per_site": BYTES_PER_SITE,
            "l2": f"inputs {2 * n_max * 64 * sets >> 20} MiB per GPU in {sets} rotating buffer set(s), far larger than "
                  "the 126 MB L2: no flush needed between steps"}


def load_pkg():
    name = "plf_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def stimulus_matrices(seed: int):
    """EV[16], P_left[64], P_right[64]: uniform(0,1) doubles cast to float, branch matrices drawn
    interleaved -- the recipe of app/src/host_mem.cpp:183-197 of the reference, seeded."""
    rng = np.random.RandomState(seed)
    ev = rng.random_sample(16).astype(np.float32)
    br = rng.random_sample(128)
    return ev, br[0::2].astype(np.float32), br[1::2].astype(np.float32)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU plf() (oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------
def host_inputs_tiled(n_sites: int):
    """host_mem.cpp:179-209 stimulus for a 1 Mi-site block, tiled to n_sites (content repeats;
    the CPU path's cost does not depend on it)."""
    import oracle
    block = min(n_sites, 1 << 20)
    ev, left, right, b1, b2, _ = oracle.host_mem_inputs(block, seed=SEED)
    x1 = np.empty((n_sites, 16), np.float32)
    x2 = np.empty((n_sites, 16), np.float32)
    for lo in range(0, n_sites, block):
        c = min(block, n_sites - lo)
        x1[lo:lo + c] = b1[:c]
        x2[lo:lo + c] = b2[:c]
    return ev, left, right, x1, x2, np.ones(n_sites, np.int32)


def cpu_reference_rate(n_sites: int, repeats: int, threads: int, lib: str | None = None):
    """Best-of-`repeats` sites/s of the reference plf() over n_sites with `threads` threads."""
    import oracle
    if lib is not None:
        ref, kind = oracle.RefOracle(lib), "reference"
    elif oracle.RefOracle.available():
        ref, kind = oracle.RefOracle(), "reference"
    else:
        ref, kind = None, "port"
        co = oracle.COracle()
    ev, left, right, x1, x2, wgt = host_inputs_tiled(n_sites)
    out = np.empty((n_sites, 16), np.float32)
    times, inc = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        if ref is not None:
            _, inc = ref.newview(x1, x2, ev, left, right, wgt, nthreads=threads, out=out)
        else:
            _, _, inc = co.newview(x1, x2, ev, left, right, wgt, nthreads=threads)
        times.append(time.perf_counter() - t0)
    assert inc == (n_sites + 3) // 4, "CPU reference produced an unexpected scaler increment"
    return kind, times


def run_reference_arm(args):
    """The reference's own CPU plf() over the WHOLE workload per step (same work, same config as our arm), every
    host thread running the unmodified function on a contiguous site range.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = args.sites or TOTAL_SITES[args.workload]
    import oracle
    ref = oracle.RefOracle() if oracle.RefOracle.available() else None
    kind = "reference" if ref is not None else "port"
    co = None if ref is not None else oracle.COracle()
    ev, left, right, x1, x2, wgt = host_inputs_tiled(total)
    out = np.empty((total, 16), np.float32)

    def one_step(m):
        if ref is not None:
            return ref.newview(x1[:m], x2[:m], ev, left, right, wgt[:m], nthreads=cores, out=out)[1]
        return co.newview(x1[:m], x2[:m], ev, left, right, wgt[:m], nthreads=cores)[2]

    for _ in range(args.warmup):
        one_step(min(total, 2 << 20))           # warm-up on a prefix: page-in, thread start-up
    t0 = time.perf_counter()
    for _ in range(args.steps):
        inc = one_step(total)
    dt = time.perf_counter() - t0
    assert inc == (total + 3) // 4, "CPU reference produced an unexpected scaler increment"
    value = total * args.steps / dt
    desc = (f"the whole {total}-site workload per step (host_mem.cpp stimulus, 1Mi-site block tiled), {cores} threads, "
            f"{'/root/reference/app/src/plf.cpp compiled in place (-O2 -ffp-contract=off)' if kind == 'reference' else 'oracle/plf_oracle.c port'}")
    line = {
        "impl": "reference", "metric": "plf_sites_per_s", "value": value, "unit": "sites/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, total, args.gpus, args.math, args.buffer_sets),
        "cpu_baseline": {"value": value, "unit": "sites/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hbm_equiv_gbs": value * BYTES_PER_SITE / 1e9,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    pkg = load_pkg()
    from plf_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PLF path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    total_sites = args.sites or TOTAL_SITES[args.workload]
    first, n = sharding.shard_for_rank(total_sites, rank, world)
    K, W = args.steps, max(args.warmup, 3)
    math_mode = pkg.MATH_FMA if args.math == "fma" else pkg.MATH_STRICT
    opts = pkg.make_opts(math_mode, args.variant, args.threads, args.blocks_per_sm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ("value") ------------------------------------------------------
    ev, left, right = stimulus_matrices(SEED)
    d_ev = torch.from_numpy(ev).to(device)
    d_pl = torch.from_numpy(left).to(device)
    d_pr = torch.from_numpy(right).to(device)
    sets = buffer_sets_for(-(-total_sites // world), args.buffer_sets)
    x1 = [torch.empty((n, 16), device=device) for _ in range(sets)]
    x2 = [torch.empty((n, 16), device=device) for _ in range(sets)]
    x3 = [torch.empty((n, 16), device=device) for _ in range(sets)]
    sc = [torch.empty(n, dtype=torch.uint8, device=device) for _ in range(sets)]
    sums = torch.zeros(K + W, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream().cuda_stream
    for s in range(sets):
        pkg.generate_device(x1[s].data_ptr(), x2[s].data_ptr(), first, n, SEED, stream)

    def step(i):
        s = i % sets
        pkg.newview_device(x1[s].data_ptr(), x2[s].data_ptr(), x3[s].data_ptr(), sc[s].data_ptr(),
                           d_ev.data_ptr(), d_pl.data_ptr(), d_pr.data_ptr(), None, n,
                           sums[i:].data_ptr(), opts, stream)

    for i in range(W):
        step(i)
    barrier()
    from tools.clocks import ClockSampler as NvmlSampler
    sampler = NvmlSampler(local_rank)
    if rank == 0:
        sampler.start()
        for i in range(W):          # keep the GPU under the same load while the sampler spins up
            step(i)
        barrier()
    else:
        barrier()
    launches0 = pkg.launch_count()
    # Two events around the K launches and nothing between them: an event record between two kernels is a stream
    # operation of its own and would break the programmatic-dependent-launch edge from one launch to the next.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record()
    for i in range(K):
        step(W + i)
    ev1.record()
    barrier()
    sampler.end()
    launches = pkg.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    per_launch_ms = [elapsed_ms / K]
    t_max_ms = sharding.max_over_ranks(elapsed_ms, device)
    launches_all = sharding.reduce_scaler_increment(launches, device)

    # correctness of what was timed: every step produced the designed scaler count, the
    # NCCL-reduced total matches, and a slice is bit-identical to the oracle.
    sums_h = sums.cpu().numpy()
    lo_scaled = (first + 3) // 4
    expect = (first + n + 3) // 4 - lo_scaled
    assert (sums_h[W:] == expect).all(), f"rank {rank}: scaler sums {sums_h[W:W + 4]} != {expect}"
    total_inc = sharding.reduce_scaler_increment(int(sums_h[-1]), device)
    assert total_inc == (total_sites + 3) // 4
    # (parity against the oracle is the job of tests/ and smoke(); here only self-consistency of the timed
    # work: the designed scaler pattern in the bytes, finite outputs, and rescaled sites back above 2^-32)
    last = (W + K - 1) % sets
    scb = sc[last][: min(n, 1 << 16)].cpu().numpy()
    want = ((np.arange(first, first + scb.size) % 4) == 0).astype(np.uint8)
    assert np.array_equal(scb, want), "scaler bytes do not follow the stimulus design"
    g3 = x3[last][: min(n, 4096)].cpu().numpy()
    assert np.isfinite(g3).all() and (np.abs(g3).max(axis=1) >= 2.0 ** -32).all(), "implausible CLV output"

    value = total_sites * K / (t_max_ms * 1e-3)
    mean_launch_ms = statistics.fmean(per_launch_ms)
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_SITE * n / (mean_launch_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    info = pkg.kernel_info(args.variant, math_mode, args.threads)

    # ---- end-to-end leg through the host API with pinned host buffers -------------------------
    del x1, x2, x3, sc
    torch.cuda.empty_cache()
    e2e = run_e2e(pkg, torch, args, local_rank, first, n, total_sites, world, math_mode, barrier,
                  sharding, device, ev, left, right)

    # ---- side workloads: the other BASELINE configs and the section-8f rows, each with its own roofline ----
    side = None
    if not args.no_side and args.workload == "cfg3" and (not args.sites or args.side_small):
        side = run_side(pkg, torch, args, rank, world, local_rank, device, math_mode, barrier, sharding, peak)

    # ---- CPU baseline (rank 0; the other ranks wait at the barrier below) --------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(total_sites, 16 << 20)
        kind, times = cpu_reference_rate(sample, 3, cores)
        _, t1 = cpu_reference_rate(1 << 20, 2, 1)
        cpu = {"value": sample / min(times), "unit": "sites/s", "cores": cores, "kind": kind,
               "sample": f"{sample} sites of the workload's stimulus recipe, best of 3, {cores} threads "
                         "each running the unmodified plf() on a contiguous site range "
                         "(-O2 -ffp-contract=off)",
               "single_thread_sites_per_s": (1 << 20) / min(t1)}
        import oracle
        if os.path.exists(oracle.LIB_REF_O0):      # the reference's own host flags: -g, no optimisation
            _, t0 = cpu_reference_rate(1 << 20, 2, 1, oracle.LIB_REF_O0)
            cpu["single_thread_sites_per_s_reference_flags_g_O0"] = (1 << 20) / min(t0)

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": "plf_sites_per_s", "value": value, "unit": "sites/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": t_max_ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, total_sites, world, args.math, args.buffer_sets),
            "detail": {"sites_per_gpu": n, "kernel": info | {"variant": args.variant}},
            "hbm_gbs": value * BYTES_PER_SITE / 1e9,
            "hbm_gbs_per_gpu": value * BYTES_PER_SITE / 1e9 / world,
            "frac_of_8TBs_per_gpu": value * BYTES_PER_SITE / 1e9 / world / NOMINAL_HBM_GBS,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ((traffic or {}).get("dram_bytes_per_site") or 0) * n or None,
                         "peak_source": peak_src, "launch_ms_mean": mean_launch_ms,
                         "launch_timing": "CUDA events on the launching stream around the K back-to-back launches; mean = elapsed / K",
                         "algorithmic_bytes_per_launch": BYTES_PER_SITE * n,
                         "traffic_note": (f"ncu dram bytes per site ({(traffic or {}).get('dram_bytes_per_site', 0):.2f}, captured at "
                                          f"{(traffic or {}).get('sites_per_launch')} sites/launch) x {n} sites of this launch; "
                                          + str((traffic or {}).get("note"))) if traffic else None},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_all, "clocks": clocks,
            "scaler_increment": total_inc, "side": side,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(pkg, torch, args, local_rank, first, n, total_sites, world, math_mode, barrier, sharding,
            device, ev, left, right):
    """Host buffers -> plf_write_* -> plf_run_async -> plf_read_* on NUM_ACCELERATORS instances."""
    inst = args.instances
    numa = sharding.bind_host_to_device(local_rank) if world > 1 and not args.no_numa_bind else {"bound": False}
    tb = pkg.TestbenchInfo(n, inst)
    if not tb.valid():
        inst, tb = 1, pkg.TestbenchInfo(n, 1)
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    bufs, frees = [], []

    def pinned(nbytes, dtype):
        a, p = pkg.host_alloc(nbytes, dtype)
        frees.append(p)
        return a

    ctx = pkg.Context(local_rank, inst, pkg.LAYOUT_COMB, pkg.INPUT_MEM)
    ctx.set_math(math_mode)
    ctx.set_tuning(args.variant, args.threads, args.blocks_per_sm)
    stream0 = torch.cuda.current_stream().cuda_stream
    header = np.concatenate([ev, left]).astype(np.float32), np.concatenate([ev, right]).astype(np.float32)
    h2d = d2h = 0
    for k in range(inst):
        cnt, lo = tb.alignments_per_instance(k), first + tb.instance_offset(k)
        lb = pinned((80 + 16 * cnt) * 4, np.float32)
        rb = pinned((80 + 16 * cnt) * 4, np.float32)
        out = pinned(16 * cnt * 4, np.float32)
        scb = pinned(cnt, np.uint8)
        # stimulus: generated on the device once, copied to the pinned host buffers (setup, untimed)
        t1 = torch.empty((cnt, 16), device=device)
        t2 = torch.empty((cnt, 16), device=device)
        pkg.generate_device(t1.data_ptr(), t2.data_ptr(), lo, cnt, SEED, stream0)
        torch.cuda.synchronize()
        lb[:80], rb[:80] = header
        torch.from_numpy(lb[80:]).copy_(t1.view(-1))
        torch.from_numpy(rb[80:]).copy_(t2.view(-1))
        del t1, t2
        ctx.instance_alloc(k, cnt)
        bufs.append((cnt, lb, rb, out, scb))
        h2d += lb.nbytes + rb.nbytes
        d2h += out.nbytes + scb.nbytes + 8
    torch.cuda.empty_cache()

    def one_call():
        for k, (cnt, lb, rb, out, scb) in enumerate(bufs):
            ctx.write_left(k, lb)
            ctx.write_right(k, rb)
            ctx.run_async(k, cnt)
            ctx.read_out(k, out)
            ctx.read_scaler(k, scb)
        total = 0
        for k in range(len(bufs)):
            total += ctx.scaler_increment(k)       # waits for the instance
        return total

    one_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        inc = one_call()
    barrier()
    dt = time.perf_counter() - t0
    dt = sharding.max_over_ranks(dt, device)
    lo_scaled = (first + 3) // 4
    assert inc == (first + n + 3) // 4 - lo_scaled, "e2e scaler increment mismatch"
    assert int(bufs[0][4][:8].sum()) == sum(1 for s in range(first, first + min(8, n)) if s % 4 == 0)
    # the streamed host path (plf_newview_stream) on the first instance's buffers: same bytes per site over
    # PCIe, chunked and triple-buffered inside the library, no packed header needed
    cnt0, lb0, rb0, out0, scb0 = bufs[0]
    ctx.newview_stream(ev, left, right, lb0[80:], rb0[80:], out0, scb0, None, n_sites=cnt0)   # untimed: allocates the chunk buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sinc = ctx.newview_stream(ev, left, right, lb0[80:], rb0[80:], out0, scb0, None, n_sites=cnt0)
    torch.cuda.synchronize()
    sdt = sharding.max_over_ranks(time.perf_counter() - t0, device)
    assert sinc == (first + cnt0 + 3) // 4 - lo_scaled, "streamed e2e scaler increment mismatch"
    stream = {"value": cnt0 * world * e2e_steps / sdt, "unit": "sites/s", "sites_per_call_per_gpu": cnt0,
              "pcie_gbs_per_gpu": 193 * cnt0 * e2e_steps / sdt / 1e9,
              "note": "plf_newview_stream over one instance-sized host range per GPU (auto chunks: n/16 clamped to 256 Ki..2 Mi sites, 3 slots)"}
    ctx.close()
    # Bare-copy yardstick of this box at this N (no kernels): every rank at once moves the same H2D : D2H byte ratio
    # between pinned host memory and its GPU with plain cudaMemcpyAsync.  The e2e leg cannot be faster than this.
    p_in, p_out = bufs[0][1], bufs[0][3]                     # instance 0's pinned left / out buffers, reused
    reps = max(2, int((2 << 30) // max(1, p_in.nbytes)))     # about 2 GiB host->device per rank
    barrier()
    psec = pkg.probe_host_link(local_rank, p_in, p_out[: max(1, int(p_in.size * 65 // 128))], reps=reps, pieces=1)
    psec = sharding.max_over_ranks(psec, device)
    probe_bytes = reps * (p_in.nbytes + p_out[: max(1, int(p_in.size * 65 // 128))].nbytes)
    probe_gbs_rank = probe_bytes / psec / 1e9
    for p in frees:
        pkg.host_free(p)
    achieved_gbs = (h2d + d2h) * e2e_steps / dt / 1e9
    link = {"bound": "pcie", "achieved": achieved_gbs * world, "peak": probe_gbs_rank * world, "unit": "GB/s",
            "frac": achieved_gbs / probe_gbs_rank, "per_gpu_achieved": achieved_gbs, "per_gpu_peak": probe_gbs_rank,
            "peak_source": f"plf_probe_host_link in this run: {world} rank(s) at once, {reps} rounds of one "
                           f"{p_in.nbytes >> 20} MiB pinned H2D copy + one {p_in.nbytes * 65 // 128 >> 20} MiB D2H copy "
                           "(the round trip's 128:65 byte ratio), no kernels, wall clock, max over ranks",
            "algorithmic_bytes_per_site": 193}
    return {"value": total_sites * e2e_steps / dt, "unit": "sites/s", "stream": stream, "roofline": link,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
            "instances": inst, "ms_per_step": dt / e2e_steps * 1e3, "host_binding": numa,
            "timing": "host wall clock around plf_write/run/read/wait, barrier + device sync on both sides, max over ranks",
            "pcie_gbs_per_gpu": (h2d + d2h) * e2e_steps / dt / 1e9}


# ---------------------------------------------------------------------------------------------
# side workloads (same run, after the headline): every other BASELINE.json config and the rows SURVEY.md
# section 8f adds, each timed with CUDA events on its launching stream and given its own roofline
# ---------------------------------------------------------------------------------------------
def _roofline(bytes_per_launch, ms, peak, note=None):
    achieved = bytes_per_launch / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "algorithmic_bytes": int(bytes_per_launch)}
    if note:
        r["note"] = note
    return r


def _stochastic(rng, *shape):
    """Row-stochastic 4x4 blocks: CLV magnitudes stay bounded up a 1024-taxon tree."""
    m = rng.random_sample(shape + (4, 4)) + 0.05
    return (m / m.sum(axis=-1, keepdims=True)).astype(np.float32)


def side_cfg2(pkg, torch, device, peak, math_mode, reps=200):
    """BASELINE configs[1]: one PLF instance, 1 Mi sites, one stream.  K back-to-back launches between TWO events (an
    event between launches would break the programmatic-dependent-launch chain)."""
    n = TOTAL_SITES["cfg2"]
    ev, left, right = stimulus_matrices(SEED)
    d_ev, d_pl, d_pr = (torch.from_numpy(a).to(device) for a in (ev, left, right))
    stream = torch.cuda.current_stream().cuda_stream
    opts = pkg.make_opts(math_mode)
    out = {}
    for label, sets in (("cold", 6), ("warm", 1)):
        x1 = [torch.empty((n, 16), device=device) for _ in range(sets)]
        x2 = [torch.empty((n, 16), device=device) for _ in range(sets)]
        x3 = [torch.empty((n, 16), device=device) for _ in range(sets)]
        sc = [torch.empty(n, dtype=torch.uint8, device=device) for _ in range(sets)]
        dsum = torch.zeros(1, dtype=torch.int64, device=device)
        for k in range(sets):
            pkg.generate_device(x1[k].data_ptr(), x2[k].data_ptr(), 0, n, SEED + k, stream)
        args = [(x1[k].data_ptr(), x2[k].data_ptr(), x3[k].data_ptr(), sc[k].data_ptr(), d_ev.data_ptr(), d_pl.data_ptr(),
                 d_pr.data_ptr(), None, n, dsum.data_ptr(), opts, stream) for k in range(sets)]
        for i in range(2 * sets + 4):
            pkg.newview_device(*args[i % sets])
        torch.cuda.synchronize()
        dsum.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(reps):
            pkg.newview_device(*args[i % sets])
        e1.record()
        issue_us = (time.perf_counter() - t0) / reps * 1e6
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert int(dsum.item()) == reps * ((n + 3) // 4), "cfg2: scaler increments do not match the stimulus design"
        out[label] = {"value": n / (ms * 1e-3), "unit": "sites/s", "us_per_launch": ms * 1e3, "launches": reps,
                      "buffer_sets": sets, "host_issue_us_per_launch": issue_us,
                      "roofline": _roofline(BYTES_PER_SITE * n, ms, peak),
                      "frac_of_8TBs": BYTES_PER_SITE * n / (ms * 1e-3) / 1e9 / NOMINAL_HBM_GBS}
        del x1, x2, x3, sc
        torch.cuda.empty_cache()
    out["workload"] = WORKLOAD_NAME["cfg2"]
    return out


def side_gen(pkg, torch, device, peak, math_mode, n, world, sharding, reps=10):
    """BASELINE configs[3]: INPUT_SRC=gen analogue -- no CLV is read.  cfg4a writes CLV + scaler (65 B/site), cfg4b
    folds the outputs into a checksum (no memory traffic: the kernel's pure issue rate)."""
    x3 = torch.empty((n, 16), device=device)
    sc = torch.empty(n, dtype=torch.uint8, device=device)
    dsum = torch.zeros(1, dtype=torch.int64, device=device)
    chk = torch.zeros(1, dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream().cuda_stream
    opts = pkg.make_opts(math_mode)
    out = {}
    for label, sink in (("cfg4a_write", pkg.GEN_WRITE), ("cfg4b_discard", pkg.GEN_DISCARD)):
        a = (x3.data_ptr(), sc.data_ptr(), n, dsum.data_ptr(), chk.data_ptr(), sink, opts, stream)
        for _ in range(3):
            pkg.newview_gen_device(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            pkg.newview_gen_device(*a)
        e1.record()
        torch.cuda.synchronize()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1) / reps, device)
        total = sharding.reduce_scaler_increment(n, device)
        row = {"value": total / (ms * 1e-3), "unit": "sites/s", "ms_per_launch": ms, "sites_per_gpu": n}
        if sink == pkg.GEN_WRITE:
            row["roofline"] = _roofline(65 * n, ms, peak, "writes only: 64 B CLV + 1 scaler byte per site")
        else:
            row["roofline"] = None
            row["note"] = "no memory traffic: arithmetic + vote + reduction only (the sink of s2mm_gen drops the stream)"
        out[label] = row
    del x3, sc
    torch.cuda.empty_cache()
    return out


def side_tree(pkg, torch, device, peak, math_mode, n, first, world, local_rank, sharding, barrier, tips=1024, reps=3):
    """BASELINE configs[4]: chained newview over a synthetic 1024-taxon tree, 1 Mi sites split over the ranks, per-site
    scaler counts accumulated up the tree, total all-reduced.  Dense tips (64 B/site) and state-code tips (1 B/site)."""
    left, right = pkg.balanced_tree(tips)
    rng = np.random.RandomState(1)
    ev = _stochastic(rng).reshape(16)
    pl = _stochastic(rng, tips - 1, 4).reshape(tips - 1, 64)
    pr = _stochastic(rng, tips - 1, 4).reshape(tips - 1, 64)
    out = {}
    for label, codes in (("dense_tips", False), ("code_tips", True)):
        t = pkg.Tree(left, right, n, device=local_rank, tip_codes=codes)
        t.set_math(math_mode)
        if codes:
            tv = (rng.random_sample((16, 4)) * np.where(np.arange(16)[:, None] % 4 == 0, 1e-10, 1.0)).astype(np.float32)
            t.write_tip_vector(tv)
            gsite = np.arange(first, first + n, dtype=np.uint64)          # a function of the GLOBAL site index and the tip,
            for tip in range(tips):                                       # so every N traverses the same alignment
                h = (gsite + np.uint64(tip) * np.uint64(0x9E3779B1)) * np.uint64(0xD6E8FEB86659FD93)
                t.write_tip_codes(tip, ((h >> np.uint64(47)) & np.uint64(15)).astype(np.uint8))
        else:
            scratch = torch.empty((n, 16), device=device)
            for tip in range(tips):      # x1-stream of the generator for even tips, x2-stream for odd ones
                a, b = (t.tip_ptr(tip), scratch.data_ptr()) if tip % 2 == 0 else (scratch.data_ptr(), t.tip_ptr(tip))
                pkg.generate_device(a, b, first + tip * 7919, n, 1000 + tip)
            torch.cuda.synchronize()
            del scratch
        t.write_matrices(ev, pl, pr)
        info = t.info()
        for _ in range(2):
            t.run_async()
        t.wait()
        times = []
        for _ in range(reps):
            barrier()
            t.run_async()
            t.wait()
            times.append(t.last_ms())
        ms = sharding.max_over_ranks(float(np.median(times)), device)
        total_scalings = sharding.reduce_scaler_increment(t.total_scalings(), device)
        total_sites = sharding.reduce_scaler_increment(n, device)
        lnl = None
        if not codes:
            diag = np.exp(-np.linspace(0.0, 1.5, 16)).astype(np.float32)
            lnl = sharding.reduce_log_likelihood(t.evaluate_root(diag), device)
        root, cnt = t.read_root(0, min(n, 1024))
        assert np.isfinite(root).all(), "cfg5: non-finite root CLV"
        out[label] = {"value": (tips - 1) * total_sites / (ms * 1e-3), "unit": "newview-sites/s", "ms_per_traversal": ms,
                      "tips": tips, "total_sites": total_sites, "sites_per_gpu": n, "levels": info["levels"],
                      "device_GiB_per_gpu": info["device_bytes"] / 2 ** 30, "total_scalings": total_scalings,
                      "root_scaler_count_max": int(cnt.max()), "log_likelihood": lnl,
                      "roofline": _roofline(info["traversal_bytes"], ms, peak,
                                            "per GPU; algorithmic bytes = plf_tree_info: 64 B/site per CLV read or written, "
                                            "1 B/site per code tip read, 4 B/site per count vector read or written")}
        t.close()
        torch.cuda.empty_cache()
    out["workload"] = "BASELINE.json configs[4]: chained PLF over a synthetic 1024-taxon balanced tree, 1Mi sites over N GPUs"
    return out


def side_evaluate(pkg, torch, device, peak, n, world, sharding, reps=10):
    """Root log-likelihood kernel over the rank's cfg3 shard: 2 CLV reads + 2 count vectors per site."""
    x1 = torch.empty((n, 16), device=device)
    x2 = torch.empty((n, 16), device=device)
    stream = torch.cuda.current_stream().cuda_stream
    pkg.generate_device(x1.data_ptr(), x2.data_ptr(), 0, n, SEED, stream)
    c1 = torch.ones(n, dtype=torch.int32, device=device)
    c2 = torch.zeros(n, dtype=torch.int32, device=device)
    diag = torch.from_numpy(np.exp(-np.linspace(0.0, 1.5, 16)).astype(np.float32)).to(device)
    lnl = torch.zeros(reps + 3, dtype=torch.float64, device=device)
    a = lambda i: (x1.data_ptr(), x2.data_ptr(), c1.data_ptr(), c2.data_ptr(), None, diag.data_ptr(), n,
                   lnl[i:].data_ptr(), stream)
    for i in range(3):
        pkg.evaluate_device(*a(i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        pkg.evaluate_device(*a(3 + i))
    e1.record()
    torch.cuda.synchronize()
    vals = lnl.cpu().numpy()
    assert np.isfinite(vals).all() and (vals == vals[0]).all(), "evaluate: the log-likelihood is not reproducible run to run"
    ms = sharding.max_over_ranks(e0.elapsed_time(e1) / reps, device)
    total = sharding.reduce_scaler_increment(n, device)
    del x1, x2, c1, c2
    torch.cuda.empty_cache()
    return {"value": total / (ms * 1e-3), "unit": "sites/s", "ms_per_launch": ms, "sites_per_gpu": n,
            "bitwise_reproducible": True, "log_likelihood_rank0": float(vals[0]),
            "roofline": _roofline(136 * n, ms, peak, "128 B of CLV + 8 B of scaler counts read per site; fp64 log per site")}


def side_protein(pkg, torch, peak, reps=5, n=2 << 20):
    """STATES=protein: the 20-state newview, strict and FMA, device-resident CLVs."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import protein_bench
    res = protein_bench.measure(pkg, torch, n, reps, shapes=[(0, 0), (4, 256)], maths=(0, 1), verbose=False)
    out = {"sites": n, "bytes_per_site": res["bytes_per_site"], "muladd_per_site": res["muladd_per_site"]}
    label = {("strict", 0): "strict", ("fma", 0): "fma", ("fma", 4): "fma_cuda_cores"}
    for row in res["rows"]:
        key = label.get((row["math"], row["variant"]))
        if key is None:
            continue
        assert row["ok"], "20-state run failed its self-checks"
        out[key] = {"value": row["gsites"] * 1e9, "unit": "sites/s", "ms_per_launch": row["ms_mean"],
                    "tmuladd_per_s": row["tmuladd_per_s"], "kernel": {k: row[k] for k in ("regs", "threads", "smem_bytes")},
                    "roofline": _roofline(res["bytes_per_site"] * n, row["ms_mean"], peak,
                                          "on the HBM / fp32 ridge: 961 B and 4800 multiply-adds per site")}
    out["fma"]["kernel"]["path"] = "tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators, TMA operand boxes (csrc/plf_protein_tc.cu)"
    out["fma_cuda_cores"]["kernel"]["path"] = "FFMA2 register tile (csrc/plf_protein.cu, variant 4 x 256 threads)"
    torch.cuda.empty_cache()
    return out


def run_side(pkg, torch, args, rank, world, local_rank, device, math_mode, barrier, sharding, peak):
    t_start = time.perf_counter()
    side = {}
    small = args.side_small
    first3, n3 = sharding.shard_for_rank((1 << 20) if small else TOTAL_SITES["cfg3"], rank, world)
    first5, n5 = sharding.shard_for_rank((1 << 14) if small else (1 << 20), rank, world)
    tips = 64 if small else 1024
    steps = [("cfg4", lambda: side_gen(pkg, torch, device, peak, math_mode, n3, world, sharding)),
             ("evaluate", lambda: side_evaluate(pkg, torch, device, peak, n3, world, sharding)),
             ("cfg5", lambda: side_tree(pkg, torch, device, peak, math_mode, n5, first5, world, local_rank, sharding, barrier,
                                        tips=tips))]
    if world == 1:          # single-GPU configurations
        steps = [("cfg2", lambda: side_cfg2(pkg, torch, device, peak, math_mode, reps=20 if small else 200)),
                 ("protein", lambda: side_protein(pkg, torch, peak, n=(1 << 17) if small else (2 << 20)))] + steps
    for name, fn in steps:
        barrier()
        t0 = time.perf_counter()
        side[name] = fn()
        side[name]["wall_s"] = round(time.perf_counter() - t0, 2)
    side["wall_s_total"] = round(time.perf_counter() - t_start, 2)
    side["note"] = ("measured in this run after the headline; CUDA events on the launching stream, max over ranks; "
                    "roofline.peak = the same measured copy bandwidth as the headline")
    return side


def run_protein_arm(args):
    """`--workload protein`: the 20-state newview on one GPU, device-resident CLVs, with the CPU
    restatement (reference plf() loop nest at S = 20, all host threads) timed beside it.  Not the headline line:
    the reference has no protein path, so there is no reference arm for it."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import protein_bench
    pkg = load_pkg()
    torch.cuda.set_device(0)
    n = args.sites or (4 << 20)
    K = max(3, args.steps)
    from tools.clocks import ClockSampler
    sampler = ClockSampler(0)
    sampler.start()
    sampler.begin()
    res = protein_bench.measure(pkg, torch, n, K, shapes=[(0, 0)], maths=(0, 1), verbose=False)
    sampler.end()
    clk = sampler.stop()
    strict, fma = res["rows"]
    assert strict["ok"] and fma["ok"], "20-state run failed its self-checks"
    cpu = None
    if not args.no_cpu_baseline:
        import oracle                                   # cpu_baseline leg
        cores = os.cpu_count() or 1
        sample = 1 << 18
        rng = np.random.RandomState(SEED)
        ev, left, right = (rng.random_sample(k).astype(np.float32) for k in (400, 1600, 1600))
        x1, x2 = pkg.generate_states_host(20, 0, sample, SEED)
        co = oracle.COracle()
        best = {}
        for thr in (1, cores):
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                _, _, inc = co.newview_states(20, x1, x2, ev, left, right, nthreads=thr)
                ts.append(time.perf_counter() - t0)
            assert inc == sample // 4
            best[thr] = sample / min(ts)
        cpu = {"value": best[cores], "unit": "sites/s", "cores": cores, "kind": "port",
               "sample": f"{sample} sites, best of 2, reference plf() loop nest with 20 states (-O2 -ffp-contract=off)",
               "single_thread_sites_per_s": best[1]}
    line = {"metric": "plf_sites_per_s", "value": strict["gsites"] * 1e9, "unit": "sites/s", "n_gpus": 1, "steps": K,
            "warmup": 2, "ms_per_step": strict["ms_mean"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "STATES=protein: 20 states x 4 rate categories, device-resident CLVs (SURVEY 8f.3)",
                       "total_sites": n, "math": "strict", "bytes_per_site": res["bytes_per_site"],
                       "muladd_per_site": res["muladd_per_site"], "kernel": {k: strict[k] for k in ("regs", "threads", "smem_bytes")},
                       "l2": f"inputs {2 * n * 320 >> 20} MiB >> 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": strict["gbs"], "peak": res["peak_gbs"], "unit": "GB/s",
                         "frac": strict["gbs"] / res["peak_gbs"], "traffic": None,
                         "note": "on the ridge: 4800 multiply-adds and 961 B per site; fp32 rate "
                                 f"{strict['tmuladd_per_s']:.1f} T mul-add/s of ~34 T/s (116/clk/SM measured)"},
            "fma_mode": {"value": fma["gsites"] * 1e9, "gbs": fma["gbs"], "frac": fma["gbs"] / res["peak_gbs"],
                         "tmuladd_per_s": fma["tmuladd_per_s"], "kernel": {k: fma[k] for k in ("regs", "threads", "smem_bytes")}},
            "cpu_baseline": cpu, "e2e": None, "gpu_launches": 2 * (K + 2) + 1, "clocks": clk}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(TOTAL_SITES) + ["protein"])
    ap.add_argument("--sites", type=int, default=0, help="override total site count")
    ap.add_argument("--math", default="strict", choices=["strict", "fma"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--buffer-sets", type=int, default=0, help="0 = auto")
    ap.add_argument("--instances", type=int, default=9, help="NUM_ACCELERATORS for the e2e leg")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the side workloads (cfg2, cfg4, cfg5, protein, evaluate)")
    ap.add_argument("--side-small", action="store_true", help="side workloads at test sizes (contract test)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin each rank to its GPU's local CPUs")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: route everything libraries print on fd 1 (e.g. NCCL's
    # version banner) to stderr and restore the real stdout only around our own print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.workload == "protein":
        if args.impl == "reference" or int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise SystemExit("--workload protein is a single-GPU b200-arm measurement")
        run_protein_arm(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
