// host/host_mem.cpp -- host test bench, INPUT_SRC=mem.  Drop-in for the reference's
// app/src/host_mem.cpp: same five positional arguments, same phases (configure -> allocate ->
// generate stimulus -> pack [EV|P|CLV] -> per call: write / run / read per instance -> host scaler
// reduction -> verify against the CPU golden -> timing tables / CSV), same compile-time switches
// NO_PRERUN_CHECK, NO_CORRECTNESS_CHECK, NO_INTERMEDIATE_RESULTS (reference Makefile:145-161).
// XRT is replaced by the C ABI of include/b200plf.h; the accelerator is one fused CUDA kernel.
//
//   host_mem.exe <config name | x.xclbin> <device: ordinal, PCI BDF, or list 0,1,..> <sites> <plf calls> <instances used>
//
// Differences from the reference, on purpose (SURVEY.md section 5 "hazards"):
//   * bad arguments are fatal (the reference prints and continues with an uninitialised field);
//   * the stimulus generator is seeded (PLF_SEED, default 42) so runs are reproducible;
//   * sizes are 64-bit; the layout comes from the configuration name and unknown names are errors;
//   * a comma-separated device list spreads the instances round-robin over several GPUs (plf_multi: one context
//     per GPU, and the kernel-fused scaler increments of the GPUs are summed with one NCCL all-reduce per call).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "b200plf.h"
#include "golden_plf.h"
#include "plfb_file.h"
#include "tb_info.h"
#include "timing_report.h"

using namespace plfhost;

namespace {

[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "host_mem: " << msg << std::endl;
    std::exit(2);
}

void check(int rc, plf_ctx *ctx, const char *what)
{
    if (rc != PLF_OK) die(std::string(what) + ": " + plf_last_error(ctx));
}

size_t parse_count(const char *s, const char *what)
{
    char *end = nullptr;
    errno = 0;
    const unsigned long long v = std::strtoull(s, &end, 10);
    if (end == s || *end != '\0' || errno != 0 || s[0] == '-') die(std::string("invalid ") + what + ": '" + s + "'");
    return static_cast<size_t>(v);
}

std::vector<int> parse_devices(const std::string &arg)
{
    std::vector<int> devs;
    std::stringstream ss(arg);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        int d = -1;
        if (plf_device_from_string(tok.c_str(), &d) != PLF_OK) die(plf_last_error(nullptr));
        devs.push_back(d);
    }
    if (devs.empty()) die("no device given");
    return devs;
}

[[maybe_unused]] bool prerun_check()
{
    // Same Y/n gate as the reference (app/src/utils.cpp:9-39); EOF counts as yes.
    for (;;) {
        std::cout << std::endl << "Ready to continue with test? [Y/n]: " << std::flush;
        std::string line;
        if (!std::getline(std::cin, line)) return true;
        if (line.empty() || line[0] == 'y' || line[0] == 'Y') return true;
        if (line[0] == 'n' || line[0] == 'N') return false;
        std::cout << "You may only type 'y' or 'n'." << std::endl;
    }
}

template <class T>
T *pinned(size_t count)
{
    void *p = nullptr;
    if (plf_host_alloc(&p, std::max<size_t>(count, 1) * sizeof(T)) != PLF_OK) die(plf_last_error(nullptr));
    return static_cast<T *>(p);
}

struct InstanceSlot {
    plf_ctx *ctx;
    unsigned local;   // instance index inside ctx
};

}  // namespace

int main(int argc, char *argv[])
{
    if (argc != 6) {
        std::cerr << "Not correct amount of parameters provided. Usage: " << argv[0]
                  << " <config name or /path/to/a.xclbin> <device ordinal | PCI BDF | list> <number of alignments>"
                     " <number of plf calls> <parallel instances used>"
                  << std::endl;
        return 2;
    }
    AcceleratorConfig cfg;
    try {
        cfg = parse_config(argv[1]);
    } catch (const std::exception &e) {
        die(e.what());
    }
    if (cfg.input_src != PLF_INPUT_MEM) die("configuration '" + cfg.name + "' is INPUT_SRC=gen: use host_gen.exe");
    const std::vector<int> devices = parse_devices(argv[2]);

    TestbenchInfo tb;
    tb.alignment_sites = parse_count(argv[3], "number of alignments");
    tb.plf_calls = parse_count(argv[4], "number of plf calls");
    tb.parallel_instances = static_cast<unsigned>(parse_count(argv[5], "parallel instances"));
    tb.window_size = cfg.window_size;
    tb.layout = cfg.layout;
    tb.set_states(cfg.n_states);                      // STATES knob: DNA (4) or AA (20), from the configuration name
    if (tb.alignment_sites == 0 || tb.plf_calls == 0 || tb.parallel_instances == 0)
        die("alignments, plf calls and instances must all be > 0");
    if (tb.parallel_instances > cfg.num_accelerators * devices.size())
        die("instances used (" + std::to_string(tb.parallel_instances) + ") exceeds NUM_ACCELERATORS (" +
            std::to_string(cfg.num_accelerators) + ") x devices (" + std::to_string(devices.size()) + ")");
    if (!tb.split_is_valid())
        die(std::to_string(tb.alignment_sites) + " alignments cannot be split over " +
            std::to_string(tb.parallel_instances) + " instances: the last instance would be empty");

    // ---- configuration tables (host_mem.cpp:45-101) -------------------------------------------
    const std::string bar(84, '=');
    auto kv = [](const std::string &k, const std::string &v) {
        std::cout << "| " << std::left << std::setw(23) << k << " | " << std::setw(54) << v << " |" << std::endl;
    };
    auto sizes = [](const std::string &k, size_t a, size_t e, size_t b) {
        std::cout << "| " << std::left << std::setw(23) << k << " | " << std::right << std::setw(16) << a << " | "
                  << std::setw(16) << e << " | " << std::setw(16) << b << " |" << std::endl;
    };
    std::cout << std::endl << bar << std::endl;
    kv("test name:", "plf (B200 / CUDA sm_100a)");
    kv("PL name:", cfg.pl_name);
    kv("AIE name:", cfg.aie_name);
    std::cout << bar << std::endl;
    kv("alignment sites:", std::to_string(tb.alignment_sites));
    kv("plf calls:", std::to_string(tb.plf_calls));
    kv("parallel plfs:", std::to_string(tb.parallel_instances));
    kv("STATES:", cfg.states + " (" + std::to_string(tb.states) + " states x 4 rate categories, " + std::to_string(tb.elements_per_alignment) + " floats per site)");
    kv("PLIO layout:", cfg.layout == PLF_LAYOUT_COMB ? "Comb  [EV|P|CLV] / [EV|P|CLV]" : "Sep  [EV|P|CLV] / [P|CLV]");
    kv("AIE window size:", std::to_string(tb.window_size) + " (ignored: no windows on the GPU)");
    std::cout << bar << std::endl;
    std::cout << "|                         |       alignments |         elements |     size (bytes) |" << std::endl;
    sizes("instance left:", tb.alignments_per_instance(), tb.instance_elements_left(), tb.instance_elements_left() * 4);
    sizes("instance right:", tb.alignments_per_instance(), tb.instance_elements_right(), tb.instance_elements_right() * 4);
    sizes("instance out:", tb.alignments_per_instance(), tb.instance_elements_out(), tb.instance_elements_out() * 4);
    sizes("total (" + std::to_string(tb.plf_calls) + " plf calls):", tb.alignment_sites * tb.plf_calls, tb.data_elements(),
          tb.data_size());
    std::cout << bar << std::endl;
    kv("RAM usage (host):", std::to_string(tb.host_mem_usage() / 1e9) + " GB");
    kv("RAM usage (B200):", std::to_string(tb.device_mem_usage() / 1e9) + " GB of 180 GB per GPU");
    std::cout << bar << std::endl;
    for (int d : devices) {
        char name[256], bdf[32];
        if (plf_device_info(d, name, sizeof name, bdf, sizeof bdf) != PLF_OK) die(plf_last_error(nullptr));
        kv("device name:", name);
        kv("device bdf:", bdf);
    }
    std::cout << bar << std::endl << std::right;

    // ---- Init: contexts (acap_info) and device buffers (xrt::bo) --------------------------------
    plf_multi *multi = nullptr;
    if (plf_multi_create_states(&multi, devices.data(), static_cast<int>(devices.size()), cfg.num_accelerators, cfg.layout, PLF_INPUT_MEM,
                                static_cast<int>(tb.states)) != PLF_OK)
        die(std::string("plf_multi_create: ") + plf_multi_last_error(nullptr));
    std::vector<plf_ctx *> ctxs;
    for (size_t r = 0; r < devices.size(); ++r) {
        plf_ctx *c = plf_multi_ctx(multi, static_cast<int>(r));
        if (const char *m = std::getenv("PLF_MATH")) check(plf_ctx_set_math(c, std::strcmp(m, "fma") == 0 ? PLF_MATH_FMA : PLF_MATH_STRICT), c, "plf_ctx_set_math");
        ctxs.push_back(c);
    }
    {
        int nccl_version = 0;
        plf_multi_info(multi, nullptr, &nccl_version, nullptr);
        if (devices.size() > 1) std::cout << "scaler increments of the " << devices.size() << " GPUs are reduced with NCCL " << nccl_version << std::endl;
    }
    // per call: sum of the kernel-fused increments of each GPU's instances, then one all-reduce over the GPUs
    auto reduce_fused = [&](const std::vector<long long> &per_gpu) {
        long long total = 0;
        if (plf_multi_reduce(multi, per_gpu.data(), nullptr, &total, nullptr) != PLF_OK)
            die(std::string("plf_multi_reduce: ") + plf_multi_last_error(multi));
        return total;
    };
    std::vector<InstanceSlot> slot(tb.parallel_instances);
    for (unsigned k = 0; k < tb.parallel_instances; ++k) {
        slot[k] = {ctxs[k % ctxs.size()], static_cast<unsigned>(k / ctxs.size())};
        check(plf_instance_alloc(slot[k].ctx, slot[k].local, tb.alignments_per_instance()), slot[k].ctx, "plf_instance_alloc");
    }
    std::cout << "alignments per instance: " << tb.alignments_per_instance() << ", padding: " << tb.alignments_padding() << std::endl;
    std::cout << "connected to " << tb.parallel_instances << " PLF instance(s) on " << devices.size() << " device(s)" << std::endl;

#if !defined(NO_PRERUN_CHECK) || NO_PRERUN_CHECK == 0
    if (!prerun_check()) return 0;
#endif

    // ---- Load data: the reference's stimulus recipe (host_mem.cpp:179-209), seeded ----------------
    std::cout << "Initialize test data ... " << std::endl;
    const char *seed_env = std::getenv("PLF_SEED");
    std::mt19937 gen(seed_env ? static_cast<uint32_t>(std::strtoul(seed_env, nullptr, 10)) : 42u);
    std::uniform_real_distribution<> dis(0.0, 1.0);
    const size_t SF = tb.elements_per_alignment, EVF = tb.ev_elements(), PF = tb.branch_elements();     // 16 / 16 / 64 for DNA
    std::vector<float> ev_v(EVF), bl_v(PF), br_v(PF);
    float *ev = ev_v.data(), *branchleft = bl_v.data(), *branchright = br_v.data();
    for (size_t j = 0; j < EVF; ++j) ev[j] = static_cast<float>(dis(gen));
    for (size_t j = 0; j < PF; ++j) {
        branchleft[j] = static_cast<float>(dis(gen));
        branchright[j] = static_cast<float>(dis(gen));
    }
    std::vector<float> alignmentsleft(tb.elements_per_plf()), alignmentsright(tb.elements_per_plf());
    // 1e-12 as in the reference for DNA; 1e-14 for AA, whose 20-term sums are 25x larger
    const float tiny_scale = tb.states == 4 ? static_cast<float>(std::pow(1.0e-12, 1)) : 1.0e-14f;
    for (size_t j = 0; j < tb.elements_per_plf(); ++j) {
        const float scale = ((j / SF) % 4 == 0) ? tiny_scale : 1.0f;    // every 4th site underflows (j % 64 < 16 for DNA)
        alignmentsleft[j] = static_cast<float>(dis(gen) * scale);
        alignmentsright[j] = static_cast<float>(dis(gen));
    }
    // PLF_LOAD_DIR=<dir>: take the stimulus from <dir>/left.plfb and <dir>/right.plfb (whole-run packed buffers,
    // host/plfb_file.h) instead of the generator
    if (const char *dir = std::getenv("PLF_LOAD_DIR")) {
        try {
            std::vector<char> lb, rb;
            const PlfbHeader hl = read_plfb(std::string(dir) + "/left.plfb", lb);
            const PlfbHeader hr = read_plfb(std::string(dir) + "/right.plfb", rb);
            if (hl.kind != PLFB_LEFT || hr.kind != PLFB_RIGHT) die("PLF_LOAD_DIR: left.plfb / right.plfb hold the wrong buffer kind");
            if (hl.sites != tb.alignment_sites || hr.sites != tb.alignment_sites)
                die("PLF_LOAD_DIR: the files hold " + std::to_string(hl.sites) + " sites, the run asks for " +
                    std::to_string(tb.alignment_sites));
            const float *l = reinterpret_cast<const float *>(lb.data()), *r = reinterpret_cast<const float *>(rb.data());
            if (tb.states != 4) die("PLF_LOAD_DIR: PLFB files hold DNA buffers");
            std::copy(l, l + 16, ev);
            std::copy(l + 16, l + 80, branchleft);
            std::copy(l + 80, l + 80 + tb.elements_per_plf(), alignmentsleft.begin());
            const size_t roff = hr.layout == 0 ? 16 : 0;
            std::copy(r + roff, r + roff + 64, branchright);
            std::copy(r + roff + 64, r + roff + 64 + tb.elements_per_plf(), alignmentsright.begin());
            std::cout << "Stimulus loaded from " << dir << std::endl;
        } catch (const std::exception &e) {
            die(e.what());
        }
    }
    std::vector<int> wgt(tb.alignment_sites, 1);

    // results per call (pinned so the reads are asynchronous)
    std::vector<float *> result(tb.plf_calls);
    std::vector<char *> scalerVector(tb.plf_calls);
    std::vector<long long> scalerIncrement(tb.plf_calls, 0), scalerIncrementFused(tb.plf_calls, 0);
    for (size_t i = 0; i < tb.plf_calls; ++i) {
        result[i] = pinned<float>(tb.elements_per_plf());
        scalerVector[i] = pinned<char>(tb.alignment_sites);
    }

    std::cout << "Prepare data for transfer ... " << std::endl;
    std::vector<float *> dataLeft(tb.parallel_instances), dataRight(tb.parallel_instances);
    auto pack = [&](unsigned k) {
        const size_t off = tb.instance_first_site(k) * SF, cnt = tb.alignments_per_instance(k) * SF;
        float *l = dataLeft[k], *r = dataRight[k];
        std::copy(ev, ev + EVF, l);                                                     // [EV | P_left | CLV]
        std::copy(branchleft, branchleft + PF, l + EVF);
        std::copy(alignmentsleft.begin() + off, alignmentsleft.begin() + off + cnt, l + EVF + PF);
        if (tb.layout == PLF_LAYOUT_COMB) {
            std::copy(ev, ev + EVF, r);
            std::copy(branchright, branchright + PF, r + EVF);
            std::copy(alignmentsright.begin() + off, alignmentsright.begin() + off + cnt, r + EVF + PF);
        } else {
            std::copy(branchright, branchright + PF, r);
            std::copy(alignmentsright.begin() + off, alignmentsright.begin() + off + cnt, r + PF);
        }
    };
    for (unsigned k = 0; k < tb.parallel_instances; ++k) {
        dataLeft[k] = pinned<float>(tb.instance_elements_left());
        dataRight[k] = pinned<float>(tb.instance_elements_right());
#if !defined(NO_INTERMEDIATE_RESULTS) || NO_INTERMEDIATE_RESULTS == 0
        pack(k);
#endif
    }

    // ---- Run ---------------------------------------------------------------------------------------
    std::cout << "Start PLF calculation on accelerator ... " << std::endl << std::endl;
    Timer t;
    auto enqueue_instance = [&](unsigned k, size_t call, bool marks) {
        plf_ctx *c = slot[k].ctx;
        const unsigned li = slot[k].local;
        const size_t first = tb.instance_first_site(k), cnt = tb.alignments_per_instance(k);
        if (marks) check(plf_mark(c, li, PLF_MARK_BEGIN), c, "plf_mark");
        check(plf_write_left(c, li, dataLeft[k], tb.instance_active_elements_left(k) * sizeof(float), 0), c, "plf_write_left");
        check(plf_write_right(c, li, dataRight[k], tb.instance_active_elements_right(k) * sizeof(float), 0), c, "plf_write_right");
        if (marks) check(plf_mark(c, li, PLF_MARK_T1), c, "plf_mark");
        check(plf_run_async(c, li, cnt), c, "plf_run_async");
        if (marks) check(plf_mark(c, li, PLF_MARK_T2), c, "plf_mark");
        check(plf_read_out(c, li, result[call] + first * SF, cnt * SF * sizeof(float), 0), c, "plf_read_out");
        check(plf_read_scaler(c, li, scalerVector[call] + first, cnt, 0), c, "plf_read_scaler");
        if (marks) check(plf_mark(c, li, PLF_MARK_END), c, "plf_mark");
    };

#if !defined(NO_INTERMEDIATE_RESULTS) || NO_INTERMEDIATE_RESULTS == 0
    std::vector<TimingData> execution_ms(tb.parallel_instances, TimingData(tb.plf_calls));
    plf_range_push("roundtrip_exec_time");            // the reference's xrt::profile::user_range (host_mem.cpp:273,282,395)
    for (size_t i = 0; i < tb.plf_calls; ++i) {
        const double call_begin = t.elapsed_ms();
        std::vector<long long> fused_per_gpu(ctxs.size(), 0);
        for (unsigned k = 0; k < tb.parallel_instances; ++k) enqueue_instance(k, i, true);
        for (unsigned k = 0; k < tb.parallel_instances; ++k) {                 // sync all instances
            plf_ctx *c = slot[k].ctx;
            check(plf_wait(c, slot[k].local), c, "plf_wait");
            float hm = 0, msm = 0, mh = 0;
            check(plf_elapsed_ms(c, slot[k].local, PLF_MARK_BEGIN, PLF_MARK_T1, &hm), c, "plf_elapsed_ms");
            check(plf_elapsed_ms(c, slot[k].local, PLF_MARK_T1, PLF_MARK_T2, &msm), c, "plf_elapsed_ms");
            check(plf_elapsed_ms(c, slot[k].local, PLF_MARK_T2, PLF_MARK_END, &mh), c, "plf_elapsed_ms");
            TimingData &d = execution_ms[k];
            d.begin[i] = call_begin;
            d.t1[i] = d.begin[i] + hm;
            d.t2[i] = d.t1[i] + msm;
            d.end[i] = d.t2[i] + mh;
            long long fused = 0;
            check(plf_scaler_increment(c, slot[k].local, &fused), c, "plf_scaler_increment");
            fused_per_gpu[k % ctxs.size()] += fused;
        }
        scalerIncrementFused[i] = reduce_fused(fused_per_gpu);
        // host-side scaler reduction, as the reference does it (host_mem.cpp:384-388)
        for (size_t j = 0; j < tb.alignment_sites; ++j) scalerIncrement[i] += static_cast<long long>(scalerVector[i][j]) * wgt[j];
    }
    plf_range_pop();
#else
    TimingData execution_ms(tb.plf_calls);
    plf_range_push("roundtrip_exec_time");
    for (size_t i = 0; i < tb.plf_calls; ++i) {
        execution_ms.begin[i] = t.elapsed_ms();
        std::vector<long long> fused_per_gpu(ctxs.size(), 0);
        for (unsigned k = 0; k < tb.parallel_instances; ++k) pack(k);             // packing is part of the round trip
        execution_ms.t1[i] = t.elapsed_ms();
        for (unsigned k = 0; k < tb.parallel_instances; ++k) enqueue_instance(k, i, false);
        for (unsigned k = 0; k < tb.parallel_instances; ++k) {
            check(plf_wait(slot[k].ctx, slot[k].local), slot[k].ctx, "plf_wait");
            long long fused = 0;
            check(plf_scaler_increment(slot[k].ctx, slot[k].local, &fused), slot[k].ctx, "plf_scaler_increment");
            fused_per_gpu[k % ctxs.size()] += fused;
        }
        scalerIncrementFused[i] = reduce_fused(fused_per_gpu);
        execution_ms.t2[i] = t.elapsed_ms();
        for (size_t j = 0; j < tb.alignment_sites; ++j) scalerIncrement[i] += static_cast<long long>(scalerVector[i][j]) * wgt[j];
        execution_ms.end[i] = t.elapsed_ms();
    }
    plf_range_pop();
#endif

    // PLF_DUMP_DIR=<dir>: the whole-run packed inputs and the outputs of call 0 as PLFB files
    if (const char *dir = std::getenv("PLF_DUMP_DIR")) {
        if (tb.states != 4) die("PLF_DUMP_DIR: PLFB files hold DNA buffers");
        try {
            const size_t n = tb.alignment_sites, roff = tb.layout == PLF_LAYOUT_COMB ? 16 : 0;
            std::vector<float> lb(80 + 16 * n), rb(roff + 64 + 16 * n);
            std::copy(ev, ev + 16, lb.begin());
            std::copy(branchleft, branchleft + 64, lb.begin() + 16);
            std::copy(alignmentsleft.begin(), alignmentsleft.end(), lb.begin() + 80);
            if (roff) std::copy(ev, ev + 16, rb.begin());
            std::copy(branchright, branchright + 64, rb.begin() + roff);
            std::copy(alignmentsright.begin(), alignmentsright.end(), rb.begin() + roff + 64);
            const uint32_t lay = tb.layout == PLF_LAYOUT_COMB ? 0u : 1u;
            write_plfb(std::string(dir) + "/left.plfb", PLFB_LEFT, lay, n, lb.data());
            write_plfb(std::string(dir) + "/right.plfb", PLFB_RIGHT, lay, n, rb.data());
            write_plfb(std::string(dir) + "/out.plfb", PLFB_OUT, lay, n, result[0]);
            write_plfb(std::string(dir) + "/scaler.plfb", PLFB_SCALER, lay, n, scalerVector[0]);
            std::cout << "Buffers written to " << dir << std::endl;
        } catch (const std::exception &e) {
            die(e.what());
        }
    }

    // ---- Check: CPU golden, exact comparison (host_mem.cpp:403-442) --------------------------------
    TimingData reference_ms(tb.plf_calls);
    int exit_code = 0;
#if !defined(NO_CORRECTNESS_CHECK) || NO_CORRECTNESS_CHECK == 0
    std::cout << "Data collected, checking for correctness ..." << std::endl;
    std::string verdict = "Passed";
    unsigned errors = 0;
    std::vector<float> cpuResult(tb.elements_per_plf());
    for (size_t i = 0; i < tb.plf_calls && errors < 20; ++i) {
        long long inc_cpu = 0;
        reference_ms.t1[i] = t.elapsed_ms();
        if (tb.states == 4)
            golden_plf(alignmentsleft.data(), alignmentsright.data(), cpuResult.data(), ev, tb.alignment_sites, branchleft,
                       branchright, wgt.data(), inc_cpu);
        else
            golden_plf_states(tb.states, alignmentsleft.data(), alignmentsright.data(), cpuResult.data(), ev, tb.alignment_sites,
                              branchleft, branchright, wgt.data(), inc_cpu);
        reference_ms.t2[i] = t.elapsed_ms();
        for (size_t j = 0; j < tb.elements_per_plf(); ++j) {
            if (cpuResult[j] != result[i][j]) {
                std::cout << "ERROR: alignment data wrong for call " << i << " at alignment " << (j / SF) << ", probability "
                          << (j % SF) << ", cpu!=b200: " << cpuResult[j] << "!=" << result[i][j] << std::endl;
                if (++errors >= 20) break;
            }
        }
        if (inc_cpu != scalerIncrement[i] || inc_cpu != scalerIncrementFused[i]) {
            std::cout << "ERROR: scalerIncrement wrong for call " << i << ", cpu / host-reduced / kernel-fused: " << inc_cpu << " / "
                      << scalerIncrement[i] << " / " << scalerIncrementFused[i] << std::endl;
            ++errors;
        }
    }
    if (errors) {
        verdict = errors >= 20 ? " Failed with more than 20 errors" : " Failed with " + std::to_string(errors) + " errors";
        exit_code = 1;
    }
    std::cout << std::endl << "Test result: " << verdict << std::endl;
    std::cout << "scalerIncrement (call 0): " << scalerIncrement[0] << std::endl;
#endif

    // ---- Result ---------------------------------------------------------------------------------------
    const double total_sites = static_cast<double>(tb.alignment_sites) * tb.plf_calls;
    const bool csv = std::getenv("PLF_WRITE_CSV") != nullptr;
    const std::string csv_name = "plf_" + cfg.aie_name + "_" + cfg.pl_name + "_plfs" + std::to_string(tb.plf_calls) + "_alignments" +
                                 std::to_string(tb.alignment_sites) + "_usedgraphs" + std::to_string(tb.parallel_instances) + ".csv";
#if !defined(NO_INTERMEDIATE_RESULTS) || NO_INTERMEDIATE_RESULTS == 0
    print_timing_data(execution_ms[0], reference_ms, static_cast<double>(tb.data_size()), total_sites, tb.plf_calls, "B200",
                      static_cast<double>(tb.alignments_per_instance(0)) * tb.plf_calls, 3.0 * SF * 4 + 1);
    if (csv) write_to_csv(csv_name, execution_ms);
#else
    {
        const std::string line(101, '=');
        std::cout << std::endl << line << std::endl;
        std::cout << "| Timing region                          | time (ms)  | bandwidth (MB/s) |         bandwidth (MA/s) |" << std::endl;
        std::cout << line << std::endl;
        print_row("Prepare input for B200:", execution_ms.hm(), static_cast<double>(tb.data_size()), total_sites);
        print_row("PLF on B200 (incl. transfers):", execution_ms.msm(), static_cast<double>(tb.data_size()), total_sites);
        print_row("scaling wgt mult:", execution_ms.mh(), static_cast<double>(tb.data_size()), total_sites);
        std::cout << line << std::endl;
    }
    if (csv) write_to_csv(csv_name, execution_ms);
#endif

    // ---- Cleanup ---------------------------------------------------------------------------------------
    for (size_t i = 0; i < tb.plf_calls; ++i) {
        plf_host_free(result[i]);
        plf_host_free(scalerVector[i]);
    }
    for (unsigned k = 0; k < tb.parallel_instances; ++k) {
        plf_host_free(dataLeft[k]);
        plf_host_free(dataRight[k]);
    }
    plf_multi_destroy(multi);
    return exit_code;
}
