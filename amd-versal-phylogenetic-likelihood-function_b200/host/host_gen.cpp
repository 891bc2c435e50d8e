// host/host_gen.cpp -- host test bench, INPUT_SRC=gen.  Drop-in for the reference's
// app/src/host_gen.cpp:11-197: same five positional arguments, no input buffers, no verification,
// timing only.  Every instance runs the gen kernel, which synthesises the movers' constant site
// pattern in registers (hls/src/mm2s{left,right}_genDNAwindowComb.cpp) instead of reading CLVs.
//
//   host_gen.exe <config name | x.xclbin> <device> <sites> <plf calls> <instances used>
//
// PLF_GEN_SINK=discard reproduces the reference's sink exactly (s2mm_genDNAwindowComb.cpp:15-53
// reads the result streams and drops them; here they are folded into a checksum); the default
// writes the CLV and scaler bytes so that the output side is exercised too.
// Note: every instance is given ceil(sites/instances) sites, as host_gen.cpp:87-99 does.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "b200plf.h"
#include "tb_info.h"
#include "timing_report.h"

using namespace plfhost;

namespace {

[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "host_gen: " << msg << std::endl;
    std::exit(2);
}

void check(int rc, plf_ctx *ctx, const char *what)
{
    if (rc != PLF_OK) die(std::string(what) + ": " + plf_last_error(ctx));
}

size_t parse_count(const char *s, const char *what)
{
    char *end = nullptr;
    const unsigned long long v = std::strtoull(s, &end, 10);
    if (end == s || *end != '\0' || s[0] == '-') die(std::string("invalid ") + what + ": '" + s + "'");
    return static_cast<size_t>(v);
}

}  // namespace

int main(int argc, char *argv[])
{
    if (argc != 6) {
        std::cerr << "Not correct amount of parameters provided. Usage: " << argv[0]
                  << " <config name or /path/to/a.xclbin> <device ordinal | PCI BDF> <number of alignments>"
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
    if (cfg.n_states != 4) die("STATES=" + cfg.states + " runs through host_states.exe (this host is the DNA drop-in)");
    if (cfg.input_src != PLF_INPUT_GEN) die("configuration '" + cfg.name + "' is INPUT_SRC=mem: use host_mem.exe");
    int device = 0;
    if (plf_device_from_string(argv[2], &device) != PLF_OK) die(plf_last_error(nullptr));

    TestbenchInfo tb;
    tb.alignment_sites = parse_count(argv[3], "number of alignments");
    tb.plf_calls = parse_count(argv[4], "number of plf calls");
    tb.parallel_instances = static_cast<unsigned>(parse_count(argv[5], "parallel instances"));
    tb.window_size = cfg.window_size;
    tb.layout = cfg.layout;
    if (tb.alignment_sites == 0 || tb.plf_calls == 0 || tb.parallel_instances == 0)
        die("alignments, plf calls and instances must all be > 0");
    if (tb.parallel_instances > cfg.num_accelerators)
        die("instances used exceeds NUM_ACCELERATORS=" + std::to_string(cfg.num_accelerators));

    const char *sink_env = std::getenv("PLF_GEN_SINK");
    const int sink = (sink_env && std::strcmp(sink_env, "discard") == 0) ? PLF_GEN_DISCARD : PLF_GEN_WRITE;

    char name[256], bdf[32];
    if (plf_device_info(device, name, sizeof name, bdf, sizeof bdf) != PLF_OK) die(plf_last_error(nullptr));
    std::cout << "| test name:        plf (B200 / CUDA sm_100a), INPUT_SRC=gen" << std::endl;
    std::cout << "| PL name:          " << cfg.pl_name << std::endl;
    std::cout << "| AIE name:         " << cfg.aie_name << std::endl;
    std::cout << "| alignment sites:  " << tb.alignment_sites << std::endl;
    std::cout << "| plf calls:        " << tb.plf_calls << std::endl;
    std::cout << "| parallel plfs:    " << tb.parallel_instances << std::endl;
    std::cout << "| sink:             " << (sink == PLF_GEN_DISCARD ? "discard (checksum only)" : "write CLV + scaler") << std::endl;
    std::cout << "| device:           " << name << " [" << bdf << "]" << std::endl;

    plf_ctx *ctx = nullptr;
    check(plf_ctx_create(&ctx, device, cfg.num_accelerators, cfg.layout, PLF_INPUT_GEN), nullptr, "plf_ctx_create");
    check(plf_ctx_set_gen_sink(ctx, sink), ctx, "plf_ctx_set_gen_sink");
    if (const char *m = std::getenv("PLF_MATH"))
        check(plf_ctx_set_math(ctx, std::strcmp(m, "fma") == 0 ? PLF_MATH_FMA : PLF_MATH_STRICT), ctx, "plf_ctx_set_math");
    const size_t per_instance = tb.alignments_per_instance();
    for (unsigned k = 0; k < tb.parallel_instances; ++k)
        check(plf_instance_alloc(ctx, k, per_instance), ctx, "plf_instance_alloc");

    std::cout << "Start PLF calculation on accelerator ... " << std::endl;
    Timer t;
    TimingData execution_ms(tb.plf_calls);
    std::vector<float> kernel_ms(tb.plf_calls, 0.0f);
    for (size_t i = 0; i < tb.plf_calls; ++i) {
        execution_ms.begin[i] = t.elapsed_ms();
        execution_ms.t1[i] = t.elapsed_ms();
        for (unsigned k = 0; k < tb.parallel_instances; ++k) {
            check(plf_mark(ctx, k, PLF_MARK_T1), ctx, "plf_mark");
            check(plf_run_async(ctx, k, per_instance), ctx, "plf_run_async");
            check(plf_mark(ctx, k, PLF_MARK_T2), ctx, "plf_mark");
        }
        for (unsigned k = 0; k < tb.parallel_instances; ++k) {
            check(plf_wait(ctx, k), ctx, "plf_wait");
            float ms = 0;
            check(plf_elapsed_ms(ctx, k, PLF_MARK_T1, PLF_MARK_T2, &ms), ctx, "plf_elapsed_ms");
            kernel_ms[i] = std::max(kernel_ms[i], ms);
        }
        execution_ms.t2[i] = t.elapsed_ms();
        execution_ms.end[i] = t.elapsed_ms();
    }

    long long inc = 0;
    check(plf_scaler_increment(ctx, 0, &inc), ctx, "plf_scaler_increment");
    std::cout << "scalerIncrement (instance 0, last call): " << inc << std::endl;
    if (sink == PLF_GEN_DISCARD) {
        double chk = 0;
        check(plf_gen_checksum(ctx, 0, &chk), ctx, "plf_gen_checksum");
        std::cout << "output checksum (instance 0, last call): " << chk << std::endl;
    }
    float best = kernel_ms[0];
    for (float v : kernel_ms) best = std::min(best, v);
    std::cout << "fastest call, device time (slowest instance): " << best << " ms = "
              << per_instance * tb.parallel_instances / (best * 1e-3) / 1e9 << " G sites/s" << std::endl;

    TimingData reference_ms;   // empty: gen mode has no CPU reference (host_gen.cpp:166-169)
    const double processed = static_cast<double>(per_instance) * tb.parallel_instances * tb.plf_calls;
    print_timing_data(execution_ms, reference_ms, processed * 64.0, processed, tb.plf_calls);
    if (std::getenv("PLF_WRITE_CSV"))
        write_to_csv("plf_" + cfg.aie_name + "_" + cfg.pl_name + "_plfs" + std::to_string(tb.plf_calls) + "_alignments" +
                         std::to_string(tb.alignment_sites) + "_usedgraphs" + std::to_string(tb.parallel_instances) + ".csv",
                     execution_ms);
    plf_ctx_destroy(ctx);
    return 0;
}
