// host/host_stream.cpp -- host test bench for the STREAMED round trip.
//
// Not a file of the reference: it is the analogue of host_mem's NO_INTERMEDIATE_RESULTS=1 round-trip mode
// (app/src/host_mem.cpp:327-382) on top of plf_newview_stream().  The host keeps the plain, unpacked
// arrays that plf() takes (plf.h:1-5); the library cuts the site range into chunks and overlaps H2D,
// kernel and D2H.  No packing pass, no device-memory limit on the site count.  Same stimulus recipe,
// same exact verification against the CPU golden as host_mem; the configuration name selects the state count
// (...DNA...: 16-float sites, ...AA...: 80-float sites).
//
//   host_stream.exe <config name> <device ordinal | PCI BDF> <sites> <plf calls> [chunk sites]
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "b200plf.h"
#include "golden_plf.h"
#include "tb_info.h"
#include "timing_report.h"

using namespace plfhost;

namespace {

[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "host_stream: " << msg << std::endl;
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

template <class T>
T *pinned(size_t count)
{
    void *p = nullptr;
    if (plf_host_alloc(&p, (count ? count : 1) * sizeof(T)) != PLF_OK) die(plf_last_error(nullptr));
    return static_cast<T *>(p);
}

}  // namespace

int main(int argc, char *argv[])
{
    if (argc != 5 && argc != 6) {
        std::cerr << "Usage: " << argv[0]
                  << " <config name> <device ordinal | PCI BDF> <number of alignments> <number of plf calls> [chunk sites]"
                  << std::endl;
        return 2;
    }
    AcceleratorConfig cfg;
    try {
        cfg = parse_config(argv[1]);
    } catch (const std::exception &e) {
        die(e.what());
    }
    const size_t S = cfg.n_states;                         // STATES knob: DNA (4) or AA (20), from the configuration name
    const size_t SF = 4 * S;                               // floats per site
    if (cfg.input_src != PLF_INPUT_MEM) die("the streamed path reads CLVs from host memory: use an INPUT_SRC=mem configuration");
    int device = 0;
    if (plf_device_from_string(argv[2], &device) != PLF_OK) die(plf_last_error(nullptr));
    const size_t n = parse_count(argv[3], "number of alignments");
    const size_t calls = parse_count(argv[4], "number of plf calls");
    const size_t chunk = argc == 6 ? parse_count(argv[5], "chunk sites") : 0;
    if (n == 0 || calls == 0) die("alignments and plf calls must be > 0");

    char name[256], bdf[32];
    if (plf_device_info(device, name, sizeof name, bdf, sizeof bdf) != PLF_OK) die(plf_last_error(nullptr));
    std::cout << "| test name:        plf streamed round trip (B200 / CUDA sm_100a)" << std::endl;
    std::cout << "| alignment sites:  " << n << "   plf calls: " << calls << "   chunk: " << (chunk ? std::to_string(chunk) : "auto")
              << std::endl;
    std::cout << "| STATES:           " << cfg.states << " (" << S << " states x 4 rate categories, " << SF << " floats per site)" << std::endl;
    std::cout << "| host CLV bytes:   " << 3.0 * n * SF * 4 / 1e9 << " GB   device: " << name << " [" << bdf << "]" << std::endl;

    plf_ctx *ctx = nullptr;
    check(plf_ctx_create_states(&ctx, device, 1, cfg.layout, PLF_INPUT_MEM, static_cast<int>(S)), nullptr, "plf_ctx_create_states");
    if (const char *m = std::getenv("PLF_MATH"))
        check(plf_ctx_set_math(ctx, std::strcmp(m, "fma") == 0 ? PLF_MATH_FMA : PLF_MATH_STRICT), ctx, "plf_ctx_set_math");

    // stimulus: the reference's recipe (host_mem.cpp:179-209), seeded, straight into pinned arrays
    const char *seed_env = std::getenv("PLF_SEED");
    std::mt19937 gen(seed_env ? static_cast<uint32_t>(std::strtoul(seed_env, nullptr, 10)) : 42u);
    std::uniform_real_distribution<> dis(0.0, 1.0);
    std::vector<float> ev_v(S * S), bl_v(4 * S * S), br_v(4 * S * S);
    float *ev = ev_v.data(), *branchleft = bl_v.data(), *branchright = br_v.data();
    for (float &v : ev_v) v = static_cast<float>(dis(gen));
    for (size_t j = 0; j < 4 * S * S; ++j) {
        branchleft[j] = static_cast<float>(dis(gen));
        branchright[j] = static_cast<float>(dis(gen));
    }
    float *x1 = pinned<float>(n * SF), *x2 = pinned<float>(n * SF), *x3 = pinned<float>(n * SF);
    char *scaler = pinned<char>(n);
    // every fourth site tiny on the left (host_mem.cpp:198-204); with 20 states the sums have 20 terms, hence 1e-14
    const float tiny_scale = S == 4 ? static_cast<float>(std::pow(1.0e-12, 1)) : 1.0e-14f;
    for (size_t j = 0; j < n * SF; ++j) {
        x1[j] = static_cast<float>(dis(gen) * ((j % (4 * SF) < SF) ? tiny_scale : 1.0f));
        x2[j] = static_cast<float>(dis(gen));
    }
    std::vector<int> wgt(n, 1);

    Timer t;
    TimingData execution_ms(calls);
    std::vector<long long> inc(calls, 0), inc_host(calls, 0);
    for (size_t i = 0; i < calls; ++i) {
        execution_ms.begin[i] = execution_ms.t1[i] = t.elapsed_ms();          // nothing to prepare: no packing pass
        check(plf_newview_stream(ctx, ev, branchleft, branchright, x1, x2, x3, scaler, nullptr, n, chunk, &inc[i]), ctx,
              "plf_newview_stream");
        execution_ms.t2[i] = t.elapsed_ms();
        for (size_t j = 0; j < n; ++j) inc_host[i] += static_cast<long long>(scaler[j]) * wgt[j];   // host_mem.cpp:384-388
        execution_ms.end[i] = t.elapsed_ms();
    }

    int exit_code = 0;
#if !defined(NO_CORRECTNESS_CHECK) || NO_CORRECTNESS_CHECK == 0
    std::vector<float> cpu(n * SF);
    long long inc_cpu = 0;
    const double g0 = t.elapsed_ms();
    if (S == 4)
        golden_plf(x1, x2, cpu.data(), ev, n, branchleft, branchright, wgt.data(), inc_cpu);
    else
        golden_plf_states(static_cast<unsigned>(S), x1, x2, cpu.data(), ev, n, branchleft, branchright, wgt.data(), inc_cpu);
    const double g1 = t.elapsed_ms();
    unsigned errors = 0;
    for (size_t j = 0; j < n * SF && errors < 20; ++j)
        if (cpu[j] != x3[j]) {
            std::cout << "ERROR: alignment data wrong at alignment " << (j / SF) << ", probability " << (j % SF) << ", cpu!=b200: "
                      << cpu[j] << "!=" << x3[j] << std::endl;
            ++errors;
        }
    if (inc_cpu != inc.back() || inc_cpu != inc_host.back()) {
        std::cout << "ERROR: scalerIncrement cpu / host-reduced / kernel-fused: " << inc_cpu << " / " << inc_host.back() << " / "
                  << inc.back() << std::endl;
        ++errors;
    }
    std::cout << std::endl << "Test result: " << (errors ? " Failed" : "Passed") << std::endl;
    std::cout << "scalerIncrement (last call): " << inc.back() << std::endl;
    std::cout << "Reference (CPU golden, 1 thread): " << (g1 - g0) << " ms" << std::endl;
    exit_code = errors ? 1 : 0;
#endif

    const double total_sites = static_cast<double>(n) * calls;
    const double bytes = total_sites * SF * 4.0;
    const std::string line(101, '=');
    std::cout << std::endl << line << std::endl;
    std::cout << "| Timing region                          | time (ms)  | bandwidth (MB/s) |         bandwidth (MA/s) |" << std::endl;
    std::cout << line << std::endl;
    print_row("Streamed round trip (H2D+PLF+D2H):", execution_ms.msm(), bytes, total_sites);
    print_row("  - fastest call:", execution_ms.min_msm(), bytes / calls, total_sites / calls);
    print_row("scaling wgt mult (host):", execution_ms.mh(), bytes, total_sites);
    std::cout << line << std::endl;
    std::cout << "| PCIe traffic, fastest call: " << (12.0 * SF + 1.0) * n / (execution_ms.min_msm() / 1e3) / 1e9 << " GB/s (" << 8 * SF
              << " B/site in, " << 4 * SF + 1 << " B/site out)"
              << std::endl;

    plf_host_free(x1);
    plf_host_free(x2);
    plf_host_free(x3);
    plf_host_free(scaler);
    plf_ctx_destroy(ctx);
    return exit_code;
}
