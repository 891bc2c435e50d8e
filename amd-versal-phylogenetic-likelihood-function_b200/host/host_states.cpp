// host/host_states.cpp -- host test bench for the STATES knob.
//
// Not a file of the reference, which builds STATES=DNA only (Makefile:31, README.md:67) and lists "Implement
// protein-based PLF" as an open to-do (README.md:202).  It is what host_mem's flow (app/src/host_mem.cpp:
// stimulus -> write -> run -> read -> host scaler reduction -> exact verification against the CPU golden) looks
// like with the state count taken from the configuration name: ...DNA... = 4 states, ...AA... = 20 states.
// Device buffers are caller-owned (plf_device_malloc) and the run step is plf_newview_states_device().
//
//   host_states.exe <config name> <device ordinal | PCI BDF> <sites> <plf calls>
//   e.g. host_states.exe plf_128x9AAwindow8192Comb_memAAwindowComb 0 1000000 3
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
    std::cerr << "host_states: " << msg << std::endl;
    std::exit(2);
}

void check(int rc, const char *what)
{
    if (rc != PLF_OK) die(std::string(what) + ": " + plf_last_error(nullptr));
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

template <class T>
T *device(int dev, size_t count)
{
    void *p = nullptr;
    if (plf_device_malloc(dev, &p, (count ? count : 1) * sizeof(T)) != PLF_OK) die(plf_last_error(nullptr));
    return static_cast<T *>(p);
}

}  // namespace

int main(int argc, char *argv[])
{
    if (argc != 5) {
        std::cerr << "Usage: " << argv[0] << " <config name> <device ordinal | PCI BDF> <number of alignments> <number of plf calls>"
                  << std::endl;
        return 2;
    }
    AcceleratorConfig cfg;
    try {
        cfg = parse_config(argv[1]);
    } catch (const std::exception &e) {
        die(e.what());
    }
    if (cfg.input_src != PLF_INPUT_MEM) die("host_states reads CLVs from host memory: use an INPUT_SRC=mem configuration");
    int dev = 0;
    if (plf_device_from_string(argv[2], &dev) != PLF_OK) die(plf_last_error(nullptr));
    const size_t n = parse_count(argv[3], "number of alignments");
    const size_t calls = parse_count(argv[4], "number of plf calls");
    if (n == 0 || calls == 0) die("alignments and plf calls must be > 0");
    const unsigned S = cfg.n_states;
    const size_t site = 4 * static_cast<size_t>(S), mat = static_cast<size_t>(S) * S;

    char name[256], bdf[32];
    if (plf_device_info(dev, name, sizeof name, bdf, sizeof bdf) != PLF_OK) die(plf_last_error(nullptr));
    std::cout << "| test name:        plf with STATES=" << cfg.states << " (" << S << " states x 4 rate categories, B200 / CUDA sm_100a)" << std::endl;
    std::cout << "| alignment sites:  " << n << "   plf calls: " << calls << "   bytes per site: " << 3 * site * 4 + 1 << std::endl;
    std::cout << "| device:           " << name << " [" << bdf << "]" << std::endl;

    // stimulus: the reference's recipe (host_mem.cpp:179-209), seeded.  The tiny factor of every 4th site's left
    // CLV is 1e-12 for DNA as in the reference and 1e-14 for AA (20-term sums are 25x larger).
    const char *seed_env = std::getenv("PLF_SEED");
    std::mt19937 gen(seed_env ? static_cast<uint32_t>(std::strtoul(seed_env, nullptr, 10)) : 42u);
    std::uniform_real_distribution<> dis(0.0, 1.0);
    std::vector<float> ev(mat), branchleft(4 * mat), branchright(4 * mat);
    for (float &v : ev) v = static_cast<float>(dis(gen));
    for (size_t j = 0; j < 4 * mat; ++j) {
        branchleft[j] = static_cast<float>(dis(gen));
        branchright[j] = static_cast<float>(dis(gen));
    }
    float *x1 = pinned<float>(n * site), *x2 = pinned<float>(n * site), *x3 = pinned<float>(n * site);
    char *scaler = pinned<char>(n);
    unsigned long long *h_sum = pinned<unsigned long long>(1);
    const float tiny_scale = S == 4 ? 1.0e-12f : 1.0e-14f;
    for (size_t j = 0; j < n * site; ++j) {
        x1[j] = static_cast<float>(dis(gen) * (((j / site) % 4 == 0) ? tiny_scale : 1.0f));
        x2[j] = static_cast<float>(dis(gen));
    }
    std::vector<int> wgt(n, 1);

    float *d1 = device<float>(dev, n * site), *d2 = device<float>(dev, n * site), *d3 = device<float>(dev, n * site);
    unsigned char *dsc = device<unsigned char>(dev, n);
    unsigned long long *dsum = device<unsigned long long>(dev, 1);
    plf_launch_opts opts;
    std::memset(&opts, 0, sizeof opts);
    if (const char *m = std::getenv("PLF_MATH")) opts.math_mode = std::strcmp(m, "fma") == 0 ? PLF_MATH_FMA : PLF_MATH_STRICT;

    Timer t;
    TimingData execution_ms(calls);
    std::vector<long long> inc_fused(calls, 0), inc_host(calls, 0);
    for (size_t i = 0; i < calls; ++i) {
        execution_ms.begin[i] = t.elapsed_ms();
        check(plf_memcpy_h2d(d1, x1, n * site * sizeof(float), nullptr), "write left");
        check(plf_memcpy_h2d(d2, x2, n * site * sizeof(float), nullptr), "write right");
        check(plf_memset_device(dsum, 0, sizeof(unsigned long long), nullptr), "clear scaler sum");
        check(plf_stream_sync(nullptr), "sync");
        execution_ms.t1[i] = t.elapsed_ms();
        check(plf_newview_states_device(static_cast<int>(S), d1, d2, d3, dsc, ev.data(), branchleft.data(), branchright.data(), nullptr, n,
                                        dsum, &opts, nullptr),
              "plf_newview_states_device");
        check(plf_stream_sync(nullptr), "run");
        execution_ms.t2[i] = t.elapsed_ms();
        check(plf_memcpy_d2h(x3, d3, n * site * sizeof(float), nullptr), "read out");
        check(plf_memcpy_d2h(scaler, dsc, n, nullptr), "read scaler");
        check(plf_memcpy_d2h(h_sum, dsum, sizeof(unsigned long long), nullptr), "read scaler sum");
        check(plf_stream_sync(nullptr), "sync");
        inc_fused[i] = static_cast<long long>(*h_sum);
        for (size_t j = 0; j < n; ++j) inc_host[i] += static_cast<long long>(scaler[j]) * wgt[j];     // host_mem.cpp:384-388
        execution_ms.end[i] = t.elapsed_ms();
    }

    int exit_code = 0;
    std::vector<float> cpu(n * site);
    long long inc_cpu = 0;
    const double g0 = t.elapsed_ms();
    golden_plf_states(S, x1, x2, cpu.data(), ev.data(), n, branchleft.data(), branchright.data(), wgt.data(), inc_cpu);
    const double g1 = t.elapsed_ms();
    unsigned errors = 0;
    const bool exact = opts.math_mode == PLF_MATH_STRICT;
    for (size_t j = 0; j < n * site && errors < 20; ++j) {
        const bool bad = exact ? cpu[j] != x3[j] : std::fabs(cpu[j] - x3[j]) > 1e-5f * std::fabs(cpu[j]);
        if (bad) {
            std::cout << "ERROR: alignment data wrong at alignment " << (j / site) << ", probability " << (j % site) << ", cpu!=b200: "
                      << cpu[j] << "!=" << x3[j] << std::endl;
            ++errors;
        }
    }
    if (inc_cpu != inc_fused.back() || inc_cpu != inc_host.back()) {
        std::cout << "ERROR: scalerIncrement cpu / host-reduced / kernel-fused: " << inc_cpu << " / " << inc_host.back() << " / "
                  << inc_fused.back() << std::endl;
        ++errors;
    }
    std::cout << std::endl << "Test result: " << (errors ? " Failed" : "Passed") << (exact ? " (exact compare)" : " (1e-5 relative)") << std::endl;
    std::cout << "scalerIncrement (last call): " << inc_fused.back() << std::endl;
    exit_code = errors ? 1 : 0;

    const double total_sites = static_cast<double>(n) * calls;
    const double bytes = total_sites * site * 4;
    const std::string line(101, '=');
    std::cout << std::endl << line << std::endl;
    std::cout << "| Timing region                          | time (ms)  | bandwidth (MB/s) |         bandwidth (MA/s) |" << std::endl;
    std::cout << line << std::endl;
    print_row("host -> device:", execution_ms.hm(), 2 * bytes, total_sites);
    print_row("PLF kernel:", execution_ms.msm(), bytes, total_sites);
    print_row("  - fastest call:", execution_ms.min_msm(), bytes / calls, total_sites / calls);
    print_row("device -> host + scaling wgt mult:", execution_ms.mh(), bytes, total_sites);
    print_row("Reference (CPU golden, 1 thread):", g1 - g0, bytes / calls, total_sites / calls);
    std::cout << line << std::endl;
    std::cout << "| kernel, fastest call: " << n / (execution_ms.min_msm() / 1e3) / 1e9 << " G sites/s, "
              << (3 * site * 4 + 1) * static_cast<double>(n) / (execution_ms.min_msm() / 1e3) / 1e9 << " GB/s of algorithmic traffic" << std::endl;

    for (void *p : {static_cast<void *>(x1), static_cast<void *>(x2), static_cast<void *>(x3), static_cast<void *>(scaler),
                    static_cast<void *>(h_sum)})
        plf_host_free(p);
    for (void *p : {static_cast<void *>(d1), static_cast<void *>(d2), static_cast<void *>(d3), static_cast<void *>(dsc),
                    static_cast<void *>(dsum)})
        plf_device_free(p);
    return exit_code;
}
