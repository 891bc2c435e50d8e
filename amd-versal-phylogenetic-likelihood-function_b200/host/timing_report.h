// host/timing_report.h -- per-call timestamps and the result tables of the host test bench.
//
// Keeps the reference's four timestamps per call (begin, t1, t2, end => host->device, execute,
// device->host; app/src/timing.h:25-99) and its table rows / CSV columns (timing.h:107-194) so
// that outputs stay comparable, and adds the two columns the B200 metric needs: achieved HBM GB/s
// at 193 algorithmic bytes per site and its fraction of the 8 TB/s roofline.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

namespace plfhost {

class Timer {
public:
    double elapsed_ms() const
    {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start_).count();
    }

private:
    std::chrono::steady_clock::time_point start_ = std::chrono::steady_clock::now();
};

struct TimingData {
    std::vector<double> begin, t1, t2, end;
    explicit TimingData(size_t calls = 0) : begin(calls, 0.0), t1(calls, 0.0), t2(calls, 0.0), end(calls, 0.0) {}
    size_t calls() const { return begin.size(); }
    double hm(size_t i) const { return t1[i] - begin[i]; }     // host -> device memory
    double msm(size_t i) const { return t2[i] - t1[i]; }       // device memory -> kernel -> device memory
    double mh(size_t i) const { return end[i] - t2[i]; }       // device memory -> host
    template <class F>
    double sum(F f) const
    {
        double s = 0;
        for (size_t i = 0; i < calls(); ++i) s += f(i);
        return s;
    }
    double hm() const { return sum([&](size_t i) { return hm(i); }); }
    double msm() const { return sum([&](size_t i) { return msm(i); }); }
    double mh() const { return sum([&](size_t i) { return mh(i); }); }
    double total() const { return calls() ? end.back() - begin.front() : 0.0; }
    double max_msm() const
    {
        double m = 0;
        for (size_t i = 0; i < calls(); ++i) m = std::max(m, msm(i));
        return m;
    }
    double min_msm() const
    {
        double m = calls() ? msm(0) : 0;
        for (size_t i = 0; i < calls(); ++i) m = std::min(m, msm(i));
        return m;
    }
};

inline double bandwidth_MBs(double ms, double bytes) { return ms > 0 ? (bytes / 1e6) / (ms / 1e3) : 0.0; }
inline double bandwidth_MAs(double ms, double sites) { return ms > 0 ? sites / (ms / 1e3) / 1e6 : 0.0; }

constexpr double kBytesPerSite = 193.0;       // 64 + 64 read, 64 + 1 written
constexpr double kHbmRooflineGBs = 8000.0;    // the figure BASELINE.json's metric normalises by

inline void print_row(const std::string &label, double ms, double bytes, double sites)
{
    std::cout << "| " << std::left << std::setw(38) << label << " | " << std::right << std::setw(10) << ms << " | "
              << std::setw(16) << bandwidth_MBs(ms, bytes) << " | " << std::setw(24) << bandwidth_MAs(ms, sites)
              << " |" << std::endl;
}

// d: accelerator timings, r: CPU golden timings (msm() only), data_size: result-CLV bytes over all
// calls (timing.h:101-103 semantics), total_sites: sites x calls.
// kernel_sites: sites x calls the TIMED instance's kernel processed (the table rows follow the reference and divide the
// whole run's sites by instance 0's time; the HBM lines must not: with several instances that overstates the traffic).
inline void print_timing_data(const TimingData &d, const TimingData &r, double data_size, double total_sites,
                              size_t calls, const char *device_label = "B200", double kernel_sites = -1.0,
                              double bytes_per_site = kBytesPerSite)
{
    if (kernel_sites < 0.0) kernel_sites = total_sites;
    const std::string bar(101, '=');
    std::cout << std::endl << bar << std::endl;
    std::cout << "| Timing region                          | time (ms)  | bandwidth (MB/s) |         bandwidth (MA/s) |"
              << std::endl << bar << std::endl;
    print_row(std::string("Host to ") + device_label + " memory:", d.hm(), data_size, total_sites);
    if (d.msm() > 0.0) {
        print_row(std::string(device_label) + " memory to kernel to memory:", d.msm(), data_size, total_sites);
        print_row("  - slowest:", d.max_msm(), data_size / calls, total_sites / calls);
        print_row("  - fastest:", d.min_msm(), data_size / calls, total_sites / calls);
    }
    print_row(std::string(device_label) + " memory to host:", d.mh(), data_size, total_sites);
    std::cout << "|----------------------------------------+------------+------------------+--------------------------|"
              << std::endl;
    print_row("Total execution time:", d.total(), data_size, total_sites);
    std::cout << bar << std::endl;
    if (d.msm() > 0.0) {
        const double gbs = bytes_per_site * kernel_sites / (d.msm() / 1e3) / 1e9;
        const double best = bytes_per_site * (kernel_sites / calls) / (d.min_msm() / 1e3) / 1e9;
        std::cout << "| HBM traffic (" << bytes_per_site << " B/site), all calls:   | " << std::setw(10) << gbs << " GB/s = " << std::setw(6)
                  << 100.0 * gbs / kHbmRooflineGBs << " % of 8 TB/s" << std::endl;
        std::cout << "| HBM traffic (" << bytes_per_site << " B/site), fastest call:| " << std::setw(10) << best << " GB/s = " << std::setw(6)
                  << 100.0 * best / kHbmRooflineGBs << " % of 8 TB/s" << std::endl;
        std::cout << bar << std::endl;
    }
    if (r.calls() && r.msm() > 0.0) {
        std::cout << std::endl << bar << std::endl;
        print_row("Reference (CPU golden, 1 thread):", r.msm(), data_size, total_sites);
        std::cout << "|----------------------------------------+------------+------------------+--------------------------|"
                  << std::endl;
        std::cout << "| Speed up (excluding pcie transfer):    | " << std::setw(56) << r.msm() / d.msm() << " |" << std::endl;
        std::cout << "| Speed up (including pcie transfer):    | " << std::setw(56) << r.msm() / d.total() << " |" << std::endl;
        std::cout << bar << std::endl;
    }
    std::cout << std::endl;
}

// Per call, per instance: hm<i>, msasm<i>, mh<i> columns (timing.h:162-194).
inline void write_to_csv(const std::string &file, const std::vector<TimingData> &d)
{
    std::ofstream out(file);
    if (!out) {
        std::cerr << "cannot write " << file << std::endl;
        return;
    }
    const char *cols[3] = {"hm", "msasm", "mh"};
    for (int c = 0; c < 3; ++c)
        for (size_t i = 0; i < d.size(); ++i)
            out << cols[c] << i << ((c == 2 && i + 1 == d.size()) ? "\n" : ",");
    for (size_t call = 0; call < (d.empty() ? 0 : d[0].calls()); ++call) {
        for (size_t i = 0; i < d.size(); ++i) out << d[i].hm(call) << ",";
        for (size_t i = 0; i < d.size(); ++i) out << d[i].msm(call) << ",";
        for (size_t i = 0; i < d.size(); ++i) out << d[i].mh(call) << (i + 1 == d.size() ? "\n" : ",");
    }
}

// Single-timeline variant (NO_INTERMEDIATE_RESULTS / gen mode): preparation,plf,scaling (timing.h:153-160).
inline void write_to_csv(const std::string &file, const TimingData &d)
{
    std::ofstream out(file);
    if (!out) {
        std::cerr << "cannot write " << file << std::endl;
        return;
    }
    out << "preparation,plf,scaling\n";
    for (size_t call = 0; call < d.calls(); ++call) out << d.hm(call) << "," << d.msm(call) << "," << d.mh(call) << "\n";
}

}  // namespace plfhost
