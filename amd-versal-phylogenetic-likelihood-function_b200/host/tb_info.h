// host/tb_info.h -- run description and size/partition math of the host test bench.
//
// Same quantities as the reference's testbench_info (app/src/include.h:150-268) and the same
// instance split rule (ceil(n/instances), last instance takes the remainder, :181-192), but every
// size is size_t: the reference's `unsigned int` byte counts wrap at 4 GiB (64 Mi sites x 64 B).
// The accelerator knobs are parsed from the configuration name, which follows the reference's
// artefact naming  plf_<AIE>_<PL>  with  AIE = 128x<N><STATES><window<W>|stream><Comb|Sep>  and
// PL = <mem|gen><STATES><window|stream><Comb|Sep>  (Makefile:26-39).  Unlike the reference's
// substring sniffing (include.h:44-75), unknown names are an error, not a silent default.
#pragma once

#include <cstddef>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "b200plf.h"

namespace plfhost {

struct AcceleratorConfig {
    std::string name;
    std::string aie_name, pl_name;
    unsigned num_accelerators = 9;    // NUM_ACCELERATORS
    std::string states = "DNA";       // STATES: "DNA" (4 states) or "AA" (20 states)
    unsigned n_states = 4;
    bool window = true;               // AIE_TYPE
    size_t window_size = 8192;        // WINDOW_SIZE (bytes per lane window; informational on GPU)
    int layout = PLF_LAYOUT_COMB;     // PLIO_LAYOUT
    int input_src = PLF_INPUT_MEM;    // INPUT_SRC
};

inline bool ends_with(const std::string &s, const std::string &suffix)
{
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// Accepts "plf_128x9DNAwindow8192Comb_memDNAwindowComb", the same with a directory and/or a
// ".xclbin" suffix (so existing command lines keep working), or the short forms "mem"/"gen".
inline AcceleratorConfig parse_config(const std::string &arg)
{
    AcceleratorConfig c;
    std::string s = arg;
    const size_t slash = s.find_last_of('/');
    if (slash != std::string::npos) s = s.substr(slash + 1);
    if (ends_with(s, ".xclbin")) s = s.substr(0, s.size() - 7);
    c.name = s;
    if (s == "mem" || s == "gen") {
        c.input_src = s == "gen" ? PLF_INPUT_GEN : PLF_INPUT_MEM;
        c.aie_name = "128x9DNAwindow8192Comb";
        c.pl_name = s + "DNAwindowComb";
        return c;
    }
    const size_t u1 = s.find('_');
    const size_t u2 = s.rfind('_');
    if (u1 == std::string::npos || u2 == u1)
        throw std::runtime_error("configuration name '" + arg + "' is not <app>_<AIE>_<PL>");
    c.aie_name = s.substr(u1 + 1, u2 - u1 - 1);
    c.pl_name = s.substr(u2 + 1);

    // PL part: <mem|gen><STATES><window|stream><Comb|Sep>
    if (c.pl_name.rfind("mem", 0) == 0) c.input_src = PLF_INPUT_MEM;
    else if (c.pl_name.rfind("gen", 0) == 0) c.input_src = PLF_INPUT_GEN;
    else throw std::runtime_error("PL name '" + c.pl_name + "' names neither mem nor gen input");
    if (ends_with(c.pl_name, "Comb")) c.layout = PLF_LAYOUT_COMB;
    else if (ends_with(c.pl_name, "Sep")) c.layout = PLF_LAYOUT_SEP;
    else throw std::runtime_error("PL name '" + c.pl_name + "' names neither Comb nor Sep layout");
    if (c.pl_name.find("window") != std::string::npos) c.window = true;
    else if (c.pl_name.find("stream") != std::string::npos) c.window = false;
    else throw std::runtime_error("PL name '" + c.pl_name + "' names neither window nor stream");
    if (c.pl_name.find("DNA") != std::string::npos) {
        c.states = "DNA";
        c.n_states = 4;
    } else if (c.pl_name.find("AA") != std::string::npos) {
        c.states = "AA";
        c.n_states = 20;
    } else {
        throw std::runtime_error("STATES must be DNA or AA (got '" + c.pl_name + "')");
    }

    // AIE part: 128x<N>DNA<window<W>|stream><Comb|Sep>
    const size_t x = c.aie_name.find('x');
    if (x == std::string::npos) throw std::runtime_error("AIE name '" + c.aie_name + "' has no 128x<N>");
    c.num_accelerators = static_cast<unsigned>(std::strtoul(c.aie_name.c_str() + x + 1, nullptr, 10));
    if (c.num_accelerators == 0) throw std::runtime_error("AIE name '" + c.aie_name + "' has no instance count");
    const size_t w = c.aie_name.find("window");
    if (w != std::string::npos) {
        c.window_size = std::strtoul(c.aie_name.c_str() + w + 6, nullptr, 10);
        if (c.window_size == 0) c.window_size = 1024;       // reference default (include.h:155)
    }
    return c;
}

struct TestbenchInfo {
    size_t alignment_sites = 0;
    size_t plf_calls = 1;
    unsigned parallel_instances = 1;
    size_t window_size = 8192;
    int layout = PLF_LAYOUT_COMB;
    unsigned states = 4;              // STATES: 4 (DNA) or 20 (AA); every size below scales with it
    size_t elements_per_alignment = PLF_SITE_FLOATS;     // 4 rate categories x states
    static constexpr size_t word_size = sizeof(float);

    void set_states(unsigned s)
    {
        states = s;
        elements_per_alignment = 4u * static_cast<size_t>(s);
    }
    size_t ev_elements() const { return static_cast<size_t>(states) * states; }
    size_t branch_elements() const { return 4u * static_cast<size_t>(states) * states; }

    size_t alignments_per_instance() const
    {
        return (alignment_sites + parallel_instances - 1) / parallel_instances;
    }
    size_t alignments_padding() const
    {
        return alignments_per_instance() * parallel_instances - alignment_sites;
    }
    size_t alignments_per_instance(unsigned k) const
    {
        return alignments_per_instance() - (k == parallel_instances - 1 ? alignments_padding() : 0);
    }
    size_t instance_first_site(unsigned k) const { return k * alignments_per_instance(); }
    bool split_is_valid() const
    {
        return alignment_sites > 0 && parallel_instances > 0 &&
               alignments_padding() < alignments_per_instance();
    }
    size_t header_left() const { return ev_elements() + branch_elements(); }          // [EV | P]: 80 floats for DNA
    size_t header_right() const { return layout == PLF_LAYOUT_COMB ? header_left() : branch_elements(); }
    size_t instance_elements_left() const { return alignments_per_instance() * elements_per_alignment + header_left(); }
    size_t instance_elements_right() const { return alignments_per_instance() * elements_per_alignment + header_right(); }
    size_t instance_elements_out() const { return alignments_per_instance() * elements_per_alignment; }
    size_t instance_active_elements_left(unsigned k) const { return alignments_per_instance(k) * elements_per_alignment + header_left(); }
    size_t instance_active_elements_right(unsigned k) const { return alignments_per_instance(k) * elements_per_alignment + header_right(); }
    size_t elements_per_plf() const { return alignment_sites * elements_per_alignment; }
    size_t data_elements() const { return elements_per_plf() * plf_calls; }
    size_t data_size() const { return data_elements() * word_size; }           // timing.h "data_size"
    size_t device_mem_usage() const
    {
        return (instance_elements_left() + instance_elements_right() + instance_elements_out()) * word_size *
                   parallel_instances +
               alignments_per_instance() * parallel_instances;
    }
    size_t host_mem_usage() const
    {
        return (instance_elements_left() + instance_elements_right()) * word_size * parallel_instances +
               plf_calls * (elements_per_plf() * word_size + alignment_sites) + alignment_sites * sizeof(int);
    }
};

}  // namespace plfhost
