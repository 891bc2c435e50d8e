// host/plfb_file.h -- the packed PLF buffers as files.
//
// Not a format of the reference (it only ever holds these buffers in memory): a 64-byte header followed by exactly
// the bytes the reference's host hands to bo.write / gets from bo.read --
//   LEFT   [EV16 | P_left64 | CLV n*16] floats                          (app/src/host_mem.cpp:231-233)
//   RIGHT  Comb: [EV16 | P_right64 | CLV]   Sep: [P_right64 | CLV]      (host_mem.cpp:234-241)
//   OUT    CLV n*16 floats                                               (host_mem.cpp:313)
//   SCALER n bytes, 0/1                                                  (host_mem.cpp:314)
// little-endian, fp32.  host_mem.exe writes them with PLF_DUMP_DIR=<dir> and reads its stimulus from them with
// PLF_LOAD_DIR=<dir>; the Python package reads and writes the same files (save_plfb / load_plfb).
#ifndef PLFHOST_PLFB_FILE_H
#define PLFHOST_PLFB_FILE_H

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace plfhost {

enum PlfbKind : uint32_t { PLFB_LEFT = 0, PLFB_RIGHT = 1, PLFB_OUT = 2, PLFB_SCALER = 3 };

struct PlfbHeader {
    char magic[4];            // "PLFB"
    uint32_t version;         // 1
    uint32_t kind;            // PlfbKind
    uint32_t layout;          // 0 Comb, 1 Sep (meaningful for RIGHT)
    uint32_t states;          // 4
    uint32_t categories;      // 4
    uint64_t sites;
    uint64_t payload_bytes;
    uint8_t reserved[24];
};
static_assert(sizeof(PlfbHeader) == 64, "PLFB header is 64 bytes");

inline size_t plfb_payload_bytes(uint32_t kind, uint32_t layout, uint64_t sites)
{
    switch (kind) {
    case PLFB_LEFT: return (80 + 16 * sites) * sizeof(float);
    case PLFB_RIGHT: return ((layout == 0 ? 80 : 64) + 16 * sites) * sizeof(float);
    case PLFB_OUT: return 16 * sites * sizeof(float);
    case PLFB_SCALER: return sites;
    default: throw std::runtime_error("plfb: unknown buffer kind " + std::to_string(kind));
    }
}

inline void write_plfb(const std::string &path, uint32_t kind, uint32_t layout, uint64_t sites, const void *payload)
{
    PlfbHeader h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.magic, "PLFB", 4);
    h.version = 1;
    h.kind = kind;
    h.layout = layout;
    h.states = 4;
    h.categories = 4;
    h.sites = sites;
    h.payload_bytes = plfb_payload_bytes(kind, layout, sites);
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("plfb: cannot open " + path + " for writing");
    f.write(reinterpret_cast<const char *>(&h), sizeof h);
    f.write(static_cast<const char *>(payload), static_cast<std::streamsize>(h.payload_bytes));
    if (!f) throw std::runtime_error("plfb: short write to " + path);
}

// Reads and validates a file; the payload lands in `bytes`.  Throws std::runtime_error with the reason.
inline PlfbHeader read_plfb(const std::string &path, std::vector<char> &bytes)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("plfb: cannot open " + path);
    PlfbHeader h;
    f.read(reinterpret_cast<char *>(&h), sizeof h);
    if (!f || std::memcmp(h.magic, "PLFB", 4) != 0) throw std::runtime_error("plfb: " + path + " is not a PLFB file");
    if (h.version != 1) throw std::runtime_error("plfb: " + path + ": unsupported version " + std::to_string(h.version));
    if (h.states != 4 || h.categories != 4) throw std::runtime_error("plfb: " + path + ": only 4 states x 4 categories");
    if (h.payload_bytes != plfb_payload_bytes(h.kind, h.layout, h.sites))
        throw std::runtime_error("plfb: " + path + ": payload size does not match kind/layout/sites");
    bytes.resize(h.payload_bytes);
    f.read(bytes.data(), static_cast<std::streamsize>(h.payload_bytes));
    if (static_cast<uint64_t>(f.gcount()) != h.payload_bytes) throw std::runtime_error("plfb: " + path + " is truncated");
    return h;
}

}  // namespace plfhost
#endif
