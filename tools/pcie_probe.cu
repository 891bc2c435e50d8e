// tools/pcie_probe.cu -- bare host<->device copy ceiling of the box the bench runs on (developer tool).
//
// The end-to-end leg of bench.py moves 128 B/site host->device and 65 B/site device->host through the C ABI.
// This program measures what the platform gives a process that does NOTHING else: plain cudaMemcpyAsync
// between pinned host memory and device memory, no kernels, one process per GPU (like torchrun), all GPUs
// released together from a shared-memory barrier.  It prints one JSON line per configuration:
//
//   pcie_probe --gpus 0,1,2,3 --mode duplex --mb 512 --reps 8 --alloc pinned|portable|wc|huge|thp
//              [--streams 1] [--ratio 128:65] [--stagger-us 0]
//
//   mode   h2d | d2h | duplex (both directions at once, byte ratio --ratio as in the PLF round trip)
//   alloc  pinned   cudaHostAlloc(default)             portable  cudaHostAllocPortable
//          wc       cudaHostAllocWriteCombined (H2D source only; the D2H target stays cacheable)
//          huge     mmap(MAP_HUGETLB) + cudaHostRegister     thp   mmap + madvise(MADV_HUGEPAGE) + cudaHostRegister
//   streams  copy streams per direction; each copy is cut into that many pieces issued round-robin
//
// One cudaMemcpyAsync per copy, nothing batched.  Build: nvcc -O2 -o build/pcie_probe tools/pcie_probe.cu
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#ifndef MAP_HUGE_2MB
#define MAP_HUGE_2MB (21 << 26)
#endif

struct Shared {
    std::atomic<int> arrived[8];
    double h2d_gbs[64], d2h_gbs[64], wall_s[64];
    int ok[64];
    char note[64][96];
};

static void barrier(Shared *sh, int phase, int n)
{
    sh->arrived[phase].fetch_add(1);
    while (sh->arrived[phase].load() < n) usleep(50);
}

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e__ = (x);                                                                   \
        if (e__ != cudaSuccess) {                                                                \
            snprintf(sh->note[idx], sizeof sh->note[idx], "%s: %s", #x, cudaGetErrorString(e__)); \
            sh->ok[idx] = 0;                                                                     \
            for (int p__ = phase_done; p__ < 3; ++p__) barrier(sh, p__, nproc);                  \
            _exit(0);                                                                            \
        }                                                                                        \
    } while (0)

static void *host_buffer(const std::string &alloc, size_t bytes, bool h2d_source, std::string *how)
{
    void *p = nullptr;
    if (alloc == "huge" || alloc == "thp") {
        const size_t two_mb = (size_t)2 << 20;
        const size_t len = (bytes + two_mb - 1) & ~(two_mb - 1);
        if (alloc == "huge") {
            p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB | MAP_HUGE_2MB, -1, 0);
            if (p == MAP_FAILED) {
                *how = "hugetlb unavailable -> thp";
                p = nullptr;
            } else {
                *how = "hugetlb";
            }
        }
        if (!p) {
            p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) return nullptr;
            madvise(p, len, MADV_HUGEPAGE);
            if (how->empty()) *how = "thp";
        }
        memset(p, 1, len);
        if (cudaHostRegister(p, len, cudaHostRegisterDefault) != cudaSuccess) return nullptr;
        return p;
    }
    unsigned flags = cudaHostAllocDefault;
    if (alloc == "portable") flags = cudaHostAllocPortable;
    if (alloc == "wc" && h2d_source) flags = cudaHostAllocWriteCombined;
    if (cudaHostAlloc(&p, bytes, flags) != cudaSuccess) return nullptr;
    *how = alloc;
    if (!(alloc == "wc" && h2d_source)) memset(p, 1, bytes);
    return p;
}

int main(int argc, char **argv)
{
    std::vector<int> gpus = {0};
    std::string mode = "duplex", alloc = "pinned";
    size_t mb = 512;
    int reps = 8, streams = 1, ratio_in = 128, ratio_out = 65, stagger_us = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--gpus") {
            gpus.clear();
            std::string v = next();
            size_t pos = 0;
            while (pos < v.size()) {
                size_t c = v.find(',', pos);
                if (c == std::string::npos) c = v.size();
                gpus.push_back(atoi(v.substr(pos, c - pos).c_str()));
                pos = c + 1;
            }
        } else if (a == "--mode") mode = next();
        else if (a == "--alloc") alloc = next();
        else if (a == "--mb") mb = (size_t)atoll(next().c_str());
        else if (a == "--reps") reps = atoi(next().c_str());
        else if (a == "--streams") streams = atoi(next().c_str());
        else if (a == "--stagger-us") stagger_us = atoi(next().c_str());
        else if (a == "--ratio") {
            std::string v = next();
            sscanf(v.c_str(), "%d:%d", &ratio_in, &ratio_out);
        } else {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    const int nproc = (int)gpus.size();
    if (nproc < 1 || nproc > 64 || streams < 1 || streams > 16) return 2;
    Shared *sh = static_cast<Shared *>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (sh == MAP_FAILED) return 3;
    memset(sh, 0, sizeof *sh);

    const bool do_in = mode != "d2h", do_out = mode != "h2d";
    const size_t in_bytes = do_in ? (mb << 20) : 0;
    const size_t out_bytes = do_out ? (mode == "duplex" ? (mb << 20) * (size_t)ratio_out / (size_t)ratio_in : (mb << 20)) : 0;

    std::vector<pid_t> kids;
    for (int idx = 0; idx < nproc; ++idx) {
        pid_t pid = fork();
        if (pid != 0) {
            kids.push_back(pid);
            continue;
        }
        // ---- child: one process per GPU, CUDA initialised after the fork ----
        int phase_done = 0;
        sh->ok[idx] = 1;
        CK(cudaSetDevice(gpus[idx]));
        std::string how_in, how_out;
        char *h_in = do_in ? static_cast<char *>(host_buffer(alloc, in_bytes, true, &how_in)) : nullptr;
        char *h_out = do_out ? static_cast<char *>(host_buffer(alloc, out_bytes, false, &how_out)) : nullptr;
        if ((do_in && !h_in) || (do_out && !h_out)) {
            snprintf(sh->note[idx], sizeof sh->note[idx], "host allocation (%s) failed", alloc.c_str());
            sh->ok[idx] = 0;
            for (int p = 0; p < 3; ++p) barrier(sh, p, nproc);
            _exit(0);
        }
        snprintf(sh->note[idx], sizeof sh->note[idx], "%s", (do_in ? how_in : how_out).c_str());
        char *d_in = nullptr, *d_out = nullptr;
        if (do_in) CK(cudaMalloc(&d_in, in_bytes));
        if (do_out) CK(cudaMalloc(&d_out, out_bytes));
        if (do_out) CK(cudaMemset(d_out, 2, out_bytes));
        std::vector<cudaStream_t> s_in(streams), s_out(streams);
        for (int k = 0; k < streams; ++k) {
            CK(cudaStreamCreateWithFlags(&s_in[k], cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&s_out[k], cudaStreamNonBlocking));
        }
        cudaEvent_t e0i, e1i, e0o, e1o;
        CK(cudaEventCreate(&e0i));
        CK(cudaEventCreate(&e1i));
        CK(cudaEventCreate(&e0o));
        CK(cudaEventCreate(&e1o));
        auto round = [&]() {
            for (int k = 0; k < streams; ++k) {
                const size_t ci = in_bytes / streams, co = out_bytes / streams;
                if (do_in) cudaMemcpyAsync(d_in + k * ci, h_in + k * ci, k == streams - 1 ? in_bytes - k * ci : ci, cudaMemcpyHostToDevice, s_in[k]);
                if (do_out) cudaMemcpyAsync(h_out + k * co, d_out + k * co, k == streams - 1 ? out_bytes - k * co : co, cudaMemcpyDeviceToHost, s_out[k]);
            }
        };
        round();                                   // warm-up: first touch of every page by the DMA engines
        CK(cudaDeviceSynchronize());
        barrier(sh, 0, nproc);
        phase_done = 1;
        if (stagger_us > 0) usleep((useconds_t)stagger_us * idx);
        const auto t0 = std::chrono::steady_clock::now();
        if (streams == 1) {
            if (do_in) cudaEventRecord(e0i, s_in[0]);
            if (do_out) cudaEventRecord(e0o, s_out[0]);
        }
        for (int r = 0; r < reps; ++r) round();
        if (streams == 1) {
            if (do_in) cudaEventRecord(e1i, s_in[0]);
            if (do_out) cudaEventRecord(e1o, s_out[0]);
        }
        CK(cudaDeviceSynchronize());
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        barrier(sh, 1, nproc);
        phase_done = 2;
        float ms_in = 0.f, ms_out = 0.f;
        if (streams == 1) {
            if (do_in) cudaEventElapsedTime(&ms_in, e0i, e1i);
            if (do_out) cudaEventElapsedTime(&ms_out, e0o, e1o);
        } else {
            ms_in = ms_out = (float)(wall * 1e3);
        }
        sh->wall_s[idx] = wall;
        sh->h2d_gbs[idx] = do_in ? (double)in_bytes * reps / (ms_in * 1e-3) / 1e9 : 0.0;
        sh->d2h_gbs[idx] = do_out ? (double)out_bytes * reps / (ms_out * 1e-3) / 1e9 : 0.0;
        // every byte arrived: the D2H target holds the device pattern
        if (do_out && (h_out[0] != 2 || h_out[out_bytes - 1] != 2)) sh->ok[idx] = 0;
        barrier(sh, 2, nproc);
        _exit(0);
    }
    int status = 0;
    for (pid_t k : kids) waitpid(k, &status, 0);

    double wall_max = 0.0, sum_in = 0.0, sum_out = 0.0;
    bool ok = true;
    for (int i = 0; i < nproc; ++i) {
        wall_max = wall_max > sh->wall_s[i] ? wall_max : sh->wall_s[i];
        ok = ok && sh->ok[i];
    }
    printf("{\"tool\": \"pcie_probe\", \"gpus\": %d, \"mode\": \"%s\", \"alloc\": \"%s\", \"alloc_effective\": \"%s\", \"mb_h2d\": %zu, \"mb_d2h\": %zu, "
           "\"reps\": %d, \"streams\": %d, \"stagger_us\": %d, \"ok\": %s, \"per_gpu\": [",
           nproc, mode.c_str(), alloc.c_str(), sh->note[0], in_bytes >> 20, out_bytes >> 20, reps, streams, stagger_us, ok ? "true" : "false");
    for (int i = 0; i < nproc; ++i) {
        printf("%s{\"gpu\": %d, \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"wall_s\": %.4f}", i ? ", " : "", gpus[i], sh->h2d_gbs[i], sh->d2h_gbs[i], sh->wall_s[i]);
        sum_in += sh->h2d_gbs[i];
        sum_out += sh->d2h_gbs[i];
    }
    const double total_gb = (double)(in_bytes + out_bytes) * reps * nproc / 1e9;
    printf("], \"aggregate_gbs_wall\": %.2f, \"sum_h2d_gbs\": %.2f, \"sum_d2h_gbs\": %.2f, \"wall_max_s\": %.4f}\n",
           wall_max > 0 ? total_gb / wall_max : 0.0, sum_in, sum_out, wall_max);
    return ok ? 0 : 1;
}
