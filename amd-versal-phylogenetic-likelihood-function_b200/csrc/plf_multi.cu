// csrc/plf_multi.cu -- several GPUs of one box from ONE host process (plf_multi_* of include/b200plf.h).
//
// The reference scales by giving contiguous site ranges to independent accelerator instances of one card
// (app/src/include.h:181-192: ceil(n / instances) sites each, the last instance takes the remainder) and sums
// the per-instance scaler bytes on the host (app/src/host_mem.cpp:384-388).  On an 8-GPU B200 box the same rule
// partitions the sites over the GPUs: sites are independent, so there is NO data-path exchange between GPUs.  The
// only collective is the optional final reduction of the path: the per-GPU scaler increments (int64) and, when the
// caller evaluates a likelihood, the per-GPU log-likelihoods (fp64), summed with ncclAllReduce over NVLink/NVSwitch.
//
// One plf_ctx per GPU (so the whole instance API applies to every rank), one host thread per GPU for the blocking
// streamed path, one NCCL communicator over the device list (ncclCommInitAll: single process, many devices).  NCCL is
// resolved with dlopen at plf_multi_create -- the library itself has no link-time dependency on it, and a process
// that already carries an NCCL (PyTorch) shares that copy.  A one-GPU plf_multi never touches NCCL.
#include "../../include/b200plf.h"
#include "plf_registry.h"

#include <nccl.h>      // types and prototypes only; the symbols come from dlopen

#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString && GetVersion; }
};

NcclApi load_nccl()
{
    NcclApi a;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) return a;
#define PLF_SYM(f) a.f = reinterpret_cast<decltype(a.f)>(dlsym(a.handle, "nccl" #f))
    PLF_SYM(CommInitAll);
    PLF_SYM(CommDestroy);
    PLF_SYM(AllReduce);
    PLF_SYM(GroupStart);
    PLF_SYM(GroupEnd);
    PLF_SYM(GetErrorString);
    PLF_SYM(GetVersion);
#undef PLF_SYM
    return a;
}

thread_local std::string g_multi_error;

}  // namespace

struct plf_multi {
    std::vector<int> devices;
    std::vector<plf_ctx *> ctx;
    std::vector<cudaStream_t> stream;           // one reduction stream per rank
    std::vector<unsigned long long *> d_inc;    // per rank: [0] scaler increment (u64)
    std::vector<double *> d_lnl;                // per rank: [0] log-likelihood
    unsigned long long *h_inc = nullptr;        // pinned staging, one slot per rank
    double *h_lnl = nullptr;
    NcclApi nccl;
    std::vector<ncclComm_t> comm;
    int nccl_version = 0;
    unsigned long long reductions = 0;
    std::string error;
};

namespace {

int mfail(plf_multi *m, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->error = buf;
    g_multi_error = buf;
    return code;
}

#define MULTI_CUDA(m, expr)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return mfail(m, e__ == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA,        \
                         "%s failed: %s", #expr, cudaGetErrorString(e__));                          \
    } while (0)

#define MULTI_NCCL(m, expr)                                                                         \
    do {                                                                                            \
        ncclResult_t r__ = (expr);                                                                  \
        if (r__ != ncclSuccess)                                                                     \
            return mfail(m, PLF_ERR_CUDA, "%s failed: %s", #expr, (m)->nccl.GetErrorString(r__));   \
    } while (0)

// In-place sum over the ranks of one u64 and one f64 per rank, both already in device memory: two all-reduces in
// one NCCL group per rank.  No-op for a single GPU.
int allreduce_device(plf_multi *m)
{
    const int g = (int)m->devices.size();
    if (g == 1) return PLF_OK;
    MULTI_NCCL(m, m->nccl.GroupStart());
    for (int r = 0; r < g; ++r) {
        MULTI_NCCL(m, m->nccl.AllReduce(m->d_inc[r], m->d_inc[r], 1, ncclUint64, ncclSum, m->comm[r], m->stream[r]));
        MULTI_NCCL(m, m->nccl.AllReduce(m->d_lnl[r], m->d_lnl[r], 1, ncclDouble, ncclSum, m->comm[r], m->stream[r]));
    }
    MULTI_NCCL(m, m->nccl.GroupEnd());
    ++m->reductions;
    return PLF_OK;
}

}  // namespace

extern "C" {

const char *plf_multi_last_error(const plf_multi *m)
{
    if (m) g_multi_error = m->error;
    return g_multi_error.c_str();
}

int plf_multi_partition(size_t n_sites, int n_parts, int part, size_t *first, size_t *count)
{
    if (n_parts < 1 || part < 0 || part >= n_parts || !first || !count)
        return mfail(nullptr, PLF_ERR_INVALID, "partition: part %d of %d", part, n_parts);
    const size_t per = (n_sites + (size_t)n_parts - 1) / (size_t)n_parts;     // include.h:181-186
    const size_t lo = per * (size_t)part;
    *first = lo < n_sites ? lo : n_sites;
    *count = lo >= n_sites ? 0 : (n_sites - lo < per ? n_sites - lo : per);   // the last part takes the remainder (:187-192)
    return PLF_OK;
}

int plf_multi_create(plf_multi **out, const int *devices, int n_devices, unsigned n_instances, int layout, int input_src)
{
    return plf_multi_create_states(out, devices, n_devices, n_instances, layout, input_src, 4);
}

int plf_multi_create_states(plf_multi **out, const int *devices, int n_devices, unsigned n_instances, int layout, int input_src,
                            int states)
{
    if (!out) return mfail(nullptr, PLF_ERR_INVALID, "NULL out-pointer");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return mfail(nullptr, PLF_ERR_INVALID, "need 1..64 devices");
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b]) return mfail(nullptr, PLF_ERR_INVALID, "device %d listed twice", devices[a]);
    plf_multi *m = new (std::nothrow) plf_multi;
    if (!m) return mfail(nullptr, PLF_ERR_NOMEM, "out of host memory");
    m->devices.assign(devices, devices + n_devices);
    auto bail = [&](int rc) {
        const std::string keep = g_multi_error;
        plf_multi_destroy(m);
        g_multi_error = keep;
        return rc;
    };
    for (int r = 0; r < n_devices; ++r) {
        plf_ctx *c = nullptr;
        int rc = plf_ctx_create_states(&c, devices[r], n_instances, layout, input_src, states);
        if (rc != PLF_OK) return bail(mfail(nullptr, rc, "rank %d (device %d): %s", r, devices[r], plf_last_error(nullptr)));
        m->ctx.push_back(c);
        cudaStream_t s = nullptr;
        unsigned long long *di = nullptr;
        double *dl = nullptr;
        cudaError_t e = cudaSetDevice(devices[r]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&di, sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&dl, sizeof(double));
        m->stream.push_back(s);
        m->d_inc.push_back(di);
        m->d_lnl.push_back(dl);
        if (e != cudaSuccess) return bail(mfail(nullptr, PLF_ERR_CUDA, "rank %d setup failed: %s", r, cudaGetErrorString(e)));
    }
    if (cudaMallocHost(&m->h_inc, n_devices * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(&m->h_lnl, n_devices * sizeof(double)) != cudaSuccess)
        return bail(mfail(nullptr, PLF_ERR_NOMEM, "pinned staging allocation failed"));
    if (n_devices > 1) {
        m->nccl = load_nccl();
        if (!m->nccl.ok())
            return bail(mfail(nullptr, PLF_ERR_STATE, "NCCL (libnccl.so.2) not found: a multi-GPU plf_multi needs it for the reduction"));
        m->nccl.GetVersion(&m->nccl_version);
        m->comm.assign(n_devices, nullptr);
        ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), n_devices, m->devices.data());
        if (r != ncclSuccess) {
            m->comm.clear();
            return bail(mfail(nullptr, PLF_ERR_CUDA, "ncclCommInitAll over %d devices failed: %s", n_devices, m->nccl.GetErrorString(r)));
        }
    }
    *out = m;
    return PLF_OK;
}

int plf_multi_destroy(plf_multi *m)
{
    if (!m) return PLF_OK;
    for (size_t r = 0; r < m->comm.size(); ++r)
        if (m->comm[r]) m->nccl.CommDestroy(m->comm[r]);
    for (size_t r = 0; r < m->devices.size(); ++r) {
        cudaSetDevice(m->devices[r]);
        if (r < m->stream.size() && m->stream[r]) {
            cudaStreamSynchronize(m->stream[r]);
            cudaStreamDestroy(m->stream[r]);
        }
        if (r < m->d_inc.size()) cudaFree(m->d_inc[r]);
        if (r < m->d_lnl.size()) cudaFree(m->d_lnl[r]);
    }
    for (plf_ctx *c : m->ctx) plf_ctx_destroy(c);
    if (m->h_inc) cudaFreeHost(m->h_inc);
    if (m->h_lnl) cudaFreeHost(m->h_lnl);
    delete m;
    return PLF_OK;
}

int plf_multi_size(const plf_multi *m) { return m ? (int)m->devices.size() : 0; }

plf_ctx *plf_multi_ctx(plf_multi *m, int rank)
{
    if (!m || rank < 0 || rank >= (int)m->ctx.size()) return nullptr;
    return m->ctx[rank];
}

int plf_multi_info(const plf_multi *m, int *n_devices, int *nccl_version, unsigned long long *reductions)
{
    if (!m) return mfail(nullptr, PLF_ERR_INVALID, "NULL multi");
    if (n_devices) *n_devices = (int)m->devices.size();
    if (nccl_version) *nccl_version = m->nccl_version;
    if (reductions) *reductions = m->reductions;
    return PLF_OK;
}

int plf_multi_reduce(plf_multi *m, const long long *increments, const double *lnl, long long *increment_total, double *lnl_total)
{
    if (!m) return mfail(nullptr, PLF_ERR_INVALID, "NULL multi");
    const int g = (int)m->devices.size();
    for (int r = 0; r < g; ++r) {
        m->h_inc[r] = increments ? (unsigned long long)increments[r] : 0ull;
        m->h_lnl[r] = lnl ? lnl[r] : 0.0;
        MULTI_CUDA(m, cudaSetDevice(m->devices[r]));
        MULTI_CUDA(m, cudaMemcpyAsync(m->d_inc[r], m->h_inc + r, sizeof(unsigned long long), cudaMemcpyHostToDevice, m->stream[r]));
        MULTI_CUDA(m, cudaMemcpyAsync(m->d_lnl[r], m->h_lnl + r, sizeof(double), cudaMemcpyHostToDevice, m->stream[r]));
    }
    int rc = allreduce_device(m);
    if (rc != PLF_OK) return rc;
    // every rank now holds the totals; read them ALL back and insist that they agree
    for (int r = 0; r < g; ++r) {
        MULTI_CUDA(m, cudaSetDevice(m->devices[r]));
        MULTI_CUDA(m, cudaMemcpyAsync(m->h_inc + r, m->d_inc[r], sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream[r]));
        MULTI_CUDA(m, cudaMemcpyAsync(m->h_lnl + r, m->d_lnl[r], sizeof(double), cudaMemcpyDeviceToHost, m->stream[r]));
    }
    for (int r = 0; r < g; ++r) {
        MULTI_CUDA(m, cudaSetDevice(m->devices[r]));
        MULTI_CUDA(m, cudaStreamSynchronize(m->stream[r]));
    }
    for (int r = 1; r < g; ++r)
        if (m->h_inc[r] != m->h_inc[0] || m->h_lnl[r] != m->h_lnl[0])
            return mfail(m, PLF_ERR_CUDA, "all-reduce results differ between rank 0 and rank %d", r);
    if (increment_total) *increment_total = (long long)m->h_inc[0];
    if (lnl_total) *lnl_total = m->h_lnl[0];
    return PLF_OK;
}

int plf_multi_newview(plf_multi *m, const float *ev, const float *p_left, const float *p_right, const float *x1, const float *x2,
                      float *x3, char *scaler, const int *wgt, size_t n_sites, long long *increment)
{
    if (!m) return mfail(nullptr, PLF_ERR_INVALID, "NULL multi");
    const int g = (int)m->devices.size();
    std::vector<int> rc(g, PLF_OK);
    std::vector<long long> inc(g, 0);
    std::vector<std::string> err(g);
    auto work = [&](int r) {
        size_t first = 0, cnt = 0;
        plf_multi_partition(n_sites, g, r, &first, &cnt);
        if (cnt == 0) return;
        const size_t sf = 4u * (size_t)plf_ctx_states(m->ctx[r]);
        rc[r] = plf_newview_stream(m->ctx[r], ev, p_left, p_right, x1 + first * sf, x2 + first * sf, x3 + first * sf,
                                   scaler ? scaler + first : nullptr, wgt ? wgt + first : nullptr, cnt, 0, &inc[r]);
        if (rc[r] != PLF_OK) err[r] = plf_last_error(m->ctx[r]);
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < g; ++r) threads.emplace_back(work, r);
    work(0);
    for (auto &t : threads) t.join();
    for (int r = 0; r < g; ++r)
        if (rc[r] != PLF_OK) return mfail(m, rc[r], "rank %d (device %d): %s", r, m->devices[r], err[r].c_str());
    if (!increment) return PLF_OK;
    return plf_multi_reduce(m, inc.data(), nullptr, increment, nullptr);
}

}  // extern "C"
