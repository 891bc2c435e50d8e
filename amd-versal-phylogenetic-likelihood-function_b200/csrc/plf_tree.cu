// csrc/plf_tree.cu -- chained newview over a rooted binary tree (plf_tree_* of include/b200plf.h).
//
// The reference ends at a single newview call; this is the layer that calls the path in a
// RAxML-like code (BASELINE.json configs[4]): a post-order traversal
// where every inner node is one newview of its two children.  Design for B200:
//   * the post-order list is cut into LEVELS (level = 1 + max level of the children); all nodes
//     of a level are independent and run in ONE launch of plf_newview_batch, so a 1024-taxon tree
//     costs ~10-40 launches instead of 1023 launch-bound ones;
//   * the level launches (+ the counter reset) are captured once into a CUDA graph and replayed;
//   * inner CLVs and their int32 scaler-count vectors live in a pool of slots that is recycled as
//     soon as a node's parent has been computed, so device memory is tips + O(widest level);
//   * per-site scaler counts are carried upwards: cnt[parent] = cnt[left] + cnt[right] + rescaled.
#include "../../include/b200plf.h"
#include "plf_kernels.cuh"
#include "plf_registry.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

struct plf_tree {
    int device = 0;
    int states = 4;                            // STATES knob: 4 (DNA, level-batched kernel) or 20 (protein, one launch per node)
    unsigned n_tips = 0, n_inner = 0;
    size_t n_sites = 0;
    int math = PLF_MATH_STRICT;
    int tune_u = 0;
    int tune_chunk = 0;
    int num_sms = 0;
    std::vector<int> left, right;              // children ids per inner node
    std::vector<std::vector<int>> levels;      // inner node indices per level, in execution order
    std::vector<int> slot;                     // pool slot of each inner node
    unsigned n_slots = 0;
    int tip_format = 0;                        // 0: dense CLVs, 1: one state code per site + tip vector table
    float *d_tips = nullptr;                   // dense: [n_tips][n_sites*16]
    unsigned char *d_codes = nullptr;          // codes: [n_tips][code_stride]
    float *d_pool = nullptr;                   // [n_slots][n_sites*16]
    int *d_counts = nullptr;                   // [n_slots][n_sites]
    float *d_mats = nullptr;                   // EV[S^2] | P_left[n_inner][4S^2] | P_right[n_inner][4S^2] | tipvec[16][4] (DNA)
    int *d_wgt = nullptr;
    bool use_wgt = false;
    float *d_tiptab = nullptr;                 // codes trees: [n_tiptip][256][16] tabulated newviews of the tip-tip nodes
    unsigned char *d_tipflag = nullptr;        //              [n_tiptip][256] rescale flags
    unsigned n_tiptip = 0;
    plf::BatchOp *d_ops = nullptr;             // all ops, level after level
    std::vector<size_t> level_op_offset;
    unsigned long long *d_sum = nullptr, *h_sum = nullptr;
    unsigned long long *d_work = nullptr;      // work-counter pair of the dynamically scheduled level launches
    bool dynamic = true;
    double *d_lnl = nullptr;                   // [0] lnL accumulator, [1..4] diag (as 16 floats)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaGraphExec_t exec = nullptr;
    int exec_math = -1, exec_u = -1, exec_chunk = -1, exec_fenced = -1;
    bool exec_wgt = false;
    bool ran = false;
    std::string error;
};

namespace {

thread_local std::string g_tree_error;

int tfail(plf_tree *t, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (t) t->error = buf;
    g_tree_error = buf;
    return code;
}

#define TREE_CUDA(t, expr)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return tfail(t, e__ == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA,      \
                         "%s failed: %s", #expr, cudaGetErrorString(e__));                        \
    } while (0)

using BatchFn = void (*)(const plf::BatchOp *, int, size_t, const int *, unsigned long long *, unsigned,
                         unsigned long long *, int);

struct BatchSel {
    BatchFn fn;
    int threads;
    size_t smem;
    int stage;
};

template <class M, class M15>
BatchSel batch_sel(int u)
{
    // u = 1: 16 consumer warps, 128-site stages x 4;  u = 2: 16 consumer warps, 256-site stages x 3;
    // u = 3: 15 consumer warps, 240-site stages x 3;  u = 4: 11 consumer warps, 4 rows per warp, 352-site stages x 2;
    // u = 0 (default): 11 consumer warps, 4 rows per warp, 352-site stages x 3.
    // Registers are allocated per 4 warps: 16 + 1 warps are charged as 20 (96 registers per thread: scalar strict
    // arithmetic, the packed one spills), 15 + 1 get 128, 11 + 1 get 170 -- room for four rows per warp in flight.
    // Four rows per warp and stage spread the per-stage bookkeeping (barrier waits, op check, addresses, the count and
    // scaler stores) over twice the sites: against u = 3 the 1024-tip traversal is 5 % faster with dense tips, 12 %
    // with compressed tips, 13 % for small levels (256 tips x 8192 sites), profiles/r01_tree.md.
    if (u == 1)
        return {plf::plf_newview_batch<M, 1, 16, 4, 1>, 17 * 32, plf::batch_smem_bytes<1, 16, 4>(), 128};
    if (u == 2)
        return {plf::plf_newview_batch<M, 2, 16, 3, 1>, 17 * 32, plf::batch_smem_bytes<2, 16, 3>(), 256};
    if (u == 3)
        return {plf::plf_newview_batch<M15, 2, 15, 3, 1>, 16 * 32, plf::batch_smem_bytes<2, 15, 3>(), 240};
    if (u == 4)
        return {plf::plf_newview_batch<M15, 4, 11, 2, 1>, 12 * 32, plf::batch_smem_bytes<4, 11, 2>(), 352};
    return {plf::plf_newview_batch<M15, 4, 11, 3, 1>, 12 * 32, plf::batch_smem_bytes<4, 11, 3>(), 352};
}

BatchSel pick_batch(int math, int u)
{
    return math == PLF_MATH_FMA ? batch_sel<plf::MathFma, plf::MathFma>(u)
                                : batch_sel<plf::MathStrictScalar, plf::MathStrict>(u);
}

size_t site_floats(const plf_tree *t) { return 4u * (size_t)t->states; }
size_t ev_floats(const plf_tree *t) { return (size_t)t->states * t->states; }
size_t p_floats(const plf_tree *t) { return 4u * (size_t)t->states * t->states; }
float *node_pl(plf_tree *t, unsigned k) { return t->d_mats + ev_floats(t) + p_floats(t) * (size_t)k; }
float *node_pr(plf_tree *t, unsigned k) { return t->d_mats + ev_floats(t) + p_floats(t) * ((size_t)t->n_inner + k); }

// tip code vectors are padded to a multiple of 16 bytes (bulk-copy granularity)
size_t code_stride(const plf_tree *t) { return (t->n_sites + 15) & ~(size_t)15; }

// dense CLV of a node; NULL for a compressed tip
float *node_clv(plf_tree *t, int node)
{
    const size_t stride = t->n_sites * site_floats(t);
    if (node < (int)t->n_tips) return t->tip_format ? nullptr : t->d_tips + (size_t)node * stride;
    return t->d_pool + (size_t)t->slot[node - t->n_tips] * stride;
}

const unsigned char *node_codes(plf_tree *t, int node)
{
    return (t->tip_format && node < (int)t->n_tips) ? t->d_codes + (size_t)node * code_stride(t) : nullptr;
}

float *tipvec_ptr(plf_tree *t) { return t->d_mats + ev_floats(t) + 2 * p_floats(t) * (size_t)t->n_inner; }

// count vectors are padded to a multiple of 4 ints so that every stage of them is a legal
// (16-byte aligned, 16-byte granular) bulk copy
size_t count_stride(const plf_tree *t) { return (t->n_sites + 3) & ~(size_t)3; }

int *node_counts(plf_tree *t, int node)
{
    return node < (int)t->n_tips ? nullptr : t->d_counts + (size_t)t->slot[node - t->n_tips] * count_stride(t);
}

int launch_level(plf_tree *t, const BatchSel &k, size_t level, cudaStream_t s)
{
    const int n_ops = (int)t->levels[level].size();
    const size_t spo = (t->n_sites + k.stage - 1) / k.stage;
    const size_t stages = spo * (size_t)n_ops;
    if (stages >= (1ull << 31)) return tfail(t, PLF_ERR_INVALID, "level %zu has too many stages (%zu)", level, stages);
    size_t grid = (size_t)t->num_sms;
    if (grid > stages) grid = stages;
    // chunk of consecutive stages dealt to one CTA at a time (see plf_newview_batch).  Auto: about 16
    // chunks per CTA keeps the level balanced while a CTA changes op (reloading the 48 constants
    // behind two dependent global loads) as rarely as possible; never longer than one op.
    size_t chunk = (size_t)t->tune_chunk;
    if (chunk == 0) chunk = std::max<size_t>(1, std::min(spo, stages / (16 * grid)));
    // programmatic dependent launch between consecutive levels: the next level's CTAs take over SMs as
    // this level drains and wait in griddepcontrol.wait for its completion (PLF_PDL=0 disables)
    static const bool use_pdl = [] {
        const char *e = getenv("PLF_PDL");
        return !(e && e[0] == '0');
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)k.threads);
    cfg.dynamicSmemBytes = k.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (use_pdl && level > 0) ? 1 : 0;      // level 0 follows a memset node: plain dependency
    const plf::BatchOp *ops = t->d_ops + t->level_op_offset[level];
    const int *wgt = t->use_wgt ? t->d_wgt : nullptr;
    unsigned long long *work = t->dynamic ? t->d_work : nullptr;
    // slot release: data dependency by default (the fence caps the compressed-tip levels, plf_kernels.cuh);
    // PLF_SAFE_RELEASE=1 selects the fenced release (read when the graph is captured)
    const int flags = plf::fenced_release(false) ? plf::kFlagFencedRelease : 0;
    TREE_CUDA(t, cudaLaunchKernelEx(&cfg, k.fn, ops, n_ops, t->n_sites, wgt, t->d_sum, (unsigned)chunk, work, flags));
    TREE_CUDA(t, cudaGetLastError());
    return PLF_OK;
}

// 20 states: one launch of the protein kernel per inner node, level after level (nodes of a level are independent, and
// the captured stream keeps the level order).  A node moves 964 B/site, so from ~30 k sites per GPU a launch is long
// enough to hide its own launch cost inside the graph.
int build_graph_states(plf_tree *t)
{
    if (t->exec) {
        cudaGraphExecDestroy(t->exec);
        t->exec = nullptr;
    }
    cudaGraph_t graph = nullptr;
    TREE_CUDA(t, cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaMemsetAsync(t->d_sum, 0, sizeof(unsigned long long), t->stream);
    int rc = PLF_OK;
    const int flags = plf::kAaReleaseUnset;          // every kernel's own default, or the process-wide override
    for (size_t l = 0; l < t->levels.size() && e == cudaSuccess && rc == PLF_OK; ++l)
        for (int k : t->levels[l]) {
            rc = plf::launch_newview_aa(node_clv(t, t->left[k]), node_clv(t, t->right[k]), node_clv(t, (int)t->n_tips + k), nullptr,
                                        t->d_mats, node_pl(t, (unsigned)k), node_pr(t, (unsigned)k), t->use_wgt ? t->d_wgt : nullptr,
                                        t->n_sites, t->d_sum, t->math, 0, 0, flags, t->stream, node_counts(t, t->left[k]),
                                        node_counts(t, t->right[k]), node_counts(t, (int)t->n_tips + k));
            if (rc != PLF_OK) break;
        }
    if (e == cudaSuccess && rc == PLF_OK)
        e = cudaMemcpyAsync(t->h_sum, t->d_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream);
    cudaError_t e2 = cudaStreamEndCapture(t->stream, &graph);
    if (rc != PLF_OK || e != cudaSuccess || e2 != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return tfail(t, rc != PLF_OK ? rc : PLF_ERR_CUDA, "graph capture of the 20-state traversal failed: %s",
                     cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    e = cudaGraphInstantiate(&t->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return tfail(t, PLF_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    t->exec_math = t->math;
    t->exec_u = t->tune_u;
    t->exec_chunk = t->tune_chunk;
    t->exec_wgt = t->use_wgt;
    t->exec_fenced = plf::fenced_release(false) ? 1 : 0;
    return PLF_OK;
}

// (Re)capture the traversal: counter reset + one launch per level + read-back of the counter.
int build_graph(plf_tree *t)
{
    if (t->states != 4) return build_graph_states(t);
    int u = t->tune_u;
    if (u == 0) {
        // 352-site stages (4 rows per warp) unless even the widest level would leave most SMs with < 4 stages
        const size_t widest = t->levels.empty() ? 1 : t->levels[0].size();
        u = ((t->n_sites + 351) / 352) * widest >= (size_t)t->num_sms * 4 ? 0 : 1;
    }
    const BatchSel k = pick_batch(t->math, u);
    TREE_CUDA(t, cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem));
    if (t->exec) {
        cudaGraphExecDestroy(t->exec);
        t->exec = nullptr;
    }
    cudaGraph_t graph = nullptr;
    TREE_CUDA(t, cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaMemsetAsync(t->d_sum, 0, sizeof(unsigned long long), t->stream);
    int rc = PLF_OK;
    if (e == cudaSuccess && t->n_tiptip) {          // the matrices may have changed since the last traversal: re-tabulate
        if (t->math == PLF_MATH_FMA) plf::plf_tiptip_tables<plf::MathFma><<<(int)t->n_inner, 256, 0, t->stream>>>(t->d_ops, (int)t->n_inner);
        else plf::plf_tiptip_tables<plf::MathStrict><<<(int)t->n_inner, 256, 0, t->stream>>>(t->d_ops, (int)t->n_inner);
        e = cudaGetLastError();
    }
    for (size_t l = 0; l < t->levels.size() && e == cudaSuccess && rc == PLF_OK; ++l) rc = launch_level(t, k, l, t->stream);
    if (e == cudaSuccess && rc == PLF_OK)
        e = cudaMemcpyAsync(t->h_sum, t->d_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream);
    cudaError_t e2 = cudaStreamEndCapture(t->stream, &graph);
    if (rc != PLF_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess || e2 != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return tfail(t, PLF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    e = cudaGraphInstantiate(&t->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return tfail(t, PLF_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    t->exec_math = t->math;
    t->exec_u = t->tune_u;
    t->exec_chunk = t->tune_chunk;
    t->exec_wgt = t->use_wgt;
    t->exec_fenced = plf::fenced_release(false) ? 1 : 0;
    return PLF_OK;
}

}  // namespace

extern "C" {

const char *plf_tree_last_error(const plf_tree *tree)
{
    if (tree) g_tree_error = tree->error;
    return g_tree_error.c_str();
}

int plf_tree_create(plf_tree **out, int device, unsigned n_tips, const int *left, const int *right, size_t n_sites)
{
    return plf_tree_create_ex(out, device, n_tips, left, right, n_sites, PLF_TIPS_DENSE);
}

int plf_tree_create_ex(plf_tree **out, int device, unsigned n_tips, const int *left, const int *right, size_t n_sites,
                       int tip_format)
{
    return plf_tree_create_states(out, device, n_tips, left, right, n_sites, tip_format, 4);
}

int plf_tree_create_states(plf_tree **out, int device, unsigned n_tips, const int *left, const int *right, size_t n_sites,
                           int tip_format, int states)
{
    if (!out) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree out-pointer");
    *out = nullptr;
    if (states != 4 && states != 20) return tfail(nullptr, PLF_ERR_INVALID, "STATES=%d is not supported (4 = DNA, 20 = protein)", states);
    if (states != 4 && tip_format != PLF_TIPS_DENSE)
        return tfail(nullptr, PLF_ERR_INVALID, "state-code tips are implemented for STATES=DNA only");
    if (n_tips < 2 || !left || !right || n_sites == 0)
        return tfail(nullptr, PLF_ERR_INVALID, "need n_tips >= 2, child arrays and n_sites > 0");
    if (n_sites > (SIZE_MAX / 1024)) return tfail(nullptr, PLF_ERR_INVALID, "n_sites too large");
    if (tip_format != PLF_TIPS_DENSE && tip_format != PLF_TIPS_CODES)
        return tfail(nullptr, PLF_ERR_INVALID, "unknown tip format %d", tip_format);
    const unsigned n_inner = n_tips - 1;
    // validate: post-order, every node except the root is used exactly once as a child
    std::vector<int> used(n_tips + n_inner, 0);
    for (unsigned k = 0; k < n_inner; ++k) {
        for (int c : {left[k], right[k]}) {
            if (c < 0 || c >= (int)(n_tips + k))
                return tfail(nullptr, PLF_ERR_INVALID, "inner node %u: child id %d is not an earlier node", k, c);
            if (used[c]++) return tfail(nullptr, PLF_ERR_INVALID, "node %d is the child of two nodes", c);
        }
        if (left[k] == right[k]) return tfail(nullptr, PLF_ERR_INVALID, "inner node %u has identical children", k);
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return tfail(nullptr, PLF_ERR_CUDA, "no CUDA device present (no CPU fallback)");
    }
    if (device < 0 || device >= count) return tfail(nullptr, PLF_ERR_INVALID, "device %d not present", device);
    TREE_CUDA(nullptr, cudaSetDevice(device));

    plf_tree *t = new (std::nothrow) plf_tree;
    if (!t) return tfail(nullptr, PLF_ERR_NOMEM, "out of host memory");
    t->device = device;
    t->n_tips = n_tips;
    t->n_inner = n_inner;
    t->n_sites = n_sites;
    t->states = states;
    t->tip_format = tip_format;
    if (const char *e = getenv("PLF_TREE_STATIC")) t->dynamic = !(e[0] == '1');
    t->left.assign(left, left + n_inner);
    t->right.assign(right, right + n_inner);
    cudaDeviceGetAttribute(&t->num_sms, cudaDevAttrMultiProcessorCount, device);

    // levels
    std::vector<int> level(n_tips + n_inner, 0);
    int height = 0;
    for (unsigned k = 0; k < n_inner; ++k) {
        level[n_tips + k] = 1 + std::max(level[left[k]], level[right[k]]);
        height = std::max(height, level[n_tips + k]);
    }
    t->levels.assign(height, {});
    for (unsigned k = 0; k < n_inner; ++k) t->levels[level[n_tips + k] - 1].push_back((int)k);

    // slot pool with recycling: a child's slot is released once its parent's level has run
    t->slot.assign(n_inner, -1);
    std::vector<int> free_slots;
    unsigned next_slot = 0;
    for (auto &lv : t->levels) {
        for (int k : lv) {
            if (!free_slots.empty()) {
                t->slot[k] = free_slots.back();
                free_slots.pop_back();
            } else {
                t->slot[k] = (int)next_slot++;
            }
        }
        for (int k : lv)
            for (int c : {t->left[k], t->right[k]})
                if (c >= (int)n_tips) free_slots.push_back(t->slot[c - n_tips]);
    }
    t->n_slots = next_slot;

    const size_t clv_bytes = n_sites * site_floats(t) * sizeof(float);
    cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&t->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&t->ev1);
    if (e == cudaSuccess && tip_format == PLF_TIPS_DENSE) e = cudaMalloc(&t->d_tips, clv_bytes * n_tips);
    if (e == cudaSuccess && tip_format == PLF_TIPS_CODES) e = cudaMalloc(&t->d_codes, code_stride(t) * n_tips);
    if (e == cudaSuccess && tip_format == PLF_TIPS_CODES) e = cudaMemset(t->d_codes, 0, code_stride(t) * n_tips);
    if (e == cudaSuccess) e = cudaMalloc(&t->d_pool, clv_bytes * t->n_slots);
    if (e == cudaSuccess) e = cudaMalloc(&t->d_counts, count_stride(t) * sizeof(int) * t->n_slots);
    if (e == cudaSuccess) e = cudaMemset(t->d_counts, 0, count_stride(t) * sizeof(int) * t->n_slots);
    if (e == cudaSuccess) e = cudaMalloc(&t->d_mats, (ev_floats(t) + 2 * p_floats(t) * (size_t)n_inner + 64) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&t->d_ops, sizeof(plf::BatchOp) * n_inner);
    if (e == cudaSuccess) e = cudaMalloc(&t->d_sum, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&t->d_work, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(t->d_work, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&t->h_sum, sizeof(unsigned long long));
    if (e != cudaSuccess) {
        cudaGetLastError();
        tfail(nullptr, e == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA,
              "tree allocation failed (%u tips, %u slots, %zu sites): %s", n_tips, t->n_slots, n_sites,
              cudaGetErrorString(e));
        plf_tree_destroy(t);
        return e == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA;
    }
    *t->h_sum = 0;

    // tip-tip nodes of a codes tree get a 256-pair table each (plf_tiptip_tables fills them at the start of a traversal)
    if (tip_format == PLF_TIPS_CODES && !getenv("PLF_NO_TIPTIP_TABLES")) {
        for (unsigned k = 0; k < n_inner; ++k) t->n_tiptip += (left[k] < (int)n_tips && right[k] < (int)n_tips);
        if (t->n_tiptip) {
            e = cudaMalloc(&t->d_tiptab, (size_t)t->n_tiptip * 256 * 16 * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&t->d_tipflag, (size_t)t->n_tiptip * 256);
            if (e != cudaSuccess) {
                cudaGetLastError();
                tfail(nullptr, PLF_ERR_NOMEM, "tip-tip table allocation failed: %s", cudaGetErrorString(e));
                plf_tree_destroy(t);
                return PLF_ERR_NOMEM;
            }
        }
    }
    unsigned next_tab = 0;

    // op descriptors, level after level
    std::vector<plf::BatchOp> ops;
    ops.reserve(n_inner);
    for (auto &lv : t->levels) {
        t->level_op_offset.push_back(ops.size());
        for (int k : lv) {
            plf::BatchOp o;
            o.x1 = reinterpret_cast<const float4 *>(node_clv(t, t->left[k]));
            o.x2 = reinterpret_cast<const float4 *>(node_clv(t, t->right[k]));
            o.x3 = reinterpret_cast<float4 *>(node_clv(t, (int)n_tips + k));
            o.cnt1 = node_counts(t, t->left[k]);
            o.cnt2 = node_counts(t, t->right[k]);
            o.cnt3 = node_counts(t, (int)n_tips + k);
            o.scaler = nullptr;
            o.tip1 = node_codes(t, t->left[k]);
            o.tip2 = node_codes(t, t->right[k]);
            o.tipvec = tip_format == PLF_TIPS_CODES ? tipvec_ptr(t) : nullptr;
            o.tiptab = nullptr;
            o.tipflag = nullptr;
            if (t->d_tiptab && o.tip1 && o.tip2) {
                o.tiptab = reinterpret_cast<const float4 *>(t->d_tiptab + (size_t)next_tab * 256 * 16);
                o.tipflag = t->d_tipflag + (size_t)next_tab * 256;
                ++next_tab;
            }
            o.ev = t->d_mats;
            o.pl = node_pl(t, (unsigned)k);
            o.pr = node_pr(t, (unsigned)k);
            ops.push_back(o);
        }
    }
    e = cudaMemcpy(t->d_ops, ops.data(), sizeof(plf::BatchOp) * ops.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        tfail(nullptr, PLF_ERR_CUDA, "op table upload failed: %s", cudaGetErrorString(e));
        plf_tree_destroy(t);
        return PLF_ERR_CUDA;
    }
    *out = t;
    return PLF_OK;
}

int plf_tree_destroy(plf_tree *t)
{
    if (!t) return PLF_OK;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    if (t->exec) cudaGraphExecDestroy(t->exec);
    cudaFree(t->d_tips);
    cudaFree(t->d_codes);
    cudaFree(t->d_pool);
    cudaFree(t->d_counts);
    cudaFree(t->d_mats);
    cudaFree(t->d_wgt);
    cudaFree(t->d_ops);
    cudaFree(t->d_tiptab);
    cudaFree(t->d_tipflag);
    cudaFree(t->d_sum);
    cudaFree(t->d_lnl);
    cudaFree(t->d_work);
    if (t->h_sum) cudaFreeHost(t->h_sum);
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
    return PLF_OK;
}

int plf_tree_set_math(plf_tree *t, int math_mode)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (math_mode != PLF_MATH_STRICT && math_mode != PLF_MATH_FMA) return tfail(t, PLF_ERR_INVALID, "unknown math mode %d", math_mode);
    t->math = math_mode;
    return PLF_OK;
}

int plf_tree_set_tuning(plf_tree *t, int u, int chunk)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (u < 0 || u > 4) return tfail(t, PLF_ERR_INVALID, "tuning u must be 0..4");
    if (chunk < 0) return tfail(t, PLF_ERR_INVALID, "chunk must be >= 0");
    t->tune_u = u;
    t->tune_chunk = chunk;
    return PLF_OK;
}

int plf_tree_tip_ptr(plf_tree *t, unsigned tip, float **clv)
{
    if (!t || !clv) return tfail(t, PLF_ERR_INVALID, "NULL argument");
    if (tip >= t->n_tips) return tfail(t, PLF_ERR_INVALID, "tip %u out of range (%u tips)", tip, t->n_tips);
    if (t->tip_format != PLF_TIPS_DENSE) return tfail(t, PLF_ERR_STATE, "tree stores tips as state codes: use plf_tree_write_tip_codes");
    *clv = node_clv(t, (int)tip);
    return PLF_OK;
}

int plf_tree_write_tip(plf_tree *t, unsigned tip, const float *clv, size_t bytes, size_t offset)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (tip >= t->n_tips) return tfail(t, PLF_ERR_INVALID, "tip %u out of range (%u tips)", tip, t->n_tips);
    if (t->tip_format != PLF_TIPS_DENSE) return tfail(t, PLF_ERR_STATE, "tree stores tips as state codes: use plf_tree_write_tip_codes");
    const size_t cap = t->n_sites * site_floats(t) * sizeof(float);
    if (offset > cap || bytes > cap - offset) return tfail(t, PLF_ERR_INVALID, "write exceeds the tip CLV (%zu bytes)", cap);
    if (bytes == 0) return PLF_OK;
    if (!clv) return tfail(t, PLF_ERR_INVALID, "NULL host buffer");
    TREE_CUDA(t, cudaSetDevice(t->device));
    TREE_CUDA(t, cudaMemcpyAsync(reinterpret_cast<char *>(node_clv(t, (int)tip)) + offset, clv, bytes,
                                 cudaMemcpyHostToDevice, t->stream));
    return PLF_OK;
}

int plf_tree_write_tip_codes(plf_tree *t, unsigned tip, const unsigned char *codes, size_t n, size_t first_site)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (t->tip_format != PLF_TIPS_CODES) return tfail(t, PLF_ERR_STATE, "tree stores dense tip CLVs: use plf_tree_write_tip");
    if (tip >= t->n_tips) return tfail(t, PLF_ERR_INVALID, "tip %u out of range (%u tips)", tip, t->n_tips);
    if (first_site > t->n_sites || n > t->n_sites - first_site) return tfail(t, PLF_ERR_INVALID, "site range out of bounds");
    if (n == 0) return PLF_OK;
    if (!codes) return tfail(t, PLF_ERR_INVALID, "NULL host buffer");
    TREE_CUDA(t, cudaSetDevice(t->device));
    TREE_CUDA(t, cudaMemcpyAsync(t->d_codes + (size_t)tip * code_stride(t) + first_site, codes, n, cudaMemcpyHostToDevice,
                                 t->stream));
    return PLF_OK;
}

int plf_tree_write_tip_vector(plf_tree *t, const float *tip_vector)
{
    if (!t || !tip_vector) return tfail(t, PLF_ERR_INVALID, "NULL argument");
    if (t->tip_format != PLF_TIPS_CODES) return tfail(t, PLF_ERR_STATE, "tree stores dense tip CLVs");
    TREE_CUDA(t, cudaSetDevice(t->device));
    TREE_CUDA(t, cudaMemcpyAsync(tipvec_ptr(t), tip_vector, 64 * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    TREE_CUDA(t, cudaStreamSynchronize(t->stream));
    return PLF_OK;
}

int plf_tree_write_matrices(plf_tree *t, const float *ev, const float *p_left, const float *p_right)
{
    if (!t || !ev || !p_left || !p_right) return tfail(t, PLF_ERR_INVALID, "NULL argument");
    TREE_CUDA(t, cudaSetDevice(t->device));
    const size_t pb = p_floats(t) * (size_t)t->n_inner * sizeof(float);
    TREE_CUDA(t, cudaMemcpyAsync(t->d_mats, ev, ev_floats(t) * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    TREE_CUDA(t, cudaMemcpyAsync(node_pl(t, 0), p_left, pb, cudaMemcpyHostToDevice, t->stream));
    TREE_CUDA(t, cudaMemcpyAsync(node_pr(t, 0), p_right, pb, cudaMemcpyHostToDevice, t->stream));
    TREE_CUDA(t, cudaStreamSynchronize(t->stream));     // host arrays may be pageable temporaries
    return PLF_OK;
}

int plf_tree_write_wgt(plf_tree *t, const int *wgt)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (!wgt) {
        t->use_wgt = false;
        return PLF_OK;
    }
    TREE_CUDA(t, cudaSetDevice(t->device));
    if (!t->d_wgt) TREE_CUDA(t, cudaMalloc(&t->d_wgt, t->n_sites * sizeof(int)));
    TREE_CUDA(t, cudaMemcpyAsync(t->d_wgt, wgt, t->n_sites * sizeof(int), cudaMemcpyHostToDevice, t->stream));
    TREE_CUDA(t, cudaStreamSynchronize(t->stream));
    t->use_wgt = true;
    return PLF_OK;
}

int plf_tree_run_async(plf_tree *t)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    TREE_CUDA(t, cudaSetDevice(t->device));
    if (!t->exec || t->exec_math != t->math || t->exec_u != t->tune_u || t->exec_chunk != t->tune_chunk || t->exec_wgt != t->use_wgt ||
        t->exec_fenced != (plf::fenced_release(false) ? 1 : 0)) {
        int rc = build_graph(t);
        if (rc != PLF_OK) return rc;
    }
    TREE_CUDA(t, cudaEventRecord(t->ev0, t->stream));
    TREE_CUDA(t, cudaGraphLaunch(t->exec, t->stream));
    TREE_CUDA(t, cudaEventRecord(t->ev1, t->stream));
    plf::count_launches((t->states == 4 ? t->levels.size() : (size_t)t->n_inner) + (t->n_tiptip ? 1 : 0));
    t->ran = true;
    return PLF_OK;
}

int plf_tree_wait(plf_tree *t)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    TREE_CUDA(t, cudaSetDevice(t->device));
    TREE_CUDA(t, cudaStreamSynchronize(t->stream));
    return PLF_OK;
}

int plf_tree_read_root(plf_tree *t, float *clv, int *scaler_counts, size_t first_site, size_t n)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (!t->ran) return tfail(t, PLF_ERR_STATE, "tree has not been run");
    if (first_site > t->n_sites || n > t->n_sites - first_site) return tfail(t, PLF_ERR_INVALID, "site range out of bounds");
    TREE_CUDA(t, cudaSetDevice(t->device));
    const int root = (int)(t->n_tips + t->n_inner - 1);
    if (clv && n)
        TREE_CUDA(t, cudaMemcpyAsync(clv, node_clv(t, root) + first_site * site_floats(t), n * site_floats(t) * sizeof(float),
                                     cudaMemcpyDeviceToHost, t->stream));
    if (scaler_counts && n)
        TREE_CUDA(t, cudaMemcpyAsync(scaler_counts, node_counts(t, root) + first_site, n * sizeof(int),
                                     cudaMemcpyDeviceToHost, t->stream));
    TREE_CUDA(t, cudaStreamSynchronize(t->stream));
    return PLF_OK;
}

int plf_tree_total_scalings(plf_tree *t, long long *total)
{
    if (!t || !total) return tfail(t, PLF_ERR_INVALID, "NULL argument");
    if (!t->ran) return tfail(t, PLF_ERR_STATE, "tree has not been run");
    int rc = plf_tree_wait(t);
    if (rc != PLF_OK) return rc;
    *total = (long long)*t->h_sum;
    return PLF_OK;
}

int plf_tree_info(plf_tree *t, unsigned *levels, unsigned *clv_slots, size_t *device_bytes, size_t *traversal_bytes)
{
    if (!t) return tfail(nullptr, PLF_ERR_INVALID, "NULL tree");
    if (levels) *levels = (unsigned)t->levels.size();
    if (clv_slots) *clv_slots = t->n_slots;
    if (device_bytes)
        *device_bytes = (t->tip_format ? code_stride(t) * (size_t)t->n_tips : t->n_sites * site_floats(t) * 4 * (size_t)t->n_tips) +
                        t->n_sites * site_floats(t) * 4 * (size_t)t->n_slots + count_stride(t) * 4 * (size_t)t->n_slots;
    if (traversal_bytes) {
        // per node: 64 B/site written, 64 B/site read per dense child (1 B/site per compressed tip),
        // + 4 B per count vector read (inner children) or written
        size_t inner_children = 0;
        for (unsigned k = 0; k < t->n_inner; ++k)
            inner_children += (t->left[k] >= (int)t->n_tips) + (t->right[k] >= (int)t->n_tips);
        const size_t tip_children = 2 * (size_t)t->n_inner - inner_children;
        const size_t clv_site = site_floats(t) * 4;
        const size_t tip_read = t->tip_format ? 1 : clv_site;
        *traversal_bytes = t->n_sites * (clv_site * (size_t)t->n_inner + clv_site * inner_children + tip_read * tip_children +
                                         4 * ((size_t)t->n_inner + inner_children));
    }
    return PLF_OK;
}

int plf_tree_evaluate_root(plf_tree *t, const float *diag, double *lnl)
{
    if (!t || !diag || !lnl) return tfail(t, PLF_ERR_INVALID, "NULL argument");
    if (!t->ran) return tfail(t, PLF_ERR_STATE, "tree has not been run");
    TREE_CUDA(t, cudaSetDevice(t->device));
    if (!t->d_lnl) TREE_CUDA(t, cudaMalloc(&t->d_lnl, 2 * sizeof(double) + 80 * sizeof(float)));
    float *d_diag = reinterpret_cast<float *>(t->d_lnl + 2);          // 16-byte aligned behind the accumulator
    TREE_CUDA(t, cudaMemsetAsync(t->d_lnl, 0, sizeof(double), t->stream));
    TREE_CUDA(t, cudaMemcpyAsync(d_diag, diag, site_floats(t) * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    const int a = t->left[t->n_inner - 1], b = t->right[t->n_inner - 1];
    if (!node_clv(t, a) || !node_clv(t, b))
        return tfail(t, PLF_ERR_STATE, "a child of the root is a compressed tip: evaluate needs two dense CLVs");
    int rc = plf::launch_evaluate(t->states, node_clv(t, a), node_clv(t, b), node_counts(t, a), node_counts(t, b),
                                  t->use_wgt ? t->d_wgt : nullptr, d_diag, t->n_sites, t->d_lnl, t->stream);
    if (rc != PLF_OK) return tfail(t, rc, "evaluate kernel launch failed");
    TREE_CUDA(t, cudaMemcpyAsync(lnl, t->d_lnl, sizeof(double), cudaMemcpyDeviceToHost, t->stream));
    TREE_CUDA(t, cudaStreamSynchronize(t->stream));
    return PLF_OK;
}

int plf_tree_last_ms(plf_tree *t, float *ms)
{
    if (!t || !ms) return tfail(t, PLF_ERR_INVALID, "NULL argument");
    if (!t->ran) return tfail(t, PLF_ERR_STATE, "tree has not been run");
    TREE_CUDA(t, cudaSetDevice(t->device));
    TREE_CUDA(t, cudaEventSynchronize(t->ev1));
    TREE_CUDA(t, cudaEventElapsedTime(ms, t->ev0, t->ev1));
    return PLF_OK;
}

}  // extern "C"
