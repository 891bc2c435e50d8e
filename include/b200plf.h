/*
 * include/b200plf.h -- C ABI of the B200-native PLF "newview" path (libb200plf.so).
 *
 * This is the drop-in boundary.  The reference (GeertRoks/AMD-Versal-phylogenetic-likelihood-
 * function) has no FFI layer of its own: its host program drives the accelerator through the
 * XRT C++ API.  Each entry point below replaces one XRT use in the reference host and cites
 * it (paths relative to /root/reference/).  Plain pointers and sizes only; no exceptions cross
 * this boundary; every call returns PLF_OK (0) or a negative plf_status, and the message is
 * available from plf_last_error().
 *
 * Threading: a plf_ctx belongs to one GPU.  Distinct instances of one ctx may be driven from
 * distinct host threads (the reference drives each instance from its own xrt::queue workers,
 * app/src/host_mem.cpp:249-260); one instance must not be driven from two threads at once.
 *
 * There is NO CPU fallback anywhere behind this header: without a CUDA device every compute
 * entry point fails with PLF_ERR_CUDA.
 */
#ifndef B200PLF_H
#define B200PLF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PLF_VERSION 100

/* Sizes fixed by STATES=DNA (Makefile:31): 4 rate categories x 4 states. */
#define PLF_SITE_FLOATS 16u      /* one site of a CLV: [category j][state l]  (plf.cpp:21-23)   */
#define PLF_EV_FLOATS 16u        /* EV[k][l]                                  (plf.cpp:47)      */
#define PLF_BRANCH_FLOATS 64u    /* P[j][k][l]                                (plf.cpp:37-38)   */
#define PLF_HEADER_COMB 80u      /* floats in front of the CLV: [EV16|P64]    (host_mem.cpp:231-233) */
#define PLF_HEADER_SEP 64u       /* Sep right buffer: [P64]                   (host_mem.cpp:239-240) */

typedef enum plf_status {
    PLF_OK = 0,
    PLF_ERR_INVALID = -1,   /* bad argument (index, size, alignment, NULL)            */
    PLF_ERR_CUDA = -2,      /* CUDA runtime / no device / kernel launch failure      */
    PLF_ERR_NOMEM = -3,     /* device or pinned-host allocation failed                */
    PLF_ERR_STATE = -4      /* call order violated (e.g. run before alloc)            */
} plf_status;

/* PLIO_LAYOUT knob (Makefile:28): how the right-child buffer is packed.
 *   COMB: left = right-format = [EV16|P64|CLV]   SEP: right = [P64|CLV]
 * (host_mem.cpp:234-241; mm2sright_memDNAwindowComb.cpp:32-42 vs ...windowSep.cpp:37-40).   */
typedef enum plf_layout { PLF_LAYOUT_COMB = 0, PLF_LAYOUT_SEP = 1 } plf_layout;

/* INPUT_SRC knob (Makefile:30).  MEM: CLVs are read from the instance buffers.
 * GEN: the kernel synthesises the constant site pattern of hls/src/mm2s{left,right}_gen*.cpp
 * in registers and reads no CLV; used to isolate kernel throughput (host_gen.cpp).            */
typedef enum plf_input_src { PLF_INPUT_MEM = 0, PLF_INPUT_GEN = 1 } plf_input_src;

/* Arithmetic mode of the fused kernel.
 *   STRICT: separately rounded fp32 multiplies and adds in the order of plf.cpp:29-52
 *           -> bit-identical to the reference's CPU plf() (the default).
 *   FMA:    contracted multiply-adds; CLVs agree to <= 1e-5 relative, scaler decisions can
 *           differ only for sites whose max |x3| is within that distance of 2^-32.          */
typedef enum plf_math { PLF_MATH_STRICT = 0, PLF_MATH_FMA = 1 } plf_math;

/* GEN sink behaviour (cfg4a / cfg4b of SURVEY.md section 8d). */
typedef enum plf_gen_sink {
    PLF_GEN_WRITE = 0,      /* write CLV + scaler bytes (65 B/site)                      */
    PLF_GEN_DISCARD = 1     /* fold outputs into a checksum, write nothing (s2mm_gen*)    */
} plf_gen_sink;

/* Timestamps the reference host takes around each call (host_mem.cpp:294,302,309,318). */
typedef enum plf_mark_id { PLF_MARK_BEGIN = 0, PLF_MARK_T1 = 1, PLF_MARK_T2 = 2, PLF_MARK_END = 3 } plf_mark_id;

typedef struct plf_ctx plf_ctx;

/* ---- device / context ------------------------------------------------------------------ */

/* Number of CUDA devices (0 and PLF_ERR_CUDA when there is none). */
int plf_device_count(int *count);
/* Device name and PCI bus id, the analogue of xrt::info::device::{name,bdf} (host_mem.cpp:88-89). */
int plf_device_info(int device, char *name, size_t name_len, char *bdf, size_t bdf_len);
/* Resolve "0000:5e:00.0"-style BDF (argv[2] of the reference host) or a decimal ordinal. */
int plf_device_from_string(const char *bdf_or_ordinal, int *device);

/* Replaces acap_info(xclbin, BDF): open device, "load" the accelerator (include.h:30-36,85-102).
 * n_instances is NUM_ACCELERATORS (Makefile:29): independent PLF instances, each with its own
 * CUDA stream.  Fails (PLF_ERR_CUDA) instead of throwing.                                     */
int plf_ctx_create(plf_ctx **ctx, int device, unsigned n_instances, int layout, int input_src);
/* The same with the STATES knob (Makefile:31; README.md:36,202): states = 4 (DNA, what plf_ctx_create gives) or 20
 * (protein).  With S states every size of the instance API scales: one site is 4*S floats, EV S*S, one child's branch
 * matrices 4*S*S, so the packed buffers are [EV S^2 | P 4S^2 | CLV n*4S] (left; right in the Comb layout) and
 * [P 4S^2 | CLV] (right, Sep) -- host_mem.cpp:231-241 with 4 replaced by S.  plf_instance_alloc, plf_write_left/right,
 * plf_run_async, plf_read_out/scaler, plf_scaler_increment and plf_newview_stream all follow the context's state count;
 * the matrices are read from the head of the device buffers, not passed per launch.  INPUT_SRC=gen exists for DNA only. */
int plf_ctx_create_states(plf_ctx **ctx, int device, unsigned n_instances, int layout, int input_src, int states);
int plf_ctx_states(const plf_ctx *ctx);
int plf_ctx_destroy(plf_ctx *ctx);
/* Last error message of this ctx (or of ctx-less calls when ctx == NULL).  Never NULL. */
const char *plf_last_error(const plf_ctx *ctx);

int plf_ctx_set_math(plf_ctx *ctx, int math_mode);          /* plf_math; default STRICT       */
int plf_ctx_set_gen_sink(plf_ctx *ctx, int sink);           /* plf_gen_sink; default WRITE    */
/* Kernel tuning: variant id and launch shape; 0 = library default.  See DESIGN.md.           */
int plf_ctx_set_tuning(plf_ctx *ctx, int variant, int threads_per_block, int blocks_per_sm);
unsigned plf_ctx_instances(const plf_ctx *ctx);

/* ---- instance buffers: xrt::bo x4 per instance (host_mem.cpp:123-133) --------------------- */

/* Allocates left, right, out and scaler device buffers for up to max_sites sites:
 * left (80+16*max_sites) floats, right (80|64 + 16*max_sites) floats, out 16*max_sites floats,
 * scaler max_sites bytes (tb.instance_size_*, include.h:173-179,222-239 -- in size_t here).   */
int plf_instance_alloc(plf_ctx *ctx, unsigned inst, size_t max_sites);
int plf_instance_free(plf_ctx *ctx, unsigned inst);

/* bo.write(host_ptr, bytes, 0) (host_mem.cpp:297-298): enqueue a host->device copy of `bytes`
 * bytes of the packed buffer at byte offset `offset` on the instance's stream.  The host buffer
 * must stay valid until plf_wait(); it is copied asynchronously when it is pinned
 * (plf_host_alloc / plf_host_register).                                                      */
int plf_write_left(plf_ctx *ctx, unsigned inst, const float *packed, size_t bytes, size_t offset);
int plf_write_right(plf_ctx *ctx, unsigned inst, const float *packed, size_t bytes, size_t offset);
/* Per-site integer weights (plf.cpp:63; host_mem.cpp:206-209 uses all ones).  wgt == NULL
 * restores the all-ones default (no weight traffic in the kernel).                          */
int plf_write_wgt(plf_ctx *ctx, unsigned inst, const int *wgt, size_t count);

/* run.set_arg(sites) + s2mm.start(); mm2sleft.start(); mm2sright.start() (host_mem.cpp:142-156,
 * 305): enqueue ONE fused kernel for `sites` sites on the instance's stream.                  */
int plf_run_async(plf_ctx *ctx, unsigned inst, size_t sites);
/* run.wait() x3 / instance_done[k].wait() (host_mem.cpp:305,323-325): block until everything
 * enqueued on the instance has finished; reports asynchronous CUDA errors.                   */
int plf_wait(plf_ctx *ctx, unsigned inst);

/* bo.read (host_mem.cpp:313-314): enqueue device->host copies of the result CLV (16 floats per
 * site) and of the per-site scaler bytes (0/1) (s2mm_memDNAwindowComb.cpp:96-97).             */
int plf_read_out(plf_ctx *ctx, unsigned inst, float *dst, size_t bytes, size_t offset);
int plf_read_scaler(plf_ctx *ctx, unsigned inst, char *dst, size_t bytes, size_t offset);
/* sum_j scaler[j]*wgt[j] (host_mem.cpp:384-388), accumulated inside the kernel.  Waits for the
 * instance, then returns the value of the LAST run.                                          */
int plf_scaler_increment(plf_ctx *ctx, unsigned inst, long long *increment);
/* GEN + DISCARD sink: checksum (sum of all outputs, fp32 accumulated in fp64) of the last run. */
int plf_gen_checksum(plf_ctx *ctx, unsigned inst, double *checksum);

/* Enqueue a timestamp on the instance's stream / read the time between two of them (ms), the
 * device-side analogue of timing_data{begin,t1,t2,end} (timing.h:25-52).                     */
int plf_mark(plf_ctx *ctx, unsigned inst, int mark_id);
int plf_elapsed_ms(plf_ctx *ctx, unsigned inst, int from_mark, int to_mark, float *ms);

/* Raw handles for callers that keep data on the device (next newview of a tree, benchmarks).
 * Any pointer argument may be NULL.  stream is a cudaStream_t.                                */
int plf_instance_device_ptrs(plf_ctx *ctx, unsigned inst, float **left, float **right,
                             float **out, unsigned char **scaler);
int plf_instance_stream(plf_ctx *ctx, unsigned inst, void **stream);

/* ---- streamed host path -------------------------------------------------
 * One newview over HOST-resident, unpacked arrays (the arguments of plf(), plf.h:1-5): the site range is
 * cut into chunks of chunk_sites (0 = auto: n/16 clamped to 256 Ki .. 2 Mi sites) that flow through three device buffers on three streams, so
 * the H2D copy of one chunk, the kernel of the previous one and the D2H copy of the one before overlap.
 * The analogue of the reference's NO_INTERMEDIATE_RESULTS=1 round-trip mode (host_mem.cpp:327-382)
 * without the packing pass, and not limited by device memory.  Blocks until x3 / scaler are complete.
 * ev[16], p_left[64], p_right[64] host; x1, x2, x3 host n_sites*16 floats (pinned => asynchronous
 * copies); scaler host n_sites chars or NULL; wgt host ints or NULL; increment may be NULL.           */
int plf_newview_stream(plf_ctx *ctx, const float *ev, const float *p_left, const float *p_right,
                       const float *x1, const float *x2, float *x3, char *scaler, const int *wgt,
                       size_t n_sites, size_t chunk_sites, long long *increment);

/* ---- pinned host memory -------------------------------------------------------------------- */
int plf_host_alloc(void **ptr, size_t bytes);
int plf_host_free(void *ptr);
int plf_host_register(void *ptr, size_t bytes);
int plf_host_unregister(void *ptr);

/* ---- the fused kernel on caller-owned device memory ---------------------------------------- */

typedef struct plf_launch_opts {
    int math_mode;          /* plf_math                                                          */
    int variant;            /* kernel variant, 0 = default.  Dynamically scheduled variants
                             * (K = 3, the default) take a work-counter pair from a per-device ring at
                             * launch time: when capturing launches into a CUDA graph that may be replayed
                             * concurrently with itself, pick a static variant (e.g. 1322).               */
    int threads_per_block;  /* 0 = default                                                       */
    int blocks_per_sm;      /* 0 = default (persistent grid = SMs x blocks_per_sm)               */
    int ev_per_category;    /* 0: ev is EV[16]; 1: ev is EV4[4][16], one matrix per category     */
    int flags;              /* plf_launch_flags, 0 = defaults                                    */
} plf_launch_opts;

/* plf_launch_opts.flags.  The ring kernels hand a shared-memory slot back to the bulk-copy engine either behind
 * fence.proxy.async (FENCED: the release as the PTX memory model words it; default of the DRAM-bound kernels) or
 * behind a data dependency on the loaded registers (DEP: no fence; default of the tree kernel, whose compressed-tip
 * levels the fence slows down, and of the tensor-core 20-state kernel, which it costs 3 %).  The environment variable PLF_SAFE_RELEASE=1 / =0 forces one of them for every
 * kernel of the process, so a suspected slot race can be bisected in the field without a rebuild.               */
typedef enum plf_launch_flags {
    PLF_LAUNCH_NO_PDL = 1,          /* do not launch with programmatic stream serialization            */
    PLF_LAUNCH_FENCED_RELEASE = 2,
    PLF_LAUNCH_DEP_RELEASE = 4,
    PLF_LAUNCH_SINGLE_CTA = 8       /* test hook: the whole launch on ONE block (its ring refills from L2 at full
                                     * speed -- the worst case for the slot-release race; tests/test_stress.py) */
} plf_launch_flags;

/* Newview of n sites.  ALL pointers are device pointers on `device`'s current context:
 *   x1,x2,x3 : 16 floats per site, 16-byte aligned; x3 may alias neither input
 *   scaler   : n bytes (0/1) or NULL
 *   ev,p_left,p_right : 16 / 64 / 64 floats (layouts of plf.cpp:37-38,47)
 *   wgt      : n ints or NULL (all ones)
 *   scaler_sum: one unsigned 64-bit counter the kernel ADDS sum(wgt over rescaled sites) to,
 *              or NULL.  The caller zeroes it.
 * opts may be NULL (defaults).  stream is a cudaStream_t (NULL = default stream).
 * Replaces plf() (plf.h:1-5) for device-resident data.                                        */
int plf_newview_device(const float *x1, const float *x2, float *x3, unsigned char *scaler,
                       const float *ev, const float *p_left, const float *p_right,
                       const int *wgt, size_t n, unsigned long long *scaler_sum,
                       const plf_launch_opts *opts, void *stream);

/* INPUT_SRC=gen analogue on caller-owned memory: no CLV is read; x3/scaler may be NULL when
 * sink == PLF_GEN_DISCARD, in which case the fp64 sum of all outputs is ADDED to *checksum
 * (device pointer, may be NULL).                                                             */
int plf_newview_gen_device(float *x3, unsigned char *scaler, size_t n,
                           unsigned long long *scaler_sum, double *checksum, int sink,
                           const plf_launch_opts *opts, void *stream);

/* The constant 16-float site patterns and the per-lane header the gen movers emit
 * (mm2sleft_genDNAwindowComb.cpp:44-49, mm2sright_genDNAwindowComb.cpp:45-50), as host arrays:
 * x1[16], x2[16], ev4[64], p_left[64], p_right[64] -- what a MEM run must be fed to reproduce
 * a GEN run.                                                                                 */
int plf_gen_pattern(float *x1, float *x2, float *ev4, float *p_left, float *p_right);

/* Synthetic stimulus generated on the device with a counter-based hash: the value
 * distribution of host_mem.cpp:198-204 (uniform(0,1); the left CLV of every 4th site times
 * 1e-12f).  Element e of x1/x2 depends only on (seed, first_site*16 + e), so any site range
 * can be generated independently on any GPU.                                                  */
int plf_generate_device(float *x1, float *x2, size_t first_site, size_t n, uint64_t seed,
                        void *stream);
/* The same generator evaluated on the host for a site range (for verification of slices). */
int plf_generate_host(float *x1, float *x2, size_t first_site, size_t n, uint64_t seed);

/* ---- chained newview over a tree (the caller of the path; SURVEY.md section 8f, BASELINE.json
 *      configs[4]).  The reference stops at one newview call (its README lists the surrounding
 *      RAxML machinery only as the origin of plf(), README.md:188-189,207-208); this is the natural
 *      next layer: a post-order traversal in which every inner node is one newview of its two
 *      children, per-site scaler COUNTS are carried upwards, and all nodes of one tree level run
 *      in a single launch.                                                                   ---- */
typedef struct plf_tree plf_tree;

/* Rooted binary tree in post-order.  Node ids: 0..n_tips-1 are tips, n_tips+k is inner node k.
 * Inner node k (0 <= k < n_tips-1) has children left[k], right[k], each a tip or an inner node
 * with a smaller index; the last inner node is the root.  Allocates all device memory:
 * n_tips tip CLVs, a recycled pool of inner CLVs + int32 scaler-count vectors, matrices.       */
int plf_tree_create(plf_tree **tree, int device, unsigned n_tips, const int *left, const int *right,
                    size_t n_sites);
/* Tip storage.  DENSE: a tip is a full CLV (64 B/site), written with
 * plf_tree_write_tip.  CODES: a tip is one state code per site (0..15, the 4-bit ambiguity code of a
 * DNA alignment) and its CLV is x[i][j][l] = tip_vector[code_i][l] for every category j (RAxML's
 * tipVector lookup) -- 1 B/site of HBM traffic and memory instead of 64.                        */
typedef enum plf_tip_format { PLF_TIPS_DENSE = 0, PLF_TIPS_CODES = 1 } plf_tip_format;
int plf_tree_create_ex(plf_tree **tree, int device, unsigned n_tips, const int *left, const int *right,
                       size_t n_sites, int tip_format);
/* STATES knob for trees: states = 4 (what plf_tree_create_ex gives) or 20.  A 20-state tree keeps dense tips (80 floats
 * per site), takes EV[400] and P_left / P_right [(n_tips-1)][1600], diag float[80] for plf_tree_evaluate_root, and runs
 * one launch of the 20-state kernel per inner node (captured in the same CUDA graph), carrying the scaler counts.   */
int plf_tree_create_states(plf_tree **tree, int device, unsigned n_tips, const int *left, const int *right,
                           size_t n_sites, int tip_format, int states);
/* CODES trees: n state codes of a tip starting at first_site; the 16 x 4 table tip_vector[code][state]. */
int plf_tree_write_tip_codes(plf_tree *tree, unsigned tip, const unsigned char *codes, size_t n,
                             size_t first_site);
int plf_tree_write_tip_vector(plf_tree *tree, const float *tip_vector);
int plf_tree_destroy(plf_tree *tree);
const char *plf_tree_last_error(const plf_tree *tree);
int plf_tree_set_math(plf_tree *tree, int math_mode);
/* u: 0 = automatic (352-site stages, 4 rows per warp, 3 deep), else the batch kernel shape (1: 128-site stages,
 * 2: 256-site, 3: 240-site stages with 2 rows per warp, 4: 352-site stages 2 deep).
 * chunk: 0 = automatic, else consecutive stages dealt to a CTA at a time (1 = fully interleaved). */
int plf_tree_set_tuning(plf_tree *tree, int u, int chunk);
/* Device pointer of a tip CLV (n_sites*16 floats) for callers that produce tips on the device. */
int plf_tree_tip_ptr(plf_tree *tree, unsigned tip, float **clv);
/* Host -> device copy into a tip CLV (bytes at byte offset), asynchronous on the tree's stream. */
int plf_tree_write_tip(plf_tree *tree, unsigned tip, const float *clv, size_t bytes, size_t offset);
/* EV[16] shared by all nodes; P_left/P_right[(n_tips-1)][64], one pair per inner node.         */
int plf_tree_write_matrices(plf_tree *tree, const float *ev, const float *p_left, const float *p_right);
/* Site weights for the total (NULL = all ones). */
int plf_tree_write_wgt(plf_tree *tree, const int *wgt);
/* Enqueue the whole traversal (one launch per tree level, replayed from a CUDA graph). */
int plf_tree_run_async(plf_tree *tree);
int plf_tree_wait(plf_tree *tree);
/* Root CLV (16 floats/site) and accumulated per-site scaler counts for sites [first, first+n). */
int plf_tree_read_root(plf_tree *tree, float *clv, int *scaler_counts, size_t first_site, size_t n);
/* Debug/verification: CLV and counts of any inner node are only defined for the root after a run
 * (inner buffers are recycled); this returns sum over all newviews of sum_i wgt[i]*rescaled(i). */
int plf_tree_total_scalings(plf_tree *tree, long long *total);
/* Schedule facts: number of levels (= launches per traversal), CLV slots in the recycled pool,
 * device bytes held, and algorithmic HBM bytes one traversal moves (205 B per site per node).  */
int plf_tree_info(plf_tree *tree, unsigned *levels, unsigned *clv_slots, size_t *device_bytes,
                  size_t *traversal_bytes);
/* Device time of the last traversal in milliseconds (events around the graph launch). */
int plf_tree_last_ms(plf_tree *tree, float *ms);

/* ---- root log-likelihood across a branch ---------------------------
 * The step after the path in RAxML-like codes (standard-RAxML evaluateGTRGAMMA; not in the
 * reference repository, which stops at newview):
 *   lnL = sum_i wgt[i] * ( log(0.25*|sum_{j,k} x1[i,j,k]*x2[i,j,k]*diag[j,k]|) + (cnt1[i]+cnt2[i])*log(2^-32) )
 * All pointers are DEVICE pointers: x1,x2 CLVs (16-byte aligned), diag float[16] = [category][state],
 * cnt1/cnt2 int32 per-site scaler counts or NULL, wgt int32 or NULL.  The result is ADDED to *lnl
 * (device double; the caller zeroes it).  fp64 products, log and accumulation.                  */
int plf_evaluate_device(const float *x1, const float *x2, const int *cnt1, const int *cnt2,
                        const int *wgt, const float *diag, size_t n, double *lnl, void *stream);
/* The same for S = 4 or 20 states: x1, x2 are n*4*S floats, diag is float[4*S] = [category][state].  The sum is
 * reproducible bit for bit from run to run (fixed-order two-stage reduction, one addition to *lnl per launch).      */
int plf_evaluate_states_device(int states, const float *x1, const float *x2, const int *cnt1, const int *cnt2,
                               const int *wgt, const float *diag, size_t n, double *lnl, void *stream);
/* The same across the ROOT branch of a traversed tree: between the two children of the last inner
 * node, with their accumulated scaler counts and the tree's weights.  diag is a HOST float[16].
 * Must follow a completed plf_tree_run_async (the children's buffers are still intact then).  */
int plf_tree_evaluate_root(plf_tree *tree, const float *diag, double *lnl);

/* ---- general state count: the STATES knob ---------------------------
 * The reference builds with STATES=DNA only (Makefile knob, README.md:67); "any other type of data
 * with more or fewer states" is anticipated (README.md:36) and "Implement protein-based PLF" is an
 * open to-do (README.md:202).  These entry points run plf() (app/src/plf.cpp:19-65) with the state
 * count S in {4 (DNA), 20 (protein)}:
 *   x1,x2,x3 : DEVICE float[n*4*S]  [site][category][state], 16-byte aligned
 *   ev       : HOST float[S*S]      [k][l]
 *   p_left/p_right : HOST float[4*S*S] [category][k][l]
 *   scaler   : DEVICE uint8[n] or NULL; wgt DEVICE int32[n] or NULL; scaler_sum DEVICE counter or NULL
 * The matrices are HOST arrays (they are host arrays in the reference's host too, host_mem.cpp:183-197): they travel
 * by value into a per-stream staging record on the device and the kernel reads them from there; they are consumed
 * before the call returns, and the call can be captured into a CUDA graph.  (The instance API of a 20-state context,
 * plf_ctx_create_states, keeps the matrices in its device buffers and uploads nothing per launch.)  A site rescales when all 4*S entries are below 2^-32.  S = 4 runs the DNA
 * kernel of plf_newview_device.  For S = 20, opts->variant is the number of sites per lane of the register tile
 * (1, 2 or 4; threads_per_block 512 / 256,384 / 128,256), or 9 = the TENSOR-CORE kernel (tcgen05.mma kind::tf32 with the
 * 3xTF32 split, fp32 accumulators in tensor memory, operands by TMA; PLF_MATH_FMA only -- tensor cores cannot reproduce
 * the reference's rounding sequence, so PLF_MATH_STRICT never uses them; <= 1e-5 relative, measured 2e-6; fp32 DENORMAL
 * operands read as zero there, the CUDA-core kernels keep them); 0 = the fastest measured kernel for the math
 * mode: the register-tile kernel in strict mode, the tensor-core kernel in FMA mode for calls of >= 16384 sites.      */
int plf_newview_states_device(int states, const float *x1, const float *x2, float *x3,
                              unsigned char *scaler, const float *ev, const float *p_left,
                              const float *p_right, const int *wgt, size_t n,
                              unsigned long long *scaler_sum, const plf_launch_opts *opts,
                              void *stream);
/* Stimulus of host_mem.cpp:198-204 for S states (uniform(0,1); the left CLV of every 4th site
 * tiny, so exactly ceil(n/4) sites rescale): identical to plf_generate_device for S = 4.        */
int plf_generate_states_device(int states, float *x1, float *x2, size_t first_site, size_t n,
                               uint64_t seed, void *stream);
int plf_generate_states_host(int states, float *x1, float *x2, size_t first_site, size_t n,
                             uint64_t seed);
/* Introspection of the 20-state kernel chosen for (math_mode, variant, threads_per_block). */
int plf_states_kernel_info(int states, int math_mode, int variant, int threads_per_block,
                           int *regs_per_thread, int *block_threads, size_t *smem_bytes,
                           int *tile_sites);

/* ---- several GPUs of one box from one host process -----------------------------------------------
 * The reference scales over independent accelerator instances by contiguous site ranges: ceil(n / parts) sites each,
 * the last part takes the remainder (app/src/include.h:181-192), and sums the per-instance scaler increments on the
 * host (app/src/host_mem.cpp:384-388).  plf_multi applies the same rule to the GPUs of a box: one plf_ctx per GPU
 * (the whole instance API applies to plf_multi_ctx(m, rank)), no data-path exchange between GPUs, and ONE collective:
 * the final sum of the per-GPU scaler increments (int64) and log-likelihoods (fp64) with ncclAllReduce over
 * NVLink/NVSwitch.  NCCL is resolved with dlopen("libnccl.so.2") at creation; a one-GPU plf_multi does not need it. */
typedef struct plf_multi plf_multi;
int plf_multi_create(plf_multi **multi, const int *devices, int n_devices, unsigned n_instances, int layout, int input_src);
int plf_multi_create_states(plf_multi **multi, const int *devices, int n_devices, unsigned n_instances, int layout,
                            int input_src, int states);
int plf_multi_destroy(plf_multi *multi);
const char *plf_multi_last_error(const plf_multi *multi);
int plf_multi_size(const plf_multi *multi);
plf_ctx *plf_multi_ctx(plf_multi *multi, int rank);
/* The reference's split rule for part `part` of `n_parts` (include.h:181-192): first site and site count. */
int plf_multi_partition(size_t n_sites, int n_parts, int part, size_t *first, size_t *count);
/* One newview over HOST arrays (the arguments of plf(), plf.h:1-5), the sites partitioned over the GPUs, every GPU
 * running the streamed path (plf_newview_stream) on its range from its own host thread.  *increment (may be NULL) is the
 * NCCL-reduced total of the per-GPU scaler increments.  Blocks until x3 / scaler are complete.                      */
int plf_multi_newview(plf_multi *multi, const float *ev, const float *p_left, const float *p_right, const float *x1,
                      const float *x2, float *x3, char *scaler, const int *wgt, size_t n_sites, long long *increment);
/* The optional final reduction for callers that drive the per-GPU contexts themselves: increments[rank] / lnl[rank]
 * (either may be NULL) are summed over the ranks with one grouped ncclAllReduce per type; every rank's result is read
 * back and must agree.                                                                                              */
int plf_multi_reduce(plf_multi *multi, const long long *increments, const double *lnl, long long *increment_total,
                     double *lnl_total);
/* Number of GPUs, NCCL version in use (0 for one GPU), number of NCCL reductions issued so far. */
int plf_multi_info(const plf_multi *multi, int *n_devices, int *nccl_version, unsigned long long *reductions);

/* ---- device memory for callers that own their buffers ------------------------------------------
 * The instance API keeps device memory inside the context (the xrt::bo role).  The *_device entry points
 * (plf_newview_device, plf_newview_states_device, plf_evaluate_device) work on caller-owned device memory;
 * a C/C++ host without its own CUDA code gets it here.  Copies are asynchronous on `stream` when the host
 * side is pinned (plf_host_alloc); stream NULL = the default stream.                                  */
int plf_device_malloc(int device, void **ptr, size_t bytes);
int plf_device_free(void *ptr);
int plf_memcpy_h2d(void *dst_device, const void *src_host, size_t bytes, void *stream);
int plf_memcpy_d2h(void *dst_host, const void *src_device, size_t bytes, void *stream);
int plf_memset_device(void *dst_device, int value, size_t bytes, void *stream);
int plf_stream_sync(void *stream);

/* Named profiler ranges (NVTX) for host code without CUDA headers: the reference brackets its run loop with
 * xrt::profile::user_range("roundtrip_exec_time") (host_mem.cpp:273,282,395); the hosts here use the same name, so an
 * nsys / ncu timeline shows the same region.                                                                      */
int plf_range_push(const char *name);
int plf_range_pop(void);

/* Bare host-link probe, the yardstick of the host-buffer path: `reps` rounds of one pinned host->device copy of
 * h2d_bytes and one device->host copy of d2h_bytes, each cut into `pieces` cudaMemcpyAsync calls, the two directions
 * on two streams at once, no kernels.  host_in / host_out are pinned host buffers of those sizes (NULL: allocated
 * and freed inside, untimed); device buffers are internal.  *seconds is the wall time of the timed rounds (one
 * untimed warm-up round first).  The PLF round trip moves 128 B/site in and 65 B/site out (the reference's own note
 * on its PCIe limit: README.md:204).                                                                              */
int plf_probe_host_link(int device, void *host_in, size_t h2d_bytes, void *host_out, size_t d2h_bytes, int reps,
                        int pieces, double *seconds);

/* Process-wide override of the ring-slot release mechanism (see plf_launch_flags): 1 = fenced everywhere,
 * 0 = data dependency everywhere, -1 = every kernel family's own default (also what PLF_SAFE_RELEASE unset means). */
int plf_set_release_mode(int mode);

/* Largest tile the library was compiled for, number of SMs etc. -- introspection for benches. */
int plf_kernel_info(int variant, int math_mode, int *regs_per_thread, int *threads_per_block,
                    int *blocks_per_sm, int *num_sms);
/* Number of kernel launches this library has issued in this process (all contexts). */
unsigned long long plf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200PLF_H */
