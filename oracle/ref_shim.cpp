// oracle/ref_shim.cpp -- extern "C" doorway to the REFERENCE's own plf().
// TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together with the reference's
// app/src/plf.cpp *where that file lies* (/root/reference, never copied into this repo) by
// oracle/Makefile into oracle/_ref/libplf_ref.so.  It exists so that tests and the CPU-baseline
// leg of bench.py can call the unmodified reference function (app/src/plf.h:1-5) via ctypes.
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>

#include "plf.h"  // resolved with -I/root/reference/app/src

extern "C" {

// One call of the unmodified reference on [0,n).  n is an int in the reference signature.
int64_t plf_ref_newview(float *x1, float *x2, float *x3, float *ev, int n,
                        float *left, float *right, int *wgt)
{
    int inc = 0;
    plf(x1, x2, x3, ev, n, left, right, wgt, inc);
    return inc;
}

// "All host cores" figure: every thread runs the unmodified function on its own contiguous
// site range (sites are independent; see SURVEY.md section 8d, CPU baseline timing).
int64_t plf_ref_newview_mt(float *x1, float *x2, float *x3, float *ev, size_t n,
                           float *left, float *right, int *wgt, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (n == 0) return 0;
    if (static_cast<size_t>(nthreads) > n) nthreads = static_cast<int>(n);
    std::vector<int> incs(nthreads, 0);
    std::vector<std::thread> pool;
    const size_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const size_t lo = static_cast<size_t>(t) * chunk;
        if (lo >= n) break;
        const size_t cnt = (n - lo < chunk) ? (n - lo) : chunk;
        pool.emplace_back([=, &incs] {
            plf(x1 + lo * 16, x2 + lo * 16, x3 + lo * 16, ev, static_cast<int>(cnt),
                left, right, wgt + lo, incs[t]);
        });
    }
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (int v : incs) total += v;
    return total;
}

}  // extern "C"
