// host/golden_plf.h -- the host test bench's CPU golden model, used ONLY in the verification
// phase of host_mem (the role app/src/plf.cpp plays in the reference host, host_mem.cpp:403-442).
// It never produces a result the program returns: it is compiled out with NO_CORRECTNESS_CHECK=1
// and nothing in libb200plf.so links it.
#pragma once
#include <cstddef>

namespace plfhost {

// Same contract as the reference's plf() (app/src/plf.h:1-5) with size_t n and a 64-bit
// increment; additionally reports per-site scaler bytes when `scaler` is non-NULL.
void golden_plf(const float *x1, const float *x2, float *x3, const float *ev, size_t n,
                const float *left, const float *right, const int *wgt, long long &scaler_increment,
                unsigned char *scaler = nullptr);

// The same loop nest with the state count as a parameter (4 = DNA, 20 = AA): x1,x2,x3 float[n*4*S]
// [site][category][state], ev float[S*S] [k][l], left/right float[4*S*S] [category][k][l].  S <= 32.
void golden_plf_states(unsigned S, const float *x1, const float *x2, float *x3, const float *ev, size_t n,
                       const float *left, const float *right, const int *wgt, long long &scaler_increment,
                       unsigned char *scaler = nullptr);

}  // namespace plfhost
