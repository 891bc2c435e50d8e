/*
 * oracle/plf_oracle.h -- CPU restatement of the reference PLF "newview" step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and there only as the checker or the reported CPU baseline.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 *   (1) bit-for-bit against the reference's own golden vectors aie/data/golden{0..3}.txt
 *       (fixture tests/golden/aie_kat.json),
 *   (2) bit-for-bit against outputs of the reference's own plf() (app/src/plf.cpp:8-68)
 *       compiled in place into oracle/_ref/ (fixtures tests/golden/ref_*.npz, and live
 *       whenever oracle/_ref/libplf_ref.so is present).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef PLF_ORACLE_H
#define PLF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Newview for n sites; follows app/src/plf.cpp:19-65 operation for operation
 * (strict fp32, accumulators start at +0.0f, l-then-k summation order, no FMA).
 *
 *   x1,x2,x3 : float[n*16], site-major [site][category j][state l]   (plf.cpp:21-23)
 *   ev       : float[16]  [k][l]                                      (plf.cpp:47)
 *   left/right: float[64] [j][k][l]                                   (plf.cpp:37-38)
 *   wgt      : int[n] or NULL (NULL == all ones, host_mem.cpp:206-209)
 *   scaler   : uint8[n] or NULL; 1 where the site was rescaled -- the per-site byte the
 *              output mover emits (hls/src/s2mm_memDNAwindowComb.cpp:77-85,97)
 * Returns sum of wgt[i] over rescaled sites (plf.cpp:63,66), in 64 bits.             */
int64_t plf_oracle_newview(const float *x1, const float *x2, float *x3,
                           const float *ev, size_t n,
                           const float *left, const float *right,
                           const int *wgt, unsigned char *scaler);

/* Same, but every category j uses its own back-transform matrix ev4[j][k][l].  This is
 * what the INPUT_SRC=gen movers feed the AIE lanes (each lane receives its own slice of
 * the constant pattern as "EV": hls/src/mm2sleft_genDNAwindowComb.cpp:44-84).          */
int64_t plf_oracle_newview_ev4(const float *x1, const float *x2, float *x3,
                               const float *ev4, size_t n,
                               const float *left, const float *right,
                               const int *wgt, unsigned char *scaler);

/* Packed-buffer front end: the device buffer format of the reference's input movers.
 *   left_buf  = [EV16 | P_left64 | CLV n*16]                  (host_mem.cpp:231-233,
 *                                                              mm2sleft_memDNAwindowComb.cpp:32-42,86)
 *   right_buf = Comb: [EV16 | P_right64 | CLV]  Sep: [P_right64 | CLV]
 *                                                             (host_mem.cpp:234-241)
 * layout: 0 = Comb, 1 = Sep.                                                           */
int64_t plf_oracle_newview_packed(const float *left_buf, const float *right_buf,
                                  int layout, size_t n, float *out,
                                  const int *wgt, unsigned char *scaler);

/* Host-side scaler reduction, host_mem.cpp:384-388. */
int64_t plf_oracle_scaler_increment(const unsigned char *scaler, const int *wgt, size_t n);

/* 4x4 transpose the input movers apply to each P matrix: out[4c+r] = in[4r+c]
 * (hls/src/transpose.cpp:6-24).  Used to read the (pre-transposed) aie/data fixtures.   */
void plf_oracle_transpose4(const float *in, float *out);

/* Multi-threaded driver for the CPU baseline: splits [0,n) into nthreads contiguous
 * ranges and runs plf_oracle_newview on each (sites are independent).                   */
int64_t plf_oracle_newview_mt(const float *x1, const float *x2, float *x3,
                              const float *ev, size_t n,
                              const float *left, const float *right,
                              const int *wgt, unsigned char *scaler, int nthreads);

/* General state count S (4 = DNA, 20 = protein; the reference's STATES knob, README.md:36,67,202):
 * plf.cpp:19-65 with "4 states" replaced by S.  x1,x2,x3: float[n*4*S] [site][category][state];
 * ev: float[S*S] [k][l]; left/right: float[4*S*S] [j][k][l].  PINNED for S = 4 (bit-identical to
 * the reference's plf(), tests/test_protein.py); PARITY UNPINNED for S = 20: the same code path, but
 * the reference has no protein implementation to pin against.  Returns -1 for an unsupported S.   */
int64_t plf_oracle_newview_states(int S, const float *x1, const float *x2, float *x3,
                                  const float *ev, size_t n,
                                  const float *left, const float *right,
                                  const int *wgt, unsigned char *scaler);
int64_t plf_oracle_newview_states_mt(int S, const float *x1, const float *x2, float *x3,
                                     const float *ev, size_t n, const float *left, const float *right,
                                     const int *wgt, unsigned char *scaler, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
