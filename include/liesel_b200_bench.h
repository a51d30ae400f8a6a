/*
 * liesel_b200_bench.h - C ABI of libliesel_b200_bench.so, the TOOLS library next to the product
 * library (bench.py's live pipe-peak measurement, tools/microbench.py, tools/tc_probe.py). Nothing
 * here is on the evaluation path; the product library does not depend on it.
 */
#ifndef LIESEL_B200_BENCH_H
#define LIESEL_B200_BENCH_H

#include "liesel_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Micro-benchmarks giving MEASURED pipe peaks (the FP32 denominators MEASURED_PEAKS.json lacks):
 * kind 0 = FP32 FMA, 1 = packed fma.rn.f32x2, 2 = MUFU ex2, 3 = shared-memory 4-byte loads,
 * 4 = FFMA2 and FFMA interleaved, 5/6 = FFMA2 with three distinct / one shared register operand,
 * 7/8 = the same for scalar FFMA. *work (host) receives the operations one launch performs.
 */
int lslb_debug_pipe_peak(int kind, int blocks, int iters, float* sink, double* work,
                         lslb_stream_t stream);
/* tcgen05 building block under test: D [128 x N] = A [128 x K] . B [N x K]^T via TMA + tcgen05.mma
 * kind::tf32 (split bit 0: three-pass hi/lo split, fp32-level accuracy; bit 1: K chunks of 16 floats
 * with the 64-byte swizzle instead of 32 floats with the 128-byte swizzle). N in {64,128,256}, K % 32 == 0. */
int lslb_debug_tc_probe(const float* A, const float* B, int N, int K, int split, float* D,
                        lslb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIESEL_B200_BENCH_H */
