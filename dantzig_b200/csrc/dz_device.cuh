// dz_device.cuh -- device helpers shared by the kernels of dantzig_b200
// (dz_kernel.cu: the general persistent kernel; dz_core.cu: the on-chip coupled-core kernel).
#ifndef DZ_DEVICE_CUH
#define DZ_DEVICE_CUH

#include "dz_internal.h"

#ifdef DZ_EMU // test-only: g++ build of this file on the SIMT emulator of tests/emu (never the product)
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace dz {

namespace {

constexpr int kMaxWarps = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double load_ref(const double *__restrict__ th, int ref) {
    if (ref < 0) return 0.0;
    // theta[0] is the constant 1.0 by contract (include/dantzig_b200.h): no load
    const double v = (ref >> 1) == 0 ? 1.0 : __ldg(th + (ref >> 1));
    return (ref & 1) ? -v : v;
}

// Total order used by every arg-max on the path: larger key first, then the
// smaller index ("first index wins", simplex.rs:432-435, linalg.rs:100-105).
__device__ __forceinline__ bool beats(double k2, int i2, double k1, int i1) {
    return i2 >= 0 && (i1 < 0 || k2 > k1 || (k2 == k1 && i2 < i1));
}

template <int N> struct Cand {
    double key[N];
    int idx[N];
};

// Block-wide arg-max of N independent (key, idx) candidates with one barrier.
// red_key/red_idx hold [2][N][kMaxWarps]; `parity` alternates between calls so
// a slot is never rewritten before every thread has read it.
// Only warps 0..nparts-1 hold candidates (warp-uniform), the others just wait
// for the result.
template <int N>
__device__ __forceinline__ void block_argmax(Cand<N> &c, double *red_key, int *red_idx,
                                             int &parity, int nparts, int tid, bool wm) {
    const int warp = tid >> 5, lane = tid & 31;
    const int nwarps = nparts;
    if (warp < nparts) {
#pragma unroll
        for (int n = 0; n < N; ++n) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double k2 = __shfl_xor_sync(kFull, c.key[n], off);
                const int i2 = __shfl_xor_sync(kFull, c.idx[n], off);
                if (beats(k2, i2, c.key[n], c.idx[n])) {
                    c.key[n] = k2;
                    c.idx[n] = i2;
                }
            }
        }
    }
    if (wm) return; // single warp: the butterfly left the result in every lane
    double *rk = red_key + (size_t)parity * N * kMaxWarps;
    int *ri = red_idx + (size_t)parity * N * kMaxWarps;
    if (lane == 0 && warp < nparts) {
#pragma unroll
        for (int n = 0; n < N; ++n) {
            rk[n * kMaxWarps + warp] = c.key[n];
            ri[n * kMaxWarps + warp] = c.idx[n];
        }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        double bk = rk[n * kMaxWarps];
        int bi = ri[n * kMaxWarps];
        for (int w = 1; w < nwarps; ++w) {
            const double k2 = rk[n * kMaxWarps + w];
            const int i2 = ri[n * kMaxWarps + w];
            if (beats(k2, i2, bk, bi)) {
                bk = k2;
                bi = i2;
            }
        }
        c.key[n] = bk;
        c.idx[n] = bi;
    }
    parity ^= 1;
}

} // namespace
} // namespace dz

#endif // DZ_DEVICE_CUH
