// simt_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A single-threaded SIMT emulator that lets g++ compile the CUDA sources of
// dantzig_b200 (dz_kernel.cu, dz_capi.cu with -DDZ_EMU) and run the very same
// kernel code on the CPU, one fiber per CUDA thread, so that the kernel LOGIC
// (and the off-by-default build variants) can be checked against the oracle in
// the CPU test suite.  It is not a backend: the product library never defines
// DZ_EMU, nothing under dantzig_b200/ builds or loads the emulated library, and
// it is orders of magnitude slower than the oracle.  Only tests/ builds it
// (tests/emu/build_emu.py) and loads it (through the DZ_LIB override).
//
// Model: a kernel launch runs its blocks one after the other; the threads of a
// block are fibers on private stacks, switched cooperatively.  A thread runs
// until it reaches a warp collective (__ballot_sync, __shfl_*_sync,
// __reduce_*_sync, __syncwarp) or __syncthreads; when every live thread of the
// warp (block) waits at the same kind of operation the results are computed and
// the threads resume.  Memory is plain host memory; atomics are plain updates.
// This yields ONE legal interleaving: it checks arithmetic, indexing and
// control flow, not data races or memory-model subtleties.
#ifndef DZ_SIMT_EMU_H
#define DZ_SIMT_EMU_H

#ifndef __x86_64__
#error "simt_emu.h: the fiber switch is written for x86-64"
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- qualifiers ---------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {

enum Op { OP_NONE = 0, OP_BALLOT, OP_SHFL, OP_RMAX_U, OP_RMIN_U, OP_RMAX_I, OP_RMIN_I, OP_SYNCWARP, OP_SYNCTHREADS, OP_GRIDSYNC, OP_YIELD };

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};

struct Thread {
    void *sp = nullptr; // saved stack pointer while switched out
    char *stack = nullptr;
    bool done = false;
    Op wait = OP_NONE;
    uint64_t in = 0, out = 0;
    int aux = 0;
    Dim3 tid;
};

struct Block {
    std::vector<Thread> th;
    unsigned char *smem = nullptr;
    Dim3 bid, bdim, gdim;
    std::function<void()> body;
};

extern Block *g_blk;   // block being executed
extern Thread *g_cur;  // thread being executed
extern void *g_sched_sp;
extern long long g_clock;

extern "C" void emu_switch(void **save_sp, void *load_sp);

inline uint64_t collective(Op op, uint64_t in, int aux) {
    Thread *t = g_cur;
    t->wait = op;
    t->in = in;
    t->aux = aux;
    emu_switch(&t->sp, g_sched_sp);
    return t->out;
}

void run_block(Block &b);
// Cooperative launch: every block of the grid is resident at once (fibers of all blocks are
// scheduled round robin) and grid_sync() is a barrier over all of them.
void run_grid(std::vector<Block> &blocks);

template <class Kern, class... Args>
int launch(Kern kern, int grid, int block, size_t smem_bytes, Args... args) {
    for (int bx = 0; bx < grid; ++bx) {
        Block b;
        b.bid.x = (unsigned)bx;
        b.bdim.x = (unsigned)block;
        b.gdim.x = (unsigned)grid;
        void *sm = nullptr;
        if (posix_memalign(&sm, 128, smem_bytes + 128) != 0) return 2;
        std::memset(sm, 0xA5, smem_bytes + 128); // shared memory starts undefined
        b.smem = static_cast<unsigned char *>(sm);
        b.th.resize((size_t)block);
        for (int t = 0; t < block; ++t) b.th[(size_t)t].tid.x = (unsigned)t;
        b.body = [=]() { kern(args...); };
        run_block(b);
        std::free(sm);
    }
    return 0;
}

template <class Kern, class... Args>
int launch_coop(Kern kern, int grid, int block, size_t smem_bytes, Args... args) {
    std::vector<Block> blocks((size_t)grid);
    std::vector<void *> mem;
    for (int bx = 0; bx < grid; ++bx) {
        Block &b = blocks[(size_t)bx];
        b.bid.x = (unsigned)bx;
        b.bdim.x = (unsigned)block;
        b.gdim.x = (unsigned)grid;
        void *sm = nullptr;
        if (posix_memalign(&sm, 128, smem_bytes + 128) != 0) return 2;
        std::memset(sm, 0xA5, smem_bytes + 128);
        mem.push_back(sm);
        b.smem = static_cast<unsigned char *>(sm);
        b.th.resize((size_t)block);
        for (int t = 0; t < block; ++t) b.th[(size_t)t].tid.x = (unsigned)t;
        b.body = [=]() { kern(args...); };
    }
    run_grid(blocks);
    for (void *m : mem) std::free(m);
    return 0;
}

inline unsigned char *dyn_smem() { return g_blk->smem; }
inline void grid_sync() { collective(OP_GRIDSYNC, 0, 0); }
// spin-wait loops on flags other threads set must give the other fibers a turn
inline void yield() { collective(OP_YIELD, 0, 0); }

} // namespace emu

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_blk->bid)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)

// ---- warp / block collectives ---------------------------------------------------
inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)emu::collective(emu::OP_BALLOT, pred ? 1 : 0, 0); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::collective(emu::OP_SYNCWARP, 0, 0); }
inline void __syncthreads() { emu::collective(emu::OP_SYNCTHREADS, 0, 0); }

template <class T> inline uint64_t emu_bits(T v) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "shuffle operand");
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> inline T emu_unbits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    (void)width;
    return emu_unbits<T>(emu::collective(emu::OP_SHFL, emu_bits(v), src & 31));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lanemask, int width = 32) {
    (void)width;
    const int lane = (int)(emu::g_cur->tid.x & 31u);
    return emu_unbits<T>(emu::collective(emu::OP_SHFL, emu_bits(v), (lane ^ lanemask) & 31));
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    (void)width;
    const int lane = (int)(emu::g_cur->tid.x & 31u);
    const int src = lane - (int)delta;
    return emu_unbits<T>(emu::collective(emu::OP_SHFL, emu_bits(v), src < 0 ? lane : src));
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return (unsigned)emu::collective(emu::OP_RMAX_U, v, 0); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return (unsigned)emu::collective(emu::OP_RMIN_U, v, 0); }
inline int __reduce_max_sync(unsigned, int v) { return (int)(int64_t)emu::collective(emu::OP_RMAX_I, (uint64_t)(int64_t)v, 0); }
inline int __reduce_min_sync(unsigned, int v) { return (int)(int64_t)emu::collective(emu::OP_RMIN_I, (uint64_t)(int64_t)v, 0); }

// ---- arithmetic and bit intrinsics (compile with -ffp-contract=off) ---------------
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __double2hiint(double d) { return (int)(uint32_t)(emu_bits(d) >> 32); }
inline int __double2loint(double d) { return (int)(uint32_t)(emu_bits(d) & 0xffffffffu); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> inline T __ldg(const T *p) { return *p; }
inline long long clock64() { return ++emu::g_clock; }
using std::isfinite;
using std::max;
using std::min;
inline int min(int a, unsigned b) { return (unsigned)a < b ? a : (int)b; }

struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }

// single-threaded execution: atomics are plain read-modify-writes
template <class T> inline T atomicAdd(T *p, T v) {
    const T old = *p;
    *p = old + v;
    return old;
}
inline unsigned atomicOr(unsigned *p, unsigned v) {
    const unsigned old = *p;
    *p = old | v;
    return old;
}
inline double __longlong_as_double(long long v) { return emu_unbits<double>((uint64_t)v); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = std::max(old, v);
    return old;
}
inline long long __double_as_longlong(double v) { return (long long)emu_bits(v); }
inline unsigned atomicExch(unsigned *p, unsigned v) {
    const unsigned old = *p;
    *p = v;
    return old;
}
inline void __threadfence() {}
inline void __threadfence_block() {}
inline int atomicMin(int *p, int v) {
    const int old = *p;
    *p = std::min(old, v);
    return old;
}
inline int atomicMax(int *p, int v) {
    const int old = *p;
    *p = std::max(old, v);
    return old;
}

// ---- the slice of the CUDA runtime API the library uses, on host memory ------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void *cudaStream_t;
struct EmuEvent {
    std::chrono::steady_clock::time_point t;
};
typedef EmuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[64];
    size_t sharedMemPerBlockOptin, sharedMemPerMultiprocessor, totalGlobalMem;
    int multiProcessorCount, major, minor, clockRate;
};
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulator error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorEmu; }
inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { // the B200 figures plan_launch sizes against
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "SIMT emulator (B200 limits)");
    p->sharedMemPerBlockOptin = 232448;
    p->sharedMemPerMultiprocessor = 233472;
    p->totalGlobalMem = (size_t)180 << 30;
    p->multiProcessorCount = 148;
    p->major = 10;
    p->minor = 0;
    p->clockRate = 1965000;
    return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) {
    void *q = nullptr;
    if (posix_memalign(&q, 256, n ? n : 1) != 0) return cudaErrorEmu;
    std::memset(q, 0xA5, n); // device memory starts undefined
    *p = static_cast<T *>(q);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    if (n) std::memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) {
    return cudaMemcpy(d, s, n, k);
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) {
    if (n) std::memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
    if (n) std::memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t *s) {
    *s = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new EmuEvent();
    return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}

#endif // DZ_SIMT_EMU_H
