// simt_emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler of the SIMT emulator
// (see simt_emu.h).  One OS thread; every CUDA thread of the running block is a
// fiber that is switched out at warp collectives and block barriers.
#include "simt_emu.h"

namespace emu {

Block *g_blk = nullptr;
Thread *g_cur = nullptr;
void *g_sched_sp = nullptr;
long long g_clock = 0;

// Save the callee-saved registers and the stack pointer of the running context,
// load another one.  A fresh fiber's stack is laid out so that the final `ret`
// enters fiber_main with the alignment a `call` would have produced.
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch, @function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch, .-emu_switch
)");

namespace {

constexpr size_t kStackBytes = 512 * 1024;

void fiber_main() {
    g_blk->body();
    Thread *t = g_cur;
    t->done = true;
    t->wait = OP_NONE;
    emu_switch(&t->sp, g_sched_sp);
    std::abort(); // a finished fiber is never resumed
}

unsigned long long g_ops[16] = {0}; // warp-level collectives resolved, per kind (EMU_STATS=1 prints them at exit)

struct StatsPrinter {
    ~StatsPrinter() {
        if (!std::getenv("EMU_STATS")) return;
        static const char *names[] = {"none", "ballot", "shfl", "rmax_u", "rmin_u", "rmax_i", "rmin_i", "syncwarp", "syncthreads"};
        std::fprintf(stderr, "simt_emu collectives:");
        for (int i = 1; i <= 8; ++i) std::fprintf(stderr, " %s=%llu", names[i], g_ops[i]);
        std::fprintf(stderr, "\n");
    }
} g_stats_printer;

void resolve_warp(Thread **lane, int n_lanes, Op op) {
    ++g_ops[op];
    uint64_t acc = 0;
    bool first = true;
    for (int l = 0; l < n_lanes; ++l) {
        Thread *t = lane[l];
        if (!t) continue;
        switch (op) {
        case OP_BALLOT: acc |= (t->in & 1u) << l; break;
        case OP_RMAX_U: acc = first ? t->in : std::max<uint64_t>(acc, t->in); break;
        case OP_RMIN_U: acc = first ? t->in : std::min<uint64_t>(acc, t->in); break;
        case OP_RMAX_I: acc = first ? t->in : (uint64_t)std::max<int64_t>((int64_t)acc, (int64_t)t->in); break;
        case OP_RMIN_I: acc = first ? t->in : (uint64_t)std::min<int64_t>((int64_t)acc, (int64_t)t->in); break;
        default: break;
        }
        first = false;
    }
    for (int l = 0; l < n_lanes; ++l) {
        Thread *t = lane[l];
        if (!t) continue;
        if (op == OP_SHFL) {
            Thread *s = (t->aux >= 0 && t->aux < n_lanes) ? lane[t->aux] : nullptr;
            t->out = s ? s->in : t->in; // an exited source lane: own value (undefined on hardware)
        } else {
            t->out = acc;
        }
    }
    for (int l = 0; l < n_lanes; ++l)
        if (lane[l]) lane[l]->wait = OP_NONE;
}

} // namespace

namespace {

void init_block(Block &b) {
    for (Thread &t : b.th) {
        void *st = nullptr;
        if (posix_memalign(&st, 64, kStackBytes) != 0) std::abort();
        t.stack = static_cast<char *>(st);
        uintptr_t top = (reinterpret_cast<uintptr_t>(t.stack) + kStackBytes) & ~(uintptr_t)15;
        void **ret_slot = reinterpret_cast<void **>(top - 16); // 16-byte aligned: entry sees rsp % 16 == 8
        ret_slot[1] = nullptr;
        ret_slot[0] = reinterpret_cast<void *>(&fiber_main);
        void **sp = ret_slot - 6; // r15 r14 r13 r12 rbx rbp
        for (int i = 0; i < 6; ++i) sp[i] = nullptr;
        t.sp = sp;
        t.done = false;
        t.wait = OP_NONE;
    }
}

size_t live_threads(const Block &b) {
    size_t n = 0;
    for (const Thread &t : b.th) n += t.done ? 0 : 1;
    return n;
}

// One scheduling round of a block: run every runnable thread up to its next collective, then
// resolve the block barrier or the warp collectives that are complete.  Returns whether anything
// moved.  Threads waiting at a grid barrier are left to run_grid.
bool step_block(Block &b) {
    g_blk = &b;
    const size_t n = b.th.size();
    bool progressed = false;
    for (Thread &t : b.th) {
        if (!t.done && t.wait == OP_YIELD) t.wait = OP_NONE; // a polling thread gets another turn
        if (t.done || t.wait != OP_NONE) continue;
        g_cur = &t;
        emu_switch(&g_sched_sp, t.sp);
        progressed = true;
    }
    if (!live_threads(b)) return progressed;
    // block barrier: every live thread has arrived
    bool all_bar = true;
    for (const Thread &t : b.th)
        if (!t.done && t.wait != OP_SYNCTHREADS) all_bar = false;
    if (all_bar) {
        ++g_ops[OP_SYNCTHREADS];
        for (Thread &t : b.th)
            if (!t.done) t.wait = OP_NONE;
        return true;
    }
    // warp collectives: every live lane of the warp waits at the same kind of operation
    for (size_t w0 = 0; w0 < n; w0 += 32) {
        Thread *lane[32] = {nullptr};
        const int nl = (int)std::min<size_t>(32, n - w0);
        Op op = OP_NONE;
        bool uniform = true, any = false;
        for (int l = 0; l < nl; ++l) {
            Thread &t = b.th[w0 + (size_t)l];
            if (t.done) continue;
            lane[l] = &t;
            if (!any) op = t.wait, any = true;
            else if (t.wait != op) uniform = false;
        }
        if (!any || op == OP_NONE || op == OP_SYNCTHREADS || op == OP_GRIDSYNC || op == OP_YIELD) continue;
        if (!uniform) {
            std::fprintf(stderr, "simt_emu: block %u warp %zu diverged across different collectives\n",
                         b.bid.x, w0 / 32);
            std::abort();
        }
        resolve_warp(lane, nl, op);
        progressed = true;
    }
    return progressed;
}

void fini_block(Block &b) {
    for (Thread &t : b.th) std::free(t.stack);
}

} // namespace

void run_block(Block &b) {
    init_block(b);
    while (live_threads(b)) {
        if (!step_block(b)) {
            std::fprintf(stderr, "simt_emu: deadlock in block %u (a barrier some threads never reach)\n", b.bid.x);
            std::abort();
        }
    }
    fini_block(b);
    g_blk = nullptr;
    g_cur = nullptr;
}

void run_grid(std::vector<Block> &blocks) {
    for (Block &b : blocks) init_block(b);
    for (;;) {
        size_t live = 0;
        bool progressed = false;
        for (Block &b : blocks) {
            if (!live_threads(b)) continue;
            progressed = step_block(b) || progressed;
            live += live_threads(b);
        }
        if (!live) break;
        if (progressed) continue;
        // nothing can move inside any block: the grid barrier, if every live thread is at it
        bool all_grid = true;
        for (const Block &b : blocks)
            for (const Thread &t : b.th)
                if (!t.done && t.wait != OP_GRIDSYNC) all_grid = false;
        if (!all_grid) {
            std::fprintf(stderr, "simt_emu: deadlock in a cooperative grid (a barrier some threads never reach)\n");
            std::abort();
        }
        for (Block &b : blocks)
            for (Thread &t : b.th)
                if (!t.done) t.wait = OP_NONE;
    }
    for (Block &b : blocks) fini_block(b);
    g_blk = nullptr;
    g_cur = nullptr;
}

} // namespace emu
