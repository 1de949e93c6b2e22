// cuda_emu.cpp -- coroutine scheduler behind cuda_emu.h (development tool, see that header).
#include "cuda_emu.h"
#include <chrono>
#include <omp.h>

thread_local uint3_emu threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local unsigned char *emu_smem = nullptr;

namespace {
constexpr size_t kStack = 64 * 1024;
struct Fiber {
    ucontext_t ctx;
    unsigned char *stack = nullptr;
    bool done = false;
    uint3_emu tid;
};
struct Worker {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int current = -1;
    const std::function<void()> *body = nullptr;
    std::vector<unsigned char> smem;
    int nthreads = 0;
    unsigned bar_cnt[64] = {0}, bar_gen[64] = {0};      // 0..15: bar.sync ids, 16..: one per warp (__syncwarp)
};
thread_local Worker *tl_worker = nullptr;

void fiber_entry() {
    Worker *w = tl_worker;
    (*w->body)();
    w->fibers[w->current].done = true;
    swapcontext(&w->fibers[w->current].ctx, &w->sched);
}
}  // namespace

static void emu_yield() {
    Worker *w = tl_worker;
    swapcontext(&w->fibers[w->current].ctx, &w->sched);
    threadIdx = w->fibers[w->current].tid;
}
// a real counting barrier: groups of a CTA may pass different numbers of barriers (persistent kernels)
void emu_barrier(int id, int nthreads) {
    Worker *w = tl_worker;
    const unsigned gen = w->bar_gen[id];
    if (++w->bar_cnt[id] == (unsigned)nthreads) {
        w->bar_cnt[id] = 0;
        w->bar_gen[id]++;
        return;
    }
    while (w->bar_gen[id] == gen) emu_yield();
}
void emu_syncthreads() { emu_barrier(0, tl_worker->nthreads); }
void emu_syncwarp() {
    Worker *w = tl_worker;
    const int warp = w->current / 32, lanes = w->nthreads - warp * 32 < 32 ? w->nthreads - warp * 32 : 32;
    emu_barrier(16 + warp, lanes);
}
// mbarrier: pending arrivals + outstanding transaction bytes of the current phase; the phase completes when both reach zero and
// the arrival count is re-armed (PTX mbarrier semantics).  Word layout (the kernels only see an opaque 64-bit slot):
//   bits 0..23 tx bytes, 24..35 pending arrivals, 36..47 arrival count of a phase, 48..63 completed phases
namespace {
inline void mbar_check(unsigned long long *bar) {
    const unsigned long long tx = *bar & 0xffffffull, pend = (*bar >> 24) & 0xfffull, cnt = (*bar >> 36) & 0xfffull;
    if (tx == 0 && pend == 0) *bar = (cnt << 24) | (cnt << 36) | ((((*bar >> 48) + 1) & 0xffffull) << 48);
}
}  // namespace
void emu_mbar_init(unsigned long long *bar, int count) { *bar = ((unsigned long long)count << 24) | ((unsigned long long)count << 36); }
void emu_mbar_arrive(unsigned long long *bar) {
    *bar -= 1ull << 24;
    mbar_check(bar);
}
void emu_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {       // mbarrier.arrive.expect_tx: one arrival + bytes to come
    *bar += bytes;
    *bar -= 1ull << 24;
    mbar_check(bar);
}
void emu_mbar_complete_tx(unsigned long long *bar, unsigned bytes) {
    *bar -= bytes;
    mbar_check(bar);
}
void emu_mbar_wait(unsigned long long *bar, unsigned parity) {
    while ((((unsigned)(*bar >> 48)) & 1u) == parity) emu_yield();
}

double emu_now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const long nblocks = (long)grid.x * grid.y * grid.z;
    const int nthreads = (int)(block.x * block.y * block.z);
#pragma omp parallel
    {
        Worker w;
        tl_worker = &w;
        w.body = &body;
        w.nthreads = nthreads;
        w.fibers.resize(nthreads);
        for (auto &f : w.fibers) f.stack = (unsigned char *)malloc(kStack);
        w.smem.assign(smem + 64, 0);
        emu_smem = w.smem.data();
        blockDim = block;
        gridDim = grid;
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < nblocks; b++) {
            blockIdx.x = (unsigned)(b % grid.x);
            blockIdx.y = (unsigned)((b / grid.x) % grid.y);
            blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
            for (int t = 0; t < nthreads; t++) {
                Fiber &f = w.fibers[t];
                f.done = false;
                f.tid.x = t % block.x;
                f.tid.y = (t / block.x) % block.y;
                f.tid.z = t / (block.x * block.y);
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = f.stack;
                f.ctx.uc_stack.ss_size = kStack;
                f.ctx.uc_link = nullptr;
                makecontext(&f.ctx, (void (*)())fiber_entry, 0);
            }
            memset(w.bar_cnt, 0, sizeof w.bar_cnt);
            int remaining = nthreads;
            while (remaining > 0) {
                for (int t = 0; t < nthreads; t++) {
                    Fiber &f = w.fibers[t];
                    if (f.done) continue;
                    w.current = t;
                    threadIdx = f.tid;
                    swapcontext(&w.sched, &f.ctx);
                    if (f.done) remaining--;
                }
            }
        }
        for (auto &f : w.fibers) free(f.stack);
        tl_worker = nullptr;
        emu_smem = nullptr;
    }
}
