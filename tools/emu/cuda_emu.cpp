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
};
thread_local Worker *tl_worker = nullptr;

void fiber_entry() {
    Worker *w = tl_worker;
    (*w->body)();
    w->fibers[w->current].done = true;
    swapcontext(&w->fibers[w->current].ctx, &w->sched);
}
}  // namespace

void emu_syncthreads() {
    Worker *w = tl_worker;
    swapcontext(&w->fibers[w->current].ctx, &w->sched);
    threadIdx = w->fibers[w->current].tid;
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
