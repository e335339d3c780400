// emu_runtime.cpp -- fiber scheduler behind emu_cuda.h (TEST INFRASTRUCTURE ONLY).
#include <sys/mman.h>

#include "emu_cuda.h"

namespace emu {

Cta g_cta;
Fiber* g_cur = nullptr;
ucontext_t g_sched;
uint3 g_blockIdx{0, 0, 0};
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
std::function<void()>* g_body = nullptr;
unsigned long long g_switches = 0;

static const size_t kStack = 256 * 1024;
static std::vector<char*> g_stacks;
static std::vector<unsigned char> g_smem_store;

void yield() {
    Fiber* me = g_cur;
    swapcontext(&me->ctx, &g_sched);
}

static void trampoline() {
    (*g_body)();
    g_cur->done = true;
    g_cta.live--;
    // a finished thread no longer takes part in __syncthreads: release waiters if
    // it was the last one they were waiting for
    if (g_cta.live > 0 && g_cta.bar_count == g_cta.live) {
        g_cta.bar_count = 0;
        g_cta.bar_gen++;
    }
    swapcontext(&g_cur->ctx, &g_sched);
}

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    int nthreads = (int)(block.x * block.y * block.z);
    while ((int)g_stacks.size() < nthreads) {
        void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) {
            fprintf(stderr, "emu: stack mmap failed\n");
            abort();
        }
        g_stacks.push_back((char*)p);
    }
    if (g_smem_store.size() < smem + 64) g_smem_store.resize(smem + 64);
    g_dyn_smem = (unsigned char*)(((uintptr_t)g_smem_store.data() + 63) & ~(uintptr_t)63);
    g_blockDim = block;
    g_gridDim = grid;
    g_body = &body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3{bx, by, bz};
                Cta& c = g_cta;
                c.fibers.assign(nthreads, Fiber());
                c.warps.assign((nthreads + 31) / 32, WarpState());
                c.bar_count = 0;
                c.bar_gen = 0;
                c.live = nthreads;
                c.nthreads = nthreads;
                int lin = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++lin) {
                            Fiber& f = c.fibers[lin];
                            f.tid = uint3{tx, ty, tz};
                            f.lin = lin;
                            f.warp = lin / 32;
                            f.lane = lin % 32;
                            f.stack = g_stacks[lin];
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack;
                            f.ctx.uc_stack.ss_size = kStack;
                            f.ctx.uc_link = &g_sched;
                            makecontext(&f.ctx, (void (*)())trampoline, 0);
                        }
                unsigned long long guard = 0;
                while (c.live > 0) {
                    for (int t = 0; t < nthreads; ++t) {
                        Fiber& f = c.fibers[t];
                        if (f.done) continue;
                        g_cur = &f;
                        g_switches++;
                        swapcontext(&g_sched, &f.ctx);
                    }
                    if (++guard > 400000000ull / (unsigned)std::max(nthreads, 1)) {
                        fprintf(stderr, "emu: kernel made no progress (deadlock?) in block (%u,%u,%u)\n",
                                bx, by, bz);
                        abort();
                    }
                }
            }
    g_body = nullptr;
    g_cur = nullptr;
}

}  // namespace emu
