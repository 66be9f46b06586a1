// cuda_emu.cpp — fiber scheduler behind tests/emu/cuda_emu.h (TEST-ONLY, see that header).
#include "cuda_emu.h"

#include <sys/mman.h>

namespace emu {
uint3_e g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char *dyn_smem = nullptr;

static const size_t STACK = 192 * 1024;
struct Fiber {
    ucontext_t ctx;
    void *stack = nullptr;
    bool done = true;
    uint3_e tid;
};
struct Warp {
    unsigned arrive = 0, depart = 0;
    unsigned long long vals[32];
    unsigned masks[32];
};
static std::vector<Fiber> fibers;
static std::vector<Warp> warps;
static ucontext_t sched_ctx;
static int cur = -1, nthreads = 0, live = 0;
static const std::function<void()> *body_ptr = nullptr;
static int bar_arrived = 0, bar_gen = 0, bar_count = 0, bar_result = 0;

static void trampoline() {
    (*body_ptr)();
    fibers[cur].done = true;
    live--;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}

void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
unsigned lane() { return (unsigned)cur & 31u; }

void syncthreads() { (void)syncthreads_count(0); }
int syncthreads_count(int pred) {
    int gen = bar_gen;
    bar_arrived++;
    if (pred) bar_count++;
    if (bar_arrived == live) {  // exited threads do not take part (matches hardware closely enough)
        bar_result = bar_count;
        bar_arrived = 0;
        bar_count = 0;
        bar_gen++;
        return bar_result;
    }
    while (bar_gen == gen) {
        if (bar_arrived >= live) {  // the threads still missing have exited meanwhile
            bar_result = bar_count;
            bar_arrived = 0;
            bar_count = 0;
            bar_gen++;
            break;
        }
        yield();
    }
    return bar_result;
}

unsigned long long collective(unsigned mask, unsigned long long v, unsigned long long *all) {
    Warp &w = warps[cur >> 5];
    unsigned me = 1u << (cur & 31);
    // lanes of the mask that do not exist (partial last warp) or have exited cannot arrive
    unsigned present = 0;
    for (int i = 0; i < 32; i++) {
        int t = (cur & ~31) + i;
        if (t < nthreads && !fibers[t].done) present |= 1u << i;
    }
    if (!(mask & me)) {
        fprintf(stderr, "emu: lane %d not in its own collective mask %08x\n", cur & 31, mask);
        abort();
    }
    mask &= present;
    while (w.arrive & me) yield();  // previous collective of this lane not fully retired
    w.vals[cur & 31] = v;
    w.masks[cur & 31] = mask;
    w.arrive |= me;
    int spins = 0;
    while ((w.arrive & mask) != mask) {
        yield();
        if (++spins > 1000000) {
            fprintf(stderr, "emu: warp collective deadlock (block %u thread %d mask %08x arrived %08x)\n", g_blockIdx.x, cur, mask, w.arrive);
            abort();
        }
    }
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1) {
            if (w.masks[i] != mask) {
                fprintf(stderr, "emu: mismatched collective masks in warp (%08x vs %08x)\n", w.masks[i], mask);
                abort();
            }
            all[i] = w.vals[i];
        } else
            all[i] = 0;
    w.depart |= me;
    if ((w.depart & mask) == mask) {
        w.arrive &= ~mask;
        w.depart &= ~mask;
    }
    return v;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    nthreads = (int)(block.x * block.y * block.z);
    if ((int)fibers.size() < nthreads) {
        size_t old = fibers.size();
        fibers.resize(nthreads);
        for (size_t i = old; i < fibers.size(); i++) {
            fibers[i].stack = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (fibers[i].stack == MAP_FAILED) { perror("mmap"); abort(); }
        }
    }
    warps.assign((nthreads + 31) / 32, Warp());
    unsigned char *smem_buf = (unsigned char *)malloc(smem ? smem : 1);
    dyn_smem = smem_buf;
    g_blockDim = block;
    g_gridDim = grid;
    body_ptr = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = uint3_e{bx, by, bz};
                for (auto &w : warps) w = Warp();
                bar_arrived = bar_count = 0;
                live = nthreads;
                for (int t = 0; t < nthreads; t++) {
                    Fiber &f = fibers[t];
                    f.done = false;
                    f.tid = uint3_e{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, trampoline, 0);
                }
                long idle_rounds = 0;
                while (live > 0) {
                    int before = live;
                    for (int t = 0; t < nthreads; t++) {
                        if (fibers[t].done) continue;
                        cur = t;
                        g_threadIdx = fibers[t].tid;
                        swapcontext(&sched_ctx, &fibers[t].ctx);
                    }
                    if (live == before) {
                        if (++idle_rounds > 50000000) { fprintf(stderr, "emu: CTA made no progress (deadlock?)\n"); abort(); }
                    } else
                        idle_rounds = 0;
                }
            }
    dyn_smem = nullptr;
    free(smem_buf);
    cur = -1;
}
}  // namespace emu
