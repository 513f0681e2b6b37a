// simt_emu.cpp - fiber scheduler behind simt_emu.h (TEST-ONLY, see that header).
#include "simt_emu.h"

namespace simt {

thread_local Block* B = nullptr;
static const size_t kStack = 256 * 1024;
static thread_local std::vector<char*> stack_pool;

static void fiber_entry() {
    Block* b = B;
    (*b->body)();
    Fiber& f = b->fibers[b->cur];
    f.done = true;
    b->live--;
    Warp& w = b->warps[f.linear >> 5];
    w.live--;
    // a thread that exits counts as arrived for barriers other threads are waiting on
    if (b->live > 0 && b->bar_count >= b->live) { b->bar_count = 0; b->bar_gen++; }
    if (w.live > 0 && w.arrived >= w.live) { w.arrived = 0; w.gen++; }
    swapcontext(&f.ctx, &b->sched);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    std::vector<unsigned char> dyn(smem + 64);
    while (stack_pool.size() < nthreads) stack_pool.push_back((char*)std::malloc(kStack));
    Block blk;
    blk.body = &body;
    blk.bdim = block;
    blk.gdim = grid;
    blk.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    blk.fibers.resize(nthreads);
    Block* saved = B;
    B = &blk;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        blk.bidx = dim3(bx, by, bz);
        blk.live = (int)nthreads;
        blk.bar_count = 0;
        blk.warps.assign((nthreads + 31) / 32, Warp());
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = blk.fibers[t];
            f.done = false;
            f.linear = t;
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            f.stack = stack_pool[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &blk.sched;
            makecontext(&f.ctx, fiber_entry, 0);
            blk.warps[t >> 5].live++;
        }
        while (blk.live > 0) {
            int before = blk.live;
            unsigned progress_probe = blk.bar_gen;
            for (unsigned t = 0; t < nthreads; ++t) {
                if (blk.fibers[t].done) continue;
                blk.cur = (int)t;
                swapcontext(&blk.sched, &blk.fibers[t].ctx);
            }
            (void)before; (void)progress_probe;
        }
    }
    B = saved;
}

}  // namespace simt
