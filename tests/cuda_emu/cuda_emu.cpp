/*
 * tests/cuda_emu/cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
 * Cooperative fiber scheduler: one block at a time, one ucontext fiber per
 * CUDA thread, round-robin; barriers and warp collectives are yield loops.
 */
#include "cuda_emu.h"

namespace cuemu {

Block* g_block = nullptr;
static const size_t kStack = 256 * 1024;
static long g_deadlock_spins = getenv("CUEMU_DEADLOCK_SPINS") ? atol(getenv("CUEMU_DEADLOCK_SPINS")) : 200000000L;

static void fiber_entry() {
	Block* b = g_block;
	b->body();
	b->fibers[b->current].done = true;
	swapcontext(&b->fibers[b->current].ctx, &b->sched);
}

void yield_now() {
	Block* b = g_block;
	swapcontext(&b->fibers[b->current].ctx, &b->sched);
}

int warp_width() {
	Block* b = g_block;
	int n = (int)b->fibers.size(), w = warp_id();
	return std::min(32, n - w * 32);
}

void warp_barrier() {
	Block* b = g_block;
	WarpSync& w = b->warps[warp_id()];
	int width = warp_width();
	int phase = w.phase;
	if (++w.arrived == width) { w.arrived = 0; w.phase ^= 1; return; }
	while (w.phase == phase) yield_now();
}

void block_barrier() {
	Block* b = g_block;
	int phase = b->bar_phase;
	if (++b->bar_arrived == (int)b->fibers.size()) { b->bar_arrived = 0; b->bar_phase ^= 1; return; }
	while (b->bar_phase == phase) yield_now();
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
	const int nthreads = (int)(block.x * block.y * block.z);
	Block blk;
	blk.bdim = block;
	blk.gdim = grid;
	blk.body = body;
	blk.fibers.resize(nthreads);
	blk.warps.resize((nthreads + 31) / 32);
	blk.dyn_smem = (char*)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
	for (int t = 0; t < nthreads; ++t) blk.fibers[t].stack = (char*)malloc(kStack);
	Block* saved = g_block;
	g_block = &blk;
	for (unsigned bz = 0; bz < grid.z; ++bz)
	for (unsigned by = 0; by < grid.y; ++by)
	for (unsigned bx = 0; bx < grid.x; ++bx) {
		blk.bid = dim3(bx, by, bz);
		blk.bar_arrived = blk.bar_phase = 0;
		for (auto& w : blk.warps) { w.arrived = 0; w.phase = 0; }
		for (int t = 0; t < nthreads; ++t) {
			Fiber& f = blk.fibers[t];
			f.done = false;
			f.linear = t;
			f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
			getcontext(&f.ctx);
			f.ctx.uc_stack.ss_sp = f.stack;
			f.ctx.uc_stack.ss_size = kStack;
			f.ctx.uc_link = &blk.sched;
			makecontext(&f.ctx, fiber_entry, 0);
		}
		int live = nthreads;
		long spins = 0;
		while (live > 0) {
			int progressed = 0;
			for (int t = 0; t < nthreads; ++t) {
				Fiber& f = blk.fibers[t];
				if (f.done) continue;
				blk.current = t;
				swapcontext(&blk.sched, &f.ctx);
				if (f.done) { --live; ++progressed; }
			}
			if (!progressed && ++spins > g_deadlock_spins) {
				fprintf(stderr, "cuda_emu: block (%u,%u,%u) appears deadlocked; live fibers:", bx, by, bz);
				for (int t = 0; t < nthreads; ++t) if (!blk.fibers[t].done) fprintf(stderr, " %d", t);
				fprintf(stderr, "\n");
				abort();
			}
		}
	}
	g_block = saved;
	for (int t = 0; t < nthreads; ++t) free(blk.fibers[t].stack);
	free(blk.dyn_smem);
}

}  // namespace cuemu
