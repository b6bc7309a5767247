// SIMT emulator runtime: TEST INFRASTRUCTURE ONLY.  See simt_emu.h.
#include "simt_emu.h"

#include <sys/mman.h>
#include <chrono>
#include <vector>

extern "C" void zg_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl zg_ctx_switch
.type zg_ctx_switch,@function
zg_ctx_switch:
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
.size zg_ctx_switch,.-zg_ctx_switch
)");

double zg_emu_now_ms() {
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

namespace zg_emu {

State g;

namespace {
enum { READY, WAIT_WARP, WAIT_CTA, DONE };
constexpr size_t kStack = 512 * 1024;

struct Fiber {
	void* sp;
	char* stack;
	int state;
	unsigned wait_mask;
	uint3 tid;
};
struct Warp {
	unsigned arrived;
	unsigned alive;
	unsigned long long slots[32];
};

std::vector<Fiber> fibers;
std::vector<char*> stack_pool;
std::vector<Warp> warps;
void* sched_sp;
int cur = -1;
int n_threads = 0, n_alive = 0, cta_arrived = 0;
const std::function<void()>* body;
int order_mode = -1;  // 0 fwd, 1 rev, 2 rand
unsigned rng = 12345;

void yield_to_sched() { zg_ctx_switch(&fibers[cur].sp, sched_sp); }

void release_cta_if_complete() {
	if (n_alive > 0 && cta_arrived == n_alive) {
		cta_arrived = 0;
		for (auto& f : fibers)
			if (f.state == WAIT_CTA) f.state = READY;
	}
}

void release_warp_if_complete(Warp& w, int wbase, unsigned mask) {
	unsigned need = mask & w.alive;
	if ((w.arrived & need) == need) {
		w.arrived &= ~mask;
		for (int l = 0; l < 32; l++)
			if ((need >> l) & 1) {
				Fiber& f = fibers[wbase + l];
				if (f.state == WAIT_WARP && f.wait_mask == mask) f.state = READY;
			}
	}
}

void fiber_entry() {
	(*body)();
	Fiber& f = fibers[cur];
	f.state = DONE;
	n_alive--;
	Warp& w = warps[cur / 32];
	w.alive &= ~(1u << (cur % 32));
	release_cta_if_complete();
	// a lane that exits can complete a pending warp rendezvous of the remaining lanes
	int wbase = (cur / 32) * 32;
	for (int l = 0; l < 32 && wbase + l < n_threads; l++) {
		Fiber& o = fibers[wbase + l];
		if (o.state == WAIT_WARP) {
			release_warp_if_complete(w, wbase, o.wait_mask);
		}
	}
	yield_to_sched();
	fprintf(stderr, "simt_emu: resumed a finished fiber\n");
	abort();
}

char* get_stack(size_t i) {
	while (stack_pool.size() <= i) {
		void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (p == MAP_FAILED) {
			perror("simt_emu mmap");
			abort();
		}
		stack_pool.push_back((char*)p);
	}
	return stack_pool[i];
}

int pick_next() {
	if (order_mode == 2) {
		// random among ready
		int cnt = 0;
		for (auto& f : fibers) cnt += f.state == READY;
		if (!cnt) return -1;
		rng = rng * 1664525u + 1013904223u;
		int k = (int)((rng >> 8) % (unsigned)cnt);
		for (int i = 0; i < n_threads; i++)
			if (fibers[i].state == READY && k-- == 0) return i;
		return -1;
	}
	int start = cur < 0 ? 0 : cur;
	for (int s = 1; s <= n_threads; s++) {
		int i = order_mode == 1 ? ((start - s) % n_threads + n_threads) % n_threads : (start + s) % n_threads;
		if (fibers[i].state == READY) return i;
	}
	return -1;
}

void run_cta() {
	n_alive = n_threads;
	cta_arrived = 0;
	for (int i = 0; i < n_threads; i++) {
		Fiber& f = fibers[i];
		f.state = READY;
		f.stack = get_stack(i);
		uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
		void** sp = (void**)top;
		*--sp = nullptr;                // fake return address of fiber_entry
		*--sp = (void*)&fiber_entry;    // `ret` target of the first switch
		for (int r = 0; r < 6; r++) *--sp = nullptr;
		f.sp = sp;
	}
	size_t nw = (n_threads + 31) / 32;
	warps.assign(nw, Warp{});
	for (size_t w = 0; w < nw; w++) {
		int lanes = std::min(32, n_threads - (int)w * 32);
		warps[w].alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
	}
	cur = -1;
	while (n_alive > 0) {
		int nx = pick_next();
		if (nx < 0) {
			fprintf(stderr, "simt_emu: DEADLOCK in CTA (%u,%u,%u): %d threads alive, none runnable\n", g.bIdx.x,
				g.bIdx.y, g.bIdx.z, n_alive);
			for (int i = 0; i < n_threads; i++)
				if (fibers[i].state != DONE)
					fprintf(stderr, "  tid %d state %d mask %08x\n", i, fibers[i].state, fibers[i].wait_mask);
			abort();
		}
		cur = nx;
		g.tIdx = fibers[cur].tid;
		zg_ctx_switch(&sched_sp, fibers[cur].sp);
	}
}
}  // namespace

unsigned lane_id() { return (unsigned)cur & 31u; }
unsigned long long* warp_slots() { return warps[cur / 32].slots; }

void warp_barrier(unsigned mask) {
	Warp& w = warps[cur / 32];
	unsigned lane = cur & 31;
	if (!((mask >> lane) & 1)) {
		fprintf(stderr, "simt_emu: lane %u called a warp collective with mask %08x that excludes it\n", lane, mask);
		abort();
	}
	w.arrived |= 1u << lane;
	unsigned need = mask & w.alive;
	if ((w.arrived & need) == need) {
		int me = cur;
		fibers[me].state = WAIT_WARP;
		fibers[me].wait_mask = mask;
		release_warp_if_complete(w, (cur / 32) * 32, mask);
		fibers[me].state = READY;
		return;
	}
	fibers[cur].state = WAIT_WARP;
	fibers[cur].wait_mask = mask;
	yield_to_sched();
	g.tIdx = fibers[cur].tid;
}

void yield() {
	fibers[cur].state = READY;
	yield_to_sched();
	g.tIdx = fibers[cur].tid;
}

void cta_barrier() {
	cta_arrived++;
	if (cta_arrived == n_alive) {
		cta_arrived = 0;
		for (auto& f : fibers)
			if (f.state == WAIT_CTA) f.state = READY;
		return;
	}
	fibers[cur].state = WAIT_CTA;
	yield_to_sched();
	g.tIdx = fibers[cur].tid;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
	if (order_mode < 0) {
		const char* e = getenv("ZG_EMU_ORDER");
		order_mode = !e ? 0 : !strcmp(e, "rev") ? 1 : !strcmp(e, "rand") ? 2 : 0;
	}
	if (cur >= 0 && n_alive > 0) {
		fprintf(stderr, "simt_emu: nested launch\n");
		abort();
	}
	n_threads = (int)(block.x * block.y * block.z);
	fibers.assign(n_threads, Fiber{});
	for (int i = 0; i < n_threads; i++) {
		fibers[i].tid.x = i % block.x;
		fibers[i].tid.y = (i / block.x) % block.y;
		fibers[i].tid.z = i / (block.x * block.y);
	}
	std::vector<unsigned char> dyn(smem + 16);
	g.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
	g.bDim = block;
	g.gDim = grid;
	body = &fn;
	for (unsigned z = 0; z < grid.z; z++)
		for (unsigned y = 0; y < grid.y; y++)
			for (unsigned x = 0; x < grid.x; x++) {
				g.bIdx = uint3{x, y, z};
				run_cta();
			}
	cur = -1;
	n_alive = 0;
}

}  // namespace zg_emu
