// tests/support/hostemu/hostemu.cpp — TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
//
// The execution model behind the emulated build of the product library:
//   * lh_emu_launch runs the blocks of a grid one after another on the calling host thread; the threads of a block are
//     fibers (ucontext) scheduled cooperatively.  A fiber runs until it reaches a rendezvous (__syncthreads, __syncwarp, a
//     shuffle, a spin-wait's __nanosleep) and then hands over; the ORDER in which fibers are resumed is a policy (ascending,
//     descending, seeded random) — a kernel whose barriers are sufficient computes the same bits under every policy, one with
//     a missing barrier does not (tests/test_hostemu.py runs the stage kernels under all three).
//   * cp.async: every fiber has a queue of commit groups.  "eager" completes a copy when it is issued, "lazy" only when a
//     wait_group forces it — the two extremes of what the hardware may do; a kernel that overwrites a source before the copy
//     has been waited for, or reads the destination before its group is complete, differs between the two.
//   * "device" memory: every allocation ends at a PROT_NONE guard page (an out-of-bounds access past the end faults at the
//     offending instruction) and is filled with a signalling pattern (NaN for doubles), so nothing can depend on fresh device
//     memory being zero.
//   * streams and events: by default everything completes inside the call that enqueues it; lh_emu_set_async turns streams into
//     deferred queues (see "streams and events" below).
//   * controls for tests: allocation-failure injection, live allocation / handle counts, device and SM count.
#include "cuda_runtime.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

// Fiber switch.  x86-64: six callee-saved registers and the stack pointer, ~10 ns (swapcontext also saves the signal mask with
// a system call per switch, and a stage kernel switches once per cell and thread).  Elsewhere: ucontext.
#if defined(__x86_64__)
#define LH_EMU_ASM_SWITCH 1
extern "C" void lhemu_switch(void** save_sp, void* load_sp);
asm(R"ASM(
    .text
    .globl lhemu_switch
    .type lhemu_switch, @function
lhemu_switch:
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
    .size lhemu_switch, .-lhemu_switch
)ASM");
struct LhEmuCtx { void* sp; };
#else
#define LH_EMU_ASM_SWITCH 0
#include <ucontext.h>
struct LhEmuCtx { ucontext_t uc; };
#endif

#include <atomic>
#include <memory>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

std::atomic<int> g_policy{0};
std::atomic<uint64_t> g_seed{1};
std::atomic<int> g_cp_lazy{0};
std::atomic<int> g_devices{1};
std::atomic<int> g_sms{148};
std::atomic<uint64_t> g_launches{0};
std::atomic<int64_t> g_live_handles{0};
std::atomic<int64_t> g_fail_alloc_in{-1};         // >= 0: that many more allocations succeed, the next one fails (then -1 again)           // streams + events created and not yet destroyed

constexpr size_t STACK_BYTES = 512 * 1024;

struct CpCopy { uint32_t dst; const void* src; };

struct Block;
struct Fiber {
    LhEmuThread t;
    LhEmuCtx ctx;
    char* stack = nullptr;
    bool done = false;
    Block* block = nullptr;
    std::vector<std::vector<CpCopy>> groups;      // committed cp.async groups, oldest first
    std::vector<CpCopy> open;                     // copies issued since the last commit
};

struct Warp { int arrived = 0, live = 0; uint64_t gen = 0; uint64_t xchg[32]; };

struct Block {
    std::vector<Fiber*> fibers;
    std::vector<Warp> warps;
    int arrived = 0, live = 0;
    uint64_t gen = 0;
    int or_acc = 0, or_result = 0;
    const std::function<void()>* body = nullptr;
    LhEmuCtx sched;
    int cursor = 0;
    uint64_t rng = 1;
};

thread_local Fiber* tl_self = nullptr;
thread_local std::vector<char*> tl_stacks;        // reused across launches of this host thread
thread_local int tl_device = 0;
thread_local LhEmuThread tl_host_thread;          // what lh_emu_self() returns outside a kernel (never dereferenced by host code)

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "[hostemu] fatal: %s\n", what);
    fflush(stderr);
    abort();
}

char* get_stack(size_t k)
{
    while (tl_stacks.size() <= k) {
        void* p = mmap(nullptr, STACK_BYTES + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) die("mmap of a fiber stack failed");
        mprotect(p, 4096, PROT_NONE);             // stack overflow faults instead of corrupting a neighbour
        tl_stacks.push_back((char*)p + 4096);
    }
    return tl_stacks[k];
}

void complete(Fiber* f, const CpCopy& c)
{
    if ((size_t)c.dst + 16 > f->t.smem_bytes) die("cp.async: destination outside the block's dynamic shared memory");
    memcpy(f->t.smem + c.dst, c.src, 16);
}

void fiber_exit(Fiber* f)
{
    // leaving with copies in flight is legal (they complete); honour them so that lazy mode is not a leak of work
    for (auto& g : f->groups) for (auto& c : g) complete(f, c);
    for (auto& c : f->open) complete(f, c);
    f->groups.clear();
    f->open.clear();
    Block* b = f->block;
    f->done = true;
    // an exited thread no longer takes part in barriers
    --b->live;
    if (b->live > 0 && b->arrived == b->live) { b->arrived = 0; b->or_result = b->or_acc; b->or_acc = 0; ++b->gen; }
    Warp& w = b->warps[f->t.warp];
    --w.live;
    if (w.live > 0 && w.arrived == w.live) { w.arrived = 0; ++w.gen; }
}

void switch_to(LhEmuCtx* from, LhEmuCtx* to)
{
#if LH_EMU_ASM_SWITCH
    lhemu_switch(&from->sp, to->sp);
#else
    if (swapcontext(&from->uc, &to->uc) != 0) die("swapcontext failed");
#endif
}

void fiber_main()
{
    Fiber* f = tl_self;
    (*f->block->body)();
    fiber_exit(f);
#if LH_EMU_ASM_SWITCH
    switch_to(&f->ctx, &f->block->sched);         // never resumed
    die("a finished fiber was resumed");
#endif
    // ucontext: returns to uc_link (the scheduler)
}

void fiber_init(Fiber* f, Block* b)
{
#if LH_EMU_ASM_SWITCH
    // what lhemu_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into fiber_main with the ABI's stack alignment
    uintptr_t top = ((uintptr_t)f->stack + STACK_BYTES) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                              // where a return address would be (fiber_main never returns)
    *--sp = (void*)fiber_main;
    for (int k = 0; k < 6; ++k) *--sp = nullptr;
    f->ctx.sp = sp;
    (void)b;
#else
    if (getcontext(&f->ctx.uc) != 0) die("getcontext failed");
    f->ctx.uc.uc_stack.ss_sp = f->stack;
    f->ctx.uc.uc_stack.ss_size = STACK_BYTES;
    f->ctx.uc.uc_link = &b->sched.uc;
    makecontext(&f->ctx.uc, fiber_main, 0);
#endif
}

// pick the next fiber to run (not done); called on the scheduler context
Fiber* pick(Block* b)
{
    const int n = (int)b->fibers.size();
    const int policy = g_policy.load(std::memory_order_relaxed);
    if (policy == 2) {
        for (int tries = 0; tries < 4 * n + 16; ++tries) {
            b->rng = b->rng * 6364136223846793005ull + 1442695040888963407ull;
            Fiber* f = b->fibers[(b->rng >> 33) % (uint64_t)n];
            if (!f->done) return f;
        }
    }
    for (int k = 0; k < n; ++k) {
        b->cursor = policy == 1 ? (b->cursor + n - 1) % n : (b->cursor + 1) % n;
        Fiber* f = b->fibers[b->cursor];
        if (!f->done) return f;
    }
    return nullptr;
}

void run_block(Block* b)
{
    b->cursor = g_policy.load() == 1 ? 0 : (int)b->fibers.size() - 1;
    for (;;) {
        Fiber* f = pick(b);
        if (!f) break;
        tl_self = f;
        switch_to(&b->sched, &f->ctx);
        tl_self = nullptr;
    }
}

inline void yield_now()
{
    Fiber* f = tl_self;
    if (!f) die("a device-side rendezvous was called outside a kernel");
    switch_to(&f->ctx, &f->block->sched);
}

// ------------------------------------------------------------------ "device" memory
struct Allocation { void* base; size_t mapped; size_t bytes; bool host; };
std::mutex g_alloc_mutex;
std::unordered_map<void*, Allocation> g_allocs;

cudaError_t guarded_alloc(void** out, size_t bytes, bool host)
{
    if (!out) return cudaErrorInvalidValue;
    *out = nullptr;
    if (g_fail_alloc_in.load() >= 0 && g_fail_alloc_in.fetch_sub(1) == 0) return cudaErrorMemoryAllocation;   // injected failure
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t need = (bytes + 15) / 16 * 16;
    const size_t body = (std::max<size_t>(need, 16) + page - 1) / page * page;
    char* base = (char*)mmap(nullptr, body + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (base == (char*)MAP_FAILED) return cudaErrorMemoryAllocation;
    if (mprotect(base + body, page, PROT_NONE) != 0) { munmap(base, body + page); return cudaErrorMemoryAllocation; }
    char* p = base + body - need;                 // the allocation ENDS at the guard page
    memset(p, 0xff, need);                        // 0xff..ff is a NaN as a double, -1 as an integer: never a plausible zero
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    g_allocs[p] = Allocation{base, body + page, bytes, host};
    *out = p;
    return cudaSuccess;
}

cudaError_t guarded_free(void* p, bool host)
{
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end() || it->second.host != host) die(host ? "cudaFreeHost of a pointer cudaMallocHost did not return" : "cudaFree of a pointer cudaMalloc did not return");
    munmap(it->second.base, it->second.mapped);
    g_allocs.erase(it);
    return cudaSuccess;
}

}  // namespace

// ---------------------------------------------------------------------- streams and events
// Synchronous mode (default): every operation completes inside the call that enqueues it.
// Asynchronous modes (lh_emu_set_async): a stream is a FIFO of deferred operations, an event completes when its record is
// executed, cudaStreamWaitEvent blocks the waiting stream's queue.  Work is executed only when something forces it — a
// stream / event synchronisation, a blocking copy, cudaFree — and then (mode 1, "lazy") only what that call needs, in
// dependency order, or (mode 2, "random") seeded random runnable operations of ANY stream until the call's condition holds.
// Both are legal executions of the same CUDA program; a host layer whose stream / event dependencies are complete computes the
// same bits in all three modes, one that relies on "the other stream has surely finished by now" does not.
struct LhEmuOp {
    enum Kind { RUN, RECORD, WAIT } kind;
    std::function<void()> fn;
    LhEmuEvent* ev = nullptr;
    uint64_t seq = 0;
};
struct LhEmuStream { int device; std::deque<LhEmuOp> q; bool visiting = false; };
struct LhEmuEvent {
    double t_ms = 0.0;
    uint64_t recorded = 0, completed = 0;            // records enqueued / executed
    std::deque<std::pair<uint64_t, LhEmuStream*>> where;   // (sequence number, stream) of the records still queued
};

namespace {
std::atomic<int> g_async{0};
std::recursive_mutex g_api;                          // asynchronous modes: one API call at a time
std::vector<LhEmuStream*> g_streams;                 // never shrinks (streams and events are leaked in asynchronous modes)
uint64_t g_async_rng = 0x2545F4914F6CDD1Dull;

double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

bool head_runnable(LhEmuStream* s)
{
    if (s->q.empty()) return false;
    const LhEmuOp& op = s->q.front();
    return op.kind != LhEmuOp::WAIT || op.ev->completed >= op.seq;
}

void exec_head(LhEmuStream* s)
{
    LhEmuOp op = std::move(s->q.front());
    s->q.pop_front();
    if (op.kind == LhEmuOp::RUN) op.fn();
    else if (op.kind == LhEmuOp::RECORD) {
        op.ev->completed = std::max(op.ev->completed, op.seq);
        op.ev->t_ms = now_ms();
        while (!op.ev->where.empty() && op.ev->where.front().first <= op.seq) op.ev->where.pop_front();
    }
}

void event_complete_lazy(LhEmuEvent* e, uint64_t seq);

// lazy: execute the operations of s in order, pulling in other streams only where a wait needs them
void stream_step_lazy(LhEmuStream* s)
{
    if (s->visiting) die("stream dependency cycle: a stream waits for an event that can only complete after the wait itself");
    s->visiting = true;
    LhEmuOp& op = s->q.front();
    if (op.kind == LhEmuOp::WAIT && op.ev->completed < op.seq) event_complete_lazy(op.ev, op.seq);
    exec_head(s);
    s->visiting = false;
}

void event_complete_lazy(LhEmuEvent* e, uint64_t seq)
{
    while (e->completed < seq) {
        LhEmuStream* t = nullptr;
        for (auto& w : e->where) if (w.first >= seq) { t = w.second; break; }
        if (!t || t->q.empty()) die("deadlock: waiting for an event whose record is not queued anywhere");
        stream_step_lazy(t);
    }
}

template <class Pred> void progress_until(Pred done, LhEmuStream* want_stream, LhEmuEvent* want_event, uint64_t want_seq)
{
    while (!done()) {
        if (g_async.load() == 2) {
            std::vector<LhEmuStream*> runnable;
            for (auto* s : g_streams) if (head_runnable(s)) runnable.push_back(s);
            if (runnable.empty()) die("deadlock: nothing is runnable but a synchronisation is outstanding");
            g_async_rng = g_async_rng * 6364136223846793005ull + 1442695040888963407ull;
            exec_head(runnable[(g_async_rng >> 33) % runnable.size()]);
        } else if (want_stream) {
            stream_step_lazy(want_stream);
        } else {
            event_complete_lazy(want_event, want_seq);
        }
    }
}

void drain_stream(LhEmuStream* s) { progress_until([&] { return s->q.empty(); }, s, nullptr, 0); }
void drain_all()
{
    for (size_t k = 0; k < g_streams.size(); ++k) drain_stream(g_streams[k]);
}

// a little background progress in random mode, so that not everything happens at the synchronisation points
void background_progress()
{
    if (g_async.load() != 2) return;
    g_async_rng = g_async_rng * 6364136223846793005ull + 1442695040888963407ull;
    int n = (int)((g_async_rng >> 40) % 4);
    while (n-- > 0) {
        std::vector<LhEmuStream*> runnable;
        for (auto* s : g_streams) if (head_runnable(s)) runnable.push_back(s);
        if (runnable.empty()) return;
        g_async_rng = g_async_rng * 6364136223846793005ull + 1442695040888963407ull;
        exec_head(runnable[(g_async_rng >> 33) % runnable.size()]);
    }
}

bool is_pinned_host(const void* p, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    for (auto& kv : g_allocs)
        if (kv.second.host && (const char*)p >= (const char*)kv.first && (const char*)p + bytes <= (const char*)kv.first + kv.second.bytes) return true;
    return false;
}
bool is_device(const void* p)
{
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    for (auto& kv : g_allocs)
        if (!kv.second.host && (const char*)p >= (const char*)kv.first && (const char*)p < (const char*)kv.first + std::max<size_t>(kv.second.bytes, 1)) return true;
    return false;
}

// enqueue fn on stream s (asynchronous modes) or run it now
void submit(LhEmuStream* s, std::function<void()> fn)
{
    if (!g_async.load() || !s) { fn(); return; }
    LhEmuOp op;
    op.kind = LhEmuOp::RUN;
    op.fn = std::move(fn);
    s->q.push_back(std::move(op));
    background_progress();
}
}  // namespace

// ---------------------------------------------------------------------- device-side hooks
extern "C" {

LhEmuThread* lh_emu_self(void) { return tl_self ? &tl_self->t : &tl_host_thread; }

void lh_emu_fault(const char* what) { die(what); }

void lh_emu_yield(void) { yield_now(); }

void lh_emu_syncthreads(void)
{
    Fiber* f = tl_self;
    if (!f) die("__syncthreads outside a kernel");
    Block* b = f->block;
    const uint64_t gen = b->gen;
    if (++b->arrived == b->live) { b->arrived = 0; b->or_result = b->or_acc; b->or_acc = 0; ++b->gen; return; }
    while (b->gen == gen) yield_now();
}

int lh_emu_syncthreads_or(int pred)
{
    Fiber* f = tl_self;
    if (!f) die("__syncthreads_or outside a kernel");
    f->block->or_acc |= pred != 0;
    lh_emu_syncthreads();
    const int r = f->block->or_result;
    lh_emu_syncthreads();                          // nobody starts the next reduction before everybody has read this one
    return r;
}

void lh_emu_syncwarp(void)
{
    Fiber* f = tl_self;
    if (!f) die("__syncwarp outside a kernel");
    Warp& w = f->block->warps[f->t.warp];
    const uint64_t gen = w.gen;
    if (++w.arrived == w.live) { w.arrived = 0; ++w.gen; return; }
    while (w.gen == gen) yield_now();
}

uint64_t lh_emu_shfl(uint64_t bits, int src_lane)
{
    Fiber* f = tl_self;
    if (!f) die("__shfl_sync outside a kernel");
    Warp& w = f->block->warps[f->t.warp];
    w.xchg[f->t.lane] = bits;
    lh_emu_syncwarp();
    const uint64_t r = w.xchg[src_lane];
    lh_emu_syncwarp();                             // the slot is not rewritten before every lane has read
    return r;
}

void lh_emu_cp_async16(uint32_t dst, const void* src)
{
    Fiber* f = tl_self;
    if (!f) die("cp.async outside a kernel");
    if ((dst & 15u) || ((uintptr_t)src & 15u)) die("cp.async 16: misaligned source or destination");
    const CpCopy c{dst, src};
    if (g_cp_lazy.load(std::memory_order_relaxed)) f->open.push_back(c);
    else {
        // eager: the bytes move now — but still fault on a source that is not readable at all
        complete(f, c);
    }
}

void lh_emu_cp_commit(void)
{
    Fiber* f = tl_self;
    if (!f) die("cp.async.commit_group outside a kernel");
    f->groups.emplace_back(std::move(f->open));
    f->open.clear();
}

void lh_emu_cp_wait(int keep_newest)
{
    Fiber* f = tl_self;
    if (!f) die("cp.async.wait_group outside a kernel");
    while ((int)f->groups.size() > keep_newest) {
        for (auto& c : f->groups.front()) complete(f, c);
        f->groups.erase(f->groups.begin());
    }
}

int lh_emu_set_schedule(int policy, uint64_t seed)
{
    const int old = g_policy.exchange(policy < 0 || policy > 2 ? 0 : policy);
    g_seed.store(seed ? seed : 1);
    return old;
}
int lh_emu_set_cp_async_lazy(int lazy) { return g_cp_lazy.exchange(lazy != 0); }
int lh_emu_set_device_count(int n) { return g_devices.exchange(n < 0 ? 0 : n); }
int lh_emu_set_sm_count(int n) { return g_sms.exchange(n < 1 ? 1 : n); }
uint64_t lh_emu_launch_count(void) { return g_launches.load(); }
int64_t lh_emu_live_handles(void) { return g_live_handles.load(); }
void lh_emu_fail_allocation_in(int64_t n) { g_fail_alloc_in.store(n); }
uint64_t lh_emu_live_allocations(void)
{
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    return (uint64_t)g_allocs.size();
}

}  // extern "C"

// ---------------------------------------------------------------------- launches
namespace {

struct Job {
    dim3 grid, block;
    size_t smem_bytes = 0;
    const std::function<void()>* body = nullptr;
    std::atomic<uint64_t> next{0};
    uint64_t total = 0;
};

// One participant (the launching host thread, or a pool worker): takes block numbers until none are left.
void run_blocks(Job& job)
{
    const dim3 grid = job.grid, block = job.block;
    const size_t smem_bytes = job.smem_bytes;
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    const size_t nwarps = nthreads / 32;
    std::vector<Fiber> fibers;
    void* smem = nullptr;
    Block b;
    for (;;) {
        const uint64_t blockno = job.next.fetch_add(1);
        if (blockno >= job.total) break;
        if (fibers.empty()) {
            fibers.resize(nthreads);
            if (posix_memalign(&smem, 128, std::max<size_t>(smem_bytes, 128)) != 0) die("out of memory (shared memory)");
            b.body = job.body;
            b.fibers.resize(nthreads);
            for (size_t k = 0; k < nthreads; ++k) { fibers[k].stack = get_stack(k); b.fibers[k] = &fibers[k]; }
        }
        const unsigned bx = (unsigned)(blockno % grid.x), by = (unsigned)((blockno / grid.x) % grid.y), bz = (unsigned)(blockno / ((uint64_t)grid.x * grid.y));
        memset(smem, 0xff, std::max<size_t>(smem_bytes, 128));      // shared memory starts out as garbage
        b.warps.assign(nwarps, Warp());
        for (auto& w : b.warps) w.live = 32;
        b.arrived = 0; b.live = (int)nthreads; b.gen = 0; b.or_acc = b.or_result = 0;
        b.rng = g_seed.load() * 0x9e3779b97f4a7c15ull + blockno + 1;
        size_t k = 0;
        for (unsigned tz = 0; tz < block.z; ++tz)
            for (unsigned ty = 0; ty < block.y; ++ty)
                for (unsigned tx = 0; tx < block.x; ++tx, ++k) {
                    Fiber& f = fibers[k];
                    f.t.tid = uint3{tx, ty, tz};
                    f.t.bid = uint3{bx, by, bz};
                    f.t.bdim = block; f.t.gdim = grid;
                    f.t.lane = (int)(k & 31); f.t.warp = (int)(k >> 5);
                    f.t.smem = (char*)smem; f.t.smem_bytes = smem_bytes;
                    f.done = false; f.block = &b;
                    f.groups.clear(); f.open.clear();
                    fiber_init(&f, &b);
                }
        run_block(&b);
    }
    free(smem);
}

// Blocks of one grid are independent (that is CUDA's contract, and every kernel here honours it), so a grid of several blocks
// is spread over a small pool of host threads.  One launch at a time uses the pool; a launch that finds it busy (several host
// threads driving their own contexts) simply runs its blocks itself.
struct Pool {
    std::mutex launch_mutex;                      // held by the launch that owns the workers
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    Job* job = nullptr;
    uint64_t generation = 0;
    int working = 0;
    std::vector<std::thread> workers;

    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            Job* j;
            {
                std::unique_lock<std::mutex> lock(m);
                cv_work.wait(lock, [&] { return generation != seen; });
                seen = generation;
                j = job;
            }
            if (j) run_blocks(*j);
            {
                std::lock_guard<std::mutex> lock(m);
                if (--working == 0) cv_done.notify_all();
            }
        }
    }
};

Pool* pool()
{
    static Pool* p = [] {
        Pool* q = new Pool();                     // never destroyed: its threads outlive main()
        int n = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("LH_EMU_THREADS")) n = atoi(e);
        n = std::max(1, std::min(n, 16));
        for (int k = 0; k + 1 < n; ++k) { q->workers.emplace_back([q] { q->worker(); }); q->workers.back().detach(); }
        return q;
    }();
    return p;
}

}  // namespace

static void launch_now(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);

void lh_emu_launch(cudaStream_t stream, dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body)
{
    if (tl_self) die("kernel launch from inside a kernel");
    if (!g_async.load()) { launch_now(grid, block, smem_bytes, body); return; }
    std::lock_guard<std::recursive_mutex> lock(g_api);
    submit(stream, [=]() { launch_now(grid, block, smem_bytes, body); });
}

static void launch_now(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body)
{
    if (tl_self) die("kernel launch from inside a kernel");
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || nthreads % 32 != 0) die("block size must be a non-zero multiple of 32, at most 1024");
    if (smem_bytes > 227 * 1024) die("more than 227 KB of dynamic shared memory");
    ++g_launches;
    Job job;
    job.grid = grid; job.block = block; job.smem_bytes = smem_bytes; job.body = &body;
    job.total = (uint64_t)grid.x * grid.y * grid.z;
    if (job.total == 0) return;
    Pool* p = job.total > 1 ? pool() : nullptr;
    if (p && !p->workers.empty() && p->launch_mutex.try_lock()) {
        {
            std::lock_guard<std::mutex> lock(p->m);
            p->job = &job;
            p->working = (int)p->workers.size();
            ++p->generation;
        }
        p->cv_work.notify_all();
        run_blocks(job);
        {
            std::unique_lock<std::mutex> lock(p->m);
            p->cv_done.wait(lock, [&] { return p->working == 0; });
            p->job = nullptr;
        }
        p->launch_mutex.unlock();
    } else {
        run_blocks(job);
    }
}

// ---------------------------------------------------------------------- runtime API
extern "C" {

const char* cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorInvalidConfiguration: return "invalid configuration argument";
    case cudaErrorNoDevice: return "no CUDA-capable device is detected";
    case cudaErrorInvalidDevice: return "invalid device ordinal";
    default: return "unknown error";
    }
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n)
{
    *n = g_devices.load();
    return *n > 0 ? cudaSuccess : cudaErrorNoDevice;
}
cudaError_t cudaSetDevice(int d)
{
    if (d < 0 || d >= g_devices.load()) return cudaErrorInvalidDevice;
    tl_device = d;
    return cudaSuccess;
}
cudaError_t cudaGetDevice(int* d) { *d = tl_device; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d)
{
    if (d < 0 || d >= g_devices.load()) return cudaErrorInvalidDevice;
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "host emulation of a %d-SM device (tests/support/hostemu)", g_sms.load());
    p->multiProcessorCount = g_sms.load();
    p->totalGlobalMem = (size_t)180 << 30;
    return cudaSuccess;
}
cudaError_t lh_emu_malloc(void** p, size_t bytes) { return guarded_alloc(p, bytes, false); }
cudaError_t cudaFree(void* p)
{
    if (g_async.load()) { std::lock_guard<std::recursive_mutex> lock(g_api); drain_all(); }    // cudaFree synchronises the device
    return guarded_free(p, false);
}
cudaError_t lh_emu_malloc_host(void** p, size_t bytes) { return guarded_alloc(p, bytes, true); }
cudaError_t cudaFreeHost(void* p)
{
    if (g_async.load()) { std::lock_guard<std::recursive_mutex> lock(g_api); drain_all(); }
    return guarded_free(p, true);
}
// Blocking copies and memsets on the legacy default stream: the product's streams are cudaStreamNonBlocking, so these do NOT
// wait for them — they happen now.
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }

static cudaError_t copy_async(cudaStream_t s, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height)
{
    auto body = [=]() { for (size_t r = 0; r < height; ++r) memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width); };
    if (!g_async.load() || !s) { body(); return cudaSuccess; }
    std::lock_guard<std::recursive_mutex> lock(g_api);
    const size_t src_span = height ? (height - 1) * spitch + width : 0, dst_span = height ? (height - 1) * dpitch + width : 0;
    const bool src_dev = is_device(src), dst_dev = is_device(dst);
    if (!src_dev && !is_pinned_host(src, src_span)) {
        // pageable source: the bytes are staged before the call returns — the caller may reuse the buffer at once
        auto staged = std::make_shared<std::vector<char>>(height * width);
        for (size_t r = 0; r < height; ++r) memcpy(staged->data() + r * width, (const char*)src + r * spitch, width);
        submit(s, [=]() { for (size_t r = 0; r < height; ++r) memmove((char*)dst + r * dpitch, staged->data() + r * width, width); });
        return cudaSuccess;
    }
    if (!dst_dev && !is_pinned_host(dst, dst_span)) {
        // pageable destination: the call returns when the data has arrived, i.e. it waits for the stream up to this copy
        drain_stream(s);
        body();
        return cudaSuccess;
    }
    submit(s, body);                                  // device <-> device, device <-> pinned host: truly asynchronous
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t s) { return copy_async(s, dst, bytes, src, bytes, bytes, 1); }
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t s)
{
    if (width > dpitch || width > spitch) return cudaErrorInvalidValue;
    return copy_async(s, dst, dpitch, src, spitch, width, height);
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t s)
{
    if (!g_async.load() || !s) { memset(p, v, bytes); return cudaSuccess; }
    std::lock_guard<std::recursive_mutex> lock(g_api);
    submit(s, [=]() { memset(p, v, bytes); });
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned)
{
    std::lock_guard<std::recursive_mutex> lock(g_api);
    *s = new LhEmuStream();
    (*s)->device = tl_device;
    g_streams.push_back(*s);                          // streams are never freed: queued operations may outlive their handle
    ++g_live_handles;
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    std::lock_guard<std::recursive_mutex> lock(g_api);
    if (s) { drain_stream(s); --g_live_handles; }     // CUDA lets the queued work finish; here it finishes now
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s)
{
    if (!g_async.load() || !s) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lock(g_api);
    drain_stream(s);
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned)
{
    if (!g_async.load() || !s) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lock(g_api);
    if (e->recorded == 0 || e->completed >= e->recorded) return cudaSuccess;      // never recorded, or already complete: no-op
    LhEmuOp op;
    op.kind = LhEmuOp::WAIT;
    op.ev = e;
    op.seq = e->recorded;                             // the most recent record at the time of THIS call
    s->q.push_back(std::move(op));
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new LhEmuEvent(); ++g_live_handles; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { if (e) --g_live_handles; return cudaSuccess; }   // memory leaked on purpose: a queued record / wait may still name it
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    std::lock_guard<std::recursive_mutex> lock(g_api);
    ++e->recorded;
    if (!g_async.load() || !s) { e->completed = e->recorded; e->t_ms = now_ms(); return cudaSuccess; }
    LhEmuOp op;
    op.kind = LhEmuOp::RECORD;
    op.ev = e;
    op.seq = e->recorded;
    e->where.emplace_back(op.seq, s);
    s->q.push_back(std::move(op));
    background_progress();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e)
{
    if (!g_async.load()) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lock(g_api);
    const uint64_t want = e->recorded;
    progress_until([&] { return e->completed >= want; }, nullptr, e, want);
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    std::lock_guard<std::recursive_mutex> lock(g_api);
    if (a->recorded == 0 || b->recorded == 0 || a->completed < a->recorded || b->completed < b->recorded) return cudaErrorInvalidValue;
    *ms = (float)(b->t_ms - a->t_ms);
    return cudaSuccess;
}
int lh_emu_set_async(int mode, uint64_t seed)
{
    std::lock_guard<std::recursive_mutex> lock(g_api);
    if (g_async.load()) drain_all();
    g_async_rng = seed * 0x9e3779b97f4a7c15ull + 0x2545F4914F6CDD1Dull;
    return g_async.exchange(mode < 0 || mode > 2 ? 0 : mode);
}

}  // extern "C"
