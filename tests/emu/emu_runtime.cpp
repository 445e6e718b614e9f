// tests/emu/emu_runtime.cpp -- host-side functional simulator for the kernels in
// aes_fhe_b200/csrc (TEST INFRASTRUCTURE ONLY; see csrc/compat.h).  A launch runs
// its blocks on a few host threads; inside a block every CUDA thread is a ucontext fiber and
// __syncthreads() yields back to a round-robin scheduler, which resumes the fibers only
// after all live ones have reached the barrier.
#include "compat.h"

#include <ucontext.h>
#include <atomic>
#include <thread>
#include <vector>
#include <cstdio>

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

// dynamic shared memory: one buffer per host worker thread (= per running block)
static size_t g_dyn_bytes = 0;
void fhe_emu_set_dyn_smem(size_t bytes) { g_dyn_bytes = bytes; }
char* fhe_emu_dyn_smem() {
    static thread_local std::vector<char> buf;
    if (buf.size() < g_dyn_bytes + 16) buf.resize(g_dyn_bytes + 16);
    return buf.data();
}

namespace {
constexpr size_t kStack = 128 * 1024;

struct BlockRunner {
    ucontext_t main_ctx;
    std::vector<ucontext_t> fibers;
    std::vector<char> stacks;
    std::vector<char> finished;
    const std::function<void()>* body = nullptr;
    int current = -1;
};
thread_local BlockRunner* g_runner = nullptr;

void trampoline() {
    BlockRunner* r = g_runner;
    int me = r->current;
    (*r->body)();
    r->finished[me] = 1;
    swapcontext(&r->fibers[me], &r->main_ctx);
}

void run_block(BlockRunner& r, dim3 b, const std::function<void()>& body) {
    const int nt = (int)(b.x * b.y * b.z);
    r.fibers.resize(nt);
    r.finished.assign(nt, 0);
    if (r.stacks.size() < (size_t)nt * kStack) r.stacks.resize((size_t)nt * kStack);
    r.body = &body;
    g_runner = &r;
    for (int t = 0; t < nt; ++t) {
        getcontext(&r.fibers[t]);
        r.fibers[t].uc_stack.ss_sp = r.stacks.data() + (size_t)t * kStack;
        r.fibers[t].uc_stack.ss_size = kStack;
        r.fibers[t].uc_link = &r.main_ctx;
        makecontext(&r.fibers[t], (void (*)())trampoline, 0);
    }
    int live = nt;
    while (live > 0) {
        live = 0;
        for (int t = 0; t < nt; ++t) {
            if (r.finished[t]) continue;
            r.current = t;
            threadIdx = dim3(t % b.x, (t / b.x) % b.y, t / (b.x * b.y));
            swapcontext(&r.main_ctx, &r.fibers[t]);
            if (!r.finished[t]) ++live;
        }
    }
}
}  // namespace

void __syncthreads() {
    BlockRunner* r = g_runner;
    int me = r->current;
    swapcontext(&r->fibers[me], &r->main_ctx);
    // resumed by the scheduler in the next round: restore identity
    r->current = me;
}

void fhe_emu_launch(dim3 g, dim3 b, const std::function<void()>& body) {
    const long nblocks = (long)g.x * g.y * g.z;
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = (int)std::min<long>(nblocks, hw ? hw : 4);
    std::atomic<long> next(0);
    auto worker = [&]() {
        BlockRunner runner;
        gridDim = g;
        blockDim = b;
        for (;;) {
            long i = next.fetch_add(1);
            if (i >= nblocks) break;
            blockIdx = dim3((unsigned)(i % g.x), (unsigned)((i / g.x) % g.y), (unsigned)(i / ((long)g.x * g.y)));
            run_block(runner, b, body);
        }
    };
    if (nthreads <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (int i = 0; i < nthreads; ++i) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
}

// Persistent kernels whose blocks wait for each other (csrc/ntt_fused.cuh) need every block
// running at once: one host thread per block.
void fhe_emu_launch_coresident(dim3 g, dim3 b, const std::function<void()>& body) {
    const long nblocks = (long)g.x * g.y * g.z;
    std::vector<std::thread> pool;
    for (long i = 0; i < nblocks; ++i)
        pool.emplace_back([&, i]() {
            BlockRunner runner;
            gridDim = g;
            blockDim = b;
            blockIdx = dim3((unsigned)(i % g.x), (unsigned)((i / g.x) % g.y), (unsigned)(i / ((long)g.x * g.y)));
            run_block(runner, b, body);
        });
    for (auto& t : pool) t.join();
}
