// emu_rt.cpp -- fiber scheduler and runtime stubs of the CPU kernel emulator.  TEST INFRASTRUCTURE ONLY.
//
// One CUDA block at a time; each thread of the block is a ucontext fiber.  A fiber runs until it reaches a
// rendezvous (__syncthreads, a *_sync warp primitive) or returns.  The scheduler releases
//   * a block barrier when every thread that has not returned waits at it (CUDA's rule), and
//   * a warp exchange when every live lane named in the mask has arrived with the same mask;
// if neither is possible and nobody can run, the kernel has a divergent barrier: the emulator says so and aborts.
// "Device" memory is host memory filled with 0xCD at allocation, so a kernel that relies on zeroed allocations
// computes garbage here too.
#include "cuda_runtime.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

// All scheduler state is per HOST thread: the ranks of an emulated decomposed run are threads that launch kernels
// concurrently (a kernel of one rank may spin on a flag that a kernel of another rank raises in peer-mapped memory).
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

// Context switch.  x86-64: a dozen instructions (callee-saved registers pushed on the outgoing stack, stack pointers
// exchanged) -- swapcontext() makes two signal-mask system calls per switch, and a 1024-thread reduction kernel switches
// ~50 000 times.  Elsewhere: ucontext.
#if defined(__x86_64__) && !defined(EMU_UCONTEXT)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void **save_sp, void *load_sp);
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
#else
#define EMU_FAST_SWITCH 0
#endif

namespace {
enum { READY = 0, AT_BARRIER = 1, AT_WARP = 2, DONE = 3 };
struct Fiber {
#if EMU_FAST_SWITCH
    void *sp;             // saved stack pointer (callee-saved registers sit on the fiber's own stack)
#else
    ucontext_t ctx;
#endif
    int state;
    unsigned wmask;
    uint64_t deposit;
    unsigned part;        // lanes that took part in the exchange this fiber was just released from
    uint3 tid;
    int lin;
    int gseq;             // marked gathers issued so far (gather statistics)
};
const size_t STACK_BYTES = 256 * 1024;
const int MAX_THREADS = 1024;
thread_local char *g_stacks = nullptr;
thread_local std::vector<Fiber> g_fibers;
#if EMU_FAST_SWITCH
thread_local void *g_sched_sp = nullptr;
#else
thread_local ucontext_t g_sched;
#endif
thread_local Fiber *g_cur = nullptr;
thread_local const std::function<void()> *g_body = nullptr;
thread_local const char *g_kernel = "?";
thread_local std::vector<unsigned char> g_dyn;
thread_local uint64_t g_wslots[32][32];          // per warp: the deposits of the last released exchange
thread_local int g_at_warp[32];                   // per warp: fibers waiting at a warp primitive
long long g_launches = 0;
// gather statistics of the block being run: (warp << 32 | sequence number, 128-byte line)
thread_local std::vector<std::pair<unsigned long long, unsigned long long>> g_gathers;
int g_gather_stats = -1;

void flush_gathers()
{
    if (g_gathers.empty()) return;
    std::sort(g_gathers.begin(), g_gathers.end());
    long long requests = 0, lines = 0;
    for (size_t k = 0; k < g_gathers.size(); k++) {
        if (k == 0 || g_gathers[k].first != g_gathers[k - 1].first) { requests++; lines++; }
        else if (g_gathers[k].second != g_gathers[k - 1].second) lines++;
    }
    __atomic_fetch_add(&sepgpu_emu_counter[1], requests, __ATOMIC_RELAXED);
    __atomic_fetch_add(&sepgpu_emu_counter[2], lines, __ATOMIC_RELAXED);
    __atomic_fetch_add(&sepgpu_emu_counter[3], (long long)g_gathers.size(), __ATOMIC_RELAXED);
    g_gathers.clear();
}

#if EMU_FAST_SWITCH
void fiber_main()
{
    (*g_body)();
    g_cur->state = DONE;
    emu_switch(&g_cur->sp, g_sched_sp);          // never resumed
    __builtin_trap();
}

void yield_to_scheduler()
{
    Fiber *me = g_cur;
    emu_switch(&me->sp, g_sched_sp);
    // resumed: the scheduler has restored threadIdx and g_cur
}

void fiber_init(Fiber &f, char *stack, size_t bytes)
{
    // a fresh stack that emu_switch can "return" into: six zeroed callee-saved registers, then fiber_main as the return
    // address; after that ret the stack pointer is 8 mod 16, as at any function entry
    uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;                             // where fiber_main would return to (it never does)
    *--sp = (void *)fiber_main;
    for (int k = 0; k < 6; k++) *--sp = nullptr;
    f.sp = sp;
}

inline void resume(Fiber &f) { emu_switch(&g_sched_sp, f.sp); }
#else
void fiber_main()
{
    (*g_body)();
    g_cur->state = DONE;
    // returning resumes uc_link == g_sched
}

void yield_to_scheduler()
{
    Fiber *me = g_cur;
    swapcontext(&me->ctx, &g_sched);
    // resumed: the scheduler has restored threadIdx and g_cur
}

void fiber_init(Fiber &f, char *stack, size_t bytes)
{
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stack;
    f.ctx.uc_stack.ss_size = bytes;
    f.ctx.uc_link = &g_sched;
    makecontext(&f.ctx, fiber_main, 0);
}

inline void resume(Fiber &f) { swapcontext(&g_sched, &f.ctx); }
#endif

void run_block(int nthreads)
{
    int alive = nthreads;
    for (int w = 0; w < 32; w++) g_at_warp[w] = 0;
    for (;;) {
        bool ran = false;
        for (int t = 0; t < nthreads; t++) {
            Fiber &f = g_fibers[t];
            if (f.state != READY) continue;
            threadIdx = f.tid;
            g_cur = &f;
            resume(f);
            ran = true;
            if (f.state == DONE) alive--;
        }
        if (alive == 0) return;
        // block barrier
        int at_bar = 0;
        for (int t = 0; t < nthreads; t++) at_bar += g_fibers[t].state == AT_BARRIER;
        bool released = false;
        if (at_bar == alive) {
            for (int t = 0; t < nthreads; t++)
                if (g_fibers[t].state == AT_BARRIER) g_fibers[t].state = READY;
            released = true;
        }
        // warp exchanges: a group (lanes waiting with the same mask) is released when every live lane it names has arrived
        for (int w0 = 0, w = 0; w0 < nthreads; w0 += 32, w++) {
            if (g_at_warp[w] == 0) continue;
            const int wn = nthreads - w0 < 32 ? nthreads - w0 : 32;
            unsigned live = 0, waiting = 0;
            for (int l = 0; l < wn; l++) {
                const int st = g_fibers[w0 + l].state;
                if (st != DONE) live |= 1u << l;
                if (st == AT_WARP) waiting |= 1u << l;
            }
            unsigned todo = waiting;
            while (todo) {
                const int l = __builtin_ctz(todo);
                const unsigned m = g_fibers[w0 + l].wmask;
                const unsigned need = m & live & (wn == 32 ? 0xffffffffu : ((1u << wn) - 1u));
                unsigned same = 0;                                   // waiting lanes with this very mask
                for (unsigned t = waiting; t; t &= t - 1) {
                    const int q = __builtin_ctz(t);
                    if (g_fibers[w0 + q].wmask == m) same |= 1u << q;
                }
                todo &= ~same;
                if ((need & same) != need) continue;                 // somebody named in the mask is still on the way
                for (unsigned t = need; t; t &= t - 1) {
                    const int q = __builtin_ctz(t);
                    g_wslots[w][q] = g_fibers[w0 + q].deposit;
                }
                for (unsigned t = need; t; t &= t - 1) {
                    Fiber &g = g_fibers[w0 + __builtin_ctz(t)];
                    g.part = need;
                    g.state = READY;
                }
                g_at_warp[w] -= __builtin_popcount(need);
                released = true;
            }
        }
        if (!ran && !released) {
            int nb = 0, nw = 0;
            for (int t = 0; t < nthreads; t++) { nb += g_fibers[t].state == AT_BARRIER; nw += g_fibers[t].state == AT_WARP; }
            fprintf(stderr, "emu: DEADLOCK in kernel %s, block (%u,%u,%u): %d threads alive, %d at __syncthreads, %d at a warp "
                            "primitive -- divergent barrier or a mask naming lanes that never arrive\n",
                    g_kernel, blockIdx.x, blockIdx.y, blockIdx.z, alive, nb, nw);
            abort();
        }
    }
}
}  // namespace

namespace emu {
dim3 d3(dim3 v) { return v; }

void *dyn_smem() { return g_dyn.data(); }

const char *self_path()
{
    static Dl_info info;
    if (!info.dli_fname) dladdr((void *)&self_path, &info);
    return info.dli_fname ? info.dli_fname : "";
}

int lane_id() { return g_cur->lin & 31; }

void record_gather(const void *p)
{
    if (g_gather_stats < 0) { const char *e = getenv("SEPGPU_EMU_GATHER_STATS"); g_gather_stats = e && e[0] == '1'; }
    if (!g_gather_stats) return;
    Fiber *me = g_cur;
    g_gathers.emplace_back(((unsigned long long)(me->lin >> 5) << 32) | (unsigned)me->gseq++, (unsigned long long)(uintptr_t)p >> 7);
}

void sync_block()
{
    g_cur->state = AT_BARRIER;
    yield_to_scheduler();
}

const uint64_t *warp_exchange(unsigned mask, uint64_t mine, unsigned *part)
{
    Fiber *me = g_cur;
    me->state = AT_WARP;
    me->wmask = mask;
    me->deposit = mine;
    g_at_warp[me->lin >> 5]++;
    yield_to_scheduler();
    *part = me->part;
    return g_wslots[me->lin >> 5];
}

void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > MAX_THREADS || grid.x == 0 || grid.y == 0 || grid.z == 0) {
        fprintf(stderr, "emu: invalid launch configuration for %s: grid (%u,%u,%u) block (%u,%u,%u)\n", name, grid.x, grid.y,
                grid.z, block.x, block.y, block.z);
        abort();
    }
    if (smem > 227 * 1024) { fprintf(stderr, "emu: %s asks for %zu bytes of shared memory (> 227 KB)\n", name, smem); abort(); }
    if (!g_stacks) {
        g_stacks = (char *)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == (char *)MAP_FAILED) { perror("emu: mmap"); abort(); }
        g_fibers.resize(MAX_THREADS);
    }
    __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED);
    g_kernel = name;
    g_body = &body;
    blockDim = block;
    gridDim = grid;
    g_dyn.assign(smem + 128, 0xCD);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                if (smem) memset(g_dyn.data(), 0xCD, smem);
                int t = 0;
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++, t++) {
                            Fiber &f = g_fibers[t];
                            fiber_init(f, g_stacks + (size_t)t * STACK_BYTES, STACK_BYTES);
                            f.state = READY;
                            f.tid.x = tx; f.tid.y = ty; f.tid.z = tz;
                            f.lin = t;
                            f.gseq = 0;
                        }
                run_block(nthreads);
                flush_gathers();
            }
    g_cur = nullptr;
}
}  // namespace emu

extern "C" long long sepgpu_emu_launches(void) { return g_launches; }
extern "C" { long long sepgpu_emu_counter[8]; }

// ---- runtime ------------------------------------------------------------------------------------------------------------
struct emuStream { int dummy; };
struct emuEvent { double t_ms; };

static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotSupported: return "operation not supported (kernel emulator)";
    default: return "unknown error";
    }
}
// SEPGPU_EMU_GUARD=1: every device allocation ends (to 32 bytes, the widest vector access) at an inaccessible page, so that a
// kernel reading or writing past its array stops the run with SIGSEGV at the offending access instead of passing by luck.
static bool guard_mode(void)
{
    static const bool on = getenv("SEPGPU_EMU_GUARD") && atoi(getenv("SEPGPU_EMU_GUARD")) != 0;
    return on;
}
static std::mutex g_guard_mu;
static std::unordered_map<void *, std::pair<void *, size_t>> g_guard_map;      // user pointer -> (mapping, length)

cudaError_t cudaMalloc(void **p, size_t bytes)
{
    if (guard_mode()) {
        const size_t page = 4096, body = ((bytes ? bytes : 32) + 31) & ~(size_t)31;
        const size_t len = ((body + page - 1) / page + 1) * page;
        char *base = (char *)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
        mprotect(base + len - page, page, PROT_NONE);
        char *q = base + len - page - body;
        memset(q, 0xCD, bytes);
        std::lock_guard<std::mutex> lk(g_guard_mu);
        g_guard_map[q] = {base, len};
        *p = q;
        return cudaSuccess;
    }
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 256)) return cudaErrorMemoryAllocation;
    memset(q, 0xCD, bytes);
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (guard_mode()) {
        if (!p) return cudaSuccess;
        std::lock_guard<std::mutex> lk(g_guard_mu);
        auto it = g_guard_map.find(p);
        if (it == g_guard_map.end()) return cudaErrorInvalidValue;
        munmap(it->second.first, it->second.second);
        g_guard_map.erase(it);
        return cudaSuccess;
    }
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes)
{
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 256)) return cudaErrorMemoryAllocation;
    memset(q, 0xCD, bytes);
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent(); (*e)->t_ms = 0; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
// "Inter-process" handles inside one process: the handle carries the pointer (ranks are threads of this process).
// SEPGPU_EMU_NO_IPC=1 makes the export fail, which sends the library down its NCCL-only path.
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
    const char *no = getenv("SEPGPU_EMU_NO_IPC");
    if (no && no[0] == '1') return cudaErrorNotSupported;
    memset(h, 0, sizeof *h);
    memcpy(h->reserved, &p, sizeof p);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return *p ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
