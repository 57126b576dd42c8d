// TEST INFRASTRUCTURE -- a small SIMT emulator: runs CUDA kernels of this repository on the CPU, one host thread per CUDA
// thread, CTAs one after the other.  __syncthreads / __syncwarp are barriers, the warp collectives exchange values through
// a per-warp slot array between two barriers, atomics are GCC atomics on plain memory, `__shared__` variables are statics
// (one CTA at a time).  Only what the emulated kernels use is implemented: full-mask collectives, one-dimensional blocks
// whose size is a multiple of 32, no inter-CTA communication.  Arithmetic is the host's IEEE float32 with
// -ffp-contract=off -- the same operations nvcc --fmad=false emits -- so results can be compared with the oracle bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <functional>
#include <semaphore>
#include <type_traits>
#include <utility>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __constant__
#define __constant__ static          // per translation unit, as nvcc treats a __constant__ defined in a header
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

// Two ways to run the CUDA threads of a CTA:
//   default      FIBERS: every CUDA thread is a ucontext fiber of ONE host thread, switched at barriers only -- no futexes, no
//                kernel, deterministic interleaving; an order of magnitude faster than host threads for barrier-heavy kernels
//   -DEMU_THREADS  real host threads from a pool, std::barrier -- what ThreadSanitizer / AddressSanitizer need
#ifdef EMU_THREADS
struct EmuBar {
    std::barrier<> bar;
    explicit EmuBar(int n) : bar(n) {}
    void wait() { bar.arrive_and_wait(); }
    void drop() { bar.arrive_and_drop(); }
};
#else
static inline void emu_yield_until(const int *gen, int g);
struct EmuBar {
    int expected, arrived = 0, gen = 0;
    explicit EmuBar(int n) : expected(n) {}
    void wait()
    {
        const int g = gen;
        if (++arrived == expected) { arrived = 0; gen++; return; }
        emu_yield_until(&gen, g);
    }
    void drop() { if (--expected > 0 && arrived == expected) { arrived = 0; gen++; } }
};
#endif
struct EmuWarp { EmuBar bar{32}; unsigned long long slot[32]; bool alive[32]; };
struct EmuCta { std::unique_ptr<EmuBar> bar; std::vector<std::unique_ptr<EmuWarp>> warps; };
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
inline thread_local EmuWarp *emu_warp = nullptr;
inline thread_local EmuCta *emu_cta = nullptr;
inline thread_local int emu_lane = 0;

#ifndef EMU_THREADS
// Context switch between fibers.  glibc's swapcontext saves the signal mask with a system call on every switch, which
// dominated the emulation; on x86-64 the switch is written out (callee-saved registers + SSE / x87 control words), anywhere
// else ucontext is used.
#if defined(__x86_64__)
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .weak emu_switch
    .type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch,.-emu_switch
)");
struct EmuCtx {
    void *sp = nullptr;
    void prepare(char *stack, size_t size, void (*entry)())
    {
        uintptr_t top = ((uintptr_t)stack + size) & ~(uintptr_t)15;
        void **p = (void **)(top - 16);                 // return-address slot: entry() starts with rsp % 16 == 8
        *p = (void *)entry;
        p -= 6;                                         // rbp rbx r12 r13 r14 r15
        for (int i = 0; i < 6; i++) p[i] = nullptr;
        p -= 1;                                         // mxcsr | x87 control word
        unsigned mx; unsigned short cw;
        asm volatile("stmxcsr %0" : "=m"(mx));
        asm volatile("fnstcw %0" : "=m"(cw));
        unsigned long long fp = (unsigned long long)mx | ((unsigned long long)cw << 32);
        memcpy(p, &fp, 8);
        sp = (void *)p;
    }
};
static inline void emu_ctx_switch(EmuCtx &from, EmuCtx &to) { emu_switch(&from.sp, to.sp); }
#else
struct EmuCtx {
    ucontext_t uc;
    void prepare(char *stack, size_t size, void (*entry)())
    { getcontext(&uc); uc.uc_stack.ss_sp = stack; uc.uc_stack.ss_size = size; uc.uc_link = nullptr; makecontext(&uc, entry, 0); }
};
static inline void emu_ctx_switch(EmuCtx &from, EmuCtx &to) { swapcontext(&from.uc, &to.uc); }
#endif

// round-robin fiber scheduler; a fiber parked on a barrier is not even switched to until the barrier's generation moves
struct EmuFibers {
    struct Fiber { EmuCtx ctx; const int *wait_gen = nullptr; int wait_val = 0; bool done = false; };
    static constexpr size_t STACK = 256 * 1024;
    EmuCtx main_ctx;
    std::vector<Fiber> fibers;
    std::vector<char *> stacks;
    std::function<void(int)> body;
    EmuCta *cta = nullptr;
    int cur = -1;
    static void trampoline();
    void run(int n, EmuCta *c, std::function<void(int)> f)
    {
        body = std::move(f); cta = c;
        while ((int)stacks.size() < n) stacks.push_back((char *)malloc(STACK));
        fibers.assign(n, Fiber());
        for (int t = 0; t < n; t++) fibers[t].ctx.prepare(stacks[t], STACK, &EmuFibers::trampoline);
        int remaining = n;
        while (remaining) {
            bool progress = false;
            for (int t = 0; t < n; t++) {
                Fiber &fb = fibers[t];
                if (fb.done || (fb.wait_gen && *fb.wait_gen == fb.wait_val)) continue;
                fb.wait_gen = nullptr;
                cur = t;
                threadIdx = make_uint3((unsigned)t, 0, 0); emu_warp = cta->warps[t / 32].get(); emu_lane = t % 32;
                emu_ctx_switch(main_ctx, fb.ctx);
                progress = true;
                if (fb.done) remaining--;
            }
            if (!progress) { fprintf(stderr, "[cuda emulation] deadlock: every live CUDA thread waits on a barrier\n"); abort(); }
        }
        cur = -1;
    }
    void park(const int *gen, int g)
    {
        Fiber &fb = fibers[cur];
        fb.wait_gen = gen; fb.wait_val = g;
        emu_ctx_switch(fb.ctx, main_ctx);
    }
};
inline EmuFibers &emu_fibers() { static EmuFibers *p = new EmuFibers(); return *p; }
inline void EmuFibers::trampoline()
{
    EmuFibers &s = emu_fibers();
    const int t = s.cur;
    s.body(t);
    s.fibers[t].done = true;
    emu_ctx_switch(s.fibers[t].ctx, s.main_ctx);        // never resumed
    abort();
}
static inline void emu_yield_until(const int *gen, int g) { emu_fibers().park(gen, g); }
#endif

static inline void __syncthreads() { emu_cta->bar->wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp->bar.wait(); }
static inline void __threadfence() {}

// publish a value, wait, let `f` read every lane's value, wait again (the slots may be overwritten afterwards)
template <class F> static inline auto emu_collective(unsigned long long mine, F f)
{
    emu_warp->slot[emu_lane] = mine;
    emu_warp->bar.wait();
    auto r = f(emu_warp->slot, emu_warp->alive);
    emu_warp->bar.wait();
    return r;
}
template <class T> static inline unsigned long long emu_bits(T v) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_from(unsigned long long u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <class T> static inline T __shfl_sync(unsigned, T v, int src)
{ return emu_collective(emu_bits(v), [&](const unsigned long long *s, const bool *) { return emu_from<T>(s[src & 31]); }); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d)
{ return emu_collective(emu_bits(v), [&](const unsigned long long *s, const bool *) { return emu_from<T>(s[emu_lane - d >= 0 ? emu_lane - d : emu_lane]); }); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d)
{ return emu_collective(emu_bits(v), [&](const unsigned long long *s, const bool *) { return emu_from<T>(s[emu_lane + d < 32 ? emu_lane + d : emu_lane]); }); }
static inline unsigned __ballot_sync(unsigned, int pred)
{ return emu_collective(pred ? 1ull : 0ull, [&](const unsigned long long *s, const bool *a) { unsigned m = 0; for (int l = 0; l < 32; l++) if (a[l] && s[l]) m |= 1u << l; return m; }); }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0u; }
// reductions honour the member mask: sub-warp groups (the octets of csrc/octbox.cuh) reduce among themselves, provided every
// live lane of the warp reaches the same call (each with its own group's mask) -- which is how the kernels use them
template <class T, class Op> static inline T emu_reduce(unsigned mask, T v, Op op)
{ return emu_collective(emu_bits(v), [&](const unsigned long long *s, const bool *a) { T r = v; for (int l = 0; l < 32; l++) if (a[l] && ((mask >> l) & 1u)) r = op(r, emu_from<T>(s[l])); return r; }); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline int __reduce_max_sync(unsigned m, int v) { return emu_reduce(m, v, [](int a, int b) { return a > b ? a : b; }); }
static inline int __reduce_min_sync(unsigned m, int v) { return emu_reduce(m, v, [](int a, int b) { return a < b ? a : b; }); }
static inline int __reduce_add_sync(unsigned, int v)
{ return emu_collective(emu_bits(v), [&](const unsigned long long *s, const bool *a) { int r = 0; for (int l = 0; l < 32; l++) if (a[l]) r += emu_from<int>(s[l]); return r; }); }
template <class T> static inline unsigned __match_any_sync(unsigned, T v)
{ return emu_collective(emu_bits(v), [&](const unsigned long long *s, const bool *a) { unsigned m = 0; for (int l = 0; l < 32; l++) if (a[l] && s[l] == emu_bits(v)) m |= 1u << l; return m; }); }

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
// kernels passed by name to the runtime's attribute / occupancy queries
template <class T> static inline cudaError_t cudaFuncSetAttribute(T *, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, T *, int, size_t) { *n = 1; return cudaSuccess; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)((((unsigned long long)hi << 32) | lo) >> (shift & 31)); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
using std::isfinite;
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMin(T *p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <class T> static inline T atomicCAS(T *p, T cmp, T val) { T e = cmp; __atomic_compare_exchange_n(p, &e, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return e; }

// cooperative launches run as ONE CTA here (the fake runtime reports one SM and one resident CTA), so a grid sync is the
// CTA barrier and `__shared__` statics stay private to the only CTA there is
namespace cooperative_groups {
struct grid_group { void sync() const { emu_cta->bar->wait(); } };
static inline grid_group this_grid() { return grid_group(); }
}

// The host threads that play the CUDA threads live in a pool (creating 512 threads per CTA dominated the run time):
// worker t sleeps on its own semaphore, a launch wakes workers 0 .. blockDim-1 once per CTA and waits for all of them.
struct EmuPool {
    struct Worker { std::thread th; std::binary_semaphore go{0}; };
    std::vector<std::unique_ptr<Worker>> workers;
    std::function<void(int)> job;
    std::atomic<int> remaining{0};
    std::binary_semaphore done{0};
    std::atomic<bool> quit{false};
    void ensure(int n)
    {
        while ((int)workers.size() < n) {
            const int id = (int)workers.size();
            workers.emplace_back(new Worker());
            Worker *w = workers.back().get();
            w->th = std::thread([this, w, id] {
                for (;;) {
                    w->go.acquire();
                    if (quit.load()) return;
                    job(id);
                    if (remaining.fetch_sub(1) == 1) done.release();
                }
            });
        }
    }
    void run(int n, std::function<void(int)> f)
    {
        ensure(n);
        job = std::move(f);
        remaining.store(n);
        for (int t = 0; t < n; t++) workers[t]->go.release();
        done.acquire();
    }
    ~EmuPool()
    {
        quit.store(true);
        for (auto &w : workers) w->go.release();
        for (auto &w : workers) w->th.join();
    }
};
inline EmuPool &emu_pool() { static EmuPool *p = new EmuPool(); return *p; }     // leaked on purpose: no join at exit

// kernel<<<grid, block>>>(args...): CTAs in order, the threads of a CTA concurrently
template <class K, class... A> static void emu_launch(K kernel, dim3 grid, dim3 block, A... args)
{
    const int nthr = (int)block.x, nwarp = nthr / 32;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        EmuCta cta;
        cta.bar.reset(new EmuBar(nthr));
        for (int w = 0; w < nwarp; w++) { cta.warps.emplace_back(new EmuWarp()); for (int l = 0; l < 32; l++) cta.warps[w]->alive[l] = true; }
        auto body = [&](int t) {
            threadIdx = make_uint3((unsigned)t, 0, 0); blockIdx = make_uint3(bx, by, bz); blockDim = block; gridDim = grid;
            emu_cta = &cta; emu_warp = cta.warps[t / 32].get(); emu_lane = t % 32;
            kernel(args...);
            EmuWarp *w = cta.warps[t / 32].get();     // (a fiber's thread-locals are only valid while it runs)
            w->alive[t % 32] = false;                 // an exited thread no longer takes part in barriers or collectives
            w->bar.drop();
            cta.bar->drop();
        };
#ifdef EMU_THREADS
        emu_pool().run(nthr, body);
#else
        blockIdx = make_uint3(bx, by, bz); blockDim = block; gridDim = grid; emu_cta = &cta;
        emu_fibers().run(nthr, &cta, body);
#endif
    }
}

// cudaLaunchCooperativeKernel((void *)kernel, grid, block, args, ...): the argument array is unpacked by the kernel's own
// parameter types
template <class... P, size_t... I> static void emu_coop_call(void (*k)(P...), void **args, std::index_sequence<I...>)
{ k(*reinterpret_cast<std::remove_reference_t<P> *>(args[I])...); }
template <class... P> static cudaError_t emu_launch_coop(void (*k)(P...), dim3 grid, dim3 block, void **args)
{
    if (grid.x * grid.y * grid.z != 1) return cudaErrorCooperativeLaunchTooLarge;
    emu_launch([=] { emu_coop_call(k, args, std::index_sequence_for<P...>()); }, grid, block);
    return cudaSuccess;
}
