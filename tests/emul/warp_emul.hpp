// tests/emul/warp_emul.hpp -- 32-fiber lockstep emulation of one CUDA warp on the CPU.
//
// TEST VEHICLE ONLY.  It lets tests run the device source of uncalled_b200/csrc/*.cuh
// (compiled with -DUNC_EMUL) on a box without a GPU, so the warp-cooperative logic can be
// compared with the oracle before any GPU time is spent.  Every warp collective is a
// rendezvous of all 32 fibers (full-mask semantics); a lane that skips a collective the
// others reach is reported as a deadlock, which is exactly the bug it would be on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
// Fiber switch: callee-saved registers + stack pointer only.  glibc's swapcontext also saves/restores the signal
// mask with a system call per switch, which dominated the emulator's run time (the fibers switch at every collective).
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".hidden emu_switch\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size emu_switch,.-emu_switch\n");
#define EMU_FAST_SWITCH 1
#else
#include <ucontext.h>
#endif

#define UNC_DEV static inline
#define UNC_DEV_NOINLINE static
#define UNC_FULL 0xffffffffu

struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
static inline uint3 make_uint3(uint32_t x, uint32_t y, uint32_t z) { uint3 r = {x, y, z}; return r; }
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }

#define EMU_MAX_THREADS 1024
struct WarpRv {      // per-warp rendezvous state
    uint64_t slot[32], result[32];
    int arrived;
    uint64_t gen;
};
struct WarpEmu {     // one CTA: n_threads fibers, 32 per warp
#ifdef EMU_FAST_SWITCH
    void *ctx[EMU_MAX_THREADS], *main_ctx;      // saved stack pointers
#else
    ucontext_t ctx[EMU_MAX_THREADS], main_ctx;
#endif
    char *stacks[EMU_MAX_THREADS];
    int n_threads;
    int cur;
    int done[EMU_MAX_THREADS];
    int n_done;
    int warp_done[EMU_MAX_THREADS / 32];   // exited threads per warp
    WarpRv rv[EMU_MAX_THREADS / 32];
    int cta_arrived;
    uint64_t cta_gen;
    int sub_arrived[16];
    uint64_t sub_gen[16];
    void (*fn)(void *);
    void *arg;
};
extern thread_local WarpEmu *g_warp;

#ifdef EMU_FAST_SWITCH
static inline void emu_swap(WarpEmu *w, int from, int to) { emu_switch(&w->ctx[from], w->ctx[to]); }
static inline void emu_swap_to_main(WarpEmu *w, int from) { emu_switch(&w->ctx[from], w->main_ctx); }
#else
static inline void emu_swap(WarpEmu *w, int from, int to) { swapcontext(&w->ctx[from], &w->ctx[to]); }
static inline void emu_swap_to_main(WarpEmu *w, int from) { swapcontext(&w->ctx[from], &w->main_ctx); }
#endif

static inline void emu_yield() {
    WarpEmu *w = g_warp;
    int from = w->cur, nxt = from;
    for (int i = 0; i < w->n_threads; i++) {
        nxt = nxt + 1 == w->n_threads ? 0 : nxt + 1;
        if (!w->done[nxt]) break;
    }
    if (nxt == from) return;
    w->cur = nxt;
    emu_swap(w, from, nxt);
}

// all-lane exchange within the calling fiber's warp: every lane posts v, then may read any
// lane's value from result[]
static inline const uint64_t *emu_exchange(uint64_t v) {
    WarpEmu *w = g_warp;
    if (w->warp_done[w->cur >> 5]) { fprintf(stderr, "warp_emul: collective reached after %d thread(s) of the warp exited\n", w->warp_done[w->cur >> 5]); abort(); }
    int lane = w->cur & 31;
    WarpRv *r = &w->rv[w->cur >> 5];
    r->slot[lane] = v;
    uint64_t mygen = r->gen;
    if (++r->arrived == 32) {
        memcpy(r->result, r->slot, sizeof(r->slot));
        r->arrived = 0;
        r->gen++;
    } else {
        long spins = 0;
        while (r->gen == mygen) {
            emu_yield();
            if (++spins > 50000000L) { fprintf(stderr, "warp_emul: deadlock at a warp collective (thread %d)\n", w->cur); abort(); }
        }
    }
    return r->result;
}

static inline int w_lane() { return g_warp->cur & 31; }
static inline int c_tid() { return g_warp->cur; }
static inline int c_nthreads() { return g_warp->n_threads; }
// __syncthreads()
static inline void c_sync() {
    WarpEmu *w = g_warp;
    if (w->n_done) { fprintf(stderr, "warp_emul: CTA barrier reached after %d thread(s) exited\n", w->n_done); abort(); }
    uint64_t mygen = w->cta_gen;
    if (++w->cta_arrived == w->n_threads) {
        w->cta_arrived = 0;
        w->cta_gen++;
    } else {
        long spins = 0;
        while (w->cta_gen == mygen) {
            emu_yield();
            if (++spins > 50000000L) { fprintf(stderr, "warp_emul: deadlock at a CTA barrier (thread %d)\n", w->cur); abort(); }
        }
    }
}
// bar.sync id, count : named barrier over `count` threads
static inline void c_sync_sub(int id, int count) {
    WarpEmu *w = g_warp;
    uint64_t mygen = w->sub_gen[id];
    if (++w->sub_arrived[id] == count) {
        w->sub_arrived[id] = 0;
        w->sub_gen[id]++;
    } else {
        long spins = 0;
        while (w->sub_gen[id] == mygen) {
            emu_yield();
            if (++spins > 50000000L) { fprintf(stderr, "warp_emul: deadlock at named barrier %d (thread %d)\n", id, w->cur); abort(); }
        }
    }
}
// bar.arrive id, count : count towards the named barrier without waiting for it
static inline void c_arrive_sub(int id, int count) {
    WarpEmu *w = g_warp;
    if (++w->sub_arrived[id] == count) {
        w->sub_arrived[id] = 0;
        w->sub_gen[id]++;
    }
}
// inside a spin-wait on shared memory written by another warp: let the others run
static inline void w_spin() { emu_yield(); }
static inline void c_fence() {}
static inline void w_sync() { emu_exchange(0); }
static inline uint32_t w_ballot(int p) {
    // copy out immediately: result[] is overwritten by the next collective
    uint64_t tmp[32];
    memcpy(tmp, emu_exchange(p ? 1 : 0), sizeof(tmp));
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (uint32_t) (tmp[i] & 1) << i;
    return m;
}
static inline uint32_t w_shfl(uint32_t v, int src) { return (uint32_t) emu_exchange(v)[src & 31]; }
static inline float w_shflf(float v, int src) {
    uint32_t b; memcpy(&b, &v, 4);
    b = w_shfl(b, src);
    float r; memcpy(&r, &b, 4);
    return r;
}
static inline uint32_t w_shfl_up(uint32_t v, int d) {
    int lane = w_lane();
    const uint64_t *r = emu_exchange(v);
    return lane >= d ? (uint32_t) r[lane - d] : v;
}
static inline uint32_t w_shfl_down(uint32_t v, int d) {
    int lane = w_lane();
    const uint64_t *r = emu_exchange(v);
    return lane + d < 32 ? (uint32_t) r[lane + d] : v;
}
static inline uint32_t w_match(uint32_t v) {
    uint64_t tmp[32];
    memcpy(tmp, emu_exchange(v), sizeof(tmp));
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) if ((uint32_t) tmp[i] == v) m |= 1u << i;
    return m;
}
static inline uint32_t d_atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t d_atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t s_atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t s_atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t s_atomic_max(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline int d_popc(uint32_t v) { return __builtin_popcount(v); }
static inline int d_popcll(uint64_t v) { return __builtin_popcountll(v); }
static inline int d_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int d_ffs(uint32_t v) { return __builtin_ffs((int) v); }
static inline int d_clzll(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
static inline int d_ctzll(uint64_t v) { return v ? __builtin_ctzll(v) : -1; }
// compiled with -ffp-contract=off: plain IEEE operations
static inline float f_mul(float a, float b) { return a * b; }
static inline float f_add(float a, float b) { return a + b; }
static inline float f_sub(float a, float b) { return a - b; }
static inline float f_div(float a, float b) { return a / b; }
static inline float f_sqrt(float a) { return sqrtf(a); }
static inline double d_mul(double a, double b) { return a * b; }
static inline double d_add(double a, double b) { return a + b; }
static inline double d_sub(double a, double b) { return a - b; }
static inline double d_div(double a, double b) { return a / b; }
static inline double d_sqrt(double a) { return sqrt(a); }
static inline uint32_t f_to_u32_x86(float v) { return (uint32_t) (long long) v; }
static inline uint64_t f_to_u64(float v) { return (uint64_t) v; }
template <typename T> static inline T d_ldg(const T *p) { return *p; }
static inline void d_prefetch(const void *p) { (void) p; }
static inline uint64_t s_load_u64(const uint64_t *p) { return *p; }
static inline void s_store_u64(uint64_t *p, uint64_t v) { *p = v; }
static inline uint4 s_load_v4(const uint4 *p) { return *p; }
static inline void s_store_v4(uint4 *p, uint4 v) { *p = v; }
static inline float u2f(uint32_t v) { float r; memcpy(&r, &v, 4); return r; }
static inline uint32_t f2u(float v) { uint32_t r; memcpy(&r, &v, 4); return r; }
static inline double d_fma(double a, double b, double c) { return fma(a, b, c); }
static inline float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
static inline double u2d(uint64_t v) { double r; memcpy(&r, &v, 8); return r; }
static inline uint64_t d2u(double v) { uint64_t r; memcpy(&r, &v, 8); return r; }
// bulk async copy: performed at issue; the barrier wait is a no-op
static inline void t_bar_init(uint64_t *bar) { *bar = 0; }
static inline void t_fence_async() {}
static inline void t_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if (((uintptr_t) dst & 15) || ((uintptr_t) src & 15) || (bytes & 15)) { fprintf(stderr, "warp_emul: misaligned bulk copy\n"); abort(); }
    memcpy(dst, src, bytes); (*bar)++;
}
static inline void t_bar_wait(uint64_t *bar, uint32_t parity) { (void) bar; (void) parity; }
static inline void a_copy16(void *dst, const void *src) { memcpy(dst, src, 16); }
static inline void a_commit() {}
static inline void a_wait_all() {}

static void emu_trampoline() {
    WarpEmu *w = g_warp;
    w->fn(w->arg);
    int me = w->cur;
    w->done[me] = 1;
    w->n_done++;
    w->warp_done[me >> 5]++;
    if (w->n_done == w->n_threads) {
        emu_swap_to_main(w, me);
    } else {
        // a thread left early: the others must not touch a collective any more (checked there)
        int nxt = me;
        for (int i = 0; i < w->n_threads; i++) { nxt = nxt + 1 == w->n_threads ? 0 : nxt + 1; if (!w->done[nxt]) break; }
        w->cur = nxt;
        emu_swap(w, me, nxt);
    }
    abort();      // a finished fiber is never resumed
}

// run fn(arg) on a CTA of n_threads (multiple of 32) fibers
static inline void emu_run_cta(void (*fn)(void *), void *arg, int n_threads) {
    WarpEmu *w = (WarpEmu *) calloc(1, sizeof(WarpEmu));
    WarpEmu *saved = g_warp;
    g_warp = w;
    w->fn = fn;
    w->arg = arg;
    w->n_threads = n_threads;
    const size_t STK = 1 << 19;
    for (int i = 0; i < n_threads; i++) {
        w->stacks[i] = (char *) malloc(STK);
#ifdef EMU_FAST_SWITCH
        // first switch into the fiber: six zeroed callee-saved registers are popped, then `ret` enters the trampoline
        // with the stack as after a call (rsp = 16n + 8); the slot above is a null return address (never used)
        uintptr_t top = ((uintptr_t) w->stacks[i] + STK) & ~(uintptr_t) 15;
        void **sp = (void **) top;
        *--sp = nullptr;
        *--sp = (void *) emu_trampoline;
        for (int k = 0; k < 6; k++) *--sp = nullptr;
        w->ctx[i] = (void *) sp;
#else
        getcontext(&w->ctx[i]);
        w->ctx[i].uc_stack.ss_sp = w->stacks[i];
        w->ctx[i].uc_stack.ss_size = STK;
        w->ctx[i].uc_link = &w->main_ctx;
        makecontext(&w->ctx[i], (void (*)()) emu_trampoline, 0);
#endif
    }
    w->cur = 0;
#ifdef EMU_FAST_SWITCH
    emu_switch(&w->main_ctx, w->ctx[0]);
#else
    swapcontext(&w->main_ctx, &w->ctx[0]);
#endif
    for (int i = 0; i < n_threads; i++) free(w->stacks[i]);
    g_warp = saved;
    free(w);
}
static inline void emu_run_warp(void (*fn)(void *), void *arg) { emu_run_cta(fn, arg, 32); }
