// cuda_runtime.h (tests/emu) -- TEST INFRASTRUCTURE, never part of the product.
//
// A host-side stand-in for the CUDA runtime + device intrinsics, just large enough to compile
// pyamg_b200/csrc/engine.cu with g++ (-DAMGB_EMU) and EXECUTE THE SAME KERNEL SOURCES on the CPU of the
// GPU-less build container.  Purpose: logic validation of the engine (tile descriptors, staging offsets,
// wave schedules, graph capture of the cycle, pointer ping-pong, epilogues, cluster programs) between GPU
// sessions.  It says nothing about performance, memory ordering or hardware limits, and it is not a fallback:
// the product (pyamg_b200/_engine.py) only ever loads libpyamg_b200.so; this library is loaded by
// tests/conftest.py when AMGB_TEST_EMU=1, and reports zero devices unless that variable is set.
//
// Execution model
//   * every CUDA thread of a launch is a fiber (own 64 KB stack, hand-written x86-64 context switch); all
//     fibers of a thread block -- of a whole cluster for cluster launches -- are scheduled round-robin on the
//     calling OS thread and switch only at synchronisation points (__syncthreads, __syncwarp, shuffles,
//     mbarrier waits, cluster barriers), so a run is deterministic;
//   * lanes of a warp therefore do NOT run in lockstep between synchronisation points: code that relies on
//     implicit warp synchrony produces wrong answers here (that is a feature);
//   * the TMA (cp.async.bulk + mbarrier complete_tx) is a memcpy that checks what the hardware requires
//     (16-byte aligned source, destination and size; source inside one device allocation; destination
//     inside the CTA's dynamic shared memory).  AMGB_EMU_TMA=lazy defers every copy to the first wait on
//     its barrier (latest legal completion: a read before the wait sees stale data); the default performs
//     it at issue (earliest completion: a refill of a stage still being read is visible);
//   * device memory is host memory poisoned with 0xFF (doubles read as NaN, ints as -1): anything computed
//     from uninitialised device memory shows up in the parity tests;
//   * stream capture records closures with by-value arguments, graph launch replays them (what baking
//     pointers into a CUDA graph means); everything else runs synchronously in stream order;
//   * AMGB_EMU_ORDER=reverse (or random:<seed>) visits warps and lanes in the opposite (a pseudo-random) order: results that change with it depend on
//     the interleaving between synchronisation points, i.e. the kernel has a race (or an order-dependent reduction);
//   * a launch in which no fiber can make progress is reported as cudaErrorLaunchFailure ("deadlock").
#pragma once
#ifndef AMGB_EMU
#error "tests/emu/cuda_runtime.h is only for the -DAMGB_EMU host build of the engine"
#endif

#include <sched.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------
// language extensions
// ------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static      /* one OS thread runs one block / cluster at a time; see emu::dyn_smem for extern arrays */

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// built-in variables: plain globals, rewritten by the scheduler on every switch into a fiber
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

// ------------------------------------------------------------------------------------------------
// runtime types
// ------------------------------------------------------------------------------------------------
enum cudaError_t {
    cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100,
    cudaErrorInvalidDevice = 101, cudaErrorLaunchFailure = 719, cudaErrorStreamCaptureInvalidated = 901
};
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1,
                             cudaStreamCaptureModeRelaxed = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
                         cudaFuncAttributeNonPortableClusterSizeAllowed = 11 };
enum cudaLaunchAttributeID { cudaLaunchAttributeCooperative = 2, cudaLaunchAttributeClusterDimension = 4,
                             cudaLaunchAttributeProgrammaticStreamSerialization = 6 };
constexpr unsigned cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0;

struct cudaLaunchAttributeValue {
    struct { unsigned x, y, z; } clusterDim;
    int programmaticStreamSerializationAllowed;
    int cooperative;
};
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };

namespace emu { struct Stream; struct Graph; }
typedef emu::Stream *cudaStream_t;
typedef emu::Graph *cudaGraph_t;
typedef emu::Graph *cudaGraphExec_t;
struct cudaEvent_st { double t_ms; };
typedef cudaEvent_st *cudaEvent_t;

struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes = 0;
    cudaStream_t stream = nullptr;
    cudaLaunchAttribute *attrs = nullptr;
    unsigned numAttrs = 0;
};
struct cudaDeviceProp {
    char name[64];
    int multiProcessorCount;
    size_t sharedMemPerMultiprocessor, sharedMemPerBlockOptin, totalGlobalMem;
    int major, minor;
};

// ------------------------------------------------------------------------------------------------
// the emulator
// ------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch,@function
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
    .size emu_switch,.-emu_switch
)");

namespace emu {

struct Stream { bool capturing = false; Graph *cap = nullptr; };
struct Graph { std::vector<std::function<void()>> nodes; };

struct Warp { int alive = 0, arrived = 0; unsigned gen = 0; uint64_t slot[32]; int pred[32]; };
struct Cluster;
struct Block {
    uint3 bid;
    int nthreads = 0, alive = 0, arrived = 0;
    unsigned gen = 0;
    std::vector<Warp> warps;
    unsigned char *smem = nullptr;
    size_t smem_bytes = 0;
    int rank = 0;
    Cluster *cl = nullptr;
};
struct Cluster { int nctas = 1; std::vector<Block> blocks; int alive = 0, arrived = 0; unsigned gen = 0; };
struct Fiber { void *sp = nullptr; uint3 tid; int lane = 0; Warp *w = nullptr; Block *blk = nullptr; bool done = false; };

struct KernelCall { virtual void run() = 0; virtual ~KernelCall() {} };

struct State {
    Fiber *cur = nullptr;
    void *sched_sp = nullptr;
    unsigned long long progress = 0;
    KernelCall *call = nullptr;
    cudaError_t last_error = cudaSuccess;
    std::map<uintptr_t, size_t> allocs;       // device allocations: base -> bytes
    std::vector<unsigned char *> stacks;      // fiber stacks (reused across launches)
    bool tma_lazy = false;
    bool reverse = false;                     // AMGB_EMU_ORDER=reverse
    bool shuffle = false;                     // AMGB_EMU_ORDER=random:<seed>
    unsigned long long rng = 1;
    size_t rot = 0;
    struct PendingCopy { unsigned long long *bar; void *dst; const void *src; unsigned bytes; };
    std::vector<PendingCopy> pending;
    long long launches = 0, fibers_run = 0;
    int device_set = 0;
    bool ext_wait = false;       // some fiber polls memory another process writes (emu::external_wait)
};
inline State g;

constexpr size_t kStackBytes = 64 * 1024;

inline void fatal(const char *msg)
{
    fprintf(stderr, "[cuda-emu] FATAL: %s\n", msg);
    fflush(stderr);
    abort();
}

inline void yield() { emu_switch(&g.cur->sp, g.sched_sp); }

inline void release_if_complete_on_exit(Fiber *f)
{
    Warp &w = *f->w;
    if (w.alive > 0 && w.arrived >= w.alive) { w.arrived = 0; w.gen++; }
    Block &b = *f->blk;
    if (b.alive > 0 && b.arrived >= b.alive) { b.arrived = 0; b.gen++; }
    Cluster &c = *b.cl;
    if (c.alive > 0 && c.arrived >= c.alive) { c.arrived = 0; c.gen++; }
}

inline void fiber_entry()
{
    Fiber *f = g.cur;
    g.call->run();
    f->done = true;
    f->w->alive--;
    f->blk->alive--;
    f->blk->cl->alive--;
    release_if_complete_on_exit(f);
    g.progress++;
    for (;;) yield();          // never scheduled again
}

inline unsigned char *get_stack(size_t i)
{
    while (g.stacks.size() <= i) {
        void *p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) fatal("cannot map a fiber stack");
        g.stacks.push_back((unsigned char *)p);
    }
    return g.stacks[i];
}

inline void prepare_fiber(Fiber &f, size_t stack_index)
{
    unsigned char *top = get_stack(stack_index) + kStackBytes;
    void **sp = (void **)top;
    *--sp = nullptr;                       // fake return address of fiber_entry (never returns)
    *--sp = (void *)&fiber_entry;          // `ret` target of the first switch
    for (int k = 0; k < 6; k++) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = sp;
}

// run one cluster (nctas consecutive blocks) to completion; false on deadlock
inline bool run_cluster(dim3 grid, dim3 block, size_t smem, unsigned first_block, int nctas)
{
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    Cluster cl;
    cl.nctas = nctas;
    cl.blocks.resize((size_t)nctas);
    cl.alive = nctas * nthreads;
    std::vector<Fiber> fibers((size_t)nctas * nthreads);
    std::vector<std::unique_ptr<unsigned char[]>> smems;
    for (int c = 0; c < nctas; c++) {
        Block &b = cl.blocks[(size_t)c];
        const unsigned lin = first_block + (unsigned)c;
        b.bid.x = lin % grid.x; b.bid.y = (lin / grid.x) % grid.y; b.bid.z = lin / (grid.x * grid.y);
        b.nthreads = b.alive = nthreads;
        b.warps.resize((size_t)nwarps);
        b.rank = c;
        b.cl = &cl;
        b.smem_bytes = smem;
        smems.emplace_back(new unsigned char[smem + 64]);
        b.smem = (unsigned char *)(((uintptr_t)smems.back().get() + 63) & ~(uintptr_t)63);
        memset(b.smem, 0xFF, smem);
        for (int t = 0; t < nthreads; t++) {
            Fiber &f = fibers[(size_t)c * nthreads + t];
            f.tid.x = (unsigned)t % block.x; f.tid.y = ((unsigned)t / block.x) % block.y; f.tid.z = (unsigned)t / (block.x * block.y);
            f.lane = t & 31;
            f.w = &b.warps[(size_t)(t >> 5)];
            f.w->alive++;
            f.blk = &b;
            prepare_fiber(f, (size_t)c * nthreads + t);
        }
    }
    gridDim = grid;
    blockDim = block;
    size_t remaining = fibers.size();
    g.fibers_run += (long long)remaining;
    // warp-granular round-robin: the lanes of one warp are cycled until none of them can move (all wait on a
    // block / cluster barrier or an mbarrier, or are done), so a warp-level synchronisation costs 32 context
    // switches, not one pass over every thread of the cluster
    const size_t nfib = fibers.size();
    double ext_since = 0.0;
    auto now_s = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    while (remaining > 0) {
        const unsigned long long before = g.progress;
        g.ext_wait = false;
        remaining = 0;
        const size_t nwarp_all = (nfib + 31) / 32;
        for (size_t wi = 0; wi < nwarp_all; wi++) {
            // AMGB_EMU_ORDER=reverse: warps and lanes are visited in the opposite order -- a correct kernel (no
            // dependence on the interleaving between synchronisation points) gives bit-identical results
            size_t wsel = g.reverse ? (nwarp_all - 1 - wi) : wi;
            if (g.shuffle) {                       // AMGB_EMU_ORDER=random:<seed>: a rotated warp order every pass,
                if (wi == 0) {                     // a rotated + strided lane order every warp visit
                    g.rng = g.rng * 6364136223846793005ull + 1442695040888963407ull;
                    g.rot = (size_t)(g.rng >> 33);
                }
                wsel = (wi + g.rot) % nwarp_all;
            }
            const size_t w0 = wsel * 32;
            const size_t w1 = std::min(nfib, w0 + 32);      // warps: blocks are multiples of 32 threads (checked in execute)
            size_t left;
            unsigned long long p;
            do {
                p = g.progress;
                left = 0;
                for (size_t ii = w0; ii < w1; ii++) {
                    size_t i = g.reverse ? (w1 - 1 - (ii - w0)) : ii;
                    if (g.shuffle && w1 - w0 == 32) i = w0 + ((ii - w0) * 13 + g.rot + wsel) % 32;   // 13 is coprime to 32
                    Fiber &f = fibers[i];
                    if (f.done) continue;
                    g.cur = &f;
                    threadIdx = f.tid;
                    blockIdx = f.blk->bid;
                    emu_switch(&g.sched_sp, f.sp);
                    if (!f.done) left++;
                }
            } while (left > 0 && g.progress != p);
            remaining += left;
        }
        if (remaining > 0 && g.progress == before && g.ext_wait) {
            // every runnable fiber waits for another PROCESS (peer flags in shared memory): not a deadlock of this
            // kernel -- give the processor away and look again (bounded: a lost peer fails the launch after 120 s)
            g.ext_wait = false;
            if (ext_since == 0.0) ext_since = now_s();
            if (now_s() - ext_since > 120.0) {
                fprintf(stderr, "[cuda-emu] timeout: waited 120 s for another process\n");
                g.cur = nullptr;
                return false;
            }
            sched_yield();
            continue;
        }
        if (g.progress != before) ext_since = 0.0;
        if (remaining > 0 && g.progress == before) {
            fprintf(stderr, "[cuda-emu] deadlock: %zu threads wait on a barrier nobody will release "
                            "(block %u of %u, %d threads per block)\n", remaining, first_block, grid.x * grid.y * grid.z, nthreads);
            g.cur = nullptr;
            return false;
        }
    }
    g.cur = nullptr;
    return true;
}

template <class... P>
struct KernelCallT : KernelCall {
    void (*fn)(P...);
    std::tuple<std::decay_t<P>...> args;
    template <class... A>
    KernelCallT(void (*f)(P...), A &&...a) : fn(f), args(std::forward<A>(a)...) {}
    void run() override { std::apply(fn, args); }      // by-value parameters are copied per thread, as on the device
};

inline cudaError_t execute(dim3 grid, dim3 block, size_t smem, int cluster, KernelCall *call)
{
    const unsigned nblocks = grid.x * grid.y * grid.z;
    if (nblocks == 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024) return g.last_error = cudaErrorInvalidValue;
    if (cluster < 1 || nblocks % (unsigned)cluster) return g.last_error = cudaErrorInvalidValue;
    if (smem > 227 * 1024) return g.last_error = cudaErrorInvalidValue;
    if (cluster > 1 && (block.x * block.y * block.z) % 32u) return g.last_error = cudaErrorInvalidValue;
    KernelCall *outer = g.call;
    g.call = call;
    g.launches++;
    bool ok = true;
    for (unsigned b0 = 0; b0 < nblocks && ok; b0 += (unsigned)cluster) ok = run_cluster(grid, block, smem, b0, cluster);
    g.call = outer;
    g.pending.clear();
    return ok ? cudaSuccess : (g.last_error = cudaErrorLaunchFailure);
}

// dynamic shared memory above 48 KB needs cudaFuncSetAttribute(kernel, MaxDynamicSharedMemorySize, >= bytes) first,
// as on the device (static shared memory is small in every kernel here: checked with cuobjdump --dump-resource-usage)
inline std::unordered_map<const void *, size_t> &smem_optin()
{
    static std::unordered_map<const void *, size_t> m;
    return m;
}
inline bool smem_allowed(const void *kernel, size_t smem)
{
    if (smem <= 48 * 1024) return true;
    auto it = smem_optin().find(kernel);
    if (it != smem_optin().end() && it->second >= smem) return true;
    fprintf(stderr, "[cuda-emu] launch with %zu bytes of dynamic shared memory without the opt-in attribute\n", smem);
    return false;
}
inline bool grid_allowed(dim3 grid)
{
    return grid.x <= 2147483647u && grid.y <= 65535u && grid.z <= 65535u;
}

template <class... P, class... A>
inline cudaError_t launch(dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster, void (*kernel)(P...), A &&...args)
{
    if (!smem_allowed(reinterpret_cast<const void *>(kernel), smem) || !grid_allowed(grid)) return g.last_error = cudaErrorInvalidValue;
    auto call = std::make_shared<KernelCallT<P...>>(kernel, std::forward<A>(args)...);
    if (s != nullptr && s->capturing) {
        s->cap->nodes.push_back([=]() { execute(grid, block, smem, cluster, call.get()); });
        return cudaSuccess;
    }
    return execute(grid, block, smem, cluster, call.get());
}

struct LaunchCfg {
    dim3 grid, block;
    size_t smem;
    cudaStream_t stream;
    LaunchCfg(dim3 g_, dim3 b_, size_t s_ = 0, cudaStream_t st = nullptr) : grid(g_), block(b_), smem(s_), stream(st) {}
};
// kernel<<<grid, block, smem, stream>>>(args...) is rewritten to this by tests/emu/build_emu.py
template <class... P, class... A>
inline void launch_kernel(const LaunchCfg &c, void (*kernel)(P...), A &&...args)
{
    launch(c.grid, c.block, c.smem, c.stream, 1, kernel, std::forward<A>(args)...);
}

// ---- device-side services used by the intrinsics below -----------------------------------------
inline void require_device_code(const char *what)
{
    if (g.cur == nullptr) { fprintf(stderr, "[cuda-emu] %s outside a kernel\n", what); abort(); }
}
inline unsigned char *dyn_smem() { require_device_code("dynamic shared memory"); return g.cur->blk->smem; }

inline void warp_barrier()
{
    Warp &w = *g.cur->w;
    if (++w.arrived >= w.alive) { w.arrived = 0; w.gen++; g.progress++; }
    else { const unsigned gen = w.gen; while (w.gen == gen) yield(); }
}
inline void block_barrier()
{
    Block &b = *g.cur->blk;
    if (++b.arrived >= b.alive) { b.arrived = 0; b.gen++; g.progress++; }
    else { const unsigned gen = b.gen; while (b.gen == gen) yield(); }
}
inline void cluster_barrier()
{
    Cluster &c = *g.cur->blk->cl;
    if (++c.arrived >= c.alive) { c.arrived = 0; c.gen++; g.progress++; }
    else { const unsigned gen = c.gen; while (c.gen == gen) yield(); }
}
// grid barrier of a cooperatively launched kernel (csrc/grid_kernel.cuh): arrive on a monotone counter in global
// memory, then wait until it reaches `target`; the arrival counts as progress for the deadlock detector
inline void grid_arrive_wait(unsigned long long *count, unsigned long long target)
{
    require_device_code("grid barrier");
    (*count)++;
    g.progress++;
    while (*count < target) yield();
}
// a spin on memory another PROCESS writes (csrc/abi_comm.cuh: peers' flags in a shared-memory block): give the
// processor away, count it as progress (the deadlock detector only knows this process), let the other fibers run
inline void external_wait()
{
    require_device_code("external wait");
    g.ext_wait = true;           // not progress of THIS process: the scheduler moves on to the other warps
    yield();
}
inline void register_allocation(void *p, size_t bytes) { g.allocs[(uintptr_t)p] = bytes; }
inline void unregister_allocation(void *p) { g.allocs.erase((uintptr_t)p); }
inline unsigned cluster_ctarank() { return (unsigned)g.cur->blk->rank; }
inline unsigned cluster_nctarank() { return (unsigned)g.cur->blk->cl->nctas; }

// distributed shared memory: (cta rank << 24) | byte offset into that CTA's dynamic shared memory
inline unsigned dsmem_addr(const void *local, unsigned cta)
{
    Block &b = *g.cur->blk;
    const ptrdiff_t off = (const unsigned char *)local - b.smem;
    if (off < 0 || (size_t)off >= b.smem_bytes) fatal("mapa: address outside the CTA's dynamic shared memory");
    if (cta >= (unsigned)b.cl->nctas) fatal("mapa: CTA rank outside the cluster");
    return (cta << 24) | (unsigned)off;
}
inline double *dsmem_ptr(unsigned addr)
{
    Cluster &c = *g.cur->blk->cl;
    Block &b = c.blocks[(size_t)(addr >> 24)];
    const unsigned off = addr & 0xFFFFFFu;
    if (off + sizeof(double) > b.smem_bytes) fatal("ld/st.shared::cluster outside the target CTA's shared memory");
    return (double *)(b.smem + off);
}

// mbarrier in 8 bytes of shared memory
struct Mbar { uint8_t phase, magic, pending, count; int32_t tx; };
static_assert(sizeof(Mbar) == 8, "mbarrier emulation must fit the 64-bit barrier word");
inline Mbar *mbar_of(unsigned long long *bar)
{
    Mbar *m = (Mbar *)bar;
    if (m->magic != 0xA5) fatal("mbarrier used before mbarrier.init");
    return m;
}
inline void mbar_check(Mbar *m)
{
    if (m->pending == 0 && m->tx == 0) { m->phase ^= 1; m->pending = m->count; g.progress++; }
}
inline void mbar_init(unsigned long long *bar, unsigned count)
{
    Mbar m;
    m.phase = 0; m.magic = 0xA5; m.pending = m.count = (uint8_t)count; m.tx = 0;
    memcpy(bar, &m, sizeof m);
}
inline void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    Mbar *m = mbar_of(bar);
    if (m->pending == 0) fatal("mbarrier.arrive on a barrier with no pending arrivals");
    m->tx += (int32_t)bytes;
    m->pending--;
    mbar_check(m);
}
inline void check_global_range(const void *p, size_t bytes, const char *what)
{
    const uintptr_t a = (uintptr_t)p;
    auto it = g.allocs.upper_bound(a);
    if (it != g.allocs.begin()) {
        --it;
        if (a >= it->first && a + bytes <= it->first + it->second) return;
    }
    fprintf(stderr, "[cuda-emu] %s: [%p, +%zu) is not inside one device allocation\n", what, p, bytes);
    abort();
}
inline void do_copy(unsigned long long *bar, void *dst, const void *src, unsigned bytes)
{
    memcpy(dst, src, bytes);
    Mbar *m = mbar_of(bar);
    m->tx -= (int32_t)bytes;
    mbar_check(m);
}
inline void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    if (bytes == 0 || (bytes & 15u)) fatal("cp.async.bulk: size must be a non-zero multiple of 16 bytes");
    if (((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u)) fatal("cp.async.bulk: source and destination must be 16-byte aligned");
    check_global_range(src, bytes, "cp.async.bulk source");
    Block &b = *g.cur->blk;
    if ((unsigned char *)dst < b.smem || (unsigned char *)dst + bytes > b.smem + b.smem_bytes)
        fatal("cp.async.bulk: destination outside the CTA's dynamic shared memory");
    if (g.tma_lazy) g.pending.push_back({bar, dst, src, bytes});
    else do_copy(bar, dst, src, bytes);
}
inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    if (g.tma_lazy) {
        for (size_t i = 0; i < g.pending.size();) {
            if (g.pending[i].bar == bar) {
                const State::PendingCopy pc = g.pending[i];
                g.pending.erase(g.pending.begin() + (ptrdiff_t)i);
                do_copy(pc.bar, pc.dst, pc.src, pc.bytes);
            } else {
                i++;
            }
        }
    }
    while (mbar_of(bar)->phase == (uint8_t)(parity & 1u)) yield();
}

template <class T>
inline T shuffle_from(T v, int src, bool in_range)
{
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    Fiber *f = g.cur;
    Warp &w = *f->w;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[f->lane] = bits;
    w.pred[f->lane] = 1;
    warp_barrier();
    T r = v;
    if (in_range && src >= 0 && src < 32) memcpy(&r, &w.slot[src], sizeof(T));
    warp_barrier();
    return r;
}

}  // namespace emu

// ------------------------------------------------------------------------------------------------
// device intrinsics
// ------------------------------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32)
{
    const int lane = emu::g.cur->lane, src = lane ^ lane_mask;
    return emu::shuffle_from(v, src, (src & ~(width - 1)) == (lane & ~(width - 1)));
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane, int width = 32)
{
    const int lane = emu::g.cur->lane;
    return emu::shuffle_from(v, (lane & ~(width - 1)) | (src_lane & (width - 1)), true);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
    const int lane = emu::g.cur->lane;
    return emu::shuffle_from(v, lane + (int)delta, (lane & (width - 1)) + (int)delta < width);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
    const int lane = emu::g.cur->lane;
    return emu::shuffle_from(v, lane - (int)delta, (lane & (width - 1)) >= (int)delta);
}
inline unsigned __ballot_sync(unsigned, int pred)
{
    emu::Fiber *f = emu::g.cur;
    emu::Warp &w = *f->w;
    w.slot[f->lane] = pred ? 1u : 0u;
    emu::warp_barrier();
    unsigned r = 0;
    const int base = (int)(f->tid.x & ~31u);
    for (int l = 0; l < 32 && base + l < f->blk->nthreads; l++) r |= (w.slot[l] ? 1u : 0u) << l;
    emu::warp_barrier();
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred)
{
    const int n = std::min(32, emu::g.cur->blk->nthreads - (int)(emu::g.cur->tid.x & ~31u));
    return __ballot_sync(m, pred) == (n == 32 ? 0xffffffffu : ((1u << n) - 1u));
}
inline double __dmul_rn(double a, double b) { return a * b; }      // the host build has no FMA contraction (-ffp-contract=off)
inline double __dadd_rn(double a, double b) { return a + b; }
inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() { __sync_synchronize(); }

// ------------------------------------------------------------------------------------------------
// runtime API (the subset the engine uses)
// ------------------------------------------------------------------------------------------------
inline const char *cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument (emulated)";
    case cudaErrorMemoryAllocation: return "out of memory (emulated)";
    case cudaErrorNoDevice: return "no CUDA-capable device is detected (emulator: AMGB_TEST_EMU is not set)";
    case cudaErrorInvalidDevice: return "invalid device ordinal (emulated)";
    case cudaErrorLaunchFailure: return "unspecified launch failure (emulated kernel deadlocked)";
    default: return "unknown error (emulated)";
    }
}
inline cudaError_t cudaGetLastError() { cudaError_t e = emu::g.last_error; emu::g.last_error = cudaSuccess; return e; }
inline cudaError_t cudaGetDeviceCount(int *n)
{
    const char *t = getenv("AMGB_TEST_EMU");
    if (!(t && t[0] == '1')) { *n = 0; return cudaErrorNoDevice; }     // never a silent CPU path
    const char *lz = getenv("AMGB_EMU_TMA");
    emu::g.tma_lazy = lz && strcmp(lz, "lazy") == 0;
    const char *ord = getenv("AMGB_EMU_ORDER");
    emu::g.reverse = ord && strcmp(ord, "reverse") == 0;
    emu::g.shuffle = ord && strncmp(ord, "random", 6) == 0;
    if (emu::g.shuffle && ord[6] == ':') emu::g.rng = strtoull(ord + 7, nullptr, 10) * 2654435761ull + 1;
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "cuda-emu (host fibers)");
    const char *s = getenv("AMGB_EMU_SMS");
    p->multiProcessorCount = (s && atoi(s) > 0) ? atoi(s) : 3;       // small persistent grids: tests stay fast
    p->sharedMemPerMultiprocessor = 233472;
    p->sharedMemPerBlockOptin = 232448;
    p->totalGlobalMem = (size_t)8 << 30;
    p->major = 10; p->minor = 0;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }

inline cudaError_t cudaMalloc(void **p, size_t bytes)
{
    void *q = nullptr;
    if (posix_memalign(&q, 256, std::max<size_t>(bytes, 256)) != 0) return emu::g.last_error = cudaErrorMemoryAllocation;
    memset(q, 0xFF, std::max<size_t>(bytes, 256));
    emu::g.allocs[(uintptr_t)q] = bytes;
    *p = q;
    return cudaSuccess;
}
template <class T> inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
// memory the "device" did not allocate itself (torch CPU tensors standing in for CUDA tensors in the distributed
// emulator test) is announced here so that the TMA source check knows it
extern "C" __attribute__((visibility("default"))) void amgb_emu_register_allocation(void *p, size_t bytes)
{
    emu::g.allocs[(uintptr_t)p] = bytes;
}
extern "C" __attribute__((visibility("default"))) void amgb_emu_unregister_allocation(void *p)
{
    emu::g.allocs.erase((uintptr_t)p);
}
inline cudaError_t cudaFree(void *p)
{
    if (p == nullptr) return cudaSuccess;
    if (emu::g.allocs.erase((uintptr_t)p) != 1) return emu::g.last_error = cudaErrorInvalidValue;
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned)
{
    *p = malloc(std::max<size_t>(bytes, 1));
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind)
{
    if (bytes) memmove(dst, src, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t s = nullptr)
{
    if (s != nullptr && s->capturing) { s->cap->nodes.push_back([=]() { if (bytes) memmove(dst, src, bytes); }); return cudaSuccess; }
    if (bytes) memmove(dst, src, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t s = nullptr)
{
    if (s != nullptr && s->capturing) { s->cap->nodes.push_back([=]() { memset(p, v, bytes); }); return cudaSuccess; }
    memset(p, v, bytes);
    return cudaSuccess;
}

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu::Stream(); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emu::Stream(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t s)
{
    return (s != nullptr && s->capturing) ? cudaErrorStreamCaptureInvalidated : cudaSuccess;   // illegal during capture
}
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
inline cudaError_t cudaStreamIsCapturing(cudaStream_t s, cudaStreamCaptureStatus *st)
{
    *st = (s != nullptr && s->capturing) ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode)
{
    if (s == nullptr || s->capturing) return cudaErrorInvalidValue;
    s->capturing = true;
    s->cap = new emu::Graph();
    return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g)
{
    if (s == nullptr || !s->capturing) { *g = nullptr; return cudaErrorInvalidValue; }
    s->capturing = false;
    *g = s->cap;
    s->cap = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long = 0)
{
    *e = new emu::Graph(*g);
    return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s)
{
    if (s != nullptr && s->capturing) return cudaErrorInvalidValue;
    for (auto &n : e->nodes) n();
    cudaError_t err = emu::g.last_error;
    emu::g.last_error = cudaSuccess;
    return err;
}

inline double emu_now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cudaEvent_st{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t_ms = emu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = (float)std::max(b->t_ms - a->t_ms, 1e-6);
    return cudaSuccess;
}

template <class F> inline cudaError_t cudaFuncSetAttribute(F *f, cudaFuncAttribute a, int v)
{
    if (a == cudaFuncAttributeMaxDynamicSharedMemorySize) {
        if (v < 0 || v > 227 * 1024) return emu::g.last_error = cudaErrorInvalidValue;
        emu::smem_optin()[reinterpret_cast<const void *>(f)] = (size_t)v;
    }
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F *, int threads, size_t smem)
{
    *n = (threads <= 1024 && smem <= 227 * 1024) ? std::max(1, 2048 / std::max(threads, 1)) : 0;
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaOccupancyMaxActiveClusters(int *n, F *, const cudaLaunchConfig_t *cfg)
{
    const size_t per_sm = 227 * 1024;
    *n = (cfg->dynamicSmemBytes <= per_sm) ? 8 : 0;
    return cudaSuccess;
}
template <class... P, class... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*kernel)(P...), A &&...args)
{
    int cluster = 1;
    for (unsigned k = 0; k < cfg->numAttrs; k++) {
        if (cfg->attrs[k].id == cudaLaunchAttributeClusterDimension) cluster = (int)cfg->attrs[k].val.clusterDim.x;
        // a cooperative grid: every block is resident at once (the device guarantees it or refuses the launch; one
        // block per SM is the rule the engine follows) -- scheduled here like one big cluster
        if (cfg->attrs[k].id == cudaLaunchAttributeCooperative && cfg->attrs[k].val.cooperative) {
            cudaDeviceProp pr;
            cudaGetDeviceProperties(&pr, 0);
            const unsigned nb = cfg->gridDim.x * cfg->gridDim.y * cfg->gridDim.z;
            if ((int)nb > pr.multiProcessorCount) return emu::g.last_error = cudaErrorInvalidValue;
            cluster = (int)nb;
        }
    }
    return emu::launch(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, cluster, kernel, std::forward<A>(args)...);
}
