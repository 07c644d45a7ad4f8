// abi_comm.cuh -- halo exchange of the row-partitioned levels as ONE kernel over NVLink peer memory
// (amgb_comm_*; included by engine.cu, same translation unit).
//
// Why not NCCL for this step: a multi-colour Gauss-Seidel cycle on partitioned levels needs one exchange of a few
// hundred KB per colour wave -- ~115 per V-cycle on the 256^3 hierarchy.  As host-issued ncclAllGather calls that is
// a latency chain of 20-30 us links that cannot be captured together with the engine's launches without capturing
// NCCL ("what's weak" 4).  Here every rank owns one IPC-exported block of device memory
//     [ flags: one monotone 64-bit counter per source rank | staging: 2 x cap doubles (double-buffered) ]
// mapped by its neighbours (cudaIpcOpenMemHandle: NVLink / NVSwitch peer access), and an exchange is one launch:
//   1. push   every CTA stores its share of the packed boundary entries v[send_idx[..]] STRAIGHT INTO THE
//             DESTINATIONS' staging buffers (st.global over NVLink), buffer (k & 1) of exchange number k;
//   2. signal the last CTA to finish pushing (device-wide counter) publishes k in every neighbour's flag slot
//             (fence.sys, then st.release.sys);
//   3. wait   one thread per source spins on the local flag (ld.acquire.sys) until it reaches k;
//   4. unpack every CTA copies its share of the local staging buffer into the halo region of v.
// No host involvement, no collective call: the whole distributed cycle becomes graph-capturable.  Exchange numbers
// live on the device (`seq`), so a replayed CUDA graph keeps counting.
//
// Safety of the double buffer: an exchange signals and waits for the SAME fixed neighbour set every time (the union
// over all levels), so a neighbour can run at most one exchange ahead: it may fill staging[(k+1) & 1] while this rank
// still unpacks staging[k & 1], and cannot start exchange k+2 before this rank has signalled k+1 -- which the stream
// orders after the unpack of k.  Flags are monotone, so "flag >= k" is the whole protocol.
//
// Launch: all CTAs must be co-resident (CTAs that have pushed wait for peers while the last one signals): a small
// grid (64 CTAs by default) ordered after the previous kernel by the stream / graph; every spin traps after a bounded
// number of polls instead of hanging the device when a peer died.
#pragma once
#include <cstring>
#ifdef AMGB_EMU
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#endif

namespace amgb {

constexpr int kCommMaxWorld = 8;
constexpr int kCommThreads = 256;
constexpr long long kCommFlagBytes = 256;        // flags region at the start of the IPC block

struct ExchArgs {
    int world, rank;
    unsigned nbr_mask;                            // ranks signalled and waited for by every exchange
    const double *v_src;                          // owned part of the vector (gather source)
    const int *send_idx;                          // packed per destination rank
    double *v_halo;                               // v + n_own: [from rank 0 | from rank 1 | ...] (recv_off layout)
    long long send_begin[kCommMaxWorld + 1];      // send_idx range of destination q
    long long dst_off[kCommMaxWorld];             // where this rank's block starts in q's staging buffer
    long long recv_total;                         // doubles this rank receives (= staging entries to unpack)
    long long send_total;
    double *stage_peer[kCommMaxWorld];            // peers' staging (buffer 0); [rank] = local
    unsigned long long *flag_peer[kCommMaxWorld]; // peers' flag arrays
    double *stage_local;
    unsigned long long *flag_local;
    long long cap;
    unsigned long long *seq;                      // exchanges completed so far (device)
    unsigned *done, *fin;                         // CTA counters of the running exchange
};

__device__ __forceinline__ unsigned long long comm_ld_acquire_sys(const unsigned long long *p)
{
#ifdef AMGB_EMU
    return *(volatile const unsigned long long *)p;
#else
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void comm_st_release_sys(unsigned long long *p, unsigned long long v)
{
#ifdef AMGB_EMU
    __sync_synchronize();
    *(volatile unsigned long long *)p = v;
#else
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

__global__ void __launch_bounds__(kCommThreads) halo_exchange_kernel(const ExchArgs a)
{
    const unsigned long long k = __ldcg(a.seq) + 1ull;          // number of this exchange (same on every rank)
    const long long boff = (long long)(k & 1ull) * a.cap;
    const long long gtid = (long long)blockIdx.x * kCommThreads + threadIdx.x;
    const long long gsz = (long long)gridDim.x * kCommThreads;
    // 1. push: four independent index -> value -> remote-store chains per thread and trip (the chain is two dependent
    //    loads of ~1 us each; a 256^2-plane halo is 65 536 entries per neighbour)
    constexpr int U = 4;
    for (long long i0 = gtid; i0 < a.send_total; i0 += gsz * U) {
        int idx[U];
        double val[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long i = i0 + (long long)u * gsz;
            idx[u] = i < a.send_total ? __ldg(a.send_idx + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; u++) val[u] = idx[u] >= 0 ? a.v_src[idx[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long i = i0 + (long long)u * gsz;
            if (idx[u] < 0) continue;
            int q = 0;
#pragma unroll
            for (int t = 1; t < kCommMaxWorld; t++) q += (t < a.world && i >= a.send_begin[t]) ? 1 : 0;
            double *dst = a.stage_peer[0];
            long long off = a.dst_off[0], beg = a.send_begin[0];
#pragma unroll
            for (int t = 1; t < kCommMaxWorld; t++)
                if (q == t) { dst = a.stage_peer[t]; off = a.dst_off[t]; beg = a.send_begin[t]; }
            dst[boff + off + (i - beg)] = val[u];
        }
    }
    // 2. signal (last CTA)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(a.done, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            for (int q = 0; q < a.world; q++)
                if ((a.nbr_mask >> q) & 1u) comm_st_release_sys(a.flag_peer[q] + a.rank, k);
        }
    }
    // 3. wait
    if (threadIdx.x < (unsigned)a.world && ((a.nbr_mask >> threadIdx.x) & 1u)) {
        unsigned long long spins = 0;
        while (comm_ld_acquire_sys(a.flag_local + threadIdx.x) < k) {
#ifdef AMGB_EMU
            ::emu::external_wait();
#else
            if (++spins > (1ull << 25)) __trap();                 // a dead peer must not hang this device
#endif
        }
        __threadfence_system();
    }
    __syncthreads();
    // 4. unpack (staging is written by peers over NVLink: read it from L2, never from a stale L1 line)
    for (long long i0 = gtid; i0 < a.recv_total; i0 += gsz * U) {
        double val[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long i = i0 + (long long)u * gsz;
            val[u] = i < a.recv_total ? __ldcg(a.stage_local + boff + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long i = i0 + (long long)u * gsz;
            if (i < a.recv_total) a.v_halo[i] = val[u];
        }
    }
    // 5. the last CTA to finish closes the exchange
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(a.fin, 1u);
        if (prev == gridDim.x - 1) {
            *a.done = 0u;
            *a.fin = 0u;
            *a.seq = k;
            __threadfence();
        }
    }
}

}  // namespace amgb

struct amgb_comm {
    int device = 0, world = 1, rank = 0;
    cudaStream_t stream = nullptr;
    long long cap = 0;
    size_t bytes = 0;
    unsigned char *base = nullptr;                         // this rank's IPC block
    unsigned char *peer[amgb::kCommMaxWorld] = {};         // mapped peer blocks ([rank] = base)
    unsigned nbr_mask = 0;
    unsigned long long *seq = nullptr;
    unsigned *done = nullptr, *fin = nullptr;
    int grid = 64;
    int coop = 0;
    long long exchanges = 0;
#ifdef AMGB_EMU
    char shm_name[64] = {};
#endif
};

// handle: 64 opaque bytes for the peers (cudaIpcMemHandle_t; a shared-memory object name on the emulator)
extern "C" int amgb_comm_create(int device, int world, int rank, int64_t cap_doubles, void *stream, amgb_comm **out,
                                unsigned char *handle64)
{
    using namespace amgb;
    if (out == nullptr || handle64 == nullptr) return fail(AMGB_EINVAL, "null pointer");
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world) return fail(AMGB_EINVAL, "world must be 1..8");
    if (cap_doubles < 1) cap_doubles = 1;
    CK(cudaSetDevice(device));
    amgb_comm *c = new amgb_comm();
    c->device = device; c->world = world; c->rank = rank; c->stream = (cudaStream_t)stream;
    c->cap = (cap_doubles + 31) & ~31ll;
    c->bytes = (size_t)kCommFlagBytes + (size_t)c->cap * 2 * sizeof(double);
    memset(handle64, 0, 64);
#ifdef AMGB_EMU
    snprintf(c->shm_name, sizeof c->shm_name, "/amgb_comm_%d_%d_%lld", (int)getppid(), rank, (long long)getpid());
    shm_unlink(c->shm_name);
    const int fd = shm_open(c->shm_name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)c->bytes) != 0) { delete c; return fail(AMGB_ECUDA, "shm_open failed"); }
    c->base = (unsigned char *)mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c->base == (unsigned char *)MAP_FAILED) { delete c; return fail(AMGB_ECUDA, "mmap failed"); }
    memset(c->base, 0, c->bytes);
    ::emu::register_allocation(c->base, c->bytes);
    strncpy((char *)handle64, c->shm_name, 63);
#else
    void *p = nullptr;
    CK(cudaMalloc(&p, c->bytes));
    c->base = (unsigned char *)p;
    CK(cudaMemset(p, 0, c->bytes));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle larger than the ABI's 64 bytes");
    CK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, sizeof h);
#endif
    void *ctrl = nullptr;
    CK(cudaMalloc(&ctrl, 256));
    CK(cudaMemset(ctrl, 0, 256));
    c->seq = (unsigned long long *)ctrl;
    c->done = (unsigned *)((unsigned char *)ctrl + 64);
    c->fin = (unsigned *)((unsigned char *)ctrl + 128);
    c->peer[rank] = c->base;
    const char *cp = getenv("AMGB_COMM_COOP");
    c->coop = (cp && cp[0] == '1') ? 1 : 0;
    const char *g = getenv("AMGB_COMM_CTAS");
    if (g && atoi(g) >= 1) c->grid = std::min(atoi(g), 64);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->grid = std::max(1, std::min(c->grid, prop.multiProcessorCount));     // cooperative: every CTA resident
    CK(cudaDeviceSynchronize());
    *out = c;
    return AMGB_OK;
}

// handles: world x 64 bytes (every rank's amgb_comm_create handle, rank order); nbr_mask: ranks this one exchanges with
// at ANY level (must be symmetric across ranks).  Collective in spirit: call after all ranks created their block.
extern "C" int amgb_comm_connect(amgb_comm *c, const unsigned char *handles, uint32_t nbr_mask)
{
    using namespace amgb;
    if (c == nullptr || handles == nullptr) return fail(AMGB_EINVAL, "null pointer");
    CK(cudaSetDevice(c->device));
    c->nbr_mask = nbr_mask & ~(1u << c->rank) & ((1u << c->world) - 1u);
    for (int q = 0; q < c->world; q++) {
        if (q == c->rank || !((c->nbr_mask >> q) & 1u) || c->peer[q] != nullptr) continue;
#ifdef AMGB_EMU
        char name[64];
        memcpy(name, handles + (size_t)q * 64, 64);
        name[63] = 0;
        const int fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) return fail(AMGB_ECUDA, "shm_open of a peer block failed");
        void *p = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) return fail(AMGB_ECUDA, "mmap of a peer block failed");
        ::emu::register_allocation(p, c->bytes);
        c->peer[q] = (unsigned char *)p;
#else
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * 64, sizeof h);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[q] = (unsigned char *)p;
#endif
    }
    return AMGB_OK;
}

// One exchange of partitioned vector v (extended layout [owned (n_own) | from rank 0 | from rank 1 | ...]):
//   send_idx (device): local indices of the entries to send, packed per destination; send_off (host, world + 1);
//   peer_off (host, world): where this rank's block starts in destination q's halo region (q's recv_off[rank]);
//   recv_total: doubles this rank receives.
extern "C" int amgb_comm_exchange(amgb_comm *c, double *v, int64_t n_own, const int32_t *send_idx,
                                  const int64_t *send_off, const int64_t *peer_off, int64_t recv_total)
{
    using namespace amgb;
    if (c == nullptr || v == nullptr || send_off == nullptr || peer_off == nullptr) return fail(AMGB_EINVAL, "null pointer");
    if (c->world == 1) return AMGB_OK;
    if (recv_total > c->cap) return fail(AMGB_EINVAL, "halo larger than the communicator's staging buffer");
    ExchArgs a;
    memset(&a, 0, sizeof a);
    a.world = c->world; a.rank = c->rank; a.nbr_mask = c->nbr_mask;
    a.v_src = v; a.send_idx = send_idx; a.v_halo = v + n_own;
    for (int q = 0; q <= c->world; q++) a.send_begin[q] = send_off[q];
    for (int q = c->world + 1; q <= kCommMaxWorld; q++) a.send_begin[q] = send_off[c->world];
    a.send_total = send_off[c->world];
    a.recv_total = recv_total;
    for (int q = 0; q < c->world; q++) {
        a.dst_off[q] = peer_off[q];
        const bool mapped = c->peer[q] != nullptr;
        if (!mapped && send_off[q + 1] > send_off[q]) return fail(AMGB_ESTATE, "exchange with a rank that is not connected");
        unsigned char *pb = mapped ? c->peer[q] : c->base;
        a.stage_peer[q] = (double *)(pb + kCommFlagBytes);
        a.flag_peer[q] = (unsigned long long *)pb;
    }
    a.stage_local = (double *)(c->base + kCommFlagBytes);
    a.flag_local = (unsigned long long *)c->base;
    a.cap = c->cap; a.seq = c->seq; a.done = c->done; a.fin = c->fin;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)c->grid);
    cfg.blockDim = dim3(kCommThreads);
    cfg.stream = c->stream;
    // All CTAs must be resident at once (CTAs that have pushed wait for peers while the last one signals).  The grid is
    // at most 64 CTAs of 256 threads on a 148-SM device and the stream / graph orders it after the previous kernel, so
    // that holds by construction; AMGB_COMM_COOP=1 additionally asks the driver to verify it (cooperative launch).
    // The emulator always co-schedules the blocks of this kernel.
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
#ifdef AMGB_EMU
    cfg.numAttrs = 1;
#else
    cfg.numAttrs = c->coop ? 1 : 0;
#endif
    CK(cudaLaunchKernelEx(&cfg, halo_exchange_kernel, a));
    c->exchanges++;
    return AMGB_OK;
}

extern "C" void amgb_comm_destroy(amgb_comm *c)
{
    if (c == nullptr) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; q++) {
        if (q == c->rank || c->peer[q] == nullptr) continue;
#ifdef AMGB_EMU
        ::emu::unregister_allocation(c->peer[q]);
        munmap(c->peer[q], c->bytes);
#else
        cudaIpcCloseMemHandle(c->peer[q]);
#endif
    }
#ifdef AMGB_EMU
    if (c->base) { ::emu::unregister_allocation(c->base); munmap(c->base, c->bytes); shm_unlink(c->shm_name); }
#else
    if (c->base) cudaFree(c->base);
#endif
    if (c->seq) cudaFree(c->seq);
    delete c;
}
