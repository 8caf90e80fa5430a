// xchg.cu -- multi-sheet merge across GPUs over NVLink peer memory, behind the C ABI (ib_xchg_*).
//
// Reference: every ice sheet contributes sparse (index, weight = wM, vals = M_unscaled x) entries on the A / E grid
// (IceCoupler.cpp:445-457); the GCM root gathers them with MPI (modele/GCMCoupler_ModelE.cpp:761-793) and
// VectorMultivec::to_dense_scale / to_dense (multivec.cpp:35-81) sums them and divides by the summed weights.
// Here one rank = one process = one GPU = one sheet, results stay in HBM, and only the cells that more than one sheet
// touches travel: a rank PUBLISHES its row of every shared cell by storing it straight into the inbox of the cell's other
// owners (their memory is mapped with CUDA IPC; posted stores over NVLink), and GATHERS, for each shared cell it holds,
// its own row and the rows in its inbox, added in ascending rank order (every rank computes the same bits), times
// 1 / (summed weights).  No NCCL call, no host round trip: two small kernels per step, capturable in a CUDA graph,
// 4 KB per (cell, other owner) at 512 fields instead of an all-reduce of every shared row through every rank.
//
// Protocol (PUSH: data and flags travel as posted stores over NVLink; every wait polls LOCAL memory, so no rank ever
// pays a load round trip over the link).  Rank r's region holds, for every peer q, an inbox [4 slots][shared rows] and
// two counters:
//   ready[q] = t : q has delivered its rows of steps < t into r's inbox        (stored remotely by q's publish kernel)
//   done[q]  = t : q has finished gathering steps < t out of ITS inbox, i.e. r may overwrite q's inbox slot again
//                                                                               (stored remotely by q's gather kernel)
//   publish(step s): wait for done[q] >= s - 3 of every peer q that shares a cell (inbox slot s % 4 was last read in
//   step s - 4: with two slots the ring publish(s) -> peers' gather(s) -> publish(s + 2) set the pace at 8 GPUs, 190 us per
//   two steps), store this rank's row of every shared cell into the inbox of every other owner, fence (system scope),
//   then set ready[me] = s + 1 in those peers;
//   gather(step s): wait for ready[q] >= s + 1 of those peers, add own row and inbox rows in ascending rank order, scale;
//   then set done[me] = s + 1 in the peers.
// Each kernel keeps its own step counter in device memory (advanced by its last block), so a captured graph replays and
// publish(s + 1) may run -- on another stream -- while gather(s) is still adding: they touch different inbox parities and
// different result buffers, and a publish starved by a concurrent SpMM then no longer delays the gathers.  Waits
// give up after XCHG_TIMEOUT_NS (a dead peer must not hang the GPU); ib_xchg_status() reports it.
//
// Spinning kernels of DIFFERENT ranks must run on different GPUs (nothing guarantees co-scheduling on one device).
// On one GPU (tests), ranks are driven one after the other with ib_xchg_publish_dev / ib_xchg_gather_dev, so every
// flag is already set when it is polled.
#include <algorithm>
#include <cstring>

#include "types.cuh"

namespace ib {

constexpr unsigned long long XCHG_TIMEOUT_NS = 4000000000ull;
constexpr int XCHG_MAX_RANKS = 64;
constexpr int XCHG_THREADS = 128;
constexpr int XCHG_DEPTH = 4;          // inbox slots per peer: publish(s + XCHG_DEPTH) is what has to wait for the peers' gather(s)

struct XHdr {
    unsigned long long ready[XCHG_MAX_RANKS], done[XCHG_MAX_RANKS];
};
static_assert(sizeof(XHdr) == 1024, "header: two counters per peer");

struct XPart {
    int64_t nrow = 0, nshared = 0, off = 0;      // off: first row of the part in the exchange buffer
    int scaled = 0;           // rows of Y arrive already divided by wlocal (a scaled matrix): only shared rows are touched
    DBuf<int32_t> loc;        // [nshared] local dense row or -1
    DBuf<int32_t> slot;       // [nrow] shared slot of a local row or -1
    DBuf<uint64_t> owners;    // [nshared] bit r: rank r holds the cell
    DBuf<double> wtot, wlocal;
};

struct Xchg {
    int rank = 0, nranks = 1, device = 0;
    int64_t total = 0, ldb = 0;
    std::vector<XPart> parts;
    void *region = nullptr;                 // XHdr + inbox[XCHG_DEPTH][nranks][total][ldb], cudaMalloc (IPC-exportable)
    size_t region_bytes = 0;
    void *peer[XCHG_MAX_RANKS] = {};        // mapped regions (peer[rank] = region)
    bool ipc_opened[XCHG_MAX_RANKS] = {};
    uint64_t reads_mask = 0, readers_mask = 0;     // ranks whose rows this rank gathers / that gather this rank's rows
    DBuf<unsigned long long> step;          // [2] steps gathered, steps published
    DBuf<unsigned> counters;                // [2] last-block counters of the two kernels
    DBuf<int> err;                          // [1] sticky: 1 = a wait timed out
    DBuf<unsigned long long> d_peer;        // [nranks] device copy of the mapped region addresses
    cudaStream_t stream = nullptr;
    bool connected = false;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double2 ld_peer(const double *p)
{
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long xtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#ifdef IB_STEP_TRACE        // tools/: per-step time stamps of the two kernels (block 0), ring of 4096 steps
__device__ unsigned long long g_x_steps[4096 * 8];       // publish: start, waited, stored, flags set; gather: start, waited, end
#define XTR(s, slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_x_steps[((s) & 4095ull) * 8 + (slot)] = xtime(); } while (0)
#define XTR_LAST(s, slot) do { g_x_steps[((s) & 4095ull) * 8 + (slot)] = xtime(); } while (0)
#else
#define XTR(s, slot) do { } while (0)
#define XTR_LAST(s, slot) do { } while (0)
#endif
// thread 0 of the block polls, in THIS rank's header, counter `field` (0 ready, 1 done) of every peer in `mask`
__device__ void xchg_wait(const XHdr *mine, uint64_t mask, int field, unsigned long long want, int *err)
{
    if (threadIdx.x == 0) {
        const unsigned long long t0 = xtime();
        for (int r = 0; r < XCHG_MAX_RANKS && (mask >> r); ++r) {
            if (!((mask >> r) & 1)) continue;
            const unsigned long long *flag = (field ? mine->done : mine->ready) + r;
            while (ld_acquire_sys(flag) < want) {
                if (xtime() - t0 > XCHG_TIMEOUT_NS) { atomicExch(err, 1); break; }
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
}

struct XPartDev {
    int64_t nrow, nshared, off;
    int scaled;
    const int32_t *loc, *slot;
    const uint64_t *owners;
    const double *wtot, *wlocal;
    double *Y;
    int64_t ldy;
};
constexpr int XCHG_MAX_PARTS = 4;
struct XArgs {
    XPartDev part[XCHG_MAX_PARTS];
    int nparts, rank, nranks, nvar;
    int64_t total, ldb;
    const unsigned long long *peers;        // [nranks] region addresses
    uint64_t reads_mask, readers_mask;
    unsigned long long *step;
    unsigned *counters;
    int *err;
};

__device__ __forceinline__ double *xchg_inbox(unsigned long long region, int parity, int nranks, int from, int64_t total, int64_t ldb)
{
    return reinterpret_cast<double *>(region + sizeof(XHdr)) + ((int64_t)parity * nranks + from) * total * ldb;
}

// The waits as kernels of their own (one warp): a publish or gather that spins for a late peer would hold the few CTA
// slots a concurrent SpMM leaves free -- the very slots the OTHER kernel of the pair needs to make that peer's wait end.
// The kernels proper repeat the wait, which is then already satisfied.
__global__ void xchg_wait_kernel(const XArgs a, int field)
{
    XHdr *mine = reinterpret_cast<XHdr *>(a.peers[a.rank]);
    if (field) {
        const unsigned long long s = a.step[1];
        if (s >= XCHG_DEPTH) xchg_wait(mine, a.readers_mask, 1, s - XCHG_DEPTH + 1, a.err);
    } else {
        xchg_wait(mine, a.reads_mask, 0, a.step[0] + 1, a.err);
    }
}

// one thread per (shared cell of any part, field pair), four items in flight per thread: this rank's row of the cell
// is stored into the inbox of every other owner
// (128-thread blocks: next to a running SpMM the two kernels share the few CTA slots it leaves free -- 20 K registers each --
// and a 256-thread publish block and a gather block did not fit one slot together, so publish and gather took turns)
__global__ void __launch_bounds__(XCHG_THREADS) xchg_publish_kernel(const XArgs a)
{
    const unsigned long long s = a.step[1];
    XHdr *mine = reinterpret_cast<XHdr *>(a.peers[a.rank]);
    XTR(s, 0);
    if (s >= XCHG_DEPTH) xchg_wait(mine, a.readers_mask, 1, s - XCHG_DEPTH + 1, a.err);      // the peers have finished reading slot s % DEPTH
    XTR(s, 1);
    const int nv2 = (a.nvar + 1) / 2;
    const int64_t work = a.total * nv2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < work; t0 += 4 * stride) {
        double2 v[4];
        uint64_t own[4];
        int64_t at[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                        // four independent loads in flight per thread
            own[u] = 0;
            const int64_t t = t0 + u * stride;
            if (t >= work) continue;
            const int64_t g = t / nv2;
            const int f = (int)(t - g * nv2) * 2;
            int p = 0;
            while (p + 1 < a.nparts && g >= a.part[p + 1].off) ++p;
            const XPartDev &P = a.part[p];
            const int32_t r = P.loc[g - P.off];
            if (r < 0) continue;                             // not one of my cells
            const double *y = P.Y + (int64_t)r * P.ldy + f;
            v[u] = make_double2(y[0], f + 1 < a.nvar ? y[1] : 0.0);
            if (P.scaled) { const double w = P.wlocal[r]; v[u].x *= w; v[u].y *= w; }      // back to this sheet's unscaled sum
            own[u] = P.owners[g - P.off] & ~(uint64_t(1) << a.rank);
            at[u] = g * a.ldb + f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            for (uint64_t o = own[u]; o; o &= o - 1) {       // posted stores over NVLink
                const int q = __ffsll((long long)o) - 1;
                *reinterpret_cast<double2 *>(xchg_inbox(a.peers[q], (int)(s % XCHG_DEPTH), a.nranks, a.rank, a.total, a.ldb) + at[u]) = v[u];
            }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(a.counters, 1u) == gridDim.x - 1) {
        a.counters[0] = 0;
        XTR_LAST(s, 2);
        __threadfence_system();
        for (uint64_t o = a.readers_mask; o; o &= o - 1) {
            const int q = __ffsll((long long)o) - 1;
            st_release_sys(reinterpret_cast<XHdr *>(a.peers[q])->ready + a.rank, s + 1);
        }
        XTR_LAST(s, 3);
        a.step[1] = s + 1;
    }
}

// one thread per (cell, field pair): a shared cell's value is the sum over its owners, in ascending rank order, of this
// rank's own row and the rows in its inbox (all local memory), times 1 / (summed weights); unscaled parts also divide
// the rows of unshared cells by their own weight
__global__ void __launch_bounds__(XCHG_THREADS) xchg_gather_kernel(const XArgs a)
{
    const unsigned long long s = *a.step;
    XHdr *mine = reinterpret_cast<XHdr *>(a.peers[a.rank]);
    XTR(s, 4);
    xchg_wait(mine, a.reads_mask, 0, s + 1, a.err);
    XTR(s, 5);
    const int nv2 = (a.nvar + 1) / 2;
    // work items: every local row of an unscaled part, only the shared cells of a scaled part
    int64_t cum[XCHG_MAX_PARTS + 1];
    cum[0] = 0;
    for (int p = 0; p < a.nparts; ++p) cum[p + 1] = cum[p] + (a.part[p].scaled ? a.part[p].nshared : a.part[p].nrow);
    const int64_t work = cum[a.nparts] * nv2;
    const double *inbox0 = xchg_inbox(a.peers[a.rank], (int)(s % XCHG_DEPTH), a.nranks, 0, a.total, a.ldb);
    const int64_t per_rank = a.total * a.ldb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t g = t / nv2;
        const int f = (int)(t - g * nv2) * 2;
        int p = 0;
        while (p + 1 < a.nparts && g >= cum[p + 1]) ++p;
        g -= cum[p];
        const XPartDev &P = a.part[p];
        int32_t k;
        if (P.scaled) {
            k = (int32_t)g;
            const int32_t r = P.loc[k];
            if (r < 0) continue;                            // not one of my cells
            g = r;
        } else {
            k = P.slot[g];
        }
        double *y = P.Y + g * P.ldy + f;
        double2 v = make_double2(y[0], f + 1 < a.nvar ? y[1] : 0.0);
        double sc;
        if (k >= 0) {
            if (P.scaled) { const double w = P.wlocal[g]; v.x *= w; v.y *= w; }
            const double2 own = v;
            const int64_t at = (P.off + k) * a.ldb + f;
            v = make_double2(0.0, 0.0);
            for (uint64_t o = P.owners[k]; o; o &= o - 1) {  // ascending rank order: every owner adds in the same order
                const int q = __ffsll((long long)o) - 1;
                const double2 c = q == a.rank ? own : ld_peer(inbox0 + q * per_rank + at);
                v.x += c.x; v.y += c.y;
            }
            sc = 1. / P.wtot[k];                            // to_dense: val * scaleE, scaleE = 1 / sum(weights)
        } else {
            sc = 1. / P.wlocal[g];
        }
        y[0] = v.x * sc;
        if (f + 1 < a.nvar) y[1] = v.y * sc;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(a.counters + 1, 1u) == gridDim.x - 1) {
        a.counters[1] = 0;
        __threadfence_system();
        for (uint64_t o = a.reads_mask; o; o &= o - 1) {
            const int q = __ffsll((long long)o) - 1;
            st_release_sys(reinterpret_cast<XHdr *>(a.peers[q])->done + a.rank, s + 1);
        }
        XTR_LAST(s, 6);
        *a.step = s + 1;
    }
}

static XArgs make_args(Xchg &x, double *const *dY, const int64_t *ldy, int nvar)
{
    if (!x.connected) fail(IB_ERR_ARG, "ib_xchg: not connected (call ib_xchg_open or ib_xchg_open_local first)");
    if (nvar <= 0 || nvar > x.ldb) fail(IB_ERR_ARG, "ib_xchg: nvar %d exceeds the %ld fields the context was created for", nvar, (long)x.ldb);
    XArgs a{};
    a.nparts = (int)x.parts.size(); a.rank = x.rank; a.nranks = x.nranks; a.nvar = nvar; a.total = x.total; a.ldb = x.ldb;
    for (int p = 0; p < a.nparts; ++p) {
        XPart &P = x.parts[p];
        if (P.nrow > 0 && !P.slot.p) fail(IB_ERR_ARG, "ib_xchg: part %d has no plan (ib_xchg_set_part)", p);
        if (!dY[p] && P.nrow > 0) fail(IB_ERR_ARG, "ib_xchg: null result buffer");
        if (ldy[p] < nvar) fail(IB_ERR_ARG, "ib_xchg: leading dimension smaller than nvar");
        a.part[p] = XPartDev{P.nrow, P.nshared, P.off, P.scaled, P.loc.p, P.slot.p, P.owners.p, P.wtot.p, P.wlocal.p, dY[p], ldy[p]};
    }
    a.peers = x.d_peer.p;
    a.reads_mask = x.reads_mask; a.readers_mask = x.readers_mask;
    a.step = x.step.p; a.counters = x.counters.p; a.err = x.err.p;
    return a;
}

void xchg_publish(Xchg &x, double *const *dY, const int64_t *ldy, int nvar, cudaStream_t s)
{
    XArgs a = make_args(x, dY, ldy, nvar);
    // few blocks, several items per thread: the kernels usually run in the CTA slots a resident SpMM grid leaves free
    const int64_t work = std::max<int64_t>(1, x.total * ((nvar + 1) / 2));
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(work, XCHG_THREADS * 4), g_tune.xchg_blocks);
    xchg_wait_kernel<<<1, 32, 0, s>>>(a, 1);
    IB_LAUNCHED();
    xchg_publish_kernel<<<grid, XCHG_THREADS, 0, s>>>(a);
    IB_LAUNCHED();
}
void xchg_gather(Xchg &x, double *const *dY, const int64_t *ldy, int nvar, cudaStream_t s)
{
    XArgs a = make_args(x, dY, ldy, nvar);
    int64_t rows = 0;
    for (auto &P : x.parts) rows += P.scaled ? P.nshared : P.nrow;
    const int64_t work = std::max<int64_t>(1, rows * ((nvar + 1) / 2));
    const bool all_scaled = std::all_of(x.parts.begin(), x.parts.end(), [](XPart const &P) { return P.scaled != 0; });
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(work, XCHG_THREADS * 4), all_scaled ? g_tune.xchg_blocks : 4 * sm_count());
    xchg_wait_kernel<<<1, 32, 0, s>>>(a, 0);
    IB_LAUNCHED();
    xchg_gather_kernel<<<grid, XCHG_THREADS, 0, s>>>(a);
    IB_LAUNCHED();
}

}   // namespace ib

using namespace ib;
struct ib_xchg { Xchg x; };
#define X_TRY try {
#define X_CATCH                                                                  \
    }                                                                            \
    catch (ib::Error const &e) { ib::set_last_error(e.what()); return e.code; }  \
    catch (std::exception const &e) { ib::set_last_error(e.what()); return IB_ERR_ARG; }

extern "C" {

int ib_xchg_create(ib_xchg **out, int rank, int nranks, int nparts, const int64_t *nshared, int32_t nvar_max)
{
    X_TRY
    if (!out || nranks < 1 || nranks > XCHG_MAX_RANKS || rank < 0 || rank >= nranks || nparts < 1 || nparts > XCHG_MAX_PARTS || !nshared || nvar_max < 1)
        fail(IB_ERR_ARG, "ib_xchg_create: bad argument");
    ensure_device();
    std::unique_ptr<ib_xchg> h(new ib_xchg);
    Xchg &x = h->x;
    x.rank = rank; x.nranks = nranks; x.device = device_index();
    x.stream = default_stream();
    x.ldb = (nvar_max + 1) & ~1;
    x.parts.resize(nparts);
    for (int p = 0; p < nparts; ++p) {
        if (nshared[p] < 0) fail(IB_ERR_ARG, "ib_xchg_create: negative cell count");
        x.parts[p].nshared = nshared[p];
        x.parts[p].off = x.total;
        x.total += nshared[p];
    }
    x.region_bytes = sizeof(XHdr) + (size_t)XCHG_DEPTH * nranks * std::max<int64_t>(x.total, 1) * x.ldb * sizeof(double);
    IB_CUDA(cudaMalloc(&x.region, x.region_bytes));
    IB_CUDA(cudaMemset(x.region, 0, x.region_bytes));
    x.peer[rank] = x.region;
    x.step.alloc(2, x.stream); x.step.zero();
    x.counters.alloc(2, x.stream); x.counters.zero();
    x.err.alloc(1, x.stream); x.err.zero();
    x.d_peer.alloc((size_t)nranks, x.stream);
    IB_CUDA(cudaStreamSynchronize(x.stream));
    *out = h.release();
    return IB_OK;
    X_CATCH
}

void ib_xchg_free(ib_xchg *h)
{
    if (!h) return;
    Xchg &x = h->x;
    cudaDeviceSynchronize();
    for (int r = 0; r < x.nranks; ++r)
        if (x.ipc_opened[r] && x.peer[r]) cudaIpcCloseMemHandle(x.peer[r]);
    if (x.region) cudaFree(x.region);
    delete h;
}

int ib_xchg_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int ib_xchg_get_handle(const ib_xchg *h, void *handle)
{
    X_TRY
    if (!h || !handle) fail(IB_ERR_ARG, "ib_xchg_get_handle: bad argument");
    cudaIpcMemHandle_t ipc;
    IB_CUDA(cudaIpcGetMemHandle(&ipc, h->x.region));
    memcpy(handle, &ipc, sizeof ipc);
    return IB_OK;
    X_CATCH
}

static void finish_connect(Xchg &x)
{
    std::vector<unsigned long long> addr((size_t)x.nranks);
    for (int r = 0; r < x.nranks; ++r) {
        if (!x.peer[r]) fail(IB_ERR_ARG, "ib_xchg: rank %d not mapped", r);
        addr[r] = (unsigned long long)(uintptr_t)x.peer[r];
    }
    x.d_peer.upload(addr.data(), addr.size());
    IB_CUDA(cudaStreamSynchronize(x.stream));
    x.connected = true;
}

/* handles: [nranks][ib_xchg_handle_bytes()] as all-gathered by the host (MPI_Allgather / torch.distributed) */
int ib_xchg_open(ib_xchg *h, const void *handles)
{
    X_TRY
    if (!h || (!handles && h->x.nranks > 1)) fail(IB_ERR_ARG, "ib_xchg_open: bad argument");
    Xchg &x = h->x;
    for (int r = 0; r < x.nranks; ++r) {
        if (r == x.rank || x.peer[r]) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const char *)handles + (size_t)r * sizeof ipc, sizeof ipc);
        IB_CUDA(cudaIpcOpenMemHandle(&x.peer[r], ipc, cudaIpcMemLazyEnablePeerAccess));
        x.ipc_opened[r] = true;
    }
    finish_connect(x);
    return IB_OK;
    X_CATCH
}

/* ranks living in ONE process (a host driving several GPUs, or the single-GPU tests): peers[r] = rank r's context */
int ib_xchg_open_local(ib_xchg *h, ib_xchg *const *peers)
{
    X_TRY
    if (!h || !peers) fail(IB_ERR_ARG, "ib_xchg_open_local: bad argument");
    Xchg &x = h->x;
    for (int r = 0; r < x.nranks; ++r) {
        if (r == x.rank) continue;
        if (!peers[r] || peers[r]->x.rank != r || peers[r]->x.total != x.total || peers[r]->x.ldb != x.ldb)
            fail(IB_ERR_ARG, "ib_xchg_open_local: context of rank %d does not match", r);
        if (peers[r]->x.device != x.device) {
            int can = 0;
            IB_CUDA(cudaDeviceCanAccessPeer(&can, x.device, peers[r]->x.device));
            if (!can) fail(IB_ERR_CUDA, "ib_xchg_open_local: device %d cannot access device %d", x.device, peers[r]->x.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->x.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) IB_CUDA(e);
            cudaGetLastError();
        }
        x.peer[r] = peers[r]->x.region;
    }
    finish_connect(x);
    return IB_OK;
    X_CATCH
}

/* Plan of part `part` (all host arrays): nrow local dense rows; for the part's shared cells k (the same list, in the same
 * order, on every rank): loc[k] = local row or -1, owners[k] = bit mask of the ranks holding the cell, wtot[k] = summed
 * weights over the owners; wlocal[nrow] = this rank's weights (wM of the matrix).  scaled = 0: Y holds the UNSCALED sums
 * (the coupler's E1vI_unscaled, IceCoupler.cpp:371-377) and every row is divided by its weight here; scaled = 1: Y comes from
 * the SCALED matrix (rows already divided by wlocal), so only the shared cells are touched: a rank publishes Y * wlocal
 * and the rows of unshared cells are final as they are. */
int ib_xchg_set_part(ib_xchg *h, int part, int64_t nrow, const int32_t *loc, const uint64_t *owners, const double *wtot, const double *wlocal,
                     int scaled)
{
    X_TRY
    if (!h || part < 0 || part >= (int)h->x.parts.size() || nrow < 0) fail(IB_ERR_ARG, "ib_xchg_set_part: bad argument");
    Xchg &x = h->x;
    XPart &P = x.parts[part];
    if ((P.nshared && (!loc || !owners || !wtot)) || (nrow && !wlocal)) fail(IB_ERR_ARG, "ib_xchg_set_part: null array");
    cudaStream_t s = x.stream;
    P.nrow = nrow;
    P.scaled = scaled != 0;
    std::vector<int32_t> slot((size_t)nrow, -1);
    for (int64_t k = 0; k < P.nshared; ++k) {
        const bool mine = loc[k] >= 0;
        if (mine != (((owners[k] >> x.rank) & 1) != 0)) fail(IB_ERR_ARG, "ib_xchg_set_part: loc and owners disagree at shared cell %ld", (long)k);
        if (owners[k] >> x.nranks) fail(IB_ERR_ARG, "ib_xchg_set_part: owner outside the rank range");
        if (!mine) continue;
        if (loc[k] >= nrow || slot[loc[k]] >= 0) fail(IB_ERR_ARG, "ib_xchg_set_part: bad local row at shared cell %ld", (long)k);
        slot[loc[k]] = (int32_t)k;
        const uint64_t others = owners[k] & ~(uint64_t(1) << x.rank);
        x.reads_mask |= others;
        x.readers_mask |= others;        // symmetric: whoever holds the cell reads every other holder
    }
    P.loc.alloc((size_t)P.nshared, s); P.loc.upload(loc, (size_t)P.nshared);
    P.owners.alloc((size_t)P.nshared, s); P.owners.upload(owners, (size_t)P.nshared);
    P.wtot.alloc((size_t)P.nshared, s); P.wtot.upload(wtot, (size_t)P.nshared);
    P.slot.alloc((size_t)nrow + 1, s); P.slot.upload(slot.data(), (size_t)nrow);
    P.wlocal.alloc((size_t)nrow + 1, s); P.wlocal.upload(wlocal, (size_t)nrow);
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    X_CATCH
}

int ib_xchg_publish_dev(ib_xchg *h, double *const *dY, const int64_t *ldy, int32_t nvar, void *stream)
{
    X_TRY
    if (!h || !dY || !ldy) fail(IB_ERR_ARG, "ib_xchg_publish_dev: bad argument");
    xchg_publish(h->x, dY, ldy, nvar, stream ? (cudaStream_t)stream : h->x.stream);
    return IB_OK;
    X_CATCH
}
int ib_xchg_gather_dev(ib_xchg *h, double *const *dY, const int64_t *ldy, int32_t nvar, void *stream)
{
    X_TRY
    if (!h || !dY || !ldy) fail(IB_ERR_ARG, "ib_xchg_gather_dev: bad argument");
    xchg_gather(h->x, dY, ldy, nvar, stream ? (cudaStream_t)stream : h->x.stream);
    return IB_OK;
    X_CATCH
}
/* publish + gather: the whole merge of one step.  dY[p]: this sheet's unscaled sums [nrow_p][ldy_p]; on completion the
 * merged, scaled values. */
int ib_xchg_merge_dev(ib_xchg *h, double *const *dY, const int64_t *ldy, int32_t nvar, void *stream)
{
    X_TRY
    if (!h || !dY || !ldy) fail(IB_ERR_ARG, "ib_xchg_merge_dev: bad argument");
    cudaStream_t s = stream ? (cudaStream_t)stream : h->x.stream;
    xchg_publish(h->x, dY, ldy, nvar, s);
    xchg_gather(h->x, dY, ldy, nvar, s);
    return IB_OK;
    X_CATCH
}
/* 0: fine; IB_ERR_CUDA: a wait on a peer timed out (synchronises the context's work first) */
int ib_xchg_status(ib_xchg *h)
{
    X_TRY
    if (!h) fail(IB_ERR_ARG, "ib_xchg_status: bad argument");
    IB_CUDA(cudaDeviceSynchronize());
    int e = 0;
    IB_CUDA(cudaMemcpy(&e, h->x.err.p, sizeof e, cudaMemcpyDeviceToHost));
    if (e) fail(IB_ERR_CUDA, "ib_xchg: a peer did not answer within %.1f s", XCHG_TIMEOUT_NS * 1e-9);
    return IB_OK;
    X_CATCH
}
int64_t ib_xchg_step(ib_xchg *h)
{
    if (!h) return -1;
    unsigned long long s = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpy(&s, h->x.step.p, sizeof s, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)s;
}

/* Host-side plan: given every rank's list of sparse ids (ids[r][n[r]], the dims[0] of its matrix, as gathered by the
 * host), the cells held by two or more ranks in ascending id order.  Call with shared == NULL for the count.
 * loc[k]: row of shared[k] in `rank`'s list or -1; owners[k]: bit mask of the ranks holding it. */
int64_t ib_xchg_plan(int nranks, const int64_t *const *ids, const int64_t *n, int64_t n_sparse, int rank, int64_t *shared, int32_t *loc,
                     uint64_t *owners)
{
    X_TRY
    if (nranks < 1 || nranks > XCHG_MAX_RANKS || !ids || !n || n_sparse < 0 || rank < 0 || rank >= nranks) fail(IB_ERR_ARG, "ib_xchg_plan: bad argument");
    std::vector<uint64_t> mask((size_t)n_sparse, 0);
    std::vector<int32_t> mine((size_t)n_sparse, -1);
    for (int r = 0; r < nranks; ++r)
        for (int64_t i = 0; i < n[r]; ++i) {
            const int64_t id = ids[r][i];
            if (id < 0 || id >= n_sparse) fail(IB_ERR_ARG, "ib_xchg_plan: id %ld outside the grid", (long)id);
            mask[id] |= uint64_t(1) << r;
            if (r == rank) mine[id] = (int32_t)i;
        }
    int64_t k = 0;
    for (int64_t id = 0; id < n_sparse; ++id) {
        if (__builtin_popcountll(mask[id]) < 2) continue;
        if (shared) { shared[k] = id; loc[k] = mine[id]; owners[k] = mask[id]; }
        ++k;
    }
    return k;
    X_CATCH
}

#ifdef IB_STEP_TRACE
int ib_debug_xchg_steps(unsigned long long *out, int n)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_x_steps, sizeof(unsigned long long) * 8 * n);
}
#endif

}   // extern "C"
