// apply.cu -- batched fp64 CSR SpMM  Y[nrow x nvar] = M * X[ncol x nvar]  with fields stored cell-major
// (the nvar values of one cell are contiguous), so every nonzero drives one contiguous row read of X.
//
// Replaces ibmisc linear::Weighted_Eigen::apply()/apply_M() (call sites merge_topo.cpp:65, modele/icebin22m.cpp:153)
// and the explicit Eigen products of the coupling step (IceCoupler.cpp:237-238, 445-446).
//
// Work decomposition ("items", built once per matrix on the device):
//   * consecutive SHORT rows (<= APPLY_CHUNK nonzeros) are packed into groups of < 2*APPLY_CHUNK nonzeros and
//     <= 32 rows, so that matrices with 2 nonzeros per row (IvE, IvA) move as much data per unit as long rows do;
//   * LONG rows are cut into chunks of APPLY_CHUNK nonzeros whose partial sums are combined, in chunk order, by a
//     small second kernel (a fused "last block adds" variant was measured 15-25 % slower: its fence + atomic round
//     trip stalls the consumers of every chunk while the ring drains).
// Main kernel (spmm_tma_kernel): one producer warp streams the X rows named by an item's column indices into a
// shared-memory ring with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers; eight
// consumer warps read each landed row once (one double2 per thread) and keep running sums in registers.  Bytes in
// flight per SM are set by the ring, not by registers.  A plain-load kernel (spmm_ldg_kernel) covers small or odd
// nvar and unaligned buffers.
//
// Compiled with -fmad=false: a row is accumulated as acc = acc + v*x over ascending columns, exactly the order and
// rounding of Eigen's column-major sparse * dense on the CPU, so rows of <= APPLY_CHUNK nonzeros agree with the
// oracle bit for bit; split rows differ only in association.
//
// Roofline: HBM.  Algorithmic bytes = 12*nnz + 4*(nrow+1) + 8*nvar*(ncol_touched + nrow)   (SURVEY 8d).
#include <cooperative_groups.h>

#include <cstdlib>

#include "gc_plan.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

constexpr int APPLY_CHUNK = 32;
constexpr int GROUP_ROWS = 32;
static int pair_chunk() { static int v = getenv("IB_PAIR_CHUNK") ? atoi(getenv("IB_PAIR_CHUNK")) : 48; return v; }

// ---- plan -----------------------------------------------------------------------------------------------
// item = (first_row, nnz_begin, nnz_end, w):  w >= 0: chunk of a long row, w = partial slot
//                                             w <  0: group of -w consecutive short rows
struct PlanLoad {
    const int32_t *rowptr;
    __device__ __forceinline__ int len(int64_t r) const { return rowptr[r + 1] - rowptr[r]; }
    __device__ __forceinline__ bool same_group(int64_t a, int64_t b) const   // a, b short rows, b = a + 1
    {
        return (rowptr[a] / APPLY_CHUNK == rowptr[b] / APPLY_CHUNK) && (a / GROUP_ROWS == b / GROUP_ROWS);
    }
    __device__ uint4 operator()(int64_t r) const
    {
        const int n = len(r);
        if (n > APPLY_CHUNK) {
            const uint32_t nch = (uint32_t)((n + APPLY_CHUNK - 1) / APPLY_CHUNK);
            return make_uint4(nch, nch, 1u, 0u);
        }
        const bool head = (r == 0) || len(r - 1) > APPLY_CHUNK || !same_group(r - 1, r);
        return make_uint4(head ? 1u : 0u, 0u, 0u, 0u);
    }
};
struct PlanApply {
    PlanLoad L;
    int64_t nrow;
    int4 *items;
    int4 *split_rows;      // (row, first partial slot, nchunks, 0)
    int32_t *slot_srow;    // partial slot -> split-row index
    __device__ void operator()(int64_t r, uint4 prefix, uint4 e) const
    {
        const int b = L.rowptr[r], end = L.rowptr[r + 1];
        if (e.z) {
            for (uint32_t q = 0; q < e.x; ++q) {
                const int lo = b + (int)q * APPLY_CHUNK;
                items[prefix.x + q] = make_int4((int)r, lo, min(end, lo + APPLY_CHUNK), (int)(prefix.y + q));
                slot_srow[prefix.y + q] = (int32_t)prefix.z;
            }
            split_rows[prefix.z] = make_int4((int)r, (int)prefix.y, (int)e.x, 0);
        } else if (e.x) {
            int64_t r2 = r + 1;
            while (r2 < nrow && L.len(r2) <= APPLY_CHUNK && L.same_group(r2 - 1, r2)) ++r2;
            items[prefix.x] = make_int4((int)r, b, L.rowptr[r2], -(int)(r2 - r));
        }
    }
};

static bool make_pair_plan(Weighted &W, cudaStream_t s);
static void make_gc_plan(Weighted &W, cudaStream_t s);
struct MaxU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
struct RowLenLoad {
    const int32_t *rowptr;
    __device__ uint32_t operator()(int64_t r) const { return (uint32_t)(rowptr[r + 1] - rowptr[r]); }
};

static void make_plan(Weighted &W, cudaStream_t s)
{
    ApplyPlan &P = W.plan;
    if (P.ready) return;
    if (W.nrow > 0) {
        ScanState<uint4> st;
        const uint4 z = make_uint4(0, 0, 0, 0);
        PlanLoad L{W.rowptr.p};
        transform_scan<uint4>(st, W.nrow, L, AddU32x4{}, z, s);
        uint4 tot;
        IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        P.nitems = tot.x; P.npartials = tot.y; P.nsplit_rows = tot.z;
        P.items.alloc((size_t)P.nitems, s);
        P.split_rows.alloc((size_t)P.nsplit_rows, s);
        P.slot_srow.alloc((size_t)P.npartials, s);
        scan_apply<uint4>(st, L, AddU32x4{}, z, PlanApply{L, W.nrow, P.items.p, P.split_rows.p, P.slot_srow.p}, s);
        {
            ScanState<uint32_t> ms;
            transform_scan<uint32_t>(ms, W.nrow, RowLenLoad{W.rowptr.p}, MaxU32{}, 0u, s);
            uint32_t mx = 0;
            IB_CUDA(cudaMemcpyAsync(&mx, ms.total.p, sizeof mx, cudaMemcpyDeviceToHost, s));
            IB_CUDA(cudaStreamSynchronize(s));
            P.max_row_len = (int)mx;
        }
        make_pair_plan(W, s);
        if (P.max_row_len <= GC_S && P.nsplit_rows == 0 && !getenv("IB_SPMM_NOGC")) make_gc_plan(W, s);
    }
    P.ready = true;
}

// ---- distinct-column groups (gc_plan.cuh): one thread per 32-row block runs the greedy grouping twice
// (count -> scan -> emit), exactly like the other plans
struct GcCountLoad {
    const int32_t *rowptr, *col;
    int64_t nrow;
    __device__ uint4 operator()(int64_t b) const
    {
        const GcCount c = gc_block<false>(rowptr, col, nullptr, nrow, b, 0u, 0u, nullptr, nullptr, nullptr);
        return make_uint4(c.items, c.recs, 0u, 0u);
    }
};
struct GcEmit {
    const int32_t *rowptr, *col;
    const double *val;
    int64_t nrow;
    GcItem *items;
    int32_t *ucol;
    GcRec *recs;
    __device__ void operator()(int64_t b, uint4 prefix, uint4) const { gc_block<true>(rowptr, col, val, nrow, b, prefix.x, prefix.y, items, ucol, recs); }
};
static_assert(sizeof(GcItem) == sizeof(int4) && sizeof(GcRec) == sizeof(int4), "GcItem / GcRec must alias int4");

static void make_gc_plan(Weighted &W, cudaStream_t s)
{
    ApplyPlan &P = W.plan;
    const int64_t nblk = cdiv(W.nrow, GC_BLOCK);
    ScanState<uint4> st;
    const uint4 z = make_uint4(0, 0, 0, 0);
    GcCountLoad L{W.rowptr.p, W.col.p, W.nrow};
    transform_scan<uint4>(st, nblk, L, AddU32x4{}, z, s);
    uint4 tot;
    IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    if ((int64_t)tot.y != W.nnz) return;          // empty rows (dummy records): left to the other kernels
    P.gc_nitems = tot.x;
    P.gc_items.alloc((size_t)tot.x, s);
    P.gc_ucol.alloc((size_t)tot.x * GC_S, s);
    P.gc_recs.alloc((size_t)tot.y + 2, s);         // the kernel's record prefetch runs two entries ahead
    P.gc_recs.zero();
    P.gc_counter.alloc(3, s);
    P.gc_counter.zero();
    scan_apply<uint4>(st, L, AddU32x4{}, z,
                      GcEmit{W.rowptr.p, W.col.p, W.val.p, W.nrow, reinterpret_cast<GcItem *>(P.gc_items.p), P.gc_ucol.p,
                             reinterpret_cast<GcRec *>(P.gc_recs.p)}, s);
    IB_CUDA(cudaStreamSynchronize(s));
    P.gc = true;
}

// ---- pair-grouped column plan ------------------------------------------------------------------------------
// EvI-like matrices have (at most) two nonzeros per COLUMN: an ice cell feeds the two elevation classes that bracket
// its elevation (IceRegridder_L0.cpp:126-143).  Row-wise processing then reads every X row twice.  Instead, columns
// are regrouped by the pair of rows (r0, r1) they feed; a chunk of <= APPLY_CHUNK such columns is one item that reads
// each X row ONCE and accumulates two partial rows (weights v0 -> r0, v1 -> r1).  Every output row is then the sum,
// in a fixed order, of the partial rows listed for it (fx_ptr/fx_slot).
__global__ void expand_rows_kernel(const int32_t *rowptr, int64_t nrow, int32_t *row)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) row[j] = (int32_t)r;
}
__global__ void u32_to_i32_copy_kernel(const uint32_t *a, int64_t n, int32_t *b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)a[i];
}
// A column with n nonzeros becomes ceil(n/2) "virtual columns", each owning two consecutive nonzeros of the column
// (ice cells on the border between GCM cells feed 4 or 8 elevation classes).  vcount -> scan -> vbase.
struct VCountLoad {
    const int32_t *colptr;
    __device__ uint32_t operator()(int64_t c) const { return (uint32_t)((colptr[c + 1] - colptr[c] + 1) / 2); }
};
// per virtual column: key = (r0 << 32 | r1 + 1), owning column, first CSC entry
struct VEmit {
    const int32_t *colptr;
    const uint32_t *perm;
    const int32_t *row;
    uint64_t *key;
    uint32_t *vid, *vcol, *vent;
    __device__ void operator()(int64_t c, uint32_t base, uint32_t nv) const
    {
        const int b = colptr[c], e = colptr[c + 1];
        for (uint32_t i = 0; i < nv; ++i) {
            const int q = b + 2 * (int)i;
            const uint32_t r0 = (uint32_t)row[perm[q]];
            const uint32_t r1p = (q + 1 < e) ? (uint32_t)row[perm[q + 1]] + 1u : 0u;
            key[base + i] = ((uint64_t)r0 << 32) | r1p;
            vid[base + i] = base + i;
            vcol[base + i] = (uint32_t)c;
            vent[base + i] = (uint32_t)q;
        }
    }
};
// position p in the regrouped stream: x = starts an item, y = (position valid)
struct PairLoad {
    const uint64_t *key;
    const uint32_t *gstart;      // group start of every position
    uint32_t chunk;
    __device__ uint32_t operator()(int64_t p) const { return ((uint32_t)p - gstart[p]) % chunk == 0 ? 1u : 0u; }
};
struct GroupStartLoad {          // max-scan of (head ? p + 1 : 0)
    const uint64_t *key;
    __device__ uint32_t operator()(int64_t p) const { return (p == 0 || key[p] != key[p - 1]) ? (uint32_t)p + 1u : 0u; }
};
struct GroupStartStore {
    const uint64_t *key;
    uint32_t *gstart;
    __device__ void operator()(int64_t p, uint32_t prefix, uint32_t mine) const { gstart[p] = (mine ? mine : prefix) - 1u; }
};
struct PairEmit {
    const uint64_t *key;
    const uint32_t *vid, *vcol, *vent;   // vid: sorted order -> virtual column
    const int32_t *colptr;
    const uint32_t *perm;
    const double *val;
    int64_t nv;
    int chunk;
    int4 *items;
    int32_t *pcol;
    double *pv0, *pv1;
    uint32_t *fxkey, *fxslot;    // [2*nitems] (row, slot) pairs for the fix-up lists; row = 0xffffffff if unused
    __device__ void operator()(int64_t p, uint32_t prefix, uint32_t start) const
    {
        const uint64_t k = key[p];
        const uint32_t v = vid[p];
        const uint32_t c = vcol[v], q = vent[v];
        pcol[p] = (int32_t)c;
        pv0[p] = val[perm[q]];
        pv1[p] = ((int)q + 1 < colptr[c + 1]) ? val[perm[q + 1]] : 0.0;
        if (!start) return;
        int64_t e = p + 1;
        while (e < nv && e - p < chunk && key[e] == k) ++e;
        const int32_t r0 = (int32_t)(k >> 32), r1 = (int32_t)(uint32_t)k - 1;
        items[prefix] = make_int4(r0, (int)p, (int)e, (int)(2 * prefix));
        fxkey[2 * prefix] = (uint32_t)r0; fxslot[2 * prefix] = 2 * prefix;
        fxkey[2 * prefix + 1] = r1 >= 0 ? (uint32_t)r1 : 0xffffffffu; fxslot[2 * prefix + 1] = 2 * prefix + 1;
    }
};

static bool make_pair_plan(Weighted &W, cudaStream_t s)
{
    ApplyPlan &P = W.plan;
    const int64_t nnz = W.nnz, ncol = W.ncol, nrow = W.nrow;
    if (getenv("IB_SPMM_NOPAIR")) return false;
    if (nnz < 1024 || 10 * nnz < 14 * ncol || nnz > 3 * ncol) return false;      // need ~2 nonzeros in most columns
    // CSC view: stable sort of the nonzeros by column
    DBuf<int32_t> row((size_t)nnz, s), csorted((size_t)nnz, s), colptr((size_t)ncol + 1, s);
    DBuf<uint32_t> k0((size_t)nnz, s), k1((size_t)nnz, s), p0((size_t)nnz, s), p1((size_t)nnz, s);
    expand_rows_kernel<<<(unsigned)cdiv(nrow, 128), 128, 0, s>>>(W.rowptr.p, nrow, row.p);
    IB_LAUNCHED();
    IB_CUDA(cudaMemcpyAsync(k0.p, W.col.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    iota_u32_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(p0.p, nnz);
    IB_LAUNCHED();
    const int w1 = radix_sort_pairs<uint32_t, uint32_t>(k0.p, p0.p, k1.p, p1.p, nnz, bits_for(ncol), s);
    const uint32_t *perm = w1 ? p1.p : p0.p;
    u32_to_i32_copy_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(w1 ? k1.p : k0.p, nnz, csorted.p);
    IB_LAUNCHED();
    colptr.zero();
    seg_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(csorted.p, nnz, ncol, colptr.p);
    IB_LAUNCHED();
    // virtual columns, regrouped by the pair of rows (r0, r1) they feed
    ScanState<uint32_t> vs;
    transform_scan<uint32_t>(vs, ncol, VCountLoad{colptr.p}, AddU32{}, 0u, s);
    uint32_t nv_u = 0;
    IB_CUDA(cudaMemcpyAsync(&nv_u, vs.total.p, sizeof nv_u, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    const int64_t nv = nv_u;
    if (nv == 0 || 10 * nv > 7 * nnz) return false;            // not enough sharing to pay for the partial rows
    DBuf<uint64_t> key0((size_t)nv, s), key1((size_t)nv, s);
    DBuf<uint32_t> vid0((size_t)nv, s), vid1((size_t)nv, s), vcol((size_t)nv, s), vent((size_t)nv, s), gstart((size_t)nv, s);
    scan_apply<uint32_t>(vs, VCountLoad{colptr.p}, AddU32{}, 0u, VEmit{colptr.p, perm, row.p, key0.p, vid0.p, vcol.p, vent.p}, s);
    const int w2 = radix_sort_pairs<uint64_t, uint32_t>(key0.p, vid0.p, key1.p, vid1.p, nv, 32 + bits_for(nrow + 1), s);
    const uint64_t *key = w2 ? key1.p : key0.p;
    const uint32_t *vid = w2 ? vid1.p : vid0.p;
    ScanState<uint32_t> gs, is;
    transform_scan<uint32_t>(gs, nv, GroupStartLoad{key}, MaxU32{}, 0u, s);
    scan_apply<uint32_t>(gs, GroupStartLoad{key}, MaxU32{}, 0u, GroupStartStore{key, gstart.p}, s);
    transform_scan<uint32_t>(is, nv, PairLoad{key, gstart.p, (uint32_t)pair_chunk()}, AddU32{}, 0u, s);
    uint32_t nitems = 0;
    IB_CUDA(cudaMemcpyAsync(&nitems, is.total.p, sizeof nitems, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    if (nitems == 0 || nv < 12 * (int64_t)nitems) return false;      // row pairs are not shared by enough columns
    P.pair_nitems = nitems;
    P.pair_items.alloc(nitems, s);
    P.pcol.alloc((size_t)nv, s);
    P.pv0.alloc((size_t)nv, s);
    P.pv1.alloc((size_t)nv, s);
    DBuf<uint32_t> fk0(2 * (size_t)nitems, s), fk1(2 * (size_t)nitems, s), fs0(2 * (size_t)nitems, s), fs1(2 * (size_t)nitems, s);
    scan_apply<uint32_t>(is, PairLoad{key, gstart.p, (uint32_t)pair_chunk()}, AddU32{}, 0u,
                         PairEmit{key, vid, vcol.p, vent.p, colptr.p, perm, W.val.p, nv, pair_chunk(), P.pair_items.p, P.pcol.p, P.pv0.p, P.pv1.p, fk0.p, fs0.p}, s);
    // fix-up lists: (row, slot) sorted by row, slots ascending inside a row (stable)
    const int64_t nfx = 2 * (int64_t)nitems;
    const int w3 = radix_sort_pairs<uint32_t, uint32_t>(fk0.p, fs0.p, fk1.p, fs1.p, nfx, 32, s);
    DBuf<int32_t> fxrow((size_t)nfx, s);
    u32_to_i32_copy_kernel<<<(unsigned)cdiv(nfx, 256), 256, 0, s>>>(w3 ? fk1.p : fk0.p, nfx, fxrow.p);   // 0xffffffff -> -1
    IB_LAUNCHED();
    // unused second slots carry row -1 after the cast and sort last (as 0xffffffff); count the valid prefix
    P.fx_ptr.alloc((size_t)nrow + 1, s);
    P.fx_ptr.zero();
    P.fx_slot.alloc((size_t)nfx, s);
    IB_CUDA(cudaMemcpyAsync(P.fx_slot.p, w3 ? fs1.p : fs0.p, nfx * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    std::vector<int32_t> hrow((size_t)nfx);
    fxrow.download(hrow.data(), (size_t)nfx);
    IB_CUDA(cudaStreamSynchronize(s));
    int64_t nvalid = nfx;
    while (nvalid > 0 && hrow[nvalid - 1] < 0) --nvalid;
    if (nvalid > 0) {
        seg_ptr_kernel<<<(unsigned)cdiv(nvalid, 256), 256, 0, s>>>(fxrow.p, nvalid, nrow, P.fx_ptr.p);
        IB_LAUNCHED();
    }
    IB_CUDA(cudaStreamSynchronize(s));
    P.pair = true;
    return true;
}

// Y[r] = sum of the partial rows listed for r (fixed order); rows without partials are zero
template <int VEC>
__global__ void __launch_bounds__(256) spmm_fixup_list_kernel(const int32_t *__restrict__ fx_ptr, const uint32_t *__restrict__ fx_slot,
                                                              int64_t nrow, const double *__restrict__ partials, int64_t ldp, int nvar,
                                                              double *__restrict__ Y, int64_t ldy)
{
    const int nv = nvar / VEC;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nv;
    int f = (int)(t - r * nv) * VEC;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched with programmatic stream serialization
    if (r >= nrow) return;
    const int b = fx_ptr[r], e = fx_ptr[r + 1];
    double s[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = 0.0;
    int q = b;
    for (; q + 4 <= e; q += 4) {
        double p[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double *pp = partials + (int64_t)fx_slot[q + u] * ldp + f;
            if (VEC == 2) { double2 tt = __ldcg(reinterpret_cast<const double2 *>(pp)); p[u][0] = tt.x; p[u][VEC - 1] = tt.y; }
            else p[u][0] = __ldcg(pp);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[k] = s[k] + p[u][k];
    }
    for (; q < e; ++q) {
        const double *pp = partials + (int64_t)fx_slot[q] * ldp + f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) s[k] = s[k] + __ldcg(pp + k);
    }
    double *out = Y + r * ldy + f;
    if (VEC == 2) *reinterpret_cast<double2 *>(out) = make_double2(s[0], s[VEC - 1]);
    else out[0] = s[0];
}

// ---- TMA (cp.async.bulk) pipeline -----------------------------------------------------------------------
constexpr int TMA_ROWS = 4;          // X rows per pipeline stage (amortises the barrier round trip)
constexpr int TMA_CONSUMERS = 256;   // consumer threads (8 warps); +1 producer warp
constexpr int TMA_MAX_TW = 2 * TMA_CONSUMERS;   // fields per tile

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct SpmmArgs {
    const int4 *items;
    int64_t nunits;
    int ntile, tw, nstage;
    const int32_t *rowptr, *col;
    const double *val, *val1, *X;     // val1 != nullptr: pair items (two output rows per column chunk)
    int64_t ldx;
    int nvar;
    double *Y;
    int64_t ldy;
    double *partials;
    int64_t ldp;
    int row_lo, row_hi;           // only items whose first row lies in [row_lo, row_hi) are processed (row-block sharding)
    // fused fix-up (cooperative launch only): 0 none, 1 split rows (contiguous partial slots), 2 slot lists
    int fix_mode;
    const int4 *split_rows;
    int64_t nfix;                 // rows to finalise
    const int32_t *fx_ptr;
    const uint32_t *fx_slot;
    // balanced static partition: CTA b owns units [ranges[b], ranges[b+1]) (equal numbers of nonzeros); nullptr: strided
    const int64_t *ranges;
};

// ranges[b] = first unit whose item starts at or after nonzero  total*b/nblk  (items are ordered by their first nonzero)
__global__ void unit_ranges_kernel(const int4 *items, int64_t nitems, int ntile, int64_t total, int nblk, int64_t *ranges)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblk) return;
    if (b == nblk) { ranges[b] = nitems * ntile; return; }
    const int64_t target = total * b / nblk;
    int64_t lo = 0, hi = nitems;            // first k with items[k].y >= target
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (items[mid].y >= target) hi = mid; else lo = mid + 1;
    }
    ranges[b] = lo * ntile;
}

// Shared-memory stage = [TMA_ROWS rows of X][header: TMA_ROWS values (+ TMA_ROWS second values)]; the producer
// also delivers the matrix values of the stage through the header, so consumers touch global memory only to read the
// (prefetched) item record and to write results.
constexpr uint32_t TMA_HDR = 2 * TMA_ROWS * 8;

template <bool PAIR>
__global__ void __launch_bounds__(TMA_CONSUMERS + 32) spmm_tma_kernel(const SpmmArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + a.nstage;
    unsigned char *data = smem_raw + 256;           // 2*nstage*8 <= 248
    const uint32_t rowbytes = (uint32_t)a.tw * 8u;
    const uint32_t stagebytes = rowbytes * TMA_ROWS + TMA_HDR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    uint32_t phase = 0;
    const int64_t ustep = a.ranges ? 1 : (int64_t)gridDim.x;
    int64_t u = a.ranges ? a.ranges[blockIdx.x] : (int64_t)blockIdx.x;
    const int64_t uend = a.ranges ? a.ranges[blockIdx.x + 1] : a.nunits;
    int4 it_next = u < uend ? a.items[u / a.ntile] : make_int4(0, 0, 0, 0);
    if (warp == TMA_CONSUMERS / 32) {
        // ===== producer warp: lanes 0..nr-1 each issue one row copy of the stage and post its matrix value(s)
        for (; u < uend; u += ustep) {
            const int64_t item = u / a.ntile;
            const int tile = (int)(u - item * a.ntile);
            const int4 it = it_next;
            if (u + ustep < uend) it_next = a.items[(u + ustep) / a.ntile];
            if (it.x < a.row_lo || it.x >= a.row_hi) continue;
            const int f0 = tile * a.tw;
            const uint32_t tb = (uint32_t)min(a.tw, a.nvar - f0) * 8u;
            for (int j = it.y; j < it.z; j += TMA_ROWS) {
                const int nr = min(TMA_ROWS, it.z - j);
                int c = 0;
                double v = 0.0, v1 = 0.0;
                if (lane < nr) {
                    c = __ldg(a.col + j + lane);
                    v = __ldg(a.val + j + lane);
                    if (PAIR) v1 = __ldg(a.val1 + j + lane);
                }
                if (lane == 0) mbar_wait(empty + stage, phase ^ 1u);
                __syncwarp();
                unsigned char *sp = data + (size_t)stage * stagebytes;
                if (lane < TMA_ROWS) {
                    double *hdr = reinterpret_cast<double *>(sp + (size_t)TMA_ROWS * rowbytes);
                    hdr[lane] = v;
                    if (PAIR) hdr[TMA_ROWS + lane] = v1;
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(full + stage, tb * (uint32_t)nr);      // arrive (release) after the header stores
                __syncwarp();
                if (lane < nr) bulk_g2s(sp + (size_t)lane * rowbytes, a.X + (int64_t)c * a.ldx + f0, tb, full + stage);
                if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== consumer warps: thread t owns fields f0 + 2t, f0 + 2t + 1
        const int t = threadIdx.x;
        for (; u < uend; u += ustep) {
            const int64_t item = u / a.ntile;
            const int tile = (int)(u - item * a.ntile);
            const int4 it = it_next;
            if (u + ustep < uend) it_next = a.items[(u + ustep) / a.ntile];
            if (it.x < a.row_lo || it.x >= a.row_hi) continue;
            const int f0 = tile * a.tw;
            const bool active = 2 * t < min(a.tw, a.nvar - f0);
            const int nrows = it.w < 0 ? -it.w : 0;
            // row ends of a group live one per lane (nrows <= 32)
            int my_end = (lane < nrows) ? __ldg(a.rowptr + it.x + lane + 1) : 0;
            int cur = 0;
            int cur_end = nrows ? __shfl_sync(0xffffffffu, my_end, 0) : it.z;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            int j = it.y;
            // leading empty rows of a group
            while (cur < nrows && cur_end == j) {
                if (active) reinterpret_cast<double2 *>(a.Y + (int64_t)(it.x + cur) * a.ldy + f0)[t] = make_double2(0.0, 0.0);
                ++cur;
                cur_end = __shfl_sync(0xffffffffu, my_end, cur & 31);
            }
            for (int jb = it.y; jb < it.z; jb += TMA_ROWS) {
                const int nr = min(TMA_ROWS, it.z - jb);
                mbar_wait(full + stage, phase);
                const unsigned char *sb = data + (size_t)stage * stagebytes;
                const double *hdr = reinterpret_cast<const double *>(sb + (size_t)TMA_ROWS * rowbytes);
                const unsigned char *sp = sb + (size_t)t * 16;
#pragma unroll
                for (int r = 0; r < TMA_ROWS; ++r) {
                    if (r < nr) {
                        if (active) {
                            const double2 x = *reinterpret_cast<const double2 *>(sp + (size_t)r * rowbytes);
                            const double v = hdr[r];
                            a0 = a0 + v * x.x;
                            a1 = a1 + v * x.y;
                            if (PAIR) {
                                const double w = hdr[TMA_ROWS + r];
                                b0 = b0 + w * x.x;
                                b1 = b1 + w * x.y;
                            }
                        }
                        ++j;
                        while (cur < nrows && cur_end == j) {      // row(s) of a group complete
                            if (active) reinterpret_cast<double2 *>(a.Y + (int64_t)(it.x + cur) * a.ldy + f0)[t] = make_double2(a0, a1);
                            a0 = 0.0; a1 = 0.0;
                            ++cur;
                            cur_end = __shfl_sync(0xffffffffu, my_end, cur & 31);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
            }
            if (it.w >= 0 && active) {  // chunk: partial sum(s), combined in a fixed order by the fix-up
                reinterpret_cast<double2 *>(a.partials + (int64_t)it.w * a.ldp + f0)[t] = make_double2(a0, a1);
                if (PAIR) reinterpret_cast<double2 *>(a.partials + (int64_t)(it.w + 1) * a.ldp + f0)[t] = make_double2(b0, b1);
            }
        }
    }
    if (a.fix_mode == 0) return;
    // ---- fused fix-up: the grid is co-resident (cooperative launch), so one grid-wide barrier replaces a kernel
    // boundary; afterwards every thread sums the partial rows of some (row, field pair) in the fixed slot order
    __threadfence();
    cooperative_groups::this_grid().sync();
    const int nv2 = a.nvar / 2;
    const int64_t total = a.nfix * nv2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / nv2;
        const int f = (int)(idx - r * nv2) * 2;
        double2 acc = make_double2(0.0, 0.0);
        int64_t orow;
        if (a.fix_mode == 1) {
            const int4 sr = a.split_rows[r];
            orow = sr.x;
            if (sr.x < a.row_lo || sr.x >= a.row_hi) continue;
            const double *pp = a.partials + (int64_t)sr.y * a.ldp + f;
            int q = 0;
            for (; q + 4 <= sr.z; q += 4) {
                double2 p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] = __ldcg(reinterpret_cast<const double2 *>(pp + (int64_t)(q + k) * a.ldp));
#pragma unroll
                for (int k = 0; k < 4; ++k) { acc.x = acc.x + p[k].x; acc.y = acc.y + p[k].y; }
            }
            for (; q < sr.z; ++q) {
                const double2 p = __ldcg(reinterpret_cast<const double2 *>(pp + (int64_t)q * a.ldp));
                acc.x = acc.x + p.x; acc.y = acc.y + p.y;
            }
        } else {
            orow = r;
            const int b = a.fx_ptr[r], e = a.fx_ptr[r + 1];
            int q = b;
            for (; q + 4 <= e; q += 4) {
                double2 p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] = __ldcg(reinterpret_cast<const double2 *>(a.partials + (int64_t)a.fx_slot[q + k] * a.ldp + f));
#pragma unroll
                for (int k = 0; k < 4; ++k) { acc.x = acc.x + p[k].x; acc.y = acc.y + p[k].y; }
            }
            for (; q < e; ++q) {
                const double2 p = __ldcg(reinterpret_cast<const double2 *>(a.partials + (int64_t)a.fx_slot[q] * a.ldp + f));
                acc.x = acc.x + p.x; acc.y = acc.y + p.y;
            }
        }
        *reinterpret_cast<double2 *>(a.Y + orow * a.ldy + f) = acc;
    }
}

// ---- X-stationary kernel: tall matrices with few columns (IvE, IvA, smoothed IvE) ------------------------------
// When the input space is small (E or A cells: ~10^3) and the output space large (ice cells: 10^5..10^7), a field slice
// of ALL of X fits in shared memory: Xs[ncol][F] with F = 8..64 fields.  A CTA keeps one slice resident and streams
// matrix rows through it: thread (row, p) walks its row's nonzeros, reads X from shared memory and writes one double2
// of Y -- X costs no L2 traffic per nonzero (the row-streaming kernel re-fetches a 4 KB X row per nonzero, which is what
// made IvE 0.42 and the 16-nnz/row smoothed IvE L2-bound).  The price is that the matrix is re-read once per slice.
// Rows are never split: every row is summed over ascending columns with mul-then-add, i.e. bit-identical to the oracle.
struct XsArgs {
    const int32_t *rowptr, *col;
    const double *val, *X;
    int64_t ldx;
    double *Y;
    int64_t ldy;
    int64_t nrow;
    int ncol, nvar, F, nslice;
    int64_t nchunk, nunits;
    int chunk_rows;
};

constexpr int XS_THREADS = 1024;
constexpr int XS_K = 4;          // rows in flight per thread (independent load chains)

__global__ void __launch_bounds__(XS_THREADS) spmm_xs_kernel(const XsArgs a)
{
    extern __shared__ __align__(16) unsigned char xs_raw[];
    double2 *Xs = reinterpret_cast<double2 *>(xs_raw);
    const int TPR = a.F >> 1;                       // threads (double2) per row
    const int rpp = blockDim.x / TPR;               // rows per pass
    const int p = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    // contiguous share of the (slice-major) unit list: a CTA changes slice at most a couple of times
    const int64_t u0 = a.nunits * blockIdx.x / gridDim.x, u1 = a.nunits * (blockIdx.x + 1) / gridDim.x;
    int cur_slice = -1;
    for (int64_t u = u0; u < u1; ++u) {
        const int slice = (int)(u / a.nchunk);
        const int64_t chunk = u - (int64_t)slice * a.nchunk;
        const int f0 = slice * a.F;
        const int fw = min(a.F, a.nvar - f0) >> 1;  // live double2 per row in this slice
        if (slice != cur_slice) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < a.ncol * TPR; idx += blockDim.x) {
                const int c = idx / TPR, q = idx - c * TPR;
                Xs[idx] = q < fw ? __ldg(reinterpret_cast<const double2 *>(a.X + (int64_t)c * a.ldx + f0) + q) : make_double2(0.0, 0.0);
            }
            __syncthreads();
            cur_slice = slice;
        }
        const int64_t r_end = min(a.nrow, (chunk + 1) * a.chunk_rows);
        for (int64_t rb = chunk * a.chunk_rows + rl; rb < r_end; rb += (int64_t)rpp * XS_K) {
            int jb[XS_K], je[XS_K];
#pragma unroll
            for (int k = 0; k < XS_K; ++k) {
                const int64_t r = rb + (int64_t)k * rpp;
                const bool ok = r < r_end;
                jb[k] = ok ? __ldg(a.rowptr + r) : 0;
                je[k] = ok ? __ldg(a.rowptr + r + 1) : 0;
            }
            // first two nonzeros of every row up front (IvE / IvA rows have 1..4): independent loads
            int c0[XS_K], c1[XS_K];
            double v0[XS_K], v1[XS_K];
#pragma unroll
            for (int k = 0; k < XS_K; ++k) {
                const bool h0 = jb[k] < je[k], h1 = jb[k] + 1 < je[k];
                c0[k] = h0 ? __ldg(a.col + jb[k]) : 0;
                v0[k] = h0 ? __ldg(a.val + jb[k]) : 0.0;
                c1[k] = h1 ? __ldg(a.col + jb[k] + 1) : 0;
                v1[k] = h1 ? __ldg(a.val + jb[k] + 1) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < XS_K; ++k) {
                const int64_t r = rb + (int64_t)k * rpp;
                if (r >= r_end) continue;
                double a0 = 0.0, a1 = 0.0;
                if (jb[k] < je[k]) { const double2 x = Xs[c0[k] * TPR + p]; a0 = a0 + v0[k] * x.x; a1 = a1 + v0[k] * x.y; }
                if (jb[k] + 1 < je[k]) { const double2 x = Xs[c1[k] * TPR + p]; a0 = a0 + v1[k] * x.x; a1 = a1 + v1[k] * x.y; }
                int j = jb[k] + 2;
                for (; j + 4 <= je[k]; j += 4) {
                    int cc[4];
                    double vv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { cc[q] = __ldg(a.col + j + q); vv[q] = __ldg(a.val + j + q); }
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const double2 x = Xs[cc[q] * TPR + p]; a0 = a0 + vv[q] * x.x; a1 = a1 + vv[q] * x.y; }
                }
                for (; j < je[k]; ++j) {
                    const int cq = __ldg(a.col + j);
                    const double vq = __ldg(a.val + j);
                    const double2 x = Xs[cq * TPR + p];
                    a0 = a0 + vq * x.x; a1 = a1 + vq * x.y;
                }
                if (p < fw) reinterpret_cast<double2 *>(a.Y + r * a.ldy + f0)[p] = make_double2(a0, a1);
            }
        }
    }
}

// ---- distinct-column group kernel: short-row matrices (IvE, IvA, EvA) --------------------------------------------
// Unit = (item of gc_plan.cuh, tile of 64*NQ fields), owned by ONE WARP; lane l owns the double2 at fields
// tile + 64q + 2l, q < NQ.  The X rows of the item's <= GC_S distinct columns are fetched ONCE, by 1-D bulk async copies
// (cp.async.bulk, one per row, issued by lanes 0..nd-1) into a warp-private shared-memory slot table, together with the
// item's records (value, slot, flags: 16 bytes per nonzero, prepared once per matrix); tables are double buffered: the
// next unit lands (completing on a warp-private mbarrier) while the current one is processed.  The inner loop is a flat
// walk over the records: broadcast-read a record, read the X row of its slot from shared memory, multiply-add, and on
// the last record of a row store the row.  Record reads run two records ahead and X reads one record ahead of the
// arithmetic.  Warps never synchronise with each other.
// History: (1) one double2 per lane, values/slots broadcast by shuffles: bound by instruction issue (58 warp
// instructions per nonzero and 64 fields); (2) NQ double2 per lane: 160 instructions per row and tile, every one
// waiting ~6 cycles on the previous (3 shuffles + convergence checks per nonzero); (3) this one.
// X traffic drops from one row per nonzero to one per distinct column (IvE 0.15x, IvA 0.06x at 5 km), which is what
// bounded the row-streaming kernels: they ran at the L2 fabric rate (reads + writes ~ 8.4 TB/s), not at HBM's.
// Rows are never split and are summed over ascending columns with mul-then-add: bit-identical to the oracle.
struct GcArgs {
    const int4 *items;
    const int32_t *ucol;
    const int4 *recs;
    int64_t nitems;
    int ntile;
    const double *X;
    int64_t ldx;
    int nvar;
    double *Y;
    int64_t ldy;
    int row_lo, row_hi;
    unsigned long long *counter;     // [0] units handed out so far (never reset), [1] its value when this launch began, [2] warps done
};
constexpr int GC_WARPS = 2;      // warps per CTA (independent; a CTA is only a container for shared memory)
constexpr int GC_RECBUF = (GC_MAXREC + 2) * 16;     // record table of one buffer (+2: the prefetch overruns)

template <int NQ>
constexpr size_t gc_smem_bytes() { return (size_t)GC_WARPS * (2 * (GC_S * 64 * NQ * 8 + GC_RECBUF) + 16); }

template <int NQ>
__global__ void __launch_bounds__(GC_WARPS * 32) spmm_gc_kernel(const GcArgs a)
{
    constexpr int TW = 64 * NQ;                        // fields per tile
    constexpr uint32_t ROWB = TW * 8;                  // bytes per slot
    constexpr uint32_t BUFB = GC_S * ROWB + GC_RECBUF; // one buffer: slot table + record table
    extern __shared__ __align__(128) unsigned char gc_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = gc_raw + (size_t)warp * (2 * BUFB);
    uint64_t *bar = reinterpret_cast<uint64_t *>(gc_raw + (size_t)GC_WARPS * (2 * BUFB)) + 2 * warp;
    const unsigned FULL = 0xffffffffu;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t nunits = a.nitems * a.ntile;

    // Units are handed out dynamically (items differ in size: 1..32 rows): unit u = item * ntile + tile.  The counter is
    // read three units ahead of the arithmetic, so its latency is never waited for.
    // the first three units of a warp are static; the counter hands out the rest.  It is never reset: the last warp
    // to leave records where the next launch starts (so a captured graph can be replayed)
    const int64_t nw = (int64_t)gridDim.x * GC_WARPS;
    const unsigned long long cbase = a.counter[1] - 3ull * (unsigned long long)nw;
    auto take_unit = [&]() -> unsigned long long {          // lane 0 holds the answer; broadcast where it is used
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(a.counter, 1ull) - cbase;
        return u;
    };
    auto bcast = [&](unsigned long long u) -> int64_t { return (int64_t)__shfl_sync(FULL, u, 0); };
    // record of a unit: the item and (lanes 0..GC_S-1) its distinct columns.  Nothing here waits for a load: the
    // row-block test is applied where the record is used.
    auto load_item = [&](int64_t u, int4 &it, int &uc) {
        it = make_int4(0, 0, 0, 0);
        uc = -1;
        if (u < nunits) {
            const int64_t item = (uint32_t)u / (uint32_t)a.ntile;        // nunits < 2^31 (checked by the host)
            it = __ldg(a.items + item);
            if (lane < GC_S) uc = __ldg(a.ucol + item * GC_S + lane);
        }
    };
    auto in_block = [&](const int4 &it) { return it.x >= a.row_lo && it.x < a.row_hi; };
    // X rows and records of a unit -> buffer b.  The buffer was last read by this warp two units ago (program order +
    // __syncwarp); units outside the row block fetch nothing.
    auto issue = [&](int64_t u, const int4 &it, int uc, int b) {
        const bool ok = in_block(it);
        const int f0 = (int)((uint32_t)u % (uint32_t)a.ntile) * TW;
        const uint32_t tb = (uint32_t)min(TW, a.nvar - f0) * 8u;
        const uint32_t rb = ok ? (uint32_t)(it.z - it.y) * 16u : 0u;
        const bool mine = ok && uc >= 0;                     // lanes >= GC_S carry uc = -1
        const int nd = __popc(__ballot_sync(FULL, mine));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of the buffer before async writes
        __syncwarp();
        unsigned char *dst = buf + (size_t)b * BUFB;
        if (lane == 0) {
            if (rb + nd > 0) mbar_expect_tx(bar + b, tb * (uint32_t)nd + rb);
            else mbar_arrive(bar + b);
        }
        __syncwarp();
        if (mine) bulk_g2s(dst + (size_t)lane * ROWB, a.X + (int64_t)uc * a.ldx + f0, tb, bar + b);
        if (lane == GC_S && rb > 0) bulk_g2s(dst + GC_S * ROWB, a.recs + it.y, rb, bar + b);
    };

    int64_t u_c = (int64_t)blockIdx.x * GC_WARPS + warp, u_n = u_c + nw, u_n2 = u_n + nw;
    unsigned long long raw_n3 = take_unit();
    int4 itc, itn, itn2;
    int ucc, ucn, ucn2;
    load_item(u_c, itc, ucc);
    load_item(u_n, itn, ucn);
    int b = 0;
    uint32_t phase = 0;                                      // bit b: parity to wait for on bar[b]
    if (u_c < nunits) issue(u_c, itc, ucc, 0);
    while (u_c < nunits) {
        // prefetch: tables of the next unit, item record of the one after, unit number of the one after that
        if (u_n < nunits) issue(u_n, itn, ucn, b ^ 1);
        load_item(u_n2, itn2, ucn2);
        const int64_t u_n3 = bcast(raw_n3);
        raw_n3 = take_unit();
        // this unit's tables have landed?
        mbar_wait(bar + b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const int n = in_block(itc) ? itc.z - itc.y : 0;
        const int f0 = (int)((uint32_t)u_c % (uint32_t)a.ntile) * TW;
        const int nq = min(NQ, (a.nvar - f0 + 63) >> 6);            // live 64-field groups of this tile
        const bool last_live = f0 + 64 * (nq - 1) + 2 * lane < a.nvar;   // ragged last group
        bool live[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) live[q] = q < nq - 1 || (q == nq - 1 && last_live);
        const double2 *xb = reinterpret_cast<const double2 *>(buf + (size_t)b * BUFB) + lane;
        const int4 *rec = reinterpret_cast<const int4 *>(buf + (size_t)b * BUFB + GC_S * ROWB);
        double2 *yp = reinterpret_cast<double2 *>(a.Y + (int64_t)itc.x * a.ldy + f0) + lane;
        const int64_t ystep = a.ldy / 2;
        double2 acc[NQ], xA[NQ], xB[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = make_double2(0.0, 0.0);
        auto fetch = [&](double2(&x)[NQ], int slot) {
            const double2 *xr = xb + (slot & (GC_S - 1)) * (TW / 2);
#pragma unroll
            for (int q = 0; q < NQ; ++q) x[q] = xr[q * 32];
        };
        // branch-free: the row sum so far goes out through its own registers when the record closes its row (the
        // accumulators are not written while a store still reads them), and the accumulators restart from zero
        auto step = [&](const int4 &r, const double2(&x)[NQ]) {
            const double v = __hiloint2double(r.y, r.x);
            const bool last = (r.w & (int)GC_LAST) != 0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double2 t;
                t.x = acc[q].x + v * x[q].x;
                t.y = acc[q].y + v * x[q].y;
                if (last && live[q]) yp[q * 32] = t;
                acc[q].x = last ? 0.0 : t.x;
                acc[q].y = last ? 0.0 : t.y;
            }
            yp += last ? ystep : 0;
        };
        // flat walk over the records, two per trip: a record names the slot of its successor, so the successor's X row
        // is fetched (and the successor read) while the record itself is processed
        int4 rA = rec[0], rB;
        fetch(xA, rA.z);
        for (int k = 0; k < n; k += 2) {
            rB = rec[k + 1];
            fetch(xB, rA.z >> 8);
            step(rA, xA);
            if (k + 1 >= n) break;
            rA = rec[k + 2];
            fetch(xA, rB.z >> 8);
            step(rB, xB);
        }
        __syncwarp();
        itc = itn; itn = itn2;
        ucn = ucn2;
        u_c = u_n; u_n = u_n2; u_n2 = u_n3;
        b ^= 1;
    }
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(a.counter + 2, 1ull) == (unsigned long long)nw - 1) {      // every other warp has taken its last unit
            a.counter[1] = atomicAdd(a.counter, 0ull);
            a.counter[2] = 0;
        }
    }
}

// ---- per-warp streaming kernel for the pair plan (EvI) -------------------------------------------------------------
// The CTA-wide ring of spmm_tma_kernel is consumer-bound: with fewer CTAs per SM it slows down in proportion (one CTA per
// SM: 0.41 of peak), so its eight consumer warps, not HBM, set the pace.  Here every warp is its own pipeline, as in
// spmm_gc_kernel: a warp owns one tile of 64*NQ fields and a CONTIGUOUS run of pair items (equal numbers of columns per
// warp group, BalanceTable), walks the regrouped column stream record by record across item boundaries, and keeps a ring
// of PW_R X-row pieces in flight: after a record is consumed its slot is refilled with the piece PW_R records ahead (one
// bulk async copy issued by lane 0, completing on the slot's own mbarrier).  Column ids and the two values of a column
// are staged through a small double-buffered shared-memory table, 32 records at a time, loaded one block ahead.  The
// warps of a CTA work on the same items (one tile each), so the pieces of an X row are requested together.
struct PwArgs {
    const int4 *items;            // pair items (r0, begin, end, first partial slot)
    const int64_t *ranges;        // [ngroup + 1] item ranges of the warp groups
    int ngroup, ntile;
    const int32_t *col;
    const double *v0, *v1, *X;
    int64_t ldx;
    int nvar;
    double *partials;
    int64_t ldp;
};
constexpr int PW_B = 4;           // records per batch (one mbarrier round trip per batch)
constexpr int PW_NB = 3;          // batches in the ring
constexpr int PW_R = PW_B * PW_NB;
constexpr int PW_WARPS = 4;       // warps per CTA

template <int NQ>
constexpr size_t pw_smem_bytes() { return (size_t)PW_WARPS * (PW_R * 64 * NQ * 8 + 2 * 32 * 24 + 64); }

template <int NQ>
__global__ void __launch_bounds__(PW_WARPS * 32) spmm_pw_kernel(const PwArgs a)
{
    constexpr int TW = 64 * NQ;
    constexpr uint32_t ROWB = TW * 8;
    constexpr uint32_t WARPB = PW_R * ROWB + 2 * 32 * 24 + 64;
    extern __shared__ __align__(128) unsigned char pw_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = pw_raw + (size_t)warp * WARPB;
    unsigned char *ring = base;
    double *tv0 = reinterpret_cast<double *>(base + PW_R * ROWB);        // [2][32]
    double *tv1 = tv0 + 64;                                              // [2][32]
    int32_t *tcol = reinterpret_cast<int32_t *>(tv1 + 64);               // [2][32]
    uint64_t *bar = reinterpret_cast<uint64_t *>(base + PW_R * ROWB + 2 * 32 * 24);      // [PW_NB]
    if (lane < PW_NB) mbar_init(bar + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int64_t gw = (int64_t)blockIdx.x * PW_WARPS + warp;
    const int group = (int)(gw / a.ntile), tile = (int)(gw % a.ntile);
    if (group >= a.ngroup) return;
    const int64_t ib = a.ranges[group], ie = a.ranges[group + 1];
    if (ib >= ie) return;
    const int f0 = tile * TW;
    const uint32_t tb = (uint32_t)min(TW, a.nvar - f0) * 8u;
    const int nq = min(NQ, (a.nvar - f0 + 63) >> 6);
    const bool last_live = f0 + 64 * (nq - 1) + 2 * lane < a.nvar;
    int4 it = __ldg(a.items + ib);
    const int64_t P0 = it.y;
    const int64_t n = __ldg(a.items + ie - 1).z - P0;                  // records of this warp
    const int64_t nb = (n + PW_B - 1) / PW_B;
    int4 it_next = ib + 1 < ie ? __ldg(a.items + ib + 1) : it;
    int64_t item = ib;

    // record blocks of 32: lane l holds record 32*j + l of block j
    auto load_block = [&](int64_t j, int &c, double &x0, double &x1) {
        const int64_t k = 32 * j + lane;
        c = 0; x0 = 0.0; x1 = 0.0;
        if (k < n) { c = __ldg(a.col + P0 + k); x0 = __ldg(a.v0 + P0 + k); x1 = __ldg(a.v1 + P0 + k); }
    };
    auto store_block = [&](int64_t j, int c, double x0, double x1) {
        const int o = (int)(j & 1) * 32 + lane;
        tcol[o] = c; tv0[o] = x0; tv1[o] = x1;
    };
    // X-row pieces of batch b -> ring batch b % PW_NB: lane r < PW_B fetches record PW_B*b + r
    auto issue_batch = [&](int64_t b) {
        const int bs = (int)(b % PW_NB);
        const int64_t k = PW_B * b + lane;
        const int cnt = (int)min((int64_t)PW_B, n - PW_B * b);
        if (lane == 0) mbar_expect_tx(bar + bs, tb * (uint32_t)cnt);
        __syncwarp();
        if (lane < cnt) {
            const int c = tcol[(int)((k >> 5) & 1) * 32 + (int)(k & 31)];
            bulk_g2s(ring + (size_t)(bs * PW_B + lane) * ROWB, a.X + (int64_t)c * a.ldx + f0, tb, bar + bs);
        }
    };
    int rc; double r0v, r1v;
    load_block(0, rc, r0v, r1v);
    store_block(0, rc, r0v, r1v);
    load_block(1, rc, r0v, r1v);
    __syncwarp();
    for (int64_t b = 0; b < nb && b < PW_NB; ++b) issue_batch(b);
    uint32_t phase = 0;
    double2 acc0[NQ], acc1[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) { acc0[q] = make_double2(0.0, 0.0); acc1[q] = make_double2(0.0, 0.0); }
    const double2 *xb = reinterpret_cast<const double2 *>(ring) + lane;
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t k0 = PW_B * b;
        if ((k0 & 31) == 0) {           // entering block j: publish block j+1 (its table half was block j-1's), load block j+2
            const int64_t j = k0 >> 5;
            __syncwarp();
            store_block(j + 1, rc, r0v, r1v);
            load_block(j + 2, rc, r0v, r1v);
            __syncwarp();
        }
        const int bs = (int)(b % PW_NB);
        mbar_wait(bar + bs, (phase >> bs) & 1u);
        phase ^= 1u << bs;
        // all table and ring reads of the batch first, then the arithmetic record by record
        const int o = (int)((k0 >> 5) & 1) * 32 + (int)(k0 & 31);          // PW_B divides 32: a batch never straddles blocks
        double v0[PW_B], v1[PW_B];
        double2 x[PW_B][NQ];
#pragma unroll
        for (int r = 0; r < PW_B; ++r) {
            v0[r] = tv0[o + r];
            v1[r] = tv1[o + r];
            const double2 *xr = xb + (bs * PW_B + r) * (TW / 2);
#pragma unroll
            for (int q = 0; q < NQ; ++q) x[r][q] = xr[q * 32];
        }
#pragma unroll
        for (int r = 0; r < PW_B; ++r) {
            if (k0 + r < n) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    acc0[q].x = acc0[q].x + v0[r] * x[r][q].x;
                    acc0[q].y = acc0[q].y + v0[r] * x[r][q].y;
                    acc1[q].x = acc1[q].x + v1[r] * x[r][q].x;
                    acc1[q].y = acc1[q].y + v1[r] * x[r][q].y;
                }
                if (P0 + k0 + r + 1 == it.z) {      // item complete: its two partial rows (summed in a fixed order by the fix-up)
                    double2 *p0 = reinterpret_cast<double2 *>(a.partials + (int64_t)it.w * a.ldp + f0) + lane;
                    double2 *p1 = reinterpret_cast<double2 *>(a.partials + (int64_t)(it.w + 1) * a.ldp + f0) + lane;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        if (q < nq - 1 || (q == nq - 1 && last_live)) { p0[q * 32] = acc0[q]; p1[q * 32] = acc1[q]; }
                        acc0[q] = make_double2(0.0, 0.0);
                        acc1[q] = make_double2(0.0, 0.0);
                    }
                    ++item;
                    it = it_next;
                    if (item + 1 < ie) it_next = __ldg(a.items + item + 1);
                }
            }
        }
        // refill the ring batch with the pieces PW_NB batches ahead (the reads above are complete: their values were used)
        if (b + PW_NB < nb) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            issue_batch(b + PW_NB);
        }
    }
}

// ---- plain-load fallback (small / odd nvar, unaligned buffers) ---------------------------------------------
// A group of L lanes (power of two) owns one (item, field-tile) unit; lane fl owns VEC consecutive fields.
template <int VEC>
__global__ void __launch_bounds__(256) spmm_ldg_kernel(const int4 *__restrict__ items, int64_t nunits, int ntile, int L,
                                                       const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                       const double *__restrict__ val, const double *__restrict__ X, int64_t ldx,
                                                       int nvar, double *__restrict__ Y, int64_t ldy, double *__restrict__ partials,
                                                       int64_t ldp, int row_lo, int row_hi)
{
    const int lane = threadIdx.x & 31;
    const int G = 32 / L;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int fl = lane % L;
    for (int64_t u = gwarp * G + lane / L; u < nunits; u += nwarps * G) {
        const int64_t item = u / ntile;
        const int tile = (int)(u - item * ntile);
        const int f0 = (tile * L + fl) * VEC;
        if (f0 >= nvar) continue;
        const int4 it = items[item];
        if (it.x < row_lo || it.x >= row_hi) continue;
        const int nrows = it.w < 0 ? -it.w : 1;
        for (int rr = 0; rr < nrows; ++rr) {
            const int jb = it.w < 0 ? __ldg(rowptr + it.x + rr) : it.y;
            const int je = it.w < 0 ? __ldg(rowptr + it.x + rr + 1) : it.z;
            double acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = 0.0;
            int j = jb;
            for (; j + 4 <= je; j += 4) {
                int c[4];
                double v[4];
                double x[4][VEC];
#pragma unroll
                for (int q = 0; q < 4; ++q) { c[q] = __ldg(col + j + q); v[q] = __ldg(val + j + q); }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double *xp = X + (int64_t)c[q] * ldx + f0;
                    if (VEC == 2) { double2 tt = __ldg(reinterpret_cast<const double2 *>(xp)); x[q][0] = tt.x; x[q][VEC - 1] = tt.y; }
                    else x[q][0] = __ldg(xp);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = acc[k] + v[q] * x[q][k];
            }
            for (; j < je; ++j) {
                const int c = __ldg(col + j);
                const double v = __ldg(val + j);
                const double *xp = X + (int64_t)c * ldx + f0;
                if (VEC == 2) { double2 tt = __ldg(reinterpret_cast<const double2 *>(xp)); acc[0] = acc[0] + v * tt.x; acc[VEC - 1] = acc[VEC - 1] + v * tt.y; }
                else acc[0] = acc[0] + v * __ldg(xp);
            }
            double *out = (it.w < 0) ? (Y + (int64_t)(it.x + rr) * ldy + f0) : (partials + (int64_t)it.w * ldp + f0);
            if (VEC == 2) *reinterpret_cast<double2 *>(out) = make_double2(acc[0], acc[VEC - 1]);
            else out[0] = acc[0];
        }
    }
}

// Y[row] = sum of the row's chunk partials, added in chunk order (deterministic)
template <int VEC>
__global__ void __launch_bounds__(256) spmm_fixup_kernel(const int4 *__restrict__ split_rows, int64_t nsplit,
                                                         const double *__restrict__ partials, int64_t ldp, int nvar,
                                                         double *__restrict__ Y, int64_t ldy, int row_lo, int row_hi)
{
    const int nv = nvar / VEC;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nv;
    int f = (int)(t - r * nv) * VEC;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched with programmatic stream serialization
    if (r >= nsplit) return;
    const int4 sr = split_rows[r];
    if (sr.x < row_lo || sr.x >= row_hi) return;
    const double *pp = partials + (int64_t)sr.y * ldp + f;
    double s[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = 0.0;   // 0 + p0 == p0 exactly
    int q = 0;
    for (; q + 8 <= sr.z; q += 8) {
        double p[8][VEC];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (VEC == 2) { double2 tt = __ldcg(reinterpret_cast<const double2 *>(pp + (int64_t)(q + u) * ldp)); p[u][0] = tt.x; p[u][VEC - 1] = tt.y; }
            else p[u][0] = __ldcg(pp + (int64_t)(q + u) * ldp);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[k] = s[k] + p[u][k];
    }
    for (; q < sr.z; ++q)
#pragma unroll
        for (int k = 0; k < VEC; ++k) s[k] = s[k] + __ldcg(pp + (int64_t)q * ldp + k);
    double *out = Y + (int64_t)sr.x * ldy + f;
    if (VEC == 2) *reinterpret_cast<double2 *>(out) = make_double2(s[0], s[VEC - 1]);
    else out[0] = s[0];
}

// ---- conservation correction (ibmisc Weighted_Eigen::apply_e; matrix_formats.rst:156-165) ---------------
// out[b][f] = sum over rows r of block b's slice of w[r] * A[r][f]
__global__ void __launch_bounds__(256) wdot_partial_kernel(const double *__restrict__ w, const double *__restrict__ A, int64_t lda,
                                                           int64_t nrow, int nvar, int64_t rows_per_block, double *__restrict__ out)
{
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(nrow, r0 + rows_per_block);
    for (int f = threadIdx.x; f < nvar; f += blockDim.x) {
        double s = 0;
        for (int64_t r = r0; r < r1; ++r) s = s + w[r] * A[r * lda + f];
        out[(int64_t)blockIdx.x * nvar + f] = s;
    }
}
__global__ void wdot_final_kernel(const double *__restrict__ pa, int nba, const double *__restrict__ pb, int nbb, int nvar,
                                  double *__restrict__ factor)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nvar) return;
    double ta = 0, tb = 0;
    for (int b = 0; b < nba; ++b) ta = ta + pa[(int64_t)b * nvar + f];
    for (int b = 0; b < nbb; ++b) tb = tb + pb[(int64_t)b * nvar + f];
    factor[f] = ta * (1. / tb);
}
__global__ void scale_copy_kernel(const double *__restrict__ X, int64_t ldx, double *__restrict__ Xs, int64_t ldxs, int64_t nrow, int nvar,
                                  const double *__restrict__ factor)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nvar;
    int f = (int)(t - r * nvar);
    if (r < nrow) Xs[r * ldxs + f] = X[r * ldx + f] * factor[f];
}
// v[c] = sum_r M[r,c] * wM[r], columns walked in CSC order (rows ascending): deterministic
__global__ void cons_v_kernel(const int32_t *colptr, const uint32_t *perm, const int32_t *row, const double *val, const double *wM,
                              int64_t ncol, double *v)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    double s = 0;
    for (int32_t j = colptr[c]; j < colptr[c + 1]; ++j) { const uint32_t e = perm[j]; s += val[e] * wM[row[e]]; }
    v[c] = s;
}
__global__ void expand_rows_kernel(const int32_t *rowptr, int64_t nrow, int32_t *row);
__global__ void u32_to_i32_copy_kernel(const uint32_t *a, int64_t n, int32_t *b);

static void compute_cons_v(Weighted &W, cudaStream_t s)
{
    if (W.cons_v.n == (size_t)W.ncol && W.ncol > 0) return;
    const int64_t nnz = W.nnz, ncol = W.ncol, nrow = W.nrow;
    W.cons_v.alloc((size_t)ncol, s);
    W.cons_v.zero();
    if (nnz == 0) return;
    DBuf<int32_t> row((size_t)nnz, s), csorted((size_t)nnz, s), colptr((size_t)ncol + 1, s);
    DBuf<uint32_t> k0((size_t)nnz, s), k1((size_t)nnz, s), p0((size_t)nnz, s), p1((size_t)nnz, s);
    expand_rows_kernel<<<(unsigned)cdiv(nrow, 128), 128, 0, s>>>(W.rowptr.p, nrow, row.p);
    IB_LAUNCHED();
    IB_CUDA(cudaMemcpyAsync(k0.p, W.col.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    iota_u32_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(p0.p, nnz);
    IB_LAUNCHED();
    const int w1 = radix_sort_pairs<uint32_t, uint32_t>(k0.p, p0.p, k1.p, p1.p, nnz, bits_for(ncol), s);
    u32_to_i32_copy_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(w1 ? k1.p : k0.p, nnz, csorted.p);
    IB_LAUNCHED();
    colptr.zero();
    seg_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(csorted.p, nnz, ncol, colptr.p);
    IB_LAUNCHED();
    cons_v_kernel<<<(unsigned)cdiv(ncol, 128), 128, 0, s>>>(colptr.p, w1 ? p1.p : p0.p, row.p, W.val.p, W.wM.p, ncol, W.cons_v.p);
    IB_LAUNCHED();
}

static void wdot(const double *w, const double *A, int64_t lda, int64_t nrow, int nvar, DBuf<double> &part, int &nb, cudaStream_t s)
{
    nb = (int)std::min<int64_t>(std::max<int64_t>(1, cdiv(nrow, 64)), 4 * sm_count());
    int64_t rpb = cdiv(std::max<int64_t>(nrow, 1), nb);
    nb = (int)cdiv(std::max<int64_t>(nrow, 1), rpb);
    part.alloc((size_t)nb * nvar, s);
    wdot_partial_kernel<<<nb, 256, 0, s>>>(w, A, lda, nrow, nvar, rpb, part.p);
    IB_LAUNCHED();
}

// Launch a fix-up kernel so that its launch overlaps the tail of the SpMM kernel before it (programmatic dependent
// launch); the kernel itself waits with griddepcontrol.wait before touching the partial sums.
template <class... KArgs, class... Args>
static void launch_dependent(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IB_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    IB_LAUNCHED();
}

void apply_dev(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, bool force_conservation,
               cudaStream_t s, int64_t row_begin, int64_t row_end)
{
    ensure_device();
    if (nvar <= 0 || W.nrow == 0) return;
    if (row_end < 0) row_end = W.nrow;
    const bool ranged = !(row_begin == 0 && row_end == W.nrow);
    if (ranged) {
        // row-block sharding (output rows split across GPUs, inputs replicated): groups of short rows never cross a
        // multiple of GROUP_ROWS, so blocks aligned to it own whole items
        if (row_begin < 0 || row_end > W.nrow || row_begin > row_end || row_begin % GROUP_ROWS || (row_end % GROUP_ROWS && row_end != W.nrow))
            fail(IB_ERR_ARG, "apply: row block [%ld, %ld) must be aligned to %d rows", (long)row_begin, (long)row_end, GROUP_ROWS);
        if (!W.conservative && force_conservation)
            fail(IB_ERR_ARG, "apply: the conservation correction needs all rows; apply row blocks with force_conservation=0");
        if (row_begin == row_end) return;
    }
    const int rlo = (int)row_begin, rhi = (int)row_end;
    // ---- conservation correction (ibmisc Weighted_Eigen::apply_e; matrix_formats.rst:156-165), fused as a prologue:
    //   B = B0 * (Mw.A)/(wM.B0)  with  wM.(M A) = (M^T wM).A =: v.A ,
    // so both dot products run over the (small) input fields, the factor is folded into a scaled copy of A, and the
    // large output is written once instead of being re-read twice.
    DBuf<double> Xscaled;
    if (!W.conservative && force_conservation) {
        compute_cons_v(W, s);
        DBuf<double> pa, pb, factor((size_t)nvar, s);
        int nba, nbb;
        wdot(W.Mw.p, dX, ldx, W.ncol, nvar, pa, nba, s);
        wdot(W.cons_v.p, dX, ldx, W.ncol, nvar, pb, nbb, s);
        wdot_final_kernel<<<(unsigned)cdiv(nvar, 128), 128, 0, s>>>(pa.p, nba, pb.p, nbb, nvar, factor.p);
        IB_LAUNCHED();
        const int64_t ldxs = (nvar + 1) & ~1;
        Xscaled.alloc((size_t)W.ncol * ldxs, s);
        const int64_t tot = W.ncol * (int64_t)nvar;
        scale_copy_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, s>>>(dX, ldx, Xscaled.p, ldxs, W.ncol, nvar, factor.p);
        IB_LAUNCHED();
        dX = Xscaled.p;
        ldx = ldxs;
    }
    if (ldx < nvar || ldy < nvar) fail(IB_ERR_ARG, "apply: leading dimension smaller than nvar");
    make_plan(W, s);
    ApplyPlan &P = W.plan;
    const bool vec2 = (nvar % 2 == 0) && (ldx % 2 == 0) && (ldy % 2 == 0) && ((uintptr_t)dX % 16 == 0) && ((uintptr_t)dY % 16 == 0);
    const int64_t ldp = (nvar + 1) & ~1;
    DBuf<double> partials((size_t)P.npartials * ldp, s);

    static const int variant = getenv("IB_SPMM_VARIANT") ? atoi(getenv("IB_SPMM_VARIANT")) : 0;   // 0 auto, 1 LDG, 2 TMA
    static const int want_bps = getenv("IB_SPMM_BPS") ? atoi(getenv("IB_SPMM_BPS")) : 4;            // TMA blocks per SM
    bool done = false, pair_done = false;
    static int max_smem_xs = 0;
    if (!max_smem_xs) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_smem_xs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        IB_CUDA(cudaFuncSetAttribute(spmm_xs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_xs));
    }
    // distinct-column groups: every row has <= GC_S nonzeros (IvE, IvA, EvA); variant 5 forces, others forbid
    static const int gc_nq = getenv("IB_GC_NQ") ? atoi(getenv("IB_GC_NQ")) : 0;             // 0: by nvar
    static const int gc_wps = getenv("IB_GC_WPS") ? atoi(getenv("IB_GC_WPS")) : 0;          // warps per SM (0: all that fit)
    if (P.gc && vec2 && nvar >= 32 && (variant == 0 || variant == 5) && P.gc_nitems * (int64_t)((nvar + 63) / 64) < (int64_t(1) << 31)) {
        static bool attr_set = false;
        if (!attr_set) {
            IB_CUDA(cudaFuncSetAttribute(spmm_gc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes<4>()));
            IB_CUDA(cudaFuncSetAttribute(spmm_gc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes<2>()));
            IB_CUDA(cudaFuncSetAttribute(spmm_gc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes<1>()));
            attr_set = true;
        }
        int nq = nvar <= 64 ? 1 : 2;
        if (gc_nq == 1 || gc_nq == 2 || gc_nq == 4) nq = gc_nq;
        const int tw = 64 * nq;
        const int ntile = (nvar + tw - 1) / tw;
        GcArgs a{P.gc_items.p, P.gc_ucol.p, P.gc_recs.p, P.gc_nitems, ntile, dX, ldx, nvar, dY, ldy, rlo, rhi, P.gc_counter.p};
        const size_t smem = nq == 4 ? gc_smem_bytes<4>() : nq == 2 ? gc_smem_bytes<2>() : gc_smem_bytes<1>();
        auto kern = nq == 4 ? spmm_gc_kernel<4> : nq == 2 ? spmm_gc_kernel<2> : spmm_gc_kernel<1>;
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GC_WARPS * 32, smem);
        occ = std::max(occ, 1);
        if (gc_wps > 0) occ = std::max(1, std::min(occ, gc_wps / GC_WARPS));
        const int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(cdiv(a.nitems * ntile, GC_WARPS), (int64_t)sm_count() * occ));
        if (a.nitems > 0) {
            kern<<<(unsigned)nblk, GC_WARPS * 32, smem, s>>>(a);
            IB_LAUNCHED();
        }
        return;
    }
    // X-stationary path: tall matrix, few columns, no long rows (variant 3 forces it, 4 forbids it)
    if (vec2 && !ranged && variant != 1 && variant != 2 && variant != 4 && W.ncol > 0 && W.nrow >= 8 * W.ncol && P.max_row_len <= 256 && nvar >= 8 &&
        (variant == 3 || 2 * W.nnz >= 3 * W.nrow)) {
        int F = 64;
        while (F > nvar) F >>= 1;
        while (F >= 8 && (size_t)W.ncol * F * 8 > (size_t)max_smem_xs - 1024) F >>= 1;
        if (F >= 8) {
            const size_t smem = (size_t)W.ncol * F * 8;
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_xs_kernel, XS_THREADS, smem);
            occ = std::max(1, std::min(occ, 4));
            XsArgs a{W.rowptr.p, W.col.p, W.val.p, dX, ldx, dY, ldy, W.nrow, (int)W.ncol, nvar, F, (nvar + F - 1) / F, 0, 0, 2048};
            a.nchunk = cdiv(W.nrow, a.chunk_rows);
            a.nunits = a.nchunk * a.nslice;
            const int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(a.nunits, (int64_t)sm_count() * occ));
            spmm_xs_kernel<<<(unsigned)nblk, XS_THREADS, smem, s>>>(a);
            IB_LAUNCHED();
            done = true;
            pair_done = true;      // nothing to fix up
        }
    }
    if (!done && vec2 && nvar >= 64 && variant != 1) {
        const int tw = std::min(nvar, TMA_MAX_TW);
        const int ntile = (nvar + tw - 1) / tw;
        const int64_t nunits = P.nitems * ntile;
        static int max_smem = 0;
        if (!max_smem) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            IB_CUDA(cudaFuncSetAttribute(spmm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
            IB_CUDA(cudaFuncSetAttribute(spmm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        }
        const size_t stagebytes = (size_t)tw * 8 * TMA_ROWS + TMA_HDR;
        const size_t budget = (size_t)(max_smem + 1024) / want_bps - 1024 - 256;
        const int nstage = (int)std::min<size_t>(15, budget / stagebytes);
        if (nstage >= 2) {
            const size_t smem = 256 + (size_t)nstage * stagebytes;
            // measured on B200: the cooperative launch + grid barrier costs a little more than the PDL-launched fix-up kernel
            static const bool fuse = getenv("IB_SPMM_FUSE") != nullptr;
            // true residency (registers, shared memory) bounds the grid of a cooperative launch
            int occ0 = 0, occ1 = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, spmm_tma_kernel<false>, TMA_CONSUMERS + 32, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, spmm_tma_kernel<true>, TMA_CONSUMERS + 32, smem);
            static const bool balance = getenv("IB_SPMM_NOBALANCE") == nullptr;
            auto launch = [&](auto kern, int occ, SpmmArgs a, int64_t nun, bool needs_fix, int64_t total_work, BalanceTable *bt) -> bool {
                const int64_t slots = (int64_t)sm_count() * std::min(want_bps, std::max(occ, 1));
                const int64_t per = std::max<int64_t>(1, cdiv(nun, slots));
                int64_t nblk = std::max<int64_t>(1, cdiv(nun, per));
                if (balance && bt && !ranged && total_work > 0) {
                    // CTAs own contiguous runs of units holding equal numbers of nonzeros (items differ in size, and with
                    // 2-4 items per CTA a strided deal leaves some CTAs a whole item short)
                    nblk = std::max<int64_t>(1, std::min<int64_t>(nun / a.ntile, slots));
                    if (bt->nblk != nblk || bt->ntile != a.ntile) {
                        bt->ranges.alloc((size_t)nblk + 1, s);
                        unit_ranges_kernel<<<(unsigned)cdiv(nblk + 1, 128), 128, 0, s>>>(a.items, nun / a.ntile, a.ntile, total_work, (int)nblk, bt->ranges.p);
                        IB_LAUNCHED();
                        bt->nblk = nblk; bt->ntile = a.ntile;
                    }
                    a.ranges = bt->ranges.p;
                }
                if (needs_fix && fuse && occ >= 1 && nblk <= (int64_t)sm_count() * occ) {
                    void *kargs[] = {&a};
                    cudaError_t e = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)nblk), dim3(TMA_CONSUMERS + 32), kargs, smem, s);
                    if (e == cudaSuccess) { IB_LAUNCHED(); return true; }
                    cudaGetLastError();      // fall through to the two-kernel path
                }
                a.fix_mode = 0;
                kern<<<(unsigned)nblk, TMA_CONSUMERS + 32, smem, s>>>(a);
                IB_LAUNCHED();
                return false;
            };
            if (P.pair && !ranged) {
                const int64_t nun = P.pair_nitems * ntile;
                DBuf<double> pp((size_t)(2 * P.pair_nitems) * ldp, s);
                SpmmArgs a{P.pair_items.p, nun, ntile, tw, nstage, W.rowptr.p, P.pcol.p, P.pv0.p, P.pv1.p, dX, ldx, nvar, dY, ldy, pp.p, ldp,
                           0, 0x7fffffff, 2, nullptr, W.nrow, P.fx_ptr.p, P.fx_slot.p, nullptr};
                static const bool skipfix = getenv("IB_SPMM_SKIPFIX") != nullptr;      // timing experiments only
                static const int pw = getenv("IB_SPMM_PW") ? atoi(getenv("IB_SPMM_PW")) : 0;       // per-warp streaming kernel
                bool main_done = false;
                if (pw > 0 && nvar >= 64) {
                    static bool pw_attr = false;
                    if (!pw_attr) {
                        IB_CUDA(cudaFuncSetAttribute(spmm_pw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw_smem_bytes<2>()));
                        IB_CUDA(cudaFuncSetAttribute(spmm_pw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw_smem_bytes<1>()));
                        pw_attr = true;
                    }
                    const int nq = pw == 1 ? 1 : 2;
                    const int ptile = (nvar + 64 * nq - 1) / (64 * nq);
                    const size_t psmem = nq == 2 ? pw_smem_bytes<2>() : pw_smem_bytes<1>();
                    auto pk = nq == 2 ? spmm_pw_kernel<2> : spmm_pw_kernel<1>;
                    int pocc = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, pk, PW_WARPS * 32, psmem);
                    pocc = std::max(pocc, 1);
                    const int64_t nwarps = (int64_t)sm_count() * pocc * PW_WARPS;
                    const int64_t ngroup = std::max<int64_t>(1, std::min<int64_t>(P.pair_nitems, nwarps / ptile));
                    BalanceTable &bt = P.bal_pw;
                    if (bt.nblk != ngroup || bt.ntile != 1) {
                        bt.ranges.alloc((size_t)ngroup + 1, s);
                        unit_ranges_kernel<<<(unsigned)cdiv(ngroup + 1, 128), 128, 0, s>>>(P.pair_items.p, P.pair_nitems, 1, (int64_t)P.pcol.n, (int)ngroup, bt.ranges.p);
                        IB_LAUNCHED();
                        bt.nblk = ngroup; bt.ntile = 1;
                    }
                    PwArgs pa{P.pair_items.p, bt.ranges.p, (int)ngroup, ptile, P.pcol.p, P.pv0.p, P.pv1.p, dX, ldx, nvar, pp.p, ldp};
                    pk<<<(unsigned)cdiv(ngroup * ptile, PW_WARPS), PW_WARPS * 32, psmem, s>>>(pa);
                    IB_LAUNCHED();
                    main_done = true;
                }
                if (main_done) {
                    if (!skipfix) {
                        const int64_t tot = W.nrow * (int64_t)(nvar / 2);
                        launch_dependent(spmm_fixup_list_kernel<2>, (unsigned)cdiv(tot, 256), 256, 0, s, P.fx_ptr.p, P.fx_slot.p, W.nrow, pp.p, ldp, nvar, dY, ldy);
                    }
                } else if (!launch(spmm_tma_kernel<true>, occ1, a, nun, true, (int64_t)P.pcol.n, &P.bal_pair) && !skipfix) {
                    const int64_t tot = W.nrow * (int64_t)(nvar / 2);
                    launch_dependent(spmm_fixup_list_kernel<2>, (unsigned)cdiv(tot, 256), 256, 0, s, P.fx_ptr.p, P.fx_slot.p, W.nrow, pp.p, ldp, nvar, dY, ldy);
                }
                pair_done = true;
            } else {
                SpmmArgs a{P.items.p, nunits, ntile, tw, nstage, W.rowptr.p, W.col.p, W.val.p, nullptr, dX, ldx, nvar, dY, ldy, partials.p, ldp,
                           rlo, rhi, 1, P.split_rows.p, P.nsplit_rows, nullptr, nullptr, nullptr};
                // equal-size chunks: the strided deal is already even (and measured 3 % faster than contiguous runs on AvI)
                static const bool balance_rows = getenv("IB_SPMM_BALANCE_ROWS") != nullptr;
                if (launch(spmm_tma_kernel<false>, occ0, a, nunits, P.nsplit_rows > 0, W.nnz, balance_rows ? &P.bal_rows : nullptr)) pair_done = true;   // fix-up already done
            }
            done = true;
        }
    }
    if (!done) {
        const int VEC = vec2 ? 2 : 1;
        const int lanes_needed = (nvar + VEC - 1) / VEC;
        int L = 1;
        while (L < 32 && L < lanes_needed) L <<= 1;
        const int ntile = (lanes_needed + L - 1) / L;
        const int64_t nunits = P.nitems * ntile;
        const int G = 32 / L;
        int64_t blocks = std::min<int64_t>(cdiv(cdiv(nunits, G), 8), (int64_t)sm_count() * 8);
        if (blocks < 1) blocks = 1;
        if (VEC == 2)
            spmm_ldg_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(P.items.p, nunits, ntile, L, W.rowptr.p, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, rlo, rhi);
        else
            spmm_ldg_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(P.items.p, nunits, ntile, L, W.rowptr.p, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, rlo, rhi);
        IB_LAUNCHED();
    }
    if (P.nsplit_rows > 0 && !pair_done && !getenv("IB_SPMM_SKIPFIX")) {
        if (vec2) {
            int64_t tot = P.nsplit_rows * (nvar / 2);
            launch_dependent(spmm_fixup_kernel<2>, (unsigned)cdiv(tot, 256), 256, 0, s, P.split_rows.p, P.nsplit_rows, partials.p, ldp, nvar, dY, ldy, rlo, rhi);
        } else {
            int64_t tot = P.nsplit_rows * (int64_t)nvar;
            launch_dependent(spmm_fixup_kernel<1>, (unsigned)cdiv(tot, 256), 256, 0, s, P.split_rows.p, P.nsplit_rows, partials.p, ldp, nvar, dY, ldy, rlo, rhi);
        }
    }
}

}   // namespace ib
