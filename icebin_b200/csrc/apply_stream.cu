// apply_stream.cu -- one pass over the fields for WIDE matrices:  [Y1; Y2] = [M1; M2] * X.
//
// The ice -> GCM half of a coupling step applies EvI and AvI to the SAME ice-grid fields
// (IceCoupler.cpp:384-458: gcm_ivalsE = E1vI * ice_ovalsI', gcm_ivalsA = A1vI * ice_ovalsI'); both matrices have a
// few nonzeros per COLUMN (an ice cell feeds the two elevation classes that bracket it, IceRegridder_L0.cpp:126-143,
// and the one GCM cell it lies in) and few rows, so the whole step is a weighted segmented reduction of the rows of
// X: 282 MB in, 5 MB out at the headline size.  Two separate SpMMs read X twice; this file reads it once.
//
// Plan (built once per matrix pair on the device).  Every column c of the stacked matrix [M1; M2] is cut into
// "records" (c, r0, r1, r2, v0, v1, v2): two consecutive nonzeros of M1's column and one of M2's (an ice cell on the
// border between GCM cells yields two to four records).  Records are sorted by (r2, r0, r1), so the records that feed
// the same output rows are contiguous, and the sorted stream is cut into R equal runs, one per CTA.  A CTA walks its
// run with three running sums in registers (rows r0, r1, r2); flags in the record say when a sum is complete
// ("close") and whether the sum of row r1 carries over as the next record's r0.  A closed sum is one "event": it
// goes straight to Y when the row has no other event, else to a partial-sum slot; the warp that delivers the last
// slot of a row adds the row's slots in slot order (= stream order, fixed at plan time: deterministic, no matter
// which warp does it) and writes Y.  No second kernel, no atomics on data.
//
// Kernel (spmm_stream_kernel): the ring of spmm_tma_kernel -- one producer warp issues one 1-D bulk async copy
// (cp.async.bulk, SASS UBLKCP) per record into a shared-memory ring of stages x 4 rows and posts the record's values
// and flags in the stage header; the consumer warps (64 fields each, one double2 per lane) never synchronise with
// each other: each one arrives on the stage's `empty` barrier by itself and runs the event protocol by itself
// (completion counters are per row and 64-field strip).  Launched with programmatic stream serialization: the
// prologue (barrier init, first record loads) overlaps the tail of the previous kernel of the stream, and nothing is
// read from X or written before griddepcontrol.wait.
//
// Compiled with -fmad=false.  A row whose records all lie in one run and that is never carried is the plain
// left-to-right sum over its columns in SORTED RECORD order, not ascending column order, so results agree with the
// oracle to rounding (1e-12 gate), not bit for bit.
//
// Roofline: HBM.  Algorithmic bytes = 12*(nnz1+nnz2) + 8*nvar*(ncol_touched + nrow1 + nrow2)   (X counted once).
#include <algorithm>
#include <mutex>

#include "radix_sort.cuh"
#include "scan.cuh"
#include "tma.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

constexpr int ST_ROWS = 4;              // records (X rows) per pipeline stage
constexpr int ST_HDR = 64 * ST_ROWS;     // stage header: the stage's records, verbatim
constexpr int ST_MAXN1 = 16, ST_MAXN2 = 8;   // nonzeros per column the plan accepts (M1, M2)
// record flags
constexpr uint32_t SF_CLOSE = 1u;       // bits 0..2: accumulator k restarts from zero after this record
constexpr uint32_t SF_SHIFT = 8u;       // bit 3: after the closes, accumulator 1 becomes accumulator 0
constexpr uint32_t SF_EVENT = 16u;      // bits 4..6: accumulator k is delivered (an event) before it restarts

// 64 bytes: values, column, flags and the (row, partial slot or -1) of the up to three sums that end with the record
struct alignas(16) StRec { double v0, v1, v2; int32_t col; uint32_t flags; int2 ev[3]; int2 pad; };
static_assert(sizeof(StRec) == 64, "StRec is four int4");

__host__ __device__ inline int64_t run_begin(int64_t n, int64_t R, int64_t b) { return n * b / R; }

// ---- plan ------------------------------------------------------------------------------------------------
struct StackView {
    const int32_t *rp1, *col1;
    const double *val1;
    int64_t nrow1, nnz1;
    const int32_t *rp2, *col2;
    const double *val2;
    int64_t nrow2, nnz2;
    __device__ __forceinline__ int32_t col(int64_t e) const { return e < nnz1 ? col1[e] : col2[e - nnz1]; }
    __device__ __forceinline__ double val(int64_t e) const { return e < nnz1 ? val1[e] : val2[e - nnz1]; }
    __device__ __forceinline__ int32_t row(int64_t e) const      // stacked row of entry e
    {
        const int32_t *rp = e < nnz1 ? rp1 : rp2;
        const int64_t nrow = e < nnz1 ? nrow1 : nrow2;
        const int32_t j = (int32_t)(e < nnz1 ? e : e - nnz1);
        int64_t lo = 0, hi = nrow;                  // largest r with rp[r] <= j
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (rp[mid] <= j) lo = mid; else hi = mid;
        }
        return (int32_t)(e < nnz1 ? lo : lo + nrow1);
    }
};

__global__ void st_count_kernel(StackView v, uint32_t *cnt1, uint32_t *cnt2)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= v.nnz1 + v.nnz2) return;
    atomicAdd((e < v.nnz1 ? cnt1 : cnt2) + v.col(e), 1u);
}
// per column: x = entries (colptr), y = records (vbase), z = max n1, w = max n2
struct StColOp {
    __host__ __device__ uint4 operator()(uint4 a, uint4 b) const
    {
        return make_uint4(a.x + b.x, a.y + b.y, a.z > b.z ? a.z : b.z, a.w > b.w ? a.w : b.w);
    }
};
struct StColLoad {
    const uint32_t *cnt1, *cnt2;
    __device__ uint4 operator()(int64_t c) const
    {
        const uint32_t n1 = cnt1[c], n2 = cnt2[c];
        const uint32_t h = (n1 + 1) / 2;
        return make_uint4(n1 + n2, h > n2 ? h : n2, n1, n2);
    }
};
struct StColStore {
    uint32_t *colptr, *vbase;
    __device__ void operator()(int64_t c, uint4 prefix, uint4) const { colptr[c] = prefix.x; vbase[c] = prefix.y; }
};
// entries of a column in arrival order (atomics); made deterministic by the sort of (row, entry) keys in st_emit_kernel
__global__ void st_scatter_kernel(StackView v, const uint32_t *colptr, uint32_t *cursor, uint64_t *ent)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= v.nnz1 + v.nnz2) return;
    const int32_t c = v.col(e);
    const uint32_t pos = colptr[c] + atomicAdd(cursor + c, 1u);
    ent[pos] = ((uint64_t)(uint32_t)v.row(e) << 32) | (uint32_t)e;
}
// one thread per column: order the column's entries by (stacked row, entry), cut them into records
__global__ void st_emit_kernel(StackView v, int64_t ncol, const uint32_t *colptr, const uint32_t *vbase, const uint32_t *cnt1,
                               const uint32_t *cnt2, const uint64_t *ent, int b1, uint64_t *ukey, uint32_t *uvid, StRec *urec)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int n1 = (int)cnt1[c], n2 = (int)cnt2[c], n = n1 + n2;
    if (n == 0) return;
    uint64_t k[ST_MAXN1 + ST_MAXN2];
    const uint64_t *src = ent + colptr[c];
    for (int i = 0; i < n; ++i) {               // insertion sort (n is 3 for almost every column)
        const uint64_t x = src[i];
        int j = i;
        while (j > 0 && k[j - 1] > x) { k[j] = k[j - 1]; --j; }
        k[j] = x;
    }
    const int nv = max((n1 + 1) / 2, n2);
    const uint32_t base = vbase[c];
    for (int i = 0; i < nv; ++i) {
        StRec r;
        r.col = (int32_t)c;
        r.flags = 0;
        r.v0 = r.v1 = r.v2 = 0.0;
        r.ev[0] = r.ev[1] = r.ev[2] = r.pad = make_int2(-1, -1);
        uint64_t r0 = 0, r1 = 0, r2 = 0;        // row + 1; 0 = none
        if (2 * i < n1) { r0 = (k[2 * i] >> 32) + 1; r.v0 = v.val((uint32_t)k[2 * i]); }
        if (2 * i + 1 < n1) { r1 = (k[2 * i + 1] >> 32) + 1; r.v1 = v.val((uint32_t)k[2 * i + 1]); }
        if (i < n2) { r2 = (k[n1 + i] >> 32) - (uint64_t)v.nrow1 + 1; r.v2 = v.val((uint32_t)k[n1 + i]); }
        ukey[base + i] = (r2 << (2 * b1)) | (r0 << b1) | r1;
        uvid[base + i] = base + i;
        urec[base + i] = r;
    }
}

// decode a sorted key into stacked rows (-1 = none)
struct StRows { int32_t r0, r1, r2; };
__device__ __forceinline__ StRows st_rows(uint64_t key, int b1, int32_t nrow1)
{
    const uint64_t m = (uint64_t(1) << b1) - 1;
    StRows r;
    r.r1 = (int32_t)(key & m) - 1;
    r.r0 = (int32_t)((key >> b1) & m) - 1;
    const int32_t a = (int32_t)(key >> (2 * b1)) - 1;
    r.r2 = a < 0 ? -1 : a + nrow1;
    return r;
}
struct StFlagLoad {
    const uint64_t *key;
    int64_t n, R;
    int b1;
    int32_t nrow1;
    __device__ uint32_t flags(int64_t p) const
    {
        const StRows c = st_rows(key[p], b1, nrow1);
        // last record of its run?  run_of(p) = ceil((p+1)*R/n) - 1
        const bool last = (p + 1 == n) || ((p + 2) * R + n - 1) / n != ((p + 1) * R + n - 1) / n;
        uint32_t f = 0;
        if (last) {
            f = 7u * SF_CLOSE;
            if (c.r0 >= 0) f |= SF_EVENT;
            if (c.r1 >= 0) f |= SF_EVENT << 1;
            if (c.r2 >= 0) f |= SF_EVENT << 2;
            return f;
        }
        const StRows x = st_rows(key[p + 1], b1, nrow1);
        const bool sh = c.r1 >= 0 && x.r0 == c.r1 && x.r0 != c.r0;
        if (x.r0 != c.r0) { f |= SF_CLOSE; if (c.r0 >= 0) f |= SF_EVENT; }
        if (sh) f |= SF_SHIFT;
        else if (x.r1 != c.r1) { f |= SF_CLOSE << 1; if (c.r1 >= 0) f |= SF_EVENT << 1; }
        if (x.r2 != c.r2) { f |= SF_CLOSE << 2; if (c.r2 >= 0) f |= SF_EVENT << 2; }
        return f;
    }
    __device__ uint32_t operator()(int64_t p) const { return (uint32_t)__popc(flags(p) >> 4); }
};
struct StFlagEmit {
    StFlagLoad L;
    const uint32_t *vid;
    const StRec *urec;
    StRec *rec;
    int32_t *cols;
    uint32_t *evrow, *evpos;        // event e: its row; where it lives (record * 4 + accumulator)
    __device__ void operator()(int64_t p, uint32_t prefix, uint32_t) const
    {
        const uint32_t f = L.flags(p);
        StRec r = urec[vid[p]];
        r.flags = f;
        rec[p] = r;
        cols[p] = r.col;
        const StRows c = st_rows(L.key[p], L.b1, L.nrow1);
        uint32_t e = prefix;
        if (f & SF_EVENT) { evrow[e] = (uint32_t)c.r0; evpos[e] = (uint32_t)p * 4u; ++e; }
        if (f & (SF_EVENT << 1)) { evrow[e] = (uint32_t)c.r1; evpos[e] = (uint32_t)p * 4u + 1u; ++e; }
        if (f & (SF_EVENT << 2)) { evrow[e] = (uint32_t)c.r2; evpos[e] = (uint32_t)p * 4u + 2u; ++e; }
    }
};
// events sorted by row: q-th sorted event; x = needs a partial slot, y = row without any event before it ... (zero rows)
struct StSlotLoad {
    const int32_t *srow, *fx_ptr;
    __device__ uint32_t operator()(int64_t q) const
    {
        const int32_t r = srow[q];
        return (fx_ptr[r + 1] - fx_ptr[r] > 1) ? 1u : 0u;
    }
};
struct StSlotEmit {
    const int32_t *srow, *fx_ptr;
    const uint32_t *spos;           // sorted event q lives at record spos[q] / 4, accumulator spos[q] % 4
    StRec *rec;
    int2 *rowinfo;                  // (first slot, number of slots) of every row that has more than one sum
    __device__ void operator()(int64_t q, uint32_t prefix, uint32_t multi) const
    {
        // slots are numbered in this (row-sorted, stream order inside a row) order: a row's slots are consecutive
        const int32_t r = srow[q];
        const uint32_t pos = spos[q];
        rec[pos >> 2].ev[pos & 3u] = make_int2(r, multi ? (int32_t)prefix : -1);
        if (multi && q == fx_ptr[r]) rowinfo[r] = make_int2((int32_t)prefix, fx_ptr[r + 1] - fx_ptr[r]);
    }
};
// rows the fix-up kernel visits: more than one sum (add the slots) or none at all (zero)
struct StFixLoad {
    const int32_t *fx_ptr;
    __device__ uint32_t operator()(int64_t r) const { return fx_ptr[r + 1] - fx_ptr[r] != 1 ? 1u : 0u; }
};
struct StFixEmit {
    const int32_t *fx_ptr;
    int32_t *fixrows;
    int2 *rowinfo;
    __device__ void operator()(int64_t r, uint32_t prefix, uint32_t z) const
    {
        if (!z) return;
        fixrows[prefix] = (int32_t)r;
        if (fx_ptr[r + 1] == fx_ptr[r]) rowinfo[r] = make_int2(-1, 0);
    }
};
static __global__ void st_u32_to_i32_kernel(const uint32_t *a, int64_t n, int32_t *b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)a[i];
}

static std::atomic<uint64_t> g_uid{1};
uint64_t next_uid() { return g_uid.fetch_add(1); }

// Builds the plan for [M1; M2] (M2 may be null) cut into R runs; returns false (and remembers it) when the pair is not
// a good fit: columns with many nonzeros, keys that do not fit 64 bits, or outputs that are not small against inputs.
static bool make_stream_plan(StreamPlan &P, Weighted &M1, Weighted *M2, int R, cudaStream_t s)
{
    const uint64_t puid = M2 ? M2->uid : 0;
    if (P.tried && P.partner_uid == puid && (!P.ok || P.R_asked == R)) return P.ok;
    P = StreamPlan();
    P.tried = true;
    P.partner_uid = puid;
    P.R = P.R_asked = R;
    const int64_t ncol = M1.ncol, nnz1 = M1.nnz, nnz2 = M2 ? M2->nnz : 0, nnz = nnz1 + nnz2;
    const int64_t nrow1 = M1.nrow, nrow2 = M2 ? M2->nrow : 0, nrowS = nrow1 + nrow2;
    const int b1 = bits_for(nrow1 + 2), b2 = bits_for(nrow2 + 2);
    if (ncol <= 0 || nnz1 < 4096 || nnz >= (int64_t(1) << 31) || nrowS >= (int64_t(1) << 30) || 2 * b1 + b2 > 64) return false;
    if (nnz > 6 * ncol || 16 * nrowS > nnz) return false;          // wide: few nonzeros per column, few rows
    StackView v{M1.rowptr.p, M1.col.p, M1.val.p, nrow1, nnz1, M2 ? M2->rowptr.p : nullptr, M2 ? M2->col.p : nullptr,
                M2 ? M2->val.p : nullptr, nrow2, nnz2};
    DBuf<uint32_t> cnt((size_t)3 * ncol, s), colptr((size_t)ncol, s), vbase((size_t)ncol, s);
    cnt.zero();
    uint32_t *cnt1 = cnt.p, *cnt2 = cnt.p + ncol, *cursor = cnt.p + 2 * ncol;
    st_count_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(v, cnt1, cnt2);
    IB_LAUNCHED();
    ScanState<uint4> cs;
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    transform_scan<uint4>(cs, ncol, StColLoad{cnt1, cnt2}, StColOp{}, z4, s);
    uint4 tot;
    IB_CUDA(cudaMemcpyAsync(&tot, cs.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    const int64_t nrec = tot.y;
    if ((int)tot.z > ST_MAXN1 || (int)tot.w > ST_MAXN2 || nrec < 4096) return false;
    R = (int)std::max<int64_t>(1, std::min<int64_t>(R, nrec / (4 * ST_ROWS)));
    P.R = R;
    scan_apply<uint4>(cs, StColLoad{cnt1, cnt2}, StColOp{}, z4, StColStore{colptr.p, vbase.p}, s);
    DBuf<uint64_t> ent((size_t)nnz, s), key0((size_t)nrec, s), key1((size_t)nrec, s);
    DBuf<uint32_t> vid0((size_t)nrec, s), vid1((size_t)nrec, s);
    DBuf<StRec> urec((size_t)nrec, s);
    st_scatter_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(v, colptr.p, cursor, ent.p);
    IB_LAUNCHED();
    st_emit_kernel<<<(unsigned)cdiv(ncol, 128), 128, 0, s>>>(v, ncol, colptr.p, vbase.p, cnt1, cnt2, ent.p, b1, key0.p, vid0.p, urec.p);
    IB_LAUNCHED();
    const int w1 = radix_sort_pairs<uint64_t, uint32_t>(key0.p, vid0.p, key1.p, vid1.p, nrec, 2 * b1 + b2, s);
    const uint64_t *key = w1 ? key1.p : key0.p;
    const uint32_t *vid = w1 ? vid1.p : vid0.p;
    // flags, events
    StFlagLoad FL{key, nrec, R, b1, (int32_t)nrow1};
    ScanState<uint32_t> es;
    transform_scan<uint32_t>(es, nrec, FL, AddU32{}, 0u, s);
    uint32_t nev = 0;
    IB_CUDA(cudaMemcpyAsync(&nev, es.total.p, sizeof nev, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    P.nrec = nrec;
    P.nev = nev;
    P.recs.alloc((size_t)nrec * 4, s);
    P.cols.alloc((size_t)nrec, s);
    if (nrec * 4 >= (int64_t(1) << 32)) return false;
    DBuf<uint32_t> er0((size_t)nev + 1, s), er1((size_t)nev + 1, s), ei0((size_t)nev + 1, s), ei1((size_t)nev + 1, s);
    scan_apply<uint32_t>(es, FL, AddU32{}, 0u, StFlagEmit{FL, vid, urec.p, reinterpret_cast<StRec *>(P.recs.p), P.cols.p, er0.p, ei0.p}, s);
    // per-row slot lists: events sorted by row, stream order inside a row (the sort is stable)
    const int w2 = radix_sort_pairs<uint32_t, uint32_t>(er0.p, ei0.p, er1.p, ei1.p, nev, bits_for(nrowS), s);
    DBuf<int32_t> srow((size_t)nev + 1, s), fx_ptr((size_t)nrowS + 1, s);
    fx_ptr.zero();
    P.rowinfo.alloc((size_t)nrowS, s);
    if (nev > 0) {
        st_u32_to_i32_kernel<<<(unsigned)cdiv(nev, 256), 256, 0, s>>>(w2 ? er1.p : er0.p, nev, srow.p);
        IB_LAUNCHED();
        seg_ptr_kernel<<<(unsigned)cdiv(nev, 256), 256, 0, s>>>(srow.p, nev, nrowS, fx_ptr.p);
        IB_LAUNCHED();
    }
    ScanState<uint32_t> ss, zs;
    transform_scan<uint32_t>(ss, nev, StSlotLoad{srow.p, fx_ptr.p}, AddU32{}, 0u, s);
    scan_apply<uint32_t>(ss, StSlotLoad{srow.p, fx_ptr.p}, AddU32{}, 0u,
                         StSlotEmit{srow.p, fx_ptr.p, w2 ? ei1.p : ei0.p, reinterpret_cast<StRec *>(P.recs.p), P.rowinfo.p}, s);
    transform_scan<uint32_t>(zs, nrowS, StFixLoad{fx_ptr.p}, AddU32{}, 0u, s);
    uint32_t nslots = 0, nfix = 0;
    IB_CUDA(cudaMemcpyAsync(&nslots, ss.total.p, sizeof nslots, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaMemcpyAsync(&nfix, zs.total.p, sizeof nfix, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    P.nslots = nslots;
    P.nfix = nfix;
    P.fixrows.alloc((size_t)nfix + 1, s);
    if (nfix) scan_apply<uint32_t>(zs, StFixLoad{fx_ptr.p}, AddU32{}, 0u, StFixEmit{fx_ptr.p, P.fixrows.p, P.rowinfo.p}, s);
    IB_CUDA(cudaStreamSynchronize(s));      // fx_ptr and the other temporaries die here
    P.nrow1 = nrow1;
    P.nrowS = nrowS;
    P.ok = true;
    return true;
}

// ---- kernel ----------------------------------------------------------------------------------------------
#ifdef IB_STREAM_TRACE      // tools/build_trace.py: per-CTA time stamps of the last launch (never in the product library)
constexpr int TR_N = 8;
__device__ unsigned long long g_stream_trace[8192 * TR_N];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(slot) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 8192) g_stream_trace[blockIdx.x * TR_N + (slot)] = gtime(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif
#ifdef IB_STEP_TRACE        // tools/build_variants.py: span of every launch (ring of 4096), for the multi-GPU step timeline
__device__ unsigned long long g_st_steps[4096 * 4];      // released (min), main end (max), fix-up released (min), fix-up end (max)
__device__ unsigned g_st_arrive[2];
__device__ __forceinline__ unsigned long long st_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define STEP_BEGIN(which, idxvar) unsigned idxvar = 0; if (threadIdx.x == 0) { idxvar = (atomicAdd(&g_st_arrive[which], 1u) / gridDim.x) & 4095u; \
    atomicMin(&g_st_steps[idxvar * 4 + 2 * (which)], st_now()); }
#define STEP_END(which, idxvar) do { if (threadIdx.x == 0) atomicMax(&g_st_steps[idxvar * 4 + 2 * (which) + 1], st_now()); } while (0)
#else
#define STEP_BEGIN(which, idxvar) do { } while (0)
#define STEP_END(which, idxvar) do { } while (0)
#endif

struct StreamArgs {
    const int4 *recs;           // StRec
    const int32_t *cols;        // column of every record (the producer needs nothing else)
    int64_t nrec;
    int R, ntile, tw, nstage;
    const double *X;
    int64_t ldx;
    int nvar;
    double *Y1;
    int64_t ldy1;
    int32_t nrow1;
    double *Y2;
    int64_t ldy2;
    double *partials;
    int64_t ldp;
};

// Second (small) kernel: every row with more than one sum is the sum of its slots, added in slot order (= stream
// order, fixed at plan time); rows without any entry are zero.  One warp per (row, 64-field strip); a row's slots are
// consecutive, so all of a row's loads are independent: one L2 round trip per row.  Launched with programmatic stream
// serialization: the warps are resident and hold their row's (first slot, count) before the main kernel has finished.
struct FixArgs {
    const int32_t *fixrows;
    const int2 *rowinfo;
    int nfix, nstrip, nvar;
    const double *partials;
    int64_t ldp;
    double *Y1;
    int64_t ldy1;
    int32_t nrow1;
    double *Y2;
    int64_t ldy2;
};
constexpr int FIX_WARPS = 8;
__global__ void __launch_bounds__(FIX_WARPS * 32) spmm_stream_fixup_kernel(const FixArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * FIX_WARPS + (threadIdx.x >> 5);
    pdl_launch_dependents();
    const bool mine = w < (int64_t)a.nfix * a.nstrip;
    int32_t r = 0;
    int2 info = make_int2(-1, 0);
    const int f = (int)(w % a.nstrip) * 64 + 2 * lane;
    if (mine) { r = __ldg(a.fixrows + w / a.nstrip); info = __ldg(a.rowinfo + r); }      // plan data: before the wait
    pdl_wait();
    STEP_BEGIN(1, st_idx);
    if (!mine || f >= a.nvar) { STEP_END(1, st_idx); return; }
    double2 t = make_double2(0.0, 0.0);
    const double *pb = a.partials + (int64_t)info.x * a.ldp + f;
    for (int q0 = 0; q0 < info.y; q0 += 8) {
        double2 p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            p[u] = q0 + u < info.y ? __ldcg(reinterpret_cast<const double2 *>(pb + (int64_t)(q0 + u) * a.ldp)) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (q0 + u < info.y) { t.x = t.x + p[u].x; t.y = t.y + p[u].y; }
    }
    double *y = (r < a.nrow1 ? a.Y1 + (int64_t)r * a.ldy1 : a.Y2 + (int64_t)(r - a.nrow1) * a.ldy2) + f;
    *reinterpret_cast<double2 *>(y) = t;
    STEP_END(1, st_idx);
}

template <int NACC>
__global__ void __launch_bounds__(288, 3) spmm_stream_kernel(const __grid_constant__ StreamArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
#ifdef IB_STEP_TRACE
    unsigned &st_idx_s = *reinterpret_cast<unsigned *>(smem_raw + ST_HDR - 8);      // the header's tail is unused
#endif
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + a.nstage;
    unsigned char *data = smem_raw + 256;           // 2*nstage*8 <= 256
    const uint32_t rowbytes = (uint32_t)a.tw * 8u;
    const uint32_t stagebytes = rowbytes * ST_ROWS + ST_HDR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (int)(blockDim.x >> 5) - 1;     // consumer warps
    pdl_launch_dependents();
    if (warp == 0) {
        TRACE(0);
#ifdef IB_STREAM_TRACE
        if (lane == 0 && blockIdx.x < 8192) { unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); g_stream_trace[blockIdx.x * TR_N + 6] = sm; }
#endif
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int run = (int)(blockIdx.x / (unsigned)a.ntile), tile = (int)(blockIdx.x % (unsigned)a.ntile);
    const int jb = (int)run_begin(a.nrec, a.R, run), je = (int)run_begin(a.nrec, a.R, run + 1);      // nrec < 2^31
    const int f0 = tile * a.tw;
    const int live = min(a.tw, a.nvar - f0);        // fields of this tile
    int stage = 0;
    uint32_t phase = 0;
    if (warp == ncw) {
        // ===== producer warp: lane r < ST_ROWS copies X row r of the stage, lane ST_ROWS the stage's records
        const uint32_t tb = (uint32_t)live * 8u;
        int c_next = (lane < ST_ROWS && jb + lane < je) ? __ldg(a.cols + jb + lane) : 0;
        pdl_wait();                                 // X may be the previous kernel's output
        TRACE(1);
#ifdef IB_STEP_TRACE
        if (lane == 0) {
            st_idx_s = (atomicAdd(&g_st_arrive[0], 1u) / gridDim.x) & 4095u;
            atomicMin(&g_st_steps[st_idx_s * 4], st_now());
        }
#endif
        for (int j = jb; j < je; j += ST_ROWS) {
            const int nr = min(ST_ROWS, je - j);
            const int c = c_next;
            const int jn = j + ST_ROWS + lane;
            if (lane < ST_ROWS && jn < je) c_next = __ldg(a.cols + jn);
            unsigned char *sp = data + (size_t)stage * stagebytes;
            if (lane == 0) {
                mbar_wait(empty + stage, phase ^ 1u);
                mbar_expect_tx(full + stage, (tb + 64u) * (uint32_t)nr);
            }
            __syncwarp();
            if (lane < nr) bulk_g2s(sp + (size_t)lane * rowbytes, a.X + (int64_t)c * a.ldx + f0, tb, full + stage);
            else if (lane == ST_ROWS) bulk_g2s(sp + (size_t)ST_ROWS * rowbytes, a.recs + 4 * (int64_t)j, 64u * (uint32_t)nr, full + stage);
            if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
        }
        TRACE(5);
        return;
    }
    // ===== consumer warps: lane owns fields f0 + 64*warp + 2*lane, +1
    const int fo = 64 * warp + 2 * lane;            // offset inside the tile (< tw: always inside the row buffer)
    const bool active = fo < live;
    pdl_wait();                                     // nothing is written before the previous kernel has completed
    double2 acc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = make_double2(0.0, 0.0);
    for (int j = jb; j < je; j += ST_ROWS) {
        const int nr = min(ST_ROWS, je - j);
        mbar_wait(full + stage, phase);
        if (warp == 0 && j == jb) TRACE(2);
        const unsigned char *sb = data + (size_t)stage * stagebytes;
        const unsigned char *hdr = sb + (size_t)ST_ROWS * rowbytes;
        const unsigned char *sp = sb + (size_t)fo * 8;
        // second quarter of every record: v2, column, flags
        int4 h1[ST_ROWS];
#pragma unroll
        for (int r = 0; r < ST_ROWS; ++r) h1[r] = *reinterpret_cast<const int4 *>(hdr + 64 * r + 16);
        uint32_t any = 0;
#pragma unroll
        for (int r = 0; r < ST_ROWS; ++r) any |= (uint32_t)h1[r].w;
        if (any == 0u && nr == ST_ROWS) {
            // common case: a full stage in the middle of its rows' sums; straight-line code
            double2 x[ST_ROWS], v01[ST_ROWS];
#pragma unroll
            for (int r = 0; r < ST_ROWS; ++r) {
                x[r] = *reinterpret_cast<const double2 *>(sp + (size_t)r * rowbytes);
                v01[r] = *reinterpret_cast<const double2 *>(hdr + 64 * r);
            }
#pragma unroll
            for (int r = 0; r < ST_ROWS; ++r) {
                acc[0].x = acc[0].x + v01[r].x * x[r].x; acc[0].y = acc[0].y + v01[r].x * x[r].y;
                acc[1].x = acc[1].x + v01[r].y * x[r].x; acc[1].y = acc[1].y + v01[r].y * x[r].y;
                if (NACC == 3) {
                    const double v2 = __hiloint2double(h1[r].y, h1[r].x);
                    acc[2].x = acc[2].x + v2 * x[r].x; acc[2].y = acc[2].y + v2 * x[r].y;
                }
            }
        } else {
            // a sum ends in this stage (or the run does)
#pragma unroll 1
            for (int r = 0; r < nr; ++r) {
                const double2 x = *reinterpret_cast<const double2 *>(sp + (size_t)r * rowbytes);
                const double2 v01 = *reinterpret_cast<const double2 *>(hdr + 64 * r);
                const int4 q1 = *reinterpret_cast<const int4 *>(hdr + 64 * r + 16);
                const double v2 = __hiloint2double(q1.y, q1.x);
                const uint32_t fl = (uint32_t)q1.w;
                acc[0].x = acc[0].x + v01.x * x.x; acc[0].y = acc[0].y + v01.x * x.y;
                acc[1].x = acc[1].x + v01.y * x.x; acc[1].y = acc[1].y + v01.y * x.y;
                if (NACC == 3) { acc[2].x = acc[2].x + v2 * x.x; acc[2].y = acc[2].y + v2 * x.y; }
                if (fl) {                           // warp-uniform
#pragma unroll 1
                    for (int k = 0; k < NACC; ++k) {
                        if (fl & (SF_EVENT << k)) {
                            const double2 sum = k == 0 ? acc[0] : (k == 1 ? acc[1] : acc[2]);
                            const int2 ev = *reinterpret_cast<const int2 *>(hdr + 64 * r + 32 + 8 * k);
                            // the row's only sum goes straight to Y, any other to its slot (added up by the fix-up kernel)
                            double *dst = ev.y >= 0 ? a.partials + (int64_t)ev.y * a.ldp
                                                    : (ev.x < a.nrow1 ? a.Y1 + (int64_t)ev.x * a.ldy1 : a.Y2 + (int64_t)(ev.x - a.nrow1) * a.ldy2);
                            if (active) *reinterpret_cast<double2 *>(dst + f0 + fo) = sum;
                        }
                    }
                    if (fl & SF_CLOSE) acc[0] = make_double2(0.0, 0.0);
                    if (fl & (SF_CLOSE << 1)) acc[1] = make_double2(0.0, 0.0);
                    if (fl & (SF_CLOSE << 2)) acc[2] = make_double2(0.0, 0.0);
                    if (fl & SF_SHIFT) { acc[0] = acc[1]; acc[1] = make_double2(0.0, 0.0); }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
        if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
    }
    if (warp == 0) { TRACE(3); TRACE(4); }
#ifdef IB_STEP_TRACE
    if (threadIdx.x == 0) atomicMax(&g_st_steps[st_idx_s * 4 + 1], st_now());
#endif
}

// ---- host side -------------------------------------------------------------------------------------------
// Applies [M1; M2] (M2 may be null) in one pass over X.  Returns false if the pair does not qualify (nothing launched).
bool apply_stream(Weighted &M1, Weighted *M2, const double *dX, int64_t ldx, int nvar, double *dY1, int64_t ldy1, double *dY2,
                  int64_t ldy2, cudaStream_t s)
{
    if (nvar < 64 || nvar % 2 || ldx % 2 || ldy1 % 2 || (uintptr_t)dX % 16 || (uintptr_t)dY1 % 16) return false;
    if (M2 && (ldy2 % 2 || (uintptr_t)dY2 % 16 || M2->ncol != M1.ncol)) return false;
    const int dev = device_index();
    static int max_smem[IB_MAX_DEVICES] = {};
    {
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        if (!max_smem[dev]) {
            int ms = 0;
            IB_CUDA(cudaDeviceGetAttribute(&ms, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            IB_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
            IB_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
            max_smem[dev] = ms;
        }
    }
    const int tw = std::min((nvar + 63) / 64 * 64, 512);
    const int ntile = (nvar + tw - 1) / tw;
    const int ncw = tw / 64;
    const size_t stagebytes = (size_t)tw * 8 * ST_ROWS + ST_HDR;
    const int bps = std::max(1, g_tune.stream_bps);
    const size_t budget = (size_t)(max_smem[dev] + 1024) / bps - 1024 - 256;
    const int nstage = (int)std::min<size_t>(16, budget / stagebytes);
    if (nstage < 2) return false;
    const size_t smem = 256 + (size_t)nstage * stagebytes;
    // one CTA per run and field tile, ALL co-resident (a second wave would double the time): ask the occupancy
    int occ = 0;
    if (M2) IB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_stream_kernel<3>, 32 * (ncw + 1), smem));
    else IB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_stream_kernel<2>, 32 * (ncw + 1), smem));
    if (occ < 1) return false;
    const int R = std::max(1, (sm_count() * std::min(bps, occ) - g_tune.stream_spare) / ntile);
    StreamPlan &P = M1.stream_plan;
    if (!make_stream_plan(P, M1, M2, R, s)) return false;
    const int64_t ldp = (int64_t)ntile * tw;
    if (P.partials.n < (size_t)P.nslots * ldp) P.partials.alloc((size_t)P.nslots * ldp + 2, s);
    StreamArgs a{P.recs.p, P.cols.p, P.nrec, P.R, ntile, tw, nstage, dX, ldx, nvar, dY1, ldy1,
                 (int32_t)P.nrow1, dY2, ldy2, P.partials.p, ldp};
    // programmatic stream serialization: the prologue of a kernel overlaps the tail of the one before it
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(P.R * ntile));
    cfg.blockDim = dim3((unsigned)(32 * (ncw + 1)));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (M2) IB_CUDA(cudaLaunchKernelEx(&cfg, spmm_stream_kernel<3>, a));
    else IB_CUDA(cudaLaunchKernelEx(&cfg, spmm_stream_kernel<2>, a));
    IB_LAUNCHED();
    if (P.nfix > 0) {
        const int nstrip = (nvar + 63) / 64;
        FixArgs fa{P.fixrows.p, P.rowinfo.p, (int)P.nfix, nstrip, nvar, P.partials.p, ldp, dY1, ldy1, (int32_t)P.nrow1, dY2, ldy2};
        cudaLaunchConfig_t fc{};
        fc.gridDim = dim3((unsigned)cdiv(P.nfix * nstrip, FIX_WARPS));
        fc.blockDim = dim3(FIX_WARPS * 32);
        fc.stream = s;
        fc.attrs = attr;
        fc.numAttrs = 1;
        IB_CUDA(cudaLaunchKernelEx(&fc, spmm_stream_fixup_kernel, fa));
        IB_LAUNCHED();
    }
    return true;
}

#ifdef IB_STEP_TRACE
extern "C" int ib_debug_stream_steps(unsigned long long *out, int n, int reset)
{
    cudaDeviceSynchronize();
    int rc = 0;
    if (out) rc = (int)cudaMemcpyFromSymbol(out, g_st_steps, sizeof(unsigned long long) * 4 * n);
    if (reset) {
        static unsigned long long init[4096 * 4];
        for (int i = 0; i < 4096; ++i) { init[4 * i] = init[4 * i + 2] = ~0ull; init[4 * i + 1] = init[4 * i + 3] = 0; }
        unsigned zero[2] = {0, 0};
        cudaMemcpyToSymbol(g_st_steps, init, sizeof init);
        cudaMemcpyToSymbol(g_st_arrive, zero, sizeof zero);
        cudaDeviceSynchronize();
    }
    return rc;
}
#endif
#ifdef IB_STREAM_TRACE
extern "C" int ib_debug_stream_trace(unsigned long long *out, int nblk)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_stream_trace, sizeof(unsigned long long) * TR_N * nblk);
}
#endif

}   // namespace ib
