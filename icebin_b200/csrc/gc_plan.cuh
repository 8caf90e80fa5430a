// gc_plan.cuh -- greedy grouping of consecutive short rows by their DISTINCT columns (apply.cu, spmm_gc_kernel).
//
// I-side matrices (IvE, IvA: IceRegridder_L0.cpp:100-155 transposed) have 1..8 nonzeros per row, and neighbouring ice
// cells feed from the same GCM cell and the same few elevation classes: in a block of 32 consecutive rows the ~70
// nonzeros of IvE name only ~10 distinct columns.  An item is a run of consecutive rows inside one 32-row aligned
// block whose nonzeros name at most GC_S distinct columns; the kernel fetches every distinct X row of an item ONCE.
// The nonzeros of an item are re-encoded as 16-byte records (value, slot of the column in the item's list, flags) in
// row order, one dummy record per empty row, at most GC_MAXREC per item, so that the kernel's inner loop is a flat walk
// over records.  This header is plain C++ so that the host unit test (tests/cpp/gc_emul.cpp) runs the very routine the
// device runs.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define IB_HD __host__ __device__
#else
#define IB_HD
#endif

namespace ib {

constexpr int GC_S = 8;          // distinct X rows per item (one shared-memory slot each)
constexpr int GC_MAXREC = 64;    // records per item
constexpr int GC_BLOCK = 32;     // items never cross a multiple of GC_BLOCK rows (row-block sharding granularity)
constexpr uint32_t GC_LAST = 1u; // record closes its row
constexpr uint32_t GC_ZERO = 2u; // dummy record of an empty row (no arithmetic)

struct GcItem { int32_t row, rbeg, rend, negrows; };     // same layout as the int4 items of the other plans
// slot: bits 0-7 the record's own slot, bits 8-15 the slot of the NEXT record of the item (so that the kernel can fetch the
// next record's X row without having read the next record)
struct alignas(16) GcRec { double v; uint32_t slot, flags; };
struct GcCount { uint32_t items, recs; };

// Runs the greedy grouping over rows [b*32, b*32+32) and returns the number of items and records.  With EMIT, writes
// items[ibase..], ucol[(ibase+k)*GC_S + q] (distinct columns in first-appearance order, -1 = unused) and the block's
// records from recs[rbase].  Every row must have <= GC_S nonzeros.
template <bool EMIT>
IB_HD inline GcCount gc_block(const int32_t *rowptr, const int32_t *col, const double *val, int64_t nrow, int64_t b,
                              uint32_t ibase, uint32_t rbase, GcItem *items, int32_t *ucol, GcRec *recs)
{
    const int64_t r0 = b * GC_BLOCK;
    const int64_t r1 = (r0 + GC_BLOCK < nrow) ? r0 + GC_BLOCK : nrow;
    int32_t ul[GC_S];
    int nu = 0;
    uint32_t n = 0, nrec = 0, rec0 = 0;        // items closed, records so far, first record of the open item
    int64_t first = r0;
    auto close = [&](int64_t rend) {
        if (EMIT) {
            GcItem it;
            it.row = (int32_t)first; it.rbeg = (int32_t)(rbase + rec0); it.rend = (int32_t)(rbase + nrec);
            it.negrows = -(int32_t)(rend - first);
            items[ibase + n] = it;
            for (int q = 0; q < GC_S; ++q) ucol[(int64_t)(ibase + n) * GC_S + q] = q < nu ? ul[q] : -1;
        }
        ++n;
    };
    for (int64_t r = r0; r < r1; ++r) {
        const int32_t jb = rowptr[r], je = rowptr[r + 1];
        const uint32_t need = je > jb ? (uint32_t)(je - jb) : 1u;
        int nu2 = nu;
        bool fits = (nrec - rec0) + need <= (uint32_t)GC_MAXREC;
        for (int32_t j = jb; j < je && fits; ++j) {
            const int32_t c = col[j];
            bool found = false;
            for (int q = 0; q < nu2; ++q) found = found || (ul[q] == c);
            if (!found) {
                if (nu2 < GC_S) ul[nu2++] = c;
                else fits = false;
            }
        }
        if (!fits) {            // close [first, r) and restart the list with row r alone (a row always fits: <= GC_S nonzeros)
            close(r);
            first = r; rec0 = nrec; nu = 0;
            for (int32_t j = jb; j < je; ++j) {
                const int32_t c = col[j];
                bool found = false;
                for (int q = 0; q < nu; ++q) found = found || (ul[q] == c);
                if (!found && nu < GC_S) ul[nu++] = c;
            }
        } else {
            nu = nu2;
        }
        if (EMIT) {
            for (int32_t j = jb; j < je; ++j) {
                const int32_t c = col[j];
                uint32_t s = 0;
                for (int q = 0; q < nu; ++q) s = (ul[q] == c) ? (uint32_t)q : s;
                GcRec rc;
                rc.v = val[j]; rc.slot = s; rc.flags = (j + 1 == je) ? GC_LAST : 0u;
                const uint32_t p = nrec + (uint32_t)(j - jb);
                recs[rbase + p] = rc;
                if (p > rec0) recs[rbase + p - 1].slot |= s << 8;       // previous record of the same item
            }
            if (je <= jb) {
                GcRec rc;
                rc.v = 0.0; rc.slot = 0; rc.flags = GC_LAST | GC_ZERO;
                recs[rbase + nrec] = rc;
            }
        }
        nrec += need;
    }
    close(r1);
    GcCount out;
    out.items = n; out.recs = nrec;
    return out;
}

}   // namespace ib
