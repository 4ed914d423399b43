/*
 * ns_hostmerge.hpp — host-side bookkeeping of the emitted rows (pure C++, no CUDA): the caller's result planes hold
 * the rows of the emitted units in increasing unit order (include/ns_b200.h).  Two producers deliver rows late:
 *   - the call-wide pool of cluster-path units (their fits finish after the chunk they came from has been copied
 *     out), merged in by ns_merge_pool_rows();
 *   - the devices of a multi-GPU call, which pull position chunks from one queue and keep their rows in private
 *     staging planes, concatenated in chunk order by ns_concat_rows().
 * The role of the reference's final `sort -k1,1V -k2,2n` over the per-region VCFs (needlestack.nf:546-553).
 * Compiled into libns_b200.so and, on its own, into the CPU test shim tests/emul/ns_hostmerge_test.cpp.
 */
#pragma once
#include <stdint.h>
#include <string.h>

struct ns_row_planes {
    double *pvalue, *qvalue, *gq, *qval_minaf, *sor, *rvsb, *fs; /* [rows][n], any may be null */
    uint8_t *gt, *status;                                        /* [rows][n] */
    double *pooled;                                              /* [rows][8] */
    int64_t *emit_unit;                                          /* [rows] */
    double *row_sigma, *row_slope, *row_max_gq;                  /* [rows] */
    uint32_t *row_flags;                                         /* [rows] */
};

/* copy `count` consecutive rows s[from..] -> d[to..]; planes missing on either side are skipped; overlapping ranges
 * of the same planes are allowed (memmove) */
static inline void ns_rows_copy(const ns_row_planes &d, int64_t to, const ns_row_planes &s, int64_t from, int64_t count,
                                int n)
{
    if (count <= 0)
        return;
    const size_t rb = (size_t)n * sizeof(double), cb = (size_t)n;
#define NS_MV(f, bytes)                                                                                                \
    if (d.f && s.f)                                                                                                    \
    memmove((char *)d.f + (size_t)to * (bytes), (const char *)s.f + (size_t)from * (bytes), (size_t)count * (bytes))
    NS_MV(pvalue, rb);
    NS_MV(qvalue, rb);
    NS_MV(gq, rb);
    NS_MV(qval_minaf, rb);
    NS_MV(sor, rb);
    NS_MV(rvsb, rb);
    NS_MV(fs, rb);
    NS_MV(gt, cb);
    NS_MV(status, cb);
    NS_MV(pooled, 8 * sizeof(double));
    NS_MV(emit_unit, sizeof(int64_t));
    NS_MV(row_sigma, sizeof(double));
    NS_MV(row_slope, sizeof(double));
    NS_MV(row_max_gq, sizeof(double));
    NS_MV(row_flags, sizeof(uint32_t));
#undef NS_MV
}

/* one late row: the unit it belongs to, how many rows of `dst` precede it (rows of smaller units already in place),
 * and where its data sits in the staging planes */
struct ns_late_row {
    int64_t unit;
    int64_t before; /* number of rows already in dst whose unit is smaller */
    int64_t src_row;
};

/* Inserts E2 late rows (sorted by unit, hence by `before`) into dst, which holds E1 rows in unit order and has room
 * for E1 + E2: walks backwards so that every row moves at most once.  emit_index (per unit, may be null) is
 * updated for the inserted rows and for every row that moved; the moved rows' units come from dst.emit_unit when
 * the caller asked for it, else from a sweep over emit_index[first_unit .. n_units).  Returns E1 + E2. */
static inline int64_t ns_merge_pool_rows(const ns_row_planes &dst, int64_t E1, const ns_row_planes &late,
                                         const ns_late_row *rows, int64_t E2, int n, int32_t *emit_index,
                                         int64_t n_units)
{
    if (E2 <= 0)
        return E1;
    if (emit_index && !dst.emit_unit) {
        /* shift of a row that was at position r = number of late rows with before <= r; the per-unit index is
         * non-decreasing over the emitted units, so one forward sweep with a running count does it */
        int64_t j = 0;
        for (int64_t u = rows[0].unit; u < n_units; u++) {
            const int32_t r = emit_index[u];
            if (r < 0)
                continue;
            while (j < E2 && rows[j].before <= (int64_t)r)
                j++;
            emit_index[u] = (int32_t)(r + j);
        }
    }
    int64_t i = E1; /* rows [0, i) of dst not yet moved */
    for (int64_t j = E2 - 1; j >= 0; j--) {
        const int64_t b = rows[j].before;
        /* light rows [b, i) go up by j + 1 */
        if (i > b) {
            ns_rows_copy(dst, b + j + 1, dst, b, i - b, n);
            if (emit_index && dst.emit_unit)
                for (int64_t r = b; r < i; r++)
                    emit_index[dst.emit_unit[r + j + 1]] = (int32_t)(r + j + 1);
            i = b;
        }
        ns_rows_copy(dst, b + j, late, rows[j].src_row, 1, n);
        if (dst.emit_unit)
            dst.emit_unit[b + j] = rows[j].unit;
        if (emit_index)
            emit_index[rows[j].unit] = (int32_t)(b + j);
    }
    return E1 + E2;
}

/* Multi-GPU gather: device d kept the rows of the chunks it processed, in increasing chunk order, in its private
 * planes; seg[c] = {device, first row, row count} for chunk c.  Rows are concatenated in chunk order into dst
 * (which is also genomic order); emit_index (per unit, may be null; holds device-private row numbers on entry) is
 * re-based per chunk.  Returns the number of rows written, or -(needed) when dst holds fewer than needed. */
struct ns_chunk_seg {
    int device;
    int64_t row0, rows;  /* in the device's private planes */
    int64_t unit0, units; /* unit range of the chunk */
};

static inline int64_t ns_concat_rows(const ns_row_planes &dst, int64_t dst_cap, const ns_row_planes *dev_planes,
                                     const ns_chunk_seg *seg, int64_t n_chunks, int n, int32_t *emit_index)
{
    int64_t need = 0;
    for (int64_t c = 0; c < n_chunks; c++)
        need += seg[c].rows;
    if (need > dst_cap)
        return -need;
    int64_t at = 0;
    for (int64_t c = 0; c < n_chunks; c++) {
        const ns_chunk_seg &s = seg[c];
        ns_rows_copy(dst, at, dev_planes[s.device], s.row0, s.rows, n);
        if (emit_index && s.rows > 0) {
            const int64_t delta = at - s.row0;
            for (int64_t u = s.unit0; u < s.unit0 + s.units; u++)
                if (emit_index[u] >= 0)
                    emit_index[u] = (int32_t)(emit_index[u] + delta);
        }
        at += s.rows;
    }
    return at;
}
