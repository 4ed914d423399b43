/*
 * tests/emul/ns_hostmerge_test.cpp — TEST INFRASTRUCTURE ONLY: C entry points around the host-side row bookkeeping of
 * needlestack_b200/csrc/ns_hostmerge.hpp (pure C++), so that the late-row merge of the cluster-path pool and the
 * chunk-order gather of a multi-GPU call can be checked on a machine without a GPU (tests/test_host_merge.py).
 */
#include "../../needlestack_b200/csrc/ns_hostmerge.hpp"
#include <vector>

extern "C" {

/* planes: only gq [rows][n], gt [rows][n], pooled [rows][8], emit_unit [rows] are exercised (the others share the code) */
int64_t hm_merge(int n, int64_t E1, double *gq, uint8_t *gt, double *pooled, int64_t *emit_unit, int use_emit_unit,
                 int64_t E2, const int64_t *late_unit, const int64_t *late_before, const int64_t *late_src,
                 double *l_gq, uint8_t *l_gt, double *l_pooled, int32_t *emit_index, int64_t n_units)
{
    ns_row_planes d = {}, l = {};
    d.gq = gq, d.gt = gt, d.pooled = pooled, d.emit_unit = use_emit_unit ? emit_unit : nullptr;
    l.gq = l_gq, l.gt = l_gt, l.pooled = l_pooled;
    std::vector<ns_late_row> rows((size_t)E2);
    for (int64_t j = 0; j < E2; j++)
        rows[(size_t)j] = ns_late_row{late_unit[j], late_before[j], late_src[j]};
    return ns_merge_pool_rows(d, E1, l, rows.data(), E2, n, emit_index, n_units);
}

/* two devices */
int64_t hm_concat(int n, int64_t cap, double *gq, int64_t *emit_unit, double *gq0, int64_t *eu0, double *gq1, int64_t *eu1,
                  int64_t n_chunks, const int *dev, const int64_t *row0, const int64_t *rows, const int64_t *unit0,
                  const int64_t *units, int32_t *emit_index)
{
    ns_row_planes d = {}, p[2] = {};
    d.gq = gq, d.emit_unit = emit_unit;
    p[0].gq = gq0, p[0].emit_unit = eu0;
    p[1].gq = gq1, p[1].emit_unit = eu1;
    std::vector<ns_chunk_seg> seg((size_t)n_chunks);
    for (int64_t c = 0; c < n_chunks; c++)
        seg[(size_t)c] = ns_chunk_seg{dev[c], row0[c], rows[c], unit0[c], units[c]};
    return ns_concat_rows(d, cap, p, seg.data(), n_chunks, n, emit_index);
}
}
