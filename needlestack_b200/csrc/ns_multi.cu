/*
 * ns_multi.cu — one workload over the GPUs of a box from ONE host process (the C ABI's ns_multi_* entry points).
 *
 * The reference parallelises by genomic region: bed_cut.r splits the BED into --nsplit pieces, one single-threaded
 * Rscript per piece, and the per-piece VCFs are concatenated and sorted at the end (needlestack.nf:421-503,
 * :546-553, bin/bed_cut.r:19-58).  Here the pieces are chunks of consecutive positions in ONE queue; one host thread
 * per GPU pulls the next chunk whenever its device is free (the work per chunk depends on the data, so a static
 * split would leave devices idle), every device copies its chunks from the caller's host planes itself, keeps the
 * rows of its emitted units in private page-locked planes, and the rows are gathered in chunk order - which is
 * genomic order - when all threads have finished.  No collective: units are independent.
 * R is single-threaded; a `.Call` into ns_multi_call_snv is how an R process reaches all GPUs.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "ns_api_internal.hpp"

struct ns_multi {
    std::vector<ns_ctx *> ctx;
    std::vector<HostRows> priv;
    std::vector<double> wall_ms;
    std::string err;
};

extern "C" {

ns_multi *ns_create_multi(const int *devices, int n_devices, char *err, size_t errlen)
{
    if (!devices || n_devices < 1 || n_devices > 64) {
        if (err && errlen)
            snprintf(err, errlen, "ns_create_multi: need 1..64 device ordinals");
        return nullptr;
    }
    ns_multi *m = new ns_multi();
    for (int i = 0; i < n_devices; i++) {
        ns_ctx *c = ns_create(devices[i], err, errlen);
        if (!c) {
            for (ns_ctx *d : m->ctx)
                ns_destroy(d);
            delete m;
            return nullptr;
        }
        m->ctx.push_back(c);
    }
    m->priv.resize((size_t)n_devices);
    m->wall_ms.assign((size_t)n_devices, 0.0);
    return m;
}

void ns_destroy_multi(ns_multi *m)
{
    if (!m)
        return;
    for (size_t i = 0; i < m->ctx.size(); i++) {
        cudaSetDevice(m->ctx[i]->device);
        m->priv[i].release();
        ns_destroy(m->ctx[i]);
    }
    delete m;
}

int ns_multi_devices(const ns_multi *m) { return m ? (int)m->ctx.size() : 0; }

const char *ns_multi_last_error(ns_multi *m) { return m ? m->err.c_str() : "null context"; }

int ns_multi_get_timings(ns_multi *m, int slot, ns_timings *t, double *wall_ms)
{
    if (!m || slot < 0 || slot >= (int)m->ctx.size() || !t)
        return NS_ERR_ARG;
    *t = m->ctx[(size_t)slot]->tm;
    if (wall_ms)
        *wall_ms = m->wall_ms[(size_t)slot];
    return NS_OK;
}

} /* extern "C" */

static int ns_multi_snv(ns_multi *m, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const int32_t *counts,
                        const uint16_t *counts16, ns_out *out)
{
    if (!m)
        return NS_ERR_ARG;
    if (!prm || !out || n < 1 || P < 0 || (P > 0 && (!ref || (!counts && !counts16)))) {
        m->err = "ns_multi_call_snv: bad argument";
        return NS_ERR_ARG;
    }
    const int D = (int)m->ctx.size();
    if (D == 1) {
        int rc = counts16 ? ns_call_snv_u16(m->ctx[0], prm, n, P, ref, counts16, out)
                          : ns_call_snv(m->ctx[0], prm, n, P, ref, counts, out);
        if (rc)
            m->err = ns_last_error(m->ctx[0]);
        m->wall_ms[0] = m->ctx[0]->tm.total_ms;
        return rc;
    }
    const int64_t U = 3 * P;
    ChunkQueue q;
    /* Chunk size.  A device's cluster-path units start when the gate of their chunk has run and take 35-130 ms each
     * whatever else the device does, so when a device's whole share is of that order (a 1 Mb panel over 8 GPUs: ~45 ms
     * of fits per device) many small chunks make the device wait for the units of its last chunk.  Measured, C2 over
     * 8 GPUs: 4 chunks per device 105 ms, 2 chunks 95-97 ms, 1 chunk 102 ms (no copy/compute overlap); over 4 GPUs:
     * 2 chunks 125 ms, 1 chunk 127 ms.  The floor is the slowest single unit (~90 ms on the device that gets it).
     * So: 2 chunks per device up to ~300 ms of estimated share, dealt evenly; 4 beyond, first come first served,
     * so that the queue balances the data-dependent work; never fewer than 16384 positions.
     * Estimate: ~4 ns per sample and position (C2 3.5, C4 6.3). */
    int64_t pc = U > 0 ? ns_pick_chunk(m->ctx[0], n, U) / 3 : 1;
    const double share_ms = 4.0e-6 * (double)n * (double)P / (double)D;
    int pieces = share_ms < 300.0 ? 2 : 4;
    const char *ep = getenv("NS_MULTI_PIECES");
    if (ep && atoi(ep) > 0)
        pieces = atoi(ep);
    int64_t per = (P + (int64_t)D * pieces - 1) / ((int64_t)D * pieces);
    if (per < 16384)
        per = 16384;
    if (pc > per)
        pc = per;
    const char *e = getenv("NS_MULTI_CHUNK_POSITIONS");
    if (e && atoll(e) > 0)
        pc = atoll(e);
    if (pc < 1)
        pc = 1;
    q.pchunk = pc;
    q.n_chunks = P > 0 ? (P + pc - 1) / pc : 0;
    q.owner.assign((size_t)q.n_chunks, -1);
    /* few chunks per device: an even deal (every thread prefetches one chunk ahead, so the threads that start first
     * would otherwise take the chunks of the last ones); many: first come, first served, which balances the work */
    q.max_per_slot = pieces <= 2 ? (q.n_chunks + D - 1) / D : 0;
    const RowSink proto = ns_sink_from_out(out);
    std::vector<ns_out> o((size_t)D, *out);
    std::vector<RowSink> sink((size_t)D);
    std::vector<int64_t> emitted((size_t)D, 0);
    std::vector<int> rcs((size_t)D, NS_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < D; d++) {
        th.emplace_back([&, d]() {
            const auto t0 = std::chrono::steady_clock::now();
            ns_ctx *c = m->ctx[(size_t)d];
            cudaSetDevice(c->device);
            ns_reset_out(c, &o[(size_t)d]);
            RowSink &s = sink[(size_t)d];
            s.want_rows = proto.want_rows;
            s.grow = &m->priv[(size_t)d];
            if (!s.grow->reserve(1024, n, 0)) {
                c->err = "page-locked allocation (private rows) failed";
                rcs[(size_t)d] = NS_ERR_NOMEM;
                q.failed.store(1);
                return;
            }
            s.p = s.grow->p;
            s.cap = s.grow->cap;
            int rc = ns_snv_host_queue(c, prm, n, P, ref, counts, counts16, &o[(size_t)d], &s, &q, d, &emitted[(size_t)d]);
            if (rc) {
                q.failed.store(1);
                ns_drain(c);
            }
            rcs[(size_t)d] = rc;
            m->wall_ms[(size_t)d] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    for (auto &t : th)
        t.join();
    for (int d = 0; d < D; d++)
        if (rcs[(size_t)d]) {
            char b[64];
            snprintf(b, sizeof b, "device slot %d: ", d);
            m->err = std::string(b) + m->ctx[(size_t)d]->err;
            return rcs[(size_t)d];
        }
    /* gather: chunk c's rows sit in its owner's private planes, located by the chunk's unit range (the private
     * rows are in unit order: every device takes its chunks in increasing order, and its pool merges by unit) */
    std::vector<ns_chunk_seg> seg((size_t)q.n_chunks);
    std::vector<ns_row_planes> planes((size_t)D);
    for (int d = 0; d < D; d++)
        planes[(size_t)d] = m->priv[(size_t)d].p;
    int64_t total = 0;
    for (int64_t c = 0; c < q.n_chunks; c++) {
        const int d = q.owner[(size_t)c];
        const int64_t p0 = c * pc, pcn = (P - p0 < pc) ? (P - p0) : pc;
        ns_chunk_seg &s = seg[(size_t)c];
        s.device = d < 0 ? 0 : d;
        s.unit0 = 3 * p0;
        s.units = 3 * pcn;
        s.row0 = s.rows = 0;
        if (d >= 0) {
            const int64_t *eu = m->priv[(size_t)d].p.emit_unit, E = emitted[(size_t)d];
            const int64_t Ec = E < m->priv[(size_t)d].cap ? E : m->priv[(size_t)d].cap;
            const int64_t *lo = std::lower_bound(eu, eu + Ec, s.unit0), *hi = std::lower_bound(eu, eu + Ec, s.unit0 + s.units);
            s.row0 = lo - eu;
            s.rows = hi - lo;
        }
        total += s.rows;
    }
    out->n_emitted = total;
    out->n_fitted = out->n_skipped = 0;
    out->n_seed = out->n_lattice = out->n_quad = 0;
    for (int d = 0; d < D; d++) {
        out->n_fitted += o[(size_t)d].n_fitted;
        out->n_seed += o[(size_t)d].n_seed;
        out->n_lattice += o[(size_t)d].n_lattice;
        out->n_quad += o[(size_t)d].n_quad;
        out->n_skipped += m->ctx[(size_t)d]->n_skipped_call;
    }
    out->n_gated = U - out->n_fitted - out->n_skipped;
    if (proto.want_rows && total > out->emit_cap) {
        m->err = "more emitted units than out->emit_cap";
        return NS_ERR_EMIT_OVERFLOW;
    }
    ns_row_planes dst = proto.p;
    if (!proto.want_rows)
        memset(&dst, 0, sizeof(dst));
    ns_concat_rows(dst, proto.want_rows ? out->emit_cap : total, planes.data(), seg.data(), q.n_chunks, n, out->emit_index);
    return NS_OK;
}

extern "C" {

int ns_multi_call_snv(ns_multi *m, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const int32_t *counts,
                      ns_out *out)
{
    return ns_multi_snv(m, prm, n, P, ref, counts, nullptr, out);
}

int ns_multi_call_snv_u16(ns_multi *m, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const uint16_t *counts,
                          ns_out *out)
{
    return ns_multi_snv(m, prm, n, P, ref, nullptr, counts, out);
}

} /* extern "C" */
