/*
 * ns_table.cpp — columnar loader for the mpileup2readcounts text table (host code, multi-threaded).
 *
 * Replaces the per-line strsplit / eval(as.name(paste(...))) / as.numeric gathers of
 * /root/reference/bin/needlestack.r:316-329 (SNV), :461-479 (DEL) and :624-643 (INS) with one
 * pass that packs the table into the layout of ns_call_snv / ns_call_units:
 *   counts int32 [P][9][n]  planes depth,A,T,C,G,a,t,c,g   (column map of needlestack.r:188-198)
 *   ref    uint8 [P]        0..3 = A,T,C,G, 4 = anything else (the reference skips such lines, :319)
 *   indel units             one per (position, unique upper-cased sequence) in first-appearance
 *                           order (sample order, then order inside the cell): Vp = count of the
 *                           entry spelled exactly SEQ, Vm = count of the entry spelled tolower(SEQ)
 * The count planes are allocated page-locked when a CUDA device is present, so that they can be
 * handed to ns_call_snv without another copy.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ns_b200.h"

struct IndelUnit {
    int64_t pos_index;
    uint8_t type; /* 1 = deletion, 2 = insertion */
    std::string seq;
    std::vector<int32_t> vp, vm;
};

struct ns_table {
    int n = 0;
    int64_t P = 0;
    int32_t *counts = nullptr;
    bool counts_pinned = false;
    size_t counts_cap = 0; /* bytes of the page-locked buffer (it may come from the pool and be larger) */
    std::vector<uint8_t> ref;
    std::vector<int64_t> loc;
    std::vector<int32_t> chrom;
    std::vector<std::string> chrom_names;
    std::vector<char> ref_text; /* the REF field as written, for the VCF (first character) */
    /* indel units, flattened */
    std::vector<int64_t> iu_pos;
    std::vector<uint8_t> iu_type;
    std::vector<std::string> iu_seq;
    std::vector<int32_t> iu_vp, iu_vm; /* [Ui][n] */
    std::vector<int64_t> indel_cov; /* [P][2] sum of all deletion / insertion counts (coverage_all_del/ins) */
    std::string err;
};

namespace {

/* A small pool of page-locked buffers: the driver parses block after block of the same size with up to three tables
 * alive at a time (one being parsed, one queued, one on the GPU), and cudaMallocHost / cudaFreeHost of a gigabyte each
 * cost ten times more than parsing it (0.4 s against 0.04 s for 0.7 GB).  A released buffer is kept for the next table
 * it fits; the smallest buffer that fits is handed out; beyond NS_PIN_POOL buffers the smallest is dropped. */
#define NS_PIN_POOL 4
std::mutex g_pool_mu;
void *g_pool_ptr[NS_PIN_POOL] = {nullptr, nullptr, nullptr, nullptr};
size_t g_pool_cap[NS_PIN_POOL] = {0, 0, 0, 0};

void *pinned_acquire(size_t bytes, size_t &cap)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (int i = 0; i < NS_PIN_POOL; i++)
            if (g_pool_ptr[i] && g_pool_cap[i] >= bytes && g_pool_cap[i] <= 2 * bytes + (64u << 20) &&
                (best < 0 || g_pool_cap[i] < g_pool_cap[best]))
                best = i;
        if (best >= 0) {
            void *p = g_pool_ptr[best];
            cap = g_pool_cap[best];
            g_pool_ptr[best] = nullptr;
            g_pool_cap[best] = 0;
            return p;
        }
    }
    void *p = nullptr;
    /* a little slack, so that blocks cut at line ends (sizes differ by a few lines) reuse one another's buffers */
    const size_t want = bytes + bytes / 16 + (1u << 20);
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    cap = want;
    return p;
}

void pinned_release(void *p, size_t cap)
{
    void *drop = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int slot = -1, smallest = 0;
        for (int i = 0; i < NS_PIN_POOL; i++) {
            if (!g_pool_ptr[i]) {
                slot = i;
                break;
            }
            if (g_pool_cap[i] < g_pool_cap[smallest])
                smallest = i;
        }
        if (slot >= 0) {
            g_pool_ptr[slot] = p;
            g_pool_cap[slot] = cap;
        } else if (g_pool_cap[smallest] < cap) {
            drop = g_pool_ptr[smallest];
            g_pool_ptr[smallest] = p;
            g_pool_cap[smallest] = cap;
        } else {
            drop = p;
        }
    }
    if (drop)
        cudaFreeHost(drop);
}

struct Chunk {
    const char *beg, *end;
    int64_t n_lines = 0, first_row = 0;
    std::vector<IndelUnit> indels;
    std::vector<std::pair<std::string, int64_t>> chroms; /* run-length: (name, first row) */
    std::string err;
};

inline const char *next_tab(const char *p, const char *end)
{
    const char *q = (const char *)memchr(p, '\t', (size_t)(end - p));
    return q ? q : end;
}

/* parses a non-negative integer field [p, q); returns false on anything else */
inline bool parse_int(const char *p, const char *q, int32_t &out)
{
    if (p == q)
        return false;
    int64_t v = 0;
    for (const char *c = p; c < q; c++) {
        if (*c < '0' || *c > '9')
            return false;
        v = v * 10 + (*c - '0');
        if (v > 2147483647LL)
            return false;
    }
    out = (int32_t)v;
    return true;
}

/* one INS or DEL cell: "NA" or "count:SEQ|count:seq|..." */
void parse_indel_cell(const char *p, const char *q, int sample, int n, uint8_t type, int64_t row,
                      std::vector<IndelUnit> &units, size_t first_unit_of_row, int64_t &cov, std::string &err)
{
    if (q - p == 2 && p[0] == 'N' && p[1] == 'A')
        return;
    while (p < q) {
        const char *bar = (const char *)memchr(p, '|', (size_t)(q - p));
        const char *e = bar ? bar : q;
        const char *colon = (const char *)memchr(p, ':', (size_t)(e - p));
        int32_t c = 0;
        if (!colon || !parse_int(p, colon, c)) {
            err = "malformed indel cell";
            return;
        }
        cov += c;
        std::string raw(colon + 1, e), up = raw, lo = raw;
        for (auto &ch : up)
            ch = (char)toupper((unsigned char)ch);
        for (auto &ch : lo)
            ch = (char)tolower((unsigned char)ch);
        /* candidate allele = upper-cased spelling, unique per (row, type), first-appearance order */
        size_t k = first_unit_of_row;
        for (; k < units.size(); k++)
            if (units[k].type == type && units[k].seq == up)
                break;
        if (k == units.size()) {
            IndelUnit u;
            u.pos_index = row;
            u.type = type;
            u.seq = up;
            u.vp.assign((size_t)n, 0);
            u.vm.assign((size_t)n, 0);
            units.push_back(std::move(u));
        }
        if (raw == up)
            units[k].vp[(size_t)sample] = c; /* forward strand: spelled in upper case */
        else if (raw == lo)
            units[k].vm[(size_t)sample] = c; /* reverse strand: spelled in lower case */
        /* mixed case: a candidate, but neither grep of needlestack.r:474-475 matches it */
        p = bar ? bar + 1 : q;
    }
}

void parse_chunk(Chunk &ck, ns_table *t)
{
    const int n = t->n;
    const char *p = ck.beg;
    int64_t row = ck.first_row;
    while (p < ck.end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(ck.end - p));
        if (!eol)
            eol = ck.end;
        const char *le = (eol > p && eol[-1] == '\r') ? eol - 1 : eol;
        if (le == p) { /* empty line */
            p = eol + 1;
            continue;
        }
        /* chr, loc, ref */
        const char *q = next_tab(p, le);
        if (ck.chroms.empty() || ck.chroms.back().first.size() != (size_t)(q - p) ||
            memcmp(ck.chroms.back().first.data(), p, (size_t)(q - p)) != 0)
            ck.chroms.emplace_back(std::string(p, q), row);
        const char *f = q + 1;
        q = next_tab(f, le);
        int64_t loc = 0;
        for (const char *c = f; c < q; c++) {
            if (*c < '0' || *c > '9') {
                ck.err = "non-numeric position in column 2";
                return;
            }
            loc = loc * 10 + (*c - '0');
        }
        t->loc[(size_t)row] = loc;
        f = q + 1;
        q = next_tab(f, le);
        uint8_t rc = 4;
        if (q - f == 1)
            rc = *f == 'A' ? 0 : *f == 'T' ? 1 : *f == 'C' ? 2 : *f == 'G' ? 3 : 4;
        t->ref[(size_t)row] = rc;
        t->ref_text[(size_t)row] = q > f ? *f : 'N';
        int32_t *dst = t->counts + (size_t)row * 9 * (size_t)n;
        const size_t first_unit = ck.indels.size();
        int64_t cov_ins = 0, cov_del = 0;
        for (int s = 0; s < n; s++) {
            /* nine count fields: hand-rolled scan (fields are 1-3 characters, memchr per field dominates) */
            const char *c = q;
            for (int k = 0; k < 9; k++) {
                if (c >= le) {
                    ck.err = "too few columns (expected 3 + 11 per sample)";
                    return;
                }
                ++c; /* the tab */
                uint32_t v = 0;
                const char *f0 = c;
                while (c < le && (unsigned)(*c - '0') <= 9u) {
                    v = v * 10u + (unsigned)(*c - '0');
                    ++c;
                }
                if (c == f0 || (c < le && *c != '\t') || c - f0 > 9) {
                    ck.err = "non-integer count field";
                    return;
                }
                dst[(size_t)k * n + s] = (int32_t)v;
            }
            q = c;
            /* insertion cell, then deletion cell (needlestack.r:197-198) */
            for (int k = 0; k < 2; k++) {
                if (q >= le) {
                    ck.err = "too few columns (expected 3 + 11 per sample)";
                    return;
                }
                f = q + 1;
                q = next_tab(f, le);
                if (rc < 4) /* the reference ignores the whole line otherwise (:319) */
                    parse_indel_cell(f, q, s, n, k == 0 ? 2 : 1, row, ck.indels, first_unit, k == 0 ? cov_ins : cov_del,
                                     ck.err);
                if (!ck.err.empty())
                    return;
            }
        }
        t->indel_cov[(size_t)row * 2] = cov_del;
        t->indel_cov[(size_t)row * 2 + 1] = cov_ins;
        /* the reference emits all deletions of a line before its insertions (:461 before :624) */
        std::stable_sort(ck.indels.begin() + (long)first_unit, ck.indels.end(),
                         [](const IndelUnit &a, const IndelUnit &b) { return a.type < b.type; });
        row++;
        p = eol + 1;
    }
}

} // namespace

extern "C" {

ns_table *ns_table_parse(const char *text, size_t len, int n_samples, int has_header, int n_threads, char *err,
                         size_t errlen)
{
    auto fail = [&](const std::string &m) -> ns_table * {
        if (err && errlen)
            snprintf(err, errlen, "ns_table_parse: %s", m.c_str());
        return nullptr;
    };
    if (!text || n_samples < 1)
        return fail("bad argument");
    const char *beg = text, *end = text + len;
    if (has_header) { /* needlestack.r:317 skips the first line */
        const char *nl = (const char *)memchr(beg, '\n', len);
        beg = nl ? nl + 1 : end;
    }
    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > 256)
        n_threads = 256;
    if ((size_t)(end - beg) < (size_t)n_threads * 4096)
        n_threads = 1;
    std::vector<Chunk> cks((size_t)n_threads);
    const char *p = beg;
    for (int k = 0; k < n_threads; k++) {
        const char *stop = k == n_threads - 1 ? end : beg + (size_t)(end - beg) * (size_t)(k + 1) / (size_t)n_threads;
        if (stop < p)
            stop = p;
        if (stop < end) {
            const char *nl = (const char *)memchr(stop, '\n', (size_t)(end - stop));
            stop = nl ? nl + 1 : end;
        }
        cks[(size_t)k].beg = p;
        cks[(size_t)k].end = stop;
        p = stop;
    }
    /* pass 1: non-empty lines per chunk */
    {
        std::vector<std::thread> th;
        for (auto &ck : cks)
            th.emplace_back([&ck]() {
                int64_t c = 0;
                const char *q = ck.beg;
                while (q < ck.end) {
                    const char *nl = (const char *)memchr(q, '\n', (size_t)(ck.end - q));
                    const char *e = nl ? nl : ck.end;
                    const char *le = (e > q && e[-1] == '\r') ? e - 1 : e;
                    if (le > q)
                        c++;
                    q = e + 1;
                }
                ck.n_lines = c;
            });
        for (auto &t : th)
            t.join();
    }
    ns_table *t = new ns_table();
    t->n = n_samples;
    int64_t P = 0;
    for (auto &ck : cks) {
        ck.first_row = P;
        P += ck.n_lines;
    }
    t->P = P;
    const size_t bytes = (size_t)P * 9 * (size_t)n_samples * sizeof(int32_t);
    if (bytes) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 &&
            (t->counts = (int32_t *)pinned_acquire(bytes, t->counts_cap)) != nullptr) {
            t->counts_pinned = true;
        } else {
            cudaGetLastError();
            t->counts = (int32_t *)malloc(bytes);
        }
        if (!t->counts) {
            delete t;
            return fail("out of memory");
        }
    }
    t->ref.assign((size_t)P, 4);
    t->ref_text.assign((size_t)P, 'N');
    t->loc.assign((size_t)P, 0);
    t->chrom.assign((size_t)P, 0);
    t->indel_cov.assign((size_t)P * 2, 0);
    /* pass 2 */
    {
        std::vector<std::thread> th;
        for (auto &ck : cks)
            th.emplace_back([&ck, t]() { parse_chunk(ck, t); });
        for (auto &x : th)
            x.join();
    }
    for (auto &ck : cks)
        if (!ck.err.empty()) {
            std::string m = ck.err;
            if (t->counts_pinned)
                pinned_release(t->counts, t->counts_cap);
            else
                free(t->counts);
            delete t;
            return fail(m);
        }
    /* chromosome dictionary in order of first appearance; indel units concatenated in input order */
    for (auto &ck : cks) {
        for (size_t r = 0; r < ck.chroms.size(); r++) {
            const std::string &c = ck.chroms[r].first;
            int32_t id = -1;
            for (int32_t k = (int32_t)t->chrom_names.size() - 1; k >= 0; k--)
                if (t->chrom_names[(size_t)k] == c) {
                    id = k;
                    break;
                }
            if (id < 0) {
                id = (int32_t)t->chrom_names.size();
                t->chrom_names.push_back(c);
            }
            const int64_t r0 = ck.chroms[r].second;
            const int64_t r1 = r + 1 < ck.chroms.size() ? ck.chroms[r + 1].second : ck.first_row + ck.n_lines;
            for (int64_t row = r0; row < r1; row++)
                t->chrom[(size_t)row] = id;
        }
        for (auto &u : ck.indels) {
            t->iu_pos.push_back(u.pos_index);
            t->iu_type.push_back(u.type);
            t->iu_seq.push_back(u.seq);
            t->iu_vp.insert(t->iu_vp.end(), u.vp.begin(), u.vp.end());
            t->iu_vm.insert(t->iu_vm.end(), u.vm.begin(), u.vm.end());
        }
    }
    return t;
}

void ns_table_free(ns_table *t)
{
    if (!t)
        return;
    if (t->counts) {
        if (t->counts_pinned)
            pinned_release(t->counts, t->counts_cap);
        else
            free(t->counts);
    }
    delete t;
}

int64_t ns_table_positions(const ns_table *t) { return t ? t->P : 0; }
int ns_table_samples(const ns_table *t) { return t ? t->n : 0; }
int ns_table_counts_pinned(const ns_table *t) { return t && t->counts_pinned; }
const int32_t *ns_table_counts(const ns_table *t) { return t ? t->counts : nullptr; }
const uint8_t *ns_table_ref(const ns_table *t) { return t ? t->ref.data() : nullptr; }
const char *ns_table_ref_text(const ns_table *t) { return t ? t->ref_text.data() : nullptr; }
const int64_t *ns_table_loc(const ns_table *t) { return t ? t->loc.data() : nullptr; }
const int32_t *ns_table_chrom(const ns_table *t) { return t ? t->chrom.data() : nullptr; }
int ns_table_n_chrom(const ns_table *t) { return t ? (int)t->chrom_names.size() : 0; }
const char *ns_table_chrom_name(const ns_table *t, int id)
{
    return (t && id >= 0 && id < (int)t->chrom_names.size()) ? t->chrom_names[(size_t)id].c_str() : nullptr;
}
const int64_t *ns_table_indel_cov(const ns_table *t) { return t ? t->indel_cov.data() : nullptr; }
int64_t ns_table_indel_units(const ns_table *t) { return t ? (int64_t)t->iu_pos.size() : 0; }
const int64_t *ns_table_indel_pos(const ns_table *t) { return t ? t->iu_pos.data() : nullptr; }
const uint8_t *ns_table_indel_type(const ns_table *t) { return t ? t->iu_type.data() : nullptr; }
const char *ns_table_indel_seq(const ns_table *t, int64_t u)
{
    return (t && u >= 0 && u < (int64_t)t->iu_seq.size()) ? t->iu_seq[(size_t)u].c_str() : nullptr;
}
const int32_t *ns_table_indel_vp(const ns_table *t) { return t ? t->iu_vp.data() : nullptr; }
const int32_t *ns_table_indel_vm(const ns_table *t) { return t ? t->iu_vm.data() : nullptr; }

} /* extern "C" */
