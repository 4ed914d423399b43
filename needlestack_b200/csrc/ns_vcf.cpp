/*
 * ns_vcf.cpp — VCF record writer (host code, multi-threaded): the text of the reference's per-variant
 * `cat(..., file=out_file, append=T)` calls, /root/reference/bin/needlestack.r:396-413 (SNV), :549-567 (deletion),
 * :712-730 (insertion), from a parsed table (ns_table) and the host result of ns_call_snv / ns_call_units on it.
 *
 * Field order and content follow the reference line by line; numbers are printed as R's cat() prints doubles under
 * options(digits = 7, scipen = 100) (bin/needlestack.r:153): the fewest significant digits (<= 7) that reproduce
 * the value rounded to 7, always in fixed notation; NA -> "NA", NaN -> "NaN", Inf -> "Inf".  The 3-bp context of
 * CONT= comes from an in-memory FASTA (the reference spawns two `samtools faidx` per variant, :389-394).
 * A record holds ~12 printed fields per sample: formatting them in Python was the slowest stage of the driver.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ns_b200.h"

struct ns_fasta {
    std::map<std::string, std::string> seq;
};

struct ns_vcf_text {
    std::string text;
    std::vector<int64_t> off; /* record e = text[off[e] .. off[e+1]) */
};

namespace {

inline void put_int(std::string &s, long long v)
{
    char b[24];
    int i = 24;
    const bool neg = v < 0;
    unsigned long long u = neg ? (unsigned long long)(-(v + 1)) + 1ULL : (unsigned long long)v;
    do {
        b[--i] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (neg)
        b[--i] = '-';
    s.append(b + i, (size_t)(24 - i));
}

/* R's cat() of one double, digits = 7, scipen = 100 (see the file header); nan_txt: "NA" where the library's NaN
 * stands for R's NA, "NaN" where the reference itself computes 0/0 */
inline void put_num(std::string &s, double x, const char *nan_txt)
{
    if (x != x) {
        s += nan_txt;
        return;
    }
    if (isinf(x)) {
        s += x > 0 ? "Inf" : "-Inf";
        return;
    }
    if (x == 0) {
        s += '0';
        return;
    }
    char b[512];
    const double ax = fabs(x);
    if (ax >= 1e-4 && ax < 1e7) {
        /* %g picks the same digits in this range and stays in fixed notation */
        const int L = snprintf(b, sizeof b, "%.7g", x);
        if (!memchr(b, 'e', (size_t)L)) {
            s.append(b, (size_t)L);
            return;
        }
    }
    snprintf(b, sizeof b, "%.6e", x);
    const char *e = strchr(b, 'e');
    const int kp = atoi(e + 1);
    int nsig = 0, seen = 0;
    for (const char *c = b; c < e; c++)
        if (*c >= '0' && *c <= '9') {
            seen++;
            if (*c != '0')
                nsig = seen;
        }
    if (nsig < 1)
        nsig = 1;
    int rgt = nsig - 1 - kp;
    if (rgt < 0)
        rgt = 0;
    if (rgt > 400)
        rgt = 400;
    const int L = snprintf(b, sizeof b, "%.*f", rgt, x);
    s.append(b, (size_t)(L < (int)sizeof b ? L : (int)sizeof b - 1));
}

const char *const GT_TXT[4] = {"0/0", "0/1", "1/1", "./."};
const char *const ST_TXT[6] = {".", "SOMATIC", "GERMLINE_CONFIRMED", "GERMLINE_UNCONFIRMED", "GERMLINE_UNCONFIRMABLE", "UNKNOWN"};

/* `samtools faidx chr:start-end | tail -n1` (1-based, inclusive) from memory; N's when there is no sequence */
std::string context(const ns_fasta *fa, const char *chrom, int64_t start, int64_t end)
{
    const int64_t want = end - start + 1 > 0 ? end - start + 1 : 0;
    if (!fa)
        return std::string((size_t)want, 'N');
    auto it = fa->seq.find(chrom);
    if (it == fa->seq.end())
        return std::string((size_t)want, 'N');
    const std::string &s = it->second;
    int64_t a = start - 1 > 0 ? start - 1 : 0, b = end > 0 ? end : 0;
    if (a > (int64_t)s.size())
        a = (int64_t)s.size();
    if (b > (int64_t)s.size())
        b = (int64_t)s.size();
    return b > a ? s.substr((size_t)a, (size_t)(b - a)) : std::string();
}

void upper(std::string &s)
{
    for (auto &c : s)
        c = (char)toupper((unsigned char)c);
}

struct Rec {
    const char *chrom;
    int64_t pos;
    std::string ref_txt, alt_txt, before, after;
    const char *type;
    const int32_t *dp, *vp, *vm, *rp, *rm; /* n each */
    int64_t row;                           /* in the emitted planes */
    double sigma, slope, max_gq;
    uint32_t flags;
};

void format_record(std::string &s, const Rec &r, int n, const ns_out *o, const char *const *names, double thr)
{
    const size_t ro = (size_t)r.row * (size_t)n;
    const double *gq = o->gq + ro, *qm = o->qval_minaf + ro, *sor = o->sor + ro, *rvsb = o->rvsb + ro, *fs = o->fs + ro;
    const uint8_t *gt = o->gt + ro, *st = o->status + ro;
    const double *pooled = o->pooled + (size_t)r.row * 8;
    int ns = 0, ncall = 0;
    double ao = 0;
    for (int i = 0; i < n; i++) {
        ns += r.dp[i] > 0;
        ncall += gq[i] >= thr;
        ao += (double)r.vp[i] + (double)r.vm[i];
    }
    const bool inv = r.flags & NS_F_REF_INV, er = r.flags & NS_F_EXTRA_ROB;
    s += r.chrom;
    s += '\t';
    put_int(s, r.pos);
    s += "\t.\t";
    s += r.ref_txt;
    s += '\t';
    s += r.alt_txt;
    s += '\t';
    put_num(s, r.max_gq, "NA");
    s += "\t.\tTYPE=";
    s += r.type;
    s += ";NS=";
    put_int(s, ns);
    s += ";AF=";
    put_num(s, ns ? (double)ncall / (double)ns : NAN, "NaN"); /* 0/0 in the reference (:397) */
    s += ";DP=";
    put_num(s, pooled[7], "NA");
    s += ";RO=";
    put_num(s, pooled[3] + pooled[4], "NA");
    s += ";AO=";
    put_num(s, ao, "NA");
    s += ";SRF=";
    put_num(s, pooled[3], "NA");
    s += ";SRR=";
    put_num(s, pooled[4], "NA");
    s += ";SAF=";
    put_num(s, pooled[5], "NA");
    s += ";SAR=";
    put_num(s, pooled[6], "NA");
    s += ";SOR=";
    put_num(s, pooled[0], "NA");
    s += ";RVSB=";
    put_num(s, pooled[1], "NA");
    s += ";FS=";
    put_num(s, pooled[2], "NA");
    s += ";ERR=";
    put_num(s, r.slope, "NA");
    s += ";SIG=";
    put_num(s, r.sigma, "NA");
    s += ";CONT=";
    s += r.before;
    s += 'x';
    s += r.after;
    if (er && !inv)
        s += ";WARN=EXTRA_ROBUST_GL";
    else if (er && inv)
        s += ";WARN=EXTRA_ROBUST_GL/INV_REF";
    else if (inv)
        s += ";WARN=INV_REF";
    s += inv ? "\tGT:QVAL_INV:DP:RO:AO:AF:SB:SOR:RVSB:FS:QVAL_minAF:STATUS" : "\tGT:QVAL:DP:RO:AO:AF:SB:SOR:RVSB:FS:QVAL_minAF:STATUS";
    /* POSSIBLE_CONTAMINATION_FROM_<samples with a GERMLINE* or UNKNOWN status> (:372-373) */
    std::string contam;
    for (int i = 0; i < n; i++)
        if (st[i] == NS_ST_POSSIBLE_CONTAMINATION) {
            contam = "POSSIBLE_CONTAMINATION_FROM_";
            bool first = true;
            for (int j = 0; j < n; j++)
                if (st[j] >= NS_ST_GERMLINE_CONFIRMED && st[j] <= NS_ST_UNKNOWN) {
                    if (!first)
                        contam += '_';
                    contam += names[j];
                    first = false;
                }
            break;
        }
    for (int i = 0; i < n; i++) {
        const long long dp = r.dp[i], vp = r.vp[i], vm = r.vm[i], rp = r.rp[i], rm = r.rm[i];
        s += '\t';
        s += GT_TXT[gt[i] & 3];
        s += ':';
        put_num(s, gq[i], "NA");
        s += ':';
        put_int(s, dp);
        s += ':';
        put_int(s, rp + rm);
        s += ':';
        put_int(s, vp + vm);
        s += ':';
        put_num(s, (double)(vp + vm) / (double)dp, "NaN"); /* IEEE: 0/0 NaN, y/0 Inf, as R */
        s += ':';
        put_int(s, rp);
        s += ',';
        put_int(s, rm);
        s += ',';
        put_int(s, vp);
        s += ',';
        put_int(s, vm);
        s += ':';
        put_num(s, sor[i], "NA");
        s += ':';
        put_num(s, rvsb[i], "NA");
        s += ':';
        put_num(s, fs[i], "NA");
        s += ':';
        put_num(s, qm[i], "NA");
        s += ':';
        if (st[i] == NS_ST_POSSIBLE_CONTAMINATION)
            s += contam;
        else
            s += ST_TXT[st[i] <= 5 ? st[i] : 0];
    }
    s += '\n';
}

/* every emitted plane present, and the unit's sigma / slope / max(GQ) / flags either per row or per unit */
bool have_planes(const ns_table *t, const ns_out *out, const char *const *sample_names)
{
    if (!t || !out || !sample_names || !out->emit_unit || !out->gq || !out->qval_minaf || !out->sor || !out->rvsb || !out->fs ||
        !out->gt || !out->status || !out->pooled)
        return false;
    const bool per_row = out->row_sigma && out->row_slope && out->row_max_gq && out->row_flags;
    const bool per_unit = out->sigma && out->slope && out->max_gq && out->flags;
    return per_row || per_unit;
}

void unit_values(Rec &r, const ns_out *out, int64_t e, int64_t u)
{
    if (out->row_sigma && out->row_slope && out->row_max_gq && out->row_flags) {
        r.sigma = out->row_sigma[e], r.slope = out->row_slope[e], r.max_gq = out->row_max_gq[e], r.flags = out->row_flags[e];
    } else {
        r.sigma = out->sigma[u], r.slope = out->slope[u], r.max_gq = out->max_gq[u], r.flags = out->flags[u];
    }
}

template <class MAKE>
ns_vcf_text *format_all(int64_t E, int n, const ns_out *o, const char *const *names, double thr, int n_threads, MAKE make)
{
    ns_vcf_text *out = new ns_vcf_text();
    out->off.assign((size_t)E + 1, 0);
    if (n_threads < 1)
        n_threads = 1;
    if ((int64_t)n_threads > E)
        n_threads = E > 0 ? (int)E : 1;
    std::vector<std::string> parts((size_t)n_threads);
    std::vector<std::vector<int64_t>> lens((size_t)n_threads);
    auto work = [&](int t) {
        const int64_t e0 = E * t / n_threads, e1 = E * (t + 1) / n_threads;
        std::string &s = parts[(size_t)t];
        s.reserve((size_t)(e1 - e0) * (size_t)(64 * n + 512));
        for (int64_t e = e0; e < e1; e++) {
            const size_t before = s.size();
            Rec r = make(e);
            format_record(s, r, n, o, names, thr);
            lens[(size_t)t].push_back((int64_t)(s.size() - before));
        }
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back(work, t);
        for (auto &x : th)
            x.join();
    }
    size_t tot = 0;
    for (auto &p : parts)
        tot += p.size();
    out->text.reserve(tot);
    int64_t e = 0, at = 0;
    for (int t = 0; t < n_threads; t++) {
        out->text += parts[(size_t)t];
        for (int64_t L : lens[(size_t)t]) {
            out->off[(size_t)e++] = at;
            at += L;
        }
    }
    out->off[(size_t)E] = at;
    return out;
}

} // namespace

extern "C" {

ns_fasta *ns_fasta_load(const char *path)
{
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f)
        return nullptr;
    ns_fasta *fa = new ns_fasta();
    std::string *cur = nullptr;
    char *line = nullptr;
    size_t cap = 0;
    ssize_t L;
    while ((L = getline(&line, &cap, f)) >= 0) {
        while (L > 0 && (line[L - 1] == '\n' || line[L - 1] == '\r' || line[L - 1] == ' ' || line[L - 1] == '\t'))
            L--;
        if (L > 0 && line[0] == '>') {
            size_t k = 1;
            while (k < (size_t)L && line[k] != ' ' && line[k] != '\t')
                k++;
            cur = &fa->seq[std::string(line + 1, k - 1)];
            cur->clear();
        } else if (cur) {
            size_t a = 0;
            while (a < (size_t)L && (line[a] == ' ' || line[a] == '\t'))
                a++;
            cur->append(line + a, (size_t)L - a);
        }
    }
    free(line);
    fclose(f);
    return fa;
}

void ns_fasta_free(ns_fasta *fa) { delete fa; }

/* SNV records of the emitted units of `out` = host result of ns_call_snv on the planes of `t` (rows in unit order) */
ns_vcf_text *ns_vcf_format_snv(const ns_table *t, const ns_out *out, const char *const *sample_names, const ns_fasta *fa,
                               double gq_threshold, int n_threads)
{
    if (!have_planes(t, out, sample_names))
        return nullptr;
    const int n = ns_table_samples(t);
    const int32_t *counts = ns_table_counts(t);
    const uint8_t *ref = ns_table_ref(t);
    const int64_t *loc = ns_table_loc(t);
    const int32_t *chrom = ns_table_chrom(t);
    static const char B[] = "ATCG";
    auto make = [&](int64_t e) {
        Rec r;
        const int64_t u = out->emit_unit[e], p = u / 3;
        const int k = (int)(u - 3 * p), rb = ref[p] & 3, alt = k + (k >= rb ? 1 : 0);
        const int32_t *c = counts + (size_t)p * 9 * (size_t)n;
        r.chrom = ns_table_chrom_name(t, chrom[p]);
        r.pos = loc[p];
        r.ref_txt.assign(1, B[rb]);
        r.alt_txt.assign(1, B[alt]);
        r.type = "snv";
        r.before = context(fa, r.chrom, r.pos - 3, r.pos - 1);
        r.after = context(fa, r.chrom, r.pos + 1, r.pos + 3);
        r.dp = c, r.vp = c + (size_t)(1 + alt) * n, r.vm = c + (size_t)(5 + alt) * n;
        r.rp = c + (size_t)(1 + rb) * n, r.rm = c + (size_t)(5 + rb) * n;
        r.row = e;
        unit_values(r, out, e, u);
        return r;
    };
    return format_all(out->n_emitted, n, out, sample_names, gq_threshold, n_threads, make);
}

/* deletion / insertion records of the emitted units of `out` = host result of ns_call_units on the indel units of `t` */
ns_vcf_text *ns_vcf_format_indel(const ns_table *t, const ns_out *out, const char *const *sample_names, const ns_fasta *fa,
                                 double gq_threshold, int n_threads)
{
    if (!have_planes(t, out, sample_names))
        return nullptr;
    const int n = ns_table_samples(t);
    const int32_t *counts = ns_table_counts(t);
    const uint8_t *ref = ns_table_ref(t);
    const int64_t *loc = ns_table_loc(t);
    const int32_t *chrom = ns_table_chrom(t);
    const int64_t *ipos = ns_table_indel_pos(t);
    const uint8_t *itype = ns_table_indel_type(t);
    const int32_t *ivp = ns_table_indel_vp(t), *ivm = ns_table_indel_vm(t);
    auto make = [&](int64_t e) {
        Rec r;
        const int64_t u = out->emit_unit[e], p = ipos[u];
        const int rb = ref[p] & 3;
        const int32_t *c = counts + (size_t)p * 9 * (size_t)n;
        const std::string seq = ns_table_indel_seq(t, u);
        const int64_t L = (int64_t)seq.size();
        r.chrom = ns_table_chrom_name(t, chrom[p]);
        r.pos = loc[p];
        if (itype[u] == 1) { /* deletion, bin/needlestack.r:541-549 */
            r.before = context(fa, r.chrom, r.pos + 1 - 3, r.pos + 1 - 1);
            r.after = context(fa, r.chrom, r.pos + 1 + L, r.pos + 1 + 3 + L - 1);
            upper(r.before), upper(r.after);
            const std::string prev = r.before.size() > 2 ? r.before.substr(2, 1) : std::string();
            r.ref_txt = prev + seq;
            r.alt_txt = prev;
            r.type = "del";
        } else { /* insertion, bin/needlestack.r:705-712 */
            r.before = context(fa, r.chrom, r.pos - 2, r.pos);
            r.after = context(fa, r.chrom, r.pos + 1, r.pos + 3);
            upper(r.before), upper(r.after);
            const std::string prev = r.before.size() > 2 ? r.before.substr(2, 1) : std::string();
            r.ref_txt = prev;
            r.alt_txt = prev + seq;
            r.type = "ins";
        }
        r.dp = c, r.vp = ivp + (size_t)u * n, r.vm = ivm + (size_t)u * n;
        r.rp = c + (size_t)(1 + rb) * n, r.rm = c + (size_t)(5 + rb) * n;
        r.row = e;
        unit_values(r, out, e, u);
        return r;
    };
    return format_all(out->n_emitted, n, out, sample_names, gq_threshold, n_threads, make);
}

const char *ns_vcf_text_data(const ns_vcf_text *v) { return v ? v->text.data() : nullptr; }
size_t ns_vcf_text_size(const ns_vcf_text *v) { return v ? v->text.size() : 0; }
const int64_t *ns_vcf_text_offsets(const ns_vcf_text *v) { return v ? v->off.data() : nullptr; }
void ns_vcf_text_free(ns_vcf_text *v) { delete v; }

/* one number as the VCF writer prints it (R's cat(), digits = 7, scipen = 100); na_as_nan != 0 prints NaN as "NaN" */
int ns_vcf_format_number(double x, int na_as_nan, char *buf, size_t buflen)
{
    std::string s;
    put_num(s, x, na_as_nan ? "NaN" : "NA");
    if (!buf || buflen == 0)
        return (int)s.size();
    const size_t L = s.size() < buflen - 1 ? s.size() : buflen - 1;
    memcpy(buf, s.data(), L);
    buf[L] = 0;
    return (int)s.size();
}

} /* extern "C" */
