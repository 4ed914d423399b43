"""Drop-in driver for the reference's `Rscript needlestack.r` step: same `--key=value` options, the
mpileup2readcounts table on STDIN, `names.txt` in the working directory, VCF appended to `--out_file`.

    samtools mpileup ... | mpileup2readcounts ... | python -m needlestack_b200.driver --out_file=x.vcf --fasta_ref=ref.fa

What is kept from bin/needlestack.r: option names and defaults (:23-128), names.txt / pairs file
handling (:155-185), the VCF header (:258-305), the order and content of every VCF field (:396-413,
:549-567, :712-730), R's number formatting (`cat`: 7 significant digits, options(scipen=100)).
What is different: the table is read in blocks and parsed in one pass per block (readcounts.load_table, C++), all
candidate alleles of a block go through the GPU batch API, the VCF records are formatted by the C++ writer of the
library (ns_vcf_format_snv / _indel; `vcf_line` below is the same formatter in Python, kept as the executable
specification the tests compare against), and the 3-bp context comes from an in-memory FASTA instead of two
`samtools faidx` processes per variant.  Three stages overlap: a reader thread parses block k+1 while block k is on the
GPU and the records of block k-1 are being written.  Plots (`--do_plots`, Gviz) are out of scope: `--do_plots` other
than NONE is accepted and ignored with a warning.
"""
import math
import os
import sys

import numpy as np

from . import _abi, api, readcounts

STATUS_TEXT = {0: ".", 1: "SOMATIC", 2: "GERMLINE_CONFIRMED", 3: "GERMLINE_UNCONFIRMED", 4: "GERMLINE_UNCONFIRMABLE",
               5: "UNKNOWN"}
GT_TEXT = ("0/0", "0/1", "1/1", "./.")


# ---- R's cat() number formatting ------------------------------------------------------------------
def rfmt(x, digits=7, nan="NA"):
    """format one number as R's cat() does with options(digits=7, scipen=100): the fewest significant
    digits (<= 7) that reproduce the value rounded to 7, fixed notation.  The library reports R's NA as
    NaN; a true NaN of the reference (0/0 allelic fraction) is printed with nan="NaN"."""
    if x is None:
        return "NA"
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if math.isnan(x):
        return nan
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0:
        return "0"
    if digits == 7 and 1e-4 <= abs(x) < 1e7:
        # C's %g makes the same choice of digits in this range and stays in fixed notation (checked against the
        # general path on 4e5 random values in tests/test_loader_driver.py); 6 x faster, and a VCF record formats
        # ~600 numbers
        t = "%.7g" % x
        if "e" not in t:
            return t
    mant, exp = f"{x:.{digits - 1}e}".split("e")
    kp = int(exp)
    nsig = len(mant.replace("-", "").replace(".", "").rstrip("0")) or 1
    rgt = max(0, nsig - 1 - kp)
    return f"{x:.{rgt}f}"


def parse_args(argv):
    """--key=value pairs as bin/needlestack.r:23-35; defaults of :64-92"""
    a = {}
    for tok in argv:
        if tok.startswith("--"):
            k, _, v = tok[2:].partition("=")
            a[k] = v if v != "" else "TRUE"
    if "help" in a or "out_file" not in a:
        raise SystemExit("Mandatory arguments: --out_file=file_name (see bin/needlestack.r --help)")

    def num(k, d):
        return float(a[k]) if k in a else d

    def flag(k, d):
        return (a[k].upper() in ("TRUE", "T", "1")) if k in a else d

    o = dict(out_file=a["out_file"], fasta_ref=a.get("fasta_ref"), samtools=a.get("samtools", "samtools"),
             SB_type=a.get("SB_type", "SOR"), SB_threshold_SNV=num("SB_threshold_SNV", 100.0),
             SB_threshold_indel=num("SB_threshold_indel", 100.0), min_coverage=num("min_coverage", 30.0),
             min_reads=num("min_reads", 3.0), min_af=num("min_af", 0.0), GQ_threshold=num("GQ_threshold", 50.0),
             output_all_SNVs=flag("output_all_SNVs", False), pairs_file=a.get("pairs_file"),
             do_plots=a.get("do_plots"), extra_rob=flag("extra_rob", False),
             min_af_extra_rob=num("min_af_extra_rob", 0.2), min_prop_extra_rob=num("min_prop_extra_rob", 0.1),
             max_prop_extra_rob=num("max_prop_extra_rob", 0.8), afmin_power=num("afmin_power", -1.0),
             sigma=num("sigma", 0.1), source_path=a.get("source_path", ""), device=int(a.get("device", 0)),
             chunk=int(a.get("chunk_lines", 0)), chunk_bytes=int(a.get("chunk_bytes", 256 << 20)),
             writer=a.get("writer", "native"))
    if o["SB_type"] not in ("SOR", "RVSB", "FS"):
        raise SystemExit("SB_type must be SOR, RVSB or FS")
    return o


def read_names(path="names.txt"):
    """names.txt: two columns, file stem and sample name (bin/needlestack.r:155-160); make.unique(sep='_')"""
    rows = [ln.split() for ln in open(path) if ln.strip()]
    if any(len(r) != 2 for r in rows):
        raise SystemExit("Error : names.txt contains only one column : samples names can't be obtained from the "
                         "bam files : use --use_file_name option")
    names, seen = [], {}
    for _, nm in rows:
        if nm in seen:
            seen[nm] += 1
            nm = f"{nm}_{seen[nm]}"
        else:
            seen[nm] = 0
        names.append(nm)
    return [r[0] for r in rows], names


def read_pairs(path, names):
    """tumour/normal pairs file (bin/needlestack.r:162-185): header with TU*/NO* columns, NA for a missing mate"""
    rows = [ln.split() for ln in open(path) if ln.strip()]
    hdr = [h.upper() for h in rows[0]]
    it = next(i for i, h in enumerate(hdr) if "TU" in h)
    inn = next(i for i, h in enumerate(hdr) if "NO" in h)
    idx = {nm: i for i, nm in enumerate(names)}
    T, N, onlyT, onlyN = [], [], [], []
    for r in rows[1:]:
        t, nrm = r[it], r[inn]
        for nm in (t, nrm):
            if nm != "NA" and nm not in idx:
                raise SystemExit("Error : sample name not present in names.txt")
        if t != "NA" and nrm != "NA":
            T.append(idx[t])
            N.append(idx[nrm])
        elif t != "NA":
            onlyT.append(idx[t])
        elif nrm != "NA":
            onlyN.append(idx[nrm])
    return dict(tindex=T, nindex=N, only_t=sorted(onlyT), only_n=sorted(onlyN))


def read_fasta(path):
    seqs, name, buf = {}, None, []
    for ln in open(path):
        if ln.startswith(">"):
            if name is not None:
                seqs[name] = "".join(buf)
            name, buf = ln[1:].split()[0], []
        else:
            buf.append(ln.strip())
    if name is not None:
        seqs[name] = "".join(buf)
    return seqs


def context(fasta, chrom, start, end):
    """samtools faidx chr:start-end | tail -n1 (1-based, inclusive); empty string when out of range"""
    if fasta is None or chrom not in fasta:
        return "N" * max(0, end - start + 1)
    s = fasta[chrom]
    return s[max(0, start - 1):max(0, end)]


# (ID, Number, Type, Description) of the header records, in the reference's order (bin/needlestack.r:271-303)
_INFO = (("TYPE", "1", "String", "The type of allele, either snp, ins or del"),
         ("NS", "1", "Integer", "Number of samples with data"),
         ("DP", "1", "Integer", "Total read depth"),
         ("RO", "1", "Integer", "Total reference allele observation count"),
         ("AO", "1", "Integer", "Total alternate allele observation count"),
         ("AF", "1", "Float", "Estimated allele frequency in the range [0,1]"),
         ("SRF", "1", "Integer", "Total number of reference observations on the forward strand"),
         ("SRR", "1", "Integer", "Total number of reference observations on the reverse strand"),
         ("SAF", "1", "Integer", "Total number of alternate observations on the forward strand"),
         ("SAR", "1", "Integer", "Total number of alternate observations on the reverse strand"),
         ("SOR", "1", "Float", "Total Symmetric Odds Ratio of 2x2 contingency table to detect strand bias"),
         ("RVSB", "1", "Float", "Total Relative Variant Strand Bias"),
         ("FS", "1", "Float", "Total Fisher Exact Test p-value for detecting strand bias (Phred-scale)"),
         ("ERR", "1", "Float", "Estimated error rate for the alternate allele"),
         ("SIG", "1", "Float", "Estimated overdispersion for the alternate allele"),
         ("CONT", "1", "String", "Context of the reference sequence"),
         ("WARN", "1", "String", "Warning message when position is processed specifically"))
_FS = ("FS", "1", "Float", "Fisher Exact Test p-value for detecting strand bias (Phred-scale)")
_FORMAT = (("GT", "1", "String", "Genotype"),
           ("QVAL", "1", "Float", "Phred-scaled qvalue for not being an error"),
           ("QVAL_INV", "1", "Float", "Phred-scaled qvalue for not being an error in position where reference is switched"),
           ("DP", "1", "Integer", "Read Depth"),
           ("RO", "1", "Integer", "Reference allele observation count"),
           ("AO", "1", "Integer", "Alternate allele observation count"),
           ("AF", "1", "Float", "Allele fraction of the alternate allele with regard to reference"),
           ("SB", "4", "Integer", "Per-sample component statistics to detect strand bias as SRF,SRR,SAF,SAR"),
           ("SOR", "1", "Float", "Symmetric Odds Ratio of 2x2 contingency table to detect strand bias"),
           ("RVSB", "1", "Float", "Relative Variant Strand Bias"),
           _FS, _FS,  # the reference writes this record twice (bin/needlestack.r:300-301); kept
           ("QVAL_minAF", "1", "Float", "Phred-scaled qvalue for an allelic fraction of minAF"),
           ("STATUS", "1", "String", "Somatic status (SOMATIC, GERMLINE_CONFIRMED, GERMLINE_UNCONFIRMED, "
                                     "GERMLINE_UNCONFIRMABLE, , or UNKNOWN) of variants"))


def vcf_header(o, names, is_tn=False):
    """the header of bin/needlestack.r:263-305, same records in the same order"""
    sb = o["SB_type"]
    h = ["##fileformat=VCFv4.1", "##fileDate=" + __import__("time").strftime("%Y%m%d"), "##source=needlestack v1.1",
         "##reference=" + (o["fasta_ref"] or ""),
         "##phasing=none",
         f"##filter=\"QVAL > {rfmt(o['GQ_threshold'])} & {sb}_SNV < {rfmt(o['SB_threshold_SNV'])} & {sb}_INDEL < "
         f"{rfmt(o['SB_threshold_indel'])} & min(AO) >= {rfmt(o['min_reads'])} & min(AF) >= {rfmt(o['min_af'])} & "
         f"min(DP) >= {rfmt(o['min_coverage'])}\""]
    h += [f"##INFO=<ID={i},Number={nn},Type={t},Description=\"{d}\">" for i, nn, t, d in _INFO]
    h += [f"##FORMAT=<ID={i},Number={nn},Type={t},Description=\"{d}\">" for i, nn, t, d in _FORMAT]
    h.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names))
    return "\n".join(h) + "\n"


def _status_text(codes, names):
    """status codes -> text; POSSIBLE_CONTAMINATION_FROM_<samples with a GERMLINE*/UNKNOWN status> (:372-373)"""
    germ = [names[i] for i, c in enumerate(codes) if c in (2, 3, 4, 5)]
    out = []
    for c in codes:
        out.append("POSSIBLE_CONTAMINATION_FROM_" + "_".join(germ) if c == 6 else STATUS_TEXT[int(c)])
    return out


def vcf_line(chrom, pos, ref_txt, alt_txt, typ, row, dp, vp, vm, rp, rm, sigma, slope, flags, max_gq, o, fasta,
             names, before, after):
    """one VCF record exactly in the reference's field order (bin/needlestack.r:396-413)"""
    thr = o["GQ_threshold"]
    gq = row["gq"]
    ns = int((dp > 0).sum())
    ma = vp + vm
    pooled = row["pooled"]
    inv = bool(flags & _abi.NS_F_REF_INV)
    er = bool(flags & _abi.NS_F_EXTRA_ROB)
    warn = ";WARN=EXTRA_ROBUST_GL" if er and not inv else (";WARN=EXTRA_ROBUST_GL/INV_REF" if er and inv else
                                                            (";WARN=INV_REF" if inv else ""))
    with np.errstate(divide="ignore", invalid="ignore"):
        af_s = ma / dp
    info = (f"TYPE={typ};NS={ns};AF={rfmt(float((gq >= thr).sum()) / ns if ns else float('nan'), nan='NaN')};DP={rfmt(pooled[7])}"
            f";RO={rfmt(pooled[3] + pooled[4])};AO={rfmt(float(ma.sum()))};SRF={rfmt(pooled[3])};SRR={rfmt(pooled[4])}"
            f";SAF={rfmt(pooled[5])};SAR={rfmt(pooled[6])};SOR={rfmt(pooled[0])};RVSB={rfmt(pooled[1])}"
            f";FS={rfmt(pooled[2])};ERR={rfmt(slope)};SIG={rfmt(sigma)};CONT={before}x{after}{warn}")
    fmt = "GT:" + ("QVAL_INV" if inv else "QVAL") + ":DP:RO:AO:AF:SB:SOR:RVSB:FS:QVAL_minAF:STATUS"
    st = _status_text(row["status"], names)
    cols = [chrom, str(pos), ".", ref_txt, alt_txt, rfmt(max_gq), ".", info, fmt]
    # plain Python numbers from here on: numpy scalars cost more to index and convert than to format
    gtl, gql, afl = np.asarray(row["gt"]).tolist(), np.asarray(gq, dtype=float).tolist(), np.asarray(af_s, dtype=float).tolist()
    dpl, rpl, rml, vpl, vml = (np.asarray(a).tolist() for a in (dp, rp, rm, vp, vm))
    sorl, rvl, fsl, qml = (np.asarray(row[k], dtype=float).tolist() for k in ("sor", "rvsb", "fs", "qval_minaf"))
    for i in range(len(dpl)):
        cols.append(":".join([GT_TEXT[gtl[i]], rfmt(gql[i]), str(dpl[i]), str(rpl[i] + rml[i]),
                              str(vpl[i] + vml[i]), rfmt(afl[i], nan="NaN"), f"{rpl[i]},{rml[i]},{vpl[i]},{vml[i]}",
                              rfmt(sorl[i]), rfmt(rvl[i]), rfmt(fsl[i]), rfmt(qml[i]), st[i]]))
    return "\t".join(cols) + "\n"


def call_table(caller, T, o, names, pairs=None, fasta=None, caller_indel=None):
    """all VCF records of one parsed table chunk, in the reference's order: per line the SNV alleles,
    then the deletions, then the insertions"""
    n = T.n
    prm = api.make_params(pairs=pairs, min_coverage=o["min_coverage"], min_reads=o["min_reads"], min_af=o["min_af"],
                          extra_rob=int(o["extra_rob"]), min_af_extra_rob=o["min_af_extra_rob"],
                          min_prop_extra_rob=o["min_prop_extra_rob"], max_prop_extra_rob=o["max_prop_extra_rob"],
                          GQ_threshold=o["GQ_threshold"], SB_type=("SOR", "RVSB", "FS").index(o["SB_type"]),
                          SB_threshold_SNV=o["SB_threshold_SNV"], SB_threshold_indel=o["SB_threshold_indel"],
                          sigma_normal=o["sigma"], afmin_power=o["afmin_power"],
                          output_all_SNVs=int(o["output_all_SNVs"]), emit_mode=api.NS_EMIT_CALLS)
    # the indel alleles of the chunk go through a second context (its own stream) while the SNV planes are on the
    # first: a chunk holds few indel units, and one cluster-path unit among them otherwise adds its whole critical path
    indel_job = None
    if len(T.indel_pos):
        dp, vp, vm, rp, rm, ind = T.indel_unit_rows()
        if caller_indel is not None:
            import threading
            box = {}

            def run_indels():
                try:
                    box["I"] = caller_indel.call_units(dp, vp, vm, rp, rm, is_indel=ind, prm=prm)
                except BaseException as e:
                    box["err"] = e

            indel_job = threading.Thread(target=run_indels)
            indel_job.start()
    native = o.get("writer", "native") == "native"
    S = caller.call_snv(T.ref, T.counts, prm=prm, per_unit=not native)  # the C++ writer reads the per-row values
    if native:
        return _native_records(caller, T, S, o, names, fasta, prm, indel_job, box if indel_job is not None else None)
    if isinstance(fasta, FastaHandle):
        fasta = fasta.as_dict()
    recs = []  # (position row, order key, text)
    keys = ("gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled")
    for e, u in enumerate(S.emit_unit):
        p, k = divmod(int(u), 3)
        r = int(T.ref[p])
        alt = [b for b in range(4) if b != r][k]
        c = T.counts[p]
        chrom, pos = T.chrom_names[T.chrom[p]], int(T.loc[p])
        before = context(fasta, chrom, pos - 3, pos - 1)
        after = context(fasta, chrom, pos + 1, pos + 3)
        row = {kk: S[kk][e] for kk in keys}
        txt = vcf_line(chrom, pos, readcounts.BASES[r], readcounts.BASES[alt], "snv", row, c[0], c[1 + alt], c[5 + alt],
                       c[1 + r], c[5 + r], S.sigma[u], S.slope[u], int(S.flags[u]), S.max_gq[u], o, fasta, names,
                       before, after)
        recs.append((p, k, txt))
    if len(T.indel_pos):
        if indel_job is not None:
            indel_job.join()
            if "err" in box:
                raise box["err"]
            I = box["I"]
        else:
            I = caller.call_units(dp, vp, vm, rp, rm, is_indel=ind, prm=prm)
        for e, u in enumerate(I.emit_unit):
            u = int(u)
            p = int(T.indel_pos[u])
            chrom, pos = T.chrom_names[T.chrom[p]], int(T.loc[p])
            seq = T.indel_seq[u]
            row = {kk: I[kk][e] for kk in keys}
            if T.indel_type[u] == 1:  # deletion, bin/needlestack.r:541-549
                before = context(fasta, chrom, pos + 1 - 3, pos + 1 - 1).upper()
                after = context(fasta, chrom, pos + 1 + len(seq), pos + 1 + 3 + len(seq) - 1).upper()
                prev_bp = before[2:3]
                ref_txt, alt_txt, typ = prev_bp + seq, prev_bp, "del"
            else:  # insertion, bin/needlestack.r:705-712
                before = context(fasta, chrom, pos - 2, pos).upper()
                after = context(fasta, chrom, pos + 1, pos + 3).upper()
                prev_bp = before[2:3]
                ref_txt, alt_txt, typ = prev_bp, prev_bp + seq, "ins"
            txt = vcf_line(chrom, pos, ref_txt, alt_txt, typ, row, dp[u], vp[u], vm[u], rp[u], rm[u], I.sigma[u],
                           I.slope[u], int(I.flags[u]), I.max_gq[u], o, fasta, names, before, after)
            recs.append((p, 3 + int(T.indel_type[u]) * 100000 + u, txt))
    recs.sort(key=lambda r: (r[0], r[1]))
    return [r[2].encode() for r in recs]


class FastaHandle:
    """reference sequences loaded once by the library (ns_fasta_load) for the C++ VCF writer"""

    def __init__(self, path):
        self.path = path
        self._L = api.load_library()
        self.h = self._L.ns_fasta_load(path.encode())
        if not self.h:
            raise SystemExit("Error : cannot read --fasta_ref")

    def as_dict(self):
        return read_fasta(self.path)

    def close(self):
        if self.h:
            self._L.ns_fasta_free(self.h)
            self.h = None


def _native_records(caller, T, S, o, names, fasta, prm, indel_job, box):
    """the records of one table block from the C++ writer: SNV records come out in unit order (= input order); indel
    records (deletions before insertions per line) are interleaved by position"""
    fh = fasta.h if isinstance(fasta, FastaHandle) else None
    thr = o["GQ_threshold"]
    txt, off = api.vcf_records(T, S, names, fh, thr, indel=False)
    if not len(T.indel_pos):
        return [txt]
    if indel_job is not None:
        indel_job.join()
        if "err" in box:
            raise box["err"]
        I = box["I"]
    else:
        dp, vp, vm, rp, rm, ind = T.indel_unit_rows()
        I = caller.call_units(dp, vp, vm, rp, rm, is_indel=ind, prm=prm)
    if I.n_emitted == 0:
        return [txt]
    itxt, ioff = api.vcf_records(T, I, names, fh, thr, indel=True)
    su = np.asarray(S.emit_unit, dtype=np.int64)
    iu = np.asarray(I.emit_unit, dtype=np.int64)
    pos = np.concatenate([su // 3, T.indel_pos[iu]])
    key = np.concatenate([su % 3, 3 + T.indel_type[iu].astype(np.int64) * 100000 + iu])
    order = np.lexsort((key, pos))
    ns_ = len(su)
    out = []
    for j in order:
        out.append(txt[off[j]:off[j + 1]] if j < ns_ else itxt[ioff[j - ns_]:ioff[j - ns_ + 1]])
    return out


def main(argv=None, stdin=None):
    o = parse_args(sys.argv[1:] if argv is None else argv)
    if o["do_plots"] not in (None, "NONE", "FALSE"):
        sys.stderr.write("needlestack-b200: plots are out of scope of the GPU path; --do_plots ignored\n")
    _, names = read_names("names.txt")
    pairs = None
    if o["pairs_file"]:
        if not os.path.exists(o["pairs_file"]):
            raise SystemExit("Error : pairs_file not located in execution directory ")
        pairs = read_pairs(o["pairs_file"], names)
        if o["afmin_power"] == -1:
            o["afmin_power"] = 0.01  # bin/needlestack.r:166
    fasta = None
    if o["fasta_ref"] and os.path.exists(o["fasta_ref"]):
        fasta = FastaHandle(o["fasta_ref"]) if o["writer"] == "native" else read_fasta(o["fasta_ref"])
    src = (stdin or sys.stdin.buffer)
    header_line = src.readline()  # bin/needlestack.r:317
    if os.path.exists(o["out_file"]):
        os.remove(o["out_file"])
    n = len(names)
    # Three stages in flight: a reader thread cuts the input into blocks of whole lines and parses block k+1 (C++,
    # outside the GIL) while block k is on the GPU, and a writer thread appends the records of block k-1.
    import queue
    import threading
    q = queue.Queue(maxsize=1)
    wq = queue.Queue(maxsize=2)

    def whole_input():
        """the rest of the input as ONE uint8 view when it is addressable without reading it: a regular file on STDIN
        (`< table`, mmap-ed) or an in-memory buffer; None for a pipe"""
        try:
            if hasattr(src, "getbuffer"):
                return np.frombuffer(src.getbuffer(), dtype=np.uint8)[src.tell():]
            import mmap
            import stat
            fd = src.fileno()
            st = os.fstat(fd)
            if stat.S_ISREG(st.st_mode) and st.st_size > 0:
                mm = mmap.mmap(fd, 0, access=mmap.ACCESS_READ)
                return np.frombuffer(mm, dtype=np.uint8)[src.tell():]
        except Exception:
            pass
        return None

    def blocks():
        """whole-line blocks of the table: --chunk_lines lines each (the reference reads line by line; a line count
        makes the block boundaries independent of the text), else ~--chunk_bytes bytes: views into the input when it is
        a file or a buffer (parsed in place), read() calls when it is a pipe"""
        whole = whole_input() if o["chunk"] <= 0 else None
        if whole is not None:
            pos, end = 0, len(whole)
            while pos < end:
                cut = min(pos + o["chunk_bytes"], end)
                if cut < end:
                    back = max(pos, cut - (4 << 20))
                    nl = np.flatnonzero(whole[back:cut] == 10)
                    if len(nl):
                        cut = back + int(nl[-1]) + 1
                    else:  # a line longer than 4 MB: extend to its end
                        nxt = np.flatnonzero(whole[cut:] == 10)
                        cut = cut + int(nxt[0]) + 1 if len(nxt) else end
                yield whole[pos:cut]
                pos = cut
            return
        if o["chunk"] > 0:
            while True:
                lines = []
                for _ in range(o["chunk"]):
                    ln = src.readline()
                    if not ln:
                        break
                    lines.append(ln)
                if not lines:
                    return
                yield b"".join(lines)
        else:
            tail = b""
            while True:
                blk = src.read(o["chunk_bytes"])
                if not blk:
                    if tail:
                        yield tail
                    return
                cut = blk.rfind(b"\n") + 1
                if cut == 0:
                    tail += blk
                    continue
                yield (tail + blk[:cut]) if tail else (blk[:cut] if cut < len(blk) else blk)
                tail = blk[cut:]

    def reader():
        try:
            for blk in blocks():
                q.put(readcounts.load_table(blk, n, has_header=False))
            q.put(None)
        except BaseException as e:  # surfaced in the main thread
            q.put(e)

    werr = []

    def writer(out):
        try:
            while True:
                recs = wq.get()
                if recs is None:
                    return
                out.writelines(recs)
        except BaseException as e:
            werr.append(e)

    th = threading.Thread(target=reader, daemon=True)
    with open(o["out_file"], "ab") as out, api.Caller(o["device"]) as caller, api.Caller(o["device"]) as caller2:
        out.write(vcf_header(o, names, pairs is not None).encode())
        wt = threading.Thread(target=writer, args=(out,), daemon=True)
        th.start()
        wt.start()
        try:
            while True:
                T = q.get()
                if T is None:
                    break
                if isinstance(T, BaseException):
                    raise T
                wq.put(call_table(caller, T, o, names, pairs, fasta, caller_indel=caller2))
                T.close()
                if werr:
                    raise werr[0]
        finally:
            wq.put(None)
            wt.join()
        if werr:
            raise werr[0]
    th.join()
    if isinstance(fasta, FastaHandle):
        fasta.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
