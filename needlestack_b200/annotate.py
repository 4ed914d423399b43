"""annotate_vcf mode: the second caller of glmrob.nb (/root/reference/bin/annotate_vcf.r), on the batch API.

    python -m needlestack_b200.annotate --input_vcf=in.vcf[.gz|.bgz] [--out_vcf=out.vcf] [--min_coverage=50] ...

Same `--key=value` options and defaults as the reference script (:19-40).  For every VCF line and every ALT
allele the DP / AD FORMAT fields of all samples become one unit (:57-69): x = DP (missing -> 0), y = AD[alt],
and, when more than half of the samples have y/x > 0.8, y = DP - sum(AD[alts]) (reference switching, :65-68).
All units of the file go through ONE `ns_call_units` call (the reference: one glmrob.nb() R call per unit);
the annotations are the reference's (:103-141): INFO ERR, SIG, WARN and FORMAT QVAL, QVAL_INV.

Not reproduced: the per-variant PDF plots (`--do_plots`; the numbers behind their contours are
`needlestack_b200.contours`), and the reference's quirk of re-writing the header for every chunk of
`chunk_size` lines (writeVcf in append mode, :143-146) - the header is written once.  VariantAnnotation's exact
number formatting could not be checked (no R in this image): values are written with 15 significant digits.
"""
import gzip
import sys

import numpy as np

from . import _abi
from .api import Caller, make_params

DEFAULTS = dict(chunk_size=1000, min_coverage=50, min_reads=5, min_af=0, GQ_threshold=50, SB_threshold=100,
                extra_rob=False, min_af_extra_rob=0.2, min_prop_extra_rob=0.1, max_prop_extra_rob=0.5)

NEW_HEADER = [
    '##INFO=<ID=ERR,Number=A,Type=Float,Description="Error rate estimated by needlestack">',
    '##INFO=<ID=SIG,Number=A,Type=Float,Description="Dispertion parameter estimated by needlestack">',
    '##INFO=<ID=WARN,Number=A,Type=Character,Description="Warning message when position is processed specifically '
    'by needlestack">',
    '##FORMAT=<ID=QVAL,Number=A,Type=Float,Description="Phred q-values computed by needlestack">',
    '##FORMAT=<ID=QVAL_INV,Number=A,Type=Float,Description="Phred q-values computed by needlestack at a position '
    'where ">',
]


def parse_args(argv):
    """annotate_vcf.r:19-40"""
    o = dict(DEFAULTS)
    raw = {}
    for a in argv:
        k, _, v = a[2:].partition("=") if a.startswith("--") else a.partition("=")
        raw[k] = v
    if "input_vcf" not in raw:
        raise SystemExit("no input VCF file")
    o["input_vcf"] = raw["input_vcf"]
    out = raw.get("out_vcf", "")
    o["out_vcf"] = out if out else o["input_vcf"].replace(".vcf.bgz", "") + "_annotated_needlestack.vcf"
    for k in ("chunk_size", "min_coverage", "min_reads", "min_af", "GQ_threshold", "SB_threshold", "min_af_extra_rob",
              "min_prop_extra_rob", "max_prop_extra_rob"):
        if k in raw:
            o[k] = float(raw[k])
    if "extra_rob" in raw:
        o["extra_rob"] = raw["extra_rob"].upper() in ("TRUE", "T")
    return o


def _open(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def read_vcf(path):
    """header lines, sample names and the records as lists of fields"""
    header, names, recs = [], [], []
    with _open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith("##"):
                header.append(line)
            elif line.startswith("#"):
                names = line.split("\t")[9:]
                header.append(line)
            elif line:
                recs.append(line.split("\t"))
    return header, names, recs


def _int(s):
    return int(s) if s not in (".", "") else None


def record_units(fields, n):
    """annotate_vcf.r:57-62: (DP[n], AD[n][L]) of one line; missing DP -> 0, missing / empty AD -> zeros;
    L = max(lengths(AD)) (2 when every AD is missing)"""
    keys = fields[8].split(":")
    i_dp = keys.index("DP") if "DP" in keys else -1
    i_ad = keys.index("AD") if "AD" in keys else -1
    dp = np.zeros(n, dtype=np.int64)
    ads = []
    for s in range(n):
        parts = fields[9 + s].split(":") if 9 + s < len(fields) else []
        v = _int(parts[i_dp]) if 0 <= i_dp < len(parts) else None
        dp[s] = v if v is not None else 0
        ad = []
        if 0 <= i_ad < len(parts) and parts[i_ad] not in (".", ""):
            ad = [_int(x) or 0 for x in parts[i_ad].split(",")]
        ads.append(ad)
    L = max([len(a) for a in ads] + [0])
    if L == 0:
        L = 2
    AD = np.zeros((n, L), dtype=np.int64)
    for s, a in enumerate(ads):
        AD[s, :len(a)] = a
    return dp, AD


def build_units(recs, n):
    """rows [U][n] for ns_call_units: depth, the allele's count, and DP - sum(alt counts) as the count the
    reference-switching test falls back to; owner[u] = (record, alt index)"""
    dps, ys, rps, owner = [], [], [], []
    for r, fields in enumerate(recs):
        dp, AD = record_units(fields, n)
        inv = dp - AD[:, 1:].sum(axis=1)
        for a in range(1, AD.shape[1]):
            dps.append(dp)
            ys.append(AD[:, a])
            rps.append(inv)
            owner.append((r, a - 1))
    if not dps:
        z = np.zeros((0, n), dtype=np.int32)
        return z, z, z, owner
    rp = np.stack(rps)
    if (rp < 0).any():
        sys.stderr.write("annotate: %d cells have sum(AD[alts]) > DP; the switched count is clamped to 0\n" % int((rp < 0).sum()))
        rp = np.maximum(rp, 0)
    return np.stack(dps).astype(np.int32), np.stack(ys).astype(np.int32), rp.astype(np.int32), owner


def fmt(v):
    if v != v:
        return "."
    s = "%.15g" % (v + 0.0)
    return s


def annotate_records(caller, recs, n, o):
    """the per-unit results and the annotated lines"""
    dp, y, rp, owner = build_units(recs, n)
    U = dp.shape[0]
    per_rec = [[] for _ in recs]
    if U:
        prm = make_params(emit_mode=_abi.NS_EMIT_ALL, min_coverage=o["min_coverage"], min_reads=o["min_reads"],
                          min_af=o["min_af"], extra_rob=int(bool(o["extra_rob"])), min_af_extra_rob=o["min_af_extra_rob"],
                          min_prop_extra_rob=o["min_prop_extra_rob"], max_prop_extra_rob=o["max_prop_extra_rob"],
                          GQ_threshold=o["GQ_threshold"])
        R = caller.call_units(dp, y, rp=rp, prm=prm, emit_cap=U)
        gq = R.dense("gq")
        for u, (r, _) in enumerate(owner):
            per_rec[r].append(dict(sigma=float(R["sigma"][u]), slope=float(R["slope"][u]), gq=gq[u],
                                   inv_ref=bool(R["flags"][u] & _abi.NS_F_REF_INV),
                                   extra_rob=bool(R["flags"][u] & _abi.NS_F_EXTRA_ROB),
                                   flags=int(R["flags"][u])))
    lines = []
    for fields, regs in zip(recs, per_rec):
        f = list(fields)
        if not regs:
            lines.append("\t".join(f))
            continue
        inv = [g["inv_ref"] for g in regs]
        ext = [g["extra_rob"] for g in regs]
        warn = None  # annotate_vcf.r:105-120, the three assignments in the reference's order
        if any(ext):
            warn = ",".join("EXTRA_ROBUST_GL" if e else "." for e in ext)
        if any(inv):
            warn = ",".join("INV_REF" if e else "." for e in inv)
        if any(inv) and all(inv):
            warn = ",".join("EXTRA_ROBUST_GL/INV_REF" for _ in inv)
        info = [] if f[7] in (".", "") else [f[7]]
        info.append("ERR=" + ",".join(fmt(g["slope"]) for g in regs))
        info.append("SIG=" + ",".join(fmt(g["sigma"]) for g in regs))
        if warn is not None:
            info.append("WARN=" + warn)
        f[7] = ";".join(info)
        f[8] = f[8] + ":QVAL:QVAL_INV"
        anyinv = any(inv)
        for s in range(n):
            # :124-135 - on lines without any switched allele BOTH fields carry the q-values (as the reference does)
            q = ",".join("." if (anyinv and iv) else fmt(g["gq"][s]) for g, iv in zip(regs, inv))
            qi = ",".join("." if (anyinv and not iv) else fmt(g["gq"][s]) for g, iv in zip(regs, inv))
            if 9 + s < len(f):
                f[9 + s] = f[9 + s] + ":" + q + ":" + qi
        lines.append("\t".join(f))
    return lines, per_rec


def annotate_vcf(caller, input_vcf, out_vcf, o=None):
    o = dict(DEFAULTS, **(o or {}))
    header, names, recs = read_vcf(input_vcf)
    n = len(names)
    lines, per_rec = annotate_records(caller, recs, n, o)
    with open(out_vcf, "w") as f:
        for h in header[:-1]:
            f.write(h + "\n")
        for h in NEW_HEADER:
            f.write(h + "\n")
        if header:
            f.write(header[-1] + "\n")
        for ln in lines:
            f.write(ln + "\n")
    return per_rec


def main(argv=None):
    o = parse_args(sys.argv[1:] if argv is None else argv)
    with Caller(0) as c:
        annotate_vcf(c, o["input_vcf"], o["out_vcf"], o)
    return 0


if __name__ == "__main__":
    sys.exit(main())
