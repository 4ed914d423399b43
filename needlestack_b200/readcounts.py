"""Columnar loader for the mpileup2readcounts text table (the STDIN of bin/needlestack.r).

`load_table` wraps ns_table_parse (C++, multi-threaded, in libns_b200.so): one pass over the text into
the [P][9][n] int32 planes that ns_call_snv takes (page-locked when a GPU is present), the reference
base codes, and the indel candidate alleles of bin/needlestack.r:461-479 / :624-643.
"""
import ctypes as C
import os

import numpy as np

from . import api

BASES = "ATCG"


class Table:
    """Parsed table.  `counts` is a VIEW into memory owned by this object (keep it alive)."""

    def __init__(self, handle, lib):
        self._h, self._L = handle, lib
        L = lib
        self.P = int(L.ns_table_positions(handle))
        self.n = int(L.ns_table_samples(handle))
        P, n = self.P, self.n

        def view(ptr, shape, dtype):
            cnt = int(np.prod(shape))
            if cnt == 0 or not ptr:
                return np.zeros(shape, dtype=dtype)
            buf = (C.c_char * (cnt * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=cnt).reshape(shape)

        self.counts = view(L.ns_table_counts(handle), (P, 9, n), np.int32)
        self.counts_pinned = bool(L.ns_table_counts_pinned(handle))
        self.ref = view(L.ns_table_ref(handle), (P,), np.uint8).copy()
        self.ref_text = view(L.ns_table_ref_text(handle), (P,), "S1").copy()
        self.loc = view(L.ns_table_loc(handle), (P,), np.int64).copy()
        self.chrom = view(L.ns_table_chrom(handle), (P,), np.int32).copy()
        self.chrom_names = [L.ns_table_chrom_name(handle, i).decode() for i in range(L.ns_table_n_chrom(handle))]
        self.indel_cov = view(L.ns_table_indel_cov(handle), (P, 2), np.int64).copy()
        Ui = int(L.ns_table_indel_units(handle))
        self.indel_pos = view(L.ns_table_indel_pos(handle), (Ui,), np.int64).copy()
        self.indel_type = view(L.ns_table_indel_type(handle), (Ui,), np.uint8).copy()
        self.indel_seq = [L.ns_table_indel_seq(handle, i).decode() for i in range(Ui)]
        self.indel_vp = view(L.ns_table_indel_vp(handle), (Ui, n), np.int32).copy()
        self.indel_vm = view(L.ns_table_indel_vm(handle), (Ui, n), np.int32).copy()

    def indel_unit_rows(self):
        """(dp, vp, vm, rp, rm, is_indel) for ns_call_units: depth and REF-base rows of each unit's position
        (the regression's DP is the plain depth column; on ref-inversion y is the REF base count,
        bin/needlestack.r:481-488)"""
        pos = self.indel_pos
        r = self.ref[pos].astype(np.int64)
        c = self.counts
        dp = c[pos, 0]
        rp = c[pos, 1 + r]
        rm = c[pos, 5 + r]
        return dp, self.indel_vp, self.indel_vm, rp, rm, np.ones(len(pos), dtype=np.uint8)

    def close(self):
        if self._h:
            self.counts = None
            self._L.ns_table_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_table(source, n_samples, has_header=True, n_threads=None):
    """source: path, bytes or str holding the whole table, or a uint8 numpy array over its text (parsed in place: a
    view into an mmap-ed file or a BytesIO buffer, no copy)"""
    keep = None
    if isinstance(source, np.ndarray):
        keep = np.ascontiguousarray(source, dtype=np.uint8)
        data, size = C.c_void_p(keep.ctypes.data), keep.size
    else:
        if isinstance(source, (bytes, bytearray)):
            data = bytes(source)
        elif isinstance(source, str) and ("\t" in source or "\n" in source):
            data = source.encode()
        else:
            with open(source, "rb") as f:
                data = f.read()
        size = len(data)
    L = api.load_library()
    err = C.create_string_buffer(512)
    h = L.ns_table_parse(data, size, int(n_samples), int(bool(has_header)),
                         int(n_threads or os.cpu_count() or 1), err, 512)
    if not h:
        raise api.NeedlestackError(err.value.decode() or "ns_table_parse failed")
    return Table(h, L)


def load_table_python(text, n_samples, has_header=True):
    """Plain-Python restatement of the same parse, following bin/needlestack.r line by line — used by the
    tests as the checker of the C++ loader (small inputs only)."""
    lines = text.decode().split("\n") if isinstance(text, (bytes, bytearray)) else text.split("\n")
    if has_header:
        lines = lines[1:]
    rows = [ln.rstrip("\r").split("\t") for ln in lines if ln.strip("\r") != ""]
    n = n_samples
    P = len(rows)
    counts = np.zeros((P, 9, n), dtype=np.int32)
    ref = np.full(P, 4, dtype=np.uint8)
    units = []
    for p, f in enumerate(rows):
        if f[2] in BASES and len(f[2]) == 1:
            ref[p] = BASES.index(f[2])
        for s in range(n):
            base = 3 + 11 * s
            counts[p, :, s] = [int(v) for v in f[base:base + 9]]
        if ref[p] > 3:
            continue
        for typ, col in ((1, 10), (2, 9)):  # deletions first (needlestack.r:461), then insertions (:624)
            cells = [f[3 + 11 * s + col] for s in range(n)]
            entries = [(s, e.split(":")) for s, cell in enumerate(cells) if cell != "NA" for e in cell.split("|")]
            uniq = []
            for _, (cnt, seq) in entries:
                if seq.upper() not in uniq:
                    uniq.append(seq.upper())
            for sq in uniq:
                vp = np.zeros(n, dtype=np.int32)
                vm = np.zeros(n, dtype=np.int32)
                for s, (cnt, seq) in entries:
                    if seq == sq:
                        vp[s] = int(cnt)
                    elif seq == sq.lower():
                        vm[s] = int(cnt)
                units.append((p, typ, sq, vp, vm))
    return dict(counts=counts, ref=ref, units=units,
                loc=np.array([int(f[1]) for f in rows], dtype=np.int64), chrom=[f[0] for f in rows])
