#!/usr/bin/env Rscript
# needlestack.r on the B200 library: drop-in for the `Rscript needlestack.r` step of the reference pipeline
# (samtools mpileup | mpileup2readcounts | Rscript needlestack.r ...).  Same command line (--key=value, names and
# defaults of bin/needlestack.r:23-92), same inputs (readcounts table on STDIN, names.txt, optional pairs file), same
# VCF (header bin/needlestack.r:258-305, one line per variant with the fields of :396-413 / :549-567 / :712-730,
# written with R's own cat(), so the number format is the reference's by construction).
#
# What changed: the table is read in blocks of --chunk_lines lines instead of one line at a time; all SNV alleles of a
# block go through ONE .Call (C_ns_call_snv, or C_ns_multi_call_snv over every GPU of the box when --gpus > 1), all
# indel alleles of the block through one C_ns_call_units; the per-allele R loop only formats the variants the
# library reports (the units that pass the gates of :381/:387).  Regression plots are out of scope of the GPU
# path (--do_plots is accepted and ignored).  NOT RUN IN THIS IMAGE (R is not installed); the .Call routines it uses
# are compiled and driven through a stub of the R API in tests/test_rshim_stub.py.

opt_defaults <- list(samtools = "samtools", SB_type = "SOR", SB_threshold_SNV = 100, SB_threshold_indel = 100,
                     min_coverage = 30, min_reads = 3, min_af = 0, GQ_threshold = 50, output_all_SNVs = FALSE,
                     pairs_file = FALSE, extra_rob = FALSE, min_af_extra_rob = 0.2, min_prop_extra_rob = 0.1,
                     max_prop_extra_rob = 0.8, afmin_power = -1, sigma = 0.1, chunk_lines = 65536, gpus = 1)
raw_args <- commandArgs(TRUE)
kv <- strsplit(sub("^--", "", raw_args), "=")
opt <- opt_defaults
for (p in kv) opt[[p[1]]] <- if (length(p) > 1) p[2] else "TRUE"
if ("help" %in% names(opt) || is.null(opt$out_file) || is.null(opt$fasta_ref)) {
  cat("needlestack.r (B200 library): mandatory --out_file=vcf --fasta_ref=fasta; optional arguments as in the reference\n",
      "(--samtools --SB_type --SB_threshold_SNV --SB_threshold_indel --min_coverage --min_reads --min_af --GQ_threshold\n",
      " --output_all_SNVs --extra_rob --min_af_extra_rob --min_prop_extra_rob --max_prop_extra_rob --pairs_file\n",
      " --afmin_power --sigma) plus --chunk_lines=65536 --gpus=1 --source_path=dir\n")
  q(save = "no")
}
for (k in names(opt_defaults)) {
  d <- opt_defaults[[k]]
  if (is.numeric(d)) opt[[k]] <- as.numeric(opt[[k]])
  if (is.logical(d) && k != "pairs_file") opt[[k]] <- as.logical(opt[[k]])
}
if (!(opt$SB_type %in% c("SOR", "RVSB", "FS"))) stop("SB_type must be SOR, RVSB or FS")
options(scipen = 100)                                        # bin/needlestack.r:153

script_dir <- function() {
  m <- regmatches(commandArgs(), regexpr("(?<=^--file=).+", commandArgs(), perl = TRUE))
  if (length(m) != 1) stop("can't determine script dir: please call the script with Rscript")
  dirname(m)
}
src <- if (is.null(opt$source_path)) paste0(script_dir(), "/") else opt$source_path
source(paste0(src, "glm_rob_nb.r"))                          # the drop-in glmrob.nb() + .ns_ctx()

# ---- samples and tumour / normal pairs (bin/needlestack.r:155-185) ----------------------------------------
samples <- read.table("names.txt", stringsAsFactors = FALSE, colClasses = "character")
if (ncol(samples) != 2) {
  cat("Error : names.txt contains only one column : samples names can't be obtained from the bam files : use --use_file_name option")
  q(save = "no")
}
sample_names <- make.unique(samples[, 2], sep = "_")
n <- length(sample_names)
tn <- NULL
afmin_power <- opt$afmin_power
if (!identical(opt$pairs_file, FALSE)) {
  if (!file.exists(opt$pairs_file)) { cat("Error : pairs_file not located in execution directory "); q(save = "no") }
  if (afmin_power == -1) afmin_power <- 0.01
  pf <- read.table(opt$pairs_file, header = TRUE, stringsAsFactors = FALSE, na.strings = "NA")
  names(pf)[grep("TU", names(pf), ignore.case = TRUE)] <- "TUMOR"
  names(pf)[grep("NO", names(pf), ignore.case = TRUE)] <- "NORMAL"
  listed <- c(pf$NORMAL[!is.na(pf$NORMAL)], pf$TUMOR[!is.na(pf$TUMOR)])
  if (!all(listed %in% sample_names)) { cat("Error : sample name not present in names.txt"); q(save = "no") }
  both <- !is.na(pf$TUMOR) & !is.na(pf$NORMAL)
  tn <- list(match(pf$TUMOR[both], sample_names) - 1L, match(pf$NORMAL[both], sample_names) - 1L,
             which(sample_names %in% pf$TUMOR[is.na(pf$NORMAL)]) - 1L,
             which(sample_names %in% pf$NORMAL[is.na(pf$TUMOR)]) - 1L)   # 0-based for the library
}

# ---- library contexts ------------------------------------------------------------------------------------
ctx <- .ns_ctx()
mctx <- if (opt$gpus > 1) .Call("C_ns_create_multi", as.integer(seq_len(opt$gpus) - 1L)) else NULL
lib_params <- list(min_coverage = opt$min_coverage, min_reads = opt$min_reads, min_af = opt$min_af,
                   extra_rob = as.integer(opt$extra_rob), min_af_extra_rob = opt$min_af_extra_rob,
                   min_prop_extra_rob = opt$min_prop_extra_rob, max_prop_extra_rob = opt$max_prop_extra_rob,
                   GQ_threshold = opt$GQ_threshold, SB_type = match(opt$SB_type, c("SOR", "RVSB", "FS")) - 1L,
                   SB_threshold_SNV = opt$SB_threshold_SNV, SB_threshold_indel = opt$SB_threshold_indel,
                   sigma_normal = opt$sigma, afmin_power = afmin_power, output_all_SNVs = as.integer(opt$output_all_SNVs))

# ---- VCF header (bin/needlestack.r:258-305) -----------------------------------------------------------------
out_file <- opt$out_file
if (file.exists(out_file)) invisible(file.remove(out_file))
emit <- function(...) cat(..., sep = "", file = out_file, append = TRUE)
rec <- function(kind, id, number, type, desc)
  emit("##", kind, "=<ID=", id, ",Number=", number, ",Type=", type, ",Description=\"", desc, "\">\n")
emit("##fileformat=VCFv4.1\n##fileDate=", format(Sys.Date(), "%Y%m%d"), "\n##source=needlestack v1.1\n##reference=",
     opt$fasta_ref, "\n##phasing=none\n##filter=\"QVAL > ", opt$GQ_threshold, " & ", opt$SB_type, "_SNV < ",
     opt$SB_threshold_SNV, " & ", opt$SB_type, "_INDEL < ", opt$SB_threshold_indel, " & min(AO) >= ", opt$min_reads,
     " & min(AF) >= ", opt$min_af, " & min(DP) >= ", opt$min_coverage, "\"\n")
info_recs <- rbind(c("TYPE", "String", "The type of allele, either snp, ins or del"),
  c("NS", "Integer", "Number of samples with data"), c("DP", "Integer", "Total read depth"),
  c("RO", "Integer", "Total reference allele observation count"), c("AO", "Integer", "Total alternate allele observation count"),
  c("AF", "Float", "Estimated allele frequency in the range [0,1]"),
  c("SRF", "Integer", "Total number of reference observations on the forward strand"),
  c("SRR", "Integer", "Total number of reference observations on the reverse strand"),
  c("SAF", "Integer", "Total number of alternate observations on the forward strand"),
  c("SAR", "Integer", "Total number of alternate observations on the reverse strand"),
  c("SOR", "Float", "Total Symmetric Odds Ratio of 2x2 contingency table to detect strand bias"),
  c("RVSB", "Float", "Total Relative Variant Strand Bias"),
  c("FS", "Float", "Total Fisher Exact Test p-value for detecting strand bias (Phred-scale)"),
  c("ERR", "Float", "Estimated error rate for the alternate allele"),
  c("SIG", "Float", "Estimated overdispersion for the alternate allele"),
  c("CONT", "String", "Context of the reference sequence"),
  c("WARN", "String", "Warning message when position is processed specifically"))
for (i in seq_len(nrow(info_recs))) rec("INFO", info_recs[i, 1], "1", info_recs[i, 2], info_recs[i, 3])
fs_desc <- "Fisher Exact Test p-value for detecting strand bias (Phred-scale)"
format_recs <- rbind(c("GT", "1", "String", "Genotype"), c("QVAL", "1", "Float", "Phred-scaled qvalue for not being an error"),
  c("QVAL_INV", "1", "Float", "Phred-scaled qvalue for not being an error in position where reference is switched"),
  c("DP", "1", "Integer", "Read Depth"), c("RO", "1", "Integer", "Reference allele observation count"),
  c("AO", "1", "Integer", "Alternate allele observation count"),
  c("AF", "1", "Float", "Allele fraction of the alternate allele with regard to reference"),
  c("SB", "4", "Integer", "Per-sample component statistics to detect strand bias as SRF,SRR,SAF,SAR"),
  c("SOR", "1", "Float", "Symmetric Odds Ratio of 2x2 contingency table to detect strand bias"),
  c("RVSB", "1", "Float", "Relative Variant Strand Bias"),
  c("FS", "1", "Float", fs_desc), c("FS", "1", "Float", fs_desc),          # written twice by the reference (:300-301): kept
  c("QVAL_minAF", "1", "Float", "Phred-scaled qvalue for an allelic fraction of minAF"),
  c("STATUS", "1", "String", "Somatic status (SOMATIC, GERMLINE_CONFIRMED, GERMLINE_UNCONFIRMED, GERMLINE_UNCONFIRMABLE, , or UNKNOWN) of variants"))
for (i in seq_len(nrow(format_recs))) rec("FORMAT", format_recs[i, 1], format_recs[i, 2], format_recs[i, 3], format_recs[i, 4])
emit("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t", paste(sample_names, collapse = "\t"), "\n")

# ---- reference context: samtools faidx, as the reference (:389-394), cached per region -------------------------
faidx <- function(chr, from, to) {
  con <- pipe(paste(opt$samtools, " faidx ", opt$fasta_ref, " ", chr, ":", from, "-", to, " | tail -n1", sep = ""))
  on.exit(close(con))
  readLines(con)
}

gt_text <- c("0/0", "0/1", "1/1", "./.")
status_text <- c(".", "SOMATIC", "GERMLINE_CONFIRMED", "GERMLINE_UNCONFIRMED", "GERMLINE_UNCONFIRMABLE", "UNKNOWN")
bases <- c("A", "T", "C", "G")                                # plane order of the table (:188-198)

# one VCF line from the library's row e of result `res` (matrices are n x E); counts are the unit's five rows
vcf_line <- function(chr, pos, ref_txt, alt_txt, type, res, e, u, DP, Vp, Vm, Rp, Rm, before, after) {
  flags <- res$flags[u]
  inv <- bitwAnd(flags, 4L) != 0
  er <- bitwAnd(flags, 2L) != 0
  gq <- res$gq[, e]
  pooled <- res$pooled[, e]
  ma <- Vp + Vm
  st_codes <- as.integer(res$status[, e])
  st <- status_text[pmin(st_codes, 5L) + 1L]
  if (any(st_codes == 6L))
    st[st_codes == 6L] <- paste0("POSSIBLE_CONTAMINATION_FROM_", paste(sample_names[st_codes %in% 2:5], collapse = "_"))
  warn <- if (er && !inv) ";WARN=EXTRA_ROBUST_GL" else if (er && inv) ";WARN=EXTRA_ROBUST_GL/INV_REF" else if (inv) ";WARN=INV_REF" else ""
  na_txt <- function(v) ifelse(is.nan(v), NA, v)             # the library reports R's NA as NaN
  emit(chr, "\t", pos, "\t.\t", ref_txt, "\t", alt_txt, "\t", na_txt(res$max_gq[u]), "\t.\t",
       "TYPE=", type, ";NS=", sum(DP > 0), ";AF=", sum(gq >= opt$GQ_threshold) / sum(DP > 0), ";DP=", pooled[8],
       ";RO=", pooled[4] + pooled[5], ";AO=", sum(ma), ";SRF=", pooled[4], ";SRR=", pooled[5], ";SAF=", pooled[6],
       ";SAR=", pooled[7], ";SOR=", pooled[1], ";RVSB=", pooled[2], ";FS=", pooled[3], ";ERR=", res$slope[u],
       ";SIG=", res$sigma[u], ";CONT=", before, "x", after, warn,
       "\tGT:", if (inv) "QVAL_INV" else "QVAL", ":DP:RO:AO:AF:SB:SOR:RVSB:FS:QVAL_minAF:STATUS")
  gt <- gt_text[as.integer(res$gt[, e]) + 1L]
  cols <- paste(gt, gq, DP, Rp + Rm, ma, ma / DP, paste(Rp, Rm, Vp, Vm, sep = ","), na_txt(res$sor[, e]),
                na_txt(res$rvsb[, e]), na_txt(res$fs[, e]), na_txt(res$qval_minaf[, e]), st, sep = ":")
  emit("\t", paste(cols, collapse = "\t"), "\n")
}

call_rows <- function(routine, handle, counts, ref_codes, cap) {
  repeat {
    res <- .Call(routine, handle, counts, ref_codes, lib_params, tn, cap)
    if (res$n_emitted <= cap) return(res)
    cap <- res$n_emitted                                       # NS_ERR_EMIT_OVERFLOW: ask again with what is needed
  }
}

# ---- main loop: blocks of lines (the reference reads one line at a time, :316) ----------------------------------
con <- file("stdin")
open(con)
invisible(readLines(con, n = 1))                             # header line (:317)
count_cols <- as.vector(outer(0:8, (seq_len(n) - 1L) * 11L + 4L, `+`))   # depth,A,T,C,G,a,t,c,g of every sample
ins_cols <- seq_len(n) * 11L - 1L + 3L
del_cols <- seq_len(n) * 11L + 3L
while (length(lines <- readLines(con, n = opt$chunk_lines, warn = FALSE)) > 0) {
  f <- do.call(rbind, strsplit(lines, "\t", fixed = TRUE))  # P x (3 + 11 n)
  P <- nrow(f)
  counts <- array(as.integer(t(f[, count_cols])), dim = c(9L, n, P))
  counts <- aperm(counts, c(2L, 1L, 3L))                    # [n, 9, P]: column-major memory == the ABI's [P][9][n]
  ref_code <- match(f[, 3], bases) - 1L
  ref_code[is.na(ref_code)] <- 4L                           # REF not in A,T,C,G: the line is skipped (:319)
  res <- call_rows(if (is.null(mctx)) "C_ns_call_snv" else "C_ns_multi_call_snv", if (is.null(mctx)) ctx else mctx,
                   counts, as.raw(ref_code), max(1024, 3 * P %/% 16))
  E <- res$n_emitted
  snv_pos <- if (E > 0) res$emit_unit[seq_len(E)] %/% 3 + 1 else numeric(0)   # 1-based line of every SNV variant
  # indel candidate alleles of the block (bin/needlestack.r:461-479, :624-643): few lines carry any
  indel_lines <- which(ref_code < 4L & (rowSums(f[, ins_cols, drop = FALSE] != "NA") > 0 |
                                        rowSums(f[, del_cols, drop = FALSE] != "NA") > 0))
  iu <- list()
  for (p in indel_lines) for (kind in c("del", "ins")) {
    cells <- f[p, if (kind == "del") del_cols else ins_cols]
    ent <- strsplit(cells[cells != "NA"], "|", fixed = TRUE)
    seqs <- unique(toupper(unlist(lapply(ent, function(x) sub("^[0-9]+:", "", x)))))
    for (s in seqs) {
      grab <- function(spell) vapply(cells, function(cell) {
        if (cell == "NA") return(0L)
        e <- strsplit(cell, "|", fixed = TRUE)[[1]]
        hit <- e[sub("^[0-9]+:", "", e) == spell]
        if (length(hit) == 0) 0L else as.integer(sub(":.*$", "", hit[1]))
      }, 0L, USE.NAMES = FALSE)
      iu[[length(iu) + 1L]] <- list(p = p, kind = kind, seq = s, vp = grab(s), vm = grab(tolower(s)))
    }
  }
  ires <- NULL
  if (length(iu) > 0) {
    ip <- vapply(iu, function(x) x$p, 0L)
    rc <- ref_code[ip] + 1L
    dpm <- vapply(seq_along(iu), function(k) counts[, 1L, ip[k]], integer(n))
    rpm <- vapply(seq_along(iu), function(k) counts[, 1L + rc[k], ip[k]], integer(n))
    rmm <- vapply(seq_along(iu), function(k) counts[, 5L + rc[k], ip[k]], integer(n))
    vpm <- vapply(iu, function(x) x$vp, integer(n))
    vmm <- vapply(iu, function(x) x$vm, integer(n))
    pr <- lib_params
    pr$emit_mode <- 0L                                         # NS_EMIT_CALLS rows only
    ires <- .Call("C_ns_call_units_rows", ctx, dpm, vpm, vmm, rpm, rmm, as.raw(rep(1L, length(iu))), pr, tn)
  }
  # write the block's variants in the reference's order: per line the SNV alleles, then deletions, then insertions
  ie <- if (is.null(ires)) integer(0) else seq_len(ires$n_emitted)
  ikey <- if (length(ie)) vapply(ires$emit_unit[ie] + 1, function(u) iu[[u]]$p + ifelse(iu[[u]]$kind == "del", 0.3, 0.6), 0) else numeric(0)
  ord <- order(c(snv_pos + 0.1 * ((res$emit_unit[seq_len(E)] %% 3) / 3), ikey))
  for (j in ord) {
    if (j <= E) {
      u <- res$emit_unit[j] + 1
      p <- (u - 1) %/% 3 + 1
      k <- (u - 1) %% 3 + 1
      r <- ref_code[p] + 1L
      a <- setdiff(1:4, r)[k]
      pos <- as.numeric(f[p, 2])
      vcf_line(f[p, 1], f[p, 2], bases[r], bases[a], "snv", res, j, u, counts[, 1, p], counts[, 1 + a, p],
               counts[, 5 + a, p], counts[, 1 + r, p], counts[, 5 + r, p],
               faidx(f[p, 1], pos - 3, pos - 1), faidx(f[p, 1], pos + 1, pos + 3))
    } else {
      e <- j - E
      x <- iu[[ires$emit_unit[e] + 1]]
      p <- x$p
      r <- ref_code[p] + 1L
      pos <- as.numeric(f[p, 2])
      L <- nchar(x$seq)
      if (x$kind == "del") {                                  # bin/needlestack.r:541-549
        before <- toupper(faidx(f[p, 1], pos + 1 - 3, pos + 1 - 1))
        after <- toupper(faidx(f[p, 1], pos + 1 + L, pos + 1 + 3 + L - 1))
        prev <- substr(before, 3, 3)
        vcf_line(f[p, 1], f[p, 2], paste0(prev, x$seq), prev, "del", ires, e, ires$emit_unit[e] + 1, counts[, 1, p], x$vp, x$vm,
                 counts[, 1 + r, p], counts[, 5 + r, p], before, after)
      } else {                                                # bin/needlestack.r:705-712
        before <- toupper(faidx(f[p, 1], pos - 2, pos))
        after <- toupper(faidx(f[p, 1], pos + 1, pos + 3))
        prev <- substr(before, 3, 3)
        vcf_line(f[p, 1], f[p, 2], prev, paste0(prev, x$seq), "ins", ires, e, ires$emit_unit[e] + 1, counts[, 1, p], x$vp, x$vm,
                 counts[, 1 + r, p], counts[, 5 + r, p], before, after)
      }
    }
  }
}
close(con)
