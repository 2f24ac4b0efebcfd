"""Helpers around the CPU oracle (oracle/) — the checker, never the thing under test."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "build", "libngsep_oracle.so")
ORACLE_BIN = os.path.join(ORACLE_DIR, "build", "ngsep_oracle")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    lib = C.CDLL(ORACLE_SO)
    ip = C.POINTER(C.c_int)
    lib.oracle_snv_column.argtypes = [ip, C.c_int, C.c_int, C.c_char, C.c_char, C.c_double, C.c_int, C.c_int, ip, ip, ip, ip, ip, C.c_char_p]
    lib.oracle_table_error.restype = C.c_double
    lib.oracle_table_error.argtypes = [C.c_int, C.c_int]
    lib.oracle_table_gt.restype = C.c_double
    lib.oracle_table_gt.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.oracle_fisher.restype = C.c_double
    lib.oracle_fisher.argtypes = [C.c_int] * 4
    lib.oracle_phred.argtypes = [C.c_double]
    lib.oracle_mononucleotide.argtypes = [C.c_char_p]
    lib.oracle_run_ssvd.restype = C.c_void_p
    lib.oracle_run_ssvd.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.POINTER(C.c_int)),
                                    C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_double)]
    lib.oracle_free.argtypes = [C.c_void_p]
    lib.oracle_run_msvd.restype = C.c_void_p
    lib.oracle_run_msvd.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.POINTER(C.c_int)),
                                    C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_double)]
    lib.oracle_run_coverage_stats.restype = C.c_void_p
    lib.oracle_run_coverage_stats.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    lib.oracle_run_relative_allele_counts.restype = C.c_void_p
    lib.oracle_run_relative_allele_counts.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.oracle_run_relative_allele_counts_regions.restype = C.c_void_p
    lib.oracle_run_relative_allele_counts_regions.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]
    lib.oracle_english_fmt.argtypes = [C.c_double, C.c_char_p]
    lib.oracle_string_hashset_order.argtypes = [C.c_int, C.POINTER(C.c_char_p), ip]
    lib.oracle_diversity_info.argtypes = [C.c_int, C.c_int, ip, ip, C.c_char_p]
    return lib


def snv_column(lib, counts, qual_char="?", n_lowq=0, ref="A", alt="C", het_rate=0.001, max_qs=30, discovery=True):
    cnt = (C.c_int * 4)(*counts)
    gt = (C.c_int * 2)()
    pl = (C.c_int * 3)()
    gq, qual, dp = C.c_int(), C.c_int(), C.c_int()
    alt_out = C.create_string_buffer(2)
    called = lib.oracle_snv_column(cnt, ord(qual_char), n_lowq, ref.encode(), alt.encode(), het_rate, max_qs, 1 if discovery else 0,
                                   gt, pl, C.byref(gq), C.byref(qual), C.byref(dp), alt_out)
    return dict(called=bool(called), gt=(gt[0], gt[1]), pl=tuple(pl), gq=gq.value, qual=qual.value, dp=dp.value,
                alt=alt_out.raw[:1].decode())


def run_ssvd(lib, fasta, sam, args=(), want_counts=False):
    """Runs the oracle detector; returns (vcf_text, counts[n,6] or None, genotyped_positions, seconds)."""
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    cptr = C.POINTER(C.c_int)()
    n_counts, gp, sec = C.c_long(0), C.c_long(0), C.c_double(0)
    res = lib.oracle_run_ssvd(fasta.encode(), sam.encode(), len(args), argv, C.byref(cptr) if want_counts else None,
                              C.byref(n_counts), C.byref(gp), C.byref(sec))
    if not res:
        raise RuntimeError("oracle run failed")
    text = C.string_at(res).decode()
    lib.oracle_free(res)
    counts = None
    if want_counts:
        counts = np.ctypeslib.as_array(cptr, shape=(n_counts.value, 6)).copy() if n_counts.value else np.zeros((0, 6), np.int32)
        lib.oracle_free(cptr)
    return text, counts, gp.value, sec.value


def run_msvd(lib, fasta, sams, args=(), want_counts=False):
    """Runs the oracle MultisampleVariantsDetector on SAM files; returns (vcf_text, pooled counts[n,6] or None, genotyped_positions, seconds)."""
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    files = (C.c_char_p * len(sams))(*[s.encode() for s in sams])
    cptr = C.POINTER(C.c_int)()
    n_counts, gp, sec = C.c_long(0), C.c_long(0), C.c_double(0)
    res = lib.oracle_run_msvd(fasta.encode(), len(sams), files, len(args), argv, C.byref(cptr) if want_counts else None,
                              C.byref(n_counts), C.byref(gp), C.byref(sec))
    if not res:
        raise RuntimeError("oracle MSVD run failed")
    text = C.string_at(res).decode()
    lib.oracle_free(res)
    counts = None
    if want_counts:
        counts = np.ctypeslib.as_array(cptr, shape=(n_counts.value, 6)).copy() if n_counts.value else np.zeros((0, 6), np.int32)
        lib.oracle_free(cptr)
    return text, counts, gp.value, sec.value


def run_coverage_stats(lib, fasta, sam, min_mq=20):
    """The oracle's CoverageStatisticsCalculator on a SAM file -> (table text, seconds)."""
    sec = C.c_double(0)
    res = lib.oracle_run_coverage_stats(fasta.encode(), sam.encode(), min_mq, C.byref(sec))
    if not res:
        raise RuntimeError("oracle CoverageStats run failed")
    text = C.string_at(res).decode()
    lib.oracle_free(res)
    return text, sec.value


def run_relative_allele_counts(lib, fasta, sam, min_rd=10, max_rd=1000, min_bq=20, secondary=False, frs=None, srs=None):
    """The oracle's RelativeAlleleCountsCalculator on a SAM file -> (report text, seconds).  frs / srs: region files to filter / select."""
    sec = C.c_double(0)
    res = lib.oracle_run_relative_allele_counts_regions(fasta.encode(), sam.encode(), min_rd, max_rd, min_bq, int(secondary), frs.encode() if frs else None,
                                                        srs.encode() if srs else None, C.byref(sec))
    if not res:
        raise RuntimeError("oracle RelativeAlleleCounts run failed")
    text = C.string_at(res).decode()
    lib.oracle_free(res)
    return text, sec.value


def diversity_info(lib, n_alleles, called, copy_numbers):
    """called: per sample (g0, g1) or (-1, -1); copy_numbers: per sample a list of n_alleles ints -> 'NS=..;AN=..;AFS=..;OH=..[;MAF=..]'."""
    n = len(called)
    c = (C.c_int * (2 * n))(*[x for g in called for x in g])
    cn = (C.c_int * (n * n_alleles))(*[x for row in copy_numbers for x in row])
    buf = C.create_string_buffer(256)
    lib.oracle_diversity_info(n_alleles, n, c, cn, buf)
    return buf.value.decode()


def english_fmt(lib, v):
    buf = C.create_string_buffer(16)
    lib.oracle_english_fmt(v, buf)
    return buf.value.decode()


def parse_vcf(text):
    """-> list of dict(chrom,pos,ref,alt,qual,info,fmt,gt,pl,gq,dp,cnt,acn) for single-sample VCF text."""
    out = []
    for line in text.splitlines():
        if not line or line[0] == "#":
            continue
        t = line.split("\t")
        s = t[9].split(":")
        out.append(dict(chrom=t[0], pos=int(t[1]), ref=t[3], alt=t[4], qual=int(t[5]), info=t[7], fmt=t[8], gt=s[0], pl=s[1],
                        gq=int(s[2]), dp=int(s[3]), cnt=s[4], acn=s[5], line=line))
    return out


# ---- the oracle over a whole workload, one process per sequence / region (the reference's only parallel mode: -querySeq) ----
_PC = {}


def _per_contig_job(job):
    """Worker (forked: the batch is inherited, not pickled): SAM text of one region's records -> oracle -> (body lines, genotyped positions)."""
    import tempfile

    from ngsepcore_b200 import simulate
    ci, first, last, multi, args = job
    batch, fa, names = _PC["batch"], _PC["fa"], _PC["names"]
    lo, hi = np.searchsorted(batch.seq_idx, ci, "left"), np.searchsorted(batch.seq_idx, ci, "right")
    f = batch.first[lo:hi]
    keep = lo + np.nonzero((f >= first - 2000) & (f <= last))[0]     # lead-in: every record that can reach `first`
    with tempfile.TemporaryDirectory() as d:
        sam = os.path.join(d, "r.sam")
        simulate.write_sam(sam, batch.select(keep))
        q = tuple(args) + ("-querySeq", names[ci], "-first", str(first), "-last", str(last))
        lib = load_oracle()
        if multi:
            text, _, gp, sec = run_msvd(lib, fa, [sam], args=q)
        else:
            text, _, gp, sec = run_ssvd(lib, fa, sam, args=q)
    return [l for l in text.splitlines() if not l.startswith("#")], gp, sec


def run_regions(genome, batch, fa, regions=None, args=(), multi=False, workers=None, piece=400000):
    """The oracle over `regions` [(contig index, first, last)] (default: every sequence of the batch, cut into pieces of
    `piece` positions), one forked process per piece; returns (body lines in genome order, genotyped positions, cpu seconds).
    A query window is exactly how the reference itself is run in parallel (APG:310-322): a piece sees every record that can reach it."""
    import multiprocessing as mp
    names = list(batch.contig_names)
    if regions is None:
        regions = []
        for ci, length in enumerate(batch.contig_lengths):
            for f in range(1, length + 1, piece):
                regions.append((ci, f, min(length, f + piece - 1)))
    _PC.update(batch=batch, fa=fa, names=names)
    workers = workers or max(1, min(len(regions), os.cpu_count() or 1))
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_per_contig_job, [(ci, f, l, multi, tuple(args)) for ci, f, l in regions], chunksize=1)
    _PC.clear()
    body, gp, sec = [], 0, 0.0
    for lines, g, s in res:
        body += lines; gp += g; sec += s
    return body, gp, sec
