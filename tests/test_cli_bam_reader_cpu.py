"""The CLI's native BAM (BGZF) reader, on CPU: a SAM file and its BAM conversion must decode to the SAME record batch
(`--dump-batch`, every array compared) and give the same host-made indel / STR records (`--host-only`: admission + realigner
pre-pass + indel caller, no device).  Mutated records (clips, =/X/N/P, secondary / NH / missing QUAL / no read group /
duplicates) are included; lower-case bases and short QUAL strings are left out of the mutation because BAM cannot hold them."""
import os
import subprocess

import numpy as np

from ngsepcore_b200 import simulate
from tests import bam_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsepcore_b200", "ngsep_b200")


def _read_dump(path):
    raw = open(path, "rb").read()
    out, p = [], 0
    while p < len(raw):
        n = int(np.frombuffer(raw, np.uint64, 1, p)[0])
        out.append(raw[p + 8:p + 8 + n])
        p += 8 + n
    return out


def test_bam_and_sam_decode_to_the_same_batch(tmp_path):
    from tests.test_gpu_sam_torture import _mutate
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 60000), ("chrB", 25000)], depth=20, seeds=(151, 152, 153), indel_rate=2e-3, n_read_groups=2)
    fa, sam0, sam, bam = (str(tmp_path / n) for n in ("r.fa", "plain.sam", "m.sam", "m.bam"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    lines = _mutate(open(sam0).readlines(), np.random.default_rng(3))
    keep = []
    for l in lines:                      # what BAM cannot represent: lower-case bases, QUAL shorter than SEQ
        if not l.startswith("@"):
            f = l.split("\t")
            if f[9] != f[9].upper() or (f[10] != "*" and len(f[10]) != len(f[9])):
                continue
        keep.append(l)
    open(sam, "w").writelines(keep)
    n = bam_util.sam_to_bam(sam, bam)
    assert n > 15000
    outs = {}
    # "sam" / "bam": one thread reads the whole file; "..._parts": ranges of about 1500 records on six threads, laid side by side afterwards
    parts = {"NGSEP_HOST_THREADS": "6", "NGSEP_BAM_PART_RECORDS": "1500"}
    for tag, path, env in (("sam", sam, {"NGSEP_HOST_THREADS": "1"}), ("sam_parts", sam, parts), ("bam", bam, {"NGSEP_HOST_THREADS": "1"}), ("bam_parts", bam, parts)):
        prefix, dump = str(tmp_path / f"out_{tag}"), str(tmp_path / f"dump_{tag}.bin")
        subprocess.check_call([CLI, "SingleSampleVariantsDetector", "-i", path, "-r", fa, "-o", prefix, "--host-only", "--dump-batch", dump],
                              stderr=subprocess.DEVNULL, env=dict(os.environ, **env))
        outs[tag] = (_read_dump(dump), open(prefix + ".vcf").read())
    names = ["seq_idx", "first", "flags", "mapq", "rg", "nh", "cigar_off", "cigar", "base_off", "bases", "quals", "name_hash"]
    assert len(outs["sam"][0]) == len(names)
    for tag in ("sam_parts", "bam", "bam_parts"):
        for name, a, b in zip(names, outs["sam"][0], outs[tag][0]):
            assert a == b, (tag, name)
        assert outs["sam"][1] == outs[tag][1]
    assert len(outs["sam"][0][1]) // 4 == n
    body = [l for l in outs["sam"][1].splitlines() if not l.startswith("#")]
    assert len(body) > 20 and outs["sam"][1] == outs["bam"][1]


def test_fasta_reader_line_forms(tmp_path):
    """The CLI's FASTA reader (one read of the file, lines found with memchr): descriptions after the name, CR LF line ends, another
    line width, text in front of the first header and no newline at the end give the same genome — the same host-made records."""
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 30000), ("chrB", 9000)], depth=15, seeds=(181, 182, 183), indel_rate=3e-3, lower_frac=0.1)
    fa, fa2, sam = (str(tmp_path / n) for n in ("r.fa", "r2.fa", "p.sam"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    with open(fa2, "wb") as o:
        o.write(b"stray text\r\n")
        for k, (name, seq) in enumerate(genome.items()):
            o.write(b">" + name.encode() + (b" a description\twith a tab" if k == 0 else b"") + b"\r\n")
            raw = bytes(seq)
            for i in range(0, len(raw), 53):
                last = k == len(genome) - 1 and i + 53 >= len(raw)
                o.write(raw[i:i + 53] + (b"" if last else b"\r\n"))
    outs = []
    for f in (fa, fa2):
        prefix = str(tmp_path / ("o_" + os.path.basename(f)))
        subprocess.check_call([CLI, "SingleSampleVariantsDetector", "-i", sam, "-r", f, "-o", prefix, "--host-only", "-ignoreLowerCaseRef"], stderr=subprocess.DEVNULL)
        outs.append(open(prefix + ".vcf").read())
    assert outs[0] == outs[1] and sum(1 for l in outs[0].splitlines() if not l.startswith("#")) > 20


def test_bam_reader_rejects_garbage(tmp_path):
    fa = str(tmp_path / "r.fa")
    simulate.write_fasta(fa, {"c": np.frombuffer(b"ACGT" * 50, np.uint8)})
    bad = str(tmp_path / "bad.bam")
    open(bad, "wb").write(b"\x1f\x8b\x08\x04garbage-not-bgzf")
    assert subprocess.call([CLI, "SingleSampleVariantsDetector", "-i", bad, "-r", fa, "-o", str(tmp_path / "o"), "--host-only"], stderr=subprocess.DEVNULL) == 2


def test_multisample_cli_with_bam_inputs(oracle, tmp_path):
    """`ngsep_b200 MultisampleVariantsDetector --host-only` over three BAM inputs (two samples each): header reading, the k-way
    merge of APG:212-289 and the population indel caller — against the oracle reading the three SAM files."""
    from tests import oracle_util as ou
    genome, batch, _ = simulate.simulate_population([("chrA", 45000), ("chrB", 15000)], n_samples=6, depth=10, mode="wgs", indel_rate=2.5e-3,
                                                    seeds=(311, 312, 313))
    fa = str(tmp_path / "r.fa")
    simulate.write_fasta(fa, genome)
    sams, bams = [], []
    for k in range(3):
        part = batch.select(np.nonzero((batch.rg // 2) == k)[0])
        sam, bam = str(tmp_path / f"p{k}.sam"), str(tmp_path / f"p{k}.bam")
        simulate.write_sam(sam, part)
        bam_util.sam_to_bam(sam, bam)
        sams.append(sam)
        bams.append(bam)
    want, _, _, _ = ou.run_msvd(oracle, fa, sams, args=("-minQuality", "30"))
    want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    out = str(tmp_path / "ms.vcf")
    subprocess.check_call([CLI, "MultisampleVariantsDetector", "-r", fa, "-o", out, "-minQuality", "30", "--host-only"] + bams, stderr=subprocess.DEVNULL)
    got = open(out).read()
    assert len(want_lines) > 15
    assert [l for l in got.splitlines() if not l.startswith("#")] == want_lines
    assert [l for l in got.splitlines() if l.startswith("#")] == [l for l in want.splitlines() if l.startswith("#")]


def test_cli_host_only_records_match_oracle_on_mutated_sam(oracle, tmp_path):
    """The SAM reader + admission + realigner pre-pass + indel caller of the CLI on the mutated-record SAM (clips, =/X/N/P,
    secondary / NH / MAPQ / unmapped / duplicate records, missing or short QUAL, lower case, no read group), without a GPU:
    its indel / STR records are the oracle's, under two option sets."""
    from tests import oracle_util as ou
    from tests.test_gpu_sam_torture import _mutate
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 70000), ("chrB", 30000)], depth=25, seeds=(161, 162, 163), indel_rate=2e-3, n_read_groups=2,
                                         lower_frac=0.02)
    fa, sam0, sam = (str(tmp_path / n) for n in ("r.fa", "plain.sam", "m.sam"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    open(sam, "w").writelines(_mutate(open(sam0).readlines(), np.random.default_rng(5)))
    for opts in ((), ("-s", "-p", "-minMQ", "5", "-maxAlnsPerStartPos", "3", "-ignore5", "2")):
        want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=opts)
        want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
        prefix = str(tmp_path / "o")
        subprocess.check_call([CLI, "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", prefix, "--host-only"] + list(opts), stderr=subprocess.DEVNULL)
        got_lines = [l for l in open(prefix + ".vcf").read().splitlines() if not l.startswith("#")]
        assert len(want_lines) > 40
        assert got_lines == want_lines


def test_bam_reader_survives_corrupted_lengths(tmp_path):
    """every length field of a BAM record (block size, l_read_name, n_cigar_op, l_seq, B-array counts, unterminated Z
    strings) and of a BGZF block (XLEN, BSIZE) is checked against the bytes that are really there.  Corrupted files must end with
    the CLI's clean error (exit code 2) or decode — never crash (a negative return code = a signal) or allocate gigabytes."""
    import struct
    import zlib
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 8000)], depth=10, seeds=(171, 172, 173), indel_rate=2e-3, n_read_groups=2)
    fa, sam, bam = (str(tmp_path / n) for n in ("r.fa", "p.sam", "p.bam"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    # a tag of every shape on the first records
    lines = open(sam).readlines()
    k = next(i for i, l in enumerate(lines) if not l.startswith("@"))
    lines[k] = lines[k].rstrip("\n") + "\tNH:i:1\tXZ:Z:some text\n"
    open(sam, "w").writelines(lines)
    bam_util.sam_to_bam(sam, bam)
    raw = open(bam, "rb").read()
    # the decompressed stream
    stream, p = bytearray(), 0
    while p < len(raw):
        bsize = struct.unpack_from("<H", raw, p + 16)[0] + 1
        stream += zlib.decompress(raw[p + 18:p + bsize - 8], -15)
        p += bsize
    l_text = struct.unpack_from("<i", stream, 4)[0]
    first_rec = 8 + l_text + 4 + 4 + len("chrA") + 1 + 4
    rng = np.random.default_rng(9)

    def run(data: bytes, name: str) -> int:
        path = str(tmp_path / name)
        with open(path, "wb") as o:
            for i in range(0, len(data), 60000):
                o.write(bam_util._bgzf_block(bytes(data[i:i + 60000])))
            o.write(bam_util._bgzf_block(b""))
        return subprocess.call([CLI, "SingleSampleVariantsDetector", "-i", path, "-r", fa, "-o", str(tmp_path / "o"), "--host-only"],
                               stderr=subprocess.DEVNULL, timeout=60)

    assert run(bytes(stream), "clean.bam") == 0
    targeted = {"block_size": (first_rec, 4), "l_read_name": (first_rec + 4 + 8, 1), "n_cigar_op": (first_rec + 4 + 12, 2), "l_seq": (first_rec + 4 + 16, 4),
                "l_text": (4, 4), "n_ref": (8 + l_text, 4), "l_name": (8 + l_text + 4, 4)}
    for name, (off, width) in targeted.items():
        for val in (0xFFFFFFFF, 0x7FFFFFFF, 0x00FFFFFF, 1, 0):
            s = bytearray(stream)
            s[off:off + width] = struct.pack("<I", val)[:width]
            rc = run(bytes(s), f"{name}_{val:x}.bam")
            assert rc in (0, 2), (name, hex(val), rc)
    for it in range(40):       # random corruption of the first records (fixed part, CIGAR, tags) and truncation
        s = bytearray(stream)
        for _ in range(int(rng.integers(1, 6))):
            s[int(rng.integers(first_rec, min(len(s), first_rec + 4000)))] = int(rng.integers(0, 256))
        if it % 4 == 0:
            s = s[:int(rng.integers(first_rec, len(s)))]
        rc = run(bytes(s), f"rand_{it}.bam")
        assert rc in (0, 2), (it, rc)
    # BGZF level: XLEN / BSIZE that point outside the file, a block cut in the middle
    for off, val in ((10, 0xFFFF), (16, 0xFFFF), (16, 3), (10, 0)):
        b = bytearray(raw)
        struct.pack_into("<H", b, off, val)
        path = str(tmp_path / f"bgzf_{off}_{val}.bam")
        open(path, "wb").write(bytes(b))
        rc = subprocess.call([CLI, "SingleSampleVariantsDetector", "-i", path, "-r", fa, "-o", str(tmp_path / "o"), "--host-only"], stderr=subprocess.DEVNULL, timeout=60)
        assert rc in (0, 2), (off, val, rc)
    open(str(tmp_path / "cut.bam"), "wb").write(raw[:len(raw) // 2])
    assert subprocess.call([CLI, "SingleSampleVariantsDetector", "-i", str(tmp_path / "cut.bam"), "-r", fa, "-o", str(tmp_path / "o"), "--host-only"],
                           stderr=subprocess.DEVNULL, timeout=60) in (0, 2)
