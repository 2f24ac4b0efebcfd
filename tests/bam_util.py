"""Minimal SAM -> BAM converter (BGZF blocks via zlib) for the tests of the CLI's native BAM reader.  Follows the SAM
specification (sections 4.1 BGZF, 4.2 BAM); test infrastructure, not part of the product."""
import re
import struct
import zlib

_SEQ = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


def _bgzf_block(data: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    cdata = comp.compress(data) + comp.flush()
    bsize = len(cdata) + 25
    return (b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + cdata +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def sam_to_bam(sam_path: str, bam_path: str, block: int = 60000) -> int:
    header, refs, recs = [], [], []
    for line in open(sam_path):
        line = line.rstrip("\n")
        if not line:
            continue
        if line[0] == "@":
            header.append(line)
            if line.startswith("@SQ"):
                d = dict(t.split(":", 1) for t in line.split("\t")[1:])
                refs.append((d["SN"], int(d["LN"])))
            continue
        recs.append(line.split("\t"))
    ref_id = {n: i for i, (n, _) in enumerate(refs)}
    text = ("\n".join(header) + "\n").encode()
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for n, l in refs:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    for f in recs:
        qname, flag, rname, pos, mapq, cigar, seq, qual = f[0], int(f[1]), f[2], int(f[3]), int(f[4]), f[5], f[9], f[10]
        ops = [] if cigar == "*" else [(int(n), _OPS[o]) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
        l_seq = 0 if seq == "*" else len(seq)
        packed = bytearray((l_seq + 1) // 2)
        for i in range(l_seq):
            packed[i >> 1] |= _SEQ.get(seq[i].upper(), 15) << (0 if i & 1 else 4)     # BAM has no lower case: upper-cased like samtools
        q = b"\xff" * l_seq if qual == "*" else bytes(min(93, ord(c) - 33) for c in qual[:l_seq]).ljust(l_seq, b"\x05")   # short QUAL cannot be stored: pad as RA:581-595 would read it ('&' = q5)
        tags = bytearray()
        for t in f[11:]:
            tag, ty, val = t.split(":", 2)
            if ty == "Z":
                tags += tag.encode() + b"Z" + val.encode() + b"\x00"
            elif ty == "i":
                v = int(val)
                tags += tag.encode() + (b"C" + struct.pack("<B", v) if 0 <= v < 256 else b"i" + struct.pack("<i", v))
        body = (struct.pack("<iiBBHHHIiii", ref_id.get(rname, -1), pos - 1, len(qname) + 1, min(mapq, 255), 4680, len(ops), flag, l_seq, -1, -1, 0) +
                qname.encode() + b"\x00" + b"".join(struct.pack("<I", (n << 4) | o) for n, o in ops) + bytes(packed) + q + bytes(tags))
        out += struct.pack("<i", len(body)) + body
    with open(bam_path, "wb") as o:
        for i in range(0, len(out), block):
            o.write(_bgzf_block(bytes(out[i:i + block])))
        o.write(_bgzf_block(b""))     # the EOF marker block
    return len(recs)
