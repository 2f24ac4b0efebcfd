"""Rendering of device SNV calls (ngsep_snv_call) into the reference's VCF record text.

Follows SingleSampleVariantsDetector.saveSequenceVariants + VCFFileWriter.printVCFRecord
(reference: src/ngsep/discovery/SingleSampleVariantsDetector.java:933-968, src/ngsep/vcf/VCFFileWriter.java:44-308):
CalledSNV stores its four log-conditionals as float (variants/CalledSNV.java:259-266), so the bi-allelic PL is
rounded from the float cast; the tri-allelic record keeps doubles (VD:170-171).
"""
from __future__ import annotations

import math

import numpy as np

from .capi import CALL_HET, CALL_HOMALT, CALL_TRIALLELIC

_B = "ACGT"


def java_round(x: float) -> int:
    f = math.floor(x)
    return int(f + 1) if x - f >= 0.5 else int(f)


def snv_record_fields(call, ploidy: int = 2, strand_bias: int | None = None, embedded: bool = False) -> dict:
    """One device call -> dict(pos, ref, alt, qual, info, fmt, sample) matching the reference's text fields."""
    ref = int(call["ref"])
    L = np.asarray(call["logcond"], np.float64).reshape(4, 4)
    kind = int(call["kind"])
    cnt = ",".join(str(int(c)) for c in call["count"])
    if kind == CALL_TRIALLELIC:
        alleles = [ref, int(call["a0"]), int(call["a1"])]
        pl = []
        for j in range(3):
            for i in range(j + 1):
                pl.append(java_round(-10.0 * float(L[alleles[i], alleles[j]])))
        gt, acn, info = "1/2", "0,1,1", "TYPE=MULTISNV"
    else:
        alt = int(call["a1"])
        alleles = [ref, alt]
        f = lambda v: float(np.float32(v))
        pl = [java_round(-10.0 * f(L[ref, ref])), java_round(-10.0 * f(L[ref, alt])), java_round(-10.0 * f(L[alt, alt]))]
        if kind == CALL_HET:
            gt, acn = "0/1", "1,1"
        elif kind == CALL_HOMALT:
            gt, acn = "1/1", "0,2"
        else:
            raise ValueError(f"unknown call kind {kind}")
        info = "TYPE=EMBEDDED" if embedded else ""
    infos = []
    if strand_bias is not None and kind != CALL_TRIALLELIC:
        infos.append(f"FS={strand_bias}")
    if info:
        infos.append(info)
    if ploidy == 1:
        gt = gt.split("/")[0] if gt[0] == gt[-1] else gt
    sample = f"{gt}:{','.join(map(str, pl))}:{int(call['gq'])}:{int(call['depth'])}:{cnt}:{acn}"
    return dict(pos=int(call["pos"]), ref=_B[ref], alt=",".join(_B[a] for a in alleles[1:]), qual=int(call["qual"]),
                info=";".join(infos) if infos else ".", fmt="GT:PL:GQ:DP:BSDP:ACN", sample=sample)


def snv_record_line(call, chrom: str, **kw) -> str:
    f = snv_record_fields(call, **kw)
    return f"{chrom}\t{f['pos']}\t.\t{f['ref']}\t{f['alt']}\t{f['qual']}\t.\t{f['info']}\t{f['fmt']}\t{f['sample']}"
