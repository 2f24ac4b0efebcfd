"""ngsep_indel_tallies (the indel / STR allele tallies on the device): the per-call part of
CountsHelper.updateCountsIndel (discovery/CountsHelper.java:253-304, 384-396) for a batch of span > 1 events against a plain Python
restatement of the same lines, bit for bit — counts, strand counts, the diagonal of logConditionalProbs, the allele every call
counts for and its fit.  (What the detector does with them — mixture terms, posteriors, records — is held to the oracle by the
whole-detector tests: every indel / STR record of test_gpu_detector_parity.py and test_gpu_baseline_configs.py goes through this
entry point.)"""
import math

import numpy as np
import pytest

from ngsepcore_b200 import capi

pytestmark = pytest.mark.gpu
LERR = math.log10(0.0001)


def _tables():
    tab = capi.host_tables(0.001)          # the very doubles the context hands to the device: [3 q] a match, [3 q + 1] a mismatch
    return {q: float(tab[3 * q]) for q in range(3, 31)}, {q: float(tab[3 * q + 1]) for q in range(3, 31)}


def _restated(alleles, calls, max_qs=30):
    match, err = _tables()
    n = len(alleles)
    counts, strand, diag, idx, fit = [0] * n, [[0, 0] for _ in range(n)], [0.0] * n, [], []
    for a, quals, neg in calls:
        index = alleles.index(a) if a in alleles else -1                      # CH:262
        lc, best = [LERR] * n, -1
        for i, al in enumerate(alleles):
            if len(al) != len(a):
                continue
            v = 0.0
            for k in range(len(al)):                                          # calculateLogCond, CH:384-396
                q = min(max_qs, ord(quals[k]) - 33)
                q = q - 256 if q > 127 else q
                if q < 3:
                    continue
                v += match[q] if al[k] == a[k] else err[q]
            lc[i] = max(LERR, v)
            if lc[i] > LERR and (best == -1 or lc[best] < lc[i]):
                best = i
        if index >= 0 and best >= 0 and best != index:
            index = min(index, best)
        elif index < 0 and best >= 0:
            index = best
        if index >= 0:
            counts[index] += 1
            strand[index][0 if neg else 1] += 1
        for i in range(n):
            diag[i] += lc[i]
        idx.append(index)
        fit.append(lc[index] if index >= 0 else LERR)
    return counts, strand, diag, idx, fit


def _random_events(rng, n_events):
    events = []
    for _ in range(n_events):
        span = int(rng.integers(2, 30))
        ref = "".join(rng.choice(list("ACGT"), span))
        alleles = [ref]
        for _ in range(int(rng.integers(0, 5))):
            kind = rng.random()
            if kind < 0.4 and span > 2:
                cut = int(rng.integers(1, span - 1)); alleles.append(ref[:1] + ref[1 + cut:])          # a deletion
            elif kind < 0.8:
                alleles.append(ref[:1] + "".join(rng.choice(list("ACGT"), int(rng.integers(1, 9)))) + ref[1:])   # an insertion
            else:
                k = int(rng.integers(0, span)); alleles.append(ref[:k] + "ACGT"[("ACGT".index(ref[k]) + 1) % 4] + ref[k + 1:])   # same length, one mismatch
        alleles = list(dict.fromkeys(alleles))
        calls = []
        for _ in range(int(rng.integers(0, 60))):
            a = alleles[int(rng.integers(0, len(alleles)))] if rng.random() < 0.8 else "".join(rng.choice(list("ACGTN"), int(rng.integers(1, span + 6))))
            if rng.random() < 0.3 and len(a) > 1:                      # sequencing errors: the call fits an allele without matching it
                k = int(rng.integers(0, len(a))); a = a[:k] + "ACGT"[int(rng.integers(0, 4))] + a[k + 1:]
            quals = "".join(chr(33 + int(q)) for q in rng.choice([0, 2, 3, 4, 10, 20, 30, 40, 41, 60], len(a)))
            calls.append((a, quals, bool(rng.random() < 0.5)))
        events.append((alleles, calls))
    events.append((["ACGT"], []))                                        # an event without calls
    events.append((["AC", "A"], [("", "", False), ("AC", "II", True)]))  # an empty call fits nothing
    return events


@pytest.mark.parametrize("max_qs", [30, 25])
def test_tallies_match_the_restated_lines(max_qs):
    events = _random_events(np.random.default_rng(7 + max_qs), 400)
    with capi.Context(max_base_qs=max_qs) as ctx:
        got = ctx.indel_tallies(events, LERR)
    assert len(got) == len(events)
    n_fits = 0
    for (alleles, calls), g in zip(events, got):
        counts, strand, diag, idx, fit = _restated(alleles, calls, max_qs)
        assert g["total"] == len(calls)
        assert g["counts"].tolist() == counts and g["strand"].tolist() == strand
        assert g["call_index"].tolist() == idx
        assert g["diag"].tolist() == diag and g["call_fit"].tolist() == fit        # bit for bit: sums of the same table values in the same order
        n_fits += sum(1 for (a, _, _), i in zip(calls, idx) if i >= 0 and alleles[i] != a)
    assert n_fits > 100                                                            # calls that count for their best-fitting allele


def test_bad_layouts_are_rejected():
    with capi.Context() as ctx:
        assert ctx.indel_tallies([], LERR) == []
        e = capi.IndelEvents(1, 1, 0, 0, None, None, None, None, None, None, None, None, LERR)
        assert ctx.lib.ngsep_indel_tallies(ctx.h, e, None, None, None, None, None, None) == -1
