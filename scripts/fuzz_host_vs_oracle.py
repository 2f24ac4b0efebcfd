"""CPU fuzz of the host pipeline against the oracle (no GPU): random simulation parameters and option sets; compares the
host-made indel / STR records (host_only mode) with the oracle's.  Round 1: 85 single-sample, 85 multi-sample and 30 CLI seeds, no
mismatch.   python scripts/fuzz_host_vs_oracle.py single|multi|cli|known|bam FIRST_SEED LAST_SEED
(bam: corrupted BAM files through the CLI's reader, one thread and many small parts on six threads: a clean error or a decode, never a crash.)"""
import sys
MODE = sys.argv.pop(1)
import sys, os, tempfile, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngsepcore_b200 import simulate, hostapi
from tests import oracle_util as ou


def single():
    lib=ou.load_oracle()
    d=tempfile.mkdtemp()
    bad=0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng=np.random.default_rng(seed)
        kw=dict(depth=float(rng.choice([6,15,30,80])), indel_rate=float(rng.choice([5e-4,2e-3,5e-3])), snv_rate=3e-3, indel_jitter=float(rng.choice([0.0,0.3,0.8])),
                ungapped_end_bp=int(rng.choice([3,6,12])), n_read_groups=int(rng.integers(1,4)), paired=bool(rng.integers(0,2)))
        opts=[(),("-minQuality","10"),("-h","0.0001","-maxBaseQS","25"),("-ignore5","5","-ignore3","7"),("-maxAlnsPerStartPos","1"),("-embeddedSNVs",),("-ploidy","1")][int(rng.integers(0,7))]
        detkw={(): {}, ("-minQuality","10"): dict(minQuality=10), ("-h","0.0001","-maxBaseQS","25"): dict(h=0.0001,maxBaseQS=25), ("-ignore5","5","-ignore3","7"): dict(ignore5=5,ignore3=7),
               ("-maxAlnsPerStartPos","1"): dict(maxAlnsPerStartPos=1), ("-embeddedSNVs",): dict(embeddedSNVs=True), ("-ploidy","1"): dict(ploidy=1)}[opts]
        genome,batch,_=simulate.simulate(contigs=[("chrA",60000),("chrB",20000)], seeds=(seed*5+1,seed*5+2,seed*5+3), **kw)
        fa,sam=os.path.join(d,"r.fa"),os.path.join(d,"r.sam")
        simulate.write_fasta(fa,genome); simulate.write_sam(sam,batch)
        want,_,_,_=ou.run_ssvd(lib,fa,sam,args=opts)
        wl=[l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
        det=hostapi.Detector(host_only=True, **detkw); det.set_genome(genome)
        gl=det.run(batch, with_header=False).splitlines()
        ok = gl==wl
        if not ok:
            bad+=1
            for a,b in zip(gl,wl):
                if a!=b: print("DIFF\n got ",a[:300],"\n want",b[:300]); break
            print("len", len(gl), len(wl))
        print(seed, kw["depth"], kw["indel_rate"], opts, len(wl), "OK" if ok else "MISMATCH", flush=True)
    print("bad", bad)


def multi():
    lib=ou.load_oracle()
    d=tempfile.mkdtemp()
    bad=0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng=np.random.default_rng(seed)
        kw=dict(n_samples=int(rng.integers(2,14)), depth=float(rng.choice([4,8,12,20])), mode=str(rng.choice(["wgs","gbs"])), n_tags=120,
                indel_rate=float(rng.choice([1e-3,3e-3])), snv_rate=3e-3, rgs_per_sample=int(rng.integers(1,3)), no_rg_frac=float(rng.choice([0,0.05])),
                indel_jitter=float(rng.choice([0.3,0.7])))
        opts=[(),("-minQuality","20"),("-h","0.01","-maxBaseQS","20"),("-ignore5","4","-ignore3","4")][int(rng.integers(0,4))]
        detkw={(): {}, ("-minQuality","20"): dict(minQuality=20), ("-h","0.01","-maxBaseQS","20"): dict(h=0.01,maxBaseQS=20), ("-ignore5","4","-ignore3","4"): dict(ignore5=4,ignore3=4)}[opts]
        cap = 100 if kw["mode"]=="gbs" else 5
        genome,batch,_=simulate.simulate_population([("chrA",40000),("chrB",15000)], seeds=(seed*7+1,seed*7+2,seed*7+3), **kw)
        fa,sam=os.path.join(d,"r.fa"),os.path.join(d,"r.sam")
        simulate.write_fasta(fa,genome); simulate.write_sam(sam,batch)
        want,_,_,_=ou.run_msvd(lib,fa,[sam],args=opts+("-maxAlnsPerStartPos",str(cap)))
        wl=[l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
        det=hostapi.MultisampleDetector(list(zip(batch.rg_names,batch.rg_samples)), host_only=True, maxAlnsPerStartPos=cap, **detkw)
        det.set_genome(genome)
        gl=[l for l in det.run(batch).splitlines() if not l.startswith("#")]
        ok = gl==wl
        if not ok:
            bad+=1
            for a,b in zip(gl,wl):
                if a!=b: print("DIFF\n got ",a[:300],"\n want",b[:300]); break
            print("len", len(gl), len(wl))
        print(seed, kw["mode"], kw["n_samples"], kw["depth"], opts, len(wl), "OK" if ok else "MISMATCH", flush=True)
    print("bad", bad)


def cli():
    """Mutated-record SAMs (tests/test_gpu_sam_torture.py::_mutate) through the CLI's SAM reader in --host-only mode."""
    import subprocess
    from tests.test_gpu_sam_torture import _mutate
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "ngsepcore_b200", "ngsep_b200")
    lib = ou.load_oracle()
    d = tempfile.mkdtemp()
    bad = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng = np.random.default_rng(seed)
        genome, batch, _ = simulate.simulate(contigs=[("chrA", 50000), ("chrB", 20000)], depth=float(rng.choice([10, 25, 50])), seeds=(seed * 3 + 1, seed * 3 + 2, seed * 3 + 3),
                                             indel_rate=float(rng.choice([1e-3, 3e-3])), n_read_groups=int(rng.integers(1, 4)), lower_frac=0.02)
        fa, sam0, sam = os.path.join(d, "r.fa"), os.path.join(d, "p.sam"), os.path.join(d, "m.sam")
        simulate.write_fasta(fa, genome); simulate.write_sam(sam0, batch)
        open(sam, "w").writelines(_mutate(open(sam0).readlines(), np.random.default_rng(seed + 1000)))
        opts = [(), ("-s", "-p"), ("-minMQ", "0", "-maxAlnsPerStartPos", "2"), ("-p", "-ignore5", "3", "-ignore3", "3"), ("-s", "-minMQ", "30")][int(rng.integers(0, 5))]
        want, _, _, _ = ou.run_ssvd(lib, fa, sam, args=opts)
        wl = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
        subprocess.check_call([exe, "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", os.path.join(d, "o"), "--host-only"] + list(opts), stderr=subprocess.DEVNULL)
        gl = [l for l in open(os.path.join(d, "o.vcf")).read().splitlines() if not l.startswith("#")]
        ok = gl == wl
        bad += not ok
        print(seed, opts, len(wl), "OK" if ok else "MISMATCH", flush=True)
    print("bad", bad)


def bam():
    """Random corruption of a BAM stream (bytes flipped anywhere in the first records / the header, truncation at a random place) through the CLI's reader
    in --host-only mode, decoded by one thread and as many small parts on six threads: exit code 0 (decoded) or 2 (the reader's error), and the same one both ways."""
    import struct, subprocess, zlib
    from tests import bam_util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "ngsepcore_b200", "ngsep_b200")
    d = tempfile.mkdtemp()
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 12000)], depth=12, seeds=(171, 172, 173), indel_rate=2e-3, n_read_groups=2)
    fa, sam, bamf = os.path.join(d, "r.fa"), os.path.join(d, "p.sam"), os.path.join(d, "p.bam")
    simulate.write_fasta(fa, genome); simulate.write_sam(sam, batch)
    bam_util.sam_to_bam(sam, bamf)
    raw = open(bamf, "rb").read()
    stream, p = bytearray(), 0
    while p < len(raw):
        bsize = struct.unpack_from("<H", raw, p + 16)[0] + 1
        stream += zlib.decompress(raw[p + 18:p + bsize - 8], -15)
        p += bsize
    bad = 0
    modes = (dict(NGSEP_HOST_THREADS="1"), dict(NGSEP_HOST_THREADS="6", NGSEP_BAM_PART_RECORDS="40"))
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng = np.random.default_rng(seed)
        s = bytearray(stream)
        kind = int(rng.integers(0, 3))
        if kind == 0:       # bytes flipped in the header and the first records
            for _ in range(int(rng.integers(1, 6))): s[int(rng.integers(0, min(len(s), 6000)))] = int(rng.integers(0, 256))
        elif kind == 1:     # bytes flipped anywhere
            for _ in range(int(rng.integers(1, 12))): s[int(rng.integers(0, len(s)))] = int(rng.integers(0, 256))
        else:               # truncation
            s = s[:int(rng.integers(1, len(s)))]
        path = os.path.join(d, "c.bam")
        with open(path, "wb") as o:
            for i in range(0, len(s), 60000): o.write(bam_util._bgzf_block(bytes(s[i:i + 60000])))
            o.write(bam_util._bgzf_block(b""))
        rcs, outs = [], []
        for env in modes:
            rc = subprocess.call([exe, "SingleSampleVariantsDetector", "-i", path, "-r", fa, "-o", os.path.join(d, "o"), "--host-only"], stderr=subprocess.DEVNULL,
                                 env=dict(os.environ, **env), timeout=120)
            rcs.append(rc); outs.append(open(os.path.join(d, "o.vcf")).read() if rc == 0 else None)
        ok = rcs[0] in (0, 2) and rcs[0] == rcs[1] and outs[0] == outs[1]
        bad += not ok
        print(seed, ("head", "any", "cut")[kind], rcs, "OK" if ok else "MISMATCH", flush=True)
    print("bad", bad)


def known():
    """-knownVariants: the host-made records (the input variants that are not SNVs) against the oracle's."""
    from tests.known_util import make_known_vcf
    lib = ou.load_oracle()
    d = tempfile.mkdtemp()
    bad = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng = np.random.default_rng(seed)
        kw = dict(depth=float(rng.choice([2, 6, 15, 40])), indel_rate=float(rng.choice([5e-4, 2e-3, 5e-3])), snv_rate=3e-3, indel_jitter=float(rng.choice([0.0, 0.3, 0.8])),
                  n_read_groups=int(rng.integers(1, 3)), paired=bool(rng.integers(0, 2)))
        opts, detkw = [((), {}), (("-minQuality", "10"), dict(minQuality=10)), (("-maxBaseQS", "25", "-h", "0.0001"), dict(maxBaseQS=25, h=0.0001)),
                       (("-ignore5", "4", "-ignore3", "6"), dict(ignore5=4, ignore3=6)), (("-ploidy", "1"), dict(ploidy=1))][int(rng.integers(0, 5))]
        contigs = [("chrA", 50000), ("chrB", 20000)]
        genome, batch, _ = simulate.simulate(contigs=contigs, seeds=(seed * 5 + 1, seed * 5 + 2, seed * 5 + 3), **kw)
        fa, sam, kv = os.path.join(d, "r.fa"), os.path.join(d, "r.sam"), os.path.join(d, "k.vcf")
        simulate.write_fasta(fa, genome); simulate.write_sam(sam, batch)
        disc, _, _, _ = ou.run_ssvd(lib, fa, sam, args=("-minQuality", "0"))
        make_known_vcf(kv, disc, genome, contigs, rng)
        want, _, _, _ = ou.run_ssvd(lib, fa, sam, args=("-knownVariants", kv) + opts)
        def snv(l):
            t = l.split("\t")
            return all(len(a) == 1 and a in "ACGT" for a in [t[3]] + t[4].split(","))
        wl = [l for l in want.splitlines() if not l.startswith("#") and not snv(l)]
        det = hostapi.Detector(host_only=True, knownVariants=kv, **detkw); det.set_genome(genome)
        gl = det.run(batch, with_header=False).splitlines()
        ok = gl == wl
        print(seed, kw["depth"], opts, len(wl), "OK" if ok else "MISMATCH", flush=True)
        bad += not ok
    print("mismatches", bad)


{"single": single, "multi": multi, "cli": cli, "known": known, "bam": bam}[MODE]()
