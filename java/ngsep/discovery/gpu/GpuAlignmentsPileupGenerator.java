/*
 * GpuAlignmentsPileupGenerator — drop-in for ngsep.discovery.AlignmentsPileupGenerator (src/ngsep/discovery/
 * AlignmentsPileupGenerator.java:38-515) that moves the per-position SNV work of the listener chain to libngsep_pileup.so.
 * Source for the reference side of the binding (INTEGRATION.md): it needs the reference's classes and a JDK to compile,
 * neither of which exists in this repository's build image; the same algorithm runs and is tested here as C++
 * (ngsepcore_b200/csrc/host/detector.cpp) against the oracle.
 *
 * Public surface = the superclass's (addListener, setGenome / setSequencesMetadata, setQuerySeq/First/Last,
 * setMaxAlnsPerStartPos, setProcessNonUniquePrimaryAlignments, setProcessSecondaryAlignments, setBasesToIgnore5P/3P, setMinMQ,
 * setKeepRunning, setLog, processFile(s), processAlignments, processAlignment + notifyEndOfAlignments: APG:71-193, 201-361,
 * 377, 447) — SingleSampleVariantsDetector only changes the class it instantiates (SSVD:131).
 *
 * How the listener SPI (PileupListener.java:27-31) is honoured:
 *   * IndelRealignerPileupListener and SingleSampleVariantPileupListener — the chain of SingleSampleVariantsDetector.findSNVs
 *     (SSVD:919-925) — are driven by THIS class: they get onSequenceStart / onSequenceEnd as always and onPileup with a real
 *     PileupRecord at every EVENT position (a position where some pending read anchors an indel: anywhere else
 *     IRPL.onPileup only sets referenceSpan = 1, IRPL:87-124, and SSVPL.onPileup runs the span-1 SNV path that the device
 *     computes).  Indel / STR calls therefore come out of the reference's own unchanged code.  The span-1 decisions of every
 *     position come back from the device as ngsep_snv_call records and are appended to SSVPL.getCalledVariants() as CalledSNV
 *     objects before onSequenceEnd, with the lastIndelEnd / embedded rule of SSVPL:146-161 applied in position order.
 *   * any OTHER PileupListener (coverage statistics, a third-party reducer...) still gets every position: for it the
 *     superclass walk runs unchanged next to the device path (the reference's own CPU code, as slow as before — there is no
 *     device-side SPI for arbitrary Java code).
 */
package ngsep.discovery.gpu;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.PriorityQueue;

import ngsep.alignments.ReadAlignment;
import ngsep.discovery.AlignmentsPileupGenerator;
import ngsep.discovery.IndelRealignerPileupListener;
import ngsep.discovery.PileupListener;
import ngsep.discovery.PileupRecord;
import ngsep.discovery.SingleSampleVariantPileupListener;
import ngsep.genome.ReferenceGenome;
import ngsep.sequences.QualifiedSequence;
import ngsep.variants.CalledGenomicVariant;
import ngsep.variants.CalledSNV;
import ngsep.variants.GenomicVariant;

public class GpuAlignmentsPileupGenerator extends AlignmentsPileupGenerator implements AutoCloseable {

	/** Region tiles of this many positions (shorter sequences are one tile); up to RING - 1 tiles are in flight. */
	public static final int TILE_LEN = 2000000;
	private static final int RING = 4;
	private static final String DNA = "ACGT";

	private final int device;
	private NgsepPileupNative dev;
	private IndelRealignerPileupListener realigner;
	private SingleSampleVariantPileupListener snvListener;
	private final List<PileupListener> chain = new ArrayList<PileupListener>();       // every listener, in registration order
	private final List<PileupListener> foreign = new ArrayList<PileupListener>();     // listeners that need the per-position walk
	private final Map<String, Integer> uploaded = new HashMap<String, Integer>();

	// reads of the sequence being processed, in admission order (the pileup order, APG:416-431)
	private QualifiedSequence sequence;
	private final List<ReadAlignment> sameStartPrimary = new ArrayList<ReadAlignment>(), sameStartSecondary = new ArrayList<ReadAlignment>();
	private int lastStart = 0;
	private final List<Read> reads = new ArrayList<Read>();

	/** One admitted read with the states the realigner gave it: a state is valid for reference positions from <= p < to. */
	private static final class Read {
		final ReadAlignment aln;
		final int admitFirst;
		final List<State> closed = new ArrayList<State>();
		int curFrom = Integer.MIN_VALUE;
		Read(ReadAlignment aln) { this.aln = aln; this.admitFirst = aln.getFirst(); }
	}
	private static final class State {
		int from, to, first, last;
		short ignStart, ignEnd;
		int[] cigar;     // len * 8 + op, RA:1180-1198
		static State of(ReadAlignment a) {
			State s = new State();
			s.first = a.getFirst(); s.last = a.getLast(); s.ignStart = a.getBasesToIgnoreStart(); s.ignEnd = a.getBasesToIgnoreEnd();
			s.cigar = new int[a.getNumCigarItems()];
			for (int i = 0; i < s.cigar.length; i++) s.cigar[i] = a.getCigarItemLength(i) * 8 + a.getCigarItemOperator(i);
			return s;
		}
		boolean same(State o) { return first == o.first && last == o.last && ignStart == o.ignStart && ignEnd == o.ignEnd && Arrays.equals(cigar, o.cigar); }
	}

	public GpuAlignmentsPileupGenerator(int device) { this.device = device; }
	public GpuAlignmentsPileupGenerator() { this(0); }

	// ---- listener registration (APG:71-76) -----------------------------------------------------------------------------------
	@Override
	public void addListener(PileupListener listener) {
		chain.add(listener);
		if (listener instanceof IndelRealignerPileupListener && realigner == null) realigner = (IndelRealignerPileupListener) listener;
		else if (listener instanceof SingleSampleVariantPileupListener && snvListener == null) snvListener = (SingleSampleVariantPileupListener) listener;
		else if (listener.getClass().getName().equals("ngsep.discovery.SingleSampleVariantsDetector")) { /* only uses onSequenceEnd (SSVD:933) */ }
		else { foreign.add(listener); super.addListener(listener); }
	}

	// ---- push mode (APG:377-403, 447-452): same admission rules, reads are kept instead of walked -----------------------------------
	@Override
	public void processAlignment(ReadAlignment aln) {
		if (!foreign.isEmpty()) super.processAlignment(aln);               // the reference's own walk for listeners the device cannot serve
		if (snvListener == null) return;
		if (sequence == null || aln.getSequenceIndex() != getSequencesMetadata().indexOf(sequence)) {
			finishSequence();
			sequence = getSequencesMetadata().get(aln.getSequenceIndex());
			for (PileupListener l : chain) if (!foreign.contains(l)) l.onSequenceStart(sequence);
			lastStart = 0;
		}
		if (aln.getFirst() != lastStart) admitSameStart();
		if (aln.isSecondary()) sameStartSecondary.add(aln); else sameStartPrimary.add(aln);
		lastStart = aln.getFirst();
	}

	@Override
	public void notifyEndOfAlignments() {
		if (!foreign.isEmpty()) super.notifyEndOfAlignments();
		finishSequence();
	}

	/** APG:407-433: per read group the first maxAlnsPerStartPos reads of a start, primaries before secondaries; ignore-5'/3' set here. */
	private void admitSameStart() {
		Map<String, Integer> perGroup = new HashMap<String, Integer>();
		int max = getMaxAlnsPerStartPos();
		for (List<ReadAlignment> lst : Arrays.asList(sameStartPrimary, sameStartSecondary)) {
			for (ReadAlignment aln : lst) {
				Integer c = perGroup.get(aln.getReadGroup());
				if (c == null) perGroup.put(aln.getReadGroup(), 1);
				else if (max <= 0 || c < max) perGroup.put(aln.getReadGroup(), c + 1);
				else continue;
				aln.setBasesToIgnore5P(getBasesToIgnore5P());
				aln.setBasesToIgnore3P(getBasesToIgnore3P());
				reads.add(new Read(aln));
			}
		}
		sameStartPrimary.clear();
		sameStartSecondary.clear();
	}

	// ---- one sequence: event pre-pass, device, replay -------------------------------------------------------------------------
	private void finishSequence() {
		if (sequence == null) return;
		admitSameStart();
		try {
			if (!reads.isEmpty()) {
				eventPrePass();
				List<CalledSNV> snvs = runDevice();
				replay(snvs);
			}
		} catch (IOException e) {
			throw new RuntimeException(e);     // the listener SPI has no checked exceptions (APG:500-504)
		}
		for (PileupListener l : chain) if (!foreign.contains(l)) l.onSequenceEnd(sequence);
		reads.clear();
		sequence = null;
	}

	/**
	 * The unchanged IndelRealignerPileupListener and SingleSampleVariantPileupListener at the sparse EVENT positions.  Every
	 * read that starts at or before an event is pending when the event is processed (APG:377-403); reads the realigner
	 * mutates get a new state from that position on (their earlier state stays valid for the positions before it).
	 */
	private void eventPrePass() {
		PriorityQueue<Integer> events = new PriorityQueue<Integer>();
		List<Read> active = new ArrayList<Read>();
		int next = 0;
		int qFirst = getQuerySeq() != null ? Math.max(1, getQueryFirst()) : 1;
		int qLast = getQuerySeq() != null ? Math.min(sequence.getLength(), getQueryLast()) : sequence.getLength();
		while (next < reads.size() || !events.isEmpty()) {
			int s = next < reads.size() ? reads.get(next).admitFirst : Integer.MAX_VALUE;
			int e = events.isEmpty() ? Integer.MAX_VALUE : events.peek();
			if (s <= e) {
				Read r = reads.get(next++);
				active.add(r);
				events.addAll(r.aln.getIndelCalls().keySet());
				continue;
			}
			while (!events.isEmpty() && events.peek() == e) events.poll();
			if (e < qFirst || e > qLast) continue;                            // no pileup outside the query window (APG:481-484)
			PileupRecord pileup = new PileupRecord(sequence.getName(), e);
			List<Read> inPileup = new ArrayList<Read>();
			boolean anchored = false;
			for (int k = active.size() - 1; k >= 0; k--) if (active.get(k).aln.getLast() < e) active.remove(k);   // APG:464-471
			for (Read r : active) {
				if (r.aln.getFirst() <= e && e <= r.aln.getLast()) { pileup.addAlignment(r.aln); inPileup.add(r); anchored |= r.aln.getIndelCall(e) != null; }
			}
			if (!anchored) continue;
			List<State> before = new ArrayList<State>();
			for (Read r : inPileup) before.add(State.of(r.aln));
			realigner.onPileup(pileup);                                         // may move indels / realign read ends (IRPL:165-526)
			for (int k = 0; k < inPileup.size(); k++) {
				Read r = inPileup.get(k);
				State now = State.of(r.aln);
				if (now.same(before.get(k))) continue;
				State old = before.get(k);
				old.from = r.curFrom; old.to = e;
				r.closed.add(old);
				r.curFrom = e;
				for (Integer p : r.aln.getIndelCalls().keySet()) if (p > e) events.add(p);
			}
			if (pileup.getReferenceSpan() > 1 || pileup.isEmbedded()) snvListener.onPileup(pileup);   // span > 1: the reference's own indel / STR call (SSVPL:257-273)
		}
	}

	/** Packs region tiles into direct buffers (layout of include/ngsep_pileup.h), keeps up to RING - 1 tiles in flight, collects the calls. */
	private List<CalledSNV> runDevice() throws IOException {
		ensureDevice();
		ReferenceGenome genome = getGenome();
		Integer seqIdx = uploaded.get(sequence.getName());
		if (seqIdx == null) {
			seqIdx = uploaded.size();
			dev.refUpload(seqIdx, genome.getReference(sequence.getName(), 1, sequence.getLength()).toString().getBytes());
			uploaded.put(sequence.getName(), seqIdx);
		}
		List<CalledSNV> out = new ArrayList<CalledSNV>();
		TileBuffers[] ring = new TileBuffers[RING];
		int inFlight = 0, slot = 0;
		ByteBuffer calls = ByteBuffer.allocateDirect(65536 * NgsepPileupNative.CALL_BYTES).order(ByteOrder.LITTLE_ENDIAN);
		int maxReach = 1, maxBack = 0;
		for (Read r : reads) {
			int last = r.aln.getLast(), first = r.aln.getFirst();
			for (State v : r.closed) { last = Math.max(last, v.last); first = Math.min(first, v.first); }
			maxReach = Math.max(maxReach, last - r.admitFirst + 1);
			maxBack = Math.max(maxBack, r.admitFirst - first);
		}
		int qFirst = getQuerySeq() != null ? Math.max(1, getQueryFirst()) : 1;
		int qLast = getQuerySeq() != null ? Math.min(sequence.getLength(), getQueryLast()) : sequence.getLength();
		for (long tf = qFirst; tf <= qLast; tf += TILE_LEN) {
			int tl = (int) Math.min(qLast, tf + TILE_LEN - 1);
			int lo = lowerBound((int) tf - maxReach + 1), hi = lowerBound(tl + maxBack + 1);
			if (lo >= hi) continue;
			while (inFlight >= RING - 1) { calls = fetch(calls, out); inFlight--; }
			if (ring[slot] == null) ring[slot] = new TileBuffers();
			TileBuffers t = ring[slot];
			t.pack(reads, lo, hi);
			dev.tileSubmit(seqIdx, (int) tf, tl, t.nRecords, t.maxSpan, t.maxBack, t.hdr, t.payload, t.nWords, t.qual, t.nQual, null);
			inFlight++;
			slot = (slot + 1) % RING;
		}
		while (inFlight > 0) { calls = fetch(calls, out); inFlight--; }
		return out;
	}

	private int lowerBound(int start) {
		int lo = 0, hi = reads.size();
		while (lo < hi) { int mid = (lo + hi) >>> 1; if (reads.get(mid).admitFirst < start) lo = mid + 1; else hi = mid; }
		return lo;
	}

	private ByteBuffer fetch(ByteBuffer calls, List<CalledSNV> out) throws IOException {
		long n = dev.tileFetch(calls, null);
		if (n * NgsepPileupNative.CALL_BYTES > calls.capacity()) {               // the tile is still at the head of the queue: larger buffer, again
			calls = ByteBuffer.allocateDirect((int) (n + 16) * NgsepPileupNative.CALL_BYTES).order(ByteOrder.LITTLE_ENDIAN);
			n = dev.tileFetch(calls, null);
		}
		for (int i = 0; i < n; i++) out.add(toCalledSNV(calls, i * NgsepPileupNative.CALL_BYTES));
		return calls;
	}

	/** struct ngsep_snv_call -> CalledSNV as VariantDiscoverySNVQAlgorithm.discoverSNV builds it (VD:181-222). */
	private CalledSNV toCalledSNV(ByteBuffer b, int o) {
		int pos = b.getInt(o + 4);
		int[] counts = {b.getInt(o + 8), b.getInt(o + 12), b.getInt(o + 16), b.getInt(o + 20)};
		int depth = b.getInt(o + 24);
		double[][] logCond = new double[4][4];
		for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) logCond[i][j] = b.getDouble(o + 64 + 8 * (4 * i + j));
		short gq = b.getShort(o + 208);
		int a0 = b.get(o + 212), a1 = b.get(o + 213), ref = b.get(o + 214), kind = b.get(o + 215);
		char refBase = DNA.charAt(ref);
		char altBase = DNA.charAt(kind == 1 ? a1 : a0);                          // het: a0 = ref, a1 = alt; hom-alt: a0 = a1 = alt; tri-allelic: a0 = alt
		byte genotype = (byte) (kind == 2 ? CalledGenomicVariant.GENOTYPE_HOMOALT : CalledGenomicVariant.GENOTYPE_HETERO);
		CalledSNV snv = new CalledSNV(sequence.getName(), pos, refBase, altBase, genotype);
		snv.setAllBaseCounts(counts);
		snv.setTotalReadDepth(depth);
		snv.setAllGenotypeLogConditionals(logCond);
		snv.setGenotypeQuality(gq);
		snv.setVariantQS(b.getShort(o + 210));
		if (kind == 3) snv.setType(GenomicVariant.TYPE_MULTISNV);
		return snv;
	}

	/** SSVPL:146-161 in position order: SNVs inside a called indel / STR are embedded (dropped unless -embeddedSNVs). */
	private void replay(List<CalledSNV> snvs) {
		List<CalledGenomicVariant> called = snvListener.getCalledVariants();
		List<int[]> spans = new ArrayList<int[]>();
		for (CalledGenomicVariant v : called) {
			if (!v.getSequenceName().equals(sequence.getName())) continue;
			if (!v.isSNV() && !v.isUndecided() && !v.isHomozygousReference()) spans.add(new int[] {v.getFirst(), v.getLast()});
		}
		snvs.sort((x, y) -> Integer.compare(x.getFirst(), y.getFirst()));
		int k = 0;
		short ploidy = snvListener.getSample() != null ? snvListener.getSample().getNormalPloidy() : 2;
		for (CalledSNV snv : snvs) {
			int p = snv.getFirst();
			while (k < spans.size() && spans.get(k)[1] < p) k++;
			boolean atIndel = k < spans.size() && spans.get(k)[0] == p;
			boolean embedded = k < spans.size() && spans.get(k)[0] < p && p <= spans.get(k)[1];
			if (atIndel) continue;                                               // the position holds the indel call itself
			if (embedded && !snvListener.isCallEmbeddedSNVs()) continue;
			if (embedded) snv.setType(GenomicVariant.TYPE_EMBEDDED_SNV);
			if (snvListener.getSample() != null) snv.setSampleId(snvListener.getSample().getId());
			snv.updateAllelesCopyNumberFromCounts(ploidy);
			called.add(snv);
		}
	}

	private void ensureDevice() throws IOException {
		if (dev != null) return;
		int flags = (snvListener.isCalcStrandBias() ? NgsepPileupNative.F_STRAND_BIAS : 0) | (snvListener.isIgnoreLowerCaseRef() ? NgsepPileupNative.F_IGNORE_LOWER_REF : 0);
		double h = snvListener.getHeterozygosityRate();
		short ploidy = snvListener.getSample() != null ? snvListener.getSample().getNormalPloidy() : 2;
		dev = new NgsepPileupNative(device, snvListener.getMaxBaseQS(), snvListener.getMinQuality(), h, ploidy, flags, 1, NgsepPileupNative.logTables(h));
	}

	@Override
	public void close() {
		if (dev != null) dev.close();
		dev = null;
	}

	/** Direct, page-locked buffers of one tile in the layout of ngsep_read_hdr / payload words / quality bytes. */
	private static final class TileBuffers {
		ByteBuffer hdr, payload, qual;
		int nRecords, maxSpan, maxBack;
		long nWords, nQual;

		private static ByteBuffer grow(ByteBuffer b, long bytes) throws IOException {
			if (b != null && b.capacity() >= bytes) { b.clear(); return b; }
			if (b != null) NgsepPileupNative.hostUnregister(b);
			ByteBuffer n = ByteBuffer.allocateDirect((int) Math.min(Integer.MAX_VALUE - 8, bytes + bytes / 4 + 4096)).order(ByteOrder.LITTLE_ENDIAN);
			NgsepPileupNative.hostRegister(n);
			return n;
		}

		void pack(List<Read> reads, int lo, int hi) throws IOException {
			long recs = 0, words = 0, quals = 0;
			for (int i = lo; i < hi; i++) {
				Read r = reads.get(i);
				int L = r.aln.getReadLength(), n = r.closed.size() + 1;
				recs += n;
				words += (long) n * (3 + 256 + (L + 15) / 16 + 2 * ((L + 31) / 32) + 1);
				quals += (L + 3) & ~3;
			}
			hdr = grow(hdr, recs * NgsepPileupNative.HDR_BYTES);
			payload = grow(payload, words * 4);
			qual = grow(qual, quals);
			nRecords = 0; maxSpan = 1; maxBack = 0;
			for (int i = lo; i < hi; i++) {
				Read r = reads.get(i);
				int qoff = qual.position() / 4;
				String q = r.aln.getQualityScores();
				int L = r.aln.getReadLength();
				for (int k = 0; k < L; k++) qual.put((byte) (q != null && k < q.length() ? q.charAt(k) : '+'));     // absent QUAL reads as '+' (RA:1025)
				while ((qual.position() & 3) != 0) qual.put((byte) '!');
				boolean plain = r.closed.isEmpty() && r.aln.getFirst() == r.admitFirst;
				if (plain) { emit(r, State.of(r.aln), false, qoff); continue; }
				for (State v : r.closed) emit(r, v, true, qoff);
				State cur = State.of(r.aln);
				cur.from = r.curFrom; cur.to = Integer.MAX_VALUE;
				emit(r, cur, true, qoff);
			}
			nWords = payload.position() / 4;
			nQual = qual.position();
		}

		private void emit(Read r, State st, boolean windowed, int qoff) {
			CharSequence seq = r.aln.getReadCharacters();
			String q = r.aln.getQualityScores();
			int L = r.aln.getReadLength(), nb = (L + 15) / 16, nw = (L + 31) / 32;
			int[] b2 = new int[nb], nm = new int[nw], lq = new int[nw];
			boolean hasN = false;
			for (int k = 0; k < L; k++) {
				int c = DNA.indexOf(Character.toUpperCase(seq.charAt(k)));
				if (c < 0) { hasN = true; nm[k >> 5] |= 1 << (k & 31); c = 0; }
				b2[k >> 4] |= c << ((k & 15) * 2);
				int qc = q != null && k < q.length() ? q.charAt(k) : '+';
				if (qc - 33 <= 3) lq[k >> 5] |= 1 << (k & 31);                      // never counted (CH:214)
			}
			int meta = L | (st.cigar.length << 16) | (r.aln.isNegativeStrand() ? NgsepPileupNative.M_NEG : 0) | (hasN ? NgsepPileupNative.M_NMASK : 0)
					| (windowed ? NgsepPileupNative.M_WINDOW : 0);
			if (st.cigar.length == 1 && !hasN && !windowed && st.cigar[0] == (L << 3 | 3)) meta |= NgsepPileupNative.M_SIMPLE;
			hdr.putInt(r.admitFirst).putInt(meta).putShort(st.ignStart).putShort(st.ignEnd).putInt(payload.position() / 4);
			if (windowed) payload.putInt(st.from).putInt(st.to).putInt(st.first);
			for (int c : st.cigar) payload.putInt(c);
			for (int w : b2) payload.putInt(w);
			if (hasN) for (int w : nm) payload.putInt(w);
			for (int w : lq) payload.putInt(w);
			payload.putInt(qoff);
			maxSpan = Math.max(maxSpan, st.last - r.admitFirst + 1);
			maxBack = Math.max(maxBack, r.admitFirst - st.first);
			nRecords++;
		}
	}
}
