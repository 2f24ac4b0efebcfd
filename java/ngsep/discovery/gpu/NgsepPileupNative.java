/*
 * JNI binding of libngsep_pileup.so (include/ngsep_pileup.h) through jni/ngsep_pileup_jni.c: one native method per exported
 * C entry point (the static methods), wrapped by a small AutoCloseable that owns one context = one GPU, one thread.
 *
 * All bulk data travel in DIRECT ByteBuffers in LITTLE_ENDIAN order laid out exactly as the C structs:
 *   hdr      16 bytes per record   struct ngsep_read_hdr {int32 first; uint32 meta; int16 ign_start; int16 ign_end; uint32 off;}
 *   payload  32-bit words          [valid_from valid_to first]? cigar[n] bases2[ceil(L/16)] nmask[ceil(L/32)]? lowq[ceil(L/32)] qoff
 *   qual     phred+33 bytes        L bytes per read padded to 4, addressed by the records' qoff
 *   group    uint16 per record     read-group number (multi-sample contexts), 0xFFFF = no sample
 *   calls    224 bytes per record  struct ngsep_snv_call   /  sites 44 + samples x 104 bytes (ngsep_pop_site, ngsep_sample_call)
 * Page-lock the buffers once with hostRegister so that tileSubmit copies asynchronously.
 */
package ngsep.discovery.gpu;

import java.io.IOException;
import java.nio.ByteBuffer;

public final class NgsepPileupNative implements AutoCloseable {
	static { System.loadLibrary("ngsep_pileup_jni"); }

	public static final int F_STRAND_BIAS = 0x1, F_IGNORE_LOWER_REF = 0x2, F_EMIT_ALL = 0x8, F_QUALITY_IN_PLACE = 0x10;
	public static final int M_NEG = 0x01000000, M_NMASK = 0x02000000, M_WINDOW = 0x04000000, M_SIMPLE = 0x10000000;
	public static final int N_TABLES = 95, CALL_BYTES = 224, SITE_BYTES = 44, SAMPLE_CALL_BYTES = 104, HDR_BYTES = 16;
	/** Indexes into the statistics array the fetch methods fill (a long[11]; a shorter array gets the leading entries). */
	public static final int ST_COVERED = 0, ST_CANDIDATES = 1, ST_CALLS = 2, ST_ALIGNED_BASES = 3, ST_KERNEL_US = 4, ST_WINDOW_KERNEL_US = 5, ST_LAUNCHES = 6,
			ST_OPEN_COLUMNS = 7, ST_H2D_BYTES = 8, ST_QUALITY_IN_PLACE = 9, ST_D2H_BYTES = 10;

	// ---- one native method per C entry point (jni/ngsep_pileup_jni.c) -----------------------------------------------------
	static native long ctxCreate(int device, int maxBaseQS, int minQuality, double hetRate, int ploidy, int flags, int nSamples, double[] tables) throws IOException;
	static native void ctxDestroy(long ctx);
	static native String lastError(long ctx);
	static native void refUpload(long ctx, int seqIdx, byte[] bases) throws IOException;
	static native void tileSubmit(long ctx, int seqIdx, int first, int last, int nReads, int maxSpan, int maxBack, ByteBuffer hdr, ByteBuffer payload, long nWords,
			ByteBuffer qual, long nQual, ByteBuffer group) throws IOException;
	static native void tileSites(long ctx, int nSites, int[] pos, short[] alleles) throws IOException;
	static native void tileAlleleCounts(long ctx, int seqIdx, int first, int last, int nReads, int maxSpan, int maxBack, ByteBuffer hdr, ByteBuffer payload, long nWords,
			ByteBuffer qual, long nQual, int minRD, int minBaseQS, ByteBuffer out) throws IOException;
	static native void indelTallies(long ctx, int[] alleleFirst, int[] callFirst, int[] alleleOff, int[] callOff, byte[] alleleChars, byte[] callChars, byte[] callQuals,
			byte[] callNeg, double logErr, int[] total, int[] counts, int[] strand, double[] diag, int[] callIndex, double[] callFit) throws IOException;
	static native long tileFetch(long ctx, ByteBuffer calls, long[] stats) throws IOException;
	static native void samplesSet(long ctx, int nGroups, int[] sampleGroupOff, short[] sampleGroups, double minAlleleDepthFrequency) throws IOException;
	static native long tileFetchPop(long ctx, ByteBuffer sites, ByteBuffer calls, long[] stats) throws IOException;
	static native long residentFetchPop(long ctx, int slot, ByteBuffer sites, ByteBuffer calls, long[] stats) throws IOException;
	static native int residentAdd(long ctx, int seqIdx, int first, int last, int nReads, int maxSpan, int maxBack, ByteBuffer hdr, ByteBuffer payload, long nWords,
			ByteBuffer qual, long nQual, ByteBuffer group) throws IOException;
	static native void residentClear(long ctx) throws IOException;
	static native void residentRun(long ctx, int slot) throws IOException;
	static native long residentFetch(long ctx, int slot, ByteBuffer calls, long[] stats) throws IOException;
	static native void residentStats(long ctx, long[] stats) throws IOException;
	static native void coverageHistogram(long ctx, int tileFirst, int tileLast, int[] first, int[] last, byte[] unique, int maxCoverage, long[] histAll, long[] histUnique) throws IOException;
	static native void synthReference(long ctx, int seqIdx, long length, long seed, double gc) throws IOException;
	static native int synthTile(long ctx, int seqIdx, int first, int last, long seed, double depth, int readLen, double snvRate, double errorRate) throws IOException;
	static native void residentDownload(long ctx, int slot, long[] shape, ByteBuffer hdr, ByteBuffer payload, ByteBuffer qual) throws IOException;
	static native void refDownload(long ctx, int seqIdx, byte[] bases) throws IOException;
	static native void timerStart(long ctx) throws IOException;
	static native float timerStop(long ctx) throws IOException;
	public static native void hostRegister(ByteBuffer buf) throws IOException;
	public static native void hostUnregister(ByteBuffer buf) throws IOException;
	public static native int abiVersion();
	public static native int builtSm();

	// ---- the context ------------------------------------------------------------------------------------------------------
	private long ctx;

	/**
	 * @param tables the 95 doubles of ngsep_pileup.h computed with the JVM's own Math.log10 / Math.pow exactly as
	 *               CountsHelper.updateProbabilitiesCache does (discovery/CountsHelper.java:149-178, 414-415): see logTables
	 */
	public NgsepPileupNative(int device, int maxBaseQS, int minQuality, double hetRate, int ploidy, int flags, int nSamples, double[] tables) throws IOException {
		ctx = ctxCreate(device, maxBaseQS, minQuality, hetRate, ploidy, flags, nSamples, tables);
	}

	/** tables[3q] = log10(1-e_q), tables[3q+1] = -0.1q - log10(3), tables[3q+2] = log10(.5(1-e_q) + .5 e_q/3), q = 3..30; [93], [94] the priors. */
	public static double[] logTables(double hetRate) {
		double[] t = new double[N_TABLES];
		for (int q = 3; q <= 30; q++) {
			double e = Math.pow(10, -0.1 * q);
			t[3 * q] = Math.log10(1 - e);
			t[3 * q + 1] = -0.1 * q - Math.log10(3);
			t[3 * q + 2] = Math.log10(0.5 * (1 - e) + 0.5 * e / 3);
		}
		t[93] = Math.log10((1 - hetRate) / 4);
		t[94] = Math.log10(hetRate / 12);
		return t;
	}

	public void refUpload(int seqIdx, byte[] bases) throws IOException { refUpload(ctx, seqIdx, bases); }
	public void samplesSet(int nGroups, int[] sampleGroupOff, short[] sampleGroups, double minAlleleDepthFrequency) throws IOException {
		samplesSet(ctx, nGroups, sampleGroupOff, sampleGroups, minAlleleDepthFrequency);
	}
	/** Asynchronous; up to four tiles may be in flight.  The buffers must stay untouched until the tileFetch that returns the tile. */
	public void tileSubmit(int seqIdx, int first, int last, int nReads, int maxSpan, int maxBack, ByteBuffer hdr, ByteBuffer payload, long nWords, ByteBuffer qual,
			long nQual, ByteBuffer group) throws IOException {
		tileSubmit(ctx, seqIdx, first, last, nReads, maxSpan, maxBack, hdr, payload, nWords, qual, nQual, group);
	}
	/** -knownVariants: the SNV sites of the NEXT tileSubmit (position; allele base indexes 2 bits each, reference first, their number in bits 8-10):
	 *  that tile is genotyped at these sites (SingleSampleVariantPileupListener.genotypeVariantSample) instead of being searched for variants. */
	public void tileSites(int nSites, int[] pos, short[] alleles) throws IOException { tileSites(ctx, nSites, pos, alleles); }
	/** RelativeAlleleCountsCalculator.onPileup over a tile: per position (largest allele count | alleles << 28, second largest count), 8 bytes each, into out. */
	public void tileAlleleCounts(int seqIdx, int first, int last, int nReads, int maxSpan, int maxBack, ByteBuffer hdr, ByteBuffer payload, long nWords, ByteBuffer qual,
			long nQual, int minRD, int minBaseQS, ByteBuffer out) throws IOException {
		tileAlleleCounts(ctx, seqIdx, first, last, nReads, maxSpan, maxBack, hdr, payload, nWords, qual, nQual, minRD, minBaseQS, out);
	}
	/** The per-call part of CountsHelper.updateCountsIndel for a batch of span > 1 events (include/ngsep_pileup.h: ngsep_indel_events): per allele counts,
	 *  strand counts and the diagonal of logConditionalProbs, per call the allele it counts for and its fit; the caller adds the mixture terms
	 *  (CountsHelper.java:290-300) with its own Math.log10 / Math.pow and goes on with getPosteriorProbabilities. */
	public void indelTallies(int[] alleleFirst, int[] callFirst, int[] alleleOff, int[] callOff, byte[] alleleChars, byte[] callChars, byte[] callQuals, byte[] callNeg,
			double logErr, int[] total, int[] counts, int[] strand, double[] diag, int[] callIndex, double[] callFit) throws IOException {
		indelTallies(ctx, alleleFirst, callFirst, alleleOff, callOff, alleleChars, callChars, callQuals, callNeg, logErr, total, counts, strand, diag, callIndex, callFit);
	}
	/** Number of ngsep_snv_call records written to calls; a value above calls.capacity()/224 means: call again with room for that many. */
	public long tileFetch(ByteBuffer calls, long[] stats) throws IOException { return tileFetch(ctx, calls, stats); }
	public long tileFetchPop(ByteBuffer sites, ByteBuffer calls, long[] stats) throws IOException { return tileFetchPop(ctx, sites, calls, stats); }
	public int residentAdd(int seqIdx, int first, int last, int nReads, int maxSpan, int maxBack, ByteBuffer hdr, ByteBuffer payload, long nWords, ByteBuffer qual,
			long nQual, ByteBuffer group) throws IOException {
		return residentAdd(ctx, seqIdx, first, last, nReads, maxSpan, maxBack, hdr, payload, nWords, qual, nQual, group);
	}
	public void residentRun(int slot) throws IOException { residentRun(ctx, slot); }
	public long residentFetch(int slot, ByteBuffer calls, long[] stats) throws IOException { return residentFetch(ctx, slot, calls, stats); }
	public long residentFetchPop(int slot, ByteBuffer sites, ByteBuffer calls, long[] stats) throws IOException { return residentFetchPop(ctx, slot, sites, calls, stats); }
	public void residentStats(long[] stats) throws IOException { residentStats(ctx, stats); }
	public void residentClear() throws IOException { residentClear(ctx); }
	public void timerStart() throws IOException { timerStart(ctx); }
	public float timerStop() throws IOException { return timerStop(ctx); }
	public String lastError() { return lastError(ctx); }

	@Override
	public void close() {
		if (ctx != 0) ctxDestroy(ctx);
		ctx = 0;
	}
}
