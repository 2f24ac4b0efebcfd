/*
 * ngsep_pileup_jni.c — JNI shim between ngsep.discovery.gpu.NgsepPileupNative (java/ngsep/discovery/gpu/) and the C ABI of
 * include/ngsep_pileup.h.  One JNI function per exported entry point; buffers are direct ByteBuffers (little endian) laid out
 * exactly as the C structs, so nothing is converted here.  Errors become java.io.IOException carrying ngsep_last_error(),
 * the exception type the reference's detectors already declare (SingleSampleVariantsDetector.findSNVs throws IOException,
 * src/ngsep/discovery/SingleSampleVariantsDetector.java:896).
 *
 * Build (needs a JDK; none exists in this repository's image, so the Makefile next to this file is gated on JAVA_HOME):
 *   cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include ngsep_pileup_jni.c \
 *      -L../ngsepcore_b200 -lngsep_pileup -Wl,-rpath,'$ORIGIN' -o libngsep_pileup_jni.so
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ngsep_pileup.h"

#define JNI_FN(name) Java_ngsep_discovery_gpu_NgsepPileupNative_##name

static void throw_io(JNIEnv* env, void* ctx, int rc, const char* what) {
    char msg[640];
    snprintf(msg, sizeof msg, "%s failed (%d): %s", what, rc, ngsep_last_error(ctx));
    jclass cls = (*env)->FindClass(env, "java/io/IOException");
    if (cls) (*env)->ThrowNew(env, cls, msg);
}
static void* direct(JNIEnv* env, jobject buf) { return buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL; }
static void fill_stats(JNIEnv* env, jlongArray out, const ngsep_tile_stats* s) {
    if (!out) return;
    /* covered, candidates, calls, aligned bases, kernel microseconds, window-kernel microseconds, launches, open columns, H2D bytes, quality in place, D2H bytes */
    jlong v[11] = {(jlong)s->covered_positions, (jlong)s->candidate_positions, (jlong)s->calls, (jlong)s->aligned_bases, (jlong)(s->kernel_ms * 1000.0f),
                   (jlong)(s->pileup_ms * 1000.0f), (jlong)s->kernel_launches, (jlong)s->open_columns, (jlong)s->h2d_bytes, (jlong)s->quality_in_place,
                   (jlong)s->d2h_bytes};
    jsize n = (*env)->GetArrayLength(env, out);
    (*env)->SetLongArrayRegion(env, out, 0, n < 11 ? n : 11, v);
}
static void fill_tile(JNIEnv* env, ngsep_tile* t, jint seq, jint first, jint last, jint nReads, jint maxSpan, jint maxBack, jobject hdr, jobject payload,
                      jlong nWords, jobject qual, jlong nQual, jobject group) {
    memset(t, 0, sizeof *t);
    t->seq_idx = seq; t->tile_first = first; t->tile_last = last; t->n_reads = nReads; t->max_span = maxSpan; t->max_back = maxBack;
    t->n_words = (uint64_t)nWords; t->n_qual = (uint64_t)nQual;
    t->hdr = (const ngsep_read_hdr*)direct(env, hdr);
    t->payload = (const uint32_t*)direct(env, payload);
    t->qual = (const uint8_t*)direct(env, qual);
    t->group = (const uint16_t*)direct(env, group);
}

/* ngsep_ctx_create */
JNIEXPORT jlong JNICALL JNI_FN(ctxCreate)(JNIEnv* env, jclass cls, jint device, jint maxBaseQS, jint minQuality, jdouble hetRate, jint ploidy, jint flags,
                                          jint nSamples, jdoubleArray tables) {
    (void)cls;
    ngsep_params p;
    memset(&p, 0, sizeof p);
    p.max_base_qs = maxBaseQS; p.min_quality = minQuality; p.het_rate = hetRate; p.ploidy = ploidy; p.flags = (uint32_t)flags; p.n_samples = nSamples;
    jsize n = (*env)->GetArrayLength(env, tables);
    jdouble* t = (*env)->GetDoubleArrayElements(env, tables, NULL);
    void* ctx = NULL;
    int rc = ngsep_ctx_create(device, &p, t, (size_t)n, &ctx);
    (*env)->ReleaseDoubleArrayElements(env, tables, t, JNI_ABORT);
    if (rc != NGSEP_OK) { throw_io(env, NULL, rc, "ngsep_ctx_create"); return 0; }
    return (jlong)(intptr_t)ctx;
}
/* ngsep_ctx_destroy */
JNIEXPORT void JNICALL JNI_FN(ctxDestroy)(JNIEnv* env, jclass cls, jlong ctx) { (void)env; (void)cls; if (ctx) ngsep_ctx_destroy((void*)(intptr_t)ctx); }
/* ngsep_last_error */
JNIEXPORT jstring JNICALL JNI_FN(lastError)(JNIEnv* env, jclass cls, jlong ctx) { (void)cls; return (*env)->NewStringUTF(env, ngsep_last_error((void*)(intptr_t)ctx)); }
/* ngsep_ref_upload */
JNIEXPORT void JNICALL JNI_FN(refUpload)(JNIEnv* env, jclass cls, jlong ctx, jint seqIdx, jbyteArray bases) {
    (void)cls;
    jsize n = (*env)->GetArrayLength(env, bases);
    jbyte* b = (*env)->GetByteArrayElements(env, bases, NULL);
    int rc = ngsep_ref_upload((void*)(intptr_t)ctx, seqIdx, (const char*)b, (int64_t)n);
    (*env)->ReleaseByteArrayElements(env, bases, b, JNI_ABORT);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_ref_upload");
}
/* ngsep_tile_submit */
JNIEXPORT void JNICALL JNI_FN(tileSubmit)(JNIEnv* env, jclass cls, jlong ctx, jint seq, jint first, jint last, jint nReads, jint maxSpan, jint maxBack,
                                          jobject hdr, jobject payload, jlong nWords, jobject qual, jlong nQual, jobject group) {
    (void)cls;
    ngsep_tile t;
    fill_tile(env, &t, seq, first, last, nReads, maxSpan, maxBack, hdr, payload, nWords, qual, nQual, group);
    int rc = ngsep_tile_submit((void*)(intptr_t)ctx, &t);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_tile_submit");
}
/* ngsep_tile_sites: the known SNV sites (-knownVariants) of the next submitted tile */
JNIEXPORT void JNICALL JNI_FN(tileSites)(JNIEnv* env, jclass cls, jlong ctx, jint nSites, jintArray pos, jshortArray alleles) {
    (void)cls;
    jint* p = (*env)->GetIntArrayElements(env, pos, NULL);
    jshort* a = (*env)->GetShortArrayElements(env, alleles, NULL);
    int rc = ngsep_tile_sites((void*)(intptr_t)ctx, nSites, (const int32_t*)p, (const uint16_t*)a);
    (*env)->ReleaseIntArrayElements(env, pos, p, JNI_ABORT);
    (*env)->ReleaseShortArrayElements(env, alleles, a, JNI_ABORT);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_tile_sites");
}
/* ngsep_tile_allele_counts: out = direct IntBuffer-backed ByteBuffer of 8 bytes per tile position */
JNIEXPORT void JNICALL JNI_FN(tileAlleleCounts)(JNIEnv* env, jclass cls, jlong ctx, jint seq, jint first, jint last, jint nReads, jint maxSpan, jint maxBack,
                                                jobject hdr, jobject payload, jlong nWords, jobject qual, jlong nQual, jint minRD, jint minBaseQS, jobject out) {
    (void)cls;
    ngsep_tile t;
    fill_tile(env, &t, seq, first, last, nReads, maxSpan, maxBack, hdr, payload, nWords, qual, nQual, NULL);
    int rc = ngsep_tile_allele_counts((void*)(intptr_t)ctx, &t, minRD, minBaseQS, (uint32_t*)direct(env, out));
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_tile_allele_counts");
}
/* ngsep_indel_tallies: the events arrive as arrays (offsets int[], strings byte[]) */
JNIEXPORT void JNICALL JNI_FN(indelTallies)(JNIEnv* env, jclass cls, jlong ctx, jintArray alleleFirst, jintArray callFirst, jintArray alleleOff, jintArray callOff,
                                            jbyteArray alleleChars, jbyteArray callChars, jbyteArray callQuals, jbyteArray callNeg, jdouble logErr,
                                            jintArray total, jintArray counts, jintArray strand, jdoubleArray diag, jintArray callIndex, jdoubleArray callFit) {
    (void)cls;
    ngsep_indel_events e;
    memset(&e, 0, sizeof e);
    e.n_events = (int32_t)(*env)->GetArrayLength(env, alleleFirst) - 1;
    e.n_alleles = (int32_t)(*env)->GetArrayLength(env, alleleOff) - 1;
    e.n_calls = (int32_t)(*env)->GetArrayLength(env, callOff) - 1;
    jint* af = (*env)->GetIntArrayElements(env, alleleFirst, NULL); jint* cf = (*env)->GetIntArrayElements(env, callFirst, NULL);
    jint* ao = (*env)->GetIntArrayElements(env, alleleOff, NULL); jint* co = (*env)->GetIntArrayElements(env, callOff, NULL);
    jbyte* ac = (*env)->GetByteArrayElements(env, alleleChars, NULL); jbyte* cc = (*env)->GetByteArrayElements(env, callChars, NULL);
    jbyte* cq = (*env)->GetByteArrayElements(env, callQuals, NULL); jbyte* cn = (*env)->GetByteArrayElements(env, callNeg, NULL);
    jint* ot = (*env)->GetIntArrayElements(env, total, NULL); jint* oc = (*env)->GetIntArrayElements(env, counts, NULL); jint* os = (*env)->GetIntArrayElements(env, strand, NULL);
    jdouble* od = (*env)->GetDoubleArrayElements(env, diag, NULL); jint* oi = (*env)->GetIntArrayElements(env, callIndex, NULL);
    jdouble* of = (*env)->GetDoubleArrayElements(env, callFit, NULL);
    e.allele_first = (const int32_t*)af; e.call_first = (const int32_t*)cf; e.allele_off = (const int32_t*)ao; e.call_off = (const int32_t*)co;
    e.allele_chars = (const char*)ac; e.call_chars = (const char*)cc; e.call_quals = (const uint8_t*)cq; e.call_neg = (const uint8_t*)cn;
    e.log_err = logErr;
    int rc = ngsep_indel_tallies((void*)(intptr_t)ctx, &e, (int32_t*)ot, (int32_t*)oc, (int32_t*)os, od, (int32_t*)oi, of);
    (*env)->ReleaseIntArrayElements(env, alleleFirst, af, JNI_ABORT); (*env)->ReleaseIntArrayElements(env, callFirst, cf, JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, alleleOff, ao, JNI_ABORT); (*env)->ReleaseIntArrayElements(env, callOff, co, JNI_ABORT);
    (*env)->ReleaseByteArrayElements(env, alleleChars, ac, JNI_ABORT); (*env)->ReleaseByteArrayElements(env, callChars, cc, JNI_ABORT);
    (*env)->ReleaseByteArrayElements(env, callQuals, cq, JNI_ABORT); (*env)->ReleaseByteArrayElements(env, callNeg, cn, JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, total, ot, 0); (*env)->ReleaseIntArrayElements(env, counts, oc, 0); (*env)->ReleaseIntArrayElements(env, strand, os, 0);
    (*env)->ReleaseDoubleArrayElements(env, diag, od, 0); (*env)->ReleaseIntArrayElements(env, callIndex, oi, 0); (*env)->ReleaseDoubleArrayElements(env, callFit, of, 0);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_indel_tallies");
}
/* ngsep_tile_fetch: returns the number of calls; a value above the buffer's capacity means "call again with a larger buffer" */
JNIEXPORT jlong JNICALL JNI_FN(tileFetch)(JNIEnv* env, jclass cls, jlong ctx, jobject calls, jlongArray stats) {
    (void)cls;
    size_t n = 0, cap = calls ? (size_t)(*env)->GetDirectBufferCapacity(env, calls) / sizeof(ngsep_snv_call) : 0;
    ngsep_tile_stats st;
    memset(&st, 0, sizeof st);
    int rc = ngsep_tile_fetch((void*)(intptr_t)ctx, (ngsep_snv_call*)direct(env, calls), cap, &n, NULL, &st);
    if (rc == NGSEP_EOVERFLOW && n > cap) return (jlong)n;
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_tile_fetch"); return -1; }
    fill_stats(env, stats, &st);
    return (jlong)n;
}
/* ngsep_samples_set */
JNIEXPORT void JNICALL JNI_FN(samplesSet)(JNIEnv* env, jclass cls, jlong ctx, jint nGroups, jintArray sampleGroupOff, jshortArray sampleGroups, jdouble minAlleleDepthFrequency) {
    (void)cls;
    jint* off = (*env)->GetIntArrayElements(env, sampleGroupOff, NULL);
    jshort* g = (*env)->GetShortArrayElements(env, sampleGroups, NULL);
    int rc = ngsep_samples_set((void*)(intptr_t)ctx, nGroups, (const int32_t*)off, (const uint16_t*)g, minAlleleDepthFrequency);
    (*env)->ReleaseIntArrayElements(env, sampleGroupOff, off, JNI_ABORT);
    (*env)->ReleaseShortArrayElements(env, sampleGroups, g, JNI_ABORT);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_samples_set");
}
static jlong fetch_pop(JNIEnv* env, jlong ctx, int slot, jobject sites, jobject calls, jlongArray stats) {
    size_t n = 0, cap = sites ? (size_t)(*env)->GetDirectBufferCapacity(env, sites) / sizeof(ngsep_pop_site) : 0;
    ngsep_tile_stats st;
    memset(&st, 0, sizeof st);
    int rc = slot < 0 ? ngsep_tile_fetch_pop((void*)(intptr_t)ctx, (ngsep_pop_site*)direct(env, sites), cap, &n, (ngsep_sample_call*)direct(env, calls), NULL, &st)
                      : ngsep_resident_fetch_pop((void*)(intptr_t)ctx, slot, (ngsep_pop_site*)direct(env, sites), cap, &n, (ngsep_sample_call*)direct(env, calls), NULL, &st);
    if (rc == NGSEP_EOVERFLOW && n > cap) return (jlong)n;
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_tile_fetch_pop"); return -1; }
    fill_stats(env, stats, &st);
    return (jlong)n;
}
/* ngsep_tile_fetch_pop */
JNIEXPORT jlong JNICALL JNI_FN(tileFetchPop)(JNIEnv* env, jclass cls, jlong ctx, jobject sites, jobject calls, jlongArray stats) { (void)cls; return fetch_pop(env, ctx, -1, sites, calls, stats); }
/* ngsep_resident_fetch_pop */
JNIEXPORT jlong JNICALL JNI_FN(residentFetchPop)(JNIEnv* env, jclass cls, jlong ctx, jint slot, jobject sites, jobject calls, jlongArray stats) { (void)cls; return fetch_pop(env, ctx, slot, sites, calls, stats); }
/* ngsep_resident_add */
JNIEXPORT jint JNICALL JNI_FN(residentAdd)(JNIEnv* env, jclass cls, jlong ctx, jint seq, jint first, jint last, jint nReads, jint maxSpan, jint maxBack,
                                           jobject hdr, jobject payload, jlong nWords, jobject qual, jlong nQual, jobject group) {
    (void)cls;
    ngsep_tile t;
    fill_tile(env, &t, seq, first, last, nReads, maxSpan, maxBack, hdr, payload, nWords, qual, nQual, group);
    int slot = 0;
    int rc = ngsep_resident_add((void*)(intptr_t)ctx, &t, &slot);
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_resident_add"); return -1; }
    return slot;
}
/* ngsep_resident_clear */
JNIEXPORT void JNICALL JNI_FN(residentClear)(JNIEnv* env, jclass cls, jlong ctx) { (void)cls; int rc = ngsep_resident_clear((void*)(intptr_t)ctx); if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_resident_clear"); }
/* ngsep_resident_run */
JNIEXPORT void JNICALL JNI_FN(residentRun)(JNIEnv* env, jclass cls, jlong ctx, jint slot) { (void)cls; int rc = ngsep_resident_run((void*)(intptr_t)ctx, slot); if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_resident_run"); }
/* ngsep_resident_fetch */
JNIEXPORT jlong JNICALL JNI_FN(residentFetch)(JNIEnv* env, jclass cls, jlong ctx, jint slot, jobject calls, jlongArray stats) {
    (void)cls;
    size_t n = 0, cap = calls ? (size_t)(*env)->GetDirectBufferCapacity(env, calls) / sizeof(ngsep_snv_call) : 0;
    ngsep_tile_stats st;
    memset(&st, 0, sizeof st);
    int rc = ngsep_resident_fetch((void*)(intptr_t)ctx, slot, (ngsep_snv_call*)direct(env, calls), cap, &n, NULL, &st);
    if (rc == NGSEP_EOVERFLOW && n > cap) return (jlong)n;
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_resident_fetch"); return -1; }
    fill_stats(env, stats, &st);
    return (jlong)n;
}
/* ngsep_resident_stats */
JNIEXPORT void JNICALL JNI_FN(residentStats)(JNIEnv* env, jclass cls, jlong ctx, jlongArray stats) {
    (void)cls;
    ngsep_tile_stats st;
    memset(&st, 0, sizeof st);
    int rc = ngsep_resident_stats((void*)(intptr_t)ctx, &st);
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_resident_stats"); return; }
    fill_stats(env, stats, &st);
}
/* ngsep_coverage_histogram: histAll / histUnique hold maxCoverage + 1 counts */
JNIEXPORT void JNICALL JNI_FN(coverageHistogram)(JNIEnv* env, jclass cls, jlong ctx, jint tileFirst, jint tileLast, jintArray first, jintArray last, jbyteArray unique,
                                                 jint maxCoverage, jlongArray histAll, jlongArray histUnique) {
    (void)cls;
    jsize n = (*env)->GetArrayLength(env, first);
    jint* f = (*env)->GetIntArrayElements(env, first, NULL);
    jint* l = (*env)->GetIntArrayElements(env, last, NULL);
    jbyte* u = (*env)->GetByteArrayElements(env, unique, NULL);
    jlong* ha = (*env)->GetLongArrayElements(env, histAll, NULL);
    jlong* hu = (*env)->GetLongArrayElements(env, histUnique, NULL);
    int rc = ngsep_coverage_histogram((void*)(intptr_t)ctx, tileFirst, tileLast, (int64_t)n, (const int32_t*)f, (const int32_t*)l, (const uint8_t*)u, maxCoverage,
                                      (uint64_t*)ha, (uint64_t*)hu);
    (*env)->ReleaseIntArrayElements(env, first, f, JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, last, l, JNI_ABORT);
    (*env)->ReleaseByteArrayElements(env, unique, u, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, histAll, ha, 0);
    (*env)->ReleaseLongArrayElements(env, histUnique, hu, 0);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_coverage_histogram");
}
/* ngsep_synth_reference */
JNIEXPORT void JNICALL JNI_FN(synthReference)(JNIEnv* env, jclass cls, jlong ctx, jint seqIdx, jlong length, jlong seed, jdouble gc) {
    (void)cls;
    int rc = ngsep_synth_reference((void*)(intptr_t)ctx, seqIdx, (int64_t)length, (uint64_t)seed, gc);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_synth_reference");
}
/* ngsep_synth_tile */
JNIEXPORT jint JNICALL JNI_FN(synthTile)(JNIEnv* env, jclass cls, jlong ctx, jint seqIdx, jint first, jint last, jlong seed, jdouble depth, jint readLen, jdouble snvRate,
                                         jdouble errorRate) {
    (void)cls;
    ngsep_synth_params p;
    memset(&p, 0, sizeof p);
    p.seed = (uint64_t)seed; p.depth = depth; p.read_len = readLen; p.snv_rate = snvRate; p.error_rate = errorRate;
    int slot = 0;
    int rc = ngsep_synth_tile((void*)(intptr_t)ctx, seqIdx, first, last, &p, &slot);
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_synth_tile"); return -1; }
    return slot;
}
/* ngsep_resident_download: shape = {n_reads, n_words, n_qual, qual_bits}; buffers may be null */
JNIEXPORT void JNICALL JNI_FN(residentDownload)(JNIEnv* env, jclass cls, jlong ctx, jint slot, jlongArray shape, jobject hdr, jobject payload, jobject qual) {
    (void)cls;
    ngsep_tile t;
    int rc = ngsep_resident_download((void*)(intptr_t)ctx, slot, &t, (ngsep_read_hdr*)direct(env, hdr), (uint32_t*)direct(env, payload), (uint8_t*)direct(env, qual));
    if (rc != NGSEP_OK) { throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_resident_download"); return; }
    jlong v[4] = {(jlong)t.n_reads, (jlong)t.n_words, (jlong)t.n_qual, (jlong)t.qual_bits};
    if (shape) (*env)->SetLongArrayRegion(env, shape, 0, 4, v);
}
/* ngsep_ref_download */
JNIEXPORT void JNICALL JNI_FN(refDownload)(JNIEnv* env, jclass cls, jlong ctx, jint seqIdx, jbyteArray bases) {
    (void)cls;
    jsize n = (*env)->GetArrayLength(env, bases);
    jbyte* b = (*env)->GetByteArrayElements(env, bases, NULL);
    int rc = ngsep_ref_download((void*)(intptr_t)ctx, seqIdx, (char*)b, (int64_t)n);
    (*env)->ReleaseByteArrayElements(env, bases, b, 0);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_ref_download");
}
/* ngsep_timer_start */
JNIEXPORT void JNICALL JNI_FN(timerStart)(JNIEnv* env, jclass cls, jlong ctx) { (void)cls; int rc = ngsep_timer_start((void*)(intptr_t)ctx); if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_timer_start"); }
/* ngsep_timer_stop */
JNIEXPORT jfloat JNICALL JNI_FN(timerStop)(JNIEnv* env, jclass cls, jlong ctx) {
    (void)cls;
    float ms = 0.f;
    int rc = ngsep_timer_stop((void*)(intptr_t)ctx, &ms);
    if (rc != NGSEP_OK) throw_io(env, (void*)(intptr_t)ctx, rc, "ngsep_timer_stop");
    return ms;
}
/* ngsep_host_register */
JNIEXPORT void JNICALL JNI_FN(hostRegister)(JNIEnv* env, jclass cls, jobject buf) {
    (void)cls;
    int rc = ngsep_host_register(direct(env, buf), (size_t)(*env)->GetDirectBufferCapacity(env, buf));
    if (rc != NGSEP_OK) throw_io(env, NULL, rc, "ngsep_host_register");
}
/* ngsep_host_unregister */
JNIEXPORT void JNICALL JNI_FN(hostUnregister)(JNIEnv* env, jclass cls, jobject buf) { (void)cls; int rc = ngsep_host_unregister(direct(env, buf)); if (rc != NGSEP_OK) throw_io(env, NULL, rc, "ngsep_host_unregister"); }
/* ngsep_abi_version */
JNIEXPORT jint JNICALL JNI_FN(abiVersion)(JNIEnv* env, jclass cls) { (void)env; (void)cls; return ngsep_abi_version(); }
/* ngsep_built_sm */
JNIEXPORT jint JNICALL JNI_FN(builtSm)(JNIEnv* env, jclass cls) { (void)env; (void)cls; return ngsep_built_sm(); }
