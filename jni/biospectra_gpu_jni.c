/*
 * biospectra_gpu_jni.c -- thin JNI shim between biospectra.gpu.BioSpectraGpu (jni/java/) and the C ABI of
 * libbiospectra_gpu.so (include/biospectra_gpu.h).  One native per C call; Java arrays are pinned or copied for the
 * duration of the call only; error codes become the exceptions the reference's classes throw today:
 *   BSP_ERR_ARG -> IllegalArgumentException (Indexer.java:65-79, Classifier.java:72-94,342-344)
 *   BSP_ERR_IO  -> java.io.IOException      (Indexer.close, IndexWriter / DirectoryReader failures)
 *   others      -> RuntimeException
 *
 * Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/biospectra_gpu_jni.c \
 *       -Lbiospectra_b200 -lbiospectra_gpu -Wl,-rpath,'$ORIGIN' -o libbiospectra_gpu_jni.so
 * This image has no JDK: tests/test_abi.py compiles this file against jni/stub/jni.h (declarations only).
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "biospectra_gpu.h"

#define H_INDEXER(h) ((bsp_indexer*)(intptr_t)(h))
#define H_CLASSIFIER(h) ((bsp_classifier*)(intptr_t)(h))

static void throw_for(JNIEnv* env, int rc, const void* h) {
    const char* cls = rc == BSP_ERR_ARG ? "java/lang/IllegalArgumentException"
                    : rc == BSP_ERR_IO  ? "java/io/IOException"
                    : rc == BSP_ERR_NOMEM ? "java/lang/OutOfMemoryError" : "java/lang/RuntimeException";
    jclass c = (*env)->FindClass(env, cls);
    const char* msg = bsp_last_error(h);
    if (c) (*env)->ThrowNew(env, c, msg ? msg : "biospectra_gpu error");
}

/* Java String -> modified UTF-8 C string; NULL stays NULL.  Release with put_utf. */
static const char* get_utf(JNIEnv* env, jstring s) { return s ? (*env)->GetStringUTFChars(env, s, NULL) : NULL; }
static void put_utf(JNIEnv* env, jstring s, const char* p) { if (s && p) (*env)->ReleaseStringUTFChars(env, s, p); }

/* ------------------------------------------------------------------ Indexer */
JNIEXPORT jlong JNICALL Java_biospectra_gpu_BioSpectraGpu_indexCreate(JNIEnv* env, jclass cls, jint kmerSize,
        jboolean minStrand, jint scoring, jint workerThreads, jint ramBuffer, jint device, jstring indexPath) {
    (void)cls;
    bsp_config conf;
    bsp_config_default(&conf);
    conf.kmer_size = kmerSize; conf.min_strand_kmer = minStrand ? 1 : 0; conf.scoring_algorithm = scoring;
    conf.worker_threads = workerThreads; conf.index_ram_buffer = ramBuffer; conf.device = device;
    const char* p = get_utf(env, indexPath);
    bsp_indexer* h = NULL;
    int rc = bsp_index_create(&conf, p, &h);
    put_utf(env, indexPath, p);
    if (rc) { throw_for(env, rc, NULL); return 0; }
    return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_indexAddEntry(JNIEnv* env, jclass cls, jlong handle,
        jstring filename, jstring header, jstring taxonomy, jbyteArray sequence) {
    (void)cls;
    if (!sequence) { (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), "sequence is null"); return; }
    const char* fn = get_utf(env, filename);
    const char* hd = get_utf(env, header);
    const char* tx = get_utf(env, taxonomy);
    jsize len = (*env)->GetArrayLength(env, sequence);
    jbyte* s = (*env)->GetByteArrayElements(env, sequence, NULL);        /* held for the call only */
    int rc = s ? bsp_index_add_entry(H_INDEXER(handle), fn ? fn : "", hd ? hd : "", tx ? tx : "", (const char*)s, (int64_t)len)
               : BSP_ERR_NOMEM;
    if (s) (*env)->ReleaseByteArrayElements(env, sequence, s, JNI_ABORT);
    put_utf(env, filename, fn); put_utf(env, header, hd); put_utf(env, taxonomy, tx);
    if (rc) throw_for(env, rc, H_INDEXER(handle));
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_indexAddFasta(JNIEnv* env, jclass cls, jlong handle,
        jstring fastaPath, jstring taxdPath) {
    (void)cls;
    const char* f = get_utf(env, fastaPath);
    const char* t = get_utf(env, taxdPath);
    int rc = bsp_index_add_fasta(H_INDEXER(handle), f, t);
    put_utf(env, fastaPath, f); put_utf(env, taxdPath, t);
    if (rc) {                                              /* the host-side reader keeps its own message */
        jclass c = (*env)->FindClass(env, rc == BSP_ERR_ARG ? "java/lang/IllegalArgumentException" : "java/io/IOException");
        if (c) (*env)->ThrowNew(env, c, bsp_host_last_error());
    }
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_indexClose(JNIEnv* env, jclass cls, jlong handle) {
    (void)cls;
    int rc = bsp_index_close(H_INDEXER(handle));          /* frees the handle whatever the outcome */
    if (rc) throw_for(env, rc, NULL);
}

/* ------------------------------------------------------------------ Classifier */
JNIEXPORT jlong JNICALL Java_biospectra_gpu_BioSpectraGpu_classifierOpen(JNIEnv* env, jclass cls, jint kmerSize,
        jint kmerSkips, jboolean minStrand, jdouble minShouldMatch, jint queryAlgorithm, jint scoring, jint device,
        jstring indexPath) {
    (void)cls;
    bsp_config conf;
    bsp_config_default(&conf);
    conf.kmer_size = kmerSize; conf.kmer_skips = kmerSkips; conf.min_strand_kmer = minStrand ? 1 : 0;
    conf.query_term_min_should_match = minShouldMatch; conf.query_algorithm = queryAlgorithm;
    conf.scoring_algorithm = scoring; conf.device = device;
    const char* p = get_utf(env, indexPath);
    bsp_classifier* h = NULL;
    int rc = bsp_classifier_open(&conf, p, &h);
    put_utf(env, indexPath, p);
    if (rc) { throw_for(env, rc, NULL); return 0; }
    return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_classifyOne(JNIEnv* env, jclass cls, jlong handle,
        jbyteArray sequence, jintArray status, jintArray nHits, jintArray topDoc, jfloatArray topScore) {
    (void)cls;
    jsize len = sequence ? (*env)->GetArrayLength(env, sequence) : 0;
    jbyte* s = sequence ? (*env)->GetByteArrayElements(env, sequence, NULL) : NULL;
    jint st = BSP_READ_ERROR, nh = 0, td[BSP_TOPN];
    jfloat ts[BSP_TOPN];
    memset(td, 0, sizeof td); memset(ts, 0, sizeof ts);
    int rc = bsp_classify_one(H_CLASSIFIER(handle), (const char*)s, (int32_t)len, &st, NULL, &nh, NULL, td, ts);
    if (s) (*env)->ReleaseByteArrayElements(env, sequence, s, JNI_ABORT);
    if (rc) { throw_for(env, rc, H_CLASSIFIER(handle)); return; }       /* null / empty sequence -> IllegalArgumentException */
    (*env)->SetIntArrayRegion(env, status, 0, 1, &st);
    (*env)->SetIntArrayRegion(env, nHits, 0, 1, &nh);
    (*env)->SetIntArrayRegion(env, topDoc, 0, BSP_TOPN, td);
    (*env)->SetFloatArrayRegion(env, topScore, 0, BSP_TOPN, ts);
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_classifyBatch(JNIEnv* env, jclass cls, jlong handle,
        jobjectArray sequences, jintArray status, jintArray nHits, jintArray topDoc, jfloatArray topScore) {
    (void)cls;
    jsize n = (*env)->GetArrayLength(env, sequences);
    /* pack the reads back to back: bsp_classify_packed takes one byte buffer + n+1 offsets */
    int64_t* off = (int64_t*)malloc(((size_t)n + 1) * sizeof *off);
    if (!off) { throw_for(env, BSP_ERR_NOMEM, NULL); return; }
    off[0] = 0;
    for (jsize i = 0; i < n; i++) {
        jbyteArray s = (jbyteArray)(*env)->GetObjectArrayElement(env, sequences, i);
        off[i + 1] = off[i] + (s ? (*env)->GetArrayLength(env, s) : 0);        /* null read: empty -> status error */
        if (s) (*env)->DeleteLocalRef(env, s);
    }
    char* bytes = (char*)malloc(off[n] ? (size_t)off[n] : 1);
    if (!bytes) { free(off); throw_for(env, BSP_ERR_NOMEM, NULL); return; }
    for (jsize i = 0; i < n; i++) {
        jbyteArray s = (jbyteArray)(*env)->GetObjectArrayElement(env, sequences, i);
        if (s) {
            (*env)->GetByteArrayRegion(env, s, 0, (jsize)(off[i + 1] - off[i]), (jbyte*)bytes + off[i]);
            (*env)->DeleteLocalRef(env, s);
        }
    }
    jint* st = (*env)->GetIntArrayElements(env, status, NULL);
    jint* nh = (*env)->GetIntArrayElements(env, nHits, NULL);
    jint* td = (*env)->GetIntArrayElements(env, topDoc, NULL);
    jfloat* ts = (*env)->GetFloatArrayElements(env, topScore, NULL);
    int rc = (st && nh && td && ts)
           ? bsp_classify_packed(H_CLASSIFIER(handle), (int32_t)n, bytes, off, (int32_t*)st, NULL, (int32_t*)nh, NULL, (int32_t*)td, ts)
           : BSP_ERR_NOMEM;
    if (st) (*env)->ReleaseIntArrayElements(env, status, st, 0);
    if (nh) (*env)->ReleaseIntArrayElements(env, nHits, nh, 0);
    if (td) (*env)->ReleaseIntArrayElements(env, topDoc, td, 0);
    if (ts) (*env)->ReleaseFloatArrayElements(env, topScore, ts, 0);
    free(bytes); free(off);
    if (rc) throw_for(env, rc, H_CLASSIFIER(handle));
}

JNIEXPORT jobjectArray JNICALL Java_biospectra_gpu_BioSpectraGpu_docFields(JNIEnv* env, jclass cls, jlong handle, jint doc) {
    (void)cls;
    const char* f[4] = {NULL, NULL, NULL, NULL};
    int rc = bsp_doc_fields(H_CLASSIFIER(handle), doc, &f[0], &f[1], &f[2], &f[3]);
    if (rc) { throw_for(env, rc, H_CLASSIFIER(handle)); return NULL; }
    jclass sc = (*env)->FindClass(env, "java/lang/String");
    jobjectArray out = sc ? (*env)->NewObjectArray(env, 4, sc, NULL) : NULL;
    if (!out) return NULL;
    for (int i = 0; i < 4; i++) {
        jstring s = (*env)->NewStringUTF(env, f[i] ? f[i] : "");
        if (!s) return NULL;
        (*env)->SetObjectArrayElement(env, out, i, s);
        (*env)->DeleteLocalRef(env, s);
    }
    return out;
}

JNIEXPORT jlongArray JNICALL Java_biospectra_gpu_BioSpectraGpu_indexStats(JNIEnv* env, jclass cls, jlong handle) {
    (void)cls;
    int64_t v[4] = {0, 0, 0, 0};
    int rc = bsp_index_stats(H_CLASSIFIER(handle), &v[0], &v[1], &v[2], &v[3]);
    if (rc) { throw_for(env, rc, H_CLASSIFIER(handle)); return NULL; }
    jlongArray out = (*env)->NewLongArray(env, 4);
    if (!out) return NULL;
    jlong jv[4] = {(jlong)v[0], (jlong)v[1], (jlong)v[2], (jlong)v[3]};
    (*env)->SetLongArrayRegion(env, out, 0, 4, jv);
    return out;
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_setMicrobatch(JNIEnv* env, jclass cls, jlong handle,
        jint maxReads, jlong maxWaitMicros) {
    (void)cls;
    int rc = bsp_classifier_set_microbatch(H_CLASSIFIER(handle), maxReads, (int64_t)maxWaitMicros);
    if (rc) throw_for(env, rc, H_CLASSIFIER(handle));
}

JNIEXPORT void JNICALL Java_biospectra_gpu_BioSpectraGpu_classifierClose(JNIEnv* env, jclass cls, jlong handle) {
    (void)env; (void)cls;
    if (handle) bsp_classifier_close(H_CLASSIFIER(handle));
}

JNIEXPORT jstring JNICALL Java_biospectra_gpu_BioSpectraGpu_version(JNIEnv* env, jclass cls) {
    (void)cls;
    return (*env)->NewStringUTF(env, bsp_version());
}
