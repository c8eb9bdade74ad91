/*
 * JNI binding of libbiospectra_gpu.so (include/biospectra_gpu.h) for BioSpectra.
 *
 * One native method per C-ABI entry point that the two patched reference classes need:
 *   biospectra.index.Indexer          (src/biospectra/index/Indexer.java:64-251)    -> index*
 *   biospectra.classify.Classifier    (src/biospectra/classify/Classifier.java:71-377) -> classif*
 * The C side is jni/biospectra_gpu_jni.c.  Nothing here touches Lucene.
 */
package biospectra.gpu;

import java.io.IOException;

public final class BioSpectraGpu {
    static {
        System.loadLibrary("biospectra_gpu_jni");          // links libbiospectra_gpu.so
    }

    private BioSpectraGpu() {
    }

    /** query_algorithm codes (biospectra.classify.QueryGenerationAlgorithm) */
    public static final int NAIVE_KMER = 0, CHAIN_PROXIMITY = 1, PAIRED_PROXIMITY = 2;
    /** scoring_algorithm codes (Configuration.getScoringAlgorithmObject, Configuration.java:169-179) */
    public static final int TFIDF = 0, BM25 = 1;
    /** per-read status */
    public static final int READ_OK = 0, READ_DROPPED = 1, READ_ERROR = 2;
    public static final int TOPN = 10;

    /** CUDA device of this JVM's handles: -Dbiospectra.gpu.device=N (configuration.json gets no new keys). */
    public static int device() {
        return Integer.getInteger("biospectra.gpu.device", 0);
    }

    /** "default" | "tfidf" | "vectorspace" | unknown -> TFIDF, "bm25" -> BM25 (Configuration.java:169-179) */
    public static int scoringCode(String scoringAlgorithm) {
        return scoringAlgorithm != null && scoringAlgorithm.equalsIgnoreCase("bm25") ? BM25 : TFIDF;
    }

    // ---- bsp_index_create / bsp_index_add_entry / bsp_index_close  (Indexer.java:64-129, 178-232, 239-251)
    /** @throws IllegalArgumentException like the Indexer constructor; wipes indexPath */
    public static native long indexCreate(int kmerSize, boolean minStrandKmer, int scoringAlgorithm, int workerThreads,
                                          int indexRamBuffer, int device, String indexPath);

    /** one FASTA entry: header without '>', whole .taxd text ("" if none), sequence as FASTAEntry.getSequence() */
    public static native void indexAddEntry(long handle, String filename, String header, String taxonomy, byte[] sequence);

    /** whole FASTA file through the library's own reader (FASTAReader#readNext semantics); taxdPath may be null */
    public static native void indexAddFasta(long handle, String fastaPath, String taxdPath) throws IOException;

    /** builds the index on the GPU, writes indexPath/biospectra.idx atomically, frees the handle */
    public static native void indexClose(long handle) throws IOException;

    // ---- bsp_classifier_open / bsp_classify_* / bsp_doc_fields / bsp_classifier_close  (Classifier.java:71-111, 341-377)
    /** @throws IllegalArgumentException like the Classifier constructor (incl. a missing index directory) */
    public static native long classifierOpen(int kmerSize, int kmerSkips, boolean minStrandKmer, double minShouldMatch,
                                             int queryAlgorithm, int scoringAlgorithm, int device, String indexPath);

    /**
     * bsp_classify_one: one read, blocking, safe from many threads at once (concurrent calls are coalesced into GPU
     * micro-batches).  Outputs: status[1], nHits[1], topDoc[10], topScore[10]; the first nHits[0] entries are the hits
     * whose score equals the maximum, rank = index (Classifier.java:358-366).
     */
    public static native void classifyOne(long handle, byte[] sequence, int[] status, int[] nHits, int[] topDoc, float[] topScore);

    /** bsp_classify_packed: n reads in one call; status/nHits are n long, topDoc/topScore n*10 */
    public static native void classifyBatch(long handle, byte[][] sequences, int[] status, int[] nHits, int[] topDoc, float[] topScore);

    /** bsp_doc_fields: {filename, header, sequence_direction, taxonomy} (IndexSearcher.doc, Classifier.java:361-363) */
    public static native String[] docFields(long handle, int doc);

    /** bsp_index_stats: {docs, terms, postings, positions} */
    public static native long[] indexStats(long handle);

    /** bsp_classifier_set_microbatch: window of classifyOne (reads per GPU batch, microseconds the first caller waits) */
    public static native void setMicrobatch(long handle, int maxReads, long maxWaitMicros);

    public static native void classifierClose(long handle);

    public static native String version();
}
