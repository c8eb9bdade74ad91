#!/usr/bin/env python
"""Writes jni/patches/Indexer.java.patch and Classifier.java.patch: the two reference classes with their Lucene bodies
replaced by calls into biospectra.gpu.BioSpectraGpu (jni/java/), as unified diffs against /root/reference/src.

The public surface of both classes is unchanged (constructors, index(...), classify(header, sequence), close());
`makeClassificationResult` (Classifier.java:263-339) and everything above the two classes stay as they are.
Run from the repo root where the reference is present:  python jni/patches/make_patches.py
Apply in a checkout of the reference:  patch -p0 < Indexer.java.patch ; patch -p0 < Classifier.java.patch
"""
import difflib
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(text, start_pat, end_pat, replacement, include_end=False):
    """Replace text[start of the line matching start_pat : line matching end_pat) by replacement."""
    a = re.search(start_pat, text, re.M)
    assert a, start_pat
    b = re.search(end_pat, text[a.end():], re.M)
    assert b, end_pat
    e = a.end() + (b.end() if include_end else b.start())
    return text[:a.start()] + replacement + text[e:]


def patch_indexer(src):
    s = src
    # imports: no Lucene, no analyzer
    for imp in ("biospectra.lucene.KmerIndexAnalyzer", "biospectra.utils.SequenceHelper", "java.util.Queue",
                "java.util.concurrent.ConcurrentLinkedQueue", "java.util.concurrent.TimeUnit", "biospectra.utils.BlockingExecutor"):
        s = s.replace("import %s;\n" % imp, "")
    s = re.sub(r"import org\.apache\.lucene\.[\w.]+;\n", "", s)
    s = s.replace("import biospectra.Configuration;\n", "import biospectra.Configuration;\nimport biospectra.gpu.BioSpectraGpu;\n")
    # fields
    s = s.replace("    private Analyzer analyzer;\n    private IndexWriter indexWriter;\n", "    private long handle;          // bsp_indexer* of libbiospectra_gpu.so\n")
    s = s.replace("    private BlockingExecutor executor;\n    private Queue<Document> freeQueue = new ConcurrentLinkedQueue<Document>();\n", "")
    s = s.replace("conf.getMinStrandKmer(), conf.getScoringAlgorithmObject(), conf.getWorkerThreads()",
                  "conf.getMinStrandKmer(), conf.getScoringAlgorithm(), conf.getWorkerThreads()")
    # initialize(): bsp_index_create creates and wipes index_path itself (Indexer.java:85-91)
    s = cut(s, r"^    private void initialize\(File indexPath", r"^    public synchronized void index\(List<File>",
            "    private void initialize(File indexPath, int kmerSize, boolean minStrandKmer, String scoringAlgorithm, int workerThreads, int ramBufferSize) throws Exception {\n"
            "        this.indexPath = indexPath;\n"
            "        this.minStrandKmer = minStrandKmer;\n"
            "        this.workerThreads = workerThreads;\n"
            "        // bsp_index_create: mkdirs + wipes index_path, like lines 85-91 did\n"
            "        this.handle = BioSpectraGpu.indexCreate(kmerSize, minStrandKmer, BioSpectraGpu.scoringCode(scoringAlgorithm),\n"
            "                workerThreads, ramBufferSize, BioSpectraGpu.device(), indexPath.getPath());\n"
            "    }\n"
            "    \n")
    # the per-entry worker: one native call, in reader order (canonical doc ids: forward 2e, reverse 2e+1)
    s = cut(s, r"^            final String f_filename = fastaDoc.getName\(\);", r"^        reader\.close\(\);",
            "            try {\n"
            "                // forward + reverse-complement documents (or one min_strand document) are made on the GPU at close()\n"
            "                BioSpectraGpu.indexAddEntry(this.handle, fastaDoc.getName(), headerLine, taxonTree, read.getSequence().getBytes(\"US-ASCII\"));\n"
            "            } catch (Exception ex) {\n"
            "                LOG.error(\"Exception occurred during index construction\", ex);\n"
            "            }\n"
            "        }\n"
            "        \n")
    # close()
    s = cut(s, r"^    public synchronized void close\(\) throws IOException \{", r"^    private void cleanUpDirectory",
            "    public synchronized void close() throws IOException {\n"
            "        if(this.handle != 0) {\n"
            "            long h = this.handle;\n"
            "            this.handle = 0;\n"
            "            BioSpectraGpu.indexClose(h);      // GPU build + atomic write of biospectra.idx\n"
            "        }\n"
            "    }\n"
            "\n")
    return s


def patch_classifier(src):
    s = src
    s = re.sub(r"import org\.apache\.lucene\.[\w.]+;\n", "", s)
    for imp in ("biospectra.index.IndexConstants", "biospectra.lucene.KmerQueryAnalyzer"):
        s = s.replace("import %s;\n" % imp, "")
    s = s.replace("import biospectra.Configuration;\n", "import biospectra.Configuration;\nimport biospectra.gpu.BioSpectraGpu;\n")
    s = s.replace("    private KmerQueryAnalyzer queryAnalyzer;\n    private IndexReader indexReader;\n    private IndexSearcher indexSearcher;\n",
                  "    private long handle;          // bsp_classifier* of libbiospectra_gpu.so\n")
    s = s.replace("conf.getQueryGenerationAlgorithm(), conf.getScoringAlgorithmObject());", "conf.getQueryGenerationAlgorithm(), conf.getScoringAlgorithm());")
    s = s.replace("QueryGenerationAlgorithm queryGenerationAlgorithm, Similarity similarity) throws Exception {",
                  "QueryGenerationAlgorithm queryGenerationAlgorithm, String scoringAlgorithm) throws Exception {")
    # initialize(): open the GPU index instead of DirectoryReader / IndexSearcher
    s = cut(s, r"^        this\.queryAnalyzer = new KmerQueryAnalyzer", r"^    private void createNaiveKmerQueryClauses",
            "        this.minShouldMatch = minShouldMatch;\n"
            "        this.queryGenerationAlgorithm = queryGenerationAlgorithm;\n"
            "        this.handle = BioSpectraGpu.classifierOpen(kmerSize, kmerSkips, minStrandKmer, minShouldMatch,\n"
            "                queryGenerationAlgorithm.ordinal(), BioSpectraGpu.scoringCode(scoringAlgorithm), BioSpectraGpu.device(), indexPath.getPath());\n"
            "    }\n"
            "\n")
    # the three clause builders, createQueryClauses and createQuery run inside the library now
    s = cut(s, r"^    private void createNaiveKmerQueryClauses", r"^    private ClassificationResult makeClassificationResult", "")
    # classify(): search + equal-to-max filter + stored fields come back from one native call
    s = cut(s, r"^        BooleanQuery q = createQuery", r"^        return classificationResult;",
            "        int[] status = new int[1], nHits = new int[1], topDoc = new int[BioSpectraGpu.TOPN];\n"
            "        float[] topScore = new float[BioSpectraGpu.TOPN];\n"
            "        // blocking, thread-safe; concurrent callers are coalesced into GPU micro-batches (bsp_classify_one)\n"
            "        BioSpectraGpu.classifyOne(this.handle, sequence.getBytes(\"US-ASCII\"), status, nHits, topDoc, topScore);\n"
            "        if(status[0] != BioSpectraGpu.READ_OK) {\n"
            "            // fewer than two k-mers or more than 10000 clauses: createQuery() threw here before (lines 223-227, 253-255)\n"
            "            throw new Exception(\"query cannot be built for the sequence\");\n"
            "        }\n"
            "        \n"
            "        if(nHits[0] > 0) {\n"
            "            List<SearchResultEntry> resultArr = new ArrayList<SearchResultEntry>();\n"
            "            for(int i=0;i<nHits[0];++i) {\n"
            "                // the hits whose score equals the maximum, rank = index in the top-10 list\n"
            "                String[] f = BioSpectraGpu.docFields(this.handle, topDoc[i]);\n"
            "                SearchResultEntry result = new SearchResultEntry();\n"
            "                result.setDocId(topDoc[i]);\n"
            "                result.setRank(i);\n"
            "                result.setFilename(f[0]);\n"
            "                result.setHeader(f[1]);\n"
            "                result.setSequenceDirection(f[2]);\n"
            "                result.setTaxonHierarchy(f[3]);\n"
            "                result.setScore(topScore[i]);\n"
            "                resultArr.add(result);\n"
            "            }\n"
            "            \n"
            "            classificationResult = makeClassificationResult(header, sequence, resultArr);\n"
            "        } else {\n"
            "            classificationResult = makeClassificationResult(header, sequence, null);\n"
            "        }\n"
            "        \n")
    s = s.replace("        this.queryAnalyzer.close();\n        this.indexReader.close();\n",
                  "        if(this.handle != 0) {\n            long h = this.handle;\n            this.handle = 0;\n            BioSpectraGpu.classifierClose(h);\n        }\n")
    return s


def main():
    for rel, fn in (("src/biospectra/index/Indexer.java", patch_indexer), ("src/biospectra/classify/Classifier.java", patch_classifier)):
        src = open(os.path.join(REF, rel)).read()
        dst = fn(src)
        assert "org.apache.lucene" not in dst and "indexWriter" not in dst and "indexSearcher" not in dst, rel
        diff = difflib.unified_diff(src.splitlines(True), dst.splitlines(True), rel, rel, n=2)
        out = os.path.join(HERE, os.path.basename(rel) + ".patch")
        with open(out, "w") as f:
            f.writelines(diff)
        print("wrote", out, len(dst.splitlines()), "lines after patching")
        with open(os.path.join("/tmp", os.path.basename(rel)), "w") as f:
            f.write(dst)


if __name__ == "__main__":
    main()
