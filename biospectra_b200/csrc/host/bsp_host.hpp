// bsp_host.hpp -- C++ mirror of the reference's host-side interface for the hot path
// (see bsp_host.cpp for the file:line map onto the reference).
#pragma once
#include "../../../include/biospectra_gpu.h"
#include "json_mini.hpp"

#include <mutex>
#include <stdexcept>
#include <memory>
#include <string>
#include <vector>

namespace bsp {

struct IllegalArgument : public std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct IOError : public std::runtime_error { using std::runtime_error::runtime_error; };
struct FASTADataError : public IOError { using IOError::IOError; };

// src/biospectra/Configuration.java: same keys, same defaults
struct Configuration {
    std::string indexPath; bool hasIndexPath = false;
    int kmerSize = 10;                        // DEFAULT_KMERSIZE
    int kmerSkips = 5;                        // DEFAULT_KMERSKIPS
    bool minStrandKmer = false;
    double queryMinShouldMatch = 0.5;
    int workerThreads = 4;
    std::string scoringAlgorithm = "default";
    int queryAlgorithm = BSP_PAIRED_PROXIMITY;
    int ramBufferSizeForIndex = 16;
    // GPU placement: not part of configuration.json (the reference's Jackson mapper rejects unknown keys)
    int device = 0, denseTables = 0, positional = 0;

    static Configuration createInstance(const std::string& json);
    static Configuration createInstanceFromFile(const std::string& path);
    int scoringId() const;
    bsp_config toStruct() const;
};

struct FastaFileFilter { static bool accept(const std::string& name); };
struct CompressedFileFilter { static bool accept(const std::string& name); };
struct FastaFileHelper {
    static std::vector<std::string> findFastaDocs(const std::string& path);
    static std::string findTaxonHierarchyDoc(const std::string& fastaPath);
};

// The bytes of an input file: a read-only mapping of a regular file (no copy: the parser threads fault the page-cache pages
// in side by side), or an owned buffer (gzip input, pipes).
class FileImage {
public:
    FileImage() {}
    explicit FileImage(std::string&& owned) : owned_(std::move(owned)), p_(owned_.data()), n_(owned_.size()) {}
    FileImage(const char* mapped, size_t n) : p_(mapped), n_(n), mapped_(true) {}
    FileImage(const FileImage&) = delete;
    FileImage& operator=(const FileImage&) = delete;
    ~FileImage();
    const char* data() const { return p_; }
    size_t size() const { return n_; }
    char operator[](size_t i) const { return p_[i]; }
private:
    std::string owned_;
    const char* p_ = nullptr;
    size_t n_ = 0;
    bool mapped_ = false;
};

struct FASTAEntry { std::string headerLine, sequence; };
class FASTAReader {
public:
    explicit FASTAReader(const std::string& path);
    // a reader over bytes [begin, end) of another reader's file image (begin at the start of a line); the image is shared
    FASTAReader(const FASTAReader& whole, size_t begin, size_t end);
    bool readNext(FASTAEntry& e);      // false at EOF; throws FASTADataError
    // Offsets at which the file image can be cut into independently parsable stripes of about `target` bytes: every
    // cut is the start of a line that begins with '>' (such a line starts an entry wherever it stands).  First element 0,
    // last element = size of the image.
    std::vector<size_t> stripeCuts(size_t target) const;
    size_t imageSize() const { return data_->size(); }
private:
    bool readLine(std::string& line);
    std::shared_ptr<const FileImage> data_;
    size_t pos_ = 0, end_ = 0;
    std::string pending_;
    bool have_pending_ = false;
};

class Indexer {
public:
    explicit Indexer(const Configuration* conf);
    ~Indexer();
    void index(const std::vector<std::string>& fastaDocs);
    void index(const std::string& fastaDoc);
    void index(const std::string& fastaDoc, const std::string& taxonDoc);
    void close();
private:
    std::mutex mu_;
    bsp_indexer* h_ = nullptr;
};

struct Taxonomy { int taxid = 0; std::string name; bool hasName = false; int parent = 0; std::string rank; bool hasRank = false; };
struct TaxonTreeDescription {
    static std::vector<Taxonomy> parse(const std::string& json);
    static std::vector<Taxonomy> classifiable(const std::vector<Taxonomy>& tree);
};

struct SearchResultEntry {
    int docId = 0, rank = 0;
    std::string filename, header, sequenceDirection, taxonHierarchy;
    double score = 0;
};
struct ClassificationResult {
    std::string queryHeader, query;
    std::vector<SearchResultEntry> result;
    std::string type = "UNKNOWN", taxonRank = "unknown", taxonName;
    bool nameIsNull = false;
    std::string toJson() const;
};
void makeClassificationResult(ClassificationResult& r);

struct ClassificationResultSummary {
    std::string queryFilename;
    long long total = 0, unknown = 0, vague = 0, classified = 0, startTime = 0, endTime = 0;
    void report(const ClassificationResult& r);
    std::string toJson() const;
};

class Classifier {
public:
    explicit Classifier(const Configuration* conf);
    Classifier(bsp_classifier* adopted, int workerThreads);     // an open handle, left open by close()
    ~Classifier();
    // false when the reference would have thrown inside classify (read dropped by its callers)
    bool classify(const std::string& header, const std::string& sequence, ClassificationResult& out);
    void classifyBatch(const std::vector<std::string>& headers, const std::vector<std::string>& sequences,
                       std::vector<ClassificationResult>& out, std::vector<int>& status);
    void close();
    bsp_classifier* handle() { return h_; }

    // ---- batch path used by LocalClassifier (the reference's per-read Runnable, LocalClassifier.java:96-118, as a batch)
    // Stored fields of a document, prepared once: the JSON fragment of its SearchResultEntry and its ranked lineage
    struct DocInfo {
        std::string json;                 // "filename":..,"header":..,"sequence_direction":..,"taxon_hierarchy":..
        bool hasTaxonomy = false, taxonomyBad = false;
        std::vector<Taxonomy> ranked;     // TaxonTreeDescription#getClassifiableTaxonomyTree
    };
    struct PackedResults {                // per read, as bsp_classify_packed returns them
        std::vector<int32_t> status, nHits, topDoc;
        std::vector<float> topScore;
        std::vector<int32_t> taxLabel, taxIndex;   // bsp_classify_packed_tax: label with taxonomy, index into hit 0's ranked lineage
    };
    void classifyPacked(const char* seqBytes, const int64_t* offsets, int32_t n, PackedResults& out);
    // JSON line of read i (no trailing newline) appended to `o`; returns false when the reference would have thrown
    // (malformed taxonomy JSON); *type gets 0 UNKNOWN / 1 VAGUE / 2 CLASSIFIED
    bool appendResultJson(std::string& o, const char* header, size_t headerLen, const char* seq, size_t seqLen,
                          const PackedResults& r, int32_t i, int* type) const;
    int workerThreads() const { return workerThreads_; }
private:
    void initDocs();
    bsp_classifier* h_ = nullptr;
    bool owned_ = true;
    std::vector<DocInfo> docs_;
    int workerThreads_ = 4;
    bool deviceLca_ = true;               // labels with taxonomy computed on the device (BSP_HOST_LCA=1: on the host)
};

class LocalClassifier {
public:
    explicit LocalClassifier(const Configuration* conf);
    LocalClassifier(bsp_classifier* adopted, int workerThreads);
    bool classify(const std::string& header, const std::string& sequence, ClassificationResult& out) {
        return classifier_.classify(header, sequence, out);
    }
    void classify(const std::string& inputFasta, const std::string& classifyOutput, const std::string& summaryOutput);
    // the reference's own control flow: one task per read on `threads` pool threads, each calling
    // Classifier.classify(header, sequence) (LocalClassifier.java:88-120); the GPU sees micro-batches
    void classifyPerRead(const std::string& inputFasta, const std::string& classifyOutput, const std::string& summaryOutput,
                         int threads /* 0 = conf.worker_threads */);
    void close() { classifier_.close(); }
    const ClassificationResultSummary& lastSummary() const { return lastSummary_; }
private:
    Classifier classifier_;
    ClassificationResultSummary lastSummary_;
};

void runIndex(const std::string& jsonConfiguration, const std::string& indexDir, const std::string& referenceDir, int device);
void runClassifyLocal(const std::string& jsonConfiguration, const std::string& indexDir, const std::string& inputDir,
                      const std::string& outputDir, int device, int perReadThreads = -1);

}  // namespace bsp
