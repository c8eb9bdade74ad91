// bsp_host.cpp -- host-side mirror of the reference's operator interface for the hot path,
// written in C++ because no JVM exists in this image (INTEGRATION.md shows the JNI stub a
// maintainer would use instead).  Same names, argument meaning and error behaviour as:
//   biospectra.Configuration            (src/biospectra/Configuration.java)
//   biospectra.index.Indexer            (src/biospectra/index/Indexer.java)
//   biospectra.classify.Classifier      (src/biospectra/classify/Classifier.java)
//   biospectra.classify.LocalClassifier (src/biospectra/classify/LocalClassifier.java)
//   biospectra.utils.FastaFileHelper / FastaFileReader / FastaFileFilter / CompressedFileFilter
//   org.yeastrc.fasta.FASTAReader       (libs/fasta-utils-1.1.jar)
//   classify.beans.*                    (result / summary JSON, lowest-common-taxon labelling)
// All k-mer / scoring work goes through the C ABI (include/biospectra_gpu.h).
#include "bsp_host.hpp"

#include <algorithm>
#include <chrono>
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

namespace bsp {

// ------------------------------------------------------------------ small utils
static std::string lower(const std::string& s) {
    std::string o = s;
    for (auto& c : o) c = (char)tolower((unsigned char)c);
    return o;
}
static bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
static bool file_exists(const std::string& p) { struct stat sb; return stat(p.c_str(), &sb) == 0; }
static bool is_directory(const std::string& p) { struct stat sb; return stat(p.c_str(), &sb) == 0 && S_ISDIR(sb.st_mode); }
static std::string base_name(const std::string& p) {
    size_t i = p.find_last_of('/');
    return i == std::string::npos ? p : p.substr(i + 1);
}
static std::string parent_dir(const std::string& p) {
    size_t i = p.find_last_of('/');
    return i == std::string::npos ? "" : p.substr(0, i);
}
static void make_dirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0777); }
        if (i < path.size()) cur.push_back(path[i]);
    }
}
static long long now_ms() {
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
static std::string read_text_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw IOError("cannot open " + path);
    std::string s;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

// ------------------------------------------------------------------ Configuration
static const char* CONF_KEYS[] = {"index_path", "kmer_size", "kmer_skips", "min_strand_kmer", "query_term_min_should_match",
                                  "worker_threads", "index_ram_buffer", "query_algorithm", "scoring_algorithm"};

Configuration Configuration::createInstance(const std::string& json) {
    if (json.empty()) throw IllegalArgument("json is empty or null");                   // Configuration.java:66-68
    jsonm::ValuePtr v = jsonm::parse(json);
    if (v->type != jsonm::Value::Object) throw IOError("configuration is not a JSON object");
    Configuration c;
    for (auto& kv : v->obj) {
        bool known = false;
        for (const char* k : CONF_KEYS) if (kv.first == k) known = true;
        if (!known) throw IOError("Unrecognized field \"" + kv.first + "\"");       // Jackson FAIL_ON_UNKNOWN_PROPERTIES
        const jsonm::Value& x = *kv.second;
        if (kv.first == "index_path") { c.indexPath = x.str; c.hasIndexPath = x.type == jsonm::Value::String; }
        else if (kv.first == "kmer_size") c.kmerSize = (int)x.inum;
        else if (kv.first == "kmer_skips") c.kmerSkips = (int)x.inum;
        else if (kv.first == "min_strand_kmer") c.minStrandKmer = x.type == jsonm::Value::Bool ? x.b : x.inum != 0;
        else if (kv.first == "query_term_min_should_match") c.queryMinShouldMatch = x.num;
        else if (kv.first == "worker_threads") c.workerThreads = (int)x.inum;
        else if (kv.first == "index_ram_buffer") c.ramBufferSizeForIndex = (int)x.inum;
        else if (kv.first == "scoring_algorithm") c.scoringAlgorithm = x.str;
        else if (kv.first == "query_algorithm") {
            if (x.str == "NAIVE_KMER") c.queryAlgorithm = BSP_NAIVE_KMER;
            else if (x.str == "CHAIN_PROXIMITY") c.queryAlgorithm = BSP_CHAIN_PROXIMITY;
            else if (x.str == "PAIRED_PROXIMITY") c.queryAlgorithm = BSP_PAIRED_PROXIMITY;
            else throw IOError("unknown query_algorithm \"" + x.str + "\"");         // enum deserialisation failure
        }
    }
    return c;
}

Configuration Configuration::createInstanceFromFile(const std::string& path) {
    return createInstance(read_text_file(path));
}

int Configuration::scoringId() const {               // Configuration.java:169-179
    return lower(scoringAlgorithm) == "bm25" ? BSP_BM25 : BSP_TFIDF;
}

bsp_config Configuration::toStruct() const {
    bsp_config c;
    bsp_config_default(&c);
    c.kmer_size = kmerSize; c.kmer_skips = kmerSkips; c.min_strand_kmer = minStrandKmer ? 1 : 0;
    c.query_algorithm = queryAlgorithm; c.scoring_algorithm = scoringId();
    c.worker_threads = workerThreads; c.index_ram_buffer = ramBufferSizeForIndex;
    c.query_term_min_should_match = queryMinShouldMatch;
    c.device = device; c.part_rank = 0; c.part_count = 1; c.dense_tables = denseTables; c.positional = positional;
    return c;
}

// ------------------------------------------------------------------ FASTA (A.1)
bool FastaFileFilter::accept(const std::string& name) {          // utils/FastaFileFilter.java:29-38
    static const char* EXT[] = {".fa", ".fa.gz", ".ffn.gz", ".ffn", ".fna.gz", ".fna", ".fasta", ".fasta.gz"};
    std::string l = lower(name);
    for (const char* e : EXT) if (ends_with(l, e)) return true;
    return false;
}
bool CompressedFileFilter::accept(const std::string& name) {     // utils/CompressedFileFilter.java:29-32
    std::string l = lower(name);
    return ends_with(l, ".gz") || ends_with(l, ".zip");
}

// utils/FastaFileHelper.java:33-60; the reference's File.listFiles() order is unspecified,
// this build sorts by path (canonical doc ids)
static void find_fasta_rec(const std::string& path, std::vector<std::string>& out) {
    if (!is_directory(path)) {
        if (FastaFileFilter::accept(base_name(path))) out.push_back(path);
        return;
    }
    DIR* d = opendir(path.c_str());
    if (!d) throw IOError("path " + path + " not exist");
    std::vector<std::string> names;
    struct dirent* de;
    while ((de = readdir(d)) != nullptr) {
        if (!strcmp(de->d_name, ".") || !strcmp(de->d_name, "..")) continue;
        names.push_back(de->d_name);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    for (auto& n : names) {
        std::string p = path + "/" + n;
        if (is_directory(p)) find_fasta_rec(p, out);
        else if (FastaFileFilter::accept(n)) out.push_back(p);
    }
}
std::vector<std::string> FastaFileHelper::findFastaDocs(const std::string& path) {
    if (!file_exists(path)) throw IOError("path " + path + " not exist");
    std::vector<std::string> out;
    find_fasta_rec(path, out);
    return out;
}
std::string FastaFileHelper::findTaxonHierarchyDoc(const std::string& fasta) { return fasta + ".taxd"; }   // :66-68

FASTAReader::FASTAReader(const std::string& path) {
    // utils/FastaFileReader.java:29-39: gzip iff the lower-cased name ends .gz/.zip
    if (CompressedFileFilter::accept(base_name(path))) {
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) throw IOError("cannot open " + path);
        gzbuffer(g, 1 << 20);
        char buf[1 << 16];
        int n;
        while ((n = gzread(g, buf, sizeof buf)) > 0) data_.append(buf, (size_t)n);
        int err = 0;
        gzerror(g, &err);
        gzclose(g);
        if (n < 0 || (err != Z_OK && err != Z_STREAM_END)) throw IOError("gzip error reading " + path);
    } else {
        data_ = read_text_file(path);
    }
}

bool FASTAReader::readLine(std::string& line) {      // BufferedReader#readLine: \n, \r\n, \r
    if (pos_ >= data_.size()) return false;
    size_t st = pos_;
    while (pos_ < data_.size() && data_[pos_] != '\n' && data_[pos_] != '\r') pos_++;
    line.assign(data_, st, pos_ - st);
    if (pos_ < data_.size()) {
        if (data_[pos_] == '\r' && pos_ + 1 < data_.size() && data_[pos_ + 1] == '\n') pos_ += 2;
        else pos_++;
    }
    return true;
}

bool FASTAReader::readNext(FASTAEntry& e) {
    std::string line;
    if (!have_pending_) {
        if (!readLine(line)) return false;
    } else { line = pending_; have_pending_ = false; }
    if (line.empty() || line[0] != '>')
        throw FASTADataError("FASTADataErrorException: expected a header line starting with '>'");
    e.headerLine = line;
    e.sequence.clear();
    bool any = false;
    while (readLine(line)) {
        if (!line.empty() && line[0] == '>') { pending_ = line; have_pending_ = true; break; }
        if (!line.empty() && line[0] == ';') continue;
        any = true;
        for (char& c : line) {
            unsigned char u = (unsigned char)c;
            if (u >= 0x80) c = '?';                  // US-ASCII decoding: undecodable byte
            else c = (char)toupper(u);               // String#toUpperCase
        }
        e.sequence += line;
    }
    if (!any) throw FASTADataError("FASTADataErrorException: header without sequence");
    // String#trim: strip chars <= ' ' at both ends
    size_t a = 0, b = e.sequence.size();
    while (a < b && (unsigned char)e.sequence[a] <= ' ') a++;
    while (b > a && (unsigned char)e.sequence[b - 1] <= ' ') b--;
    if (a > 0 || b < e.sequence.size()) e.sequence = e.sequence.substr(a, b - a);
    return true;
}

// ------------------------------------------------------------------ Indexer
static std::string last_err() { const char* m = bsp_last_error(nullptr); return m ? m : ""; }
static void check_rc(int rc) {
    if (rc == BSP_OK) return;
    if (rc == BSP_ERR_ARG) throw IllegalArgument(last_err());
    if (rc == BSP_ERR_IO) throw IOError(last_err());
    throw std::runtime_error(last_err());
}

Indexer::Indexer(const Configuration* conf) {
    if (!conf) throw IllegalArgument("conf is null");                                  // Indexer.java:65-67
    if (!conf->hasIndexPath) throw IllegalArgument("indexPath is null");               // :69-71
    bsp_config c = conf->toStruct();
    check_rc(bsp_index_create(&c, conf->indexPath.c_str(), &h_));                      // validates k, threads; wipes dir
}
Indexer::~Indexer() { try { close(); } catch (...) {} }

void Indexer::index(const std::vector<std::string>& fastaDocs) {                       // Indexer.java:131-139
    for (auto& f : fastaDocs) index(f);
}
void Indexer::index(const std::string& fastaDoc) {                                     // :141-148
    index(fastaDoc, FastaFileHelper::findTaxonHierarchyDoc(fastaDoc));
}
void Indexer::index(const std::string& fastaDoc, const std::string& taxonDoc) {        // :150-236
    std::lock_guard<std::mutex> lk(mu_);
    if (!h_) throw std::runtime_error("indexer is closed");
    std::string taxonTree;
    if (!taxonDoc.empty() && file_exists(taxonDoc)) taxonTree = read_text_file(taxonDoc);   // :157-161
    FASTAReader reader(fastaDoc);
    FASTAEntry read;
    std::string filename = base_name(fastaDoc);                                        // fastaDoc.getName()
    while (reader.readNext(read)) {
        std::string header = read.headerLine;
        if (!header.empty() && header[0] == '>') header = header.substr(1);            // :167-170
        int rc = bsp_index_add_entry(h_, filename.c_str(), header.c_str(), taxonTree.c_str(), read.sequence.data(),
                                     (int64_t)read.sequence.size());
        if (rc != BSP_OK)                                                              // :227-229 log and continue
            fprintf(stderr, "ERROR Exception occurred during index construction: %s\n", last_err().c_str());
    }
}
void Indexer::close() {                                                                // :239-251
    std::lock_guard<std::mutex> lk(mu_);
    if (!h_) return;
    bsp_indexer* h = h_;
    h_ = nullptr;
    check_rc(bsp_index_close(h));
}

// ------------------------------------------------------------------ taxonomy beans
std::vector<Taxonomy> TaxonTreeDescription::parse(const std::string& json) {           // TaxonTreeDescription.java:36-53
    if (json.empty()) throw IllegalArgument("json is empty or null");
    jsonm::ValuePtr v;
    const jsonm::Value* tree = nullptr;
    if (json.front() == '[' && json.back() == ']') { v = jsonm::parse(json); tree = v.get(); }
    else if (json.front() == '{' && json.back() == '}') { v = jsonm::parse(json); tree = v->get("tree"); }
    else throw IOError("cannot create instance from wrong json format");
    std::vector<Taxonomy> out;
    if (!tree || tree->type != jsonm::Value::Array) return out;
    for (auto& t : tree->arr) {
        Taxonomy x;
        if (const jsonm::Value* a = t->get("taxid")) x.taxid = (int)a->inum;
        if (const jsonm::Value* a = t->get("name")) { x.name = a->str; x.hasName = a->type == jsonm::Value::String; }
        if (const jsonm::Value* a = t->get("parent")) x.parent = (int)a->inum;
        if (const jsonm::Value* a = t->get("rank")) { x.rank = a->str; x.hasRank = a->type == jsonm::Value::String; }
        out.push_back(x);
    }
    return out;
}
std::vector<Taxonomy> TaxonTreeDescription::classifiable(const std::vector<Taxonomy>& tree) {   // :85-98
    std::vector<Taxonomy> out;
    for (auto& t : tree)
        if (t.hasRank && !t.rank.empty() && lower(t.rank) != "no rank") out.push_back(t);
    return out;
}

// classify/Classifier.java:263-339
void makeClassificationResult(ClassificationResult& r) {
    auto& res = r.result;
    if (res.empty()) { r.type = "UNKNOWN"; r.taxonRank = "unknown"; r.taxonName = ""; return; }
    if (res.size() == 1) {
        r.type = "CLASSIFIED"; r.taxonRank = "unknown"; r.taxonName = "";
        if (!res[0].taxonHierarchy.empty()) {
            auto ranked = TaxonTreeDescription::classifiable(TaxonTreeDescription::parse(res[0].taxonHierarchy));
            if (!ranked.empty()) { r.taxonRank = ranked[0].rank; r.taxonName = ranked[0].name; r.nameIsNull = !ranked[0].hasName; }
        }
        return;
    }
    std::vector<std::vector<Taxonomy>> trees;
    for (auto& e : res) {
        if (e.taxonHierarchy.empty()) { r.type = "VAGUE"; r.taxonRank = "unknown"; r.taxonName = ""; return; }   // quick fail
        trees.push_back(TaxonTreeDescription::classifiable(TaxonTreeDescription::parse(e.taxonHierarchy)));
    }
    for (auto& tax : trees[0]) {
        bool common = true;
        for (size_t j = 1; j < trees.size() && common; j++) {
            bool found = false;
            for (auto& t2 : trees[j]) if (t2.taxid == tax.taxid) { found = true; break; }
            if (!found) common = false;
        }
        if (common) { r.type = "CLASSIFIED"; r.taxonRank = tax.rank; r.taxonName = tax.name; r.nameIsNull = !tax.hasName; return; }
    }
    r.type = "VAGUE"; r.taxonRank = "unknown"; r.taxonName = "";
}

std::string ClassificationResult::toJson() const {   // beans/ClassificationResult.java:59-117, SearchResultEntry.java:53-121
    std::string o;
    o.reserve(256 + query.size());
    o += "{\"query_header\":"; jsonm::write_string(o, queryHeader);
    o += ",\"query\":"; jsonm::write_string(o, query);
    o += ",\"result\":[";
    for (size_t i = 0; i < result.size(); i++) {
        const SearchResultEntry& e = result[i];
        if (i) o += ",";
        o += "{\"docid\":" + std::to_string(e.docId) + ",\"rank\":" + std::to_string(e.rank);
        o += ",\"filename\":"; jsonm::write_string(o, e.filename);
        o += ",\"header\":"; jsonm::write_string(o, e.header);
        o += ",\"sequence_direction\":"; jsonm::write_string(o, e.sequenceDirection);
        o += ",\"taxon_hierarchy\":"; jsonm::write_string(o, e.taxonHierarchy);
        o += ",\"score\":" + jsonm::java_double(e.score) + "}";
    }
    o += "],\"type\":\"" + type + "\",\"taxon_rank\":";
    jsonm::write_string(o, taxonRank);
    o += ",\"taxon_name\":";
    if (nameIsNull) o += "null"; else jsonm::write_string(o, taxonName);
    o += "}";
    return o;
}

std::string ClassificationResultSummary::toJson() const {   // beans/ClassificationResultSummary.java:42-133
    std::string o = "{\"query_filename\":";
    jsonm::write_string(o, queryFilename);
    o += ",\"start_time\":" + std::to_string(startTime) + ",\"end_time\":" + std::to_string(endTime);
    o += ",\"total\":" + std::to_string(total) + ",\"unknown\":" + std::to_string(unknown) + ",\"vague\":" + std::to_string(vague);
    o += ",\"classified\":" + std::to_string(classified) + "}";
    return o;
}
void ClassificationResultSummary::report(const ClassificationResult& r) {   // :118-133
    total++;
    if (r.type == "VAGUE") vague++;
    else if (r.type == "CLASSIFIED") classified++;
    else unknown++;
}

// ------------------------------------------------------------------ Classifier
Classifier::Classifier(const Configuration* conf) {
    if (!conf) throw IllegalArgument("conf is null");                                  // Classifier.java:72-74
    if (!conf->hasIndexPath) throw IllegalArgument("indexPath is null");               // :76-78
    bsp_config c = conf->toStruct();
    check_rc(bsp_classifier_open(&c, conf->indexPath.c_str(), &h_));                   // validates k, skips, dir
}
Classifier::~Classifier() { close(); }
void Classifier::close() {
    if (h_) { bsp_classifier_close(h_); h_ = nullptr; }
}

void Classifier::classifyBatch(const std::vector<std::string>& headers, const std::vector<std::string>& sequences,
                               std::vector<ClassificationResult>& out, std::vector<int>& status) {
    const int32_t n = (int32_t)sequences.size();
    out.assign(n, ClassificationResult());
    status.assign(n, BSP_READ_ERROR);
    std::vector<int64_t> off((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; i++) off[i + 1] = off[i] + (int64_t)sequences[i].size();
    std::string bytes;
    bytes.reserve((size_t)off[n]);
    for (auto& s : sequences) bytes += s;
    std::vector<int32_t> st(n), label(n), nhits(n), ntop(n), tdoc((size_t)n * BSP_TOPN);
    std::vector<float> tscore((size_t)n * BSP_TOPN);
    check_rc(bsp_classify_packed(h_, n, bytes.data(), off.data(), st.data(), label.data(), nhits.data(), ntop.data(), tdoc.data(),
                                 tscore.data()));
    for (int32_t i = 0; i < n; i++) {
        status[i] = st[i];
        if (sequences[i].empty()) status[i] = BSP_READ_ERROR;       // Classifier.java:342-344 IllegalArgumentException
        if (status[i] != BSP_READ_OK) continue;
        ClassificationResult& r = out[i];
        r.queryHeader = headers[i];
        r.query = sequences[i];
        for (int j = 0; j < nhits[i]; j++) {                        // Classifier.java:358-366: hits equal to the max
            SearchResultEntry e;
            e.docId = tdoc[(size_t)i * BSP_TOPN + j];
            e.rank = j;
            e.score = (double)tscore[(size_t)i * BSP_TOPN + j];
            const char *fn, *hd, *dir, *tax;
            check_rc(bsp_doc_fields(h_, e.docId, &fn, &hd, &dir, &tax));
            e.filename = fn; e.header = hd; e.sequenceDirection = dir; e.taxonHierarchy = tax;
            r.result.push_back(std::move(e));
        }
        try { makeClassificationResult(r); }
        catch (const std::exception& ex) { status[i] = BSP_READ_ERROR; fprintf(stderr, "ERROR Exception occurred during search: %s\n", ex.what()); }
    }
}

bool Classifier::classify(const std::string& header, const std::string& sequence, ClassificationResult& out) {
    if (sequence.empty()) throw IllegalArgument("sequence is null or empty");           // Classifier.java:342-344
    std::vector<ClassificationResult> r;
    std::vector<int> st;
    classifyBatch({header}, {sequence}, r, st);
    if (st[0] != BSP_READ_OK) return false;      // the reference throws (NPE / TooManyClauses); callers log + skip
    out = std::move(r[0]);
    return true;
}

// ------------------------------------------------------------------ LocalClassifier
LocalClassifier::LocalClassifier(const Configuration* conf) : classifier_(conf) {}

void LocalClassifier::classify(const std::string& inputFasta, const std::string& classifyOutput, const std::string& summaryOutput) {
    if (inputFasta.empty()) throw IllegalArgument("inputFasta is null");               // LocalClassifier.java:61-63
    if (classifyOutput.empty()) throw IllegalArgument("classifyOutput is null");       // :65-67
    std::string pd = parent_dir(classifyOutput);
    if (!pd.empty() && !file_exists(pd)) make_dirs(pd);                                // :69-71
    if (!summaryOutput.empty()) { std::string sd = parent_dir(summaryOutput); if (!sd.empty() && !file_exists(sd)) make_dirs(sd); }
    FASTAReader reader(inputFasta);
    FILE* fw = fopen(classifyOutput.c_str(), "wb");
    if (!fw) throw IOError("cannot create " + classifyOutput);
    std::vector<char> wbuf(1 << 20);
    setvbuf(fw, wbuf.data(), _IOFBF, wbuf.size());
    ClassificationResultSummary summary;
    summary.queryFilename = base_name(inputFasta);
    summary.startTime = now_ms();
    const size_t BATCH_READS = 262144, BATCH_BYTES = (size_t)128 << 20;
    std::vector<std::string> headers, seqs;
    size_t bytes = 0;
    auto flush = [&]() {
        if (seqs.empty()) return;
        std::vector<ClassificationResult> res;
        std::vector<int> st;
        // reads with an empty sequence throw in Classifier.classify and are skipped (LocalClassifier.java:112-114)
        classifier_.classifyBatch(headers, seqs, res, st);
        for (size_t i = 0; i < res.size(); i++) {
            if (st[i] != BSP_READ_OK) continue;
            std::string js = res[i].toJson();
            summary.report(res[i]);
            fwrite(js.data(), 1, js.size(), fw);
            fputc('\n', fw);
        }
        headers.clear(); seqs.clear(); bytes = 0;
    };
    try {
        FASTAEntry read;
        while (reader.readNext(read)) {
            bytes += read.sequence.size();
            headers.push_back(read.headerLine);                                        // header INCLUDING '>' (:94)
            seqs.push_back(std::move(read.sequence));
            if (seqs.size() >= BATCH_READS || bytes >= BATCH_BYTES) flush();
        }
        flush();
    } catch (...) { fclose(fw); throw; }
    fclose(fw);
    summary.endTime = now_ms();
    if (!summaryOutput.empty()) {                                                      // :128-130
        FILE* fs = fopen(summaryOutput.c_str(), "wb");
        if (!fs) throw IOError("cannot create " + summaryOutput);
        std::string js = summary.toJson();
        fwrite(js.data(), 1, js.size(), fs);
        fclose(fs);
    }
    lastSummary_ = summary;
}

// ------------------------------------------------------------------ drivers (BioSpectra.java:110-161)
void runIndex(const std::string& jsonConfiguration, const std::string& indexDir, const std::string& referenceDir, int device) {
    Configuration conf;
    if (!jsonConfiguration.empty()) conf = Configuration::createInstanceFromFile(jsonConfiguration);
    else { conf.indexPath = indexDir; conf.hasIndexPath = !indexDir.empty(); }
    conf.device = device;
    Indexer indexer(&conf);
    for (auto& f : FastaFileHelper::findFastaDocs(referenceDir)) {
        long long t0 = now_ms();
        indexer.index(f);
        fprintf(stderr, "INFO indexing %s finished - %lld milliseconds\n", f.c_str(), now_ms() - t0);
    }
    indexer.close();
}

void runClassifyLocal(const std::string& jsonConfiguration, const std::string& indexDir, const std::string& inputDir,
                      const std::string& outputDir, int device) {
    Configuration conf;
    if (!jsonConfiguration.empty()) conf = Configuration::createInstanceFromFile(jsonConfiguration);
    else { conf.indexPath = indexDir; conf.hasIndexPath = !indexDir.empty(); }
    conf.device = device;
    LocalClassifier classifier(&conf);
    if (!file_exists(outputDir)) make_dirs(outputDir);
    for (auto& f : FastaFileHelper::findFastaDocs(inputDir)) {
        std::string name = base_name(f);
        classifier.classify(f, outputDir + "/" + name + ".result", outputDir + "/" + name + ".result.sum");
        fprintf(stderr, "INFO classifying %s finished in %lld millisec\n", name.c_str(),
                classifier.lastSummary().endTime - classifier.lastSummary().startTime);
    }
    classifier.close();
}

}  // namespace bsp

// ------------------------------------------------------------------ C entry points
#define HOST_API extern "C" __attribute__((visibility("default")))
static thread_local std::string g_host_error;

HOST_API const char* bsp_host_last_error(void) { return g_host_error.c_str(); }

// BioSpectra.index (BioSpectra.java:110-134): json_conf_or_null XOR index_dir
HOST_API int bsp_host_index(const char* json_conf, const char* index_dir, const char* reference_dir, int32_t device) {
    try {
        bsp::runIndex(json_conf ? json_conf : "", index_dir ? index_dir : "", reference_dir ? reference_dir : "", device);
        return BSP_OK;
    } catch (const bsp::IllegalArgument& e) { g_host_error = e.what(); return BSP_ERR_ARG; }
    catch (const bsp::IOError& e) { g_host_error = e.what(); return BSP_ERR_IO; }
    catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_STATE; }
}

// BioSpectra.classifyLocal (BioSpectra.java:136-161)
HOST_API int bsp_host_classify_local(const char* json_conf, const char* index_dir, const char* input_dir, const char* output_dir,
                                     int32_t device) {
    try {
        bsp::runClassifyLocal(json_conf ? json_conf : "", index_dir ? index_dir : "", input_dir ? input_dir : "",
                              output_dir ? output_dir : "", device);
        return BSP_OK;
    } catch (const bsp::IllegalArgument& e) { g_host_error = e.what(); return BSP_ERR_ARG; }
    catch (const bsp::IOError& e) { g_host_error = e.what(); return BSP_ERR_IO; }
    catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_STATE; }
}

// Pure host helpers exported for CPU-side tests (no GPU needed)
HOST_API int bsp_host_label_json(const char* const* taxon_hierarchies, int32_t n, char* out, int32_t cap) {
    try {
        bsp::ClassificationResult r;
        for (int32_t i = 0; i < n; i++) { bsp::SearchResultEntry e; e.taxonHierarchy = taxon_hierarchies[i]; e.rank = i; r.result.push_back(e); }
        bsp::makeClassificationResult(r);
        std::string js = r.toJson();
        if ((int32_t)js.size() + 1 > cap) return BSP_ERR_ARG;
        memcpy(out, js.c_str(), js.size() + 1);
        return BSP_OK;
    } catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_STATE; }
}

// parse a FASTA file with A.1 semantics; writes "header\tsequence\n" records
HOST_API int bsp_host_read_fasta(const char* path, char* out, int64_t cap, int64_t* needed) {
    try {
        bsp::FASTAReader rd(path);
        bsp::FASTAEntry e;
        std::string o;
        while (rd.readNext(e)) { o += e.headerLine; o.push_back('\t'); o += e.sequence; o.push_back('\n'); }
        if (needed) *needed = (int64_t)o.size() + 1;
        if ((int64_t)o.size() + 1 > cap) return BSP_ERR_NOMEM;
        memcpy(out, o.c_str(), o.size() + 1);
        return BSP_OK;
    } catch (const bsp::FASTADataError& e) { g_host_error = e.what(); return BSP_ERR_IO; }
    catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_IO; }
}

HOST_API int bsp_host_double_to_string(double d, char* out, int32_t cap) {
    std::string s = jsonm::java_double(d);
    if ((int32_t)s.size() + 1 > cap) return BSP_ERR_ARG;
    memcpy(out, s.c_str(), s.size() + 1);
    return BSP_OK;
}

HOST_API int bsp_host_parse_config(const char* json, bsp_config* out, char* index_path, int32_t cap) {
    try {
        bsp::Configuration c = bsp::Configuration::createInstance(json ? json : "");
        *out = c.toStruct();
        if (index_path && cap > 0) { strncpy(index_path, c.indexPath.c_str(), (size_t)cap - 1); index_path[cap - 1] = 0; }
        return BSP_OK;
    } catch (const bsp::IllegalArgument& e) { g_host_error = e.what(); return BSP_ERR_ARG; }
    catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_IO; }
}
