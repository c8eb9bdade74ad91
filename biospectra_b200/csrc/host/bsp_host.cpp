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
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <condition_variable>
#include <deque>
#include <atomic>
#include <thread>

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
static long long env_ll(const char* name, long long dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoll(v) : dflt;
}
static long long now_ms() {
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
static std::string read_text_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw IOError("cannot open " + path);
    std::string s;
    struct stat sb;
    if (fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
        // a regular file: one allocation, large unbuffered reads straight into it
        s.resize((size_t)sb.st_size);
        const int fd = fileno(f);
        const size_t total = s.size();
        auto read_range = [&](size_t a, size_t b) {           // pread until [a, b) is filled or the file ends
            while (a < b) {
                const ssize_t n = pread(fd, &s[a], std::min<size_t>(b - a, (size_t)32 << 20), (off_t)a);
                if (n <= 0) return a;
                a += (size_t)n;
            }
            return a;
        };
        size_t got = total;
        if (total >= ((size_t)32 << 20)) {                    // a large file: four ranges side by side (page-cache copies overlap)
            const int R = 4;
            size_t ends[R];
            std::vector<std::thread> th;
            for (int i = 0; i < R; i++) th.emplace_back([&, i] { ends[i] = read_range(total * i / R, total * (i + 1) / R); });
            for (auto& t : th) t.join();
            for (int i = R - 1; i >= 0; i--) if (ends[i] < total * (i + 1) / R) got = ends[i];      // first short range ends the data
        } else got = read_range(0, total);
        s.resize(got);
        if (got == total) lseek(fd, (off_t)total, SEEK_SET);  // the buffered reads below continue behind it
        else { fclose(f); return s; }
    }
    char buf[1 << 16];                                   // whatever follows (the file grew, or it is not a regular file)
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

FASTAReader::FASTAReader(const FASTAReader& whole, size_t begin, size_t end)
    : data_(whole.data_), pos_(std::min(begin, whole.data_->size())), end_(std::min(end, whole.data_->size())) {}

std::vector<size_t> FASTAReader::stripeCuts(size_t target) const {
    const FileImage& d = *data_;
    std::vector<size_t> cuts{0};
    if (target == 0) target = 1;
    size_t at = target;
    while (at < d.size()) {
        // the next '>' that starts a line (after \n or \r: BufferedReader#readLine ends a line at either)
        const char* q = (const char*)memchr(d.data() + at, '>', d.size() - at);
        if (!q) break;
        const size_t p = (size_t)(q - d.data());
        if (p > 0 && (d[p - 1] == '\n' || d[p - 1] == '\r')) { cuts.push_back(p); at = p + target; }
        else at = p + 1;
    }
    cuts.push_back(d.size());
    return cuts;
}

FASTAReader::FASTAReader(const std::string& path) {
    std::string data_local;
    std::string& data_ = data_local;
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
        // a regular file is mapped, not copied (BSP_FASTA_MMAP=0: read into a buffer like everything else)
        const int fd = env_ll("BSP_FASTA_MMAP", 1) != 0 ? open(path.c_str(), O_RDONLY) : -1;
        if (fd >= 0) {
            struct stat sb;
            if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
                void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) {
                    madvise(m, (size_t)sb.st_size, MADV_WILLNEED);
                    close(fd);
                    this->end_ = (size_t)sb.st_size;
                    this->data_ = std::make_shared<const FileImage>((const char*)m, (size_t)sb.st_size);
                    return;
                }
            }
            close(fd);
        }
        data_ = read_text_file(path);
    }
    this->end_ = data_local.size();
    this->data_ = std::make_shared<const FileImage>(std::move(data_local));
}

FileImage::~FileImage() { if (mapped_ && p_) munmap(const_cast<char*>(p_), n_); }

bool FASTAReader::readLine(std::string& line) {      // BufferedReader#readLine: \n, \r\n, \r
    const FileImage& d = *data_;
    if (pos_ >= end_) return false;
    size_t st = pos_;
    while (pos_ < end_ && d[pos_] != '\n' && d[pos_] != '\r') pos_++;
    line.assign(d.data() + st, pos_ - st);
    if (pos_ < end_) {
        if (d[pos_] == '\r' && pos_ + 1 < d.size() && d[pos_ + 1] == '\n') pos_ += 2;
        else pos_++;
    }
    return true;
}

// FASTAHeader.<init> (fasta-utils-1.1.jar): a header part must match ^(\S+)\s*$ or ^(\S+)\s+(\S+.*)$, i.e. it is
// non-empty and starts with a non-blank character (\s = [ \t\n\x0B\f\r]).
static bool fasta_header_part_ok(const char* p, size_t n) {
    if (n == 0) return false;
    const unsigned char c = (unsigned char)p[0];
    return !(c == ' ' || (c >= 9 && c <= 13));
}

bool FASTAReader::readNext(FASTAEntry& e) {
    // org.yeastrc.fasta.FASTAReader#readNext, statement for statement (tests/golden/fasta_reader.json holds what the
    // reference's own class file returns), scanning the file image in place
    static unsigned char UP[256];
    static bool up_init = false;
    if (!up_init) {
        for (int c = 0; c < 256; c++) UP[c] = c >= 0x80 ? (unsigned char)c       // kept: written as U+FFFD (US-ASCII decoding)
                                              : (unsigned char)toupper(c);          // String#toUpperCase
        up_init = true;
    }
    std::string line;
    if (!have_pending_) {
        if (!readLine(line)) return false;
    } else { line = pending_; have_pending_ = false; }
    if (line.empty() || line[0] != '>')
        throw FASTADataError("FASTADataErrorException: Expected header line, but line did not start with \">\".");
    {   // the header minus '>' is split on Ctrl-A (String#split: trailing empty parts removed) and every part parsed
        const char* h = line.data() + 1;
        size_t hn = line.size() - 1;
        if (!memchr(h, 1, hn)) {
            if (!fasta_header_part_ok(h, hn)) throw FASTADataError("FASTADataErrorException: Could not parse FASTA header line");
        } else {
            while (hn > 0 && h[hn - 1] == 1) hn--;                    // trailing empties
            size_t st = 0;
            while (hn > 0 && st <= hn) {
                const char* q = (const char*)memchr(h + st, 1, hn - st);
                size_t en = q ? (size_t)(q - h) : hn;
                if (!fasta_header_part_ok(h + st, en - st)) throw FASTADataError("FASTADataErrorException: Could not parse FASTA header line");
                if (!q) break;
                st = en + 1;
            }
        }
    }
    e.headerLine = line;
    e.sequence.clear();
    const char* d = data_->data();
    const size_t n = end_;
    // `while (line.startsWith(";"))` right after the header dereferences null at EOF; then a '>' line is an error
    bool first = true;
    while (true) {
        if (pos_ >= n) {
            // (a stripe that ends behind a header line: the next stripe starts with '>', the reference's reader sees that line)
            if (first && n < data_->size()) throw FASTADataError("FASTADataErrorException: Did not get a sequence line after a header line");
            if (first) throw FASTADataError("NullPointerException: header line at end of file");
            break;
        }
        const char c0 = d[pos_];
        if (c0 == '>') {
            if (first) throw FASTADataError("FASTADataErrorException: Did not get a sequence line after a header line");
            readLine(pending_);                                          // next entry: leave it for the next call
            have_pending_ = true;
            break;
        }
        size_t st = pos_;                                // one line: up to \n, \r\n or \r (BufferedReader#readLine)
        const char* nl = (const char*)memchr(d + st, '\n', n - st);
        size_t end = nl ? (size_t)(nl - d) : n;
        const char* cr = (const char*)memchr(d + st, '\r', end - st);
        size_t stop = cr ? (size_t)(cr - d) : end;
        if (stop < n) pos_ = (d[stop] == '\r' && stop + 1 < n && d[stop + 1] == '\n') ? stop + 2 : stop + 1;
        else pos_ = n;
        if (c0 == ';' && stop > st) continue;            // comment lines do not end the "first line" check
        first = false;
        const size_t old = e.sequence.size();
        e.sequence.resize(old + (stop - st));
        char* o = &e.sequence[old];
        for (size_t i = st; i < stop; i++) o[i - st] = (char)UP[(unsigned char)d[i]];
    }
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
// Indexer.index(File fastaDoc, File taxonDoc) (Indexer.java:150-236) on a C-ABI indexer handle
void index_fasta_into(bsp_indexer* h, const std::string& fastaDoc, const std::string& taxonDoc) {
    std::string taxonTree;
    if (!taxonDoc.empty() && file_exists(taxonDoc)) taxonTree = read_text_file(taxonDoc);   // :157-161
    FASTAReader reader(fastaDoc);
    FASTAEntry read;
    std::string filename = base_name(fastaDoc);                                        // fastaDoc.getName()
    while (reader.readNext(read)) {
        std::string header = read.headerLine;
        if (!header.empty() && header[0] == '>') header = header.substr(1);            // :167-170
        int rc = bsp_index_add_entry(h, filename.c_str(), header.c_str(), taxonTree.c_str(), read.sequence.data(),
                                     (int64_t)read.sequence.size());
        if (rc != BSP_OK)                                                              // :227-229 log and continue
            fprintf(stderr, "ERROR Exception occurred during index construction: %s\n", last_err().c_str());
    }
}
void Indexer::index(const std::string& fastaDoc, const std::string& taxonDoc) {        // :150-236
    std::lock_guard<std::mutex> lk(mu_);
    if (!h_) throw std::runtime_error("indexer is closed");
    index_fasta_into(h_, fastaDoc, taxonDoc);
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
    o += "{\"query_header\":"; jsonm::write_fasta_string(o, queryHeader);
    o += ",\"query\":"; jsonm::write_fasta_string(o, query);
    o += ",\"result\":[";
    for (size_t i = 0; i < result.size(); i++) {
        const SearchResultEntry& e = result[i];
        if (i) o += ",";
        o += "{\"docid\":" + std::to_string(e.docId) + ",\"rank\":" + std::to_string(e.rank);
        o += ",\"filename\":"; jsonm::write_string(o, e.filename);
        o += ",\"header\":"; jsonm::write_fasta_string(o, e.header);
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
    workerThreads_ = conf->workerThreads > 0 ? conf->workerThreads : 1;
    initDocs();
}

// an already open handle (not closed by this object): the batch driver over an index that is resident
Classifier::Classifier(bsp_classifier* adopted, int workerThreads) : h_(adopted), owned_(false) {
    if (!adopted) throw IllegalArgument("handle is null");
    workerThreads_ = workerThreads > 0 ? workerThreads : 1;
    initDocs();
}

void Classifier::initDocs() {
    // stored fields of every document (IndexSearcher.doc, Classifier.java:361-363), rendered and parsed once
    int64_t nDocs = 0;
    check_rc(bsp_index_stats(h_, &nDocs, nullptr, nullptr, nullptr));
    docs_.resize((size_t)nDocs);
    for (int64_t d = 0; d < nDocs; d++) {
        const char *fn, *hd, *dir, *tax;
        check_rc(bsp_doc_fields(h_, (int32_t)d, &fn, &hd, &dir, &tax));
        DocInfo& di = docs_[(size_t)d];
        di.json = "\"filename\":"; jsonm::write_string(di.json, fn);
        di.json += ",\"header\":"; jsonm::write_fasta_string(di.json, hd);
        di.json += ",\"sequence_direction\":"; jsonm::write_string(di.json, dir);
        di.json += ",\"taxon_hierarchy\":"; jsonm::write_string(di.json, tax);
        di.hasTaxonomy = tax[0] != 0;
        if (di.hasTaxonomy) {
            try { di.ranked = TaxonTreeDescription::classifiable(TaxonTreeDescription::parse(tax)); }
            catch (const std::exception&) { di.taxonomyBad = true; }
        }
    }
    // ranked lineages interned for the device: makeClassificationResult's taxid comparisons
    // (Classifier.java:296-327) then run next to the merge kernel and the formatter only looks names up
    const char* hostLca = getenv("BSP_HOST_LCA");
    deviceLca_ = !(hostLca && atoi(hostLca) != 0) && nDocs > 0;
    if (deviceLca_) {
        std::vector<int64_t> off((size_t)nDocs + 1, 0);
        std::vector<int32_t> tax;
        std::vector<uint8_t> flags((size_t)nDocs, 0);
        for (int64_t d = 0; d < nDocs; d++) {
            const DocInfo& di = docs_[(size_t)d];
            flags[(size_t)d] = (uint8_t)((di.hasTaxonomy ? 1 : 0) | (di.taxonomyBad ? 2 : 0));
            for (const Taxonomy& t : di.ranked) tax.push_back((int32_t)t.taxid);
            off[(size_t)d + 1] = (int64_t)tax.size();
        }
        check_rc(bsp_classifier_set_lineages(h_, (int32_t)nDocs, off.data(), tax.empty() ? nullptr : tax.data(), flags.data()));
    }
}

void Classifier::classifyPacked(const char* seqBytes, const int64_t* offsets, int32_t n, PackedResults& out) {
    out.status.assign((size_t)n, BSP_READ_ERROR);
    out.nHits.assign((size_t)n, 0);
    out.topDoc.assign((size_t)n * BSP_TOPN, -1);
    out.topScore.assign((size_t)n * BSP_TOPN, 0.0f);
    if (n == 0) return;
    if (deviceLca_) {
        out.taxLabel.assign((size_t)n, 0);
        out.taxIndex.assign((size_t)n, -1);
        check_rc(bsp_classify_packed_tax(h_, n, seqBytes, offsets, out.status.data(), nullptr, out.nHits.data(), nullptr,
                                         out.topDoc.data(), out.topScore.data(), out.taxLabel.data(), out.taxIndex.data()));
    } else {
        out.taxLabel.clear(); out.taxIndex.clear();
        check_rc(bsp_classify_packed(h_, n, seqBytes, offsets, out.status.data(), nullptr, out.nHits.data(), nullptr, out.topDoc.data(),
                                     out.topScore.data()));
    }
    for (int32_t i = 0; i < n; i++)
        if (offsets[i + 1] == offsets[i]) out.status[(size_t)i] = BSP_READ_ERROR;       // Classifier.java:342-344
}

// Classifier.classify's tail (Classifier.java:356-371) + makeClassificationResult (:263-339) + the Jackson layout of
// ClassificationResult / SearchResultEntry, on the prepared per-document data
bool Classifier::appendResultJson(std::string& o, const char* header, size_t headerLen, const char* seq, size_t seqLen,
                                  const PackedResults& r, int32_t i, int* type) const {
    const int nh = r.nHits[(size_t)i];
    const int32_t* docs = &r.topDoc[(size_t)i * BSP_TOPN];
    const float* scores = &r.topScore[(size_t)i * BSP_TOPN];
    // label
    int ty = 0;
    const Taxonomy* tax = nullptr;
    if (!r.taxLabel.empty()) {                               // computed on the device (taxonomy_label_kernel)
        ty = r.taxLabel[(size_t)i];
        if (ty < 0) return false;                            // malformed taxonomy among the hits: the reference throws
        const int ti = r.taxIndex[(size_t)i];
        if (ti >= 0 && nh > 0) tax = &docs_[(size_t)docs[0]].ranked[(size_t)ti];
    } else if (nh == 1) {
        ty = 2;
        const DocInfo& d = docs_[(size_t)docs[0]];
        if (d.hasTaxonomy) { if (d.taxonomyBad) return false; if (!d.ranked.empty()) tax = &d.ranked[0]; }
    } else if (nh >= 2) {
        bool all = true;
        for (int j = 0; j < nh && all; j++) {
            const DocInfo& d = docs_[(size_t)docs[j]];
            if (!d.hasTaxonomy) all = false;                 // quick fail: VAGUE (the earlier entries were already parsed)
            else if (d.taxonomyBad) return false;
        }
        ty = 1;
        if (all) {
            for (const Taxonomy& t : docs_[(size_t)docs[0]].ranked) {
                bool common = true;
                for (int j = 1; j < nh && common; j++) {
                    bool found = false;
                    for (const Taxonomy& t2 : docs_[(size_t)docs[j]].ranked) if (t2.taxid == t.taxid) { found = true; break; }
                    common = found;
                }
                if (common) { ty = 2; tax = &t; break; }
            }
        }
    }
    *type = ty;
    o += "{\"query_header\":"; jsonm::write_fasta_string_n(o, header, headerLen);
    o += ",\"query\":"; jsonm::write_fasta_string_n(o, seq, seqLen);
    o += ",\"result\":[";
    for (int j = 0; j < nh; j++) {
        if (j) o.push_back(',');
        o += "{\"docid\":"; jsonm::append_int(o, docs[j]);
        o += ",\"rank\":"; jsonm::append_int(o, j);
        o.push_back(',');
        o += docs_[(size_t)docs[j]].json;
        o += ",\"score\":";
        jsonm::append_java_double(o, (double)scores[j]);
        o.push_back('}');
    }
    static const char* TYPE[3] = {"UNKNOWN", "VAGUE", "CLASSIFIED"};
    o += "],\"type\":\""; o += TYPE[ty]; o += "\",\"taxon_rank\":";
    if (tax) jsonm::write_string(o, tax->rank); else o += "\"unknown\"";
    o += ",\"taxon_name\":";
    if (tax) { if (tax->hasName) jsonm::write_string(o, tax->name); else o += "null"; }
    else o += "\"\"";
    o.push_back('}');
    return true;
}
Classifier::~Classifier() { close(); }
void Classifier::close() {
    if (h_ && owned_) bsp_classifier_close(h_);
    h_ = nullptr;
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

// Thread-safe like the reference's (LocalClassifier.java:96-118 and ClassifierServer.java:70-98 call it from
// worker_threads pool threads at once): concurrent calls are coalesced into GPU micro-batches by bsp_classify_one.
bool Classifier::classify(const std::string& header, const std::string& sequence, ClassificationResult& out) {
    if (sequence.empty()) throw IllegalArgument("sequence is null or empty");           // Classifier.java:342-344
    int32_t st = BSP_READ_ERROR, label = 0, nhits = 0, ntop = 0, tdoc[BSP_TOPN];
    float tscore[BSP_TOPN];
    check_rc(bsp_classify_one(h_, sequence.data(), (int32_t)sequence.size(), &st, &label, &nhits, &ntop, tdoc, tscore));
    if (st != BSP_READ_OK) return false;         // the reference throws (NPE / TooManyClauses); callers log + skip
    ClassificationResult r;
    r.queryHeader = header;
    r.query = sequence;
    for (int j = 0; j < nhits; j++) {                                   // Classifier.java:358-366: hits equal to the max
        SearchResultEntry e;
        e.docId = tdoc[j];
        e.rank = j;
        e.score = (double)tscore[j];
        const char *fn, *hd, *dir, *tax;
        check_rc(bsp_doc_fields(h_, e.docId, &fn, &hd, &dir, &tax));
        e.filename = fn; e.header = hd; e.sequenceDirection = dir; e.taxonHierarchy = tax;
        r.result.push_back(std::move(e));
    }
    makeClassificationResult(r);
    out = std::move(r);
    return true;
}

// ------------------------------------------------------------------ LocalClassifier
LocalClassifier::LocalClassifier(const Configuration* conf) : classifier_(conf) {}

// One batch of reads, packed back to back (what bsp_classify_packed takes) plus its results.
namespace {
struct ReadBatch {
    std::string headers, seqs;              // header lines (with '>') and sequences, concatenated
    std::vector<size_t> hoff{0};
    std::vector<int64_t> soff{0};
    Classifier::PackedResults res;
    size_t size() const { return soff.size() - 1; }
};

template <typename T>
class BoundedQueue {                        // bounded hand-off between pipeline stages (any number of producers / consumers)
public:
    explicit BoundedQueue(size_t cap) : cap_(cap) {}
    void push(T v) {
        std::unique_lock<std::mutex> lk(mu_);
        not_full_.wait(lk, [&] { return q_.size() < cap_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    void close() { std::lock_guard<std::mutex> lk(mu_); closed_ = true; not_empty_.notify_all(); }
    bool pop(T& v) {
        std::unique_lock<std::mutex> lk(mu_);
        not_empty_.wait(lk, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
private:
    std::mutex mu_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};
}  // namespace

// LocalClassifier.classify(File, File, File) (LocalClassifier.java:60-131).  The reference runs one Runnable per read on
// `worker_threads` threads and serialises the writer; here three stages overlap: a reader thread parses FASTA into
// batches, the calling thread classifies each batch on the GPU, a writer thread formats the JSON lines of a batch on
// `worker_threads` threads and writes them in read order.
void LocalClassifier::classify(const std::string& inputFasta, const std::string& classifyOutput, const std::string& summaryOutput) {
    if (inputFasta.empty()) throw IllegalArgument("inputFasta is null");               // LocalClassifier.java:61-63
    if (classifyOutput.empty()) throw IllegalArgument("classifyOutput is null");       // :65-67
    std::string pd = parent_dir(classifyOutput);
    if (!pd.empty() && !file_exists(pd)) make_dirs(pd);                                // :69-71
    if (!summaryOutput.empty()) { std::string sd = parent_dir(summaryOutput); if (!sd.empty() && !file_exists(sd)) make_dirs(sd); }
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    FASTAReader reader(inputFasta);
    const double file_load_ms = ms_since(t_begin);
    double reader_done_ms = 0, format_ms = 0, write_ms = 0;          // stage clocks (gpu_metrics.json)
    FILE* fw = fopen(classifyOutput.c_str(), "wb");
    if (!fw) throw IOError("cannot create " + classifyOutput);
    ClassificationResultSummary summary;
    summary.queryFilename = base_name(inputFasta);
    summary.startTime = now_ms();
    const size_t BATCH_READS = (size_t)std::max<long long>(1024, env_ll("BSP_HOST_BATCH_READS", 65536)), BATCH_BYTES = (size_t)64 << 20;
    typedef std::unique_ptr<ReadBatch> BatchPtr;
    BoundedQueue<BatchPtr> parsed(3), classified(3);
    std::exception_ptr reader_error, writer_error;
    std::atomic<long long> unscored_reads{0};   // non-empty reads the GPU could not score (e.g. positional index dropped)

    // Reader stage, striped: the file image is cut at entry starts into stripes of ~8 MB, `parsers` threads parse stripes
    // side by side (each with its own FASTAReader over its byte range), and the reader thread hands their entries on in file
    // order, packed into batches.  A malformed entry ends the stream where the reference's reader would have thrown: the
    // entries before it are delivered, everything behind it is dropped, the error is rethrown at the end.
    struct Stripe { BatchPtr batch; std::exception_ptr error; bool done = false; };
    const std::vector<size_t> cuts = reader.stripeCuts((size_t)std::max<long long>(1 << 16, env_ll("BSP_FASTA_STRIPE_BYTES", 8 << 20)));
    const size_t n_stripes = cuts.size() - 1;
    std::vector<Stripe> stripes(n_stripes);
    std::mutex smu;
    std::condition_variable scv;
    size_t next_stripe = 0, delivered = 0;
    bool stop_parsing = false;
    const long long hw = (long long)std::max(1u, std::thread::hardware_concurrency());
    const int parsers = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(n_stripes, env_ll("BSP_FASTA_PARSERS", 12)),
                                                                           std::max<long long>(1, hw / 2)));
    auto parse_stripes = [&] {
        while (true) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(smu);
                // at most 2 x parsers stripes ahead of the hand-off (bounded memory)
                scv.wait(lk, [&] { return stop_parsing || next_stripe >= n_stripes || next_stripe < delivered + 2 * (size_t)parsers; });
                if (stop_parsing || next_stripe >= n_stripes) return;
                i = next_stripe++;
            }
            Stripe st;
            st.batch.reset(new ReadBatch());
            try {
                FASTAReader part(reader, cuts[i], cuts[i + 1]);
                FASTAEntry read;
                ReadBatch& b = *st.batch;
                b.headers.reserve((cuts[i + 1] - cuts[i]) / 2);
                b.seqs.reserve(cuts[i + 1] - cuts[i]);
                while (part.readNext(read)) {
                    b.headers += read.headerLine;                                      // header INCLUDING '>' (:94)
                    b.hoff.push_back(b.headers.size());
                    b.seqs += read.sequence;
                    b.soff.push_back((int64_t)b.seqs.size());
                }
            } catch (...) { st.error = std::current_exception(); }
            {
                std::lock_guard<std::mutex> lk(smu);
                stripes[i].batch = std::move(st.batch);
                stripes[i].error = st.error;
                stripes[i].done = true;
            }
            scv.notify_all();
        }
    };
    std::thread reader_thread([&] {
        std::vector<std::thread> pool;
        for (int t = 0; t < parsers; t++) pool.emplace_back(parse_stripes);
        try {
            BatchPtr b(new ReadBatch());
            for (size_t i = 0; i < n_stripes; i++) {
                Stripe st;
                {
                    std::unique_lock<std::mutex> lk(smu);
                    scv.wait(lk, [&] { return stripes[i].done; });
                    st = std::move(stripes[i]);
                    delivered = i + 1;
                }
                scv.notify_all();
                ReadBatch& p = *st.batch;
                if (b->size() == 0 && !st.error && p.size() >= BATCH_READS / 2) {
                    parsed.push(std::move(st.batch));                                  // large enough on its own: no copy
                } else {
                    const size_t h0 = b->headers.size();
                    const int64_t s0 = (int64_t)b->seqs.size();
                    b->headers += p.headers;
                    b->seqs += p.seqs;
                    for (size_t k = 1; k < p.hoff.size(); k++) b->hoff.push_back(h0 + p.hoff[k]);
                    for (size_t k = 1; k < p.soff.size(); k++) b->soff.push_back(s0 + p.soff[k]);
                    if (b->size() >= BATCH_READS || b->seqs.size() >= BATCH_BYTES) { parsed.push(std::move(b)); b.reset(new ReadBatch()); }
                }
                if (st.error) { reader_error = st.error; break; }
            }
            if (b->size()) parsed.push(std::move(b));
        } catch (...) { reader_error = std::current_exception(); }
        { std::lock_guard<std::mutex> lk(smu); stop_parsing = true; }
        scv.notify_all();
        for (auto& th : pool) th.join();
        reader_done_ms = ms_since(t_begin);
        parsed.close();
    });

    const Classifier& cls = classifier_;
    const int T = (int)std::max<long long>(1, std::min<long long>(cls.workerThreads(), env_ll("BSP_HOST_FORMATTERS", 1 << 20)));
    // formatted JSON of one batch (one string per formatter thread, in read order) on its way to the file
    typedef std::unique_ptr<std::vector<std::string>> ChunkPtr;
    BoundedQueue<ChunkPtr> formatted(2);
    std::mutex spare_mu;
    std::vector<ChunkPtr> spare_chunks;             // written chunks on their way back to the formatter
    std::exception_ptr file_error;
    std::thread file_thread([&] {
        try {
            // striped writer: the chunks of a batch land at their file offsets from several threads at once
            const int fdw = fileno(fw);
            off_t file_off = 0;
            ChunkPtr c;
            while (formatted.pop(c)) {
                const auto t_wr = std::chrono::steady_clock::now();
                const size_t nc = c->size();
                std::vector<off_t> at(nc + 1, file_off);
                for (size_t i = 0; i < nc; i++) at[i + 1] = at[i] + (off_t)(*c)[i].size();
                std::atomic<size_t> next{0};
                std::atomic<bool> failed{false};
                auto put = [&] {
                    for (size_t i = next++; i < nc; i = next++) {
                        const std::string& o = (*c)[i];
                        size_t done = 0;
                        while (done < o.size()) {
                            const ssize_t n = pwrite(fdw, o.data() + done, o.size() - done, at[i] + (off_t)done);
                            if (n <= 0) { failed = true; return; }
                            done += (size_t)n;
                        }
                    }
                };
                const int W = (int)std::min<size_t>(nc, (size_t)std::max<long long>(1, std::min<long long>(env_ll("BSP_HOST_WRITERS", 4), hw / 4)));
                std::vector<std::thread> th;
                for (int t = 1; t < W; t++) th.emplace_back(put);
                put();
                for (auto& t : th) t.join();
                if (failed) throw IOError("write failed: " + classifyOutput);
                file_off = at[nc];
                write_ms += ms_since(t_wr);
                // the buffers go back to the formatter (their pages stay mapped: no page faults for the next batch's text)
                { std::lock_guard<std::mutex> lk(spare_mu); if (spare_chunks.size() < 4) spare_chunks.push_back(std::move(c)); }
            }
        } catch (...) {
            file_error = std::current_exception();
            ChunkPtr drop;
            while (formatted.pop(drop)) {}
        }
    });
    std::thread writer_thread([&] {
        try {
            BatchPtr b;
            while (classified.pop(b)) {
                const size_t n = b->size();
                const int nt = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, n / 4096));
                ChunkPtr chunk_holder;
                { std::lock_guard<std::mutex> lk(spare_mu); if (!spare_chunks.empty()) { chunk_holder = std::move(spare_chunks.back()); spare_chunks.pop_back(); } }
                if (!chunk_holder) chunk_holder.reset(new std::vector<std::string>());
                if (chunk_holder->size() > (size_t)nt) {           // keep the strings with the largest capacities
                    std::sort(chunk_holder->begin(), chunk_holder->end(), [](const std::string& a, const std::string& b) { return a.capacity() > b.capacity(); });
                }
                chunk_holder->resize((size_t)nt);
                for (std::string& o : *chunk_holder) o.clear();
                std::vector<std::string>& chunk = *chunk_holder;
                std::vector<long long> counts((size_t)nt * 3, 0);
                auto work = [&](int t) {
                    const size_t i0 = n * (size_t)t / (size_t)nt, i1 = n * (size_t)(t + 1) / (size_t)nt;
                    std::string& o = chunk[(size_t)t];
                    o.reserve((i1 - i0) * 700);
                    for (size_t i = i0; i < i1; i++) {
                        // reads the reference's classify() throws on (empty sequence, < 2 k-mers, > 10000 clauses) are
                        // logged and skipped: no line, not counted (LocalClassifier.java:112-114)
                        if (b->res.status[i] != BSP_READ_OK) {
                            // ... but a non-empty read with an error status is one the reference WOULD have classified
                            // (no scoring path resident for it): never silent — counted, reported on stderr and in gpu_metrics
                            if (b->res.status[i] == BSP_READ_ERROR && b->soff[i + 1] > b->soff[i]) unscored_reads++;
                            continue;
                        }
                        const size_t mark = o.size();
                        int type = 0;
                        if (!cls.appendResultJson(o, b->headers.data() + b->hoff[i], b->hoff[i + 1] - b->hoff[i],
                                                  b->seqs.data() + b->soff[i], (size_t)(b->soff[i + 1] - b->soff[i]), b->res,
                                                  (int32_t)i, &type)) {
                            o.resize(mark);
                            fprintf(stderr, "ERROR Exception occurred during search: malformed taxonomy\n");
                            continue;
                        }
                        o.push_back('\n');
                        counts[(size_t)t * 3 + (size_t)type]++;
                    }
                };
                const auto t_fmt = std::chrono::steady_clock::now();
                std::vector<std::thread> pool;
                for (int t = 1; t < nt; t++) pool.emplace_back(work, t);
                work(0);
                for (auto& th : pool) th.join();
                format_ms += ms_since(t_fmt);
                for (int t = 0; t < nt; t++) {
                    summary.unknown += counts[(size_t)t * 3]; summary.vague += counts[(size_t)t * 3 + 1];
                    summary.classified += counts[(size_t)t * 3 + 2];
                }
                formatted.push(std::move(chunk_holder));       // the file thread writes it while the next batch is formatted
            }
        } catch (...) {
            writer_error = std::current_exception();
            BatchPtr drop;
            while (classified.pop(drop)) {}
        }
        formatted.close();
    });

    std::exception_ptr gpu_error;
    int64_t gpu_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double gpu_score_ms = 0, gpu_pipe_ms = 0, gpu_wall_ms = 0;
    long long gpu_batches = 0;
    try {
        BatchPtr b;
        while (parsed.pop(b)) {
            const auto t0 = std::chrono::steady_clock::now();
            classifier_.classifyPacked(b->seqs.data(), b->soff.data(), (int32_t)b->size(), b->res);
            gpu_wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            int64_t c[8]; double sm = 0, pm = 0;
            if (bsp_last_counters(classifier_.handle(), c) == BSP_OK) for (int i = 0; i < 8; i++) gpu_counters[i] += c[i];
            if (bsp_last_timings(classifier_.handle(), &sm, &pm) == BSP_OK) { gpu_score_ms += sm; gpu_pipe_ms += pm; }
            gpu_batches++;
            classified.push(std::move(b));
        }
    } catch (...) {
        gpu_error = std::current_exception();
        BatchPtr drop;
        while (parsed.pop(drop)) {}
    }
    classified.close();
    reader_thread.join();
    writer_thread.join();
    file_thread.join();
    fclose(fw);
    if (gpu_error) std::rethrow_exception(gpu_error);
    if (writer_error) std::rethrow_exception(writer_error);
    if (file_error) std::rethrow_exception(file_error);
    if (reader_error) std::rethrow_exception(reader_error);
    summary.total = summary.unknown + summary.vague + summary.classified;
    summary.endTime = now_ms();
    if (!summaryOutput.empty()) {                                                      // :128-130
        FILE* fs = fopen(summaryOutput.c_str(), "wb");
        if (!fs) throw IOError("cannot create " + summaryOutput);
        std::string js = summary.toJson();
        fwrite(js.data(), 1, js.size(), fs);
        fclose(fs);
    }
    lastSummary_ = summary;
    // GPU-side counters next to the reference's files (not a reference output; SURVEY.md section 5 "metrics")
    {
        FILE* fm = fopen((classifyOutput + ".gpu_metrics.json").c_str(), "wb");
        if (fm) {
            fprintf(fm, "{\"reads\":%lld,\"gpu_batches\":%lld,\"reads_filter_kernel\":%lld,\"reads_exact_dense_kernel\":%lld,"
                        "\"reads_positional_kernel\":%lld,\"reads_handed_over_by_filter\":%lld,\"docs_rescored_exactly\":%lld,"
                        "\"exception_cells\":%lld,\"kernel_launches\":%lld,\"score_kernel_ms\":%.3f,\"device_pipeline_ms\":%.3f,"
                        "\"classify_call_wall_ms\":%.3f,\"file_to_file_ms\":%lld,\"reads_not_scored_gpu_error\":%lld,"
                        "\"stage_ms\":{\"file_load\":%.1f,\"reader_done_at\":%.1f,\"format\":%.1f,\"write\":%.1f,\"parser_threads\":%d}}\n",
                    (long long)gpu_counters[0], gpu_batches, (long long)gpu_counters[3], (long long)gpu_counters[1],
                    (long long)gpu_counters[2], (long long)gpu_counters[7], (long long)gpu_counters[5], (long long)gpu_counters[6],
                    (long long)gpu_counters[4], gpu_score_ms, gpu_pipe_ms, gpu_wall_ms, summary.endTime - summary.startTime,
                    unscored_reads.load(), file_load_ms, reader_done_ms, format_ms, write_ms, parsers);
            fclose(fm);
        }
    }
    if (unscored_reads.load())
        fprintf(stderr, "ERROR %lld reads were not scored (status error: no scoring path resident for them, e.g. the positional "
                        "index was dropped for lack of device memory); they are missing from %s\n",
                unscored_reads.load(), classifyOutput.c_str());
}

// The reference's LocalClassifier.classify as written (LocalClassifier.java:88-120): the calling thread reads FASTA
// entries and hands one task per read to a pool of `threads` workers behind a queue of 2 x threads (BlockingExecutor);
// a worker calls Classifier.classify(header, sequence), serialises the result and writes the line under a lock, so the
// line order is the completion order, as in the reference.  Nothing here batches: bsp_classify_one coalesces the
// workers' concurrent calls on the GPU side.
void LocalClassifier::classifyPerRead(const std::string& inputFasta, const std::string& classifyOutput, const std::string& summaryOutput,
                                      int threads) {
    if (inputFasta.empty()) throw IllegalArgument("inputFasta is null");
    if (classifyOutput.empty()) throw IllegalArgument("classifyOutput is null");
    std::string pd = parent_dir(classifyOutput);
    if (!pd.empty() && !file_exists(pd)) make_dirs(pd);
    if (!summaryOutput.empty()) { std::string sd = parent_dir(summaryOutput); if (!sd.empty() && !file_exists(sd)) make_dirs(sd); }
    FASTAReader reader(inputFasta);
    FILE* fw = fopen(classifyOutput.c_str(), "wb");
    if (!fw) throw IOError("cannot create " + classifyOutput);
    ClassificationResultSummary summary;
    summary.queryFilename = base_name(inputFasta);
    summary.startTime = now_ms();
    const int T = std::max(1, threads > 0 ? threads : classifier_.workerThreads());
    typedef std::pair<std::string, std::string> Task;              // (header line, sequence)
    BoundedQueue<Task> tasks((size_t)T * 2);
    std::mutex out_mu;
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) pool.emplace_back([&] {
        Task task;
        std::string line;
        while (tasks.pop(task)) {
            try {
                ClassificationResult result;
                if (!classifier_.classify(task.first, task.second, result)) continue;   // the reference logs + skips
                line = result.toJson();
                line.push_back('\n');
                std::lock_guard<std::mutex> lk(out_mu);
                summary.report(result);
                fwrite(line.data(), 1, line.size(), fw);
            } catch (const std::exception& ex) {
                fprintf(stderr, "ERROR Exception occurred during search: %s\n", ex.what());
            }
        }
    });
    std::exception_ptr reader_error;
    try {
        FASTAEntry read;
        while (reader.readNext(read)) tasks.push(Task(read.headerLine, read.sequence));
    } catch (...) { reader_error = std::current_exception(); }
    tasks.close();
    for (auto& th : pool) th.join();
    fclose(fw);
    if (reader_error) std::rethrow_exception(reader_error);
    summary.total = summary.unknown + summary.vague + summary.classified;
    summary.endTime = now_ms();
    if (!summaryOutput.empty()) {
        FILE* fs = fopen(summaryOutput.c_str(), "wb");
        if (!fs) throw IOError("cannot create " + summaryOutput);
        std::string js = summary.toJson();
        fwrite(js.data(), 1, js.size(), fs);
        fclose(fs);
    }
    lastSummary_ = summary;
}

// ------------------------------------------------------------------ drivers (BioSpectra.java:110-161)
LocalClassifier::LocalClassifier(bsp_classifier* adopted, int workerThreads) : classifier_(adopted, workerThreads) {}

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
                      const std::string& outputDir, int device, int perReadThreads) {
    Configuration conf;
    if (!jsonConfiguration.empty()) conf = Configuration::createInstanceFromFile(jsonConfiguration);
    else { conf.indexPath = indexDir; conf.hasIndexPath = !indexDir.empty(); }
    conf.device = device;
    LocalClassifier classifier(&conf);
    if (!file_exists(outputDir)) make_dirs(outputDir);
    for (auto& f : FastaFileHelper::findFastaDocs(inputDir)) {
        std::string name = base_name(f);
        if (perReadThreads >= 0) classifier.classifyPerRead(f, outputDir + "/" + name + ".result", outputDir + "/" + name + ".result.sum", perReadThreads);
        else classifier.classify(f, outputDir + "/" + name + ".result", outputDir + "/" + name + ".result.sum");
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

// Same job with the reference's control flow: `threads` pool threads (0 = conf.worker_threads) each call
// Classifier.classify(header, sequence) per read (LocalClassifier.java:88-120); lines are written in completion order.
HOST_API int bsp_host_classify_local_per_read(const char* json_conf, const char* index_dir, const char* input_dir,
                                              const char* output_dir, int32_t device, int32_t threads) {
    try {
        if (threads < 0) throw bsp::IllegalArgument("threads must be >= 0");
        bsp::runClassifyLocal(json_conf ? json_conf : "", index_dir ? index_dir : "", input_dir ? input_dir : "",
                              output_dir ? output_dir : "", device, threads);
        return BSP_OK;
    } catch (const bsp::IllegalArgument& e) { g_host_error = e.what(); return BSP_ERR_ARG; }
    catch (const bsp::IOError& e) { g_host_error = e.what(); return BSP_ERR_IO; }
    catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_STATE; }
}

// Pure host helpers exported for CPU-side tests (no GPU needed)
HOST_API int bsp_host_classify_file(bsp_classifier* h, const char* input_fasta, const char* result_path, const char* summary_path,
                                    int32_t worker_threads) {
    try {
        if (!h || !input_fasta || !result_path) throw bsp::IllegalArgument("null argument");
        bsp::LocalClassifier lc(h, worker_threads);
        lc.classify(input_fasta, result_path, summary_path ? summary_path : "");
        return BSP_OK;
    } catch (const bsp::IllegalArgument& e) { g_host_error = e.what(); return BSP_ERR_ARG; }
    catch (const bsp::IOError& e) { g_host_error = e.what(); return BSP_ERR_IO; }
    catch (const std::exception& e) { g_host_error = e.what(); return BSP_ERR_STATE; }
}

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

// Indexer.index(File, File) behind the plain C ABI (what the JNI shim's indexAddFasta binds): the library's own
// FASTAReader + one bsp_index_add_entry per entry.  Entries read before a data error stay indexed, like the reference's.
HOST_API int bsp_index_add_fasta(bsp_indexer* h, const char* fasta_path, const char* taxd_path_or_null) {
    if (!h || !fasta_path) { g_host_error = "fastaDoc is null"; return BSP_ERR_ARG; }      // Indexer.java:151-153
    try {
        bsp::index_fasta_into(h, fasta_path, taxd_path_or_null ? std::string(taxd_path_or_null) : std::string());
        return BSP_OK;
    } catch (const std::exception& ex) { g_host_error = ex.what(); return BSP_ERR_IO; }
}

// parse a FASTA file with A.1 semantics; writes "header\x1Fsequence\x1E" records (unit / record separators: headers
// and sequences may hold tabs and, before the final trim, blanks).  On a data error the records read before
// it are still written (the reference's callers have processed those entries by the time readNext throws) and the
// call returns BSP_ERR_IO.
HOST_API int bsp_host_read_fasta(const char* path, char* out, int64_t cap, int64_t* needed) {
    std::string o;
    int rc = BSP_OK;
    try {
        bsp::FASTAReader rd(path);
        bsp::FASTAEntry e;
        try {
            while (rd.readNext(e)) { o += e.headerLine; o.push_back('\x1F'); o += e.sequence; o.push_back('\x1E'); }
        } catch (const bsp::FASTADataError& ex) { g_host_error = ex.what(); rc = BSP_ERR_IO; }
    } catch (const std::exception& ex) { g_host_error = ex.what(); return BSP_ERR_IO; }
    if (needed) *needed = (int64_t)o.size() + 1;
    if ((int64_t)o.size() + 1 > cap) return BSP_ERR_NOMEM;
    memcpy(out, o.c_str(), o.size() + 1);
    return rc;
}

// the same through stripes of `stripe_bytes` (what LocalClassifier's reader stage does on several threads): must equal
// bsp_host_read_fasta byte for byte, error code included (the entries before a malformed one are kept)
HOST_API int bsp_host_read_fasta_striped(const char* path, int64_t stripe_bytes, char* out, int64_t cap, int64_t* needed) {
    std::string o;
    int rc = BSP_OK;
    try {
        bsp::FASTAReader whole(path);
        const std::vector<size_t> cuts = whole.stripeCuts((size_t)std::max<int64_t>(stripe_bytes, 1));
        for (size_t i = 0; i + 1 < cuts.size() && rc == BSP_OK; i++) {
            bsp::FASTAReader rd(whole, cuts[i], cuts[i + 1]);
            bsp::FASTAEntry e;
            try {
                while (rd.readNext(e)) { o += e.headerLine; o.push_back('\x1F'); o += e.sequence; o.push_back('\x1E'); }
            } catch (const bsp::FASTADataError& ex) { g_host_error = ex.what(); rc = BSP_ERR_IO; }
        }
    } catch (const std::exception& ex) { g_host_error = ex.what(); return BSP_ERR_IO; }
    if (needed) *needed = (int64_t)o.size() + 1;
    if ((int64_t)o.size() + 1 > cap) return BSP_ERR_NOMEM;
    memcpy(out, o.c_str(), o.size() + 1);
    return rc;
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
