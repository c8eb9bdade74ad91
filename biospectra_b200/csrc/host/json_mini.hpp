// json_mini.hpp -- minimal JSON value/parser/writer for the host-side mirror
// (configuration.json, .taxd lineages, .result / .result.sum output).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace jsonm {

struct Value;
typedef std::shared_ptr<Value> ValuePtr;

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0;
    bool is_int = false;
    long long inum = 0;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;   // insertion order kept

    const Value* get(const std::string& key) const {
        for (auto& kv : obj) if (kv.first == key) return kv.second.get();
        return nullptr;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s), p_(0) {}
    ValuePtr parse() {
        ValuePtr v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_;
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("JSON parse error: ") + m + " at " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++; }
    ValuePtr value() {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        char c = s_[p_];
        auto v = std::make_shared<Value>();
        if (c == '{') {
            v->type = Value::Object; p_++; ws();
            if (p_ < s_.size() && s_[p_] == '}') { p_++; return v; }
            while (true) {
                ws();
                if (p_ >= s_.size() || s_[p_] != '"') fail("expected key");
                std::string k = string();
                ws();
                if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
                p_++;
                v->obj.emplace_back(k, value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
                if (p_ < s_.size() && s_[p_] == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v->type = Value::Array; p_++; ws();
            if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
            while (true) {
                v->arr.push_back(value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
                if (p_ < s_.size() && s_[p_] == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v->type = Value::String; v->str = string();
        } else if (!s_.compare(p_, 4, "true")) { v->type = Value::Bool; v->b = true; p_ += 4; }
        else if (!s_.compare(p_, 5, "false")) { v->type = Value::Bool; v->b = false; p_ += 5; }
        else if (!s_.compare(p_, 4, "null")) { v->type = Value::Null; p_ += 4; }
        else {
            size_t st = p_;
            bool isint = true;
            if (p_ < s_.size() && (s_[p_] == '-' || s_[p_] == '+')) p_++;
            while (p_ < s_.size() && (isdigit((unsigned char)s_[p_]) || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E' || s_[p_] == '-' || s_[p_] == '+')) {
                if (s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E') isint = false;
                p_++;
            }
            if (p_ == st) fail("unexpected character");
            std::string t = s_.substr(st, p_ - st);
            v->type = Value::Number; v->num = strtod(t.c_str(), nullptr); v->is_int = isint;
            v->inum = isint ? strtoll(t.c_str(), nullptr, 10) : (long long)v->num;
        }
        return v;
    }
    static void put_utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o.push_back((char)cp);
        else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
        else { o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    std::string string() {
        std::string o;
        p_++;   // opening quote
        while (true) {
            if (p_ >= s_.size()) fail("unterminated string");
            char c = s_[p_++];
            if (c == '"') break;
            if (c != '\\') { o.push_back(c); continue; }
            if (p_ >= s_.size()) fail("bad escape");
            char e = s_[p_++];
            switch (e) {
                case '"': o.push_back('"'); break;
                case '\\': o.push_back('\\'); break;
                case '/': o.push_back('/'); break;
                case 'b': o.push_back('\b'); break;
                case 'f': o.push_back('\f'); break;
                case 'n': o.push_back('\n'); break;
                case 'r': o.push_back('\r'); break;
                case 't': o.push_back('\t'); break;
                case 'u': {
                    if (p_ + 4 > s_.size()) fail("bad \\u escape");
                    unsigned cp = (unsigned)strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
                    p_ += 4;
                    if (cp >= 0xD800 && cp < 0xDC00 && p_ + 6 <= s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                        unsigned lo = (unsigned)strtoul(s_.substr(p_ + 2, 4).c_str(), nullptr, 16);
                        if (lo >= 0xDC00 && lo < 0xE000) { cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); p_ += 6; }
                    }
                    put_utf8(o, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
        return o;
    }
};

inline ValuePtr parse(const std::string& s) { return Parser(s).parse(); }

// Jackson-style string escaping (JsonStringEncoder): ", \ and control chars
inline void write_string_n(std::string& o, const char* s, size_t n);
inline void write_string(std::string& o, const std::string& s) { write_string_n(o, s.data(), s.size()); }
// Text that came out of a FASTA file: the reference decodes those as US-ASCII (FASTAReaderConstants), so every byte
// >= 0x80 is the one character U+FFFD, which Jackson writes as the UTF-8 bytes EF BF BD.  The bytes are kept raw in
// memory (one byte per character keeps offsets and lengths right) and widened here.
inline void write_fasta_string_n(std::string& o, const char* sp, size_t n) {
    bool high = false;
    for (size_t i = 0; i < n; i++) if ((unsigned char)sp[i] >= 0x80) { high = true; break; }
    if (!high) { write_string_n(o, sp, n); return; }
    std::string t;
    t.reserve(n + 16);
    for (size_t i = 0; i < n; i++) {
        if ((unsigned char)sp[i] >= 0x80) t += "\xEF\xBF\xBD"; else t.push_back(sp[i]);
    }
    write_string_n(o, t.data(), t.size());
}
inline void write_fasta_string(std::string& o, const std::string& s) { write_fasta_string_n(o, s.data(), s.size()); }
inline void write_string_n(std::string& o, const char* sp, size_t n) {
    o.push_back('"');
    // fast path: runs of characters that need no escaping are appended in one go
    size_t run = 0;
    for (size_t i = 0; i < n; i++) {
        const unsigned char c = (unsigned char)sp[i];
        if (c >= 0x20 && c != '"' && c != '\\') continue;
        o.append(sp + run, i - run);
        run = i + 1;
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default: { char b[8]; snprintf(b, sizeof b, "\\u%04X", c); o += b; }
        }
    }
    o.append(sp + run, n - run);
    o.push_back('"');
}

// java.lang.Double#toString layout: shortest digits that round-trip; plain decimal for
// 1e-3 <= |d| < 1e7, otherwise d.dddE[-]n
// appends the text to `o` without intermediate strings (the result formatter calls this for every hit)
inline void append_java_double(std::string& o, double d) {
    if (std::isnan(d)) { o += "NaN"; return; }
    if (std::isinf(d)) { o += d > 0 ? "Infinity" : "-Infinity"; return; }
    if (d == 0) { o += std::signbit(d) ? "-0.0" : "0.0"; return; }
    char buf[64];
    // shortest digits that round-trip (std::to_chars), scientific layout: [-]d[.ddddd]e[+-]xx
    auto tc = std::to_chars(buf, buf + sizeof buf - 1, d, std::chars_format::scientific);
    const char* p = buf;
    const char* end = tc.ptr;
    if (*p == '-') { o.push_back('-'); p++; }
    const char* e = p;
    while (e < end && *e != 'e') e++;
    int exp10 = 0;
    {
        const char* q = e + 1;
        bool eneg = false;
        if (q < end && (*q == '+' || *q == '-')) { eneg = *q == '-'; q++; }
        for (; q < end; q++) exp10 = exp10 * 10 + (*q - '0');
        if (eneg) exp10 = -exp10;
    }
    char digits[32];
    int nd = 0;
    for (const char* q = p; q < e; q++) if (*q != '.') digits[nd++] = *q;
    const double ad = std::fabs(d);
    if (ad >= 1e-3 && ad < 1e7) {
        if (exp10 >= 0) {
            const int ni = exp10 + 1;                       // digits before the point
            for (int i = 0; i < ni; i++) o.push_back(i < nd ? digits[i] : '0');
            o.push_back('.');
            if (nd > ni) o.append(digits + ni, (size_t)(nd - ni)); else o.push_back('0');
        } else {
            o += "0.";
            o.append((size_t)(-exp10 - 1), '0');
            o.append(digits, (size_t)nd);
        }
    } else {
        o.push_back(digits[0]);
        o.push_back('.');
        if (nd > 1) o.append(digits + 1, (size_t)(nd - 1)); else o.push_back('0');
        o.push_back('E');
        char eb[16];
        auto te = std::to_chars(eb, eb + sizeof eb, exp10);
        o.append(eb, (size_t)(te.ptr - eb));
    }
}
inline std::string java_double(double d) {
    std::string o;
    append_java_double(o, d);
    return o;
}
inline void append_int(std::string& o, long long v) {
    char b[24];
    auto t = std::to_chars(b, b + sizeof b, v);
    o.append(b, (size_t)(t.ptr - b));
}

}  // namespace jsonm
