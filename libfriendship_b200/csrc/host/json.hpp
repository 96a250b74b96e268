// json.hpp — the small JSON subset of the reference's on-disk / wire format (serde_json):
// objects keep member order; integers are kept exactly (u64); serialisation is compact like serde_json::to_vec.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace frb {
namespace host {

struct Json {
    enum Type { Null, Bool, Int, Float, String, Array, Object } type = Null;
    bool b = false;
    uint64_t u = 0;          // magnitude for Int
    bool neg = false;
    double f = 0.0;
    std::string s;
    std::vector<Json> a;
    std::vector<std::pair<std::string, Json>> o;

    static Json null() { return Json(); }
    static Json integer(uint64_t v) { Json j; j.type = Int; j.u = v; return j; }
    static Json string(const std::string& v) { Json j; j.type = String; j.s = v; return j; }
    static Json array() { Json j; j.type = Array; return j; }
    static Json object() { Json j; j.type = Object; return j; }
    Json& add(const std::string& k, Json v) { o.emplace_back(k, std::move(v)); return *this; }
    const Json* get(const std::string& k) const {
        if (type != Object) return nullptr;
        for (auto& kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Json& at(const std::string& k) const {
        const Json* p = get(k);
        if (!p) throw std::runtime_error("missing field `" + k + "`");
        return *p;
    }
    uint64_t as_u64() const {
        if (type != Int || neg) throw std::runtime_error("expected an unsigned integer");
        return u;
    }
    const std::string& as_string() const {
        if (type != String) throw std::runtime_error("expected a string");
        return s;
    }
    const std::vector<Json>& as_array() const {
        if (type != Array) throw std::runtime_error("expected an array");
        return a;
    }

    // ---- serialise (compact) ----
    void write(std::string& out) const {
        switch (type) {
            case Null: out += "null"; break;
            case Bool: out += b ? "true" : "false"; break;
            case Int: if (neg) out += '-'; out += std::to_string(u); break;
            case Float: { char buf[40]; snprintf(buf, sizeof buf, "%.17g", f); out += buf; break; }
            case String: write_string(s, out); break;
            case Array:
                out += '[';
                for (size_t i = 0; i < a.size(); i++) { if (i) out += ','; a[i].write(out); }
                out += ']';
                break;
            case Object:
                out += '{';
                for (size_t i = 0; i < o.size(); i++) {
                    if (i) out += ',';
                    write_string(o[i].first, out);
                    out += ':';
                    o[i].second.write(out);
                }
                out += '}';
                break;
        }
    }
    std::string dump() const { std::string out; write(out); return out; }
    static void write_string(const std::string& s, std::string& out) {
        out += '"';
        for (unsigned char c : s) {
            switch (c) {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\r': out += "\\r"; break;
                case '\t': out += "\\t"; break;
                case '\b': out += "\\b"; break;
                case '\f': out += "\\f"; break;
                default:
                    if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                    else out += (char)c;
            }
        }
        out += '"';
    }

    // ---- parse ----
    static Json parse(const std::string& text) {
        size_t p = 0;
        Json j = parse_value(text, p, 0);
        skip_ws(text, p);
        if (p != text.size()) throw std::runtime_error("trailing characters after JSON value");
        return j;
    }

private:
    static void skip_ws(const std::string& t, size_t& p) {
        while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\r' || t[p] == '\t')) p++;
    }
    // nesting is bounded like serde_json's parser, which the reference loads effect files with (recursion limit 128):
    // a hostile file must not overflow the stack
    static constexpr int kMaxDepth = 128;
    static Json parse_value(const std::string& t, size_t& p, int depth) {
        if (depth > kMaxDepth) throw std::runtime_error("recursion limit exceeded");
        skip_ws(t, p);
        if (p >= t.size()) throw std::runtime_error("unexpected end of JSON");
        char c = t[p];
        if (c == '{') {
            Json j = object();
            p++;
            skip_ws(t, p);
            if (p < t.size() && t[p] == '}') { p++; return j; }
            for (;;) {
                skip_ws(t, p);
                if (p >= t.size() || t[p] != '"') throw std::runtime_error("expected object key");
                std::string k = parse_string(t, p);
                skip_ws(t, p);
                if (p >= t.size() || t[p] != ':') throw std::runtime_error("expected ':'");
                p++;
                j.o.emplace_back(k, parse_value(t, p, depth + 1));
                skip_ws(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == '}') { p++; return j; }
                throw std::runtime_error("expected ',' or '}'");
            }
        }
        if (c == '[') {
            Json j = array();
            p++;
            skip_ws(t, p);
            if (p < t.size() && t[p] == ']') { p++; return j; }
            for (;;) {
                j.a.push_back(parse_value(t, p, depth + 1));
                skip_ws(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == ']') { p++; return j; }
                throw std::runtime_error("expected ',' or ']'");
            }
        }
        if (c == '"') { Json j; j.type = String; j.s = parse_string(t, p); return j; }
        if (t.compare(p, 4, "null") == 0) { p += 4; return Json(); }
        if (t.compare(p, 4, "true") == 0) { p += 4; Json j; j.type = Bool; j.b = true; return j; }
        if (t.compare(p, 5, "false") == 0) { p += 5; Json j; j.type = Bool; return j; }
        // number
        size_t q = p;
        bool neg = false, is_float = false;
        if (t[q] == '-') { neg = true; q++; }
        size_t digits = q;
        while (q < t.size() && t[q] >= '0' && t[q] <= '9') q++;
        if (q == digits) throw std::runtime_error("invalid JSON value");
        if (q < t.size() && (t[q] == '.' || t[q] == 'e' || t[q] == 'E')) {
            is_float = true;
            while (q < t.size() && (t[q] == '.' || t[q] == 'e' || t[q] == 'E' || t[q] == '+' || t[q] == '-' || (t[q] >= '0' && t[q] <= '9'))) q++;
        }
        Json j;
        if (is_float) { j.type = Float; j.f = std::stod(t.substr(p, q - p)); }
        else { j.type = Int; j.neg = neg; j.u = std::stoull(t.substr(digits, q - digits)); }
        p = q;
        return j;
    }
    static std::string parse_string(const std::string& t, size_t& p) {
        std::string out;
        p++;   // opening quote
        while (p < t.size() && t[p] != '"') {
            char c = t[p++];
            if (c != '\\') { out += c; continue; }
            if (p >= t.size()) break;
            char e = t[p++];
            switch (e) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    unsigned cp = std::stoul(t.substr(p, 4), nullptr, 16);
                    p += 4;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                    else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += e;
            }
        }
        if (p >= t.size()) throw std::runtime_error("unterminated string");
        p++;   // closing quote
        return out;
    }
};

}  // namespace host
}  // namespace frb
