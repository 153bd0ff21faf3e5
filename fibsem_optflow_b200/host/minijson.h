// minijson.h -- a small JSON value + parser + writer for the job files of fibsem-optflow.
//
// The reference parses its job file with jsoncpp in non-strict mode (reference
// src/optflow.cpp:32-58: Json::Reader::parse(..., collectComments=false)), which accepts
// C and C++ comments; its own docs/example.json relies on that.  jsoncpp is not available
// here, so this header provides what the driver needs: objects keep their members in
// alphabetical order (jsoncpp's std::map order, which fixes the order in which the reference
// walks "rois", src/optflow.cpp:339), numbers remember whether they were integers, and the
// writer prints doubles with 17 significant digits like jsoncpp's StreamWriter.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mj {

struct Value;
using Array = std::vector<Value>;
using Object = std::map<std::string, Value>;

struct Value {
    enum Type { Null, Bool, Int, Real, String, Arr, Obj } type = Null;
    bool b = false;
    long long i = 0;
    double d = 0.0;
    std::string s;
    std::shared_ptr<Array> a;
    std::shared_ptr<Object> o;

    Value() {}
    Value(bool v) : type(Bool), b(v) {}
    Value(int v) : type(Int), i(v), d((double)v) {}
    Value(long long v) : type(Int), i(v), d((double)v) {}
    Value(double v) : type(Real), i((long long)v), d(v) {}
    Value(const char* v) : type(String), s(v) {}
    Value(const std::string& v) : type(String), s(v) {}
    static Value array() { Value v; v.type = Arr; v.a = std::make_shared<Array>(); return v; }
    static Value object() { Value v; v.type = Obj; v.o = std::make_shared<Object>(); return v; }

    bool isNull() const { return type == Null; }
    bool isObject() const { return type == Obj; }
    bool isArray() const { return type == Arr; }
    bool isMember(const std::string& k) const { return type == Obj && o->count(k) != 0; }
    size_t size() const { return type == Arr ? a->size() : (type == Obj ? o->size() : 0); }

    // jsoncpp-like conversions (lenient)
    bool asBool() const { return type == Bool ? b : (type == Int ? i != 0 : (type == Real ? d != 0.0 : false)); }
    long long asInt() const { return type == Int ? i : (type == Real ? (long long)d : (type == Bool ? (long long)b : 0)); }
    double asDouble() const { return type == Real ? d : (type == Int ? (double)i : (type == Bool ? (double)b : 0.0)); }
    float asFloat() const { return (float)asDouble(); }
    std::string asString() const
    {
        if (type == String) return s;
        if (type == Null) return "";
        std::ostringstream os;
        if (type == Bool) os << (b ? "true" : "false");
        else if (type == Int) os << i;
        else if (type == Real) os << d;
        return os.str();
    }

    const Value& get(const std::string& k, const Value& dflt) const
    {
        if (type != Obj) return dflt;
        auto it = o->find(k);
        return it == o->end() ? dflt : it->second;
    }
    const Value& at(const std::string& k) const
    {
        static const Value null_value;
        return get(k, null_value);
    }
    Value& operator[](const std::string& k)
    {
        if (type != Obj) { *this = object(); }
        return (*o)[k];
    }
    const Value& operator[](size_t k) const { return (*a)[k]; }
    Value& operator[](size_t k) { return (*a)[k]; }
    void append(const Value& v)
    {
        if (type != Arr) { *this = array(); }
        a->push_back(v);
    }
};

class Parser {
public:
    explicit Parser(const std::string& text) : t(text), p(0) {}
    Value parse()
    {
        Value v = value();
        ws();
        if (p != t.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& t;
    size_t p;
    [[noreturn]] void fail(const std::string& m) const
    {
        size_t line = 1;
        for (size_t k = 0; k < p && k < t.size(); k++) line += t[k] == '\n';
        throw std::runtime_error("JSON: " + m + " (line " + std::to_string(line) + ")");
    }
    void ws()
    {
        for (;;) {
            while (p < t.size() && (t[p] == ' ' || t[p] == '\t' || t[p] == '\n' || t[p] == '\r')) p++;
            if (p + 1 < t.size() && t[p] == '/' && t[p + 1] == '*') {
                size_t e = t.find("*/", p + 2);
                if (e == std::string::npos) fail("unterminated comment");
                p = e + 2;
            } else if (p + 1 < t.size() && t[p] == '/' && t[p + 1] == '/') {
                while (p < t.size() && t[p] != '\n') p++;
            } else {
                return;
            }
        }
    }
    Value value()
    {
        ws();
        if (p >= t.size()) fail("unexpected end");
        const char c = t[p];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') return Value(string());
        if (t.compare(p, 4, "true") == 0) { p += 4; return Value(true); }
        if (t.compare(p, 5, "false") == 0) { p += 5; return Value(false); }
        if (t.compare(p, 4, "null") == 0) { p += 4; return Value(); }
        return number();
    }
    Value object()
    {
        Value v = Value::object();
        p++;
        for (;;) {
            ws();
            if (p < t.size() && t[p] == '}') { p++; return v; }
            if (p >= t.size() || t[p] != '"') fail("expected a member name");
            const std::string k = string();
            ws();
            if (p >= t.size() || t[p] != ':') fail("expected ':'");
            p++;
            (*v.o)[k] = value();
            ws();
            if (p < t.size() && t[p] == ',') { p++; continue; }   // also tolerates a trailing comma
            if (p < t.size() && t[p] == '}') { p++; return v; }
            fail("expected ',' or '}'");
        }
    }
    Value array()
    {
        Value v = Value::array();
        p++;
        for (;;) {
            ws();
            if (p < t.size() && t[p] == ']') { p++; return v; }
            v.a->push_back(value());
            ws();
            if (p < t.size() && t[p] == ',') { p++; continue; }
            if (p < t.size() && t[p] == ']') { p++; return v; }
            fail("expected ',' or ']'");
        }
    }
    std::string string()
    {
        std::string r;
        p++;
        while (p < t.size() && t[p] != '"') {
            if (t[p] == '\\' && p + 1 < t.size()) {
                const char e = t[++p];
                switch (e) {
                    case 'n': r += '\n'; break;
                    case 't': r += '\t'; break;
                    case 'r': r += '\r'; break;
                    case 'b': r += '\b'; break;
                    case 'f': r += '\f'; break;
                    case 'u': {
                        if (p + 4 >= t.size()) fail("bad \\u escape");
                        const unsigned cp = (unsigned)std::strtoul(t.substr(p + 1, 4).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) r += (char)cp;
                        else if (cp < 0x800) { r += (char)(0xC0 | (cp >> 6)); r += (char)(0x80 | (cp & 0x3F)); }
                        else { r += (char)(0xE0 | (cp >> 12)); r += (char)(0x80 | ((cp >> 6) & 0x3F)); r += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: r += e;
                }
                p++;
            } else {
                r += t[p++];
            }
        }
        if (p >= t.size()) fail("unterminated string");
        p++;
        return r;
    }
    Value number()
    {
        const size_t s0 = p;
        bool real = false;
        if (p < t.size() && (t[p] == '-' || t[p] == '+')) p++;
        while (p < t.size() && (isdigit((unsigned char)t[p]) || t[p] == '.' || t[p] == 'e' || t[p] == 'E' || t[p] == '-' || t[p] == '+')) {
            if (t[p] == '.' || t[p] == 'e' || t[p] == 'E') real = true;
            p++;
        }
        if (p == s0) fail("unexpected character");
        const std::string tok = t.substr(s0, p - s0);
        if (real) return Value(std::strtod(tok.c_str(), nullptr));
        return Value((long long)std::strtoll(tok.c_str(), nullptr, 10));
    }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

inline void write(const Value& v, std::string& out, const std::string& indent, int depth)
{
    auto nl = [&](int d) {
        if (indent.empty()) return;
        out += '\n';
        for (int k = 0; k < d; k++) out += indent;
    };
    char buf[64];
    switch (v.type) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Int: std::snprintf(buf, sizeof(buf), "%lld", v.i); out += buf; break;
        case Value::Real:
            if (std::isfinite(v.d)) {
                std::snprintf(buf, sizeof(buf), "%.17g", v.d);
                out += buf;
                if (!std::strpbrk(buf, ".eEn")) out += ".0";
            } else {
                out += "null";
            }
            break;
        case Value::String:
            out += '"';
            for (char ch : v.s) {
                if (ch == '"' || ch == '\\') { out += '\\'; out += ch; }
                else if (ch == '\n') out += "\\n";
                else if (ch == '\t') out += "\\t";
                else out += ch;
            }
            out += '"';
            break;
        case Value::Arr:
            if (v.a->empty()) { out += "[]"; break; }
            out += '[';
            for (size_t k = 0; k < v.a->size(); k++) {
                if (k) out += ',';
                nl(depth + 1);
                write((*v.a)[k], out, indent, depth + 1);
            }
            nl(depth);
            out += ']';
            break;
        case Value::Obj:
            if (v.o->empty()) { out += "{}"; break; }
            out += '{';
            {
                bool first = true;
                for (const auto& kv : *v.o) {
                    if (!first) out += ',';
                    first = false;
                    nl(depth + 1);
                    out += '"' + kv.first + "\" : ";
                    write(kv.second, out, indent, depth + 1);
                }
            }
            nl(depth);
            out += '}';
            break;
    }
}

inline std::string dump(const Value& v, const std::string& indent = "   ")
{
    std::string out;
    write(v, out, indent, 0);
    return out;
}

}  // namespace mj
