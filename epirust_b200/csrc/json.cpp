#include "json.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace epi {

namespace {
struct Parser {
    const std::string& s;
    size_t i = 0;
    explicit Parser(const std::string& text) : s(text) {}
    [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("JSON: " + what + " at byte " + std::to_string(i)); }
    void ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
    }
    char peek() {
        ws();
        if (i >= s.size()) fail("unexpected end of input");
        return s[i];
    }
    void expect(char c) {
        if (peek() != c) fail(std::string("expected '") + c + "'");
        ++i;
    }
    bool literal(const char* lit) {
        size_t n = 0;
        while (lit[n]) ++n;
        if (s.compare(i, n, lit) == 0) { i += n; return true; }
        return false;
    }
    std::string string() {
        expect('"');
        std::string out;
        while (true) {
            if (i >= s.size()) fail("unterminated string");
            char c = s[i++];
            if (c == '"') break;
            if (c == '\\') {
                if (i >= s.size()) fail("bad escape");
                char e = s[i++];
                switch (e) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        if (i + 4 > s.size()) fail("bad \\u escape");
                        unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
                        i += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: fail("bad escape");
                }
            } else {
                out += c;
            }
        }
        return out;
    }
    JsonValue value() {
        JsonValue v;
        char c = peek();
        if (c == '{') {
            ++i;
            v.kind = JsonValue::Object;
            if (peek() == '}') { ++i; return v; }
            while (true) {
                ws();
                std::string key = string();
                expect(':');
                v.obj.emplace_back(std::move(key), value());
                char d = peek();
                ++i;
                if (d == '}') break;
                if (d != ',') fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            ++i;
            v.kind = JsonValue::Array;
            if (peek() == ']') { ++i; return v; }
            while (true) {
                v.arr.push_back(value());
                char d = peek();
                ++i;
                if (d == ']') break;
                if (d != ',') fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = JsonValue::String;
            v.str = string();
        } else if (literal("true")) {
            v.kind = JsonValue::Bool; v.b = true;
        } else if (literal("false")) {
            v.kind = JsonValue::Bool; v.b = false;
        } else if (literal("null")) {
            v.kind = JsonValue::Null;
        } else {
            const char* start = s.c_str() + i;
            char* end = nullptr;
            v.num = std::strtod(start, &end);
            if (end == start) fail("unexpected character");
            v.kind = JsonValue::Number;
            i += (size_t)(end - start);
        }
        return v;
    }
};
}  // namespace

JsonValue json_parse(const std::string& text) {
    Parser p(text);
    JsonValue v = p.value();
    p.ws();
    if (p.i != text.size()) p.fail("trailing characters");
    return v;
}

std::string json_read_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

}  // namespace epi
