#include "ph_ron.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace ph {

const RonValue* RonValue::get(const std::string& name) const {
    for (const auto& kv : fields)
        if (kv.first == name) return kv.second.get();
    return nullptr;
}

namespace {

struct Parser {
    const std::string& s;
    size_t i = 0;
    explicit Parser(const std::string& text) : s(text) {}

    [[noreturn]] void fail(const std::string& msg) {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); k++)
            if (s[k] == '\n') line++;
        throw std::runtime_error("line " + std::to_string(line) + ": " + msg);
    }

    void ws() {
        while (i < s.size()) {
            char c = s[i];
            if (c == ' ' || c == '\t' || c == '\r' || c == '\n') i++;
            else if (c == '/' && i + 1 < s.size() && s[i + 1] == '/') {
                while (i < s.size() && s[i] != '\n') i++;
            } else if (c == '/' && i + 1 < s.size() && s[i + 1] == '*') {
                size_t e = s.find("*/", i + 2);
                if (e == std::string::npos) fail("unterminated block comment");
                i = e + 2;
            } else break;
        }
    }
    char peek() {
        ws();
        return i < s.size() ? s[i] : '\0';
    }
    void expect(char c) {
        if (peek() != c) fail(std::string("expected `") + c + "`");
        i++;
    }
    static bool id_start(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
    static bool id_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }
    std::string ident() {
        size_t j = i;
        while (j < s.size() && id_char(s[j])) j++;
        std::string out = s.substr(i, j - i);
        i = j;
        return out;
    }

    RonPtr string_lit() {
        auto v = std::make_shared<RonValue>();
        v->kind = RonValue::String;
        size_t j = i + 1;
        while (true) {
            if (j >= s.size()) fail("unterminated string");
            char c = s[j];
            if (c == '"') break;
            if (c == '\\') {
                if (j + 1 >= s.size()) fail("bad escape");
                char e = s[j + 1];
                switch (e) {
                    case 'n': v->s += '\n'; break;
                    case 't': v->s += '\t'; break;
                    case 'r': v->s += '\r'; break;
                    case '0': v->s += '\0'; break;
                    case '\\': case '"': case '\'': v->s += e; break;
                    case 'u': {
                        size_t close = s.find('}', j);
                        if (close == std::string::npos) fail("bad \\u escape");
                        unsigned long cp = std::strtoul(s.substr(j + 3, close - j - 3).c_str(), nullptr, 16);
                        // UTF-8 encode
                        if (cp < 0x80) v->s += char(cp);
                        else if (cp < 0x800) { v->s += char(0xC0 | (cp >> 6)); v->s += char(0x80 | (cp & 0x3F)); }
                        else if (cp < 0x10000) { v->s += char(0xE0 | (cp >> 12)); v->s += char(0x80 | ((cp >> 6) & 0x3F)); v->s += char(0x80 | (cp & 0x3F)); }
                        else { v->s += char(0xF0 | (cp >> 18)); v->s += char(0x80 | ((cp >> 12) & 0x3F)); v->s += char(0x80 | ((cp >> 6) & 0x3F)); v->s += char(0x80 | (cp & 0x3F)); }
                        j = close - 1;
                        break;
                    }
                    default: fail(std::string("unknown escape \\") + e);
                }
                j += 2;
                continue;
            }
            v->s += c;
            j++;
        }
        i = j + 1;
        return v;
    }

    RonPtr raw_string() {
        size_t j = i + 1;
        size_t hashes = 0;
        while (j < s.size() && s[j] == '#') { hashes++; j++; }
        if (j >= s.size() || s[j] != '"') fail("bad raw string");
        std::string end = "\"" + std::string(hashes, '#');
        size_t e = s.find(end, j + 1);
        if (e == std::string::npos) fail("unterminated raw string");
        auto v = std::make_shared<RonValue>();
        v->kind = RonValue::String;
        v->s = s.substr(j + 1, e - j - 1);
        i = e + end.size();
        return v;
    }

    RonPtr number() {
        size_t j = i;
        if (s[j] == '+' || s[j] == '-') j++;
        bool is_float = false;
        if (s.compare(j, 3, "inf") == 0) {
            auto v = std::make_shared<RonValue>();
            v->kind = RonValue::Float;
            v->f = s[i] == '-' ? -INFINITY : INFINITY;
            i = j + 3;
            return v;
        }
        while (j < s.size() && (std::isdigit((unsigned char)s[j]) || s[j] == '_' || s[j] == '.' || s[j] == 'e' || s[j] == 'E' ||
                                ((s[j] == '+' || s[j] == '-') && (s[j - 1] == 'e' || s[j - 1] == 'E')))) {
            if (s[j] == '.' || s[j] == 'e' || s[j] == 'E') is_float = true;
            j++;
        }
        std::string tok;
        for (size_t k = i; k < j; k++)
            if (s[k] != '_') tok += s[k];
        if (tok.empty() || tok == "+" || tok == "-") fail("bad number");
        auto v = std::make_shared<RonValue>();
        if (is_float) {
            v->kind = RonValue::Float;
            v->f = std::strtod(tok.c_str(), nullptr);  // correctly rounded, like Rust's f64 parser
        } else {
            v->kind = RonValue::Int;
            v->i = std::strtoll(tok.c_str(), nullptr, 10);
        }
        i = j;
        return v;
    }

    // '(' ... ')' -> fills either fields (struct) or items (tuple); returns true when struct
    bool paren_body(RonValue& out) {
        expect('(');
        size_t save = i;
        bool is_struct = false;
        if (id_start(peek())) {
            ident();
            if (peek() == ':') is_struct = true;
        }
        i = save;
        while (true) {
            if (peek() == ')') { i++; break; }
            if (is_struct) {
                ws();
                std::string key = ident();
                if (key.empty()) fail("expected field name");
                expect(':');
                out.fields.emplace_back(key, value());
            } else {
                out.items.push_back(value());
            }
            if (peek() == ',') i++;
        }
        return is_struct;
    }

    RonPtr value() {
        char c = peek();
        if (c == '\0') fail("unexpected end of input");
        if (c == '(') {
            auto v = std::make_shared<RonValue>();
            bool st = paren_body(*v);
            v->kind = st ? RonValue::Struct : RonValue::List;
            return v;
        }
        if (c == '[') {
            i++;
            auto v = std::make_shared<RonValue>();
            v->kind = RonValue::List;
            while (true) {
                if (peek() == ']') { i++; break; }
                v->items.push_back(value());
                if (peek() == ',') i++;
            }
            return v;
        }
        if (c == '{') {
            i++;
            auto v = std::make_shared<RonValue>();
            v->kind = RonValue::Map;
            while (true) {
                if (peek() == '}') { i++; break; }
                RonPtr k = value();
                expect(':');
                v->map.emplace_back(k, value());
                if (peek() == ',') i++;
            }
            return v;
        }
        if (c == '"') return string_lit();
        if (c == 'r' && i + 1 < s.size() && (s[i + 1] == '"' || s[i + 1] == '#')) return raw_string();
        if (id_start(c)) {
            std::string name = ident();
            auto v = std::make_shared<RonValue>();
            if (name == "true" || name == "false") {
                v->kind = RonValue::Bool;
                v->b = name == "true";
                return v;
            }
            if (name == "inf" || name == "NaN") {
                v->kind = RonValue::Float;
                v->f = name == "inf" ? INFINITY : NAN;
                return v;
            }
            if (peek() == '(') {
                RonValue body;
                bool st = paren_body(body);
                if (name == "Some") {
                    if (st || body.items.size() != 1) fail("Some(..) takes exactly one value");
                    return body.items[0];
                }
                v->kind = RonValue::Tagged;
                v->s = name;
                v->tagged_struct = st;
                v->items = std::move(body.items);
                v->fields = std::move(body.fields);
                return v;
            }
            if (name == "None") return v;  // Null
            v->kind = RonValue::Tagged;
            v->s = name;
            return v;
        }
        if (std::isdigit((unsigned char)c) || c == '+' || c == '-' || c == '.') return number();
        fail(std::string("unexpected character `") + c + "`");
    }
};

}  // namespace

RonPtr ron_parse(const std::string& text, std::string& err) {
    try {
        Parser p(text);
        RonPtr v = p.value();
        if (p.peek() != '\0') p.fail("trailing characters");
        return v;
    } catch (const std::exception& e) {
        err = e.what();
        return nullptr;
    }
}

}  // namespace ph
