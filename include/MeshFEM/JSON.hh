// Minimal JSON value + parser for .material / .bc files.  (The reference uses nlohmann/json
// 3.1.2, which is not available offline; only the operations those two file formats need are
// provided: objects, arrays, strings, numbers, booleans, null; lookup, iteration, dump.)
#ifndef MESHFEM_B200_JSON_HH
#define MESHFEM_B200_JSON_HH
#include <cstdio>
#include <cstdlib>
#include <istream>
#include <iterator>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mjson {

class json {
public:
    enum Type { Null, Bool, Number, String, Array, Object };
    json() : m_type(Null) {}
    json(bool b) : m_type(Bool), m_bool(b) {}
    json(double d) : m_type(Number), m_num(d) {}
    json(int d) : m_type(Number), m_num(d) {}
    json(const std::string &s) : m_type(String), m_str(s) {}
    json(const char *s) : m_type(String), m_str(s) {}
    json(const std::vector<double> &v) : m_type(Array) { for (double d : v) m_arr.emplace_back(d); }
    static json array() { json j; j.m_type = Array; return j; }
    static json object() { json j; j.m_type = Object; return j; }

    Type type() const { return m_type; }
    bool is_null() const { return m_type == Null; }
    bool is_string() const { return m_type == String; }
    bool is_number() const { return m_type == Number; }
    bool is_boolean() const { return m_type == Bool; }
    bool is_array() const { return m_type == Array; }
    bool is_object() const { return m_type == Object; }

    size_t size() const { return m_type == Array ? m_arr.size() : (m_type == Object ? m_obj.size() : (m_type == Null ? 0 : 1)); }
    size_t count(const std::string &k) const { return (m_type == Object && m_obj.count(k)) ? 1 : 0; }

    const json &operator[](const std::string &k) const {
        if (m_type != Object) throw std::runtime_error("json: not an object (key '" + k + "')");
        auto it = m_obj.find(k);
        if (it == m_obj.end()) throw std::runtime_error("json: key '" + k + "' not found");
        return it->second;
    }
    json &operator[](const std::string &k) {
        if (m_type == Null) m_type = Object;
        if (m_type != Object) throw std::runtime_error("json: not an object");
        return m_obj[k];
    }
    const json &operator[](size_t i) const {
        if (m_type != Array || i >= m_arr.size()) throw std::runtime_error("json: bad array index");
        return m_arr[i];
    }
    void push_back(const json &j) {
        if (m_type == Null) m_type = Array;
        if (m_type != Array) throw std::runtime_error("json: not an array");
        m_arr.push_back(j);
    }
    std::vector<json>::const_iterator begin() const { return m_arr.begin(); }
    std::vector<json>::const_iterator end() const { return m_arr.end(); }
    const std::map<std::string, json> &items() const { return m_obj; }

    double number() const {
        if (m_type != Number) throw std::runtime_error("json: type must be number");
        return m_num;
    }
    const std::string &str() const {
        if (m_type != String) throw std::runtime_error("json: type must be string");
        return m_str;
    }
    bool boolean() const {
        if (m_type != Bool) throw std::runtime_error("json: type must be boolean");
        return m_bool;
    }
    operator double() const { return number(); }
    operator std::string() const { return str(); }

    bool value(const std::string &k, bool def) const { return count(k) ? (*this)[k].boolean() : def; }
    std::string value(const std::string &k, const char *def) const { return count(k) ? (*this)[k].str() : std::string(def); }
    double value(const std::string &k, double def) const { return count(k) ? (*this)[k].number() : def; }

    bool operator==(const json &o) const {
        if (m_type != o.m_type) return false;
        switch (m_type) {
            case Null: return true;
            case Bool: return m_bool == o.m_bool;
            case Number: return m_num == o.m_num;
            case String: return m_str == o.m_str;
            case Array: return m_arr == o.m_arr;
            case Object: return m_obj == o.m_obj;
        }
        return false;
    }
    bool operator!=(const json &o) const { return !(*this == o); }

    std::string dump() const {
        std::ostringstream os;
        m_dump(os);
        return os.str();
    }

    static json parse(const std::string &text) {
        size_t pos = 0;
        json j = m_parseValue(text, pos);
        m_skipWs(text, pos);
        if (pos != text.size()) throw std::runtime_error("json: trailing characters at offset " + std::to_string(pos));
        return j;
    }
    static json parse(std::istream &is) {
        std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
        return parse(text);
    }

private:
    Type m_type;
    bool m_bool = false;
    double m_num = 0.0;
    std::string m_str;
    std::vector<json> m_arr;
    std::map<std::string, json> m_obj;

    void m_dump(std::ostream &os) const {
        switch (m_type) {
            case Null: os << "null"; break;
            case Bool: os << (m_bool ? "true" : "false"); break;
            case Number: {
                char buf[40];
                if (m_num == (long long)m_num && std::abs(m_num) < 1e15) std::snprintf(buf, sizeof(buf), "%lld", (long long)m_num);
                else std::snprintf(buf, sizeof(buf), "%.17g", m_num);
                os << buf;
                break;
            }
            case String: os << '"' << m_str << '"'; break;
            case Array: {
                os << '[';
                for (size_t i = 0; i < m_arr.size(); ++i) { if (i) os << ','; m_arr[i].m_dump(os); }
                os << ']';
                break;
            }
            case Object: {
                os << '{';
                bool first = true;
                for (const auto &kv : m_obj) { if (!first) os << ','; first = false; os << '"' << kv.first << "\":"; kv.second.m_dump(os); }
                os << '}';
                break;
            }
        }
    }
    static void m_skipWs(const std::string &s, size_t &p) {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) ++p;
    }
    static std::runtime_error m_err(const std::string &what, size_t p) {
        return std::runtime_error("json parse error: " + what + " at offset " + std::to_string(p));
    }
    static std::string m_parseString(const std::string &s, size_t &p) {
        if (s[p] != '"') throw m_err("expected string", p);
        ++p;
        std::string out;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\') {
                ++p;
                if (p >= s.size()) throw m_err("bad escape", p);
                switch (s[p]) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (p + 4 >= s.size()) throw m_err("bad \\u escape", p);
                        unsigned code = std::stoul(s.substr(p + 1, 4), nullptr, 16);
                        if (code < 0x80) out += char(code);
                        else if (code < 0x800) { out += char(0xC0 | (code >> 6)); out += char(0x80 | (code & 0x3F)); }
                        else { out += char(0xE0 | (code >> 12)); out += char(0x80 | ((code >> 6) & 0x3F)); out += char(0x80 | (code & 0x3F)); }
                        p += 4;
                        break;
                    }
                    default: out += s[p];
                }
                ++p;
            } else out += s[p++];
        }
        if (p >= s.size()) throw m_err("unterminated string", p);
        ++p;
        return out;
    }
    static json m_parseValue(const std::string &s, size_t &p) {
        m_skipWs(s, p);
        if (p >= s.size()) throw m_err("unexpected end", p);
        const char c = s[p];
        if (c == '{') {
            json j = object();
            ++p; m_skipWs(s, p);
            if (p < s.size() && s[p] == '}') { ++p; return j; }
            while (true) {
                m_skipWs(s, p);
                std::string key = m_parseString(s, p);
                m_skipWs(s, p);
                if (p >= s.size() || s[p] != ':') throw m_err("expected ':'", p);
                ++p;
                j.m_obj[key] = m_parseValue(s, p);
                m_skipWs(s, p);
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == '}') { ++p; break; }
                throw m_err("expected ',' or '}'", p);
            }
            return j;
        }
        if (c == '[') {
            json j = array();
            ++p; m_skipWs(s, p);
            if (p < s.size() && s[p] == ']') { ++p; return j; }
            while (true) {
                j.m_arr.push_back(m_parseValue(s, p));
                m_skipWs(s, p);
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == ']') { ++p; break; }
                throw m_err("expected ',' or ']'", p);
            }
            return j;
        }
        if (c == '"') return json(m_parseString(s, p));
        if (s.compare(p, 4, "true") == 0) { p += 4; return json(true); }
        if (s.compare(p, 5, "false") == 0) { p += 5; return json(false); }
        if (s.compare(p, 4, "null") == 0) { p += 4; return json(); }
        const char *start = s.c_str() + p;
        char *end = nullptr;
        const double d = std::strtod(start, &end);
        if (end == start) throw m_err("unexpected character", p);
        p += size_t(end - start);
        return json(d);
    }
};

}  // namespace mjson
#endif
