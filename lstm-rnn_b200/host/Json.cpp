#include "Json.hpp"
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace helpers {

bool JsonValue::HasMember(const std::string &name) const
{
    for (const auto &m : m_members) if (m.first == name) return true;
    return false;
}

const JsonValue &JsonValue::operator[](const std::string &name) const
{
    for (const auto &m : m_members) if (m.first == name) return m.second;
    throw std::runtime_error("Missing JSON member '" + name + "'");
}

JsonValue &JsonValue::member(const std::string &name)
{
    if (m_type == Null) m_type = Object;
    for (auto &m : m_members) if (m.first == name) return m.second;
    m_members.emplace_back(name, JsonValue());
    return m_members.back().second;
}

double JsonValue::GetDouble() const
{
    if (m_type != Number) throw std::runtime_error("JSON value is not a number");
    return m_num;
}

int JsonValue::GetInt() const
{
    if (m_type != Number) throw std::runtime_error("JSON value is not a number");
    return (int)m_num;
}

const std::string &JsonValue::GetString() const
{
    if (m_type != String) throw std::runtime_error("JSON value is not a string");
    return m_str;
}

static void writeString(std::string &out, const std::string &s)
{
    out += '"';
    for (char ch : s) {
        switch (ch) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\t': out += "\\t"; break;
            case '\r': out += "\\r"; break;
            default: out += ch;
        }
    }
    out += '"';
}

void JsonValue::write(std::string &out, bool pretty, int indent) const
{
    auto nl = [&](int ind) { if (pretty) { out += '\n'; out.append((size_t)ind * 4, ' '); } };
    switch (m_type) {
        case Null: out += "null"; break;
        case Bool: out += m_bool ? "true" : "false"; break;
        case Number: {
            char buf[64];
            // the reference's writer prints 6 significant digits ("%g"), which does not round-trip an fp32 weight; 9 digits do, and
            // the reference's reader parses them unchanged (tests/test_gpu_weightfile.py feeds this output to it)
            if (m_isInt) snprintf(buf, sizeof(buf), "%d", (int)m_num);
            else snprintf(buf, sizeof(buf), "%.9g", m_num);
            out += buf;
            break;
        }
        case String: writeString(out, m_str); break;
        case Array:
            out += '[';
            for (size_t i = 0; i < m_array.size(); ++i) {
                if (i) out += ',';
                nl(indent + 1);
                m_array[i].write(out, pretty, indent + 1);
            }
            if (!m_array.empty()) nl(indent);
            out += ']';
            break;
        case Object:
            out += '{';
            for (size_t i = 0; i < m_members.size(); ++i) {
                if (i) out += ',';
                nl(indent + 1);
                writeString(out, m_members[i].first);
                out += pretty ? ": " : ":";
                m_members[i].second.write(out, pretty, indent + 1);
            }
            if (!m_members.empty()) nl(indent);
            out += '}';
            break;
    }
}

std::string JsonValue::serialize(bool pretty) const
{
    std::string out;
    write(out, pretty, 0);
    return out;
}

class JsonParser {
public:
    explicit JsonParser(const std::string &t) : s(t), i(0) {}
    JsonValue parseDocument()
    {
        JsonValue v = parseValue();
        skip();
        if (i != s.size()) fail("trailing characters");
        return v;
    }
private:
    const std::string &s;
    size_t i;
    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("JSON parse error at offset ") + std::to_string(i) + ": " + what); }
    void skip() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; }
    JsonValue parseValue()
    {
        skip();
        if (i >= s.size()) fail("unexpected end");
        char c = s[i];
        if (c == '{') return parseObject();
        if (c == '[') return parseArray();
        if (c == '"') { JsonValue v; v.m_type = JsonValue::String; v.m_str = parseString(); return v; }
        if (!s.compare(i, 4, "true"))  { i += 4; JsonValue v; v.m_type = JsonValue::Bool; v.m_bool = true; return v; }
        if (!s.compare(i, 5, "false")) { i += 5; JsonValue v; v.m_type = JsonValue::Bool; v.m_bool = false; return v; }
        if (!s.compare(i, 4, "null"))  { i += 4; return JsonValue(); }
        return parseNumber();
    }
    JsonValue parseNumber()
    {
        const char *b = s.c_str() + i; char *e = nullptr;
        double d = strtod(b, &e);
        if (e == b) fail("bad value");
        bool isInt = true;
        for (const char *p = b; p < e; ++p) if (*p == '.' || *p == 'e' || *p == 'E') isInt = false;
        i += (size_t)(e - b);
        return JsonValue::makeNumber(d, isInt);
    }
    std::string parseString()
    {
        std::string out;
        ++i;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') {
                if (++i >= s.size()) fail("bad escape");
                switch (s[i]) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': { if (i + 4 >= s.size()) fail("bad \\u"); unsigned cp = (unsigned)strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16); out += (char)(cp < 128 ? cp : '?'); i += 4; break; }
                    default: out += s[i];
                }
                ++i;
            } else out += s[i++];
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    JsonValue parseArray()
    {
        JsonValue v = JsonValue::makeArray();
        ++i; skip();
        if (i < s.size() && s[i] == ']') { ++i; return v; }
        for (;;) {
            v.m_array.push_back(parseValue());
            skip();
            if (i >= s.size()) fail("unterminated array");
            if (s[i] == ',') { ++i; continue; }
            if (s[i] == ']') { ++i; return v; }
            fail("expected , or ]");
        }
    }
    JsonValue parseObject()
    {
        JsonValue v = JsonValue::makeObject();
        ++i; skip();
        if (i < s.size() && s[i] == '}') { ++i; return v; }
        for (;;) {
            skip();
            if (i >= s.size() || s[i] != '"') fail("expected member name");
            std::string name = parseString();
            skip();
            if (i >= s.size() || s[i] != ':') fail("expected :");
            ++i;
            v.m_members.emplace_back(name, parseValue());
            skip();
            if (i >= s.size()) fail("unterminated object");
            if (s[i] == ',') { ++i; continue; }
            if (s[i] == '}') { ++i; return v; }
            fail("expected , or }");
        }
    }
};

JsonDocument parseJson(const std::string &text) { return JsonParser(text).parseDocument(); }

int safeJsonGetInt(const JsonValue &val, const char *name)
{
    return val.HasMember(name) ? val[name].GetInt() : 0;
}

} // namespace helpers
