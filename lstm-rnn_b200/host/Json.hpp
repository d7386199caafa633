// Minimal JSON DOM for the network description / weight file (the reference vendors rapidjson v0.11;
// this is an independent implementation of just what the file format needs: helpers/JsonClasses.hpp,
// NeuralNetwork.cpp:37-130, 192-235).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace helpers {

class JsonValue {
public:
    enum Type { Null, Bool, Number, String, Array, Object };

    JsonValue() : m_type(Null), m_num(0), m_isInt(false), m_bool(false) {}
    static JsonValue makeArray()  { JsonValue v; v.m_type = Array;  return v; }
    static JsonValue makeObject() { JsonValue v; v.m_type = Object; return v; }
    static JsonValue makeNumber(double d, bool isInt = false) { JsonValue v; v.m_type = Number; v.m_num = d; v.m_isInt = isInt; return v; }
    static JsonValue makeString(const std::string &s) { JsonValue v; v.m_type = String; v.m_str = s; return v; }
    static JsonValue makeBool(bool b) { JsonValue v; v.m_type = Bool; v.m_bool = b; return v; }

    Type type() const { return m_type; }
    bool IsObject() const { return m_type == Object; }
    bool IsArray()  const { return m_type == Array; }
    bool IsNumber() const { return m_type == Number; }
    bool IsInt()    const { return m_type == Number && m_isInt; }
    bool IsString() const { return m_type == String; }
    bool IsBool()   const { return m_type == Bool; }
    bool GetBool() const { return m_bool; }

    bool HasMember(const std::string &name) const;
    const JsonValue &operator[](const std::string &name) const;   // throws if missing
    JsonValue &member(const std::string &name);                   // creates if missing (objects keep insertion order)
    const std::vector<std::pair<std::string, JsonValue>> &members() const { return m_members; }

    size_t Size() const { return m_array.size(); }
    const JsonValue &at(size_t i) const { return m_array.at(i); }
    JsonValue &at(size_t i) { return m_array.at(i); }
    void PushBack(const JsonValue &v) { m_array.push_back(v); }
    void Reserve(size_t n) { m_array.reserve(n); }

    double GetDouble() const;
    int GetInt() const;
    const std::string &GetString() const;
    void SetInt(int v) { m_type = Number; m_num = v; m_isInt = true; }

    // numbers are written like rapidjson v0.11's Writer::WriteDouble: "%g" (6 significant digits)
    std::string serialize(bool pretty = true) const;

private:
    void write(std::string &out, bool pretty, int indent) const;
    friend class JsonParser;
    Type m_type;
    double m_num;
    bool m_isInt, m_bool;
    std::string m_str;
    std::vector<JsonValue> m_array;
    std::vector<std::pair<std::string, JsonValue>> m_members;
};

typedef JsonValue JsonDocument;

// throws std::runtime_error("JSON parse error ...")
JsonDocument parseJson(const std::string &text);

int safeJsonGetInt(const JsonValue &val, const char *name);       // helpers/JsonClasses.cpp:32-38

} // namespace helpers
