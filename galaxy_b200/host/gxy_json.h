// gxy_json.h -- a small JSON reader for Galaxy state files, volume headers and partition documents.
// The reference parses these with rapidjson (third-party/rapidjson, e.g. Application::OpenJSONFile,
// src/framework/Application.cpp); only the read-side subset it uses is provided here: objects, arrays,
// strings, numbers (kept as double, with an "is integer" note), true/false/null, // and /* */ comments are
// NOT accepted (rapidjson's default flags reject them too).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace gxy {
namespace json {

class Value {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };
  Type type = Null;
  bool b = false;
  double num = 0.0;
  bool is_int = false;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;  // insertion order is kept (operators are positional)

  bool IsNull() const { return type == Null; }
  bool IsBool() const { return type == Bool; }
  bool IsNumber() const { return type == Number; }
  bool IsString() const { return type == String; }
  bool IsArray() const { return type == Array; }
  bool IsObject() const { return type == Object; }
  bool HasMember(const std::string &k) const { return Find(k) != nullptr; }
  const Value *Find(const std::string &k) const {
    if (type != Object) return nullptr;
    for (const auto &kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Value &operator[](const std::string &k) const {
    const Value *v = Find(k);
    if (!v) throw std::runtime_error("JSON: missing member \"" + k + "\"");
    return *v;
  }
  const Value &operator[](size_t i) const {
    if (type != Array || i >= arr.size()) throw std::runtime_error("JSON: array index out of range");
    return arr[i];
  }
  size_t Size() const { return type == Array ? arr.size() : type == Object ? obj.size() : 0; }
  double GetDouble() const {
    if (type != Number) throw std::runtime_error("JSON: number expected");
    return num;
  }
  int GetInt() const { return (int)GetDouble(); }
  bool GetBool() const {
    if (type == Bool) return b;
    if (type == Number) return num != 0.0;
    throw std::runtime_error("JSON: bool expected");
  }
  const std::string &GetString() const {
    if (type != String) throw std::runtime_error("JSON: string expected");
    return str;
  }
};

class Parser {
 public:
  explicit Parser(const std::string &text) : s(text), p(0) {}
  Value Parse() {
    Value v = ParseValue();
    SkipWs();
    if (p != s.size()) Fail("trailing characters");
    return v;
  }

 private:
  const std::string &s;
  size_t p;
  [[noreturn]] void Fail(const char *what) const {
    std::ostringstream o;
    o << "JSON parse error at offset " << p << ": " << what;
    throw std::runtime_error(o.str());
  }
  void SkipWs() {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++;
  }
  Value ParseValue() {
    SkipWs();
    if (p >= s.size()) Fail("unexpected end");
    const char c = s[p];
    if (c == '{') return ParseObject();
    if (c == '[') return ParseArray();
    if (c == '"') { Value v; v.type = Value::String; v.str = ParseString(); return v; }
    if (c == 't' || c == 'f' || c == 'n') return ParseLiteral();
    return ParseNumber();
  }
  Value ParseLiteral() {
    Value v;
    if (!s.compare(p, 4, "true")) { v.type = Value::Bool; v.b = true; p += 4; }
    else if (!s.compare(p, 5, "false")) { v.type = Value::Bool; v.b = false; p += 5; }
    else if (!s.compare(p, 4, "null")) { v.type = Value::Null; p += 4; }
    else Fail("bad literal");
    return v;
  }
  Value ParseNumber() {
    const char *b = s.c_str() + p;
    char *e = nullptr;
    const double d = strtod(b, &e);
    if (e == b) Fail("bad number");
    Value v;
    v.type = Value::Number;
    v.num = d;
    v.is_int = true;
    for (const char *q = b; q < e; q++)
      if (*q == '.' || *q == 'e' || *q == 'E') v.is_int = false;
    p += (size_t)(e - b);
    return v;
  }
  std::string ParseString() {
    std::string out;
    p++;  // opening quote
    while (true) {
      if (p >= s.size()) Fail("unterminated string");
      const char c = s[p++];
      if (c == '"') break;
      if (c != '\\') { out.push_back(c); continue; }
      if (p >= s.size()) Fail("bad escape");
      const char e = s[p++];
      switch (e) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          if (p + 4 > s.size()) Fail("bad \\u escape");
          const unsigned cp = (unsigned)strtoul(s.substr(p, 4).c_str(), nullptr, 16);
          p += 4;
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xc0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3f))); }
          else { out.push_back((char)(0xe0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3f))); out.push_back((char)(0x80 | (cp & 0x3f))); }
          break;
        }
        default: Fail("bad escape");
      }
    }
    return out;
  }
  Value ParseArray() {
    Value v;
    v.type = Value::Array;
    p++;
    SkipWs();
    if (p < s.size() && s[p] == ']') { p++; return v; }
    while (true) {
      v.arr.push_back(ParseValue());
      SkipWs();
      if (p >= s.size()) Fail("unterminated array");
      if (s[p] == ',') { p++; continue; }
      if (s[p] == ']') { p++; break; }
      Fail("',' or ']' expected");
    }
    return v;
  }
  Value ParseObject() {
    Value v;
    v.type = Value::Object;
    p++;
    SkipWs();
    if (p < s.size() && s[p] == '}') { p++; return v; }
    while (true) {
      SkipWs();
      if (p >= s.size() || s[p] != '"') Fail("member name expected");
      std::string k = ParseString();
      SkipWs();
      if (p >= s.size() || s[p] != ':') Fail("':' expected");
      p++;
      v.obj.push_back(std::make_pair(k, ParseValue()));
      SkipWs();
      if (p >= s.size()) Fail("unterminated object");
      if (s[p] == ',') { p++; continue; }
      if (s[p] == '}') { p++; break; }
      Fail("',' or '}' expected");
    }
    return v;
  }
};

inline Value ParseString(const std::string &text) { return Parser(text).Parse(); }
inline Value ParseFile(const std::string &path) {
  std::ifstream in(path.c_str());
  if (!in) throw std::runtime_error("cannot open " + path);
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();
  return Parser(text).Parse();
}

}  // namespace json
}  // namespace gxy
