// mini_json.hpp — the subset of JSON the rt.resynth job and preset files use (objects, strings, numbers, booleans,
// null, arrays), parsed into a small tree. The reference reads these files with nlohmann::json
// (source/rt.resynth.lib.params.cpp:3, :57-150, :183-247); nothing else of that library is needed here.
#pragma once

#include <cctype>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::map<std::string, Value> obj;

  const Value& at(const std::string& key) const {
    if (kind != Object) throw std::runtime_error("json: not an object (looking for '" + key + "')");
    auto it = obj.find(key);
    if (it == obj.end()) throw std::runtime_error("json: key '" + key + "' not found");
    return it->second;
  }
  const Value* find(const std::string& key) const {
    if (kind != Object) return nullptr;
    auto it = obj.find(key);
    return it == obj.end() ? nullptr : &it->second;
  }
  const std::string& as_string() const {
    if (kind != String) throw std::runtime_error("json: not a string");
    return str;
  }
  double as_number() const {
    if (kind == Bool) return b ? 1.0 : 0.0;
    if (kind != Number) throw std::runtime_error("json: not a number");
    return num;
  }
};

class Parser {
public:
  explicit Parser(const std::string& text) : s_(text) {}
  Value parse() {
    Value v = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const std::string& what) const {
    throw std::runtime_error("json: " + what + " at offset " + std::to_string(i_));
  }
  void ws() {
    while (i_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[i_]))) ++i_;
  }
  bool eat(char c) {
    ws();
    if (i_ < s_.size() && s_[i_] == c) {
      ++i_;
      return true;
    }
    return false;
  }
  Value value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    const char c = s_[i_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') {
      Value v;
      v.kind = Value::String;
      v.str = string();
      return v;
    }
    if (s_.compare(i_, 4, "true") == 0) { i_ += 4; Value v; v.kind = Value::Bool; v.b = true; return v; }
    if (s_.compare(i_, 5, "false") == 0) { i_ += 5; Value v; v.kind = Value::Bool; v.b = false; return v; }
    if (s_.compare(i_, 4, "null") == 0) { i_ += 4; return Value{}; }
    char* end = nullptr;
    const double d = std::strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) fail("unexpected character");
    i_ = static_cast<size_t>(end - s_.c_str());
    Value v;
    v.kind = Value::Number;
    v.num = d;
    return v;
  }
  std::string string() {
    if (!eat('"')) fail("expected string");
    std::string out;
    while (i_ < s_.size() && s_[i_] != '"') {
      char c = s_[i_++];
      if (c == '\\') {
        if (i_ >= s_.size()) fail("bad escape");
        const char e = s_[i_++];
        switch (e) {
          case 'n': c = '\n'; break;
          case 't': c = '\t'; break;
          case 'r': c = '\r'; break;
          case 'b': c = '\b'; break;
          case 'f': c = '\f'; break;
          case 'u': {  // file names in job files are ASCII / UTF-8; \uXXXX below 0x80 only
            if (i_ + 4 > s_.size()) fail("bad \\u escape");
            const unsigned long cp = std::strtoul(s_.substr(i_, 4).c_str(), nullptr, 16);
            i_ += 4;
            if (cp >= 0x80) fail("non-ASCII \\u escape not supported");
            c = static_cast<char>(cp);
            break;
          }
          default: c = e;  // \" \\ \/
        }
      }
      out.push_back(c);
    }
    if (i_ >= s_.size()) fail("unterminated string");
    ++i_;
    return out;
  }
  Value object() {
    Value v;
    v.kind = Value::Object;
    eat('{');
    if (eat('}')) return v;
    do {
      ws();
      std::string key = string();
      if (!eat(':')) fail("expected ':'");
      v.obj[key] = value();
    } while (eat(','));
    if (!eat('}')) fail("expected '}'");
    return v;
  }
  Value array() {
    Value v;
    v.kind = Value::Array;
    eat('[');
    if (eat(']')) return v;
    do {
      v.arr.push_back(value());
    } while (eat(','));
    if (!eat(']')) fail("expected ']'");
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace mini_json
