// Minimal JSON reader for the scene schema of Extensions/Scene/Loader.fs
// (System.Text.Json stands behind it in the reference).  Numbers keep their
// source text so float32 fields are parsed straight from decimal to binary32
// (strtof), as System.Text.Json does for `float32` record fields.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace bnjson {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  std::string text;  // number literal or string contents
  std::vector<ValuePtr> items;
  std::vector<std::pair<std::string, ValuePtr>> members;

  const Value* get(const std::string& key) const {
    if (kind != Object) return nullptr;
    for (auto& m : members)
      if (m.first == key) return m.second->kind == Null ? nullptr : m.second.get();
    return nullptr;
  }
  float as_f32() const {
    if (kind != Number) throw std::runtime_error("JSON: number expected");
    return std::strtof(text.c_str(), nullptr);
  }
  int as_int() const {
    if (kind != Number) throw std::runtime_error("JSON: number expected");
    return (int)std::strtol(text.c_str(), nullptr, 10);
  }
  bool as_bool() const {
    if (kind != Bool) throw std::runtime_error("JSON: bool expected");
    return b;
  }
  const std::string& as_string() const {
    if (kind != String) throw std::runtime_error("JSON: string expected");
    return text;
  }
  size_t size() const { return items.size(); }
  const Value& operator[](size_t i) const { return *items.at(i); }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {
    // UTF-8 BOM (Asset/cbox.json ships with one)
    if (s_.size() >= 3 && (unsigned char)s_[0] == 0xEF && (unsigned char)s_[1] == 0xBB && (unsigned char)s_[2] == 0xBF) p_ = 3;
  }
  ValuePtr parse() {
    ValuePtr v = value();
    ws();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;

  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error(std::string("JSON parse error at byte ") + std::to_string(p_) + ": " + what);
  }
  void ws() {
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
  }
  bool eat(char c) {
    ws();
    if (p_ < s_.size() && s_[p_] == c) { ++p_; return true; }
    return false;
  }
  ValuePtr value() {
    ws();
    if (p_ >= s_.size()) fail("unexpected end");
    char c = s_[p_];
    auto v = std::make_shared<Value>();
    if (c == '{') {
      ++p_;
      v->kind = Value::Object;
      if (eat('}')) return v;
      do {
        ws();
        if (p_ >= s_.size() || s_[p_] != '"') fail("member name expected");
        std::string k = string();
        if (!eat(':')) fail("':' expected");
        v->members.emplace_back(k, value());
      } while (eat(','));
      if (!eat('}')) fail("'}' expected");
    } else if (c == '[') {
      ++p_;
      v->kind = Value::Array;
      if (eat(']')) return v;
      do v->items.push_back(value()); while (eat(','));
      if (!eat(']')) fail("']' expected");
    } else if (c == '"') {
      v->kind = Value::String;
      v->text = string();
    } else if (s_.compare(p_, 4, "true") == 0) {
      v->kind = Value::Bool; v->b = true; p_ += 4;
    } else if (s_.compare(p_, 5, "false") == 0) {
      v->kind = Value::Bool; v->b = false; p_ += 5;
    } else if (s_.compare(p_, 4, "null") == 0) {
      v->kind = Value::Null; p_ += 4;
    } else {
      size_t b = p_;
      while (p_ < s_.size() && (std::isdigit((unsigned char)s_[p_]) || s_[p_] == '-' || s_[p_] == '+' || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E')) ++p_;
      if (b == p_) fail("value expected");
      v->kind = Value::Number;
      v->text = s_.substr(b, p_ - b);
    }
    return v;
  }
  std::string string() {
    ++p_;  // opening quote
    std::string out;
    while (p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if (c == '\\') {
        if (p_ >= s_.size()) fail("bad escape");
        char e = s_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (p_ + 4 > s_.size()) fail("bad \\u escape");
            unsigned cp = (unsigned)std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (p_ >= s_.size()) fail("unterminated string");
    ++p_;
    return out;
  }
};

inline ValuePtr parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace bnjson
