// pm_json.hpp — a small JSON reader for the discovery wire format (host side only).
// Just enough of RFC 8259 for `GET {discovery}/api/pool/{id}` bodies
// (crates/orchestrator/src/discovery/monitor.rs:109-193, shared/src/models/node.rs:552-570).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace pmjson {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;                       // String, and the raw text of a Number
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;

  const Value* get(const char* key) const {
    if (kind != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool is_null() const { return kind == Null; }
};

class Parser {
 public:
  Parser(const char* s, size_t n) : p_(s), end_(s + n) {}
  bool parse(Value* out, std::string* err) {
    skip();
    if (!value(out, 0)) {
      if (err) *err = err_.empty() ? "invalid JSON" : err_;
      return false;
    }
    skip();
    if (p_ != end_) {
      if (err) *err = "trailing characters after JSON value";
      return false;
    }
    return true;
  }

 private:
  const char* p_;
  const char* end_;
  std::string err_;
  void skip() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_;
  }
  bool lit(const char* w) {
    size_t n = std::char_traits<char>::length(w);
    if ((size_t)(end_ - p_) < n || std::char_traits<char>::compare(p_, w, n) != 0) return false;
    p_ += n;
    return true;
  }
  bool string(std::string* out) {
    if (p_ >= end_ || *p_ != '"') return false;
    ++p_;
    while (p_ < end_ && *p_ != '"') {
      char c = *p_++;
      if (c == '\\') {
        if (p_ >= end_) return false;
        char e = *p_++;
        switch (e) {
          case '"': out->push_back('"'); break;
          case '\\': out->push_back('\\'); break;
          case '/': out->push_back('/'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'n': out->push_back('\n'); break;
          case 'r': out->push_back('\r'); break;
          case 't': out->push_back('\t'); break;
          case 'u': {
            if (end_ - p_ < 4) return false;
            unsigned cp = 0;
            for (int i = 0; i < 4; ++i) {
              char h = *p_++;
              cp <<= 4;
              if (h >= '0' && h <= '9') cp |= unsigned(h - '0');
              else if (h >= 'a' && h <= 'f') cp |= unsigned(h - 'a' + 10);
              else if (h >= 'A' && h <= 'F') cp |= unsigned(h - 'A' + 10);
              else return false;
            }
            if (cp < 0x80) out->push_back(char(cp));
            else if (cp < 0x800) { out->push_back(char(0xC0 | (cp >> 6))); out->push_back(char(0x80 | (cp & 0x3F))); }
            else { out->push_back(char(0xE0 | (cp >> 12))); out->push_back(char(0x80 | ((cp >> 6) & 0x3F))); out->push_back(char(0x80 | (cp & 0x3F))); }
            break;
          }
          default: return false;
        }
      } else {
        out->push_back(c);
      }
    }
    if (p_ >= end_) return false;
    ++p_;
    return true;
  }
  bool value(Value* v, int depth) {
    if (depth > 64) { err_ = "JSON nested too deeply"; return false; }
    skip();
    if (p_ >= end_) return false;
    const char c = *p_;
    if (c == '{') {
      ++p_;
      v->kind = Value::Object;
      skip();
      if (p_ < end_ && *p_ == '}') { ++p_; return true; }
      for (;;) {
        skip();
        std::string k;
        if (!string(&k)) return false;
        skip();
        if (p_ >= end_ || *p_ != ':') return false;
        ++p_;
        Value child;
        if (!value(&child, depth + 1)) return false;
        v->obj.emplace_back(std::move(k), std::move(child));
        skip();
        if (p_ < end_ && *p_ == ',') { ++p_; continue; }
        if (p_ < end_ && *p_ == '}') { ++p_; return true; }
        return false;
      }
    }
    if (c == '[') {
      ++p_;
      v->kind = Value::Array;
      skip();
      if (p_ < end_ && *p_ == ']') { ++p_; return true; }
      for (;;) {
        Value child;
        if (!value(&child, depth + 1)) return false;
        v->arr.push_back(std::move(child));
        skip();
        if (p_ < end_ && *p_ == ',') { ++p_; continue; }
        if (p_ < end_ && *p_ == ']') { ++p_; return true; }
        return false;
      }
    }
    if (c == '"') { v->kind = Value::String; return string(&v->str); }
    if (lit("true")) { v->kind = Value::Bool; v->b = true; return true; }
    if (lit("false")) { v->kind = Value::Bool; v->b = false; return true; }
    if (lit("null")) { v->kind = Value::Null; return true; }
    const char* s = p_;
    if (p_ < end_ && (*p_ == '-' || *p_ == '+')) ++p_;
    while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) ++p_;
    if (p_ == s) return false;
    v->kind = Value::Number;
    v->str.assign(s, p_);
    v->num = std::strtod(v->str.c_str(), nullptr);
    return true;
  }
};

}  // namespace pmjson
