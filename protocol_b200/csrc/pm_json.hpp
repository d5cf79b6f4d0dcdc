// pm_json.hpp — a small JSON reader for the discovery wire format (host side only).
// Just enough of RFC 8259 for `GET {discovery}/api/pool/{id}` bodies
// (crates/orchestrator/src/discovery/monitor.rs:109-193, shared/src/models/node.rs:552-570).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
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

 public:
  // ---- streaming interface: the caller walks objects/arrays itself and pulls or skips values --------------
  const std::string& error() const { return err_; }
  bool at_end() { skip(); return p_ == end_; }
  char peek() { skip(); return p_ < end_ ? *p_ : '\0'; }
  bool consume(char c) { skip(); if (p_ < end_ && *p_ == c) { ++p_; return true; } return false; }
  bool read_value(Value* out) { return value(out, 0); }
  bool read_string(std::string* out) { skip(); out->clear(); return string(out); }
  // an object key: points into the input when the key has no escape (the usual case), else into *scratch
  bool read_key(std::string* scratch, const char** begin, size_t* len) {
    skip();
    if (p_ >= end_ || *p_ != '"') return false;
    const char* q = p_ + 1;
    while (q < end_ && *q != '"' && *q != '\\') ++q;
    if (q < end_ && *q == '"') {
      *begin = p_ + 1;
      *len = size_t(q - p_ - 1);
      p_ = q + 1;
      return true;
    }
    scratch->clear();
    if (!string(scratch)) return false;
    *begin = scratch->data();
    *len = scratch->size();
    return true;
  }
  // skips one value of any shape without building it
  bool skip_value(int depth = 0) {
    if (depth > 64) { err_ = "JSON nested too deeply"; return false; }
    skip();
    if (p_ >= end_) return false;
    const char c = *p_;
    if (c == '{' || c == '[') {
      const char close = c == '{' ? '}' : ']';
      ++p_;
      skip();
      if (p_ < end_ && *p_ == close) { ++p_; return true; }
      for (;;) {
        if (c == '{') {
          skip();
          if (!skip_string()) return false;
          skip();
          if (p_ >= end_ || *p_ != ':') return false;
          ++p_;
        }
        if (!skip_value(depth + 1)) return false;
        skip();
        if (p_ < end_ && *p_ == ',') { ++p_; continue; }
        if (p_ < end_ && *p_ == close) { ++p_; return true; }
        return false;
      }
    }
    if (c == '"') return skip_string();
    if (c == 't') return lit("true");
    if (c == 'f') return lit("false");
    if (c == 'n') return lit("null");
    const char* s = p_;
    if (p_ < end_ && (*p_ == '-' || *p_ == '+')) ++p_;
    while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) ++p_;
    return p_ != s;
  }

 private:
  const char* p_;
  const char* end_;
  std::string err_;
  bool skip_string() {   // same acceptance as string(), nothing stored
    if (p_ >= end_ || *p_ != '"') return false;
    ++p_;
    while (p_ < end_ && *p_ != '"') {
      if (*p_++ == '\\') {
        if (p_ >= end_) return false;
        const char e = *p_++;
        if (e == 'u') {
          if (end_ - p_ < 4) return false;
          for (int i = 0; i < 4; ++i) {
            const char h = *p_++;
            if (!((h >= '0' && h <= '9') || (h >= 'a' && h <= 'f') || (h >= 'A' && h <= 'F'))) return false;
          }
        } else if (e != '"' && e != '\\' && e != '/' && e != 'b' && e != 'f' && e != 'n' && e != 'r' && e != 't') {
          return false;
        }
      }
    }
    if (p_ >= end_) return false;
    ++p_;
    return true;
  }
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
    for (;;) {
      const char* run = p_;   // a run without escapes is appended in one piece
      while (p_ < end_ && *p_ != '"' && *p_ != '\\') ++p_;
      out->append(run, p_);
      if (p_ >= end_) return false;
      if (*p_ == '"') { ++p_; return true; }
      ++p_;   // the backslash
      if (p_ >= end_) return false;
      const char e = *p_++;
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
    }
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
    if (c == 't') { if (!lit("true")) return false; v->kind = Value::Bool; v->b = true; return true; }
    if (c == 'f') { if (!lit("false")) return false; v->kind = Value::Bool; v->b = false; return true; }
    if (c == 'n') { if (!lit("null")) return false; v->kind = Value::Null; return true; }
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
