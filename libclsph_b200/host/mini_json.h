// mini_json.h -- just enough JSON for fluid_properties/*.json and simulation_properties/*.json:
// objects, numbers, booleans, strings, null, arrays. Like the parser the reference vendors, it
// reads ONE value and ignores whatever follows (the shipped files end in "};").
#pragma once

#include <cstdlib>
#include <istream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace clsph_host {

class json_value {
 public:
  enum kind_t { k_null, k_bool, k_number, k_string, k_object, k_array };
  kind_t kind = k_null;
  bool boolean = false;
  double number = 0.0;
  std::string text;
  std::map<std::string, json_value> members;
  std::vector<json_value> items;

  const json_value& at(const std::string& key) const {
    if (kind != k_object) throw std::runtime_error("JSON: not an object while looking for \"" + key + "\"");
    auto it = members.find(key);
    if (it == members.end()) throw std::runtime_error("JSON: missing key \"" + key + "\"");
    return it->second;
  }
  double as_number(const std::string& what) const {
    if (kind != k_number) throw std::runtime_error("JSON: \"" + what + "\" is not a number");
    return number;
  }
  bool as_bool(const std::string& what) const {
    if (kind != k_bool) throw std::runtime_error("JSON: \"" + what + "\" is not a boolean");
    return boolean;
  }
};

class json_reader {
 public:
  explicit json_reader(const std::string& s) : s_(s), at_(0) {}

  json_value parse() {
    json_value v = value();
    return v;  // trailing characters (";") are not looked at
  }

 private:
  const std::string& s_;
  size_t at_;

  void skip() {
    while (at_ < s_.size() && (s_[at_] == ' ' || s_[at_] == '\t' || s_[at_] == '\n' || s_[at_] == '\r')) ++at_;
  }
  char peek() {
    skip();
    if (at_ >= s_.size()) throw std::runtime_error("JSON: unexpected end of input");
    return s_[at_];
  }
  void expect(char c) {
    if (peek() != c) throw std::runtime_error(std::string("JSON: expected '") + c + "' at offset " + std::to_string(at_));
    ++at_;
  }
  bool literal(const char* word) {
    size_t n = std::char_traits<char>::length(word);
    if (s_.compare(at_, n, word) == 0) {
      at_ += n;
      return true;
    }
    return false;
  }
  std::string string_body() {
    expect('"');
    std::string out;
    while (at_ < s_.size() && s_[at_] != '"') {
      char c = s_[at_++];
      if (c == '\\' && at_ < s_.size()) {
        char e = s_[at_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          default: out += e; break;  // \" \\ \/ (and \uXXXX left as-is: not needed for settings)
        }
      } else {
        out += c;
      }
    }
    if (at_ >= s_.size()) throw std::runtime_error("JSON: unterminated string");
    ++at_;
    return out;
  }
  json_value value() {
    json_value v;
    char c = peek();
    if (c == '{') {
      ++at_;
      v.kind = json_value::k_object;
      if (peek() == '}') { ++at_; return v; }
      while (true) {
        std::string key = string_body();
        expect(':');
        v.members[key] = value();
        if (peek() == ',') { ++at_; continue; }
        expect('}');
        return v;
      }
    }
    if (c == '[') {
      ++at_;
      v.kind = json_value::k_array;
      if (peek() == ']') { ++at_; return v; }
      while (true) {
        v.items.push_back(value());
        if (peek() == ',') { ++at_; continue; }
        expect(']');
        return v;
      }
    }
    if (c == '"') {
      v.kind = json_value::k_string;
      v.text = string_body();
      return v;
    }
    if (literal("true")) { v.kind = json_value::k_bool; v.boolean = true; return v; }
    if (literal("false")) { v.kind = json_value::k_bool; v.boolean = false; return v; }
    if (literal("null")) { v.kind = json_value::k_null; return v; }
    const char* begin = s_.c_str() + at_;
    char* end = nullptr;
    double d = std::strtod(begin, &end);
    if (end == begin) throw std::runtime_error("JSON: unexpected character at offset " + std::to_string(at_));
    at_ += static_cast<size_t>(end - begin);
    v.kind = json_value::k_number;
    v.number = d;
    return v;
  }
};

inline json_value parse_json_stream(std::istream& in) {
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  return json_reader(text).parse();
}

}  // namespace clsph_host
