// xml_lite.h — minimal XML reader for the MJCF subset (tinyxml2 is not available in the image).
// Handles elements, attributes (single or double quoted), comments, <?...?> and <!...> declarations,
// self-closing tags and the five predefined entities. Text nodes are ignored (MJCF carries no text).
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace b2 {

struct XmlElem {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlElem>> children;
  int line = 0;

  const char* attr(const char* key) const {
    for (auto& kv : attrs)
      if (kv.first == key) return kv.second.c_str();
    return nullptr;
  }
  bool has(const char* key) const { return attr(key) != nullptr; }
  void set(const std::string& key, const std::string& val) {
    for (auto& kv : attrs)
      if (kv.first == key) { kv.second = val; return; }
    attrs.emplace_back(key, val);
  }
  const XmlElem* child(const char* n) const {
    for (auto& c : children)
      if (c->name == n) return c.get();
    return nullptr;
  }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& text) : s_(text) {}

  std::unique_ptr<XmlElem> parse() {
    skip_misc();
    auto root = parse_elem();
    if (!root) fail("no root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  int line_ = 1;

  [[noreturn]] void fail(const std::string& msg) const {
    throw std::runtime_error("XML parse error (line " + std::to_string(line_) + "): " + msg);
  }
  bool eof() const { return p_ >= s_.size(); }
  char cur() const { return s_[p_]; }
  void adv() {
    if (s_[p_] == '\n') line_++;
    p_++;
  }
  bool starts(const char* lit) const { return s_.compare(p_, std::char_traits<char>::length(lit), lit) == 0; }
  void skip_ws() {
    while (!eof() && (cur() == ' ' || cur() == '\t' || cur() == '\n' || cur() == '\r')) adv();
  }
  void skip_until(const char* lit) {
    while (!eof() && !starts(lit)) adv();
    if (eof()) fail(std::string("unterminated construct, expected ") + lit);
    for (size_t i = 0; lit[i]; i++) adv();
  }
  // whitespace, text, comments, declarations between elements
  void skip_misc() {
    for (;;) {
      while (!eof() && cur() != '<') adv();
      if (eof()) return;
      if (starts("<!--")) skip_until("-->");
      else if (starts("<?")) skip_until("?>");
      else if (starts("<!")) skip_until(">");
      else return;
    }
  }
  static bool name_char(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' ||
           c == ':' || c == '.';
  }
  std::string parse_name() {
    size_t b = p_;
    while (!eof() && name_char(cur())) adv();
    if (p_ == b) fail("expected a name");
    return s_.substr(b, p_ - b);
  }
  static std::string unescape(const std::string& v) {
    if (v.find('&') == std::string::npos) return v;
    std::string o;
    for (size_t i = 0; i < v.size(); i++) {
      if (v[i] != '&') { o += v[i]; continue; }
      auto rep = [&](const char* ent, char c) {
        size_t n = std::char_traits<char>::length(ent);
        if (v.compare(i, n, ent) == 0) { o += c; i += n - 1; return true; }
        return false;
      };
      if (rep("&amp;", '&') || rep("&lt;", '<') || rep("&gt;", '>') || rep("&quot;", '"') || rep("&apos;", '\'')) continue;
      o += v[i];
    }
    return o;
  }

  std::unique_ptr<XmlElem> parse_elem() {
    if (eof() || cur() != '<') return nullptr;
    adv();
    auto e = std::make_unique<XmlElem>();
    e->line = line_;
    e->name = parse_name();
    for (;;) {
      skip_ws();
      if (eof()) fail("unterminated tag <" + e->name);
      if (cur() == '/') {
        adv();
        if (eof() || cur() != '>') fail("expected '>' after '/'");
        adv();
        return e;
      }
      if (cur() == '>') { adv(); break; }
      std::string key = parse_name();
      skip_ws();
      if (eof() || cur() != '=') fail("expected '=' after attribute " + key);
      adv();
      skip_ws();
      if (eof() || (cur() != '"' && cur() != '\'')) fail("expected quoted value for " + key);
      char q = cur();
      adv();
      size_t b = p_;
      while (!eof() && cur() != q) adv();
      if (eof()) fail("unterminated attribute value");
      e->attrs.emplace_back(key, unescape(s_.substr(b, p_ - b)));
      adv();
    }
    // children until the matching close tag
    for (;;) {
      skip_misc();
      if (eof()) fail("missing </" + e->name + ">");
      if (starts("</")) {
        adv(); adv();
        std::string n = parse_name();
        if (n != e->name) fail("mismatched close tag </" + n + "> for <" + e->name + ">");
        skip_ws();
        if (eof() || cur() != '>') fail("expected '>'");
        adv();
        return e;
      }
      e->children.push_back(parse_elem());
    }
  }
};

}  // namespace b2
