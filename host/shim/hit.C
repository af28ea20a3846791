#include "hit.h"

#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>

#include "marlin_b200.h"

namespace hit {

std::string Node::fullpath() const {
  if (!parent) return "";
  const std::string p = parent->fullpath();
  return p.empty() ? name : p + "/" + name;
}

Node *Node::find(const std::string &path) { return const_cast<Node *>(static_cast<const Node *>(this)->find(path)); }
const Node *Node::find(const std::string &path) const {
  const size_t s = path.find('/');
  const std::string head = path.substr(0, s);
  for (const auto &c : children)
    if (c->name == head) {
      if (s == std::string::npos) return c.get();
      if (const Node *r = c->find(path.substr(s + 1))) return r;
    }
  return nullptr;
}

static std::vector<std::string> split_ws(const std::string &s) {
  std::istringstream is(s);
  std::vector<std::string> out;
  std::string w;
  while (is >> w) out.push_back(w);
  return out;
}

std::vector<Node *> Node::sections() const {
  const Node *act = field("active"), *inact = field("inactive");
  std::vector<std::string> a = act ? split_ws(act->value) : std::vector<std::string>();
  std::vector<std::string> ia = inact ? split_ws(inact->value) : std::vector<std::string>();
  std::vector<Node *> out;
  for (const auto &c : children) {
    if (!c->is_section) continue;
    bool keep = true;
    if (act && !(a.size() == 1 && a[0] == "__all__")) {
      keep = false;
      for (const auto &n : a) keep = keep || n == c->name;
    }
    for (const auto &n : ia) keep = keep && n != c->name;
    if (keep) out.push_back(c.get());
  }
  return out;
}
std::vector<Node *> Node::fields() const {
  std::vector<Node *> out;
  for (const auto &c : children)
    if (!c->is_section) out.push_back(c.get());
  return out;
}
const Node *Node::field(const std::string &key) const {
  const Node *hit = nullptr;
  for (const auto &c : children)
    if (!c->is_section && c->name == key) hit = c.get();  // the last assignment wins
  return hit;
}

namespace {
struct Lexer {
  const std::string &s;
  const std::string &fname;
  size_t i = 0;
  int line = 1;
  Lexer(const std::string &text, const std::string &f) : s(text), fname(f) {}
  [[noreturn]] void fail(const std::string &m) const { throw std::runtime_error(fname + ":" + std::to_string(line) + ": " + m); }
  void skip() {
    while (i < s.size()) {
      if (s[i] == '\n') {
        ++line;
        ++i;
      } else if (std::isspace((unsigned char)s[i])) {
        ++i;
      } else if (s[i] == '#') {
        while (i < s.size() && s[i] != '\n') ++i;
      } else {
        break;
      }
    }
  }
};

Node *add(Node *parent, bool section, const std::string &name, int line) {
  auto n = std::make_unique<Node>();
  n->is_section = section;
  n->name = name;
  n->line = line;
  n->parent = parent;
  parent->children.push_back(std::move(n));
  return parent->children.back().get();
}

// `a/b/c` section headers create (or re-open) nested sections
Node *open_section(Node *cur, const std::string &path, int line) {
  std::string rest = path;
  while (!rest.empty()) {
    const size_t s = rest.find('/');
    const std::string head = rest.substr(0, s);
    rest = s == std::string::npos ? "" : rest.substr(s + 1);
    if (head.empty() || head == ".") continue;
    Node *next = nullptr;
    for (auto &c : cur->children)
      if (c->is_section && c->name == head) next = c.get();
    cur = next ? next : add(cur, true, head, line);
  }
  return cur;
}

void parse_into(Node *root, const std::string &text, const std::string &fname) {
  Lexer lx(text, fname);
  Node *cur = root;
  std::vector<int> depth_stack;  // how many levels each open header pushed
  while (true) {
    lx.skip();
    if (lx.i >= text.size()) break;
    if (text[lx.i] == '[') {
      const size_t e = text.find(']', lx.i);
      if (e == std::string::npos) lx.fail("unterminated section header");
      std::string name = text.substr(lx.i + 1, e - lx.i - 1);
      lx.i = e + 1;
      // trim
      while (!name.empty() && std::isspace((unsigned char)name.back())) name.pop_back();
      while (!name.empty() && std::isspace((unsigned char)name.front())) name.erase(name.begin());
      if (name.empty() || name == "../" || name == "..") {
        if (depth_stack.empty()) lx.fail("unmatched section close '[]'");
        for (int k = 0; k < depth_stack.back(); ++k) cur = cur->parent;
        depth_stack.pop_back();
      } else {
        if (name.rfind("./", 0) == 0) name = name.substr(2);
        Node *before = cur;
        cur = open_section(cur, name, lx.line);
        int d = 0;
        for (Node *p = cur; p != before; p = p->parent) ++d;
        depth_stack.push_back(d);
      }
      continue;
    }
    // field: key (=|:=) value
    size_t k = lx.i;
    while (k < text.size() && (std::isalnum((unsigned char)text[k]) || strchr("_./:<>+-*", text[k])) && !(text[k] == ':' && k + 1 < text.size() && text[k + 1] == '=')) ++k;
    std::string key = text.substr(lx.i, k - lx.i);
    if (key.empty()) lx.fail(std::string("unexpected character '") + text[lx.i] + "'");
    lx.i = k;
    while (lx.i < text.size() && (text[lx.i] == ' ' || text[lx.i] == '\t')) ++lx.i;
    if (lx.i < text.size() && text[lx.i] == ':' && lx.i + 1 < text.size() && text[lx.i + 1] == '=') ++lx.i;
    if (lx.i >= text.size() || text[lx.i] != '=') lx.fail("expected '=' after '" + key + "'");
    ++lx.i;
    while (lx.i < text.size() && (text[lx.i] == ' ' || text[lx.i] == '\t')) ++lx.i;
    const int line = lx.line;
    std::string val;
    bool quoted = false;
    // a value is one or more adjacent quoted strings (concatenated), or a bare run to end of line
    if (lx.i < text.size() && (text[lx.i] == '\'' || text[lx.i] == '"')) {
      quoted = true;
      while (lx.i < text.size() && (text[lx.i] == '\'' || text[lx.i] == '"')) {
        const char q = text[lx.i++];
        const size_t e = text.find(q, lx.i);
        if (e == std::string::npos) lx.fail("unterminated string");
        std::string piece = text.substr(lx.i, e - lx.i);
        for (char ch : piece)
          if (ch == '\n') ++lx.line;
        val += piece;
        lx.i = e + 1;
        size_t j = lx.i;  // look ahead for a continuation string on the following lines
        int nl = 0;
        while (j < text.size() && std::isspace((unsigned char)text[j])) {
          if (text[j] == '\n') ++nl;
          ++j;
        }
        if (j < text.size() && (text[j] == '\'' || text[j] == '"') && nl > 0) {
          lx.line += nl;
          lx.i = j;
        } else {
          break;
        }
      }
    } else {
      size_t e = lx.i;
      int brace = 0;
      while (e < text.size() && (text[e] != '\n' || brace > 0)) {
        if (text[e] == '$' && e + 1 < text.size() && text[e + 1] == '{') ++brace;
        if (text[e] == '}' && brace > 0) --brace;
        if (text[e] == '#' && brace == 0) break;
        if (text[e] == '\n') ++lx.line;
        ++e;
      }
      val = text.substr(lx.i, e - lx.i);
      while (!val.empty() && std::isspace((unsigned char)val.back())) val.pop_back();
      lx.i = e;
    }
    // keys may carry a path: a/b/key = value
    Node *holder = cur;
    const size_t slash = key.rfind('/');
    if (slash != std::string::npos) {
      holder = open_section(cur, key.substr(0, slash), line);
      key = key.substr(slash + 1);
    }
    Node *f = add(holder, false, key, line);
    f->value = val;
    f->quoted = quoted;
  }
  if (!depth_stack.empty()) lx.fail("missing closing '[]' at end of input");
}

// ${...} expansion ----------------------------------------------------------------------
const Node *lookup(const Node *from, const std::string &name, const Node *self = nullptr) {
  // search the enclosing scopes from the innermost outwards (hit semantics)
  if (name.find('/') != std::string::npos) {
    const Node *root = from;
    while (root->parent) root = root->parent;
    const Node *n = root->find(name);
    return (n && !n->is_section) ? n : nullptr;
  }
  // the field being expanded never resolves to itself: `dt = ${dt}` inside a block refers to the
  // enclosing scope's dt (MOOSE inputs rely on this, e.g. [Executioner] dt = ${dt})
  for (const Node *s = from; s; s = s->parent)
    if (const Node *f = s->field(name))
      if (f != self) return f;
  return nullptr;
}

std::string fmt_num(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.17g", v);
  return buf;
}

struct Expander {
  const std::string &fname;
  std::map<const Node *, int> state;  // 1 = in progress, 2 = done
  const Node *current = nullptr;      // field whose value is being expanded
  explicit Expander(const std::string &f) : fname(f) {}

  std::string expand_text(const Node *scope, const std::string &v, int line) {
    std::string out;
    size_t i = 0;
    while (i < v.size()) {
      if (v[i] == '$' && i + 1 < v.size() && v[i + 1] == '{') {
        int depth = 1;
        size_t j = i + 2;
        while (j < v.size() && depth > 0) {
          if (v[j] == '{') ++depth;
          if (v[j] == '}') --depth;
          ++j;
        }
        if (depth) throw std::runtime_error(fname + ":" + std::to_string(line) + ": unterminated '${'");
        out += eval_brace(scope, expand_text(scope, v.substr(i + 2, j - i - 3), line), line);
        i = j;
      } else {
        out += v[i++];
      }
    }
    return out;
  }

  std::string eval_brace(const Node *scope, const std::string &body, int line) {
    std::istringstream is(body);
    std::string cmd;
    is >> cmd;
    auto err = [&](const std::string &m) { return std::runtime_error(fname + ":" + std::to_string(line) + ": " + m); };
    if (cmd == "fparse") {
      std::string expr;
      std::getline(is, expr);
      // names inside the expression refer to other input variables
      std::string resolved;
      for (size_t i = 0; i < expr.size();) {
        if (std::isalpha((unsigned char)expr[i]) || expr[i] == '_') {
          size_t j = i;
          while (j < expr.size() && (std::isalnum((unsigned char)expr[j]) || expr[j] == '_')) ++j;
          const std::string id = expr.substr(i, j - i);
          const bool is_call = j < expr.size() && expr[j] == '(';
          const Node *f = is_call ? nullptr : lookup(scope, id, current);
          resolved += (f && id != "pi" && id != "e") ? "(" + value_of(f) + ")" : id;
          i = j;
        } else {
          resolved += expr[i++];
        }
      }
      double v = 0;
      if (mrl_expr_constant(resolved.c_str(), 0, nullptr, nullptr, &v)) throw err("fparse '" + expr + "': " + mrl_last_error());
      return fmt_num(v);
    }
    if (cmd == "units") {  // ${units 1.5 m -> mm}: unit conversion is not needed by the Marlin inputs; keep the number
      std::string num;
      is >> num;
      return num;
    }
    if (cmd == "raw") {
      std::string w, out;
      while (is >> w) out += w;
      return out;
    }
    if (cmd == "env") {
      std::string n;
      is >> n;
      const char *e = getenv(n.c_str());
      return e ? e : "";
    }
    std::string extra;
    if (is >> extra) throw err("unknown brace command '" + cmd + "'");
    const Node *f = lookup(scope, cmd, current);
    if (!f) throw err("no variable '" + cmd + "' found for substitution");
    return value_of(f);
  }

  std::string value_of(const Node *f) {
    int &st = state[f];
    if (st == 1) throw std::runtime_error(fname + ":" + std::to_string(f->line) + ": circular variable reference '" + f->name + "'");
    if (st == 0) {
      st = 1;
      const Node *outer = current;
      current = f;
      const std::string v = expand_text(f->parent, f->value, f->line);
      current = outer;
      const_cast<Node *>(f)->value = v;
      state[f] = 2;
    }
    return f->value;
  }

  void run(Node *n) {
    for (auto &c : n->children) {
      if (c->is_section)
        run(c.get());
      else
        value_of(c.get());
    }
  }
};
// `!include other.i` on a line of its own pulls that file in (path relative to the including file)
std::string splice_includes(const std::string &text, const std::string &fname, int depth) {
  if (text.find("!include") == std::string::npos) return text;
  if (depth > 16) throw std::runtime_error(fname + ": !include nested too deeply");
  const size_t sl = fname.rfind('/');
  const std::string dir = sl == std::string::npos ? "" : fname.substr(0, sl + 1);
  std::istringstream is(text);
  std::string line, out;
  int lineno = 0;
  while (std::getline(is, line)) {
    ++lineno;
    const size_t b = line.find_first_not_of(" \t");
    if (b != std::string::npos && line.compare(b, 8, "!include") == 0) {
      std::vector<std::string> w = split_ws(line.substr(b + 8));
      if (w.size() != 1) throw std::runtime_error(fname + ":" + std::to_string(lineno) + ": !include takes one file name");
      const std::string path = w[0][0] == '/' ? w[0] : dir + w[0];
      std::ifstream in(path);
      if (!in) throw std::runtime_error(fname + ":" + std::to_string(lineno) + ": cannot open included file '" + path + "'");
      std::stringstream ss;
      ss << in.rdbuf();
      out += splice_includes(ss.str(), path, depth + 1);
      out += "\n";
    } else {
      out += line;
      out += "\n";
    }
  }
  return out;
}
}  // namespace

std::unique_ptr<Node> parse(const std::string &text, const std::string &fname, const std::vector<std::string> &overrides) {
  auto root = std::make_unique<Node>();
  root->name = "";
  parse_into(root.get(), splice_includes(text, fname, 0), fname);
  for (const std::string &o : overrides) {
    const size_t eq = o.find('=');
    if (eq == std::string::npos) throw std::runtime_error("command line override '" + o + "' is not of the form path/key=value");
    std::string path = o.substr(0, eq), val = o.substr(eq + 1);
    if (path.size() && path.back() == ':') path.pop_back();
    Node *holder = root.get();
    const size_t slash = path.rfind('/');
    if (slash != std::string::npos) {
      holder = open_section(root.get(), path.substr(0, slash), 0);
      path = path.substr(slash + 1);
    }
    Node *f = add(holder, false, path, 0);
    if (val.size() >= 2 && (val.front() == '\'' || val.front() == '"') && val.back() == val.front()) val = val.substr(1, val.size() - 2);
    f->value = val;
  }
  Expander ex(fname);
  ex.run(root.get());
  return root;
}

}  // namespace hit
