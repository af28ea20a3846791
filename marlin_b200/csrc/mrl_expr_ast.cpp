// Expression front end: tokenizer, recursive-descent parser, printer, simplifier,
// substitution, symbolic differentiation and scalar evaluation.  See mrl_expr_ast.h for the
// grammar and the reference rules each function follows.
#include "mrl_expr_ast.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>

namespace mrlx {

// ------------------------------------------------------------------------------ constructors
[[maybe_unused]] static P mk(Kind k) {
  auto n = std::make_shared<Node>();
  n->k = k;
  return n;
}
P num(double v) {
  auto n = std::make_shared<Node>();
  n->k = Kind::Num;
  n->v = v;
  return n;
}
static P named(Kind k, const std::string &s, std::vector<P> a = {}) {
  auto n = std::make_shared<Node>();
  n->k = k;
  n->s = s;
  n->a = std::move(a);
  return n;
}
P var(const std::string &n) { return named(Kind::Var, n); }
P cst(const std::string &n) { return named(Kind::Const, n); }
P bin(const std::string &op, P l, P r) { return named(Kind::Bin, op, {l, r}); }
P un(const std::string &op, P x) { return named(Kind::Un, op, {x}); }
P cmp(const std::string &op, P l, P r) { return named(Kind::Cmp, op, {l, r}); }
P lgc(const std::string &op, P l, P r) { return named(Kind::Log, op, {l, r}); }
P call(const std::string &f, std::vector<P> args) { return named(Kind::Call, f, std::move(args)); }
P let(std::vector<std::string> names, std::vector<P> vals, P body) {
  auto n = std::make_shared<Node>();
  n->k = Kind::Let;
  n->names = std::move(names);
  n->a = std::move(vals);
  n->a.push_back(body);
  return n;
}
static bool is_num(const P &e) { return e->k == Kind::Num; }
static bool is_num(const P &e, double v) { return e->k == Kind::Num && e->v == v; }

// ------------------------------------------------------------------------------ tokenizer
namespace {
struct Tok {
  enum T { NUM, ID, OP, END } t;
  std::string s;
  size_t pos;
};

std::vector<Tok> tokenize(const std::string &s) {
  std::vector<Tok> out;
  size_t i = 0;
  const size_t n = s.size();
  auto digit = [&](size_t j) { return j < n && s[j] >= '0' && s[j] <= '9'; };
  while (true) {
    while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
    if (i >= n) break;
    const char c = s[i];
    if (digit(i)) {
      size_t j = i;
      while (digit(j)) ++j;
      if (j < n && s[j] == '.' && digit(j + 1)) {
        ++j;
        while (digit(j)) ++j;
      }
      if (j < n && (s[j] == 'e' || s[j] == 'E')) {
        size_t k = j + 1;
        if (k < n && (s[k] == '+' || s[k] == '-')) ++k;
        if (digit(k)) {
          while (digit(k)) ++k;
          j = k;
        }
      }
      out.push_back({Tok::NUM, s.substr(i, j - i), i});
      i = j;
    } else if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') {
      size_t j = i;
      while (j < n && ((s[j] >= 'a' && s[j] <= 'z') || (s[j] >= 'A' && s[j] <= 'Z') || s[j] == '_' || digit(j))) ++j;
      out.push_back({Tok::ID, s.substr(i, j - i), i});
      i = j;
    } else {
      static const char *two[] = {":=", "<=", ">=", "==", "!="};
      std::string op;
      for (const char *t : two)
        if (s.compare(i, 2, t) == 0) op = t;
      if (op.empty()) {
        if (std::string("-+*/%^()<>,;!&|").find(c) == std::string::npos)
          throw ParseError("Line 1:" + std::to_string(i + 1) + ": syntax error, unexpected '" + std::string(1, c) + "'.");
        op = std::string(1, c);
      }
      out.push_back({Tok::OP, op, i});
      i += op.size();
    }
  }
  out.push_back({Tok::END, "", n});
  return out;
}

struct Parser {
  std::vector<Tok> t;
  size_t i = 0;
  const std::set<std::string> &constants;
  Parser(const std::string &text, const std::set<std::string> &c) : t(tokenize(text)), constants(c) {}

  const Tok &peek(size_t k = 0) const { return t[std::min(i + k, t.size() - 1)]; }
  bool isop(const char *op, size_t k = 0) const { return peek(k).t == Tok::OP && peek(k).s == op; }
  bool isany(std::initializer_list<const char *> ops) const {
    for (const char *o : ops)
      if (isop(o)) return true;
    return false;
  }
  Tok eat() { return t[i++]; }
  [[noreturn]] void fail(const Tok &tk, const std::string &what) const {
    throw ParseError("Line 1:" + std::to_string(tk.pos + 1) + ": syntax error, " + what);
  }
  void expect(const char *op) {
    if (!isop(op)) fail(peek(), std::string("expecting '") + op + "'.");
    eat();
  }

  P statements() {
    std::vector<std::string> names;
    std::vector<P> vals;
    while (peek().t == Tok::ID && isop(":=", 1)) {
      names.push_back(eat().s);
      eat();
      vals.push_back(logical());
      expect(";");
    }
    P body = logical();
    if (peek().t != Tok::END) fail(peek(), "unexpected '" + peek().s + "'.");
    return names.empty() ? body : let(names, vals, body);
  }
  P logical() {
    P e = comparison();
    while (isany({"|", "&"})) {
      const std::string op = eat().s;
      e = lgc(op, e, comparison());
    }
    return e;
  }
  P comparison() {
    P e = additive();
    if (isany({"<=", ">=", "==", "!=", "<", ">"})) {
      const std::string op = eat().s;
      e = cmp(op, e, additive());
    }
    return e;
  }
  P additive() {
    P e = multitive();
    while (isany({"+", "-"})) {
      const std::string op = eat().s;
      e = bin(op, e, multitive());
    }
    return e;
  }
  P multitive() {
    P e = unary();
    while (isany({"*", "/", "%"})) {
      const std::string op = eat().s;
      e = bin(op, e, unary());
    }
    return e;
  }
  P unary() {
    if (isany({"-", "!"})) {
      const std::string op = eat().s;
      return un(op, unary());
    }
    return power();
  }
  P power() {
    P e = primary();
    if (isop("^")) {
      eat();
      e = bin("^", e, power());
    }
    return e;
  }
  P primary() {
    const Tok tk = peek();
    if (tk.t == Tok::ID) {
      if (isop("(", 1)) {
        eat();
        eat();
        std::vector<P> args;
        if (!isop(")")) {
          args.push_back(logical());
          while (isop(",")) {
            eat();
            args.push_back(logical());
          }
        }
        expect(")");
        return call(tk.s, args);
      }
      if (isop(":=", 1)) fail(tk, "unexpected ':='.");
      eat();
      return constants.count(tk.s) ? cst(tk.s) : var(tk.s);
    }
    if (tk.t == Tok::NUM) {
      eat();
      return num(strtod(tk.s.c_str(), nullptr));
    }
    if (isop("(")) {
      eat();
      P e = logical();
      expect(")");
      return e;
    }
    fail(tk, (tk.t != Tok::END ? "unexpected '" + tk.s + "', " : std::string()) + "expecting <IDENTIFIER>, <NUMBER>, '('.");
  }
};
}  // namespace

P parse(const std::string &text, const std::set<std::string> &constants) {
  Parser p(text, constants);
  return p.statements();
}

// ------------------------------------------------------------------------------ printer
std::string to_string(const P &e) {
  switch (e->k) {
    case Kind::Num: {
      char buf[400];
      snprintf(buf, sizeof buf, "%f", e->v);  // std::to_string(double)
      return buf;
    }
    case Kind::Var:
    case Kind::Const: return e->s;
    case Kind::Bin:
    case Kind::Cmp:
    case Kind::Log: return "(" + to_string(e->a[0]) + " " + e->s + " " + to_string(e->a[1]) + ")";
    case Kind::Un: return "(" + e->s + to_string(e->a[0]) + ")";
    case Kind::Call: {
      std::string r = e->s + "(";
      for (size_t i = 0; i < e->a.size(); ++i) r += (i ? ", " : "") + to_string(e->a[i]);
      return r + ")";
    }
    case Kind::Let: {
      std::string r;
      for (size_t i = 0; i < e->names.size(); ++i) r += e->names[i] + ":=" + to_string(e->a[i]) + "; ";
      return r + to_string(e->a.back());
    }
  }
  return "";
}

// ------------------------------------------------------------------------------ folding helpers
static double round_half_away(double v) { return std::round(v); }
static bool fold1(const std::string &f, double x, double &out) {
  static const std::map<std::string, double (*)(double)> tab = {
      {"sin", std::sin},     {"cos", std::cos},     {"tan", std::tan},     {"sinh", std::sinh},   {"cosh", std::cosh},
      {"tanh", std::tanh},   {"asin", std::asin},   {"acos", std::acos},   {"atan", std::atan},   {"asinh", std::asinh},
      {"acosh", std::acosh}, {"atanh", std::atanh}, {"exp", std::exp},     {"log", std::log},     {"log10", std::log10},
      {"log2", std::log2},   {"sqrt", std::sqrt},   {"abs", std::fabs},    {"ceil", std::ceil},   {"floor", std::floor},
      {"round", round_half_away}, {"trunc", std::trunc}};
  auto it = tab.find(f);
  if (it == tab.end()) return false;
  out = it->second(x);
  return true;
}
static bool fold2(const std::string &f, double a, double b, double &out) {
  if (f == "min") out = std::min(a, b);
  else if (f == "max") out = std::max(a, b);
  else if (f == "atan2") out = std::atan2(a, b);
  else if (f == "hypot") out = std::hypot(a, b);
  else if (f == "pow") out = std::pow(a, b);
  else return false;
  return true;
}

// ------------------------------------------------------------------------------ simplify
P simplify(const P &e) {
  switch (e->k) {
    case Kind::Num:
    case Kind::Var:
    case Kind::Const: return e;
    case Kind::Bin: {
      const std::string &op = e->s;
      P l = simplify(e->a[0]), r = simplify(e->a[1]);
      if (is_num(l) && is_num(r)) {
        const double a = l->v, b = r->v;
        if (op == "+") return num(a + b);
        if (op == "-") return num(a - b);
        if (op == "*") return num(a * b);
        if (op == "/") return num(a / b);
        if (op == "^") return num(std::pow(a, b));
        if (op == "%") return num(std::fmod(a, b));
      }
      if (op == "+") {
        if (is_num(l, 0.0)) return r;
        if (is_num(r, 0.0)) return l;
      } else if (op == "-") {
        if (is_num(r, 0.0)) return l;
        if (is_num(l, 0.0)) return simplify(un("-", r));
      } else if (op == "*") {
        if (is_num(l, 0.0) || is_num(r, 0.0)) return num(0.0);
        if (is_num(l, 1.0)) return r;
        if (is_num(r, 1.0)) return l;
        if (is_num(l, -1.0)) return simplify(un("-", r));
        if (is_num(r, -1.0)) return simplify(un("-", l));
      } else if (op == "/") {
        if (is_num(l, 0.0)) return num(0.0);
        if (is_num(r, 1.0)) return l;
      } else if (op == "^") {
        if (is_num(r, 0.0)) return num(1.0);
        if (is_num(r, 1.0)) return l;
        if (is_num(l, 1.0)) return num(1.0);
      }
      return bin(op, l, r);
    }
    case Kind::Un: {
      P x = simplify(e->a[0]);
      if (is_num(x)) return e->s == "-" ? num(-x->v) : num(x->v == 0.0 ? 1.0 : 0.0);
      return un(e->s, x);
    }
    case Kind::Cmp: {
      P l = simplify(e->a[0]), r = simplify(e->a[1]);
      if (is_num(l) && is_num(r)) {
        const double a = l->v, b = r->v;
        const std::string &op = e->s;
        const bool res = op == "<" ? a < b : op == ">" ? a > b : op == "<=" ? a <= b : op == ">=" ? a >= b : op == "==" ? a == b : a != b;
        return num(res ? 1.0 : 0.0);
      }
      return cmp(e->s, l, r);
    }
    case Kind::Log: {
      P l = simplify(e->a[0]), r = simplify(e->a[1]);
      if (is_num(l) && is_num(r)) {
        const bool a = l->v != 0.0, b = r->v != 0.0;
        return num((e->s == "&" ? (a && b) : (a || b)) ? 1.0 : 0.0);
      }
      if (e->s == "&") {
        if (is_num(l, 0.0) || is_num(r, 0.0)) return num(0.0);
      } else {
        if ((is_num(l) && l->v != 0.0) || (is_num(r) && r->v != 0.0)) return num(1.0);
      }
      return lgc(e->s, l, r);
    }
    case Kind::Call: {
      std::vector<P> args;
      bool all = true;
      for (const P &x : e->a) {
        args.push_back(simplify(x));
        all = all && is_num(args.back());
      }
      if (all && !args.empty()) {
        double out;
        if (args.size() == 1 && fold1(e->s, args[0]->v, out)) return num(out);
        if (args.size() == 2 && fold2(e->s, args[0]->v, args[1]->v, out)) return num(out);
        if (e->s == "if" && args.size() == 3) return num(args[0]->v != 0.0 ? args[1]->v : args[2]->v);
      }
      return call(e->s, args);
    }
    case Kind::Let: {
      std::vector<P> vals;
      for (size_t i = 0; i < e->names.size(); ++i) vals.push_back(simplify(e->a[i]));
      P body = simplify(e->a.back());
      return e->names.empty() ? body : let(e->names, vals, body);
    }
  }
  return e;
}

// ------------------------------------------------------------------------------ substitute
P substitute(const P &e, const std::string &v, const P &rep) {
  switch (e->k) {
    case Kind::Var: return e->s == v ? rep : e;
    case Kind::Num:
    case Kind::Const: return e;
    case Kind::Bin:
    case Kind::Cmp:
    case Kind::Log:
    case Kind::Un:
    case Kind::Call: {
      std::vector<P> a;
      for (const P &x : e->a) a.push_back(substitute(x, v, rep));
      return named(e->k, e->s, a);
    }
    case Kind::Let: {
      std::vector<P> vals;
      bool shadowed = false;
      for (size_t i = 0; i < e->names.size(); ++i) {
        vals.push_back(substitute(e->a[i], v, rep));
        shadowed = shadowed || e->names[i] == v;
      }
      return let(e->names, vals, shadowed ? e->a.back() : substitute(e->a.back(), v, rep));
    }
  }
  return e;
}

// ------------------------------------------------------------------------------ differentiate
P differentiate(const P &e, const std::string &v) {
  auto D = [&](const P &x) { return differentiate(x, v); };
  auto B = [](const char *op, P l, P r) { return bin(op, l, r); };
  auto C1 = [](const char *f, P a) { return call(f, {a}); };
  switch (e->k) {
    case Kind::Num:
    case Kind::Const: return num(0.0);
    case Kind::Var: return num(e->s == v ? 1.0 : 0.0);
    case Kind::Bin: {
      const std::string &op = e->s;
      const P &l = e->a[0], &r = e->a[1];
      P dl = D(l), dr = D(r);
      if (op == "+" || op == "-") return bin(op, dl, dr);
      if (op == "*") return B("+", B("*", dl, r), B("*", l, dr));
      if (op == "/") return B("/", B("-", B("*", dl, r), B("*", l, dr)), B("^", r, num(2.0)));
      if (op == "^") {
        if (is_num(r)) return B("*", B("*", r, B("^", l, num(r->v - 1.0))), dl);
        return B("*", B("^", l, r), B("+", B("*", dr, C1("log", l)), B("*", r, B("/", dl, l))));
      }
      return dl;  // '%'
    }
    case Kind::Un: return e->s == "-" ? un("-", D(e->a[0])) : num(0.0);
    case Kind::Cmp:
    case Kind::Log: return num(0.0);
    case Kind::Call: {
      const std::string &f = e->s;
      if (e->a.empty()) return num(0.0);
      const P &a = e->a[0];
      P da = D(a);
      P one = num(1.0), two = num(2.0);
      if (f == "sin") return B("*", C1("cos", a), da);
      if (f == "cos") return B("*", un("-", C1("sin", a)), da);
      if (f == "tan") {
        P c = C1("cos", a);
        return B("/", da, B("*", c, c));
      }
      if (f == "sinh") return B("*", C1("cosh", a), da);
      if (f == "cosh") return B("*", C1("sinh", a), da);
      if (f == "tanh") {
        P c = C1("cosh", a);
        return B("/", da, B("*", c, c));
      }
      if (f == "exp") return B("*", C1("exp", a), da);
      if (f == "exp2") return B("*", B("*", C1("exp2", a), C1("log", two)), da);
      if (f == "log") return B("/", da, a);
      if (f == "log10") return B("/", da, B("*", a, C1("log", num(10.0))));
      if (f == "log2") return B("/", da, B("*", a, C1("log", two)));
      if (f == "sqrt") return B("/", da, B("*", two, C1("sqrt", a)));
      if (f == "rsqrt") return B("*", un("-", B("/", C1("rsqrt", a), B("*", two, a))), da);
      if (f == "asin") return B("/", da, C1("sqrt", B("-", one, B("*", a, a))));
      if (f == "acos") return un("-", B("/", da, C1("sqrt", B("-", one, B("*", a, a)))));
      if (f == "atan") return B("/", da, B("+", one, B("*", a, a)));
      if (f == "asinh") return B("/", da, C1("sqrt", B("+", B("*", a, a), one)));
      if (f == "acosh") return B("/", da, C1("sqrt", B("-", B("*", a, a), one)));
      if (f == "atanh") return B("/", da, B("-", one, B("*", a, a)));
      if (f == "abs") return B("*", B("/", a, e), da);
      if (e->a.size() == 2) {
        const P &a2 = e->a[1];
        P da2 = D(a2);
        if (f == "hypot") return B("+", B("*", B("/", a, e), da), B("*", B("/", a2, e), da2));
        if (f == "atan2") return B("/", B("-", B("*", a2, da), B("*", a, da2)), B("+", B("*", a2, a2), B("*", a, a)));
        if (f == "pow") return B("*", e, B("+", B("*", a2, B("/", da, a)), B("*", C1("log", a), da2)));
        if (f == "min") return call("if", {cmp("<", a, a2), da, da2});
        if (f == "max") return call("if", {cmp(">", a, a2), da, da2});
      }
      if (f == "if" && e->a.size() == 3) return call("if", {e->a[0], D(e->a[1]), D(e->a[2])});
      if (f == "round" || f == "ceil" || f == "floor" || f == "trunc") return num(0.0);
      throw std::runtime_error("Derivative not implemented for function: " + f);
    }
    case Kind::Let: {
      // every binding b gets a companion binding `db` = total derivative (chain rule through the
      // earlier bindings); the body picks the chain terms up through the same `d<name>` symbols
      std::vector<std::string> names, seen;
      std::vector<P> vals;
      for (size_t i = 0; i < e->names.size(); ++i) {
        const P &x = e->a[i];
        names.push_back(e->names[i]);
        vals.push_back(x);
        P dx = D(x);
        for (const std::string &earlier : seen) {
          P part = differentiate(x, earlier);
          if (is_num(part, 0.0)) continue;
          dx = B("+", dx, B("*", part, var("d" + earlier)));
        }
        names.push_back("d" + e->names[i]);
        vals.push_back(dx);
        seen.push_back(e->names[i]);
      }
      P dbody = D(e->a.back());
      for (const std::string &n : e->names) {
        P part = differentiate(e->a.back(), n);
        if (is_num(part, 0.0)) continue;
        dbody = B("+", dbody, B("*", part, var("d" + n)));
      }
      return let(names, vals, dbody);
    }
  }
  return num(0.0);
}

// ------------------------------------------------------------------------------ scalar evaluation
static double py_remainder(double a, double b) {  // aten::remainder: sign follows the divisor
  double r = std::fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}
double eval_scalar(const P &e, const std::map<std::string, double> &env) {
  auto E = [&](const P &x) { return eval_scalar(x, env); };
  switch (e->k) {
    case Kind::Num: return e->v;
    case Kind::Var:
    case Kind::Const: {
      auto it = env.find(e->s);
      if (it == env.end()) throw std::runtime_error("Variable '" + e->s + "' not found in variable list");
      return it->second;
    }
    case Kind::Bin: {
      const double a = E(e->a[0]), b = E(e->a[1]);
      const std::string &op = e->s;
      if (op == "+") return a + b;
      if (op == "-") return a - b;
      if (op == "*") return a * b;
      if (op == "/") return a / b;
      if (op == "^") return std::pow(a, b);
      return py_remainder(a, b);
    }
    case Kind::Un: {
      const double a = E(e->a[0]);
      return e->s == "-" ? -a : (a == 0.0 ? 1.0 : 0.0);
    }
    case Kind::Cmp: {
      const double a = E(e->a[0]), b = E(e->a[1]);
      const std::string &op = e->s;
      const bool r = op == "<" ? a < b : op == ">" ? a > b : op == "<=" ? a <= b : op == ">=" ? a >= b : op == "==" ? a == b : a != b;
      return r ? 1.0 : 0.0;
    }
    case Kind::Log: {
      const bool a = E(e->a[0]) != 0.0, b = E(e->a[1]) != 0.0;
      return (e->s == "&" ? (a && b) : (a || b)) ? 1.0 : 0.0;
    }
    case Kind::Call: {
      std::vector<double> v;
      for (const P &x : e->a) v.push_back(E(x));
      double out;
      if (v.size() == 1) {
        if (e->s == "exp2") return std::exp2(v[0]);
        if (e->s == "rsqrt") return 1.0 / std::sqrt(v[0]);
        if (e->s == "round") return std::nearbyint(v[0]);  // aten::round: half to even
        if (fold1(e->s, v[0], out)) return out;
      }
      if (v.size() == 2 && fold2(e->s, v[0], v[1], out)) return out;
      if (e->s == "if" && v.size() == 3) return v[0] != 0.0 ? v[1] : v[2];
      throw std::runtime_error("Unknown or unsupported function: " + e->s);
    }
    case Kind::Let: {
      std::map<std::string, double> scope = env;
      for (size_t i = 0; i < e->names.size(); ++i) scope[e->names[i]] = eval_scalar(e->a[i], scope);
      return eval_scalar(e->a.back(), scope);
    }
  }
  return 0.0;
}

void collect_symbols(const P &e, std::set<std::string> &out) {
  if (e->k == Kind::Var || e->k == Kind::Const) out.insert(e->s);
  for (const P &x : e->a) collect_symbols(x, out);
}

}  // namespace mrlx
