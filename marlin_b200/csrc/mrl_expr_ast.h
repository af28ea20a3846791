// Expression front end behind ParsedCompute: parse -> differentiate -> simplify -> (CUDA source).
//
// Grammar re-derived from the PEG text of the reference (include/utils/MarlinExpressionParser.h:383-427):
//   STATEMENTS <- (IDENT ':=' LOGICAL ';')* LOGICAL
//   LOGICAL    <- COMPARISON (('|' / '&') COMPARISON)*
//   COMPARISON <- ADDITIVE (('<=' / '>=' / '==' / '!=' / '<' / '>') ADDITIVE)?
//   ADDITIVE   <- MULTITIVE (('+' / '-') MULTITIVE)*
//   MULTITIVE  <- UNARY (('*' / '/' / '%') UNARY)*
//   UNARY      <- ('-' / '!') UNARY / POWER
//   POWER      <- PRIMARY ('^' POWER)?          (right associative, binds tighter than unary minus)
//   PRIMARY    <- IDENT '(' args ')' / IDENT / NUMBER / '(' LOGICAL ')'
//   NUMBER     <- [0-9]+ ('.' [0-9]+)? ([eE] [+-]? [0-9]+)?      (no leading '.')
// Rewrite rules follow src/utils/MarlinExpressionParser.C:51-141 (binary simplify), :144-203
// (binary derivative), :252-309 (unary), :317-505 (comparison / logical), :517-601 (function
// folding), :604-860 (function derivatives), :972-1104 (let-bindings and their `d<name>`
// derivative bindings).  Host only; no device code in this header.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace mrlx {

struct ParseError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

enum class Kind { Num, Var, Const, Bin, Un, Cmp, Log, Call, Let };

struct Node;
using P = std::shared_ptr<const Node>;
struct Node {
  Kind k;
  double v = 0;                                  // Num
  std::string s;                                 // Var/Const name, operator, function name
  std::vector<P> a;                              // operands / arguments / [bindings..., body]
  std::vector<std::string> names;                // Let: binding names (a[i] is binding i, a.back() the body)
};

P num(double v);
P var(const std::string &n);
P cst(const std::string &n);
P bin(const std::string &op, P l, P r);
P un(const std::string &op, P x);
P cmp(const std::string &op, P l, P r);
P lgc(const std::string &op, P l, P r);
P call(const std::string &f, std::vector<P> args);
P let(std::vector<std::string> names, std::vector<P> vals, P body);

P parse(const std::string &text, const std::set<std::string> &constants);
std::string to_string(const P &e);
P simplify(const P &e);
P substitute(const P &e, const std::string &v, const P &rep);
P differentiate(const P &e, const std::string &v);
// scalar evaluation with IEEE / libm semantics (constant_expressions, all-constant expressions)
double eval_scalar(const P &e, const std::map<std::string, double> &env);
void collect_symbols(const P &e, std::set<std::string> &out);

}  // namespace mrlx
