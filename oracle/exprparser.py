"""Oracle restatement of Marlin's expression parser (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows the PEG grammar in include/utils/MarlinExpressionParser.h:383-427 and the
simplify / differentiate / substitute / toString rules in
src/utils/MarlinExpressionParser.C:51-141 (BinaryOp::simplify), :144-203 (BinaryOp d/dx),
:252-309 (UnaryOp), :317-505 (Comparison, LogicalOp), :517-601 (FunctionCall::simplify),
:604-860 (FunctionCall::differentiate), :972-1104 (LetExpression), and the evaluation
semantics of the TorchScript graph built in :213-240, :862-952 (aten op per node).

AST nodes are plain tuples:
  ("num", v) ("var", name) ("const", name) ("bin", op, l, r) ("un", op, x)
  ("cmp", op, l, r) ("log", op, l, r) ("call", name, [args]) ("let", [(name, expr)], body)
"""
import math
import re

import torch

_TOKEN = re.compile(
    r"\s*(?:(?P<num>[0-9]+(?:\.[0-9]+)?(?:[eE][+-]?[0-9]+)?)|(?P<id>[a-zA-Z_][a-zA-Z0-9_]*)"
    r"|(?P<op>:=|<=|>=|==|!=|[-+*/%^()<>,;!&|]))"
)


class ParseError(ValueError):
    pass


def _tokenize(s):
    pos, out = 0, []
    while True:
        m = re.compile(r"\s*").match(s, pos)
        pos = m.end()
        if pos >= len(s):
            break
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise ParseError(f"syntax error at column {pos + 1}: {s[pos:pos + 10]!r}")
        if m.group("num") is not None:
            out.append(("num", m.group("num"), m.start("num"), m.end()))
        elif m.group("id") is not None:
            out.append(("id", m.group("id"), m.start("id"), m.end()))
        else:
            out.append(("op", m.group("op"), m.start("op"), m.end()))
        pos = m.end()
    out.append(("end", "", len(s), len(s)))
    return out


class _P:
    """Recursive descent over the token list; one method per PEG rule."""

    def __init__(self, text, constants):
        self.t = _tokenize(text)
        self.i = 0
        self.constants = set(constants)
        self.text = text

    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def isop(self, *ops, k=0):
        tk = self.peek(k)
        return tk[0] == "op" and tk[1] in ops

    def eat(self):
        tk = self.t[self.i]
        self.i += 1
        return tk

    def expect(self, op):
        if not self.isop(op):
            tk = self.peek()
            raise ParseError(f"Line 1:{tk[2] + 1}: syntax error, expecting '{op}'.")
        self.eat()

    def statements(self):
        bindings = []
        # ASSIGNMENT <- IDENTIFIER ':=' LOGICAL, each followed by ';'
        while self.peek()[0] == "id" and self.isop(":=", k=1):
            name = self.eat()[1]
            self.eat()
            e = self.logical()
            self.expect(";")
            bindings.append((name, e))
        body = self.logical()
        if self.peek()[0] != "end":
            tk = self.peek()
            raise ParseError(f"Line 1:{tk[2] + 1}: syntax error, unexpected '{tk[1]}'.")
        return ("let", bindings, body) if bindings else body

    def logical(self):
        e = self.comparison()
        while self.isop("|", "&"):
            op = self.eat()[1]
            e = ("log", op, e, self.comparison())
        return e

    def comparison(self):
        e = self.additive()
        if self.isop("<=", ">=", "==", "!=", "<", ">"):
            op = self.eat()[1]
            e = ("cmp", op, e, self.additive())
        return e

    def additive(self):
        e = self.multitive()
        while self.isop("+", "-"):
            op = self.eat()[1]
            e = ("bin", op, e, self.multitive())
        return e

    def multitive(self):
        e = self.unary()
        while self.isop("*", "/", "%"):
            op = self.eat()[1]
            e = ("bin", op, e, self.unary())
        return e

    def unary(self):
        if self.isop("-", "!"):
            op = self.eat()[1]
            return ("un", op, self.unary())
        return self.power()

    def power(self):
        e = self.primary()
        if self.isop("^"):
            self.eat()
            e = ("bin", "^", e, self.power())  # right associative; exponent is POWER, not UNARY
        return e

    def primary(self):
        tk = self.peek()
        if tk[0] == "id":
            if self.isop("(", k=1):
                name = self.eat()[1]
                self.eat()
                args = []
                if not self.isop(")"):
                    args.append(self.logical())
                    while self.isop(","):
                        self.eat()
                        args.append(self.logical())
                self.expect(")")
                return ("call", name, args)
            if self.isop(":=", k=1):
                raise ParseError(f"Line 1:{tk[2] + 1}: syntax error, unexpected ':='.")
            self.eat()
            return ("const", tk[1]) if tk[1] in self.constants else ("var", tk[1])
        if tk[0] == "num":
            self.eat()
            return ("num", float(tk[1]))
        if self.isop("("):
            self.eat()
            e = self.logical()
            self.expect(")")
            return e
        what = f"unexpected '{tk[1]}', " if tk[0] != "end" else ""
        raise ParseError(
            f"Line 1:{tk[2] + 1}: syntax error, {what}expecting <IDENTIFIER>, <NUMBER>, '('."
        )


def parse(text, constants=()):
    return _P(text, constants).statements()


# ----------------------------------------------------------------------------- toString
def to_string(e):
    k = e[0]
    if k == "num":
        return "%f" % e[1]  # std::to_string(double)
    if k in ("var", "const"):
        return e[1]
    if k in ("bin", "cmp", "log"):
        return "(" + to_string(e[2]) + " " + e[1] + " " + to_string(e[3]) + ")"
    if k == "un":
        return "(" + e[1] + to_string(e[2]) + ")"
    if k == "call":
        return e[1] + "(" + ", ".join(to_string(a) for a in e[2]) + ")"
    if k == "let":
        return "".join(n + ":=" + to_string(x) + "; " for n, x in e[1]) + to_string(e[2])
    raise AssertionError(k)


# ----------------------------------------------------------------------------- simplify
def _num(v):
    return ("num", float(v))


def _isnum(e, v=None):
    return e[0] == "num" and (v is None or e[1] == v)


def _cdiv(a, b):
    if b == 0.0:
        if a == 0.0 or a != a:
            return float("nan")
        return math.copysign(float("inf"), a) * math.copysign(1.0, b)
    return a / b


def _cpow(a, b):
    try:
        return math.pow(a, b)
    except (OverflowError, ValueError):
        return float(torch.pow(torch.tensor(a, dtype=torch.float64), b))


_FOLD1 = {
    "sin": math.sin, "cos": math.cos, "tan": math.tan, "sinh": math.sinh, "cosh": math.cosh,
    "tanh": math.tanh, "asin": math.asin, "acos": math.acos, "atan": math.atan,
    "asinh": math.asinh, "acosh": math.acosh, "atanh": math.atanh, "exp": math.exp,
    "log": math.log, "log10": math.log10, "log2": math.log2, "sqrt": math.sqrt, "abs": abs,
    "ceil": lambda v: float(math.ceil(v)), "floor": lambda v: float(math.floor(v)),
    "round": lambda v: math.copysign(math.floor(abs(v) + 0.5), v),  # std::round: half away
    "trunc": lambda v: float(math.trunc(v)),
}
_FOLD2 = {
    "min": min, "max": max, "atan2": math.atan2, "hypot": math.hypot, "pow": _cpow,
}


def simplify(e):
    k = e[0]
    if k in ("num", "var", "const"):
        return e
    if k == "bin":
        op, l, r = e[1], simplify(e[2]), simplify(e[3])
        if _isnum(l) and _isnum(r):
            a, b = l[1], r[1]
            if op == "+":
                return _num(a + b)
            if op == "-":
                return _num(a - b)
            if op == "*":
                return _num(a * b)
            if op == "/":
                return _num(_cdiv(a, b))
            if op == "^":
                return _num(_cpow(a, b))
            if op == "%":
                return _num(math.fmod(a, b) if b != 0 else float("nan"))
        if op == "+":
            if _isnum(l, 0.0):
                return r
            if _isnum(r, 0.0):
                return l
        elif op == "-":
            if _isnum(r, 0.0):
                return l
            if _isnum(l, 0.0):
                return simplify(("un", "-", r))
        elif op == "*":
            if _isnum(l, 0.0) or _isnum(r, 0.0):
                return _num(0.0)
            if _isnum(l, 1.0):
                return r
            if _isnum(r, 1.0):
                return l
            if _isnum(l, -1.0):
                return simplify(("un", "-", r))
            if _isnum(r, -1.0):
                return simplify(("un", "-", l))
        elif op == "/":
            if _isnum(l, 0.0):
                return _num(0.0)
            if _isnum(r, 1.0):
                return l
        elif op == "^":
            if _isnum(r, 0.0):
                return _num(1.0)
            if _isnum(r, 1.0):
                return l
            if _isnum(l, 1.0):
                return _num(1.0)
        return ("bin", op, l, r)
    if k == "un":
        x = simplify(e[2])
        if _isnum(x):
            return _num(-x[1]) if e[1] == "-" else _num(1.0 if x[1] == 0.0 else 0.0)
        return ("un", e[1], x)
    if k == "cmp":
        l, r = simplify(e[2]), simplify(e[3])
        if _isnum(l) and _isnum(r):
            a, b = l[1], r[1]
            res = {"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b, "==": a == b, "!=": a != b}
            return _num(1.0 if res[e[1]] else 0.0)
        return ("cmp", e[1], l, r)
    if k == "log":
        l, r = simplify(e[2]), simplify(e[3])
        if _isnum(l) and _isnum(r):
            a, b = l[1] != 0.0, r[1] != 0.0
            return _num(1.0 if ((a and b) if e[1] == "&" else (a or b)) else 0.0)
        if e[1] == "&":
            if _isnum(l, 0.0) or _isnum(r, 0.0):
                return _num(0.0)
        else:
            if (_isnum(l) and l[1] != 0.0) or (_isnum(r) and r[1] != 0.0):
                return _num(1.0)
        return ("log", e[1], l, r)
    if k == "call":
        args = [simplify(a) for a in e[2]]
        if all(_isnum(a) for a in args):
            v = [a[1] for a in args]
            try:
                if e[1] in _FOLD1 and len(v) == 1:
                    return _num(_FOLD1[e[1]](v[0]))
                if e[1] in _FOLD2 and len(v) == 2:
                    return _num(_FOLD2[e[1]](v[0], v[1]))
            except (ValueError, OverflowError):
                return _num(float("nan"))
            if e[1] == "if" and len(v) == 3:
                return _num(v[1] if v[0] != 0.0 else v[2])
        return ("call", e[1], args)
    if k == "let":
        b = [(n, simplify(x)) for n, x in e[1]]
        body = simplify(e[2])
        return ("let", b, body) if b else body
    raise AssertionError(k)


# ----------------------------------------------------------------------------- substitute
def substitute(e, var, rep):
    k = e[0]
    if k == "var":
        return rep if e[1] == var else e
    if k in ("num", "const"):
        return e
    if k in ("bin", "cmp", "log"):
        return (k, e[1], substitute(e[2], var, rep), substitute(e[3], var, rep))
    if k == "un":
        return (k, e[1], substitute(e[2], var, rep))
    if k == "call":
        return (k, e[1], [substitute(a, var, rep) for a in e[2]])
    if k == "let":
        b = [(n, substitute(x, var, rep)) for n, x in e[1]]
        shadowed = any(n == var for n, _ in e[1])
        return ("let", b, e[2] if shadowed else substitute(e[2], var, rep))
    raise AssertionError(k)


# ----------------------------------------------------------------------------- differentiate
def _b(op, l, r):
    return ("bin", op, l, r)


def _call(name, *args):
    return ("call", name, list(args))


def differentiate(e, var):
    k = e[0]
    D = lambda x: differentiate(x, var)  # noqa: E731
    if k in ("num", "const"):
        return _num(0.0)
    if k == "var":
        return _num(1.0 if e[1] == var else 0.0)
    if k == "bin":
        op, l, r = e[1], e[2], e[3]
        dl, dr = D(l), D(r)
        if op in "+-":
            return _b(op, dl, dr)
        if op == "*":
            return _b("+", _b("*", dl, r), _b("*", l, dr))
        if op == "/":
            return _b("/", _b("-", _b("*", dl, r), _b("*", l, dr)), _b("^", r, _num(2.0)))
        if op == "^":
            if _isnum(r):
                return _b("*", _b("*", r, _b("^", l, _num(r[1] - 1.0))), dl)
            return _b("*", _b("^", l, r),
                      _b("+", _b("*", dr, _call("log", l)), _b("*", r, _b("/", dl, l))))
        if op == "%":
            return dl
    if k == "un":
        return ("un", "-", D(e[2])) if e[1] == "-" else _num(0.0)
    if k in ("cmp", "log"):
        return _num(0.0)
    if k == "call":
        name, args = e[1], e[2]
        if not args:
            return _num(0.0)
        a = args[0]
        da = D(a)
        one, two = _num(1.0), _num(2.0)
        if name == "sin":
            return _b("*", _call("cos", a), da)
        if name == "cos":
            return _b("*", ("un", "-", _call("sin", a)), da)
        if name == "tan":
            c = _call("cos", a)
            return _b("/", da, _b("*", c, c))
        if name == "sinh":
            return _b("*", _call("cosh", a), da)
        if name == "cosh":
            return _b("*", _call("sinh", a), da)
        if name == "tanh":
            c = _call("cosh", a)
            return _b("/", da, _b("*", c, c))
        if name == "exp":
            return _b("*", _call("exp", a), da)
        if name == "exp2":
            return _b("*", _b("*", _call("exp2", a), _call("log", two)), da)
        if name == "log":
            return _b("/", da, a)
        if name == "log10":
            return _b("/", da, _b("*", a, _call("log", _num(10.0))))
        if name == "log2":
            return _b("/", da, _b("*", a, _call("log", two)))
        if name == "sqrt":
            return _b("/", da, _b("*", two, _call("sqrt", a)))
        if name == "rsqrt":
            return _b("*", ("un", "-", _b("/", _call("rsqrt", a), _b("*", two, a))), da)
        if name == "asin":
            return _b("/", da, _call("sqrt", _b("-", one, _b("*", a, a))))
        if name == "acos":
            return ("un", "-", _b("/", da, _call("sqrt", _b("-", one, _b("*", a, a)))))
        if name == "atan":
            return _b("/", da, _b("+", one, _b("*", a, a)))
        if name == "asinh":
            return _b("/", da, _call("sqrt", _b("+", _b("*", a, a), one)))
        if name == "acosh":
            return _b("/", da, _call("sqrt", _b("-", _b("*", a, a), one)))
        if name == "atanh":
            return _b("/", da, _b("-", one, _b("*", a, a)))
        if name == "abs":
            return _b("*", _b("/", a, e), da)
        if len(args) == 2:
            a2 = args[1]
            da2 = D(a2)
            if name == "hypot":
                return _b("+", _b("*", _b("/", a, e), da), _b("*", _b("/", a2, e), da2))
            if name == "atan2":  # atan2(y, x)
                return _b("/", _b("-", _b("*", a2, da), _b("*", a, da2)),
                          _b("+", _b("*", a2, a2), _b("*", a, a)))
            if name == "pow":
                return _b("*", e, _b("+", _b("*", a2, _b("/", da, a)),
                                     _b("*", _call("log", a), da2)))
            if name == "min":
                return _call("if", ("cmp", "<", a, a2), da, da2)
            if name == "max":
                return _call("if", ("cmp", ">", a, a2), da, da2)
        if name == "if" and len(args) == 3:
            return _call("if", args[0], D(args[1]), D(args[2]))
        if name in ("round", "ceil", "floor", "trunc"):
            return _num(0.0)
        raise ValueError("Derivative not implemented for function: " + name)
    if k == "let":
        newb, names = [], []
        for n, x in e[1]:
            newb.append((n, x))
            dx = D(x)
            for earlier in names:
                part = differentiate(x, earlier)
                if _isnum(part, 0.0):
                    continue
                dx = _b("+", dx, _b("*", part, ("var", "d" + earlier)))
            newb.append(("d" + n, dx))
            names.append(n)
        dbody = D(e[2])
        for n, _ in e[1]:
            part = differentiate(e[2], n)
            if _isnum(part, 0.0):
                continue
            dbody = _b("+", dbody, _b("*", part, ("var", "d" + n)))
        return ("let", newb, dbody)
    raise AssertionError(k)


# ----------------------------------------------------------------------------- evaluation
_T1 = {
    "sin": torch.sin, "cos": torch.cos, "tan": torch.tan, "sinh": torch.sinh,
    "cosh": torch.cosh, "tanh": torch.tanh, "asin": torch.asin, "acos": torch.acos,
    "atan": torch.atan, "asinh": torch.asinh, "acosh": torch.acosh, "atanh": torch.atanh,
    "exp": torch.exp, "exp2": torch.exp2, "log": torch.log, "log10": torch.log10,
    "log2": torch.log2, "sqrt": torch.sqrt, "rsqrt": torch.rsqrt, "abs": torch.abs,
    "ceil": torch.ceil, "floor": torch.floor, "round": torch.round, "trunc": torch.trunc,
}
_T2 = {"min": torch.minimum, "max": torch.maximum, "atan2": torch.atan2, "hypot": torch.hypot,
       "pow": torch.pow}


def _t(v):
    return v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=torch.float64)


def evaluate(e, env):
    """Evaluate with aten semantics; python floats play the role of JIT scalar constants
    (src/utils/MarlinExpressionParser.C:213-240, :862-952)."""
    k = e[0]
    if k == "num":
        return e[1]
    if k in ("var", "const"):
        if e[1] not in env:
            raise KeyError(f"Variable '{e[1]}' not found in variable list")
        return env[e[1]]
    if k == "bin":
        a, b = evaluate(e[2], env), evaluate(e[3], env)
        op = e[1]
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
                return _cdiv(a, b)
            return a / b
        if op == "^":
            if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
                return _cpow(a, b)
            return torch.pow(a, b)
        if op == "%":
            return torch.remainder(_t(a), b) if not isinstance(b, torch.Tensor) else \
                torch.remainder(a, b)
    if k == "un":
        a = evaluate(e[2], env)
        if e[1] == "-":
            return -a
        return torch.logical_not(_t(a))
    if k == "cmp":
        a, b = evaluate(e[2], env), evaluate(e[3], env)
        if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
            a = _t(a)
        return {"<": lambda: a < b, ">": lambda: a > b, "<=": lambda: a <= b,
                ">=": lambda: a >= b, "==": lambda: a == b, "!=": lambda: a != b}[e[1]]()
    if k == "log":
        a, b = _t(evaluate(e[2], env)), _t(evaluate(e[3], env))
        return torch.logical_and(a, b) if e[1] == "&" else torch.logical_or(a, b)
    if k == "call":
        name = e[1]
        v = [evaluate(a, env) for a in e[2]]
        if name in _T1 and len(v) == 1:
            return _T1[name](_t(v[0]))
        if name in _T2 and len(v) == 2:
            if name == "pow":
                if isinstance(v[0], torch.Tensor) or isinstance(v[1], torch.Tensor):
                    return torch.pow(v[0], v[1])
                return _cpow(v[0], v[1])
            return _T2[name](_t(v[0]), _t(v[1]))
        if name == "if" and len(v) == 3:
            return torch.where(_t(v[0]).bool() if not (isinstance(v[0], torch.Tensor)
                                                        and v[0].dtype == torch.bool) else v[0],
                               v[1], v[2])
        raise ValueError("Unknown or unsupported function: " + name)
    if k == "let":
        scope = dict(env)
        for n, x in e[1]:
            scope[n] = evaluate(x, scope)
        return evaluate(e[2], scope)
    raise AssertionError(k)


class ParsedTensor:
    """Restates ParsedJITTensor (src/utils/ParsedJITTensor.C:22-156): parse, differentiate,
    compile (= simplify), eval."""

    def __init__(self, expression, variables, constants=None):
        self.variables = list(variables)
        self.constants = dict(constants or {})
        self.ast = parse(expression, self.constants.keys())

    def differentiate(self, var):
        self.ast = differentiate(self.ast, var)

    def compile(self):
        self.ast = simplify(self.ast)

    def eval(self, params):
        if len(params) != len(self.variables):
            raise ValueError("Parameter count mismatch")
        env = dict(zip(self.variables, params))
        env.update(self.constants)
        out = evaluate(self.ast, env)
        if not isinstance(out, torch.Tensor):  # all-constant expression -> 0-d tensor
            out = torch.tensor(float(out), dtype=torch.float64)
        return out

    def __str__(self):
        return to_string(self.ast)
