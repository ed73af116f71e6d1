"""zig2py.parser — tokenizer + recursive-descent parser for the Zig subset the reference is written in.

TEST INFRASTRUCTURE.  Nothing of the reference is stored here: the parser is applied to the reference's
source text (/root/reference/src/**.zig) at test time (tests/ref_transpile.py).

The subset (everything nsfisis/RayTracingInOneWeekend.zig uses, nothing else): file-level `const` / `fn`,
`struct` / `union(enum)` containers with fields, decls and methods, generic functions returning a type,
`const`/`var` locals, assignment and compound assignment, `if`/`else` (statement and expression), `while`
with a continue expression, `for (slice) |x|`, `switch` on tagged unions with payload captures, `return`,
`continue`, `break`, `defer`, `try`, anonymous and typed struct literals, enum literals, pointers
(`&x`, `p.*`), indexing, calls, `@builtin(...)` calls, and the usual arithmetic/comparison/boolean operators.

AST nodes are plain tuples: (kind, ...).
"""
import re

KEYWORDS = {
    "const", "var", "pub", "fn", "return", "if", "else", "while", "for", "switch", "and", "or", "try", "defer",
    "continue", "break", "struct", "union", "enum", "comptime", "undefined", "true", "false", "anytype",
}

_TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<num>0x[0-9a-fA-F_]+|[0-9][0-9_]*(?:\.[0-9][0-9_]*)?(?:[eE][+-]?[0-9]+)?)
  | (?P<str>"(?:\\.|[^"\\])*")
  | (?P<builtin>@[A-Za-z_][A-Za-z0-9_]*)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>\.\*|\.\?|=>|==|!=|<=|>=|\+=|-=|\*=|/=|%=|<<|>>|[-+*/%=<>!&|^~.,;:(){}\[\]?])
""", re.VERBOSE)


class Tok:
    __slots__ = ("kind", "val", "line")

    def __init__(self, kind, val, line):
        self.kind, self.val, self.line = kind, val, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}@{self.line}"


def tokenize(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = _TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"zig2py: cannot tokenize at line {line}: {src[pos:pos + 30]!r}")
        kind = m.lastgroup
        text = m.group(kind)
        if kind != "ws":
            if kind == "id" and text in KEYWORDS:
                kind = "kw"
            toks.append(Tok(kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(Tok("eof", "", line))
    return toks


ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%="}


class Parser:
    def __init__(self, src, fname="<zig>"):
        self.t = tokenize(src)
        self.i = 0
        self.fname = fname

    # ---- token helpers ---------------------------------------------------------------------------
    @property
    def cur(self):
        return self.t[self.i]

    def peek(self, k=1):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def at(self, val, kind=None):
        c = self.cur
        return c.val == val and c.kind != "str" and (kind is None or c.kind == kind)

    def accept(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.at(val):
            self.err(f"expected {val!r}, got {self.cur.val!r}")
        self.i += 1

    def ident(self):
        if self.cur.kind != "id":
            self.err(f"expected identifier, got {self.cur.val!r}")
        v = self.cur.val
        self.i += 1
        return v

    def err(self, msg):
        raise SyntaxError(f"zig2py {self.fname}:{self.cur.line}: {msg}")

    # ---- file / containers -----------------------------------------------------------------------
    def parse_file(self):
        members = []
        while self.cur.kind != "eof":
            members.append(self.member(file_level=True))
        return ("container", "file", members)

    def member(self, file_level=False):
        """field | [pub] const/var decl | [pub] fn"""
        self.accept("pub")
        if self.at("const") or self.at("var"):
            d = self.decl()
            return d
        if self.at("fn"):
            return self.fn()
        if file_level:
            self.err(f"unexpected token {self.cur.val!r} at file level")
        # container field:  name: Type [= default],
        name = self.ident()
        self.expect(":")
        ty = self.type_expr()
        default = None
        if self.accept("="):
            default = self.expr()
        self.accept(",")
        return ("field", name, ty, default)

    def container_body(self, kind):
        self.expect("{")
        members = []
        while not self.at("}"):
            members.append(self.member())
        self.expect("}")
        return ("container", kind, members)

    def fn(self):
        line = self.cur.line
        self.expect("fn")
        name = self.ident()
        self.expect("(")
        params = []
        while not self.at(")"):
            self.accept("comptime")
            pname = self.ident()
            self.expect(":")
            pty = self.type_expr()
            params.append((pname, pty))
            if not self.accept(","):
                break
        self.expect(")")
        ret = self.type_expr()
        body = self.block()
        return ("fn", name, params, ret, body, line)

    # ---- types -----------------------------------------------------------------------------------
    def type_expr(self):
        if self.accept("!"):
            return ("errunion", self.type_expr())
        if self.accept("?"):
            return ("optional", self.type_expr())
        if self.accept("*"):
            self.accept("const")
            return ("ptr", self.type_expr())
        if self.at("["):
            self.expect("[")
            if self.accept("]"):
                self.accept("const")
                return ("slice", self.type_expr())
            n = self.expr()
            self.expect("]")
            return ("array", n, self.type_expr())
        if self.at("anytype"):
            self.i += 1
            return ("name", "anytype")
        if self.at("struct") or self.at("union"):
            return self.primary()
        # named / dotted / generic call — no struct-literal postfix here
        e = ("name", self.ident())
        while True:
            if self.at(".") and self.peek().kind == "id":
                self.i += 1
                e = ("field", e, self.ident())
            elif self.at("("):
                e = ("call", e, self.call_args())
            else:
                return e

    # ---- statements ------------------------------------------------------------------------------
    def block(self):
        self.expect("{")
        stmts = []
        while not self.at("}"):
            stmts.append(self.stmt())
        self.expect("}")
        return ("block", stmts)

    def decl(self):
        line = self.cur.line
        kw = self.cur.val
        self.i += 1
        name = self.ident()
        ty = None
        if self.accept(":"):
            ty = self.type_expr()
        self.expect("=")
        val = self.expr()
        self.expect(";")
        return ("decl", kw, name, ty, val, line)

    def stmt_or_block(self):
        return self.block() if self.at("{") else self.stmt()

    def stmt(self):
        c = self.cur
        if c.kind == "kw":
            if c.val in ("const", "var"):
                return self.decl()
            if c.val == "return":
                self.i += 1
                val = None if self.at(";") else self.expr()
                self.expect(";")
                return ("return", val, c.line)
            if c.val == "continue":
                self.i += 1
                self.expect(";")
                return ("continue",)
            if c.val == "break":
                self.i += 1
                self.expect(";")
                return ("break",)
            if c.val == "defer":
                self.i += 1
                inner = self.stmt_or_block()
                return ("defer", inner)
            if c.val == "if":
                self.i += 1
                self.expect("(")
                cond = self.expr()
                self.expect(")")
                then = self.stmt_or_block()
                other = None
                if self.accept("else"):
                    other = self.stmt_or_block()
                return ("if", cond, then, other)
            if c.val == "while":
                self.i += 1
                self.expect("(")
                cond = self.expr()
                self.expect(")")
                cont = None
                if self.accept(":"):
                    self.expect("(")
                    cont = self.assign_or_expr()
                    self.expect(")")
                body = self.block()
                return ("while", cond, cont, body)
            if c.val == "for":
                self.i += 1
                self.expect("(")
                it = self.expr()
                self.expect(")")
                self.expect("|")
                cap = self.ident()
                self.expect("|")
                body = self.block()
                return ("for", it, cap, body)
        if self.at("{"):
            return self.block()
        s = self.assign_or_expr()
        self.expect(";")
        return s

    def assign_or_expr(self):
        line = self.cur.line
        lhs = self.expr()
        if self.cur.kind == "op" and self.cur.val in ASSIGN_OPS:
            op = self.cur.val
            self.i += 1
            rhs = self.expr()
            return ("assign", op, lhs, rhs, line)
        return ("exprstmt", lhs, line)

    # ---- expressions (precedence climbing) ----------------------------------------------------------
    def expr(self):
        return self.p_or()

    def p_or(self):
        e = self.p_and()
        while self.at("or"):
            self.i += 1
            e = ("bin", "or", e, self.p_and())
        return e

    def p_and(self):
        e = self.p_cmp()
        while self.at("and"):
            self.i += 1
            e = ("bin", "and", e, self.p_cmp())
        return e

    def p_cmp(self):
        e = self.p_bits()
        if self.cur.kind == "op" and self.cur.val in ("==", "!=", "<", ">", "<=", ">="):
            op = self.cur.val
            self.i += 1
            e = ("bin", op, e, self.p_bits())
        return e

    def p_bits(self):
        e = self.p_shift()
        while self.cur.kind == "op" and self.cur.val in ("&", "^", "|") and not (self.cur.val == "|" and self._capture_ahead()):
            op = self.cur.val
            self.i += 1
            e = ("bin", op, e, self.p_shift())
        return e

    def _capture_ahead(self):
        # `|name|` after a switch prong / for header is a capture, never a bit-or in this subset
        return self.peek().kind == "id" and self.peek(2).val == "|"

    def p_shift(self):
        e = self.p_add()
        while self.cur.kind == "op" and self.cur.val in ("<<", ">>"):
            op = self.cur.val
            self.i += 1
            e = ("bin", op, e, self.p_add())
        return e

    def p_add(self):
        e = self.p_mul()
        while self.cur.kind == "op" and self.cur.val in ("+", "-"):
            op = self.cur.val
            self.i += 1
            e = ("bin", op, e, self.p_mul())
        return e

    def p_mul(self):
        e = self.p_unary()
        while self.cur.kind == "op" and self.cur.val in ("*", "/", "%"):
            op = self.cur.val
            self.i += 1
            e = ("bin", op, e, self.p_unary())
        return e

    def p_unary(self):
        c = self.cur
        if c.kind == "op" and c.val in ("-", "!", "&", "~"):
            self.i += 1
            return ("un", c.val, self.p_unary())
        if c.kind == "kw" and c.val == "try":
            self.i += 1
            return self.p_unary()
        return self.postfix(self.primary())

    def call_args(self):
        self.expect("(")
        args = []
        while not self.at(")"):
            args.append(self.expr())
            if not self.accept(","):
                break
        self.expect(")")
        return args

    def _struct_lit_ahead(self):
        # `{` that opens a typed struct literal: `{}` or `{ .name = ...`
        if not self.at("{"):
            return False
        n1, n2, n3 = self.peek(1), self.peek(2), self.peek(3)
        return n1.val == "}" or (n1.val == "." and n2.kind == "id" and n3.val == "=")

    def postfix(self, e):
        while True:
            if self.at(".") and self.peek().kind in ("id", "kw"):
                self.i += 1
                name = self.cur.val
                self.i += 1
                e = ("field", e, name)
            elif self.at(".*"):
                self.i += 1
                e = ("deref", e)
            elif self.at(".?"):
                self.i += 1
            elif self.at("("):
                e = ("call", e, self.call_args())
            elif self.at("["):
                self.i += 1
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            elif e[0] in ("name", "field", "call") and self._struct_lit_ahead():
                e = ("structlit", e, self.lit_fields())
            else:
                return e

    def lit_fields(self):
        self.expect("{")
        fields = []
        while not self.at("}"):
            self.expect(".")
            name = self.cur.val
            self.i += 1
            self.expect("=")
            fields.append((name, self.expr()))
            if not self.accept(","):
                break
        self.expect("}")
        return fields

    def primary(self):
        c = self.cur
        if c.kind == "num":
            self.i += 1
            return ("num", c.val.replace("_", ""))
        if c.kind == "str":
            self.i += 1
            return ("str", c.val)
        if c.kind == "builtin":
            self.i += 1
            return ("builtin", c.val[1:], self.call_args())
        if c.kind == "id":
            self.i += 1
            return ("name", c.val)
        if c.kind == "kw":
            if c.val in ("true", "false", "undefined"):
                self.i += 1
                return ("lit", c.val)
            if c.val == "if":
                self.i += 1
                self.expect("(")
                cond = self.expr()
                self.expect(")")
                a = self.expr()
                self.expect("else")
                b = self.expr()
                return ("ifexpr", cond, a, b)
            if c.val == "switch":
                return self.switch()
            if c.val == "struct":
                self.i += 1
                return self.container_body("struct")
            if c.val == "union":
                self.i += 1
                self.expect("(")
                self.expect("enum")
                self.expect(")")
                return self.container_body("union")
            if c.val == "anytype":
                self.i += 1
                return ("name", "anytype")
        if self.at("("):
            self.i += 1
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if self.at("."):
            if self.peek().val == "{":
                n2, n3, n4 = self.peek(2), self.peek(3), self.peek(4)
                if n2.val == "}" or (n2.val == "." and n3.kind in ("id", "kw") and n4.val == "="):
                    self.i += 1
                    return ("anonlit", self.lit_fields())
                # tuple literal  .{ a, b }
                self.i += 2
                items = []
                while not self.at("}"):
                    items.append(self.expr())
                    if not self.accept(","):
                        break
                self.expect("}")
                return ("tuple", items)
            if self.peek().kind in ("id", "kw"):
                self.i += 1
                name = self.cur.val
                self.i += 1
                return ("enumlit", name)
        self.err(f"unexpected token {c.val!r} in expression")

    def switch(self):
        self.expect("switch")
        self.expect("(")
        subj = self.expr()
        self.expect(")")
        self.expect("{")
        prongs = []
        while not self.at("}"):
            if self.accept("else"):
                tag = None
            else:
                self.expect(".")
                tag = self.cur.val
                self.i += 1
            self.expect("=>")
            cap = None
            if self.accept("|"):
                cap = self.ident()
                self.expect("|")
            body = self.expr()
            prongs.append((tag, cap, body))
            if not self.accept(","):
                break
        self.expect("}")
        return ("switch", subj, prongs)


def parse(src, fname="<zig>"):
    return Parser(src, fname).parse_file()
