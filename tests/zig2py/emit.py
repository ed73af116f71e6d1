"""zig2py.emit — turns the AST of zig2py.parser into Python source that runs on zig2py.runtime.

TEST INFRASTRUCTURE (see parser.py).  The translation is purely mechanical — statement for statement, operator for
operator, in the reference's own evaluation order — so that running the result IS running the reference's
arithmetic in IEEE f64 (Python floats), with these mappings:

  struct / union(enum)      -> Python classes on runtime.Struct / runtime.Union (fields typed lazily)
  .{ .a = x }               -> runtime.Anon, coerced to the type the context declares (field, parameter, return
                               type, typed local, pointee), exactly where Zig's result-location typing applies
  *T, &x, p.*               -> Python object identity for structs; runtime.FieldPtr for `&obj.field`
  var x = y  (struct)       -> a copy (Zig value semantics)
  a / b, @sqrt, @min ...    -> runtime helpers with IEEE semantics (x/0 = inf, sqrt(-1) = nan, minNum)
  switch (u) { .t => |p| e} -> runtime.switch over the union's tag
"""
import keyword

PRIMS = {"f64", "f32", "u8", "u16", "u32", "u64", "i32", "i64", "usize", "isize", "bool", "void", "type", "anytype",
         "comptime_int", "comptime_float"}

BIN_PY = {"and": "and", "or": "or", "==": "==", "!=": "!=", "<": "<", ">": ">", "<=": "<=", ">=": ">=", "+": "+",
          "-": "-", "*": "*", "%": "%", "&": "&", "|": "|", "^": "^", "<<": "<<", ">>": ">>"}


def py_name(n):
    return n + "_" if keyword.iskeyword(n) else n


class Scope:
    def __init__(self, parent=None, container=None):
        self.parent = parent
        self.names = set()
        self.types = {}          # local name -> declared type AST
        self.container = container  # (python class name, set of decl names) when this scope is a container

    def declare(self, n, ty=None):
        self.names.add(n)
        if ty is not None:
            self.types[n] = ty

    def lookup(self, n):
        """-> ('local', None) | ('member', class_name) | None (module global / builtin)"""
        s = self
        while s is not None:
            if s.container is not None:
                if n in s.container[1]:
                    return ("member", s.container[0])
            elif n in s.names:
                return ("local", None)
            s = s.parent
        return None

    def declared_type(self, n):
        s = self
        while s is not None:
            if n in s.names:
                return s.types.get(n)
            s = s.parent
        return None


class Emitter:
    def __init__(self):
        self.lines = []
        self.ind = 0
        self.anon_count = 0
        self.cont_stack = []   # continue-expressions of the enclosing while loops
        self.ret_stack = []    # return type AST of the enclosing functions

    def w(self, s):
        self.lines.append("    " * self.ind + s)

    # ---- containers ------------------------------------------------------------------------------
    def emit_file(self, ast):
        _, _, members = ast
        scope = Scope()
        # pre-declare nothing: file-level names are module globals (resolved at call time)
        for m in members:
            if m[0] == "decl":
                _, kw, name, ty, val, line = m
                if val[0] == "container":
                    self.emit_container(py_name(name), val, scope)
                else:
                    # lazy: imports and aliases may be cyclic between files (vec.zig <-> rand.zig)
                    self.w(f"_lazy({py_name(name)!r}, lambda: {self.typed(ty, val, scope)})")
            elif m[0] == "fn":
                self.emit_fn(m, scope, generic=self.returns_type(m))
            else:
                raise SyntaxError(f"zig2py: unexpected file member {m[0]}")
        return "\n".join(self.lines) + "\n"

    @staticmethod
    def returns_type(fn):
        return fn[3] == ("name", "type")

    def emit_container(self, cls, node, scope):
        _, kind, members = node
        fields = [m for m in members if m[0] == "field"]
        decl_names = {m[2] for m in members if m[0] == "decl"} | {m[1] for m in members if m[0] == "fn"}
        cscope = Scope(scope, container=(cls, decl_names))
        base = "_rt.Struct" if kind == "struct" else "_rt.Union"
        self.w(f"class {cls}({base}):")
        self.ind += 1
        self.w(f"_names = ({''.join(repr(py_name(f[1])) + ', ' for f in fields)})")
        tys = "".join(self.ty(f[2], cscope) + ", " for f in fields)
        self.w(f"_types_thunk = staticmethod(lambda: ({tys}))")
        post = []
        for m in members:
            if m[0] == "decl":
                _, kw, name, ty, val, line = m
                if val[0] == "container":
                    self.emit_container(py_name(name), val, cscope)
                elif val == ("builtin", "This", []):
                    post.append(f"{cls}.{py_name(name)} = {cls}")
                else:
                    self.w(f"{py_name(name)} = {self.typed(ty, val, cscope)}")
            elif m[0] == "fn":
                self.emit_fn(m, cscope, generic=False)
        self.ind -= 1
        for p in post:
            self.w(p)
        self.w("")

    def emit_fn(self, fn, scope, generic):
        _, name, params, ret, body, line = fn
        fscope = Scope(scope)
        pnames = []
        for pname, pty in params:
            pn = py_name(pname)
            if pname == "_":
                pn = f"_unused{len(pnames)}"
            fscope.declare(pname, pty)
            pnames.append(pn)
        if generic:
            self.w("@_rt.generic")
        self.w(f"def {py_name(name)}({', '.join(pnames)}):")
        self.ind += 1
        for (pname, pty), pn in zip(params, pnames):
            if pname == "_" or pty in (("name", "anytype"), ("name", "type")):
                continue
            self.w(f"{pn} = _rt.co({self.ty(pty, fscope)}, {pn})")
        self.ret_stack.append((ret, generic, name))
        n0 = len(self.lines)
        self.emit_block(body, fscope)
        if len(self.lines) == n0:
            self.w("pass")
        self.ret_stack.pop()
        self.ind -= 1
        self.w("")

    # ---- types -----------------------------------------------------------------------------------
    def ty(self, t, scope):
        if t is None:
            return "None"
        k = t[0]
        if k in ("errunion", "optional"):
            return self.ty(t[1], scope)
        if k == "ptr":
            return f"_rt.PtrT({self.ty(t[1], scope)})"
        if k == "slice":
            return f"_rt.SliceT({self.ty(t[1], scope)})"
        if k == "array":
            return f"_rt.ArrT({self.ex(t[1], scope)}, {self.ty(t[2], scope)})"
        return self.ex(t, scope)

    # ---- statements ------------------------------------------------------------------------------
    def emit_block(self, block, scope):
        for s in block[1]:
            self.emit_stmt(s, scope)

    def emit_body(self, node, scope):
        self.ind += 1
        n0 = len(self.lines)
        if node[0] == "block":
            self.emit_block(node, scope)
        else:
            self.emit_stmt(node, scope)
        if len(self.lines) == n0:
            self.w("pass")
        self.ind -= 1

    def typed(self, ty, val, scope):
        """expression `val` evaluated with result type `ty` (None = inferred)"""
        if val == ("lit", "undefined"):
            return f"_rt.undefined({self.ty(ty, scope)})"
        e = self.ex(val, scope)
        if ty is None:
            return e
        return f"_rt.co({self.ty(ty, scope)}, {e})"

    def emit_stmt(self, s, scope):
        k = s[0]
        if k == "decl":
            _, kw, name, ty, val, line = s
            if val[0] == "container":
                scope.declare(name)
                self.emit_container(py_name(name), val, scope)
                return
            e = self.typed(ty, val, scope)
            if kw == "var" and val != ("lit", "undefined"):
                e = f"_rt.cp({e})"  # Zig value semantics: a `var` initialised from a struct is a copy
            scope.declare(name, ty)
            self.w(f"{py_name(name)} = {e}")
        elif k == "assign":
            _, op, lhs, rhs, line = s
            if lhs == ("name", "_"):
                if rhs[0] not in ("name", "field"):
                    self.w(self.ex(rhs, scope))
                return
            self.emit_assign(op, lhs, rhs, scope)
        elif k == "exprstmt":
            self.w(self.ex(s[1], scope))
        elif k == "return":
            ret, generic, fname = self.ret_stack[-1]
            if s[1] is None:
                self.w("return")
            elif s[1][0] == "container":
                self.anon_count += 1
                cls = f"_{fname}_T{self.anon_count}"
                self.emit_container(cls, s[1], scope)
                self.w(f"return {cls}")
            else:
                self.w(f"return _rt.co({self.ty(ret, scope)}, {self.ex(s[1], scope)})")
        elif k == "if":
            _, cond, then, other = s
            self.w(f"if {self.ex(cond, scope)}:")
            self.emit_body(then, scope)
            while other is not None and other[0] == "if":
                _, cond, then, other = other
                self.w(f"elif {self.ex(cond, scope)}:")
                self.emit_body(then, scope)
            if other is not None:
                self.w("else:")
                self.emit_body(other, scope)
        elif k == "while":
            _, cond, cont, body = s
            self.w(f"while {self.ex(cond, scope)}:")
            self.cont_stack.append(cont)
            self.emit_body(body, scope)
            self.cont_stack.pop()
            if cont is not None:
                self.ind += 1
                self.emit_stmt(cont, scope)
                self.ind -= 1
        elif k == "for":
            _, it, cap, body = s
            scope.declare(cap)
            self.w(f"for {py_name(cap)} in {self.ex(it, scope)}:")
            self.cont_stack.append(None)
            self.emit_body(body, scope)
            self.cont_stack.pop()
        elif k == "continue":
            if self.cont_stack and self.cont_stack[-1] is not None:
                self.emit_stmt(self.cont_stack[-1], scope)
            self.w("continue")
        elif k == "break":
            self.w("break")
        elif k == "defer":
            pass  # the reference's defers are deinit / leak asserts: nothing the arithmetic depends on
        elif k == "block":
            self.emit_block(s, scope)
        else:
            raise SyntaxError(f"zig2py: unknown statement {k}")

    def emit_assign(self, op, lhs, rhs, scope):
        r = self.ex(rhs, scope)
        if lhs[0] == "deref":  # p.* = v
            assert op == "="
            self.w(f"_rt.store({self.ex(lhs[1], scope)}, {r})")
            return
        if lhs[0] == "name":
            n = lhs[1]
            where = scope.lookup(n)
            target = py_name(n) if where is None or where[0] == "local" else f"{where[1]}.{py_name(n)}"
            if op == "=":
                dty = scope.declared_type(n)
                if dty is not None:
                    r = f"_rt.co({self.ty(dty, scope)}, {r})"
                self.w(f"{target} = {r}")
            else:
                self.w(f"{target} = {self.binop(op[0], target, r)}")
            return
        if lhs[0] == "field":
            obj = self.ex(lhs[1], scope)
            f = py_name(lhs[2])
            if op == "=":
                if rhs[0] in ("anonlit", "num"):
                    self.w(f"_rt.setf({obj}, {f!r}, {r})")  # result-location typing from the declared field type
                else:
                    self.w(f"{obj}.{f} = {r}")
            else:
                self.w(f"_o = {obj}; _o.{f} = {self.binop(op[0], '_o.' + f, r)}")
            return
        if lhs[0] == "index":
            arr, idx = self.ex(lhs[1], scope), self.ex(lhs[2], scope)
            if op == "=":
                self.w(f"{arr}[{idx}] = {r}")
            else:
                self.w(f"_a = {arr}; _i = {idx}; _a[_i] = {self.binop(op[0], '_a[_i]', r)}")
            return
        raise SyntaxError(f"zig2py: cannot assign to {lhs[0]}")

    # ---- expressions -----------------------------------------------------------------------------
    def binop(self, op, a, b):
        if op == "/":
            return f"_rt.div({a}, {b})"
        return f"({a} {BIN_PY[op]} {b})"

    def ex(self, e, scope):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "str":
            return e[1]
        if k == "lit":
            return {"true": "True", "false": "False", "undefined": "None"}[e[1]]
        if k == "name":
            n = e[1]
            where = scope.lookup(n)
            if where is not None:
                return py_name(n) if where[0] == "local" else f"{where[1]}.{py_name(n)}"
            if n in PRIMS:
                return f"_rt.{n}"
            return py_name(n)
        if k == "paren":
            return f"({self.ex(e[1], scope)})"
        if k == "field":
            return f"{self.ex(e[1], scope)}.{py_name(e[2])}"
        if k == "deref":
            return f"_rt.load({self.ex(e[1], scope)})"
        if k == "index":
            return f"{self.ex(e[1], scope)}[{self.ex(e[2], scope)}]"
        if k == "call":
            args = ", ".join(self.ex(a, scope) for a in e[2])
            return f"{self.ex(e[1], scope)}({args})"
        if k == "builtin":
            return self.builtin(e[1], e[2], scope)
        if k == "bin":
            return self.binop(e[1], self.ex(e[2], scope), self.ex(e[3], scope))
        if k == "un":
            op, x = e[1], e[2]
            if op == "&":
                if x[0] == "field":
                    return f"_rt.FieldPtr({self.ex(x[1], scope)}, {py_name(x[2])!r})"
                return self.ex(x, scope)  # &local_struct: object identity
            v = self.ex(x, scope)
            return {"-": f"(-{v})", "!": f"(not {v})", "~": f"(~{v})"}[op]
        if k == "ifexpr":
            return f"({self.ex(e[2], scope)} if {self.ex(e[1], scope)} else {self.ex(e[3], scope)})"
        if k == "anonlit":
            inner = ", ".join(f"{py_name(n)}={self.ex(v, scope)}" for n, v in e[1])
            return f"_rt.Anon({inner})"
        if k == "structlit":
            inner = ", ".join(f"{py_name(n)}={self.ex(v, scope)}" for n, v in e[2])
            return f"_rt.co({self.ex(e[1], scope)}, _rt.Anon({inner}))"
        if k == "tuple":
            return "(" + "".join(self.ex(x, scope) + ", " for x in e[1]) + ")"
        if k == "enumlit":
            return f"_rt.EnumLit({e[1]!r})"
        if k == "switch":
            parts = []
            for tag, cap, body in e[2]:
                inner = Scope(scope)
                if cap:
                    inner.declare(cap)
                parts.append(f"{tag!r}: lambda {py_name(cap) if cap else '_p=None'}: {self.ex(body, inner)}")
            return f"_rt.switch({self.ex(e[1], scope)}, {{{', '.join(parts)}}})"
        if k == "container":
            raise SyntaxError("zig2py: container expression outside a declaration/return")
        raise SyntaxError(f"zig2py: unknown expression {k}")

    def builtin(self, name, args, scope):
        a = [self.ex(x, scope) for x in args]
        if name == "import":
            return f"_imp({a[0]})"
        if name == "This":
            where = scope
            while where is not None and where.container is None:
                where = where.parent
            return where.container[0]
        if name == "as":
            return f"_rt.as_({self.ty(args[0], scope)}, {a[1]})"
        simple = {"sqrt": "sqrt", "abs": "abs_", "min": "min_", "max": "max_", "floor": "floor", "sin": "sin", "cos": "cos",
                  "tan": "tan", "intFromFloat": "int_from_float", "floatFromInt": "float_from_int", "intCast": "int_cast",
                  "divTrunc": "div_trunc"}
        if name in simple:
            return f"_rt.{simple[name]}({', '.join(a)})"
        raise SyntaxError(f"zig2py: builtin @{name} is not in the subset")


def transpile(ast):
    return Emitter().emit_file(ast)
