"""zig2py.runtime — what transpiled reference code runs on.  TEST INFRASTRUCTURE.

Two kinds of things live here:

1. Language semantics (structs with value copies, tagged unions, pointers, result-location coercion of anonymous
   literals, IEEE-754 behaviour of `/`, @sqrt, @min ...).  Python floats are IEEE binary64, `math.sqrt` is correctly
   rounded and +,-,*,/ are single IEEE operations, so f64 arithmetic transpiled statement by statement gives the bits a
   Zig build without FMA contraction gives (Zig never contracts a*b+c unless asked to with @mulAdd).

2. Shims of what the reference imports from OUTSIDE its own tree — Zig std (`std.Random.DefaultPrng`, `std.math`,
   `std.ArrayList`, allocators, `std.debug`) and zigimg (`Image`).  These are restated from the published algorithms
    and are therefore NOT pinned by the reference's source text: `Random.float(f64)` and
   `uintLessThan` are the recalled Zig 0.14 algorithms, std.math functions are libm's (glibc here, musl-derived ports
   in Zig: differences of an ulp are possible and irrelevant to every tolerance in the tests).
"""
import math
import os
import sys

sys.setrecursionlimit(20000)

INF = float("inf")
NAN = float("nan")


# ---- primitive types ---------------------------------------------------------------------------------
class Prim:
    def __init__(self, name, is_float=False, is_int=False):
        self.name, self.is_float, self.is_int = name, is_float, is_int

    def __repr__(self):
        return self.name


f64 = Prim("f64", is_float=True)
f32 = Prim("f32", is_float=True)
comptime_float = Prim("comptime_float", is_float=True)
u8, u16, u32, u64, usize = (Prim(n, is_int=True) for n in ("u8", "u16", "u32", "u64", "usize"))
i32, i64, isize, comptime_int = (Prim(n, is_int=True) for n in ("i32", "i64", "isize", "comptime_int"))
bool = Prim("bool")  # noqa: A001  (the Zig type name)
void = Prim("void")
type = Prim("type")  # noqa: A001
anytype = Prim("anytype")

import builtins as _b  # the Python builtins shadowed just above


class PtrT:
    def __init__(self, t):
        self.t = t


class SliceT:
    def __init__(self, t):
        self.t = t


class ArrT:
    def __init__(self, n, t):
        self.n, self.t = n, t


# ---- values ------------------------------------------------------------------------------------------
class Anon:
    """`.{ .a = x, ... }` before its result type is known."""
    __slots__ = ("f",)

    def __init__(self, **kw):
        self.f = kw

    def __getattr__(self, n):  # tuple-ish structs that never get a type (zigimg pixel values)
        try:
            return self.f[n]
        except KeyError:
            raise AttributeError(n)


class EnumLit:
    def __init__(self, name):
        self.name = name

    def __eq__(self, o):
        return isinstance(o, EnumLit) and o.name == self.name

    def __hash__(self):
        return hash(self.name)


class Struct:
    _names = ()
    _types_thunk = staticmethod(lambda: ())
    _types_cache = None

    @classmethod
    def _types(cls):
        if cls.__dict__.get("_types_cache") is None:
            cls._types_cache = dict(zip(cls._names, cls._types_thunk()))
        return cls._types_cache

    @classmethod
    def _from_anon(cls, a):
        tys = cls._types()
        extra = set(a.f) - set(cls._names)
        if extra:
            raise TypeError(f"{cls.__name__} has no field(s) {sorted(extra)}")
        o = cls.__new__(cls)
        d = o.__dict__
        for n in cls._names:
            d[n] = co(tys[n], a.f[n]) if n in a.f else None
        return o

    @classmethod
    def _undefined(cls):
        o = cls.__new__(cls)
        for n in cls._names:
            o.__dict__[n] = None
        return o

    def _copy(self):
        o = self.__class__.__new__(self.__class__)
        o.__dict__.update(self.__dict__)
        return o

    def __repr__(self):
        return f"{self.__class__.__name__}({', '.join(f'{n}={self.__dict__.get(n)!r}' for n in self._names)})"


class Union(Struct):
    """union(enum): _names = tags, _types = payload types; instance = (tag, payload)."""

    @classmethod
    def _from_anon(cls, a):
        if len(a.f) != 1:
            raise TypeError(f"union {cls.__name__} initialised with {len(a.f)} fields")
        (tag, val), = a.f.items()
        tys = cls._types()
        if tag not in tys:
            raise TypeError(f"union {cls.__name__} has no tag {tag}")
        o = cls.__new__(cls)
        o.tag = tag
        o.payload = co(tys[tag], val)
        return o

    @classmethod
    def _undefined(cls):
        o = cls.__new__(cls)
        o.tag = None
        o.payload = None
        return o

    def __repr__(self):
        return f"{self.__class__.__name__}.{self.tag}({self.payload!r})"


class Ptr:
    """Result of allocator.create(T): a pointer to one heap T.  Attribute access auto-dereferences (Zig's `p.field`).

    get/clone/deinit make a bare `*const Material` usable where the reference stores it into an `Rc(Material)` field —
    scenes 1-5 of main.zig do that and therefore do not type-check; this is the minimal repair
    ("wrap the material in an Rc") applied at run time instead of to the text.
    """

    def __init__(self, t, pointee):
        object.__setattr__(self, "_t", t)
        object.__setattr__(self, "_pointee", pointee)

    def deref(self):
        return self._pointee

    def __getattr__(self, n):
        return getattr(self._pointee, n)

    def __setattr__(self, n, v):
        setattr(self._pointee, n, v)

    def get(self):
        return self

    def clone(self):
        return self

    def deinit(self):
        pass


class FieldPtr:
    """`&obj.field`"""

    def __init__(self, obj, name):
        object.__setattr__(self, "_obj", obj)
        object.__setattr__(self, "_name", name)

    def deref(self):
        return getattr(self._obj, self._name)

    def __getattr__(self, n):
        return getattr(self.deref(), n)

    def __setattr__(self, n, v):
        setattr(self.deref(), n, v)


def _is_ptr(x):
    return x.__class__ is FieldPtr or x.__class__ is Ptr


def co(T, x):
    """coerce x to the declared type T (Zig result-location typing + implicit int->float of comptime literals)"""
    tx = x.__class__
    if tx is float:
        return x
    if tx is Anon:
        if isinstance(T, _b.type) and issubclass(T, Struct):
            return T._from_anon(x)
        if hasattr(T, "_from_anon"):
            return T._from_anon(x)
        return x
    if tx is int:
        if T.__class__ is Prim and T.is_float:
            return float(x)
        return x
    if tx is FieldPtr or tx is Ptr:
        if T.__class__ is PtrT or T is anytype or T is None:
            return x
        if tx is Ptr and getattr(T, "_generic_name", None) == "Rc":
            return x  # D4 repair, see Ptr
        return co(T, x.deref())  # a value is expected: implicit dereference (method call on a pointer)
    return x


def cp(x):
    if isinstance(x, Struct):
        return x._copy()
    return x


def undefined(T):
    if isinstance(T, ArrT):
        return [undefined(T.t) for _ in range(T.n)]
    if isinstance(T, _b.type) and issubclass(T, Struct):
        return T._undefined()
    return None


def load(p):
    if _is_ptr(p):
        return cp(p.deref())
    return cp(p)


def store(p, v):
    if p.__class__ is Ptr:
        object.__setattr__(p, "_pointee", cp(co(p._t, v)))
    elif p.__class__ is FieldPtr:
        setf(p._obj, p._name, v)
    elif isinstance(p, Struct):  # pointer to a local struct = the object itself
        v = co(p.__class__, v)
        p.__dict__.clear()
        p.__dict__.update(v.__dict__)
    else:
        raise TypeError(f"cannot store through {p!r}")


def setf(obj, name, v):
    target = obj.deref() if _is_ptr(obj) else obj
    if isinstance(target, Struct):
        t = target.__class__._types().get(name)
        if t is not None:
            v = co(t, v)
    setattr(target, name, v)


def switch(u, table):
    if _is_ptr(u):
        u = u.deref()
    fn = table.get(u.tag)
    if fn is None:
        fn = table[None]
    return fn(u.payload)


def generic(fn):
    cache = {}

    def wrapper(*args):
        key = tuple(id(a) for a in args)
        if key not in cache:
            cls = fn(*args)
            cls._generic_name = fn.__name__
            cls._generic_args = args
            cls.__name__ = f"{fn.__name__}({', '.join(getattr(a, '__name__', repr(a)) for a in args)})"
            cache[key] = cls
        return cache[key]

    wrapper.__name__ = fn.__name__
    return wrapper


# ---- arithmetic with IEEE semantics ---------------------------------------------------------------------
def div(a, b):
    if a.__class__ is int and b.__class__ is int:
        raise TypeError("integer '/' is not in the subset")
    try:
        return a / b
    except ZeroDivisionError:
        a = float(a)
        if a != a or a == 0.0:
            return NAN
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, float(b)) < 0)
        return -INF if neg else INF


def sqrt(x):
    if x < 0:
        return NAN  # IEEE sqrt of a negative number; -0.0 falls through to math.sqrt -> -0.0
    return math.sqrt(x)


def abs_(x):
    return _b.abs(x)


def min_(a, b):  # @min on floats = minNum: a NaN operand loses
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def max_(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def floor(x):
    if x != x or x in (INF, -INF):
        return x
    return float(math.floor(x))


def _libm(fn):
    def f(x):
        try:
            return fn(x)
        except (ValueError, OverflowError):
            return NAN
    return f


sin = _libm(math.sin)
cos = _libm(math.cos)
tan = _libm(math.tan)


def int_from_float(x):
    return int(x)  # truncation toward zero, as @intFromFloat


def float_from_int(x):
    return float(x)


def int_cast(x):
    return x


def div_trunc(a, b):
    q = div(a, b)
    return float(math.trunc(q)) if isinstance(q, float) and q == q and q not in (INF, -INF) else q


def as_(T, x):
    if T.__class__ is Prim:
        if T.is_float:
            return float(x)
        if T.is_int:
            return int(x)
        return x
    return co(T, x)


# ---- shim: Zig std ------------------------------------------------------------------------------------------
M64 = (1 << 64) - 1


class Xoshiro256:
    """std.Random.Xoshiro256 (xoshiro256++), seeded through SplitMix64 — SURVEY.md Appendix E."""

    def __init__(self, seed):
        s = []
        sm = seed & M64
        for _ in range(4):
            sm = (sm + 0x9E3779B97F4A7C15) & M64
            z = sm
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            s.append(z ^ (z >> 31))
        self.s = s
        self.draws = 0

    @staticmethod
    def init(seed):
        return Xoshiro256(seed)

    def next(self):
        s = self.s
        self.draws += 1
        x = (s[0] + s[3]) & M64
        r = ((((x << 23) | (x >> 41)) & M64) + s[0]) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) | (s[3] >> 19)) & M64
        return r

    def random(self):
        return Random(self)


class Random:
    """std.Random: float(f64) (52 mantissa bits + geometric exponent) and intRangeLessThan (Lemire) of Zig 0.14."""

    DefaultPrng = Xoshiro256

    def __init__(self, gen):
        self.gen = gen

    def float(self, T):
        assert T is f64
        r = self.gen.next()
        lz = 64 - r.bit_length()
        if lz >= 12:
            lz = 12
            while True:
                a = 64 - self.gen.next().bit_length()
                lz += a
                if a != 64:
                    break
                if lz >= 1022:
                    lz = 1022
                    break
        bits = ((1022 - lz) << 52) | (r & 0xFFFFFFFFFFFFF)
        import struct
        return struct.unpack("<d", struct.pack("<Q", bits))[0]

    def uintLessThan(self, T, less_than):
        x = self.gen.next()
        m = x * less_than
        lo = m & M64
        if lo < less_than:
            t = (-less_than) & M64
            t %= less_than
            while lo < t:
                x = self.gen.next()
                m = x * less_than
                lo = m & M64
        return m >> 64

    def intRangeLessThan(self, T, at_least, less_than):
        return at_least + self.uintLessThan(T, less_than - at_least)


class ZList(list):
    @property
    def len(self):
        return len(self)


@generic
def ArrayList(T):
    class _ArrayList:
        def __init__(self):
            self.items = ZList()

        @staticmethod
        def init(allocator):
            return _ArrayList()

        @staticmethod
        def initCapacity(allocator, n):
            return _ArrayList()

        def append(self, x):
            self.items.append(cp(co(T, x)))

        def deinit(self):
            pass

    return _ArrayList


class Allocator:
    def create(self, T):
        return Ptr(T, undefined(T))

    def destroy(self, p):
        pass


class _Gpa:
    @staticmethod
    def _from_anon(a):
        return _Gpa()

    def allocator(self):
        return Allocator()

    def deinit(self):
        return EnumLit("ok")


class _Ns:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _pow(T, a, b):
    try:
        return math.pow(a, b)
    except (ValueError, OverflowError):
        return NAN


def _clamp(v, lo, hi):  # std.math.clamp: @max(lower, @min(val, upper))
    return max_(lo, min_(v, hi))


def _acos(x):
    try:
        return math.acos(x)
    except ValueError:
        return NAN


std = _Ns(
    Random=Random,
    ArrayList=ArrayList,
    mem=_Ns(Allocator=Allocator),
    heap=_Ns(GeneralPurposeAllocator=lambda cfg: _Gpa),
    debug=_Ns(print=lambda *a: None, assert_=lambda c: None),
    math=_Ns(pi=math.pi, atan2=math.atan2, acos=_acos, pow=_pow, inf=lambda T: INF, clamp=_clamp, sin=sin, cos=cos),
)
setattr(std.debug, "assert", std.debug.assert_)


# ---- shim: zigimg -----------------------------------------------------------------------------------------------
class _Pixels:
    def __init__(self, raw=None, rgb24=None):
        self._raw = raw
        self.rgb24 = rgb24

    def asBytes(self):
        return self._raw


class Image:
    asset_root = None   # set by the loader: the reference's repo root (main.zig opens "assets/sekaichizu.png")
    created = []        # images made by Image.create, in order (main.zig's output image)

    def __init__(self, width, height, pixels):
        self.width, self.height, self.pixels = width, height, pixels

    @staticmethod
    def fromFilePath(allocator, path):
        from PIL import Image as PILImage
        im = PILImage.open(os.path.join(Image.asset_root, path)).convert("RGBA")
        return Image(im.width, im.height, _Pixels(raw=im.tobytes()))

    @staticmethod
    def create(allocator, width, height, fmt):
        img = Image(width, height, _Pixels(rgb24=[None] * (width * height)))
        Image.created.append(img)
        return img

    def deinit(self):
        pass

    def writeToFilePath(self, path, opts):
        pass  # the harness reads Image.created[-1] instead of a file


zigimg = _Ns(Image=Image, PixelFormat=_Ns(rgb24="rgb24"))
