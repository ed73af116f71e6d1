"""zig2py — a Zig-subset to Python transpiler used to RUN the reference's own source text in the tests.

TEST INFRASTRUCTURE ONLY (tests/ref_transpile.py is the entry point).  Nothing of the reference is stored in this
repository: the transpiler reads /root/reference/src/**.zig at test time.
"""
from .parser import parse  # noqa: F401
from .emit import transpile  # noqa: F401
from . import runtime  # noqa: F401
