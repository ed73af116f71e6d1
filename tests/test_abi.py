"""The C-ABI library loads and exports every symbol include/rtw_cuda.h declares; the ctypes mirror matches the
header's struct layout.  No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtw_cuda.h")


def test_header_symbols_are_exported(rtw):
    L = rtw.cuda_lib.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(rtw_cuda_[a-z0-9_]+)\s*\(", text))
    assert declared == set(rtw.abi.CUDA_SYMBOLS), declared ^ set(rtw.abi.CUDA_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.rtw_cuda_abi_version() == rtw.abi.RTW_ABI_VERSION


def test_ctypes_mirror_matches_header_layout(rtw):
    abi = rtw.abi
    structs = {"rtw_prim": abi.Prim, "rtw_xform": abi.Xform, "rtw_material": abi.Material, "rtw_texture": abi.Texture,
               "rtw_image": abi.Image, "rtw_perlin": abi.Perlin, "rtw_scene_desc": abi.SceneDesc,
               "rtw_camera": abi.Camera, "rtw_render_params": abi.RenderParams, "rtw_stats": abi.Stats}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "rtw_cuda.h"\nint main(void){' + "".join(lines) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "layout.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "layout")
        subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_enums_match_header(rtw):
    text = open(HEADER).read()
    abi = rtw.abi

    def val(name):
        return int(re.search(name + r"\s*=\s*(\d+)", text).group(1))
    assert [val(n) for n in ("RTW_PRIM_SPHERE", "RTW_PRIM_MOVING_SPHERE", "RTW_PRIM_XY_RECT", "RTW_PRIM_XZ_RECT", "RTW_PRIM_YZ_RECT")] == \
        [abi.PRIM_SPHERE, abi.PRIM_MOVING_SPHERE, abi.PRIM_XY_RECT, abi.PRIM_XZ_RECT, abi.PRIM_YZ_RECT]
    assert [val(n) for n in ("RTW_MAT_DIFFUSE", "RTW_MAT_METAL", "RTW_MAT_DIELECTRIC", "RTW_MAT_DIFFUSE_LIGHT")] == [0, 1, 2, 3]
    assert [val(n) for n in ("RTW_TEX_SOLID", "RTW_TEX_CHECKER", "RTW_TEX_NOISE", "RTW_TEX_IMAGE")] == [0, 1, 2, 3]
    assert [val(n) for n in ("RTW_VARIANT_AUTO", "RTW_VARIANT_MEGA_FLAT", "RTW_VARIANT_MEGA_BVH", "RTW_VARIANT_WAVEFRONT")] == [0, 1, 2, 3]
    assert [val(n) for n in ("RTW_BVH_BUILDER_SAH", "RTW_BVH_BUILDER_LBVH")] == [abi.BVH_BUILDER_SAH, abi.BVH_BUILDER_LBVH]
    assert int(re.search(r"RTW_FLAG_COUNT_EVENTS\s*=\s*(\d+)u", text).group(1)) == abi.FLAG_COUNT_EVENTS
    assert int(re.search(r"RTW_FLAG_DETERMINISTIC\s*=\s*(\d+)u", text).group(1)) == abi.FLAG_DETERMINISTIC
    assert int(re.search(r"#define RTW_ABI_VERSION (\d+)u", text).group(1)) == abi.RTW_ABI_VERSION


def test_zig_mirror_keeps_the_header_field_order(rtw):
    """zig/rtw_cuda.zig is uncompiled here (no toolchain): at least its extern structs must list the header's fields in
    the header's order, and its ABI version must match."""
    zig = open(os.path.join(ROOT, "raytracinginoneweekend.zig_b200", "zig", "rtw_cuda.zig")).read()
    abi = rtw.abi
    assert int(re.search(r"ABI_VERSION: u32 = (\d+);", zig).group(1)) == abi.RTW_ABI_VERSION
    for zname, cls in (("Camera", abi.Camera), ("RenderParams", abi.RenderParams), ("Stats", abi.Stats),
                       ("SceneDesc", abi.SceneDesc), ("Prim", abi.Prim), ("Xform", abi.Xform), ("Material", abi.Material),
                       ("Texture", abi.Texture), ("Image", abi.Image), ("Perlin", abi.Perlin)):
        m = re.search(r"pub const " + zname + r" = extern struct \{(.*?)\};", zig, re.S)
        assert m, zname
        body = re.sub(r"//[^\n]*", "", m.group(1))
        fields = re.findall(r"([a-z_][a-z0-9_]*)\s*:", body)
        assert fields == [f for f, _ in cls._fields_], (zname, fields)


def test_product_does_not_touch_the_oracle():
    """The product path must never import, link or execute oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "raytracinginoneweekend.zig_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp", ".zig")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_binding" not in text and "liboracle" not in text and "rtw_oracle" not in text, f


def test_create_fails_loudly_without_a_gpu(rtw):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtw.RtwCudaError, match="no CUDA device"):
        rtw.Context(0)
