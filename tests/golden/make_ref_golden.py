"""Writes tests/golden/ref_golden.json: outputs of the REFERENCE ITSELF (its source text executed through
tests/ref_transpile.py) on the seeded inputs of tests/ref_cases.py.  Needs /root/reference; run from the repo root:

    python tests/golden/make_ref_golden.py

The fixture is what pins the oracle on machines without the reference (tests/test_ref_golden.py) — the inputs are
regenerated from the same seeds there, so only outputs are stored.  Python's json round-trips f64 exactly (repr).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import numpy as np  # noqa: E402

import ref_cases as rc  # noqa: E402

# (family, parameters) — small enough for a fixture of a few hundred KB, same generators as the live pin
PLAN = dict(
    main=rc.MAIN_CASES,
    hits=[("random:11", 250, 1), ("random:12", 250, 2), ("builtin:1", 200, 3), ("builtin:6", 200, 4), ("builtin:5", 100, 5),
          ("builtin:2", 60, 6), ("builtin:4", 60, 7)],
    boxes=["random:11", "builtin:1", "builtin:6"],
    aabb=(600, 3), helpers=(400, 4), textures=(120, 6), perlin=(500, 7), scatter=(700, 8), camera=(300, 9),
    ray_color=[("builtin:1", (0.7, 0.8, 1.0), 60), ("builtin:6", (0.0, 0.0, 0.0), 40), ("builtin:3", (0.7, 0.8, 1.0), 40)],
    samplers=(77, 300),
)


def evaluate(side):
    """side = 'ref' or 'orc': the same plan through the reference or through the oracle"""
    f = (lambda name: getattr(rc, f"{side}_{name}"))
    out = {}
    out["main"] = [f("main")(*c) for c in PLAN["main"]]
    out["hits"] = [f("hits")(k, rc.rays_for(k, n, seed)) for k, n, seed in PLAN["hits"]]
    out["boxes"] = [f("boxes")(k) for k in PLAN["boxes"]]
    out["aabb"] = f("aabb")(*rc.aabb_inputs(*PLAN["aabb"]))
    out["helpers"] = f("helpers")(*rc.helper_inputs(*PLAN["helpers"]))
    out["textures"] = f("textures")(rc.texture_inputs(*PLAN["textures"]))
    out["perlin"] = f("perlin")(rc.perlin_inputs(*PLAN["perlin"]))
    out["perlin_tables"] = list(rc.ref_perlin_tables(42) if side == "ref" else rc.orc_perlin_tables())
    out["scatter"] = f("scatter")(rc.scatter_inputs(*PLAN["scatter"]))
    out["camera"] = f("camera")(rc.camera_inputs(*PLAN["camera"]))
    out["ray_color"] = [f("ray_color")(k, rc.rays_for(k, n, 21), bg, 50, np.arange(1000, 1000 + n)) for k, bg, n in PLAN["ray_color"]]
    out["samplers"] = f("samplers")(*PLAN["samplers"])
    return out


if __name__ == "__main__":
    import ref_transpile
    assert ref_transpile.available(), "the reference sources are needed to make this fixture"
    data = evaluate("ref")
    path = os.path.join(HERE, "ref_golden.json")
    with open(path, "w") as fh:
        json.dump(dict(about="outputs of /root/reference/src executed via tests/ref_transpile.py; inputs = tests/ref_cases.py seeds",
                       data=data), fh, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
