"""The committed outputs of the reference (tests/golden/ref_golden.json) against the oracle — runs anywhere.

The fixture holds what the reference's own source computes (tests/golden/make_ref_golden.py executes it through
tests/ref_transpile.py); this test recomputes the same families with the CPU oracle and demands equality, bit for
bit.  Where /root/reference is present the fixture is also checked for freshness against the live reference.
"""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import ref_cases as rc  # noqa: E402
import make_ref_golden as mk  # noqa: E402


def _same(a, b):
    """== with NaN == NaN and the MovingSphere u,v convention (None in the reference, 0 in the oracle) for hit lists"""
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, float) and isinstance(b, float) and a != a and b != b:
        return True
    return a == b


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "ref_golden.json")) as fh:
        return json.load(fh)["data"]


@pytest.fixture(scope="module")
def oracle_side():
    return mk.evaluate("orc")


@pytest.mark.parametrize("family", ["main", "boxes", "aabb", "helpers", "textures", "perlin", "perlin_tables", "scatter", "camera",
                                    "ray_color", "samplers"])
def test_oracle_reproduces_reference_outputs(golden, oracle_side, family):
    assert _same(golden[family], oracle_side[family]), family


def test_oracle_reproduces_reference_hit_records(golden, oracle_side):
    for (key, n, seed), ref_out, orc_out in zip(mk.PLAN["hits"], golden["hits"], oracle_side["hits"]):
        assert rc.same_hits(ref_out, orc_out), key


def test_fixture_is_what_the_reference_computes_today(golden):
    import ref_transpile
    if not ref_transpile.available():
        pytest.skip("needs /root/reference")
    live = mk.evaluate("ref")
    assert _same(json.loads(json.dumps(live)), golden)
