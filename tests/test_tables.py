"""Lookup tables: the committed golden dump equals what the generator produces, the product's copy and packed header
are in sync with it, and the invariants of SURVEY.md Appendix B hold."""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "lookup_tables.json")
PKG = os.path.join(ROOT, "gridapembedded.jl_b200")


def test_generator_reproduces_golden(tmp_path):
    import gen_lookup_tables as g
    out = tmp_path / "t.json"
    doc = g.main(str(out))
    gold = json.load(open(GOLD))
    assert doc["tables"] == gold["tables"]
    assert doc["simplexify"] == gold["simplexify"]


def test_product_copy_equals_golden():
    a = json.load(open(GOLD))
    b = json.load(open(os.path.join(PKG, "data", "lookup_tables.json")))
    assert a["tables"] == b["tables"] and a["simplexify"] == b["simplexify"]


def test_packed_header_in_sync(tmp_path):
    out = tmp_path / "gecut_tables.h"
    subprocess.check_call([sys.executable, os.path.join(PKG, "tools", "pack_tables.py"),
                           os.path.join(PKG, "data", "lookup_tables.json"), str(out)], stdout=subprocess.DEVNULL)
    assert out.read_text() == open(os.path.join(PKG, "csrc", "gecut_tables.h")).read()


def test_table_shapes_and_counts():
    t = json.load(open(GOLD))["tables"]
    tet = t["TET"]
    assert tet["num_cases"] == 16
    nsub = [len(c) for c in tet["case_to_subcell_to_points"]]
    npts = [len(c) for c in tet["case_to_point_to_coordinates"]]
    nfac = [len(c) for c in tet["case_to_subfacet_to_points"]]
    # 1 vs 3 vertices -> 4 subcells / 7 points / 1 facet; 2 vs 2 -> 6 / 8 / 2
    for c in range(16):
        k = bin(c).count("1")
        exp = {0: (1, 4, 0), 4: (1, 4, 0), 1: (4, 7, 1), 3: (4, 7, 1), 2: (6, 8, 2)}[k]
        assert (nsub[c], npts[c], nfac[c]) == exp
    tri = t["TRI"]
    for c in range(8):
        k = bin(c).count("1")
        exp = {0: (1, 3, 0), 3: (1, 3, 0), 1: (3, 5, 1), 2: (3, 5, 1)}[k]
        assert (len(tri["case_to_subcell_to_points"][c]), len(tri["case_to_point_to_coordinates"][c]),
                len(tri["case_to_subfacet_to_points"][c])) == exp
    # n-cubes: only the IN/OUT/CUT classification is used (LookupTables.jl:289-302)
    for name, nv in (("QUAD", 4), ("HEX", 8)):
        ioc = t[name]["case_to_inoutcut"]
        assert ioc[0] == -1 and ioc[-1] == 1 and all(v == 0 for v in ioc[1:-1]) and len(ioc) == 2 ** nv


def test_invariants():
    import gen_lookup_tables as g
    for t in json.load(open(GOLD))["tables"].values():
        g.check_invariants(t)


def test_simplexify_positive_orientation():
    doc = json.load(open(GOLD))
    for name in ("QUAD", "HEX"):
        V = np.asarray(doc["tables"][name]["vertices"], float)
        raw, pos = doc["simplexify"][name]["raw"], doc["simplexify"][name]["positive"]
        assert len(raw) == (2 if name == "QUAD" else 6)
        vol = 0.0
        for r, p in zip(raw, pos):
            assert sorted(r) == sorted(p)
            X = V[[i - 1 for i in p]]
            d = np.linalg.det((X[1:] - X[0]).T)
            assert d > 0
            vol += d
        assert abs(vol - (2 if name == "QUAD" else 6)) < 1e-14  # sum of D! * simplex volumes = D! * 1
