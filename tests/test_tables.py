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


def _julia_style_dump(doc):
    """What `JSON3.write(LookupTable(p))` would hold for our tables (field names of LookupTables.jl:2-13; VectorValue as
    {"data": [...]}, orientations as Float64, plus the normals the struct carries and our JSON does not)."""
    import gen_lookup_tables as g
    dump = {}
    for name, t in doc["tables"].items():
        jt = {"num_cases": t["num_cases"], "case_to_inoutcut": t["case_to_inoutcut"]}
        if "case_to_subcell_to_points" in t:
            jt.update({k: t[k] for k in ("case_to_subcell_to_points", "case_to_subcell_to_inout", "case_to_num_subcells_in",
                                         "case_to_num_subcells_out", "case_to_subfacet_to_points")})
            jt["case_to_subfacet_to_orientation"] = [[float(o) for o in case] for case in t["case_to_subfacet_to_orientation"]]
            jt["case_to_point_to_coordinates"] = [[{"data": list(x)} for x in case] for case in t["case_to_point_to_coordinates"]]
            jt["case_to_subfacet_to_normal"] = [
                [{"data": [float(v) for v in g._unit_normal_from_points([coords[q - 1] for q in f])]} for f in facets]
                for coords, facets in zip(t["case_to_point_to_coordinates"], t["case_to_subfacet_to_points"])]
        else:  # n-cubes: the reference allocates empty per-case vectors
            for k in ("case_to_subcell_to_points", "case_to_subcell_to_inout", "case_to_subfacet_to_points",
                      "case_to_subfacet_to_normal", "case_to_subfacet_to_orientation", "case_to_point_to_coordinates"):
                jt[k] = [[] for _ in range(t["num_cases"])]
            jt["case_to_num_subcells_in"] = [0] * t["num_cases"]
            jt["case_to_num_subcells_out"] = [0] * t["num_cases"]
        dump[name] = jt
    return dump


def test_julia_dump_loader_round_trip():
    """A reference-held `LookupTable` dump drops in without code changes: a synthetic dump in the reference's field layout
    goes through oracle/load_julia_tables.py and yields exactly the document the oracle and the library consume."""
    import load_julia_tables as lj
    gold = json.load(open(GOLD))
    dump = json.loads(json.dumps(_julia_style_dump(gold)))  # through text, as a file would
    doc = lj.convert(dump)
    assert doc["tables"] == gold["tables"]
    assert doc["simplexify"] == gold["simplexify"]


def test_julia_dump_loader_takes_the_dump_not_the_generator():
    """The loader really uses the dump's connectivity (another valid table gives another document) and rejects dumps
    that violate the table invariants."""
    import copy
    import pytest
    import load_julia_tables as lj
    gold = json.load(open(GOLD))
    dump = _julia_style_dump(gold)
    alt = copy.deepcopy(dump)
    # another valid ordering of TRI case 2 and of its mirror case: swap the first two subcells (and their states)
    for c in (1, 6):
        sc, io = alt["TRI"]["case_to_subcell_to_points"][c], alt["TRI"]["case_to_subcell_to_inout"][c]
        sc[0], sc[1] = sc[1], sc[0]
        io[0], io[1] = io[1], io[0]
    doc = lj.convert(alt)
    assert doc["tables"]["TRI"]["case_to_subcell_to_points"][1] != gold["tables"]["TRI"]["case_to_subcell_to_points"][1]
    assert doc["tables"]["TET"] == gold["tables"]["TET"]
    bad = copy.deepcopy(dump)
    a = bad["TET"]["case_to_subcell_to_points"][1][0]
    a[0], a[1] = a[1], a[0]  # negative Jacobian
    with pytest.raises(AssertionError):
        lj.convert(bad)
    bad2 = copy.deepcopy(dump)
    del bad2["HEX"]
    with pytest.raises(ValueError):
        lj.convert(bad2)
