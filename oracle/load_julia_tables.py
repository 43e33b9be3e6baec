#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE -- not part of the shipped product path.

Loader for a reference-held dump of the marching-simplex tables.  The connectivity inside a table case is "parity
unpinned" (DESIGN.md section 5): the reference builds its tables at run time with MiniQhull and no test stores them.  The
day a Julia run is available, a maintainer produces the dump with

    using GridapEmbedded.LevelSetCutters: LookupTable
    using Gridap.ReferenceFEs, JSON3
    open("lookup_tables_julia.json", "w") do io
      JSON3.write(io, Dict(string(n) => LookupTable(p) for (n, p) in
                           (:SEGMENT => SEGMENT, :TRI => TRI, :TET => TET, :QUAD => QUAD, :HEX => HEX)))
    end

(`struct LookupTable`, `/root/reference/src/LevelSetCutters/LookupTables.jl:2-13`: JSON3 writes the fields by name, a
`VectorValue{D,T}` as `{"data": [...]}`), and `python oracle/load_julia_tables.py lookup_tables_julia.json` turns it into the
JSON document both the oracle and the CUDA library consume (tests/golden/lookup_tables.json -> data/lookup_tables.json ->
gecut_tables.h via tools/pack_tables.py) -- no code changes.  The polytope numbering (vertices, edges, facets, simplexify)
is not part of `LookupTable`; it stays what oracle/gen_lookup_tables.py states (assumptions A1-A3).
"""
import json
import sys

import gen_lookup_tables as g

FULL = ("SEGMENT", "TRI", "TET")
FIELDS = ("case_to_subcell_to_points", "case_to_subcell_to_inout", "case_to_num_subcells_in", "case_to_num_subcells_out",
          "case_to_subfacet_to_points", "case_to_subfacet_to_orientation", "case_to_point_to_coordinates")


def _vv(x):
    """VectorValue{D,T}: JSON3 writes the struct field `data`; StaticArrays-style dumps are plain lists."""
    if isinstance(x, dict):
        x = x["data"]
    return [float(v) for v in x]


def convert_table(name, jt):
    """One `LookupTable` dump (dict of its fields) -> the table entry of our JSON document."""
    p = g.POLYTOPES[name]
    nv = len(p["vertices"])
    ncases = int(jt["num_cases"])
    if ncases != 2 ** nv:
        raise ValueError(f"{name}: num_cases {ncases} != 2^{nv}")
    t = dict(polytope=name, dim=p["dim"], num_cases=ncases, vertices=p["vertices"], edges=p["edges"],
             case_to_inoutcut=[int(v) for v in jt["case_to_inoutcut"]])
    if len(t["case_to_inoutcut"]) != ncases:
        raise ValueError(f"{name}: case_to_inoutcut has {len(t['case_to_inoutcut'])} entries")
    if name not in FULL:
        return t
    t["simplex_facets"] = p["simplex_facets"]
    t["case_to_subcell_to_points"] = [[[int(q) for q in sc] for sc in case] for case in jt["case_to_subcell_to_points"]]
    t["case_to_subcell_to_inout"] = [[int(v) for v in case] for case in jt["case_to_subcell_to_inout"]]
    t["case_to_num_subcells_in"] = [int(v) for v in jt["case_to_num_subcells_in"]]
    t["case_to_num_subcells_out"] = [int(v) for v in jt["case_to_num_subcells_out"]]
    t["case_to_subfacet_to_points"] = [[[int(q) for q in sf] for sf in case] for case in jt["case_to_subfacet_to_points"]]
    # Julia stores the orientation as T (+-1.0)
    t["case_to_subfacet_to_orientation"] = [[int(round(float(v))) for v in case] for case in jt["case_to_subfacet_to_orientation"]]
    t["case_to_point_to_coordinates"] = [[_vv(x) for x in case] for case in jt["case_to_point_to_coordinates"]]
    for k in FIELDS:
        if len(t[k]) != ncases:
            raise ValueError(f"{name}: {k} has {len(t[k])} cases")
    for c in range(ncases):
        ins = sum(1 for s in t["case_to_subcell_to_inout"][c] if s == g.IN)
        outs = sum(1 for s in t["case_to_subcell_to_inout"][c] if s == g.OUT)
        if ins != t["case_to_num_subcells_in"][c] or outs != t["case_to_num_subcells_out"][c]:
            raise ValueError(f"{name} case {c + 1}: IN/OUT counts disagree with case_to_subcell_to_inout")
        if len(t["case_to_subfacet_to_points"][c]) != len(t["case_to_subfacet_to_orientation"][c]):
            raise ValueError(f"{name} case {c + 1}: subfacet arrays of different lengths")
    g.check_invariants(t)  # positive sub-Jacobians, partition of the simplex, orientation IN -> OUT, mirror symmetry
    return t


def convert(dump):
    """{polytope name: LookupTable dump} -> JSON document in the layout of tests/golden/lookup_tables.json."""
    missing = [n for n in ("SEGMENT", "TRI", "TET", "QUAD", "HEX") if n not in dump]
    if missing:
        raise ValueError(f"dump lacks the tables of {missing}")
    return dict(generator="oracle/load_julia_tables.py",
                provenance="LookupTable(p) of the reference, dumped from Julia (fields of LookupTables.jl:2-13)",
                tables={n: convert_table(n, dump[n]) for n in ("SEGMENT", "TRI", "TET", "QUAD", "HEX")},
                simplexify={n: g.build_simplexify(n) for n in ("SEGMENT", "TRI", "TET", "QUAD", "HEX")})


if __name__ == "__main__":
    src = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else "lookup_tables.json"
    doc = convert(json.load(open(src)))
    with open(out, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print(f"wrote {out}: replace tests/golden/lookup_tables.json and gridapembedded.jl_b200/data/lookup_tables.json with it, "
          f"re-run gridapembedded.jl_b200/tools/pack_tables.py and oracle/emit_tables_inc.py, rebuild")
