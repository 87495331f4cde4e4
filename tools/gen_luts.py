#!/usr/bin/env python3
"""Generate the marching-cubes lookup data used by the product kernels and by the oracle.

The 256-case triangle tables are *interoperability data*: bit-exact face connectivity with the
reference is only possible with the same case -> triangle-list mapping, so the numbers are read
from the reference headers (include/mc/nagae.cuh:19-234, include/mc/lorensen.cuh:20-192) when
/root/reference is present, validated independently, and re-emitted in our own encodings:

  isoext_b200/csrc/mc_luts.inc   one uint64 per case: nibble k (k<15) = k-th edge id (0xF = none),
                                 top nibble = number of triangles          (product, device code)
  oracle/oracle_luts.h           plain int8 [256][16] rows, -1 terminated   (CPU oracle)

Validation done here (none of it uses reference code):
  * every triangle of case c uses only edges whose two corners differ in sign under c
    (the cell/edge convention of include/shared_luts.cuh:3-38), and every such edge is used
    -> the reference's edge_status_table (src/shared_luts.cu:6-71) is redundant: it equals the
       sign-change mask, which the kernels compute with XORs;
  * each case's triangle multiset is the image of a base case of luts/mc_methods/*.json under one of
    the 24 proper cube rotations (Lorensen: also complement + winding flip), with the rotations
    generated here from the three axis quarter-turns.

Run:  python tools/gen_luts.py            (rewrites both files; needs /root/reference)
      python tools/gen_luts.py --check    (re-validates the committed files without the reference)
"""
from __future__ import annotations

import argparse
import itertools
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

# corner i = (x,y,z) = (i>>2&1, i>>1&1, i&1); edge e joins EDGES[e][0] < EDGES[e][1]
EDGES = [(0, 1), (1, 3), (2, 3), (0, 2), (4, 5), (5, 7), (6, 7), (4, 6), (0, 4), (1, 5), (3, 7), (2, 6)]


def parse_header(path: Path) -> list[list[int]]:
    txt = path.read_text()
    m = re.search(r"tri_table\[\]\s*=\s*\{(.*?)\};", txt, re.S)
    nums = [int(t) for t in re.findall(r"-?\d+", m.group(1))]
    assert len(nums) % 256 == 0
    w = len(nums) // 256
    return [nums[c * w:(c + 1) * w] for c in range(256)]


def sign_change_mask(case: int) -> int:
    mask = 0
    for e, (a, b) in enumerate(EDGES):
        if ((case >> a) ^ (case >> b)) & 1:
            mask |= 1 << e
    return mask


def tris_of(row: list[int]) -> list[tuple[int, int, int]]:
    out = []
    for k in range(0, len(row) - 2, 3):
        if row[k] < 0:
            break
        out.append((row[k], row[k + 1], row[k + 2]))
    return out


def cube_rotations() -> list[list[int]]:
    """The 24 proper rotations as corner permutations, generated from quarter turns."""
    def corner(i):
        return (i >> 2 & 1, i >> 1 & 1, i & 1)

    def idx(p):
        return p[0] << 2 | p[1] << 1 | p[2]

    def rot_x(p):  # (x,y,z) -> (x, 1-z, y)
        return (p[0], 1 - p[2], p[1])

    def rot_y(p):
        return (p[2], p[1], 1 - p[0])

    def rot_z(p):
        return (1 - p[1], p[0], p[2])

    gens = [[idx(f(corner(i))) for i in range(8)] for f in (rot_x, rot_y, rot_z)]
    seen = {tuple(range(8))}
    frontier = [tuple(range(8))]
    while frontier:
        nxt = []
        for perm in frontier:
            for g in gens:
                q = tuple(g[perm[i]] for i in range(8))
                if q not in seen:
                    seen.add(q)
                    nxt.append(q)
        frontier = nxt
    assert len(seen) == 24
    return [list(p) for p in sorted(seen)]


def validate(name: str, table: list[list[int]], base_cases: dict | None, use_reflection: bool) -> None:
    for c in range(256):
        tris = tris_of(table[c])
        used = 0
        for t in tris:
            for e in t:
                assert 0 <= e < 12
                used |= 1 << e
        assert used == sign_change_mask(c), f"{name}: case {c} edge set mismatch"
        if c in (0, 255):
            assert not tris
    if base_cases is None:
        return
    rots = cube_rotations()
    edge_index = {e: i for i, e in enumerate(EDGES)}
    images: dict[int, set] = {}
    for case_str, tris in base_cases.items():
        bits = [int(s) for s in case_str]  # string position p <-> corner 7-p (format(i,'08b'))
        case = sum(bits[7 - i] << i for i in range(8))
        tris = [tuple(t) for t in tris if t]
        for perm in rots:
            new_case = sum(((case >> i) & 1) << perm[i] for i in range(8))
            emap = [edge_index[tuple(sorted((perm[a], perm[b])))] for a, b in EDGES]
            rt = [tuple(emap[e] for e in t) for t in tris]
            images.setdefault(new_case, set()).add(canon(rt))
            if use_reflection:
                images.setdefault(new_case ^ 0xFF, set()).add(canon([t[::-1] for t in rt]))
    for c in range(256):
        assert c in images, f"{name}: case {c} unreachable from base cases"
        assert canon(tris_of(table[c])) in images[c], f"{name}: case {c} is not an image of a base case"


def canon(tris):
    """Triangle multiset up to cyclic rotation of each triangle (orientation preserved)."""
    def cyc(t):
        return min((t[0], t[1], t[2]), (t[1], t[2], t[0]), (t[2], t[0], t[1]))
    return tuple(sorted(cyc(t) for t in tris))


def pack(row: list[int]) -> int:
    word = 0
    n = len(tris_of(row))
    for k in range(15):
        e = row[k] if k < len(row) and k < 3 * n else 0xF
        word |= (e & 0xF) << (4 * k)
    return word | (n << 60)


def unpack(word: int) -> list[int]:
    n = word >> 60
    return [(word >> (4 * k)) & 0xF for k in range(3 * n)]


def emit(tables: dict[str, list[list[int]]]) -> None:
    inc = ["// GENERATED by tools/gen_luts.py -- do not edit.",
           "// One uint64 per case: nibble k (k < 15) = k-th edge id of the triangle list (0xF = none);",
           "// bits 60..63 = number of triangles.  The includer defines ISX_LUT_QUAL (e.g. __device__).",
           "#ifndef ISX_LUT_QUAL", "#define ISX_LUT_QUAL static", "#endif"]
    for name in ("nagae", "lorensen"):
        words = [pack(r) for r in tables[name]]
        inc.append(f"ISX_LUT_QUAL const unsigned long long kTriWords_{name}[256] = {{")
        for i in range(0, 256, 4):
            inc.append("    " + ", ".join(f"0x{w:016x}ull" for w in words[i:i + 4]) + ",")
        inc.append("};")
    (ROOT / "isoext_b200/csrc/mc_luts.inc").write_text("\n".join(inc) + "\n")

    h = ["/* GENERATED by tools/gen_luts.py -- do not edit.  Oracle-side copy (test infrastructure). */",
         "#pragma once",
         "/* row c: edge ids of the triangles of case c, 3 per triangle, -1 terminated */"]
    for name in ("nagae", "lorensen"):
        h.append(f"static const signed char ORC_TRI_{name.upper()}[256][16] = {{")
        for r in tables[name]:
            n = len(tris_of(r))
            row = [r[k] if k < 3 * n else -1 for k in range(16)]
            h.append("    {" + ",".join(str(v) for v in row) + "},")
        h.append("};")
    (ROOT / "oracle/oracle_luts.h").write_text("\n".join(h) + "\n")


def read_committed() -> dict[str, list[list[int]]]:
    txt = (ROOT / "isoext_b200/csrc/mc_luts.inc").read_text()
    out = {}
    for name in ("nagae", "lorensen"):
        m = re.search(rf"kTriWords_{name}\[256\]\s*=\s*\{{(.*?)\}};", txt, re.S)
        words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", m.group(1))]
        assert len(words) == 256
        out[name] = [unpack(w) + [-1] for w in words]
    return out


BASE_CASE_FILES = {"nagae": "luts/mc_methods/nagae.json", "lorensen": "luts/mc_methods/lorensen.json"}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    if args.check or not REF.exists():
        tables = read_committed()
        for name, t in tables.items():
            validate(name, t, None, False)
        print("committed LUTs valid (edge-set check)")
        return 0
    tables = {"nagae": parse_header(REF / "include/mc/nagae.cuh"),
              "lorensen": parse_header(REF / "include/mc/lorensen.cuh")}
    for name, t in tables.items():
        meta = json.loads((REF / BASE_CASE_FILES[name]).read_text())
        validate(name, t, meta["base_cases"], meta.get("use_reflection", True))
        print(f"{name}: {sum(len(tris_of(r)) for r in t)} triangles over 256 cases, "
              f"max {max(len(tris_of(r)) for r in t)} per case -- valid")
    emit(tables)
    rt = read_committed()
    for name in tables:
        for c in range(256):
            assert [e for tri in tris_of(tables[name][c]) for e in tri] == [e for e in rt[name][c] if e >= 0]
    print("wrote isoext_b200/csrc/mc_luts.inc and oracle/oracle_luts.h")
    return 0


if __name__ == "__main__":
    sys.exit(main())
