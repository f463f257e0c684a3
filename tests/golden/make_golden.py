#!/usr/bin/env python3
"""Generate tests/golden/reference_kats.json from the reference's own unit tests.

The reference (/root/reference, Rust) cannot be executed in this image (no rustc/cargo), so
its golden vectors are TRANSCRIBED from the literal matrices and expected arrays in its
``#[cfg(test)]`` module by parsing the source text.  Run in the build container only
(``python tests/golden/make_golden.py``); the GPU box never reads /root/reference.

Every entry records the reference file:line range it was taken from.
"""
import json
import os
import re
import sys

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def fn_body(src_lines, name):
    """(first_line_no, last_line_no, text) of `fn name()` up to its closing brace."""
    start = None
    for i, l in enumerate(src_lines):
        if re.search(r"\bfn\s+%s\s*\(\s*\)" % re.escape(name), l):
            start = i
            break
    if start is None:
        raise KeyError(name)
    depth = 0
    seen = False
    for j in range(start, len(src_lines)):
        depth += src_lines[j].count("{") - src_lines[j].count("}")
        if "{" in src_lines[j]:
            seen = True
        if seen and depth == 0:
            return start + 1, j + 1, "".join(src_lines[start:j + 1])
    raise ValueError(name)


def strip_comments(text):
    return re.sub(r"//[^\n]*", "", text)


def matrices(text, subst=None):
    """All `&[ &[...], &[...] ]` literals (lists of lists of ints) in order of appearance."""
    text = strip_comments(text)
    out = []
    for m in re.finditer(r"&\[\s*((?:&\[[^\]]*\]\s*,?\s*)+)\]", text):
        rows = []
        for r in re.finditer(r"&\[([^\]]*)\]", m.group(1)):
            cells = [c.strip() for c in r.group(1).split(",") if c.strip()]
            rows.append([int(subst.get(c, c)) if subst else int(c) for c in cells])
        out.append(rows)
    return out


def flat_arrays(text, subst=None):
    """`vec![..]` / `&[..]` flat int arrays on assert lines, keyed by the field they check."""
    text = strip_comments(text)
    res = {}
    for m in re.finditer(r"assert_eq!\(\s*(?:\w+)\.(v|col_index|row_index)(?:\.as_slice\(\))?\s*,\s*(?:vec!|&)\[([^\]]*)\]", text):
        cells = [c.strip() for c in m.group(2).split(",") if c.strip()]
        res[m.group(1)] = [int(subst.get(c, c)) if subst else int(c) for c in cells]
    return res


def main():
    sparse = open(os.path.join(REF, "sparse.rs")).readlines()
    dense = open(os.path.join(REF, "dense.rs")).readlines()
    kats = {"_source": "transcribed from /root/reference/src/{sparse,dense}.rs #[cfg(test)] by tests/golden/make_golden.py",
            "structure": [], "mul_dense": [], "mul_vector": [], "dense": []}

    # --- structure KATs: Csr::from_data -> exact v / col_index / row_index -------------
    for name, subst in (("example_mat_0", None), ("example_mat_1", None), ("example_mat_2", None),
                        ("csr_with_empty_row_top", {"a": "11", "b": "12", "c": "13"}),
                        ("csr_with_empty_row_middle", None)):
        a, b, text = fn_body(sparse, name)
        mats = matrices(text, subst)
        arrs = flat_arrays(text, subst)
        assert len(mats) == 1 and set(arrs) == {"v", "col_index", "row_index"}, (name, mats, arrs)
        kats["structure"].append({"name": name, "ref": f"src/sparse.rs:{a}-{b}", "data": mats[0], **arrs})

    # create_mat_by_insert: inserts 5,6,7 at row 0 of a 3x3, equals from_data([[5,6,7],[0..],[0..]])
    a, b, text = fn_body(sparse, "create_mat_by_insert")
    mats = matrices(text)
    kats["structure"].append({"name": "create_mat_by_insert", "ref": f"src/sparse.rs:{a}-{b}",
                              "inserts": [[5, 0, 0], [6, 0, 1], [7, 0, 2]], "dims": [3, 3],
                              "equals_from_data": mats[-1]})

    # --- hot path KATs ----------------------------------------------------------------------
    for name in ("test_dense_mul", "test_nnz"):
        a, b, text = fn_body(sparse, name)
        mats = matrices(text)
        t = strip_comments(text)
        # order of appearance differs between the two tests; identify by the constructor
        order = [m.group(1) for m in re.finditer(r"(Dense|Csr)::from_data\(", t)]
        assert len(order) == 3 and len(mats) == 3, (name, order)
        dense_cols = mats[order.index("Dense")]
        csr_idx = [i for i, o in enumerate(order) if o == "Csr"]
        entry = {"name": name, "ref": f"src/sparse.rs:{a}-{b}",
                 "dense_columns": dense_cols,          # Dense::from_data rows are COLUMNS (dense.rs:21-29)
                 "csr_rows": mats[csr_idx[0]], "output_rows": mats[csr_idx[1]]}
        m = re.search(r"get_nnz\(\)\s*,\s*(\d+)", t)
        if m:
            entry["nnz"] = int(m.group(1))
        kats["mul_dense"].append(entry)

    a, b, text = fn_body(sparse, "test_mul_vector")
    mats = matrices(text)
    assert len(mats) == 3
    kats["mul_vector"] = {"ref": f"src/sparse.rs:{a}-{b}", "v": [0, 1, 2, 3, 4],
                          "bad_dims_matrix": mats[0], "identity": mats[1],
                          "matrix": mats[2], "expected": [16, 7]}
    t = strip_comments(text)
    assert "vec![16,7]" in t.replace(" ", "") and "vec![0,1,2,3,4]" in t.replace(" ", "")

    # --- Dense KATs -----------------------------------------------------------------------------
    a, b, text = fn_body(dense, "init")
    mats = matrices(text)
    kats["dense"].append({"name": "init", "ref": f"src/dense.rs:{a}-{b}", "new_default_with_dims": [5, 7],
                          "equals_from_data_columns": mats[0]})
    a, b, text = fn_body(dense, "get_col")
    mats = matrices(text)
    kats["dense"].append({"name": "get_col", "ref": f"src/dense.rs:{a}-{b}", "columns": mats[0],
                          "get_col_2": [7, 8, 9]})
    assert "&[7,8,9]" in strip_comments(text).replace(" ", "")

    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
