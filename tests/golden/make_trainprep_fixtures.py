"""Generates tests/golden/trainprep.json (run in the BUILD container: imports the reference's own functions from /root/reference).

Randomised SAM records (both strands; hard / soft clips; M I D N P = X operations; alignments that begin or end with a non-match
operation, which exercises the reference's trimming loops and the end_clipped_bases quirk of alignutils.py:137) are pushed through
the REFERENCE implementation of
    alignutils.parse_sam_record, preprocessing.clean_read_map_ref, preprocessing.fix_raw_starts_for_clipped_bases,
    preprocessing.get_base_label / get_base_color, nanorevtrainutils.get_trainning_input (AST-extracted: its module imports
    pandas / albacore / keras at the top)
and inputs + outputs are written as one JSON document that tests/test_trainprep.py replays against nanoreviser_b200/trainprep.py.
Deterministic (seeded); the output is committed.
"""
import ast
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
from nanorevutils import alignutils, preprocessing  # noqa: E402


def extract(src_path, names, extra_globals):
    tree = ast.parse(open(src_path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    g = dict(extra_globals)
    exec(compile(ast.Module(body=body, type_ignores=[]), src_path, "exec"), g)
    return g


rng = np.random.default_rng(20261017)
ACGT = "ACGT"


def rand_seq(n):
    return "".join(rng.choice(list(ACGT), size=n))


def make_case(case_id):
    """An alignment built in READ orientation, then expressed as a SAM record (reference orientation)."""
    strand = "-" if rng.random() < 0.5 else "+"
    n_ops = int(rng.integers(1, 14))
    body_types = list(rng.choice(list("MMMM=XIDNP"), size=n_ops))
    if not any(t in "M=X" for t in body_types):
        body_types[int(rng.integers(0, n_ops))] = "M"
    if rng.random() < 0.6:                         # most cases: a proper alignment that starts and ends on a match
        body_types[0] = "M"; body_types[-1] = "M"
    ops = [(int(rng.integers(1, 9)), t) for t in body_types]
    q, t = [], []                                    # query (read orientation) and target (read orientation)
    for n, typ in ops:
        if typ in "M=X":
            ref = rand_seq(n)
            qs = list(ref)
            for i in range(n):
                if typ == "X" or (typ == "M" and rng.random() < 0.2):
                    qs[i] = rng.choice([c for c in ACGT if c != ref[i]])
            q.append("".join(qs)); t.append(ref)
        elif typ in "IP":
            q.append(rand_seq(n))
        else:
            t.append(rand_seq(n))
    q, t = "".join(q), "".join(t)
    s0 = int(rng.integers(0, 6)) if rng.random() < 0.5 else 0
    s1 = int(rng.integers(0, 6)) if rng.random() < 0.5 else 0
    h0 = int(rng.integers(1, 5)) if rng.random() < 0.25 else 0
    h1 = int(rng.integers(1, 5)) if rng.random() < 0.25 else 0
    full_ops = ([(h0, "H")] if h0 else []) + ([(s0, "S")] if s0 else []) + ops + ([(s1, "S")] if s1 else []) + ([(h1, "H")] if h1 else [])
    q_full = rand_seq(s0) + q + rand_seq(s1)
    pre, post = rand_seq(int(rng.integers(0, 30))), rand_seq(int(rng.integers(0, 30)))
    if strand == "+":
        cigar = "".join("%d%s" % o for o in full_ops)
        seq, g = q_full, t
    else:
        cigar = "".join("%d%s" % o for o in full_ops[::-1])
        seq, g = alignutils.rev_comp(q_full), alignutils.rev_comp(t)
    chrom = "chr%d" % (case_id % 3)
    genome = {chrom: pre + g + post, "other": rand_seq(10)}
    rec = {"qName": "read%d" % case_id, "flag": "16" if strand == "-" else "0", "rName": chrom, "pos": str(len(pre) + 1),
           "mapq": "40", "cigar": cigar, "rNext": "*", "pNext": "0", "tLen": "0", "seq": seq, "qual": "*"}
    return rec, genome


def main():
    cases = []
    for cid in range(400):
        rec, genome = make_case(cid)
        out = {"sam": rec, "genome": genome}
        try:
            rv, fv, mv, loc, sc, ec = alignutils.parse_sam_record(dict(rec), genome)
            out["parse"] = {"read": "".join(rv), "ref": "".join(fv), "map": "".join(mv), "loc": loc, "sc": int(sc), "ec": int(ec)}
            if len(rv) == len(fv) == len(mv) and len(rv) >= 1:
                c = preprocessing.clean_read_map_ref(rv, mv, fv)
                out["clean"] = ["".join(x) for x in c]
                out["labels"] = [[int(preprocessing.get_base_label(ch)) for ch in c[2]], [int(preprocessing.get_base_label(ch)) for ch in c[3]],
                                 [int(preprocessing.get_base_color(ch)) for ch in c[0]]]
        except AssertionError as e:
            out["error"] = "AssertionError"
        except Exception as e:                     # noqa: BLE001
            out["error"] = type(e).__name__
        cases.append(out)
    # fix_raw_starts_for_clipped_bases
    fixes = []
    for k in range(40):
        n = int(rng.integers(12, 60))
        length = rng.choice([2., 3., 5., 10., 15.], size=n)
        starts = np.concatenate([[0], np.cumsum(length[:-1])]).astype(np.int64)
        a0 = int(rng.integers(0, 500))
        evm = rng.normal(100, 10, n).astype(np.float32); evs = rng.random(n).astype(np.float32)
        sc, ec = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        o = preprocessing.fix_raw_starts_for_clipped_bases(sc, ec, starts.copy(), length.copy(), a0, evm.copy(), evs.copy())
        fixes.append({"sc": sc, "ec": ec, "starts": starts.tolist(), "length": length.tolist(), "a0": a0, "evm": evm.tolist(), "evs": evs.tolist(),
                      "out": [np.asarray(o[0]).tolist(), np.asarray(o[1]).tolist(), int(o[2]), np.asarray(o[3]).tolist(), np.asarray(o[4]).tolist()]})
    # get_trainning_input on three synthetic per-read .npz files
    g = extract(os.path.join(REF, "nanorevutils/nanorevtrainutils.py"), {"get_trainning_input"}, {"np": np, "os": os})
    tensors = []
    for W in (5, 13):
        with tempfile.TemporaryDirectory() as d:
            reads = []
            for k in range(3):
                n = int(rng.integers(20, 40))
                a = dict(refvals=rng.integers(0, 6, n), refvals2=rng.integers(1, 6, n), readVals=rng.choice([250, 180, 100, 30, 0], n),
                         signal_mean=rng.normal(500, 50, n), signal_std=rng.random(n) * 20, signal_len=rng.choice([2., 5., 10.], n),
                         ab_mean=rng.normal(100, 10, n).astype(np.float32), ab_std=rng.random(n).astype(np.float32),
                         signal_x=rng.normal(0, 1, (n, 50)), mapvals=np.array(list("M" * n)), starts=np.arange(n) * 5,
                         scale=np.float64(rng.integers(20, 70)), shift=np.float64(rng.integers(400, 600)))
                np.savez(os.path.join(d, "r%d" % k), **a)
                reads.append({kk: np.asarray(v).tolist() for kk, v in a.items() if kk not in ("mapvals",)})
            order = [f for f in os.listdir(d)]
            x, sx, y, y2 = g["get_trainning_input"](True, d, W)
        tensors.append({"W": W, "order": order, "reads": reads, "x_sum": float(x.sum()), "x_shape": list(x.shape),
                        "x_first": x[0].tolist(), "x_last": x[-1].tolist(), "sx_shape": list(sx.shape), "sx_sum": float(sx.sum()),
                        "sx_last_row": sx[-1, -1].tolist(), "y": y[:, 0].tolist(), "y2": y2[:, 0].tolist()})
    doc = {"cases": cases, "fixes": fixes, "tensors": tensors}
    with open(os.path.join(HERE, "trainprep.json"), "w") as fp:
        json.dump(doc, fp)
    n_err = sum(1 for c in cases if "error" in c)
    print("cases %d (reference raised on %d), fixes %d, tensors %d" % (len(cases), n_err, len(fixes), len(tensors)))


if __name__ == "__main__":
    main()
