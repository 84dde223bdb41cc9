"""Golden vectors for the raw power ingest, made by EXECUTING the reference's own code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_ingest.py
Writes tests/golden/ingest_vectors.npz (committed).

``ParseEK.pad_shorter_ping`` (convert/parse_base.py:686-730, a numpy-only staticmethod) and the module constant
``INDEX2POWER`` (parse_base.py:24) are lifted out of the module with ``ast`` and executed unmodified; the power
scaling statement of ``_parse_and_pad_datagram`` (parse_base.py:302, ``padded_arr.astype("float32") * INDEX2POWER``)
is applied to the padded array exactly as written there.  Nothing from the reference is copied into this
repository - only the numeric outputs are stored.
"""

import ast
import os

import numpy as np

REF = "/root/reference/echopype/convert/parse_base.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift():
    tree = ast.parse(open(REF).read())
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(getattr(t, "id", None) == "INDEX2POWER" for t in node.targets):
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "ParseEK":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "pad_shorter_ping":
                    item.decorator_list = []
                    item.returns = None
                    exec(compile(ast.fix_missing_locations(ast.Module(body=[item], type_ignores=[])), REF, "exec"), ns)
    return ns


def main():
    ns = lift()
    rng = np.random.default_rng(20261017)
    out = {"INDEX2POWER": np.float64(ns["INDEX2POWER"])}
    # ragged pings (shorter pings get NaN-padded) and equal-length pings (no padding branch)
    for name, lens in (("ragged", [37, 64, 64, 5, 50, 64, 1, 63]), ("equal", [48] * 6)):
        pings = [rng.integers(-32767, 32767, size=n, endpoint=True).astype(np.int16) for n in lens]
        pings[0][:4] = [-32767, 32767, 0, -1]  # extremes of the count range
        padded = ns["pad_shorter_ping"](pings)
        power = padded.astype("float32") * ns["INDEX2POWER"]  # parse_base.py:302, as written
        out[f"{name}_lens"] = np.array(lens)
        out[f"{name}_counts"] = np.concatenate(pings)
        out[f"{name}_power"] = power
    np.savez_compressed(os.path.join(HERE, "ingest_vectors.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
