"""Writes tests/golden/tiny_examples.json: a 3-entry sample history (2, 0 and 1 samples) and its bincode bytes as the
ORACLE encodes them (oracle/learn.hpp: bincode 1.3.1 + ndarray 0.13 serde layout restated; the reference holds no
.examples fixture, so this pins product == oracle == the independent parser in tests/test_learn_cpu.py, not the crates).
Run from the repo root:  python tests/golden/make_examples_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as ge  # noqa: E402

ge.build_oracle()
import oracle_api as o  # noqa: E402

rng = np.random.default_rng(20261018)
counts = [2, 0, 1]
n = sum(counts)
boards = (rng.random((n, 2, 6, 7)) < 0.3).astype(np.float32)
pis = rng.random((n, 7)).astype(np.float32)
pis /= pis.sum(1, keepdims=True)
vs = np.array([1.0, -1.0, 1.0], np.float32)
blob = o.examples_encode(counts, boards, pis, vs)
json.dump({"counts": counts, "boards": boards.reshape(n, -1).tolist(), "pis": pis.tolist(), "vs": vs.tolist(),
           "bincode_hex": blob.hex()}, open(os.path.join(HERE, "tiny_examples.json"), "w"))
print(len(blob), "bytes")
