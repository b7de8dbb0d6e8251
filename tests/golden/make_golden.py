"""Generate tests/golden/*.npz from the CPU oracle.

The reference cannot be imported in this image (`mlx` absent), so these
fixtures pin the ORACLE (regression + cross-implementation anchor), not the reference
binary: inputs are stored alongside the expected outputs so they do not depend on the
numpy RNG stream.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import datasets, vs_oracle  # noqa: E402

OUT = Path(__file__).resolve().parent


def case(name, db, q, k):
    blob = {"db": db, "q": q, "k": np.int32(k)}
    for metric in ("cosine", "euclidean", "dot_product"):
        ids, scores, _ = vs_oracle.search(q, db, k, metric)
        blob[f"ids_{metric}"] = ids
        blob[f"scores_{metric}"] = scores
    np.savez_compressed(OUT / f"{name}.npz", **blob)
    print(name, db.shape, q.shape, k)


if __name__ == "__main__":
    case("normal_512x48", datasets.make_db(512, 48, "normal"), datasets.make_queries(4, 48, "normal"), 10)
    case("uniform_300x100", datasets.make_db(300, 100, "uniform"), datasets.make_queries(3, 100, "uniform"), 7)
    db, q = datasets.make_adversarial(257, 64)
    case("adversarial_257x64", db, q, 12)
