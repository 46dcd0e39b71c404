"""Helpers to read tests/golden/*.npz (produced by oracle/gen_golden.py from the unmodified reference)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.name = name

    def __getitem__(self, k):
        return self.z[k]

    def has(self, k):
        return k in self.z.files

    @property
    def k(self):
        return self.meta["parts"]

    def iter_stats(self):
        """rows: (frame, alpha, E, |g|^2); the first row of a frame has alpha=0 ("after initX")."""
        rows = []
        for line in str(self.z["iterStats"]).strip().splitlines():
            t = line.split()
            rows.append([float(v) for v in t])
        return np.asarray(rows)

    def states(self):
        out = sorted({f.split("/")[0] for f in self.z.files if f.startswith("frame") or f.startswith("kernel")})
        return out


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
