"""Copies the INPUT columns of the reference's only data file (discourse_westbrook/westbrook.csv:
1 438 basketball shots, columns x and result) into tests/golden/westbrook_xy.npz.  Data, not code;
the reference directory does not exist on the GPU box.  x has only 1 073 unique values, so the exact-GP
Gram matrix of models/westbrook_exact.stan:17-21 is singular without its jitter."""
import csv
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rows = list(csv.DictReader(open("/root/reference/discourse_westbrook/westbrook.csv")))
x = np.array([float(r["x"]) for r in rows])
y = np.array([1.0 if r["result"] == "made" else 0.0 for r in rows])
np.savez(os.path.join(HERE, "westbrook_xy.npz"), x=x, y=y)
print(len(x), len(np.unique(x)), x.min(), x.max(), y.mean())
