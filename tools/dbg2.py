import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from test_gpu_parity import teacher_forced_errors
r, mm, frames = teacher_forced_errors()
free = r["contacts"] == 0; hit = ~free
print("frames", frames, "mismatch", mm, "free", free.sum(), "hit", hit.sum())
print("free max pos/vel/dist", r["pos"][free].max(), r["vel"][free].max(), r["dist"][free].max())
for q in (0.5, 0.9, 0.99, 0.999, 1.0):
    print("hit q", q, np.quantile(r["pos"][hit], q), np.quantile(r["vel"][hit], q), np.quantile(r["dist"][hit], q))
scale = 1.0 + 35.0 / np.maximum(r["dmin"][hit], 1e-3)
print("scaled max pos", (r["pos"][hit]/scale).max(), "vel", (r["vel"][hit]/scale).max())
i = np.argsort(r["pos"][hit])[-8:]
print(np.c_[r["pos"][hit][i], r["vel"][hit][i], r["dmin"][hit][i], r["contacts"][hit][i]])
