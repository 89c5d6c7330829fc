"""Writes tests/golden/pairs_c3_sha256.json: SHA-256 of the reference's eval_indices (int64 [P,2], row-major bytes)
for the C3 point set (torch.manual_seed(0); torch.rand(100000, 3); r = (150/(4 pi N))^(1/3)).

The pair list is computed twice on the CPU and must agree bit for bit before the hash is written:
  (1) oracle.eval_indices_torch -- the reference's own ATen ops (quadconv.py:178-183) on 2048-row blocks
      (the unblocked reference needs a 120 GB [No,Ni,3] tensor at this size);
  (2) oracle.eval_indices_c     -- the plain-C brute force (fmaf/sqrtf form of vector_norm).
Run from the repo root:  python tests/golden/make_pairs_c3_hash.py     (~2-3 minutes on 8 cores)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import quadconv_oracle as O  # noqa: E402

N = 100_000
torch.manual_seed(0)
pts = torch.rand(N, 3)
r = (150 / (4 * np.pi * N)) ** (1 / 3)
t0 = time.time()
pc = O.eval_indices_c(pts.numpy(), pts.numpy(), r)
t1 = time.time()
pt = O.eval_indices_torch(pts, pts, r).numpy()
t2 = time.time()
assert pc.shape == pt.shape and np.array_equal(pc, pt), "C oracle and blocked ATen restatement disagree"
# pairs whose distance is within 1 ulp of the radius (SURVEY hard part 1 expects ~1 at this size)
d = np.linalg.norm((pts[pc[:, 0]] - pts[pc[:, 1]]).numpy().astype(np.float64), axis=1)
r32 = float(np.float32(r))
near = int(np.sum(np.abs(d - r32) <= np.spacing(np.float32(r32))))
out = {"N": N, "d": 3, "seed": 0, "radius": r, "P": int(pc.shape[0]),
       "sha256": hashlib.sha256(np.ascontiguousarray(pc).tobytes()).hexdigest(),
       "pairs_within_1ulp_of_radius": near,
       "seconds": {"c_oracle": round(t1 - t0, 1), "aten_blocked": round(t2 - t1, 1)}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "pairs_c3_sha256.json"), "w"), indent=1)
print(out)
