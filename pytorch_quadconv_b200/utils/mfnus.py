"""
Host-side restatement of the reference's node-elimination coarsening (torch_quadconv/utils/MFNUS2.py:11-45,
wrapped by utils/quadrature.py:190-194).  Setup-time only: it runs once in
MeshHandler.construct and its elimination map is the input contract of Mesh_MaxPool.

Algorithm: lexsort the nodes bottom-up, find K nearest neighbours, walk the sorted nodes; a node that
is still alive eliminates every later neighbour closer than fc x (its nearest-neighbour distance).
Returned: kept points (in sorted order) and {kept original index: [eliminated original indices]}.
"""
from __future__ import annotations

import numpy as np
from sklearn.neighbors import NearestNeighbors


def mfnus_eliminate(points, fc: float = 1.5, K: int = 10):
    xy = np.asarray(points)
    if xy.shape[0] < xy.shape[1]:
        xy = xy.T
    n = xy.shape[0]
    order = np.lexsort(xy.T, axis=0)            # MFNUS2.py:20 (last coordinate is the primary key)
    pts = xy[order]
    dist, nbr = NearestNeighbors(n_neighbors=K + 1, algorithm="auto").fit(pts).kneighbors(pts)
    alive = np.ones(n, dtype=bool)
    elim_map = {}
    for k in range(n):                           # MFNUS2.py:29-38
        if not alive[k]:
            continue
        close = np.nonzero(dist[k, 1:] < fc * dist[k, 1])[0]
        victims = nbr[k, close + 1]
        victims = victims[victims >= k]          # only nodes above the present one
        alive[victims] = False
        elim_map[order[k]] = [order[v] for v in victims]
    kept = pts[alive]
    return kept, elim_map


def mfnus(input_points, num_points, appx_ds: float = 1.65, K: int = 10):
    """Same call signature and return triple as utils/quadrature.py:190-194."""
    pts = input_points.detach().cpu().numpy() if hasattr(input_points, "detach") else np.asarray(input_points)
    new_xy, elim_map = mfnus_eliminate(pts, fc=appx_ds, K=K)
    return new_xy, None, elim_map
