"""Quadrature / coarsening maps selectable by name from MeshHandler.construct(quad_map=...)
(reference torch_quadconv/utils/quadrature.py).  Only `mfnus` is usable from construct() in the
reference today (every other map returns torch tensors and fails at mesh_handler.py:196, SURVEY.md
section 2 row 6), so only `mfnus` is provided; asking for another name raises AttributeError exactly
as getattr() on the reference module would for an unknown map."""
from .mfnus import mfnus  # noqa: F401


def param_quad(*args, **kwargs):
    raise TypeError("quad_map='param_quad' cannot build a coarser level (it fails upstream at "
                    "mesh_handler.py:196 as well); use quad_map='mfnus' or a single-level construct([N])")
