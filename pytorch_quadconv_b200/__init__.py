"""
pytorch_quadconv_b200 -- B200-native (sm_100a) implementation of the QuadConv layer hot path behind
the reference's own nn.Module API (torch_quadconv/__init__.py:4-6).

    from pytorch_quadconv_b200 import QuadConv, Mesh_MaxPool, MeshHandler

(The distribution is named `pytorch-quadconv_b200`; the importable package uses an underscore because
Python module names cannot contain '-'.)
"""
from .quadconv import QuadConv
from .mesh_pool import Mesh_MaxPool, MeshMaxPool
from .mesh_handler import MeshHandler
from .graphs import graphed

__all__ = ["QuadConv", "Mesh_MaxPool", "MeshMaxPool", "MeshHandler", "graphed"]
