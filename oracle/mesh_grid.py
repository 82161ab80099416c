"""CPU stand-in for ``models/mesh_grid.py::MeshGrid`` - TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Same protocol as the reference class (``compute_distance``, ``get_vertices_torch``, ``get_vertex_normal_torch``,
``get_number_of_vertices``; ``models/mesh_grid.py:45-150``), neighbour search by the exact-KNN restatement of FRNN
(``oracle/knn.py``), blend by ``oracle.field.mesh_distance`` (``mesh_grid.py:121-144``, differentiable in the indicator
vector / weight exactly as the reference).  Used by the tests and by ``bench.py --workload train --impl reference`` to run
the reference's training-step arithmetic on the host cores; the product never imports it."""
from __future__ import annotations

import numpy as np
import torch


class OracleMeshGrid:
    def __init__(self, mesh):
        self.mesh = mesh
        self.vertices = torch.as_tensor(np.asarray(mesh.vertices), dtype=torch.float32)
        self.vertex_normals = torch.as_tensor(np.asarray(mesh.vertex_normals), dtype=torch.float32)
        self.distance_method = "frnn"

    def get_number_of_vertices(self):
        return self.vertices.shape[0]

    def get_vertex_normal_torch(self):
        return self.vertex_normals

    def get_vertices_torch(self):
        return self.vertices

    def compute_distance(self, xyz, indicator_vector=None, indicator_weight=0.1, K=8):
        from oracle.field import mesh_distance
        ind = self.vertex_normals if indicator_vector is None else indicator_vector
        return mesh_distance(xyz, self.vertices, ind, indicator_weight, K)
