"""CPU oracle for the NeuMesh rendering hot path - TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this package; ``neumesh_b200`` never does.
"""
