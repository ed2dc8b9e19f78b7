"""TEST INFRASTRUCTURE: CPU oracle of distance3d_b200 (see oracle/src/d3d_oracle.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.
"""
