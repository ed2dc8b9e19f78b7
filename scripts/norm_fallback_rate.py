import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from distance3d_b200 import _lib
from oracle import cpu_oracle as O
from test_norm_gpu import _vectors
v = _vectors()
got = _lib.debug_norm(v, 0); ref = O.norm(v)
bad = got != ref
print("vectors", len(v), "mismatches without fallback:", int(bad.sum()))
rs = np.random.RandomState(1)
w = rs.randn(4000000, 3)
got = _lib.debug_norm(w, 0); ref = O.norm(w)
print("random 4M: mismatches without fallback:", int((got != ref).sum()))
