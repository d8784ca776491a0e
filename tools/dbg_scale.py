import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import test_scale_gpu as t
from oracle import tt_oracle as orc
from tensor_networks_b200 import TensorTrain
eps = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5
y = t._decaying([64] * 4, parts=8, r_part=32, decade=1.5, seed=77)
tt = TensorTrain.from_cores(y).round(eps)
print(tt.ranks(), tt.last_round)
ref, _ = orc.svd_round([c.copy() for c in y], eps)
print(orc.ranks_of(ref))
