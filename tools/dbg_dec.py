import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import oracle_lib, product_lib
from cases import c2_raster, smooth_field
import lerc_b200
prod, orc = product_lib(), oracle_lib()
h, w = 513, 2049
rng = np.random.default_rng(h * 7919 + w)
arr = (smooth_field(h, w) + rng.normal(0, 0.5, (h, w))).astype(np.float32)
for mz in (1.0, 0.01):
    s, b, _ = orc.encode(arr, mz)
    s0 = lerc_b200.stats()
    st, d, _ = prod.decode(b)
    _, d_o, _ = orc.decode(b)
    print(mz, "status", st, "fastdec", lerc_b200.stats()[4] - s0[4], "equal", np.array_equal(d.view(np.uint8), d_o.view(np.uint8)), len(b))
