import sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import oracle_lib, product_lib, ref_lib
from cases import c2_raster
import lerc_b200
prod = product_lib(); ref = ref_lib() or oracle_lib()
img = c2_raster(4096, 4096)
t=time.time(); s, b_r, _ = ref.encode(img, 0.01); print("ref enc", time.time()-t)
s0 = lerc_b200.stats()
s, b_p, _ = prod.encode(img, 0.01)
print("status", s, "equal", b_p == b_r, len(b_p), len(b_r), "fast", lerc_b200.stats()[3]-s0[3])
_, d_r, _ = ref.decode(b_r)
t=time.time(); st, d_p, _ = prod.decode(b_r); print("dec", st, time.time()-t, np.array_equal(d_p.view(np.uint8), d_r.view(np.uint8)), "fastdec", lerc_b200.stats()[4])
