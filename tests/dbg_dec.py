import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import oracle_lib, product_lib
from cases import c2_raster
import lerc_b200
prod, orc = product_lib(), oracle_lib()
for shape in [(65, 2049), (257, 300), (513, 2049), (513, 2050), (1024, 1024), (4096, 4096)]:
    img = c2_raster(*shape)
    s, b, _ = orc.encode(img, 0.01)
    s0 = lerc_b200.stats()
    st, d, _ = prod.decode(b)
    _, d_o, _ = orc.decode(b)
    print(shape, "status", st, "fastdec", lerc_b200.stats()[4] - s0[4], "equal", np.array_equal(d.view(np.uint8), d_o.view(np.uint8)))
