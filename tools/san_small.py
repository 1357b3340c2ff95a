import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib
from cases import c2_raster, c4_raster
prod = product_lib()
rng = np.random.default_rng(3)
for name, arr, mz, kw in [("f32", c2_raster(264, 523), 0.01, {}), ("f32b", c2_raster(512, 1024), 0.01, {}),
                          ("i16", np.clip(c2_raster(200, 333) * 3 - 2000, -32768, 32767).astype(np.int16), 0, {}),
                          ("masked", c2_raster(128, 256), 0.01, {"mask": (rng.random((128, 256)) > 0.1).astype(np.uint8)}),
                          ("u8x3", c4_raster(96, 128), 0, {"n_depth": 3}), ("f64", c2_raster(64, 200).astype(np.float64), 0.001, {})]:
    st, blob, _ = prod.encode(arr, mz, **kw)
    st2, d, m = prod.decode(blob)
    print(name, st, st2, len(blob))
