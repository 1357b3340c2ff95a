"""small battery for `compute-sanitizer --tool memcheck|racecheck python tools/san_small.py`: every round-2 kernel once at shapes with
ragged edges (single-pass encoder incl. several packing passes, stream decoder, tile batch on the TMA encoder and on the older kernel,
8-bit Huffman kernels, parallel mask RLE both ways, masked offsets + verification)."""
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib, encode_tiles, decode_tiles
from cases import c2_raster, c4_raster
prod = product_lib()
rng = np.random.default_rng(3)
yy, xx = np.mgrid[0:300, 0:1024]
coast = (xx > 350 + 120 * np.sin(yy / 37.0) + 25 * np.sin(yy / 5.0)).astype(np.uint8)
for name, arr, mz, kw in [("f32", c2_raster(264, 523), 0.01, {}), ("f32b", c2_raster(512, 1024), 0.01, {}), ("f32_wide", (c2_raster(64, 2056) * 37).astype(np.float32), 1e-5, {}),
                          ("i16", np.clip(c2_raster(200, 333) * 3 - 2000, -32768, 32767).astype(np.int16), 0, {}),
                          ("masked", c2_raster(128, 256), 0.01, {"mask": (rng.random((128, 256)) > 0.1).astype(np.uint8)}),
                          ("coast", c2_raster(300, 1024), 0.01, {"mask": coast}),
                          ("coast_i16", np.clip(c2_raster(300, 1024) * 3, -32768, 32767).astype(np.int16), 0, {"mask": coast}),
                          ("u8x3", c4_raster(96, 128), 0, {"n_depth": 3}), ("u8x3_ragged", c4_raster(45, 1100), 0, {"n_depth": 3}), ("u8", c4_raster(70, 4128)[..., 0].copy(), 0, {}),
                          ("f64", c2_raster(64, 200).astype(np.float64), 0.001, {})]:
    st, blob, _ = prod.encode(arr, mz, **kw)
    st2, d, m = prod.decode(blob)
    print(name, st, st2, len(blob))
for name, arr, tr, tc in [("tiles_tma", c2_raster(512, 768), 256, 256), ("tiles_old", c2_raster(200, 330), 64, 64)]:
    st, blobs, _ = encode_tiles(prod, arr, tr, tc, 0.01)
    st2, dec = decode_tiles(prod, blobs, np.float32, arr.shape[0], arr.shape[1], tr, tc)
    print(name, st, st2, len(blobs))
