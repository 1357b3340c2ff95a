"""Seeded synthetic rasters shared by the oracle-pinning tests (CPU) and the CUDA parity tests (GPU).

Each case is (name, array, maxZErr, kwargs for LercLib.encode).  The generators follow SURVEY.md 8(d):
smooth sinusoid fields plus noise for the float configs, clipped integer versions of the same for the
integer types, plus the distributions that flip an image-global decision of the encoder
(all-integer floats, pre-rounded floats, constant images, uniform noise, masks, NaNs, LUT-friendly data).
"""
import numpy as np


def smooth_field(h, w, phase=0.0):
    yy, xx = np.mgrid[0:h, 0:w]
    return 1000 + 300 * np.sin(xx / 97 + phase) * np.cos(yy / 131) + 50 * np.sin(xx / 13 + yy / 17 + phase)


def c2_raster(h, w, seed=1234, phase=0.0):
    """BASELINE config 2/3/5 generator: float32 terrain-like field with N(0, 0.5) noise."""
    rng = np.random.default_rng(seed)
    return (smooth_field(h, w, phase) + rng.normal(0, 0.5, (h, w))).astype(np.float32)


def c4_raster(h, w, seed=7, sigma=2.0):
    """BASELINE config 4 generator: uint8, nDepth=3 (bluemarble-like)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    r = 128 + 100 * np.sin(xx / 53) * np.cos(yy / 71)
    g = 128 + 90 * np.sin(xx / 31 + 1) * np.cos(yy / 47)
    b = 100 + 80 * np.cos(xx / 91) * np.sin(yy / 23)
    img = np.stack([r, g, b], axis=-1) + rng.normal(0, sigma, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def all_cases(h=257, w=300, seed=1):
    rng = np.random.default_rng(seed)
    smooth = smooth_field(h, w)
    f32 = (smooth + rng.normal(0, 0.5, (h, w))).astype(np.float32)
    cases = []

    def add(name, arr, mz, **kw):
        cases.append((name, arr, mz, kw))

    add("f32_noisy_0.01", f32, 0.01)
    add("f32_noisy_0.001", f32, 0.001)
    add("f32_noisy_1", f32, 1.0)
    add("f32_noisy_100", f32, 100.0)
    add("f64_noisy_0.01", f32.astype(np.float64) + 1e-7, 0.01)
    add("f32_rounded_0.1", np.round(f32, 1), 0.01)
    add("f32_allint", np.round(f32), 0.01)
    add("f32_allint_3.7", np.round(f32), 3.7)
    add("f32_const", np.full((h, w), 3.25, np.float32), 0.01)
    add("f32_zero", np.zeros((h, w), np.float32), 0.01)
    add("f32_negzero_min", np.where(rng.random((h, w)) < 0.01, np.float32(-0.0), np.abs(f32)).astype(np.float32), 0.01)
    add("f32_huge_range", (f32 * np.float32(1e6)).astype(np.float32), 1e-4)
    for dt in [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32]:
        info = np.iinfo(dt)
        name = np.dtype(dt).name
        a = np.clip(smooth / 8 + rng.normal(0, 2, (h, w)), info.min, info.max).astype(dt)
        add(f"{name}_smooth_lossless", a, 0)
        add(f"{name}_smooth_lossy2", a, 2)
        add(f"{name}_uniform_lossless", rng.integers(info.min, int(info.max) + 1, (h, w), dtype=np.int64).astype(dt), 0)
    mask = (rng.random((h, w)) > 0.3).astype(np.uint8)
    mask2 = np.ones((h, w), np.uint8)
    mask2[50:120, 30:200] = 0
    add("f32_masked_random", f32, 0.01, mask=mask)
    add("f32_masked_rect", f32, 0.01, mask=mask2)
    add("u8_masked_rect", np.clip(smooth / 8, 0, 255).astype(np.uint8), 0, mask=mask2)
    add("u8_masked_random", np.clip(smooth / 8, 0, 255).astype(np.uint8), 0, mask=mask)
    add("f32_all_masked", f32, 0.01, mask=np.zeros((h, w), np.uint8))
    rgb = c4_raster(h, w, seed=7)
    add("u8_depth3_lossless", rgb, 0, n_depth=3)
    add("u8_depth3_lossless_masked", rgb, 0, n_depth=3, mask=mask2)
    add("i8_depth3_lossless", (rgb.astype(np.int16) - 128).astype(np.int8), 0, n_depth=3)
    add("u16_depth3_lossless", rgb.astype(np.uint16) * 7, 0, n_depth=3)
    add("i32_depth3_lossless", rgb.astype(np.int32) * 1000 - 5000, 0, n_depth=3)
    add("f32_depth3_0.01", rgb.astype(np.float32) * np.float32(1.37), 0.01, n_depth=3)
    bands = np.stack([f32, f32 * 2, f32 + 5])
    add("f32_3bands", bands, 0.01, n_bands=3)
    add("f32_3bands_3masks", bands, 0.01, n_bands=3, mask=np.stack([mask, mask2, mask]))
    add("f32_3bands_1mask", bands, 0.01, n_bands=3, mask=mask2)
    fn = f32.copy()
    fn[10:20, 10:50] = np.nan
    add("f32_nan", fn, 0.01)
    add("f32_nan_3bands", np.stack([f32, fn, f32]), 0.01, n_bands=3)
    add("f32_stepped_lut", (np.floor(smooth / 40) * 40).astype(np.float32), 0.01)
    add("u8_classes_lut", (rng.integers(0, 4, (h, w)) * 60).astype(np.uint8), 0)
    cls = np.repeat(np.repeat(rng.integers(0, 5, (h // 4 + 1, w // 4 + 1)), 4, 0), 4, 1)[:h, :w]
    add("u16_class_blocks_lut", (cls * 1000).astype(np.uint16), 0)
    add("i16_class_blocks_lut", (cls * 1000 - 2000).astype(np.int16), 0)
    add("f32_lossless_raw", f32, 0)
    add("f64_lossless_raw", f32.astype(np.float64) * 1.0000001, 0)
    add("f32_tiny_3x5", f32[:3, :5].copy(), 0.01)
    add("u8_tiny_1x1", np.array([[7]], np.uint8), 0)
    add("f32_8x8", f32[:8, :8].copy(), 0.01)
    add("i16_smooth_low_bitrate", (smooth / 300).astype(np.int16), 0)   # < 1.5 bpp -> 16x16 blocks
    return cases


def tile_cases(seed=5):
    """(name, raster, tileRows, tileCols, maxZErr) for the tile batch entry points: ordinary tiles plus tiles that flip an
    image-global decision of the encoder (constant, all-integer, pre-rounded, NaN, LUT-friendly, flat -> 16x16 / one sweep)."""
    rng = np.random.default_rng(seed)
    cases = []
    f = c2_raster(200, 330)
    cases.append(("f32_64x64_0.01", f, 64, 64, 0.01))
    cases.append(("f32_40x56_edge_0.01", f, 40, 56, 0.01))
    cases.append(("f32_256x256_single", c2_raster(256, 256), 256, 256, 0.01))
    cases.append(("f32_one_tile_bigger", f, 512, 512, 0.01))
    cases.append(("f32_8x8_tiles", f[:40, :48].copy(), 8, 8, 0.01))
    cases.append(("f64_48x48_0.001", f.astype(np.float64) + 1e-7, 48, 48, 0.001))
    cases.append(("i16_64x64_lossless", np.clip(smooth_field(200, 330) / 8 + rng.normal(0, 2, (200, 330)), -3e4, 3e4).astype(np.int16), 64, 64, 0))
    cases.append(("u16_64x80_lossy2", (smooth_field(130, 250) + rng.normal(0, 3, (130, 250))).astype(np.uint16), 64, 80, 2))
    cases.append(("i32_32x32_lossless", (smooth_field(100, 100) * 1000 + rng.normal(0, 50, (100, 100))).astype(np.int32), 32, 32, 0))
    cases.append(("u8_64x64_lossless", np.clip(smooth_field(128, 192) / 8 + rng.normal(0, 2, (128, 192)), 0, 255).astype(np.uint8), 64, 64, 0))
    mixed = c2_raster(192, 256)
    mixed[0:64, 0:64] = 3.25                                   # constant tile
    mixed[0:64, 64:128] = np.round(mixed[0:64, 64:128])        # all-integer tile
    mixed[0:64, 128:192] = np.round(mixed[0:64, 128:192], 1)   # on a 0.1 grid: maxZError raised
    mixed[64:128, 0:64] = (np.floor(mixed[64:128, 0:64] / 40) * 40)     # LUT friendly
    mixed[64:128, 64:128] = 0.0                                # all zero
    mixed[70:80, 130:150] = np.nan                             # NaN -> mask
    mixed[128:192, 0:64] = np.floor(smooth_field(64, 64) / 300)         # flat: 16x16 blocks
    mixed[128:192, 64:128] = rng.random((64, 64)).astype(np.float32) * 1e30   # raw / one sweep
    cases.append(("f32_mixed_decisions", mixed, 64, 64, 0.01))
    cases.append(("f32_lossless", f[:100, :100].copy(), 50, 50, 0))
    whole = np.round(f)                                        # every tile all-integer: second fused pass at max(0.5, floor(maxZErr))
    cases.append(("f32_all_integer_0.01", whole, 64, 64, 0.01))
    cases.append(("f32_all_integer_3.7", whole, 64, 64, 3.7))
    whole0 = whole.copy()
    whole0[0:64, 0:64] = 7.0                                   # ... but one of them constant
    cases.append(("f32_all_integer_one_const", whole0, 64, 64, 0.01))
    cases.append(("f64_all_integer_0.3", np.round(f.astype(np.float64) * 3), 48, 48, 0.3))
    # constant tiles (header-only blobs, decided on the device): values on and off the decimal grids maxZError may be raised to,
    # integers, zeros of both signs, a tile mixing +0 and -0, constant tiles at the ragged edge
    cst = c2_raster(192, 330)
    for k, v in enumerate([3.25, 0.1, 7.0, 0.0, -0.0, 1e-3, 123.456, -2.5]):
        y, x = divmod(k, 4)
        cst[y * 64:(y + 1) * 64, x * 64:(x + 1) * 64] = v
    cst[128:192, 0:64] = 0.0
    cst[128:192:2, 0:64] = -0.0
    cst[128:192, 320:330] = 42.5
    for mz in (0.01, 0.3, 0.0004):
        cases.append((f"f32_const_tiles_{mz}", cst, 64, 64, mz))
    cases.append(("f64_const_tiles_0.01", cst.astype(np.float64), 64, 64, 0.01))
    ci = np.clip(smooth_field(128, 200) / 8 + rng.normal(0, 2, (128, 200)), -3e4, 3e4).astype(np.int16)
    ci[0:64, 0:64] = -17
    ci[64:128, 128:192] = 0
    ci[64:128, 192:200] = 5
    cases.append(("i16_const_tiles_lossless", ci, 64, 64, 0))
    cases.append(("i32_const_tiles_lossy3", ci.astype(np.int32) * 5, 64, 64, 3))
    # rasters that are whole multiples of the tile size with tiles made of whole micro-blocks: the TMA-staged persistent encoder
    # (lerc_encode_tile.cuh, BATCH): several block rows per encoder tile, every pixel type, tiles as wide as an encoder tile, values
    # wider than 16 bits (several packing passes), decisions flipped per image
    g = c2_raster(512, 768)
    cases.append(("f32_256x256_config5_shape", g, 256, 256, 0.01))
    cases.append(("f32_128x256_0.1", g, 128, 256, 0.1))
    cases.append(("f32_16x16_tiles", g[:64, :96].copy(), 16, 16, 0.01))
    cases.append(("f32_8x1024_wide_tiles", c2_raster(24, 2048), 8, 1024, 0.01))
    cases.append(("f32_264x512_partial_last_encoder_tile", c2_raster(264, 1024), 264, 512, 0.01))
    cases.append(("f32_wide_values_1e-5", (g[:256, :256] * 37.0 + rng.normal(0, 30, (256, 256))).astype(np.float32), 128, 128, 1e-5))
    cases.append(("f64_64x64_exact", g[:192, :256].astype(np.float64) + 1e-9, 64, 64, 0.001))
    cases.append(("f64_64x512_wide", g[:128, :1024 - 256].astype(np.float64) * 1.0000001, 64, 512, 1e-6))
    gi = np.clip(smooth_field(256, 384) * 3 + rng.normal(0, 4, (256, 384)), -3e4, 3e4)
    cases.append(("i16_128x128_exact", gi.astype(np.int16), 128, 128, 0))
    cases.append(("u16_64x128_exact_lossy", (gi + 32768).astype(np.uint16), 64, 128, 3))
    cases.append(("i32_128x64_exact", (gi * 1000).astype(np.int32), 128, 64, 0))
    cases.append(("u32_64x64_exact_lossy", (gi * 100 + 4e6).astype(np.uint32), 64, 64, 2))
    mix2 = c2_raster(256, 512)
    mix2[0:128, 0:128] = -7.5                                   # constant
    mix2[0:128, 128:256] = np.round(mix2[0:128, 128:256])       # all-integer
    mix2[128:256, 0:128] = np.floor(mix2[128:256, 0:128] / 40) * 40      # LUT friendly
    mix2[130:140, 300:310] = np.nan
    mix2[128:256, 384:512] = rng.random((128, 128)).astype(np.float32) * 1e30
    cases.append(("f32_exact_mixed_decisions", mix2, 128, 128, 0.01))
    cases.append(("f32_exact_all_integer", np.round(g[:256, :256]), 128, 128, 0.01))
    cases.append(("f32_every_tile_const", np.repeat(np.repeat(rng.integers(0, 4, (3, 5)).astype(np.float32) * np.float32(1.5), 32, 0), 32, 1), 32, 32, 0.01))
    return cases


def nodata_cases(seed=11):
    """(name, array [nBands?][rows][cols][nDepth?], maxZErr, kwargs incl. uses_no_data / no_data) for the _4D calls:
    noData values beside valid values of the same pixel (nDepth > 1), pixels that are noData in every depth, NaN together
    with noData, noData close to / far from the valid range, integer and float types, all-integer floats, several bands."""
    rng = np.random.default_rng(seed)
    h, w = 48, 60
    cases = []

    def add(name, arr, mz, **kw):
        cases.append((name, arr, mz, kw))

    base = (rng.random((h, w, 3)) * 100 + 50).astype(np.float32)
    a = base.copy(); a[3:9, 5:20, 1] = -9999; a[20:24, 10:14, :] = -9999
    add("f32_d3_nodata_far", a, 0.01, n_depth=3, uses_no_data=[1], no_data=[-9999.0])
    add("f32_d3_nodata_far_lossless", a, 0, n_depth=3, uses_no_data=[1], no_data=[-9999.0])
    b = base.copy(); b[3:9, 5:20, 1] = 49.9995; b[20:24, 10:14, :] = 49.9995
    add("f32_d3_nodata_close", b, 0.01, n_depth=3, uses_no_data=[1], no_data=[49.9995])
    c = base.copy(); c[3:9, 5:20, 1] = 1e30
    add("f32_d3_nodata_above", c, 0.01, n_depth=3, uses_no_data=[1], no_data=[1e30])
    d = base.copy(); d[3:9, 5:20, 1] = -9999; d[30:33, 2:9, 2] = np.nan; d[40:42, 40:44, :] = np.nan
    add("f32_d3_nodata_and_nan", d, 0.01, n_depth=3, uses_no_data=[1], no_data=[-9999.0])
    e = np.round(base); e[3:9, 5:20, 0] = -32768
    add("f32_d3_allint_nodata", e.astype(np.float32), 0.01, n_depth=3, uses_no_data=[1], no_data=[-32768.0])
    add("f32_d3_allint_nodata_3.2", e.astype(np.float32), 3.2, n_depth=3, uses_no_data=[1], no_data=[-32768.0])
    f = base[:, :, 0].copy(); f[5:15, 5:25] = -1
    add("f32_d1_nodata", f, 0.01, uses_no_data=[1], no_data=[-1.0])
    g = (rng.random((h, w, 3)) * 100 + 50).astype(np.float64); g[1:4, 1:30, 2] = -1e300
    add("f64_d3_nodata_huge", g, 0.001, n_depth=3, uses_no_data=[1], no_data=[-1e300])
    i16 = (rng.integers(100, 2000, (h, w, 3))).astype(np.int16); i16[3:9, 5:20, 1] = -32768; i16[20:24, 10:14, :] = -32768
    add("i16_d3_nodata", i16, 0, n_depth=3, uses_no_data=[1], no_data=[-32768.0])
    add("i16_d3_nodata_lossy3", i16, 3, n_depth=3, uses_no_data=[1], no_data=[-32768.0])
    u8 = (rng.integers(1, 255, (h, w, 3))).astype(np.uint8); u8[3:9, 5:20, 1] = 0
    add("u8_d3_nodata_0_adjacent", u8, 0, n_depth=3, uses_no_data=[1], no_data=[0.0])
    u8b = (rng.integers(0, 200, (h, w, 3))).astype(np.uint8); u8b[3:9, 5:20, 1] = 255
    add("u8_d3_nodata_255_min0", u8b, 0, n_depth=3, uses_no_data=[1], no_data=[255.0])
    i32 = (rng.integers(-1000, 1000, (h, w, 2))).astype(np.int32); i32[0:5, 0:9, 0] = 2147483647
    add("i32_d2_nodata_max", i32, 2, n_depth=2, uses_no_data=[1], no_data=[2147483647.0])
    bands = np.stack([a, base, d])
    add("f32_d3_3bands_mixed", bands, 0.01, n_depth=3, n_bands=3, uses_no_data=[1, 0, 1], no_data=[-9999.0, 0.0, -9999.0])
    m = np.ones((h, w), np.uint8); m[0:10, 0:10] = 0
    add("f32_d3_nodata_masked", a, 0.01, n_depth=3, mask=m, uses_no_data=[1], no_data=[-9999.0])
    add("f32_d3_flag_without_values", base, 0.01, n_depth=3, uses_no_data=[1], no_data=None)          # WrongParam
    add("u8_d3_nodata_out_of_range", u8, 0, n_depth=3, uses_no_data=[1], no_data=[300.0])              # WrongParam
    add("f32_d3_all_nodata", np.full((h, w, 3), -5.0, np.float32), 0.01, n_depth=3, uses_no_data=[1], no_data=[-5.0])
    return cases


def fpl_cases(seed=21):
    """(name, array, kwargs) of float rasters for maxZError = 0: the reference codes these with its lossless FPL codec
    (IEM_DeltaDeltaHuffman, fpl_*.cpp) when that is >= 10 % smaller than raw blocks.  Used for DECODE parity: the blobs the
    reference makes from them are committed in tests/golden/fpl_ref.npz."""
    rng = np.random.default_rng(seed)
    f = c2_raster(97, 131)
    cases = [("f32_noisy", f, {}), ("f64_noisy", f.astype(np.float64) * 1.0000001, {}),
             ("f32_smooth", smooth_field(80, 133).astype(np.float32), {}),
             ("f32_const_rows", np.repeat(rng.random((60, 1)).astype(np.float32), 90, 1), {}),
             ("f32_const_cols", np.repeat(rng.random((1, 90)).astype(np.float32), 60, 0), {})]
    m = np.ones((97, 131), np.uint8)
    m[20:60, 30:100] = 0
    cases.append(("f32_masked", f, {"mask": m}))
    fn = f.copy()
    fn[10:20, 10:50] = np.nan
    cases.append(("f32_nan", fn, {}))
    cases.append(("f32_depth3", (c4_raster(40, 50).astype(np.float32) * np.float32(1.37)), {"n_depth": 3}))
    cases.append(("f64_depth2", rng.random((30, 40, 2)), {"n_depth": 2}))
    cases.append(("f32_3bands", np.stack([f, f * 2, f + 5]), {"n_bands": 3}))
    cases.append(("f32_eighths", (np.round(f * 8) / 8).astype(np.float32), {}))
    cases.append(("f32_huge_noise", (rng.random((64, 70)) * 1e30).astype(np.float32), {}))
    cases.append(("f32_one_row", f[:1, :].copy(), {}))
    cases.append(("f32_one_col", f[:, :1].copy(), {}))
    cases.append(("f32_big", c2_raster(300, 517), {}))
    return cases


def fpl_encode_cases(seed=23):
    """more float rasters for maxZError = 0, ENCODE parity with the oracle (itself pinned to the reference's blobs on fpl_cases):
    every predictor (none / row / row + column), byte-delta levels > 0, one-value, PackBits, stored and Huffman planes, inputs
    shorter than one 8 KB sample, nDepth > 1, long runs that cross PackBits' 129-byte chunks"""
    rng = np.random.default_rng(seed)
    cases = [
        ("f32_sparse", np.where(rng.random((300, 400)) < 0.01, 1.5, 0).astype(np.float32), {}),
        ("f32_ramp", np.arange(500 * 600, dtype=np.float32).reshape(500, 600) * np.float32(0.37), {}),
        ("f64_ramp", np.arange(300 * 200, dtype=np.float64).reshape(300, 200) * 0.001, {}),
        ("f32_sin", np.sin(np.arange(512 * 512, dtype=np.float32).reshape(512, 512) / 97).astype(np.float32), {}),
        ("f32_depth4", rng.random((100, 120, 4)).astype(np.float32) * np.float32(3), {"n_depth": 4}),
        ("f32_tiny", rng.random((5, 7)).astype(np.float32), {}),
        ("f32_8191", rng.random((1, 8191)).astype(np.float32), {}),
        ("f32_8192", rng.random((1, 8192)).astype(np.float32), {}),
        ("f32_thin", rng.random((9000, 3)).astype(np.float32), {}),
        ("f32_quarters", (np.round(smooth_field(400, 400)) / 4).astype(np.float32), {}),
        ("f32_long_runs", np.repeat(np.repeat(rng.integers(0, 3, (4, 5)).astype(np.float32), 100, 0), 200, 1), {}),
        ("f64_smooth", smooth_field(300, 500) * 1.000001, {}),
        ("f32_1024x2048", c2_raster(1024, 2048), {}),         # 16 test blocks / snippets, 2 M values per plane
    ]
    m = np.ones((300, 400), np.uint8)
    m[50:200, 100:300] = 0
    nan = c2_raster(300, 400)
    nan[100:120, 50:350] = np.nan                              # NaN at valid and at masked pixels
    cases.append(("f32_nan_and_mask", nan, {"mask": m}))
    return cases


def fpl_fuzz_cases(seed=29, n=40):
    """seeded random float rasters for maxZError = 0: degenerate shapes (one row / column, fewer than 7 values, wider than one
    8 KB test block, exactly 8192 + 1 values), nDepth 1..5, masks, NaNs, value patterns that hit every plane coding"""
    rng = np.random.default_rng(seed)
    shapes = [(1, 2), (2, 1), (3, 3), (1, 9000), (9000, 1), (2, 5000), (4, 20000), (7, 7), (64, 129), (130, 64), (8, 1024), (1024, 8),
              (1, 8193), (3, 2731), (2731, 3), (5, 1639), (90, 91), (17, 4000)]
    cases = []
    for it in range(n):
        h, w = shapes[rng.integers(len(shapes))]
        dt = np.float32 if rng.random() < 0.7 else np.float64
        nd = 1 if rng.random() < 0.7 else int(rng.integers(2, 6))
        if nd > 1 and h * w > 20000:
            h, w = min(h, 50), min(w, 60)
        kind = int(rng.integers(6))
        shp = (h, w) if nd == 1 else (h, w, nd)
        if kind == 0:
            a = rng.random(shp)
        elif kind == 1:
            a = np.cumsum(rng.normal(0, 1, shp), axis=1)
        elif kind == 2:
            a = np.where(rng.random(shp) < 0.02, rng.random(shp), 0)
        elif kind == 3:
            a = np.round(rng.random(shp) * 4) / 4 + 1
        elif kind == 4:
            a = (np.arange(int(np.prod(shp))).reshape(shp) % int(rng.integers(2, 300))) * 0.5
        else:
            a = rng.integers(0, 2, shp) * 1e20 + rng.random(shp)
        a = a.astype(dt)
        kw = {}
        if nd > 1:
            kw["n_depth"] = nd
        if rng.random() < 0.25:
            kw["mask"] = (rng.random((h, w)) < 0.8).astype(np.uint8)
        if rng.random() < 0.2 and nd == 1:
            a[rng.random((h, w)) < 0.05] = np.nan
        cases.append((f"{it:02d}_{dt.__name__}_{h}x{w}x{nd}_k{kind}", a, kw))
    return cases


def bitplane_cases(seed=31):
    """(name, array, kwargs) for maxZErr == 777, the reference's "cheat code" for the integer bit-plane mode (Lerc2.cpp:210-217,
    :1071-1229): low bit planes that look like noise are dropped by raising maxZError to half of the last plane kept."""
    rng = np.random.default_rng(seed)
    h, w = 120, 160
    smooth = smooth_field(h, w)
    cases = []
    noisy_low = (smooth.astype(np.int64) * 64 + rng.integers(0, 64, (h, w))).astype(np.uint16)        # 6 noise planes under a smooth signal
    cases.append(("u16_six_noise_planes", noisy_low, {}))
    cases.append(("i16_six_noise_planes", (noisy_low.astype(np.int32) - 20000).astype(np.int16), {}))
    cases.append(("u8_smooth_no_noise", np.clip(smooth / 8, 0, 255).astype(np.uint8), {}))
    cases.append(("u8_all_noise", rng.integers(0, 256, (h, w)).astype(np.uint8), {}))
    cases.append(("i32_ten_noise_planes", (smooth.astype(np.int64) * 1024 + rng.integers(0, 1024, (h, w)) - 500000).astype(np.int32), {}))
    cases.append(("u32_three_noise_planes", (smooth.astype(np.int64) * 8 + rng.integers(0, 8, (h, w))).astype(np.uint32), {}))
    m = np.ones((h, w), np.uint8)
    m[10:60, 20:100] = 0
    cases.append(("u16_masked", noisy_low, {"mask": m}))
    d3 = np.stack([noisy_low, noisy_low // 2, (noisy_low * 3) & 0xffff], axis=-1).astype(np.uint16)
    cases.append(("u16_depth3", d3, {"n_depth": 3}))
    cases.append(("u16_too_small", noisy_low[:50, :60].copy(), {}))                                    # fewer than 5000 pixels: lossless
    cases.append(("f32_is_refused", smooth.astype(np.float32), {}))                                    # float types: Failed
    cases.append(("u16_2bands", np.stack([noisy_low, noisy_low[::-1].copy()]), {"n_bands": 2}))
    return cases
