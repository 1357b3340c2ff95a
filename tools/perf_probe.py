"""Wall-clock probe of lerc_encode / lerc_decode (device pointers) on several data kinds; prints ms and whether the
fused paths were taken.  Not a test; used to find slow paths."""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib, DT_CODE
from cases import c2_raster, smooth_field
import lerc_b200
prod = product_lib()
enc, dec = prod.f["encode"], prod.f["decode"]

def probe(name, arr, mz, mask=None, reps=3):
    h, w = arr.shape
    dt = DT_CODE[arr.dtype]
    d_img = torch.from_numpy(arr).cuda()
    d_mask = torch.from_numpy(mask).cuda() if mask is not None else None
    cap = arr.nbytes + arr.nbytes // 4 + 65536
    d_blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_img)
    d_mout = torch.empty((h, w), dtype=torch.uint8, device="cuda") if mask is not None else None
    n = C.c_uint(0)
    te = td = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = enc(d_img.data_ptr(), dt, 1, w, h, 1, 1 if mask is not None else 0, d_mask.data_ptr() if mask is not None else None, mz,
                 d_blob.data_ptr(), cap, C.addressof(n))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        s0 = lerc_b200.stats()
        st2 = dec(d_blob.data_ptr(), n.value, 1 if mask is not None else 0, d_mout.data_ptr() if mask is not None else None, 1, w, h, 1, dt, d_out.data_ptr())
        torch.cuda.synchronize(); t2 = time.perf_counter()
        s1 = lerc_b200.stats()
        te, td = min(te, t1 - t0), min(td, t2 - t1)
    err = float((d_out.double() - d_img.double()).abs().max().item()) if mask is None else float(((d_out.double() - d_img.double()).abs() * d_mask).max().item())
    print(f"{name:28s} st {st}/{st2} blob {n.value/1e6:8.2f} MB ratio {arr.nbytes/max(n.value,1):5.2f}  enc {te*1e3:8.3f} ms  dec {td*1e3:8.3f} ms  fastdec {s1[4]-s0[4]}  maxerr {err:.4g}")

rng = np.random.default_rng(1)
f = c2_raster(4096, 4096)
probe("f32 4096^2 mz0.01", f, 0.01)
probe("f32 4096^2 mz1.0", f, 1.0)
probe("f32 4096^2 mz0.0001", f, 0.0001)
dem = np.clip(smooth_field(4096, 4096) * 3 + rng.normal(0, 2, (4096, 4096)), -32768, 32767).astype(np.int16)
probe("i16 4096^2 lossless", dem, 0)
probe("u16 4096^2 lossy 4", (dem.astype(np.int32) + 2000).astype(np.uint16), 4)
probe("i32 4096^2 lossless", dem.astype(np.int32) * 1000, 0)
sea = f.copy(); sea[:, :1500] = 0
probe("f32 4096^2 40% zero blocks", sea, 0.01)
m = np.ones((4096, 4096), np.uint8); m[1000:2000, 500:3000] = 0
probe("f32 4096^2 masked rect", f, 0.01, mask=m)
lerc_b200.kernel_times()
lerc_b200.profile(True)
probe("f32 4096^2 masked rect (prof)", f, 0.01, mask=m, reps=1)
lerc_b200.profile(False)
print({k: (v[0], round(v[1], 3)) for k, v in sorted(lerc_b200.kernel_times().items(), key=lambda kv: -kv[1][1])})
mr = (rng.random((4096, 4096)) > 0.2).astype(np.uint8)
probe("f32 4096^2 masked random 20%", f, 0.01, mask=mr)
big = c2_raster(16384, 16384)
probe("f32 16384^2 mz0.001", big, 0.001, reps=3)
lerc_b200.profile(True)
probe("f32 16384^2 mz0.001 (prof)", big, 0.001, reps=1)
lerc_b200.profile(False)
print({k: (v[0], round(v[1], 3)) for k, v in lerc_b200.kernel_times().items()})
