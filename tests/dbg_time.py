import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib
from cases import c2_raster
import lerc_b200
prod = product_lib()
enc, dec = prod.f["encode"], prod.f["decode"]
img = c2_raster(4096, 4096)
d_img = torch.from_numpy(img).cuda()
cap = img.nbytes
d_blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_img)
n = C.c_uint(0)
def sync(): torch.cuda.synchronize()
for rep in range(3):
    sync(); t0 = time.perf_counter()
    st = enc(d_img.data_ptr(), 6, 1, 4096, 4096, 1, 0, None, 0.01, d_blob.data_ptr(), 26000000, C.addressof(n)); sync()
    t1 = time.perf_counter()
    st2 = dec(d_blob.data_ptr(), n.value, 0, None, 1, 4096, 4096, 1, 6, d_out.data_ptr()); sync()
    t2 = time.perf_counter()
    print("enc ms", (t1 - t0) * 1e3, "dec ms", (t2 - t1) * 1e3, st, st2, n.value, lerc_b200.stats())
lerc_b200.profile(True)
for rep in range(3):
    st2 = dec(d_blob.data_ptr(), n.value, 0, None, 1, 4096, 4096, 1, 6, d_out.data_ptr()); sync()
print(lerc_b200.kernel_times())
