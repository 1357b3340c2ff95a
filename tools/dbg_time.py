import sys, os, time, ctypes as C, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib
from cases import c2_raster
import lerc_b200
prod = product_lib()
enc, dec = prod.f["encode"], prod.f["decode"]
imgs = [torch.from_numpy(c2_raster(4096, 4096, seed=1234 + i)).cuda() for i in range(4)]
cap = 27000000
blobs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(4)]
outs = [torch.empty_like(imgs[0]) for _ in range(4)]
n = C.c_uint(0)
def step(k):
    assert enc(imgs[k].data_ptr(), 6, 1, 4096, 4096, 1, 0, None, 0.01, blobs[k].data_ptr(), cap, C.addressof(n)) == 0
    assert dec(blobs[k].data_ptr(), n.value, 0, None, 1, 4096, 4096, 1, 6, outs[k].data_ptr()) == 0
for i in range(4): step(i)
torch.cuda.synchronize()
lerc_b200.profile(True)
for i in range(12): step(i % 4)
torch.cuda.synchronize()
lerc_b200.profile(False)
kt = lerc_b200.kernel_times()
print(os.environ.get("LERC_B200_ENC_OCC"), {k: round(v[1] / v[0] * 1e3, 1) for k, v in kt.items() if "encode" in k or "dec_blocks" in k})
