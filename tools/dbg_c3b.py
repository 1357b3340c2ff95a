import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib, oracle_lib, ref_lib
from cases import c2_raster
import lerc_b200
prod = product_lib(); chk = ref_lib() or oracle_lib()
# --- what test_c2_full_size_bit_exact does
img = c2_raster(4096, 4096)
s_r, b_r, _ = chk.encode(img, 0.01)
s_p, b_p, _ = prod.encode(img, 0.01)
print("c2 enc equal", b_p == b_r)
st, nsz = prod.compute_size(img, 0.01)
st, d_p, _ = prod.decode(b_r)
print("c2 dec", st, lerc_b200.stats())
# --- what test_c3 does
enc, dec = prod.f["encode"], prod.f["decode"]
n = 16384
torch.manual_seed(5)
xx = torch.arange(n, device="cuda", dtype=torch.float32)
img = (1000 + 300 * torch.sin(xx[None, :] / 97) * torch.cos(xx[:, None] / 131) + 50 * torch.sin(xx[None, :] / 13 + xx[:, None] / 17)
       + 0.5 * torch.randn(n, n, device="cuda")).contiguous()
cap = n * n * 4 + (1 << 20)
blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
out = torch.full_like(img, -7.0)
nb = C.c_uint(0)
s0 = lerc_b200.stats()
st = enc(img.data_ptr(), 6, 1, n, n, 1, 0, None, 0.001, blob.data_ptr(), cap, C.addressof(nb))
s1 = lerc_b200.stats()
print("enc", st, nb.value, "fastenc", s1[3] - s0[3])
st = dec(blob.data_ptr(), nb.value, 0, None, 1, n, n, 1, 6, out.data_ptr())
torch.cuda.synchronize()
s2 = lerc_b200.stats()
err = (out.double() - img.double()).abs()
print("dec", st, "stats", s2, "fastdec", s2[4] - s1[4], "launches", s2[0] - s1[0], "maxerr", float(err.max().item()), "untouched", int((out == -7.0).sum().item()))
info = np.zeros(11, np.uint32)
print("blobinfo", prod.f["getBlobInfo"](blob.data_ptr(), nb.value, info.ctypes.data, None, 11, 0), info)
