import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib
import lerc_b200
prod = product_lib()
enc, dec = prod.f["encode"], prod.f["decode"]
n = 16384
torch.manual_seed(5)
xx = torch.arange(n, device="cuda", dtype=torch.float32)
img = (1000 + 300 * torch.sin(xx[None, :] / 97) * torch.cos(xx[:, None] / 131) + 50 * torch.sin(xx[None, :] / 13 + xx[:, None] / 17)
       + 0.5 * torch.randn(n, n, device="cuda")).contiguous()
cap = n * n * 4 + (1 << 20)
blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
nb = C.c_uint(0)
s0 = lerc_b200.stats()
st = enc(img.data_ptr(), 6, 1, n, n, 1, 0, None, 0.001, blob.data_ptr(), cap, C.addressof(nb))
print("enc", st, nb.value, "fastenc", lerc_b200.stats()[3] - s0[3])
for rep in range(3):
    out = torch.full_like(img, -7.0)
    s0 = lerc_b200.stats()
    st = dec(blob.data_ptr(), nb.value, 0, None, 1, n, n, 1, 6, out.data_ptr())
    torch.cuda.synchronize()
    err = (out.double() - img.double()).abs()
    bad = (err > 0.0011)
    nbad = int(bad.sum().item())
    print("dec", st, "fastdec", lerc_b200.stats()[4] - s0[4], "maxerr", float(err.max().item()), "bad px", nbad, "untouched", int((out == -7.0).sum().item()))
    if nbad:
        idx = bad.nonzero()
        print("  first bad", idx[0].tolist(), "last bad", idx[-1].tolist(), "rows with bad:", int(bad.any(dim=1).sum().item()))
