"""development aid: a california-shaped masked 4096 x 4096 float32 raster (one third invalid, ragged coast) through lerc_encode /
lerc_decode with device pointers: wall time per call and the kernel table."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import lerc_b200
from cases import c2_raster
from lercapi import product_lib, ref_lib, oracle_lib
prod = product_lib()
H = W = 4096
img = c2_raster(H, W)
yy, xx = np.mgrid[0:H, 0:W]
coast = 1400 + 500 * np.sin(yy / 300.0) + 120 * np.sin(yy / 37.0) + 25 * np.sin(yy / 5.0)
mask = (xx > coast).astype(np.uint8)
chk = ref_lib() or oracle_lib()
s_r, b_r, _ = chk.encode(img, 0.01, mask=mask)
d_img, d_mask = torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda()
d_out = torch.zeros(len(b_r) + 4096, dtype=torch.uint8, device="cuda")
d_dec = torch.empty_like(d_img); d_dm = torch.empty_like(d_mask)
n = C.c_uint(0)
enc, dec = prod.f["encode"], prod.f["decode"]
for it in range(4):
    if it == 3: lerc_b200.profile(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = enc(d_img.data_ptr(), 6, 1, W, H, 1, 1, d_mask.data_ptr(), 0.01, d_out.data_ptr(), d_out.numel(), C.addressof(n))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    st2 = dec(d_out.data_ptr(), n.value, 1, d_dm.data_ptr(), 1, W, H, 1, 6, d_dec.data_ptr())
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"iter {it}: encode {1e3*(t1-t0):.3f} ms  decode {1e3*(t2-t1):.3f} ms  status {st} {st2}  bytes {n.value} (reference {len(b_r)})")
kt = lerc_b200.kernel_times()
for k, (c, ms) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {k:44s} x{c}  {ms:.4f} ms")
print("blob equals the reference's:", d_out[: n.value].cpu().numpy().tobytes() == b_r, " mask round trip:", bool((d_dm == d_mask).all()))
