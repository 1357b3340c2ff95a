"""Wall-clock probe of the 8-bit Huffman path (BASELINE configs[3] shape): uint8, nDepth = 3, lossless."""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import product_lib, oracle_lib
from cases import c4_raster
import lerc_b200
prod = product_lib()
enc, dec = prod.f["encode"], prod.f["decode"]
for n in (1024, 4096, 8192):
    img = c4_raster(n, n)
    d_img = torch.from_numpy(img).cuda()
    cap = img.nbytes + 65536
    d_blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_img)
    nb = C.c_uint(0)
    te = td = 1e9
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = enc(d_img.data_ptr(), 1, 3, n, n, 1, 0, None, 0.0, d_blob.data_ptr(), cap, C.addressof(nb))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        s0 = lerc_b200.stats()
        if rep == 1: lerc_b200.kernel_times(); lerc_b200.profile(True)
        st2 = dec(d_blob.data_ptr(), nb.value, 0, None, 3, n, n, 1, 1, d_out.data_ptr())
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if rep == 1: lerc_b200.profile(False)
        s1 = lerc_b200.stats()
        te, td = min(te, t1 - t0), min(td, t2 - t1)
    ok = bool((d_out == d_img).all().item())
    print(f"u8x3 {n}^2 st {st}/{st2} blob {nb.value/1e6:.2f} MB ({img.nbytes/nb.value:.2f}x)  enc {te*1e3:.3f} ms  dec {td*1e3:.3f} ms  fastdec {s1[4]-s0[4]}  lossless {ok}")
    print("   ", {k: (v[0], round(v[1], 3)) for k, v in sorted(lerc_b200.kernel_times().items(), key=lambda kv: -kv[1][1])[:8]})
    if n == 1024:
        orc = oracle_lib()
        blob = bytes(d_blob[: nb.value].cpu().numpy().tobytes())
        s_o, b_o, _ = orc.encode(img, 0, n_depth=3)
        print("    blob equals oracle:", blob == b_o)
