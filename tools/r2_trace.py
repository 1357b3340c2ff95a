"""development aid: host-side timeline of one pipelined host-buffer encode + decode (LERC_B200_TRACE=1)"""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from cases import c2_raster
from lercapi import product_lib
prod = product_lib()
img = torch.from_numpy(c2_raster(4096, 4096)).pin_memory()
out = torch.empty(26_000_000, dtype=torch.uint8).pin_memory()
back = torch.empty_like(img).pin_memory()
n = C.c_uint(0)
enc, dec = prod.f["encode"], prod.f["decode"]
for it in range(4):
    if it == 3: sys.stderr.write("---- traced iteration\n")
    t0 = time.perf_counter()
    st = enc(img.data_ptr(), 6, 1, 4096, 4096, 1, 0, None, 0.01, out.data_ptr(), out.numel(), C.addressof(n))
    t1 = time.perf_counter()
    st2 = dec(out.data_ptr(), n.value, 0, None, 1, 4096, 4096, 1, 6, back.data_ptr())
    t2 = time.perf_counter()
    sys.stderr.write(f"iter {it}: encode {1e3*(t1-t0):.3f} ms  decode {1e3*(t2-t1):.3f} ms  status {st} {st2} bytes {n.value}\n")
# plain copies for comparison
d = torch.empty((4096, 4096), dtype=torch.float32, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(img, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    back.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
sys.stderr.write(f"plain 64 MB H2D {1e3*(t1-t0):.3f} ms ({64*1.048576/(1e3*(t1-t0)):.1f} GB/s)  D2H {1e3*(t2-t1):.3f} ms\n")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(img, non_blocking=True)
with torch.cuda.stream(s2): back.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
sys.stderr.write(f"both directions at once, 64 MB each: {1e3*(t1-t0):.3f} ms\n")
