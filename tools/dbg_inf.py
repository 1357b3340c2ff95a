import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from lercapi import oracle_lib, product_lib
from cases import smooth_field
import lerc_b200
prod, orc = product_lib(), oracle_lib()
h,w=128,256
rng=np.random.default_rng(5)
f32=(smooth_field(h,w)+rng.normal(0,0.5,(h,w))).astype(np.float32)
arr=np.where(f32>1290,np.float32(np.inf),f32).astype(np.float32)
so,bo,_=orc.encode(arr,0.01)
s0=lerc_b200.stats()
sp,bp,_=prod.encode(arr,0.01)
print("fast taken", lerc_b200.stats()[3]-s0[3], so, sp, len(bo), len(bp))
a=np.frombuffer(bo,np.uint8);b=np.frombuffer(bp,np.uint8)
n=min(len(a),len(b))
d=np.nonzero(a[:n]!=b[:n])[0]
print(len(d), d[:40])
for i in d[:8]: print(i, a[i], b[i])
