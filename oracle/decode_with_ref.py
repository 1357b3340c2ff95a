#!/usr/bin/env python3
"""oracle/decode_with_ref.py <blob file> <out.npz> -- TEST INFRASTRUCTURE: decodes a blob with the unmodified reference
(oracle/_ref/libLerc_ref.so) and stores pixels + mask.  Used by `make -C oracle ref` for testData/world.lerc1, a Lerc1 blob the
product does not read (ingested through the reference, then round-tripped through the product)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from lercapi import ref_lib  # noqa: E402

ref = ref_lib()
assert ref is not None, "oracle/_ref/libLerc_ref.so missing"
blob = open(sys.argv[1], "rb").read()
st, info = ref.blob_info(blob)
assert st == 0, st
st, px, mask = ref.decode(blob)
assert st == 0, st
np.savez_compressed(sys.argv[2], pixels=px, mask=mask if mask is not None else np.zeros(0, np.uint8), info=np.array([info[k] for k in sorted(info)], dtype=np.float64))
print("decoded", sys.argv[1], px.shape, px.dtype, "valid" if mask is None else int(mask.sum()))
