"""Generates the golden fixtures of tests/golden/ from the UNMODIFIED reference (run in the build
container, where /root/reference exists and oracle/_ref/libLerc_ref.so has been built by
`make -C oracle ref`).  Outputs (all small, committed):

  <name>.lerc2          the reference's own shipped test blobs, byte for byte (testData/*.lerc2), and the
                        inline golden blob of OtherLanguages/js/tests/sanity.mjs:6
  <name>.npz            what the reference decodes from them: pixels, mask, blob info, data ranges
  synthetic_ref.npz     for a few seeded synthetic rasters (tests/cases.py): SHA-256 of the blob the
                        reference encodes and of the pixels it decodes, so the oracle stays pinned to the
                        reference on machines where /root/reference is absent (the GPU box)
"""
import hashlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from lercapi import ref_lib  # noqa: E402
from cases import all_cases, bitplane_cases, fpl_cases, nodata_cases  # noqa: E402

REF = "/root/reference"


def main():
    ref = ref_lib()
    assert ref is not None, "build oracle/_ref first: make -C oracle ref"
    blobs = {}
    for f in ["california_400_400_1_float.lerc2", "bluemarble_256_256_3_byte.lerc2"]:
        blobs[f[:-6]] = open(os.path.join(REF, "testData", f), "rb").read()
    js = open(os.path.join(REF, "OtherLanguages/js/tests/sanity.mjs")).read()
    m = re.search(r'const data4D =\s*"([0-9,]+)"', js)
    blobs["js_sanity_30_20_3_byte"] = bytes(int(v) for v in m.group(1).split(","))
    for name, blob in blobs.items():
        open(os.path.join(HERE, name + ".lerc2"), "wb").write(blob)
        st, info = ref.blob_info(blob)
        assert st == 0, (name, st)
        st, data, mask = ref.decode(blob)
        assert st == 0, (name, st)
        st, mins, maxs = ref.data_ranges(blob, info["nDepth"], info["nBands"])
        assert st == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=data, mask=mask if mask is not None else np.zeros(0, np.uint8),
                            info=np.array([info[k] for k in sorted(info)], dtype=np.float64), info_keys=np.array(sorted(info)),
                            mins=mins, maxs=maxs)
        print(name, info)
    # synthetic: hashes only
    names, enc, dec, sizes = [], [], [], []
    for name, arr, mz, kw in all_cases():
        st, blob, _ = ref.encode(arr, mz, **kw)
        if st != 0:
            continue
        st, data, mask = ref.decode(blob)
        assert st == 0
        names.append(name)
        sizes.append(len(blob))
        enc.append(hashlib.sha256(blob).hexdigest())
        h = hashlib.sha256(data.tobytes())
        if mask is not None:
            h.update(mask.tobytes())
        dec.append(h.hexdigest())
    np.savez_compressed(os.path.join(HERE, "synthetic_ref.npz"), names=np.array(names), sizes=np.array(sizes), enc=np.array(enc), dec=np.array(dec))
    print("synthetic cases:", len(names))
    # the same rasters through lerc_encodeForVersion(2..5) (Lerc::EncodeInternal_v5, Lerc.cpp:526-624): status, blob and pixel hashes
    keys, status, enc, dec, sizes = [], [], [], [], []
    for v in (2, 3, 4, 5):
        for name, arr, mz, kw in all_cases():
            st, blob, _ = ref.encode(arr, mz, version=v, **kw)
            keys.append(f"{name}|{v}")
            status.append(st)
            if st != 0:
                sizes.append(0); enc.append(""); dec.append("")
                continue
            st2, data, mask = ref.decode(blob)
            assert st2 == 0
            sizes.append(len(blob))
            enc.append(hashlib.sha256(blob).hexdigest())
            h = hashlib.sha256(data.tobytes())
            if mask is not None:
                h.update(mask.tobytes())
            dec.append(h.hexdigest())
    np.savez_compressed(os.path.join(HERE, "synthetic_ref_versions.npz"), keys=np.array(keys), status=np.array(status), sizes=np.array(sizes),
                        enc=np.array(enc), dec=np.array(dec))
    print("old-version cases:", len(keys))
    # the _4D calls with per-band noData values (Lerc.cpp:1241-1374, :1378-1618, :1046-1076)
    names, status, enc, dec, sizes, uses, vals = [], [], [], [], [], [], []
    for name, arr, mz, kw in nodata_cases():
        st, blob = ref.encode_4d(arr, mz, **kw)
        names.append(name); status.append(st)
        if st != 0:
            sizes.append(0); enc.append(""); dec.append(""); uses.append(""); vals.append("")
            continue
        st2, data, mask, u, v = ref.decode_4d(blob)
        assert st2 == 0
        sizes.append(len(blob)); enc.append(hashlib.sha256(blob).hexdigest())
        h = hashlib.sha256(data.tobytes())
        if mask is not None:
            h.update(mask.tobytes())
        dec.append(h.hexdigest()); uses.append(u.tobytes().hex()); vals.append(v.tobytes().hex())
    np.savez_compressed(os.path.join(HERE, "nodata_ref.npz"), names=np.array(names), status=np.array(status), sizes=np.array(sizes),
                        enc=np.array(enc), dec=np.array(dec), uses=np.array(uses), vals=np.array(vals))
    print("noData cases:", len(names))
    # lossless float: the blobs the reference makes with its FPL codec (decode parity; the blobs themselves are the fixtures)
    out = {}
    for name, arr, kw in fpl_cases():
        st, blob, _ = ref.encode(arr, 0, **kw)
        assert st == 0
        st, data, mask = ref.decode(blob)
        assert st == 0
        h = hashlib.sha256(data.tobytes())
        if mask is not None:
            h.update(mask.tobytes())
        out["blob_" + name] = np.frombuffer(blob, np.uint8)
        out["hash_" + name] = np.array(h.hexdigest())
    np.savez_compressed(os.path.join(HERE, "fpl_ref.npz"), **out)
    print("FPL cases:", len(out) // 2)
    # maxZErr == 777: the integer bit-plane mode (Lerc2.cpp:210-217, :1071-1229)
    names, status, enc, sizes, mz = [], [], [], [], []
    for name, arr, kw in bitplane_cases():
        st, blob, _ = ref.encode(arr, 777, **kw)
        names.append(name); status.append(st); sizes.append(len(blob)); enc.append(hashlib.sha256(blob).hexdigest() if st == 0 else "")
        mz.append(ref.blob_info(blob)[1]["maxZErrUsed"] if st == 0 else -1.0)
    np.savez_compressed(os.path.join(HERE, "bitplane_ref.npz"), names=np.array(names), status=np.array(status), sizes=np.array(sizes),
                        enc=np.array(enc), maxzerr=np.array(mz))
    print("bit-plane cases:", len(names))


if __name__ == "__main__":
    main()
