#!/usr/bin/env python3
"""tools/cusim/prep.py -- DEVELOPMENT TOOL, NOT PRODUCT.

Copies the product sources into tools/cusim/_build/tree/ with the two CUDA-only syntaxes rewritten so that g++ can
compile them against the shim in tools/cusim/cuda_runtime.h:
  kernel<<<grid, block, smem, stream>>>(args);   ->  cusim::launch(grid, block, smem, stream, [&]() { kernel(args); });
  extern __shared__ [align] T name[];            ->  T* name = (T*)cusim::dynSmem();
The product sources themselves are not modified."""
import os, re, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tools", "cusim", os.environ.get("CUSIM_BUILD_DIR", "_build"), "tree")

LAUNCH = re.compile(r"([A-Za-z_][\w:<>, ]*?)<<<(.+?)>>>\((.*?)\);")
DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:]+(?:\s+[\w:]+)*?)\s+(\w+)\[\];")


def rewrite(text):
    text = LAUNCH.sub(lambda m: "cusim::launch(%s, [&]() { %s(%s); });" % (m.group(2), m.group(1).strip(), m.group(3)), text)
    text = DYN.sub(lambda m: "%s* %s = (%s*)cusim::dynSmem();" % (m.group(1), m.group(2), m.group(1)), text)
    text = text.replace("__noinline__", "CUSIM_NOINLINE")
    assert "<<<" not in text, "unrewritten launch"
    assert not re.search(r"extern\s+__shared__", text), "unrewritten dynamic shared memory"
    return text


def main():
    for sub in ("lerc_b200/csrc", "include"):
        src, dst = os.path.join(ROOT, sub), os.path.join(OUT, sub)
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            if not name.endswith((".cu", ".cuh", ".cpp", ".h")):
                continue
            with open(os.path.join(src, name)) as f:
                text = rewrite(f.read())
            path = os.path.join(dst, name[:-3] + ".cpp" if name.endswith(".cu") else name)
            old = open(path).read() if os.path.exists(path) else None
            if old != text:
                with open(path, "w") as f:
                    f.write(text)


if __name__ == "__main__":
    main()
