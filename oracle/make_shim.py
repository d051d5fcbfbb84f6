#!/usr/bin/env python
"""oracle/make_shim.py -- compile the MINIMAL shim of INTEGRATION.md: the reference's own sources with ONE function body replaced.

The reference tree is copied to a scratch directory (nothing of it enters the repository), the body of
`Network::ContractIndices` (/root/reference/src/Network.h:876-971) is cut out by brace matching and oracle/shim/contract_indices_body.inc
-- a call sequence on the C ABI of include/qtorch_b200.h -- is put in its place, `#include "qtorch_b200.h"` is added, and the reference's
UNMODIFIED main.cpp is compiled against the result into oracle/_ref/qtorch_shim (links libqtorch_b200.so).  tests/test_gpu_cli.py runs it
next to the unmodified reference binary.  Test infrastructure; only possible where the reference tree exists."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")


def patched_network_h(text, body):
    sig = "inline void Network::ContractIndices("
    at = text.index(sig)
    open_brace = text.index("{", text.index(")", text.index("nodeC", at)))
    depth, i = 0, open_brace
    while True:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    out = text[:open_brace + 1] + "\n" + body + "\n    " + text[i:]
    first_include = out.index("#include")
    return out[:first_include] + '#include "qtorch_b200.h"\n' + out[first_include:]


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print("reference tree %s absent: keeping prebuilt oracle/_ref/qtorch_shim" % REF)
        return 0
    body = open(os.path.join(HERE, "shim", "contract_indices_body.inc")).read()
    out = os.path.join(HERE, "_ref", "qtorch_shim")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="qtb_shim_") as d:
        shutil.copytree(os.path.join(REF, "src"), os.path.join(d, "src"))
        path = os.path.join(d, "src", "Network.h")
        os.chmod(path, 0o644)
        text = open(path).read()
        open(path, "w").write(patched_network_h(text, body))
        lib = os.path.join(ROOT, "qtorch_b200")
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++11", "-pthread", "-w", "-I" + d, "-I" + os.path.join(ROOT, "include"), "-o", out,
               os.path.join(d, "src", "main.cpp"), "-L" + lib, "-lqtorch_b200", "-Wl,-rpath,$ORIGIN/../../qtorch_b200"]
        subprocess.run(cmd, check=True)
    print("built", out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
