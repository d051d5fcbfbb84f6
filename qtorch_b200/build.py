"""Build every native artefact of qtorch_b200 in-tree (so it travels to the GPU box with the snapshot).

  qtorch_b200/libqtorch_b200.so   CUDA engine + C ABI (include/qtorch_b200.h), sm_100a only
  qtorch_b200/bin/qtorch          drop-in `qtorch <script.inp>` front-end (host mirror, C++14)
  qtorch_b200/bin/qtb_harness     full-precision driver used by the parity tests

`python -m qtorch_b200.build` or `qtorch_b200.build.build_all()`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libqtorch_b200.so")
BIN = os.path.join(HERE, "bin")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(sub, exts):
    d = os.path.join(HERE, sub)
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(exts)]


def build_engine(force=False, verbose=False):
    srcs = _sources("csrc", (".cu", ".cuh", ".h", ".inl")) + [os.path.join(ROOT, "include", "qtorch_b200.h")]
    if not force and not _newer(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(HERE, "csrc", "engine.cu"), "-ldl"]
    subprocess.run(cmd, check=True)
    return LIB


def build_host(force=False):
    os.makedirs(BIN, exist_ok=True)
    hdrs = _sources("host", (".h", ".hpp")) + [os.path.join(ROOT, "include", "qtorch_b200.h"), LIB]
    out = []
    for name, src in (("qtb_harness", "qtb_harness.cpp"), ("qtorch", "qtorch_main.cpp"), ("maxcutQAOA", "maxcut_main.cpp")):
        target = os.path.join(BIN, name)
        source = os.path.join(HERE, "apps", src)
        if force or _newer(target, hdrs + [source]):
            cmd = ["g++", "-std=c++14", "-O2", "-o", target, source, "-L" + HERE, "-lqtorch_b200",
                   "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
            subprocess.run(cmd, check=True)
        out.append(target)
    target = os.path.join(HERE, "libqtorch_host.so")
    source = os.path.join(HERE, "apps", "host_capi.cpp")
    if force or _newer(target, hdrs + [source]):
        subprocess.run(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-o", target, source, "-L" + HERE, "-lqtorch_b200",
                        "-Wl,-rpath,$ORIGIN", "-lpthread"], check=True)
    out.append(target)
    return out


def build_all(force=False):
    build_engine(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB)
