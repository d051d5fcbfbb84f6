"""Build every native artefact of qtorch_b200 in-tree (so it travels to the GPU box with the snapshot).

  qtorch_b200/libqtorch_b200.so   CUDA engine + C ABI (include/qtorch_b200.h), sm_100a only
  qtorch_b200/bin/qtorch          drop-in `qtorch <script.inp>` front-end (host mirror, C++14)
  qtorch_b200/bin/qtb_harness     full-precision driver used by the parity tests
  qtorch_b200/bin/maxcutQAOA      drop-in `maxcutQAOA <graph> <p> <mode> ...` front-end; links NLopt's COBYLA (bin/_nlopt, built from the
                                  tree the reference vendors) when that tree is present at build time

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


NLOPT_SRC = os.environ.get("QTB_NLOPT_SRC", "/root/reference/nlopt-2.4.2")
NLOPT = os.path.join(BIN, "_nlopt")


def build_nlopt():
    """NLopt 2.4.2 -- the third-party optimiser the reference vendors for maxcutQAOA (src/maxcut.cpp:211-213, LN_COBYLA) -- compiled
    OUT OF TREE from the sources where they lie (nothing of it enters the repository) into bin/_nlopt (static library + headers;
    git-ignored, travels to the GPU box).  Returns the install prefix, or None when the tree is absent and nothing was built before:
    maxcutQAOA then falls back to its Nelder-Mead ascent."""
    lib = os.path.join(NLOPT, "lib", "libnlopt.a")
    if os.path.exists(lib) and os.path.exists(os.path.join(NLOPT, "include", "nlopt.hpp")):
        return NLOPT
    if not os.path.exists(os.path.join(NLOPT_SRC, "configure")):
        return None
    import tempfile
    with tempfile.TemporaryDirectory(prefix="qtb_nlopt_") as d:
        try:
            for cmd in ([os.path.join(NLOPT_SRC, "configure"), "--prefix=" + NLOPT, "--disable-shared", "--without-octave", "--without-python",
                         "--without-guile", "--without-matlab"], ["make", "-j8"], ["make", "install"]):
                subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            return None
    return NLOPT if os.path.exists(lib) else None


def build_host(force=False):
    os.makedirs(BIN, exist_ok=True)
    nlopt = build_nlopt()
    hdrs = _sources("host", (".h", ".hpp")) + [os.path.join(ROOT, "include", "qtorch_b200.h"), LIB]
    out = []
    for name, src in (("qtb_harness", "qtb_harness.cpp"), ("qtorch", "qtorch_main.cpp"), ("maxcutQAOA", "maxcut_main.cpp")):
        target = os.path.join(BIN, name)
        source = os.path.join(HERE, "apps", src)
        if force or _newer(target, hdrs + [source]):
            cmd = ["g++", "-std=c++14", "-O2", "-o", target, source, "-L" + HERE, "-lqtorch_b200",
                   "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
            if name == "maxcutQAOA" and nlopt:
                cmd += ["-DQTB_HAVE_NLOPT", "-I" + os.path.join(nlopt, "include"), os.path.join(nlopt, "lib", "libnlopt.a"), "-lm"]
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
