"""In-tree native build driver (nvcc / g++ invoked directly, no JIT cache).

Two shared objects are produced next to this file:

  libmxb200_host.so  -- C++17 runtime: dependency engine, storage pools, .params IO,
                        RecordIO + threaded batch loader, membership coordinator / van,
                        kvstore key registry + host reducers.  No CUDA needed to load.
  libmxb200_cuda.so  -- sm_100a kernels (tcgen05/TMA GEMM, fused BN/ReLU, softmax-CE,
                        fused optimizer, P2P/NVLS allreduce+update) and their launchers.

Both export a flat C ABI (see csrc/c_api.h) that python binds with ctypes -- the same
layering as the reference's include/mxnet/c_api.h + python/mxnet/base.py, which keeps
compile times in seconds (no torch headers in any translation unit).

A content hash of the sources is stored in ``<lib>.stamp`` so a library that travelled
to the GPU box is reused without rebuilding, and a stale one is rebuilt (or fails loudly).
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")

NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
# NOTE: deliberately NOT $CXX: this image exports CXX=/opt/gcc/bin/g++, a toolchain that links libstdc++ statically;
# a second libstdc++ inside a dlopen'ed library next to libtorch's crashes in iostream / std::thread code.
CXX = os.environ.get("MXB200_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")

CUDA_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v",
              "--use_fast_math", "-I", CSRC]
# hidden visibility: only the MXB_API entry points are exported
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-pthread", "-fvisibility=hidden",
              "-fvisibility-inlines-hidden", "-I", CSRC]


# MXB200_SANITIZE=address|undefined|thread: instrument the host runtime (the reference's USE_ASAN cmake switch,
# CMakeLists.txt:304-309).  The flags are part of the source hash, so switching rebuilds.  Load the result with
# LD_PRELOAD=$(gcc -print-file-name=libasan.so) python ...
_SAN = os.environ.get("MXB200_SANITIZE", "")
if _SAN:
    HOST_FLAGS = HOST_FLAGS + ["-fsanitize=" + _SAN, "-fno-omit-frame-pointer", "-g"]
HOST_LINK_FLAGS = (["-fsanitize=" + _SAN] if _SAN else [])


def _sources(kind):
    out = []
    for root, _d, files in os.walk(CSRC):
        for f in sorted(files):
            p = os.path.join(root, f)
            if kind == "cuda" and f.endswith(".cu"):
                out.append(p)
            if kind == "host" and f.endswith(".cc"):
                out.append(p)
    return out


def _headers():
    out = []
    for root, _d, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".h", ".cuh", ".hpp")):
                out.append(os.path.join(root, f))
    return out


def source_hash(kind):
    h = hashlib.sha256()
    for p in _sources(kind) + _headers():
        h.update(os.path.relpath(p, CSRC).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(CUDA_FLAGS if kind == "cuda" else HOST_FLAGS).encode())
    return h.hexdigest()


def lib_path(kind):
    return os.path.join(HERE, "libmxb200_%s.so" % kind)


def is_fresh(kind):
    lib, stamp = lib_path(kind), lib_path(kind) + ".stamp"
    if not (os.path.exists(lib) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == source_hash(kind)


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("build failed:\n$ %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _obj_for(src, kind):
    rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
    sub = kind + ("_" + _SAN.replace(",", "_") if (_SAN and kind == "host") else "")   # instrumented objects never mix in
    return os.path.join(OBJ, sub, rel + ".o")


def _obj_fresh(src, obj, hdr_mtime):
    return os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime)


def build(kind, force=False, verbose=False):
    """Compile every source of ``kind`` ('host' | 'cuda') and link the shared object."""
    if not force and is_fresh(kind):
        return lib_path(kind)
    srcs = _sources(kind)
    if not srcs:
        raise RuntimeError("no %s sources under %s" % (kind, CSRC))
    os.makedirs(os.path.dirname(_obj_for(srcs[0], kind)), exist_ok=True)
    hdr_mtime = max([os.path.getmtime(h) for h in _headers()] + [os.path.getmtime(__file__)])
    log = []
    jobs = []
    for s in srcs:
        o = _obj_for(s, kind)
        if not force and _obj_fresh(s, o, hdr_mtime):
            continue
        if kind == "cuda":
            jobs.append([NVCC] + CUDA_FLAGS + ["-c", s, "-o", o])
        else:
            jobs.append([CXX] + HOST_FLAGS + ["-c", s, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 4))) as ex:
        for out in ex.map(lambda c: _run(c, log), jobs):
            if verbose:
                sys.stdout.write(out)
    objs = [_obj_for(s, kind) for s in srcs]
    lib = lib_path(kind)
    if kind == "cuda":
        _run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared",
                                                    "-Xcompiler", "-fPIC", "-lpthread", "-ldl"], log)
    else:
        _run([CXX, "-shared", "-o", lib] + objs + ["-pthread", "-ldl", "-lrt"] + HOST_LINK_FLAGS, log)
    with open(lib + ".stamp", "w") as f:
        f.write(source_hash(kind))
    with open(os.path.join(OBJ, "build_%s.log" % kind), "w") as f:
        f.write("\n".join(log))
    return lib


def build_all(force=False, verbose=False):
    return [build("host", force, verbose), build("cuda", force, verbose)]


if __name__ == "__main__":
    force = "--force" in sys.argv
    for p in build_all(force=force, verbose="-v" in sys.argv):
        print("built", p)
