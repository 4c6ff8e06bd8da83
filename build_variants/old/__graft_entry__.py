"""Driver entry points: build() compiles every CUDA extension for sm_100a, smoke() runs one tiny render on cuda:0."""
import importlib.util
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "cg-raytracer_b200")
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.environ.get("CGRT_LIB") or os.path.join(PKG_DIR, "libcgrt_b200.so")

# Strict arithmetic: no FMA contraction, IEEE division / square root, no flush-to-zero.
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
]
SOURCES = ["csrc/cgrt_kernels.cu", "csrc/cgrt_capi.cu", "csrc/bvh_build.cpp", "host/host_api.cpp", "host/obj_loader.cpp"]


def load_package():
    """Import the hyphen-named package directory as module `cg_raytracer_b200`."""
    name = "cg_raytracer_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    """Compile libcgrt_b200.so (CUDA, sm_100a) in-tree, the oracle's C++ restatement and — when /root/reference is
    present — the verbatim-reference checker oracle/_ref (building the checker is not using it)."""
    srcs = [os.path.join(PKG_DIR, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "cgrt_b200.h")]
    host_dir = os.path.join(PKG_DIR, "host")
    if os.path.isdir(host_dir):
        deps += [os.path.join(host_dir, f) for f in os.listdir(host_dir)]
    if force or _newer(LIB, deps):
        extra = os.environ.get("CGRT_NVCC_EXTRA", "").split()  # experiment switches (-DCGRT_...); speed only
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-I", os.path.join(ROOT, "include"), "-I", host_dir, "-I", os.path.join(host_dir, "compat"), "-shared", "-o", LIB] + srcs + \
              ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, cwd=PKG_DIR, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libcgrt_b200.so")
    # headless C++ harness over the reference-named host interface (cgrt_host.h) linked against the library
    cli = os.path.join(PKG_DIR, "cgrt_cli")
    cli_src = os.path.join(host_dir, "cgrt_cli.cpp")
    if os.path.exists(cli_src) and (force or _newer(cli, [cli_src, LIB, os.path.join(host_dir, "cgrt_host.h")])):
        r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", host_dir, "-I", os.path.join(host_dir, "compat"),
                            "-I", os.path.join(ROOT, "include"), cli_src, "-o", cli, "-L", PKG_DIR, "-lcgrt_b200",
                            "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building cgrt_cli")
    # CPU checkers (test infrastructure)
    mk = os.path.join(ROOT, "oracle", "Makefile")
    if os.path.exists(mk):
        r = subprocess.run(["make", "-s", "-f", mk, "all"], cwd=os.path.join(ROOT, "oracle"), capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("oracle build failed")
    pkg = load_package()
    pkg.capi.load_library()  # dlopen + resolve every exported symbol
    return None


def smoke():
    """One small invocation of the hot path on cuda:0 (Cornell-like box scene, 64x64, 1 light, trace limit 2),
    checked against the CPU oracle (test infrastructure; the product path never touches it)."""
    import numpy as np
    pkg = load_package()
    capi = pkg.capi
    if capi.device_count() < 1:
        raise RuntimeError("smoke(): no CUDA device visible; the product has no CPU fallback")
    sys.path.insert(0, ROOT)
    from oracle import bindings as ob
    flat = ob.random_soup(3000, seed=7, scale=0.08, n_meshes=3)
    lights = np.array([[0.0, 0.9, 0.0, 1, 1, 1]], np.float32)
    W = H = 64
    scene = capi.Scene(flat, lights=lights, device=0)
    cam = capi.make_camera(W, H)
    rgb, stats = scene.render(cam, W, H, trace_limit=2)
    rays = ob.random_rays(4096, seed=11)
    hits = scene.intersect(rays)
    try:
        chk = ob.RefLib()
    except (FileNotFoundError, OSError):
        chk = ob.OracleLib()
    cs = chk.scene(flat, lights)
    cb = cs.bvh(mode=1)
    ocam = ob.default_camera(W, H)
    orgb, ocnt = cb.render(ocam, W, H, trace_limit=2)
    ohits = cb.intersect(rays)
    canon = flat.canonical_ids()
    gid = np.where(hits["tri"] >= 0, canon[np.maximum(hits["tri"], 0)], hits["tri"])
    id_match = float((gid == ohits["tri"]).mean())
    max_px = float(np.abs(rgb - orgb).max())
    print(f"[smoke] rays: primary={stats['primary']} shadow={stats['shadow']} bounce={stats['bounce']} "
          f"launches={stats['kernel_launches']} id_match={id_match:.6f} max|dpix|={max_px:.3g}")
    assert stats["primary"] == ocnt["primary"] and stats["shadow"] == ocnt["shadow"] and stats["bounce"] == ocnt["bounce"], (stats, ocnt)
    assert id_match >= 0.9999, id_match
    assert np.array_equal(hits["t"].view(np.uint32), ohits["t"].view(np.uint32)) or np.allclose(hits["t"], ohits["t"], rtol=1e-4)
    assert max_px <= 1.0 / 255.0, max_px
    scene.close()
    return None


if __name__ == "__main__":
    build(verbose="-v" in sys.argv)
    if "smoke" in sys.argv:
        smoke()
