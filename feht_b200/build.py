"""Build libfeht_cuda.so (sm_100a) in-tree with nvcc, and the CPU oracle with gcc.

The shared library is plain C ABI (include/feht_cuda.h); there is no torch extension.
`python -m feht_b200.build` or `__graft_entry__.build()` runs this.  nvcc cross-compiles
without a GPU.  The .so lands next to this file so that it travels with the tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "feht_b200" / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = ROOT / "feht_b200" / "libfeht_cuda.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", str(ROOT / "include")]
# per translation unit extras
EXTRA = {
    # the reference p-value arithmetic is one IEEE rounding per operation: never contract a*b+c
    "fet.cu": ["-fmad=false"],
    "choose_table.cpp": ["-Xcompiler", "-ffp-contract=off"],
}


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; libfeht_cuda has no CPU fallback and cannot be built")
    return cand


HOST = ROOT / "feht_b200" / "host"
CLI = ROOT / "feht_b200" / "bin" / "feht-b200"


def _sources() -> list[Path]:
    # kernels + C ABI, and the C++ host mirror (everything except the CLI's main) in one shared library
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + [HOST / "feht_host.cpp"])


def _stamp(src: Path, flags: list[str]) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for dep in sorted(list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))):
        h.update(dep.read_bytes())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build_cuda(verbose: bool = False, ptxas_info: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    objs = []
    rebuilt = False
    for src in _sources():
        flags = ARCH + COMMON + EXTRA.get(src.name, []) + os.environ.get("FEHT_EXTRA_NVCC_FLAGS", "").split()
        if ptxas_info and src.suffix == ".cu":
            flags = flags + ["-Xptxas", "-v"]
        obj = OBJ / (src.name + ".o")
        stamp_file = OBJ / (src.name + ".stamp")
        stamp = _stamp(src, flags)
        if not (obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp):
            cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
            stamp_file.write_text(stamp)
            rebuilt = True
        objs.append(str(obj))
    if rebuilt or not LIB.exists():
        cmd = [nvcc, *ARCH, "-shared", "-o", str(LIB), *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    # the command line (feht's own options) linked against the library next to it
    main_src = HOST / "feht_main.cpp"
    if rebuilt or not CLI.exists() or CLI.stat().st_mtime < main_src.stat().st_mtime:
        CLI.parent.mkdir(parents=True, exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-I", str(ROOT / "include"), str(main_src), "-o", str(CLI),
               "-L", str(LIB.parent), "-lfeht_cuda", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath-link," + str(LIB.parent)]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


def build_oracle(verbose: bool = False) -> Path:
    """The checker (tests / smoke / cpu_baseline only)."""
    out = subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=not verbose)
    del out
    return ROOT / "oracle" / "libfeht_oracle.so"


def main(argv: list[str]) -> int:
    verbose = "-v" in argv
    lib = build_cuda(verbose=verbose, ptxas_info="--ptxas" in argv)
    ora = build_oracle(verbose=verbose)
    print(f"built {lib}\nbuilt {ora}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
