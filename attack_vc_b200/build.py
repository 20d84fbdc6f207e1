"""Build libavc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libavc_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas=-v",
]


VERSION_PREFIX = b"avc_b200 0.2 (sm_100a) src "


def sources():
    return sorted(CSRC.glob("*.cu"))


def source_hash() -> str:
    """sha256 over every file the library is compiled from (names + bytes) and the compiler flags."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))) + [HERE.parent / "include" / "avc_b200.h"]
    for d in deps:
        h.update(d.name.encode() + b"\0" + d.read_bytes() + b"\0")
    h.update(" ".join(NVCC_FLAGS + os.environ.get("AVC_NVCC_EXTRA", "").split()).encode())
    return h.hexdigest()[:16]


def built_hash(lib: Path = LIB):
    """The source hash baked into a built library (the tail of avc_version()), read from its bytes: no dlopen."""
    if not lib.exists():
        return None
    data = lib.read_bytes()
    i = data.find(VERSION_PREFIX)
    if i < 0:
        return None
    return data[i + len(VERSION_PREFIX): i + len(VERSION_PREFIX) + 16].decode(errors="replace")


def _stale() -> bool:
    # content, not mtime: a checkout that reorders timestamps must not leave a stale binary in use
    return built_hash() != source_hash()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into attack_vc_b200/libavc_b200.so.  Raises on failure (no fallback)."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libavc_b200.so cannot be built and there is no CPU fallback")
    extra = os.environ.get("AVC_NVCC_EXTRA", "").split()      # experiments only (e.g. -DAVC_SMALL_MINB=2)
    cmd = [nvcc, *NVCC_FLAGS, *extra, f'-DAVC_SRC_HASH="{source_hash()}"', *map(str, sources()), "-o", str(LIB)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libavc_b200.so")
    (HERE / "build_ptxas.log").write_text(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
