"""Build recipe for the oracle (TEST INFRASTRUCTURE — never imported by the product path).

Two artefacts:

* ``oracle/liboracle.so``            — our plain-C restatement of the reference search
  (``oracle/hnsw_oracle.c``), always buildable (gcc only).
* ``oracle/_ref/libusearch_ref_{parity,perf}.so`` — the UNMODIFIED reference headers under
  ``/root/reference`` compiled together with ``oracle/ref_driver.cpp``. Only built where
  ``/root/reference`` exists (the development container); the GPU box uses the prebuilt files that
  travel with the snapshot. Reference sources are compiled where they lie, never copied.

  - ``parity``: ``-O2 -ffp-contract=off`` and no ``-ffast-math`` so the pinned metrics are exact.
  - ``perf``:   the reference's production GNU flags ``-O3 -ffast-math -march=native``
    (/root/reference/CMakeLists.txt:232-254) — used as the timed CPU baseline.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("USEARCH_REFERENCE_DIR", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError("oracle build failed")


def build_port(force: bool = False) -> str:
    out = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, "hnsw_oracle.c"), os.path.join(HERE, "metrics_pinned.h")]
    if not force and _newer(out, srcs):
        return out
    _run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
          "-o", out, srcs[0], "-lm", "-lpthread"])
    return out


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "include", "usearch"))


def build_reference(flavour: str, force: bool = False) -> str | None:
    """Compile the reference where it lies. Returns the .so path, or the prebuilt one, or None."""
    out = os.path.join(REF_OUT, f"libusearch_ref_{flavour}.so")
    if not reference_available():
        return out if os.path.exists(out) else None
    os.makedirs(REF_OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref_driver.cpp"), os.path.join(HERE, "metrics_pinned.h")]
    if not force and _newer(out, srcs):
        return out
    common = [
        "-fPIC", "-DUSEARCH_USE_SIMSIMD=1", "-DUSEARCH_USE_FP16LIB=0", "-DUSEARCH_USE_OPENMP=0",
        "-DSIMSIMD_NATIVE_F16=0", "-DSIMSIMD_NATIVE_BF16=0", "-DSIMSIMD_DYNAMIC_DISPATCH=1",
        f"-I{REF}/include", f"-I{REF}/simsimd/include", f"-I{REF}/fp16/include", f"-I{HERE}", "-w",
    ]
    if flavour == "parity":
        opt = ["-O2", "-ffp-contract=off", "-march=x86-64-v3"]
    elif flavour == "perf":
        opt = ["-O3", "-ffast-math", "-march=native"]
    else:
        raise ValueError(flavour)
    obj_cpp = os.path.join(REF_OUT, f"ref_driver_{flavour}.o")
    obj_c = os.path.join(REF_OUT, f"simsimd_{flavour}.o")
    _run(["g++", "-std=c++17", *opt, *common, "-c", srcs[0], "-o", obj_cpp])
    # SimSIMD's dynamic-dispatch translation unit: each kernel carries its own target attribute.
    _run(["gcc", "-std=c11", "-O3", *common, "-c", f"{REF}/simsimd/c/lib.c", "-o", obj_c])
    _run(["g++", "-shared", "-o", out, obj_cpp, obj_c, "-lpthread", "-lm"])
    return out


def build_all(force: bool = False) -> dict:
    built = {"port": build_port(force)}
    for flavour in ("parity", "perf"):
        built[flavour] = build_reference(flavour, force)
    return built


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
