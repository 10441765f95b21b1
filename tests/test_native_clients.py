"""The drop-in boundary exercised from C99 and C++11, the way Go/C#/C and header-level C++ callers would bind it."""
import os
import subprocess
import sys

import numpy as np
import pytest

import common
from oracle import bindings

NATIVE = os.path.join(common.ROOT, "tests", "native")
INCLUDE = os.path.join(common.ROOT, "include")
LIBDIR = os.path.join(common.ROOT, "usearch_b200")


def _compile(tmp_path):
    from usearch_b200 import build
    build.build()
    c_bin, cpp_bin = str(tmp_path / "test_c_abi"), str(tmp_path / "test_cpp_mirror")
    link = [f"-L{LIBDIR}", "-lusearch_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"]
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{INCLUDE}",
                    os.path.join(NATIVE, "test_c_abi.c"), "-o", c_bin, *link], check=True, capture_output=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{INCLUDE}",
                    os.path.join(NATIVE, "test_cpp_mirror.cpp"), "-o", cpp_bin, *link], check=True, capture_output=True)
    return c_bin, cpp_bin


def test_native_clients_compile_against_the_headers(tmp_path):
    """C99 / C++11, -Wall -Wextra -Werror: the headers are plain C ABI (no CUDA, no torch, no C++ in the .h)."""
    _compile(tmp_path)


@pytest.mark.gpu
def test_native_clients_run(tmp_path):
    c_bin, cpp_bin = _compile(tmp_path)
    g = np.load(os.path.join(common.GOLDEN, "cos_f32_n2000_d64.npz"))
    index_path = str(tmp_path / "golden.usearch")
    np.asarray(g["blob"], dtype=np.uint8).tofile(index_path)
    queries = np.ascontiguousarray(g["queries"], dtype=np.float32)
    k = 10
    keys, dist, counts, _, _ = bindings.PortIndex(g["blob"], 64).search(queries, k, threads=2)
    cases = str(tmp_path / "cases.bin")
    with open(cases, "wb") as f:
        np.array([queries.shape[0], queries.shape[1], k], dtype=np.uint64).tofile(f)
        queries.tofile(f)
        keys.tofile(f)
        dist.tofile(f)
        counts.astype(np.uint64).tofile(f)
    for binary, marker in ((c_bin, "C_ABI_OK"), (cpp_bin, "CPP_MIRROR_OK")):
        out = subprocess.run([binary, index_path, cases], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and marker in out.stdout, out.stdout + out.stderr


def test_half_words_summation_order(tmp_path):
    """Host emulation of the 4-lane, by-accumulator split of the f16/bf16 metrics (metrics.cuh *_halfw_t): the bits
    of the pinned oracle for l2sq / ip / cos."""
    exe = tmp_path / "half_words"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-std=c99", "-I", os.path.join(common.ROOT, "oracle"),
                    os.path.join(common.ROOT, "tests", "native", "test_half_words_order.c"), "-o", str(exe), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "HALF_WORDS_ORDER_OK" in out


def test_key_map_native(tmp_path):
    """The host-side key -> slot table (contains / count / get / remove / rename of the C ABI) as a plain C++11 unit test:
    a multi-index with removed entries, growth from a wrong size hint, erase and re-probe, against std::multimap."""
    exe = tmp_path / "key_map"
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(common.ROOT, "usearch_b200", "csrc"),
                    os.path.join(NATIVE, "test_key_map.cpp"), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "KEY_MAP_OK" in out
