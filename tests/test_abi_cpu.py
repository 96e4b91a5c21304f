"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol of include/aos_gpu.h."""
import ctypes
import os
import re

from aos_gpu import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "aos_gpu.h")).read()
    return sorted(set(re.findall(r"AOS_API\s+[\w\s\*]+?\b(aos_\w+)\s*\(", hdr)))


def test_library_built_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    L = ctypes.CDLL(lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/aos_gpu.h but not exported"
    assert sorted(lib.EXPORTED_SYMBOLS) == declared


def test_grid_geometry_matches_baseline_sizes():
    """generateOccupancyGrid dims (seed_gen:587-596): pure host code, no GPU needed."""
    from aos_gpu import synth
    for name, wh in (("C1", (1000, 600)), ("C2", (2000, 1200)), ("C3", (20000, 20000)), ("C4", (40000, 40000))):
        spec = synth.config(name)
        gi = lib.grid_geometry(lib.SeedParams(grid_resolution=spec.grid_resolution, polygon=spec.polygon))
        assert (gi.width, gi.height) == wh
    gi = lib.grid_geometry(lib.SeedParams(polygon=synth.REFERENCE_POLYGON))
    assert (gi.width, gi.height) == (1546, 296)  # the field map of the default polygon


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    try:
        lib.Context(0)
    except lib.AosError:
        return
    raise AssertionError("Context() must raise without a CUDA device (no CPU fallback)")


def test_header_is_plain_c(tmp_path):
    """include/aos_gpu.h must be consumable from C (the reference's nodes are C++, but the boundary is a C-ABI):
    compile a C99 translation unit that takes the address of every declared entry point."""
    import subprocess
    names = _declared_symbols()
    src = tmp_path / "abi.c"
    body = "\n".join(f"  p[{i}] = (void (*)(void))&{n};" for i, n in enumerate(names))
    src.write_text('#include "aos_gpu.h"\n'
                   f"void (*p[{len(names)}])(void);\n"
                   "int main(void) {\n" + body + "\n  aos_seed_params sp; aos_gvd_graph g; aos_band b; aos_batch_item it;\n"
                   "  (void)sp; (void)g; (void)b; (void)it;\n  return p[0] == 0;\n}\n")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        "-c", str(src), "-o", str(tmp_path / "abi.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_example_links_against_the_library(tmp_path):
    """examples/map_to_graph.cpp is what a node-side caller looks like: it must compile as C++17 and link against
    libaos_gpu.so (running it needs a GPU: tests/test_gvd_gpu.py::test_cpp_example_runs)."""
    import subprocess
    import __graft_entry__ as g
    g.build()
    exe = tmp_path / "map_to_graph"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "map_to_graph.cpp"), "-L", os.path.dirname(lib.LIB_PATH), "-laos_gpu",
                        "-Wl,-rpath," + os.path.dirname(lib.LIB_PATH), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
