"""The header-only C++ mirror compiles against include/rfgpu.h and its host-side logic (element widening by value,
HashableChar semantics) runs without a GPU; the GPU entry points are only instantiated here
(tests/test_gpu_parity.py::test_cpp_host_mirror_known_answers runs them on the device)."""
import os
import subprocess


def test_cpp_mirror_host_logic(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "rapidfuzz-rs_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "librfgpu.so")):
        import __graft_entry__ as g
        g.build()
    exe = str(tmp_path / "test_cpp_host")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "include"),
                           "-I", os.path.join(root, "rapidfuzz-rs_b200", "cpp"), os.path.join(root, "tests", "cpp", "test_cpp_host.cpp"),
                           "-o", exe, "-L", libdir, "-lrfgpu", "-Wl,-rpath," + libdir])
    assert subprocess.run([exe], timeout=60).returncode == 0
