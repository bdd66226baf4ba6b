"""Host logic of the C++ wrapper (include/ssw_cpp.h, csrc/ssw_cpp.cpp) on CPU: tests/cpp_wrapper/driver.cpp is built
against our wrapper with the oracle shim answering ssw.h / ssw_batch.h, and must print what the same driver printed
when built against the UNMODIFIED reference wrapper and ssw.c (tests/golden/cpp_wrapper.txt).
The GPU suite repeats the comparison over libssw.so."""
import os
import subprocess

import common as C


def build_driver(out, extra):
    src = [os.path.join(C.ROOT, "tests", "cpp_wrapper", "driver.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-DWITH_BATCH", "-I" + os.path.join(C.ROOT, "include"), "-o", out] + src + extra
                   + ["-lm"], check=True)


def test_wrapper_matches_reference_wrapper(tmp_path):
    subprocess.run(["make", "-s", "-C", C.ORACLE_DIR, "all"], check=True)
    shim = str(tmp_path / "shim.o")
    subprocess.run(["gcc", "-O2", "-c", "-o", shim, os.path.join(C.ROOT, "tests", "cli_shim", "shim.c")], check=True)
    exe = str(tmp_path / "driver_cpu")
    build_driver(exe, [os.path.join(C.PKG, "csrc", "ssw_cpp.cpp"), shim, "-L" + C.ORACLE_DIR, "-lssw_oracle",
                       "-Wl,-rpath," + C.ORACLE_DIR])
    got = subprocess.run([exe], capture_output=True, text=True, timeout=300, check=True).stdout
    want = open(os.path.join(C.GOLDEN, "cpp_wrapper.txt")).read()
    assert got == want
    assert want.count("\n") > 100
