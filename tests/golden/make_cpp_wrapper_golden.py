#!/usr/bin/env python
"""Freeze the output of tests/cpp_wrapper/driver.cpp built against the UNMODIFIED reference C++ wrapper
(/root/reference/src/ssw_cpp.{h,cpp} + ssw.c compiled where they lie, products in a temp dir).
Build container only.  The same driver built against include/ssw_cpp.h + our implementation must print the same text
(tests/test_cpp_wrapper_host.py on CPU through the oracle shim, tests/test_gpu_parity.py on the GPU)."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
tmp = tempfile.mkdtemp()
exe = os.path.join(tmp, "driver_ref")
subprocess.run(["gcc", "-O2", "-c", "-o", os.path.join(tmp, "ssw.o"), os.path.join(REF, "ssw.c")], check=True)
subprocess.run(["g++", "-O2", "-std=c++17", "-I" + REF, "-o", exe, os.path.join(ROOT, "tests", "cpp_wrapper", "driver.cpp"),
                os.path.join(REF, "ssw_cpp.cpp"), os.path.join(tmp, "ssw.o"), "-lm"], check=True)
out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
with open(os.path.join(HERE, "cpp_wrapper.txt"), "w") as f:
    f.write(out)
print("froze", len(out.splitlines()), "lines")
