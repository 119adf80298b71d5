"""Builds tests/cpp/vfw_caller (a CodecInst-shaped C++ caller of include/screencodec_b200.h) with g++ against libscpr_b200.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "vfw_caller.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "build", "vfw_caller")


def build() -> str:
    libdir = os.path.join(ROOT, "screenpressor_b200")
    deps = [SRC, os.path.join(ROOT, "include", "screencodec_b200.h"), os.path.join(ROOT, "include", "scpr_c.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L" + libdir,
                        "-lscpr_b200", "-Wl,-rpath," + libdir], check=True)
    return EXE
