"""Builds the C++ host mirror of the reference's API (host/) against libbos_b200.so:
libproj02_b200.so (State, observations, parse_g2o, triangulate_landmarks, Solver) and the headless harness
bearing_only_slam.  g++ only; works without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libproj02_b200.so")
EXE = os.path.join(HERE, "bearing_only_slam")
SYNTH = os.path.join(os.path.dirname(HERE), "synth")   # input generator of the harness' --synth mode: its own library
SOURCES = ["framework/state.cpp", "utils/g2o_utils.cpp", "slam/triangulation.cpp", "slam/solver.cpp"]
HEADERS = ["framework/linalg.hpp", "framework/definitions.hpp", "framework/state.hpp", "framework/observation.hpp", "utils/g2o_utils.hpp",
           "slam/triangulation.hpp", "slam/solver.hpp", "../../include/bos_b200.h"]
CXXFLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wextra"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def build(force=False):
    srcs = [os.path.join(HOST, s) for s in SOURCES]
    deps = srcs + [os.path.join(HOST, h) for h in HEADERS] + [os.path.abspath(__file__)]
    link = ["-L" + HERE, "-lbos_b200", "-Wl,-rpath,$ORIGIN"]
    if force or _stale(LIB, deps):
        _run(["g++"] + CXXFLAGS + ["-shared", "-o", LIB] + srcs + link)
    exe_src = os.path.join(HOST, "executables", "bearing_only_slam.cpp")
    import sys as _sys
    if os.path.dirname(HERE) not in _sys.path:
        _sys.path.insert(0, os.path.dirname(HERE))
    from synth import build as build_synth
    synth_lib = build_synth()
    if force or _stale(EXE, deps + [exe_src, LIB, synth_lib]):
        _run(["g++"] + CXXFLAGS + ["-I" + SYNTH, "-o", EXE, exe_src, "-L" + HERE, "-lproj02_b200", "-lbos_b200", "-L" + SYNTH, "-lbos_synth",
                                   "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../synth"])
    return LIB, EXE


def build_test(src, out):
    """Compiles a C++ test program against the host library (used by tests/)."""
    _run(["g++"] + CXXFLAGS + ["-I" + HOST, "-o", out, src, "-L" + HERE, "-lproj02_b200", "-lbos_b200", "-Wl,-rpath," + HERE])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
