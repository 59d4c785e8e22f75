"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's occupancy-query + mesh-extraction path, used as the
checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing in icon_b200/ or lib/ may import from here.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile oracle/*.c (gcc, fp32, no implicit contraction, OpenMP) -> liboracle.so."""
    srcs = [os.path.join(_HERE, f) for f in ("sdf_oracle.c", "mc_oracle.c")
            if os.path.exists(os.path.join(_HERE, f))]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fopenmp", "-mfma", "-mavx2", "-ffp-contract=off",
           "-fno-fast-math", "-o", _SO] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib
