"""ctypes loader for oracle/clik_oracle.c (TEST / BASELINE INFRASTRUCTURE ONLY; see the header of
clik_oracle.c).  Importable from tests/ and bench.py only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libclik_oracle.so")


class Joint(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("xyz", ctypes.c_double * 3), ("R", ctypes.c_double * 9),
                ("axis", ctypes.c_double * 3)]


_lib = None


def load(build_if_missing=True):
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "clik_oracle.c")
        if (not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src)) and build_if_missing:
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = ctypes.CDLL(LIB)
        _lib.clik_ref_pinv_track.restype = ctypes.c_int
        _lib.clik_ref_pinv_track.argtypes = [ctypes.POINTER(Joint), ctypes.c_int, ctypes.c_int,
                                             ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
                                             ctypes.c_int]
        _lib.clik_ref_max_threads.restype = ctypes.c_int
    return _lib


def chain_table(chain):
    """clik_oracle.load_chain(...) output -> C array of joints."""
    arr = (Joint * len(chain))()
    for k, (jt, xyz, Ro, axis, _, _) in enumerate(chain):
        arr[k].type = 1 if jt in ("revolute", "continuous") else 0
        arr[k].xyz[:] = list(map(float, xyz))
        arr[k].R[:] = list(map(float, np.asarray(Ro).reshape(-1)))
        a = np.asarray(axis, dtype=float)
        arr[k].axis[:] = list(a / np.linalg.norm(a))
    return arr


def max_threads():
    """All host cores this process may use.  Deliberately not omp_get_max_threads(): torchrun
    exports OMP_NUM_THREADS=1, which would silently turn the CPU baseline into a 1-core run."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def pinv_track(chain, q, y, gain=1.0, lam=1e-7, threads=0):
    """q (n, N), y (3, N) float64 coordinate-major -> qdot (n, N), threads used."""
    lib = load()
    q = np.ascontiguousarray(q, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, N = q.shape
    out = np.empty_like(q)
    tab = chain_table(chain)
    used = lib.clik_ref_pinv_track(tab, len(chain), n, N, q.ctypes.data, y.ctypes.data,
                                   float(gain), float(lam), out.ctypes.data, int(threads))
    return out, used
