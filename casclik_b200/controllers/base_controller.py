"""Shared pieces of the controllers (the reference's BaseController only defines __repr__,
casclik/controllers/base_controller.py:1-6; the batch plumbing below is new)."""
import ctypes

import numpy as np

from .. import sym as cs
from .. import runtime


class BaseController(object):
    controller_type = "BaseController"

    def __repr__(self):
        return "%s<%s>" % (self.controller_type, self.skill_spec.label)


def resolve_devices(devices):
    """`devices` argument of solve_batch -> list of CUDA ordinals, or None for "the default device".
    "all": every visible GPU; an int n: devices 0..n-1; a sequence: those ordinals."""
    if devices is None:
        return None
    if isinstance(devices, str):
        if devices != "all":
            raise ValueError('devices must be None, "all", a count or a list of CUDA ordinals')
        n = runtime.device_count()
        if n < 1:
            runtime.require_device()
        return list(range(n))
    if isinstance(devices, int):
        if devices < 1 or devices > runtime.device_count():
            raise ValueError("devices=%d but %d CUDA devices are visible" % (devices, runtime.device_count()))
        return list(range(devices))
    devs = [int(d) for d in devices]
    if not devs or len(set(devs)) != len(devs):
        raise ValueError("devices must be a non-empty list of distinct CUDA ordinals")
    return devs


def skill_handle_array(skills):
    arr = (ctypes.c_void_p * len(skills))(*[s.handle for s in skills])
    return arr


def as_vector(value, n, what):
    """float / list / ndarray / DM -> float64 array of length n."""
    if isinstance(value, cs.GenericMatrixCommon):
        value = value.toarray()
    a = np.asarray(value, dtype=np.float64).reshape(-1)
    if a.size != n:
        raise ValueError("%s has %d entries, the skill expects %d" % (what, a.size, n))
    return a


def dm_column(a):
    return cs.DM(np.asarray(a, dtype=np.float64).reshape(-1, 1))


class Batch(object):
    """Validated structure-of-arrays view of (t, q, x, y) for N instances.

    Device path: torch CUDA float64 tensors, q of shape (n_robot, N) (coordinate-major, so
    instance i of coordinate j sits at q[j, i]); t a tensor of shape (N,) or a Python float.
    Host path: the same shapes as NumPy arrays (copied through the pipelined *_host ABI).
    """

    def __init__(self, n_rob, n_virt, n_in, t, q, x, y):
        self.on_device = runtime._is_torch(q)
        if self.on_device:
            import torch
            self.torch = torch
            if q.dim() != 2 or q.shape[0] != n_rob:
                raise ValueError("robot_var batch must have shape (%d, N), got %s" % (n_rob, tuple(q.shape)))
            self.N = int(q.shape[1])
            self.device = q.device
            if not torch.is_tensor(t):
                t = torch.full((1,), float(t), dtype=torch.float64, device=q.device)
            self.t_stride = 0 if t.numel() == 1 else 1
            if self.t_stride and t.numel() != self.N:
                raise ValueError("time_var batch must have N entries")
            if t.device != q.device:
                raise ValueError("time_var lives on %s, robot_var on %s" % (t.device, q.device))
            self.t = t
            self.tp = runtime.dev_ptr(t, "f64", t.numel(), "time_var")
            self.qp = runtime.dev_ptr(q, "f64", n_rob * self.N, "robot_var")
            self.xp = self._dev(x, n_virt, "virtual_var")
            self.yp = self._dev(y, n_in, "input_var")
            self.keep = (t, q, x, y)
        else:
            q = np.ascontiguousarray(q, dtype=np.float64)
            if q.ndim != 2 or q.shape[0] != n_rob:
                raise ValueError("robot_var batch must have shape (%d, N), got %s" % (n_rob, q.shape))
            self.N = q.shape[1]
            t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(-1))
            self.t_stride = 0 if t.size == 1 else 1
            if self.t_stride and t.size != self.N:
                raise ValueError("time_var batch must have N entries")
            self.tp = ctypes.c_void_p(t.ctypes.data)
            self.qp = ctypes.c_void_p(q.ctypes.data)
            self.xp, x = self._host(x, n_virt, "virtual_var")
            self.yp, y = self._host(y, n_in, "input_var")
            self.keep = (t, q, x, y)

    def _dev(self, a, rows, what):
        if rows == 0:
            return None
        if a is not None and (not runtime._is_torch(a) or a.device != self.device):
            raise ValueError("%s must be a CUDA tensor on %s like robot_var" % (what, self.device))
        if a is None:
            a = self.torch.zeros((rows, self.N), dtype=self.torch.float64, device=self.device)
            self._zeros = getattr(self, "_zeros", []) + [a]
        if a.dim() != 2 or a.shape[0] != rows:
            raise ValueError("%s batch must have shape (%d, N)" % (what, rows))
        return runtime.dev_ptr(a, "f64", rows * self.N, what)

    def _host(self, a, rows, what):
        if rows == 0:
            return None, None
        if a is None:
            a = np.zeros((rows, self.N))
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (rows, self.N):
            raise ValueError("%s batch must have shape (%d, N), got %s" % (what, rows, a.shape))
        return ctypes.c_void_p(a.ctypes.data), a

    def empty(self, rows, dtype="f64"):
        """Output array (rows, N) on the same side as the inputs."""
        if self.on_device:
            dt = {"f64": self.torch.float64, "i32": self.torch.int32}[dtype]
            shape = (rows, self.N) if rows > 0 else (self.N,)
            return self.torch.empty(shape, dtype=dt, device=self.device)
        dt = {"f64": np.float64, "i32": np.int32}[dtype]
        return np.empty((rows, self.N) if rows > 0 else (self.N,), dtype=dt)

    def ptr(self, arr):
        if arr is None:
            return None
        if self.on_device:
            return ctypes.c_void_p(arr.data_ptr())
        return ctypes.c_void_p(arr.ctypes.data)

    def out_ptr(self, arr, rows, dtype, what):
        """Validated pointer of an output buffer (rows, N) — (N,) when rows == 0 — that the kernel
        writes in place: dtype, element count, contiguity and (device path) the device must match,
        because a wrong buffer would be an out-of-bounds write, not an exception."""
        if arr is None:
            return None
        numel = (rows if rows > 0 else 1) * self.N
        if self.on_device:
            if not runtime._is_torch(arr):
                raise runtime.ClikError("%s must be a CUDA tensor when the inputs are CUDA tensors" % what)
            if arr.device != self.device:
                raise runtime.ClikError("%s lives on %s, the inputs on %s" % (what, arr.device, self.device))
            return runtime.dev_ptr(arr, {"f64": "f64", "i32": "i32"}[dtype], numel, what)
        return runtime.host_out_ptr(arr, {"f64": np.float64, "i32": np.int32}[dtype], numel, what)

    @property
    def device_index(self):
        """CUDA ordinal the batch runs on: the tensors' device, or the default for host arrays."""
        if self.on_device:
            return self.device.index if self.device.index is not None else self.torch.cuda.current_device()
        return runtime.current_device()

    def stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
