"""ctypes binding of ``libgwin_b200.so`` (C ABI declared in ``include/gwin_b200.h``).

There is no CPU fallback: if the shared library is missing or no sm_100 device is
present every entry point raises.  ``build()`` compiles the library in-tree with
nvcc for sm_100a (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("GWIN_B200_LIB") or os.path.join(_HERE, "libgwin_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", "gwin_cuda.cu"),
           os.path.join(_HERE, "csrc", "gwin_ensemble.cu")]
HEADER = os.path.join(ROOT, "include", "gwin_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]

GWIN_NPARAM = 13
GWIN_MAX_DET = 4
MODE_GAUSSIAN = 0
MODE_MARG_PHASE = 1
MODE_MARG_DIST = 2
MODE_MARG_DIST_PHASE = 3
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_RECURRENCE, KERNEL_TABLE = 0, 1, 2, 3

# Every symbol include/gwin_b200.h declares (checked by tests/test_abi.py).
SYMBOLS = (
    "gwin_plan_create", "gwin_plan_destroy", "gwin_plan_cutoffs",
    "gwin_plan_lognl", "gwin_plan_configure", "gwin_plan_set_distance_marg",
    "gwin_plan_set_calibration", "gwin_plan_cover_host", "gwin_plan_cover_device",
    "gwin_plan_table_info", "gwin_plan_slow_path", "gwin_loglr_batch_calib",
    "gwin_loglr_batch_calib_host", "gwin_generate_calib_host",
    "gwin_loglr_batch",
    "gwin_loglr_batch_host", "gwin_loglr_batch_strided",
    "gwin_loglr_batch_host_strided", "gwin_generate_host",
    "gwin_loglr_batch_points_host", "gwin_loglr_batch_points_submit", "gwin_loglr_batch_points_wait",
    "gwin_host_alloc", "gwin_host_free",
    "gwin_plan_last_launches",
    "gwin_plan_kernel_timing", "gwin_plan_kernel_timing_detail", "gwin_plan_table_blocks",
    "gwin_ens_propose", "gwin_ens_prior", "gwin_ens_prior_calib", "gwin_ens_accept", "gwin_ens_swap",
    "gwin_ens_swap_all", "gwin_ens_record", "gwin_ens_tick",
    "gwin_fp64_peak", "gwin_fp64_probe", "gwin_last_error", "gwin_version",
)

ENS_MAXDIM = 64
PRIOR_UNIFORM, PRIOR_UNIFORM_ANGLE, PRIOR_SIN_ANGLE, PRIOR_COS_ANGLE, \
    PRIOR_UNIFORM_RADIUS, PRIOR_GAUSSIAN, PRIOR_NONE = range(7)
TARGET_MCHIRP, TARGET_Q, TARGET_CALIB0 = 100, 101, 200


class PriorSpec(C.Structure):
    """``gwin_prior_spec`` of include/gwin_b200.h."""
    _fields_ = [("ndim", C.c_int32),
                ("type", C.c_int32 * ENS_MAXDIM),
                ("target", C.c_int32 * ENS_MAXDIM),
                ("lo", C.c_double * ENS_MAXDIM),
                ("hi", C.c_double * ENS_MAXDIM),
                ("mu", C.c_double * ENS_MAXDIM),
                ("sigma", C.c_double * ENS_MAXDIM),
                ("fixed", C.c_double * GWIN_NPARAM),
                ("derived_masses", C.c_int32),
                ("derived_type", C.c_int32 * 2),
                ("derived_lo", C.c_double * 2), ("derived_hi", C.c_double * 2),
                ("derived_mu", C.c_double * 2), ("derived_sigma", C.c_double * 2)]


_lock = threading.Lock()
_lib = None


class GwinCudaError(RuntimeError):
    """Raised when the CUDA library reports a failure (or is missing)."""


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES + [HEADER]
               if os.path.exists(s))


def build(force=False, verbose=False):
    """Compile ``libgwin_b200.so`` in-tree for sm_100a."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise GwinCudaError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


def load():
    """Load the shared library (building it first if the source is newer)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:
                build()
            except (OSError, GwinCudaError) as e:
                raise GwinCudaError(
                    "libgwin_b200.so is not built and could not be compiled "
                    "(%s); gwin_b200 has no CPU fallback" % e)
        lib = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        lib.gwin_plan_create.restype = C.c_int
        lib.gwin_plan_create.argtypes = [
            C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_double,
            C.c_double, C.c_double, C.c_double, C.c_double,
            dp, dp, dp, dp, C.c_double]
        lib.gwin_plan_destroy.restype = None
        lib.gwin_plan_destroy.argtypes = [C.c_void_p]
        lib.gwin_plan_cutoffs.restype = C.c_int
        lib.gwin_plan_cutoffs.argtypes = [C.c_void_p, C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int64)]
        lib.gwin_plan_lognl.restype = C.c_int
        lib.gwin_plan_lognl.argtypes = [C.c_void_p, dp, dp]
        lib.gwin_plan_set_distance_marg.restype = C.c_int
        lib.gwin_plan_set_distance_marg.argtypes = [C.c_void_p, C.c_int64,
                                                    C.c_void_p, C.c_void_p]
        lib.gwin_plan_cover_host.restype = C.c_int
        lib.gwin_plan_cover_host.argtypes = [C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_int64, C.c_int64]
        lib.gwin_plan_cover_device.restype = C.c_int
        lib.gwin_plan_cover_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p,
                                               C.c_int64, C.c_int64, C.c_void_p]
        lib.gwin_plan_table_info.restype = C.c_int
        lib.gwin_plan_table_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4
        lib.gwin_plan_slow_path.restype = C.c_int
        lib.gwin_plan_slow_path.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.gwin_plan_set_calibration.restype = C.c_int
        lib.gwin_plan_set_calibration.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.gwin_loglr_batch_calib.restype = C.c_int
        lib.gwin_loglr_batch_calib.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
            C.c_void_p, C.c_int64, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_loglr_batch_calib_host.restype = C.c_int
        lib.gwin_loglr_batch_calib_host.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
            C.c_void_p, C.c_int64, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_generate_calib_host.restype = C.c_int
        lib.gwin_generate_calib_host.argtypes = [C.c_void_p, dp, C.c_void_p, dp,
                                                 C.POINTER(C.c_int64)]
        lib.gwin_plan_configure.restype = C.c_int
        lib.gwin_plan_configure.argtypes = [C.c_void_p, C.c_int, C.c_double,
                                            C.c_double]
        lib.gwin_loglr_batch.restype = C.c_int
        lib.gwin_loglr_batch.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_loglr_batch_host.restype = C.c_int
        lib.gwin_loglr_batch_host.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_loglr_batch_strided.restype = C.c_int
        lib.gwin_loglr_batch_strided.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_loglr_batch_host_strided.restype = C.c_int
        lib.gwin_loglr_batch_host_strided.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_loglr_batch_points_host.restype = C.c_int
        lib.gwin_loglr_batch_points_host.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gwin_loglr_batch_points_submit.restype = C.c_int
        lib.gwin_loglr_batch_points_submit.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        lib.gwin_loglr_batch_points_wait.restype = C.c_int
        lib.gwin_loglr_batch_points_wait.argtypes = [C.c_void_p, C.c_int]
        lib.gwin_host_alloc.restype = C.c_int
        lib.gwin_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        lib.gwin_host_free.restype = None
        lib.gwin_host_free.argtypes = [C.c_void_p]
        lib.gwin_generate_host.restype = C.c_int
        lib.gwin_generate_host.argtypes = [C.c_void_p, dp, dp,
                                           C.POINTER(C.c_int64)]
        lib.gwin_plan_last_launches.restype = C.c_int
        lib.gwin_plan_last_launches.argtypes = [C.c_void_p]
        lib.gwin_plan_kernel_timing.restype = C.c_int
        lib.gwin_plan_kernel_timing.argtypes = [C.c_void_p, C.c_int, dp,
                                                C.POINTER(C.c_int)]
        lib.gwin_plan_kernel_timing_detail.restype = C.c_int
        lib.gwin_plan_kernel_timing_detail.argtypes = [C.c_void_p, dp]
        lib.gwin_plan_table_blocks.restype = C.c_int
        lib.gwin_plan_table_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        lib.gwin_fp64_peak.restype = C.c_int
        lib.gwin_fp64_peak.argtypes = [C.c_int, dp]
        vp, u64, u32, ci, cd = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
        lib.gwin_ens_propose.restype = ci
        lib.gwin_ens_propose.argtypes = [u64, u32, vp, ci, ci, ci, ci, cd, ci, ci, vp, vp, vp, vp]
        lib.gwin_ens_prior.restype = ci
        lib.gwin_ens_prior.argtypes = [C.POINTER(PriorSpec), C.c_int64, vp, vp, vp, vp, vp]
        lib.gwin_ens_prior_calib.restype = ci
        lib.gwin_ens_prior_calib.argtypes = [C.POINTER(PriorSpec), C.c_int64, vp, vp, vp, vp, vp, ci, vp]
        lib.gwin_ens_accept.restype = ci
        lib.gwin_ens_accept.argtypes = [u64, u32, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp,
                                        vp, vp, vp, vp, vp]
        lib.gwin_ens_swap.restype = ci
        lib.gwin_ens_swap.argtypes = [u64, u32, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.gwin_ens_swap_all.restype = ci
        lib.gwin_ens_swap_all.argtypes = [u64, u32, vp, ci, ci, ci, vp, vp, C.c_int64,
                                          vp, vp, vp, vp, vp]
        lib.gwin_ens_record.restype = ci
        lib.gwin_ens_record.argtypes = [C.c_int64, ci, C.c_int64, u32, vp,
                                        vp, vp, vp, vp, vp, vp, vp]
        lib.gwin_ens_tick.restype = ci
        lib.gwin_ens_tick.argtypes = [vp, vp]
        lib.gwin_fp64_probe.restype = C.c_int
        lib.gwin_fp64_probe.argtypes = [C.c_int, C.c_int, dp]
        lib.gwin_last_error.restype = C.c_char_p
        lib.gwin_version.restype = C.c_char_p
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load().gwin_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise GwinCudaError("gwin_b200 error %d: %s" % (rc, msg))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (``gwin_host_alloc``): the host entry points copy
    from / into such buffers directly instead of through their staging buffers."""
    import weakref
    lib = load()
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = C.c_void_p()
    check(lib.gwin_host_alloc(C.byref(ptr), nbytes))
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.gwin_host_free, ptr.value)
    return arr


class PendingCall(object):
    """A queued ``gwin_loglr_batch_points_submit``: ``wait()`` blocks until the results are in
    host memory (``gwin_loglr_batch_points_wait``) and returns them.  The input buffers are kept
    alive until then."""

    def __init__(self, plan, ticket, out, keep=()):
        self._plan, self._ticket, self._out, self._keep = plan, ticket, out, keep

    def wait(self):
        if self._ticket is not None:
            check(self._plan._lib.gwin_loglr_batch_points_wait(self._plan._h, self._ticket))
            self._ticket, self._keep = None, ()
        return self._out

    def __del__(self):
        # never leave a slot busy (and the device writing into freed host memory)
        try:
            if self._ticket is not None and self._plan._h:
                self._plan._lib.gwin_loglr_batch_points_wait(self._plan._h, self._ticket)
        except Exception:        # noqa: BLE001 -- interpreter shutdown
            pass


class Plan(object):
    """Owns one ``gwin_plan`` (device tables + workspace) on one GPU."""

    def __init__(self, device, delta_f, epoch, f_lower, f_upper, f_lower_wf,
                 det_response, det_location, data, psd, norm):
        lib = load()
        data = np.ascontiguousarray(data, dtype=np.complex128)
        n_det, n_f = data.shape
        resp = np.ascontiguousarray(det_response, dtype=np.float64).reshape(n_det, 9)
        loc = np.ascontiguousarray(det_location, dtype=np.float64).reshape(n_det, 3)
        if psd is not None:
            psd = np.ascontiguousarray(psd, dtype=np.float64)
            if psd.shape != (n_det, n_f):
                raise ValueError("psd shape %r does not match data %r"
                                 % (psd.shape, data.shape))
        handle = C.c_void_p()
        rc = lib.gwin_plan_create(
            C.byref(handle), int(device), int(n_det), int(n_f), float(delta_f),
            float(epoch), float(f_lower or 0.0), float(f_upper or 0.0),
            float(f_lower_wf), _dptr(resp), _dptr(loc),
            _dptr(data.view(np.float64)),
            _dptr(psd) if psd is not None else None,
            float(norm) if norm is not None else 0.0)
        check(rc)
        self._h = handle
        self._lib = lib
        self.device = int(device)
        self.n_det = int(n_det)
        self.n_cal = 0
        self.n_f = int(n_f)
        kmin, kmax = C.c_int64(), C.c_int64()
        check(lib.gwin_plan_cutoffs(self._h, C.byref(kmin), C.byref(kmax)))
        self.kmin, self.kmax = kmin.value, kmax.value
        tot = C.c_double()
        per = np.zeros(n_det)
        check(lib.gwin_plan_lognl(self._h, C.byref(tot), _dptr(per)))
        self.lognl = tot.value
        self.det_lognl = per

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gwin_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, kernel=KERNEL_AUTO, phase_tol=0.0, mchirp_min=0.0):
        check(self._lib.gwin_plan_configure(self._h, int(kernel),
                                            float(phase_tol), float(mchirp_min)))

    def set_distance_marg(self, dist, weight):
        """Distance grid and logsumexp weights (``deltad * dist_prior``) of the
        distance-marginalised modes; see ``gwin_plan_set_distance_marg``."""
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        if dist.shape != weight.shape or dist.ndim != 1:
            raise ValueError("dist and weight must be 1-D arrays of one length")
        check(self._lib.gwin_plan_set_distance_marg(
            self._h, dist.size, dist.ctypes.data, weight.ctypes.data))

    def cover(self, params=None, param_major=False, stream=None):
        """Declare what the moment tables must cover (``gwin_plan_cover_host`` /
        ``_device``): ``params`` is a [W, 13] host array or CUDA float64 tensor of
        points spanning the batches to come (e.g. prior draws), [13, W] with
        ``param_major``; ``None`` returns to automatic coverage."""
        if params is None:
            check(self._lib.gwin_plan_cover_host(self._h, 0, None, 0, 0))
            return
        if hasattr(params, "is_cuda"):
            import torch
            if not params.is_cuda or params.dtype != torch.float64:
                raise ValueError("params must be a CUDA float64 tensor")
            params = params.contiguous()
            W = params.shape[1] if param_major else params.shape[0]
            sw, si = (1, W) if param_major else (GWIN_NPARAM, 1)
            if stream is None:
                stream = torch.cuda.current_stream(params.device).cuda_stream
            check(self._lib.gwin_plan_cover_device(self._h, W, params.data_ptr(),
                                                   sw, si, stream))
            return
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim != 2 or params.shape[1 - int(param_major)] != GWIN_NPARAM:
            raise ValueError("params must be [W, %d]" % GWIN_NPARAM)
        W = params.shape[1] if param_major else params.shape[0]
        sw, si = (1, W) if param_major else (GWIN_NPARAM, 1)
        check(self._lib.gwin_plan_cover_host(self._h, W, params.ctypes.data, sw, si))

    def table_info(self):
        """dict(subblocks, references, bytes, k_split) of the current moment tables."""
        v = [C.c_int64() for _ in range(4)]
        check(self._lib.gwin_plan_table_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("subblocks", "references", "bytes", "k_split"),
                        (x.value for x in v)))

    def slow_path_walkers(self):
        """Walkers evaluated by the slow path (outside the block schedule's or the
        tables' coverage: right, but ~20x slower) since the last call."""
        n = C.c_int64()
        check(self._lib.gwin_plan_slow_path(self._h, C.byref(n)))
        return n.value

    def set_calibration(self, spline_freqs):
        """``spline_freqs`` [n_det, n_points]: ``CubicSpline.spline_points`` of
        every detector, plan detector order; ``None`` clears.  See
        ``gwin_plan_set_calibration``."""
        if spline_freqs is None:
            check(self._lib.gwin_plan_set_calibration(self._h, 0, None))
            self.n_cal = 0
            return
        sf = np.ascontiguousarray(spline_freqs, dtype=np.float64)
        if sf.ndim != 2 or sf.shape[0] != self.n_det:
            raise ValueError("spline_freqs must be [n_det, n_points]")
        check(self._lib.gwin_plan_set_calibration(self._h, sf.shape[1],
                                                  sf.ctypes.data))
        self.n_cal = sf.shape[1]

    @property
    def last_launches(self):
        return self._lib.gwin_plan_last_launches(self._h)

    def kernel_timing(self, enable=True):
        """(mean ms, launches) of the loglr kernel since the last call; see
        ``gwin_plan_kernel_timing``."""
        ms, n = C.c_double(), C.c_int()
        check(self._lib.gwin_plan_kernel_timing(self._h, int(bool(enable)),
                                                C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_timing_detail(self):
        """Mean ms of the launches inside the bracket ``kernel_timing`` last reported:
        dict(recurrence, table, table_fine, recurrence_tail)."""
        ms = np.zeros(4)
        check(self._lib.gwin_plan_kernel_timing_detail(self._h, _dptr(ms)))
        return dict(zip(("recurrence", "table", "table_fine", "recurrence_tail"), ms.tolist()))

    def table_blocks(self):
        """The moment tables' schedule: int32 arrays (start, m, references, fine_min)."""
        n = C.c_int64()
        check(self._lib.gwin_plan_table_blocks(self._h, 0, None, None, None, None, C.byref(n)))
        out = [np.zeros(n.value, dtype=np.int32) for _ in range(4)]
        check(self._lib.gwin_plan_table_blocks(self._h, n.value, *[a.ctypes.data for a in out], C.byref(n)))
        return tuple(out)

    def loglr_host(self, params, mode=MODE_GAUSSIAN, skip=None,
                   param_major=False, calib=None):
        """params: [W, 13] float64 host array, or [13, W] with
        ``param_major=True``; ``calib``: optional calibration values in the same
        orientation, [W, n_det * 2 * n_points] (or its transpose).  Returns
        loglr[W], hd[W, n_det] complex128, hh[W, n_det]."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if param_major:
            if params.ndim != 2 or params.shape[0] != GWIN_NPARAM:
                raise ValueError("params must be [%d, W]" % GWIN_NPARAM)
            W = params.shape[1]
            sw, si = 1, W
        else:
            if params.ndim != 2 or params.shape[1] != GWIN_NPARAM:
                raise ValueError("params must be [W, %d]" % GWIN_NPARAM)
            W = params.shape[0]
            sw, si = GWIN_NPARAM, 1
        loglr = np.empty(W)
        hd = np.empty((W, self.n_det), dtype=np.complex128)
        hh = np.empty((W, self.n_det))
        sk = None
        if skip is not None:
            sk = np.ascontiguousarray(skip, dtype=np.uint8)
            if sk.shape != (W,):
                raise ValueError("skip must be [W]")
        if calib is not None:
            nv = self.n_det * 2 * self.n_cal
            calib = np.ascontiguousarray(calib, dtype=np.float64)
            if calib.shape != ((nv, W) if param_major else (W, nv)) or nv == 0:
                raise ValueError("calib must be [W, %d] (transposed with "
                                 "param_major) after set_calibration" % nv)
            csw, csv = (1, W) if param_major else (nv, 1)
            check(self._lib.gwin_loglr_batch_calib_host(
                self._h, int(mode), W, params.ctypes.data, sw, si,
                calib.ctypes.data, csw, csv,
                sk.ctypes.data if sk is not None else None,
                loglr.ctypes.data, hd.ctypes.data, hh.ctypes.data))
            return loglr, hd, hh
        check(self._lib.gwin_loglr_batch_host_strided(
            self._h, int(mode), W, params.ctypes.data, sw, si,
            sk.ctypes.data if sk is not None else None,
            loglr.ctypes.data, hd.ctypes.data, hh.ctypes.data))
        return loglr, hd, hh

    def loglr_points_host(self, points, colmap, fixed, mode=MODE_GAUSSIAN, skip=None,
                          stats=True, col_major=False, defer=False):
        """``gwin_loglr_batch_points_host``: ``points`` is the [W, ndim] array of SAMPLED values
        (``[ndim, W]`` with ``col_major``), ``colmap[i]`` the column holding kernel parameter i
        (-1: ``fixed[i]``).  Page-locked ``points`` (``pinned_empty``) are copied from directly.
        ``stats=False`` leaves hd / hh on the device: returns (loglr, None, None).
        ``defer=True`` only queues the call (``gwin_loglr_batch_points_submit``) and returns a
        ``PendingCall``; its ``wait()`` gives the same tuple.  Two calls may be pending per plan."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        if points.ndim != 2:
            raise ValueError("points must be 2-D")
        ndim, W = points.shape if col_major else points.shape[::-1]
        sw, si = (1, W) if col_major else (ndim, 1)
        cm = np.ascontiguousarray(colmap, dtype=np.int32)
        fx = np.ascontiguousarray(fixed, dtype=np.float64)
        if cm.shape != (GWIN_NPARAM,) or fx.shape != (GWIN_NPARAM,):
            raise ValueError("colmap and fixed must have %d entries" % GWIN_NPARAM)
        loglr = np.empty(W)
        hd = hh = None
        if stats:
            hd = np.empty((W, self.n_det), dtype=np.complex128)
            hh = np.empty((W, self.n_det))
        sk = None
        if skip is not None:
            sk = np.ascontiguousarray(skip, dtype=np.uint8)
            if sk.shape != (W,):
                raise ValueError("skip must be [W]")
        args = (self._h, int(mode), W, points.ctypes.data, ndim, sw, si, cm.ctypes.data,
                fx.ctypes.data, sk.ctypes.data if sk is not None else None, loglr.ctypes.data,
                hd.ctypes.data if stats else None, hh.ctypes.data if stats else None)
        if defer:
            ticket = C.c_int(-1)
            check(self._lib.gwin_loglr_batch_points_submit(*(args + (C.byref(ticket),))))
            return PendingCall(self, ticket.value, (loglr, hd, hh), keep=(points, cm, fx, sk))
        check(self._lib.gwin_loglr_batch_points_host(*args))
        return loglr, hd, hh

    def loglr_device(self, params, mode=MODE_GAUSSIAN, skip=None, out=None,
                     stream=None, calib=None):
        """Device-resident variant: ``params`` is a CUDA float64 torch tensor
        [W, 13] (``calib``: optional [W, n_det * 2 * n_points]); returns torch
        tensors (loglr, hd[W,n_det,2], hh) on the same device.  Asynchronous on
        ``stream`` (default: torch's current stream).  torch is only the owner
        of the device memory here."""
        import torch
        if not params.is_cuda or params.dtype != torch.float64:
            raise ValueError("params must be a CUDA float64 tensor")
        params = params.contiguous()
        W = params.shape[0]
        if out is None:
            loglr = torch.empty(W, dtype=torch.float64, device=params.device)
            hd = torch.empty((W, self.n_det, 2), dtype=torch.float64,
                             device=params.device)
            hh = torch.empty((W, self.n_det), dtype=torch.float64,
                             device=params.device)
        else:
            loglr, hd, hh = out
        if stream is None:
            stream = torch.cuda.current_stream(params.device).cuda_stream
        sk = None
        if skip is not None:
            sk = skip.to(torch.uint8).contiguous()
        if calib is not None:
            nv = self.n_det * 2 * self.n_cal
            if (not calib.is_cuda or calib.dtype != torch.float64
                    or tuple(calib.shape) != (W, nv) or nv == 0):
                raise ValueError("calib must be a CUDA float64 tensor [W, %d] "
                                 "after set_calibration" % nv)
            calib = calib.contiguous()
            check(self._lib.gwin_loglr_batch_calib(
                self._h, int(mode), W, params.data_ptr(), GWIN_NPARAM, 1,
                calib.data_ptr(), nv, 1,
                sk.data_ptr() if sk is not None else None,
                loglr.data_ptr(), hd.data_ptr(), hh.data_ptr(), stream))
            return loglr, hd, hh
        check(self._lib.gwin_loglr_batch(
            self._h, int(mode), W, params.data_ptr(),
            sk.data_ptr() if sk is not None else None,
            loglr.data_ptr(), hd.data_ptr(), hh.data_ptr(), stream))
        return loglr, hd, hh

    def generate(self, params, calib=None):
        """One detector-frame waveform set: returns (h[n_det, n_f] complex128,
        length).  ``calib``: optional [n_det * 2 * n_points] calibration values."""
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(GWIN_NPARAM)
        h = np.zeros((self.n_det, self.n_f), dtype=np.complex128)
        n = C.c_int64()
        cv = None
        if calib is not None:
            cv = np.ascontiguousarray(calib, dtype=np.float64).reshape(
                self.n_det * 2 * self.n_cal)
        check(self._lib.gwin_generate_calib_host(
            self._h, _dptr(p), cv.ctypes.data if cv is not None else None,
            _dptr(h.view(np.float64)), C.byref(n)))
        return h, n.value


def fp64_peak(device=0):
    """Measured FP64 FMA lane-ops/s of the device (independent DFMA chains)."""
    r = C.c_double()
    check(load().gwin_fp64_peak(int(device), C.byref(r)))
    return r.value


def fp64_probe(device=0, variant=0):
    """Diagnostic FP64 rate probes, see ``gwin_fp64_probe``."""
    r = C.c_double()
    check(load().gwin_fp64_probe(int(device), int(variant), C.byref(r)))
    return r.value


def version():
    return load().gwin_version().decode()
