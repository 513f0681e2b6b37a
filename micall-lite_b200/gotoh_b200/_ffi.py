"""ctypes binding of libgotoh_b200.so (include/gotoh_b200.h).

The default loader opens exactly one file, ``micall-lite_b200/lib/libgotoh_b200.so`` (built by
``__graft_entry__.build()`` / ``csrc/build.py`` with nvcc for sm_100a).  There is no CPU
fallback: if the library is missing the import of the public API fails, and if no CUDA device
is visible every compute call raises ``GotohError(ENODEVICE)``.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(os.path.dirname(_PKG), "lib", "libgotoh_b200.so")

OK, EINVAL, EEMPTY, EDOMAIN, ESENTINEL, ERANGE, ENODEVICE, ECUDA, ENOMEM, ETRACEBACK, ECAPACITY = 0, -1, -2, -3, -4, -5, -6, -7, -8, -9, -10
NT, HIV25, AA_RB = 0, 1, 2

# every symbol include/gotoh_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "gotoh_b200_version", "gotoh_b200_last_error", "gotoh_b200_device_count",
    "gotoh_b200_pairscore_table", "gotoh_b200_align_batch", "gotoh_b200_plan_create",
    "gotoh_b200_plan_run", "gotoh_b200_plan_fetch", "gotoh_b200_plan_destroy",
    "gotoh_b200_plan_stat", "gotoh_b200_host_alloc", "gotoh_b200_host_free", "gotoh_b200_int_peak",
    "gotoh_b200_release_cache", "gotoh_b200_gotoh2_align_batch", "gotoh_b200_gotoh2_last_stats", "gotoh_b200_edit_distance_batch",
    "gotoh_b200_align_batch_tight", "gotoh_b200_align_batch_compact", "gotoh_b200_d2h_probe",
)


class GotohError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libgotoh_b200 error %d: %s" % (code, message))
        self.code = code


class GotohInputError(GotohError, ValueError):
    """Input outside the reference's defined domain."""


class GotohCapacityError(GotohError):
    """Tight / compact result forms: the caller's result buffer was too small (retry with the worst-case bound)."""


_vp, _i32, _i64, _u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32


class Library:
    def __init__(self, path=None):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise ImportError(
                "%s not found: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
                "gotoh_b200 has no CPU fallback." % self.path)
        lib = ctypes.CDLL(self.path)
        lib.gotoh_b200_version.restype = _i32
        lib.gotoh_b200_last_error.restype = ctypes.c_char_p
        lib.gotoh_b200_device_count.restype = _i32
        lib.gotoh_b200_pairscore_table.restype = _i32
        lib.gotoh_b200_pairscore_table.argtypes = [_i32, _vp]
        lib.gotoh_b200_align_batch.restype = _i32
        lib.gotoh_b200_align_batch.argtypes = [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32,
                                               _vp, _vp, _vp, _vp, _vp, _u32]
        lib.gotoh_b200_align_batch_tight.restype = _i32
        lib.gotoh_b200_align_batch_tight.argtypes = [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32,
                                                     _vp, _vp, _i64, _vp, _vp, _vp, _u32]
        lib.gotoh_b200_align_batch_compact.restype = _i32
        lib.gotoh_b200_align_batch_compact.argtypes = [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32,
                                                       _vp, _vp, _i64, _vp, _u32]
        lib.gotoh_b200_d2h_probe.restype = _i32
        lib.gotoh_b200_d2h_probe.argtypes = [_i32, _vp, _i64, _i32, ctypes.POINTER(ctypes.c_double)]
        lib.gotoh_b200_plan_create.restype = _i32
        lib.gotoh_b200_plan_create.argtypes = [_i32, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32,
                                               _vp, ctypes.POINTER(_vp)]
        lib.gotoh_b200_plan_run.restype = _i32
        lib.gotoh_b200_plan_run.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
        lib.gotoh_b200_plan_fetch.restype = _i32
        lib.gotoh_b200_plan_fetch.argtypes = [_vp, _vp, _vp, _vp, _vp]
        lib.gotoh_b200_plan_destroy.restype = None
        lib.gotoh_b200_plan_destroy.argtypes = [_vp]
        lib.gotoh_b200_plan_stat.restype = _i64
        lib.gotoh_b200_plan_stat.argtypes = [_vp, _i32]
        lib.gotoh_b200_host_alloc.restype = _vp
        lib.gotoh_b200_host_alloc.argtypes = [_i64]
        lib.gotoh_b200_host_free.restype = None
        lib.gotoh_b200_host_free.argtypes = [_vp]
        lib.gotoh_b200_gotoh2_align_batch.restype = _i32
        lib.gotoh_b200_gotoh2_align_batch.argtypes = [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                                      ctypes.c_char_p, _vp, _vp, _vp, _vp, _vp, _vp, _i32]
        lib.gotoh_b200_edit_distance_batch.restype = _i32
        lib.gotoh_b200_edit_distance_batch.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp, _i32]
        lib.gotoh_b200_gotoh2_last_stats.restype = _i32
        lib.gotoh_b200_gotoh2_last_stats.argtypes = [_vp, _i32]
        lib.gotoh_b200_release_cache.restype = None
        lib.gotoh_b200_release_cache.argtypes = []
        lib.gotoh_b200_int_peak.restype = _i32
        lib.gotoh_b200_int_peak.argtypes = [_i32, _i32, ctypes.POINTER(ctypes.c_double)]
        self.lib = lib

    def check(self, rc):
        if rc == OK:
            return
        msg = (self.lib.gotoh_b200_last_error() or b"").decode("utf-8", "replace")
        if rc in (EEMPTY, EDOMAIN, ESENTINEL, ERANGE):
            raise GotohInputError(rc, msg)
        if rc == ECAPACITY:
            raise GotohCapacityError(rc, msg)
        raise GotohError(rc, msg)

    def device_count(self):
        return int(self.lib.gotoh_b200_device_count())

    def version(self):
        return int(self.lib.gotoh_b200_version())


_default = None


def default_library():
    """The one product library.  Raises ImportError when it has not been built."""
    global _default
    if _default is None:
        _default = Library()
    return _default
