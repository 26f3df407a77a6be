"""ctypes binding of libeaqhm_b200.so (the C ABI declared in include/eaqhm_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
visible, the product path raises.  ``build()`` compiles the library in-tree with
nvcc for sm_100a (works without a GPU).
"""
import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# EAQHM_LIB: load another build of the same sources (kernel experiments: tools/build_variant.sh); default in-tree
LIB_PATH = os.environ.get("EAQHM_LIB") or os.path.join(_HERE, "libeaqhm_b200.so")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class Tables(C.Structure):
    """eaqhm_tables (include/eaqhm_b200.h)."""
    _fields_ = [("T", C.c_int64), ("maxlen", C.c_int64), ("G", C.c_int32), ("U", C.c_int32),
                ("Kmax", C.c_int32), ("step", C.c_int32), ("fs", C.c_double),
                ("pos", C.c_void_p), ("own", C.c_void_p), ("utt", C.c_void_p), ("N", C.c_void_p),
                ("K0", C.c_void_p), ("f0", C.c_void_p), ("ustart", C.c_void_p), ("ulen", C.c_void_p),
                ("gfirst", C.c_void_p), ("gcount", C.c_void_p), ("f0stale", C.c_void_p),
                ("f0min", C.c_void_p), ("active", C.c_void_p), ("stat_lo", C.c_void_p), ("stat_hi", C.c_void_p),
                ("xoff", C.c_void_p), ("alist", C.c_void_p), ("nA", C.c_int32), ("flist", C.c_void_p),
                ("nF", C.c_int32)]


class Tracks(C.Structure):
    """eaqhm_tracks."""
    _fields_ = [("am", C.c_void_p), ("ph", C.c_void_p), ("fmc", C.c_void_p), ("zc", C.c_void_p), ("zs", C.c_void_p)]


class Knots(C.Structure):
    """eaqhm_knots."""
    _fields_ = [("am", C.c_void_p), ("ph", C.c_void_p), ("fm", C.c_void_p), ("a0", C.c_void_p)]


class EaqhmError(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force=False, verbose=False):
    """Compile every CUDA source of the package into libeaqhm_b200.so for sm_100a."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(_HERE), "include", "eaqhm_b200.h")]
    if not force and os.path.exists(LIB_PATH) and \
            all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise EaqhmError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None

_SIGNATURES = {
    "eaqhm_version": (C.c_int, []),
    "eaqhm_last_error": (C.c_char_p, []),
    "eaqhm_device_count": (C.c_int, []),
    "eaqhm_profile_enable": (C.c_int, [C.c_int]),
    "eaqhm_profile_reset": (C.c_int, []),
    "eaqhm_profile_classes": (C.c_int, []),
    "eaqhm_profile_name": (C.c_char_p, [C.c_int]),
    "eaqhm_profile_read": (C.c_int, [c_f64p, c_i64p, C.c_int32]),
    "eaqhm_analysis_scratch_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "eaqhm_analysis_pipeline_sets": (C.c_int, []),
    "eaqhm_analysis": (C.c_int, [C.POINTER(Tables), C.c_void_p, C.POINTER(Tracks), C.POINTER(Knots), C.c_void_p,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "eaqhm_interp_work_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "eaqhm_interpolate": (C.c_int, [C.POINTER(Tables), C.POINTER(Knots), C.POINTER(Tracks), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_void_p]),
    "eaqhm_synth_work_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "eaqhm_synthesize": (C.c_int, [C.POINTER(Tables), C.POINTER(Tracks), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "eaqhm_snapshot": (C.c_int, [C.POINTER(Tables), C.c_void_p, C.POINTER(Tracks), C.POINTER(Knots), C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Knots), C.c_void_p]),
    "eaqhm_iqhmLS_complexamps": (C.c_int, [c_f64p, C.c_int32, c_f64p, C.c_int32, c_f64p, C.c_double, c_f64p, c_f64p]),
    "eaqhm_eaqhmLS_complexamps": (C.c_int, [c_f64p, C.c_int32, c_f64p, c_f64p, C.c_int32, c_f64p, C.c_double,
                                            c_f64p, c_f64p]),
    "eaqhm_phase_integr_interpolation": (C.c_int, [c_f64p, c_f64p, C.c_int64, c_i64p, C.c_int32, c_f64p]),
    "eaqhm_filtfilt_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eaqhm_vuv_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "eaqhm_swipe_loudness": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_double,
                                       C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                       C.c_void_p]),
    "eaqhm_swipe_strength": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "eaqhm_swipe_accumulate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "eaqhm_swipe_argmax": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """The loaded library (raises EaqhmError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EaqhmError("libeaqhm_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; "
                             "g.build()'`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise EaqhmError("eaqhm error %d: %s" % (rc, lib().eaqhm_last_error().decode()))


def profile_read():
    """{kernel class: (milliseconds, launches)} accumulated since the last profile_reset()."""
    import numpy as np
    L = lib()
    n = L.eaqhm_profile_classes()
    ms = np.zeros(n)
    cnt = np.zeros(n, np.int64)
    check(L.eaqhm_profile_read(ms.ctypes.data_as(c_f64p), cnt.ctypes.data_as(c_i64p), n))
    return {L.eaqhm_profile_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
