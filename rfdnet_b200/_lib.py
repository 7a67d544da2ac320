"""ctypes binding of librfdnet_b200.so (the C ABI declared in include/rfdnet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librfdnet_b200.so")

_vp, _i, _f, _ll, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES); must match include/rfdnet_b200.h
SIGNATURES = {
    "rfd_abi_version": [],
    "rfd_status_string": [_i],
    "rfd_last_error": [],
    "rfd_device_info": [_vp, _vp, _vp],
    "rfd_launch_count": [],
    "rfd_furthest_point_sampling": [_vp, _i, _i, _i, _vp, _vp],
    "rfd_furthest_point_sampling_xyz": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "rfd_fps_prefix_check": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "rfd_furthest_point_sampling_cond": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "rfd_gather_points": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "rfd_gather_points_grad": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "rfd_ball_query": [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp],
    "rfd_group_points": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "rfd_group_points_grad": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "rfd_query_and_group": [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp],
    "rfd_query_and_group_rotated": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp],
    "rfd_stn_apply": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "rfd_three_nn": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "rfd_three_interpolate": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "rfd_three_interpolate_grad": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "rfd_three_nn_interpolate": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "rfd_pointwise_mlp_f32": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "rfd_make_3d_grid": [_i, _f, _vp, _vp],
    "rfd_occupancy_bits": [_vp, _i, _i, _f, _vp, _vp, _vp],
    "rfd_extract_mesh": [_vp, _i, _i, ctypes.c_double, ctypes.c_double, _vp, _i, _vp, _ll, _ll, _vp, _vp, _vp],
    "rfd_mlp_chain_packed_bytes": [_i, _i, _i, _i, _i, _i],
    "rfd_mlp_chain_pack": [_i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp],
    "rfd_mlp_chain": [_i, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "rfd_mlp_chain_ex": [_i, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _vp],
    "rfd_mlp_chain_rows": [_i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp],
    "rfd_sa_mlp_chain": [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _i, _i, _i, _vp, _vp, _vp],
    "rfd_transpose_features": [_vp, _i, _i, _i, _i, _vp, _vp],
    "rfd_onet_packed_bytes": [_i],
    "rfd_onet_pack_weights": [_vp, _i, _vp, _vp],
    "rfd_onet_aff_floats": [],
    "rfd_onet_cbn_tables": [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp],
    "rfd_onet_decode": [_vp, _ll, _i, _i, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp],
    "rfd_onet_decode_f32": [_vp, _ll, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _sz, _vp],
    "rfd_onet_decode_set_cluster": [_i],
    "rfd_onet_decode_traced": [_vp, _ll, _i, _i, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp],
    "rfd_umma_selftest": [_vp, _vp, _vp, _vp],
    "rfd_umma_selftest_ts": [_vp, _vp, _vp, _vp],
}
_RESTYPES = {"rfd_status_string": ctypes.c_char_p, "rfd_last_error": ctypes.c_char_p,
             "rfd_launch_count": _ll, "rfd_onet_packed_bytes": _sz, "rfd_onet_aff_floats": _sz, "rfd_mlp_chain_packed_bytes": _sz}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(rfdnet_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.rfd_status_string(status).decode()
        detail = lib.rfd_last_error().decode()
        raise RuntimeError(f"{what} failed: {msg}" + (f" [{detail}]" if detail else ""))


def launch_count():
    return int(load().rfd_launch_count())


# ---------------------------------------------------------------------------------------------------
# optional per-call device timing (bench.py): CUDA events recorded on the launching stream around a call.
TIMERS = None  # None = off; list = collect (name, start_event, end_event, work) tuples


class timed:
    """`with timed("name", work): launch(...)` -- no-op unless TIMERS is a list.  Never synchronises."""

    def __init__(self, name, work=0.0):
        self.name, self.work = name, work

    def __enter__(self):
        if TIMERS is not None:
            import torch
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if TIMERS is not None:
            self.e.record()
            TIMERS.append((self.name, self.s, self.e, self.work))
        return False
