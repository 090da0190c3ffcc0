"""ctypes binding of the C-ABI in include/cda_b200.h (libcda_b200.so, built in-tree by nvcc).

There is NO CPU fallback: if the shared library is missing or was not built for this machine
the import of the env fails loudly.  `build()` cross-compiles on a box without a GPU.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("CDA_B200_LIB") or os.path.join(CSRC, "libcda_b200.so")   # CDA_B200_LIB: a variant build (tools/variant_bench.py)
SOURCES = ("cda_b200.cu", "cda_kernels.cuh", "cda_zig_tables.cuh", "cda_dec128.cuh", "cda_twin.cuh")
HEADER = os.path.join(_ROOT, "include", "cda_b200.h")
TESTING_HEADER = os.path.join(_ROOT, "include", "cda_b200_testing.h")   # test / measurement entries (not product ABI)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
]


class CdaConfig(ctypes.Structure):
    _fields_ = [
        ("num_agents", ctypes.c_int32), ("n_hist", ctypes.c_int32), ("max_step", ctypes.c_int32),
        ("tick_size", ctypes.c_int32), ("init_cash", ctypes.c_int64), ("min_size", ctypes.c_int32),
        ("mkt_max_size", ctypes.c_int32), ("limit_size_multiple", ctypes.c_int32),
        ("initial_price_min", ctypes.c_int32), ("initial_price_max", ctypes.c_int32),
        ("order_capacity", ctypes.c_int32), ("fill_capacity", ctypes.c_int32),
        ("order_penalty", ctypes.c_double), ("trade_penalty", ctypes.c_double),
        ("drawdown_penalty", ctypes.c_double), ("passive_bonus", ctypes.c_double),
        ("loss_multiplier", ctypes.c_double), ("decimal_ledger", ctypes.c_int32), ("fill_tape", ctypes.c_int32),
    ]


INFO_FIELDS = (
    "cash", "cash_on_hold", "cost_basis", "nav", "prev_nav", "max_nav", "net_position",
    "position_val", "num_trades", "num_trades_step", "num_passive_fills_step",
    "order_step_placed", "num_rejected_step", "is_pass_action", "market",
)
INFO_MARKET_COLS = ("last_price", "best_bid", "best_ask", "time", "next_order_id", "t_step",
                    "done_mask", "status")

EXPORTS = (
    "cda_create", "cda_destroy", "cda_reset", "cda_step", "cda_step_host", "cda_step_host_ring", "cda_reset_host_ring", "cda_step_host_window", "cda_reset_host_window", "cda_window_bind", "cda_step_window", "cda_step_planes", "cda_reset_planes", "cda_serve_bind", "cda_serve_step", "cda_serve_stop", "cda_serve_launches", "cda_rollout_random",
    "cda_gather_create", "cda_gather_connect", "cda_gather_publish", "cda_step_gather", "cda_gather_wait", "cda_gather_pos", "cda_gather_row_words", "cda_gather_record_parity", "cda_get_info", "cda_get_info_all", "cda_get_fills", "cda_set_action_log", "cda_dump_market", "cda_state_bytes", "cda_save_state",
    "cda_load_state", "cda_num_markets", "cda_record_bytes", "cda_obs_dim", "cda_order_capacity",
    "cda_kernel_launches", "cda_strerror", "cda_last_cuda_error", "cda_build_info",
    "cda_seed_to_pcg64", "cda_state_layout", "cda_twin_sync", "cda_status_flag", "cda_status_flag_clear",
)
TESTING_EXPORTS = ("cda_debug_phase_buffer", "cda_debug_dec_op", "cda_debug_dec_op_device", "cda_debug_set_window_mode", "cda_debug_restart_count", "cda_debug_serve_timeline")


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [HEADER, TESTING_HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -> csrc/libcda_b200.so (sm_100a).  Works without a GPU (cross-compile)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcda_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(_ROOT, "include"), "-I", CSRC,
                                 "-o", SO_PATH, os.path.join(CSRC, "cda_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return SO_PATH


_lib = None


def lib():
    """Load the shared library (fails loudly when it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the env step)")
    L = ctypes.CDLL(SO_PATH)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
    variant = bool(os.environ.get("CDA_B200_LIB"))   # a variant / older build may lack newer entry points (tools/variant_bench.py)

    def sig(name, argtypes=None, restype=None):
        if variant and not hasattr(L, name):
            return
        f = getattr(L, name)
        if argtypes is not None:
            f.argtypes = argtypes
        if restype is not None:
            f.restype = restype

    sig("cda_create", [ctypes.POINTER(CdaConfig), i32, i32, ctypes.POINTER(vp)])
    sig("cda_destroy", [vp])
    sig("cda_reset", [vp, vp, vp, vp, vp])
    sig("cda_step", [vp] * 11)
    sig("cda_step_host", [vp] * 11)
    sig("cda_step_host_ring", [vp] * 10 + [i64, vp])
    sig("cda_reset_host_ring", [vp, vp, vp, vp, vp])
    sig("cda_step_host_window", [vp] * 7 + [i32, i32, vp, i32, vp])
    sig("cda_reset_host_window", [vp, vp, vp, vp, i32, vp])
    sig("cda_window_bind", [vp, vp, i32, vp, vp])
    sig("cda_step_window", [vp, vp, i32, i32])
    sig("cda_step_planes", [vp, vp, vp, i32, i32, vp])
    sig("cda_reset_planes", [vp, vp, vp, vp, i32, i32, i32, vp])
    sig("cda_serve_bind", [vp, vp, i32, i32])
    sig("cda_serve_step", [vp, vp, i32, vp])
    sig("cda_serve_stop", [vp])
    sig("cda_serve_launches", [vp], i64)
    sig("cda_rollout_random", [vp, i32, u64, vp, vp, vp, vp, vp])
    sig("cda_gather_create", [vp, i32, i32, vp, ctypes.POINTER(vp), ctypes.POINTER(u64)])
    sig("cda_gather_connect", [vp, vp])
    sig("cda_step_gather", [vp, vp, vp, vp, vp, vp, vp])
    sig("cda_get_info", [vp, i32, vp, vp])
    sig("cda_get_info_all", [vp, vp, vp])
    sig("cda_get_fills", [vp, vp, vp, vp])
    sig("cda_set_action_log", [vp, vp])
    sig("cda_dump_market", [vp, i32, vp, vp, vp, vp, i32, vp, vp])
    sig("cda_state_bytes", [vp], ctypes.c_size_t)
    sig("cda_save_state", [vp, vp, vp])
    sig("cda_load_state", [vp, vp, vp])
    for name in ("cda_num_markets", "cda_obs_dim", "cda_order_capacity", "cda_record_bytes"):
        sig(name, [vp], i32)
    sig("cda_kernel_launches", [vp], i64)
    for name in ("cda_last_cuda_error", "cda_build_info"):
        sig(name, None, ctypes.c_char_p)
    sig("cda_strerror", [i32], ctypes.c_char_p)
    sig("cda_seed_to_pcg64", [u64, ctypes.POINTER(u64)])
    sig("cda_state_layout", [vp, ctypes.POINTER(i32)])
    sig("cda_twin_sync", [vp, vp, vp])
    sig("cda_gather_publish", [vp, vp])
    sig("cda_gather_wait", [vp, vp])
    sig("cda_gather_pos", [vp], i32)
    sig("cda_gather_row_words", [vp], i32)
    sig("cda_gather_record_parity", [vp], i32)
    sig("cda_status_flag", [vp], ctypes.POINTER(ctypes.c_uint32))
    sig("cda_status_flag_clear", [vp])
    sig("cda_debug_set_window_mode", [i32], None)
    sig("cda_debug_restart_count", None, i64)
    sig("cda_debug_phase_buffer", None, ctypes.c_void_p)
    sig("cda_debug_serve_timeline", [i32], ctypes.c_void_p)
    sig("cda_debug_dec_op", [i32, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i32, ctypes.POINTER(i32)])
    sig("cda_debug_dec_op_device", [i32, i32, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p, i32, ctypes.POINTER(i32)])
    _lib = L
    return L


class CdaError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        L = lib()
        msg = L.cda_strerror(rc).decode()
        if rc == -2:
            msg += ": " + L.cda_last_cuda_error().decode()
        if rc == -1:
            raise ValueError("cda_b200: " + msg)
        raise CdaError("cda_b200: " + msg)
