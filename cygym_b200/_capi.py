"""ctypes binding of libcygym_b200.so (the C-ABI of include/cygym_b200.h).

There is no CPU path: importing this module without the compiled CUDA library raises, and every
entry point needs a CUDA device.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

from . import draw_tables as DT

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CYGYM_B200_LIB") or os.path.join(_HERE, "libcygym_b200.so")  # override: profiling builds only
_SRC = os.path.join(_HERE, "csrc", "cyg_kernels.cu")
_DEPS = [_SRC, os.path.join(_HERE, "csrc", "cyg_core.cuh"), os.path.join(_HERE, "csrc", "cyg_tables.h"),
         os.path.join(os.path.dirname(_HERE), "include", "cygym_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--split-compile", "0"]  # split-compile: ptxas of the kernels in parallel

NSCAL = 16
ATYPE_NONE = 0x80
MODE_DEFENDER, MODE_ATTACKER = 0, 1
STEP_GROUPED, STEP_SKIP_WORK = 1, 2
BASE_LINES = {"Nash": 0, "No Defense": 1, "Preset": 2, "No Attack": 3}
E_INVAL, E_NOMEM, E_CUDA, E_STATE = -22, -12, -5, -71

# scalar slots (CYG_S_*)
(S_STEP, S_EPOCH, S_FLAGS, S_PREV_X, S_DEF_STEP, S_ATT_STEP, S_LOGS, S_COMPCNT, S_WORK, S_DEFCOST,
 S_CLEANCOST, S_SCAN, S_REVERT, S_CKPT, S_EBLK, S_EADD) = range(16)
FL_ERR_MASK = 0xF0


class CygConfig(C.Structure):
    """struct cyg_config"""
    _fields_ = [
        ("M", C.c_int32), ("E", C.c_int32), ("X", C.c_int32), ("n_exploits", C.c_int32), ("xcap", C.c_int32),
        ("num_of_device", C.c_int32), ("min_network_size", C.c_int32), ("evolve_period", C.c_int32),
        ("wl_period_base", C.c_int32), ("wl_period_max", C.c_int32), ("wl_cap", C.c_int32),
        ("scaling_vulnerability", C.c_int32), ("turbo", C.c_int32), ("zero_day", C.c_int32),
        ("zero_day_mask", C.c_uint32), ("att_space_n", C.c_int32), ("def_space_n", C.c_int32),
        ("default_high", C.c_int32), ("n_app_ids", C.c_int32), ("base_line", C.c_int32), ("tri_high", C.c_int32),
        ("reserved0", C.c_int32),
        ("work_scale", C.c_float), ("comp_scale", C.c_float), ("def_scale", C.c_float), ("gamma", C.c_float),
        ("thr_p_add", C.c_uint64), ("thr_p_attacker", C.c_uint64),
        ("poisson_tab", C.c_uint32 * 16), ("tri_tab", C.c_uint32 * 8), ("seed", C.c_uint64),
        ("turbo_frac_clients", C.c_double), ("turbo_frac_servers", C.c_double), ("turbo_max_clients", C.c_int32),
        ("turbo_max_servers", C.c_int32), ("turbo_ramp_steps", C.c_int32), ("reserved1", C.c_int32),
    ]


class CygNetwork(C.Structure):
    """struct cyg_network (host pointers)"""
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("mult", C.c_void_p), ("dev_static", C.c_void_p),
                ("os_val", C.c_void_p), ("ver_val", C.c_void_p)]


class CygState(C.Structure):
    """struct cyg_state (device pointers, canonical layout)"""
    _fields_ = [("dev", C.c_void_p), ("ckpt", C.c_void_p), ("blocked", C.c_void_p), ("extra", C.c_void_p),
                ("scal", C.c_void_p)]


class CygActions(C.Structure):
    """struct cyg_actions"""
    _fields_ = [("hdr", C.c_void_p), ("mask", C.c_void_p), ("order", C.c_void_p), ("order_stride", C.c_int32),
                ("n_groups", C.c_int32)]


class CygStepOut(C.Structure):
    """struct cyg_step_out"""
    _fields_ = [("raw_reward", C.c_void_p), ("shaped_reward", C.c_void_p), ("done", C.c_void_p),
                ("pre_masks", C.c_void_p), ("obs", C.c_void_p), ("obs_mode", C.c_int32)]


EXPORTS = ["cyg_version", "cyg_last_error", "cyg_create", "cyg_destroy", "cyg_set_base_line", "cyg_set_base_line_per_env", "cyg_internal_words",
           "cyg_bind", "cyg_import_state", "cyg_export_state", "cyg_step", "cyg_randomize", "cyg_sample_actions",
           "cyg_observe", "cyg_launch_count", "cyg_set_debug_cycles"]


class CygError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cygym_b200 error {code}: {msg}")
        self.code = code


def build(force=False, verbose=False):
    """Compile cygym_b200/csrc/cyg_kernels.cu for sm_100a into cygym_b200/libcygym_b200.so."""
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in _DEPS):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, _SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(cygym_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.cyg_version.restype = C.c_int
        L.cyg_last_error.restype = C.c_char_p
        L.cyg_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(CygConfig), C.POINTER(CygNetwork), C.c_int32, C.c_int32, C.c_int32]
        L.cyg_destroy.argtypes = [C.c_void_p]
        L.cyg_set_base_line.argtypes = [C.c_void_p, C.c_int32]
        L.cyg_set_base_line_per_env.argtypes = [C.c_void_p, C.c_void_p]
        L.cyg_internal_words.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.cyg_bind.argtypes = [C.c_void_p, C.c_void_p]
        L.cyg_import_state.argtypes = [C.c_void_p, C.POINTER(CygState), C.c_void_p]
        L.cyg_export_state.argtypes = [C.c_void_p, C.POINTER(CygState), C.c_void_p]
        L.cyg_step.argtypes = [C.c_void_p, C.POINTER(CygActions), C.c_uint32, C.POINTER(CygStepOut), C.c_void_p]
        L.cyg_randomize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.cyg_sample_actions.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cyg_observe.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.cyg_launch_count.argtypes = [C.c_void_p]
        L.cyg_launch_count.restype = C.c_int64
        L.cyg_set_debug_cycles.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise CygError(rc, lib().cyg_last_error().decode(errors="replace"))


def make_config(cfg, E, seed=0, xcap=16, base_line="Nash", tri_mode=2, tri_high=5):
    """cfg: the attribute dict of a Network (network.py) -> struct cyg_config."""
    c = CygConfig()
    c.M, c.E, c.X, c.n_exploits, c.xcap = cfg["M"], E, cfg["X"], cfg["n_exploits"], xcap
    c.num_of_device, c.min_network_size = cfg["numOfDevice"], cfg["Min_network_size"]
    c.evolve_period = cfg["evolve_period"]
    c.wl_period_base, c.wl_period_max, c.wl_cap = cfg["workload_period_base"], cfg["workload_period_max"], cfg["workload_cap"]
    c.scaling_vulnerability, c.turbo, c.zero_day = cfg["scaling_vulnerability"], cfg["turbo"], cfg["zero_day"]
    c.zero_day_mask = cfg["zero_day_mask"]
    c.att_space_n, c.def_space_n, c.default_high = cfg["att_space_n"], cfg["def_space_n"], cfg["default_high"]
    c.n_app_ids = cfg.get("n_app_ids", 0)
    c.base_line = BASE_LINES.get(base_line, 4)
    c.tri_high = tri_high
    c.work_scale, c.comp_scale, c.def_scale, c.gamma = cfg["work_scale"], cfg["comp_scale"], cfg["def_scale"], cfg["gamma"]
    c.thr_p_add = DT.bernoulli_threshold(cfg["p_add"])
    c.thr_p_attacker = DT.bernoulli_threshold(cfg["p_attacker"])
    for i, t in enumerate(DT.poisson_table(cfg["lambda_events"])):
        c.poisson_tab[i] = t
    for i, t in enumerate(DT.triangular_ceil_table(tri_mode, tri_high)):
        c.tri_tab[i] = t
    c.seed = seed
    c.turbo_frac_clients = cfg.get("turbo_fraction_clients", 0.05)
    c.turbo_frac_servers = cfg.get("turbo_fraction_servers", 0.02)
    c.turbo_max_clients = cfg.get("turbo_max_clients", 200)
    c.turbo_max_servers = cfg.get("turbo_max_servers", 40)
    c.turbo_ramp_steps = cfg.get("turbo_ramp_steps", 200)
    return c
