"""ctypes binding of libadb200.so (the C-ABI declared in include/adb200.h) and its build recipe.

The library is the only compute path: there is no CPU or PyTorch fallback. `lib()` raises
`RuntimeError` if the shared object is missing or cannot be loaded, and every wrapper in
`ops.py` raises if a call returns a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

_PKG = Path(__file__).resolve().parent
CSRC = _PKG / "csrc"
LIB_DIR = _PKG / "lib"
# ADB_LIB_PATH: load another build of the library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = Path(os.environ["ADB_LIB_PATH"]) if os.environ.get("ADB_LIB_PATH") else LIB_DIR / "libadb200.so"
HEADER = _PKG.parent / "include" / "adb200.h"

SOURCES = ["host.cu", "conv_igemm.cu", "attention.cu", "attention2.cu", "groupnorm.cu", "elementwise.cu", "moments.cu",
           "attention_bwd.cu", "attention_bwd_fused.cu", "backward.cu", "attention_sd.cu", "sd_ops.cu", "inception_ops.cu"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Wno-deprecated-gpu-targets",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", HEADER]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into autodiffusion_b200/lib/libadb200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


# ---- C structs (mirror include/adb200.h) -------------------------------------------------


class ConvSeg(C.Structure):
    _fields_ = [("act", C.c_void_p), ("cin", C.c_int), ("taps", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("h", C.c_int),
        ("w", C.c_int),
        ("cout", C.c_int),
        ("cout_pad", C.c_int),
        ("nseg", C.c_int),
        ("seg", ConvSeg * 3),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("res_mode", C.c_int),
        ("out", C.c_void_p),
        ("out_mode", C.c_int),
        ("stats_out", C.c_void_p),
        ("stats2_out", C.c_void_p),
        ("stats2_cpg", C.c_int),
        ("stats2_choff", C.c_int),
        ("bias_stride", C.c_int),
        ("seg_stride", C.c_int * 3),
        ("gnb_x", C.c_void_p),
        ("gnb_stats", C.c_void_p),
        ("gnb_gamma", C.c_void_p),
        ("gnb_beta", C.c_void_p),
        ("gnb_scale_shift", C.c_void_p),
        ("gnb_ss_stride", C.c_int),
        ("gnb_eps", C.c_float),
        ("gnb_silu", C.c_int),
        ("gnb_bstats", C.c_void_p),
    ]


class AttnSdDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_width", C.c_int), ("q_col0", C.c_int),
        ("kv", C.c_void_p), ("kv_width", C.c_int), ("k_col0", C.c_int), ("v_col0", C.c_int),
        ("out", C.c_void_p),
        ("b", C.c_int), ("heads", C.c_int), ("d_head", C.c_int), ("d_pad", C.c_int),
        ("tq", C.c_int), ("tk_rows", C.c_int), ("tk_valid", C.c_int), ("v_ones", C.c_int),
    ]


class GnDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("h", C.c_int),
        ("w", C.c_int),
        ("src0", C.c_void_p),
        ("c0", C.c_int),
        ("src1", C.c_void_p),
        ("c1", C.c_int),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("eps", C.c_float),
        ("scale_shift", C.c_void_p),
        ("ss_stride", C.c_int),
        ("silu", C.c_int),
        ("resample", C.c_int),
        ("out", C.c_void_p),
        ("stats", C.c_void_p),
        ("stats_ready", C.c_int),
    ]


class GnBwdDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("h", C.c_int),
        ("w", C.c_int),
        ("c", C.c_int),
        ("x", C.c_void_p),
        ("stats", C.c_void_p),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("eps", C.c_float),
        ("scale_shift", C.c_void_p),
        ("ss_stride", C.c_int),
        ("silu", C.c_int),
        ("resample", C.c_int),
        ("dout", C.c_void_p),
        ("add", C.c_void_p),
        ("add_mode", C.c_int),
        ("dx", C.c_void_p),
        ("bstats", C.c_void_p),
        ("bstats_ready", C.c_int),
    ]


RES_NONE, RES_SAME, RES_AVGPOOL2, RES_NEAREST2 = 0, 1, 2, 3
OUT_BF16_NHWC, OUT_F32_NCHW = 0, 1
RESAMPLE_NONE, RESAMPLE_AVGPOOL2, RESAMPLE_NEAREST2 = 0, 1, 2

# every symbol include/adb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I = C.c_int
SYMBOLS = {
    "adb_last_error": (C.c_char_p, []),
    "adb_version": (_I, []),
    "adb_device_check": (_I, []),
    "adb_plan_create": (_P, []),
    "adb_plan_destroy": (None, [_P]),
    "adb_plan_num_ops": (_I, [_P]),
    "adb_plan_run": (_I, [_P, _P]),
    "adb_plan_op_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "adb_plan_run_profiled": (_I, [_P, _P, C.POINTER(C.c_float), _I]),
    "adb_conv_block_n": (_I, [_I]),
    "adb_conv_gnb_supported": (_I, [_I, _I, _I, _I]),
    "adb_conv_igemm": (_I, [_P, C.POINTER(ConvDesc), _P]),
    "adb_attention": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "adb_groupnorm": (_I, [_P, C.POINTER(GnDesc), _P]),
    "adb_resample2x": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "adb_stem_conv": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "adb_timestep_embedding": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "adb_linear": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "adb_ddim_step": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, C.POINTER(C.c_float), _I, _P]),
    "adb_pack_uint8": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "adb_moments_accumulate": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "adb_memset0": (_I, [_P, _P, C.c_size_t, _P]),
    "adb_stem_im2col": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "adb_split_bf16": (_I, [_P, _P, _P, _P, C.c_size_t, _I, _P]),
    "adb_attention_lse": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "adb_attention_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "adb_attention_backward_ws": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "adb_set_attention_backward_fused": (_I, [_I]),
    "adb_gn_backward": (_I, [_P, C.POINTER(GnBwdDesc), _P]),
    "adb_pool_prepare": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "adb_pool_attention": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "adb_pool_attention_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "adb_pool_merge": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "adb_logsoftmax_grad": (_I, [_P, _P, _P, _P, _I, _I, C.c_float, _P]),
    "adb_attention_sd": (_I, [_P, C.POINTER(AttnSdDesc), _P]),
    "adb_layernorm": (_I, [_P, _P, _P, _P, _P, _I, _I, C.c_float, _P]),
    "adb_geglu": (_I, [_P, _P, _P, _I, _I, _P]),
    "adb_cfg_ddim_step": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, C.c_float, C.POINTER(C.c_float), _P]),
    "adb_pad_context": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "adb_timestep_embedding_f32": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "adb_dpm_x0": (_I, [_P, _P, _P, _P, C.c_size_t, _I, C.c_float, C.c_float, C.c_float, _P]),
    "adb_dpm_update": (_I, [_P, _P, _P, _P, _P, C.c_size_t, _I, C.c_float, C.c_float, C.c_float, C.c_float, _P]),
    "adb_cfg_combine": (_I, [_P, _P, _P, C.c_size_t, _I, C.c_float, _P]),
    "adb_plms_update": (_I, [_P, _P, _P, _P, _P, _P, _I, C.POINTER(C.c_float), _P, _P, C.c_size_t, _P]),
    "adb_resize_bilinear_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "adb_gather_patches": (_I, [_P, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "adb_pool3x3": (_I, [_P, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "adb_global_avgpool": (_I, [_P, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _P, _I, _I, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libadb200.so; fail loudly if it is not there (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(autodiffusion_b200 has no CPU or PyTorch fallback)"
        )
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def last_error() -> str:
    msg = lib().adb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


class AdbError(RuntimeError):
    pass


def check(status: int, what: str) -> int:
    if status < 0:
        raise AdbError(f"{what} failed (status {status}): {last_error()}")
    return status
