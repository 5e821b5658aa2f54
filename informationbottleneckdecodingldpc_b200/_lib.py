"""ctypes binding of libibldpc.so (include/ibldpc.h).

The CUDA library is the only execution path of this package: if it is missing or cannot be
loaded every decoder / quantizer call raises.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IBLDPC_LIB") or os.path.join(HERE, "libibldpc.so")   # IBLDPC_LIB: A/B builds of the same sources
UNITS = ["ibldpc.cu", "ib_fast_cn.cu", "ib_fast_vn.cu", "ib_n4_cn_v2.cu", "ib_n4_cn_pair.cu", "ib_n4_vn_pair.cu", "ib_n4_vn_v2.cu", "ib_n4_vn_v4.cu", "ib_n4_vn3.cu", "ib_n4_coop.cu",
         "llr_f32.cu", "llr_f64.cu", "encoder.cu", "nccl_abi.cu", "ib_phase.cu", "ib_phase_wlan.cu", "ib_phase_dvbs2.cu", "ib_phase_reg36.cu", "ib_phase_reg36_tri.cu", "ib_perframe.cu", "llr_layered.cu", "ib_t32.cu", "ib_t32_cn.cu", "ib_t32_vn.cu", "ib_t32_out.cu", "ib_t32_phase.cu"]   # compiled in parallel
SOURCES = [os.path.join(HERE, "csrc", f) for f in UNITS + ["ib_kernels.cuh", "ib_kernels_n4.cuh", "ib_coop_n4.cuh", "llr_kernels.cuh", "kernel_tables.h", "ibldpc_internal.h", "ib_phase_n4.cuh", "ib_phase_sets.h", "ib_kernels_t32.cuh", "ib_triple_n4.cuh"]]
HEADER = os.path.join(os.path.dirname(HERE), "include", "ibldpc.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128"]   # 128: unreachable loops in degree-specialised templates

ALGO_MINSUM, ALGO_BP = 0, 1
F32, F64 = 32, 64


class CodeDesc(C.Structure):
    _fields_ = [("n_var", C.c_int32), ("n_chk", C.c_int32), ("n_edge", C.c_int32),
                ("inbox_start_chk", C.c_void_p), ("degree_chk", C.c_void_p), ("target_cells_chk", C.c_void_p),
                ("inbox_start_var", C.c_void_p), ("degree_var", C.c_void_p), ("target_cells_var", C.c_void_p)]


class EncoderDesc(C.Structure):
    _fields_ = [("n_var", C.c_int32), ("n_info", C.c_int32), ("method", C.c_int32),
                ("a_rowptr", C.c_void_p), ("a_col", C.c_void_p),
                ("eq", C.c_void_p), ("var", C.c_void_p), ("oth_ptr", C.c_void_p), ("oth", C.c_void_p),
                ("dense_inverse", C.c_void_p)]


ENC_SUBSTITUTION, ENC_DENSE = 1, 2


class LutDesc(C.Structure):
    _fields_ = [("card_channel", C.c_int32), ("card_decoder", C.c_int32), ("imax", C.c_int32),
                ("cn_degree", C.c_int32), ("vn_degree", C.c_int32),
                ("cn_lut", C.c_void_p), ("cn_lut_len", C.c_int64),
                ("vn_lut", C.c_void_p), ("vn_lut_len", C.c_int64),
                ("cn_match", C.c_void_p), ("cn_match_len", C.c_int64),
                ("vn_match", C.c_void_p), ("vn_match_len", C.c_int64)]


# every symbol include/ibldpc.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
SIGNATURES = {
    "ibldpc_create": (_i, [C.POINTER(CodeDesc), _i, C.POINTER(_vp)]),
    "ibldpc_set_luts": (_i, [_vp, C.POINTER(LutDesc)]),
    "ibldpc_decode_ib": (_i, [_vp, _vp, _i64, _i, _i, _vp, C.POINTER(C.c_int32), _vp]),
    "ibldpc_decode_ib_host": (_i, [_vp, _vp, _i64, _i, _i, _vp, C.POINTER(C.c_int32)]),
    "ibldpc_decode_ib_perframe": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "ibldpc_last_i_num": (_i, [_vp, C.POINTER(C.c_int32)]),
    "ibldpc_decode_ib_host_i32": (_i, [_vp, _vp, _i64, _i, _i, _vp, C.POINTER(C.c_int32)]),
    "ibldpc_decode_ib_host_packed": (_i, [_vp, _vp, _i64, _i, _i, _vp, _i64, C.POINTER(C.c_int32)]),
    "ibldpc_decode_llr": (_i, [_vp, _i, _i, _vp, _i64, _i, _i, _vp, C.POINTER(C.c_int32), _vp]),
    "ibldpc_decode_llr_layered": (_i, [_vp, _i, _i, _vp, _i64, _i, _i, _vp, C.POINTER(C.c_int32), _vp]),
    "ibldpc_layer_count": (_i, [_vp, C.POINTER(C.c_int32)]),
    "ibldpc_count_errors_u8": (_i, [_i, _vp, _i64, _i64, _i, _vp, C.POINTER(C.c_int64), _vp]),
    "ibldpc_count_errors_u8_async": (_i, [_i, _vp, _i64, _i64, _i, _vp, _vp, _vp]),
    "ibldpc_count_errors_llr": (_i, [_i, _vp, _i, _i64, _i64, _vp, C.POINTER(C.c_int64), _vp]),
    "ibldpc_count_errors_llr_async": (_i, [_i, _vp, _i, _i64, _i64, _vp, _vp, _vp]),
    "ibldpc_nccl_unique_id": (_i, [_vp]),
    "ibldpc_nccl_init": (_i, [_vp, _vp, _i, _i]),
    "ibldpc_allreduce_counters": (_i, [_vp, _vp, _i, _vp]),
    "ibldpc_nccl_finalize": (_i, [_vp]),
    "ibldpc_quantize": (_i, [_i, _vp, _i64, _vp, _i, _vp, _vp]),
    "ibldpc_quantize_llr": (_i, [_i, _vp, _i64, _vp, _i, _vp, _i, _vp, _vp]),
    "ibldpc_sample_direct": (_i, [_i, _vp, _i, _u64, _u64, _i64, _vp, _vp]),
    "ibldpc_sample_direct_llr": (_i, [_i, _vp, _i, _vp, _u64, _u64, _i64, _i, _vp, _vp]),
    "ibldpc_uniform": (_i, [_i, _u64, _u64, _i64, _vp, _vp]),
    "ibldpc_info": (_i, [_vp, C.POINTER(C.c_int32)]),
    "ibldpc_set_profiling": (_i, [_vp, _i]),
    "ibldpc_phase_times": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "ibldpc_set_host_chunk": (_i, [_vp, _i]),
    "ibldpc_encoder_create": (_i, [C.POINTER(EncoderDesc), _i, C.POINTER(_vp)]),
    "ibldpc_encode": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "ibldpc_encoder_destroy": (_i, [_vp]),
    "ibldpc_random_bits": (_i, [_i, _u64, _u64, _i64, _vp, _vp]),
    "ibldpc_awgn": (_i, [_i, _vp, _vp, _i64, C.c_double, _u64, _u64, _vp, _vp]),
    "ibldpc_plan_geometry": (_i, [_i64, _i, _i, _i, C.POINTER(C.c_int32)]),
    "ibldpc_host_chunk_schedule": (_i, [_i64, _i64, _i64, _i, C.POINTER(C.c_int64), _i]),
    "ibldpc_last_error": (C.c_char_p, []),
    "ibldpc_destroy": (_i, [_vp]),
}


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ibldpc.cu for sm_100a into the in-tree libibldpc.so (nvcc cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library."""
    deps = SOURCES + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in deps):
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    build_dir = os.path.join(HERE, "csrc", "build")
    os.makedirs(build_dir, exist_ok=True)

    def compile_unit(unit):
        obj = os.path.join(build_dir, unit.replace(".cu", ".o"))
        src = os.path.join(HERE, "csrc", unit)
        hdrs = [s for s in deps if not s.endswith(".cu")]
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(s) for s in [src] + hdrs):
            return obj
        cmd = ["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
        objs = list(pool.map(compile_unit, UNITS))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB_PATH


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B build of the same sources with extra -D defines into libibldpc_<name>.so (select it with IBLDPC_LIB=<path>).
    Used for measured kernel experiments; the product library is build_library()'s."""
    from concurrent.futures import ThreadPoolExecutor
    build_dir = os.path.join(HERE, "csrc", f"build_{name}")
    os.makedirs(build_dir, exist_ok=True)
    out = os.path.join(HERE, f"libibldpc_{name}.so")

    def compile_unit(unit):
        obj = os.path.join(build_dir, unit.replace(".cu", ".o"))
        cmd = ["nvcc"] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", os.path.join(HERE, "csrc", unit), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
        objs = list(pool.map(compile_unit, UNITS))
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-ldl"])
    return out


_lib = None


def lib() -> C.CDLL:
    """Load libibldpc.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ibldpc_last_error()
        raise RuntimeError(f"libibldpc error {rc}: {msg.decode() if msg else '?'}")
