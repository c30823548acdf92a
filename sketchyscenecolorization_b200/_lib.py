"""ctypes binding of libfgcolor.so (include/fgcolor.h) and its in-tree build.

The shared library is built next to this file with plain nvcc for sm_100a
(`python -m sketchyscenecolorization_b200._lib` or `__graft_entry__.build()`); there is no JIT and no
CPU implementation: loading fails loudly when the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libfgcolor.so")
SOURCES = ["api.cu", "elementwise.cu", "text.cu", "sn.cu", "loss.cu", "conv_simple.cu", "conv_small.cu", "conv_tc.cu", "conv_api.cu", "input.cu", "trunk.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fgcolor.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libfgcolor.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


class FgcSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int), ("ups", C.c_int), ("patch", C.c_void_p)]


_P, _I, _LL, _F, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "fgc_last_error": [],
    "fgc_version": [],
    "fgc_launch_count": [],
    "fgc_crc32c": [_P, _SZ, C.c_uint],
    "fgc_conv2d_ws_bytes": [C.POINTER(C.c_int), _I, _I, _I, _I],
    "fgc_set_conv_impl": [_I],
    "fgc_set_conv_flags": [_I, _I],
    "fgc_debug_conv_counts": [_P],
    "fgc_im2col_small": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "fgc_debug_set_trace": [_P, _I],
    "fgc_debug_keep_packed": [_I],
    "fgc_conv2d_fwd": [C.POINTER(FgcSrc), _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "fgc_conv2d_fwd_acc": [C.POINTER(FgcSrc), _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "fgc_conv2d_fwd_phase": [C.POINTER(FgcSrc), _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "fgc_split_term": [_P, _LL, _I, _P, _P, _P],
    "fgc_tapsum_w": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P],
    "fgc_conv2d_dgrad": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P],
    "fgc_conv2d_wgrad": [C.POINTER(FgcSrc), _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "fgc_chan_stats": [_P, _I, _LL, _I, _P, _P, _P],
    "fgc_cbn_act_fwd": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P],
    "fgc_cbn_act_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "fgc_prelu_fwd": [_P, _I, _LL, _P, _P, _P],
    "fgc_prelu_bwd": [_P, _P, _I, _LL, _I, _P, _P, _P, _P, _P],
    "fgc_prelu_bwd_acc": [_P, _P, _I, _LL, _P, _P, _P, _P],
    "fgc_colsum": [_P, _I, _LL, _I, _P, _P],
    "fgc_minmax_fwd": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "fgc_minmax_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "fgc_act_bwd": [_P, _P, _I, _LL, _I, _P, _P],
    "fgc_gate_fma_fwd": [_P, _P, _P, _I, _LL, _P, _P],
    "fgc_gate_fma_bwd": [_P, _P, _P, _I, _LL, _P, _P, _P],
    "fgc_gate_prelu_fwd": [_P, _P, _P, _I, _LL, _P, _P, _P],
    "fgc_gate_prelu_bwd": [_P, _P, _P, _P, _I, _LL, _P, _P, _P, _I, _P, _P, _P],
    "fgc_mul_up_fwd": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_mul_up_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "fgc_blend_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_blend_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "fgc_addpool_fwd": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_unpool_bwd": [_P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_upsample2x": [_P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_sum2x2": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "fgc_axpy": [_P, _P, _I, _I, _LL, _F, _P],
    "fgc_spatial_mean_fwd": [_P, _I, _I, _I, _I, _P, _P],
    "fgc_spatial_mean_bwd": [_P, _I, _I, _I, _I, _P, _P],
    "fgc_nchw_to_nhwc": [_P, _I, _I, _I, _I, _P, _I, _P],
    "fgc_nhwc_to_nchw": [_P, _I, _I, _I, _I, _P, _I, _P],
    "fgc_cast": [_P, _I, _P, _I, _LL, _P],
    "fgc_tanh_fwd": [_P, _I, _LL, _P, _P],
    "fgc_space_to_depth": [_P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_depth_to_space": [_P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_copy_rect": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "fgc_phase_weights": [_P, _I, _I, _I, _P, _P],
    "fgc_phase_wgrad": [_P, _I, _I, _I, _P, _P],
    "fgc_paired_input": [_P, _P, _I, _I, _I, _I, _I, C.c_ulonglong, _I, _P, _P, _P, _P],
    "fgc_l2norm_rows_fwd": [_P, _I, _I, _P, _P, _P],
    "fgc_l2norm_rows_bwd": [_P, _P, _P, _I, _I, _P, _P],
    "fgc_embedding_fwd": [_P, _P, _I, _I, _I, _I, _P, _P],
    "fgc_embedding_bwd": [_P, _P, _I, _I, _I, _I, _P, _P],
    "fgc_embedding_all_fwd": [_P, _P, _I, _I, _I, _P, _P],
    "fgc_embedding_all_bwd": [_P, _P, _I, _I, _I, _P, _P],
    "fgc_lstm_seq_fwd": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "fgc_lstm_seq_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "fgc_lstm_cell_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "fgc_lstm_cell_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "fgc_rows_group_sum": [_P, _I, _I, _I, _P, _P],
    "fgc_atanh_relu_fwd": [_P, _LL, _P, _P],
    "fgc_atanh_relu_bwd": [_P, _P, _LL, _P, _P],
    "fgc_sn_fwd": [_P, _P, _I, _I, _P, _P, _P],
    "fgc_sn_bwd": [_P, _P, _P, _I, _I, _P, _P, _P, _P],
    "fgc_softplus_mean": [_P, _I, _LL, _F, _P, _I, _P, _P],
    "fgc_ce_loss": [_P, _I, _P, _I, _I, _I, _F, _P, _I, _P, _P],
    "fgc_smooth_l1": [_P, _P, _I, _LL, _F, _P, _I, _P, _P],
    "fgc_reg_loss": [_P, _P, _P, _P, _I, _P, _I, _P],
    "fgc_adam_step": [_P, _P, _P, _P, _P, _P, _I, _F, _P, _F, _F, _I, _P],
    "fgc_opt_step": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P, _I, _P],
    "fgc_affine_act": [_P, _I, _LL, _I, _P, _P, _P, _P, _P, _I, _P, _P],
    "fgc_pad_cast_rows": [_P, _I, _LL, _I, _I, _P, _I, _P],
    "fgc_maxpool3x3s2": [_P, _I, _I, _I, _I, _I, _P, _P],
    "fgc_space_to_batch": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "fgc_batch_to_space": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "fgc_resize_bilinear": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
}
_RESTYPE = {"fgc_last_error": C.c_char_p, "fgc_launch_count": C.c_longlong, "fgc_conv2d_ws_bytes": C.c_size_t}

_lib = None


def load():
    """dlopen libfgcolor.so and attach prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libfgcolor.so is missing (%s): run `python -m sketchyscenecolorization_b200._lib` "
                           "or __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


class FgcError(RuntimeError):
    pass


def check(code, what=""):
    if code != 0:
        msg = load().fgc_last_error()
        raise FgcError("%s failed (%d): %s" % (what, code, msg.decode() if msg else "?"))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
