"""ctypes binding of the C-ABI CUDA library (include/mcm_b200.h).

There is deliberately no fallback: if `libmcm_b200.so` is missing or cannot be loaded the import of
the product path raises, and every entry point raises `McmError` with the library's own message on
a non-zero status.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCM_LIB_PATH: development aid -- load an alternative build of the same library (kernel variant A/B runs)
LIB_PATH = os.environ.get("MCM_LIB_PATH") or os.path.join(_HERE, "libmcm_b200.so")


class McmError(RuntimeError):
    pass


class McmConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "input_feats", "seq_len", "latent_dim", "time_embed_dim", "ffn_dim", "text_latent_dim", "num_heads",
        "num_layers", "num_ctrl_blocks", "ctrl_cond_feats", "max_batch", "max_text_tokens", "precise_all")]


_FP = ctypes.POINTER(ctypes.c_float)
_IP = ctypes.POINTER(ctypes.c_int)


class McmSampler(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("n_steps", ctypes.c_int), ("eta", ctypes.c_float),
                ("timestep_map", _IP), ("alphas_cumprod", _FP), ("alphas_cumprod_prev", _FP),
                ("sqrt_recip_alphas_cumprod", _FP), ("sqrt_recipm1_alphas_cumprod", _FP),
                ("posterior_mean_coef1", _FP), ("posterior_mean_coef2", _FP),
                ("posterior_log_variance_clipped", _FP), ("seed", ctypes.c_ulonglong),
                ("model_mean_type", ctypes.c_int)]


class McmRepaint(ctypes.Structure):
    _fields_ = [("n_times", ctypes.c_int), ("times", _IP), ("betas", _FP), ("gt", ctypes.c_void_p),
                ("keep_mask", ctypes.c_void_p), ("noise_seq", ctypes.c_void_p), ("n_draws", ctypes.c_longlong),
                ("overlap_len", ctypes.c_int), ("add_blend", ctypes.c_int), ("blend_w", ctypes.c_void_p)]


# name -> (restype, argtypes); exactly the symbols include/mcm_b200.h declares
_VP, _I, _LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
SIGNATURES = {
    "mcm_create": (_I, [ctypes.POINTER(McmConfig), ctypes.POINTER(_VP)]),
    "mcm_destroy": (None, [_VP]),
    "mcm_set_option": (_I, [_VP, ctypes.c_char_p, _I]),
    "mcm_set_param": (_I, [_VP, ctypes.c_char_p, _VP, _LL]),
    "mcm_finalize_params": (_I, [_VP, _VP]),
    "mcm_prepare_conditions": (_I, [_VP, _I, _VP, _I, _VP, _VP, _I, _VP]),
    "mcm_denoise": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP]),
    "mcm_block_forward": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP]),
    "mcm_layers_forward": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "mcm_sample": (_I, [_VP, ctypes.POINTER(McmSampler), _I, _VP, _VP, _VP, _VP]),
    "mcm_sample_host": (_I, [_VP, ctypes.POINTER(McmSampler), _I, _VP, _VP, _VP, _VP]),
    "mcm_sample_repaint": (_I, [_VP, ctypes.POINTER(McmSampler), ctypes.POINTER(McmRepaint), _I, _VP, _VP, _VP]),
    "mcm_handoff_smplx": (_I, [_VP, _I, _I, _VP, _VP, _VP, _I, _VP, _I, _VP, _I, _VP, _I, _VP, _VP, _VP, _VP]),
    "mcm_handoff_denorm": (_I, [_VP, _VP, _VP, _LL, _I, _I, _VP, _VP, _VP]),
    "mcm_handoff_align_faces": (_I, [_VP, _VP, _LL, _I, _VP]),
    "mcm_cfg_combine": (_I, [_VP, _VP, ctypes.c_double, ctypes.c_double, _VP, _LL, _VP]),
    "mcm_part_mix": (_I, [_VP, _VP, _VP, _LL, _I, _I, _VP]),
    "mcm_sffn_forward": (_I, [_I] * 6 + [_VP] * 14),
    "mcm_stma_mix": (_I, [_I] * 8 + [_VP] * 19),
    "mcm_test_randn": (_I, [_VP, _LL, ctypes.c_ulonglong, ctypes.c_ulonglong, _VP]),
    "mcm_test_linear": (_I, [_I, _I, _I, _VP, _VP, _VP, _VP, _I, _VP]),
    "mcm_timing_enable": (None, [_I]),
    "mcm_timing_collect": (_I, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong),
                                ctypes.POINTER(ctypes.c_double)]),
    "mcm_debug_read": (_I, [ctypes.POINTER(ctypes.c_ulonglong), _I]),
    "mcm_debug_copy": (_I, [_VP, _I, _VP, _LL]),
    "mcm_debug_read32": (_I, [ctypes.POINTER(ctypes.c_ulonglong), _I]),
    "mcm_last_error": (ctypes.c_char_p, []),
    "mcm_gemm_launches": (ctypes.c_ulonglong, []),
    "mcm_kernel_launches": (ctypes.c_ulonglong, []),
    "mcm_version": (ctypes.c_char_p, []),
}

_lib = None


def load():
    """Load the shared library once; raises McmError if it is absent (run `python -c 'import
    __graft_entry__ as g; g.build()'` or ./build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McmError(f"{LIB_PATH} not found: the CUDA extension is not built (./build.sh). "
                       "motioncraft_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().mcm_last_error()
        raise McmError(msg.decode() if msg else f"mcm status {status}")


def kernel_launches():
    return int(load().mcm_kernel_launches())


def gemm_launches():
    return int(load().mcm_gemm_launches())


def timing_enable(on):
    load().mcm_timing_enable(1 if on else 0)


def timing_collect():
    """-> dict(gemm=dict(ms, launches, flops), row=dict(...)); clears the record."""
    ms = (ctypes.c_double * 3)()
    n = (ctypes.c_ulonglong * 3)()
    fl = (ctypes.c_double * 3)()
    check(load().mcm_timing_collect(ms, n, fl))
    return {"gemm": dict(ms=ms[0], launches=int(n[0]), flops=fl[0]),
            "row": dict(ms=ms[1], launches=int(n[1]), flops=fl[1]),
            "fused": dict(ms=ms[2], launches=int(n[2]), flops=fl[2])}
