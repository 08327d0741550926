"""ctypes binding of the C-ABI in include/pvae_b200.h (libpvae_b200.so, built in-tree by nvcc).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libpvae_b200.so")
SOURCES = ["capi.cu", "sampler.cu", "fused_fwd.cu", "step.cu", "gemm_tc.cu", "lin_step.cu", "peer_allreduce.cu", "conv_tc.cu",
           "conv_misc.cu", "conv_net.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libpvae_b200.so (nvcc cross-compiles without a GPU).  One object per
    source, compiled in parallel and re-used while the source and the shared headers are unchanged, then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(_HERE), "include", "pvae_b200.h"))
    hdrs.append(os.path.abspath(__file__))  # the flags live here
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in srcs + hdrs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), newest_hdr):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, srcs))
    if verbose:
        print("".join(log for _, log in results))
    cmd = [nvcc] + LINK_FLAGS + ["-o", LIB_PATH] + [obj for obj, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB_PATH


_f, _i, _i64, _u64, _p = C.c_float, C.c_int, C.c_int64, C.c_uint64, C.c_void_p

class Cell(C.Structure):
    """pvae_cell of include/pvae_b200.h"""
    _fields_ = ([("kind", C.c_int), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ci", C.c_int), ("Co", C.c_int),
                 ("eps_res", C.c_float), ("precise", C.c_int)] +
                [(n, C.c_void_p) for n in ("w1", "ln1", "b1", "w2", "ln2", "b2", "se_w1", "se_b1", "se_w2", "se_b2", "skip_w",
                                           "skip_ln", "skip_b", "g_w1", "g_ln1", "g_b1", "g_w2", "g_ln2", "g_b2", "g_se_w1",
                                           "g_se_b1", "g_se_w2", "g_se_b2", "g_skip_w", "g_skip_ln", "g_skip_b")])


class LinStepDesc(C.Structure):
    """pvae_lin_step_desc of include/pvae_b200.h"""
    _fields_ = [("B", C.c_int64), ("B_global", C.c_int64), ("row_offset", C.c_int64), ("K", C.c_int64), ("D", C.c_int64),
                ("has_bias", C.c_int), ("n_exp", C.c_int), ("indicator", C.c_int), ("exc_only", C.c_int), ("precise", C.c_int),
                ("exit_logit", C.c_float), ("temp", C.c_float), ("beta", C.c_float),
                ("x", C.c_void_p), ("x_lo", C.c_void_p), ("params", C.c_void_p), ("params_lo", C.c_void_p), ("grads", C.c_void_p),
                ("ws", C.c_void_p), ("loss3", C.c_void_p), ("rate_max", C.c_void_p),
                ("seed", C.c_uint64), ("offset", C.c_uint64), ("offset_dev", C.c_void_p), ("events", C.c_void_p),
                ("update", C.c_int), ("step", C.c_int), ("exp_avg", C.c_void_p), ("exp_inf", C.c_void_p),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("lr_sched", C.c_void_p), ("n_sched", C.c_int), ("ring_n", C.c_int), ("state", C.c_void_p),
                ("philox_increment", C.c_uint64), ("rmax_ring", C.c_void_p),
                ("world", C.c_int), ("rank", C.c_int), ("token", C.c_uint32), ("peer_grads", C.c_void_p), ("peer_flags", C.c_void_p)]


_SIGNATURES = {
    "pvae_version": (C.c_int, []),
    "pvae_last_error": (C.c_char_p, []),
    "pvae_set_pdl": (_i, [_i]),
    "pvae_set_carveout": (_i, [_i]),
    "pvae_launch_count": (_i64, []),
    "pvae_num_partials": (_i64, [_i64]),
    "pvae_sampler_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _f, _i, _i, _f, _u64, _u64, _p, _p]),
    "pvae_sampler_bwd": (_i, [_p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _i, _i64, _i64, _i64, _i, _f, _i, _i, _f, _u64, _u64, _p, _p]),
    "pvae_rsample_fwd": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i, _f, _i, _f, _f, _u64, _u64, _p, _p]),
    "pvae_rsample_bwd": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i, _f, _i, _f, _f, _u64, _u64, _p, _p]),
    "pvae_kl_fwd": (_i, [_p, _p, _p, _i64, _i64, _p]),
    "pvae_kl_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i64, _i64, _p]),
    "pvae_gemm_ws_floats": (_i64, [_i64, _i64, _i]),
    "pvae_set_gemm_variant": (_i, [_i]),
    "pvae_gemm_tf32": (_i, [_p, _p, _p, _i64, _i64, _i64, _i, _i, _i, _i, _p, _p]),
    "pvae_recon_residual": (_i, [_p, _p, _p, _p, _i64, _i64, _f, _p]),
    "pvae_loss_assemble": (_i, [_p, _p, _p, _i64, _f, _i64, _i, _p]),
    "pvae_gemm_residual_parts": (_i64, [_i64]),
    "pvae_gemm_tf32_residual": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _i, _p]),
    "pvae_grad_norm_ws_floats": (_i64, []),
    "pvae_grad_norm": (_i, [_p, _i64, _f, _p, _p, _p]),
    "pvae_adamax_from_partials": (_i, [_p, _p, _p, _p, _p, _i, _f, _i, _f, _f, _f, _f, _p]),
    "pvae_gemm_effective_splits": (_i, [_i64, _i]),
    "pvae_sampler_bwd_partial_rows": (_i64, [_i64]),
    "pvae_sampler_bwd_bias_offset_rows": (_i64, [_i64]),
    "pvae_lin_step_train": (_i, [_p] * 8 + [_i64, _i64, _i64, _i, _i, _f, _f, _i, _i, _f, _i, _u64, _u64, _p, _p, _f, _i, _f, _f, _f, _f, _p]),
    "pvae_adamax_step_dev": (_i, [_p, _p, _p, _p, _i64, _p, _i, _p, _f, _f, _f, _f, _p, _u64, _p, _p, _i, _p]),
    "pvae_gather": (_i, [_p, _p, _i, _p, _p]),
    "pvae_adamax_step": (_i, [_p, _p, _p, _p, _i64, _f, _i, _f, _f, _f, _f, _p, _p]),
    "pvae_allreduce_adamax": (_i, [_p, _p, _p, _i64, _i64, _p, _p, _i, _i, C.c_uint32, _p, _p, _f, _i, _f, _f, _f, _f, _p]),
    "pvae_lin_step_ws_floats": (_i64, [_i64, _i64, _i64]),
    "pvae_lin_step_ws_offsets": (_i, [_i64, _i64, _i64, _p]),
    "pvae_lin_step": (_i, [_p, _p]),
    "pvae_lin_step_desc_size": (_i64, []),
    "pvae_gemm_tf32_planes": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _i, _i, _p, _p]),
    "pvae_tf32_lo": (_i, [_p, _p, _i64, _p]),
    "pvae_set_step_overlap": (_i, [_i]),
    "pvae_set_step_trace": (_i, [_p, _i]),
    "pvae_set_step_coef": (_i, [_i]),
    "pvae_set_step_fused": (_i, [_i]),
    "pvae_kl_diag_ws_floats": (_i64, [_i64, _i64]),
    "pvae_kl_diag": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i, _p]),
    "pvae_set_fused_debug": (_i, [_i]),
    "pvae_set_fused_timeline": (_i, [_p]),
    "pvae_set_gemm_timeline": (_i, [_p]),
    "pvae_fused_fwd_supported": (_i, [_i64, _i64, _i64]),
    "pvae_fused_fwd": (_i, [_p] * 14 + [_i64, _i64, _i64, _i64, _i, _f, _i, _i, _f, _f, _u64, _u64, _p, _p]),
    "pvae_lin_step_fwd_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i, _i, _f, _f, _i, _i, _f, _i,
                                   _u64, _u64, _p, _p, _p]),
    "pvae_conv_weight_prep": (_i, [_p] * 6 + [_i] * 7 + [_p]),
    "pvae_conv_weight_grad": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "pvae_conv_igemm": (_i, [_p] * 8 + [_i] * 14 + [_p, _p, _p, _i, _i, _i, _p]),
    "pvae_conv_wgrad_ws_floats": (_i64, [_i, _i, _i]),
    "pvae_conv_wgrad": (_i, [_p] * 8 + [_i] * 9 + [_p, _p, _p, _i, _i, _i, _i, _p]),
    "pvae_cell_saved_floats": (_i64, [_p, _i]),
    "pvae_cell_scratch_floats": (_i64, [_p]),
    "pvae_cell_fwd": (_i, [_p] * 6 + [_p]),
    "pvae_cell_bwd": (_i, [_p] * 7 + [_p]),
    "pvae_set_cell_overlap": (_i, [_i]),
    "pvae_set_fuse_reduce": (_i, [_i]),
    "pvae_set_dec_splits": (_i, [_i]),
    "pvae_recon_residual_parts": (_i, [_p, _i, _p, _p, _p, _i64, _i64, _f, _p]),
    "pvae_set_conv_variant": (_i, [_i]),
    "pvae_reduce_ws_floats": (_i64, [_i64]),
    "pvae_silu_fwd": (_i, [_p, _p, _i64, _p]),
    "pvae_add": (_i, [_p, _p, _p, _i64, _p]),
    "pvae_stem_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pvae_stem_wgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pvae_colsum": (_i, [_p, _p, _p, _i64, _i, _p]),
    "pvae_small_tn": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p]),
    "pvae_se_fwd": (_i, [_p] * 6 + [_f] + [_p] * 5 + [_i] * 7 + [_p]),
    "pvae_se_bwd": (_i, [_p] * 6 + [_f] + [_p] * 3 + [_i] * 5 + [_p]),
    "pvae_se_param_grads": (_i, [_p] * 9 + [_i, _i, _i, _p]),
    "pvae_upsample_fwd": (_i, [_p, _p] + [_i] * 6 + [_p]),
    "pvae_upsample_bwd": (_i, [_p, _p, _p] + [_i] * 6 + [_p]),
    "pvae_permute_cp": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _p]),
    "pvae_head_fwd": (_i, [_p] * 5 + [_i64, _i, _i, _p]),
    "pvae_head_bwd": (_i, [_p] * 7 + [_i64, _i, _i, _p]),
    "pvae_avgpool_fwd": (_i, [_p, _p, _i, _i, _i, _p]),
    "pvae_pool_silu_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "pvae_bias_act": (_i, [_p, _p, _p, _i64, _i, _i, _f, _u64, _u64, _p]),
    "pvae_relu_bwd": (_i, [_p, _p, _i64, _f, _p]),
    "pvae_layernorm_fwd": (_i, [_p] * 7 + [_i64, _i, _f, _p]),
    "pvae_layernorm_bwd": (_i, [_p] * 7 + [_i64, _i, _p]),
    "pvae_set_phase_a_events": (_i, [_i]),
    "pvae_set_absorb_exit": (_i, [_i]),
    "pvae_set_pool": (_i, [_i]),
    "pvae_set_wave_factor": (_i, [_f]),
    "pvae_set_rows_per_block": (_i, [_i]),
    "pvae_debug_log_check": (_i, [_p, _p]),
    "pvae_debug_draws": (_i, [_p, _p, _i64, _i64, _i64, _i, _u64, _u64, _p, _p]),
}

EXPORTS = tuple(_SIGNATURES)


def load():
    """Load libpvae_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(poissonvae_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pvae_last_error().decode()
        if rc == -1:
            raise ValueError(msg or what)
        if rc == -2:
            raise NotImplementedError(msg or what)
        raise RuntimeError(msg or what)
