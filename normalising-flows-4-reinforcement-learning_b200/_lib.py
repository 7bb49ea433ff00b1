"""ctypes binding of libnf_b200.so (include/nf_b200.h).  No CPU fallback: if the CUDA
library is missing the import of this module fails loudly."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnf_b200.so")

PREC = {"fp32_simt": 0, "fp32": 1, "tf32x3": 1, "tf32": 3}

NF_P_SLOTS = 25
NF_P_NET_SLOTS = 10
SLOT_NAMES = ["l.s", "l.U", "l.L", "l.P", "l.P_inv"] + [
    f"{net}.{k}" for net in ("t", "s")
    for k in ("0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias", "5.weight", "5.bias",
              "6.weight", "6.bias")]
NF_PREP_FWD, NF_PREP_INV = 1, 2


class NfFlowDesc(Structure):
    _fields_ = [("D", c_int32), ("C", c_int32), ("H1", c_int32), ("R", c_int32), ("n_blocks", c_int32),
                ("ln_eps", c_float), ("precision", c_int32), ("reserved", c_int32)]


class NfEncoderDesc(Structure):
    _fields_ = [("in_features", c_int32), ("hidden", c_int32), ("out_features", c_int32), ("n_hidden", c_int32),
                ("ln_eps", c_float), ("precision", c_int32)]


class NfTarflowDesc(Structure):
    _fields_ = [("A", c_int32), ("C", c_int32), ("R", c_int32), ("head_dim", c_int32), ("n_blocks", c_int32),
                ("n_layers", c_int32), ("expansion", c_int32), ("ln_eps", c_float), ("logscale_clamp", c_float),
                ("precision", c_int32)]


NF_TARFLOW_LAYOUT_SLOTS = 22


class NfParamLayout(Structure):
    _fields_ = [("slot_off", c_int64 * NF_P_SLOTS), ("slot_numel", c_int64 * NF_P_SLOTS),
                ("block_stride", c_int64), ("net_stride", c_int64), ("total", c_int64)]


class NfError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        # Build in-tree on first use (nvcc cross-compiles for sm_100a without a GPU).
        from .build import build
        build()
    lib = ctypes.CDLL(LIB_PATH)
    D = POINTER(NfFlowDesc)
    E = POINTER(NfEncoderDesc)
    T = POINTER(NfTarflowDesc)
    fp, vp, i64 = c_void_p, c_void_p, c_int64   # device pointers are passed as integers
    sigs = {
        "nf_abi_version": (c_int, []),
        "nf_last_error": (c_char_p, []),
        "nf_param_layout": (c_int, [D, POINTER(NfParamLayout)]),
        "nf_flow_packed_bytes": (c_int64, [D]),
        "nf_flow_stash_bytes": (c_int64, [D, i64]),
        "nf_flow_workspace_bytes": (c_int64, [D, i64]),
        "nf_flow_prepare": (c_int, [D, fp, vp, c_int, vp, i64, vp]),
        "nf_flow_forward": (c_int, [D, fp, vp, fp, fp, i64, fp, fp, fp, vp, vp, i64, vp]),
        "nf_flow_inverse": (c_int, [D, fp, vp, fp, fp, i64, fp, fp, fp, vp, i64, vp]),
        "nf_flow_backward": (c_int, [D, fp, vp, vp, fp, i64, fp, fp, fp, fp, fp, vp, i64, vp]),
        "nf_flow_backward_range": (c_int, [D, fp, vp, vp, fp, i64, fp, fp, fp, fp, fp, c_int32, c_int32, vp, i64, vp]),
        "nf_flow_inverse_backward": (c_int, [D, fp, vp, vp, fp, i64, fp, fp, fp, fp, fp, vp, i64, vp]),
        "nf_flow_backward_dp": (c_int, [D, fp, vp, vp, fp, i64, fp, fp, fp, fp, fp, c_int32, vp, i64, vp]),
        "nf_dp_unique_id": (c_int, [vp]),
        "nf_dp_init": (c_int, [vp, c_int32, c_int32]),
        "nf_dp_world": (c_int, []),
        "nf_dp_allreduce": (c_int, [fp, i64, c_int32, vp]),
        "nf_dp_broadcast": (c_int, [fp, i64, c_int32, vp]),
        "nf_dp_p2p_alloc": (c_int, [i64, vp]),
        "nf_dp_p2p_connect": (c_int, [vp, c_int32, c_int32]),
        "nf_dp_p2p_enable": (c_int, [c_int32]),
        "nf_dp_p2p_status": (c_int, []),
        "nf_dp_p2p_timeline": (c_int, [vp]),
        "nf_dp_shutdown": (c_int, []),
        "nf_affine_logdet_fwd": (c_int, [fp, fp, fp, fp, fp, c_float, i64, c_int32, fp, fp, fp, vp]),
        "nf_affine_logdet_inv": (c_int, [fp, fp, fp, fp, fp, c_float, i64, c_int32, fp, fp, vp]),
        "nf_affine_logdet_bwd": (c_int, [fp, fp, fp, fp, i64, c_int32, fp, fp, fp, vp]),
        "nf_plu_build": (c_int, [fp, fp, fp, fp, fp, c_int32, fp, fp, fp, vp, i64, vp]),
        "nf_linear": (c_int, [fp, i64, fp, i64, fp, fp, i64, i64, c_int32, c_int32, c_int32, vp]),
        "nf_linear_wgrad": (c_int, [fp, i64, fp, i64, fp, i64, i64, c_int32, c_int32, c_int32, vp]),
        "nf_prof_enable": (c_int, [c_int]),
        "nf_prof_collect": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int64), c_int]),
        "nf_launch_count": (c_int64, [c_int]),
        "nf_stdnormal_logprob": (c_int, [fp, i64, c_int32, fp, vp]),
        "nf_stdnormal_logprob_bwd": (c_int, [fp, fp, i64, c_int32, fp, vp]),
        "nf_logprob_group_argmin": (c_int, [fp, fp, fp, i64, i64, c_int32, fp, fp, fp, vp]),
        "nf_noise_augment": (c_int, [fp, i64, c_float, ctypes.c_uint64, ctypes.c_uint64, fp, vp]),
        "nf_nll_forward": (c_int, [fp, fp, i64, c_int32, fp, vp]),
        "nf_nll_cotangents": (c_int, [fp, fp, i64, c_int32, fp, fp, vp]),
        "nf_adamw_step": (c_int, [D, fp, fp, fp, fp, fp, c_float, c_float, c_float, c_float, c_float, i64, vp]),
        "nf_encoder_param_layout": (c_int64, [E, POINTER(c_int64)]),
        "nf_encoder_stash_bytes": (c_int64, [E, i64]),
        "nf_encoder_workspace_bytes": (c_int64, [E, i64]),
        "nf_encoder_forward": (c_int, [E, fp, fp, i64, fp, vp, vp, i64, vp]),
        "nf_encoder_backward": (c_int, [E, fp, vp, i64, fp, fp, fp, vp, i64, vp]),
        "nf_tarflow_param_layout": (c_int64, [T, POINTER(c_int64)]),
        "nf_tarflow_stash_bytes": (c_int64, [T, i64]),
        "nf_tarflow_workspace_bytes": (c_int64, [T, i64]),
        "nf_tarflow_forward": (c_int, [T, fp, fp, fp, i64, fp, fp, fp, vp, vp, i64, vp]),
        "nf_tarflow_inverse": (c_int, [T, fp, fp, fp, i64, fp, fp, vp, i64, vp]),
        "nf_tarflow_backward": (c_int, [T, fp, vp, i64, fp, fp, fp, fp, fp, vp, i64, vp]),
        "nf_set_option": (c_int, [c_char_p, c_int]),
        "nf_get_option": (c_int, [c_char_p]),
        "nf_graphs_enable": (c_int, [c_int]),
        "nf_graphs_clear": (c_int, []),
        "nf_graphs_stats": (c_int, [POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sigs)


lib, EXPORTS = _load()


def check(rc: int, what: str):
    if rc != 0:
        msg = lib.nf_last_error().decode(errors="replace")
        raise NfError(f"{what} failed (code {rc}): {msg}")


NF_FLAG_LN_FAST_VARIANCE = 1


def make_desc(D, C, H1, R, n, ln_eps, precision, flags=0) -> NfFlowDesc:
    return NfFlowDesc(int(D), int(C), int(H1), int(R), int(n), float(ln_eps), int(precision), int(flags))


def param_layout(desc: NfFlowDesc) -> NfParamLayout:
    out = NfParamLayout()
    check(lib.nf_param_layout(ctypes.byref(desc), ctypes.byref(out)), "nf_param_layout")
    return out


def set_option(name: str, value: int):
    check(lib.nf_set_option(name.encode(), int(value)), "nf_set_option")


def get_option(name: str) -> int:
    return lib.nf_get_option(name.encode())


def graph_stats():
    """(replays, captures, eager runs) of the library's launch-sequence cache."""
    a, b, c = c_int64(), c_int64(), c_int64()
    lib.nf_graphs_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


PROF_KINDS = ["tc_gemm", "tc_wgrad", "simt_gemm", "ln_fwd", "ln_bwd", "affine", "param", "row_tail", "row_bwd_head",
              "row_bwd_tail", "col_reduce", "tf_row", "tf_attn"]


def prof_collect():
    """{kind: (ms, work, launches)} since the last call (see nf_prof_collect)."""
    n = len(PROF_KINDS)
    ms = (ctypes.c_double * n)()
    work = (ctypes.c_double * n)()
    cnt = (c_int64 * n)()
    lib.nf_prof_collect(ms, work, cnt, n)
    return {k: (ms[i], work[i], cnt[i]) for i, k in enumerate(PROF_KINDS)}
