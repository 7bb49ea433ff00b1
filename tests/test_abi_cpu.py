"""CPU-side checks: the C-ABI library loads and exports every symbol include/nf_b200.h declares,
argument validation works without a GPU, and the host-side module mirrors the reference's
parameter tree (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import nf_b200
from nf_b200 import _lib
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nf_b200.h")).read()
    return sorted(set(re.findall(r"NF_API\s+[\w\s\*]+?\b(nf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTS) == syms, "ctypes binding and header disagree"
    assert _lib.lib.nf_abi_version() == 1


def test_bad_descriptor_is_rejected_with_message():
    d = _lib.make_desc(1, 32, 32, 4, 2, 1e-5, 0)   # D=1 has nothing to transform
    out = _lib.NfParamLayout()
    rc = _lib.lib.nf_param_layout(ctypes.byref(d), ctypes.byref(out))
    assert rc == -1
    assert b"D=1" in _lib.lib.nf_last_error()
    assert _lib.lib.nf_flow_packed_bytes(ctypes.byref(d)) == -1
    d2 = _lib.make_desc(8, 32, 32, 4, 2, 1e-5, 7)
    assert _lib.lib.nf_param_layout(ctypes.byref(d2), ctypes.byref(out)) == -1


@pytest.mark.parametrize("D,C,R,n,H1", [(9, 32, 16, 3, None), (2, 16, 8, 2, 24), (8, 64, 0, 1, None)])
def test_param_layout_matches_state_dict(D, C, R, n, H1):
    m = nf_b200.RealNVP(D, C, R, n, None, first_hidden=H1)
    lay = m._layout
    sd = m.state_dict()
    keys = list(sd.keys())
    ref_sd = synth.make_state_dict(D, C, R, n, H1=H1)
    assert keys == list(ref_sd.keys())
    for k in keys:
        assert tuple(sd[k].shape) == ref_sd[k].shape, k
    # slot order == state_dict order inside a block, offsets 16-byte aligned and non-overlapping
    per_block = len(keys) // n
    assert per_block == _lib.NF_P_SLOTS
    end = 0
    for slot in range(_lib.NF_P_SLOTS):
        assert keys[slot] == "blocks.0." + _lib.SLOT_NAMES[slot]
        assert lay.slot_off[slot] % 4 == 0 and lay.slot_off[slot] >= end
        assert lay.slot_numel[slot] == sd[keys[slot]].numel()
        end = lay.slot_off[slot] + lay.slot_numel[slot]
    assert lay.block_stride >= end and lay.total == n * lay.block_stride
    assert nf_b200.count_parameters(m) == sum(v.size for v in ref_sd.values())


def test_parameters_are_views_of_one_arena_and_survive_load_state_dict():
    D, C, R, n = 9, 32, 16, 2
    m = nf_b200.RealNVP(D, C, R, n, None)
    assert m._is_flat()
    sd = {k: torch.tensor(v) for k, v in synth.make_state_dict(D, C, R, n, seed=7).items()}
    m.load_state_dict(sd)
    assert m._is_flat()
    lay = m._layout
    flat = m._flat
    w = m.blocks[1].s[3].weight
    off = lay.block_stride + lay.slot_off[5 + 10 + 4]
    assert torch.equal(flat[off:off + w.numel()].view_as(w), sd["blocks.1.s.3.weight"])
    # frozen permutation factors, trainable everything else (torch_fcnf.py:30-31)
    assert not m.blocks[0].l.P.requires_grad and not m.blocks[0].l.P_inv.requires_grad
    assert m.blocks[0].l.L.requires_grad
    # in-place optimiser-style update is visible through the arena and bumps the version
    v0 = m._params_version()
    with torch.no_grad():
        w.add_(1.0)
    assert m._params_version() != v0
    assert torch.equal(flat[off:off + w.numel()].view_as(w), sd["blocks.1.s.3.weight"] + 1.0)


def test_same_seed_gives_the_reference_initialisation():
    """The containers are built with the same torch calls in the same order as torch_fcnf.py:15-31,61-93, so a
    seed reproduces the reference's initial weights BIT FOR BIT: the refinit fixtures hold the state_dict of the real
    reference module built after torch.manual_seed(0) (tests/golden/make_golden.py:make_refinit perturbs only the
    zero heads and the LayerNorm affine parameters afterwards).  The heads are zero (s = t = 0 at init)."""
    import os
    from tests.helpers import GOLDEN_DIR
    for name in ("refinit_d9", "refinit_d2", "refinit_d8"):
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        D, C, R, n = [int(v) for v in g["cfg"][:4]]
        torch.manual_seed(0)
        m = nf_b200.RealNVP(D, C, R, n, None)
        sd = m.state_dict()
        checked = 0
        for k in sd:
            leaf = k.split(".", 2)[2]
            if leaf.startswith("l.") or leaf[2:] in ("0.weight", "0.bias", "3.weight", "3.bias"):
                np.testing.assert_array_equal(sd[k].numpy(), g["sd/" + k], err_msg=f"{name}: {k}")
                checked += 1
        assert checked == n * (5 + 8)
        for blk in m.blocks:
            assert float(blk.s[6].weight.abs().max()) == 0.0 and float(blk.t[6].bias.abs().max()) == 0.0
            assert float((blk.s[2].weight - 1).abs().max()) == 0.0 and float(blk.t[5].bias.abs().max()) == 0.0
            W = blk.l.P @ (torch.tril(blk.l.L, -1) + torch.eye(D)) @ (torch.triu(blk.l.U, 1) + torch.diag(blk.l.s))
            assert torch.allclose(W @ W.T, torch.eye(D), atol=1e-5)   # orthogonal init


def test_default_precision_is_the_tensor_core_path():
    """A constructor call with the reference's exact positional signature (torch_nfbc_ant.py:147) must land on the
    tcgen05 path, not on the FFMA checker."""
    m = nf_b200.RealNVP(8, 32, 16, 2, None)
    assert m.precision == "fp32" and m._desc.precision == _lib.PREC["fp32"] == 1
    assert nf_b200.GEncoder(8, 4, hidden=16, n_hidden=1).precision == "fp32"


def test_cpu_tensors_are_refused():
    m = nf_b200.RealNVP(8, 32, 16, 2, None)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(4, 8), torch.zeros(4, 16))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.reverse(torch.zeros(4, 8), torch.zeros(4, 16))


def test_flops_formula():
    # SURVEY.md section 8(d) table
    assert abs(synth.flops_fwd_per_sample(**synth.CONFIGS["C2a"]) / 1e6 - 2.016) < 0.01
    assert abs(synth.flops_fwd_per_sample(**synth.CONFIGS["C4"]) / 1e6 - 37.97) < 0.05
    assert abs(synth.flops_fwd_per_sample(**synth.CONFIGS["C5"]) / 1e6 - 129.02) < 0.05


# ------------------------------------------------------------------ conditioning encoder (row f2)
@pytest.mark.parametrize("input_size,rep_size,hidden,n_hidden", [(58, 64, 512, 4), (29, 10, 64, 2), (600, 1024, 512, 4)])
def test_encoder_layout_matches_reference_state_dict(input_size, rep_size, hidden, n_hidden):
    m = nf_b200.GEncoder(input_size, rep_size, hidden=hidden, n_hidden=n_hidden)
    ref_sd = synth.make_encoder_state_dict(input_size, rep_size, hidden=hidden, n_hidden=n_hidden)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())            # fc1.., fc_out, norm1..: the reference's order
    for k in sd:
        assert tuple(sd[k].shape) == ref_sd[k].shape, k
    # slots are 16-byte aligned, disjoint, in arena order, and the parameters are views of the arena
    end = 0
    for p, off in m._slots:
        assert off % 4 == 0 and off >= end
        assert p.data_ptr() == m._flat.data_ptr() + 4 * off
        end = off + p.numel()
    assert end <= m._total
    # load_state_dict keeps the views and copies the values
    m.load_state_dict({k: torch.tensor(v) for k, v in ref_sd.items()})
    assert m._is_flat()
    np.testing.assert_array_equal(m.fc_out.weight.detach().numpy(), ref_sd["fc_out.weight"])
    d = m._desc
    assert _lib.lib.nf_encoder_stash_bytes(ctypes.byref(d), 100) >= 4 * (100 * input_size + n_hidden * 2 * 100 * hidden)
    assert _lib.lib.nf_encoder_workspace_bytes(ctypes.byref(d), 100) >= 4 * 2 * 100 * hidden


def test_encoder_bad_arguments():
    bad = _lib.NfEncoderDesc(58, 510, 64, 4, 1e-5, 1)          # hidden not a multiple of 4
    assert _lib.lib.nf_encoder_param_layout(ctypes.byref(bad), None) == -1
    assert b"hidden" in _lib.lib.nf_last_error()
    assert _lib.lib.nf_encoder_stash_bytes(ctypes.byref(bad), 8) == -1
    deep = _lib.NfEncoderDesc(58, 512, 64, 9, 1e-5, 1)          # more hidden layers than the ABI allows
    assert _lib.lib.nf_encoder_workspace_bytes(ctypes.byref(deep), 8) == -1
    with pytest.raises(RuntimeError, match="CUDA only"):
        nf_b200.GEncoder(8, 4, hidden=16, n_hidden=1)(torch.zeros(2, 8))
    with pytest.raises(ValueError):
        nf_b200.GEncoder(8, 4, precision="bf16")


def test_reduced_precision_mode_is_rejected_not_silently_downgraded():
    with pytest.raises(ValueError):
        nf_b200.RealNVP(8, 32, 16, 2, None, precision="bf16")
    d = _lib.make_desc(8, 32, 32, 16, 2, 1e-5, 2)              # NF_PREC_BF16 through the raw ABI
    out = _lib.NfParamLayout()
    assert _lib.lib.nf_param_layout(ctypes.byref(d), ctypes.byref(out)) == -4   # NF_E_UNSUPPORTED


def test_peer_memory_allreduce_entry_points_without_a_gpu():
    """The peer-memory all-reduce (csrc/dp.cu) is for 2..8 ranks of one box: the host side declines anything else
    before it touches CUDA or the process group, and the C ABI rejects bad arguments with an error code."""
    from nf_b200.module import _connect_peer_memory
    cpu = torch.device("cpu")
    assert _connect_peer_memory(None, cpu, 0, 1) is False      # single process: nothing to exchange
    assert _connect_peer_memory(None, cpu, 0, 9) is False      # more ranks than one NVSwitch box: NCCL keeps the job
    assert _lib.lib.nf_dp_p2p_alloc(0, None) < 0               # bad arguments: error code, nothing thrown
    assert _lib.lib.nf_dp_p2p_connect(None, 0, 2) < 0
    assert _lib.lib.nf_dp_p2p_connect(ctypes.c_char_p(b"\0" * 64 * 9), 0, 9) < 0
    assert b"2..8" in _lib.lib.nf_last_error()
