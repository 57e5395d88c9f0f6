"""Host-side logic that needs no GPU: layout arithmetic, config packing, schedule choice, trl patching."""
import sys
import types

import pytest
import torch

import swh_trl_b200 as S
from swh_trl_b200 import ops


def test_collapse_rows():
    x = torch.empty(4, 7, 32)
    assert ops.collapse_rows(tuple(x.shape), tuple(x.stride())) == 32
    big = torch.empty(4, 9, 32)
    sl = big[:, :-1][:, -7:]  # the slice of grpo_trainer.py:1252-1254: not one arithmetic progression
    assert ops.collapse_rows(tuple(sl.shape), tuple(sl.stride())) is None
    wide = torch.empty(6, 40)[:, :32]  # padded rows
    assert ops.collapse_rows(tuple(wide.shape), tuple(wide.stride())) == 40
    one = torch.empty(1, 5, 32)[:, 1:]
    assert ops.collapse_rows(tuple(one.shape), tuple(one.stride())) == 32
    tr = torch.empty(32, 4).t()  # vocab dim not dense
    assert ops.collapse_rows(tuple(tr.shape), tuple(tr.stride())) is None
    assert ops.collapse_rows((32,), (1,)) == 32
    # two-level (batch, row) layouts: the slice the trainer takes from the model output is read in place
    assert ops.collapse_rows2(tuple(sl.shape), tuple(sl.stride())) == (32, 7, 9 * 32)
    assert ops.collapse_rows2(tuple(x.shape), tuple(x.stride())) == (32, 0, 0)
    assert ops.collapse_rows2(tuple(wide.shape), tuple(wide.stride())) == (40, 0, 0)
    four = torch.empty(2, 3, 9, 32)[:, :, 1:8]  # leading dims merge: still (batch, row)
    assert ops.collapse_rows2(tuple(four.shape), tuple(four.stride())) == (32, 7, 288)
    three = torch.empty(2, 5, 9, 32)[:, 1:4, 1:8]  # three levels: needs a copy
    assert ops.collapse_rows2(tuple(three.shape), tuple(three.stride())) is None
    merged = torch.empty(2, 3, 7, 32)  # dense: merges to flat
    assert ops.collapse_rows2(tuple(merged.shape), tuple(merged.stride())) == (32, 0, 0)
    assert ops.collapse_rows2(tuple(tr.shape), tuple(tr.stride())) is None


def test_make_cfg():
    cfg = ops.make_cfg(0.04, 0.2, 0.28, 2.0, "dr_grpo", "sequence", 256, 0.5)
    assert cfg.clip_low == torch.tensor(1 - 0.2, dtype=torch.float32).item()
    assert cfg.clip_high == torch.tensor(1 + 0.28, dtype=torch.float32).item()
    assert (cfg.has_delta, cfg.loss_type, cfg.is_level) == (1, 2, 1)
    assert ops.make_cfg(0, 0.2, 0.2, None, "bnpo", "token", 8).has_delta == 0
    with pytest.raises(ValueError, match="Unknown loss type"):  # grpo_trainer.py:2137
        ops.make_cfg(0, 0.2, 0.2, None, "x", "token", 8)
    with pytest.raises(ValueError, match="Unknown importance sampling level"):  # :2106-2109
        ops.make_cfg(0, 0.2, 0.2, None, "bnpo", "x", 8)


def test_schedule_choice():
    assert S.GRPOLoss().schedule(has_old=True) == "fused"
    assert S.GRPOLoss(importance_sampling_level="sequence").schedule(has_old=False) == "fused"
    assert S.GRPOLoss(importance_sampling_level="sequence").schedule(has_old=True) == "two-phase"
    assert S.GRPOLoss(top_entropy_quantile=0.2).schedule(has_old=False) == "two-phase"


def test_patch_trl_rebinds_every_importer():
    """Ten reference modules bind selective_log_softmax at import time."""
    def orig(logits, index):
        return "reference"
    pkg = types.ModuleType("faketrl")
    utils = types.ModuleType("faketrl.trainer.utils")
    utils.selective_log_softmax = orig
    utils.entropy_from_logits = orig
    grpo_mod = types.ModuleType("faketrl.trainer.grpo_trainer")
    grpo_mod.selective_log_softmax = orig  # `from .utils import selective_log_softmax`
    grpo_mod.get_high_entropy_mask = orig

    class GRPOTrainer:
        def _compute_loss(self, model, inputs):
            return "reference"
    grpo_mod.GRPOTrainer = GRPOTrainer
    grpo_mod.is_liger_kernel_available = lambda: False  # liger-kernel not installed: the name below does not exist
    core = types.ModuleType("faketrl.core")
    core.masked_whiten = orig
    ppo_mod = types.ModuleType("faketrl.trainer.ppo_trainer")  # ppo_trainer.py:54-71 binds the mask helpers too
    ppo_mod.first_true_indices = orig
    ppo_mod.truncate_response = orig
    other = types.ModuleType("faketrl.other")
    mods = {"faketrl": pkg, "faketrl.trainer.utils": utils, "faketrl.trainer.grpo_trainer": grpo_mod,
            "faketrl.core": core, "faketrl.trainer.ppo_trainer": ppo_mod, "faketrl.other": other}
    sys.modules.update(mods)
    try:
        report = S.patch_trl("faketrl")
    finally:
        for k in mods:
            sys.modules.pop(k)
    # the leaf bindings keep the reference's dtype contract (out_dtype defaults to the logits dtype) and hand CPU /
    # fp64 inputs to the reference's own function instead of raising
    for bound in (utils.selective_log_softmax, grpo_mod.selective_log_softmax, utils.entropy_from_logits):
        assert bound._b200trl_patched and bound is not orig
        assert bound(torch.zeros(2, 4), torch.zeros(2, dtype=torch.long)) == "reference"
    assert utils._trl_original_selective_log_softmax is orig
    assert grpo_mod.get_high_entropy_mask is S.get_high_entropy_mask
    assert core.masked_whiten is S.masked_whiten
    assert GRPOTrainer._compute_loss is S.compute_loss
    # the seam: `use_liger_loss=True` now constructs the B200 operator with the reference's keyword arguments
    assert grpo_mod.LigerFusedLinearGRPOLoss is S.B200FusedLinearGRPOLoss and grpo_mod.is_liger_kernel_available()
    op = grpo_mod.LigerFusedLinearGRPOLoss(beta=0.04, epsilon_low=0.2, epsilon_high=0.28, temperature=0.9,
                                           use_ref_model=True, loss_type="dr_grpo", max_completion_length=64)
    assert (op.beta, op.epsilon_high, op.temperature, op.loss_type) == (0.04, 0.28, 0.9, "dr_grpo")
    assert GRPOTrainer._trl_original_compute_loss(None, None, None) == "reference"
    assert ppo_mod.first_true_indices is S.first_true_indices and ppo_mod.truncate_response is S.truncate_response
    assert ppo_mod._trl_original_truncate_response is orig
    assert "faketrl.other" not in report and set(report) == {"faketrl.trainer.utils", "faketrl.trainer.grpo_trainer",
                                                              "faketrl.core", "faketrl.trainer.ppo_trainer"}


def test_graphed_step_has_no_cpu_path():
    """GraphedStep validates its static buffers before touching CUDA: CPU tensors raise (no fallback)."""
    import swh_trl_b200 as S
    with pytest.raises(ValueError, match="no CPU fallback"):
        S.GraphedStep(lambda s: {"y": s["x"]}, {"x": torch.zeros(4)})
    with pytest.raises(ValueError):
        S.GraphedStep(lambda s: {}, {})
    with pytest.raises(TypeError):
        S.GraphedStep(lambda s: {}, {"x": 3})


def test_private_workspaces_scope():
    """Inside ops.private_workspaces the operators' scratch buffers live in the caller's dict (a GraphedStep's), keyed
    by kernel family, grow on demand, and the scope unwinds on exit -- also when the body raises."""
    cpu = torch.device("cpu")
    store = {}
    with ops.private_workspaces(store) as s:
        assert s is store
        a = ops._workspace(cpu, 100, "family_a", zero=True)
        assert a.numel() == 256 and int(a.sum()) == 0
        assert ops._workspace(cpu, 200, "family_a", zero=True) is a          # fits: the same buffer
        b = ops._workspace(cpu, 1000, "family_a", zero=False)                # grows: replaced inside the store only
        assert b.numel() == 1000 and store[(cpu, "family_a")] is b
        with ops.private_workspaces({}) as inner:                             # nests
            ops._workspace(cpu, 10, "family_b", zero=False)
            assert (cpu, "family_b") in inner and (cpu, "family_b") not in store
        assert ops._ws_private is store
    assert ops._ws_private is None
    with pytest.raises(RuntimeError):
        with ops.private_workspaces({}):
            raise RuntimeError("boom")
    assert ops._ws_private is None
    assert not any(k[-1] in ("family_a", "family_b") for k in ops._ws_cache)   # nothing leaked into the shared cache
