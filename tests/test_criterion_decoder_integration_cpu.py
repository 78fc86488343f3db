"""CPU: decoder -> criterion, wired as bench.py --criterion / a trainer would (BASELINE config 3): the decoder module
(native ops replaced by their CPU stand-ins, tests/test_host_logic_cpu.py::install_cpu_ops) feeds the device criterion
(sampling kernels emulated in host memory, oracle matcher) with its real output structure -- query slices of the
collected mask logits, dn_out with dn_args -- and the weighted loss back-propagates into the decoder's parameters.
Checked against the criterion oracle on the same outputs and seed."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402
import test_criterion_host_logic_cpu as T  # noqa: E402
from oracle import criterion_oracle as CO  # noqa: E402
from oracle import matcher_oracle as MO  # noqa: E402
from test_host_logic_cpu import build_decoder, install_cpu_ops  # noqa: E402

PTS = 60


class _Matcher(torch.nn.Module):
    def forward(self, outputs, targets):
        return MO.hungarian_match(outputs, targets, PTS, 2.0, 5.0, 5.0)[0]

    def __repr__(self, _repr_indent=4):
        return "Matcher oracle"


def test_decoder_outputs_through_criterion_and_back(monkeypatch):
    from mp_former_b200 import native, workload
    from mp_former_b200.criterion import SetCriterion
    install_cpu_ops(monkeypatch.setattr)
    monkeypatch.setattr(native, "point_sample_rows", T._emu_sample)
    monkeypatch.setattr(native, "point_sample_rows_bwd", T._emu_sample_bwd)
    monkeypatch.setattr(native, "topk_gather_rows", T._emu_topk_gather)
    monkeypatch.setattr(native, "MaskLossRows", T._EmuMaskLossRows)
    dec = build_decoder().train()
    x, mf = cases.decoder_inputs()
    targets = cases.dn_targets()
    torch.manual_seed(0)
    out = dec(x, mf, None, {"tgt": targets, "scalar": 2, "noise_scale": 0.0})
    assert out["dn_out"] is not None and out["dn_out"]["dn_args"]["pad_size"] > 0
    K, L = cases.DEC_CFG["num_classes"], cases.DEC_CFG["dec_layers"]

    # the recipe's weight dict (maskformer_model.py:123-132), built by the bench helper, on a CPU-capable matcher
    crit_dev, weighted_sum = workload.build_criterion(num_classes=K, dec_layers=L + 1, num_points=PTS, device="cpu")
    crit = SetCriterion(K, matcher=_Matcher(), weight_dict=crit_dev.weight_dict, eos_coef=0.1,
                        losses=["labels", "masks"], num_points=PTS, oversample_ratio=3.0,
                        importance_sample_ratio=0.75).train(True)
    assert len(out["aux_outputs"]) == L and all(f"loss_dice_dn_{i}" in crit.weight_dict for i in range(L))
    torch.manual_seed(7)
    losses = crit(out, targets)
    torch.manual_seed(7)
    ref = CO.set_criterion(out, targets, num_classes=K, eos_coef=0.1, losses=["labels", "masks"], num_points=PTS,
                           oversample_ratio=3.0, importance_sample_ratio=0.75, cost_class=2.0, cost_mask=5.0,
                           cost_dice=5.0, training=True)
    assert sorted(losses) == sorted(ref) and set(losses) == set(crit.weight_dict)
    for k in ref:
        assert torch.allclose(losses[k], ref[k], rtol=1e-5, atol=1e-6), k
    g_ref = torch.autograd.grad(weighted_sum(ref), [p for p in dec.parameters() if p.requires_grad],
                                retain_graph=True, allow_unused=True)
    weighted_sum(losses).backward()
    n = 0
    for p, g in zip([p for p in dec.parameters() if p.requires_grad], g_ref):
        if g is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0
            continue
        assert torch.allclose(p.grad, g, rtol=1e-3, atol=1e-6)
        n += 1
    assert n > 20
