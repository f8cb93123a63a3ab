"""Pin of the third-party piece the reference delegates its sparse layer to: ``deepspeed.moe.sharded_moe.top2gating`` of
``deepspeed==0.9.5`` (requirements.txt:9; call site llava_qwen1_5_moe.py:536-546).  DeepSpeed is not installable in the build container
(no network, not in the wheelhouse), so the oracle's restatement (oracle/restated.py::top2gating) is "parity
unpinned" -- THIS test turns it into a pin the moment a real DeepSpeed is importable (e.g. shipped by the driver under baseline/_ref):
it feeds both implementations the same fp32 logits and the same Gumbel noise (DeepSpeed draws its noise inside the function; the
sampler is patched to hand out ours) and requires identical l_aux, combine weights, dispatch mask and expert counts.
Skipped, loudly, while DeepSpeed is absent."""
import importlib
import os
import sys

import pytest
import torch

from oracle import restated as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _import_deepspeed_gating():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        return importlib.import_module("deepspeed.moe.sharded_moe")
    except Exception as e:                                  # noqa: BLE001 -- any import failure means "no usable DeepSpeed here"
        pytest.skip("deepspeed is not importable here (%s: %s): the MoE gate oracle stays 'parity unpinned'" % (type(e).__name__, e))


@pytest.mark.parametrize("S,E,cf,min_cap", [(64, 4, 1.5, 0), (333, 4, 1.0, 4), (2048, 4, 1.5, 0), (500, 8, 0.5, 0), (16, 2, 2.0, 8)])
def test_restated_top2gating_equals_deepspeed(S, E, cf, min_cap):
    sm = _import_deepspeed_gating()
    g = torch.Generator().manual_seed(S * 31 + E)
    logits = torch.randn(S, E, generator=g) * 2.0
    noise = R.gumbel_noise((S, E), g)
    # deepspeed/moe/sharded_moe.py top2gating: `logits_w_noise = logits + gumbel_rsample(logits.shape, device=logits.device)`
    orig = sm.gumbel_rsample
    sm.gumbel_rsample = lambda shape, device=None: noise.to(device if device is not None else "cpu")
    try:
        out = sm.top2gating(logits.clone(), cf, min_cap)
    finally:
        sm.gumbel_rsample = orig
    l_aux, combine, dispatch, exp_counts = out[0], out[1], out[2], out[3]
    mine = R.top2gating(logits, noise, cf, min_cap)
    assert combine.shape == mine["combine"].shape, (combine.shape, mine["combine"].shape)       # [S, E, C]: same capacity rule
    assert torch.equal(dispatch.bool(), mine["dispatch"])                                          # which token sits in which slot: bit exact
    torch.testing.assert_close(combine.float(), mine["combine"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(l_aux.float(), mine["l_aux"], rtol=1e-6, atol=1e-7)
    assert torch.equal(torch.as_tensor(exp_counts).long().cpu(), mine["exp_counts"].long())


def test_restated_moe_layer_equals_deepspeed_moe_module():
    """deepspeed.moe.layer.MoE end to end (gate + dispatch einsum + Experts + combine einsum) on CPU, ep_size 1, against
    oracle/restated.py::moe_layer with the module's own weights."""
    sm = _import_deepspeed_gating()
    try:
        from deepspeed.moe.layer import MoE
    except Exception as e:                                  # noqa: BLE001
        pytest.skip("deepspeed.moe.layer.MoE not importable: %s" % e)
    H, I, E, S = 32, 48, 4, 96

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj = torch.nn.Linear(H, I, bias=False)
            self.up_proj = torch.nn.Linear(H, I, bias=False)
            self.down_proj = torch.nn.Linear(I, H, bias=False)

        def forward(self, x):
            return self.down_proj(torch.nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))

    torch.manual_seed(0)
    try:
        moe = MoE(H, expert=MLP(), num_experts=E, ep_size=1, k=2, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                  use_residual=False)
    except Exception as e:                                  # noqa: BLE001 -- needs an initialised process group in some versions
        pytest.skip("deepspeed MoE could not be constructed without a distributed backend: %s" % e)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, S, H, generator=g)
    noise = R.gumbel_noise((S, E), g)
    orig = sm.gumbel_rsample
    sm.gumbel_rsample = lambda shape, device=None: noise
    try:
        out, l_aux, counts = moe(x)
    finally:
        sm.gumbel_rsample = orig
    sd = {"m." + k.replace("deepspeed_moe.", ""): v.detach() for k, v in moe.state_dict().items()}
    cfg = R.LMCfg(hidden=H, inter=I, layers=1, heads=2, kv_heads=2, vocab=8, moe_layers=[0], num_experts=E, capacity_factor=1.5, min_capacity=0)
    y, la, cnt = R.moe_layer(sd, "m.", cfg, x, noise)
    torch.testing.assert_close(out, y, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(l_aux, la, rtol=1e-6, atol=1e-7)
