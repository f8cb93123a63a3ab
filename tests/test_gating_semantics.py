"""The one piece of the oracle that cannot be pinned on reference code (DeepSpeed 0.9.5 `top2gating` is an absent third-party dependency):
a second, independent statement of the published algorithm -- a token-by-token Python simulation of "seat first choices in token order, then
second choices behind ALL first choices of that expert, drop what does not fit the capacity, renormalise the kept gate values" -- must agree
with the vectorised restatement in oracle/restated.py (which the CUDA router is checked against bit for bit) on every integer and to fp32
rounding on the weights and the auxiliary loss."""
import math

import pytest
import torch

from oracle import restated as R


def simulate(logits, noise, capacity_factor, min_capacity):
    S, E = logits.shape
    gates = torch.softmax(logits, dim=1)
    C = max(math.ceil(S / E * capacity_factor * 2), min_capacity)
    first = [int(torch.argmax(gates[s])) for s in range(S)]
    second = []
    for s in range(S):
        row = (logits[s] + noise[s]).clone()
        row[first[s]] = float("-inf")
        second.append(int(torch.argmax(row)))
    n_first = [first.count(e) for e in range(E)]
    seat1, seat2, seen1, seen2 = [], [], [0] * E, [0] * E
    for s in range(S):
        seat1.append(seen1[first[s]])
        seen1[first[s]] += 1
    for s in range(S):
        seat2.append(n_first[second[s]] + seen2[second[s]])
        seen2[second[s]] += 1
    keep1 = [p < C for p in seat1]
    keep2 = [p < C for p in seat2]
    w1, w2 = [], []
    eps = torch.finfo(torch.float32).eps
    for s in range(S):
        a = float(gates[s, first[s]]) if keep1[s] else 0.0
        b = float(gates[s, second[s]]) if keep2[s] else 0.0
        d = max(a + b, eps)
        w1.append(a / d)
        w2.append(b / d)
    me = gates.mean(0)
    ce = torch.tensor([n / S for n in n_first])
    l_aux = float((me * ce).mean() * E * E)
    return dict(C=C, first=first, second=second, seat1=seat1, seat2=seat2, keep1=keep1, keep2=keep2, w1=w1, w2=w2, l_aux=l_aux, n_first=n_first)


@pytest.mark.parametrize("S,E,cf,min_cap,seed,skew", [(64, 4, 1.5, 0, 0, 0.0), (50, 4, 1.0, 4, 1, 2.0), (33, 8, 1.25, 0, 2, 3.0), (7, 4, 1.5, 0, 3, 0.0),
                                                       (128, 4, 0.5, 0, 4, 4.0), (16, 2, 2.0, 0, 5, 1.0)])
def test_vectorised_gating_equals_token_by_token_simulation(S, E, cf, min_cap, seed, skew):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(S, E, generator=g)
    logits[:, 0] += skew                                        # skewed routers overflow expert 0 and exercise the drops
    noise = R.gumbel_noise((S, E), g)
    r = R.top2gating(logits, noise, cf, min_cap)
    sim = simulate(logits, noise, cf, min_cap)
    assert r["capacity"] == sim["C"] == R.moe_capacity(S, E, cf, min_cap, 2)
    assert r["idx1"].tolist() == sim["first"] and r["idx2"].tolist() == sim["second"]
    assert r["exp_counts"].tolist() == sim["n_first"]
    assert r["keep1"].tolist() == sim["keep1"] and r["keep2"].tolist() == sim["keep2"]
    for s in range(S):                                          # seats only matter (and are only defined) for kept tokens
        if sim["keep1"][s]:
            assert int(r["slot1"][s]) == sim["seat1"][s]
        if sim["keep2"][s]:
            assert int(r["slot2"][s]) == sim["seat2"][s]
    assert torch.allclose(r["g1"], torch.tensor(sim["w1"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(r["g2"], torch.tensor(sim["w2"]), rtol=1e-6, atol=1e-7)
    assert abs(float(r["l_aux"]) - sim["l_aux"]) < 1e-6
    # the dense combine tensor says the same thing: one (expert, seat) per kept choice, no seat used twice
    comb = r["combine"]
    assert comb.shape == (S, E, sim["C"])
    used = (comb > 0).sum(0)
    assert int(used.max()) <= 1
    if skew >= 2.0:
        assert not all(sim["keep1"]) or not all(sim["keep2"])   # the skewed cases really drop tokens
