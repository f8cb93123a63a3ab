"""Host-side pieces of the eval path (no GPU): nucleus filtering against transformers' TopPLogitsWarper, and KeywordsStoppingCriteria
against the reference's own class (llavamod/mm_utils.py:73-105, loaded in a subprocess because it needs the `llavamod` package name)."""
import json
import os
import subprocess
import sys
import types

import pytest
import torch

from tests.golden.make_data_golden import load_tokenizer

REF = os.environ.get("LLAVAMOD_REFERENCE", "/root/reference")


def test_top_p_filter_matches_transformers_warper():
    from transformers.generation.logits_process import TopPLogitsWarper
    from llavamod.model.generation import _top_p_filter
    g = torch.Generator().manual_seed(0)
    for top_p in (0.1, 0.5, 0.9, 0.999):
        logits = torch.randn(3, 200, generator=g) * 3
        want = TopPLogitsWarper(top_p=top_p)(None, logits.clone())
        assert torch.equal(_top_p_filter(logits.clone(), top_p), want)


CASES = [("USER: hi ASSISTANT: There are two birds.<|endoftext|>", ["<|endoftext|>"]),
         ("USER: hi ASSISTANT: There are two birds", ["<|endoftext|>"]),
         ("USER: hi ASSISTANT: A small red square", ["red square", "zzz"]),
         ("USER: hi ASSISTANT: A", ["red square"])]


def _mine(tok):
    from llavamod.mm_utils import KeywordsStoppingCriteria
    out = []
    for text, kws in CASES:
        ids = torch.tensor([tok(text).input_ids])
        start = ids[:, :5]
        crit = KeywordsStoppingCriteria(kws, tok, start)
        out.append([bool(crit(ids[:, :n], None)) for n in range(6, ids.shape[1] + 1)])
    return out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "llavamod")), reason="reference tree not present (GPU box)")
def test_keywords_stopping_criteria_matches_reference_class(golden_dir):
    tok_path = os.path.join(golden_dir, "tiny_tokenizer.json")
    code = r'''
import json, os, sys, types, torch
sys.path.insert(0, %r)
from tests.golden.make_data_golden import load_tokenizer
base = os.path.join(%r, "llavamod")
m = types.ModuleType("llavamod"); m.__path__ = [base]; sys.modules["llavamod"] = m
from llavamod.mm_utils import KeywordsStoppingCriteria
tok = load_tokenizer(%r)
out = []
for text, kws in %r:
    ids = torch.tensor([tok(text).input_ids])
    crit = KeywordsStoppingCriteria(kws, tok, ids[:, :5])
    out.append([bool(crit(ids[:, :n], None)) for n in range(6, ids.shape[1] + 1)])
print("RESULT" + json.dumps(out))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), REF, tok_path, CASES)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    want = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT")][0][6:])
    got = _mine(load_tokenizer(tok_path))
    assert got == want
    assert any(any(row) for row in want) and not all(all(row) for row in want)      # the cases exercise both outcomes
