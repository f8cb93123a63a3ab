"""N>1 host logic on CPU (gloo, world_size 2): flat gradient arena packing, the student-gradient all-reduce (the only
data-path collective), per-rank sharding of the synthetic batches and rank-0-only reference arm."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "llava-mod_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from llavamod.model import synthetic as S
    from llavamod.train.engine import TrainState
    import bench
    student = S.make_student("tiny", "tiny", device="cpu", dtype=torch.float32, seed=1)
    st = TrainState(student, lr=1e-3)
    assert st.world == world
    # every trainable parameter is a view of the flat arena and its grad a view of the flat grad buffer
    n = 0
    for name, p in student.named_parameters():
        if p.requires_grad:
            arena = st.w32 if p.dtype == torch.float32 and "wg" in name else (st.w16 if st.n16 else st.w32)
            assert p.grad is not None and p.grad.shape == p.shape
            n += p.numel()
    assert n <= st.n16 + st.n32 and n > 0
    # rank-dependent gradients -> all-reduce sums them
    st.zero_grad()
    for i, (name, p) in enumerate(sorted((k, v) for k, v in student.named_parameters() if v.requires_grad)):
        p.grad.fill_(float(rank + 1) * (i + 1))
    st.allreduce_grads()
    ok = True
    for i, (name, p) in enumerate(sorted((k, v) for k, v in student.named_parameters() if v.requires_grad)):
        ok &= bool(torch.all(p.grad == float(sum(range(1, world + 1))) * (i + 1)))
    # fused buffers really alias the parameters the reference exposes
    lay = student.model.layers[1].mlp
    ok &= lay.gate_proj.weight.data_ptr() == lay.gu_weight.data_ptr()
    ex = student.model.layers[0].mlp.deepspeed_moe.experts
    ok &= ex.deepspeed_experts[2].down_proj.weight.data_ptr() == ex.dn_weight[2].data_ptr()
    # data sharding: ranks draw different synthetic samples
    b = bench.synth_batch(bench.WORKLOADS["tiny"], rank, 0, 512)
    ids = [torch.zeros_like(b["input_ids"]) for _ in range(world)]
    dist.all_gather(ids, b["input_ids"])
    ok &= not torch.equal(ids[0], ids[1])
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {0: True, 1: True}


def test_reference_arm_other_ranks_exit_quietly(monkeypatch, capsys):
    import bench
    monkeypatch.setenv("RANK", "1")
    import types
    bench.run_reference(types.SimpleNamespace(workload="tiny", steps=1, warmup=0, gpus=2))
    assert capsys.readouterr().out == ""
