"""CPU, world_size 2, gloo: the N>1 host logic of the batch-sharded path — shard boundaries,
max-over-ranks timing, gathering of per-shard answers.  The per-shard arithmetic is the CPU
oracle here (the CUDA path is covered per shard by tests/test_model_gpu.py)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from helpers import build_ours
    from oracle import sdnet_oracle
    from ruart_b200 import dist_utils, synth
    r, w = dist_utils.init("gloo")
    assert (r, w) == (rank, world)
    net, opt = build_ours("tiny", BERT_num_layers=2)
    sd = net.state_dict()
    batch = synth.make_batch(dict(B=4, n_ocr=6, n_od=3, max_ocr_num=100, max_od_num=30), seed=5, ragged=True)
    shard = dist_utils.shard_for_rank(batch)
    assert len(shard[1]["num_cnt"]) == 2
    probs, _, _ = sdnet_oracle.sdnet_forward(sd, opt, *shard)
    picks = synth.select_answers(probs, shard[1]["num_cnt"])
    all_picks = dist_utils.gather_picks(picks)
    ms = dist_utils.max_over_ranks([10.0 + rank, 5.0 - rank])
    # gradient all-reduce (mean) of the optional training step on a toy parameter set
    lin = torch.nn.Linear(3, 2)
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    for p_ in lin.parameters():
        p_.grad = torch.full_like(p_, float(rank + 1))
    n_red = dist_utils.allreduce_mean_grads(list(lin.parameters()) + [frozen])
    assert n_red == 8 and all(torch.allclose(p_.grad, torch.full_like(p_, 1.5)) for p_ in lin.parameters())
    q.put((rank, picks, all_picks, ms, float(probs.sum())))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, p0, all0, ms0, s0), (r1, p1, all1, ms1, s1) = res
    assert all0 == all1 == p0 + p1 and len(all0) == 4
    assert ms0 == ms1 == [11.0, 5.0]
    assert abs(s0 - 2.0) < 1e-4 and abs(s1 - 2.0) < 1e-4   # two questions per shard, rows sum to 1
