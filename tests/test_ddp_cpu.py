"""CPU (gloo, world_size 2): the N>1 host logic — utterance sharding and the bucketed gradient all-reduce."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, overlap, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from multimodal_av_model_b200 import ddp
    r, _, w = ddp.init_distributed("gloo")
    torch.manual_seed(0)                                   # identical init on both ranks
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    unused = torch.nn.Linear(4, 4)                          # never receives a gradient (cf. cross_attn_visual)
    if rank == 1:
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)                                 # diverge on purpose; broadcast must repair it
    ddp.broadcast_module(model)
    params = list(model.parameters()) + list(unused.parameters())
    red = ddp.GradBucketReducer(params, bucket_bytes=64, overlap=overlap, never_used=list(unused.parameters()))
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 6, generator=g); y_all = torch.randn(8, 3, generator=g)
    lo, hi = ddp.shard_range(8, r, w)
    for step in range(3):
        red.zero_grad()
        loss = ((model(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean()
        loss.backward()
        red.finish()
        # .grad lives inside the communication buckets (no copy in, no copy back)
        assert all(any(p.grad.data_ptr() >= b["flat"].data_ptr() and
                       p.grad.data_ptr() < b["flat"].data_ptr() + 4 * b["flat"].numel() for b in red.buckets)
                   for p in model.parameters())
    grads = [p.grad.clone() for p in model.parameters()]
    # single-process truth: average of the two shard gradients
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    acc = [torch.zeros_like(p) for p in ref.parameters()]
    for rr in range(w):
        a, b = ddp.shard_range(8, rr, w)
        ref.zero_grad()
        ((ref(x_all[a:b]) - y_all[a:b]) ** 2).mean().backward()
        for t, p in zip(acc, ref.parameters()):
            t += p.grad / w
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(grads, acc)) and all(p.grad is None for p in unused.parameters())
    q.put((rank, bool(ok), red.grad_bytes()))
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_bucketed_allreduce_equals_averaged_shard_gradients(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + int(overlap)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_shard_range_partitions_everything():
    sys.path.insert(0, ROOT)
    from multimodal_av_model_b200 import ddp
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            spans = [ddp.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _buffer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from multimodal_av_model_b200 import ddp
    ddp.init_distributed("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv1d(2, 4, 3), torch.nn.BatchNorm1d(4))
    net.train()
    g = torch.Generator().manual_seed(10 + rank)                 # every rank sees its own batches
    for _ in range(3):
        net(torch.randn(5, 2, 9, generator=g))
    mine = net[1].running_mean.clone()
    w_before = net[0].weight.clone()
    ddp.broadcast_buffers(net)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, net[1].running_mean)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    q.put((rank, bool(same), bool(torch.equal(mine, net[1].running_mean)), bool(torch.equal(w_before, net[0].weight)),
           int(net[1].num_batches_tracked)))
    dist.destroy_process_group()


def test_broadcast_buffers_aligns_running_statistics_only():
    """BatchNorm running statistics drift per rank in train mode; broadcast_buffers (used by evaluate()) makes every
    rank adopt rank 0's, and leaves parameters alone."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_buffer_worker, args=(r, 2, 29621, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(same for _, same, _, _, _ in res), res
    assert res[0][2] is True and res[1][2] is False          # rank 0 kept its statistics, rank 1 adopted them
    assert all(w_ok for _, _, _, w_ok, _ in res) and all(n == 3 for *_, n in res)


def _failure_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from multimodal_av_model_b200 import ddp
    r, _, w = ddp.init_distributed("gloo")
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    red = ddp.GradBucketReducer(list(model.parameters()), bucket_bytes=64)
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 6, generator=g); y_all = torch.randn(8, 3, generator=g)
    lo, hi = ddp.shard_range(8, r, w)
    out = {}
    for step, fail_rank in enumerate((None, 1, None)):
        red.zero_grad()
        ok = True
        try:
            h = model[1](model[0](x_all[lo:hi]))
            if fail_rank == rank:
                raise RuntimeError("synthetic failure between forward and backward")
            ((model[2](h) - y_all[lo:hi]) ** 2).mean().backward()
        except RuntimeError:
            ok = False
        red.finish(ok=ok)
        out[step] = [p.grad.clone() for p in model.parameters()]
    # truth: step 1 = rank 0's own gradient on every rank (the failed rank is left out of the average);
    # step 2 = the plain average again (no stale bucket state survived the failed step)
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    per_rank = []
    for rr in range(w):
        a, b = ddp.shard_range(8, rr, w)
        ref.zero_grad()
        ((ref(x_all[a:b]) - y_all[a:b]) ** 2).mean().backward()
        per_rank.append([p.grad.clone() for p in ref.parameters()])
    avg = [(a + b) / 2 for a, b in zip(*per_rank)]
    ok1 = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(out[1], per_rank[0]))
    ok02 = all(torch.allclose(a, b, atol=1e-6) for s_ in (0, 2) for a, b in zip(out[s_], avg))
    q.put((rank, bool(ok1), bool(ok02)))
    dist.destroy_process_group()


def test_failed_rank_is_left_out_and_every_rank_gets_the_same_gradients():
    """a step that raises on one rank must neither desynchronise the collectives nor leak stale
    bucket state into the next step."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_failure_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(a and b for _, a, b in res), res


def _layerdrop_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from multimodal_av_model_b200 import ddp
    r, _, w = ddp.init_distributed("gloo")
    torch.manual_seed(0)
    a, skipped, c = torch.nn.Linear(6, 6), torch.nn.Linear(6, 6), torch.nn.Linear(6, 3)
    params = list(a.parameters()) + list(skipped.parameters()) + list(c.parameters())
    red = ddp.GradBucketReducer(params, bucket_bytes=64)
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 6, generator=g)
    lo, hi = ddp.shard_range(8, r, w)
    res = []
    for step, use in enumerate((False, True, False, True)):      # wav2vec2 LayerDrop: a trainable layer is skipped in some steps
        red.zero_grad()
        h = a(x_all[lo:hi])
        if use:
            h = skipped(h)
        c(h).pow(2).mean().backward()
        red.finish()
        gs = skipped.weight.grad
        res.append(None if gs is None else float(gs.abs().sum()))
    ok = res[0] == 0.0 and res[2] == 0.0 and res[1] > 0 and res[3] > 0 and abs(res[1] - res[3]) < 1e-6
    q.put((rank, bool(ok), res))
    dist.destroy_process_group()


def test_parameter_without_gradient_in_some_steps():
    """Round 2: LayerDrop leaves a trainable layer without a gradient in some steps; the reducer neither raises nor
    desynchronises, the skipped layer's reduced gradient is zero in that step (torch DDP semantics) and right afterwards."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_layerdrop_worker, args=(r, 2, 29651, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
