"""Utterance-sharded data parallelism: one process per GPU, gradients all-reduced with NCCL over NVLink.

The reference is single-process (no torch.distributed anywhere); north_star adds data parallelism
by batch item.  Every hot-path op is per-sample except the fusion resample (batch-max length) and InfoNCE
(mixes frames across the local batch), so an N-GPU step equals "N independent reference micro-batches with
averaged gradients"  — that is what tests/test_ddp_cpu.py checks on gloo.

GradBucketReducer: parameters are packed (reverse registration order = backward order) into flat fp32 buckets and
every `.grad` IS a view into its bucket, so autograd accumulates straight into the communication buffer: no
per-parameter copy in, no copy back.  A post-accumulate-grad hook only counts; when a bucket is complete (and
every earlier bucket has been launched — the collective order is the bucket index on every rank) an asynchronous
all_reduce(SUM) goes out on a side stream under the rest of backward.  finish() launches what is left, waits, and
scales all buckets with ONE foreach multiply by 1 / (number of ranks whose step succeeded) — that count travels
in a spare slot of the last bucket, so a rank whose forward/backward raised contributes zeros and every rank
still applies the same update (see finish()).  Parameters that are outside every step's graph by construction
(cross_attn_visual, never used by the reference) are declared `never_used`: no slot, `.grad` stays None, as in a
single process.  A parameter that happens to get no gradient in ONE step (wav2vec2's LayerDrop skips a trainable layer
with probability 0.1 per call) contributes zeros to the sum, as under torch's DistributedDataParallel; its bucket, and
the ones behind it in launch order, then leave from finish() instead of from the hooks.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*). Returns
    (rank, local_rank, world_size); a no-op single-process triple when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(n_items, rank, world):
    """Contiguous utterance shard [lo, hi) of rank (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_module(module, src=0, group=None):
    """Make parameters and buffers identical on every rank (initial sync; also the lazily created
    projection layer of trainer.py:105-106 and the BatchNorm running stats of the frozen visual encoder)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def broadcast_buffers(module, src=0, group=None):
    """Buffers only (BatchNorm running statistics, num_batches_tracked): every rank adopts rank `src`'s values, as
    torch's DistributedDataParallel(broadcast_buffers=True) keeps them.  The frozen visual encoder runs in train mode
    (trainer.py:54), so its running statistics follow each rank's own batches; evaluation and checkpoints must not
    depend on which rank they came from."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in module.buffers():
        dist.broadcast(t.data, src=src, group=group)


class GradBucketReducer:
    def __init__(self, params, bucket_bytes=None, group=None, overlap=True, never_used=()):
        if bucket_bytes is None:
            bucket_bytes = int(os.environ.get("AVCTC_BUCKET_MB", "24")) << 20
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # `never_used`: parameters that are outside every step's graph by construction (cross_attn_visual, which the
        # reference builds and never calls, fusion_module.py:14,61): they get no slot, their .grad stays None
        skip = {id(p) for p in never_used}
        self.params = [p for p in params if p.requires_grad and id(p) not in skip]
        self.overlap = overlap and self.world > 1
        self.buckets = []          # dict(flat, items=[(param, offset, numel)], ready=set(), work)
        self._slot = {}
        self._next = 0             # buckets [0, _next) have been launched this step
        self._handles = []
        self.stream = None
        if not self.params:
            return
        dev = self.params[0].device
        # Buckets in backward order.  The gradients produced LAST (the first trainable layer) cannot hide behind any
        # compute, so the final `bucket_bytes` of the order are cut into quarter-size buckets: what is still on the wire
        # when backward returns is then a small bucket, not a full one (measured at N=8: 1.3 ms of the step were the
        # exposed tail with uniform 24 MB buckets).
        total = sum(p.numel() * 4 for p in self.params)
        cur, cur_bytes, seen = [], 0, 0
        for p in reversed(self.params):          # backward produces gradients roughly in reverse order
            cur.append(p)
            cur_bytes += p.numel() * 4
            seen += p.numel() * 4
            limit = bucket_bytes if total - seen > bucket_bytes else max(bucket_bytes // 4, 1)
            if cur_bytes >= limit:
                self._add_bucket(cur, dev)
                cur, cur_bytes = [], 0
        if cur:
            self._add_bucket(cur, dev)
        # spare slot behind the last bucket: the number of ranks whose step succeeded
        last = self.buckets[-1]
        flat = torch.zeros(last["flat"].numel() + 1, dtype=torch.float32, device=dev)
        last["flat"] = flat
        self._ok = flat[-1:]
        self._scale = torch.ones((), dtype=torch.float32, device=dev)
        self.stream = torch.cuda.Stream(device=dev) if (dev.type == "cuda" and self.overlap) else None
        for p in self.params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.zero_grad()

    def _add_bucket(self, plist, dev):
        total = sum(p.numel() for p in plist)
        b = dict(flat=torch.zeros(total, dtype=torch.float32, device=dev), items=[], work=None, ready=set())
        off = 0
        for p in plist:
            b["items"].append((p, off, p.numel()))
            self._slot[p] = len(self.buckets)
            off += p.numel()
        self.buckets.append(b)

    # -------------------------------------------------------------------------------------------- step protocol
    def reset(self):
        """Forget a step that did not reach finish(): wait for collectives that are still in flight (every rank
        launched them, so they complete), clear the per-step bookkeeping."""
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
            b["ready"] = set()
        self._next = 0

    def zero_grad(self):
        """Start of a step (instead of optimizer.zero_grad()): one fill per bucket, and every parameter's .grad is
        (again) the view into its bucket that autograd accumulates into."""
        self.reset()
        for b in self.buckets:
            b["flat"].zero_()
            flat = b["flat"]
            for p, off, n in b["items"]:
                g = p.grad
                if g is None or g.data_ptr() != flat.data_ptr() + 4 * off:
                    p.grad = flat[off:off + n].view_as(p)

    def _launch(self, b):
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(b["flat"].device))
            with torch.cuda.stream(self.stream):
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _launch_ready(self):
        last = len(self.buckets) - 1            # the last bucket carries the ok count: it leaves from finish()
        while self._next < last and len(self.buckets[self._next]["ready"]) == len(self.buckets[self._next]["items"]):
            self._launch(self.buckets[self._next])
            self._next += 1

    def _on_grad(self, p):
        b = self.buckets[self._slot[p]]
        b["ready"].add(p)
        if self.overlap and self.world > 1:
            self._launch_ready()

    def finish(self, ok=True):
        """Call after backward() — also when the step raised (ok=False): every rank launches every bucket exactly
        once per step, in index order, whatever happened to its own step.  A failed rank contributes zeros for the
        buckets that had not left yet and 0 to the ok count; gradients are divided by the ok count, so the ranks
        whose step worked are averaged over themselves and ALL ranks (the failed one included) end up with the same
        gradients and must apply the same optimizer step."""
        if self.world == 1 or not self.buckets:
            return
        self._ok.fill_(1.0 if ok else 0.0)
        while self._next < len(self.buckets):
            b = self.buckets[self._next]
            if not ok:
                b["flat"][:sum(n for _, _, n in b["items"])].zero_()
            self._launch(b)
            self._next += 1
        for b in self.buckets:
            b["work"].wait()
            b["work"] = None
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        torch.reciprocal(self._ok.clamp(min=1.0).reshape(()), out=self._scale)
        torch._foreach_mul_([b["flat"] for b in self.buckets], self._scale)
        for b in self.buckets:
            b["ready"] = set()
        self._next = 0

    def grad_bytes(self):
        return sum(b["flat"].numel() * 4 for b in self.buckets)

    def close(self):
        for h in getattr(self, "_handles", []):
            h.remove()
        self._handles = []
