"""Utterance-sharded data parallelism: one process per GPU, gradients all-reduced with NCCL over NVLink.

The reference is single-process (no torch.distributed anywhere); north_star adds data parallelism
by batch item.  Every hot-path op is per-sample except the fusion resample (batch-max length) and InfoNCE
(mixes frames across the local batch), so an N-GPU step equals "N independent reference micro-batches with
averaged gradients"  — that is what tests/test_ddp_cpu.py checks on gloo.

GradBucketReducer: parameters are packed (reverse registration order = backward order) into flat fp32 buckets;
a post-accumulate-grad hook copies each gradient into its bucket and, when a bucket is complete, launches an
asynchronous all_reduce on a side stream so communication overlaps the rest of backward.  finish() waits,
divides by world size and copies the averages back into .grad.  Parameters that received no gradient in a step
(e.g. cross_attn_visual, never used by the reference) are skipped: their bucket slots stay zero on every rank.
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
    def __init__(self, params, bucket_bytes=64 << 20, group=None, overlap=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.overlap = overlap and self.world > 1
        self.buckets = []          # list of dict(flat, items=[(param, offset, numel)], pending, work)
        self._slot = {}
        if not self.params:
            return
        dev = self.params[0].device
        cur, cur_bytes = [], 0
        for p in reversed(self.params):          # backward produces gradients roughly in reverse order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._add_bucket(cur, dev)
                cur, cur_bytes = [], 0
        if cur:
            self._add_bucket(cur, dev)
        self.stream = torch.cuda.Stream(device=dev) if (dev.type == "cuda" and self.overlap) else None
        self._handles = []
        if self.overlap:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _add_bucket(self, plist, dev):
        total = sum(p.numel() for p in plist)
        b = dict(flat=torch.zeros(total, dtype=torch.float32, device=dev), items=[], pending=0, work=None, ready=set())
        off = 0
        for p in plist:
            b["items"].append((p, off, p.numel()))
            self._slot[p] = (len(self.buckets), off)
            off += p.numel()
        self.buckets.append(b)

    def _launch(self, b):
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(b["flat"].device))
            with torch.cuda.stream(self.stream):
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, p):
        bi, off = self._slot[p]
        b = self.buckets[bi]
        b["flat"][off:off + p.numel()].copy_(p.grad.reshape(-1))
        b["ready"].add(p)
        if len(b["ready"]) == len(b["items"]) and b["work"] is None:
            self._launch(b)

    def finish(self):
        """Call after backward(): completes every bucket and writes averaged gradients back."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b["work"] is None:                 # hooks disabled, or some params got no gradient this step
                for p, off, n in b["items"]:
                    if p not in b["ready"]:
                        if p.grad is not None:
                            b["flat"][off:off + n].copy_(p.grad.reshape(-1))
                        else:
                            b["flat"][off:off + n].zero_()
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        inv = 1.0 / self.world
        for b in self.buckets:
            for p, off, n in b["items"]:
                if p.grad is not None:
                    p.grad.copy_((b["flat"][off:off + n] * inv).view_as(p.grad))
            b["work"] = None
            b["ready"] = set()

    def grad_bytes(self):
        return sum(b["flat"].numel() * 4 for b in self.buckets)

    def close(self):
        for h in getattr(self, "_handles", []):
            h.remove()
        self._handles = []
