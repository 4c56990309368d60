"""Training-side plumbing for BASELINE config C5.

The reference trains with threaded ``nn.DataParallel`` (trainer.py:180-186, data_parallel.py): every step it
broadcasts all 21.18 M parameters (84.7 MB) to every GPU and reduces the gradients onto GPU 0.  The B200 way is one
process per GPU with the parameters replicated once and ONE gradient all-reduce per step over NCCL / NVLink-5
(NVSwitch: full bandwidth between any pair, in-switch reduction when NVLS is available) -- bucketed so that the
reduction of the gradients produced first (the last layers) overlaps the rest of the backward pass.

``BucketedGradAllReduce`` is that collective, backend-agnostic (nccl on GPUs, gloo in the CPU tests).  It registers
post-accumulate-grad hooks; a bucket whose last gradient has arrived is flattened and all-reduced asynchronously (on a
side stream under nccl); ``finish()`` waits, averages and scatters the result back into ``.grad``.  Train-mode BN
stays per replica, as in the reference (no SyncBN).

What is native here: the DCN module's backward (deft_dcnv2_backward, dcn_v2._DCNFunction) and the AFE forward / loss
values (deft_afe_forward_scores / deft_afe_loss).  The backward of the plain convolutions, BN and the AFE head is not
written yet: config C5's frames/s is therefore not reported by bench.py.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, params, bucket_bytes=25 * 1024 * 1024, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # buckets in REVERSE registration order: autograd produces the last layers' gradients first
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._where = {}
        for b, ps in enumerate(self.buckets):
            for p in ps:
                self._where[p] = b
        self._pending = [len(b) for b in self.buckets]
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._side = None
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def bytes_per_step(self):
        return sum(p.numel() * p.element_size() for p in self.params)

    def _on_grad(self, p):
        b = self._where[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        ps = self.buckets[b]
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        self._flat[b] = flat
        if self.world == 1:
            return
        if flat.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=flat.device)
            self._side.wait_stream(torch.cuda.current_stream(flat.device))     # the bucket's gradients are complete
            with torch.cuda.stream(self._side):
                self._work[b] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self._work[b] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket, average, write back into .grad; buckets whose hooks never fired (unused parameters) are
        reduced as zeros so that all ranks issue the same collectives."""
        for b, n in enumerate(self._pending):
            if n != 0:
                for p in self.buckets[b]:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(b)
        for b, ps in enumerate(self.buckets):
            if self._work[b] is not None:
                self._work[b].wait()
            flat = self._flat[b]
            if flat.is_cuda and self._side is not None:
                torch.cuda.current_stream(flat.device).wait_stream(self._side)
            if self.world > 1:
                flat.div_(self.world)
            off = 0
            for p in ps:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        self._pending = [len(b) for b in self.buckets]
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)

    def remove(self):
        for h in self._handles:
            h.remove()
