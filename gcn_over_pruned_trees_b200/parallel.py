"""Sentence-sharded data parallelism: one process per GPU, parameters replicated, one gradient all-reduce per step.

The reference is single-device: every sentence's tree, CSR, aggregation and pooling live on one GPU,
so the only exchange in a data-parallel step is the gradient all-reduce (NCCL over NVLink 5 / NVSwitch on the GPU
box, gloo in the CPU tests).  Loss terms are batch means (trainer.py:94-100), so gradients are *averaged*.

Small dense gradients travel in one flat bucket (a single collective: latency-bound), the word-embedding gradient
-- the one large message, [V, emb_dim] -- is reduced in place without a staging copy.  Parameters whose gradient is
None on every rank (deprel_emb in 'regular' mode; ner_emb on SemEval) are skipped consistently because the set is
decided by the architecture, not by the data.
"""
import os

import torch
import torch.distributed as dist

LARGE_NUMEL = 1 << 20


def init_from_env(backend=None):
    """Join the process group torchrun describes (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); no-op for 1 process."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_batch(batch, rank, world):
    """Rows rank::world of a loader batch (the unit of work is a sentence; no tree spans two ranks), re-padded to the
    shard's own longest sentence, as the loader would have padded it (data/loader.py:167-174)."""
    idx = torch.arange(rank, batch[0].shape[0], world)
    width = int((~batch[1][idx]).sum(1).max()) if idx.numel() else 0
    fields = [t[idx][:, :width].contiguous() if t.dim() >= 2 else t[idx] for t in batch[:-1]]
    return tuple(fields) + ([batch[-1][i] for i in idx.tolist()],)


class GradAllReducer(object):
    """Average ``.grad`` of ``params`` over the process group after ``loss.backward()``, before clipping."""

    def __init__(self, params, group=None):
        seen, self.params = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        self.group = group
        self._flat = None

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def grads(self):
        """The gradient tensors reduce() would act on right now (None entries skipped)."""
        return [p.grad for p in self.params if p.grad is not None]

    def reduce(self, grads=None):
        """Average gradients in place.  ``grads`` pins the tensors (a CUDA-graph engine passes the buffers its graph
        writes, which need not be what ``p.grad`` currently points at); default: the parameters' current ``.grad``."""
        world = self.world
        if world == 1:
            return
        grads = self.grads() if grads is None else grads
        small = [g for g in grads if g.numel() < LARGE_NUMEL]
        large = [g for g in grads if g.numel() >= LARGE_NUMEL]
        handles = []
        for g in large:                      # in place, asynchronously, while the bucket is being packed
            handles.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        if small:
            n = sum(g.numel() for g in small)
            if self._flat is None or self._flat.numel() != n or self._flat.device != small[0].device:
                self._flat = torch.empty(n, dtype=small[0].dtype, device=small[0].device)
            views = []
            off = 0
            for g in small:
                k = g.numel()
                views.append(self._flat[off:off + k].view_as(g))
                off += k
            torch._foreach_copy_(views, small)
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
            self._flat.div_(world)
            torch._foreach_copy_(small, views)
        for h, g in zip(handles, large):
            h.wait()
            g.div_(world)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device):
    """Max of a python float over all ranks (step time is the slowest rank's)."""
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class PeerExchange(object):
    """The exchange regions of all ranks of one node, mapped into this process (csrc/dp.cu, K8).

    Construction is collective: every rank allocates its region, the 64-byte cudaIpc handles travel through the
    process group once (all_gather_object), peers are opened, and a barrier guarantees that every region is
    initialised before the first push.  After that the step path never calls NCCL: ranks meet through flags in
    peer memory, inside the kernels.
    """

    def __init__(self, cap_rows, emb_dim, vocab, n_flat, group=None):
        from . import ops
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise ValueError('PeerExchange is a single-node (<= 8 GPU) exchange')
        self.multicast = None           # NVSwitch multicast address of the W regions, when the fabric offers one
        self.symm = None
        self.region = None
        if self.world > 1 and os.environ.get('GPT_DP_SYMM', '1') != '0':
            self._try_symmetric_memory(cap_rows, emb_dim, vocab, n_flat, group)
        if self.region is None:
            # cudaMalloc + cudaIpc: every rank allocates, the 64-byte handles travel through the process group once
            self.region = ops.ExchangeRegion(self.world, cap_rows, emb_dim, vocab, n_flat)
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, self.region.handle, group=group)
            self.ptrs = [self.region.ptr if r == self.rank else ops.open_peer_region(handles[r])
                         for r in range(self.world)]
        self.shape = self.region.shape
        self.n_partials = self.region.n_partials
        if self.world > 1:
            dist.barrier(group=group)

    def _try_symmetric_memory(self, cap_rows, emb_dim, vocab, n_flat, group):
        """Regions in torch symmetric memory (plumbing: cuMemCreate + fabric handles + cuMulticast*): same peer pointers
        as the cudaIpc path, plus a multicast address when the devices sit behind an NVSwitch.  Collective; every rank
        must reach the same verdict, so the outcome is agreed on with an all-reduce before it is used."""
        from . import ops
        ok, buf, hdl = 1, None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = ops.ExchangeRegion.bytes_for(self.world, cap_rows, emb_dim, vocab, n_flat)
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=torch.device('cuda', torch.cuda.current_device()))
            hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world or ptrs[self.rank] != buf.data_ptr():
                ok = 0
        except Exception:               # no symmetric-memory support in this build / on this fabric
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) != 1:
            return
        self.symm = (buf, hdl)
        self.ptrs = ptrs
        self.region = ops.ExchangeRegion(self.world, cap_rows, emb_dim, vocab, n_flat, buffer=buf)
        mc = int(hdl.multicast_ptr) if getattr(hdl, 'has_multicast_support', False) else 0
        use = torch.tensor([1 if (mc and os.environ.get('GPT_DP_MULTICAST', '1') != '0') else 0], dtype=torch.int32,
                           device='cuda')
        dist.all_reduce(use, op=dist.ReduceOp.MIN, group=group)
        self.multicast = mc if int(use) == 1 else None
