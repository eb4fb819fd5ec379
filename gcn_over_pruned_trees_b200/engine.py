"""Whole-step execution engine: one CUDA-graph replay per training step.

The reference drives a step from Python as five separate calls (/root/reference/train.py:213-227: zero_grad,
``trainer.update``, ``loss.backward()``, ``clip_grad_norm_``, ``optimizer.step()``), and its forward leaves the device
six times (gcn.py:96-110).  At TACRED batch sizes (50 sentences, ~1 800 tokens) the kernels of this package need a
few hundred microseconds in total, so launch latency and interpreter time dominate an eager step.  Because the new
forward never synchronises with the host and every buffer has a static shape for a given (batch, width), the whole
step -- K1 prune, embeddings, both GCN layers, pooling, MLP, loss, the complete backward, global-norm clipping and
the SGD update -- is captured once per batch shape into a CUDA graph and replayed.

``GraphedTrainStep`` keeps the reference's semantics: same parameters, same ``.grad`` tensors (so checkpoints and
the trainer API keep working), same loss value.  Data-parallel runs split the step into two graphs around the
gradient all-reduce (backward | all-reduce | clip + update).
"""
import torch

from .model.trainer import unpack_batch


class GraphedTrainStep(object):
    """``step(batch) -> loss tensor`` (device scalar, valid until the next call) for loader-shaped batches.

    batch tensors may live on the host (pinned or not) or on the device; they are copied into static device
    buffers, which is the only per-step work outside the graph replay.
    """

    def __init__(self, trainer, max_grad_norm=None, reducer=None, warmup=3):
        self.trainer = trainer
        self.model = trainer.model
        self.opt = trainer.opt
        self.max_grad_norm = trainer.opt['max_grad_norm'] if max_grad_norm is None else max_grad_norm
        self.reducer = reducer if (reducer is not None and reducer.world > 1) else None
        self.warmup = warmup
        self.params = list(self.model.parameters())
        self.sparse = self._make_sparse_state()
        self._graphs = {}       # (B, T, n_fields) -> dict(static inputs, graphs, loss)
        self._lr = self.trainer.optimizer.param_groups[0]['lr']
        self._seen = {}
        self.replays = 0
        self.kernels_per_replay = {}

    def _make_sparse_state(self):
        """Row-sparse word-embedding update (csrc/embed.cu) when it is arithmetic-identical to the dense one: a
        single-process run with the plain SGD the reference ships (no momentum / weight decay / nesterov)."""
        from . import ops
        opt = self.trainer.optimizer
        emb = self.model.gcn_model.emb.weight
        gcn = self.model.gcn_model.gcn
        plain_sgd = isinstance(opt, torch.optim.SGD) and all(
            g.get('momentum', 0) == 0 and g.get('weight_decay', 0) == 0 and not g.get('nesterov', False) and
            not g.get('maximize', False) for g in opt.param_groups)
        if not (plain_sgd and emb.requires_grad and emb.is_cuda and self.reducer is None) or self.opt.get('rnn', False):
            return None
        state = ops.SparseEmbeddingState(emb.data, self.opt['topn'])
        gcn.sparse_embedding = state
        self.emb_weight = emb
        return state

    # -- the step itself, used both for eager warm-up and under capture --------------------------------------------
    def _fwd_bwd(self, inputs, labels):
        # grads are re-created by every backward (no zero-fill kernels, no accumulate-adds); under capture they live
        # in the graph's private pool at fixed addresses
        self.trainer.optimizer.zero_grad(set_to_none=True)
        logits, pooling_output = self.model(inputs)
        loss = self.trainer._loss(logits, pooling_output, labels)
        loss.backward()
        return loss.detach()

    def _update(self):
        if self.sparse is None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
            self.trainer.optimizer.step()
            return
        from . import ops
        st = self.sparse
        dense = [p.grad for p in self.params if p.grad is not None and p is not self.emb_weight]
        st.sq.zero_()
        ops.embed_rows_sqnorm(st)                                   # word-embedding share of the global norm
        total_sq = st.sq + torch.stack(torch._foreach_norm(dense)).pow(2).sum()
        coef = (self.max_grad_norm / (total_sq.sqrt() + 1e-6)).clamp(max=1.0)   # clip_grad_norm_ semantics
        torch._foreach_mul_(dense, coef.reshape(()))
        self.trainer.optimizer.step()                               # emb.weight.grad is None here: skipped
        ops.embed_rows_sgd(st, self.emb_weight.data, total_sq, self.max_grad_norm,
                           self.trainer.optimizer.param_groups[0]['lr'])

    def _eager(self, inputs, labels):
        loss = self._fwd_bwd(inputs, labels)
        if self.reducer is not None:
            self.reducer.reduce()
        self._update()
        return loss

    def _capture(self, key, inputs, labels):
        from . import _lib
        static_in = [torch.empty_like(t) for t in inputs]
        static_lab = torch.empty_like(labels)
        for s, t in zip(static_in, inputs):
            s.copy_(t)
        static_lab.copy_(labels)
        entry = {'inputs': static_in, 'labels': static_lab}
        torch.cuda.synchronize()
        n0 = _lib.lib().gpt_launch_count()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            loss = self._fwd_bwd(static_in, static_lab)
            if self.reducer is None:
                self._update()
        entry['g1'], entry['loss'] = g1, loss
        entry['grads'] = [p.grad for p in self.reducer.params if p.grad is not None] if self.reducer else None
        if self.reducer is not None:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._update()
            entry['g2'] = g2
        self.kernels_per_replay[key] = int(_lib.lib().gpt_launch_count() - n0)
        self._graphs[key] = entry
        return entry

    def __call__(self, batch):
        lr = self.trainer.optimizer.param_groups[0]['lr']
        if lr != self._lr:              # the learning rate is baked into the captured update: re-capture
            self._graphs.clear()
            self._seen = {k: self.warmup for k in self._seen}
            self._lr = lr
        fields, labels = batch[:-2], batch[-2]
        key = (tuple(fields[0].shape), len(fields))
        entry = self._graphs.get(key)
        if entry is None:
            inputs, labels = unpack_batch(batch, self.opt['cuda'])[:2]
            seen = self._seen.get(key, 0)
            self._seen[key] = seen + 1
            if seen < self.warmup:      # the first steps of a new shape run eagerly (allocates .grad, warms caches)
                return self._eager(inputs, labels)
            entry = self._capture(key, inputs, labels)
        else:                           # host (pinned) or device source, straight into the static buffers
            for s, t in zip(entry['inputs'], fields):
                s.copy_(t, non_blocking=True)
            entry['labels'].copy_(labels, non_blocking=True)
        entry['g1'].replay()
        if self.reducer is not None:
            self.reducer.reduce(entry['grads'])     # the buffers this shape's graph writes
            entry['g2'].replay()
        self.replays += 1
        return entry['loss']

    def launches_per_replay(self, batch):
        """Number of this library's kernels inside one replay for the batch's shape (0 before capture)."""
        key = (tuple(batch[0].shape), len(batch) - 2)
        return self.kernels_per_replay.get(key, 0)
