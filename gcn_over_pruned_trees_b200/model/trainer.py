"""Trainer API of the reference (/root/reference/model/trainer.py:14-127) on the B200 model.

``GCNTrainer(opt, emb_matrix)``, ``.update(batch) -> loss tensor (caller runs backward/clip/step, train.py:220-227)``,
``.predict(batch, unsort=True) -> (predictions, probs, loss)``, ``.save/.load`` with the same checkpoint dict
``{'model': state_dict, 'config': opt}``, ``.update_lr``, ``.get_deprel_emb``; attributes ``model, criterion,
parameters, optimizer, opt``.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .gcn import GCNClassifier
try:
    from .. import torch_utils
except ImportError:
    from gcn_over_pruned_trees_b200 import torch_utils


class Trainer(object):
    def __init__(self, opt, emb_matrix=None):
        raise NotImplementedError

    def update(self, batch):
        raise NotImplementedError

    def predict(self, batch):
        raise NotImplementedError

    def update_lr(self, new_lr):
        torch_utils.change_lr(self.optimizer, new_lr)

    def load(self, filename):
        try:
            checkpoint = torch.load(filename, map_location=None if torch.cuda.is_available() else 'cpu')
        except BaseException:
            print("Cannot load model from {}".format(filename))
            exit()
        self.model.load_state_dict(checkpoint['model'])
        self.opt = checkpoint['config']

    def save(self, filename, epoch):
        params = {'model': self.model.state_dict(), 'config': self.opt}
        try:
            torch.save(params, filename)
            print("model saved to {}".format(filename))
        except BaseException:
            print("[Warning: Saving failed... continuing anyway.]")


def unpack_batch(batch, cuda):
    """Loader tuple -> (inputs, labels, tokens, head, subj_pos, obj_pos, lens); TACRED 10-tuple or SemEval 9-tuple."""
    if cuda:
        inputs = [b.cuda(non_blocking=True) for b in batch[:-2]]
        labels = batch[-2].cuda(non_blocking=True)
    else:
        inputs = list(batch[:-2])
        labels = batch[-2]
    tokens = batch[0]
    off = 5 if len(batch) >= 10 else 4
    head, subj_pos, obj_pos = batch[off], batch[off + 1], batch[off + 2]
    lens = batch[1].eq(0).long().sum(1).squeeze()
    return inputs, labels, tokens, head, subj_pos, obj_pos, lens


class GCNTrainer(Trainer):
    def __init__(self, opt, emb_matrix=None):
        self.opt = opt
        self.emb_matrix = emb_matrix
        self.model = GCNClassifier(opt, emb_matrix=emb_matrix)
        self.criterion = nn.CrossEntropyLoss()
        self.parameters = [p for p in self.model.parameters() if p.requires_grad]
        if opt['cuda']:
            self.model.cuda()
            self.criterion.cuda()
        self.optimizer = torch_utils.get_optimizer(opt['optim'], self.parameters, opt['lr'],
                                                    capturable=bool(opt['cuda']))
        # update() / loss.backward() / optimizer.step() on captured graphs where the configuration allows it
        # (engine.FastUpdate); False, or GPT_FAST_UPDATE=0, keeps every call on the per-op autograd path
        self.fast_update = os.environ.get('GPT_FAST_UPDATE', '1') != '0'
        self._fast = None
        # per-op path: classifier head + loss through K6 (ops.head_loss) instead of nn.Linear / CrossEntropyLoss kernels
        self.fused_head = os.environ.get('GPT_FUSED_HEAD', '1') != '0'

    def _loss(self, logits, pooling_output, labels):
        loss = self.criterion(logits, labels)
        if self.opt.get('conv_l2', 0) > 0:
            loss = loss + self.model.conv_l2() * self.opt['conv_l2']
        if self.opt.get('pooling_l2', 0) > 0:
            loss = loss + self.opt['pooling_l2'] * (pooling_output ** 2).sum(1).mean()
        return loss

    def update(self, batch):
        fast = self._fast_path()
        if fast is not None:
            return fast.update(batch)
        inputs, labels = unpack_batch(batch, self.opt['cuda'])[:2]
        return self._forward_loss(inputs, labels)

    def _forward_loss(self, inputs, labels):
        """Forward + loss of the per-op path (also what engine.GraphedTrainStep captures)."""
        if self.fused_head and self.opt['cuda'] and self.model.training and torch.is_grad_enabled():
            # the head and its loss in K6's two launches (still autograd: the caller's loss.backward() works as before)
            fused = self.model.loss_fused(inputs, labels, self.opt.get('pooling_l2', 0))
            if fused is not None:
                loss = fused[0]
                if self.opt.get('conv_l2', 0) > 0:
                    loss = loss + self.model.conv_l2() * self.opt['conv_l2']
                return loss
        logits, pooling_output = self.model(inputs)
        return self._loss(logits, pooling_output, labels)

    def _predict_path(self, batch):
        if not (self.fast_update and self.opt['cuda']) or batch[0].dim() != 2:
            return None
        if self.model.gcn_model.gcn.injected_masks is not None:
            return None
        fp = getattr(self, '_fused_predict', None)
        if fp is None:
            try:
                from ..engine import FusedPredict, FusedTrainStep
            except ImportError:
                from gcn_over_pruned_trees_b200.engine import FusedPredict, FusedTrainStep
            if FusedTrainStep.unsupported_reason(self, training=False) is not None:
                self._fused_predict = False
                return None
            fp = self._fused_predict = FusedPredict(self)
        return fp or None

    def _fast_path(self):
        """engine.FastUpdate when this call can take it: a training-mode forward with autograd on, on a configuration
        FusedTrainStep covers (regular GCN, plain SGD, CUDA); None otherwise."""
        if not (self.fast_update and self.opt['cuda'] and self.model.training and torch.is_grad_enabled()):
            return None
        if self.model.gcn_model.gcn.injected_masks is not None:
            return None
        if self._fast is None:
            try:
                from ..engine import FastUpdate
            except ImportError:
                from gcn_over_pruned_trees_b200.engine import FastUpdate
            if FastUpdate.unsupported_reason(self) is not None:
                self.fast_update = False
                return None
            self._fast = FastUpdate(self)
        elif self._fast.trainer.optimizer is not self.optimizer:
            return None                     # the caller installed an optimizer of its own
        return self._fast

    def predict(self, batch, unsort=True):
        self.model.eval()
        fused = self._predict_path(batch)
        if fused is not None:           # one graph replay + one device-to-host copy (engine.FusedPredict, K11)
            return fused.predict(batch, unsort)
        inputs, labels = unpack_batch(batch, self.opt['cuda'])[:2]
        orig_idx = batch[-1]
        with torch.no_grad():
            logits, _ = self.model(inputs)
            loss = self.criterion(logits, labels)
            probs = F.softmax(logits, 1).cpu().numpy().tolist()
            predictions = np.argmax(logits.cpu().numpy(), axis=1).tolist()
        if unsort:
            _, predictions, probs = [list(t) for t in zip(*sorted(zip(orig_idx, predictions, probs)))]
        return predictions, probs, loss.item()

    def train_step(self, batch, reducer=None):
        """One whole optimisation step (zero_grad, forward, loss, backward, clip, optimizer step -- the five calls of
        train.py:213-227) as a single CUDA-graph replay; returns the loss as a device scalar.  Not part of the
        reference API: the reference-compatible path is update() + caller-owned backward/clip/step."""
        if getattr(self, '_graphed', None) is None:
            try:
                from ..engine import FusedTrainStep, GraphedTrainStep
            except ImportError:
                from gcn_over_pruned_trees_b200.engine import FusedTrainStep, GraphedTrainStep
            single = reducer is None or reducer.world == 1
            if single and FusedTrainStep.unsupported_reason(self) is None:
                self._graphed = FusedTrainStep(self)         # library kernels only, no autograd
            else:
                self._graphed = GraphedTrainStep(self, reducer=reducer)
        return self._graphed(batch)

    def get_deprel_emb(self):
        return self.model.get_deprel_emb()
