"""ORACLE (test infrastructure, not product code) -- CPU/PyTorch-fp32 restatement of the reference's GCN classifier
(``adj_type`` 'regular', 'full_deprel', 'diagonal_deprel') and of the loss ``GCNTrainer.update`` builds.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.

It deliberately keeps the reference's *dense* formulation -- a ``[B,T,T]`` float adjacency built sentence by
sentence on the host, ``adj.bmm(h)``, two ``Linear`` calls per layer, three ``masked_fill`` + ``max`` passes --
so that timing it measures the reference's CPU algorithm, and so that the CUDA path (CSR gather, one GEMM per
layer) is checked against an independent formulation.

Also restated: the fork's relation-aware layers ``adj_type='full_deprel'`` and
``'diagonal_deprel'`` with edge dropout, relation forgetting, ``deprel_max_depth``, ``deprel_directed`` and
``deprel_self_loop``.

Reference lines restated:
  module/parameter layout      /root/reference/model/gcn.py:15-22, 38-68, 128-176
  relation-aware layers        /root/reference/model/gcn.py:272-294 (diagonal), 296-386 + 400-434 (full),
                               436-449 (edge dropout), 451-470 (relation forgetting)
  embedding concat + in_drop   /root/reference/model/gcn.py:235-247
  BiLSTM encoder (C-GCN)       /root/reference/model/gcn.py:186-197, 250-253, 485-492
  adjacency binarise/denom     /root/reference/model/gcn.py:260-265
  layer loop                   /root/reference/model/gcn.py:266-271, 390-393
  pooling                      /root/reference/model/gcn.py:116-122, 473-483
  loss                         /root/reference/model/trainer.py:93-100

Pinning: ``tests/golden/make_golden.py`` and ``make_deprel_golden.py`` run the real reference (imported from
/root/reference) on seeded inputs and store logits / loss / gradients; ``tests/test_oracle_golden.py`` checks this
module against them.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import tree_oracle

NEG_FILL = -1e12       # /root/reference/utils/constant.py:35
N_POS, N_NER, N_DEPREL = 47, 15, 85


class _DenseGCN(nn.Module):
    def __init__(self, opt, tables):
        super().__init__()
        self.opt = opt
        self.emb, self.pos_emb, self.ner_emb, self.deprel_emb = tables
        self.tacred = opt['dataset'] == 'tacred'
        width = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if self.tacred else 0)
        if opt.get('rnn', False):
            self.rnn = nn.LSTM(width, opt['rnn_hidden'], opt['rnn_layers'], batch_first=True,
                               dropout=opt['rnn_dropout'], bidirectional=True)
            width = 2 * opt['rnn_hidden']
        hidden = opt['hidden_dim']
        self.adj_type = opt.get('adj_type', 'regular')
        if self.adj_type == 'diagonal_deprel':        # gcn.py:153-155: no per-layer weights at all
            self.preprocessor = nn.Linear(width, hidden)
        elif self.adj_type == 'full_deprel':          # gcn.py:164-167: ONE Linear(in, D*H) shared by every layer
            self.W = nn.Linear(width, opt['deprel_emb_dim'] * hidden)
        else:
            self.W = nn.ModuleList(nn.Linear(width if l == 0 else hidden, hidden)
                                   for l in range(opt['num_layers']))

    # ---- relation-aware layers ------------------------------------------------------------------------------
    def _edge_keep(self, a, name, masks):
        """gcn.py:436-449: Bernoulli(edge_keep_prob) over the DENSE matrix, drawn per direction and per layer."""
        if masks is not None and name in masks:
            return a * masks[name]
        p = self.opt.get('edge_keep_prob', 1.0)
        if self.training and p < 1.0:
            return torch.empty_like(a).bernoulli_(p) * a
        return a

    def _forget(self, e, name, masks):
        """gcn.py:451-470: with probability 1 - deprel_keep_prop a token's relation vector becomes all ones."""
        if masks is not None and name in masks:
            keep = masks[name]
        else:
            p = self.opt.get('deprel_keep_prop', 1.0)
            if not (self.training and p < 1.0):
                return e
            keep = torch.empty((e.size(0), e.size(1), 1)).bernoulli_(p)
        return torch.where(keep.expand_as(e) == 1, e, torch.ones_like(e))

    def _relation_layer(self, adj, x, deprel, l, masks):
        """One layer's pre-normalisation sum for the two relation-aware modes.  ``adj`` holds the relation ids:
        (0,42) parent->child, (42,84) child->parent, 84 self loop (tree.py:184-192)."""
        fwd = ((adj > 0) & (adj < 42)).float()
        rev = ((adj > 42) & (adj < 84)).float()
        e_f = self.deprel_emb(deprel)                  # keyed by the token's OWN incoming relation ...
        e_r = self.deprel_emb(deprel + 42)             # ... in both directions (gcn.py:315, 349)
        e_s = self.deprel_emb.weight[84]
        if self.adj_type == 'diagonal_deprel':         # gcn.py:272-294: no dropout of edges, no forgetting
            return fwd.bmm(e_f * x) + rev.bmm(e_r * x) + x * e_s
        D, H = self.opt['deprel_emb_dim'], self.opt['hidden_dim']
        w = self.W.weight.reshape(D, -1, H)            # a reshape of [D*H, in], NOT a permute (gcn.py:301)
        b = self.W.bias.reshape(D, H)
        deep = l >= self.opt['deprel_max_depth']       # gcn.py:324-325, 355-356, 376-379

        def traverse(e):                               # gcn.py:400-415
            return torch.einsum('bnd,bnk,dkh->bnh', e, x, w) + e @ b

        fwd = self._edge_keep(fwd, 'edge_f%d' % l, masks)
        e_f = self._forget(e_f, 'forget_f%d' % l, masks)
        total = fwd.bmm(traverse(torch.ones_like(e_f) if deep else e_f))
        if not self.opt['deprel_directed']:
            rev = self._edge_keep(rev, 'edge_r%d' % l, masks)
            e_r = self._forget(e_r, 'forget_r%d' % l, masks)
            total = total + rev.bmm(traverse(torch.ones_like(e_r) if deep else e_r))
        if self.opt['deprel_self_loop']:               # gcn.py:369-386, 417-434: every token, kept or not
            e = torch.ones_like(e_s) if deep else e_s
            total = total + x @ torch.einsum('d,dkh->kh', e, w) + e @ b
        return total

    def conv_l2(self):
        return sum(p.pow(2).sum() for lin in self.W for p in (lin.weight, lin.bias))

    def _drop(self, x, p, name, masks):
        if masks is not None and name in masks:       # injected, already scaled by 1/(1-p)
            return x * masks[name]
        return F.dropout(x, p, self.training)

    def forward(self, adj, inputs, masks=None):
        if self.tacred:
            words, pad, pos, ner, deprel, head, subj_pos, obj_pos = inputs
        else:
            words, pad, pos, deprel, head, subj_pos, obj_pos = inputs
            ner = None
        parts = [self.emb(words)]
        if self.opt['pos_dim'] > 0:
            parts.append(self.pos_emb(pos))
        if self.opt['ner_dim'] > 0 and self.tacred:
            parts.append(self.ner_emb(ner))
        x = self._drop(torch.cat(parts, dim=2), self.opt['input_dropout'], 'in', masks)
        if self.opt.get('rnn', False):
            lens = pad.eq(0).long().sum(1)
            zeros = torch.zeros(2 * self.opt['rnn_layers'], words.size(0), self.opt['rnn_hidden'])
            packed = nn.utils.rnn.pack_padded_sequence(x, lens.cpu(), batch_first=True, enforce_sorted=False)
            y, _ = self.rnn(packed, (zeros, zeros))
            y, _ = nn.utils.rnn.pad_packed_sequence(y, batch_first=True)
            x = self._drop(y, self.opt['rnn_dropout'], 'rnn', masks)
        if self.adj_type == 'diagonal_deprel':
            x = self.preprocessor(x)                   # gcn.py:255-257
        a = (adj != 0).float()
        denom = a.sum(2, keepdim=True) + 1
        not_in_tree = (a.sum(2) + a.sum(1)).eq(0).unsqueeze(2)
        if self.opt.get('no_adj', False):
            a = torch.zeros_like(a)
        last = self.opt['num_layers'] - 1
        for l in range(self.opt['num_layers']):
            if self.adj_type == 'regular':
                lin = self.W[l]
                z = (lin(a.bmm(x)) + lin(x)) / denom  # bias enters twice, self term twice
            else:
                z = self._relation_layer(adj, x, deprel, l, masks) / denom
            x = F.relu(z)
            if l < last:
                x = self._drop(x, self.opt['gcn_dropout'], 'gcn%d' % l, masks)
        return x, not_in_tree


def masked_pool(h, mask, kind):
    if kind == 'max':
        return h.masked_fill(mask, NEG_FILL).max(1)[0]
    h = h.masked_fill(mask, 0)
    if kind == 'avg':
        return h.sum(1) / (mask.size(1) - mask.float().sum(1))
    return h.sum(1)


class _DenseRelationModel(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.emb = nn.Embedding(opt['vocab_size'], opt['emb_dim'], padding_idx=0)
        self.pos_emb = nn.Embedding(N_POS, opt['pos_dim']) if opt['pos_dim'] > 0 else None
        self.ner_emb = nn.Embedding(N_NER, opt['ner_dim']) if opt['ner_dim'] > 0 else None
        adj_type = opt.get('adj_type', 'regular')    # gcn.py:48-57: width H (diagonal), D (full), dummy 1 (regular)
        side = {'regular': 1, 'diagonal_deprel': opt['hidden_dim']}.get(adj_type, opt.get('deprel_emb_dim', 1))
        self.deprel_emb = nn.Embedding(N_DEPREL, side, padding_idx=0)
        self.emb.weight.data[1:].uniform_(-1.0, 1.0)
        self.gcn = _DenseGCN(opt, (self.emb, self.pos_emb, self.ner_emb, self.deprel_emb))
        hidden = opt['hidden_dim']
        mlp = [nn.Linear(3 * hidden, hidden), nn.ReLU()]
        for _ in range(opt['mlp_layers'] - 1):
            mlp += [nn.Linear(hidden, hidden), nn.ReLU()]
        self.out_mlp = nn.Sequential(*mlp)

    def adjacency(self, inputs):
        """Host-side, per-sentence, dense -- the reference's inputs_to_tree_reps (gcn.py:96-110)."""
        tacred = self.opt['dataset'] == 'tacred'
        pad = inputs[1]
        deprel, head, subj_pos, obj_pos = inputs[4:8] if tacred else inputs[3:7]
        lens = (pad.numpy() == 0).astype(np.int64).sum(1)
        adj = tree_oracle.batch_adjacency(head.numpy(), subj_pos.numpy(), obj_pos.numpy(), deprel.numpy(),
                                          lens, self.opt['prune_k'], int(lens.max()))
        return torch.from_numpy(adj)

    def forward(self, inputs, masks=None, adj=None):
        if adj is None:
            adj = self.adjacency(inputs)
        h, not_in_tree = self.gcn(adj, inputs, masks)
        subj_pos, obj_pos = inputs[-2], inputs[-1]
        kind = self.opt['pooling']
        pooled = [masked_pool(h, m, kind) for m in
                  (not_in_tree, subj_pos.ne(0).unsqueeze(2), obj_pos.ne(0).unsqueeze(2))]
        return self.out_mlp(torch.cat(pooled, dim=1)), pooled[0]


class DenseClassifier(nn.Module):
    """Same state_dict keys as the reference's GCNClassifier (SURVEY.md §8b checkpoint row)."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.gcn_model = _DenseRelationModel(opt)
        self.classifier = nn.Linear(opt['hidden_dim'], opt['num_class'])

    def forward(self, inputs, masks=None, adj=None):
        rep, h_out = self.gcn_model(inputs, masks, adj)
        return self.classifier(rep), h_out

    def loss(self, batch, masks=None):
        """trainer.py:87-100 -- CE + conv_l2 * sum(W^2, b^2) + pooling_l2 * mean_b sum_h h_out^2."""
        inputs, labels = list(batch[:-2]), batch[-2]
        logits, h_out = self(inputs, masks)
        loss = F.cross_entropy(logits, labels)
        if self.opt.get('conv_l2', 0) > 0:
            loss = loss + self.gcn_model.gcn.conv_l2() * self.opt['conv_l2']
        if self.opt.get('pooling_l2', 0) > 0:
            loss = loss + self.opt['pooling_l2'] * (h_out ** 2).sum(1).mean()
        return loss, logits


def train_step(model, optimizer, batch, max_grad_norm=5.0):
    """One optimisation step as train.py:213-227 runs it (zero_grad, update, backward, clip, step)."""
    optimizer.zero_grad()
    loss, _ = model.loss(batch)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    return loss
