"""GCN relation classifier on the B200 kernels -- same classes, constructor arguments, forward signatures and
``state_dict`` keys as /root/reference/model/gcn.py (GCNClassifier :15-36, GCNRelationModel :38-126, GCN :128-470,
pool :473-483, rnn_zero_state :485-492): ``adj_type='regular'`` and the fork's relation-aware ``'full_deprel'`` /
``'diagonal_deprel'`` layers (gcn.py:272-386, 400-470).

What changed underneath:
  * the per-sentence numpy loop + dense [B,T,T] adjacency (gcn.py:96-110) is one kernel launch that leaves a CSR
    on the device (ops.prune_csr -> csrc/prune_csr.cu); nothing is copied to the host inside forward()
  * each layer is one projection GEMM + one fused gather/normalise/ReLU/dropout kernel (ops.gcn_layer) instead
    of bmm + 2 Linear + div + relu + dropout (gcn.py:269-271, 390-393)
  * the three pool() passes + cat (gcn.py:116-121) are one kernel (ops.pool3)
  * relation-aware layers: one projection GEMM shared by the forward / reverse / self-loop traversals + a relation
    mix kernel + a direction-aware CSR gather (ops.relation_layer_full / relation_layer_diag -> csrc/deprel.cu)
    instead of three [B,T,D,K] outer products, three [D,K,H] contractions and two dense bmm per layer
There is no CPU implementation: inputs must live on a CUDA device.
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

try:                                    # imported as gcn_over_pruned_trees_b200.model.gcn
    from .. import constant, ops, torch_utils
except ImportError:                     # imported as top-level `model.gcn` (package dir on PYTHONPATH, like the reference)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from gcn_over_pruned_trees_b200 import constant, ops, torch_utils


class GCNClassifier(nn.Module):
    """GCNRelationModel + linear classifier (reference gcn.py:15-36)."""

    def __init__(self, opt, emb_matrix=None):
        super().__init__()
        self.gcn_model = GCNRelationModel(opt, emb_matrix=emb_matrix)
        self.classifier = nn.Linear(opt['hidden_dim'], opt['num_class'])
        self.opt = opt

    def conv_l2(self):
        return self.gcn_model.gcn.conv_l2()

    def forward(self, inputs):
        outputs, pooling_output = self.gcn_model(inputs)
        return self.classifier(outputs), pooling_output

    def loss_fused(self, inputs, labels, pooling_l2):
        """Training-time ``CrossEntropy(classifier(out_mlp(pooled))) + pooling_l2 * mean_b |h_out_b|^2`` (gcn.py:122 +
        trainer.py:94-100) with the whole head -- forward, loss and every gradient of it -- in K6's two launches instead
        of ~45 ATen kernels, still under autograd (ops.head_loss).  Returns (loss, logits); None when K6 does not take
        the configuration."""
        mlp = [m for m in self.gcn_model.out_mlp if isinstance(m, nn.Linear)]
        H = self.opt['hidden_dim']
        if H % 4 != 0 or len(mlp) > 4 or not labels.is_cuda:
            return None
        pooled = self.gcn_model.pooled(inputs)
        wb = [t for m in mlp for t in (m.weight, m.bias)]
        return ops.head_loss(pooled, labels, float(pooling_l2), self.classifier.weight, self.classifier.bias, *wb)

    def get_deprel_emb(self):
        return self.gcn_model.get_deprel_embedding()

    def get_gcn_parameters(self):
        return self.gcn_model.get_gcn_parameters()


class GCNRelationModel(nn.Module):
    def __init__(self, opt, emb_matrix=None):
        super().__init__()
        adj_type = opt.get('adj_type', 'regular')
        if adj_type not in ('regular', 'full_deprel', 'diagonal_deprel'):
            # 'concat_deprel' builds weights in the reference but its forward raises ValueError (gcn.py:388)
            raise NotImplementedError("adj_type=%r is not a working mode of the reference (SURVEY.md 10-3)" % adj_type)
        self.opt = opt
        self.emb_matrix = emb_matrix
        # tables are registered here AND on self.gcn, as in the reference (gcn.py:45-57,138): the checkpoint
        # carries both key sets, pointing at shared storage
        self.emb = nn.Embedding(opt['vocab_size'], opt['emb_dim'], padding_idx=constant.PAD_ID)
        self.pos_emb = nn.Embedding(constant.NUM_POS, opt['pos_dim']) if opt['pos_dim'] > 0 else None
        self.ner_emb = nn.Embedding(constant.NUM_NER, opt['ner_dim']) if opt['ner_dim'] > 0 else None
        # relation vectors (gcn.py:48-57): width hidden_dim ('diagonal_deprel'), deprel_emb_dim ('full_deprel'), or a
        # dummy column that 'regular' never reads
        side = {'regular': 1, 'diagonal_deprel': opt['hidden_dim']}.get(adj_type, opt.get('deprel_emb_dim', 1))
        self.deprel_emb = nn.Embedding(constant.NUM_DEPREL, side, padding_idx=0)
        self.init_embeddings()
        self.gcn = GCN(opt, (self.emb, self.pos_emb, self.ner_emb, self.deprel_emb), opt['hidden_dim'],
                       opt['num_layers'])
        hidden = opt['hidden_dim']
        layers = [nn.Linear(3 * hidden, hidden), nn.ReLU()]
        for _ in range(opt['mlp_layers'] - 1):
            layers += [nn.Linear(hidden, hidden), nn.ReLU()]
        self.out_mlp = nn.Sequential(*layers)
        self.last_csr = None

    def get_deprel_embedding(self):
        return self.deprel_emb.weight

    def init_embeddings(self):
        if self.emb_matrix is None:
            self.emb.weight.data[1:, :].uniform_(-1.0, 1.0)
        else:
            self.emb_matrix = torch.from_numpy(self.emb_matrix)
            self.emb.weight.data.copy_(self.emb_matrix)
        topn = self.opt['topn']
        if topn <= 0:
            print("Do not finetune word embedding layer.")
            self.emb.weight.requires_grad = False
        elif topn < self.opt['vocab_size']:
            print("Finetune top {} word embeddings.".format(topn))
            self.emb.weight.register_hook(lambda g: torch_utils.keep_partial_grad(g, int(topn)))
        else:
            print("Finetune all embeddings.")

    def pooled(self, inputs):
        """[B, 3H]: the sentence / subject / object pools of the last GCN layer (gcn.py:96-121 up to the cat)."""
        if self.opt['dataset'] == 'tacred':
            words, masks, pos, ner, deprel, head, subj_pos, obj_pos = inputs
        else:
            words, masks, pos, deprel, head, subj_pos, obj_pos = inputs
        # K1: pruned trees -> CSR, on the device, no host round trip
        csr = ops.prune_csr(head, subj_pos, obj_pos, deprel, masks, self.opt['prune_k'])
        self.last_csr = csr
        h, _ = self.gcn(csr, inputs)
        # K4: sentence / subject / object pools in one pass
        return ops.pool3(h, csr, self.opt['pooling'])

    def forward(self, inputs):
        pooled = self.pooled(inputs)
        h_out = pooled[:, :self.opt['hidden_dim']]
        return self.out_mlp(pooled), h_out

    def get_gcn_parameters(self):
        return self.gcn.get_gcn_parameters()


class GCN(nn.Module):
    """GCN / C-GCN over the pruned-tree CSR (reference gcn.py:128-470)."""

    def __init__(self, opt, embeddings, mem_dim, num_layers):
        super().__init__()
        self.opt = opt
        self.layers = num_layers
        self.use_cuda = opt['cuda']
        self.mem_dim = mem_dim
        self.in_dim = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if opt['dataset'] == 'tacred' else 0)
        self.emb, self.pos_emb, self.ner_emb, self.deprel_emb = embeddings
        if opt.get('rnn', False):      # C-GCN encoder stays on cuDNN (north_star)
            self.rnn = nn.LSTM(self.in_dim, opt['rnn_hidden'], opt['rnn_layers'], batch_first=True,
                               dropout=opt['rnn_dropout'], bidirectional=True)
            self.in_dim = opt['rnn_hidden'] * 2
            self.rnn_drop = nn.Dropout(opt['rnn_dropout'])
        self.in_drop = nn.Dropout(opt['input_dropout'])
        self.gcn_drop = nn.Dropout(opt['gcn_dropout'])   # kept for API parity; the fused kernel draws its own mask
        if opt.get('emb_dropout', 0.0) > 0:
            raise NotImplementedError('emb_dropout > 0 (EmbeddingDropout) is outside the built path (SURVEY.md 2-#8)')
        self.adj_type = opt.get('adj_type', 'regular')
        if self.adj_type == 'diagonal_deprel':       # gcn.py:153-155: a preprocessor, no per-layer weights at all
            self.preprocessor = nn.Linear(self.in_dim, mem_dim)
            self.in_dim = mem_dim
        elif self.adj_type == 'full_deprel':         # gcn.py:164-167: ONE Linear(in, D*H) shared by every layer
            if num_layers > 1 and self.in_dim != mem_dim:
                raise ValueError('full_deprel shares one Linear(in_dim, D*H) between the layers (gcn.py:164-167,301): the '
                                 'GCN input width %d must equal hidden_dim %d (the reference fails inside einsum)'
                                 % (self.in_dim, mem_dim))
            self.W = nn.Linear(self.in_dim, opt['deprel_emb_dim'] * mem_dim, bias=True)
        else:
            self.W = nn.ModuleList(nn.Linear(self.in_dim if l == 0 else mem_dim, mem_dim) for l in range(num_layers))
        if self.adj_type != 'regular' and (opt.get('edge_keep_prob', 1.0) <= 0 or opt.get('deprel_keep_prop', 1.0) < 0):
            raise ValueError('edge_keep_prob must be in (0, 1], deprel_keep_prop in [0, 1]')
        # projection arithmetic: 'tf32x3' = tcgen05 3xTF32 (fp32-grade, passes the 1e-5 parity tests; FFMA fallback for
        # shapes TMA cannot describe), 'fp32' = FFMA everywhere, 'tf32' = one TF32 pass (~1e-3)
        self.gemm_mode = opt.get('gemm_mode', 'tf32x3')
        # {seed, step} consumed by the in-kernel Philox dropout; int64 storage, read as uint64 by the kernel
        self.register_buffer('rng_state', torch.tensor([torch.initial_seed() & 0x7fffffffffffffff, 0],
                                                       dtype=torch.int64), persistent=False)
        self.injected_masks = None      # tests only: {'in': m, 'rnn': m, 'gcn0': m, ...}, pre-scaled by 1/(1-p)
        self.sparse_embedding = None    # ops.SparseEmbeddingState while engine.GraphedTrainStep drives the step

    def conv_l2(self):
        if self.adj_type == 'full_deprel':
            # the reference iterates `for w in self.W` over a single nn.Linear and raises TypeError (gcn.py:182,
            # SURVEY.md 10-3); the penalty it evidently meant is the one over that shared layer
            return self.W.weight.pow(2).sum() + self.W.bias.pow(2).sum()
        return sum(p.pow(2).sum() for lin in self.W for p in (lin.weight, lin.bias))   # 'diagonal_deprel' has no W

    def get_gcn_parameters(self):
        return self.W

    def _host_dropout(self, x, module, name):
        if self.injected_masks is not None and name in self.injected_masks:
            return x * self.injected_masks[name]
        return module(x)

    def encode_with_rnn(self, rnn_inputs, masks, batch_size):
        """BiLSTM over the padded batch (reference gcn.py:186-197: pack_padded_sequence -> nn.LSTM -> pad_packed_sequence),
        same arithmetic WITHOUT leaving the device: the reference needs the lengths on the host to pack, which costs a
        synchronisation and makes cuDNN's descriptors data dependent (no CUDA-graph capture).  An LSTM is causal, so a
        forward pass over the padded rows is exact on the real tokens; the backward direction runs as a forward LSTM
        (with the ``_reverse`` weights) over each sentence reversed inside its own length -- a gather with an index
        built on the device -- and is un-reversed by the same gather.  Padded positions are zeroed, as
        pad_packed_sequence does.  Still cuDNN (torch._VF.lstm), two unidirectional calls per layer."""
        B, T, _ = rnn_inputs.shape
        H, L = self.opt['rnn_hidden'], self.opt['rnn_layers']
        lens = masks.eq(constant.PAD_ID).sum(1, keepdim=True)                        # [B,1], stays on the device
        t = torch.arange(T, device=rnn_inputs.device).unsqueeze(0)                   # [1,T]
        valid = t < lens                                                             # [B,T]
        rev = torch.where(valid, lens - 1 - t, t)                                    # an involution on every row
        h0 = rnn_inputs.new_zeros((1, B, H))                                         # rnn_zero_state, gcn.py:485-492
        x = rnn_inputs
        for l in range(L):
            names = ['weight_ih_l%d', 'weight_hh_l%d', 'bias_ih_l%d', 'bias_hh_l%d']
            w_f = [getattr(self.rnn, n % l) for n in names]
            w_b = [getattr(self.rnn, (n % l) + '_reverse') for n in names]
            idx = rev.unsqueeze(2).expand(-1, -1, x.size(2))
            out_f = torch._VF.lstm(x, (h0, h0), w_f, True, 1, 0.0, self.training, False, True)[0]
            out_b = torch._VF.lstm(x.gather(1, idx), (h0, h0), w_b, True, 1, 0.0, self.training, False, True)[0]
            out_b = out_b.gather(1, rev.unsqueeze(2).expand(-1, -1, H))
            x = torch.cat([out_f, out_b], dim=2) * valid.unsqueeze(2).to(out_f.dtype)
            if l < L - 1 and self.training and self.opt['rnn_dropout'] > 0:          # nn.LSTM's inter-layer dropout
                x = F.dropout(x, self.opt['rnn_dropout'], True)
        return x

    def encode_with_rnn_packed(self, rnn_inputs, masks, batch_size):
        """The reference's own sequence of calls (gcn.py:186-197); kept for the parity test of encode_with_rnn."""
        seq_lens = masks.eq(constant.PAD_ID).long().sum(1).cpu()
        h0, c0 = rnn_zero_state(batch_size, self.opt['rnn_hidden'], self.opt['rnn_layers'],
                                use_cuda=rnn_inputs.is_cuda)
        packed = nn.utils.rnn.pack_padded_sequence(rnn_inputs, seq_lens, batch_first=True, enforce_sorted=False)
        out, _ = self.rnn(packed, (h0, c0))
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=rnn_inputs.size(1))
        return out

    def _relation_layers(self, x, csr, deprel, drop_p, rng=None):
        """The layer loop of the relation-aware modes (gcn.py:272-386 + 390-393).  `no_adj` has no effect here, as in the
        reference (these branches re-derive their masks from `adj`, gcn.py:276,308)."""
        opt = self.opt
        emb = self.deprel_emb.weight
        injected = self.injected_masks
        rng = self.rng_state if rng is None else rng      # the forward's frozen {seed, step} (see forward())
        full = self.adj_type == 'full_deprel'
        if full:
            D, H = opt['deprel_emb_dim'], self.mem_dim
            # weight_l = W.weight.reshape(D, in, H) is a reshape of the [D*H, in] matrix, not a permute (gcn.py:301);
            # the projection GEMM wants the [D*H, in] matrix whose row d*H+h is weight_l[d, :, h]
            wmat = self.W.weight.reshape(D, -1, H).permute(0, 2, 1).reshape(D * H, -1).contiguous()
            ws = ops.weight_prep(wmat.detach(), self.gemm_mode)
            # only rows some pool can see are projected and mixed (a quarter of a TACRED-shaped batch at prune_k = 1): the
            # reference computes all of them and masks the rest away in pool() (gcn.py:116-120, 262)
            live = ops.LiveRows(csr.flags) if os.environ.get('GPT_K10_COMPACT', '1') != '0' else None
        else:
            x = ops.linear(x, self.preprocessor.weight, self.preprocessor.bias, self.gemm_mode)    # gcn.py:255-257
        B, T = csr.B, csr.T
        for l in range(self.layers):
            last = l == self.layers - 1
            mask = None if injected is None else injected.get('gcn%d' % l)
            cfg = ops.RelationLayerConfig(l, drop_p=0.0 if (last or mask is not None) else drop_p,
                                          drop_mask=None if last else mask, rng_state=rng,
                                          gemm_mode=self.gemm_mode, live=live if full else None)
            if not full:
                x = ops.relation_layer_diag(x, emb, csr, deprel, cfg)
                continue
            cfg.deep = l >= opt['deprel_max_depth']
            cfg.directed, cfg.self_loop = bool(opt['deprel_directed']), bool(opt['deprel_self_loop'])
            if injected is not None and ('edge_f%d' % l) in injected:
                # tests: dense 0/1 [B,T,T] masks per direction, as maybe_drop_edges draws them (gcn.py:436-449)
                cfg.keep_edges = tuple(injected['edge_%s%d' % (d, l)].to(torch.uint8).contiguous() for d in 'fr')
            elif self.training and opt.get('edge_keep_prob', 1.0) < 1.0:
                cfg.edge_keep = opt['edge_keep_prob']
            if injected is not None and ('forget_f%d' % l) in injected:
                cfg.keep_tokens = tuple(injected['forget_%s%d' % (d, l)].reshape(-1).to(torch.uint8).contiguous()
                                        for d in 'fr')
            elif self.training and opt.get('deprel_keep_prop', 1.0) < 1.0:
                cfg.keep_tokens = ops.relation_keep_tokens(rng, B * T, l, opt['deprel_keep_prop'])
            x = ops.relation_layer_full(x, wmat, self.W.bias, emb, csr, deprel, cfg, ws)
        return x

    def forward(self, adj, inputs):
        if not isinstance(adj, ops.TreeCSR):
            raise TypeError('GCN.forward takes the TreeCSR produced by ops.prune_csr, not a dense adjacency')
        if self.opt['dataset'] == 'tacred':
            words, masks, pos, ner, deprel, head, subj_pos, obj_pos = inputs
        else:
            words, masks, pos, deprel, head, subj_pos, obj_pos = inputs
            ner = None
        use_ner = self.opt['ner_dim'] > 0 and self.opt['dataset'] == 'tacred'
        rng = self.rng_state
        if self.training and self.injected_masks is None:
            self.rng_state[1] += 1         # new dropout streams every training forward (graph-capture safe)
            # this forward's {seed, step}, frozen: the backward kernels re-derive the Philox masks from it, and another
            # training forward may advance the live buffer before this one's backward runs (two losses summed before
            # one backward, activation checkpointing, ...).  A device-side copy: capture-safe.
            rng = self.rng_state.clone()
        if words.dim() > 2 or self.injected_masks is not None:
            # pre-computed token vectors (BERT path of the loader) or injected test masks: plain lookups
            embs = [words if words.dim() > 2 else self.emb(words)]
            if self.opt['pos_dim'] > 0:
                embs.append(self.pos_emb(pos))
            if use_ner:
                embs.append(self.ner_emb(ner))
            x = self._host_dropout(torch.cat(embs, dim=2), self.in_drop, 'in')
        else:
            # K5: gather + concat + input dropout in one kernel; its backward scatters straight into the tables
            x = ops.embed_concat(words, pos, ner if use_ner else None, self.emb.weight,
                                 self.pos_emb.weight if self.opt['pos_dim'] > 0 else None,
                                 self.ner_emb.weight if use_ner else None,
                                 drop_p=self.opt['input_dropout'] if self.training else 0.0, rng_state=rng,
                                 subseq=0xE0, flags=None if self.opt.get('rnn', False) else adj.flags,
                                 topn=self.opt['topn'], sparse=self.sparse_embedding)
        if self.opt.get('rnn', False):
            x = self._host_dropout(self.encode_with_rnn(x, masks, words.size(0)), self.rnn_drop, 'rnn')
        drop_p = self.opt['gcn_dropout'] if self.training else 0.0
        if self.adj_type != 'regular':
            return self._relation_layers(x, adj, deprel, drop_p, rng), adj.pool_mask()
        use_adj = not self.opt.get('no_adj', False)
        for l, lin in enumerate(self.W):
            last = l == self.layers - 1
            mask = None if self.injected_masks is None else self.injected_masks.get('gcn%d' % l)
            p = 0.0 if (last or mask is not None) else drop_p
            x = ops.gcn_layer(x, lin.weight, lin.bias, adj, use_adj=use_adj, drop_p=p, rng_state=rng,
                              subseq=l, drop_mask=None if last else mask, gemm_mode=self.gemm_mode)
        return x, adj.pool_mask()


def pool(h, mask, type='max'):
    """Single masked pool with the reference's semantics (gcn.py:473-483); the model itself uses ops.pool3."""
    if type == 'max':
        return h.masked_fill(mask, -constant.INFINITY_NUMBER).max(1)[0]
    h = h.masked_fill(mask, 0)
    if type == 'avg':
        return h.sum(1) / (mask.size(1) - mask.float().sum(1))
    return h.sum(1)


def rnn_zero_state(batch_size, hidden_dim, num_layers, bidirectional=True, use_cuda=True):
    shape = (num_layers * (2 if bidirectional else 1), batch_size, hidden_dim)
    h0 = c0 = torch.zeros(*shape, device='cuda' if use_cuda else 'cpu')
    return h0, c0
