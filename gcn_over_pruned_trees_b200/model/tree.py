"""``head_to_tree`` / ``tree_to_adj`` with the reference's signatures (/root/reference/model/tree.py:58,167), served
by the batched CUDA kernel K1 (csrc/prune_csr.cu) instead of per-sentence Python.

The model itself never builds ``Tree`` objects (it consumes the CSR directly); these wrappers exist for callers
of the old per-sentence API such as /root/reference/data/tree_structures.py:7.  Inputs are numpy arrays as in the
reference; they are moved to the current CUDA device, one kernel runs, and the pruned tree is rebuilt from the CSR.
"""
import numpy as np
import torch

try:
    from .. import constant, ops
except ImportError:
    from gcn_over_pruned_trees_b200 import constant, ops


class Tree(object):
    """Pruned-tree node: ``idx, deprel, head, token, parent, children, num_children`` (reference tree.py:16-56)."""

    def __init__(self):
        self.parent = None
        self.num_children = 0
        self.children = []
        self.idx = self.deprel = self.head = self.token = None
        self.dist = None            # not carried by the CSR

    def add_child(self, child):
        child.parent = self
        self.num_children += 1
        self.children.append(child)

    def size(self):
        return 1 + sum(c.size() for c in self.children)

    def depth(self):
        return 1 + max(c.depth() for c in self.children) if self.children else 0

    def __iter__(self):
        yield self
        for c in self.children:
            for x in c:
                yield x


def head_to_tree(head, tokens, len_, prune, subj_pos, obj_pos, deprel):
    """One sentence -> root ``Tree`` of the (pruned) dependency tree; raises on trees the reference cannot build."""
    n = int(len_)
    dev = torch.device('cuda')

    def row(a, fill=0):
        out = np.full((1, n), fill, dtype=np.int64)
        out[0] = np.asarray(a)[:n]
        return torch.from_numpy(out).to(dev)

    masks = torch.zeros((1, n), dtype=torch.bool, device=dev)
    csr = ops.prune_csr(row(head), row(subj_pos), row(obj_pos), row(deprel), masks, int(prune))
    csr.check()
    rowptr, col, val = csr.rowptr[0].cpu().numpy(), csr.col[0].cpu().numpy(), csr.val[0].cpu().numpy()
    in_tree = (csr.flags[0].cpu().numpy() & 1) != 0
    nodes = {}
    for i in np.nonzero(in_tree)[0]:
        t = Tree()
        t.idx, t.deprel, t.head, t.token = int(i), int(deprel[i]), int(head[i]), tokens[i]
        nodes[int(i)] = t
    root = None
    for i, t in nodes.items():
        is_child = False
        for e in range(rowptr[i], rowptr[i + 1]):
            v = int(val[e])
            if 0 < v < constant.DEPREL_FORWARD_BOUND:                      # forward entry: i -> child col[e]
                t.add_child(nodes[int(col[e])])
            elif constant.DEPREL_FORWARD_BOUND <= v < constant.SELF_LOOP_INDEX:   # reverse entry: i hangs under col[e]
                is_child = True
        if not is_child:
            root = t
    if root is None:                 # single-node tree: no adjacency entries at all
        root = Tree()
        kept = np.nonzero(in_tree)[0]
        assert len(kept) == 0
    root._csr = (n, rowptr, col, val)
    return root


def tree_to_adj(sent_len, tree, directed=True, self_loop=False):
    """Dense float32 [sent_len, sent_len] adjacency with relation-id values (reference tree.py:167-204)."""
    ret = np.zeros((sent_len, sent_len), dtype=np.float32)
    n, rowptr, col, val = tree._csr
    for i in range(n):
        for e in range(rowptr[i], rowptr[i + 1]):
            v = int(val[e])
            if v == constant.SELF_LOOP_INDEX:
                if self_loop:
                    ret[i, col[e]] = v
            elif v >= constant.DEPREL_FORWARD_BOUND:
                if not directed:
                    ret[i, col[e]] = v
            else:
                ret[i, col[e]] = v
    return ret
