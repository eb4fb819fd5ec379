"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference's pruned-tree adjacency.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product path (``gcn_over_pruned_trees_b200``) never does.

What it restates
----------------
``head_to_tree`` (/root/reference/model/tree.py:58-165) followed by
``tree_to_adj(directed=False, self_loop=True)`` (/root/reference/model/tree.py:167-204) exactly as
``inputs_to_tree_reps`` calls them (/root/reference/model/gcn.py:102-110): one sentence in, one dense
``float32 [maxlen, maxlen]`` adjacency out whose entries are dependency-relation ids

    A[parent, child] = deprel[child]            tree.py:184
    A[child, parent] = deprel[child] + 42       tree.py:186-188
    A[i, i]          = 84 for every endpoint    tree.py:190-192

The restatement is set-wise instead of building ``Tree`` objects, but it keeps the same
per-sentence Python-loop cost model as the reference so that it is a fair CPU baseline.

Pinning
-------
Parity is pinned by ``tests/golden/make_golden.py``, which imports the real reference from
``/root/reference`` (with the ``Tree.head = None`` shim that ``prune_k < 0`` needs) and stores
its outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them and against
the SURVEY.md §8c hashes of the bundled ``dataset/tacred/*.json`` sample.
"""
import numpy as np

FORWARD_BOUND = 42   # /root/reference/utils/constant.py:13
SELF_LOOP_ID = 84    # /root/reference/utils/constant.py:16
DIST_INF = 10000     # /root/reference/model/tree.py:144


class MalformedTree(Exception):
    """Inputs on which the reference raises or never returns."""


def _chain(i, parent, limit):
    """[i, parent(i), ..., root]; the reference loops forever on a cycle (tree.py:91-94), we raise."""
    out = [i]
    while parent[out[-1]] >= 0:
        out.append(parent[out[-1]])
        if len(out) > limit:
            raise MalformedTree('cycle in head[]')
    return out


def kept_nodes_and_root(head, subj_pos, obj_pos, length, prune_k):
    """Return (kept: bool[length], root: int) of the pruned tree.

    prune_k < 0  -> tree.py:67-79: every token hangs under its head; the *last* token with head 0 wins as
                    root; only that root's component is ever visited by tree_to_adj's BFS.
    prune_k >= 0 -> tree.py:80-162: common ancestors of all entity tokens, lowest one = LCA, path nodes =
                    (all entity ancestor chains) minus the common ancestors plus the LCA, distance of every
                    other token = number of steps up to its first path-node ancestor (10000 if it walks
                    past the root), keep distance <= prune_k.
    """
    n = int(length)
    parent = [int(h) - 1 for h in head[:n]]
    for i, p in enumerate(parent):
        if p >= n or p < -1 or p == i:
            raise MalformedTree('head out of range')

    if prune_k < 0:
        roots = [i for i in range(n) if parent[i] < 0]
        if not roots:
            raise MalformedTree('no root')          # tree.py:164 assert
        root = roots[-1]                             # tree.py:76-77, later roots overwrite
        kept = np.zeros(n, dtype=bool)
        for i in range(n):
            kept[i] = _chain(i, parent, n)[-1] == root
        return kept, root

    subj = [i for i in range(n) if subj_pos[i] == 0]  # tree.py:82
    obj = [i for i in range(n) if obj_pos[i] == 0]    # tree.py:83
    if not subj:
        raise MalformedTree('empty subject span')    # tree.py:109 / :113 fail on cas=None
    on_some_chain = set()
    common = None
    for e in subj + obj:                              # tree.py:87-109
        chain = _chain(e, parent, n)
        on_some_chain.update(chain)
        common = set(chain) if common is None else common & set(chain)
    if not common:
        raise MalformedTree('entities in different components')   # tree.py:121-127 UnboundLocalError
    # the lowest common ancestor is the common ancestor with no common-ancestor child (tree.py:112-124)
    has_common_child = {parent[c] for c in common if parent[c] in common}
    lca = next(c for c in common if c not in has_common_child)
    path = (on_some_chain - common) | {lca}           # tree.py:126-127

    dist = np.full(n, DIST_INF, dtype=np.int64)       # tree.py:130-144
    for i in range(n):
        j, steps = i, 0
        while j >= 0 and j not in path:
            j = parent[j]
            steps += 1
            if steps > n:
                raise MalformedTree('cycle in head[]')
        if j >= 0:
            dist[i] = steps
    return dist <= prune_k, lca                        # tree.py:147,162


def pruned_adjacency(head, subj_pos, obj_pos, deprel, length, prune_k, maxlen):
    """Dense float32 [maxlen, maxlen] adjacency of one sentence, bit-for-bit what the reference builds."""
    kept, root = kept_nodes_and_root(head, subj_pos, obj_pos, length, prune_k)
    adj = np.zeros((maxlen, maxlen), dtype=np.float32)
    for c in range(int(length)):
        p = int(head[c]) - 1
        # tree.py:158-160: a kept token hangs under its head unless it is the pruned tree's root
        if not kept[c] or c == root or p < 0:
            continue
        if not kept[p]:
            raise MalformedTree('kept token under a pruned head')   # tree.py:159 assert
        adj[p, c] = deprel[c]
        adj[c, p] = deprel[c] + FORWARD_BOUND
        adj[p, p] = SELF_LOOP_ID
        adj[c, c] = SELF_LOOP_ID
    return adj


def batch_adjacency(head, subj_pos, obj_pos, deprel, lengths, prune_k, maxlen=None):
    """[B, maxlen, maxlen] float32, the tensor inputs_to_tree_reps returns (gcn.py:105-108)."""
    head, subj_pos, obj_pos, deprel = (np.asarray(a) for a in (head, subj_pos, obj_pos, deprel))
    lengths = np.asarray(lengths)
    if maxlen is None:
        maxlen = int(lengths.max())               # gcn.py:97
    out = [pruned_adjacency(head[b], subj_pos[b], obj_pos[b], deprel[b], lengths[b], prune_k, maxlen)[None]
           for b in range(len(lengths))]
    return np.concatenate(out, axis=0)
