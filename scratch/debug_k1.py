import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import numpy as np, torch
import cases
from gcn_over_pruned_trees_b200 import ops, synth
from oracle import tree_oracle
g = np.load('tests/golden/adjacency.npz')
batch = cases.batch_from_npz(g, 'train')
for k in (-1, 0, 1):
    dep, head, sp, op = [t.cuda() for t in batch[4:8]]
    csr = ops.prune_csr(head, sp, op, dep, batch[1].cuda(), k)
    got = csr.to_dense().numpy()
    want = g['train/adj_k%d' % k].astype(np.float32)
    print('k', k, 'err', csr.err.tolist())
    bad = np.argwhere(got != want)
    print('n mismatches', len(bad))
    for (b, i, j) in bad[:12]:
        print('  b%d (%d,%d) got %g want %g' % (b, i, j, got[b, i, j], want[b, i, j]))
    if len(bad):
        b = bad[0][0]
        n = int(csr.lens[b])
        rp = csr.rowptr[b].cpu().numpy(); col = csr.col[b].cpu().numpy(); val = csr.val[b].cpu().numpy()
        print('  head', batch[5][b, :n].tolist())
        for i in range(n):
            print('   row', i, list(zip(col[rp[i]:rp[i+1]].tolist(), val[rp[i]:rp[i+1]].tolist())), ' want', [(int(j), int(want[b,i,j])) for j in np.nonzero(want[b,i])[0]])
