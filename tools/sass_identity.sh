#!/bin/bash
# Are the device instruction streams of two revisions of csrc/ identical?  Used after host-only edits (the
# GPT_HOST_EMULATION seams of tests/emu) to show that the library a GPU run verified is the library in the tree.
#   tools/sass_identity.sh <git-rev-that-was-verified-on-the-GPU> [<rev, default: working tree>]
set -e
repo=$(cd "$(dirname "$0")/.." && pwd)
a=${1:?usage: sass_identity.sh REV_A [REV_B]}; b=${2:-WORKTREE}
tmp=$(mktemp -d)
for side in a b; do
  rev=$([ $side = a ] && echo "$a" || echo "$b")
  mkdir -p $tmp/$side/csrc $tmp/$side/out
  if [ "$rev" = WORKTREE ]; then cp $repo/gcn_over_pruned_trees_b200/csrc/*.cu $repo/gcn_over_pruned_trees_b200/csrc/*.cuh $tmp/$side/csrc/
  else git -C $repo archive "$rev" gcn_over_pruned_trees_b200/csrc | tar -x -C $tmp/$side --strip-components=1; fi
  ( cd $tmp/$side/csrc
    for f in *.cu; do
      ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-O2 -c $f -o ../out/${f%.cu}.o 2>/dev/null &&
        cuobjdump -sass ../out/${f%.cu}.o | grep -o "^\s*/\*[0-9a-f]\{4\}\*/.*;" | sed 's#/\*[0-9a-f]*\*/##g' > ../out/${f%.cu}.ins ) &
    done; wait )
done
rc=0
for f in $tmp/b/out/*.ins; do
  n=$(basename $f .ins)
  if [ ! -f $tmp/a/out/$n.ins ]; then echo "$n: new file"; elif cmp -s $tmp/a/out/$n.ins $f; then echo "$n: identical ($(wc -l < $f) instructions)"; else echo "$n: DIFFERS"; rc=1; fi
done
rm -rf $tmp
exit $rc
