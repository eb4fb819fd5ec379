#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for ONE training step
(steps are delimited by the prune_csr kernel; without a step index the last of the shortest steps, i.e. a fused
one, is taken)."""
import collections
import csv
import sys


def main():
    src = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    i_name, i_val, i_id = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
    recs = [(r[i_name], float(r[i_val].replace(',', ''))) for r in rows[1:] if r[i_id].isdigit()]
    marks = [i for i, (n, v) in enumerate(recs) if 'prune_csr' in n]
    segs = [(marks[i], marks[i + 1]) for i in range(len(marks) - 1)]
    if len(sys.argv) > 2:
        s, e = segs[which]
    else:       # the fused step: the shortest segments (eager / autograd steps of the same run launch several times more)
        fewest = min(e - s for s, e in segs)
        s, e = [se for se in segs if se[1] - se[0] == fewest][-1]
    step = recs[s:e]
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n[:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    print('# kernels in step: %d, total %.1f us (cold-cache, serialised: compare shares)' % (len(step), tot / 1000))
    print('kernel,launches,total_us,share')
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('"%s",%d,%.1f,%.3f' % (n, c, v / 1000, v / tot))


if __name__ == '__main__':
    main()
