#!/usr/bin/env python
"""The roofline leg of bench.py alone (K2 forward / backward variants, K1 and K4 forward at the large512 shape):
prints the K1 / K4 sub-records and K2's fraction of the measured HBM peak.  ~6 s on a B200."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

r = bench.aggregation_roofline(types.SimpleNamespace(roofline_batch=4096), bench.load_peaks())
print(json.dumps({k: v for k, v in r['other'].items() if k.startswith('k')}), r['frac'])
