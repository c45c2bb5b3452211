#!/usr/bin/env python
"""Compare two accumulation buffers saved by tools/time_path.py --save: pixels that differ, and by how much."""
import sys
import numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
d = np.abs(a[..., :3].astype(np.float64) - b[..., :3]).max(axis=2)
print(f"{sys.argv[1]} vs {sys.argv[2]}: {int((d > 0).sum())} of {d.size} pixels differ; max |sum difference| {d.max():.0f} "
      f"(samples per pixel {a[..., 3].max():.0f}); counts equal: {bool(np.array_equal(a[..., 3], b[..., 3]))}")
