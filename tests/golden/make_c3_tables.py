#!/usr/bin/env python
"""BASELINE.json configs[3] topology: seeded synthetic 100-node ring + 200 chords (300 links), k = 10 shortest
paths by NetworkX (utils.get_k_shortest_paths equivalent, ~1 minute).  The resulting flat tables are committed
(topo_c3_ring100_chords200_k10.npz) so that the GPU tests do not repeat the host pre-processing:
    python tests/golden/make_c3_tables.py
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "optical-rl-gym_b200"))
from optical_rl_gym_b200.topology import synthetic_ring_chords  # noqa: E402

t0 = time.time()
t = synthetic_ring_chords(num_nodes=100, num_chords=200, k_paths=10, seed=1)
path = os.path.join(HERE, "topo_c3_ring100_chords200_k10.npz")
t.save(path)
print("%s: %d nodes, %d links, %d paths, max hops %d, %.1f KB, %.0f s" % (
    t.name, t.num_nodes, t.num_links, t.num_paths, int(t.path_hops.max()), os.path.getsize(path) / 1024, time.time() - t0))
