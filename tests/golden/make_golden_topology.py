#!/usr/bin/env python
"""Golden tables for SURVEY.md row f3 (host topology pipeline), produced by the LIVE reference.

Build container only:   python tests/golden/make_golden_topology.py
* writes two small topology files of our own in the reference's two input formats
  (tests/golden/topo_small.txt, tests/golden/topo_small.xml -- synthetic, not reference data),
* runs the reference's examples/create_topology.py:get_topology on them and on its shipped
  germany50.xml, and stores the flattened result (TopologyTables.from_graph) as
  topo_small_txt_tables.npz / topo_small_xml_tables.npz / topo_germany50_tables.npz.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "optical-rl-gym_b200"))

import ref_harness as rh  # noqa: E402

from optical_rl_gym_b200.topology import TopologyTables  # noqa: E402

SMALL_TXT = """# synthetic 9-node mesh in the reference's .txt format: nodes, links, then "src dst km" per link
9
14
1 2 300
2 3 450
3 4 300
4 5 700
5 6 300
6 1 520
1 7 410
7 4 380
2 8 300
8 5 640
3 9 450
9 6 450
7 8 150
8 9 150
"""

_NODES = [("Aveiro", -8.65, 40.64), ("Braga", -8.43, 41.55), ("Coimbra", -8.41, 40.21), ("Evora", -7.91, 38.57),
          ("Faro", -7.93, 37.02), ("Guarda", -7.27, 40.54), ("Lisboa", -9.14, 38.72), ("Porto", -8.61, 41.15),
          ("Viseu", -7.91, 40.66), ("Leiria", -8.81, 39.74)]
_LINKS = [("Porto", "Braga"), ("Porto", "Aveiro"), ("Aveiro", "Coimbra"), ("Coimbra", "Leiria"), ("Leiria", "Lisboa"),
          ("Lisboa", "Evora"), ("Evora", "Faro"), ("Lisboa", "Faro"), ("Coimbra", "Viseu"), ("Viseu", "Guarda"),
          ("Guarda", "Evora"), ("Braga", "Viseu"), ("Aveiro", "Viseu"), ("Guarda", "Coimbra"), ("Porto", "Viseu")]


def small_xml():
    out = ['<?xml version="1.0" encoding="ISO-8859-1"?>', '<network xmlns="http://sndlib.zib.de/network" version="1.0">',
           ' <networkStructure>', '  <nodes coordinatesType="geographical">']
    for name, x, y in _NODES:
        out += ['   <node id="%s">' % name, '    <coordinates>', '     <x>%s</x>' % x, '     <y>%s</y>' % y,
                '    </coordinates>', '   </node>']
    out += ['  </nodes>', '  <links>']
    for i, (a, b) in enumerate(_LINKS):
        out += ['   <link id="L%d">' % i, '    <source>%s</source>' % a, '    <target>%s</target>' % b, '   </link>']
    out += ['  </links>', ' </networkStructure>', '</network>', '']
    return "\n".join(out)


def main():
    rh.setup()
    sys.path.insert(0, os.path.join(rh.REFERENCE, "examples"))
    import create_topology as ct          # the reference's own pipeline (module-level code only defines the modulations)

    mods = getattr(ct, "modulations", None)
    if mods is None:                       # the tuple is a local of the __main__ block in some revisions: rebuild it
        from optical_rl_gym.utils import Modulation
        mods = (Modulation("BPSK", 100_000, 1, 12.6, -14), Modulation("QPSK", 2_000, 2, 12.6, -17),
                Modulation("8QAM", 1_000, 3, 18.6, -20), Modulation("16QAM", 500, 4, 22.4, -23),
                Modulation("32QAM", 250, 5, 26.4, -26), Modulation("64QAM", 125, 6, 30.4, -29))
    with open(os.path.join(HERE, "topo_small.txt"), "w") as f:
        f.write(SMALL_TXT)
    with open(os.path.join(HERE, "topo_small.xml"), "w") as f:
        f.write(small_xml())
    jobs = [(os.path.join(HERE, "topo_small.txt"), "topo_small_txt", 4),
            (os.path.join(HERE, "topo_small.xml"), "topo_small_xml", 3),
            (os.path.join(rh.REFERENCE, "examples", "topologies", "germany50.xml"), "topo_germany50", 5)]
    for path, name, k in jobs:
        with contextlib.redirect_stdout(io.StringIO()):     # get_topology prints every path
            g = ct.get_topology(path, name, mods, k)
        t = TopologyTables.from_graph(g)
        out = os.path.join(HERE, name + "_tables.npz")
        t.save(out)
        print("%-20s nodes %3d links %3d paths %5d  %.1f KB" % (name, t.num_nodes, t.num_links, t.num_paths,
                                                               os.path.getsize(out) / 1024.0))


if __name__ == "__main__":
    main()
