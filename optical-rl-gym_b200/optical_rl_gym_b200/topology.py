"""Host-side topology pre-processing (done once, uploaded to HBM by the native library).

Mirrors what the reference keeps in ``topology.graph`` after
``examples/create_topology.py:96-147`` (``get_topology``): for every unordered node
pair the k shortest simple paths by length, each with hops, length and the most
spectrally-efficient modulation whose reach covers it (``utils.py:84-96``).  Here
the result is a set of flat numpy tables instead of a dict of dataclasses:

* ``pair_first[src*N+dst]`` / ``pair_count[...]``  -> rows of the path tables
  (both orders of a pair map to the same rows, as ``create_topology.py:136-137``)
* ``path_hops, path_length, path_se, path_mod`` and the CSR
  ``path_link_ptr/path_links`` (link *index* of every hop, ``graph_utils.py:106-113``)
* modulation table ``mod_se, mod_reach, mod_osnr, mod_xt``.
"""
from __future__ import annotations

import dataclasses
from itertools import islice
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

# Modulation set of the reference's shipped pickles (examples/create_topology.py:47-92):
# (name, reach km, spectral efficiency, min OSNR dB, in-band XT dB)
DEFAULT_MODULATIONS = (
    ("BPSK", 100_000, 1, 12.6, -14.0),
    ("QPSK", 2_000, 2, 12.6, -17.0),
    ("8QAM", 1_000, 3, 18.6, -20.0),
    ("16QAM", 500, 4, 22.4, -23.0),
    ("32QAM", 250, 5, 26.4, -26.0),
    ("64QAM", 125, 6, 30.4, -29.0),
)

# NSFNET as used by DeepRMSA (14 nodes, 22 links; node ids 1-based, length in km).
NSFNET_LINKS = (
    (1, 2, 1050), (1, 3, 1500), (1, 8, 2400), (2, 3, 600), (2, 4, 750), (3, 6, 1800),
    (4, 5, 600), (4, 11, 1950), (5, 6, 1200), (5, 7, 600), (6, 10, 1050), (6, 14, 1800),
    (7, 8, 750), (7, 10, 1350), (8, 9, 750), (9, 10, 750), (9, 12, 300), (9, 13, 300),
    (11, 12, 600), (11, 13, 750), (12, 14, 300), (13, 14, 150),
)


@dataclasses.dataclass
class TopologyTables:
    name: str
    num_nodes: int
    num_links: int
    k_paths: int
    node_names: Tuple[str, ...]
    link_nodes: np.ndarray      # int32 [E,2] node indices (low, high) of each link index
    link_length: np.ndarray     # float64 [E]
    pair_first: np.ndarray      # int32 [N*N] first path row of (src,dst); -1 on the diagonal
    pair_count: np.ndarray      # int32 [N*N] number of candidate paths (<= k)
    path_hops: np.ndarray       # int32 [P]
    path_length: np.ndarray     # float64 [P]
    path_se: np.ndarray         # int32 [P] spectral efficiency of best_modulation
    path_mod: np.ndarray        # int32 [P] index of best_modulation in the modulation table
    path_link_ptr: np.ndarray   # int32 [P+1]
    path_links: np.ndarray      # int32 [sum hops] link indices, in hop order
    path_nodes: np.ndarray      # int32 [sum hops + P] node indices (CSR with ptr+row)
    mod_names: Tuple[str, ...]
    mod_se: np.ndarray          # int32 [M]
    mod_reach: np.ndarray       # float64 [M]
    mod_osnr: np.ndarray        # float64 [M]
    mod_xt: np.ndarray          # float64 [M]
    link_order: Optional[np.ndarray] = None   # int32 [E] link indices in nx.Graph.edges() iteration order (np.mean over links)

    def __post_init__(self):
        if self.link_order is None:
            self.link_order = np.arange(self.num_links, dtype=np.int32)

    @property
    def num_paths(self) -> int:
        return int(self.path_hops.shape[0])

    def links_of(self, path_row: int) -> np.ndarray:
        return self.path_links[self.path_link_ptr[path_row]:self.path_link_ptr[path_row + 1]]

    def rows_of(self, src: int, dst: int) -> range:
        first = int(self.pair_first[src * self.num_nodes + dst])
        return range(first, first + int(self.pair_count[src * self.num_nodes + dst]))

    # ------------------------------------------------------------------ persistence
    def save(self, file) -> None:
        d = {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}
        for key in ("node_names", "mod_names"):
            d[key] = np.array(d[key], dtype=np.str_)
        d["name"] = np.array(d["name"], dtype=np.str_)
        np.savez_compressed(file, **d)

    @classmethod
    def load(cls, file) -> "TopologyTables":
        with np.load(file, allow_pickle=False) as z:
            d = {k: z[k] for k in z.files}
        d.setdefault("link_order", None)
        d["name"] = str(d["name"])
        d["node_names"] = tuple(str(x) for x in d["node_names"])
        d["mod_names"] = tuple(str(x) for x in d["mod_names"])
        for key in ("num_nodes", "num_links", "k_paths"):
            d[key] = int(d[key])
        return cls(**d)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_graph(cls, graph) -> "TopologyTables":
        """From a reference-style ``nx.Graph`` (``graph.graph['ksp']`` etc.), e.g. an
        un-pickled ``*.h5`` of the reference.  Duck-typed: nothing of the reference is imported."""
        g = graph.graph
        node_names = tuple(g["node_indices"])
        nidx = {n: i for i, n in enumerate(node_names)}
        n = len(node_names)
        e = graph.number_of_edges()
        link_nodes = np.zeros((e, 2), np.int32)
        link_length = np.zeros(e, np.float64)
        for u, v, data in graph.edges(data=True):
            link_nodes[data["index"]] = (min(nidx[u], nidx[v]), max(nidx[u], nidx[v]))
            link_length[data["index"]] = float(data.get("length", 0.0))
        mods = tuple(g.get("modulations") or ())
        mod_key = [(m.name, m.spectral_efficiency) for m in mods]
        pair_first = np.full(n * n, -1, np.int32)
        pair_count = np.zeros(n * n, np.int32)
        hops, length, se, mod, ptr, links, nodes = [], [], [], [], [0], [], []
        for i in range(n):
            for j in range(i + 1, n):
                plist = g["ksp"][node_names[i], node_names[j]]
                first = len(hops)
                for p in plist:
                    hops.append(len(p.node_list) - 1)
                    length.append(float(p.length))
                    bm = p.best_modulation
                    se.append(int(bm.spectral_efficiency) if bm is not None else 1)
                    mod.append(mod_key.index((bm.name, bm.spectral_efficiency)) if bm is not None and mods else 0)
                    for a, b in zip(p.node_list[:-1], p.node_list[1:]):
                        links.append(int(graph[a][b]["index"]))
                    nodes.extend(nidx[x] for x in p.node_list)
                    ptr.append(len(links))
                for key in (i * n + j, j * n + i):
                    pair_first[key] = first
                    pair_count[key] = len(plist)
        return cls(
            name=str(g.get("name", "topology")), num_nodes=n, num_links=e, k_paths=int(g["k_paths"]),
            node_names=tuple(str(x) for x in node_names), link_nodes=link_nodes, link_length=link_length,
            pair_first=pair_first, pair_count=pair_count,
            path_hops=np.array(hops, np.int32), path_length=np.array(length, np.float64),
            path_se=np.array(se, np.int32), path_mod=np.array(mod, np.int32),
            path_link_ptr=np.array(ptr, np.int32), path_links=np.array(links, np.int32),
            path_nodes=np.array(nodes, np.int32),
            mod_names=tuple(m.name for m in mods),
            mod_se=np.array([m.spectral_efficiency for m in mods], np.int32),
            mod_reach=np.array([m.maximum_length for m in mods], np.float64),
            mod_osnr=np.array([m.minimum_osnr if m.minimum_osnr is not None else np.nan for m in mods], np.float64),
            mod_xt=np.array([m.inband_xt if m.inband_xt is not None else np.nan for m in mods], np.float64),
            link_order=np.array([data["index"] for _, _, data in graph.edges(data=True)], np.int32),
        )

    @classmethod
    def from_links(cls, name: str, num_nodes: int, links: Iterable[Tuple[int, int, float]], k_paths: int = 5,
                   modulations: Optional[Sequence[Tuple[str, float, int, float, float]]] = DEFAULT_MODULATIONS,
                   one_based: bool = True) -> "TopologyTables":
        """k-shortest-path pre-processing from a plain link list over nodes 1..N (or 0..N-1)."""
        off = 1 if one_based else 0
        node_names = tuple(str(i + off) for i in range(num_nodes))
        return cls.from_named_links(name, node_names, [(str(u), str(v), l) for u, v, l in links], k_paths, modulations)

    @classmethod
    def from_named_links(cls, name: str, node_names: Sequence[str], links: Sequence[Tuple[str, str, float]],
                         k_paths: int = 5,
                         modulations: Optional[Sequence[Tuple[str, float, int, float, float]]] = DEFAULT_MODULATIONS,
                         ) -> "TopologyTables":
        """``get_topology`` (examples/create_topology.py:96-147) on a node list + link list: NetworkX Yen's
        algorithm by length (``utils.get_k_shortest_paths``), path length as ``np.sum`` over the hops, the most
        spectrally efficient modulation whose reach covers it (``utils.get_best_modulation_format``).  Nodes and
        links are inserted in the given order, which is what fixes the tie-breaking between equal-length paths
        and the ``topology.edges()`` iteration order (``link_order``)."""
        import networkx as nx

        node_names = tuple(str(x) for x in node_names)
        nidx = {nm: i for i, nm in enumerate(node_names)}
        graph = nx.Graph()
        for nm in node_names:            # same insertion order as graph_utils.read_txt_file / read_sndlib_topology
            graph.add_node(nm)
        links = [(str(u), str(v), l) for u, v, l in links]
        for idx, (u, v, length) in enumerate(links):
            if u not in nidx or v not in nidx:
                raise ValueError("link %d joins an unknown node (%s, %s)" % (idx, u, v))
            graph.add_edge(u, v, index=idx, length=length)
        mods = tuple(modulations or ())
        by_se = sorted(range(len(mods)), key=lambda m: mods[m][2], reverse=True)
        n = len(node_names)
        pair_first = np.full(n * n, -1, np.int32)
        pair_count = np.zeros(n * n, np.int32)
        hops, length, se, mod, ptr, plinks, nodes = [], [], [], [], [0], [], []
        for i in range(n):
            for j in range(i + 1, n):
                paths = list(islice(nx.shortest_simple_paths(graph, node_names[i], node_names[j], weight="length"), k_paths))
                first = len(hops)
                for p in paths:
                    plen = float(np.sum([graph[a][b]["length"] for a, b in zip(p[:-1], p[1:])]))
                    best = 0
                    if mods:
                        best = next((m for m in by_se if plen <= mods[m][1]), None)
                        if best is None:
                            raise ValueError("It was not possible to find a suitable MF for a path with %s km" % plen)
                    hops.append(len(p) - 1)
                    length.append(plen)
                    se.append(int(mods[best][2]) if mods else 1)
                    mod.append(best)
                    plinks.extend(int(graph[a][b]["index"]) for a, b in zip(p[:-1], p[1:]))
                    nodes.extend(nidx[x] for x in p)
                    ptr.append(len(plinks))
                for key in (i * n + j, j * n + i):
                    pair_first[key] = first
                    pair_count[key] = len(paths)
        return cls(
            name=name, num_nodes=n, num_links=len(links), k_paths=k_paths, node_names=node_names,
            link_nodes=np.array([(min(nidx[u], nidx[v]), max(nidx[u], nidx[v])) for u, v, _ in links], np.int32).reshape(-1, 2),
            link_length=np.array([l for _, _, l in links], np.float64),
            pair_first=pair_first, pair_count=pair_count,
            path_hops=np.array(hops, np.int32), path_length=np.array(length, np.float64),
            path_se=np.array(se, np.int32), path_mod=np.array(mod, np.int32),
            path_link_ptr=np.array(ptr, np.int32), path_links=np.array(plinks, np.int32),
            path_nodes=np.array(nodes, np.int32),
            mod_names=tuple(m[0] for m in mods),
            mod_se=np.array([m[2] for m in mods], np.int32),
            mod_reach=np.array([m[1] for m in mods], np.float64),
            mod_osnr=np.array([m[3] for m in mods], np.float64),
            mod_xt=np.array([m[4] for m in mods], np.float64),
            link_order=np.array([data["index"] for _, _, data in graph.edges(data=True)], np.int32),
        )


# ---------------------------------------------------------------------- topology files
def read_txt_file(file) -> Tuple[Tuple[str, ...], list]:
    """The reference's ``.txt`` format (examples/graph_utils.py:89-116): lines starting with ``#`` are comments,
    then the node count, the link count, and one ``src dst length_km`` line per link; nodes are named "1".."N"
    and link ``index`` is the line order.  Returns (node names, [(src, dst, length)])."""
    with open(file, "r") as fh:
        lines = [ln for ln in fh if not ln.startswith("#")]
    names: Tuple[str, ...] = ()
    links = []
    for idx, line in enumerate(lines):
        if idx == 0:
            names = tuple(str(i) for i in range(1, int(line) + 1))
        elif idx == 1:
            int(line)                      # declared link count (the reference reads and ignores it)
        elif len(line) > 1:
            info = line.replace("\n", "").split(" ")
            links.append((info[0], info[1], int(info[2])))
    return names, links


def _geographical_distance(latlong1, latlong2) -> float:
    """examples/graph_utils.py:10-28 (haversine, R = 6373 km; note the reference feeds (x, y) = (lon, lat))."""
    import math

    R = 6373.0
    lat1, lon1 = math.radians(latlong1[0]), math.radians(latlong1[1])
    lat2, lon2 = math.radians(latlong2[0]), math.radians(latlong2[1])
    dlon, dlat = lon2 - lon1, lat2 - lat1
    a = math.sin(dlat / 2) ** 2 + math.cos(lat1) * math.cos(lat2) * math.sin(dlon / 2) ** 2
    c = 2 * math.atan2(math.sqrt(a), math.sqrt(1 - a))
    return R * c


def read_sndlib_topology(file) -> Tuple[Tuple[str, ...], list]:
    """SNDlib native XML (examples/graph_utils.py:31-86): nodes in document order with (x, y) coordinates, link
    length = haversine distance (``coordinatesType="geographical"``) or the Euclidean one, rounded to 3 decimals."""
    import math
    import xml.dom.minidom

    doc = xml.dom.minidom.parse(file).documentElement
    ctype = doc.getElementsByTagName("nodes")[0].getAttribute("coordinatesType")
    pos, names = {}, []
    for node in doc.getElementsByTagName("node"):
        x = float(node.getElementsByTagName("x")[0].childNodes[0].data)
        y = float(node.getElementsByTagName("y")[0].childNodes[0].data)
        names.append(node.getAttribute("id"))
        pos[names[-1]] = (x, y)
    links = []
    for link in doc.getElementsByTagName("link"):
        s = link.getElementsByTagName("source")[0].childNodes[0].data
        t = link.getElementsByTagName("target")[0].childNodes[0].data
        if ctype == "geographical":
            length = np.around(_geographical_distance(pos[s], pos[t]), 3)
        else:
            length = np.around(math.sqrt((pos[s][0] - pos[t][0]) ** 2 + (pos[s][1] - pos[t][1]) ** 2), 3)
        links.append((s, t, float(length)))
    return tuple(names), links


def get_topology(file_name, topology_name: Optional[str] = None, modulations=DEFAULT_MODULATIONS, k_paths: int = 5,
                 cache_dir: Optional[str] = None) -> TopologyTables:
    """``examples/create_topology.py:get_topology`` -> flat tables.  ``cache_dir`` keeps the k-shortest-path
    result on disk (keyed by file contents, k and the modulation table): Yen's algorithm on a 100-node graph
    takes about a minute of host time."""
    import hashlib
    import os

    file_name = str(file_name)
    if file_name.endswith(".xml"):
        names, links = read_sndlib_topology(file_name)
    elif file_name.endswith(".txt"):
        names, links = read_txt_file(file_name)
    else:
        raise ValueError("Supplied topology is unknown")
    name = topology_name or os.path.splitext(os.path.basename(file_name))[0]
    cache = None
    if cache_dir:
        cache_dir = os.path.expanduser(str(cache_dir))
        with open(file_name, "rb") as fh:
            key = hashlib.sha256(fh.read() + repr((k_paths, tuple(modulations or ()))).encode()).hexdigest()[:16]
        cache = os.path.join(cache_dir, "%s_k%d_%s.npz" % (name, k_paths, key))
        if os.path.exists(cache):
            return TopologyTables.load(cache)
    tables = TopologyTables.from_named_links(name, names, links, k_paths=k_paths, modulations=modulations)
    if cache:
        os.makedirs(cache_dir, exist_ok=True)
        tables.save(cache)
    return tables


def load_reference_pickle(file) -> TopologyTables:
    """Un-pickles a topology written by the reference's ``create_topology.py`` (the shipped ``*.h5`` files are
    pickles of an ``nx.Graph`` whose ``ksp`` entries are ``optical_rl_gym.utils.Path`` / ``Modulation``
    dataclasses, utils.py:14-59).  The reference package is not needed: stand-in classes with the same module
    path are registered for the duration of the load."""
    import pickle
    import sys
    import types

    stubs = {}
    if "optical_rl_gym.utils" not in sys.modules:
        @dataclasses.dataclass
        class Modulation:
            name: str
            maximum_length: float
            spectral_efficiency: int
            minimum_osnr: Optional[float] = None
            inband_xt: Optional[float] = None

        @dataclasses.dataclass
        class Path:
            path_id: int
            node_list: Tuple[str]
            hops: int
            length: float
            best_modulation: Optional[Modulation] = None
            current_modulation: Optional[Modulation] = None

        pkg = sys.modules.get("optical_rl_gym") or types.ModuleType("optical_rl_gym")
        utils = types.ModuleType("optical_rl_gym.utils")
        utils.Modulation, utils.Path = Modulation, Path
        Modulation.__module__ = Path.__module__ = "optical_rl_gym.utils"
        stubs = {"optical_rl_gym": pkg, "optical_rl_gym.utils": utils}
        added = [k for k in stubs if k not in sys.modules]
        sys.modules.update({k: stubs[k] for k in added})
    else:
        added = []
    try:
        with open(file, "rb") as fh:
            graph = pickle.load(fh)
    finally:
        for k in added:
            sys.modules.pop(k, None)
    return TopologyTables.from_graph(graph)


def nsfnet(k_paths: int = 5) -> TopologyTables:
    """NSFNET (14 nodes / 22 links) with the reference's 6-modulation table; equals the
    tables extracted from the reference's ``nsfnet_chen_5-paths_6-modulations.h5``
    (checked in tests/test_topology.py against tests/golden/nsfnet_tables.npz)."""
    return TopologyTables.from_links("nsfnet_chen", 14, NSFNET_LINKS, k_paths=k_paths)


def synthetic_ring_chords(num_nodes: int = 100, num_chords: int = 200, k_paths: int = 10, seed: int = 1,
                          min_len: int = 100, max_len: int = 1000) -> TopologyTables:
    """Seeded 2-edge-connected synthetic graph (ring + random chords) for the large
    config of BASELINE.json (100 nodes / 300 links)."""
    rng = np.random.default_rng(seed)
    links, seen = [], set()
    for i in range(num_nodes):
        j = (i + 1) % num_nodes
        links.append((i, j, int(rng.integers(min_len, max_len + 1))))
        seen.add((min(i, j), max(i, j)))
    while len(links) < num_nodes + num_chords:
        i, j = (int(x) for x in rng.integers(0, num_nodes, 2))
        key = (min(i, j), max(i, j))
        if i == j or key in seen:
            continue
        seen.add(key)
        links.append((i, j, int(rng.integers(min_len, max_len + 1))))
    return TopologyTables.from_links("ring%d_chords%d" % (num_nodes, num_chords), num_nodes, links,
                                     k_paths=k_paths, one_based=False)
