"""Pure-Python stand-in for python-igraph (TEST INFRASTRUCTURE ONLY).

Implements the small directed-multigraph surface the reference touches
(CDSimulatorComponents.py:613-748, volt_typhoon_env.py:456-511,
CyberDefenseEnv.py:583-875).  Behavioural conventions kept from python-igraph:

* vertices are 0..n-1; `neighbors()` / `get_adjlist()` return neighbour vertex
  ids in ascending order and repeat a neighbour once per parallel edge;
* `get_eid(error=False)` returns -1 when the edge is missing;
* `Barabasi(n, m, directed=True)` grows a graph where each new vertex sends up
  to `m` edges to distinct earlier vertices chosen with probability
  proportional to (in-degree + 1) -- the "psumtree" rule without multi-edges.

The stand-in only decides WHICH topology gets generated (an input of the step
path).  The step arithmetic itself always runs in the unmodified reference.
"""
import random as _random


class _Vertex:
    __slots__ = ("_g", "index")

    def __init__(self, g, index):
        self._g = g
        self.index = index

    def __getitem__(self, key):
        return self._g._vattr[key][self.index]

    def __setitem__(self, key, value):
        col = self._g._vattr.get(key)
        if col is None:
            col = self._g._vattr[key] = [None] * self._g._n
        col[self.index] = value

    def get(self, key, default=None):
        col = self._g._vattr.get(key)
        if col is None:
            return default
        v = col[self.index]
        return default if v is None else v

    def attributes(self):
        return {k: col[self.index] for k, col in self._g._vattr.items()}


class _VertexSeq:
    def __init__(self, g):
        self._g = g

    def __len__(self):
        return self._g._n

    def __iter__(self):
        for i in range(self._g._n):
            yield _Vertex(self._g, i)

    def __getitem__(self, key):
        if isinstance(key, str):
            return list(self._g._vattr[key])
        if isinstance(key, int):
            if key < 0 or key >= self._g._n:
                raise IndexError(key)
            return _Vertex(self._g, key)
        raise TypeError(key)

    def __setitem__(self, key, values):
        if not isinstance(key, str):
            raise TypeError(key)
        values = list(values)
        if len(values) != self._g._n:
            raise ValueError("attribute list length mismatch")
        self._g._vattr[key] = values

    def attributes(self):
        return list(self._g._vattr.keys())

    def find(self, name=None, **kw):
        col = self._g._vattr.get("name")
        if col is None:
            raise ValueError("no such vertex")
        for i, v in enumerate(col):
            if v == name:
                return _Vertex(self._g, i)
        raise ValueError("no such vertex")


class _Edge:
    __slots__ = ("_g", "index")

    def __init__(self, g, index):
        self._g = g
        self.index = index

    def __getitem__(self, key):
        return self._g._eattr[key][self.index]

    def __setitem__(self, key, value):
        col = self._g._eattr.get(key)
        if col is None:
            col = self._g._eattr[key] = [None] * len(self._g._edges)
        col[self.index] = value

    @property
    def tuple(self):
        return self._g._edges[self.index]

    @property
    def source(self):
        return self._g._edges[self.index][0]

    @property
    def target(self):
        return self._g._edges[self.index][1]


class _EdgeSlice:
    def __init__(self, g, idxs):
        self._g = g
        self._idxs = idxs

    def __setitem__(self, key, values):
        col = self._g._eattr.get(key)
        if col is None:
            col = self._g._eattr[key] = [None] * len(self._g._edges)
        values = list(values)
        for i, v in zip(self._idxs, values):
            col[i] = v


class _EdgeSeq:
    def __init__(self, g):
        self._g = g

    def __len__(self):
        return len(self._g._edges)

    def __iter__(self):
        for i in range(len(self._g._edges)):
            yield _Edge(self._g, i)

    def __getitem__(self, key):
        if isinstance(key, str):
            return list(self._g._eattr[key])
        if isinstance(key, slice):
            return _EdgeSlice(self._g, list(range(*key.indices(len(self._g._edges)))))
        if isinstance(key, int):
            if key < 0:
                key += len(self._g._edges)
            if key < 0 or key >= len(self._g._edges):
                raise IndexError(key)
            return _Edge(self._g, key)
        raise TypeError(key)

    def __setitem__(self, key, values):
        if not isinstance(key, str):
            raise TypeError(key)
        values = list(values)
        if len(values) != len(self._g._edges):
            raise ValueError("attribute list length mismatch")
        self._g._eattr[key] = values

    def attributes(self):
        return list(self._g._eattr.keys())


class Graph:
    def __init__(self, n=0, edges=None, directed=False):
        self._n = int(n)
        self._directed = bool(directed)
        self._edges = []
        self._vattr = {}
        self._eattr = {}
        self._adj = {}          # mode -> cached sorted adjacency (dropped on mutation)
        if edges:
            self.add_edges(edges)

    # -- construction ------------------------------------------------------
    @classmethod
    def Barabasi(cls, n, m=1, directed=False, **_):
        g = cls(n=n, directed=directed)
        indeg = [0] * n
        edges = []
        for v in range(1, n):
            k = min(m, v)
            weights = [indeg[u] + 1 for u in range(v)]
            chosen = []
            cand = list(range(v))
            for _ in range(k):
                tot = sum(weights)
                r = _random.random() * tot
                acc = 0.0
                pick = len(cand) - 1
                for j, w in enumerate(weights):
                    acc += w
                    if r < acc:
                        pick = j
                        break
                chosen.append(cand.pop(pick))
                weights.pop(pick)
            for u in chosen:
                edges.append((v, u))
                indeg[u] += 1
        g.add_edges(edges)
        return g

    def is_directed(self):
        return self._directed

    def vcount(self):
        return self._n

    def ecount(self):
        return len(self._edges)

    @property
    def vs(self):
        return _VertexSeq(self)

    @property
    def es(self):
        return _EdgeSeq(self)

    def _vid(self, v):
        if isinstance(v, _Vertex):
            return v.index
        if isinstance(v, bool):
            raise TypeError(v)
        if isinstance(v, int):
            if v < 0 or v >= self._n:
                raise ValueError(f"vertex index out of range: {v}")
            return v
        if hasattr(v, "__index__"):
            return self._vid(int(v))
        # name lookup (strings)
        col = self._vattr.get("name")
        if col is not None:
            for i, nm in enumerate(col):
                if nm == v:
                    return i
        raise ValueError(f"no such vertex: {v!r}")

    def add_vertices(self, n):
        if isinstance(n, int):
            names = None
            cnt = n
        else:
            names = list(n)
            cnt = len(names)
        start = self._n
        self._n += cnt
        for col in self._vattr.values():
            col.extend([None] * cnt)
        if names is not None:
            col = self._vattr.get("name")
            if col is None:
                col = self._vattr["name"] = [None] * self._n
            for i, nm in enumerate(names):
                col[start + i] = nm

    def add_edges(self, pairs):
        pairs = [(self._vid(u), self._vid(v)) for (u, v) in pairs]
        self._adj = {}
        self._edges.extend(pairs)
        for col in self._eattr.values():
            col.extend([None] * len(pairs))

    def add_edge(self, u, v, **kw):
        self.add_edges([(u, v)])
        for k, val in kw.items():
            self.es[len(self._edges) - 1][k] = val

    def delete_edges(self, eids):
        if isinstance(eids, int):
            eids = [eids]
        dead = set(int(e) for e in eids)
        if not dead:
            return
        self._adj = {}
        keep = [i for i in range(len(self._edges)) if i not in dead]
        self._edges = [self._edges[i] for i in keep]
        for k, col in self._eattr.items():
            self._eattr[k] = [col[i] for i in keep]

    # -- queries -----------------------------------------------------------
    @staticmethod
    def _mode(mode):
        if isinstance(mode, str):
            return mode.lower()
        return {1: "out", 2: "in", 3: "all"}.get(mode, "all")

    def neighbors(self, vertex, mode="all"):
        v = self._vid(vertex)
        return list(self._adjlist(self._mode(mode))[v])

    def get_adjlist(self, mode="out"):
        return [list(r) for r in self._adjlist(self._mode(mode))]

    def _adjlist(self, mode):
        adj = self._adj.get(mode)
        if adj is not None and len(adj) == self._n:
            return adj
        adj = [[] for _ in range(self._n)]
        for (a, b) in self._edges:
            if mode in ("out", "all"):
                adj[a].append(b)
            if mode in ("in", "all"):
                adj[b].append(a)
        for row in adj:
            row.sort()
        self._adj[mode] = adj
        return adj

    def incident(self, vertex, mode="out"):
        v = self._vid(vertex)
        mode = self._mode(mode)
        out = []
        for i, (a, b) in enumerate(self._edges):
            if (mode in ("out", "all") and a == v) or (mode in ("in", "all") and b == v):
                out.append(i)
        return out

    def degree(self, vertices=None, mode="all", loops=True):
        mode = self._mode(mode)
        deg = [0] * self._n
        for (a, b) in self._edges:
            if mode in ("out", "all"):
                deg[a] += 1
            if mode in ("in", "all"):
                deg[b] += 1
        if vertices is None:
            return deg
        if isinstance(vertices, (list, tuple)):
            return [deg[self._vid(v)] for v in vertices]
        return deg[self._vid(vertices)]

    def get_eid(self, v1, v2, directed=True, error=True):
        try:
            a, b = self._vid(v1), self._vid(v2)
        except ValueError:
            if error:
                raise
            return -1
        for i, (x, y) in enumerate(self._edges):
            if (x == a and y == b) or (not (directed and self._directed) and x == b and y == a):
                return i
        if error:
            raise ValueError("no such edge")
        return -1

    def get_edgelist(self):
        return list(self._edges)

    def copy(self):
        g = Graph(n=self._n, directed=self._directed)
        g._edges = list(self._edges)
        g._vattr = {k: list(v) for k, v in self._vattr.items()}
        g._eattr = {k: list(v) for k, v in self._eattr.items()}
        g._adj = {}
        return g
