"""Dag = one GSL query (graphlearn/python/gsl/dag.py:25-130): a registry of aliased
nodes plus the root; ``set_ready`` freezes it into an executable plan."""
from __future__ import annotations

import threading

_LOCK = threading.Lock()
_COUNT = [0]


def get_dag_name():
    with _LOCK:
        _COUNT[0] += 1
        return "dag_%d" % _COUNT[0]


class Dag(object):
    def __init__(self, graph):
        self.graph = graph
        self.name = get_dag_name()
        self.root = None
        self._alias_to_node = {}
        self._aliases = []
        self._ready = False
        self._value_func = None
        self._next_id = 0

    def next_id(self):
        self._next_id += 1
        return self._next_id

    def add_node(self, alias, node, temp=False):
        if alias in self._alias_to_node and self._alias_to_node[alias] is not node:
            raise ValueError("alias {!r} is already used in this query.".format(alias))
        self._alias_to_node[alias] = node
        if alias not in self._aliases:
            self._aliases.append(alias)

    def get_node(self, alias):
        node = self._alias_to_node.get(alias)
        if node is None:
            raise ValueError("alias {!r} not found in the query.".format(alias))
        return node

    def list_alias(self):
        return list(self._aliases)

    @property
    def node_types(self):
        return sorted({n.type for n in self._alias_to_node.values() if isinstance(n.type, str)})

    def is_ready(self):
        return self._ready

    def set_ready(self, func=lambda x: x):
        if self.root is None:
            raise ValueError("query has no root (start it with g.V() or g.E()).")
        self._ready = True
        self._value_func = func

    @property
    def value_func(self):
        return self._value_func

    def __str__(self):
        return "Dag(%s, aliases=%s)" % (self.name, self._aliases)
