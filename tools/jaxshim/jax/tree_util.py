from ._core import tree_leaves, tree_map  # noqa: F401


def register_pytree_node_class(cls):
    return cls


def register_pytree_node(*a, **k):
    pass
