"""The host-tree bindings live in the package (montecarlotreesearch_b200/hosttree.py); the tests use them from here."""
from montecarlotreesearch_b200.hosttree import CB, Session, load  # noqa: F401
