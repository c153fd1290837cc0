"""Importable alias of the `no-time-to-train_b200/` package (its directory name is not an identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("no-time-to-train_b200")
