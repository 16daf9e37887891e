"""Minimal `pyrootutils` stand-in: setup_root(..., pythonpath=True) puts the project root on sys.path.
The reference tree ships no `.project-root` marker, so the root falls back to
$SEEDSTORY_PROJECT_ROOT or the current working directory."""
import os
import sys


def find_root(search_from=".", indicator=".project-root"):
    d = os.path.abspath(search_from if os.path.isdir(search_from) else os.path.dirname(search_from))
    while True:
        if os.path.exists(os.path.join(d, indicator)):
            return d
        parent = os.path.dirname(d)
        if parent == d:
            return None
        d = parent


def setup_root(search_from=".", indicator=".project-root", pythonpath=True, cwd=False, dotenv=False, **kw):
    root = os.environ.get("SEEDSTORY_PROJECT_ROOT") or find_root(search_from, indicator) or os.getcwd()
    if pythonpath and root not in sys.path:
        sys.path.insert(0, root)
    if cwd:
        os.chdir(root)
    os.environ.setdefault("PROJECT_ROOT", root)
    return root
