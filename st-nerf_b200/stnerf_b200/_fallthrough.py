"""Fall-through of the facade packages to the reference tree for the submodules they do not replace.

`st-nerf_b200/{utils,layers,engine}` carry the reference's import names, so with `st-nerf_b200/` ahead of the reference root
on `sys.path` they shadow the reference's packages of the same name.  The reference's demos also import submodules the hot
path does not touch (`engine.layered_trainer`, `utils.metrics`, `utils.render_helpers`, ... demo/taekwondo_demo.py:16-23).
Each facade package therefore appends the reference's directory of the same name to its `__path__`: a submodule that exists
here wins, anything else resolves to the reference's file.

The reference root is `$STNERF_REFERENCE_ROOT` if set, else the first `sys.path` entry that holds
`modeling/layered_rfrender.py` (the demos put it there themselves: `sys.path.append('.')`, demo/taekwondo_demo.py:15).
Without a reference tree the facade packages simply stand alone.
"""
from __future__ import annotations

import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # .../st-nerf_b200
_MARKER = os.path.join("modeling", "layered_rfrender.py")
_registered = {}                                                              # package name -> its __path__ list


def reference_root():
    """Directory of the reference checkout this process can see, or None."""
    env = os.environ.get("STNERF_REFERENCE_ROOT")
    if env:
        return os.path.abspath(env) if os.path.isfile(os.path.join(env, _MARKER)) else None
    for entry in sys.path:
        d = os.path.abspath(entry or ".")
        if d == _PKG_ROOT:
            continue
        if os.path.isfile(os.path.join(d, _MARKER)):
            return d
    return None


def extend(name: str, path: list) -> list:
    """Append `<reference root>/<name>` to a facade package's `__path__` (idempotent).  Returns `path`."""
    _registered[name] = path
    root = reference_root()
    if root is not None:
        d = os.path.join(root, name)
        if os.path.isdir(d) and d not in path:
            path.append(d)
    return path


def refresh():
    """Re-run `extend` for every facade package (for callers that add the reference root to sys.path after importing them)."""
    for name, path in list(_registered.items()):
        extend(name, path)
