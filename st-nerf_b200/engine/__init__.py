"""Reference import name `engine`: `render` (engine/render.py) runs on the native path; `engine.layered_trainer` and anything
else falls through to the reference tree when one is on sys.path (stnerf_b200/_fallthrough.py) -- training is out of scope."""
from stnerf_b200 import _fallthrough

_fallthrough.extend("engine", __path__)

from .render import render  # noqa: E402

__all__ = ["render"]
