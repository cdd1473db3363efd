from .render import render

__all__ = ["render"]
