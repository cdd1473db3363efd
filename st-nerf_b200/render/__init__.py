"""`render` as the reference names it (`from render import LayeredNeuralRenderer`, demo/taekwondo_demo.py:23)."""
from .renderer import LayeredNeuralRenderer

__all__ = ["LayeredNeuralRenderer"]
