"""`render` as the reference names it (`from render import LayeredNeuralRenderer`, demo/taekwondo_demo.py:23)."""
from .layered_neural_renderer import LayeredNeuralRenderer

__all__ = ["LayeredNeuralRenderer"]
