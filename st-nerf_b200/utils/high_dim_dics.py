"""utils/high_dim_dics.py of the reference: nested-dict setters used by render/neural_renderer.py."""


def add_two_dim_dict(adic, key_a, key_b, val):
    adic.setdefault(key_a, {})[key_b] = val


def add_three_dim_dict(adic, key_a, key_b, key_c, val):
    adic.setdefault(key_a, {}).setdefault(key_b, {})[key_c] = val
