"""Deterministic parameter values shared by tests/golden/make_golden_exp1.py (which loads them into the LIVE reference
modules) and the tests (which load them into the B200 modules), so the fixture does not have to carry state dicts."""
import numpy as np


def make_state(shapes, seed):
    """shapes: {state_dict key: shape} -> {key: float32 array}.  Keys that name a shared tensor (the word embedding
    of every text encoder, the category embedding of every element encoder) get the same array."""
    rng = np.random.default_rng(seed)
    shared = {}
    out = {}
    for k in sorted(shapes):
        shape = tuple(shapes[k])
        leaf = k.rsplit(".", 2)[-2] + "." + k.rsplit(".", 1)[-1] if k.endswith("embedding.weight") else None
        if leaf is not None:
            if leaf not in shared:
                a = (rng.standard_normal(shape) * 0.5).astype(np.float32)
                a[0] = 0.0                                     # padding_idx row
                shared[leaf] = a
            out[k] = shared[leaf]
        elif k.endswith("position_embedding") or k.endswith("attention_query_vector"):
            out[k] = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        elif len(shape) == 2:
            out[k] = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        else:
            out[k] = rng.uniform(-0.06, 0.06, size=shape).astype(np.float32)
    return out
