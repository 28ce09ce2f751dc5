"""TEST INFRASTRUCTURE: makes every dropout of a PyTorch model deterministic and order-dependent, so that the oracle's restatement
of WHERE the reference applies dropout can be pinned against the real reference (oracle/make_golden.py, train-mode fixtures):
``torch.nn.functional.dropout`` -- which nn.Dropout, F.multi_head_attention_forward and HuggingFace's eager attention all call -- is
replaced by a function whose k-th call draws its mask from a generator seeded with (seed, k).  Two implementations produce the same
outputs under it only if they call dropout the same number of times, in the same order, on tensors of the same shape and layout."""
import contextlib

import torch
import torch.nn.functional as F


@contextlib.contextmanager
def deterministic_dropout(seed):
    state = {"n": 0, "log": []}
    orig = F.dropout

    def dropout(input, p=0.5, training=True, inplace=False):
        if not training or p <= 0.0:
            return input
        k = state["n"]
        state["n"] += 1
        g = torch.Generator().manual_seed(seed * 7919 + k)
        keep = (torch.rand(input.shape, generator=g) >= p).to(input.dtype).to(input.device)
        state["log"].append((k, tuple(input.shape), float(p)))
        return input * keep / (1.0 - p)

    F.dropout = dropout
    try:
        yield state
    finally:
        F.dropout = orig
