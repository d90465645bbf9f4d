"""CPU tests of host-side logic that needs no kernel: the level-gradient accumulator of a training clip
(lgu-slam_b200/corr.py: LevelGradAccumulator) -- ownership by identity, lazy zero buffers, hand-over, and no strong
reference from the accumulator to the pyramid (the autograd graph points at the accumulator)."""
import gc
import weakref
from importlib import import_module

import torch


def _acc():
    import lgu_slam_b200  # noqa: F401  (registers the package alias; importing corr does not load the CUDA library)
    return import_module("lgu-slam_b200.corr").LevelGradAccumulator()


def test_accumulator_owns_exactly_its_pyramid():
    acc = _acc()
    pyr = tuple(torch.zeros(2, 3, 4, 3 >> 0, 4 >> 0) for _ in range(4))
    assert not acc.owns(pyr)                       # nothing registered yet
    acc.levels = pyr
    assert acc.owns(pyr)
    clones = tuple(t.clone() for t in pyr)         # equal values, different tensors (CorrBlock.cat / __getitem__)
    assert not acc.owns(clones)
    assert not acc.owns(pyr[:3] + (clones[3],))


def test_accumulator_buffers_are_lazy_zeroed_and_handed_over_once():
    acc = _acc()
    pyr = tuple(torch.randn(1, 2, 2, 2 >> l, 2 >> l) for l in range(2))
    acc.levels = pyr
    assert acc.grads is None
    bufs = acc.buffers()
    assert [b.shape for b in bufs] == [p.shape for p in pyr] and all(float(b.abs().sum()) == 0 for b in bufs)
    bufs[0] += 1
    assert acc.buffers()[0] is bufs[0]             # the same persistent buffers on every call of the clip
    taken = acc.take()
    assert taken[0] is bufs[0] and acc.grads is None and acc.take() is None
    assert float(acc.buffers()[0].abs().sum()) == 0    # a second backward (retain_graph) starts from zero again


def test_accumulator_does_not_keep_the_pyramid_alive():
    acc = _acc()
    pyr = [torch.zeros(1, 1, 1, 1, 1) for _ in range(4)]
    refs = [weakref.ref(t) for t in pyr]
    acc.levels = tuple(pyr)
    del pyr
    gc.collect()
    assert all(r() is None for r in refs)
    assert not acc.owns(tuple(torch.zeros(1) for _ in range(4)))
