"""Tuning knobs of the tick must not change a single bit of its results (run last: nothing here is on the parity path)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_prefetch_mode_does_not_change_results(cuda_device):
    """``t2d_set_prefetch``: the early L2 prefetch of the tick's inputs on, off, or left to the library's policy - five ticks
    each from the same state with the same actions give identical state, flags, hit indices, status and done."""
    import torch

    from tactics2d_b200 import BatchedWorld, synthetic
    from tactics2d_b200._lib import T2DError

    sc = synthetic.config2(512, 64, seed=61)
    act = torch.from_numpy(synthetic.random_actions(62, sc.shape)).to(cuda_device)
    seen = {}
    for mode in (0, 1, -1):
        w = BatchedWorld(*sc.shape, sc.table, device=cuda_device)
        w.set_map(sc.segments, sc.bounds)
        w.set_state(sc.x, sc.y, sc.heading, sc.speed, type_id=sc.type_id)
        w.set_prefetch(mode)
        for _ in range(5):
            r = w.step(act)
        torch.cuda.synchronize()
        st = w.state_numpy()
        seen[mode] = [st[k].copy() for k in ("x", "y", "heading", "speed", "vx", "vy")] + \
                     [getattr(r, k).cpu().numpy().copy() for k in ("flags", "hit_index", "hit_segment", "status", "done")]
        if mode == -1:
            with pytest.raises(T2DError):
                w.set_prefetch(2)
        w.close()
    for mode in (1, -1):
        for a, b in zip(seen[0], seen[mode]):
            assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)
    assert seen[0][6].any()          # the comparison saw events, not an empty scene
