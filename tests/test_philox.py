"""Known-answer tests of the Philox4x32-10 restatement in the oracle (Random123 kat_vectors) and of the
uniform-stream layout the CUDA kernels implement (DESIGN.md "Philox layout")."""
import numpy as np
import torch

from oracle import ref_oracle as O


def _hex(a):
    return [f"{int(x):08x}" for x in a]


def test_random123_known_answers():
    z = O.philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    assert _hex(z) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    f = O.philox4x32_10(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    assert _hex(f) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    p = O.philox4x32_10(np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32),
                        np.array([0xA4093822, 0x299F31D0], np.uint32))
    assert _hex(p) == ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_uniform_stream_properties():
    for dt in (torch.float32, torch.float64):
        u = O.philox_uniform(123, 0, 0, 20000, 5, dt)
        assert u.dtype == dt and u.shape == (20000, 5)
        assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
        assert abs(float(u.double().mean()) - 0.5) < 0.01 and abs(float(u.double().var()) - 1 / 12) < 0.005
        # pure function of (seed, call, row, d): row ranges compose, calls and seeds differ
        assert torch.equal(u[700:900], O.philox_uniform(123, 0, 700, 200, 5, dt))
        assert not torch.equal(u, O.philox_uniform(123, 1, 0, 20000, 5, dt))
        assert not torch.equal(u, O.philox_uniform(124, 0, 0, 20000, 5, dt))
    # rows beyond 2^32 use the high counter word
    a = O.philox_uniform(1, 0, 2**32 - 2, 4, 3, torch.float32)
    b = O.philox_uniform(1, 0, 2**32, 2, 3, torch.float32)
    assert torch.equal(a[2:], b)
