"""CPU: the host mirror of Gym.Spaces -- the reference's own BoxTest (tests/Gym.Tests/Spaces/BoxTest.cs:15-42)
restated, plus Discrete semantics (src/Gym/Spaces/Discrete.cs)."""
import numpy as np
import pytest

from gymnet_b200 import Box, Discrete


def test_box_bounded_test():   # BoxTest.TestBoxBoundedTest
    box = Box(-5.0, 5.0, dtype=np.float32)
    assert box.IsBounded("Both")
    box = Box(-np.inf, 5.0, dtype=np.float32)
    assert not box.IsBounded("Below") and box.IsBounded("Above") and not box.IsBounded("Both")
    box = Box(5.0, np.inf, dtype=np.float32)
    assert not box.IsBounded("Above") and box.IsBounded("Below") and not box.IsBounded("Both")
    box = Box(-np.inf, np.inf, dtype=np.float32)
    assert not box.IsBounded("Above") and not box.IsBounded("Below") and not box.IsBounded("Both")


def test_box_bounded_sampling():   # BoxTest.TestBoxBoundedSampling
    box = Box(-5.0, 5.0)
    sample = float(box.Sample(None)[0])
    assert -5.0 <= sample <= 5.0
    with pytest.raises(NotImplementedError):
        box.Sample(mask=np.ones(1))          # Box.cs:70-73


def test_box_contains_and_mixed_bounds():
    box = Box(np.array([-1.0, -np.inf, 0.0], np.float32), np.array([1.0, 0.0, np.inf], np.float32))
    box.Seed(0)
    s = box.Sample()
    assert s.dtype == np.float32 and -1 <= s[0] <= 1
    assert box.Contains(np.array([0.0, -3.0, 9.0], np.float32))
    assert not box.Contains(np.array([2.0, -3.0, 9.0], np.float32))
    with pytest.raises(NotImplementedError):
        box.Contains([0.0, 0.0, 0.0])        # NotSupportedException (Box.cs:95)


def test_discrete():
    d = Discrete(4, seed=1)
    xs = [d.Sample() for _ in range(200)]
    assert set(xs) == {0, 1, 2, 3}
    assert d.Contains(0) and d.Contains(3) and not d.Contains(4) and not d.Contains(-1)
    assert d.Sample(mask=np.array([0, 0, 1, 0])) == 2 and d.Sample(mask=np.zeros(4)) == 0   # Discrete.cs:19-25
    with pytest.raises(NotImplementedError):
        d.Contains(1.5)
    assert Discrete(3, start=10, seed=0).Sample() in (10, 11, 12)
