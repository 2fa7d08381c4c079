"""Host-side mirror of Gym.Spaces (src/Gym/Spaces/Space.cs, Box.cs, Discrete.cs).

Metadata only: the spaces describe ONE env instance, exactly like the reference's VecEnv keeps
the single-env ActionSpace/ObservationSpace (src/Gym/Envs/VecEnv.cs:13-19).  The random policy on
the hot path is sampled inside the CUDA rollout kernel; Sample() here is the host convenience
with the reference's distribution semantics (its NumSharp stream itself is not reproducible).
"""
import numpy as np


class Space:
    """Gym.Spaces.Space (Space.cs:5-18): Shape, DType, Sample(mask), Contains(x), Seed(seed)."""

    def __init__(self, shape, dtype):
        self.Shape = tuple(shape) if shape is not None else None
        self.DType = np.dtype(dtype)

    def Sample(self, mask=None):
        raise NotImplementedError

    def Contains(self, x):
        raise NotImplementedError

    def Seed(self, seed):
        raise NotImplementedError


class Box(Space):
    """Gym.Spaces.Box (Box.cs:15-100)."""

    BOTH, BELOW, ABOVE = "Both", "Below", "Above"   # BoundedMannerEnum (Box.cs:9-14)

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=-1, random_state=None):
        dtype = np.dtype(dtype)
        if np.isscalar(low) and np.isscalar(high):
            if shape is None:
                shape = (1,)
            low = np.full(shape, low, dtype)    # Box.cs:28-29
            high = np.full(shape, high, dtype)
        else:
            low = np.asarray(low).astype(dtype)  # Box.cs:40-41
            high = np.asarray(high).astype(dtype)
            assert low.shape == high.shape       # Box.cs:38
        super().__init__(low.shape, dtype)
        self.Low, self.High = low, high
        self.RandomState = (np.random.RandomState(seed) if seed != -1
                            else (random_state if random_state is not None else np.random))
        # CheckBounded (Box.cs:46-51)
        self.BoundedLow = self.Low > -np.inf
        self.BoundedHigh = self.High < np.inf

    def IsBounded(self, manner="Both"):
        below, above = bool(np.all(self.BoundedLow)), bool(np.all(self.BoundedHigh))
        if manner == Box.BOTH:
            return below and above
        if manner == Box.ABOVE:
            return above
        if manner == Box.BELOW:
            return below
        raise ValueError("Unsupported BoundedMannerEnum value.")

    def Sample(self, mask=None):
        if mask is not None:
            raise NotImplementedError("Box.sample cannot be provided a mask.")   # Box.cs:70-73
        bl, bh = self.BoundedLow, self.BoundedHigh
        unbounded, upp, low, both = ~bl & ~bh, ~bl & bh, bl & ~bh, bl & bh
        rs = self.RandomState
        sample = np.empty(self.Shape, np.float64)
        sample[unbounded] = rs.normal(0.5, 1.0, int(unbounded.sum()))                      # Box.cs:81
        sample[low] = rs.exponential(1.0, int(low.sum())) + self.Low[low]                  # Box.cs:82
        sample[upp] = rs.exponential(1.0, int(upp.sum())) + self.High[upp]                 # Box.cs:83
        sample[both] = rs.uniform(self.Low[both], self.High[both])                         # Box.cs:84
        if self.DType.kind in "iu":
            sample = np.floor(sample)
        return sample.astype(self.DType)

    def SampleBatch(self, count, seed=0, index=0, device=0):
        """`count` samples of this box drawn ON THE GPU (gymcuda_box_sample): the same four-way split per component
        (Box.cs:81-84) from the engine's Philox stream -- sample c, component j is a pure function of (seed, index + c, j).
        Returns [count, *Shape] in this box's dtype.  No CPU fallback: raises without a CUDA device."""
        import ctypes as C
        from . import _native as N
        dim = int(np.prod(self.Shape))
        low = np.ascontiguousarray(self.Low, np.float32).reshape(dim)
        high = np.ascontiguousarray(self.High, np.float32).reshape(dim)
        out = np.empty((int(count), dim), np.float32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().gymcuda_box_sample(int(device), int(seed) & (2**64 - 1), int(index), vp(low), vp(high), dim, int(count),
                                           1 if self.DType.kind in "iu" else 0, vp(out)))
        return out.reshape((int(count),) + self.Shape).astype(self.DType)

    def Contains(self, x):
        if not isinstance(x, np.ndarray):
            raise NotImplementedError(repr(x))   # NotSupportedException (Box.cs:95)
        return x.shape == self.Shape and bool(np.all(x >= self.Low)) and bool(np.all(x <= self.High))

    def Seed(self, seed):
        self.RandomState = np.random.RandomState(seed)

    def __eq__(self, other):
        return isinstance(other, Box) and np.array_equal(self.Low, other.Low) and np.array_equal(self.High, other.High)

    def __repr__(self):
        return "Box" + str(self.Shape)


class Discrete(Space):
    """Gym.Spaces.Discrete (Discrete.cs:5-48)."""

    def __init__(self, n, dtype=np.float32, seed=-1, start=0, random_state=None):
        super().__init__((n,), dtype)
        self.N, self.Start = int(n), int(start)
        self.RandomState = (np.random.RandomState(seed) if seed != -1
                            else (random_state if random_state is not None else np.random))

    def Sample(self, mask=None):
        if mask is not None:
            # Discrete.cs:19-25.  The reference casts the ARRAY of valid indices to one int and draws
            # `choice(int)` = randint(0, that int), which can return a masked-out action; implemented here (and in
            # the device sampler, kernels.cuh sample_kernel) is what its comment and upstream gym say: uniform over
            # the entries equal to 1, `Start` when there is none.
            valid = np.nonzero(np.asarray(mask) == 1)[0]
            if valid.size:
                return self.Start + int(self.RandomState.choice(valid))
            return self.Start
        return self.Start + int(self.RandomState.randint(0, self.N))   # Discrete.cs:27

    def Contains(self, x):
        if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
            return 0 <= int(x) < self.N                      # Discrete.cs:38-40
        raise NotImplementedError(repr(x))                   # NotSupportedException (Discrete.cs:35)

    def Seed(self, seed):
        self.RandomState = np.random.RandomState(seed)

    def __repr__(self):
        return "Discrete(%d)" % self.N
