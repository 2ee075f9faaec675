"""Minimal host containers that define the INPUT LAYOUT of the hot path (reference src/Data/datatypes.jl, configs.jl).

Only what the path reads is mirrored: nested per-atom descriptor vectors, energies, forces, Configuration, DataSet,
and the accessors the hot-path callers use. Everything else of src/Data (units, AtomsBase systems, threaded ACE
descriptor evaluation) is out of scope and stays Julia.
"""
from __future__ import annotations

import numpy as np


class Energy:
    """Energy (datatypes.jl:38)."""

    def __init__(self, d: float, u=None):
        self.d = float(d)
        self.u = u


class LocalDescriptors:
    """LocalDescriptors (datatypes.jl:72): natoms vectors of length K. Stored as one natoms x K array (row = atom)."""

    def __init__(self, b):
        self.b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
        if self.b.ndim != 2:
            raise ValueError("LocalDescriptors needs natoms vectors of equal length")

    def __len__(self):
        return self.b.shape[0]


class Forces:
    """Forces (datatypes.jl:119): natoms force 3-vectors."""

    def __init__(self, f):
        self.f = np.ascontiguousarray(np.asarray(f, dtype=np.float64))
        if self.f.ndim != 2:
            raise ValueError("Forces needs natoms 3-vectors")

    def __len__(self):
        return self.f.shape[0]


class ForceDescriptors:
    """ForceDescriptors (datatypes.jl:153): natoms x 3 vectors of length K, stored natoms x 3 x K.
    reduce(vcat, get_values(fd)) (linear-learning-problem.jl:81) is then b.reshape(3*natoms, K): atom-major, xyz inner."""

    def __init__(self, b):
        self.b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
        if self.b.ndim != 3 or self.b.shape[1] != 3:
            raise ValueError("ForceDescriptors needs natoms x 3 vectors of equal length")

    def __len__(self):
        return self.b.shape[0]


def get_values(x):
    """get_values (datatypes.jl:63,92,127,138,173)."""
    if isinstance(x, Energy):
        return x.d
    if isinstance(x, LocalDescriptors):
        return x.b
    if isinstance(x, Forces):
        return x.f
    if isinstance(x, ForceDescriptors):
        return x.b
    raise TypeError(type(x))


class Configuration:
    """Configuration = Dict{DataType, ...} (configs.jl:2-4). `+` merges two configurations (configs.jl:24-26)."""

    def __init__(self, *data):
        self.data = {}
        for d in data:
            if d is not None:
                self.data[type(d)] = d

    def __add__(self, other: "Configuration") -> "Configuration":
        c = Configuration()
        c.data = {**self.data, **other.data}
        return c


class DataSet:
    """DataSet = Vector{Configuration} (configs.jl:96-107)."""

    def __init__(self, configurations):
        self.Configurations = list(configurations)

    def __len__(self):
        return len(self.Configurations)

    def __iter__(self):
        return iter(self.Configurations)

    def __getitem__(self, i):
        if isinstance(i, (slice, list, np.ndarray)):
            idx = range(len(self))[i] if isinstance(i, slice) else i
            return DataSet([self.Configurations[int(j)] for j in idx])
        return self.Configurations[i]


def _get(c: Configuration, T):
    try:
        return c.data[T]
    except KeyError:
        raise KeyError(f"Configuration has no {T.__name__}") from None


def get_energy(c):
    return _get(c, Energy)


def get_local_descriptors(c):
    return _get(c, LocalDescriptors)


def get_forces(c):
    return _get(c, Forces)


def get_force_descriptors(c):
    return _get(c, ForceDescriptors)
