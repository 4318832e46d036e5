"""Symbolic tensor descriptor: a shape + dtype record with just enough algebra (view / transpose / expand / squeeze)
to describe operator signatures without allocating anything.  The IR stores plain tuples internally; this class is the
user-facing way to build and inspect shapes (graph capture, ``analysis_op_info`` consumers, custom emitters).

Parity target: /root/reference/simumax/core/tensor.py (``TensorSize`` a.k.a. ``FakeTensor``, ``Float8Tensor``, ``BPE``).
"""
from __future__ import annotations

import itertools
import math
from typing import Iterable, List, Optional, Sequence, Tuple, Union

BPE = {"bf16": 2, "fp16": 2, "fp32": 4, "fp8": 1, "int32": 4, "int64": 8}
_ids = itertools.count()


def _norm_dim(dim: int, ndim: int) -> int:
    if not -ndim <= dim < ndim:
        raise IndexError(f"dim {dim} out of range for a {ndim}-d shape")
    return dim % ndim


class TensorSize:
    """Immutable-ish shape record. Every transforming method returns a new descriptor."""

    __slots__ = ("shape", "dtype", "id", "parents")

    def __init__(self, shape: Sequence[int], dtype: str = "bf16", parents: Iterable["TensorSize"] = ()):
        if dtype not in BPE:
            raise ValueError(f"unknown dtype {dtype!r}; expected one of {sorted(BPE)}")
        self.shape: List[int] = [int(s) for s in shape]
        self.dtype = dtype
        self.id = next(_ids)
        self.parents: Tuple[TensorSize, ...] = tuple(parents)

    # ---- inspection
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def tensors(self) -> List["TensorSize"]:
        return [self]

    @property
    def element_size(self) -> int:
        return BPE[self.dtype]

    def size(self, index: Optional[int] = None) -> Union[int, List[int]]:
        return list(self.shape) if index is None else self.shape[_norm_dim(index, self.ndim)]

    def numel(self) -> int:
        return math.prod(self.shape) if self.shape else 0

    def get_memory_size(self) -> int:
        return self.numel() * self.element_size

    mem_size = get_memory_size

    # ---- algebra
    def _child(self, shape, dtype=None) -> "TensorSize":
        return type(self)(shape, dtype or self.dtype, parents=(self,))

    def new(self, dtype: Optional[str] = None) -> "TensorSize":
        return self._child(self.shape, dtype)

    def new_with_dim(self, dim: int, value: int) -> "TensorSize":
        s = list(self.shape)
        s[_norm_dim(dim, self.ndim)] = int(value)
        return self._child(s)

    def view(self, *shape: int) -> "TensorSize":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if shape.count(-1) > 1:
            raise ValueError("only one dimension may be inferred")
        known = math.prod(s for s in shape if s != -1)
        if -1 in shape:
            if known == 0 or self.numel() % known:
                raise ValueError(f"cannot view {self.shape} as {shape}")
            shape[shape.index(-1)] = self.numel() // known
        elif math.prod(shape) != self.numel():
            raise ValueError(f"cannot view {self.shape} ({self.numel()} elements) as {shape}")
        return self._child(shape)

    reshape = view

    def transpose(self, d0: int, d1: int) -> "TensorSize":
        s = list(self.shape)
        a, b = _norm_dim(d0, self.ndim), _norm_dim(d1, self.ndim)
        s[a], s[b] = s[b], s[a]
        return self._child(s)

    def permute(self, *dims: int) -> "TensorSize":
        if sorted(_norm_dim(d, self.ndim) for d in dims) != list(range(self.ndim)):
            raise ValueError(f"bad permutation {dims} for {self.ndim} dims")
        return self._child([self.shape[_norm_dim(d, self.ndim)] for d in dims])

    def expand(self, *shape: int) -> "TensorSize":
        if len(shape) < self.ndim:
            raise ValueError("expand cannot drop dimensions")
        lead = len(shape) - self.ndim
        out = []
        for i, s in enumerate(shape):
            cur = 1 if i < lead else self.shape[i - lead]
            if s == -1:
                s = cur
            if cur not in (1, s):
                raise ValueError(f"cannot expand {self.shape} to {list(shape)}")
            out.append(s)
        return self._child(out)

    def unsqueeze(self, dim: int) -> "TensorSize":
        d = dim if dim >= 0 else dim + self.ndim + 1
        s = list(self.shape)
        s.insert(d, 1)
        return self._child(s)

    unsqeeze = unsqueeze      # the reference's spelling, kept so ported scripts run

    def squeeze(self, dim: Optional[int] = None) -> "TensorSize":
        if dim is None:
            return self._child([s for s in self.shape if s != 1])
        d = _norm_dim(dim, self.ndim)
        s = list(self.shape)
        if s[d] == 1:
            s.pop(d)
        return self._child(s)

    def split(self, sizes: Union[int, Sequence[int]], dim: int = -1) -> List["TensorSize"]:
        d = _norm_dim(dim, self.ndim)
        if isinstance(sizes, int):
            full, rem = divmod(self.shape[d], sizes)
            sizes = [sizes] * full + ([rem] if rem else [])
        if sum(sizes) != self.shape[d]:
            raise ValueError(f"split sizes {list(sizes)} do not add up to {self.shape[d]}")
        return [self.new_with_dim(d, s) for s in sizes]

    @staticmethod
    def cat(tensors: Sequence["TensorSize"], dim: int = -1) -> "TensorSize":
        first = tensors[0]
        d = _norm_dim(dim, first.ndim)
        for t in tensors[1:]:
            if t.ndim != first.ndim or any(a != b for i, (a, b) in enumerate(zip(t.shape, first.shape)) if i != d):
                raise ValueError("cat: shapes differ outside the concatenation dim")
        s = list(first.shape)
        s[d] = sum(t.shape[d] for t in tensors)
        return type(first)(s, first.dtype, parents=tuple(tensors))

    # ---- dunder
    def __eq__(self, other):
        return isinstance(other, TensorSize) and self.shape == other.shape and self.dtype == other.dtype

    def __hash__(self):
        return hash((tuple(self.shape), self.dtype))

    def __iter__(self):
        return iter(self.shape)

    def __len__(self):
        return self.ndim

    def __getitem__(self, i):
        return self.shape[i]

    def __repr__(self):
        return f"{type(self).__name__}(shape={self.shape}, dtype={self.dtype})"


FakeTensor = TensorSize


class Float8Tensor(TensorSize):
    """fp8 payload + one fp32 scale (per-tensor scaling, TE ``Float8Quantizer`` style): 1 byte per element."""

    def __init__(self, shape: Sequence[int], dtype: str = "fp8", parents: Iterable[TensorSize] = ()):
        super().__init__(shape, "fp8", parents)

    def get_memory_size(self) -> int:
        return self.numel()

    mem_size = get_memory_size
