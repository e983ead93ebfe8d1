"""KeyedJaggedTensor / JaggedTensor / KeyedTensor value classes.

The reference takes these from torchrec ([EXT] torchrec.sparse.jagged_tensor, imported at
tzrec/datasets/utils.py:21 and tzrec/modules/embedding.py) and only ever uses the subset below
(SURVEY.md §7.1 step 0).  Layout contract (App. A.1): `values` are key-major; `lengths[f*B + b]`;
`offsets = [0, cumsum(lengths)]`; `stride = B`.

Device-side derived data (offsets) is produced by the tzk scan kernel when the tensors live on a CUDA device;
host-side construction (`from_lengths_sync` on CPU tensors) uses torch.cumsum because the reference builds
its batches on the host too (tzrec/datasets/data_parser.py:526-594).
"""

from typing import Dict, List, Optional, Sequence

import torch


def _offsets_from_lengths(lengths: torch.Tensor) -> torch.Tensor:
    if lengths.is_cuda:
        from .kernels import default_kernels

        return default_kernels().lengths_to_offsets(lengths.to(torch.int32).contiguous())
    out = torch.zeros(lengths.numel() + 1, dtype=torch.int64)
    torch.cumsum(lengths.to(torch.int64), 0, out=out[1:])
    return out


class JaggedTensor:
    """values [sum(len), ...] + lengths [B] (+ lazily offsets [B+1])."""

    def __init__(self, values: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 offsets: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None) -> None:
        assert lengths is not None or offsets is not None
        self._values, self._lengths, self._offsets, self._weights = values, lengths, offsets, weights

    def values(self) -> torch.Tensor:
        return self._values

    def weights_or_none(self) -> Optional[torch.Tensor]:
        return self._weights

    def lengths(self) -> torch.Tensor:
        if self._lengths is None:
            self._lengths = (self._offsets[1:] - self._offsets[:-1]).to(torch.int32)
        return self._lengths

    def offsets(self) -> torch.Tensor:
        if self._offsets is None:
            self._offsets = _offsets_from_lengths(self._lengths)
        return self._offsets

    def to(self, device, non_blocking: bool = False) -> "JaggedTensor":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return JaggedTensor(mv(self._values), mv(self._lengths), mv(self._offsets), mv(self._weights))

    def to_padded_dense(self, desired_length: Optional[int] = None, padding_value: float = 0.0) -> torch.Tensor:
        """[B, T, D]; row b holds its first min(len_b, T) rows (App. A.14; embedding.py:1429,1480)."""
        from .functional import jagged_to_padded_dense

        if desired_length is None:
            desired_length = int(self.lengths().max().item()) if self.lengths().numel() else 0
        out = jagged_to_padded_dense(self._values, self.offsets(), desired_length)
        if padding_value != 0.0:       # (the reference only pads with zeros; kept for API parity with torchrec)
            lens = self.lengths().to(torch.int64).clamp(max=desired_length)
            pad = torch.arange(desired_length, device=out.device)[None, :] >= lens[:, None]
            out = out.masked_fill(pad.view(pad.shape + (1,) * (out.dim() - 2)), padding_value)
        return out


class KeyedJaggedTensor:
    """Key-major jagged ids for F keys x B samples (tzrec/datasets/utils.py:299-342, row A0 of SURVEY §8a)."""

    def __init__(self, keys: Sequence[str], values: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 offsets: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None,
                 stride: Optional[int] = None) -> None:
        assert lengths is not None or offsets is not None
        self._keys = list(keys)
        self._values, self._lengths, self._offsets, self._weights = values, lengths, offsets, weights
        n_bags = lengths.numel() if lengths is not None else offsets.numel() - 1
        self._stride = stride if stride is not None else (n_bags // len(self._keys) if self._keys else 0)
        self._length_per_key: Optional[List[int]] = None

    # ---- constructors -------------------------------------------------------------------------------
    @staticmethod
    def from_lengths_sync(keys: Sequence[str], values: torch.Tensor, lengths: torch.Tensor,
                          weights: Optional[torch.Tensor] = None) -> "KeyedJaggedTensor":
        return KeyedJaggedTensor(keys, values, lengths=lengths.to(torch.int32), weights=weights)

    @staticmethod
    def from_offsets_sync(keys: Sequence[str], values: torch.Tensor, offsets: torch.Tensor,
                          weights: Optional[torch.Tensor] = None) -> "KeyedJaggedTensor":
        return KeyedJaggedTensor(keys, values, offsets=offsets.to(torch.int64), weights=weights)

    @staticmethod
    def empty(device=None) -> "KeyedJaggedTensor":
        return KeyedJaggedTensor([], torch.zeros(0, dtype=torch.int64, device=device),
                                 lengths=torch.zeros(0, dtype=torch.int32, device=device), stride=0)

    # ---- accessors ----------------------------------------------------------------------------------
    def keys(self) -> List[str]:
        return self._keys

    def values(self) -> torch.Tensor:
        return self._values

    def weights_or_none(self) -> Optional[torch.Tensor]:
        return self._weights

    def stride(self) -> int:
        return self._stride

    def lengths(self) -> torch.Tensor:
        if self._lengths is None:
            self._lengths = (self._offsets[1:] - self._offsets[:-1]).to(torch.int32)
        return self._lengths

    def offsets(self) -> torch.Tensor:
        if self._offsets is None:
            self._offsets = _offsets_from_lengths(self._lengths)
        return self._offsets

    def length_per_key(self) -> List[int]:
        """Host list (forces a sync, like torchrec's)."""
        if self._length_per_key is None:
            F, B = len(self._keys), self._stride
            self._length_per_key = self.lengths().view(F, B).sum(dim=1).tolist() if F and B else [0] * F
        return self._length_per_key

    def device(self) -> torch.device:
        return self._values.device

    # ---- transforms ---------------------------------------------------------------------------------
    def to(self, device, non_blocking: bool = False) -> "KeyedJaggedTensor":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        out = KeyedJaggedTensor(self._keys, mv(self._values), mv(self._lengths), mv(self._offsets),
                                mv(self._weights), self._stride)
        out._length_per_key = self._length_per_key
        return out

    def pin_memory(self) -> "KeyedJaggedTensor":
        pm = lambda t: None if t is None else t.pin_memory()
        out = KeyedJaggedTensor(self._keys, pm(self._values), pm(self._lengths), pm(self._offsets),
                                pm(self._weights), self._stride)
        out._length_per_key = self._length_per_key
        return out

    def record_stream(self, stream) -> None:
        for t in (self._values, self._lengths, self._offsets, self._weights):
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    def to_dict(self) -> Dict[str, JaggedTensor]:
        off = self.offsets()
        B = self._stride
        bounds = off[:: B].tolist() if B else [0] * (len(self._keys) + 1)  # host sync, as in torchrec
        out = {}
        for f, k in enumerate(self._keys):
            s, e = bounds[f], bounds[f + 1]
            out[k] = JaggedTensor(self._values[s:e], lengths=self.lengths()[f * B:(f + 1) * B],
                                  weights=None if self._weights is None else self._weights[s:e])
        return out

    def permute(self, indices: Sequence[int]) -> "KeyedJaggedTensor":
        """Key permutation / selection (keys may repeat).  [EXT] fbgemm::permute_2D_sparse_data (K2)."""
        from .functional import kjt_permute

        return kjt_permute(self, list(indices))

    def __repr__(self) -> str:
        return f"KeyedJaggedTensor(keys={self._keys}, stride={self._stride}, nnz={self._values.numel()})"


class KeyedTensor:
    """Dense [B, sum(length_per_key)] with named column blocks ([EXT] torchrec KeyedTensor; App. A.2)."""

    def __init__(self, keys: Sequence[str], length_per_key: Sequence[int], values: torch.Tensor, key_dim: int = 1):
        assert key_dim == 1
        self._keys, self._length_per_key, self._values = list(keys), list(length_per_key), values

    def keys(self) -> List[str]:
        return self._keys

    def length_per_key(self) -> List[int]:
        return self._length_per_key

    def values(self) -> torch.Tensor:
        return self._values

    def offset_per_key(self) -> List[int]:
        out = [0]
        for n in self._length_per_key:
            out.append(out[-1] + n)
        return out

    def to(self, device, non_blocking: bool = False) -> "KeyedTensor":
        return KeyedTensor(self._keys, self._length_per_key, self._values.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "KeyedTensor":
        return KeyedTensor(self._keys, self._length_per_key, self._values.pin_memory())

    def record_stream(self, stream) -> None:
        if self._values.is_cuda:
            self._values.record_stream(stream)

    def to_dict(self) -> Dict[str, torch.Tensor]:
        off = self.offset_per_key()
        return {k: self._values[:, off[i]:off[i + 1]] for i, k in enumerate(self._keys)}

    @staticmethod
    def from_tensor_list(keys: Sequence[str], tensors: List[torch.Tensor], key_dim: int = 1,
                         cat_dim: int = 1) -> "KeyedTensor":
        """[EXT] torchrec KeyedTensor.from_tensor_list (used by tzrec/datasets/utils.py and every embedding test):
        concatenates `[B, d_i]` blocks along the key dimension."""
        assert key_dim == 1 and cat_dim == 1 and len(keys) == len(tensors)
        return KeyedTensor(list(keys), [int(t.shape[1]) for t in tensors], torch.cat(list(tensors), dim=1))

    @staticmethod
    def regroup_as_dict(keyed_tensors: List["KeyedTensor"], groups: List[List[str]], keys: List[str]
                        ) -> Dict[str, torch.Tensor]:
        """{group name: [B, sum D]} — call site tzrec/modules/embedding.py:972-976 (K6, App. A.13)."""
        from .functional import regroup

        return dict(zip(keys, regroup(keyed_tensors, groups)))

    @staticmethod
    def regroup(keyed_tensors: List["KeyedTensor"], groups: List[List[str]]) -> List[torch.Tensor]:
        from .functional import regroup

        return regroup(keyed_tensors, groups)
