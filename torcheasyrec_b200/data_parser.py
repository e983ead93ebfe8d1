"""DataParser: parsed feature columns -> `Batch`, built straight in its final layout.

The reference's `DataParser.to_batch` (tzrec/datasets/data_parser.py:402-500) receives one tensor per feature column
(`<feature>.values`, `<feature>.lengths`, optional `.weights` / `.key_lengths`; dense `<feature>.values`; labels) and
builds one KeyedJaggedTensor per data group with 2-3 `torch.cat`s over the group's features (`_to_sparse_features`
:526-594), one KeyedTensor per dense group (`_to_dense_features` :502-524); the DataLoader then pins every tensor and
`Batch.to` issues one H2D copy per tensor (tzrec/datasets/utils.py:344-463).

Here the same contract has two entry points:
  * `to_batch(input_data)`            — the reference's semantics, tensor for tensor (the parity baseline);
  * `to_batch_into(input_data, arena)`— every column is written ONCE, at its final offset, into ONE pinned host arena
    laid out exactly like the device arena ([values | lengths | weights | dense | labels] per data group, 16-B aligned
    blocks); `arena.to_device()` is then a SINGLE cudaMemcpyAsync per step and the device `Batch` is a set of views.
    No per-feature cat temporaries, no per-tensor pin, no per-tensor copy; lengths -> offsets happens on the device
    (tzk_lengths_to_offsets) when a collection first asks for it.
The feature-generation / Arrow side of the parser (pyfg, FG_DAG) is out of scope like the rest of the reference's IO.
"""
from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import torch

from .batch import Batch
from .features import BaseFeature
from .sparse import JaggedTensor, KeyedJaggedTensor, KeyedTensor


class DataParser:
    """Key partition as tzrec/datasets/data_parser.py:107-132."""

    def __init__(self, features: Sequence[BaseFeature], labels: Optional[List[str]] = None,
                 sample_weights: Optional[List[str]] = None) -> None:
        self._features = list(features)
        self._labels = list(labels or [])
        self._sample_weights = list(sample_weights or [])
        self.dense_keys: Dict[str, List[str]] = defaultdict(list)
        self.dense_length_per_key: Dict[str, List[int]] = defaultdict(list)
        self.sparse_keys: Dict[str, List[str]] = defaultdict(list)
        self.sequence_mulval_sparse_keys: Dict[str, List[str]] = defaultdict(list)
        self.sequence_dense_keys: List[str] = []
        self.has_weight_keys: Dict[str, List[str]] = defaultdict(list)
        for f in self._features:
            if getattr(f, "stub_type", False):
                continue
            if f.is_sequence:
                if f.is_sparse:
                    self.sparse_keys[f.data_group].append(f.name)
                    if f.value_dim != 1:
                        self.sequence_mulval_sparse_keys[f.data_group].append(f.name)
                else:
                    self.sequence_dense_keys.append(f.name)
            elif f.is_sparse:
                self.sparse_keys[f.data_group].append(f.name)
            else:
                self.dense_keys[f.data_group].append(f.name)
                self.dense_length_per_key[f.data_group].append(f.value_dim)
            if f.is_weighted:
                self.has_weight_keys[f.data_group].append(f.name)

    # ---- reference semantics ----------------------------------------------------------------------------------------
    def _lengths_of(self, input_data, dg: str, key: str):
        """(bag lengths, seq_length | None, key_length | None) — multi-value sequences flatten (data_parser.py:556-566)."""
        length = input_data[f"{key}.lengths"]
        if key in self.sequence_mulval_sparse_keys[dg]:
            seq_length, key_length = length, input_data[f"{key}.key_lengths"]
            length = torch.segment_reduce(key_length.float(), "sum", lengths=seq_length).to(length.dtype)
            return length, seq_length, key_length
        return length, None, None

    def to_batch(self, input_data: Dict[str, torch.Tensor]) -> Batch:
        batch = Batch()
        for dg, keys in self.dense_keys.items():
            batch.dense_features[dg] = KeyedTensor(keys, self.dense_length_per_key[dg],
                                                   torch.cat([input_data[f"{k}.values"] for k in keys], dim=-1))
        for dg, keys in self.sparse_keys.items():
            values, lengths, weights = [], [], []
            mv_keys, mv_seq, mv_key = [], [], []
            has_w = self.has_weight_keys[dg]
            for key in keys:
                values.append(input_data[f"{key}.values"])
                length, seq_l, key_l = self._lengths_of(input_data, dg, key)
                if seq_l is not None:
                    mv_keys.append(key)
                    mv_seq.append(seq_l)
                    mv_key.append(key_l)
                lengths.append(length)
                if has_w:
                    weights.append(input_data[f"{key}.weights"] if key in has_w
                                   else torch.ones_like(input_data[f"{key}.values"], dtype=torch.float32))
            kjt = KeyedJaggedTensor(keys, torch.cat(values, dim=-1), lengths=torch.cat(lengths, dim=-1).to(torch.int32),
                                    weights=torch.cat(weights, dim=-1) if has_w else None, stride=lengths[0].size(0))
            kjt._length_per_key = [int(v.numel()) for v in values]
            batch.sparse_features[dg] = kjt
            if mv_keys:
                batch.sequence_mulval_lengths[dg] = KeyedJaggedTensor(
                    mv_keys, torch.cat(mv_key, dim=-1), lengths=torch.cat(mv_seq, dim=-1).to(torch.int32))
        for key in self.sequence_dense_keys:
            batch.sequence_dense_features[key] = JaggedTensor(input_data[f"{key}.values"],
                                                              lengths=input_data[f"{key}.lengths"])
        for name in self._labels:
            batch.labels[name] = input_data[name]
        for name in self._sample_weights:
            batch.sample_weights[name] = input_data[name]
        return batch

    # ---- final-layout build -------------------------------------------------------------------------------------------
    def make_arena(self, batch_size: int, max_ids: Dict[str, int], device=None) -> "BatchArena":
        """`max_ids[data_group]` = most ids a batch of the group carries (capacity of its values block)."""
        return BatchArena(self, batch_size, max_ids, device)

    def to_batch_into(self, input_data: Dict[str, torch.Tensor], arena: "BatchArena") -> Batch:
        """Writes every column at its final place in the pinned arena; returns the HOST batch (views of the arena).
        `arena.to_device()` afterwards gives the device batch with one copy."""
        B = arena.B
        host = Batch()
        for dg, keys in self.sparse_keys.items():
            blk = arena.blocks[dg]
            has_w = bool(self.has_weight_keys[dg])
            o, lpk = 0, []
            for f, key in enumerate(keys):
                v = input_data[f"{key}.values"]
                n = int(v.numel())
                if o + n > blk["cap"]:
                    raise RuntimeError(f"data group {dg}: {o + n} ids exceed the arena capacity {blk['cap']}")
                blk["h_values"][o:o + n].copy_(v)
                length, _, _ = self._lengths_of(input_data, dg, key)
                blk["h_lengths"][f * B:(f + 1) * B].copy_(length)
                if has_w:
                    if key in self.has_weight_keys[dg]:
                        blk["h_weights"][o:o + n].copy_(input_data[f"{key}.weights"])
                    else:
                        blk["h_weights"][o:o + n].fill_(1.0)
                o += n
                lpk.append(n)
            blk["nnz"] = o
            kjt = KeyedJaggedTensor(keys, blk["h_values"][:o], lengths=blk["h_lengths"],
                                    weights=blk["h_weights"][:o] if has_w else None, stride=B)
            kjt._length_per_key = lpk
            host.sparse_features[dg] = kjt
            mv = self.sequence_mulval_sparse_keys[dg]
            if mv:    # rare (not in the BASELINE configs): kept as separate small tensors
                host.sequence_mulval_lengths[dg] = KeyedJaggedTensor(
                    mv, torch.cat([input_data[f"{k}.key_lengths"] for k in mv]),
                    lengths=torch.cat([input_data[f"{k}.lengths"] for k in mv]).to(torch.int32))
        for dg, keys in self.dense_keys.items():
            blk = arena.blocks[dg]
            c = 0
            for key, d in zip(keys, self.dense_length_per_key[dg]):
                blk["h_dense"][:, c:c + d].copy_(input_data[f"{key}.values"].reshape(B, d))
                c += d
            host.dense_features[dg] = KeyedTensor(keys, self.dense_length_per_key[dg], blk["h_dense"])
        for name in self._labels:
            arena.h_labels[name].copy_(input_data[name])
            host.labels[name] = arena.h_labels[name]
        for name in self._sample_weights:
            arena.h_labels[name].copy_(input_data[name])
            host.sample_weights[name] = arena.h_labels[name]
        return host


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class BatchArena:
    """One pinned host buffer + (optionally) its device twin, laid out per data group as
    [values int64 x cap | lengths int32 x F*B | weights f32 x cap | dense f32 x B*sum(dim)] then the labels."""

    def __init__(self, parser: DataParser, batch_size: int, max_ids: Dict[str, int], device=None) -> None:
        self.parser, self.B = parser, int(batch_size)
        B = self.B
        plan, o = {}, 0

        def take(nbytes):
            nonlocal o
            s = o
            o = _align(o + nbytes)
            return s

        groups = list(dict.fromkeys(list(parser.sparse_keys) + list(parser.dense_keys)))
        for dg in groups:
            d = {}
            F = len(parser.sparse_keys.get(dg, []))
            if F:
                cap = int(max_ids[dg])
                d.update(cap=cap, F=F, values=take(cap * 8), lengths=take(F * B * 4))
                if parser.has_weight_keys.get(dg):
                    d["weights"] = take(cap * 4)
            if parser.dense_keys.get(dg):
                d["dense_dim"] = sum(parser.dense_length_per_key[dg])
                d["dense"] = take(B * d["dense_dim"] * 4)
            plan[dg] = d
        lab = {name: take(B * 4) for name in parser._labels + parser._sample_weights}
        self.nbytes = o
        pin = torch.cuda.is_available()
        self.host = torch.empty(max(o, 16), dtype=torch.uint8, pin_memory=pin)
        self.device = torch.device(device) if device is not None else None
        self.dev = torch.empty(max(o, 16), dtype=torch.uint8, device=self.device) if self.device is not None else None
        self._plan, self._lab = plan, lab
        self.blocks = {dg: self._views(self.host, d, "h_") for dg, d in plan.items()}
        for dg in self.blocks:
            self.blocks[dg].update(cap=plan[dg].get("cap", 0), nnz=0)
        self.h_labels = {n: self.host[s:s + B * 4].view(torch.float32) for n, s in lab.items()}

    def _views(self, buf: torch.Tensor, d: dict, pre: str) -> dict:
        B, out = self.B, {}
        if "values" in d:
            out[pre + "values"] = buf[d["values"]:d["values"] + d["cap"] * 8].view(torch.int64)
            out[pre + "lengths"] = buf[d["lengths"]:d["lengths"] + d["F"] * B * 4].view(torch.int32)
            if "weights" in d:
                out[pre + "weights"] = buf[d["weights"]:d["weights"] + d["cap"] * 4].view(torch.float32)
        if "dense" in d:
            out[pre + "dense"] = buf[d["dense"]:d["dense"] + B * d["dense_dim"] * 4].view(torch.float32).view(B, d["dense_dim"])
        return out

    def to_device(self, host_batch: Batch, non_blocking: bool = True) -> Batch:
        """ONE host->device copy of the whole arena (the bytes in use are a prefix-free layout, so the full buffer
        travels: padding included it is what the per-tensor copies moved, minus their launch overhead)."""
        assert self.dev is not None, "arena built without a device"
        self.dev.copy_(self.host, non_blocking=non_blocking)
        B, p = self.B, self.parser
        out = Batch()
        for dg, hk in host_batch.sparse_features.items():
            v = self._views(self.dev, self._plan[dg], "d_")
            nnz = self.blocks[dg]["nnz"]
            kjt = KeyedJaggedTensor(hk.keys(), v["d_values"][:nnz], lengths=v["d_lengths"],
                                    weights=v["d_weights"][:nnz] if "d_weights" in v else None, stride=B)
            kjt._length_per_key = hk._length_per_key
            out.sparse_features[dg] = kjt
        for dg, hk in host_batch.dense_features.items():
            v = self._views(self.dev, self._plan[dg], "d_")
            out.dense_features[dg] = KeyedTensor(hk.keys(), hk.length_per_key(), v["d_dense"])
        for name, s in self._lab.items():
            t = self.dev[s:s + B * 4].view(torch.float32)
            (out.labels if name in p._labels else out.sample_weights)[name] = t
        for dg, kj in host_batch.sequence_mulval_lengths.items():
            out.sequence_mulval_lengths[dg] = kj.to(self.device, non_blocking=non_blocking)
        return out
