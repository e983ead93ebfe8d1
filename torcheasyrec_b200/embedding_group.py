"""EmbeddingGroup: feature groups -> {EmbeddingBagCollection, EmbeddingCollection} -> grouped tensors.

Host-side mirror of tzrec/modules/embedding.py (`EmbeddingGroup` :167-547, `EmbeddingGroupImpl` :681-978,
`SequenceEmbeddingGroupImpl` :993-1498) with the same constructor, query methods and output dict, built on
this repo's collections (embedding_modules.py) instead of torchrec's.  Table-set rules restated from the
reference:
  * table name = `embedding_name` or `{feature}_emb`; WIDE groups append `_wide`; a group suffix appends `_{suffix}`
  * WIDE embedding dim = `wide_embedding_dim or 4`
  * same-name configs must agree (rows, dim, pooling, init) and are merged, feature lists unioned
  * a feature mapped to more than one distinct table is emitted under the key `feature@table`
Out of scope here (raise): input-tile modes, managed-collision (zch) tables, dense "autodis/mlp" embeddings,
multi-value ids inside sequences (K8), which none of the five BASELINE configs use.
"""

from collections import OrderedDict, defaultdict
from typing import Dict, List, NamedTuple, Optional

import torch
from torch import nn

from .config import Message
from .embedding_modules import EmbeddingBagCollection, EmbeddingBagConfig, EmbeddingCollection, EmbeddingConfig
from .features import BaseFeature, ParameterConstraints, create_init_fn
from .sparse import JaggedTensor, KeyedJaggedTensor, KeyedTensor

SEQ_TYPES = ("SEQUENCE", "JAGGED_SEQUENCE")


def _check_emb_name_suffix_collisions(feat_to_group_to_emb_name, group_name_to_suffix) -> None:
    """embedding.py:546-573."""
    emb_name_to_suffixes = defaultdict(set)
    for group_to_emb_name in feat_to_group_to_emb_name.values():
        for grp_name, emb_name in group_to_emb_name.items():
            emb_name_to_suffixes[emb_name].add(group_name_to_suffix.get(grp_name, ""))
    for emb_name, suffixes in emb_name_to_suffixes.items():
        if len(suffixes) > 1:
            raise ValueError(
                f"embedding table name {emb_name!r} is produced by groups with different embedding_name_suffix "
                f"values {sorted(suffixes)}; this would silently merge tables that were meant to be independent.")


def _init_repr(fn) -> str:
    return getattr(fn, "__repr_str__", repr(fn))


def _add_embedding_config(configs: Dict[str, object], cfg, check_pooling: bool) -> None:
    """embedding.py:576-627 (_add_embedding_bag_config / _add_embedding_config)."""
    if cfg.name in configs:
        old = configs[cfg.name]
        same = (cfg.num_embeddings == old.num_embeddings and cfg.embedding_dim == old.embedding_dim
                and _init_repr(cfg.init_fn) == _init_repr(old.init_fn)
                and (not check_pooling or cfg.pooling == old.pooling))
        assert same, f"there is a mismatch between {cfg} and {old}, can not share embedding."
        for f in cfg.feature_names:
            if f not in old.feature_names:
                old.feature_names.append(f)
    else:
        configs[cfg.name] = cfg


class EmbeddingGroupImpl(nn.Module):
    """Non-sequence groups of one data group (embedding.py:681-978)."""

    def __init__(self, features: List[BaseFeature], feature_groups: List[Message],
                 wide_embedding_dim: Optional[int] = None, wide_init_fn: Optional[str] = None,
                 device: Optional[torch.device] = None) -> None:
        super().__init__()
        name_to_feature = {x.name: x for x in features}
        emb_bag_configs: Dict[str, EmbeddingBagConfig] = OrderedDict()
        self._emb_bag_constraints: Dict[str, ParameterConstraints] = OrderedDict()
        self.has_sparse = self.has_dense = False
        self.has_mc_sparse = self.has_sparse_user = self.has_mc_sparse_user = self.has_dense_embedding = False
        self._group_to_feature_names = OrderedDict()
        self._group_to_shared_feature_names = OrderedDict()
        self._group_total_dim: Dict[str, int] = {}
        self._group_feature_output_dims: Dict[str, Dict[str, int]] = {}

        feat_to_group_to_emb_name = defaultdict(dict)
        group_name_to_suffix = {fg.group_name: fg.embedding_name_suffix for fg in feature_groups}
        for fg in feature_groups:
            for fname in fg.feature_names:
                feature = name_to_feature[fname]
                if feature.is_sparse:
                    emb_name = feature.emb_bag_config.name
                    if fg.group_type == "WIDE":
                        emb_name += "_wide"
                    if fg.embedding_name_suffix:
                        emb_name += "_" + fg.embedding_name_suffix
                    feat_to_group_to_emb_name[fname][fg.group_name] = emb_name
        _check_emb_name_suffix_collisions(feat_to_group_to_emb_name, group_name_to_suffix)
        shared_flag = {f: len(set(g2e.values())) > 1 for f, g2e in feat_to_group_to_emb_name.items()}

        for fg in feature_groups:
            total_dim, out_dims, shared_names = 0, OrderedDict(), []
            is_wide = fg.group_type == "WIDE"
            for name in fg.feature_names:
                feature = name_to_feature[name]
                shared_name = name
                if feature.is_sparse:
                    output_dim = feature.output_dim
                    cfg = feature.emb_bag_config
                    feature.mc_module(device)
                    if is_wide:
                        cfg.embedding_dim = output_dim = wide_embedding_dim or 4
                        if wide_init_fn:
                            cfg.init_fn = create_init_fn(wide_init_fn)
                    cfg.name = feat_to_group_to_emb_name[name][fg.group_name]
                    const = feature.parameter_constraints(cfg)
                    _add_embedding_config(emb_bag_configs, cfg, check_pooling=True)
                    self.has_sparse = True
                    if const is not None:
                        self._emb_bag_constraints[cfg.name] = const
                    if shared_flag[name]:
                        shared_name = shared_name + "@" + cfg.name
                else:
                    if is_wide:
                        raise ValueError(f"dense feature [{name}] should not be configured in wide group.")
                    output_dim = feature.output_dim
                    self.has_dense = True
                total_dim += output_dim
                out_dims[name] = output_dim
                shared_names.append(shared_name)
            self._group_to_feature_names[fg.group_name] = list(fg.feature_names)
            if shared_names:
                self._group_to_shared_feature_names[fg.group_name] = shared_names
            self._group_total_dim[fg.group_name] = total_dim
            self._group_feature_output_dims[fg.group_name] = out_dims
        self.ebc = EmbeddingBagCollection(list(emb_bag_configs.values()), device=device)

    def group_dims(self, group_name: str) -> List[int]:
        return list(self._group_feature_output_dims[group_name].values())

    def group_feature_dims(self, group_name: str) -> Dict[str, int]:
        return self._group_feature_output_dims[group_name]

    def group_total_dim(self, group_name: str) -> int:
        return self._group_total_dim[group_name]

    def parameter_constraints(self, prefix: str = "") -> Dict[str, ParameterConstraints]:
        return {f"{prefix}ebc.{k}": v for k, v in self._emb_bag_constraints.items()}

    def forward(self, sparse_feature: Optional[KeyedJaggedTensor], dense_feature: Optional[KeyedTensor],
                sparse_feature_user=None, tile_size: int = -1) -> Dict[str, torch.Tensor]:
        kts: List[KeyedTensor] = []
        if self.has_sparse:
            kts.append(self.ebc(sparse_feature))     # embedding.py:930  <- the drop-in boundary
        if self.has_dense:
            kts.append(dense_feature)
        return KeyedTensor.regroup_as_dict(kts, list(self._group_to_shared_feature_names.values()),
                                           list(self._group_to_shared_feature_names.keys()))  # :972-976


class _SeqInfo(NamedTuple):
    name: str
    raw_name: str
    is_sparse: bool
    pooling: str
    value_dim: int
    is_sequence: bool


class SequenceEmbeddingGroupImpl(nn.Module):
    """SEQUENCE / JAGGED_SEQUENCE groups and nested sequence_groups (embedding.py:993-1498)."""

    def __init__(self, features: List[BaseFeature], feature_groups: List[Message],
                 device: Optional[torch.device] = None) -> None:
        super().__init__()
        name_to_feature = {x.name: x for x in features}
        dim_to_emb_configs = defaultdict(OrderedDict)
        self._dim_to_emb_constraints = defaultdict(OrderedDict)
        self.has_sparse = self.has_dense = self.has_sequence_dense = self.has_mulval_seq = False
        self.has_mc_sparse = self.has_sparse_user = self.has_mc_sparse_user = self.has_mulval_seq_user = False
        self._group_to_shared_query = OrderedDict()
        self._group_to_shared_sequence = OrderedDict()
        self._group_total_dim: Dict[str, int] = {}
        self._group_output_dims: Dict[str, List[int]] = {}
        self._group_to_is_jagged: Dict[str, bool] = {}
        self._group_to_sequence_length = OrderedDict()

        feat_to_group_to_emb_name = defaultdict(dict)
        group_name_to_suffix = {fg.group_name: fg.embedding_name_suffix for fg in feature_groups}
        for fg in feature_groups:
            self._group_to_is_jagged[fg.group_name] = (fg._spec("group_type") is not None
                                                       and fg.group_type == "JAGGED_SEQUENCE")
            for fname in fg.feature_names:
                feature = name_to_feature[fname]
                if feature.is_sparse:
                    emb_name = feature.emb_config.name
                    if fg.embedding_name_suffix:
                        emb_name += "_" + fg.embedding_name_suffix
                    feat_to_group_to_emb_name[fname][fg.group_name] = emb_name
        _check_emb_name_suffix_collisions(feat_to_group_to_emb_name, group_name_to_suffix)
        shared_flag = {f: len(set(g2e.values())) > 1 for f, g2e in feat_to_group_to_emb_name.items()}

        emb_name_to_feature_to_shared_name = defaultdict(dict)
        for fg in feature_groups:
            gname = fg.group_name
            q_dims, s_dims, o_dims, shared_q, shared_s = [], [], [], [], []
            group_seq_len = None
            for name in fg.feature_names:
                feature = name_to_feature[name]
                shared_name = name
                if feature.is_sparse:
                    cfg = feature.emb_config
                    feature.mc_module(device)
                    cfg.name = feat_to_group_to_emb_name[name][gname]
                    if shared_flag[name]:
                        shared_name = shared_name + "@" + cfg.name
                    emb_name_to_feature_to_shared_name[cfg.name][name] = shared_name
                    const = feature.parameter_constraints(cfg)
                    _add_embedding_config(dim_to_emb_configs[cfg.embedding_dim], cfg, check_pooling=False)
                    self.has_sparse = True
                    if const is not None:
                        self._dim_to_emb_constraints[cfg.embedding_dim][cfg.name] = const
                    if feature.is_sequence and feature.value_dim != 1:
                        self.has_mulval_seq = True      # embedding.py:1150-1155: pooled per step in forward
                else:
                    if feature.is_sequence:
                        self.has_sequence_dense = True
                    else:
                        self.has_dense = True
                info = _SeqInfo(shared_name, name, feature.is_sparse, feature.pooling_type.value.lower()
                                if feature.is_sparse else "sum", feature.value_dim, feature.is_sequence)
                out_dim = feature.output_dim
                if feature.is_sequence:
                    shared_s.append(info)
                    s_dims.append(out_dim)
                    group_seq_len = feature.sequence_length
                else:
                    shared_q.append(info)
                    q_dims.append(out_dim)
                o_dims.append(out_dim)
            self._group_to_shared_query[gname] = shared_q
            self._group_to_shared_sequence[gname] = shared_s
            self._group_to_sequence_length[gname] = group_seq_len
            self._group_total_dim[f"{gname}.query"] = sum(q_dims)
            self._group_total_dim[f"{gname}.sequence"] = sum(s_dims)
            self._group_output_dims[f"{gname}.query"] = q_dims
            self._group_output_dims[f"{gname}.sequence"] = s_dims
            self._group_output_dims[gname] = o_dims

        self.ec_dict = nn.ModuleDict()
        for k, cfgs in dim_to_emb_configs.items():   # one collection per embedding_dim (:1193-1197)
            self.ec_dict[str(k)] = EmbeddingCollection(list(cfgs.values()), device=device)
        self.ec_dict_features = OrderedDict()
        for k, ec in self.ec_dict.items():
            pairs = []
            for cfg, out_keys in zip(ec.embedding_configs(), ec.embedding_names_by_table()):
                for fname, okey in zip(cfg.feature_names, out_keys):
                    pairs.append((okey, emb_name_to_feature_to_shared_name[cfg.name].get(fname, fname)))
            self.ec_dict_features[k] = pairs

    def group_dims(self, group_name: str) -> List[int]:
        return self._group_output_dims[group_name]

    def group_total_dim(self, group_name: str) -> int:
        if "." in group_name:
            return self._group_total_dim[group_name]
        return self._group_total_dim[f"{group_name}.query"] + self._group_total_dim[f"{group_name}.sequence"]

    def all_group_total_dim(self) -> Dict[str, int]:
        return self._group_total_dim

    def parameter_constraints(self, prefix: str = "") -> Dict[str, ParameterConstraints]:
        out = {}
        for dim, consts in self._dim_to_emb_constraints.items():
            for k, v in consts.items():
                out[f"{prefix}ec_dict.{dim}.{k}"] = v
        return out

    def has_group(self, group_name: str) -> bool:
        return group_name.split(".")[0] in self._group_output_dims

    def forward(self, sparse_feature: Optional[KeyedJaggedTensor], dense_feature: Optional[KeyedTensor],
                sequence_dense_features: Optional[Dict[str, JaggedTensor]] = None,
                sequence_mulval_lengths: Optional[KeyedJaggedTensor] = None, *unused) -> Dict[str, torch.Tensor]:
        jt_dict: Dict[str, JaggedTensor] = {}
        if self.has_sparse:
            for pairs, ec in zip(self.ec_dict_features.values(), self.ec_dict.values()):
                d_jt = ec(sparse_feature)                      # embedding.py:1301 <- drop-in boundary
                for okey, shared in pairs:
                    jt_dict[shared] = d_jt[okey]
        dense_t = dense_feature.to_dict() if self.has_dense else {}
        mulval_len = sequence_mulval_lengths.to_dict() if (self.has_mulval_seq and sequence_mulval_lengths is not None) \
            else {}
        done = set()
        for infos in self._group_to_shared_sequence.values():
            for info in infos:
                if info.name in done:
                    continue
                done.add(info.name)
                if not info.is_sparse:
                    jt_dict[info.name] = sequence_dense_features[info.name]
                elif info.value_dim != 1:
                    # several ids per sequence step (embedding.py:1354-1366): the un-pooled rows of the step's ids are
                    # reduced with the same ATen op the reference calls; `values` of the length JT = ids per step,
                    # its `lengths` = steps per sample
                    if info.raw_name not in mulval_len:
                        raise ValueError(f"sequence feature {info.raw_name} is multi-value: the batch must carry "
                                         "sequence_mulval_lengths for it")
                    length_jt = mulval_len[info.raw_name]
                    jt = jt_dict[info.name]
                    vals = torch.segment_reduce(jt.values(), info.pooling, lengths=length_jt.values().to(torch.int64))
                    if info.pooling == "mean":
                        vals = torch.nan_to_num(vals, nan=0.0)
                    jt_dict[info.name] = JaggedTensor(values=vals, lengths=length_jt.lengths())

        results: Dict[str, torch.Tensor] = {}
        query_cache: Dict[str, torch.Tensor] = {}
        for gname, infos in self._group_to_shared_query.items():
            parts = []
            for info in infos:
                if info.name in query_cache:
                    q = query_cache[info.name]
                elif info.is_sparse:
                    jt = jt_dict[info.name]
                    if info.value_dim == 1:
                        q = jt.to_padded_dense(1).squeeze(1)                      # :1429
                    else:
                        # multi-value id (value_dim 0 is the non-sequence default): ATen segment_reduce,
                        # the same op the reference calls at :1432-1436 (K8)
                        q = torch.segment_reduce(jt.values(), info.pooling, lengths=jt.lengths().to(torch.int64))
                        if info.pooling == "mean":
                            q = torch.nan_to_num(q, nan=0.0)
                    query_cache[info.name] = q
                else:
                    q = dense_t[info.name]
                parts.append(q)
            if parts:
                results[f"{gname}.query"] = torch.cat(parts, dim=1)
        for gname, infos in self._group_to_shared_sequence.items():
            parts = []
            T = 1
            rows_only = self._group_to_is_jagged[gname] or gname in getattr(self, "_jagged_for_attention", ())
            for i, info in enumerate(infos):
                jt = jt_dict[info.name]
                if i == 0:
                    seq_len = jt.lengths()
                    results[f"{gname}.sequence_length"] = seq_len
                    if gname in getattr(self, "_jagged_for_attention", ()):
                        # the consumer (DIN attention, csrc/tzk_din.cu) works on the gather's rows as
                        # they are: no longest-length host read (:1468), no padded [B, T, D] tensor (:1480)
                        results[f"{gname}.sequence_offsets"] = jt.offsets()
                    elif not rows_only:
                        T = int(torch.max(seq_len).item()) if seq_len.numel() else 0   # host sync as in :1468
                parts.append(jt.values() if rows_only else jt.to_padded_dense(T))  # :1475-1480
            if parts:
                results[f"{gname}.sequence"] = torch.cat(parts, dim=-1)
        return results


class EmbeddingGroup(nn.Module):
    """Applies embedding lookup transformation for feature groups (embedding.py:167-547).

    Args mirror the reference: features, feature_groups, wide_embedding_dim, wide_init_fn, device.
    `seq_encoder_factory(cfg, group_total_dim) -> nn.Module` builds DEEP-group sequence encoders.
    """

    def __init__(self, features: List[BaseFeature], feature_groups: List[Message],
                 wide_embedding_dim: Optional[int] = None, wide_init_fn: Optional[str] = None,
                 device: Optional[torch.device] = None, seq_encoder_factory=None) -> None:
        super().__init__()
        if device is None:
            device = torch.device("meta")
        self._features = features
        self._feature_groups = feature_groups
        self._name_to_feature = {x.name: x for x in features}
        self._name_to_feature_group = OrderedDict((x.group_name, x) for x in feature_groups)
        self.emb_impls = nn.ModuleDict()
        self.seq_emb_impls = nn.ModuleDict()
        impl_to_feat_groups, impl_to_seq_groups = defaultdict(list), defaultdict(list)
        self._group_name_to_impl_key = {}
        self._group_name_to_seq_encoder_configs = defaultdict(list)
        self._grouped_features_keys: List[str] = []

        seq_group_names: List[str] = []
        for fg in feature_groups:
            gname = fg.group_name
            self._inspect_and_supplement_feature_group(fg, seq_group_names)
            by_data_group = defaultdict(list)
            for fname in fg.feature_names:
                by_data_group[self._name_to_feature[fname].data_group].append(fname)
            for sg in fg.sequence_groups:
                for fname in sg.feature_names:
                    by_data_group[self._name_to_feature[fname].data_group].append(fname)
            if len(by_data_group) > 1:
                info = [",".join(v) for v in by_data_group.values()]
                raise ValueError(f"Feature {info} should not belong to same feature group.")
            impl_key = list(by_data_group.keys())[0]
            self._group_name_to_impl_key[gname] = impl_key
            if fg.group_type in SEQ_TYPES:
                impl_to_seq_groups[impl_key].append(fg)
                self._grouped_features_keys += [gname + ".query", gname + ".sequence", gname + ".sequence_length"]
            else:
                impl_to_feat_groups[impl_key].append(fg)
                for sg in fg.sequence_groups:
                    sg_copy = Message("SeqGroupConfig")
                    sg_copy.CopyFrom(sg)
                    if fg.embedding_name_suffix and not sg_copy.embedding_name_suffix:
                        sg_copy.embedding_name_suffix = fg.embedding_name_suffix
                    impl_to_seq_groups[impl_key].append(sg_copy)
                if len(fg.sequence_encoders) > 0:
                    self._group_name_to_seq_encoder_configs[gname] = list(fg.sequence_encoders)
                self._grouped_features_keys.append(gname)

        for k, v in impl_to_feat_groups.items():
            self.emb_impls[k] = EmbeddingGroupImpl(features, v, wide_embedding_dim, wide_init_fn, device)
        for k, v in impl_to_seq_groups.items():
            self.seq_emb_impls[k] = SequenceEmbeddingGroupImpl(features, v, device)

        self._group_name_to_seq_encoders = nn.ModuleDict()
        for gname, enc_cfgs in self._group_name_to_seq_encoder_configs.items():
            if seq_encoder_factory is None:
                raise NotImplementedError("sequence_encoders inside DEEP groups need a seq_encoder_factory")
            seq_emb = self.seq_emb_impls[self._group_name_to_impl_key[gname]]
            self._group_name_to_seq_encoders[gname] = nn.ModuleList(
                [seq_encoder_factory(c, seq_emb.all_group_total_dim()) for c in enc_cfgs])

        self._group_feature_dims = OrderedDict()
        for fg in feature_groups:
            gname = fg.group_name
            if fg.group_type not in SEQ_TYPES:
                dims = OrderedDict()
                dims.update(self.emb_impls[self._group_name_to_impl_key[gname]].group_feature_dims(gname))
                if gname in self._group_name_to_seq_encoders:
                    for i, enc in enumerate(self._group_name_to_seq_encoders[gname]):
                        dims[f"{gname}_seq_encoder_{i}"] = enc.output_dim()
                self._group_feature_dims[gname] = dims
        self._grouped_features_keys.sort()

    # ---- validation (embedding.py:308-372) ------------------------------------------------------------------
    def _inspect_and_supplement_feature_group(self, fg: Message, seq_group_names: List[str]) -> None:
        gname = fg.group_name
        sgs, encs = list(fg.sequence_groups), list(fg.sequence_encoders)
        if fg.group_type == "DEEP":
            if not sgs and not encs:
                return
            if sgs and not encs:
                raise ValueError(f"{gname} group has sequence_groups,but no sequence_encoders ")
            if encs and not sgs:
                raise ValueError(f"{gname} group has sequence_encoders,but no sequence_groups ")
            if len(sgs) > 1:
                for sg in sgs:
                    if not sg.HasField("group_name"):
                        raise ValueError(f"{gname} has many sequence_groups, every sequence_group must has group_name")
            elif not sgs[0].HasField("group_name"):
                sgs[0].group_name = gname
            for sg in sgs:
                if sg.group_name in seq_group_names:
                    raise ValueError(f"has repeat sequences groups_name: {sg.group_name}")
                seq_group_names.append(sg.group_name)
            has_enc = {sg.group_name: False for sg in sgs}
            for enc in encs:
                kinds = [k for k in enc._values]
                sc = getattr(enc, kinds[0])
                if not sc.HasField("input") and len(sgs) == 1:
                    sc.input = sgs[0].group_name
                if not sc.HasField("input"):
                    raise ValueError(f"{gname} group has multi sequence_groups, so sequence_encoders must has input")
                if sc.input not in has_enc:
                    raise ValueError(f"{gname} sequence_encoder input {sc.input} not in sequence_groups")
                has_enc[sc.input] = True
            for k, v in has_enc.items():
                if not v:
                    raise ValueError(f"{gname} sequence_groups {k} not has seq_encoder")
        elif sgs or encs:
            raise ValueError(f"{gname} group group_type is not DEEP, sequence_groups and sequence_encoders "
                             "must configured in DEEP")

    # ---- queries (embedding.py:374-445) ---------------------------------------------------------------------
    def grouped_features_keys(self) -> List[str]:
        return self._grouped_features_keys

    def group_names(self) -> List[str]:
        return list(self._name_to_feature_group.keys())

    def _true(self, group_name: str) -> str:
        return group_name.split(".")[0] if "." in group_name else group_name

    def group_dims(self, group_name: str) -> List[int]:
        true = self._true(group_name)
        impl_key = self._group_name_to_impl_key[true]
        if self._name_to_feature_group[true].group_type in SEQ_TYPES:
            return self.seq_emb_impls[impl_key].group_dims(group_name)
        dims = self.emb_impls[impl_key].group_dims(group_name)
        if group_name in self._group_name_to_seq_encoders:
            dims += [enc.output_dim() for enc in self._group_name_to_seq_encoders[group_name]]
        return dims

    def group_total_dim(self, group_name: str) -> int:
        true = self._true(group_name)
        if self._name_to_feature_group[true].group_type in SEQ_TYPES:
            return self.seq_emb_impls[self._group_name_to_impl_key[true]].group_total_dim(group_name)
        return sum(self._group_feature_dims[group_name].values())

    def group_feature_dims(self, group_name: str) -> Dict[str, int]:
        if self._name_to_feature_group[self._true(group_name)].group_type in SEQ_TYPES:
            raise ValueError("not support sequence group")
        return self._group_feature_dims[group_name]

    def group_type(self, group_name: str) -> str:
        return self._name_to_feature_group[self._true(group_name)].group_type

    def has_group(self, group_name: str) -> bool:
        return self._true(group_name) in self._name_to_feature_group

    def parameter_constraints(self, prefix: str = "") -> Dict[str, ParameterConstraints]:
        out = {}
        for k, impl in self.emb_impls.items():
            out.update(impl.parameter_constraints(f"{prefix}emb_impls.{k}."))
        for k, impl in self.seq_emb_impls.items():
            out.update(impl.parameter_constraints(f"{prefix}seq_emb_impls.{k}."))
        return out

    def set_jagged_for_attention(self, group_names) -> None:
        """The named sequence groups emit `<g>.sequence` as the un-pooled gather's rows [N, D] plus `<g>.sequence_offsets`
        [B + 1] instead of the padded [B, T, D] tensor (consumers: jagged DIN attention)."""
        names = set(group_names)
        for impl in self.seq_emb_impls.values():
            impl._jagged_for_attention = names & set(impl._group_to_shared_sequence)

    def sparse_collections(self):
        """All arena-backed collections (what BaseModel.sparse_parameters discovers, models/model.py:162-201)."""
        for m in self.modules():
            if isinstance(m, (EmbeddingBagCollection, EmbeddingCollection)):
                yield m

    # ---- forward (embedding.py:447-536) ---------------------------------------------------------------------
    def forward(self, batch) -> Dict[str, torch.Tensor]:
        result: Dict[str, torch.Tensor] = {}
        for key, impl in self.emb_impls.items():
            result.update(impl(batch.sparse_features[key] if impl.has_sparse else None,
                               batch.dense_features[key] if impl.has_dense else None))
        for key, impl in self.seq_emb_impls.items():
            result.update(impl(batch.sparse_features[key] if impl.has_sparse else None,
                               batch.dense_features[key] if impl.has_dense else None,
                               batch.sequence_dense_features,
                               batch.sequence_mulval_lengths.get(key) if impl.has_mulval_seq else None))
        for gname, encs in self._group_name_to_seq_encoders.items():
            new = torch.cat([enc(result) for enc in encs], dim=-1)
            result[gname] = torch.cat([result[gname], new], dim=-1) if gname in result else new
        return result

    def predict(self, batch) -> List[torch.Tensor]:
        grouped = self.forward(batch)
        return [grouped[k] for k in self._grouped_features_keys]
