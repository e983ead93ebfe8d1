"""Feature objects: FeatureConfig -> table specs (EmbeddingBagConfig / EmbeddingConfig) and constraints.

Mirrors the slice of tzrec/features/ that the hot path consumes: `BaseFeature`
(feature.py:380), `IdFeature` (id_feature.py:25-87), `RawFeature` (raw_feature.py:25-66) incl. their sequence
variants and grouped `sequence_feature` sub-features (create_features, feature.py:1161-1220).  Feature
generation (pyfg, string ops, Arrow parsing) is host-side IO and out of scope (§8 "out of scope"): batches
reach the engine already bucketised (row A0).
"""

from typing import Any, Dict, List, Optional

from .config import Message
from .embedding_modules import DataType, EmbeddingBagConfig, EmbeddingConfig, PoolingType

BASE_DATA_GROUP = "__BASE__"
NEG_DATA_GROUP = "__NEG__"


class ParameterConstraints:
    """[EXT] torchrec ParameterConstraints subset used by tzrec (feature.proto:6-13, feature.py:359-371)."""

    def __init__(self, sharding_types: Optional[List[str]] = None, compute_kernels: Optional[List[str]] = None):
        self.sharding_types = list(sharding_types) if sharding_types else None
        self.compute_kernels = list(compute_kernels) if compute_kernels else None

    def __repr__(self) -> str:
        return f"ParameterConstraints(sharding_types={self.sharding_types}, compute_kernels={self.compute_kernels})"


def build_embedding_constraints(cfg: Message) -> ParameterConstraints:
    return ParameterConstraints(list(cfg.sharding_types), list(cfg.compute_kernels))


def create_init_fn(spec: str):
    """tzrec/utils/init_util.py: "nn.init.uniform_,a=-0.01,b=0.01" -> callable(tensor)."""
    import torch  # noqa: F401
    from torch import nn  # noqa: F401

    parts = [p.strip() for p in spec.split(",")]
    fn = eval(parts[0], {"nn": nn, "torch": torch})  # same mechanism as the reference (trusted config)
    kwargs = {}
    for p in parts[1:]:
        k, v = p.split("=")
        kwargs[k.strip()] = float(v) if any(c in v for c in ".e") else int(v)
    out = lambda t, fn=fn, kwargs=kwargs: fn(t, **kwargs)
    out.__repr_str__ = spec
    return out


class BaseFeature:
    """tzrec/features/feature.py:380 — only the properties the embedding path reads."""

    def __init__(self, feature_config: Message, fg_mode: str = "FG_NONE", is_sequence: bool = False,
                 sequence_name: Optional[str] = None, sequence_delim: Optional[str] = None,
                 sequence_length: Optional[int] = None, sequence_pk: Optional[str] = None, **kwargs: Any) -> None:
        fc_type = feature_config.WhichOneof("feature")
        self._feature_config = feature_config
        self.config = getattr(feature_config, fc_type)
        self.fg_mode = fg_mode
        self._is_sequence = is_sequence
        self._is_grouped_seq = False
        self._data_group = BASE_DATA_GROUP
        self._is_user_feat: Optional[bool] = None
        self.sequence_name = self.sequence_delim = self.sequence_length = self.sequence_pk = None
        if is_sequence:
            if sequence_name is None:
                self.sequence_delim = self.config.sequence_delim
                self.sequence_length = self.config.sequence_length
            else:
                self._is_grouped_seq = True
                self.sequence_name, self.sequence_delim = sequence_name, sequence_delim
                self.sequence_length = sequence_length
                self.sequence_pk = sequence_pk or f"user:{sequence_name}"

    # ---- identity ---------------------------------------------------------------------------------------
    @property
    def name(self) -> str:
        prefix = f"{self.sequence_name}__" if self._is_grouped_seq else ""
        return f"{prefix}{self.config.feature_name}"

    @property
    def feature_config(self) -> Message:
        return self._feature_config

    @property
    def data_group(self) -> str:
        return self._data_group

    @data_group.setter
    def data_group(self, v: str) -> None:
        self._data_group = v

    @property
    def is_user_feat(self) -> bool:
        if self._is_user_feat is None:
            if self._is_grouped_seq:
                return True
            expr = self.config.expression if self.config.HasField("expression") else ""
            return expr.split(":")[0] == "user"
        return self._is_user_feat

    @property
    def is_sequence(self) -> bool:
        return self._is_sequence

    @property
    def is_grouped_sequence(self) -> bool:
        return self._is_grouped_seq

    @property
    def is_weighted(self) -> bool:
        return False

    @property
    def is_sparse(self) -> bool:
        raise NotImplementedError

    @property
    def value_dim(self) -> int:
        raise NotImplementedError

    @property
    def output_dim(self) -> int:
        raise NotImplementedError

    @property
    def num_embeddings(self) -> int:
        raise NotImplementedError

    @property
    def has_embedding(self) -> bool:
        return self.is_sparse

    @property
    def dense_emb_config(self):
        return None

    def mc_module(self, device):  # zero-collision hash is a "next" row (§8f N4)
        if self.is_sparse and self.config._spec("zch") is not None and self.config.HasField("zch"):
            raise NotImplementedError(f"feature {self.name}: zch (managed collision) is not supported yet")
        return None

    # ---- embedding configs (feature.py:586-662) ------------------------------------------------------------
    @property
    def pooling_type(self) -> PoolingType:
        p = self.config.pooling.upper()
        assert p in {"SUM", "MEAN"}, "available pooling type is SUM | MEAN"
        return getattr(PoolingType, p)

    @property
    def _embedding_dim(self) -> int:
        if self.has_embedding:
            assert self.config.embedding_dim > 0, (
                f"embedding_dim of {self.__class__.__name__}[{self.name}] should be greater than 0.")
        return self.config.embedding_dim

    def _common_emb_kwargs(self) -> Dict[str, Any]:
        init_fn = create_init_fn(self.config.init_fn) if self.config.HasField("init_fn") else None
        return dict(num_embeddings=self.num_embeddings, embedding_dim=self._embedding_dim,
                    name=self.config.embedding_name or f"{self.name}_emb", feature_names=[self.name],
                    init_fn=init_fn, data_type=getattr(DataType, self.config.data_type.upper()))

    @property
    def emb_bag_config(self) -> Optional[EmbeddingBagConfig]:
        if not self.is_sparse:
            return None
        cfg = EmbeddingBagConfig(pooling=self.pooling_type, **self._common_emb_kwargs())
        cfg.trainable = self.config.trainable
        cfg.use_dynamicemb = False
        return cfg

    @property
    def emb_config(self) -> Optional[EmbeddingConfig]:
        if not self.is_sparse:
            return None
        cfg = EmbeddingConfig(**self._common_emb_kwargs())
        cfg.trainable = self.config.trainable
        cfg.use_dynamicemb = False
        return cfg

    def parameter_constraints(self, emb_config=None) -> Optional[ParameterConstraints]:
        """feature.py:832-845."""
        if self.config.HasField("embedding_constraints"):
            return build_embedding_constraints(self.config.embedding_constraints)
        return None

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}[{self.name}]"


class IdFeature(BaseFeature):
    """tzrec/features/id_feature.py:25-87."""

    @property
    def is_weighted(self) -> bool:
        return bool(self.config.weighted) if self.config._spec("weighted") else False

    @property
    def value_dim(self) -> int:
        if self.config.HasField("value_dim"):
            return self.config.value_dim
        return 1 if self.is_sequence else 0

    @property
    def output_dim(self) -> int:
        return self.config.embedding_dim

    @property
    def is_sparse(self) -> bool:
        return True

    @property
    def num_embeddings(self) -> int:
        c = self.config
        if c.HasField("zch"):
            return c.zch.zch_size
        if c.HasField("hash_bucket_size"):
            return c.hash_bucket_size
        if c.HasField("num_buckets"):
            return c.num_buckets
        if len(c.vocab_list) > 0:
            # default_value and <OOV> buckets are prepended unless default_bucketize_value is set
            return len(c.vocab_list) + (0 if c.HasField("default_bucketize_value") else 2)
        raise ValueError(f"{self.__class__.__name__}[{self.name}] must set hash_bucket_size"
                         " or num_buckets or vocab_list or vocab_dict or zch.zch_size")


class RawFeature(BaseFeature):
    """tzrec/features/raw_feature.py:25-66 (bucketised when `boundaries` is set, dense otherwise)."""

    @property
    def value_dim(self) -> int:
        return self.config.value_dim if self.config.HasField("value_dim") else 1

    @property
    def is_sparse(self) -> bool:
        return len(self.config.boundaries) > 0

    @property
    def output_dim(self) -> int:
        return self.config.embedding_dim if self.has_embedding else self.value_dim

    @property
    def num_embeddings(self) -> int:
        return len(self.config.boundaries) + 1

    @property
    def has_embedding(self) -> bool:
        if self.is_sparse:
            return True
        if not self._is_sequence and self.config.WhichOneof("dense_emb") is not None:
            raise NotImplementedError(f"feature {self.name}: autodis/mlp dense embeddings are out of scope")
        return False


_FEATURE_CLASSES = {"IdFeature": IdFeature, "RawFeature": RawFeature}


def _feature_class(msg_type: str, name: str):
    if msg_type not in _FEATURE_CLASSES:
        raise NotImplementedError(
            f"feature type {msg_type} ({name}) is outside the hot-path scope (SURVEY.md §2 row 5): only id / raw "
            "features and their sequence variants reach the embedding engine")
    return _FEATURE_CLASSES[msg_type]


def create_features(feature_configs: List[Message], fg_mode: str = "FG_NONE", **kwargs: Any) -> List[BaseFeature]:
    """tzrec/features/feature.py:1161-1220 (without the pyfg DAG pass)."""
    features: List[BaseFeature] = []
    for fc in feature_configs:
        ftype = fc.WhichOneof("feature")
        sub = getattr(fc, ftype)
        if sub._type == "SequenceFeature":
            for sfc in sub.features:
                st = sfc.WhichOneof("feature")
                cls = _feature_class(getattr(sfc, st)._type, st)
                features.append(cls(sfc, fg_mode=fg_mode, is_sequence=True, sequence_name=sub.sequence_name,
                                    sequence_delim=sub.sequence_delim, sequence_length=sub.sequence_length,
                                    sequence_pk=sub.sequence_pk))
        else:
            cls = _feature_class(sub._type, ftype)
            features.append(cls(fc, fg_mode=fg_mode, is_sequence=ftype.startswith("sequence_")))
    return features
