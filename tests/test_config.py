"""Config surface: the reference's examples/*.config load unchanged, and the generated configs are equivalent."""
import glob
import os

import pytest

from torcheasyrec_b200 import example_configs
from torcheasyrec_b200.config import config_to_kwargs, edit_config, load_pipeline_config, parse_text
from torcheasyrec_b200.features import create_features

REF_EXAMPLES = "/root/reference/examples"
have_ref = pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present (GPU box)")


@have_ref
def test_all_reference_examples_parse():
    files = sorted(glob.glob(os.path.join(REF_EXAMPLES, "*.config")))
    assert len(files) >= 17
    for p in files:
        cfg = load_pipeline_config(p)
        assert cfg.model_config.WhichOneof("model") is not None, p
        assert len(cfg.feature_configs) > 0


@have_ref
@pytest.mark.parametrize("name", list(example_configs.GENERATORS))
def test_generated_config_equals_reference_example(name):
    ref = load_pipeline_config(os.path.join(REF_EXAMPLES, name + ".config"))
    ours = parse_text(example_configs.GENERATORS[name]())
    assert ours.to_dict() == ref.to_dict()


@pytest.mark.parametrize("name,groups", [
    ("dlrm_criteo", {"dense": 13, "sparse": 26}),
    ("deepfm_criteo", {"wide": 26, "fm": 26, "deep": 39}),
    ("mmoe_taobao", {"all": 16}),
    ("multi_tower_din_taobao", {"deep": 16, "seq": 6}),
])
def test_generated_configs_structure(name, groups):
    cfg = parse_text(example_configs.GENERATORS[name]())
    got = {g.group_name: len(g.feature_names) for g in cfg.model_config.feature_groups}
    assert got == groups
    assert cfg.train_config.sparse_optimizer.WhichOneof("optimizer") == "adagrad_optimizer"
    assert abs(cfg.train_config.sparse_optimizer.adagrad_optimizer.lr - 0.001) < 1e-12
    assert cfg.train_config.cuda_matmul_allow_tf32 is False  # train.proto field 14 default
    feats = create_features(list(cfg.feature_configs), fg_mode=cfg.data_config.fg_mode)
    assert len(feats) == sum(1 if fc.WhichOneof("feature") != "sequence_feature" else
                             len(fc.sequence_feature.features) for fc in cfg.feature_configs)


def test_text_format_features():
    cfg = parse_text('''
      # comment
      model_dir: 'a' "b"
      train_config { num_steps: 10 sparse_optimizer { sgd_optimizer { lr: 1e-2 } } }
      model_config {
        feature_groups { group_name: "g" feature_names: ["x", "y"] group_type: WIDE }
        dlrm { final { hidden_units: 8 hidden_units: 4 } }
      }
      unknown_block < inner: 3 inner: 4 >
    ''')
    assert cfg.model_dir == "ab"
    assert cfg.train_config.num_steps == 10 and cfg.train_config.save_checkpoints_steps == 1000
    assert cfg.train_config.sparse_optimizer.sgd_optimizer.lr == 0.01
    g = cfg.model_config.feature_groups[0]
    assert g.feature_names == ["x", "y"] and g.group_type == "WIDE"
    assert config_to_kwargs(cfg.model_config.dlrm.final)["hidden_units"] == [8, 4]
    assert cfg.model_config.dlrm.arch_with_sparse is True and not cfg.model_config.dlrm.HasField("dense_mlp")
    assert cfg.unknown_block.inner == [3, 4]
    edit_config(cfg, {"train_config.num_steps": 5, "model_config.feature_groups[0].group_name": "h"})
    assert cfg.train_config.num_steps == 5 and cfg.model_config.feature_groups[0].group_name == "h"
