"""CPU: HyperPyYAML-free recipe reader and the reference's checkpoint-directory layout."""
import os

import pytest
import torch

from avse_challenge_b200 import CONFIGS, init_state_dicts, checkpoint

RECIPE = """# Generated 2024-06-17 from:
seed: 1234
__set_seed: !apply:torch.manual_seed [1234]
num_spks: 2 # set to 3 for wsj0-3mix
sample_rate: 8000
speed_changes: &id001 [95, 100, 105]
speed_perturb: !new:speechbrain.augment.time_domain.SpeedPerturb
  orig_freq: 8000
  speeds: *id001
# Encoder parameters
N_encoder_out: {N}
out_channels: {D}
kernel_size: 16
kernel_stride: !ref <kernel_size> // 2
bidirectional: {bidir}
n_mamba: {n}
ssm_dim: 16
mamba_expand: 2
mamba_conv: 4
fused_add_norm: False
rms_norm: True
residual_in_fp32: False
Encoder: !new:speechbrain.lobes.models.dual_path.Encoder
    kernel_size: !ref <kernel_size>
    out_channels: !ref <N_encoder_out>
MaskNet: !new:{cls}
    enc_dim: !ref <N_encoder_out>
    n_mamba: !ref <n_mamba>
modules:
    encoder: !ref <Encoder>
    masknet: !ref <MaskNet>
"""


def _write(tmp_path, **kw):
    args = dict(N=256, D=256, n=16, bidir="True", cls="modules.mamba_masknet.MaskNet")
    args.update(kw)
    p = tmp_path / "hyperparams.yaml"
    p.write_text(RECIPE.format(**args))
    return str(p)


def test_recipe_reader_resolves_refs_and_matches_shipped_config(tmp_path):
    hp = checkpoint.read_hparams_yaml(_write(tmp_path), name="S")
    assert hp == CONFIGS["S"]
    y = checkpoint.read_yaml_scalars(_write(tmp_path))
    assert y["kernel_stride"] == 8 and y["num_spks"] == 2 and y["bidirectional"] is True
    assert "Encoder" not in y and "speed_perturb" not in y and "modules" not in y


def test_recipe_reader_refuses_what_is_not_built(tmp_path):
    assert checkpoint.read_hparams_yaml(_write(tmp_path, bidir="False"), name="S_causal") == CONFIGS["S"].causal()
    with pytest.raises(NotImplementedError):
        checkpoint.read_hparams_yaml(_write(tmp_path, cls="speechbrain.lobes.models.dual_path.Dual_Path_Model"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/Mamba-TasNet/hparams"), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["XS", "S", "M", "L"])
def test_recipe_reader_on_the_reference_recipes(name):
    for p in (f"/root/reference/Mamba-TasNet/hparams/WSJ0Mix/mambatasnet_{name}.yaml",
              f"/root/reference/Mamba-TasNet/ckpts/WSJ0Mix/mambatasnet_{name}/1234/hyperparams.yaml"):
        assert checkpoint.read_hparams_yaml(p, name=name) == CONFIGS[name]


def test_checkpoint_dir_round_trip(tmp_path):
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 5)
    save = tmp_path / "save"
    checkpoint.save_checkpoint_dir(sds, str(save / "CKPT+2024-03-03+18-23-45+00"))
    (save / "CKPT+2024-01-01+00-00-00+00").mkdir()          # an older, incomplete one is ignored
    got = checkpoint.load_checkpoint_dir(str(save))
    assert set(got) == {"encoder", "decoder", "masknet"}
    for m in got:
        assert got[m].keys() == sds[m].keys()
        assert all(torch.equal(got[m][k], sds[m][k]) for k in got[m])
    with pytest.raises(FileNotFoundError):
        checkpoint.find_checkpoint_dir(str(tmp_path / "nothing"))
