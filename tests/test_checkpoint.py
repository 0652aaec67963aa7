"""CPU: HyperPyYAML-free recipe reader and the reference's checkpoint-directory layout."""
import os

import pytest
import torch

from avse_challenge_b200 import CONFIGS, init_state_dicts, checkpoint
from avse_challenge_b200.hparams import DP_CONFIGS, DPHParams

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
        checkpoint.read_hparams_yaml(_write(tmp_path, cls="speechbrain.lobes.models.dual_path.SepformerWrapper"))
    with pytest.raises(KeyError):              # a dual-path MaskNet without the dual-path scalars
        checkpoint.read_hparams_yaml(_write(tmp_path, cls="speechbrain.lobes.models.dual_path.Dual_Path_Model"))


def test_recipe_reader_accepts_every_option_the_engines_implement(tmp_path):
    """rms_norm False, mask_nonlinear softmax and num_spks 3 are built, so their recipes must load."""
    import dataclasses
    txt = RECIPE.format(N=256, D=256, n=16, bidir="True", cls="modules.mamba_masknet.MaskNet")
    txt = txt.replace("rms_norm: True", "rms_norm: False").replace("num_spks: 2", "num_spks: 3")
    txt = txt.replace("    n_mamba: !ref <n_mamba>\n", "    n_mamba: !ref <n_mamba>\n    mask_nonlinear: softmax  # mamba_masknet.py:133\n")
    p = tmp_path / "h.yaml"
    p.write_text(txt)
    hp = checkpoint.read_hparams_yaml(str(p), name="S")
    assert hp == dataclasses.replace(CONFIGS["S"], rms_norm=False, mask_nonlinear="softmax", n_spk=3)


DP_RECIPE = """num_spks: 2
sample_rate: 8000
N_encoder_out: 128
out_channels: 128
kernel_size: 16
kernel_stride: !ref <kernel_size> // 2
n_dp: 8
chunk_size: 250
skip_n_block: {skip}
skip_around_intra: False
bidirectional: True
n_mamba_dp: 2
ssm_dim: 16
mamba_expand: 2
mamba_conv: 4
fused_add_norm: False
rms_norm: True
Mambaintra: &id002 !new:modules.mamba_blocks.MambaBlocksSequential
    n_mamba: !ref <n_mamba_dp> // 2
    d_model: !ref <out_channels>
MaskNet: &id006 !new:{cls}

    num_spks: !ref <num_spks>
    in_channels: !ref <N_encoder_out>
    num_layers: !ref <n_dp>
    K: !ref <chunk_size>
    intra_model: *id002
    norm: {norm}
    linear_layer_after_inter_intra: {lin}
    skip_around_intra: !ref <skip_around_intra>
Decoder: !new:speechbrain.lobes.models.dual_path.Decoder
    in_channels: !ref <N_encoder_out>
"""


def test_recipe_reader_dpmamba(tmp_path):
    """dpmamba_*.yaml:108-174 (and the anchored copy speechbrain saves) -> DPHParams."""
    def write(**kw):
        a = dict(cls="speechbrain.lobes.models.dual_path.Dual_Path_Model", norm="ln", lin="False", skip=0)
        a.update(kw)
        p = tmp_path / "dp.yaml"
        p.write_text(DP_RECIPE.format(**a))
        return str(p)
    assert checkpoint.read_hparams_yaml(write(), name="dp_XS") == DP_CONFIGS["XS"]
    hp = checkpoint.read_hparams_yaml(write(cls="modules.dual_path.Dual_Path_Model_Skip", skip=2), name="dp_XS")
    assert isinstance(hp, DPHParams) and hp.skip_n_block == 2
    for bad in (dict(norm="gln"), dict(lin="True")):
        with pytest.raises(NotImplementedError):
            checkpoint.read_hparams_yaml(write(**bad))


@pytest.mark.skipif(not os.path.isdir("/root/reference/Mamba-TasNet/hparams"), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["XS", "S", "M", "L"])
def test_recipe_reader_on_the_reference_dpmamba_recipes(name):
    for p in (f"/root/reference/Mamba-TasNet/hparams/WSJ0Mix/dpmamba_{name}.yaml",
              f"/root/reference/Mamba-TasNet/ckpts/WSJ0Mix/dpmamba_{name}/1234/hyperparams.yaml"):
        assert checkpoint.read_hparams_yaml(p, name="dp_" + name) == DP_CONFIGS[name]


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
