"""Loading the reference's artefacts without HyperPyYAML / speechbrain.

The reference restores a model with ``load_hyperpyyaml(hparams file)`` + ``mod.load_state_dict(torch.load(CKPT_PATH/
name.ckpt))`` for ``name`` in ``modules: {encoder, decoder, masknet}`` (``Mamba-TasNet/inference.ipynb`` cells 0-1; the
checkpointer writes those files into ``<save_folder>/CKPT+<timestamp>/``, ``hparams/WSJ0Mix/mambatasnet_S.yaml:179-186``).
Here:

* ``read_hparams_yaml`` resolves the handful of scalar keys the model graph needs (``N_encoder_out``, ``out_channels``,
  ``kernel_size``, ``n_mamba``, ``ssm_dim``, ``mamba_expand``, ``mamba_conv``, ``num_spks``, ``sample_rate`` and the
  norm / direction switches) from such a yaml, including ``!ref <key>`` indirections with integer arithmetic
  (``kernel_stride: !ref <kernel_size> // 2``), and refuses configurations this build does not implement;
* ``load_checkpoint_dir`` reads the three state_dicts; ``separator_from_checkpoint`` builds the drop-in module.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Dict, Optional

import torch

from .hparams import DPHParams, HParams

_SCALAR = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*):\s*(.*?)\s*$")
_REF = re.compile(r"<([A-Za-z_][A-Za-z0-9_]*)>")
MODULE_FILES = ("encoder", "decoder", "masknet")            # yaml `modules:` keys = checkpoint file stems


def _strip_comment(v: str) -> str:
    out, quote = [], None
    for ch in v:
        if quote:
            quote = None if ch == quote else quote
        elif ch in "'\"":
            quote = ch
        elif ch == "#":
            break
        out.append(ch)
    return "".join(out).strip()


def _parse_value(v: str):
    low = v.lower()
    if low in ("true", "false"):
        return low == "true"
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v.strip("'\"")


def read_yaml_scalars(path: str) -> Dict[str, object]:
    """Top-level ``key: scalar`` entries of a HyperPyYAML file, with ``!ref`` expressions over other top-level scalars
    resolved.  Object tags (``!new:`` / ``!name:`` / ``!apply:``) and nested mappings are skipped."""
    raw: Dict[str, str] = {}
    with open(path) as f:
        for line in f:
            if not line or line[0] in " \t#\n-":
                continue                      # nested, comment or list item
            m = _SCALAR.match(line.rstrip("\n"))
            if not m:
                continue
            val = _strip_comment(m.group(2))
            if val == "" or val.startswith(("!new:", "!name:", "!apply:", "&", "*", "[", "{")):
                continue
            raw[m.group(1)] = val
    resolved: Dict[str, object] = {}

    def resolve(key, depth=0):
        if key in resolved:
            return resolved[key]
        if key not in raw or depth > 20:
            raise KeyError(key)
        v = raw[key]
        if v.startswith("!ref"):
            expr = v[4:].strip()
            names = _REF.findall(expr)
            if len(names) == 1 and expr == f"<{names[0]}>":
                out = resolve(names[0], depth + 1)
            else:
                for n in names:
                    expr = expr.replace(f"<{n}>", repr(resolve(n, depth + 1)))
                if not re.fullmatch(r"[0-9eE\.\s\+\-\*/\(\)]*", expr):
                    raise ValueError(f"{key}: unsupported !ref expression {v!r}")
                out = eval(expr, {"__builtins__": {}})          # digits and arithmetic only (checked above)
        else:
            out = _parse_value(v)
        resolved[key] = out
        return out

    for k in list(raw):
        try:
            resolve(k)
        except (KeyError, ValueError):
            pass                                # a reference into a nested object: not a model scalar
    return resolved


def _masknet_class(path: str) -> Optional[str]:
    with open(path) as f:
        for line in f:
            m = re.match(r"^MaskNet:\s*(?:&\S+\s+)?!new:(\S+)", line)   # the saved copy carries yaml anchors (&id006)
            if m:
                return m.group(1)
    return None


def read_object_block(path: str, key: str, scalars: Optional[Dict[str, object]] = None) -> Dict[str, object]:
    """The ``name: value`` lines nested under a top-level ``key: !new:Class`` entry (constructor keyword arguments, e.g.
    ``MaskNet: ... mask_nonlinear: softmax``); ``!ref <x>`` values are looked up in ``scalars``, references to other
    objects (``intra_model: !ref <Mambaintra>``) are returned as the referenced key's name."""
    scalars = scalars or {}
    out: Dict[str, object] = {}
    inside = False
    with open(path) as f:
        for line in f:
            if not inside:
                inside = re.match(rf"^{re.escape(key)}:\s*(?:&\S+\s+)?!new:", line) is not None
                continue
            if line.strip() == "" or line.lstrip().startswith("#"):
                continue
            if line[0] not in " \t":
                break                                   # next top-level entry
            m = _SCALAR.match(line.strip())
            if not m:
                continue
            val = _strip_comment(m.group(2))
            if val.startswith("!ref"):
                names = _REF.findall(val)
                expr = val[4:].strip()
                if len(names) == 1 and expr == f"<{names[0]}>":
                    out[m.group(1)] = scalars.get(names[0], names[0])
                elif all(n in scalars for n in names):
                    for n in names:
                        expr = expr.replace(f"<{n}>", repr(scalars[n]))
                    if re.fullmatch(r"[0-9eE\.\s\+\-\*/\(\)]*", expr):
                        out[m.group(1)] = eval(expr, {"__builtins__": {}})   # digits and arithmetic only
            elif val != "":
                out[m.group(1)] = _parse_value(val)
    return out


MASKNET_TASNET = ("modules.mamba_masknet.MaskNet",)
MASKNET_DP = ("speechbrain.lobes.models.dual_path.Dual_Path_Model", "modules.dual_path.Dual_Path_Model_Skip")


def read_hparams_yaml(path: str, name: Optional[str] = None):
    """``HParams`` of a Mamba-TasNet recipe yaml (``hparams/WSJ0Mix/mambatasnet_*.yaml:108-155``) or ``DPHParams`` of a
    DPMamba one (``hparams/WSJ0Mix/dpmamba_*.yaml:108-174``: ``MaskNet: !new:...Dual_Path_Model`` over two
    ``MambaBlocksSequential``), also from the ``hyperparams.yaml`` copy speechbrain stores next to the checkpoints.
    Everything the engines implement is accepted (``rms_norm: False``, ``mask_nonlinear: softmax``, ``num_spks: 3``,
    ``bidirectional: False``, ``fused_add_norm: True`` -- the latter only picks between two identical reference code paths,
    ``modules/mamba_blocks.py:195-210``); what they do not implement raises ``NotImplementedError``."""
    cls = _masknet_class(path)
    if cls is not None and cls not in MASKNET_TASNET + MASKNET_DP:
        raise NotImplementedError(f"{path}: MaskNet is {cls}; built: {MASKNET_TASNET + MASKNET_DP}")
    y = read_yaml_scalars(path)
    dp = cls in MASKNET_DP
    need = ["N_encoder_out", "out_channels", "kernel_size"] + (["n_dp", "chunk_size"] if dp else ["n_mamba"])
    missing = [k for k in need if k not in y]
    if missing:
        raise KeyError(f"{path}: missing {missing}")
    stride = y.get("kernel_stride", y["kernel_size"] // 2)
    if stride != y["kernel_size"] // 2:
        raise NotImplementedError(f"kernel_stride {stride} != kernel_size // 2")
    blk = read_object_block(path, "MaskNet", y) if cls is not None else {}
    nm = name or os.path.splitext(os.path.basename(path))[0]
    common = dict(kernel_size=int(y["kernel_size"]), d_state=int(y.get("ssm_dim", 16)), expand=int(y.get("mamba_expand", 2)),
                  d_conv=int(y.get("mamba_conv", 4)), n_spk=int(blk.get("num_spks", blk.get("n_spk", y.get("num_spks", 2)))),
                  sample_rate=int(y.get("sample_rate", 8000)))
    if not dp:
        return HParams(nm, int(y["N_encoder_out"]), int(y["out_channels"]), int(y["n_mamba"]),
                       mask_nonlinear=str(blk.get("mask_nonlinear", "relu")), rms_norm=bool(y.get("rms_norm", True)),
                       bidirectional=bool(y.get("bidirectional", True)), **common)
    # speechbrain Dual_Path_Model defaults: norm="ln", linear_layer_after_inter_intra=True, use_global_pos_enc=False
    if str(blk.get("norm", "ln")) != "ln" or bool(blk.get("linear_layer_after_inter_intra", True)) or \
            bool(blk.get("use_global_pos_enc", False)):
        raise NotImplementedError(f"{path}: Dual_Path_Model with norm != ln / linear_layer_after_inter_intra / "
                                  "use_global_pos_enc is not built")
    if not y.get("bidirectional", True) or not y.get("rms_norm", True):
        raise NotImplementedError(f"{path}: the dual-path stacks are built bidirectional with RMSNorm (every shipped recipe)")
    n_mamba_dp = int(y.get("n_mamba_dp", 2))
    if n_mamba_dp < 2 or n_mamba_dp % 2:
        raise NotImplementedError(f"{path}: n_mamba_dp = {n_mamba_dp} (intra and inter stacks take n_mamba_dp // 2 layers each)")
    return DPHParams(nm, int(y["N_encoder_out"]), int(y["out_channels"]), int(y["n_dp"]), bool(y.get("skip_around_intra", True)),
                     chunk_size=int(y["chunk_size"]), n_mamba_dp=n_mamba_dp,
                     skip_n_block=int(y.get("skip_n_block", 0)) if cls == MASKNET_DP[1] else 0, **common)


def find_checkpoint_dir(save_folder: str) -> str:
    """Newest ``CKPT+*`` directory of a speechbrain ``save_folder`` (or the folder itself if it holds the files)."""
    if all(os.path.exists(os.path.join(save_folder, f"{m}.ckpt")) for m in MODULE_FILES):
        return save_folder
    cands = sorted(glob.glob(os.path.join(save_folder, "CKPT+*")))
    cands = [c for c in cands if all(os.path.exists(os.path.join(c, f"{m}.ckpt")) for m in MODULE_FILES)]
    if not cands:
        raise FileNotFoundError(f"no CKPT+* directory with {MODULE_FILES} .ckpt files under {save_folder}")
    return cands[-1]


def load_checkpoint_dir(ckpt_dir: str) -> Dict[str, dict]:
    """``{encoder, decoder, masknet}`` state_dicts as the reference saved them (``torch.save(module.state_dict())``)."""
    d = find_checkpoint_dir(ckpt_dir)
    return {m: torch.load(os.path.join(d, f"{m}.ckpt"), map_location="cpu", weights_only=True) for m in MODULE_FILES}


def save_checkpoint_dir(sds: Dict[str, dict], ckpt_dir: str) -> str:
    """Write state_dicts in the reference's layout (used by tests and to hand weights back to the reference)."""
    os.makedirs(ckpt_dir, exist_ok=True)
    for m in MODULE_FILES:
        torch.save({k: v.detach().cpu() for k, v in sds[m].items()}, os.path.join(ckpt_dir, f"{m}.ckpt"))
    return ckpt_dir


def separator_from_checkpoint(hparams_yaml: str, ckpt_dir: str, mode: str = "fp32", use_graph: bool = True,
                              device="cuda"):
    """The reference's ``inference.ipynb`` cells 0-1 in one call: yaml + CKPT dir -> ``MambaTasNetSeparator`` (or
    ``DPMambaSeparator`` for a dpmamba recipe) with the weights loaded (``strict=True``) on ``device``."""
    from .modules import DPMambaSeparator, MambaTasNetSeparator
    hp = read_hparams_yaml(hparams_yaml)
    cls = DPMambaSeparator if isinstance(hp, DPHParams) else MambaTasNetSeparator
    sep = cls.from_hparams(hp, mode=mode, use_graph=use_graph)
    sep.load_reference_state_dicts(load_checkpoint_dir(ckpt_dir), strict=True)
    return sep.to(device)
