"""CPU: host-side logic -- state_dict contract, hparams, synthetic data, C-ABI exports (no GPU compute)."""
import ctypes
import os
import re

import pytest
import torch

from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture, pit_si_snr
from avse_challenge_b200 import modules, _lib
from tests.helpers import load_golden_forward, hp_from_sds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,params", [("XS", 2210176), ("S", 7926528), ("M", 15647488), ("L", 58951168)])
def test_param_counts_match_reference(name, params):
    hp = CONFIGS[name]
    sep = modules.MambaTasNetSeparator.from_hparams(hp)
    n = sum(p.numel() for p in sep.masknet.parameters())
    n += sum(p.numel() for p in sep.encoder.parameters()) + sum(p.numel() for p in sep.decoder.parameters())
    assert n == params  # SURVEY.md section 0 (probe counts of the reference modules)


@pytest.mark.parametrize("tag", ["tiny_refinit", "tiny_trained"])
def test_reference_state_dict_loads_strict(golden_dir, tag):
    sds, _, _ = load_golden_forward(os.path.join(golden_dir, f"forward_{tag}.npz"))
    hp = hp_from_sds(sds)
    sep = modules.MambaTasNetSeparator.from_hparams(hp)
    sep.load_reference_state_dicts(sds, strict=True)
    for grp, mod in (("encoder", sep.encoder), ("masknet", sep.masknet), ("decoder", sep.decoder)):
        mine = mod.state_dict()
        assert set(mine) == set(sds[grp])
        for k, v in sds[grp].items():
            assert mine[k].shape == v.shape, k
            assert torch.equal(mine[k], v), k


def test_init_state_dicts_has_reference_keys():
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 3)
    sep = modules.MambaTasNetSeparator.from_hparams(hp)
    sep.load_reference_state_dicts(sds, strict=True)


def test_causal_modules_have_mamba_ssm_state_dict_keys():
    """bidirectional=False -> mamba_ssm.Mamba parameter set (no *_b keys), strict load of a causal state_dict."""
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 3)
    assert not any("_b" in k.split(".")[-2] + k.split(".")[-1] for k in sds["masknet"] if "mixer" in k)
    sep = modules.MambaTasNetSeparator.from_hparams(hp)
    sep.load_reference_state_dicts(sds, strict=True)
    with pytest.raises(RuntimeError):      # a bidirectional checkpoint does not fit a causal model
        sep.load_reference_state_dicts(init_state_dicts(CONFIGS["tiny"], 3), strict=True)
    with pytest.raises(NotImplementedError):   # step() is the unidirectional mixer's interface
        modules.Mamba(64).step(torch.zeros(1, 1, 64), torch.zeros(1, 128, 4), torch.zeros(1, 128, 16))


def test_unknown_mask_nonlinear_raises_like_the_reference():
    with pytest.raises(ValueError):                       # modules/mamba_masknet.py:138
        modules.MaskNet(64, 64, n_mamba=1, d_model=64, mask_nonlinear="tanh")
    modules.MaskNet(64, 64, n_mamba=1, d_model=64, mask_nonlinear="softmax")


@pytest.mark.parametrize("kw", [dict(d_state=8),
                                dict(d_conv=3)])
def test_unsupported_options_raise(kw):
    with pytest.raises(NotImplementedError):
        modules.MaskNet(64, 64, n_mamba=1, d_model=64, **kw)


def test_no_cpu_fallback():
    hp = CONFIGS["tiny"]
    sep = modules.MambaTasNetSeparator.from_hparams(hp)
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    with pytest.raises(_lib.MtnError):
        sep(torch.zeros(1, 160))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "avse_challenge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "oracle/" not in src and "selscan_ref" not in src, f


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "mtn_b200.h")).read()
    declared = set(re.findall(r"\b(mtn_[a-z0-9_]+)\s*\(", header))
    declared -= {"mtn_stream_t"}
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mtn_abi_version.restype = ctypes.c_int
    assert lib.mtn_abi_version() >= 1


def test_struct_layouts_match_header(built_lib):
    """ctypes Structures must mirror the C structs field for field."""
    header = open(os.path.join(ROOT, "include", "mtn_b200.h")).read()
    for cname, cls in (("mtn_gemm_args", _lib.GemmArgs), ("mtn_scan_args", _lib.ScanArgs),
                       ("mtn_gn_apply_args", _lib.GnApplyArgs)):
        body = re.search(r"typedef struct \{([^}]*)\} " + cname + ";", header).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(void|float|int)\s*\*?\s*", "", decl)
            names += [n.strip().lstrip("*").strip() for n in decl.split(",")]
        assert names == [f[0] for f in cls._fields_], cname


def test_integration_stub_matches_the_header(built_lib):
    """The ctypes stub printed in INTEGRATION.md section 2 is what a maintainer pastes: its struct must have the fields of
    mtn_scan_args in the header's order and the size the built library reports (a short struct makes the kernel read
    sum_delta / L_last / dtp from whatever follows it in memory)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    snippet = re.search(r"```python\n(import ctypes, torch.*?)```", doc, re.S).group(1)
    cls_src = re.search(r"(class MtnScanArgs\(ctypes\.Structure\):.*?)\n\n", snippet, re.S).group(1)
    ns = {"ctypes": ctypes}
    exec(cls_src, ns)
    stub = ns["MtnScanArgs"]
    assert [f[0] for f in stub._fields_] == [f[0] for f in _lib.ScanArgs._fields_]
    assert [f[1] for f in stub._fields_] == [f[1] for f in _lib.ScanArgs._fields_]
    lib = ctypes.CDLL(built_lib)
    lib.mtn_sizeof_scan_args.restype = ctypes.c_size_t
    assert ctypes.sizeof(stub) == lib.mtn_sizeof_scan_args() == ctypes.sizeof(_lib.ScanArgs)
    lib.mtn_abi_version.restype = ctypes.c_int
    m = re.search(r"mtn_abi_version\(\) == (\d+)", snippet)
    assert m and int(m.group(1)) == lib.mtn_abi_version() == _lib.EXPECTED_ABI
    header = open(os.path.join(ROOT, "include", "mtn_b200.h")).read()
    assert int(re.search(r"#define MTN_ABI_VERSION (\d+)", header).group(1)) == _lib.EXPECTED_ABI
    # every name the stub's call passes as a keyword is a field
    call = re.search(r"a = MtnScanArgs\((.*?)\)\n    rc", snippet, re.S).group(1)
    assert set(re.findall(r"(\w+)=", call)) == {f[0] for f in stub._fields_}


def test_loader_refuses_a_library_of_another_abi(built_lib, monkeypatch):
    """A stale / experiment build (git-ignored .so, rebuilt on mtimes only) must not be handed this binding's structs."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "EXPECTED_ABI", _lib.EXPECTED_ABI + 1)
    with pytest.raises(_lib.MtnError, match="ABI"):
        _lib.load()
    monkeypatch.undo()
    _lib._lib = None
    assert _lib.load().mtn_abi_version() == _lib.EXPECTED_ABI


def test_shape_caches_are_bounded():
    """per-shape workspaces / graphs must not grow without bound over a variable-length test set."""
    from avse_challenge_b200._cache import LRUDict
    evicted = []
    ws = LRUDict(3, on_evict=lambda k, v: evicted.append(k))
    for T in range(10):
        ws[(1, T)] = object()
        ws.get((1, 0))                      # keep one shape hot
    assert len(ws) == 3 and (1, 0) in ws and (1, 9) in ws
    assert evicted == [(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 7)]


def test_synth_is_deterministic_and_scaled():
    mix, src = synth_mixture(2, 4000, seed=5)
    mix2, _ = synth_mixture(2, 4000, seed=5)
    assert torch.equal(mix, mix2)
    assert torch.allclose(src.pow(2).mean(dim=1).sqrt(), torch.full((2, 2), 0.05), rtol=1e-3)
    assert torch.allclose(mix, src.sum(-1))
    s = pit_si_snr(src.flip(-1) * 1.7, src)
    assert (s > 60).all()


def test_stream_fragment_packer_matches_mma_layout():
    """stream_fused.pack_fragments lays a weight slice out as mma.sync.m16n8k16 A fragments (row-major 16 x 16 tile: lane
    (g, t) holds a0 = (g, 2t..2t+1), a1 = (g+8, 2t..), a2 = (g, 2t+8..), a3 = (g+8, 2t+8..)); rebuild the tiles from the packed
    tensor exactly as the tensor core reads them and compare with hi + lo of the weight."""
    from avse_challenge_b200.stream_fused import pack_fragments, eligible
    from avse_challenge_b200 import CONFIGS
    g = torch.Generator().manual_seed(0)
    W = torch.randn(96, 80, generator=g)
    rows = torch.cat([torch.arange(16, 48), torch.arange(64, 96)])
    ks = torch.arange(16, 80)
    fr = pack_fragments(W, rows, ks)                          # [ct, ks, 2, 32, 8]
    assert fr.dtype == torch.bfloat16 and tuple(fr.shape) == (4, 4, 2, 32, 8)
    Wt = W[rows][:, ks]
    hi = Wt.to(torch.bfloat16)
    lo = (Wt - hi.float()).to(torch.bfloat16)
    for plane, want in ((0, hi), (1, lo)):
        for ct in range(4):
            for s in range(4):
                tile = torch.zeros(16, 16, dtype=torch.bfloat16)
                for lane in range(32):
                    gg, t = lane // 4, lane % 4
                    regs = fr[ct, s, plane, lane].view(4, 2)
                    tile[gg, 2 * t:2 * t + 2] = regs[0]
                    tile[gg + 8, 2 * t:2 * t + 2] = regs[1]
                    tile[gg, 2 * t + 8:2 * t + 10] = regs[2]
                    tile[gg + 8, 2 * t + 8:2 * t + 10] = regs[3]
                assert torch.equal(tile, want[16 * ct:16 * ct + 16, 16 * s:16 * s + 16])
    assert eligible(CONFIGS["S"].causal(), "fp32") and eligible(CONFIGS["tiny"].causal(), "fp32")
    assert not eligible(CONFIGS["S"], "fp32") and not eligible(CONFIGS["S"].causal(), "bf16")
