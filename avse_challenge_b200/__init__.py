"""B200-native Mamba-TasNet separator forward (drop-in for ``Mamba-TasNet/modules`` of shangfuu/avse_challenge).

Public surface:
    Encoder, MaskNet, Decoder, MambaBlocksSequential, Block, Mamba   -- reference-named nn.Module drop-ins
    MambaTasNetSeparator                                              -- fused compute_forward equivalent
    SeparatorEngine                                                   -- the kernel plan itself
    CONFIGS / HParams / init_state_dicts                              -- the four shipped configurations
    checkpoint (recipe yaml + CKPT dir loader), scoring (SI-SNR / PIT on device, test_results.csv)
The compute path is ``libmtn_b200.so`` (C ABI in ``include/mtn_b200.h``); importing this package does not
load it, calling any op without it raises.
"""
from .hparams import CONFIGS, DP_CONFIGS, DPHParams, HParams, init_dp_state_dicts, init_state_dicts  # noqa: F401
from .synth import synth_mixture, si_snr, pit_si_snr  # noqa: F401


def __getattr__(name):
    if name in ("Encoder", "MaskNet", "Decoder", "MambaBlocksSequential", "Block", "Mamba", "MambaTasNetSeparator",
                "RMSNorm", "ChannelwiseLayerNorm", "Dual_Path_Model", "Dual_Path_Model_Skip", "Dual_Computation_Block",
                "DPMambaSeparator"):
        from . import modules
        return getattr(modules, name)
    if name == "SeparatorEngine":
        from .engine import SeparatorEngine
        return SeparatorEngine
    if name in ("ShardedSeparator", "SequenceParallelSeparator"):
        from . import parallel
        return getattr(parallel, name)
    if name == "StreamingSeparator":
        from .streaming import StreamingSeparator
        return StreamingSeparator
    raise AttributeError(name)
