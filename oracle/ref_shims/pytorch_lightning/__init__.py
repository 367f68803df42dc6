"""Stand-in for `pytorch_lightning` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY.  The reference uses pytorch_lightning purely as base classes
(/root/reference/rbnet/pcfg.py:44 `pl.LightningModule`, /root/reference/rbnet/util.py:213
`pl.LightningDataModule`); no arithmetic goes through it.
"""
import torch


class LightningModule(torch.nn.Module):
    pass


class LightningDataModule:
    def __init__(self, *args, **kwargs):
        pass
