"""Identity stand-in for `dropblock.DropBlock2D` — TEST INFRASTRUCTURE ONLY.

The reference uses DropBlock only inside the encoder's purifier
(`networks/pemp_stage1.py:73-80`), which is outside the hot path and inactive in
eval mode.  Nothing under ``pemp_b200/`` may import this module.
"""
import torch.nn as nn


class DropBlock2D(nn.Module):
    def __init__(self, drop_prob=0.0, block_size=1):
        super().__init__()
        self.drop_prob = drop_prob
        self.block_size = block_size

    def forward(self, x):
        return x
