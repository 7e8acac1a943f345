"""``InputPadder`` with the reference's arithmetic (benchmark/utils.py:57-80): centred replicate padding
so that height and width become multiples of ``divisor``."""
import torch.nn.functional as F


class InputPadder:
    def __init__(self, dims, divisor=16):
        self.ht, self.wd = dims[-2:]
        extra_h = (-self.ht) % divisor
        extra_w = (-self.wd) % divisor
        self._pad = [extra_w // 2, extra_w - extra_w // 2, extra_h // 2, extra_h - extra_h // 2]

    def pad(self, *inputs):
        padded = [F.pad(x, self._pad, mode='replicate') for x in inputs]
        return padded[0] if len(padded) == 1 else padded

    def unpad(self, *inputs):
        l, r, t, b = self._pad
        cropped = [x[..., t:x.shape[-2] - b, l:x.shape[-1] - r] for x in inputs]
        return cropped[0] if len(cropped) == 1 else cropped
