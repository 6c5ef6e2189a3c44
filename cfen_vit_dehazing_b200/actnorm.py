"""ActNorm2d parameter holder (mirror of reference models/actnorm.py:8-78).

In the B200 build an ActNorm2d is a per-channel affine ``(x + bias) * exp(weight)`` that the
CUDA kernels fold into the epilogue of the convolution in front of it; this class only owns
the three state_dict entries (``weight``, ``bias``, ``initialized``) so that checkpoints
round-trip bit-for-bit.  The data-dependent first-forward initialisation of the reference
(models/actnorm.py:25-37) is reproduced on the GPU by the first ``dec_ipt.forward`` (C ABI ``cfen_forward_calibrate``).
"""
import torch
import torch.nn as nn

__all__ = ["ActNorm2d"]


class ActNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-12):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        # reference allocates these uninitialised (models/actnorm.py:14-15); zeros are the
        # identity affine and are overwritten by calibration or by load_state_dict.
        self.weight = nn.Parameter(torch.zeros(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("initialized", torch.tensor(0))

    def forward(self, x):  # pragma: no cover - not on the product path
        raise RuntimeError(
            "ActNorm2d is fused into the sm_100a kernels of dec_ipt.forward; "
            "it has no standalone (CPU/eager) path in cfen_vit_dehazing_b200")

    def extra_repr(self):
        return str(self.num_features)
