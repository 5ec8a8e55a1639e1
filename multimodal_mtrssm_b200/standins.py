"""Shape-compatible stand-ins for `cnn.Encoder` / `cnn.Decoder` (absent third-party package, cnn 3.1.1 @ c669849,
`uv.lock:442-444`; assumption A7 of SURVEY.md §8(c)).  They sit OUTSIDE the rollout kernel (north_star: encoders and
decoders stay PyTorch/cuDNN at the kernel boundary) and exist so the end-to-end configs can run where `cnn` is not
installed: encoder [*, 1, 32, 32] -> [*, linear_sizes[-1]], decoder [*, F] -> [*, 1, 32, 32] with the configured
output activation.  Residual blocks / coord-conv of the real package are not reproduced."""

from __future__ import annotations

import math

import torch
from torch import Tensor, nn


def _act(name: str) -> nn.Module:
    return getattr(nn, name)()


def _nhwc(stack: nn.Module, x: Tensor) -> Tensor:
    """On CUDA the conv stacks run channels-last (weights re-strided once, in place: same Parameters, same state_dict): cuDNN's
    tensor-core kernels are NHWC, and with NCHW tensors 29 % of a B = 256 training step was nchw<->nhwc conversion kernels."""
    if not x.is_cuda:
        return x
    if not getattr(stack, "_nhwc_done", False):
        stack.to(memory_format=torch.channels_last)
        stack._nhwc_done = True  # noqa: SLF001
    return x.contiguous(memory_format=torch.channels_last)


class Encoder(nn.Module):
    def __init__(self, config: dict) -> None:
        super().__init__()
        chans = [1, *config["channels"]]
        layers: list[nn.Module] = []
        for i, (k, s, p) in enumerate(zip(config["kernel_sizes"], config["strides"], config["paddings"])):
            layers += [nn.Conv2d(chans[i], chans[i + 1], k, s, p), _act(config["activation_name"])]
        self.conv = nn.Sequential(*layers)
        self.head = nn.LazyLinear(config["linear_sizes"][-1])
        self.out_act = _act(config.get("out_activation_name", "Identity"))

    def forward(self, x: Tensor) -> Tensor:
        lead = x.shape[:-3]
        y = self.conv(_nhwc(self.conv, x.reshape(-1, *x.shape[-3:]))).flatten(1)
        return self.out_act(self.head(y)).reshape(*lead, -1)


class Decoder(nn.Module):
    def __init__(self, config: dict) -> None:
        super().__init__()
        act = config["activation_name"]
        sizes = config["linear_sizes"]
        self.conv_in_shape = tuple(config["conv_in_shape"])
        lin: list[nn.Module] = [nn.LazyLinear(sizes[0]), _act(act)]
        for a, b in zip(sizes[:-1], sizes[1:]):
            lin += [nn.Linear(a, b), _act(act)]
        if sizes[-1] != math.prod(self.conv_in_shape):
            lin += [nn.Linear(sizes[-1], math.prod(self.conv_in_shape)), _act(act)]
        self.lin = nn.Sequential(*lin)
        chans = [self.conv_in_shape[0], *config["channels"]]
        n = len(config["channels"])
        deconv: list[nn.Module] = []
        for i, (k, s, p, op) in enumerate(zip(config["kernel_sizes"], config["strides"], config["paddings"], config["output_paddings"])):
            deconv.append(nn.ConvTranspose2d(chans[i], chans[i + 1], k, s, p, op))
            deconv.append(_act(config["out_activation_name"] if i == n - 1 else act))
        self.deconv = nn.Sequential(*deconv)

    def forward(self, x: Tensor) -> Tensor:
        lead = x.shape[:-1]
        y = self.lin(x.reshape(-1, x.shape[-1])).reshape(-1, *self.conv_in_shape)
        y = self.deconv(_nhwc(self.deconv, y))
        return y.reshape(*lead, *y.shape[-3:])


def materialize(model: nn.Module, feature_dim: int, device: torch.device | str = "cpu") -> None:
    """Run the lazy layers once so parameters exist (before building an optimiser / DDP-style bucket)."""
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, Encoder):
                m(torch.zeros(1, 1, 32, 32, device=device))
            elif isinstance(m, Decoder):
                m(torch.zeros(1, feature_dim, device=device))
