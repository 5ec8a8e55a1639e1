"""Stand-in for `torchrl.modules.MLP` (absent; torchrl 0.10.1, `uv.lock:3347-3348`) -- assumption A6 of SURVEY.md §8(c):
`MLP(in, out, num_cells=h, depth=1, activation_class=act, activate_last_layer=False)` is
`Sequential(Linear(in, h), act(), Linear(h, out))` with default `nn.Tanh` and sub-module indices 0 and 2."""

from __future__ import annotations

from torch import nn


class MLP(nn.Sequential):
    def __init__(
        self,
        in_features: int,
        out_features: int,
        num_cells: int,
        depth: int = 1,
        activation_class: type[nn.Module] | str = nn.Tanh,
        activate_last_layer: bool = False,
    ) -> None:
        if depth != 1 or activate_last_layer:
            msg = "this stand-in covers the configurations the reference uses: depth=1, activate_last_layer=False"
            raise NotImplementedError(msg)
        if isinstance(activation_class, str):  # YAML: "torch.nn.ELU"
            activation_class = getattr(nn, activation_class.rsplit(".", 1)[-1])
        super().__init__(nn.Linear(in_features, num_cells), activation_class(), nn.Linear(num_cells, out_features))
        self.in_features, self.out_features, self.num_cells = in_features, out_features, num_cells
