"""state_dict key <-> C-ABI weight-field mapping for the two rollouts (SURVEY.md §8(b) `state_dict` keys)."""

from __future__ import annotations

from typing import Mapping

from torch import Tensor

from ._lib import MR_WEIGHT_FIELDS, MT_WEIGHT_FIELDS


def _mlp(prefix: str) -> list[str]:
    return [f"{prefix}.0.weight", f"{prefix}.0.bias", f"{prefix}.2.weight", f"{prefix}.2.bias"]


# order == _lib.MR_WEIGHT_FIELDS
MR_STATE_KEYS: tuple[str, ...] = (
    *_mlp("transition.action_state_projector"),
    "transition.rnn_cell.weight_ih", "transition.rnn_cell.weight_hh", "transition.rnn_cell.bias_ih", "transition.rnn_cell.bias_hh",
    *_mlp("transition.rnn_to_prior_projector"),
    *_mlp("audio_representation.rnn_to_post_projector"),
    *_mlp("vision_representation.rnn_to_post_projector"),
)

# order == _lib.MT_WEIGHT_FIELDS
MT_STATE_KEYS: tuple[str, ...] = (
    "l_rnn._d2h.weight", "l_rnn._d2h.bias", "l_rnn._input2h.weight", "l_rnn._input2h.bias",
    "h_rnn._d2h.weight", "h_rnn._d2h.bias", "h_rnn._input2h.weight", "h_rnn._input2h.bias",
    *_mlp("l_prior"), *_mlp("h_prior"), *_mlp("h_posterior"),
    *_mlp("audio_representation.rnn_to_post_projector"),
    *_mlp("vision_representation.rnn_to_post_projector"),
)

assert len(MR_STATE_KEYS) == len(MR_WEIGHT_FIELDS) and len(MT_STATE_KEYS) == len(MT_WEIGHT_FIELDS)


def mrssm_weight_list(params: Mapping[str, Tensor]) -> list[Tensor]:
    return [params[k] for k in MR_STATE_KEYS]


def mtrssm_weight_list(params: Mapping[str, Tensor]) -> list[Tensor]:
    return [params[k] for k in MT_STATE_KEYS]
