"""Episode store + loader for the rollout's training step.

The reference keeps ONE `.pt` FILE PER EPISODE PER MODALITY (`act_000.pt`, `audio_obs_000.pt`, `vision_obs_000.pt`, written by
`EpisodeDataModule._process_episode_data`, `models/mrssm/dataset.py:105-134`), reads them item by item through six
`EpisodeDataset`s (`models/dataset.py:45-66`: `transform(load_tensor(path))`) zipped by a `StackDataset` (`models/mrssm/dataset.py:
155-183`), and collates them in 4 worker processes with `prefetch_factor=1` (`models/dataset.py:335-342`).  With the rollout fused
this becomes the end-to-end limiter: a batch of 256 costs 1536 `torch.load` calls + a collate copy + an unpinned H2D copy.

Here the same data lives in ONE CONTIGUOUS PINNED TENSOR PER MODALITY, `[N, T, ...]`, built once from the reference's own
processed directory (same file names, same sorted order, same 80/20 split as `split_path_list`, `models/dataset.py:69-81`):

* `EpisodeStore.from_processed_dir(dir)`  -- reads the per-episode files once; `save(path)` / `load(path)` keep the packed form
  (`<path>/{act,audio_obs,vision_obs}.pt`, three files instead of 3 N);
* `EpisodeStore.batch(indices)`           -- the 6-tuple `(action_in, audio_in, vision_in, action_tgt, audio_tgt, vision_tgt)` of
  `StackDataset` + default collate for those episodes, gathered with three `index_select`s into pinned staging buffers and passed
  through the same six transforms (`per_episode=True` applies them episode by episode exactly like `EpisodeDataset.__getitem__`;
  the default applies each transform once to the whole `[B, T, ...]` batch, which is the same function for the elementwise /
  per-frame transforms of `transform.py`);
* `PinnedEpisodeLoader`                   -- the `train_dataloader()` replacement: shuffled (or sequential) index batches, gathered
  on the host while the previous batch computes, copied to the GPU by `dp.PinnedPrefetcher` on a side stream (double buffered),
  yielding DEVICE 6-tuples.  `drop_last=False`, `shuffle` and the batch size follow `models/dataset.py:335-342`.

Nothing here touches the numerics of the path: batches are bit-identical to the reference's loader for the same indices
(tests/test_episode_store_cpu.py builds both from the same files).
"""

from __future__ import annotations

from pathlib import Path
from typing import Callable, Iterator, Sequence

import torch
from torch import Tensor

Transform = Callable[[Tensor], Tensor]
MODALITIES = ("act", "audio_obs", "vision_obs")  # file-name prefixes of the reference's processed directory


def _identity(x: Tensor) -> Tensor:
    return x


def _load_tensor(path: Path) -> Tensor:
    """models/dataset.py:45-66 (`load_tensor`): `.npy` or a `.pt` holding one Tensor."""
    if path.suffix == ".npy":
        import numpy as np

        return torch.Tensor(np.load(path))
    if path.suffix == ".pt":
        t = torch.load(path, weights_only=False)
        if isinstance(t, Tensor):
            return t
    msg = f"Unknown file extension: {path.suffix}"
    raise ValueError(msg)


def split_indices(n: int, train_ratio: float = 0.8) -> tuple[range, range]:
    """`split_path_list` (models/dataset.py:69-81) on the sorted episode order: the first int(n * ratio) episodes train."""
    k = int(n * train_ratio)
    return range(0, k), range(k, n)


class EpisodeStore:
    """One contiguous (pinned when CUDA is available) tensor per modality: `act [N,T,A]`, `audio_obs`, `vision_obs [N,T,C,H,W]`."""

    def __init__(self, act: Tensor, audio_obs: Tensor, vision_obs: Tensor, pin: bool | None = None) -> None:
        n = act.shape[0]
        if audio_obs.shape[0] != n or vision_obs.shape[0] != n:
            msg = f"modalities disagree on the number of episodes: {act.shape[0]}, {audio_obs.shape[0]}, {vision_obs.shape[0]}"
            raise ValueError(msg)
        pin = torch.cuda.is_available() if pin is None else pin
        prep = lambda t: t.contiguous().pin_memory() if pin else t.contiguous()  # noqa: E731
        self.tensors = {"act": prep(act), "audio_obs": prep(audio_obs), "vision_obs": prep(vision_obs)}
        self.pinned = pin

    def __len__(self) -> int:
        return self.tensors["act"].shape[0]

    # ---- construction ---------------------------------------------------------------------------------------------------
    @classmethod
    def from_processed_dir(cls, directory: str | Path, pin: bool | None = None) -> "EpisodeStore":
        """Reads the reference's processed directory (one file per episode per modality, sorted glob order as in
        `EpisodeDataModule.setup`, models/mrssm/dataset.py:155-160).  Episodes of one modality must share a shape."""
        directory = Path(directory)
        stacked = {}
        for name in MODALITIES:
            paths = sorted(directory.glob(f"{name}*"))
            if not paths:
                msg = f"no `{name}*` files in {directory} (expected the layout written by EpisodeDataModule._process_episode_data)"
                raise FileNotFoundError(msg)
            stacked[name] = torch.stack([_load_tensor(p) for p in paths])
        return cls(stacked["act"], stacked["audio_obs"], stacked["vision_obs"], pin=pin)

    def save(self, path: str | Path) -> None:
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        for name, t in self.tensors.items():
            torch.save(t.clone(), path / f"{name}.pt")  # clone: do not serialise the pinned storage flag

    @classmethod
    def load(cls, path: str | Path, pin: bool | None = None) -> "EpisodeStore":
        path = Path(path)
        return cls(*(torch.load(path / f"{name}.pt", weights_only=True) for name in MODALITIES), pin=pin)

    # ---- access ---------------------------------------------------------------------------------------------------------------
    def batch(self, indices: Sequence[int] | Tensor, transforms: Sequence[Transform] | None = None, per_episode: bool = False,
              out: dict[str, Tensor] | None = None) -> tuple[Tensor, ...]:
        """The collated 6-tuple of `StackDataset(act_in, audio_in, vision_in, act_tgt, audio_tgt, vision_tgt)[indices]`.
        `transforms`: the six transforms in that order (default: identity).  `out`: optional staging buffers per modality
        (`{name: [B, ...]}`, e.g. pinned) the gathers write into."""
        idx = torch.as_tensor(indices, dtype=torch.long)
        tf = list(transforms) if transforms is not None else [_identity] * 6
        if len(tf) != 6:  # noqa: PLR2004
            msg = f"expected 6 transforms (input x3, target x3), got {len(tf)}"
            raise ValueError(msg)
        gathered = {}
        for name, t in self.tensors.items():
            dst = None if out is None else out[name][: idx.numel()]
            gathered[name] = torch.index_select(t, 0, idx, out=dst) if dst is not None else torch.index_select(t, 0, idx)
        order = [gathered["act"], gathered["audio_obs"], gathered["vision_obs"]] * 2
        if per_episode:
            return tuple(torch.stack([f(x[i]) for i in range(x.shape[0])]) for f, x in zip(tf, order))
        return tuple(f(x) for f, x in zip(tf, order))


class PinnedEpisodeLoader:
    """`train_dataloader()` / `val_dataloader()` replacement (models/dataset.py:316-363) over an `EpisodeStore`: yields DEVICE
    6-tuples; the host gather of batch i+1 and its H2D copy (side stream, `dp.PinnedPrefetcher`) overlap batch i's step.

    `episodes`: the subset to draw from (e.g. `split_indices(len(store))[0]`); `transforms`: the six transforms, applied ON THE
    DEVICE to the copied batch (three H2D copies instead of six: inputs and targets share their source tensors)."""

    def __init__(self, store: EpisodeStore, batch_size: int, device: torch.device | str, episodes: Sequence[int] | None = None,
                 shuffle: bool = True, transforms: Sequence[Transform] | None = None, drop_last: bool = False,
                 generator: torch.Generator | None = None) -> None:
        self.store, self.batch_size, self.device = store, batch_size, torch.device(device)
        self.episodes = torch.as_tensor(list(episodes) if episodes is not None else range(len(store)), dtype=torch.long)
        self.shuffle, self.drop_last, self.generator = shuffle, drop_last, generator
        self.transforms = list(transforms) if transforms is not None else [_identity] * 6
        if self.device.type != "cuda":
            msg = "PinnedEpisodeLoader feeds a CUDA device (use EpisodeStore.batch on the host)"
            raise RuntimeError(msg)
        if not store.pinned:
            msg = "PinnedEpisodeLoader needs a pinned store (EpisodeStore(..., pin=True))"
            raise RuntimeError(msg)
        from .dp import PinnedPrefetcher

        # two pinned staging sets (one being copied, one being gathered) and the prefetcher's two device sets
        shape = lambda t: (batch_size, *t.shape[1:])  # noqa: E731
        self._staging = [{k: torch.empty(shape(t), dtype=t.dtype).pin_memory() for k, t in store.tensors.items()} for _ in range(2)]
        self._copied = [torch.cuda.Event() for _ in range(2)]
        self._pre = PinnedPrefetcher(self._staging[0], self.device)

    def __len__(self) -> int:
        n = self.episodes.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _index_batches(self) -> list[Tensor]:
        order = self.episodes[torch.randperm(self.episodes.numel(), generator=self.generator)] if self.shuffle else self.episodes
        chunks = list(order.split(self.batch_size))
        if self.drop_last and chunks and chunks[-1].numel() < self.batch_size:
            chunks.pop()
        return chunks

    def _submit(self, k: int, idx: Tensor) -> None:
        stage = self._staging[k & 1]
        self._copied[k & 1].synchronize()  # the copy that last read this staging set has finished
        for name, t in self.store.tensors.items():
            torch.index_select(t, 0, idx, out=stage[name][: idx.numel()])
        self._pre.submit(stage)
        self._copied[k & 1].record(self._pre.stream)

    def __iter__(self) -> Iterator[tuple[Tensor, ...]]:
        chunks = self._index_batches()
        if not chunks:
            return
        self._submit(0, chunks[0])
        for k, idx in enumerate(chunks):
            slot, dev = self._pre.next()
            if k + 1 < len(chunks):
                self._submit(k + 1, chunks[k + 1])
            n = idx.numel()
            order = [dev["act"][:n], dev["audio_obs"][:n], dev["vision_obs"][:n]] * 2
            yield tuple(f(x) for f, x in zip(self.transforms, order))
            self._pre.release(slot)  # the consumer's work queued so far was the last use of this device buffer
