"""the packed episode store against the reference's per-episode-file loader.

The comparator below IS the reference's loading scheme restated with torch's own classes (`EpisodeDataset.__getitem__` =
`transform(load_tensor(path))`, models/dataset.py:45-66,84-118; six of them zipped by `StackDataset`, models/mrssm/dataset.py:
155-183; default collate, `DataLoader(batch_size, shuffle)`, models/dataset.py:335-342), run on files written exactly like
`EpisodeDataModule._process_episode_data` writes them (models/mrssm/dataset.py:105-134)."""

from __future__ import annotations

import pytest
import torch
from torch.utils.data import DataLoader, Dataset, StackDataset

from multimodal_mtrssm_b200.episode_store import EpisodeStore, PinnedEpisodeLoader, split_indices


class RefEpisodeDataset(Dataset):
    def __init__(self, paths, transform):
        self.paths, self.transform = paths, transform

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, i):
        return self.transform(torch.load(self.paths[i], weights_only=False))


def write_processed_dir(d, n=13, T=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    for i in range(n):
        torch.save(torch.randn(T, 6, generator=g), d / f"act_{i:03d}.pt")
        torch.save(torch.rand(T, 1, 32, 32, generator=g) * 2 - 1, d / f"audio_obs_{i:03d}.pt")
        torch.save(torch.rand(T, 1, 32, 32, generator=g) * 2 - 1, d / f"vision_obs_{i:03d}.pt")


TRANSFORMS = [lambda a: a * 2.0, lambda x: x + 0.25, lambda x: x.clamp(-0.5, 0.5), lambda a: a, lambda x: x, lambda x: -x]


def reference_dataset(d, episodes):
    lists = [sorted(d.glob(p + "*")) for p in ("act", "audio_obs", "vision_obs")]
    pick = lambda paths: [paths[i] for i in episodes]  # noqa: E731
    sets = [RefEpisodeDataset(pick(lists[j % 3]), TRANSFORMS[j]) for j in range(6)]
    return StackDataset(*sets)


def test_store_batches_equal_the_reference_loader(tmp_path):
    write_processed_dir(tmp_path)
    store = EpisodeStore.from_processed_dir(tmp_path, pin=False)
    assert len(store) == 13
    train, val = split_indices(len(store))
    assert (len(train), len(val)) == (10, 3)  # split_path_list(paths, 0.8): int(13 * 0.8) = 10
    ref = DataLoader(reference_dataset(tmp_path, list(val)), batch_size=2, shuffle=False)
    chunks = [list(val)[i:i + 2] for i in range(0, len(val), 2)]
    for want, idx in zip(ref, chunks):
        for per_episode in (False, True):
            got = store.batch(idx, TRANSFORMS, per_episode=per_episode)
            assert len(got) == 6
            for a, b in zip(got, want):
                assert torch.equal(a, b)
    # packed form round trip: three files instead of 3 N
    store.save(tmp_path / "packed")
    again = EpisodeStore.load(tmp_path / "packed", pin=False)
    assert sorted(p.name for p in (tmp_path / "packed").iterdir()) == ["act.pt", "audio_obs.pt", "vision_obs.pt"]
    for k in store.tensors:
        assert torch.equal(store.tensors[k], again.tensors[k])


def test_store_errors():
    with pytest.raises(ValueError, match="disagree"):
        EpisodeStore(torch.zeros(3, 2, 6), torch.zeros(2, 2, 1, 4, 4), torch.zeros(3, 2, 1, 4, 4), pin=False)
    s = EpisodeStore(torch.zeros(3, 2, 6), torch.zeros(3, 2, 1, 4, 4), torch.zeros(3, 2, 1, 4, 4), pin=False)
    with pytest.raises(ValueError, match="6 transforms"):
        s.batch([0], [lambda x: x])
    with pytest.raises(RuntimeError, match="CUDA"):
        PinnedEpisodeLoader(s, 2, "cpu")


def test_missing_modality_is_reported(tmp_path):
    torch.save(torch.zeros(2, 6), tmp_path / "act_000.pt")
    with pytest.raises(FileNotFoundError, match="audio_obs"):
        EpisodeStore.from_processed_dir(tmp_path, pin=False)


@pytest.mark.gpu
def test_pinned_loader_delivers_every_episode_once_on_the_device(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    write_processed_dir(tmp_path, n=23)
    store = EpisodeStore.from_processed_dir(tmp_path)
    train, _ = split_indices(len(store))
    loader = PinnedEpisodeLoader(store, batch_size=4, device="cuda", episodes=train, shuffle=True, transforms=TRANSFORMS,
                                 generator=torch.Generator().manual_seed(5))
    assert len(loader) == 5  # 18 training episodes, drop_last=False
    seen = []
    for epoch in range(2):
        for batch in loader:
            assert len(batch) == 6 and all(t.is_cuda for t in batch)
            act_tgt = batch[3].cpu()  # identity transform of the action: identifies the episodes of this batch
            for row in act_tgt:
                match = [i for i in train if torch.equal(store.tensors["act"][i], row)]
                assert len(match) == 1
                seen.append(match[0])
            ids = seen[-act_tgt.shape[0]:]
            want = store.batch(ids, TRANSFORMS)
            for a, b in zip(batch, want):
                assert torch.equal(a.cpu(), b)
    assert sorted(seen) == sorted(list(train) * 2)
