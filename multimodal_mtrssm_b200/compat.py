"""Drop-in glue: expose this package's classes under the reference's module paths and instantiate its YAML configs.

`install()` registers alias modules so that the class paths the reference's configs / scripts use resolve to the
B200 implementation:

    multimodal_rssm.models.mrssm.mopoe_mrssm.MoPoE_MRSSM          (mopoe_mrssm/configs/default.yaml:5)
    multimodal_rssm.models.mmtrssm.mopoe_mmtrssm.MoPoE_MMTRSSM     (mopoe_mmtrssm/configs/default.yaml:5)
    multimodal_rssm.models.networks.{Representation,Transition}, multimodal_rssm.models.state.{State,...},
    multimodal_rssm.models.mmtrssm.{MTState,...}, multimodal_rssm.models.objective.likelihood, ...

and, ONLY when the real packages are not importable, stand-ins for the third-party names those configs mention
(`distribution_extension.MultiOneHotFactory`, `torchrl.modules.MLP`, `cnn.Encoder/Decoder`).
`instantiate()` / `load_model()` are a small `class_path` / `init_args` loader (the subset of jsonargparse the
`model:` section of the reference's YAML needs)."""

from __future__ import annotations

import importlib
import importlib.util
import sys
import types
from pathlib import Path
from typing import Any


def _module(name: str, **attrs: Any) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__dict__.setdefault("__path__", [])  # behaves as a package for sub-module aliases
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    return mod


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install(*, override_reference: bool = True) -> None:
    """Register the aliases (idempotent).  `override_reference=False` keeps an installed `multimodal_rssm`."""
    from . import core, distribution, mlp, mopoe_mmtrssm, mopoe_mrssm, mtstate, networks, objective, standins, state

    if _missing("distribution_extension"):
        de = _module(
            "distribution_extension", Distribution=distribution.Distribution, MultiOneHotFactory=distribution.MultiOneHotFactory,
            MultiOneHot=distribution.MultiOneHot, kl_divergence=distribution.kl_divergence,
        )
        de.utils = _module("distribution_extension.utils", stack_distribution=distribution.stack_distribution,
                           cat_distribution=distribution.cat_distribution)
    if _missing("torchrl"):
        _module("torchrl")
        _module("torchrl.modules", MLP=mlp.MLP)
    if _missing("cnn"):
        _module("cnn", Encoder=standins.Encoder, Decoder=standins.Decoder)

    if not override_reference and not _missing("multimodal_rssm"):
        return
    common = dict(Representation=networks.Representation, Transition=networks.Transition, likelihood=objective.likelihood)
    st = dict(State=state.State, cat_states=state.cat_states, stack_states=state.stack_states)
    mt = dict(MTState=mtstate.MTState, cat_mtstates=mtstate.cat_mtstates, stack_mtstates=mtstate.stack_mtstates)
    _module("multimodal_rssm")
    _module("multimodal_rssm.models", **common, **st)
    _module("multimodal_rssm.models.core", BaseRSSM=core.BaseRSSM)
    _module("multimodal_rssm.models.networks", Representation=networks.Representation, Transition=networks.Transition)
    _module("multimodal_rssm.models.state", **st)
    _module("multimodal_rssm.models.objective", likelihood=objective.likelihood)
    _module("multimodal_rssm.models.mrssm")
    _module("multimodal_rssm.models.mrssm.mopoe_mrssm", MoPoE_MRSSM=mopoe_mrssm.MoPoE_MRSSM, **common, **st)
    _module("multimodal_rssm.models.mrssm.mopoe_mrssm.core", MoPoE_MRSSM=mopoe_mrssm.MoPoE_MRSSM)
    _module("multimodal_rssm.models.mmtrssm", **mt)
    _module("multimodal_rssm.models.mmtrssm.state", **mt)
    _module("multimodal_rssm.models.mmtrssm.mopoe_mmtrssm", MTRNN=mopoe_mmtrssm.MTRNN, MoPoE_MMTRSSM=mopoe_mmtrssm.MoPoE_MMTRSSM,
            **common, **mt)
    _module("multimodal_rssm.models.mmtrssm.mopoe_mmtrssm.core", MTRNN=mopoe_mmtrssm.MTRNN, MoPoE_MMTRSSM=mopoe_mmtrssm.MoPoE_MMTRSSM)


def resolve(path: str) -> Any:
    """'pkg.mod.Name' -> object."""
    mod_name, _, attr = path.rpartition(".")
    return getattr(importlib.import_module(mod_name), attr)


def instantiate(node: Any) -> Any:
    """Recursively build `{class_path, init_args}` trees (dicts without class_path and lists are walked)."""
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "class_path" not in node:
        return {k: instantiate(v) for k, v in node.items()}
    cls = resolve(node["class_path"])
    kwargs = {}
    for k, v in (node.get("init_args") or {}).items():
        if k == "activation_class" and isinstance(v, str):
            v = resolve(v)
        elif k == "config" and isinstance(v, dict) and "class_path" not in v:
            pass  # plain dict config (cnn.Encoder / cnn.Decoder)
        else:
            v = instantiate(v)
        kwargs[k] = v
    return cls(**kwargs)


def load_model(yaml_path: str | Path):  # noqa: ANN201
    """Instantiate the `model:` section of a reference-style LightningCLI YAML with this package's classes."""
    import yaml

    install()
    cfg = yaml.safe_load(Path(yaml_path).read_text())
    return instantiate(cfg["model"])
