"""Backend registry and factory, same surface as ``python/minisgl/attention/__init__.py:15-75``.

Inside a mini-sglang process ``SUPPORTED_ATTENTION_BACKENDS`` *is* the reference's registry and
importing this module adds the name ``"b200"`` to it (so ``--attn b200`` and ``--attn b200,fi``
validate at argparse time, reference server/args.py:188-195).  Stand-alone, an equivalent
registry is provided.
"""

from __future__ import annotations

from .base import BaseAttnBackend, BaseAttnMetadata, HybridBackend

BACKEND_NAME = "b200"

try:  # pragma: no cover - only when the reference is importable
    from minisgl.attention import SUPPORTED_ATTENTION_BACKENDS

    IN_MINISGL = True
except ImportError:
    from ..utils import Registry

    SUPPORTED_ATTENTION_BACKENDS = Registry("Attention Backend")
    IN_MINISGL = False


def create_b200_backend(config) -> BaseAttnBackend:
    from .backend import B200AttnBackend

    return B200AttnBackend(config)


if BACKEND_NAME not in SUPPORTED_ATTENTION_BACKENDS.supported_names():
    SUPPORTED_ATTENTION_BACKENDS.register(BACKEND_NAME)(create_b200_backend)


def validate_attn_backend(backend: str, allow_auto: bool = True) -> str:
    if backend != "auto":
        SUPPORTED_ATTENTION_BACKENDS.assert_supported(backend.split(",") if "," in backend else [backend])
    elif not allow_auto:
        raise AssertionError("auto is not allowed here")
    return backend


def create_attention_backend(backend: str, config) -> BaseAttnBackend:
    """``"b200"`` or ``"p,d"`` (hybrid prefill/decode) -- reference __init__.py:52-66."""
    validate_attn_backend(backend, allow_auto=False)
    if "," in backend:
        if backend.count(",") != 1:
            raise AssertionError("Only one comma is allowed in hybrid backend")
        p_name, d_name = backend.split(",", 1)
        if p_name != d_name:
            return HybridBackend(
                create_attention_backend(p_name, config), create_attention_backend(d_name, config)
            )
        backend = p_name
    return SUPPORTED_ATTENTION_BACKENDS[backend](config)


__all__ = [
    "BACKEND_NAME",
    "BaseAttnBackend",
    "BaseAttnMetadata",
    "HybridBackend",
    "SUPPORTED_ATTENTION_BACKENDS",
    "create_attention_backend",
    "validate_attn_backend",
]
