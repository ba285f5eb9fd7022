"""echoseal_b200 — B200-native (sm_100a) implementation of EchoSeal's receive/verify hot path and the
embed-side spreading, behind the reference's own entry points (rtwm/__init__.py:9-12).

    from echoseal_b200 import WatermarkEmbedder, WatermarkDetector

The CUDA kernels live in libechoseal_b200.so (C ABI: include/echoseal_b200.h); there is no CPU
fallback — importing works anywhere, calling a compute entry point without the library or a GPU raises."""

__all__ = ["WatermarkEmbedder", "WatermarkDetector", "TxParams"]


def __getattr__(name):
    if name == "WatermarkDetector":
        from .detector import WatermarkDetector
        return WatermarkDetector
    if name in ("WatermarkEmbedder", "TxParams"):
        from . import embedder
        return getattr(embedder, name)
    raise AttributeError(name)
