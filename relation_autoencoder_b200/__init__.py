"""relation-autoencoder on B200: the fused forward/backward/update training step of the discrete-state relation
autoencoder (reference: boromir674/relation-autoencoder, learning/OieModel.py + decoders + Optimizers.py) as
hand-written sm_100a CUDA kernels behind a C ABI (include/rae.h, librae.so).

Importing the package is cheap; ``Engine`` (and torch) load on first use.
"""

__all__ = ["Engine", "build"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "build":
        from .build import build
        return build
    raise AttributeError(name)
