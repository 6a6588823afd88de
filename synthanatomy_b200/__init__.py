"""synthanatomy_b200 -- B200 (sm_100a) native implementation of SynthAnatomy's VQ-VAE / Performer hot paths.

The compute lives in ``lib/libsynthanatomy_b200.so`` (hand-written CUDA, C ABI in ``include/synthanatomy_b200.h``);
this package is the host-side mirror of the reference's ``src/networks`` plugin surface.
"""
__version__ = "0.1.0"
