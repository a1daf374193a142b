"""deepmod_b200: B200-native implementation of DeepMod's `detect` hot path.

Python host code over hand-written sm_100a CUDA kernels behind a C ABI
(``include/deepmod_b200.h`` -> ``deepmod_b200/libdeepmod_b200.so``).
"""
__version__ = "0.1.0"
