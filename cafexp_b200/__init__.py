"""cafexp_b200 — B200-native (sm_100a) per-family birth-death likelihood engine behind a C ABI.

`hostio` prepares flat arrays from CAFE text inputs; `engine` binds the CUDA library
(cafexp_b200/libcafe_b200.so, built by __graft_entry__.build()).  There is no CPU fallback: any
compute entry point raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
