"""vnext_b200 — B200-native (sm_100a) hot path of VNext (SeqFormer / IDOL).

The CUDA library `vnext_b200/lib/libvnext_b200.so` (C ABI: include/vnext_b200.h) is the product;
this package is the host-side mirror of the reference's operator interfaces on top of it.
There is NO CPU fallback: every op raises if the library is missing or a tensor is not on a GPU.
"""
__version__ = "0.1.0"
