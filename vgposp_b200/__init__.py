"""vgposp_b200 -- B200-native implementation of VGPosp's GP placement hot path.

Public surface mirrors the reference's two hot-path modules:

    vgposp_b200.placement_algorithm2   greedy mutual-information placement (reference placement_algorithm2.py)
    vgposp_b200.gp_functions           kernel / GP / VGP helpers            (reference gp_functions.py)

Both call hand-written sm_100a CUDA kernels through the C-ABI of include/vgposp.h (ctypes binding in
vgposp_b200._ffi).  There is no CPU fallback.
"""
__version__ = "0.1.0"
