"""b200ring: B200-native RNS NTT / NTT-domain multiply behind Lattigo's ring.Ring / ring.Poly API.

Import it as `b200ring` (see ../b200ring/__init__.py).  Layout:
  csrc/      hand-written CUDA for sm_100a + the C ABI (include/b200ring.h)
  capi.py    ctypes binding of the C ABI (fails loudly when lib/libb200ring.so is missing)
  ring.py    host-side mirror of the reference's ring.Ring / ring.SubRing / ring.Poly method set
  sharding.py limb-wise partition over the GPUs of one box (torch.distributed plumbing)
"""
