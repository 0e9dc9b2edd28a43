import sys, os
sys.path.insert(0, os.getcwd())
import torch, b200ring
sys.path.insert(0, "tools")
from bench_configs import QI60, uniform, timed
N, polys = 1 << 15, 128      # 4N = 2^17 divides q - 1 for Qi60 (q = 1 mod 2^18)
moduli = QI60[:24]
for typ, n in ((b200ring.ConjugateInvariant, N), (b200ring.Standard, N), (b200ring.Standard, 2 * N)):
    ring = b200ring.NewRingFromType(n, moduli, typ)
    ring.SetStream(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    data = uniform(ring, moduli, polys, n, gen)
    ref = data.clone()
    p = b200ring.Poly.from_torch(ring, data)
    ms = timed(lambda: (ring.NTT(p, p), ring.INTT(p, p)), 10)
    assert torch.equal(ref, data)
    prof = ring.Profile(lambda: (ring.NTT(p, p), ring.INTT(p, p)))
    print(str(typ), "N=%d" % n, "%.3f ms/step" % ms, "%.0f limb-transforms/s" % (2 * polys * 24 / ms * 1e3), {k: round(v["ms"], 3) for k, v in prof.items()})
