import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "active-orchard-slam_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from aos_gpu import lib
ctx = lib.Context(0)
L = ctx.L
def run(w, h, R):
    rng = np.random.default_rng(1)
    img = (rng.random((h, w)) < 0.002)
    bits = torch.from_numpy(lib.pack_bits(img).view(np.int32)).cuda()
    out = torch.zeros_like(bits); outb = torch.zeros_like(bits)
    for name, fn in (("inflate", lambda: L.aos_inflate_bits(ctx.h, bits.data_ptr(), out.data_ptr(), outb.data_ptr(), w, h, R)),
                     ("open", lambda: L.aos_open_bits(ctx.h, out.data_ptr(), outb.data_ptr(), w, h)),
                     ("thin", lambda: L.aos_thin_bits(ctx.h, outb.data_ptr(), w, h, None, None))):
        rc = fn(); rc2 = L.aos_synchronize(ctx.h)
        print(w, h, R, name, rc, rc2, L.aos_last_error(ctx.h).decode() if (rc or rc2) else "ok", flush=True)
        if rc or rc2: sys.exit(1)
for (w, h, R) in [(4096, 256, 16), (2000, 1200, 16), (400, 240, 16), (400, 240, 0), (31, 7, 3)]:
    run(w, h, R)
