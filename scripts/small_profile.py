"""Phase timeline of conv_small_kernel inside the graphed e2e iteration (debug build -DAVC_SMALL_PROFILE).
Usage on the GPU box:  AVC_NVCC_EXTRA=-DAVC_SMALL_PROFILE python -m attack_vc_b200.build --force && python scripts/small_profile.py
Columns per launch (CTA 0): gap since the previous conv_small exit (ns, globaltimer), entry->dependency wait released,
wait->windows ready, main loop, split-K reduce + store (SM cycles), kernel entry->exit (ns)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from attack_vc_b200 import Engine  # noqa: E402
from attack_vc_b200._lib import load  # noqa: E402
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs  # noqa: E402

lib = load()
fn = lib.avc_debug_small_profile
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to("cuda"))
inp = make_inputs("e2e", 1, 256, seed=1)
x, at, src, w0 = (inp[k].cuda() for k in ("vc_tgt", "adv_tgt", "vc_src", "w0"))
buf = np.zeros((8192, 12), dtype=np.uint64)
eng.attack("e2e", x, at, 0.1, 64, vc_src=src, w0=w0)
fn(buf.ctypes.data, 8192)                       # drop warm-up
eng.attack("e2e", x, at, 0.1, 48, vc_src=src, w0=w0)
n = fn(buf.ctypes.data, 8192)
b = buf[:n].astype(np.int64)
# keep the steady-state part: last 16 iterations
per_it = 52
rows = b[-16 * per_it:]
rows = rows.reshape(16, per_it, 12)
t_in, t_out = rows[:, :, 0], rows[:, :, 6]
gap = np.zeros_like(t_in)
gap[:, 1:] = t_in[:, 1:] - t_out[:, :-1]
ph = lambda i, j: np.median(rows[:, :, j] - rows[:, :, i], axis=0)
meta = rows[0, :, 7]
print("iteration span (first conv entry -> last conv exit), us:", np.median(t_out[:, -1] - t_in[:, 0]) / 1e3)
print(" #  grid     pro epi | gap_ns | ent->wait  wait->win  mainloop  reduce+st (cycles) | kernel_ns")
tot = np.zeros(6)
for k in range(per_it):
    gx, gy, pm, em = meta[k] >> 32, (meta[k] >> 16) & 0xffff, (meta[k] >> 4) & 0xf, meta[k] & 0xf
    vals = [np.median(gap[:, k]), ph(1, 2)[k], ph(2, 3)[k], ph(3, 4)[k], ph(4, 5)[k], np.median(t_out[:, k] - t_in[:, k])]
    extra = f"  fold: wait->cp {ph(2, 10)[k]:.0f} sync {ph(10, 8)[k]:.0f} combine {ph(8, 9)[k]:.0f} sync {ph(9, 11)[k]:.0f} transform {ph(11, 3)[k]:.0f}" if pm else ""
    tot += np.array(vals)
    print(f"{k:2d}  {gx:3d}x{gy:<3d}  {pm}   {em}  | {vals[0]:6.0f} | {vals[1]:8.0f} {vals[2]:9.0f} {vals[3]:9.0f} {vals[4]:9.0f} | {vals[5]:7.0f}{extra}")
print("sum: gap %.1f us | ent->wait %.0f  wait->win %.0f  main %.0f  reduce %.0f cycles | in-kernel %.1f us" % (tot[0] / 1e3, tot[1], tot[2], tot[3], tot[4], tot[5] / 1e3))
