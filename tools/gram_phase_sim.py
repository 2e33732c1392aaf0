"""Host-side estimate of how much of the work-list Gram's panel traffic is shared through L2 (no GPU needed):
   python tools/gram_phase_sim.py [m] [n]

Calls lb2_gram_wl_plan_sharing (csrc/gram_wl.cu), which replays the ACTUAL schedule on a common clock — every CTA advances
at its tiles' cost rate — and counts the distinct (panel, row window) requests, with the pieces walked from their first
row (phase 0) and with the phase-aligned cyclic walk (phase 1, the default)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from lobpcg_b200 import api

m = int(sys.argv[1]) if len(sys.argv) > 1 else 896
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096000
for ph in (0, 1):
    for w in (16, 64, 256):
        share = C.c_double(0)
        rc = api.lib().lb2_gram_wl_plan_sharing(m, m, 1, n, 148, 32, ph, w, 400, C.byref(share))
        print(f"m={m} n={n} phase={ph} window={w} chunks: rc={rc} distinct share of panel requests {share.value:.3f}")
