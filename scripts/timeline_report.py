"""Per-class timeline of the tiled kernel's CTAs from a -DTILED_TIMELINE dump ($MCS_TILED_TIMELINE)."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.float64)
t0 = a[:, 0].min()
a = np.where(a > 0, a - t0, np.nan) / 1e3   # us
names = ["start", "prologue", "seg0 BAND", "seg1 FAST", "seg2 WARP", "seg3 COPY", "seg4 ZERO", "end"]
for i, n in enumerate(names):
    col = a[:, i]
    if np.all(np.isnan(col)):
        continue
    print("%-10s min %8.1f  median %8.1f  max %8.1f us" % (n, np.nanmin(col), np.nanmedian(col), np.nanmax(col)))
print("CTA busy time: min %.1f median %.1f max %.1f us" % tuple(f(a[:, 7] - a[:, 0]) for f in (np.nanmin, np.nanmedian, np.nanmax)))
