import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gridpp_b200 as gpp
from util import golden
g = golden("statistics")
f = g["field"]
for hw in (0, 1, 3):
    for name, st in (("std", gpp.Std), ("variance", gpp.Variance)):
        want = g["nbh_hw%d__%s" % (hw, name)]
        got = gpp.neighbourhood(f, hw, st)
        a = np.isnan(want) & ~np.isnan(got)
        b = ~np.isnan(want) & np.isnan(got)
        print(hw, name, "want nan/got num:", a.sum(), "got nan/want num:", b.sum(), "nan in f:", np.isnan(f).sum())
        idx = np.argwhere(a)[:5]
        for (y, x) in idx:
            print("   at", y, x, "f=", f[y, x], "got=", got[y, x], "want=", want[y, x], "mean=", gpp.neighbourhood(f, hw, gpp.Mean)[y, x])
        idx = np.argwhere(b)[:3]
        for (y, x) in idx:
            print("   (b) at", y, x, "f=", f[y, x], "got=", got[y, x], "want=", want[y, x])
