import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gridpp_b200 as gpp
f = np.arange(600, dtype=np.float32).reshape(20, 30)
for hw in (0, 1, 7):
    for st in (gpp.Mean, gpp.Min):
        print(hw, st, gpp.neighbourhood(f, hw, st)[0, :3])
