"""GPU probe: LDS.64 rate of the lane-replicated rare-frequency table against compact layouts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hibag_b200 import api
api.set_device(0)
clk = api.device_info()["clock_khz"]*1e3; sm = api.device_info()["sm_count"]
for w, name in ((0,'popc32'),(4,'lds64 lane-replicated'),(8,'lds64 compact window8'),(9,'lds64 compact same row'),(10,'lds64 compact window32')):
    ops, ms = api.pipe_peak(w)
    print("%-28s %.3e /s  %.2f /clk/SM  %.3f ms" % (name, ops, ops/clk/sm, ms))
