import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from agi_lidar_slam_b200 import _cabi, synth
from agi_lidar_slam_b200.replay import LioReplay, MeasureGroup, ReplayConfig
from oracle import pyoracle as orc
from replay_oracle import OracleReplay
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seq = synth.sequence(n, 2002)
ctx = _cabi.Context(0, max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=1 << 21)
gpu = LioReplay(ctx, ReplayConfig(max_iteration=3)); cpu = OracleReplay(orc, max_iteration=3)
for k, m in enumerate(seq):
    xg0, Pg0 = gpu.x.copy(), gpu.P.copy()
    a = gpu.process(MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"]))
    b = cpu.process(m)
    ga = gpu.log[-1] if gpu.log else {}; gb = cpu.log[-1] if cpu.log else {}
    dx = np.abs(gpu.x - cpu.x).max(); dP = np.abs(gpu.P - cpu.P).max()
    print(k, ga.get("status"), "m", ga.get("m"), gb.get("m"), "valid", ga.get("n_valid"), gb.get("n_valid"), "passes", ga.get("n_passes"), gb.get("n_passes"),
          "counts", ga.get("counts"), gb.get("counts"), "|dx| %.2e |dP| %.2e" % (dx, dP), "map", ctx.map_size()[1] if gpu.map_built else 0, cpu.map.size() if cpu.map else 0)
