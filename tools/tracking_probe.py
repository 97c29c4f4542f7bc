#!/usr/bin/env python
"""One tracked frame through the handle API (bench.bench_tracking_frame), for ncu launch lists of the latency path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth
print(json.dumps(bench.bench_tracking_frame(vo, synth)))
