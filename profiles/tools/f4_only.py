"""bench.py's `f4` rows alone (calibration loop and AAD on the device type): python profiles/tools/f4_only.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import bench_configs as bc  # noqa: E402

pkg = bc.graft.load_package()
pkg.native.init(0)
print(json.dumps(bc.next_rows(pkg)))
