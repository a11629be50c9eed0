"""Prints the Encoder's error against the fp32 torch reference for every precision mode (GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.path.join(ROOT, "tests", "_encoder_child.py")

for mode in (0, 1, 2):
    for case in ("golden", "4", "131", "16384"):
        out = subprocess.run([sys.executable, CHILD, case, str(mode)], capture_output=True, text=True, timeout=600)
        tail = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:]
        print(json.dumps({"mode": mode, "case": case, "rc": out.returncode}), tail, flush=True)
