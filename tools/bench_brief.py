"""One-line digest of bench.py's JSON line (stdin): value, ms/step, phase times, launches."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d.get('roofline') or {}
print(f"{d['value'] / 1e9:.3f} G  {d['ms_per_step']:.4f} ms  {r.get('kernel_ms')}  launches {d.get('gpu_launches')}  "
      f"frac {((r.get('substep') or {}).get('frac'))}")
