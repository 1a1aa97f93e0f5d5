"""One-line digest of bench.py's JSON line (file argument, else stdin): value, ms/step, phase times, launches."""
import json
import sys

text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads(text.strip().splitlines()[-1])
r = d.get('roofline') or {}
print(f"{d['value'] / 1e9:.3f} G  {d['ms_per_step']:.4f} ms  {r.get('kernel_ms')}  launches {d.get('gpu_launches')}  "
      f"frac {((r.get('substep') or {}).get('frac'))}")
