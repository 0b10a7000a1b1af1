"""Print the SM pipe micro-benchmarks (sapt_measure_pipes) as JSON; `python devtools/pipes.py [out.json]`."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapt_b200 import _lib

d = _lib.measure_pipes(0)
d2 = _lib.measure_pipes(0)   # second pass: clocks settled
out = dict(first=d, second=d2)
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s)
