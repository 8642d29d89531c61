"""Run the flip-aware full-network parity check (tests/test_capnet_gpu.py) over extra (mode, B, N, seed) cases and
print one line per case -- a diagnostic for how often / why a case exceeds the per-parameter gradient bar."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import test_capnet_gpu as T  # noqa: E402

cases = []
for spec in sys.argv[1:]:
    mode, B, N, seed = spec.split(",")[:4]
    name = "x_%s_B%s_N%s_s%s" % (mode, B, N, seed)
    T.CASES[name] = (mode, int(B), int(N), False, 150, None, int(seed))
    cases.append(name)
for name in cases:
    try:
        T.test_capnet_forward_backward_parity(name)
        print("OK  ", name, T.BRANCHES[name], flush=True)
    except AssertionError as e:
        print("FAIL", name, str(e)[:700], flush=True)
