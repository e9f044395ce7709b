"""Generates tests/golden/dubins_golden.json from the independent pure-Python transliteration of
/root/reference/src/dubins.rs (oracle/dubins_py.py, glibc libm through ctypes).  The reference ships no
golden vectors and cannot be compiled here (no rustc), so these are restatement-derived known answers:
the C oracle must reproduce them bit for bit (tests/test_oracle_dubins.py), the CUDA path to 1e-9.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dubins_py as P  # noqa: E402

R45 = 45.0 * (math.pi / 180.0)
CASES = [
    ("bench benches/all.rs:102-111", (1.0, 1.0, R45, -3.0, -3.0, -R45, 1.0, 0.1)),
    ("example conf1 examples/dubins/src/main.rs:25-34", (1.0, 1.0, R45, -3.0, -3.0, -R45, 0.5, 0.01)),
    ("example conf2 examples/dubins/src/main.rs:36-45", (-3.0, -3.0, -R45, 1.0, 1.0, R45, 0.5, 0.01)),
    ("ccc", (0.0, 0.0, 0.0, 0.5, 0.5, math.pi, 1.0, 0.1)),
    ("straight Q4/Q9", (0.0, 0.0, 0.0, 5.0, 0.0, 0.0, 1.0, 0.1)),
    ("same pose Q7", (2.0, 3.0, 0.7, 2.0, 3.0, 0.7, 1.0, 0.1)),
    ("rrt-like Q14", (10.0, 10.0, 2.356194490192345, 5.0, 15.0, 1.0, 0.8, 0.1)),
    ("u-turn", (0.0, 0.0, 0.0, 0.0, 2.0, math.pi, 1.0, 0.1)),
    ("backwards goal", (0.0, 0.0, 0.0, -4.0, 0.0, 0.0, 1.0, 0.25)),
    ("large radius", (3.0, -2.0, 1.0, 40.0, 25.0, -2.0, 7.5, 0.05)),
]
rnd = random.Random(20261018)
for k in range(60):
    span = 3.0 if k % 2 == 0 else 40.0
    CASES.append((f"random {k}", (rnd.uniform(-span, span), rnd.uniform(-span, span), rnd.uniform(-math.pi, math.pi),
                                  rnd.uniform(-span, span), rnd.uniform(-span, span), rnd.uniform(-math.pi, math.pi),
                                  rnd.choice([0.5, 0.8, 1.0, 2.0]), rnd.choice([0.05, 0.1, 0.3]))))


def hx(v):
    return float(v).hex()


out = []
for name, c in CASES:
    r = P.dubins_path_planning(*c)
    rec = {"name": name, "in": [hx(v) for v in c]}
    if r is None:
        rec["none"] = True
    else:
        px, py, pyaw, word, cost, tpq, n_point = r
        idxs = sorted(set(list(range(0, len(px), 10)) + ([1, len(px) - 1] if len(px) > 1 else [])))
        rec.update({"word": P.WORD_NAMES[word], "cost": hx(cost), "tpq": [hx(v) for v in tpq], "n_point": n_point,
                    "count": len(px), "sample_idx": idxs,
                    "samples": [[hx(px[i]), hx(py[i]), hx(pyaw[i])] for i in idxs]})
    out.append(rec)
# six-word table for a few (alpha, beta, d)
words = []
for (a, b, d) in [(math.pi, math.pi / 2, 5.65685424949238), (0.3, 5.9, 0.7), (2.0, 2.0, 3.0), (4.0, 1.0, 12.5), (0.0, 0.0, 5.0)]:
    row = {"abd": [hx(a), hx(b), hx(d)], "words": []}
    for f in P.ALL_PLANNERS:
        t, p, q, _ = f(a, b, d)
        row["words"].append(None if t is None else [hx(t), hx(p), hx(q)])
    words.append(row)
json.dump({"generator": "oracle/dubins_py.py (pure-Python transliteration of src/dubins.rs, glibc libm)",
           "paths": out, "words": words},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dubins_golden.json"), "w"), indent=0)
print(len(out), "paths")
