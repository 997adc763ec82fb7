"""Every parity figure the -m gpu tests measure is also RECORDED: one JSON object per test case in
gpurun_out/r02_parity.json under the repository root (gpurun merges gpurun_out/ back; the copy under
profiles/ is the committed one).  Asserting a tolerance and printing the number are different things."""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "r02_parity.json")


def record(case: str, **figures) -> None:
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        data = {}
        if os.path.exists(PATH):
            with open(PATH) as f:
                data = json.load(f)
        figures = {k: (round(float(v), 8) if isinstance(v, (int, float)) and not isinstance(v, bool) else v)
                   for k, v in figures.items()}
        figures["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
        data[case] = figures
        with open(PATH, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass      # a read-only checkout must not fail a parity test
