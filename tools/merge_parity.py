"""Merge the parity records of the LAST gpurun call (gpurun_out/r02_parity.json: only the tests that ran in that call, because
every call starts from a fresh snapshot without gpurun_out/) into the committed log profiles/r02_parity.json.
usage: python tools/merge_parity.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(ROOT, "gpurun_out", "r02_parity.json"), os.path.join(ROOT, "profiles", "r02_parity.json")
new = json.load(open(src))
old = json.load(open(dst)) if os.path.exists(dst) else {}
old.update(new)
json.dump(old, open(dst, "w"), indent=1, sort_keys=True)
print(f"{dst}: {len(old)} records ({len(new)} from the last call)")
