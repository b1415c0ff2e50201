"""Parity-margin log: every GPU parity check records how far it landed from its bound, so a test that passes at 90 %
of its tolerance is visible before an unlucky seed fails it.  One JSON object per line in
gpurun_out/parity_margins.jsonl (copied to profiles/ with the round tag by scripts/collect_profiles.py)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, "gpurun_out", "parity_margins.jsonl")


def record(what, **kv):
    try:
        os.makedirs(os.path.dirname(_PATH), exist_ok=True)
        rec = {"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "what": what}
        rec.update({k: (float(v) if isinstance(v, (int, float)) else v) for k, v in kv.items()})
        with open(_PATH, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
