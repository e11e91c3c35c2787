"""Runs the UNMODIFIED reference tracker (tools/nusc_shasta/pub_tracker_merged.py) on synthetic detection sequences
and writes tests/golden/tracker_*.json. Run in the authoring container (needs /root/reference)."""
import contextlib
import copy
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("SHASTA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "tools", "nusc_shasta"))

from oracle import tracker_oracle as TO  # noqa: E402


def main():
    import pub_tracker_merged as R  # the reference module, imported from its own directory
    for seed, max_age in ((1, 0), (2, 3), (3, 3)):
        seq = TO.synthetic_sequence(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            trk = R.PubTrackerMerged(hungarian=False, max_age=max_age)
        outs = []
        for frame in copy.deepcopy(seq):
            outs.append(TO.summarize(trk.step_centertrack(frame, 0.5)))
        path = os.path.join(ROOT, "tests", "golden", "tracker_seed%d_age%d.json" % (seed, max_age))
        with open(path, "w") as f:
            json.dump({"seed": seed, "max_age": max_age, "time_lag": 0.5, "outputs": outs}, f)
        print(path, sum(len(o) for o in outs), "tracks over", len(outs), "frames")


if __name__ == "__main__":
    main()
