#!/usr/bin/env python3
"""Import the reference's INPUT DATA (cone coordinates, vehicle parameters) into this repo.

The lap-time hot path is defined on named inputs ("Buckmore track + TBR18 vehicle", BASELINE.json),
and /root/reference does not exist on the GPU box, so the numbers have to travel with the repo.
Only data is imported -- no reference source code.  The files keep the reference's JSON schema
(`{"name", "left": {"x","y"}, "right": {"x","y"}}` read by reference src/track.py:52-60, and the vehicle
schemas read by src/vehicle.py:13-22 / src/vehicleMX5.py:46-79) so that a user's own files in that
schema load unchanged; MX5.json in the reference carries `//` comments, which we strip here (our
loader also accepts them).

Usage (in the build container only):  python tools/make_data.py
"""
import json
import os
import re
import sys

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "lap_time_optimization_b200", "data")


def strip_comments(text):
    text = re.sub(r"//.*", "", text)
    return re.sub(r"/\*.*?\*/", "", text, flags=re.DOTALL)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference data not available here; data/ is already committed")
    for sub in ("tracks", "vehicles"):
        os.makedirs(os.path.join(OUT, sub), exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(REF, sub))):
            if not fn.endswith(".json"):
                continue
            with open(os.path.join(REF, sub, fn)) as f:
                obj = json.loads(strip_comments(f.read()))
            with open(os.path.join(OUT, sub, fn), "w") as f:
                json.dump(obj, f, separators=(",", ":"))
                f.write("\n")
            print("wrote", sub, fn)


if __name__ == "__main__":
    main()
