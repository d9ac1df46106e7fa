#!/usr/bin/env python
"""Conformance check against the ITU-R BS.1387 test items -- the counterpart of the
reference's src/checkconformanceresults.sh (lines 5-39): for each of the 16 items
run the CLI on (<item with 'cod' -> 'ref'>.wav, <item>.wav), take the printed
"Distortion Index:" and compare it, as a three-decimal string, with the DI the
reference itself prints for that item (tests/golden/conformance_di.json, restated
from the reference's doc/conformance_*_table.xml, column "Actual DI").

The ITU WAV set is not redistributable and is not in this repository: like the
reference's script this exits 77 (skipped) when CONFORMANCEDATADIR is unset or
missing.  Items must be 48 kHz WAVs (the CLI does not resample).

  CONFORMANCEDATADIR=/path/to/itu python scripts/check_conformance.py [--peaq PATH] [--mode basic|advanced|both]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def di_of(peaq, mode, ref, cod):
    env = dict(os.environ, LC_ALL="C")
    p = subprocess.run([peaq, "--" + mode, ref, cod], capture_output=True, text=True, env=env)
    if p.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (peaq, p.returncode, p.stderr.strip() or p.stdout.strip()))
    for line in p.stdout.splitlines():
        if line.startswith("Distortion Index:"):
            return line.split(" ")[2]
    raise RuntimeError("no 'Distortion Index:' line in the CLI output")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--peaq", default=os.path.join(ROOT, "gstpeaq_b200", "peaq"))
    ap.add_argument("--mode", default="both", choices=["basic", "advanced", "both"])
    ap.add_argument("--table", default=os.path.join(ROOT, "tests", "golden", "conformance_di.json"))
    args = ap.parse_args()
    data_dir = os.environ.get("CONFORMANCEDATADIR")
    if not data_dir:
        print("CONFORMANCEDATADIR not set, conformance test NOT run.")
        return 77
    if not os.path.isdir(data_dir):
        print("Reference data not found, conformance test NOT run.")
        return 77
    table = json.load(open(args.table))["items"]
    for mode in (["basic", "advanced"] if args.mode == "both" else [args.mode]):
        print("%s version:" % mode.capitalize())
        for row in table[mode]:
            item = row["item"]
            cod = os.path.join(data_dir, item + ".wav")
            ref = os.path.join(data_dir, item.replace("cod", "ref", 1) + ".wav")
            di = di_of(args.peaq, mode, ref, cod)
            ok = di == row["reference_di"]
            print(item, di, row["reference_di"], "OK" if ok else "FAILED")
            if not ok:
                return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
