"""scripts/check_conformance.py -- the counterpart of the reference's
src/checkconformanceresults.sh: skip code 77 without the ITU data set, string
comparison of the printed DI at three decimals otherwise (exercised with a stand-in
CLI; the real WAV set is not redistributable)."""
import json
import os
import stat
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "scripts", "check_conformance.py")


def run(env_extra, *args):
    env = {k: v for k, v in os.environ.items() if k != "CONFORMANCEDATADIR"}
    env.update(env_extra)
    return subprocess.run([sys.executable, TOOL] + list(args), capture_output=True, text=True, env=env)


def test_skips_like_the_reference_without_data(tmp_path):
    p = run({})
    assert p.returncode == 77 and "NOT run" in p.stdout
    p = run({"CONFORMANCEDATADIR": str(tmp_path / "missing")})
    assert p.returncode == 77


def test_table_has_the_sixteen_items_of_both_modes():
    t = json.load(open(os.path.join(ROOT, "tests", "golden", "conformance_di.json")))["items"]
    for mode in ("basic", "advanced"):
        assert len(t[mode]) == 16
        assert all(len(r["item"]) == 7 and "cod" in r["item"] for r in t[mode])
    assert t["basic"][0] == {"item": "acodsna", "itu_di": 1.304, "reference_di": "1.297"}


def test_compares_the_printed_di(tmp_path):
    table = {"items": {"basic": [{"item": "acodsna", "itu_di": 1.304, "reference_di": "1.297"},
                                 {"item": "bcodtri", "itu_di": 1.949, "reference_di": "1.973"}],
                       "advanced": []}}
    tpath = tmp_path / "t.json"
    tpath.write_text(json.dumps(table))
    fake = tmp_path / "peaq"
    # prints 1.297 for acodsna and a wrong value for anything else; checks the ref/cod naming
    fake.write_text("#!/bin/sh\ncase \"$3\" in *acodsna.wav) [ \"$2\" = \"%s/arefsna.wav\" ] || exit 3; "
                    "echo 'Objective Difference Grade: -1.000'; echo 'Distortion Index: 1.297';; "
                    "*) echo 'Objective Difference Grade: -1.000'; echo 'Distortion Index: 9.999';; esac\n" % tmp_path)
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    p = run({"CONFORMANCEDATADIR": str(tmp_path)}, "--peaq", str(fake), "--mode", "basic", "--table", str(tpath))
    assert p.returncode == 1, p.stdout + p.stderr
    assert "acodsna 1.297 1.297 OK" in p.stdout and "bcodtri 9.999 1.973 FAILED" in p.stdout
