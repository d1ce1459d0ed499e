#!/usr/bin/env python3
"""python tools/seal_kat/compare.py seal_kat_out.json — diff the output of seal_kat (real SEAL 3.7) against
tests/golden/seal_kat_expected.json (the CPU oracle).  Exit code 0 = the oracle's restatement of SEAL is pinned."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[2]


def main():
    got = json.loads(pathlib.Path(sys.argv[1]).read_text())
    exp = json.loads((ROOT / "tests" / "golden" / "seal_kat_expected.json").read_text())
    bad = 0
    for name, e in exp["configs"].items():
        g = got["configs"].get(name)
        if g is None:
            print(f"{name}: missing from the SEAL run")
            bad += 1
            continue
        for key, val in e.items():
            if g.get(key) != val:
                print(f"{name}.{key}: SEAL {g.get(key)!r} != oracle {val!r}")
                bad += 1
    print("SEAL", got.get("seal_version"), "-", "ALL MATCH: oracle pinned" if not bad else f"{bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
