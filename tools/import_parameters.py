#!/usr/bin/env python3
"""Collect the reference's parameter sets into one fixture.

Reads every /root/reference/parameters/*.json (APSU's PSUParams JSON schema,
common/apsu/psu_params.cpp:290-374) and writes tests/golden/parameters.json as
{file name: parsed object}.  Run in the build container only (the GPU box has no
/root/reference); the output is committed.
"""
import json, pathlib, sys

src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/parameters")
out = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "parameters.json"
table = {p.name: json.loads(p.read_text()) for p in sorted(src.glob("*.json"))}
out.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
print(f"wrote {len(table)} parameter sets to {out}")
