#!/usr/bin/env python3
"""Extracts the reference's literal lookup tables as test vectors ->
tests/golden/ref_tables.json.  Run in the build container only (needs the
read-only reference checkout); the JSON is committed because /root/reference
does not exist on the GPU box.

    python tools/extract_ref_tables.py [/root/reference]
"""
import json
import re
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def table(text: str, name: str):
    m = re.search(name + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, flags=re.S)
    if not m:
        raise SystemExit(f"table {name} not found")
    return [int(x) for x in re.findall(r"-?\d+", m.group(1))]


ise = strip_comments((ref / "ASTC_IntegerSequenceEncoding.hlsl").read_text(encoding="utf-8", errors="replace"))
tab = strip_comments((ref / "ASTC_Table.hlsl").read_text(encoding="utf-8", errors="replace"))
out = {
    "source": "niepp/astc_encoder ASTC_IntegerSequenceEncoding.hlsl:5-71, ASTC_Table.hlsl:3-66",
    "bits_trits_quints_table": table(ise, "bits_trits_quints_table"),
    "integer_from_trits": table(ise, "integer_from_trits"),
    "integer_from_quints": table(ise, "integer_from_quints"),
    "scramble_table": table(tab, "scramble_table"),
}
assert len(out["bits_trits_quints_table"]) == 63 and len(out["integer_from_trits"]) == 243
assert len(out["integer_from_quints"]) == 125 and len(out["scramble_table"]) == 12 * 32
dst = Path(__file__).resolve().parents[1] / "tests" / "golden" / "ref_tables.json"
dst.write_text(json.dumps(out, separators=(",", ":")) + "\n")
print("wrote", dst)
