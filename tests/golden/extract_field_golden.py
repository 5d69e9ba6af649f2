"""Extracts the BN254 golden rows from the reference's own fixture file
(crates/jolt-field/tests/golden_bytes.rs:68-329) into field_golden.json.
Run in the authoring container only (needs /root/reference):
    python tests/golden/extract_field_golden.py
The JSON is committed; nothing at test time reads /root/reference."""
import json, re, pathlib

SRC = pathlib.Path("/root/reference/crates/jolt-field/tests/golden_bytes.rs")
WANTED = ["FIX_BN254_FR", "FIX_BN254_FQ", "FIX_BN254_FR_CHALLENGE", "FIX_BN254_FR_SCALAR_CHALLENGE",
          "FIX_BN254_FQ_CHALLENGE", "FIX_BN254_FQ_SCALAR_CHALLENGE"]
text = SRC.read_text()
out = {"source": "a16z/jolt crates/jolt-field/tests/golden_bytes.rs @ ff9f8c13"}
for name in WANTED:
    m = re.search(r"const %s: &\[\(&str, &str\)\] = &\[(.*?)\n\];" % name, text, re.S)
    rows = re.findall(r'\(\s*"([0-9a-f]*)",\s*"([0-9a-f]*)",?\s*\)', m.group(1))
    assert rows, name
    out[name] = [list(r) for r in rows]
path = pathlib.Path(__file__).with_name("field_golden.json")
path.write_text(json.dumps(out, indent=1) + "\n")
print({k: len(v) for k, v in out.items() if k != "source"})
