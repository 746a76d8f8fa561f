"""CPU tier: the evidence index stays in sync with the files — every file under profiles/ is described in
profiles/README.md, every file the READMEs / DESIGN.md cite under profiles/, tools/ and tests/ exists."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*p):
    return open(os.path.join(ROOT, *p), encoding="utf-8").read()


def test_every_profile_is_indexed_and_every_indexed_profile_exists():
    idx = _read("profiles", "README.md")
    listed = set(re.findall(r"`([A-Za-z0-9_.\-]+\.(?:txt|json))`", idx))
    present = {f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith((".txt", ".json"))}
    assert present <= listed, f"not described in profiles/README.md: {sorted(present - listed)}"
    cited = {f for f in listed if f.startswith(("r1_", "r2_", "roofline_"))}
    assert cited <= present, f"described but missing: {sorted(cited - present)}"


def test_files_cited_by_the_documents_exist():
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md"), os.path.join("tools", "README.md")):
        text = _read(doc)
        for m in re.finditer(r"`((?:profiles|tools|tests|oracle|include|macaw-llm_b200)/[A-Za-z0-9_./\-]+\.(?:py|sh|txt|json|cu|cuh|h|md))", text):
            if not os.path.exists(os.path.join(ROOT, m.group(1))):
                missing.append((doc, m.group(1)))
        if doc == "DESIGN.md":  # bare profile names in the review table / results section
            for m in re.finditer(r"`(r[12]_[A-Za-z0-9_]+\.txt|roofline_traffic\.json)`", text):
                if not os.path.exists(os.path.join(ROOT, "profiles", m.group(1))):
                    missing.append((doc, "profiles/" + m.group(1)))
    assert not missing, missing
