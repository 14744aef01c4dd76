"""Documentation drift checks (no GPU): every run-time switch DESIGN.md §7b lists exists in the sources, every getenv("HB_...") of the
library is listed there, and every evidence file DESIGN.md cites under profiles/ is committed."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sources():
    out = ""
    for d in ("hala_b200/csrc", "hala_b200", "scripts", "tests"):
        for f in os.listdir(os.path.join(ROOT, d)):
            if f.endswith((".cu", ".cuh", ".py", ".hpp", ".cpp")):
                out += open(os.path.join(ROOT, d, f), errors="ignore").read()
    return out


def test_runtime_switches_are_documented_and_exist():
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    table = design[design.index("## 7b. Run-time switches"):design.index("## 8. Out of scope")]
    documented = set(re.findall(r"`(HB_[A-Z0-9_]+)", table))
    src = _sources()
    lib = ""
    for f in os.listdir(os.path.join(ROOT, "hala_b200", "csrc")):
        if f.endswith((".cu", ".cuh")):
            lib += open(os.path.join(ROOT, "hala_b200", "csrc", f)).read()
    used = set(re.findall(r'getenv\("(HB_[A-Z0-9_]+)"\)', lib)) | set(re.findall(r'env_int\("(HB_[A-Z0-9_]+)"', lib))
    assert documented, "switch table not found"
    missing_in_sources = {v for v in documented if v not in src}
    assert not missing_in_sources, f"documented but not in the sources: {sorted(missing_in_sources)}"
    undocumented = used - documented
    assert not undocumented, f"read by the library but not in DESIGN.md §7b: {sorted(undocumented)}"


def test_cited_profiles_exist():
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for path in set(re.findall(r"`(profiles/[A-Za-z0-9_./+\-]+\.(?:json|jsonl|csv|log|txt))`", text)):
            assert os.path.exists(os.path.join(ROOT, path)), f"{doc} cites {path}, which is not in the tree"
