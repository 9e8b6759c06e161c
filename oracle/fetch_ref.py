"""TEST / BASELINE INFRASTRUCTURE ONLY - makes the unmodified reference hot-path files travel to the GPU box.

    python oracle/fetch_ref.py            # /root/reference -> baseline/_ref/ (git-ignored, NOT gpurun-ignored)

`/root/reference` exists only in the build container.  The reference is pure Python (nothing to compile), so
"building" `_ref` is a verbatim copy of the files `oracle/ref_loader.py` loads by path: the seven pruner modules and the
vendored PEFT LoRA module with its `utils` package.  Nothing is edited; `baseline/_ref/MANIFEST.json` records the
sha256 of every file so that a test (tests/test_reference_arm.py) can check the copy is byte-identical to the source
tree whenever both are present.  The directory is listed in .gitignore: reference sources never enter the history.
`__graft_entry__.build()` calls fetch() when /root/reference is present; bench.py's `--impl reference` / `cpu_baseline`
legs and the live differential tests read the copy through ref_loader (VLMC_REFERENCE_ROOT overrides the location).
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [f"lavis/compression/pruners/{n}.py" for n in (
    "utils", "base_pruner", "layer_single_base_pruner", "wanda_pruner", "sparsegpt_pruner", "dsnot_pruner",
    "global_pruner")] + ["lavis/peft/src/peft/tuners/lora.py"] + [f"lavis/peft/src/peft/utils/{n}.py" for n in (
        "__init__", "adapters_utils", "config", "other", "save_and_load")]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def fetch(src=SRC, dst=DST):
    """Copies FILES from src to dst (same relative paths).  Returns the manifest, or None when src is absent."""
    if not os.path.isfile(os.path.join(src, FILES[0])):
        return None
    manifest = {}
    for rel in FILES:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        manifest[rel] = _sha(out)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1)
    return manifest


def verify(dst=DST):
    """True when every file of the manifest is present in dst with the recorded digest."""
    path = os.path.join(dst, "MANIFEST.json")
    if not os.path.isfile(path):
        return False
    with open(path) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    m = fetch()
    print("reference tree absent: nothing copied" if m is None else f"copied {len(m)} files to {DST}")
