"""TEST / BENCH INFRASTRUCTURE ONLY — archives the four UNMODIFIED reference scripts into oracle/_ref/.

    python -m oracle.build_ref        (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no native code to compile), so its "build" for the CPU baseline arm is a byte-for-byte
zip of the scripts that hold the hot path (SURVEY.md abbreviations RV, RW, LG, LS) plus a manifest of their SHA-256.
oracle/_ref/ is listed in .gitignore — no reference source enters the repository history — and not in
.gpurunignore, so the archive reaches the GPU box with the snapshot, where `bench.py --impl reference` and the
`cpu_baseline` leg load `ptReplica.likelihood_func` from it through oracle/ref_loader.py (kind: "reference").
"""
from __future__ import annotations

import hashlib
import json
import os
import zipfile

from . import ref_loader


def build(quiet: bool = False) -> bool:
    if not ref_loader.reference_available():
        if not quiet:
            print("oracle.build_ref: /root/reference absent — keeping the existing archive" if ref_loader.archive_available()
                  else "oracle.build_ref: /root/reference absent and no archive: the CPU arm will time the oracle port")
        return ref_loader.archive_available()
    os.makedirs(os.path.dirname(ref_loader.ARCHIVE), exist_ok=True)
    manifest = {}
    tmp = ref_loader.ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for abbrev, rel in sorted(ref_loader.VARIANTS.items()):
            src = os.path.join(ref_loader.REFERENCE_ROOT, rel)
            data = open(src, "rb").read()
            manifest[abbrev] = dict(path=rel, sha256=hashlib.sha256(data).hexdigest(), bytes=len(data))
            info = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))   # fixed timestamp: reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(info, data)
        z.writestr(zipfile.ZipInfo("MANIFEST.json", date_time=(2020, 1, 1, 0, 0, 0)), json.dumps(manifest, indent=1))
    os.replace(tmp, ref_loader.ARCHIVE)
    if not quiet:
        print(f"oracle.build_ref: wrote {ref_loader.ARCHIVE} ({os.path.getsize(ref_loader.ARCHIVE)} bytes, {len(manifest)} scripts)")
    return True


if __name__ == "__main__":
    build()
